"""Regenerates tests/golden/*.npz from the REAL reference (oracle/_ref/libref_shim_serial.so,
compiled from /root/reference by oracle/Makefile).  Run in the build container:
    python tests/golden/make_golden.py
The fixtures are small (a few hundred KB) and committed, because /root/reference does not
exist on the GPU box."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import harness as H  # noqa: E402

FORMATS = ["csr", "csc", "ell", "dia", "jad", "bsr"]


def main():
    H.ensure_built()
    ref = H.ref_shim("serial")
    assert ref is not None, "build oracle/_ref first (needs /root/reference)"
    cases = {
        "poisson1d_300": (H.poisson1d(300), False),
        "poisson3d_7pt_9x8x7_sorted": (H.poisson3d_7pt(9, 8, 7), True),      # spmvtest3.c sorts the rows
        "poisson3d_27pt_6x6x5": (H.poisson3d_27pt(6, 6, 5), False),
        "random_400": (H.random_csr(400, 6, 101, values="wide"), False),
    }
    for name, ((ptr, idx, val), sort_rows) in cases.items():
        n = len(ptr) - 1
        x = H.rand_vec(n, 7, "wide")
        out = dict(ptr=ptr, idx=idx, val=val, x=x, sort_rows=np.array(sort_rows))
        for fmt in FORMATS:
            if fmt == "dia" and name.startswith("random"):
                continue
            out[f"y_{fmt}"], _ = ref.spmv(fmt, ptr, idx, val, x, bnr=2, bnc=2, sort_rows=sort_rows)
        np.savez_compressed(os.path.join(HERE, f"spmv_{name}.npz"), **out)
    # solver fixtures: test/test3.c-style system, b = A*1, serial reference
    ptr, idx, val = H.poisson3d_7pt(12, 12, 12)
    n = len(ptr) - 1
    b, _ = ref.spmv("csr", ptr, idx, val, np.ones(n))
    out = dict(ptr=ptr, idx=idx, val=val, b=b)
    for tag, opts in {"cg_jacobi": "-i cg -p jacobi", "cg_ssor": "-i cg -p ssor", "bicgstab_ssor": "-i bicgstab -p ssor",
                      "bicgstab_jacobi": "-i bicgstab -p jacobi", "gmres30_jacobi": "-i gmres -restart 30 -p jacobi",
                      "gmres5_ssor": "-i gmres -restart 5 -p ssor -ssor_omega 1.2"}.items():
        r = ref.solve(ptr, idx, val, b, opts)
        assert r["status"] == 0, (tag, r)
        out[f"iter_{tag}"] = np.array(r["iter"]); out[f"opts_{tag}"] = np.array(opts); out[f"rhist_{tag}"] = r["rhistory"]
        out[f"x_{tag}"] = r["x"]
        print(tag, r["iter"], r["resid"])
    np.savez_compressed(os.path.join(HERE, "solve_poisson3d_12.npz"), **out)
    ptr, idx, val = H.random_csr(1500, 7, 202, band=50)
    b = H.rand_vec(1500, 203)
    out = dict(ptr=ptr, idx=idx, val=val, b=b)
    for tag, opts in {"bicgstab_ssor": "-i bicgstab -p ssor", "gmres20_jacobi": "-i gmres -restart 20 -p jacobi",
                      "bicgstab_none": "-i bicgstab"}.items():
        r = ref.solve(ptr, idx, val, b, opts)
        assert r["status"] == 0, (tag, r)
        out[f"iter_{tag}"] = np.array(r["iter"]); out[f"opts_{tag}"] = np.array(opts); out[f"rhist_{tag}"] = r["rhistory"]
        out[f"x_{tag}"] = r["x"]
        print(tag, r["iter"], r["resid"])
    np.savez_compressed(os.path.join(HERE, "solve_unsym_1500.npz"), **out)

    # BiCG (default solver): reference outputs for the GPU box
    out = {}
    for key, (ptr, idx, val) in {"p7": H.poisson3d_7pt(10, 9, 8), "unsym": H.random_csr(900, 6, 303, band=40)}.items():
        n = len(ptr) - 1
        b, _ = ref.spmv("csr", ptr, idx, val, np.ones(n))
        out[f"ptr_{key}"], out[f"idx_{key}"], out[f"val_{key}"], out[f"b_{key}"] = ptr, idx, val, b
        for pre in ("none", "jacobi"):
            opts = f"-i bicg -p {pre}"
            r = ref.solve(ptr, idx, val, b, opts)
            assert r["status"] == 0
            tag = f"{key}_{pre}"
            out[f"iter_{tag}"] = np.array(r["iter"]); out[f"opts_{tag}"] = np.array(opts)
            out[f"rhist_{tag}"] = r["rhistory"]; out[f"x_{tag}"] = r["x"]
            print("bicg", tag, r["iter"], r["resid"])
    np.savez_compressed(os.path.join(HERE, "solve_bicg.npz"), **out)


EXT = ["cgs", "crs", "cr", "cocg", "cocr", "bicr", "bicrstab", "tfqmr", "gpbicg", "gpbicr", "bicgsafe", "bicrsafe", "orthomin",
       "minres", "fgmres", "bicgstabl", "idrs", "idr1", "jacobi", "gs", "sor"]
EXT_SYMMETRIC_ONLY = ("cr", "cocg", "cocr", "minres")
EXT_STATIONARY = ("jacobi", "gs", "sor")
BASE_WITH_NEW_PRECONS = ["cg", "bicg", "bicgstab", "gmres"]        # the north-star solvers with ILU / transposed SSOR


def ext_solvers():
    """solve_ext.npz: every further solver of the reference on two systems; per case the
    iteration counts of the OpenMP reference at 1, 2, 4 and 8 threads (its own spread), and the
    serial run's first residuals (its solution is the vector of ones to 1e-8, asserted here)."""
    H.ensure_built()
    ref, omp = H.ref_shim("serial"), H.ref_shim("omp")
    out = {}
    systems = {"p7": H.poisson3d_7pt(12, 11, 10), "unsym": H.random_csr(1500, 7, 404, band=50)}
    for key, (ptr, idx, val) in systems.items():
        n = len(ptr) - 1
        b, _ = ref.spmv("csr", ptr, idx, val, np.ones(n))
        out[f"ptr_{key}"], out[f"idx_{key}"], out[f"val_{key}"], out[f"b_{key}"] = ptr, idx, val, b
        for sv in EXT + BASE_WITH_NEW_PRECONS:
            if key == "unsym" and (sv in EXT_SYMMETRIC_ONLY or sv == "cg"):
                continue
            if sv in BASE_WITH_NEW_PRECONS:
                precs = ("ilu", "ilu -ilu_fill 1") + (("ssor",) if sv == "bicg" else ())
            for pre in precs if sv in BASE_WITH_NEW_PRECONS else ("none",) if sv in EXT_STATIONARY else ("none", "jacobi", "ssor", "ilu"):
                if key == "unsym" and sv == "sor":
                    continue                           # SOR(1.9) diverges on this matrix
                opts = f"-i {sv} -p {pre}" + (" -maxiter 4000" if sv in EXT_STATIONARY else "")
                r = ref.solve(ptr, idx, val, b, opts)
                its = [r["iter"]]
                if pre.split()[0] not in ("ssor", "ilu"):         # block SSOR / block ILU change the preconditioner itself with the thread count
                    for t in (2, 4, 8):
                        omp.set_threads(t)
                        its.append(omp.solve(ptr, idx, val, b, opts)["iter"])
                tag = f"{key}_{sv}_{pre}".replace(" -ilu_fill ", "")
                out[f"opts_{tag}"] = np.array(opts); out[f"status_{tag}"] = np.array(r["status"]); out[f"iters_{tag}"] = np.array(its)
                out[f"rhist_{tag}"] = r["rhistory"][:12]
                assert r["status"] == 0 and np.abs(r["x"] - 1.0).max() < 1e-8, tag
                print(tag, r["status"], its, r["resid"])
    np.savez_compressed(os.path.join(HERE, "solve_ext.npz"), **out)

    # one application of M^-1 and M^-H: bitwise fixtures (the triangular solves keep the reference's
    # summation order, so the GPU must reproduce these exactly); serial and 4-thread block variants
    out = {}
    for key, (ptr, idx, val) in {"p7": H.poisson3d_7pt(9, 8, 7), "unsym": H.random_csr(600, 6, 41, band=30),
                                 "p27": H.poisson3d_27pt(6, 5, 4)}.items():
        n = len(ptr) - 1
        b = H.rand_vec(n, 5)
        out[f"ptr_{key}"], out[f"idx_{key}"], out[f"val_{key}"], out[f"b_{key}"] = ptr, idx, val, b
        for pi, pre in enumerate(("ilu", "ilu -ilu_fill 1", "ilu -ilu_fill 3", "ssor", "ssor -ssor_omega 1.3")):
            for t in (1, 4):
                shim = ref if t == 1 else omp
                shim.set_threads(t)
                for tr in (0, 1):
                    tag = f"{key}_{pi}_{t}_{tr}"
                    out[f"opts_{tag}"] = np.array("-p " + pre)
                    out[f"x_{tag}"] = shim.psolve(ptr, idx, val, b, "-p " + pre, transposed=bool(tr))
    omp.set_threads(1)
    np.savez_compressed(os.path.join(HERE, "psolve_ilu_ssor.npz"), **out)


ESOLVE_CASES = ["pi|-e pi -emaxiter 400", "ii|-e ii", "rqi|-e rqi", "cg|-e cg", "cr|-e cr", "crs|-e cr -shift 0.5",
                "si|-e si -ss 3 -ie ii -emaxiter 60", "sipi|-e si -ss 2 -ie pi -emaxiter 300", "li|-e li -ss 3", "lirv|-e li -ss 4 -rval true",
                "ai|-e ai -ss 3", "aicr|-e ai -ss 2 -ie cr"]
ESOLVE_INITS = {"default": "", "cgjac": "-i cg -p jacobi"}


def esolvers():
    """esolve.npz: the compiled serial reference's eigensolvers (tests/esolve_worker.py drives
    shim_esolve), one process per case -- the reference does not survive several lis_esolve calls of
    different kinds in one process -- for the default inner linear solver and for "-i cg -p jacobi"
    on the command line.  Keys: <init>_<case>_<matrix>_{rc,d,rh,x,ev,er,ei}."""
    import subprocess
    import tempfile
    out = {}
    ROOT = os.path.dirname(os.path.dirname(HERE))
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_shim_serial.so")
    worker = os.path.join(ROOT, "tests", "esolve_worker.py")
    for iname, init in ESOLVE_INITS.items():
        for case in ESOLVE_CASES:
            if iname != "default" and case.split("|")[0] not in ("ii", "rqi", "cg", "cr", "li", "ai"):
                continue
            with tempfile.TemporaryDirectory() as d:
                path = os.path.join(d, "o.npz")
                r = subprocess.run([sys.executable, worker, ref, init, path, case], capture_output=True, text=True)
                assert r.returncode == 0, (case, r.stderr[-2000:])
                z = np.load(path)
                for k in z.files:
                    if case.startswith("lirv") and not (k.endswith("_ev") or k.endswith("_d")):
                        continue            # -rval true: only the Ritz values are defined (the rest is uninitialised memory there)
                    out[f"{iname}_{k}"] = z[k][:1] if (case.startswith("lirv") and k.endswith("_d")) else z[k]
    np.savez_compressed(os.path.join(HERE, "esolve.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ext":
        ext_solvers()
    elif len(sys.argv) > 1 and sys.argv[1] == "esolve":
        esolvers()
    else:
        main()
        ext_solvers()
