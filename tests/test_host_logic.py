"""CPU-side tests (no GPU needed): the C-ABI library loads and exports what include/*.h
declares, the host logic behind lis.h (conversion layouts, assembly, option parsing, Matrix
Market input, sorting, error codes) matches the compiled reference, and compute entry points
fail loudly -- never fall back -- when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import harness as H
import lis_b200

INCLUDE = os.path.join(H.ROOT, "include")


def declared_functions(header):
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#define[^\n]*(\\\n[^\n]*)*", "", text)
    names = set()
    for m in re.finditer(r"\b(?:LIS_INT|int|double|void|const char \*|void \*)\s*\*?\s*(\w+)\s*\([^;{]*\)\s*;", text):
        names.add(m.group(1))
    return {n for n in names if not n.startswith("LIS_") or n == "LIS_MATVEC"} - {"conj"}


def test_library_exports_every_declared_symbol(built):
    lib = lis_b200.load_library()
    missing = []
    total = 0
    for header in ("lis.h", "lislib.h", "lis_b200_kernels.h"):
        for name in sorted(declared_functions(header)):
            total += 1
            if not hasattr(lib, name):
                missing.append(f"{header}:{name}")
    assert total > 150, total
    assert not missing, f"declared but not exported: {missing}"


def test_kernel_abi_has_no_torch_types():
    text = open(os.path.join(INCLUDE, "lis_b200_kernels.h")).read()
    assert "torch" not in text and "at::" not in text and "extern \"C\"" in text


def test_no_cpu_fallback_without_device(b200):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the compute path runs")
    with pytest.raises(RuntimeError, match="Lis error 7"):       # LIS_ERR_DEVICE
        b200.vec_op("axpy", np.ones(4), np.ones(4), 1.0)
    ptr, idx, val = H.poisson1d(10)
    with pytest.raises(RuntimeError, match="Lis error 7"):
        b200.spmv("csr", ptr, idx, val, np.ones(10))


def test_product_does_not_reference_oracle():
    """the product (library sources, package, headers) must not include, link or load oracle/"""
    for root in ("lis_b200", "include"):
        for dp, dn, fn in os.walk(os.path.join(H.ROOT, root)):
            if "_build" in dp or "_lib" in dp or "__pycache__" in dp:
                continue
            for f in fn:
                if f.endswith((".c", ".h", ".cu", ".cuh", ".py")):
                    text = open(os.path.join(dp, f)).read()
                    assert "lis_oracle" not in text and "oracle/_ref" not in text and "orc_" not in text, os.path.join(dp, f)
    mk = open(os.path.join(H.ROOT, "Makefile")).read()
    assert "oracle" not in mk


# ------------------------------------------------------------------ conversion layouts (host only)
MATS = {
    "poisson1d": lambda: H.poisson1d(101),
    "poisson3d_unsorted": lambda: H.poisson3d_7pt(6, 5, 7),
    "poisson27": lambda: H.poisson3d_27pt(4, 5, 6),
    "random": lambda: H.random_csr(257, 5, 3, values="wide"),
    "random_empty": lambda: H.random_csr(130, 4, 4, empty_rows=True, diag_dominant=False),
}


@pytest.mark.parametrize("fmt", ["csr", "csc", "ell", "dia", "jad", "bsr"])
def test_conversion_layouts_match_oracle_builders(b200, oracle, fmt):
    """lis_matrix_convert produces the reference's SERIAL layouts (they are what the kernels read)"""
    for name, mk in MATS.items():
        ptr, idx, val = mk()
        if fmt == "dia" and name.startswith("random"):
            continue
        got = b200.convert(fmt, ptr, idx, val, bnr=2, bnc=2)
        if fmt == "csr":
            exp = dict(ptr=ptr, index=idx, value=val)
        elif fmt == "bsr":
            exp = oracle.to_bsr(ptr, idx, val, 2, 2)
        else:
            exp = {"ell": oracle.to_ell, "dia": oracle.to_dia, "jad": oracle.to_jad, "csc": oracle.to_csc}[fmt](ptr, idx, val)
        for key in ("ptr", "index", "value", "row", "bptr", "bindex"):
            if key in got and key in exp:
                a = np.asarray(got[key]); e = np.ascontiguousarray(np.asarray(exp[key])[:len(a)], a.dtype)
                assert np.array_equal(a.view(np.uint8), e.view(np.uint8)), f"{fmt}/{name}/{key}"
        for key in ("maxnzr", "nnd", "bnnz"):
            if key in exp:
                assert got[key] == exp[key], f"{fmt}/{name}/{key}"


@pytest.mark.parametrize("fmt", ["csc", "ell", "dia", "jad", "bsr"])
def test_conversion_layouts_match_reference(b200, ref_serial, fmt):
    for name, mk in MATS.items():
        ptr, idx, val = mk()
        if fmt == "dia" and name.startswith("random"):
            continue
        got = b200.convert(fmt, ptr, idx, val, bnr=3, bnc=2)
        ref = ref_serial.convert(fmt, ptr, idx, val, bnr=3, bnc=2)
        for key in ("maxnzr", "nnd", "bnnz", "nr", "bnr", "bnc"):
            assert got[key] == ref[key], f"{fmt}/{name}/{key}"
        for key in ("ptr", "bptr", "bindex") + (() if fmt == "jad" else ("index", "value")):
            if key in ref:
                assert np.array_equal(np.asarray(got[key]).view(np.uint8), np.asarray(ref[key]).view(np.uint8)), f"{fmt}/{name}/{key}"
        if fmt == "jad":
            # the order of equal-length rows is a quicksort artefact in the reference: compare per row
            def rows_of(d):
                out = {}
                for i, r in enumerate(d["row"]):
                    ks = [int(d["ptr"][j]) + i for j in range(d["maxnzr"]) if i < d["ptr"][j + 1] - d["ptr"][j]]
                    out[int(r)] = (tuple(d["index"][ks]), tuple(d["value"][ks]))
                return out
            assert rows_of(got) == rows_of(ref), f"jad/{name}"
            lens = np.diff(ptr)
            assert list(lens[got["row"]]) == sorted(lens, reverse=True)


@pytest.mark.parametrize("fmt", ["msr", "coo", "bsc", "vbr", "dns"])
def test_other_format_layouts_match_reference(b200, ref_serial, fmt):
    """the five formats outside the named path (host/lis_formats_ext.c): same arrays as the reference's
    serial builders, byte for byte (BSC goes CSR -> CSC -> BSC there; VBR derives its own partition)"""
    for name, mk in MATS.items():
        ptr, idx, val = mk()
        if fmt == "dns" and len(ptr) - 1 > 1500:
            continue
        got = b200.convert(fmt, ptr, idx, val, bnr=2, bnc=2)
        ref = ref_serial.convert(fmt, ptr, idx, val, bnr=2, bnc=2)
        for key in ("nnz", "ndz", "bnnz", "nr", "nc") + (("bnr", "bnc") if fmt == "bsc" else ()):
            assert got[key] == ref[key], f"{fmt}/{name}/{key}: {got[key]} vs {ref[key]}"
        for key in ("row", "col", "ptr", "bptr", "bindex", "index", "value"):
            if key not in ref:
                continue
            a, b = np.asarray(got[key]).copy(), np.asarray(ref[key]).copy()
            if fmt == "msr":
                n = got["n"]
                a[n] = b[n] = 0                      # value[n] is never written by the reference (malloc'ed), index[n] compared through nnz
                if key == "value" and got["ndz"]:
                    continue                          # ... nor the diagonal slot of a row that stores none
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"{fmt}/{name}/{key}"


def test_other_formats_convert_back_to_csr(b200, ref_serial):
    """<fmt> -> CSR: explicit zeros of the dense blocks dropped, MSR diagonal first; same arrays as the
    reference except COO, whose rows keep the order of k (the reference's unstable quicksort does not)"""
    import ctypes as C
    for name, mk in MATS.items():
        ptr, idx, val = mk()
        n = len(ptr) - 1
        x = H.rand_vec(n, 12, "wide")
        for fmt in ("msr", "coo", "bsc", "vbr") + (("dns",) if n <= 1500 else ()):
            if fmt == "msr" and name == "random_empty":
                continue                  # rows without a diagonal entry: the reference's MSR round trip corrupts its heap (glibc abort)
            for shim in (b200, ref_serial):
                L = shim.lib
                L.shim_roundtrip_open.argtypes = [C.c_int, C.c_int, np.ctypeslib.ndpointer(np.int32), np.ctypeslib.ndpointer(np.int32),
                                                  np.ctypeslib.ndpointer(np.float64), C.c_int, C.c_int]
            hs = [s.lib.shim_roundtrip_open(lis_b200.FMT[fmt], n, ptr, idx, val, 2, 2) for s in (b200, ref_serial)]
            assert min(hs) >= 0, (fmt, name, hs)
            got, ref = b200._grab_handle(hs[0], "csr"), ref_serial._grab_handle(hs[1], "csr")
            assert np.array_equal(got["ptr"], ref["ptr"]), f"{fmt}/{name}"
            if fmt == "coo":
                for i in range(n):
                    a = sorted(zip(got["index"][got["ptr"][i]:got["ptr"][i + 1]], got["value"][got["ptr"][i]:got["ptr"][i + 1]]))
                    b = sorted(zip(ref["index"][ref["ptr"][i]:ref["ptr"][i + 1]], ref["value"][ref["ptr"][i]:ref["ptr"][i + 1]]))
                    assert a == b, f"coo/{name}/row {i}"
                assert np.array_equal(got["index"], idx) and np.array_equal(got["value"], val)        # k order kept
            else:
                assert np.array_equal(got["index"], ref["index"]) and np.array_equal(got["value"].view(np.uint8), ref["value"].view(np.uint8)), f"{fmt}/{name}"


def test_set_value_assembly_matches_reference(b200, ref_serial):
    rng = np.random.default_rng(5)
    n = 60
    rows = rng.integers(0, n, 900); cols = rng.integers(0, n, 900); vals = rng.standard_normal(900)
    flags = rng.integers(0, 2, 900)                       # LIS_INS_VALUE / LIS_ADD_VALUE, duplicates included
    for fmt in ("csr", "ell", "csc"):
        a = b200.assemble(n, rows, cols, vals, flags, fmt)
        r = ref_serial.assemble(n, rows, cols, vals, flags, fmt)
        for key in ("ptr", "index", "value"):
            if key in r:
                assert np.array_equal(np.asarray(a[key]).view(np.uint8), np.asarray(r[key]).view(np.uint8)), f"{fmt}/{key}"


def test_matrix_market_input_matches_reference(b200, ref_serial, tmp_path):
    ptr, idx, val = H.random_csr(40, 4, 8)
    n = 40
    b = H.rand_vec(n, 9)
    lines = ["%%MatrixMarket matrix coordinate real general", "% comment line", f"{n} {n} {ptr[-1]} 1 0"]
    rng = np.random.default_rng(1)
    entries = [(i, idx[j], val[j]) for i in range(n) for j in range(ptr[i], ptr[i + 1])]
    for k in rng.permutation(len(entries)):
        i, j, v = entries[k]
        lines.append(f"{i + 1} {j + 1} {v:.20e}")
    lines += [f"{i + 1} {b[i]:.20e}" for i in range(n)]
    path = tmp_path / "a.mtx"
    path.write_text("\n".join(lines) + "\n")
    sym = tmp_path / "s.mtx"
    sym.write_text("%%MatrixMarket matrix coordinate real symmetric\n3 3 4\n1 1 2.0\n2 1 -1.0\n3 2 -1.5\n3 3 4.0\n")
    for p in (path, sym):
        for fmt in ("csr", "ell"):
            ga, gb, gx = b200.input_mm(str(p), fmt)
            ra, rb, rx = ref_serial.input_mm(str(p), fmt)
            for key in ("ptr", "index", "value"):
                if key in ra:
                    assert np.array_equal(np.asarray(ga[key]).view(np.uint8), np.asarray(ra[key]).view(np.uint8)), (p.name, fmt, key)
            assert (gb is None) == (rb is None) and (gx is None) == (rx is None)
            if rb is not None:
                H.assert_bits_equal(gb, rb, "rhs")


def _write_hb(path, ptr, idx, val, ptrfmt=(8, 10), indfmt=(8, 10), valfmt=(3, 26, 18), rhs=False, exponent="E"):
    """a Harwell-Boeing RUA file of the matrix given as CSR (written column by column)"""
    import scipy.sparse as sp
    n = len(ptr) - 1
    csc = sp.csr_matrix((val, idx, ptr), shape=(n, n)).tocsc()
    csc.sort_indices()

    def cards(items, per, fmt):
        lines = []
        for k in range(0, len(items), per):
            lines.append("".join(fmt(v) for v in items[k:k + per]))
        return lines
    pl = cards(list(csc.indptr + 1), ptrfmt[0], lambda v: f"{v:{ptrfmt[1]}d}")
    il = cards(list(csc.indices + 1), indfmt[0], lambda v: f"{v:{indfmt[1]}d}")
    vl = cards(list(csc.data), valfmt[0], lambda v: f"{v:{valfmt[1]}.{valfmt[2]}E}".replace("E", exponent))
    rl = cards([1.0] * n, valfmt[0], lambda v: f"{v:{valfmt[1]}.{valfmt[2]}E}") if rhs else []
    with open(path, "w") as f:
        f.write(f"{'lis_b200 test matrix':<72}{'KEY':<8}\n")
        f.write(f"{len(pl) + len(il) + len(vl) + len(rl):14d}{len(pl):14d}{len(il):14d}{len(vl):14d}{len(rl):14d}\n")
        f.write(f"{'RUA':<14}{n:14d}{n:14d}{csc.nnz:14d}{0:14d}\n")
        vf = f"({valfmt[0]}E{valfmt[1]}.{valfmt[2]})"
        f.write(f"{f'({ptrfmt[0]}I{ptrfmt[1]})':<16}{f'({indfmt[0]}I{indfmt[1]})':<16}{vf:<20}{(vf if rhs else ''):<20}\n")
        if rhs:
            f.write(f"{'F':<14}{1:14d}{0:14d}\n")
        f.write("\n".join(pl + il + vl + rl) + "\n")


def test_harwell_boeing_input_matches_reference(b200, ref_serial, tmp_path):
    """lis_input on Harwell-Boeing RUA files (src/system/lis_input_hb.c): same arrays as the compiled
    reference in CSR, CSC and ELL; several field layouts, a right-hand-side block (skipped by both),
    Fortran D exponents (both read the mantissa only: atof)"""
    cases = [("p7", H.poisson3d_7pt(5, 4, 3), {}),
             ("rand", H.random_csr(150, 6, 3, values="wide"), {"ptrfmt": (13, 6), "indfmt": (16, 5), "valfmt": (4, 20, 12)}),
             ("rhs", H.random_csr(40, 4, 5), {"rhs": True}),
             ("dexp", H.random_csr(30, 3, 6), {"exponent": "D", "valfmt": (3, 26, 17)})]
    for name, (ptr, idx, val), kw in cases:
        path = tmp_path / f"{name}.rua"
        _write_hb(str(path), ptr, idx, val, **kw)
        for fmt in ("csr", "csc", "ell"):
            ga, gb, gx = b200.input_mm(str(path), fmt)
            ra, rb, rx = ref_serial.input_mm(str(path), fmt)
            assert ga["n"] == ra["n"] == len(ptr) - 1
            for key in ("ptr", "index", "value"):
                if key in ra:
                    assert np.array_equal(np.asarray(ga[key]).view(np.uint8), np.asarray(ra[key]).view(np.uint8)), (name, fmt, key)
            assert gb is None and rb is None and gx is None and rx is None
        if name != "dexp":
            ga, _, _ = b200.input_mm(str(path), "csr")
            import scipy.sparse as sp
            n = len(ptr) - 1
            want = sp.csr_matrix((val, idx, ptr), shape=(n, n)); want.sort_indices()
            got = sp.csr_matrix((ga["value"], ga["index"], ga["ptr"]), shape=(n, n)); got.sort_indices()
            assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
            assert np.allclose(got.data, want.data, rtol=1e-11 if name == "rand" else 1e-15, atol=0)


def test_harwell_boeing_rejects_what_the_reference_rejects(b200, ref_serial, tmp_path):
    ptr, idx, val = H.poisson1d(6)
    path = tmp_path / "sym.rsa"
    _write_hb(str(path), ptr, idx, val)
    text = path.read_text().replace("RUA", "RSA")
    path.write_text(text)
    for shim in (b200, ref_serial):
        with pytest.raises(RuntimeError):
            shim.input_mm(str(path), "csr")


def test_lis_output_files_identical_to_reference(b200, ref_serial, tmp_path):
    """lis_output (matrix [+ b, x] as Matrix Market, ASCII and Lis' binary variant) and lis_output_vector
    write byte-identical files, and lis_input reads them back to the same arrays"""
    import ctypes as C
    ptr, idx, val = H.random_csr(60, 5, 17, values="wide")
    n = len(ptr) - 1
    bvec = H.rand_vec(n, 1, "wide"); xvec = H.rand_vec(n, 2, "wide")
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C"); f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    files = {}
    for tag, shim in (("b200", b200), ("ref", ref_serial)):
        shim.lib.shim_output.argtypes = [C.c_int, C.c_int, i32p, i32p, f64p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p,
                                         C.c_int, C.c_char_p]
        for case, fmt, bb, xx, format in (("csr_mm", 1, bvec, xvec, 2), ("csr_mm_nob", 1, None, None, 2), ("ell_mm_b", 5, bvec, None, 2),
                                          ("csr_mmb", 1, bvec, xvec, 8), ("bsr_mmb_nov", 7, None, None, 8)):
            path = tmp_path / f"{tag}_{case}.mtx"
            for vformat, vname in ((1, "plain"), (2, "mm"), (3, "lis")):
                vpath = tmp_path / f"{tag}_{case}_{vname}.vec"
                rc = shim.lib.shim_output(fmt, n, ptr, idx, val, bb.ctypes.data if bb is not None else None,
                                          xx.ctypes.data if xx is not None else None, format, str(path).encode(),
                                          vformat, str(vpath).encode())
                assert rc == 0, (tag, case, rc)
                if bb is not None:
                    files[(tag, case, vname)] = vpath.read_bytes()
            files[(tag, case)] = path.read_bytes()
    for key, data in files.items():
        if key[0] != "b200":
            continue
        ref = files[("ref",) + key[1:]]
        if key[1:] == ("csr_mm",):
            # the reference's serial ASCII writer prints b a second time where x belongs
            # (src/system/lis_output_mm.c:303 reads b->value[i] in the x loop); lis_b200 writes x.
            gl, rl = data.split(b"\n"), ref.split(b"\n")
            assert len(gl) == len(rl) and gl[:-n - 1] == rl[:-n - 1], key
            assert rl[-n - 1:-1] == rl[-2 * n - 1:-n - 1], "reference quirk gone?"
            assert [float(t.split()[1]) for t in gl[-n - 1:-1]] == list(xvec)
        elif key[1:] == ("csr_mmb",):
            # binary vector records {int i; <4 bytes of padding>; double value}: the reference writes
            # whatever its stack held into the padding, lis_b200 zeros
            nv = 2 * n * 16
            assert data[:-nv] == ref[:-nv], key
            ga = np.frombuffer(data[-nv:], np.uint8).reshape(-1, 16); ra = np.frombuffer(ref[-nv:], np.uint8).reshape(-1, 16)
            assert np.array_equal(ga[:, :4], ra[:, :4]) and np.array_equal(ga[:, 8:], ra[:, 8:]) and not ga[:, 4:8].any(), key
        else:
            assert data == ref, key
    # read back (ASCII and binary) with both libraries
    for case in ("csr_mm", "csr_mmb", "ell_mm_b"):
        for tag, shim in (("b200", b200), ("ref", ref_serial)):
            a, rb, rx = shim.input_mm(str(tmp_path / f"b200_{case}.mtx"), "csr")
            assert np.array_equal(a["ptr"], ptr) and np.array_equal(a["index"], idx), (case, tag)
            H.assert_bits_equal(a["value"], val, f"{case} {tag}")
            H.assert_bits_equal(rb, bvec, f"{case} {tag} b")
            if case != "ell_mm_b":
                H.assert_bits_equal(rx, xvec, f"{case} {tag} x")
    # and the reference's own ASCII file (x block = b, see above) reads back the same way in both
    for tag, shim in (("b200", b200), ("ref", ref_serial)):
        a, rb, rx = shim.input_mm(str(tmp_path / "ref_csr_mm.mtx"), "csr")
        H.assert_bits_equal(rb, bvec, tag); H.assert_bits_equal(rx, bvec, tag)


def test_lis_array_matches_reference(built):
    """lis_array_* (the public dense helpers, src/array/lis_array.c): every function, n = 1..6 and the
    three `op` modes, bit for bit against the compiled reference -- incl. the written-out small cases"""
    import ctypes as C
    ref_path = os.path.join(H.ROOT, "oracle", "_ref", "libref_shim_serial.so")
    if not os.path.exists(ref_path):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ref = C.CDLL(ref_path); our = lis_b200.load_library()
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    rng = np.random.default_rng(12)

    def both(name, argtypes, make_args, outs):
        res = []
        args0 = make_args()
        for lib in (ref, our):
            fn = getattr(lib, name); fn.argtypes = argtypes; fn.restype = C.c_int
            args = [a.copy() if isinstance(a, np.ndarray) else a for a in args0]
            rc = fn(*[a if not isinstance(a, np.ndarray) else a for a in args])
            res.append((rc, [args[k].copy() for k in outs]))
        assert res[0][0] == res[1][0], name
        for a, b in zip(res[0][1], res[1][1]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (name, a, b)

    ci, cd = C.c_int, C.c_double
    for n in (1, 2, 3, 4, 6):
        v = lambda k=n: rng.standard_normal(k) * 10.0 ** rng.integers(-3, 3, k)
        m = lambda k=n: (rng.standard_normal((k, k)) + 3 * np.eye(k)).ravel()
        for name in ("lis_array_swap", "lis_array_copy"):
            both(name, [ci, f64p, f64p], lambda: [n, v(), v()], (1, 2))
        both("lis_array_axpy", [ci, cd, f64p, f64p], lambda: [n, 0.37, v(), v()], (3,))
        both("lis_array_xpay", [ci, f64p, cd, f64p], lambda: [n, v(), -1.7, v()], (3,))
        both("lis_array_axpyz", [ci, cd, f64p, f64p, f64p], lambda: [n, 2.5, v(), v(), v()], (4,))
        both("lis_array_scale", [ci, cd, f64p], lambda: [n, -0.3, v()], (2,))
        for name in ("lis_array_pmul", "lis_array_pdiv"):
            both(name, [ci, f64p, f64p, f64p], lambda: [n, v(), v(), v()], (3,))
        both("lis_array_set_all", [ci, cd, f64p], lambda: [n, 4.25, v()], (2,))
        for name in ("lis_array_abs", "lis_array_reciprocal", "lis_array_conjugate"):
            both(name, [ci, f64p], lambda: [n, v()], (1,))
        both("lis_array_shift", [ci, cd, f64p], lambda: [n, 0.125, v()], (2,))
        for name in ("lis_array_dot", "lis_array_nhdot"):
            both(name, [ci, f64p, f64p, f64p], lambda: [n, v(), v(), np.zeros(1)], (3,))
        for name in ("lis_array_nrm1", "lis_array_nrm2", "lis_array_nrmi", "lis_array_sum"):
            both(name, [ci, f64p, f64p], lambda: [n, v(), np.zeros(1)], (2,))
        for op in (0, 1, 2):                       # LIS_INS_VALUE, LIS_ADD_VALUE, LIS_SUB_VALUE
            for name in ("lis_array_matvec", "lis_array_matvech"):
                both(name, [ci, f64p, f64p, f64p, ci], lambda: [n, m(), v(), v(), op], (3,))
            both("lis_array_matvec_ns", [ci, ci, f64p, ci, f64p, f64p, ci], lambda: [n, n, m(n + 1)[:(n + 1) * n], n + 1, v(), v(), op], (5,))
            both("lis_array_matmat", [ci, f64p, f64p, f64p, ci], lambda: [n, m(), m(), m(), op], (3,))
            both("lis_array_matmat_ns", [ci, ci, ci, f64p, ci, f64p, ci, f64p, ci, ci],
                 lambda: [n, n, n, m(), n, m(), n, m(), n, op], (7,))
        both("lis_array_ge", [ci, f64p], lambda: [n, m()], (1,))
        both("lis_array_solve", [ci, f64p, f64p, f64p, f64p], lambda: [n, m(), v(), v(), m()], (3, 4))
        for name in ("lis_array_cgs", "lis_array_mgs"):
            both(name, [ci, f64p, f64p, f64p], lambda: [n, m(), m(), m()], (1, 2, 3))
    # QR iteration on a symmetric tridiagonal matrix (what Lanczos hands it)
    n = 5
    t = np.zeros((n, n)); t[np.arange(n), np.arange(n)] = [4, 3, 5, 2, 6]; t[np.arange(n - 1), np.arange(1, n)] = 1; t += np.triu(t, 1).T
    for lib_out in ([], []):
        pass
    outs = []
    for lib in (ref, our):
        a = t.ravel().copy(); q = np.zeros(n * n); r = np.zeros(n * n); it = C.c_int(0); er = C.c_double(0)
        lib.lis_array_qr.argtypes = [ci, f64p, f64p, f64p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        assert lib.lis_array_qr(n, a, q, r, C.byref(it), C.byref(er)) == 0
        outs.append((a, it.value, er.value))
    assert outs[0][1] == outs[1][1] and outs[0][2] == outs[1][2] and np.array_equal(outs[0][0].view(np.uint8), outs[1][0].view(np.uint8))
    assert np.allclose(np.sort(np.diag(outs[1][0].reshape(n, n))), np.linalg.eigvalsh(t))


def test_reference_fixture_testmat(b200):
    """test/testmat.mtx of the reference, when its tree is present"""
    path = "/root/reference/test/testmat.mtx"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    a, b, x = b200.input_mm(path)
    assert a["n"] == 100 and int(a["ptr"][-1]) == 460 and b is not None and x is None


def test_option_parsing_matches_reference(b200, ref_serial):
    texts = ["", "-i cg -p jacobi -maxiter 77 -tol 1e-9", "-i 4 -p 3 -ssor_omega 1.35 -print all",
             "-I BiCGSTAB -P SSOR -Print MEM", "-i gmres -restart 25 -storage ell -conv_cond nrm2_b",
             "-initx_zeros false -p none -print 2", "-i 9 -storage 5 -tol 1.0e-8 -maxiter 5"]
    for t in texts:
        assert b200.parse_options(t) == ref_serial.parse_options(t), t
    for bad in ("-i nosuchsolver", "-p nosuchprecon", "-print loud", "-storage xyz"):
        assert b200.parse_options(bad)[0] == ref_serial.parse_options(bad)[0] == 1, bad     # LIS_ERR_ILL_ARG


def test_sort_id_matches_reference(b200, ref_serial):
    """also among EQUAL keys: the order the satellites end in is a property of the reference's partition scheme
    (src/system/lis_sort.c:90-118), and CSR -> DIA / VBR keep the copy that ends last"""
    rng = np.random.default_rng(2)
    for n in (0, 1, 2, 7, 64, 65, 500):
        keys = rng.permutation(n * 3)[:n]
        vals = rng.standard_normal(n)
        gk, gv = b200.sort_id(keys, vals)
        rk, rv = ref_serial.sort_id(keys, vals)
        assert np.array_equal(gk, rk) and np.array_equal(gv, rv)
    for t in range(400):
        n = int(rng.integers(1, 300))
        keys = rng.integers(0, int(rng.integers(1, 400)), n).astype(np.int32)
        if t % 5 == 0:
            keys = np.sort(keys)
        if t % 7 == 0:
            keys = np.sort(keys)[::-1].copy()
        vals = rng.standard_normal(n)
        gk, gv = b200.sort_id(keys, vals)
        rk, rv = ref_serial.sort_id(keys, vals)
        assert np.array_equal(gk, rk) and np.array_equal(gv, rv), (t, n)


def test_error_codes_host_side(built):
    lib = lis_b200.load_library()
    lib.lis_initialize(None, None)
    A = C.c_void_p()
    assert lib.lis_matrix_create(1, C.byref(A)) == 0
    assert lib.lis_matrix_set_size(A, 5, 3) == 1                  # local > global: LIS_ERR_ILL_ARG
    assert lib.lis_matrix_set_size(A, -1, 0) == 1
    assert lib.lis_matrix_set_size(A, 0, 0) == 1
    assert lib.lis_matrix_set_size(A, 0, 8) == 0
    assert lib.lis_matrix_set_type(A, 99) == 1
    assert lib.lis_matrix_set_type(A, 5) == 0
    lib.lis_matrix_set_value.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
    assert lib.lis_matrix_set_value(0, 9, 0, 1.0, A) == 1         # row out of range
    assert lib.lis_matrix_set_value(0, 0, 0, 1.0, A) == 0
    assert lib.lis_matrix_assemble(A) == 0
    assert lib.lis_matrix_set_value(0, 1, 1, 1.0, A) == 1         # already assembled
    assert lib.lis_is_malloc(A) == 1
    assert lib.lis_matrix_destroy(A) == 0
    assert lib.lis_is_malloc(A) == 0
    v = C.c_void_p()
    assert lib.lis_vector_create(1, C.byref(v)) == 0
    assert lib.lis_vector_is_null(v) == 1
    assert lib.lis_vector_set_size(v, 0, 6) == 0
    assert lib.lis_vector_is_null(v) == 0
    lib.lis_vector_set_value.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p]
    assert lib.lis_vector_set_value(0, 6, 1.0, v) == 1            # index out of range
    assert lib.lis_vector_destroy(v) == 0


def test_public_api_coverage(built):
    """every function the reference's include/lis.h declares is exported by liblis_b200.so, except three names the reference
    declares but never defines (lis_gesolve is there for B = NULL; the generalized problem returns LIS_ERR_NOT_IMPLEMENTED)"""
    import subprocess
    hdr = "/root/reference/include/lis.h"
    if not os.path.exists(hdr):
        pytest.skip("reference tree not present")
    want = set(re.findall(r"extern [A-Za-z_ ]*\*?\s*(lis_[a-z0-9_]*)\(", open(hdr).read()))
    out = subprocess.run(["nm", "-D", os.path.join(lis_b200.LIB_DIR, "liblis_b200.so")], capture_output=True, text=True).stdout
    have = {ln.split()[2] for ln in out.splitlines() if len(ln.split()) == 3 and ln.split()[1] == "T"}
    missing = sorted(want - have)
    assert len(want) > 170
    assert missing == ["lis_iesolver_destroy", "lis_matrix_set_value_csr", "lis_matrix_set_value_new"], missing


def test_psd_update_and_vbr_partition(built):
    """lis_matrix_psd_set_value rewrites a stored entry of an assembled CSR matrix in place (INS / ADD; an entry that is
    not stored is left alone); lis_matrix_get_vbr_rowcol returns the partition lis_matrix_convert would use"""
    lib = lis_b200.load_library()
    ptr, idx, val = H.poisson1d(12)
    n = len(ptr) - 1
    A = C.c_void_p()
    libc = C.CDLL("libc.so.6"); libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]

    def dup(a):
        q = libc.malloc(max(a.nbytes, 8)); C.memmove(q, a.ctypes.data, a.nbytes); return q
    assert lib.lis_matrix_create(0, C.byref(A)) == 0 and lib.lis_matrix_set_size(A, 0, n) == 0
    lib.lis_matrix_set_csr.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.lis_matrix_set_csr(int(ptr[-1]), dup(np.asarray(ptr, np.int32)), dup(np.asarray(idx, np.int32)), dup(np.asarray(val, np.float64)), A) == 0
    assert lib.lis_matrix_assemble(A) == 0
    lib.lis_matrix_psd_set_value.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
    assert lib.lis_matrix_psd_set_value(0, 3, 3, 7.5, A) == 0            # LIS_INS_VALUE
    assert lib.lis_matrix_psd_set_value(1, 3, 4, 0.25, A) == 0           # LIS_ADD_VALUE
    assert lib.lis_matrix_psd_set_value(0, 3, 9, 1.0, A) == 0            # not stored: ignored
    assert lib.lis_matrix_psd_set_value(0, 3, 99, 1.0, A) == 1           # out of range: LIS_ERR_ILL_ARG
    d = np.zeros(n); v = C.c_void_p()
    assert lib.lis_vector_duplicate(A, C.byref(v)) == 0 and lib.lis_matrix_get_diagonal(A, v) == 0
    lib.lis_vector_gather.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(np.float64)]
    assert lib.lis_vector_gather(v, d) == 0 and d[3] == 7.5 and d[2] == 2.0
    nr, nc = C.c_int(), C.c_int(); row, col = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
    assert lib.lis_matrix_get_vbr_rowcol(A, C.byref(nr), C.byref(nc), C.byref(row), C.byref(col)) == 0
    cuts = [row[k] for k in range(nr.value + 1)]
    assert nr.value == nc.value and cuts[0] == 0 and cuts[-1] == n and cuts == sorted(set(cuts))
    lib.lis_vector_destroy(v); lib.lis_matrix_destroy(A)


def test_descriptor_hand_over_between_ranks(built):
    """host/lis_peer.c: the file descriptors of the exportable inbox blocks travel as SCM_RIGHTS messages over abstract unix
    datagram sockets, one socket per rank.  Here: two 'ranks' in one process, a pipe's read end handed from rank 0 to rank 1
    and from rank 1 to rank 0; what is written into the pipes comes out of the received descriptors."""
    import ctypes as C
    import lis_b200
    lib = lis_b200.load_library()
    lib.lisd_fd_socket.argtypes = [C.c_char_p, C.c_int]
    lib.lisd_fd_send.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.lisd_fd_recv.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    job = f"pytest-{os.getpid()}".encode()
    s0, s1 = lib.lisd_fd_socket(job, 0), lib.lisd_fd_socket(job, 1)
    assert s0 >= 0 and s1 >= 0
    assert lib.lisd_fd_socket(job, 1) < 0                      # the name is taken: a second bind must fail
    r0, w0 = os.pipe(); r1, w1 = os.pipe()
    try:
        assert lib.lisd_fd_send(s0, job, 1, 0, r0) == 1 and lib.lisd_fd_send(s1, job, 0, 1, r1) == 1
        frm, fd = C.c_int(-1), C.c_int(-1)
        assert lib.lisd_fd_recv(s1, C.byref(frm), C.byref(fd), 2000) == 1 and frm.value == 0 and fd.value not in (r0, -1)
        os.write(w0, b"from rank 0")
        assert os.read(fd.value, 64) == b"from rank 0"
        os.close(fd.value)
        assert lib.lisd_fd_recv(s0, C.byref(frm), C.byref(fd), 2000) == 1 and frm.value == 1
        os.write(w1, b"from rank 1")
        assert os.read(fd.value, 64) == b"from rank 1"
        os.close(fd.value)
        assert lib.lisd_fd_recv(s0, C.byref(frm), C.byref(fd), 50) == 0          # nothing queued: times out
        assert lib.lisd_peer_available() == 0                   # no CUDA device here: the exchange stays on the fallback
    finally:
        for f in (r0, w0, r1, w1, s0, s1):
            os.close(f)
