/*
 * lis_krylov.c -- the three Krylov drivers of the hot path: lis_cg, lis_bicgstab, lis_gmres.
 *
 * Host loops with exactly the reference's operation order and scalar arithmetic
 * (src/solver/lis_solver_cg.c:129-235, lis_solver_bicgstab.c:137-315,
 * lis_solver_gmres.c:135-343, quirks included); every vector operation is a kernel launch on
 * the library stream.  Elementwise kernels are only enqueued (lisd_*); a reduction
 * (lis_vector_dot / nrm2) is the one point per step where the host waits, because its scalar
 * feeds the next step.
 *
 * BiCGSTAB fuses its vector updates with the norms that follow them (bit-identical to the separate
 * calls: same element -> thread map and summation tree); GMRES fuses the Gram-Schmidt chain.
 * CG has a fused path (default) for CSR/ELL/... + none/Jacobi: psolve+dot, SpMV+dot and
 * axpy+axpy+nrm2 run as three kernels per iteration instead of eight.  The fused kernels
 * produce the same vector bits as the unfused sequence; dot values may differ in the last
 * bits because the summation tree differs (both are deterministic).  Set
 * LIS_B200_FUSE=0 in the environment to run the call-for-call sequence.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lis_device.h"
#include "lis_host.h"
#include "lis_b200_kernels.h"

#define CHK(e) do { LIS_INT e_ = (e); if (e_) return e_; } while (0)

static void record(LIS_SOLVER solver, LIS_INT output, LIS_INT iter, LIS_REAL nrm2)
{
    if (output) {
        if (output & LIS_PRINT_MEM) solver->rhistory[iter] = nrm2;
        if (output & LIS_PRINT_OUT) lis_host_print_rhistory(iter, nrm2);
    }
}

static int fuse_enabled(void)
{
    const char *e = getenv("LIS_B200_FUSE");
    return !(e && e[0] == '0');
}

/* ================================================================== CG */
LIS_INT lis_cg(LIS_SOLVER solver)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR x = solver->x;
    LIS_VECTOR z = solver->work[0], q = solver->work[1], r = solver->work[2], p = solver->work[3];
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER];
    const LIS_INT output = solver->options[LIS_OPTIONS_OUTPUT];
    LIS_SCALAR alpha, beta = 0.0, rho, rho_old = 1.0, dot_pq;
    LIS_REAL bnrm2, nrm2 = 0.0, tol;
    LIS_INT iter;
    double time, ptime = 0.0;
    const LIS_INT ptype = solver->precon->precon_type;
    const int fused = fuse_enabled() && solver->options[LIS_OPTIONS_CONV_COND] != LIS_CONV_COND_NRM1_B &&
                      (ptype == LIS_PRECON_TYPE_NONE || ptype == LIS_PRECON_TYPE_JACOBI);

    {
        LIS_INT e = lis_solver_get_initial_residual(solver, NULL, NULL, r, &bnrm2);
        if (e == LIS_FAILS) return LIS_SUCCESS;
        if (e) return e;
    }
    tol = solver->tol;
    CHK(lisd_set_all(0.0, p));

    /* Jacobi: the update that ends iteration k also forms z = M^-1 r and <r,z> for iteration k+1 in the
     * same pass over r (LIS_B200_CG=split: separate launches, same bits) */
    const int carry = fused && ptype == LIS_PRECON_TYPE_JACOBI && !(getenv("LIS_B200_CG") && strcmp(getenv("LIS_B200_CG"), "split") == 0);
    int have_rho = 0;
    LIS_SCALAR rho_next = 0.0;

    for (iter = 1; iter <= maxiter; iter++) {
        /* z = M^-1 r ; rho = <r,z> */
        time = lis_wtime();
        if (have_rho) {
            rho = rho_next;
            have_rho = 0;
        } else if (fused && ptype == LIS_PRECON_TYPE_JACOBI) {
            CHK(lisd_jacobi_dot(r, solver->precon->D, z, &rho));
            ptime += lis_wtime() - time;
        } else {
            CHK(lis_psolve(solver, r, z));
            ptime += lis_wtime() - time;
            CHK(lis_vector_dot(r, z, &rho));
        }
        beta = rho / rho_old;
        /* p = z + beta*p */
        CHK(lisd_xpay(z, beta, p));
        /* q = A p ; dot_pq = <p,q> */
        if (fused) CHK(lisd_matvec_dot(A, p, q, &dot_pq));
        else { CHK(lisd_matvec(A, p, q)); CHK(lis_vector_dot(p, q, &dot_pq)); }
        if (dot_pq == 0.0) {
            solver->retcode = LIS_BREAKDOWN; solver->iter = iter; solver->resid = nrm2;
            return LIS_BREAKDOWN;
        }
        alpha = rho / dot_pq;
        /* x += alpha*p ; r -= alpha*q ; nrm2 = ||r|| * bnrm */
        if (carry) {
            CHK(lisd_cg_update_jacobi(alpha, p, q, x, r, solver->precon->D, z, &nrm2, &rho_next, &have_rho));
            if (!have_rho) CHK(lisd_cg_update(alpha, p, q, x, r, &nrm2));
            nrm2 = nrm2 * solver->bnrm;
        } else if (fused) {
            CHK(lisd_cg_update(alpha, p, q, x, r, &nrm2));
            nrm2 = nrm2 * solver->bnrm;
        } else {
            CHK(lisd_axpy(alpha, p, x));
            CHK(lisd_axpy(-alpha, q, r));
            CHK(lis_host_solver_residual(solver, r, &nrm2));
        }
        record(solver, output, iter, nrm2);
        if (tol >= nrm2) {
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        rho_old = rho;
    }
    solver->retcode = LIS_MAXITER; solver->iter = iter; solver->resid = nrm2;
    return LIS_MAXITER;
}

/* ================================================================== BiCGSTAB */
LIS_INT lis_bicgstab(LIS_SOLVER solver)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR x = solver->x;
    LIS_VECTOR rtld = solver->work[0], r = solver->work[1], s = solver->work[1], t = solver->work[2];
    LIS_VECTOR p = solver->work[3], v = solver->work[4], phat = solver->work[5], shat = solver->work[6];
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER];
    const LIS_INT output = solver->options[LIS_OPTIONS_OUTPUT];
    LIS_SCALAR alpha = 1.0, beta, omega = 1.0, rho, rho_old = 1.0, tmpdot1, tmpdot2;
    LIS_REAL bnrm2, nrm2 = 0.0, tol;
    LIS_INT iter;
    double time, ptime = 0.0;
    const int fused = fuse_enabled();
    /* vector updates fused with each other and with the norm that follows (LIS_B200_BICGSTAB=calls: one
     * launch per reference call); the norm must be the 2-norm the fused kernels compute */
    const int fusedv = fused && !(getenv("LIS_B200_BICGSTAB") && strcmp(getenv("LIS_B200_BICGSTAB"), "calls") == 0);
    const int fusedn = fusedv && solver->options[LIS_OPTIONS_CONV_COND] != LIS_CONV_COND_NRM1_B;

    CHK(lisd_set_all(0.0, p));
    CHK(lisd_set_all(0.0, phat));
    CHK(lisd_set_all(0.0, s));
    CHK(lisd_set_all(0.0, shat));
    {
        LIS_INT e = lis_solver_get_initial_residual(solver, NULL, NULL, r, &bnrm2);
        if (e == LIS_FAILS) return LIS_SUCCESS;
        if (e) return e;
    }
    tol = solver->tol;
    CHK(lis_host_solver_shadow_residual(solver, r, rtld));

    for (iter = 1; iter <= maxiter; iter++) {
        CHK(lis_vector_dot(rtld, r, &rho));
        if (rho == 0.0) {
            solver->retcode = LIS_BREAKDOWN; solver->iter = iter; solver->resid = nrm2;
            return LIS_BREAKDOWN;
        }
        if (iter == 1) {
            CHK(lisd_copy(r, p));
        } else {
            beta = (rho / rho_old) * (alpha / omega);
            if (fusedv) CHK(lisd_bicgstab_p(omega, beta, v, r, p));     /* p = r + beta*(p - omega*v), one pass */
            else {
                CHK(lisd_axpy(-omega, v, p));
                CHK(lisd_xpay(r, beta, p));
            }
        }
        time = lis_wtime();
        CHK(lis_psolve(solver, p, phat));
        ptime += lis_wtime() - time;
        CHK(lisd_matvec(A, phat, v));
        CHK(lis_vector_dot(rtld, v, &tmpdot1));
        alpha = rho / tmpdot1;
        if (fusedn) {                           /* s = r - alpha*v ; ||s|| */
            CHK(lisd_axpy_nrm2(-alpha, v, r, &nrm2));
            nrm2 = nrm2 * solver->bnrm;
        } else {
            CHK(lisd_axpy(-alpha, v, r));
            CHK(lis_host_solver_residual(solver, s, &nrm2));
        }
        if (nrm2 <= tol) {
            record(solver, output, iter, nrm2);
            CHK(lisd_axpy(alpha, phat, x));
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        time = lis_wtime();
        CHK(lis_psolve(solver, s, shat));
        ptime += lis_wtime() - time;
        CHK(lisd_matvec(A, shat, t));
        if (fused) {
            LIS_SCALAR d2[2];
            CHK(lisd_dot2(t, s, d2));           /* <t,s>, <t,t> in one pass */
            tmpdot1 = d2[0]; tmpdot2 = d2[1];
        } else {
            CHK(lis_vector_dot(t, s, &tmpdot1));
            CHK(lis_vector_dot(t, t, &tmpdot2));
        }
        omega = tmpdot1 / tmpdot2;
        if (fusedn) {                           /* x += alpha*phat + omega*shat ; r -= omega*t ; ||r|| */
            CHK(lisd_bicgstab_update(alpha, omega, phat, shat, t, x, r, &nrm2));
            nrm2 = nrm2 * solver->bnrm;
        } else {
            CHK(lisd_axpy(alpha, phat, x));
            CHK(lisd_axpy(omega, shat, x));
            CHK(lisd_axpy(-omega, t, r));
            CHK(lis_host_solver_residual(solver, r, &nrm2));
        }
        record(solver, output, iter, nrm2);
        if (tol >= nrm2) {
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        if (omega == 0.0) {
            solver->retcode = LIS_BREAKDOWN; solver->iter = iter; solver->resid = nrm2;
            return LIS_BREAKDOWN;
        }
        rho_old = rho;
    }
    solver->retcode = LIS_MAXITER; solver->iter = iter; solver->resid = nrm2;
    return LIS_MAXITER;
}

/* ================================================================== modified Gram-Schmidt
 * w = v[i] against v[0..i): hcol[k] = <w,v_k>; w -= hcol[k] v_k; *nrm = ||w||_2
 * (src/solver/lis_solver_gmres.c:225-236, lis_solver_fgmres.c the same loop).
 * The reference waits for every dot.  On one rank the whole chain stays on the device: each axpy
 * reads its coefficient from the slot the dot before it wrote, and shares its pass over w with the
 * next dot (or with the norm that ends the chain): i+1 passes over w and one host wait per Krylov
 * step instead of 2i+1 passes and i+1 waits.  Same arithmetic in the same order: same bits
 * (LIS_B200_MGS=chain keeps axpy and dot as separate launches, LIS_B200_FUSE=0 also the waits). */
LIS_INT lis_host_mgs(LIS_VECTOR *v, LIS_INT i, LIS_SCALAR *hcol, LIS_REAL *nrm)
{
    const char *mode = getenv("LIS_B200_MGS");
    const int on_device = fuse_enabled() && lisd_nranks() == 1 && i <= LISD_NSCALARS;
    LIS_VECTOR w = v[i];
    LIS_SCALAR t;
    LIS_INT k;
    if (on_device && !(mode && strcmp(mode, "chain") == 0)) {
        CHK(lisd_dot_to_slot(w, v[0], 0));
        for (k = 0; k + 1 < i; k++) CHK(lisd_mgs_step((int)k, -1.0, v[k], w, v[k + 1], (int)k + 1, NULL));
        CHK(lisd_dev_scalars_fetch(0, (int)i));
        CHK(lisd_mgs_step((int)i - 1, -1.0, v[i - 1], w, NULL, 0, nrm));          /* waits for the stream */
        for (k = 0; k < i; k++) hcol[k] = lisd_fetched((int)k);
    } else if (on_device) {
        for (k = 0; k < i; k++) {
            CHK(lisd_dot_to_slot(w, v[k], (int)k));
            CHK(lisd_axpy_from_slot((int)k, -1.0, v[k], w));
        }
        CHK(lisd_dev_scalars_fetch(0, (int)i));
        CHK(lis_vector_nrm2(w, nrm));                                              /* waits for the stream */
        for (k = 0; k < i; k++) hcol[k] = lisd_fetched((int)k);
    } else if (fuse_enabled() && lisd_nranks() > 1 && i >= 1 && !(mode && strcmp(mode, "chain") == 0)) {
        /* row-partitioned: every coefficient has to be combined across the ranks on the host, so each link waits --
         * but the axpy still shares its pass over w with the dot of the next link (or the closing norm): i+1 passes
         * over w instead of 2i+1, same kernels and trees as the separate calls, same bits */
        CHK(lis_vector_dot(w, v[0], &t));
        hcol[0] = t;
        for (k = 0; k + 1 < i; k++) {
            CHK(lisd_axpy_dot(-hcol[k], v[k], w, v[k + 1], &t));
            hcol[k + 1] = t;
        }
        CHK(lisd_axpy_nrm2(-hcol[i - 1], v[i - 1], w, nrm));
    } else {
        for (k = 0; k < i; k++) {
            CHK(lis_vector_dot(w, v[k], &t));
            hcol[k] = t;
            CHK(lisd_axpy(-t, v[k], w));
        }
        CHK(lis_vector_nrm2(w, nrm));
    }
    return LIS_SUCCESS;
}

/* ================================================================== GMRES(m) */
LIS_INT lis_gmres(LIS_SOLVER solver)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR b = solver->b, x = solver->x;
    LIS_VECTOR r = solver->work[1], z = solver->work[2], *v = &solver->work[3];
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER];
    const LIS_INT output = solver->options[LIS_OPTIONS_OUTPUT];
    const LIS_INT m = solver->options[LIS_OPTIONS_RESTART];
    const LIS_INT h_dim = m + 1;
    const LIS_INT cs = (m + 1) * h_dim, sn = (m + 2) * h_dim;
    LIS_SCALAR aa, bb, rr, a2, b2, t;
    LIS_REAL bnrm2, nrm2 = 0.0, tol, rnorm;
    LIS_INT iter, i, j, k, ii = 0, i1 = 0, iih, jj;
    LIS_INT err = LIS_SUCCESS;
    double time, ptime = 0.0;

    LIS_SCALAR *h = (LIS_SCALAR *)lis_malloc(sizeof(LIS_SCALAR) * (size_t)(h_dim + 1) * (size_t)(h_dim + 2), "lis_gmres::h");
    LIS_SCALAR *s = (LIS_SCALAR *)lis_calloc(sizeof(LIS_SCALAR) * (size_t)(m + 2), "lis_gmres::s");
    if (h == NULL || s == NULL) { lis_free2(2, h, s); LIS_SETERR_MEM(sizeof(LIS_SCALAR) * (h_dim + 1) * (h_dim + 2)); return LIS_ERR_OUT_OF_MEMORY; }
#define GCHK(e) do { err = (e); if (err) goto fail; } while (0)

    /* r = M^-1 (b - A x): computed into v[0] and then overwritten by the initial residual,
     * exactly like the reference (:179-184) */
    GCHK(lisd_matvec(A, x, z));
    GCHK(lisd_xpay(b, -1.0, z));
    GCHK(lis_psolve(solver, z, v[0]));
    err = lis_solver_get_initial_residual(solver, NULL, NULL, v[0], &bnrm2);
    if (err == LIS_FAILS) { lis_free2(2, h, s); return LIS_SUCCESS; }
    if (err) goto fail;
    tol = solver->tol;

    iter = 0;
    while (iter < maxiter) {
        /* v[0] = r / ||r|| ; s = ||r|| e_1 */
        GCHK(lis_vector_nrm2(v[0], &rnorm));
        GCHK(lisd_scale(1.0 / rnorm, v[0]));
        for (k = 0; k < m + 1; k++) s[k] = 0.0;
        s[0] = rnorm;
        i = 0;
        do {
            iter++; i++;
            ii = i - 1; i1 = i; iih = (i - 1) * h_dim;
            /* z = M^-1 v ; w = A z */
            time = lis_wtime();
            GCHK(lis_psolve(solver, v[ii], z));
            ptime += lis_wtime() - time;
            GCHK(lisd_matvec(A, z, v[i1]));
            /* modified Gram-Schmidt: h[k] = <w,v_k>; w -= h[k] v_k.  On one rank the i dot->axpy
             * pairs are chained on the device (the axpy reads its coefficient from the slot the dot
             * wrote) and the host collects h[0..i) with the norm that follows: one wait per Krylov
             * step instead of i+1.  Same kernels, same bits. */
            GCHK(lis_host_mgs(v, i, h + iih, &t));
            h[i1 + iih] = t;
            GCHK(lisd_scale(1.0 / t, v[i1]));
            /* Givens rotations on the new Hessenberg column */
            for (k = 1; k <= ii; k++) {
                jj = k - 1;
                t = h[jj + iih];
                aa = h[jj + cs] * t;
                aa += h[jj + sn] * h[k + iih];
                bb = -h[jj + sn] * t;
                bb += h[jj + cs] * h[k + iih];
                h[jj + iih] = aa;
                h[k + iih] = bb;
            }
            aa = h[ii + iih];
            bb = h[i1 + iih];
            a2 = aa * aa;
            b2 = bb * bb;
            rr = sqrt(a2 + b2);
            if (rr == 0.0) rr = 1.0e-17;
            h[ii + cs] = aa / rr;
            h[ii + sn] = bb / rr;
            s[i1] = -h[ii + sn] * s[ii];
            s[ii] = h[ii + cs] * s[ii];
            aa = h[ii + cs] * h[ii + iih];
            aa += h[ii + sn] * h[i1 + iih];
            h[ii + iih] = aa;
            nrm2 = fabs(s[i1]) * bnrm2;
            record(solver, output, iter, nrm2);
            if (tol >= nrm2) break;
        } while (i < m && iter < maxiter);

        /* solve H y = s (upper triangular after the rotations) */
        s[ii] = s[ii] / h[ii + iih];
        for (k = 1; k <= ii; k++) {
            jj = ii - k;
            t = s[jj];
            for (j = jj + 1; j <= ii; j++) t -= h[jj + j * h_dim] * s[j];
            s[jj] = t / h[jj + jj * h_dim];
        }
        /* z = sum_j y_j v_j  (z[k] = y_0*v0[k], then axpys; :291-300) */
        GCHK(lisd_copy(v[0], z));
        GCHK(lisd_scale(s[0], z));
        for (j = 1; j <= ii; j++) GCHK(lisd_axpy(s[j], v[j], z));
        /* r = M^-1 z ; x += r */
        time = lis_wtime();
        GCHK(lis_psolve(solver, z, r));
        ptime += lis_wtime() - time;
        GCHK(lisd_axpy(1.0, r, x));
        if (tol >= nrm2) {
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            lis_free2(2, h, s);
            return LIS_SUCCESS;
        }
        /* restart residual rebuilt from the basis (:321-333) */
        for (j = 1; j <= i; j++) {
            jj = i1 - j + 1;
            s[jj - 1] = -h[jj - 1 + sn] * s[jj];
            s[jj] = h[jj - 1 + cs] * s[jj];
        }
        for (j = 0; j <= i1; j++) {
            t = s[j];
            if (j == 0) t = t - 1.0;
            GCHK(lisd_axpy(t, v[j], v[0]));
        }
    }
    solver->retcode = LIS_MAXITER; solver->iter = iter + 1; solver->resid = nrm2;
    lis_free2(2, h, s);
    return LIS_MAXITER;
fail:
    lis_free2(2, h, s);
    return err;
#undef GCHK
}

/* ================================================================== BiCG
 * src/solver/lis_solver_bicg.c:137-290 -- the reference's DEFAULT solver (test1 testmat.mtx 0, the
 * `make check` case).  First row of SURVEY.md section 8(f): same kernels plus the transposed product.
 * q aliases z and qtld aliases ztld (work[2], work[3]) as in the reference. */
LIS_INT lis_bicg(LIS_SOLVER solver)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR x = solver->x;
    LIS_VECTOR r = solver->work[0], rtld = solver->work[1], z = solver->work[2], ztld = solver->work[3];
    LIS_VECTOR p = solver->work[4], ptld = solver->work[5], q = solver->work[2], qtld = solver->work[3];
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER];
    const LIS_INT output = solver->options[LIS_OPTIONS_OUTPUT];
    LIS_SCALAR alpha, beta, rho, rho_old = 1.0, tmpdot1;
    LIS_REAL bnrm2, nrm2 = 0.0, tol;
    LIS_INT iter;
    double time, ptime = 0.0;

    {
        LIS_INT e = lis_solver_get_initial_residual(solver, NULL, NULL, r, &bnrm2);
        if (e == LIS_FAILS) return LIS_SUCCESS;
        if (e) return e;
    }
    tol = solver->tol;
    CHK(lis_host_solver_shadow_residual(solver, r, rtld));
    CHK(lisd_set_all(0.0, p));
    CHK(lisd_set_all(0.0, ptld));

    for (iter = 1; iter <= maxiter; iter++) {
        /* z = M^-1 r ; ztld = M^-H rtld */
        time = lis_wtime();
        CHK(lis_psolve(solver, r, z));
        CHK(lis_psolveh(solver, rtld, ztld));
        ptime += lis_wtime() - time;
        CHK(lis_vector_dot(rtld, z, &rho));
        if (rho == 0.0) {
            solver->retcode = LIS_BREAKDOWN; solver->iter = iter; solver->resid = nrm2;
            return LIS_BREAKDOWN;
        }
        beta = rho / rho_old;
        CHK(lisd_xpay(z, beta, p));              /* p    = z    + beta*p    */
        CHK(lisd_matvec(A, p, q));               /* q    = A p               */
        CHK(lisd_xpay(ztld, beta, ptld));        /* ptld = ztld + beta*ptld */
        CHK(lisd_matvech(A, ptld, qtld));        /* qtld = A^H ptld          */
        CHK(lis_vector_dot(ptld, q, &tmpdot1));
        if (tmpdot1 == 0.0) {
            solver->retcode = LIS_BREAKDOWN; solver->iter = iter; solver->resid = nrm2;
            return LIS_BREAKDOWN;
        }
        alpha = rho / tmpdot1;
        CHK(lisd_axpy(alpha, p, x));
        CHK(lisd_axpy(-alpha, q, r));
        CHK(lis_host_solver_residual(solver, r, &nrm2));
        record(solver, output, iter, nrm2);
        if (tol >= nrm2) {
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        CHK(lisd_axpy(-alpha, qtld, rtld));      /* rtld -= conj(alpha) qtld */
        rho_old = rho;
    }
    solver->retcode = LIS_MAXITER; solver->iter = iter; solver->resid = nrm2;
    return LIS_MAXITER;
}
