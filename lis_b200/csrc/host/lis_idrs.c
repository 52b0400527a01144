/*
 * lis_idrs.c -- IDR(s) and IDR(1) (src/solver/lis_solver_idrs.c:223, :526 of the reference) on the
 * lis_b200 kernels, plus what they need around them: the MT19937 generator that fills the shadow
 * space P (the reference seeds it with init_by_array {0x123,0x234,0x345,0x456}, so P is the same
 * numbers here), its Gram-Schmidt orthonormalisation and the small dense solve M c = m.
 *
 * The reference updates dX/dR with hand-written element loops  h = om*av[i]; h -= dX_j[i]*c_j ...
 * Those are expressed with the elementwise kernels (scale, then one axpy per j, then copy): the
 * same multiplications and subtractions in the same order, so the bits agree.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include "lis_device.h"
#include "lis_host.h"

/* ------------------------------------------------------------------ MT19937 (Matsumoto & Nishimura, 2002) */
#define MT_N 624
#define MT_M 397
static uint32_t mt[MT_N];
static int mti = MT_N + 1;

static void mt_seed(uint32_t s)
{
    mt[0] = s;
    for (mti = 1; mti < MT_N; mti++) mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
}

static void mt_seed_array(const uint32_t *key, int len)
{
    int i = 1, j = 0, k;
    mt_seed(19650218u);
    for (k = MT_N > len ? MT_N : len; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = MT_N - 1; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
    }
    mt[0] = 0x80000000u;
}

static uint32_t mt_next(void)
{
    static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
    uint32_t y;
    if (mti >= MT_N) {
        int kk;
        if (mti == MT_N + 1) mt_seed(5489u);
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < MT_N - 1; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
        mti = 0;
    }
    y = mt[mti++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

static double mt_real1(void) { return mt_next() * (1.0 / 4294967295.0); }      /* [0,1] */

/* ------------------------------------------------------------------ helpers */
#define CHK(e) do { LIS_INT e_ = (e); if (e_) return e_; } while (0)
#define W(k) (solver->work[k])

/* the first s*n uniforms of the reference's seed, n per vector, written into P[0..s) */
LIS_INT lis_host_fill_mt19937(LIS_INT s, LIS_INT n, LIS_VECTOR *P)
{
    static const uint32_t key[4] = {0x123, 0x234, 0x345, 0x456};
    LIS_SCALAR *buf = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(n > 0 ? n : 1));
    if (!buf) { LIS_SETERR_MEM(n * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    mt_seed_array(key, 4);
    for (LIS_INT k = 0; k < s; k++) {
        for (LIS_INT i = 0; i < n; i++) buf[i] = mt_real1();
        LIS_INT err = n > 0 ? lis_vector_set_values2(LIS_INS_VALUE, P[k]->is + P[k]->origin, n, buf, P[k]) : LIS_SUCCESS;
        if (err) { free(buf); return err; }
    }
    free(buf);
    return LIS_SUCCESS;
}

/* P[k][i] = uniform numbers in generation order k-major, then orthonormalised (lis_idrs_orth) */
static LIS_INT shadow_space(LIS_INT s, LIS_INT n, LIS_VECTOR *P)
{
    CHK(lis_host_fill_mt19937(s, n, P));
    for (LIS_INT j = 0; j < s; j++) {
        LIS_REAL r;
        LIS_SCALAR d;
        CHK(lis_vector_nrm2(P[j], &r));
        r = 1.0 / r;
        CHK(lisd_scale(r, P[j]));
        for (LIS_INT i = j + 1; i < s; i++) {
            CHK(lis_vector_dot(P[j], P[i], &d));
            CHK(lisd_axpy(-d, P[j], P[i]));
        }
    }
    return LIS_SUCCESS;
}

/* x = A^-1 b for a small column-major matrix, LU without pivoting with reciprocal pivots:
 * lis_array_solve, src/array/lis_array.c:960 (the 1x1 and 2x2 cases are spelled out there) */
static void small_solve(LIS_INT n, const LIS_SCALAR *a, const LIS_SCALAR *b, LIS_SCALAR *x, LIS_SCALAR *w)
{
    LIS_INT i, j, k;
    LIS_SCALAR t;
    for (i = 0; i < n * n; i++) w[i] = a[i];
    if (n == 1) { x[0] = b[0] / w[0]; return; }
    if (n == 2) {
        w[0] = 1.0 / w[0];
        w[1] *= w[0];
        w[3] -= w[1] * w[2];
        w[3] = 1.0 / w[3];
        x[0] = b[0];
        x[1] = b[1] - w[1] * x[0];
        x[1] *= w[3];
        x[0] -= w[2] * x[1];
        x[0] *= w[0];
        return;
    }
    for (k = 0; k < n; k++) {
        w[k + k * n] = 1.0 / w[k + k * n];
        for (i = k + 1; i < n; i++) {
            t = w[i + k * n] * w[k + k * n];
            for (j = k + 1; j < n; j++) w[i + j * n] -= t * w[k + j * n];
            w[i + k * n] = t;
        }
    }
    for (i = 0; i < n; i++) {
        x[i] = b[i];
        for (j = 0; j < i; j++) x[i] -= w[i + j * n] * x[j];
    }
    for (i = n - 1; i >= 0; i--) {
        for (j = i + 1; j < n; j++) x[i] -= w[i + j * n] * x[j];
        x[i] *= w[i + i * n];
    }
}

#define RECORD_AT(it) do { if (output) { if (output & LIS_PRINT_MEM) solver->rhistory[it] = nrm2; \
                                         if (output & LIS_PRINT_OUT) lis_host_print_rhistory(it, nrm2); } } while (0)

/* ================================================================== IDR(s) */
LIS_INT lis_idrs(LIS_SOLVER solver)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR x = solver->x;
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER], output = solver->options[LIS_OPTIONS_OUTPUT];
    const LIS_INT s = solver->options[LIS_OPTIONS_IDRS_RESTART], n = A->n;
    LIS_VECTOR r = W(0), t = W(1), v = W(2), av = W(3), *dX = &W(4), *P = &W(4 + s), *dR = &W(4 + 2 * s);
    LIS_SCALAR om = 0.0, h;
    LIS_REAL bnrm2, nrm2 = 0.0, tol;
    LIS_INT i, j, k, oldest, iter, err;
    double ptime = 0.0;
    LIS_SCALAR *buf = (LIS_SCALAR *)lis_calloc(sizeof(LIS_SCALAR) * (size_t)(2 * s + 2 * s * s + 4), "lis_idrs::buf");
    if (!buf) { LIS_SETERR_MEM(sizeof(LIS_SCALAR) * (2 * s + 2 * s * s)); return LIS_ERR_OUT_OF_MEMORY; }
    LIS_SCALAR *m = buf, *c = m + s, *M = c + s, *MM = M + s * s;
#define ICHK(e) do { err = (e); if (err) { lis_free(buf); return err; } } while (0)
#define IPSOLVE(b_, x_) do { const double t0_ = lis_wtime(); ICHK(lis_psolve(solver, b_, x_)); ptime += lis_wtime() - t0_; } while (0)
    err = lis_solver_get_initial_residual(solver, NULL, NULL, r, &bnrm2);
    if (err) { lis_free(buf); return err == LIS_FAILS ? LIS_SUCCESS : err; }
    tol = solver->tol;
    ICHK(shadow_space(s, n, P));
    /* s start-up steps of minimal residual type */
    for (k = 0; k < s; k++) {
        IPSOLVE(r, dX[k]);
        ICHK(lisd_matvec(A, dX[k], dR[k]));
        ICHK(lis_vector_dot(dR[k], dR[k], &h));
        ICHK(lis_vector_dot(dR[k], r, &om));
        om = om / h;
        ICHK(lisd_scale(om, dX[k]));
        ICHK(lisd_scale(-om, dR[k]));
        ICHK(lisd_axpy(1.0, dX[k], x));
        ICHK(lisd_axpy(1.0, dR[k], r));
        ICHK(lis_host_solver_residual(solver, r, &nrm2));
        RECORD_AT(k + 1);
        if (tol >= nrm2) {
            lis_free(buf);
            solver->retcode = LIS_SUCCESS; solver->iter = k + 1; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        for (i = 0; i < s; i++) ICHK(lis_vector_dot(P[i], dR[k], &M[k * s + i]));
    }
    iter = s;
    oldest = 0;
    for (i = 0; i < s; i++) ICHK(lis_vector_dot(P[i], r, &m[i]));
    while (iter <= maxiter) {
        small_solve(s, M, m, c, MM);                                  /* M c = m */
        ICHK(lisd_copy(r, v));
        for (j = 0; j < s; j++) ICHK(lisd_axpy(-c[j], dR[j], v));     /* v = r - dR c */
        if ((iter % (s + 1)) == s) {
            IPSOLVE(v, av);
            ICHK(lisd_matvec(A, av, t));
            ICHK(lis_vector_dot(t, t, &h));
            ICHK(lis_vector_dot(t, v, &om));
            om = om / h;
            ICHK(lisd_scale(om, av));                                 /* dX_old = om*av - dX c  */
            for (j = 0; j < s; j++) ICHK(lisd_axpy(-c[j], dX[j], av));
            ICHK(lisd_scale(-om, t));                                 /* dR_old = -om*t - dR c  */
            for (j = 0; j < s; j++) ICHK(lisd_axpy(-c[j], dR[j], t));
            ICHK(lisd_copy(av, dX[oldest]));
            ICHK(lisd_copy(t, dR[oldest]));
        } else {
            IPSOLVE(v, av);
            ICHK(lisd_scale(om, av));
            for (j = 0; j < s; j++) ICHK(lisd_axpy(-c[j], dX[j], av));
            ICHK(lisd_copy(av, dX[oldest]));
            ICHK(lisd_matvec(A, dX[oldest], dR[oldest]));
            ICHK(lisd_scale(-1.0, dR[oldest]));
        }
        ICHK(lisd_axpy(1.0, dR[oldest], r));
        ICHK(lisd_axpy(1.0, dX[oldest], x));
        iter++;
        ICHK(lis_host_solver_residual(solver, r, &nrm2));
        RECORD_AT(iter);
        if (tol >= nrm2) {
            lis_free(buf);
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        for (i = 0; i < s; i++) {
            ICHK(lis_vector_dot(P[i], dR[oldest], &h));
            m[i] += h;
            M[oldest * s + i] = h;
        }
        oldest++;
        if (oldest == s) oldest = 0;
    }
    lis_free(buf);
    solver->retcode = LIS_MAXITER; solver->iter = iter; solver->resid = nrm2;
    return LIS_MAXITER;
#undef ICHK
#undef IPSOLVE
}

/* ================================================================== IDR(1): two steps per sweep */
LIS_INT lis_idr1(LIS_SOLVER solver)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR x = solver->x;
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER], output = solver->options[LIS_OPTIONS_OUTPUT];
    const LIS_INT n = A->n;
    LIS_VECTOR r = W(0), t = W(1), v = W(2), av = W(3), *P = &W(4), *dX = &W(5), *dR = &W(6);
    LIS_SCALAR om, h, M, m, c;
    LIS_REAL bnrm2, nrm2 = 0.0, tol;
    LIS_INT iter;
    double ptime = 0.0;
#define PSOLVE1(b_, x_) do { const double t0_ = lis_wtime(); CHK(lis_psolve(solver, b_, x_)); ptime += lis_wtime() - t0_; } while (0)
    {
        LIS_INT e = lis_solver_get_initial_residual(solver, NULL, NULL, r, &bnrm2);
        if (e == LIS_FAILS) return LIS_SUCCESS;
        if (e) return e;
    }
    tol = solver->tol;
    CHK(shadow_space(1, n, P));
    PSOLVE1(r, dX[0]);
    CHK(lisd_matvec(A, dX[0], dR[0]));
    CHK(lis_vector_dot(dR[0], dR[0], &h));
    CHK(lis_vector_dot(dR[0], r, &om));
    om = om / h;
    CHK(lisd_scale(om, dX[0]));
    CHK(lisd_scale(-om, dR[0]));
    CHK(lisd_axpy(1.0, dX[0], x));
    CHK(lisd_axpy(1.0, dR[0], r));
    CHK(lis_host_solver_residual(solver, r, &nrm2));
    RECORD_AT(1);
    if (tol >= nrm2) {
        solver->retcode = LIS_SUCCESS; solver->iter = 1; solver->resid = nrm2; solver->ptime = ptime;
        return LIS_SUCCESS;
    }
    CHK(lis_vector_dot(P[0], dR[0], &M));
    iter = 1;
    CHK(lis_vector_dot(P[0], r, &m));
    while (iter <= maxiter) {
        /* first half: new omega */
        c = m / M;
        CHK(lisd_axpyz(-c, dR[0], r, v));                /* v = r - c*dR */
        PSOLVE1(v, av);
        CHK(lisd_matvec(A, av, t));
        CHK(lis_vector_dot(t, t, &h));
        CHK(lis_vector_dot(t, v, &om));
        om = om / h;
        CHK(lisd_scale(om, av));                         /* dX = om*av - c*dX */
        CHK(lisd_axpy(-c, dX[0], av));
        CHK(lisd_copy(av, dX[0]));
        CHK(lisd_scale(-om, t));                         /* dR = -om*t - c*dR */
        CHK(lisd_axpy(-c, dR[0], t));
        CHK(lisd_copy(t, dR[0]));
        CHK(lisd_axpy(1.0, dR[0], r));
        CHK(lisd_axpy(1.0, dX[0], x));
        iter++;
        CHK(lis_host_solver_residual(solver, r, &nrm2));
        RECORD_AT(iter);
        if (tol >= nrm2) {
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        CHK(lis_vector_dot(P[0], dR[0], &h));
        m += h;
        M = h;
        /* second half: same omega */
        c = m / M;
        CHK(lisd_axpyz(-c, dR[0], r, v));
        PSOLVE1(v, av);
        CHK(lisd_scale(om, av));
        CHK(lisd_axpy(-c, dX[0], av));
        CHK(lisd_copy(av, dX[0]));
        CHK(lisd_matvec(A, dX[0], dR[0]));
        CHK(lisd_scale(-1.0, dR[0]));
        CHK(lisd_axpy(1.0, dR[0], r));
        CHK(lisd_axpy(1.0, dX[0], x));
        iter++;
        CHK(lis_host_solver_residual(solver, r, &nrm2));
        RECORD_AT(iter);
        if (tol >= nrm2) {
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        CHK(lis_vector_dot(P[0], dR[0], &h));
        m += h;
        M = h;
    }
    solver->retcode = LIS_MAXITER; solver->iter = iter; solver->resid = nrm2;
    return LIS_MAXITER;
#undef PSOLVE1
}
