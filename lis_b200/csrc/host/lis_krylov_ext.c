/*
 * lis_krylov_ext.c -- further Krylov drivers of the reference on the same kernels
 * (SURVEY.md section 8(f), row 1): host loops only, every vector operation one of the CUDA
 * kernels behind lis_b200_kernels.h.  Operation order and scalar arithmetic follow
 *   CGS  src/solver/lis_solver_cgs.c:134     CRS  lis_solver_cgs.c:805
 *   CR   src/solver/lis_solver_cg.c:821      COCG lis_solver_cg.c:632     COCR lis_solver_cg.c:1155
 *   BiCR src/solver/lis_solver_bicg.c:788
 * so that, on the mock device of tests/hostcheck (sequential reductions), each reproduces the
 * serial reference bit for bit -- iteration count, residual history, solution.
 * Real scalars: conj() is the identity, lis_vector_nhdot == lis_vector_dot.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lis_device.h"
#include "lis_host.h"

#define CHK(e) do { LIS_INT e_ = (e); if (e_) return e_; } while (0)
#define W(k) (solver->work[k])

/* shared epilogues */
#define STOP(code) do { solver->retcode = (code); solver->iter = iter; solver->resid = nrm2; return (code); } while (0)
#define BREAKDOWN_IF(cond) do { if (cond) STOP(LIS_BREAKDOWN); } while (0)
#define RECORD() do { if (output) { if (output & LIS_PRINT_MEM) solver->rhistory[iter] = nrm2; \
                                    if (output & LIS_PRINT_OUT) lis_host_print_rhistory(iter, nrm2); } } while (0)
#define CONVERGED_IF_TOL() do { if (tol >= nrm2) { solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; \
                                                   solver->ptime = ptime; return LIS_SUCCESS; } } while (0)
#define PSOLVE(b, x) do { const double t0_ = lis_wtime(); CHK(lis_psolve(solver, b, x)); ptime += lis_wtime() - t0_; } while (0)
#define PSOLVEH(b, x) do { const double t0_ = lis_wtime(); CHK(lis_psolveh(solver, b, x)); ptime += lis_wtime() - t0_; } while (0)
#define RESIDUAL(r) CHK(lis_host_solver_residual(solver, r, &nrm2))
#define INITIAL_RESIDUAL(r)                                                                  \
    do {                                                                                     \
        LIS_INT e_ = lis_solver_get_initial_residual(solver, NULL, NULL, r, &bnrm2);         \
        if (e_ == LIS_FAILS) return LIS_SUCCESS;                                             \
        if (e_) return e_;                                                                   \
        tol = solver->tol;                                                                   \
    } while (0)
#define COMMON_LOCALS \
    LIS_MATRIX A = solver->A; LIS_VECTOR x = solver->x; \
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER], output = solver->options[LIS_OPTIONS_OUTPUT]; \
    LIS_REAL bnrm2, nrm2 = 0.0, tol; LIS_INT iter; double ptime = 0.0; (void)bnrm2

/* ================================================================== CGS */
LIS_INT lis_cgs(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    LIS_VECTOR r = W(0), rtld = W(1), p = W(2), phat = W(3), q = W(4), qhat = W(5), u = W(5), uhat = W(6), vhat = W(6);
    LIS_SCALAR alpha, beta, rho, rho_old = 1.0, tmpdot1;
    INITIAL_RESIDUAL(r);
    CHK(lis_host_solver_shadow_residual(solver, r, rtld));
    CHK(lisd_set_all(0.0, q));
    CHK(lisd_set_all(0.0, p));
    for (iter = 1; iter <= maxiter; iter++) {
        CHK(lis_vector_dot(rtld, r, &rho));
        BREAKDOWN_IF(rho == 0.0);
        beta = rho / rho_old;
        CHK(lisd_axpyz(beta, q, r, u));          /* u = r + beta*q              */
        CHK(lisd_xpay(q, beta, p));              /* p = u + beta*(q + beta*p)   */
        CHK(lisd_xpay(u, beta, p));
        PSOLVE(p, phat);
        CHK(lisd_matvec(A, phat, vhat));
        CHK(lis_vector_dot(rtld, vhat, &tmpdot1));
        BREAKDOWN_IF(tmpdot1 == 0.0);
        alpha = rho / tmpdot1;
        CHK(lisd_axpyz(-alpha, vhat, u, q));     /* q = u - alpha*vhat          */
        CHK(lisd_axpyz(1.0, u, q, phat));        /* phat = u + q                */
        PSOLVE(phat, uhat);
        CHK(lisd_axpy(alpha, uhat, x));
        CHK(lisd_matvec(A, uhat, qhat));
        CHK(lisd_axpy(-alpha, qhat, r));
        RESIDUAL(r);
        RECORD();
        CONVERGED_IF_TOL();
        rho_old = rho;
    }
    STOP(LIS_MAXITER);
}

/* ================================================================== CRS */
LIS_INT lis_crs(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    LIS_VECTOR r = W(0), rtld = W(1), p = W(2), z = W(3), u = W(3), uq = W(3), q = W(4), ap = W(4), map = W(5), auq = W(5);
    LIS_SCALAR alpha, beta, rho, rho_old, tmpdot1;
    INITIAL_RESIDUAL(r);
    CHK(lis_host_solver_shadow_residual(solver, r, p));
    CHK(lisd_matvech(A, p, rtld));               /* rtld = A^H r0 */
    rho_old = 1.0;
    CHK(lisd_set_all(0.0, q));
    CHK(lisd_set_all(0.0, p));
    for (iter = 1; iter <= maxiter; iter++) {
        PSOLVE(r, z);
        CHK(lis_vector_dot(rtld, z, &rho));
        BREAKDOWN_IF(rho == 0.0);
        beta = rho / rho_old;
        CHK(lisd_axpyz(beta, q, z, u));
        CHK(lisd_xpay(q, beta, p));
        CHK(lisd_xpay(u, beta, p));
        CHK(lisd_matvec(A, p, ap));
        PSOLVE(ap, map);
        CHK(lis_vector_dot(rtld, map, &tmpdot1));
        BREAKDOWN_IF(tmpdot1 == 0.0);
        alpha = rho / tmpdot1;
        CHK(lisd_axpyz(-alpha, map, u, q));
        CHK(lisd_axpyz(1.0, u, q, uq));
        CHK(lisd_matvec(A, uq, auq));
        CHK(lisd_axpy(alpha, uq, x));
        CHK(lisd_axpy(-alpha, auq, r));
        RESIDUAL(r);
        RECORD();
        CONVERGED_IF_TOL();
        rho_old = rho;
    }
    STOP(LIS_MAXITER);
}

/* ================================================================== CR and COCR (identical for real scalars) */
LIS_INT lis_cr(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    LIS_VECTOR z = W(0), q = W(1), r = W(2), p = W(3), qtld = W(4), az = W(5);
    LIS_SCALAR alpha, beta, rho, dot_rq, dot_zq;
    INITIAL_RESIDUAL(r);
    PSOLVE(r, p);
    CHK(lisd_matvec(A, p, q));
    CHK(lisd_copy(p, z));
    for (iter = 1; iter <= maxiter; iter++) {
        PSOLVE(q, qtld);
        CHK(lis_vector_dot(qtld, q, &rho));
        BREAKDOWN_IF(rho == 0.0);
        CHK(lis_vector_dot(r, qtld, &dot_rq));
        alpha = dot_rq / rho;
        CHK(lisd_axpy(alpha, p, x));
        CHK(lisd_axpy(-alpha, q, r));
        RESIDUAL(r);
        RECORD();
        CONVERGED_IF_TOL();
        CHK(lisd_axpy(-alpha, qtld, z));
        CHK(lisd_matvec(A, z, az));
        CHK(lis_vector_dot(az, qtld, &dot_zq));
        beta = -dot_zq / rho;
        CHK(lisd_xpay(z, beta, p));
        CHK(lisd_xpay(az, beta, q));
    }
    STOP(LIS_MAXITER);
}

LIS_INT lis_cocr(LIS_SOLVER solver) { return lis_cr(solver); }

/* ================================================================== COCG (CG with the unconjugated dot product) */
LIS_INT lis_cocg(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    LIS_VECTOR z = W(0), q = W(1), r = W(2), p = W(3);
    LIS_SCALAR alpha, beta, rho, rho_old = 1.0, dot_pq;
    INITIAL_RESIDUAL(r);
    CHK(lisd_set_all(0.0, p));
    for (iter = 1; iter <= maxiter; iter++) {
        PSOLVE(r, z);
        CHK(lis_vector_nhdot(r, z, &rho));
        beta = rho / rho_old;
        CHK(lisd_xpay(z, beta, p));
        CHK(lisd_matvec(A, p, q));
        CHK(lis_vector_nhdot(p, q, &dot_pq));
        BREAKDOWN_IF(dot_pq == 0.0);
        alpha = rho / dot_pq;
        CHK(lisd_axpy(alpha, p, x));
        CHK(lisd_axpy(-alpha, q, r));
        RESIDUAL(r);
        RECORD();
        CONVERGED_IF_TOL();
        rho_old = rho;
    }
    STOP(LIS_MAXITER);
}

/* ================================================================== BiCR */
LIS_INT lis_bicr(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    LIS_VECTOR r = W(0), rtld = W(1), z = W(2), ztld = W(3), p = W(4), ptld = W(5), ap = W(6), az = W(7), map = W(8), aptld = W(9);
    LIS_SCALAR alpha, beta, rho, rho_old, tmpdot1;
    INITIAL_RESIDUAL(r);
    CHK(lis_host_solver_shadow_residual(solver, r, rtld));
    CHK(lis_psolve(solver, r, z));
    CHK(lis_psolveh(solver, rtld, ztld));
    CHK(lisd_copy(z, p));
    CHK(lisd_copy(ztld, ptld));
    CHK(lisd_matvec(A, z, ap));
    CHK(lis_vector_dot(ztld, ap, &rho_old));
    for (iter = 1; iter <= maxiter; iter++) {
        CHK(lisd_matvech(A, ptld, aptld));
        PSOLVE(ap, map);
        CHK(lis_vector_dot(aptld, map, &tmpdot1));
        BREAKDOWN_IF(tmpdot1 == 0.0);
        alpha = rho_old / tmpdot1;
        CHK(lisd_axpy(alpha, p, x));
        CHK(lisd_axpy(-alpha, ap, r));
        RESIDUAL(r);
        RECORD();
        CONVERGED_IF_TOL();
        CHK(lisd_axpy(-alpha, aptld, rtld));
        CHK(lisd_axpy(-alpha, map, z));
        PSOLVEH(rtld, ztld);
        CHK(lisd_matvec(A, z, az));
        CHK(lis_vector_dot(ztld, az, &rho));
        BREAKDOWN_IF(rho == 0.0);
        beta = rho / rho_old;
        CHK(lisd_xpay(z, beta, p));
        CHK(lisd_xpay(ztld, beta, ptld));
        CHK(lisd_xpay(az, beta, ap));
        rho_old = rho;
    }
    STOP(LIS_MAXITER);
}

/* ================================================================== BiCRSTAB   src/solver/lis_solver_bicgstab.c:951 */
LIS_INT lis_bicrstab(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    LIS_VECTOR rtld = W(0), r = W(1), s = W(2), ms = W(3), ams = W(4), p = W(5), ap = W(6), map = W(7), z = W(8);
    LIS_SCALAR alpha, beta, omega, rho, rho_old, tmpdot1, tmpdot2;
    INITIAL_RESIDUAL(r);
    CHK(lis_host_solver_shadow_residual(solver, r, p));
    CHK(lisd_matvech(A, p, rtld));               /* rtld = A^H r0 */
    PSOLVE(r, z);
    CHK(lisd_copy(z, p));
    CHK(lis_vector_dot(rtld, z, &rho_old));
    for (iter = 1; iter <= maxiter; iter++) {
        CHK(lisd_matvec(A, p, ap));
        PSOLVE(ap, map);
        CHK(lis_vector_dot(rtld, map, &tmpdot1));
        alpha = rho_old / tmpdot1;
        CHK(lisd_axpyz(-alpha, ap, r, s));       /* s = r - alpha*ap */
        RESIDUAL(s);
        if (nrm2 <= tol) {                        /* early exit on the half step */
            RECORD();
            CHK(lisd_axpy(alpha, p, x));
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        CHK(lisd_axpyz(-alpha, map, z, ms));     /* ms = z - alpha*map */
        CHK(lisd_matvec(A, ms, ams));
        CHK(lis_vector_dot(ams, s, &tmpdot1));
        CHK(lis_vector_dot(ams, ams, &tmpdot2));
        omega = tmpdot1 / tmpdot2;
        CHK(lisd_axpy(alpha, p, x));
        CHK(lisd_axpy(omega, ms, x));
        CHK(lisd_axpyz(-omega, ams, s, r));      /* r = s - omega*ams */
        RESIDUAL(r);
        RECORD();
        CONVERGED_IF_TOL();
        PSOLVE(r, z);
        CHK(lis_vector_dot(rtld, z, &rho));
        BREAKDOWN_IF(rho == 0.0);
        beta = (rho / rho_old) * (alpha / omega);
        CHK(lisd_axpy(-omega, map, p));
        CHK(lisd_xpay(z, beta, p));
        rho_old = rho;
    }
    STOP(LIS_MAXITER);
}

/* ================================================================== TFQMR   src/solver/lis_solver_qmr.c:113
 * the residual estimate tau*sqrt(1+m)*bnrm2 replaces a computed norm; only the m == 0 half step
 * is recorded in the history */
LIS_INT lis_tfqmr(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    LIS_VECTOR r = W(0), rtld = W(1), u = W(2), p = W(3), d = W(4), t = W(5), t1 = W(6), q = W(7), v = W(8);
    LIS_SCALAR alpha, beta, rho, rhoold, s, eta = 0.0;
    LIS_REAL tau, w, wold, ww, theta = 0.0, c;
    INITIAL_RESIDUAL(r);
    CHK(lis_host_solver_shadow_residual(solver, r, rtld));
    CHK(lisd_copy(r, p));
    CHK(lisd_copy(r, u));
    CHK(lisd_set_all(0.0, d));
    PSOLVE(p, t);
    CHK(lisd_matvec(A, t, v));
    CHK(lis_vector_dot(r, rtld, &rhoold));
    CHK(lis_vector_nrm2(r, &tau));
    wold = tau;
    iter = 1;
    while (iter <= maxiter) {
        CHK(lis_vector_dot(v, rtld, &s));
        BREAKDOWN_IF(s == 0.0);
        alpha = rhoold / s;
        CHK(lisd_axpyz(-alpha, v, u, q));        /* q = u - alpha*v */
        CHK(lisd_axpyz(1.0, u, q, t));           /* t = u + q       */
        PSOLVE(t, t1);
        CHK(lisd_matvec(A, t1, v));
        CHK(lisd_axpy(-alpha, v, r));
        CHK(lis_vector_nrm2(r, &w));
        for (int m = 0; m < 2; m++) {
            if (m == 0) {
                ww = sqrt(w * wold);
                CHK(lisd_xpay(u, theta * theta * eta / alpha, d));
            } else {
                ww = w;
                CHK(lisd_xpay(q, theta * theta * eta / alpha, d));
            }
            theta = ww / tau;
            c = 1.0 / sqrt(1.0 + theta * theta);
            eta = c * c * alpha;
            tau = tau * theta * c;
            PSOLVE(d, t1);
            CHK(lisd_axpy(eta, t1, x));
            nrm2 = tau * sqrt(1.0 + m) * bnrm2;
            if (m == 0) RECORD();
            CONVERGED_IF_TOL();
        }
        CHK(lis_vector_dot(r, rtld, &rho));
        BREAKDOWN_IF(rho == 0.0);
        beta = rho / rhoold;
        CHK(lisd_axpyz(beta, q, r, u));          /* u = r + beta*q            */
        CHK(lisd_xpay(q, beta, p));              /* p = u + beta*(q + beta*p) */
        CHK(lisd_xpay(u, beta, p));
        PSOLVE(p, t1);
        CHK(lisd_matvec(A, t1, v));
        rhoold = rho;
        wold = w;
        iter++;
    }
    STOP(LIS_MAXITER);
}

/* ================================================================== GPBiCG   src/solver/lis_solver_gpbicg.c:145
 * relies on the work vectors starting at zero (mr, u, z are read before they are first written) */
static LIS_INT gpbi(LIS_SOLVER solver, int cr)      /* cr: the GPBiCR variant, lis_solver_gpbicg.c:1349 */
{
    COMMON_LOCALS;
    LIS_VECTOR rtld = W(0), r = W(1), mr = W(2), p = W(3), ap = W(4), map = W(5), t = W(6), mt = W(7), amt = W(8),
               u = W(9), y = W(10), w = W(11), z = W(12), mt_old = W(13);
    LIS_SCALAR alpha, beta = 0.0, rho, rho_old, qsi, eta, tmp, d[5];
    INITIAL_RESIDUAL(r);
    if (cr) {
        CHK(lis_host_solver_shadow_residual(solver, r, p));
        CHK(lisd_matvech(A, p, rtld));           /* rtld = A^H r0 */
        PSOLVE(r, p);
        CHK(lis_vector_dot(rtld, p, &rho_old));
    } else {
        CHK(lis_host_solver_shadow_residual(solver, r, rtld));
        PSOLVE(r, p);
        CHK(lis_vector_dot(rtld, r, &rho_old));
    }
    CHK(lisd_set_all(0.0, t));
    CHK(lisd_set_all(0.0, w));
    for (iter = 1; iter <= maxiter; iter++) {
        CHK(lisd_matvec(A, p, ap));
        PSOLVE(ap, map);
        CHK(lis_vector_dot(rtld, cr ? map : ap, &d[0]));
        BREAKDOWN_IF(d[0] == 0.0);
        alpha = rho_old / d[0];
        CHK(lisd_axpyz(-1.0, w, ap, y));         /* y = t - r - alpha*w + alpha*ap */
        CHK(lisd_xpay(t, alpha, y));
        CHK(lisd_axpy(-1.0, r, y));
        CHK(lisd_axpyz(-alpha, ap, r, t));       /* t = r - alpha*ap */
        RESIDUAL(t);
        if (nrm2 <= tol) {
            RECORD();
            CHK(lisd_axpy(alpha, p, x));
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        CHK(lisd_axpyz(-alpha, map, mr, mt));    /* mt = mr - alpha*map */
        CHK(lisd_matvec(A, mt, amt));
        CHK(lis_vector_dot(y, y, &d[0]));
        CHK(lis_vector_dot(amt, t, &d[1]));
        CHK(lis_vector_dot(y, t, &d[2]));
        CHK(lis_vector_dot(amt, y, &d[3]));
        CHK(lis_vector_dot(amt, amt, &d[4]));
        if (iter == 1) {
            qsi = d[1] / d[4];
            eta = 0.0;
        } else {
            tmp = d[4] * d[0] - d[3] * d[3];
            qsi = (d[0] * d[1] - d[2] * d[3]) / tmp;
            eta = (d[4] * d[2] - d[3] * d[1]) / tmp;
        }
        CHK(lisd_xpay(mt_old, beta, u));         /* u = qsi*map + eta*(mt_old - mr + beta*u) */
        CHK(lisd_axpy(-1.0, mr, u));
        CHK(lisd_scale(eta, u));
        CHK(lisd_axpy(qsi, map, u));
        CHK(lisd_scale(eta, z));                 /* z = qsi*mr + eta*z - alpha*u */
        CHK(lisd_axpy(qsi, mr, z));
        CHK(lisd_axpy(-alpha, u, z));
        CHK(lisd_axpy(alpha, p, x));             /* x = x + alpha*p + z */
        CHK(lisd_axpy(1.0, z, x));
        CHK(lisd_axpyz(-qsi, amt, t, r));        /* r = t - eta*y - qsi*amt */
        CHK(lisd_axpy(-eta, y, r));
        RESIDUAL(r);
        RECORD();
        CONVERGED_IF_TOL();
        PSOLVE(r, mr);
        CHK(lis_vector_dot(rtld, cr ? mr : r, &rho));
        BREAKDOWN_IF(rho == 0.0);
        beta = (rho / rho_old) * (alpha / qsi);
        CHK(lisd_axpyz(beta, ap, amt, w));       /* w = amt + beta*ap */
        CHK(lisd_axpy(-1.0, u, p));              /* p = mr + beta*(p - u) */
        CHK(lisd_xpay(mr, beta, p));
        CHK(lisd_copy(mt, mt_old));
        rho_old = rho;
    }
    STOP(LIS_MAXITER);
}

LIS_INT lis_gpbicg(LIS_SOLVER solver) { return gpbi(solver, 0); }
LIS_INT lis_gpbicr(LIS_SOLVER solver) { return gpbi(solver, 1); }

/* the 2x2 least-squares step shared by the GPBi / Safe families */
static void qsi_eta(LIS_INT iter, const LIS_SCALAR d[5], LIS_SCALAR *qsi, LIS_SCALAR *eta)
{
    if (iter == 1) {
        *qsi = d[1] / d[4];
        *eta = 0.0;
    } else {
        const LIS_SCALAR tmp = d[4] * d[0] - d[3] * d[3];
        *qsi = (d[0] * d[1] - d[2] * d[3]) / tmp;
        *eta = (d[4] * d[2] - d[3] * d[1]) / tmp;
    }
}

/* ================================================================== BiCGSafe   src/solver/lis_solver_bicgsafe.c:145
 * y, u, z start at zero (fresh work vectors) */
LIS_INT lis_bicgsafe(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    LIS_VECTOR rtld = W(0), r = W(1), mr = W(2), amr = W(3), p = W(4), ap = W(5), t = W(6), mt = W(7), y = W(8), u = W(9),
               z = W(10), au = W(11);
    LIS_SCALAR alpha, beta = 0.0, rho, rho_old, qsi, eta, d[5];
    INITIAL_RESIDUAL(r);
    CHK(lis_host_solver_shadow_residual(solver, r, rtld));
    PSOLVE(r, mr);
    CHK(lisd_matvec(A, mr, amr));
    CHK(lis_vector_dot(rtld, r, &rho_old));
    CHK(lisd_copy(amr, ap));
    CHK(lisd_copy(mr, p));
    for (iter = 1; iter <= maxiter; iter++) {
        CHK(lis_vector_dot(rtld, ap, &d[0]));
        alpha = rho_old / d[0];
        CHK(lis_vector_dot(y, y, &d[0]));
        CHK(lis_vector_dot(amr, r, &d[1]));
        CHK(lis_vector_dot(y, r, &d[2]));
        CHK(lis_vector_dot(amr, y, &d[3]));
        CHK(lis_vector_dot(amr, amr, &d[4]));
        qsi_eta(iter, d, &qsi, &eta);
        CHK(lisd_copy(y, t));                    /* t = qsi*ap + eta*y ; mt = M^-1 t */
        CHK(lisd_scale(eta, t));
        CHK(lisd_axpy(qsi, ap, t));
        PSOLVE(t, mt);
        CHK(lisd_xpay(mt, eta * beta, u));       /* u = mt + eta*beta*u */
        CHK(lisd_matvec(A, u, au));
        CHK(lisd_scale(eta, z));                 /* z = qsi*mr + eta*z - alpha*u */
        CHK(lisd_axpy(qsi, mr, z));
        CHK(lisd_axpy(-alpha, u, z));
        CHK(lisd_scale(eta, y));                 /* y = qsi*amr + eta*y - alpha*au */
        CHK(lisd_axpy(qsi, amr, y));
        CHK(lisd_axpy(-alpha, au, y));
        CHK(lisd_axpy(alpha, p, x));
        CHK(lisd_axpy(1.0, z, x));
        CHK(lisd_axpy(-alpha, ap, r));
        CHK(lisd_axpy(-1.0, y, r));
        RESIDUAL(r);
        RECORD();
        CONVERGED_IF_TOL();
        CHK(lis_vector_dot(rtld, r, &rho));
        BREAKDOWN_IF(rho == 0.0);
        beta = (rho / rho_old) * (alpha / qsi);
        PSOLVE(r, mr);
        CHK(lisd_matvec(A, mr, amr));
        CHK(lisd_axpy(-1.0, u, p));              /* p  = mr  + beta*(p  - u)  */
        CHK(lisd_xpay(mr, beta, p));
        CHK(lisd_axpy(-1.0, au, ap));            /* ap = amr + beta*(ap - au) */
        CHK(lisd_xpay(amr, beta, ap));
        rho_old = rho;
    }
    STOP(LIS_MAXITER);
}

/* ================================================================== BiCRSafe   src/solver/lis_solver_bicgsafe.c:1048 */
LIS_INT lis_bicrsafe(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    LIS_VECTOR rtld = W(0), r = W(1), mr = W(2), amr = W(3), p = W(4), ap = W(5), map = W(6), my = W(7), y = W(8), u = W(9),
               z = W(10), au = W(11), artld = W(12);
    LIS_SCALAR alpha, beta = 0.0, rho, rho_old, qsi, eta, d[5];
    INITIAL_RESIDUAL(r);
    CHK(lis_host_solver_shadow_residual(solver, r, rtld));
    CHK(lisd_matvech(A, rtld, artld));
    PSOLVE(r, mr);
    CHK(lisd_matvec(A, mr, amr));
    CHK(lis_vector_dot(rtld, amr, &rho_old));
    CHK(lisd_copy(amr, ap));
    CHK(lisd_copy(mr, p));
    for (iter = 1; iter <= maxiter; iter++) {
        PSOLVE(ap, map);
        CHK(lis_vector_dot(artld, map, &d[0]));
        alpha = rho_old / d[0];
        CHK(lis_vector_dot(y, y, &d[0]));
        CHK(lis_vector_dot(amr, r, &d[1]));
        CHK(lis_vector_dot(y, r, &d[2]));
        CHK(lis_vector_dot(amr, y, &d[3]));
        CHK(lis_vector_dot(amr, amr, &d[4]));
        qsi_eta(iter, d, &qsi, &eta);
        CHK(lisd_scale(eta * beta, u));          /* u = qsi*map + eta*my + eta*beta*u */
        CHK(lisd_axpy(qsi, map, u));
        CHK(lisd_axpy(eta, my, u));
        CHK(lisd_matvec(A, u, au));
        CHK(lisd_scale(eta, z));
        CHK(lisd_axpy(qsi, mr, z));
        CHK(lisd_axpy(-alpha, u, z));
        CHK(lisd_scale(eta, y));
        CHK(lisd_axpy(qsi, amr, y));
        CHK(lisd_axpy(-alpha, au, y));
        PSOLVE(y, my);
        CHK(lisd_axpy(alpha, p, x));
        CHK(lisd_axpy(1.0, z, x));
        CHK(lisd_axpy(-alpha, ap, r));
        CHK(lisd_axpy(-1.0, y, r));
        RESIDUAL(r);
        RECORD();
        CONVERGED_IF_TOL();
        CHK(lisd_axpy(-alpha, map, mr));         /* mr = mr - alpha*map - my */
        CHK(lisd_axpy(-1.0, my, mr));
        CHK(lisd_matvec(A, mr, amr));
        CHK(lis_vector_dot(rtld, amr, &rho));
        BREAKDOWN_IF(rho == 0.0);
        beta = (rho / rho_old) * (alpha / qsi);
        CHK(lisd_axpy(-1.0, u, p));
        CHK(lisd_xpay(mr, beta, p));
        CHK(lisd_axpy(-1.0, au, ap));
        CHK(lisd_xpay(amr, beta, ap));
        rho_old = rho;
    }
    STOP(LIS_MAXITER);
}

/* ================================================================== Orthomin(m)   src/solver/lis_solver_orthomin.c:124
 * left-preconditioned: rtld = M^-1 r is carried along; m = -restart directions are kept */
LIS_INT lis_orthomin(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    const LIS_INT m = solver->options[LIS_OPTIONS_RESTART];
    LIS_VECTOR r = W(0), rtld = W(1), *p = &W(2), *ap = &W((m + 1) + 2), *aptld = &W(2 * (m + 1) + 2);
    LIS_SCALAR alpha, beta;
    LIS_SCALAR *dotsave = (LIS_SCALAR *)lis_calloc(sizeof(LIS_SCALAR) * (size_t)(m + 1), "lis_orthomin::dotsave");
    if (dotsave == NULL) { LIS_SETERR_MEM(sizeof(LIS_SCALAR) * (m + 1)); return LIS_ERR_OUT_OF_MEMORY; }
    LIS_INT err;
#define OCHK(e) do { err = (e); if (err) { lis_free(dotsave); return err; } } while (0)
    err = lis_solver_get_initial_residual(solver, solver->precon, r, rtld, &bnrm2);
    if (err) { lis_free(dotsave); return err == LIS_FAILS ? LIS_SUCCESS : err; }
    tol = solver->tol;
    iter = 1;
    while (iter <= maxiter) {
        const LIS_INT ip = (iter - 1) % (m + 1);
        OCHK(lisd_copy(rtld, p[ip]));
        OCHK(lisd_matvec(A, p[ip], ap[ip]));
        { const double t0 = lis_wtime(); OCHK(lis_psolve(solver, ap[ip], aptld[ip])); ptime += lis_wtime() - t0; }
        const LIS_INT lmax = _min(m, iter - 1);
        for (LIS_INT l = 1; l <= lmax; l++) {
            const LIS_INT ip0 = (ip + m + 1 - l) % (m + 1);
            OCHK(lis_vector_dot(aptld[ip], aptld[ip0], &beta));
            beta = -beta * dotsave[l - 1];
            OCHK(lisd_axpy(beta, p[ip0], p[ip]));
            OCHK(lisd_axpy(beta, ap[ip0], ap[ip]));
            OCHK(lisd_axpy(beta, aptld[ip0], aptld[ip]));
        }
        for (LIS_INT l = m - 1; l > 0; l--) dotsave[l] = dotsave[l - 1];
        OCHK(lis_vector_dot(aptld[ip], aptld[ip], &dotsave[0]));
        if (dotsave[0] == 0.0) { lis_free(dotsave); STOP(LIS_BREAKDOWN); }
        dotsave[0] = 1.0 / dotsave[0];
        OCHK(lis_vector_dot(rtld, aptld[ip], &alpha));
        alpha = alpha * dotsave[0];
        OCHK(lisd_axpy(alpha, p[ip], x));
        OCHK(lisd_axpy(-alpha, ap[ip], r));
        OCHK(lisd_axpy(-alpha, aptld[ip], rtld));
        OCHK(lis_host_solver_residual(solver, r, &nrm2));
        RECORD();
        if (tol >= nrm2) {
            lis_free(dotsave);
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        iter++;
    }
#undef OCHK
    lis_free(dotsave);
    STOP(LIS_MAXITER);
}

/* ================================================================== MINRES   src/solver/lis_solver_minres.c:121
 * its own residual bookkeeping: ||r_k|| / ||r_0|| from the Lanczos recurrences, tested against
 * the raw -tol; the initial guess is used as given (no lis_solver_get_initial_residual) */
LIS_INT lis_minres(LIS_SOLVER solver)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR b = solver->b, x = solver->x;
    LIS_VECTOR v1 = W(0), v2 = W(1), v3 = W(2), v4 = W(3), w0 = W(4), w1 = W(5), w2 = W(6);
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER], output = solver->options[LIS_OPTIONS_OUTPUT];
    const LIS_REAL tol = solver->params[LIS_PARAMS_RESID - LIS_OPTIONS_LEN];
    LIS_REAL nrm2, beta2, beta3, r0_euc, r_euc;
    LIS_SCALAR alpha, gamma1, gamma2, gamma3, delta, eta, sigma1, sigma2, sigma3, rho1, rho2, rho3;
    LIS_INT iter;
    double ptime = 0.0;

    CHK(lisd_matvec(A, x, v2));
    CHK(lisd_xpay(b, -1.0, v2));
    PSOLVE(v2, v3);
    CHK(lisd_copy(v3, v2));
    CHK(lis_vector_nrm2(v2, &r_euc));
    eta = beta2 = r0_euc = r_euc;
    gamma2 = gamma1 = 1.0;
    sigma2 = sigma1 = 0.0;
    CHK(lisd_set_all(0.0, v1));
    CHK(lisd_set_all(0.0, w0));
    CHK(lisd_set_all(0.0, w1));
    nrm2 = r_euc / r0_euc;
    for (iter = 1; iter <= maxiter; iter++) {
        /* Lanczos step */
        CHK(lisd_scale(1.0 / beta2, v2));
        CHK(lisd_matvec(A, v2, v3));
        PSOLVE(v3, v4);
        CHK(lis_vector_dot(v2, v4, &alpha));
        CHK(lisd_axpy(-alpha, v2, v4));
        CHK(lisd_axpy(-beta2, v1, v4));
        CHK(lis_vector_nrm2(v4, &beta3));
        /* Givens rotations on the tridiagonal */
        delta = gamma2 * alpha - gamma1 * sigma2 * beta2;
        rho1 = sqrt(delta * delta + beta3 * beta3);
        rho2 = sigma2 * alpha + gamma1 * gamma2 * beta2;
        rho3 = sigma1 * beta2;
        gamma3 = delta / rho1;
        sigma3 = beta3 / rho1;
        CHK(lisd_axpyz(-rho3, w0, v2, w2));
        CHK(lisd_axpy(-rho2, w1, w2));
        CHK(lisd_scale(1.0 / rho1, w2));
        CHK(lisd_axpy(gamma3 * eta, w2, x));
        r_euc *= fabs(sigma3);
        nrm2 = r_euc / r0_euc;
        RECORD();
        if (nrm2 <= tol) {
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            return LIS_SUCCESS;
        }
        eta *= -sigma3;
        CHK(lisd_copy(v2, v1));
        CHK(lisd_copy(v4, v2));
        CHK(lisd_copy(w1, w0));
        CHK(lisd_copy(w2, w1));
        beta2 = beta3;
        gamma1 = gamma2; gamma2 = gamma3;
        sigma1 = sigma2; sigma2 = sigma3;
    }
    STOP(LIS_MAXITER);
}

/* ================================================================== FGMRES(m)   src/solver/lis_solver_gmres.c:1128
 * flexible variant: the preconditioned vectors z_j are kept, the restart residual is recomputed
 * from b - A x; the residual estimate is |s_{i+1}| of the system normalised by ||r_0|| */
LIS_INT lis_fgmres(LIS_SOLVER solver)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR b = solver->b, x = solver->x;
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER], output = solver->options[LIS_OPTIONS_OUTPUT];
    const LIS_INT m = solver->options[LIS_OPTIONS_RESTART];
    const LIS_INT h_dim = m + 1, cs = (m + 1) * h_dim, sn = (m + 2) * h_dim;
    LIS_VECTOR *z = &W(2), *v = &W(m + 2);
    LIS_SCALAR aa, bb, rr, a2, b2, t;
    LIS_REAL bnrm2, nrm2 = 0.0, tol, rnorm;
    LIS_INT iter, i, j, k, ii = 0, i1 = 0, iih, jj, err;
    double ptime = 0.0;
    LIS_SCALAR *h = (LIS_SCALAR *)lis_malloc(sizeof(LIS_SCALAR) * (size_t)(h_dim + 1) * (size_t)(h_dim + 2), "lis_fgmres::h");
    LIS_SCALAR *s = (LIS_SCALAR *)lis_calloc(sizeof(LIS_SCALAR) * (size_t)(m + 2), "lis_fgmres::s");
    if (!h || !s) { lis_free2(2, h, s); LIS_SETERR_MEM(sizeof(LIS_SCALAR) * (h_dim + 1) * (h_dim + 2)); return LIS_ERR_OUT_OF_MEMORY; }
#define FCHK(e) do { err = (e); if (err) { lis_free2(2, h, s); return err; } } while (0)
    err = lis_solver_get_initial_residual(solver, NULL, NULL, v[0], &bnrm2);
    if (err) { lis_free2(2, h, s); return err == LIS_FAILS ? LIS_SUCCESS : err; }
    tol = solver->tol;
    rnorm = 1.0 / bnrm2;
    iter = 0;
    while (iter < maxiter) {
        FCHK(lisd_scale(bnrm2, v[0]));
        for (k = 0; k < m + 1; k++) s[k] = 0.0;
        s[0] = rnorm;
        i = 0;
        do {
            iter++; i++;
            ii = i - 1; i1 = i; iih = (i - 1) * h_dim;
            { const double t0 = lis_wtime(); FCHK(lis_psolve(solver, v[ii], z[ii])); ptime += lis_wtime() - t0; }
            FCHK(lisd_matvec(A, z[ii], v[i1]));
            FCHK(lis_host_mgs(v, i, h + iih, &t));
            h[i1 + iih] = t;
            FCHK(lisd_scale(1.0 / t, v[i1]));
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
            nrm2 = fabs(s[i1]);
            RECORD();
            if (tol >= nrm2) break;
        } while (i < m && iter < maxiter);
        s[ii] = s[ii] / h[ii + iih];
        for (k = 1; k <= ii; k++) {
            jj = ii - k;
            t = s[jj];
            for (j = jj + 1; j <= ii; j++) t -= h[jj + j * h_dim] * s[j];
            s[jj] = t / h[jj + jj * h_dim];
        }
        for (j = 0; j <= ii; j++) FCHK(lisd_axpy(s[j], z[j], x));
        if (tol >= nrm2) {
            solver->retcode = LIS_SUCCESS; solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
            lis_free2(2, h, s);
            return LIS_SUCCESS;
        }
        FCHK(lisd_matvec(A, x, v[0]));
        FCHK(lisd_xpay(b, -1.0, v[0]));
        FCHK(lis_vector_nrm2(v[0], &rnorm));
        bnrm2 = 1.0 / rnorm;
    }
#undef FCHK
    lis_free2(2, h, s);
    solver->retcode = LIS_MAXITER; solver->iter = iter + 1; solver->resid = nrm2;
    return LIS_MAXITER;
}

/* ================================================================== BiCGSTAB(l)   src/solver/lis_solver_bicgstabl.c:123
 * l = -ell (default 2).  Right preconditioning through the accumulated update: x collects the
 * un-preconditioned corrections and is finished as x = M^-1 x + x_0 on every exit. */
LIS_INT lis_bicgstabl(LIS_SOLVER solver)
{
    COMMON_LOCALS;
    const LIS_INT l = solver->options[LIS_OPTIONS_ELL], z_dim = l + 1;
    LIS_VECTOR rtld = W(0), xp = W(1), bp = W(2), t = W(3), *r = &W(4), *u = &W(l + 1 + 4);
    LIS_SCALAR alpha = 0.0, beta, omega = 1.0, rho0 = 1.0, rho1, nu;
    LIS_REAL rnorm0, rnorm, normx, normr;
    LIS_INT i, j, err;
    LIS_SCALAR *tau = (LIS_SCALAR *)lis_calloc(sizeof(LIS_SCALAR) * (size_t)z_dim * (size_t)(4 + l + 1), "lis_bicgstabl::tau");
    if (tau == NULL) { LIS_SETERR_MEM(sizeof(LIS_SCALAR) * z_dim * (4 + l + 1)); return LIS_ERR_OUT_OF_MEMORY; }
    LIS_SCALAR *gamma = &tau[z_dim * z_dim], *gamma1 = &gamma[z_dim], *gamma2 = &gamma1[z_dim], *sigma = &gamma2[z_dim];
#define LCHK(e) do { err = (e); if (err) { lis_free(tau); return err; } } while (0)
#define FINISH_X() do { const double t0_ = lis_wtime(); LCHK(lis_psolve(solver, x, t)); LCHK(lisd_copy(t, x)); \
                        ptime += lis_wtime() - t0_; LCHK(lisd_axpy(1.0, xp, x)); } while (0)
#define LEAVE(code) do { FINISH_X(); solver->retcode = (code); solver->iter = iter; solver->resid = nrm2; \
                         solver->ptime = ptime; lis_free(tau); return (code); } while (0)
    err = lis_solver_get_initial_residual(solver, NULL, NULL, r[0], &bnrm2);
    if (err) { lis_free(tau); return err == LIS_FAILS ? LIS_SUCCESS : err; }
    tol = solver->tol;
    LCHK(lis_host_solver_shadow_residual(solver, r[0], rtld));
    LCHK(lisd_copy(r[0], bp));
    LCHK(lisd_copy(x, xp));
    LCHK(lisd_set_all(0.0, u[0]));
    LCHK(lis_vector_nrm2(r[0], &rnorm0));
    rnorm = normx = normr = rnorm0;
    iter = 0;
    while (iter <= maxiter) {
        /* BiCG part */
        rho0 = -omega * rho0;
        for (j = 0; j < l; j++) {
            iter++;
            LCHK(lis_vector_dot(rtld, r[j], &rho1));
            if (rho1 == 0.0) LEAVE(LIS_BREAKDOWN);
            beta = alpha * (rho1 / rho0);
            rho0 = rho1;
            for (i = 0; i <= j; i++) LCHK(lisd_xpay(r[i], -beta, u[i]));       /* u_i = r_i - beta*u_i */
            { const double t0 = lis_wtime(); LCHK(lis_psolve(solver, u[j], t)); ptime += lis_wtime() - t0; }
            LCHK(lisd_matvec(A, t, u[j + 1]));
            LCHK(lis_vector_dot(rtld, u[j + 1], &nu));
            if (nu == 0.0) LEAVE(LIS_BREAKDOWN);
            alpha = rho1 / nu;
            LCHK(lisd_axpy(alpha, u[0], x));
            for (i = 0; i <= j; i++) LCHK(lisd_axpy(-alpha, u[i + 1], r[i]));
            LCHK(lis_host_solver_residual(solver, r[0], &nrm2));
            if (iter % l != 0) RECORD();
            if (tol >= nrm2) { RECORD(); LEAVE(LIS_SUCCESS); }
            { const double t0 = lis_wtime(); LCHK(lis_psolve(solver, r[j], t)); ptime += lis_wtime() - t0; }
            LCHK(lisd_matvec(A, t, r[j + 1]));
            LCHK(lis_vector_nrm2(r[0], &rnorm));
            normx = _max(normx, rnorm);
            normr = _max(normr, rnorm);
        }
        /* minimal-residual part: modified Gram-Schmidt on r_1..r_l, then the polynomial update */
        for (j = 1; j <= l; j++) {
            for (i = 1; i <= j - 1; i++) {
                LCHK(lis_vector_dot(r[j], r[i], &nu));
                nu = nu / sigma[i];
                tau[i * z_dim + j] = nu;
                LCHK(lisd_axpy(-nu, r[i], r[j]));
            }
            LCHK(lis_vector_dot(r[j], r[j], &sigma[j]));
            LCHK(lis_vector_dot(r[0], r[j], &nu));
            gamma1[j] = nu / sigma[j];
        }
        gamma[l] = gamma1[l];
        omega = gamma[l];
        for (j = l - 1; j >= 1; j--) {
            nu = 0.0;
            for (i = j + 1; i <= l; i++) nu += tau[j * z_dim + i] * gamma[i];
            gamma[j] = gamma1[j] - nu;
        }
        for (j = 1; j <= l - 1; j++) {
            nu = 0.0;
            for (i = j + 1; i <= l - 1; i++) nu += tau[j * z_dim + i] * gamma[i + 1];
            gamma2[j] = gamma[j + 1] + nu;
        }
        LCHK(lisd_axpy(gamma[1], r[0], x));
        LCHK(lisd_axpy(-gamma1[l], r[l], r[0]));
        LCHK(lisd_axpy(-gamma[l], u[l], u[0]));
        for (j = 1; j <= l - 1; j++) {
            LCHK(lisd_axpy(-gamma[j], u[j], u[0]));
            LCHK(lisd_axpy(gamma2[j], r[j], x));
            LCHK(lisd_axpy(-gamma1[j], r[j], r[0]));
        }
        LCHK(lis_host_solver_residual(solver, r[0], &nrm2));
        RECORD();
        if (tol >= nrm2) LEAVE(LIS_SUCCESS);
    }
    (void)normx; (void)normr; (void)rnorm;
    lis_free(tau);
#undef LCHK
#undef FINISH_X
#undef LEAVE
    STOP(LIS_MAXITER);
}

/* ================================================================== stationary methods
 * src/solver/lis_solver_jacobi.c:113, lis_solver_gs.c:113, lis_solver_sor.c:123.  They iterate on
 * x directly, use the raw -tol against ||b - A M^-1 x|| / ||b||, and finish with x = M^-1 x.
 * (With a preconditioner the reference first rescales the system, lis_solver.c:676-690; that
 * path is not carried over, so lis_solve accepts these three with -p none only.) */
static LIS_INT stationary_run(LIS_SOLVER solver, int kind, LIS_MATRIX S)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR b = solver->b, x = solver->x, r = W(0), t = W(1), s = W(2);
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER], output = solver->options[LIS_OPTIONS_OUTPUT];
    const LIS_REAL tol = solver->params[LIS_PARAMS_RESID - LIS_OPTIONS_LEN];
    LIS_REAL bnrm2, nrm2 = 0.0;
    LIS_INT iter;
    double ptime = 0.0;
    CHK(lis_vector_nrm2(b, &bnrm2));
    bnrm2 = 1.0 / bnrm2;
    if (kind == 0) {                                   /* Jacobi: d = 1/diag(A) */
        CHK(lis_matrix_get_diagonal(A, W(3)));
        CHK(lis_vector_reciprocal(W(3)));
    } else {
        CHK(lis_matrix_split(S));
        if (kind == 1) CHK(lis_host_set_wd(S, 1.0, 0, LIS_SOLVER_GS));                                            /* WD = 1/D     */
        else CHK(lis_host_set_wd(S, 1.0 / solver->params[LIS_PARAMS_OMEGA - LIS_OPTIONS_LEN], 1, LIS_SOLVER_SOR));  /* WD = 1/(D/w) */
    }
    for (iter = 1; iter <= maxiter; iter++) {
        PSOLVE(x, s);
        CHK(lisd_matvec(A, s, t));
        CHK(lisd_axpyz(-1.0, t, b, r));              /* r = b - A M^-1 x */
        CHK(lis_vector_nrm2(r, &nrm2));
        if (kind == 0) {
            CHK(lisd_pmul(r, W(3), r));
            CHK(lisd_axpy(1.0, r, x));
        } else {
            CHK(lis_matrix_solve(S, r, t, LIS_MATRIX_LOWER));
            CHK(lisd_axpy(1.0, t, x));
        }
        nrm2 = nrm2 * bnrm2;
        RECORD();
        if (tol >= nrm2) break;
    }
    PSOLVE(x, s);
    CHK(lisd_copy(s, x));
    solver->iter = iter; solver->resid = nrm2; solver->ptime = ptime;
    solver->retcode = iter <= maxiter ? LIS_SUCCESS : LIS_MAXITER;
    return solver->retcode;
}

/* Gauss-Seidel / SOR sweep on the (D + L) part: in CSR storage the solver's own matrix is split in place like
 * the reference does; in the other scalar formats the sweep runs on a private CSR copy (same D and L) while the
 * products stay in the chosen format.  The block formats sweep block-wise in the reference: not offered. */
static LIS_INT stationary(LIS_SOLVER solver, int kind)
{
    LIS_MATRIX A = solver->A, S = A;
    LIS_INT err;
    if (kind != 0 && A->matrix_type != LIS_MATRIX_CSR && solver->precon && solver->precon->is_copy && solver->precon->A &&
        solver->precon->A->matrix_type == LIS_MATRIX_CSR && solver->precon->A->is_splited) {
        /* -p ssor already sweeps on a private split copy: the reference has ONE split matrix for both (so the two share
         * WD, whoever set it first under the same tag: lis_solver_sor.c, lis_precon_ssor.c) -- share it here too */
        return stationary_run(solver, kind, solver->precon->A);
    }
    if (kind != 0 && A->matrix_type != LIS_MATRIX_CSR) {
        if (A->matrix_type == LIS_MATRIX_BSR || A->matrix_type == LIS_MATRIX_BSC || A->matrix_type == LIS_MATRIX_VBR) {
            LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "Gauss-Seidel / SOR on block storage (BSR/BSC/VBR) is not available; use a scalar format\n");
            return LIS_ERR_NOT_IMPLEMENTED;
        }
        err = lis_matrix_duplicate(A, &S);
        if (err) return err;
        err = lis_matrix_set_type(S, LIS_MATRIX_CSR);
        if (!err) err = lis_matrix_convert(A, S);
        if (err) { lis_matrix_destroy(S); return err; }
    }
    err = stationary_run(solver, kind, S);
    if (S != A) lis_matrix_destroy(S);
    return err;
}

LIS_INT lis_jacobi(LIS_SOLVER solver) { return stationary(solver, 0); }
LIS_INT lis_gs(LIS_SOLVER solver) { return stationary(solver, 1); }
LIS_INT lis_sor(LIS_SOLVER solver) { return stationary(solver, 2); }
