/*
 * lis_sptrsv.c -- a strictly triangular CSR factor prepared for the one-launch solve kernel
 * (lisb200_sweep_sell): rows grouped by dependency level on the host, once per factor, the
 * factor uploaded permuted into that order.  Users: the ILU(k) apply (lis_precon_ilu.c) and the
 * transposed SSOR sweep (lis_precon.c).  The row sums run in the storage order of the CSR given
 * here, which is how the callers pin the reference's summation order.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lis_device.h"
#include "lis_host.h"
#include "lis_b200_kernels.h"

void lisd_tri_free(lisd_tri *T)
{
    if (T == NULL) return;
    lisd_perm_free(&T->p);
    lisd_free(T->d_ticket);
    free(T);
}

LIS_INT lisd_tri_build(int n, const LIS_INT *ptr, const LIS_INT *idx, const LIS_SCALAR *val, lisd_tri **out)
{
    lisd_tri *T = (lisd_tri *)calloc(1, sizeof(lisd_tri));
    int *lvl = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *rows = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *lptr = NULL;
    LIS_INT err = LIS_OUT_OF_MEMORY;
    int lower = 0, upper = 0, nlev = 0;
    *out = NULL;
    if (!T || !lvl || !rows) { LIS_SETERR_MEM(n * sizeof(int)); goto fail; }
    for (int i = 0; i < n; i++)
        for (LIS_INT j = ptr[i]; j < ptr[i + 1]; j++) {
            if (idx[j] < i) lower = 1;
            else if (idx[j] > i && idx[j] < n) upper = 1;
            else { LIS_SETERR2(LIS_ERR_ILL_ARG, "triangular factor: entry (%D,%D) is not strictly triangular\n", i, idx[j]); err = LIS_ERR_ILL_ARG; goto fail; }
        }
    if (lower && upper) { LIS_SETERR(LIS_ERR_ILL_ARG, "triangular factor: entries on both sides of the diagonal\n"); err = LIS_ERR_ILL_ARG; goto fail; }
    /* level of a row = 1 + the deepest row it reads */
    for (int s = 0; s < n; s++) {
        const int i = upper ? n - 1 - s : s;
        int l = 0;
        for (LIS_INT j = ptr[i]; j < ptr[i + 1]; j++)
            if (lvl[idx[j]] + 1 > l) l = lvl[idx[j]] + 1;
        lvl[i] = l;
        if (l + 1 > nlev) nlev = l + 1;
    }
    lptr = lisd_order_by_level(n, lvl, nlev, rows);
    if (!lptr) { LIS_SETERR_MEM(nlev * sizeof(int)); goto fail; }
    T->n = n; T->nlev = nlev;
    err = lisd_perm_build(&T->p, n, nlev, lptr, rows, ptr, idx, val, NULL, NULL);
    if (!err) err = lisd_malloc((void **)&T->d_ticket, 64);
    if (err) goto fail;
    free(lvl); free(rows); free(lptr);
    *out = T;
    return LIS_SUCCESS;
fail:
    free(lvl); free(rows); free(lptr);
    lisd_tri_free(T);
    return err;
}

LIS_INT lisd_tri_solve(const lisd_tri *T, int mode, const double *d_wd, const double *d_in, double *d_out, const char *what)
{
    lisd_mark_busy();
    return lisd_check(lisd_perm_sweep(&T->p, mode, T->n, d_wd, d_in, d_out, T->d_ticket), what);
}
