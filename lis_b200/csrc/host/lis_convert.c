/*
 * lis_convert.c -- storage-format conversion (host, done once before the hot path):
 * lis_matrix_copy, lis_matrix_convert, lis_matrix_convert_self.
 *
 * Every conversion routes through CSR like the reference's hub
 * (src/matrix/lis_matrix_ops.c:127-322).  The layouts produced are the ones the reference's
 * SERIAL build produces (nprocs == 1 in its builders), which are also the coalesced layouts
 * the GPU kernels want:
 *   ELL  value[j*n+i], pad (0.0, column i)          src/matrix/lis_matrix_ell.c:957-1070
 *   DIA  value[j*n+i], offsets ascending, sorts Ain  src/matrix/lis_matrix_dia.c:1190-1305
 *   JAD  rows by descending length                   src/matrix/lis_matrix_jad.c:1590-1770
 *   BSR  blocks column-major, first-seen order       src/matrix/lis_matrix_bsr.c:350-545
 *   CSC  counting transpose                          src/matrix/lis_matrix_csc.c:1000-1080
 * Back-conversions drop explicit zeros where the reference does (ELL, DIA, BSR).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lis_device.h"
#include "lis_host.h"

/* ------------------------------------------------------------------ CSR -> X */
static LIS_INT finish(LIS_MATRIX Aout, LIS_INT err)
{
    if (err) return err;
    err = lis_matrix_assemble(Aout);
    if (err) lis_matrix_storage_destroy(Aout);
    return err;
}

static LIS_INT csr2csr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, nnz = Ain->ptr[n];
    LIS_INT *ptr, *index;
    LIS_SCALAR *value;
    LIS_INT err = lis_matrix_malloc_csr(n, nnz, &ptr, &index, &value);
    if (err) return err;
    memcpy(ptr, Ain->ptr, ((size_t)n + 1) * sizeof(LIS_INT));
    memcpy(index, Ain->index, (size_t)nnz * sizeof(LIS_INT));
    memcpy(value, Ain->value, (size_t)nnz * sizeof(LIS_SCALAR));
    err = lis_matrix_set_csr(nnz, ptr, index, value, Aout);
    if (err) { lis_free2(3, ptr, index, value); return err; }
    Aout->is_sorted = Ain->is_sorted;
    return finish(Aout, LIS_SUCCESS);
}

static LIS_INT csr2ell(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n;
    LIS_INT maxnzr = 0, err;
    LIS_INT *index;
    LIS_SCALAR *value;
    for (LIS_INT i = 0; i < n; i++)
        if (Ain->ptr[i + 1] - Ain->ptr[i] > maxnzr) maxnzr = Ain->ptr[i + 1] - Ain->ptr[i];
    err = lis_matrix_malloc_ell(n, maxnzr, &index, &value);
    if (err) return err;
    for (LIS_INT j = 0; j < maxnzr; j++) {
        LIS_INT *ij = index + (size_t)j * n;
        LIS_SCALAR *vj = value + (size_t)j * n;
        for (LIS_INT i = 0; i < n; i++) { vj[i] = 0.0; ij[i] = i; }
    }
    for (LIS_INT i = 0; i < n; i++) {
        size_t k = (size_t)i;
        for (LIS_INT j = Ain->ptr[i]; j < Ain->ptr[i + 1]; j++, k += (size_t)n) {
            value[k] = Ain->value[j];
            index[k] = Ain->index[j];
        }
    }
    err = lis_matrix_set_ell(maxnzr, index, value, Aout);
    if (err) { lis_free2(2, index, value); return err; }
    Aout->nnz = Ain->ptr[n];
    return finish(Aout, LIS_SUCCESS);
}

static LIS_INT csr2dia(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, np = Ain->np;
    LIS_INT nnd = 0, err;
    LIS_INT *index;
    LIS_SCALAR *value;
    lis_matrix_sort_csr(Ain);                                  /* :1217, mutates Ain like the reference */
    /* distinct offsets column-row in (-n, np): a presence map instead of sorting nnz integers */
    const size_t span = (size_t)n + (size_t)np;
    unsigned char *seen = (unsigned char *)calloc(span > 0 ? span : 1, 1);
    if (seen == NULL) { LIS_SETERR_MEM(span); return LIS_OUT_OF_MEMORY; }
    for (LIS_INT i = 0; i < n; i++)
        for (LIS_INT j = Ain->ptr[i]; j < Ain->ptr[i + 1]; j++) seen[(size_t)(Ain->index[j] - i + n)] = 1;
    for (size_t k = 0; k < span; k++) nnd += seen[k];
    err = lis_matrix_malloc_dia(n, nnd, &index, &value);
    if (err) { free(seen); return err; }
    nnd = 0;
    for (size_t k = 0; k < span; k++)
        if (seen[k]) index[nnd++] = (LIS_INT)((long long)k - n);
    free(seen);
    memset(value, 0, (size_t)n * (size_t)nnd * sizeof(LIS_SCALAR));
    for (LIS_INT i = 0; i < n; i++) {
        LIS_INT k = 0;
        for (LIS_INT j = Ain->ptr[i]; j < Ain->ptr[i + 1]; j++) {
            const LIS_INT off = Ain->index[j] - i;
            while (index[k] != off) k++;
            value[(size_t)k * n + i] = Ain->value[j];
        }
    }
    err = lis_matrix_set_dia(nnd, index, value, Aout);
    if (err) { lis_free2(2, index, value); return err; }
    Aout->nnz = Ain->ptr[n];
    return finish(Aout, LIS_SUCCESS);
}

static LIS_INT csr2jad(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, nnz = Ain->ptr[n];
    LIS_INT maxnzr = 0, err;
    LIS_INT *perm, *ptr, *index;
    LIS_SCALAR *value;
    for (LIS_INT i = 0; i < n; i++)
        if (Ain->ptr[i + 1] - Ain->ptr[i] > maxnzr) maxnzr = Ain->ptr[i + 1] - Ain->ptr[i];
    err = lis_matrix_malloc_jad(n, nnz, maxnzr, &perm, &ptr, &index, &value);
    if (err) return err;
    /* rows by descending length (counting sort, stable in the row number) */
    LIS_INT *cnt = (LIS_INT *)calloc((size_t)maxnzr + 2, sizeof(LIS_INT));
    if (cnt == NULL) { lis_free2(4, perm, ptr, index, value); LIS_SETERR_MEM(maxnzr); return LIS_OUT_OF_MEMORY; }
    memset(ptr, 0, ((size_t)maxnzr + 1) * sizeof(LIS_INT));
    for (LIS_INT i = 0; i < n; i++) {
        const LIS_INT len = Ain->ptr[i + 1] - Ain->ptr[i];
        cnt[maxnzr - len + 1]++;
        for (LIS_INT j = 0; j < len; j++) ptr[j + 1]++;
    }
    for (LIS_INT l = 0; l <= maxnzr; l++) cnt[l + 1] += cnt[l];
    for (LIS_INT i = 0; i < n; i++) perm[cnt[maxnzr - (Ain->ptr[i + 1] - Ain->ptr[i])]++] = i;
    free(cnt);
    for (LIS_INT j = 0; j < maxnzr; j++) ptr[j + 1] += ptr[j];
    for (LIS_INT i = 0; i < n; i++) {
        const LIS_INT js = Ain->ptr[perm[i]], je = Ain->ptr[perm[i] + 1];
        for (LIS_INT j = js; j < je; j++) {
            const LIS_INT l = ptr[j - js] + i;
            value[l] = Ain->value[j];
            index[l] = Ain->index[j];
        }
    }
    err = lis_matrix_set_jad(nnz, maxnzr, perm, ptr, index, value, Aout);
    if (err) { lis_free2(4, perm, ptr, index, value); return err; }
    return finish(Aout, LIS_SUCCESS);
}

static LIS_INT csr2bsr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, np = Ain->np;
    const LIS_INT bnr = Aout->conv_bnr, bnc = Aout->conv_bnc, bs = bnr * bnc;
    /* row-partitioned: n local rows, np = n + halo columns (src/matrix/lis_matrix_bsr.c:371-375 under USE_MPI) */
    const LIS_INT nr = 1 + (n - 1) / bnr, nc = 1 + (np - 1) / bnc;
    LIS_INT *bptr = NULL, *bindex = NULL, err;
    LIS_SCALAR *value = NULL;
    LIS_INT *pos = (LIS_INT *)calloc((size_t)nc, sizeof(LIS_INT));       /* 1 + block slot, 0 = unseen */
    LIS_INT *cntptr = (LIS_INT *)malloc(((size_t)nr + 1) * sizeof(LIS_INT));
    LIS_INT *list = (LIS_INT *)malloc((size_t)nc * sizeof(LIS_INT));
    if (!pos || !cntptr || !list) { free(pos); free(cntptr); free(list); LIS_SETERR_MEM(nc); return LIS_OUT_OF_MEMORY; }
    cntptr[0] = 0;
    for (LIS_INT bi = 0; bi < nr; bi++) {
        LIS_INT cnt = 0;
        for (LIS_INT ii = 0; ii < bnr && bi * bnr + ii < n; ii++)
            for (LIS_INT j = Ain->ptr[bi * bnr + ii]; j < Ain->ptr[bi * bnr + ii + 1]; j++) {
                const LIS_INT bj = Ain->index[j] / bnc;
                if (!pos[bj]) { pos[bj] = 1; list[cnt++] = bj; }
            }
        for (LIS_INT k = 0; k < cnt; k++) pos[list[k]] = 0;
        cntptr[bi + 1] = cntptr[bi] + cnt;
    }
    free(list);
    const LIS_INT bnnz = cntptr[nr];
    err = lis_matrix_malloc_bsr(n, bnr, bnc, bnnz, &bptr, &bindex, &value);
    if (err) { free(pos); free(cntptr); return err; }
    memcpy(bptr, cntptr, ((size_t)nr + 1) * sizeof(LIS_INT));
    free(cntptr);
    for (LIS_INT bi = 0; bi < nr; bi++) {
        LIS_INT kk = bptr[bi];
        for (LIS_INT ii = 0; ii < bnr && bi * bnr + ii < n; ii++)
            for (LIS_INT k = Ain->ptr[bi * bnr + ii]; k < Ain->ptr[bi * bnr + ii + 1]; k++) {
                const LIS_INT bj = Ain->index[k] / bnc, j = Ain->index[k] % bnc;
                if (pos[bj] == 0) {
                    const size_t kv = (size_t)kk * bs;
                    pos[bj] = kk + 1;
                    bindex[kk] = bj;
                    for (LIS_INT q = 0; q < bs; q++) value[kv + q] = 0.0;
                    value[kv + (size_t)j * bnr + ii] = Ain->value[k];
                    kk++;
                } else {
                    value[(size_t)(pos[bj] - 1) * bs + (size_t)j * bnr + ii] = Ain->value[k];
                }
            }
        for (LIS_INT j = bptr[bi]; j < bptr[bi + 1]; j++) pos[bindex[j]] = 0;
    }
    free(pos);
    err = lis_matrix_set_bsr(bnr, bnc, bnnz, bptr, bindex, value, Aout);
    if (err) { lis_free2(3, bptr, bindex, value); return err; }
    Aout->nnz = Ain->ptr[n];
    return finish(Aout, LIS_SUCCESS);
}

/* counting transpose shared by csr2csc and csc2csr (n rows, ncols columns of the input) */
static LIS_INT transpose(LIS_INT n, LIS_INT ncols, const LIS_INT *ptr, const LIS_INT *index, const LIS_SCALAR *value,
                         LIS_INT **optr, LIS_INT **oindex, LIS_SCALAR **ovalue)
{
    const LIS_INT nnz = ptr[n];
    LIS_INT err = lis_matrix_malloc_csr(ncols, nnz, optr, oindex, ovalue);
    if (err) return err;
    LIS_INT *iw = (LIS_INT *)calloc((size_t)ncols + 1, sizeof(LIS_INT));
    if (iw == NULL) { lis_free2(3, *optr, *oindex, *ovalue); LIS_SETERR_MEM(ncols); return LIS_OUT_OF_MEMORY; }
    for (LIS_INT i = 0; i < n; i++)
        for (LIS_INT j = ptr[i]; j < ptr[i + 1]; j++) iw[index[j]]++;
    (*optr)[0] = 0;
    for (LIS_INT i = 0; i < ncols; i++) { (*optr)[i + 1] = (*optr)[i] + iw[i]; iw[i] = (*optr)[i]; }
    for (LIS_INT i = 0; i < n; i++)
        for (LIS_INT j = ptr[i]; j < ptr[i + 1]; j++) {
            const LIS_INT l = iw[index[j]]++;
            (*ovalue)[l] = value[j];
            (*oindex)[l] = i;
        }
    free(iw);
    return LIS_SUCCESS;
}

LIS_INT lis_host_transpose(LIS_INT n, LIS_INT ncols, const LIS_INT *ptr, const LIS_INT *index, const LIS_SCALAR *value,
                           LIS_INT **optr, LIS_INT **oindex, LIS_SCALAR **ovalue)
{
    return transpose(n, ncols, ptr, index, value, optr, oindex, ovalue);
}

static LIS_INT csr2csc(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    /* row-partitioned: np = n + halo columns, each with its own (possibly empty) column list
     * (src/matrix/lis_matrix_csc.c: ptr has np+1 entries under USE_MPI) */
    LIS_INT *ptr, *index;
    LIS_SCALAR *value;
    LIS_INT err = transpose(Ain->n, Ain->np, Ain->ptr, Ain->index, Ain->value, &ptr, &index, &value);
    if (err) return err;
    err = lis_matrix_set_csc(Ain->ptr[Ain->n], ptr, index, value, Aout);
    if (err) { lis_free2(3, ptr, index, value); return err; }
    return finish(Aout, LIS_SUCCESS);
}

/* ------------------------------------------------------------------ X -> CSR */
static LIS_INT install_csr(LIS_MATRIX Aout, LIS_INT nnz, LIS_INT *ptr, LIS_INT *index, LIS_SCALAR *value)
{
    LIS_INT err = lis_matrix_set_csr(nnz, ptr, index, value, Aout);
    if (err) { lis_free2(3, ptr, index, value); return err; }
    return finish(Aout, LIS_SUCCESS);
}

static LIS_INT csc2csr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    LIS_INT *ptr, *index;
    LIS_SCALAR *value;
    LIS_INT err = transpose(Ain->np, Ain->n, Ain->ptr, Ain->index, Ain->value, &ptr, &index, &value);       /* np columns back into n rows */
    if (err) return err;
    return install_csr(Aout, Ain->ptr[Ain->n], ptr, index, value);
}

static LIS_INT ell2csr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, maxnzr = Ain->maxnzr;
    LIS_INT *ptr, *index, err;
    LIS_SCALAR *value;
    LIS_INT *iw = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    if (iw == NULL) { LIS_SETERR_MEM(n); return LIS_OUT_OF_MEMORY; }
    for (LIS_INT j = 0; j < maxnzr; j++)
        for (LIS_INT i = 0; i < n; i++)
            if (Ain->value[(size_t)j * n + i] != 0.0) iw[i]++;
    LIS_INT nnz = 0;
    for (LIS_INT i = 0; i < n; i++) nnz += iw[i];
    err = lis_matrix_malloc_csr(n, nnz, &ptr, &index, &value);
    if (err) { free(iw); return err; }
    ptr[0] = 0;
    for (LIS_INT i = 0; i < n; i++) { ptr[i + 1] = ptr[i] + iw[i]; iw[i] = ptr[i]; }
    for (LIS_INT j = 0; j < maxnzr; j++)
        for (LIS_INT i = 0; i < n; i++) {
            const size_t k = (size_t)j * n + i;
            if (Ain->value[k] != 0.0) { value[iw[i]] = Ain->value[k]; index[iw[i]] = Ain->index[k]; iw[i]++; }
        }
    free(iw);
    return install_csr(Aout, nnz, ptr, index, value);
}

static LIS_INT dia2csr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, np = Ain->np, nnd = Ain->nnd;
    LIS_INT *ptr, *index, err;
    LIS_SCALAR *value;
    LIS_INT *iw = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    if (iw == NULL) { LIS_SETERR_MEM(n); return LIS_OUT_OF_MEMORY; }
    for (LIS_INT j = 0; j < nnd; j++) {
        const LIS_INT off = Ain->index[j];
        const LIS_INT is = off < 0 ? -off : 0, ie = np - off < n ? np - off : n;
        for (LIS_INT i = is; i < ie; i++)
            if (Ain->value[(size_t)j * n + i] != 0.0) iw[i]++;
    }
    LIS_INT nnz = 0;
    for (LIS_INT i = 0; i < n; i++) nnz += iw[i];
    err = lis_matrix_malloc_csr(n, nnz, &ptr, &index, &value);
    if (err) { free(iw); return err; }
    ptr[0] = 0;
    for (LIS_INT i = 0; i < n; i++) { ptr[i + 1] = ptr[i] + iw[i]; iw[i] = ptr[i]; }
    for (LIS_INT j = 0; j < nnd; j++) {
        const LIS_INT off = Ain->index[j];
        const LIS_INT is = off < 0 ? -off : 0, ie = np - off < n ? np - off : n;
        for (LIS_INT i = is; i < ie; i++) {
            const LIS_SCALAR v = Ain->value[(size_t)j * n + i];
            if (v != 0.0) { value[iw[i]] = v; index[iw[i]] = i + off; iw[i]++; }
        }
    }
    free(iw);
    return install_csr(Aout, nnz, ptr, index, value);
}

static LIS_INT jad2csr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, maxnzr = Ain->maxnzr, nnz = Ain->nnz;
    LIS_INT *ptr, *index, err;
    LIS_SCALAR *value;
    err = lis_matrix_malloc_csr(n, nnz, &ptr, &index, &value);
    if (err) return err;
    memset(ptr, 0, ((size_t)n + 1) * sizeof(LIS_INT));
    for (LIS_INT j = 0; j < maxnzr; j++) {
        const LIS_INT len = Ain->ptr[j + 1] - Ain->ptr[j];
        for (LIS_INT i = 0; i < len; i++) ptr[Ain->row[i] + 1]++;
    }
    for (LIS_INT i = 0; i < n; i++) ptr[i + 1] += ptr[i];
    for (LIS_INT j = 0; j < maxnzr; j++) {
        const LIS_INT s = Ain->ptr[j], len = Ain->ptr[j + 1] - s;
        for (LIS_INT i = 0; i < len; i++) {
            const LIS_INT l = ptr[Ain->row[i]] + j;
            value[l] = Ain->value[s + i];
            index[l] = Ain->index[s + i];
        }
    }
    return install_csr(Aout, nnz, ptr, index, value);
}

static LIS_INT bsr2csr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, nr = Ain->nr, bnr = Ain->bnr, bnc = Ain->bnc, bs = bnr * bnc;
    LIS_INT *ptr, *index, err, nnz = 0;
    LIS_SCALAR *value;
    LIS_INT *iw = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    if (iw == NULL) { LIS_SETERR_MEM(n); return LIS_OUT_OF_MEMORY; }
    for (LIS_INT bi = 0; bi < nr; bi++)
        for (LIS_INT bc = Ain->bptr[bi]; bc < Ain->bptr[bi + 1]; bc++)
            for (LIS_INT j = 0; j < bnc; j++)
                for (LIS_INT i = 0; i < bnr; i++)
                    if (bi * bnr + i < n && Ain->value[(size_t)bc * bs + (size_t)j * bnr + i] != 0.0) iw[bi * bnr + i]++;
    for (LIS_INT i = 0; i < n; i++) nnz += iw[i];
    err = lis_matrix_malloc_csr(n, nnz, &ptr, &index, &value);
    if (err) { free(iw); return err; }
    ptr[0] = 0;
    for (LIS_INT i = 0; i < n; i++) { ptr[i + 1] = ptr[i] + iw[i]; iw[i] = ptr[i]; }
    for (LIS_INT bi = 0; bi < nr; bi++)
        for (LIS_INT i = 0; i < bnr && bi * bnr + i < n; i++) {
            const LIS_INT r = bi * bnr + i;
            for (LIS_INT bc = Ain->bptr[bi]; bc < Ain->bptr[bi + 1]; bc++)
                for (LIS_INT j = 0; j < bnc; j++) {
                    const LIS_SCALAR v = Ain->value[(size_t)bc * bs + (size_t)j * bnr + i];
                    if (v != 0.0) { value[iw[r]] = v; index[iw[r]] = Ain->bindex[bc] * bnc + j; iw[r]++; }
                }
        }
    free(iw);
    return install_csr(Aout, nnz, ptr, index, value);
}

/* ------------------------------------------------------------------ public */
static void inherit_partition(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    Aout->np = Ain->np;
    Aout->pad = Ain->pad;
    Aout->is_comm = Ain->is_comm;
    Aout->is_pmat = Ain->is_pmat;
}

static LIS_INT to_csr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    inherit_partition(Ain, Aout);
    switch (Ain->matrix_type) {
    case LIS_MATRIX_CSR: return csr2csr(Ain, Aout);
    case LIS_MATRIX_CSC: return csc2csr(Ain, Aout);
    case LIS_MATRIX_ELL: return ell2csr(Ain, Aout);
    case LIS_MATRIX_DIA: return dia2csr(Ain, Aout);
    case LIS_MATRIX_JAD: return jad2csr(Ain, Aout);
    case LIS_MATRIX_BSR: return bsr2csr(Ain, Aout);
    default: {
        int handled = 0;                                   /* MSR, COO, BSC, VBR, DNS: lis_formats_ext.c */
        LIS_INT err = lis_host_ext_to_csr(Ain, Aout, &handled);
        if (handled) return err;
        LIS_SETERR1(LIS_ERR_NOT_IMPLEMENTED, "conversion from storage format %D is not available\n", Ain->matrix_type);
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    }
}

static LIS_INT from_csr(LIS_MATRIX Acsr, LIS_MATRIX Aout)
{
    inherit_partition(Acsr, Aout);
    {
        /* LIS_B200_CONVERT=device: rearranged in HBM by kernels/convert.cu (lis_convert_dev.c); same arrays */
        int done = 0;
        LIS_INT err = lisd_convert_from_csr(Acsr, Aout, &done);
        if (err || done) return err;
    }
    switch (Aout->matrix_type) {
    case LIS_MATRIX_CSR: return csr2csr(Acsr, Aout);
    case LIS_MATRIX_CSC: return csr2csc(Acsr, Aout);
    case LIS_MATRIX_ELL: return csr2ell(Acsr, Aout);
    case LIS_MATRIX_DIA: return csr2dia(Acsr, Aout);
    case LIS_MATRIX_JAD: return csr2jad(Acsr, Aout);
    case LIS_MATRIX_BSR: return csr2bsr(Acsr, Aout);
    default: {
        int handled = 0;
        LIS_INT err = lis_host_ext_from_csr(Acsr, Aout, &handled);
        if (handled) return err;
        LIS_SETERR1(LIS_ERR_NOT_IMPLEMENTED, "conversion to storage format %D is not available\n", Aout->matrix_type);
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    }
}

LIS_INT lis_matrix_copy(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    LIS_INT err = lis_host_matrix_check_input(Ain);
    if (err) return err;
    if (Ain->matrix_type == LIS_MATRIX_CSR) { inherit_partition(Ain, Aout); return csr2csr(Ain, Aout); }
    /* same-format copy of the other layouts: round trip through CSR keeps the layout rules */
    LIS_MATRIX T;
    const LIS_INT type = Ain->matrix_type;
    err = lis_matrix_duplicate(Ain, &T);
    if (err) return err;
    err = to_csr(Ain, T);
    if (err) { lis_matrix_destroy(T); return err; }
    Aout->matrix_type = type;
    if (type == LIS_MATRIX_BSR || type == LIS_MATRIX_BSC) { Aout->conv_bnr = Ain->bnr; Aout->conv_bnc = Ain->bnc; }
    err = from_csr(T, Aout);
    lis_matrix_destroy(T);
    return err;
}

LIS_INT lis_matrix_convert(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    LIS_INT err = lis_host_matrix_check_input(Ain);
    if (err) return err;
    if (!lis_is_malloc(Aout)) { LIS_SETERR(LIS_ERR_ILL_ARG, "matrix Aout is undefined\n"); return LIS_ERR_ILL_ARG; }
    err = lis_matrix_merge(Ain);
    if (err) return err;
    const LIS_INT target = Aout->matrix_type;
    if (Ain->matrix_type == target && !Ain->is_block) return lis_matrix_copy(Ain, Aout);
    if (Ain->matrix_type == LIS_MATRIX_CSR) return from_csr(Ain, Aout);
    if (target == LIS_MATRIX_CSR) return to_csr(Ain, Aout);
    LIS_MATRIX T;
    err = lis_matrix_duplicate(Ain, &T);
    if (err) return err;
    err = to_csr(Ain, T);
    if (!err) err = from_csr(T, Aout);
    lis_matrix_destroy(T);
    return err;
}

/* -storage fmt: converts the solver's matrix IN PLACE (src/matrix/lis_matrix_ops.c:325-368) */
LIS_INT lis_matrix_convert_self(LIS_SOLVER solver)
{
    LIS_MATRIX A = solver->A, B;
    const LIS_INT storage = solver->options[LIS_OPTIONS_STORAGE];
    const LIS_INT block = solver->options[LIS_OPTIONS_STORAGE_BLOCK];
    if (storage > 0 && A->matrix_type != storage) {
        LIS_INT err = lis_matrix_duplicate(A, &B);
        if (err) return err;
        lis_matrix_set_blocksize(B, block, block, NULL, NULL);
        lis_matrix_set_type(B, storage);
        err = lis_matrix_convert(A, B);
        if (err) { lis_matrix_destroy(B); return err; }
        lis_host_matrix_adopt(A, B);
    }
    return LIS_SUCCESS;
}
