/*
 * lis_convert_dev.c -- CSR -> ELL / DIA / JAD / BSR on the device (kernels/convert.cu), behind
 * lis_matrix_convert.  The reference builds these layouts with host loops
 * (src/matrix/lis_matrix_ell.c:957, lis_matrix_dia.c:1190, lis_matrix_jad.c:1590,
 * lis_matrix_bsr.c:350); host/lis_convert.c restates them.  Here the CSR mirror in HBM is
 * rearranged by kernels INTO MANAGED MEMORY that is at once the public arrays of Aout (the lis.h
 * struct exposes them, the host may read them -- pages migrate on demand -- and lis_matrix_destroy
 * frees them through lis_free like any array the library allocated) and the device mirror of
 * Aout: nothing is downloaded, the first lis_matvec on Aout uploads nothing.  (Round 1 downloaded
 * the result into freshly malloc'ed pageable arrays: 5.4 s for ELL at 512^3, no faster than the
 * host builder.)  Where managed memory is refused the old download path runs.
 *
 * Default; LIS_B200_CONVERT=host selects the host builders (host/lis_convert.c).  Any case the
 * kernels do not cover (more than 255 entries in a row for JAD, more than 64 blocks in a block row
 * for BSR, row-partitioned BSR) reports *done = 0 and the host builder runs.  Output arrays are
 * identical to the host builder's, entry for entry (tests/test_z1_gpu_parity2.py, tests/test_emu_kernels.py).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lis_device.h"
#include "lis_host.h"
#include "lis_b200_kernels.h"

int lisd_convert_on_device(void)
{
    const char *e = getenv("LIS_B200_CONVERT");
    return !(e && strcmp(e, "host") == 0) && lisd_available();
}

#define CK(call, what) do { err = lisd_check((call), what); if (err) goto out; } while (0)
#define CKE(call) do { err = (call); if (err) goto out; } while (0)

static LIS_INT lis_host_malloc_array(void **p, size_t bytes)
{
    *p = lis_malloc(bytes ? bytes : 1, "lis_convert_dev::array");
    if (*p == NULL) { LIS_SETERR_MEM(bytes); return LIS_OUT_OF_MEMORY; }
    return LIS_SUCCESS;
}

static LIS_INT dmalloc_pad(void **p, size_t bytes, size_t pad)
{
    LIS_INT err = lisd_malloc(p, bytes + pad);
    if (!err && pad) err = lisd_memset((char *)*p + bytes, 0, pad);
    return err;
}

/* an output array of `bytes` (+pad readable zero bytes): managed memory doubling as the public array when
 * available (*pub = *dev), else plain device memory (*pub = NULL: the caller downloads into a host array) */
static LIS_INT out_alloc(void **dev, void **pub, size_t bytes, size_t pad)
{
    static int off = -1;
    if (off < 0) { const char *e = getenv("LIS_B200_CONVERT_SHARED"); off = (e && e[0] == '0') ? 1 : 0; }
    *pub = off ? NULL : lisd_shared_alloc(bytes + pad);
    if (*pub) { *dev = *pub; return pad ? lisd_memset((char *)*dev + bytes, 0, pad) : LIS_SUCCESS; }
    return dmalloc_pad(dev, bytes, pad);
}

static LIS_INT finish_dev(LIS_MATRIX Aout, lisd_matrix *M)
{
    LIS_INT err = lis_matrix_assemble(Aout);
    if (err) { lis_matrix_storage_destroy(Aout); lisd_mirror_free(M); return err; }
    lisd_matrix_drop(Aout);
    M->type = Aout->matrix_type; M->n = Aout->n; M->np = Aout->np; M->splited = 0;
    Aout->b200_dev = M;
    return LIS_SUCCESS;
}

static LIS_INT max_row_len(lisd_matrix *S, int n, int *out)
{
    int *d = NULL;
    LIS_INT err = lisd_malloc((void **)&d, 16);
    if (err) return err;
    lisd_mark_busy();
    err = lisd_check(lisb200_csr_max_row_len(n, S->csr.ptr, d, lisd_stream()), "row length scan");
    if (!err) err = lisd_download(out, d, sizeof(int));
    lisd_free(d);
    return err;
}

static LIS_INT dev_csr2ell(LIS_MATRIX Ain, lisd_matrix *S, LIS_MATRIX Aout)
{
    const int n = Ain->n;
    int maxnzr = 0;
    LIS_INT err, *index = NULL;
    LIS_SCALAR *value = NULL;
    lisd_matrix *M = (lisd_matrix *)calloc(1, sizeof(lisd_matrix));
    if (!M) { LIS_SETERR_MEM(sizeof(lisd_matrix)); return LIS_OUT_OF_MEMORY; }
    CKE(max_row_len(S, n, &maxnzr));
    const size_t cnt = (size_t)n * (size_t)maxnzr;
    CKE(out_alloc((void **)&M->idx, (void **)&index, cnt * sizeof(int), 16));
    if (index) M->shared |= LISD_SH_IDX;
    CKE(out_alloc((void **)&M->val, (void **)&value, cnt * sizeof(double), 16));
    if (value) M->shared |= LISD_SH_VAL;
    M->maxnzr = maxnzr; M->ld = n;
    lisd_mark_busy();
    CK(lisb200_csr2ell(n, maxnzr, n, S->csr.ptr, S->csr.idx, S->csr.val, M->idx, M->val, lisd_stream()), "csr2ell");
    if (!index) { CKE(lis_host_malloc_array((void **)&index, cnt * sizeof(int))); CKE(lisd_download(index, M->idx, cnt * sizeof(int))); }
    if (!value) { CKE(lis_host_malloc_array((void **)&value, cnt * sizeof(double))); CKE(lisd_download(value, M->val, cnt * sizeof(double))); }
    CKE(lisd_sync());
    CKE(lis_matrix_set_ell(maxnzr, index, value, Aout));
    index = NULL; value = NULL;
    Aout->nnz = Ain->ptr[n];
    return finish_dev(Aout, M);
out:
    lis_free2(2, index, value);
    lisd_mirror_free(M);
    return err;
}

static LIS_INT dev_csr2dia(LIS_MATRIX Ain, lisd_matrix *S, LIS_MATRIX Aout)
{
    const int n = Ain->n, np = Ain->np;
    const int nseg = lisb200_dia_segments(n, np);
    LIS_INT err, *index = NULL;
    LIS_SCALAR *value = NULL;
    unsigned char *d_flags = NULL;
    int *d_cnt = NULL, *d_base = NULL, *h_cnt = NULL;
    int nnd = 0;
    lisd_matrix *M = (lisd_matrix *)calloc(1, sizeof(lisd_matrix));
    if (!M) { LIS_SETERR_MEM(sizeof(lisd_matrix)); return LIS_OUT_OF_MEMORY; }
    h_cnt = (int *)malloc(sizeof(int) * (size_t)(nseg > 0 ? nseg : 1));
    if (!h_cnt) { LIS_SETERR_MEM(nseg * sizeof(int)); err = LIS_OUT_OF_MEMORY; goto out; }
    CKE(lisd_malloc((void **)&d_flags, (size_t)n + (size_t)np));
    CKE(lisd_malloc((void **)&d_cnt, sizeof(int) * (size_t)nseg));
    CKE(lisd_malloc((void **)&d_base, sizeof(int) * (size_t)nseg));
    lisd_mark_busy();
    CK(lisb200_csr2dia_mark(n, np, S->csr.ptr, S->csr.idx, d_flags, d_cnt, lisd_stream()), "csr2dia (offsets)");
    CKE(lisd_download(h_cnt, d_cnt, sizeof(int) * (size_t)nseg));
    for (int s = 0; s < nseg; s++) { const int c = h_cnt[s]; h_cnt[s] = nnd; nnd += c; }      /* exclusive prefix */
    CKE(lisd_upload(d_base, h_cnt, sizeof(int) * (size_t)nseg));
    const size_t cnt = (size_t)n * (size_t)nnd;
    CKE(dmalloc_pad((void **)&M->off, (size_t)nnd * sizeof(int), 16));
    CKE(out_alloc((void **)&M->val, (void **)&value, cnt * sizeof(double), 16));
    if (value) M->shared |= LISD_SH_VAL;
    M->nnd = nnd; M->ld = n;
    lisd_mark_busy();
    CK(lisb200_csr2dia_fill(n, np, nnd, n, S->csr.ptr, S->csr.idx, S->csr.val, d_flags, d_base, d_cnt, M->off, M->val, lisd_stream()), "csr2dia");
    CKE(lis_host_malloc_array((void **)&index, (size_t)nnd * sizeof(int)));
    CKE(lisd_download(index, M->off, (size_t)nnd * sizeof(int)));
    if (!value) { CKE(lis_host_malloc_array((void **)&value, cnt * sizeof(double))); CKE(lisd_download(value, M->val, cnt * sizeof(double))); }
    CKE(lisd_sync());
    CKE(lis_matrix_set_dia(nnd, index, value, Aout));
    index = NULL; value = NULL;
    Aout->nnz = Ain->ptr[n];
    lisd_free(d_flags); lisd_free(d_cnt); lisd_free(d_base); free(h_cnt);
    return finish_dev(Aout, M);
out:
    lisd_free(d_flags); lisd_free(d_cnt); lisd_free(d_base); free(h_cnt);
    lis_free2(2, index, value);
    lisd_mirror_free(M);
    return err;
}

static LIS_INT dev_csr2jad(LIS_MATRIX Ain, lisd_matrix *S, LIS_MATRIX Aout, int *done)
{
    const int n = Ain->n, nnz = Ain->ptr[n];
    const int bins = lisb200_jad_bins(), nctas = lisb200_jad_ctas(n);
    int maxnzr = 0;
    LIS_INT err, *perm = NULL, *ptr = NULL, *index = NULL;
    LIS_SCALAR *value = NULL;
    int *d_tab = NULL, *h_tab = NULL;
    lisd_matrix *M = (lisd_matrix *)calloc(1, sizeof(lisd_matrix));
    if (!M) { LIS_SETERR_MEM(sizeof(lisd_matrix)); return LIS_OUT_OF_MEMORY; }
    CKE(max_row_len(S, n, &maxnzr));
    if (maxnzr >= bins) { *done = 0; err = LIS_SUCCESS; goto out; }           /* host builder */
    const size_t tab = (size_t)nctas * (size_t)bins;
    h_tab = (int *)malloc(sizeof(int) * (tab > 0 ? tab : 1));
    if (!h_tab) { LIS_SETERR_MEM(tab * sizeof(int)); err = LIS_OUT_OF_MEMORY; goto out; }
    CKE(lisd_malloc((void **)&d_tab, sizeof(int) * tab));
    lisd_mark_busy();
    CK(lisb200_csr2jad_hist(n, maxnzr, S->csr.ptr, d_tab, lisd_stream()), "csr2jad (histogram)");
    CKE(lisd_download(h_tab, d_tab, sizeof(int) * tab));
    CKE(lis_host_malloc_array((void **)&ptr, ((size_t)maxnzr + 1) * sizeof(int)));
    {
        /* rows per bin -> jagged-diagonal pointers: diagonal j holds the rows longer than j;
         * start position of (bin, cta): bins ascending (= length descending), CTAs in order */
        long long *tot = (long long *)calloc((size_t)bins, sizeof(long long));
        if (!tot) { LIS_SETERR_MEM(bins * sizeof(long long)); err = LIS_OUT_OF_MEMORY; goto out; }
        for (int c = 0; c < nctas; c++)
            for (int b = 0; b < bins; b++) tot[b] += h_tab[(size_t)c * bins + b];
        ptr[0] = 0;
        for (int j = 0; j < maxnzr; j++) {
            long long rows = 0;                                   /* rows with length >= j+1 <=> bin <= maxnzr-j-1 */
            for (int b = 0; b <= maxnzr - j - 1; b++) rows += tot[b];
            ptr[j + 1] = ptr[j] + (LIS_INT)rows;
        }
        long long start = 0;
        for (int b = 0; b < bins; b++) {
            long long run = start;
            for (int c = 0; c < nctas; c++) {
                const int k = h_tab[(size_t)c * bins + b];
                h_tab[(size_t)c * bins + b] = (int)run;
                run += k;
            }
            start += tot[b];
        }
        free(tot);
    }
    CKE(lisd_upload(d_tab, h_tab, sizeof(int) * tab));
    CKE(dmalloc_pad((void **)&M->jptr, ((size_t)maxnzr + 1) * sizeof(int), 16));
    CKE(out_alloc((void **)&M->perm, (void **)&perm, (size_t)n * sizeof(int), 16));
    if (perm) M->shared |= LISD_SH_PERM;
    CKE(out_alloc((void **)&M->idx, (void **)&index, (size_t)nnz * sizeof(int), 16));
    if (index) M->shared |= LISD_SH_IDX;
    CKE(out_alloc((void **)&M->val, (void **)&value, (size_t)nnz * sizeof(double), 16));
    if (value) M->shared |= LISD_SH_VAL;
    M->maxnzr = maxnzr;
    CKE(lisd_upload(M->jptr, ptr, ((size_t)maxnzr + 1) * sizeof(int)));
    lisd_mark_busy();
    CK(lisb200_csr2jad_fill(n, maxnzr, S->csr.ptr, S->csr.idx, S->csr.val, d_tab, M->jptr, M->perm, M->idx, M->val, lisd_stream()), "csr2jad");
    if (!perm) { CKE(lis_host_malloc_array((void **)&perm, (size_t)n * sizeof(int))); CKE(lisd_download(perm, M->perm, (size_t)n * sizeof(int))); }
    if (!index) { CKE(lis_host_malloc_array((void **)&index, (size_t)nnz * sizeof(int))); CKE(lisd_download(index, M->idx, (size_t)nnz * sizeof(int))); }
    if (!value) { CKE(lis_host_malloc_array((void **)&value, (size_t)nnz * sizeof(double))); CKE(lisd_download(value, M->val, (size_t)nnz * sizeof(double))); }
    CKE(lisd_sync());
    CKE(lis_matrix_set_jad(nnz, maxnzr, perm, ptr, index, value, Aout));
    perm = NULL; ptr = NULL; index = NULL; value = NULL;
    lisd_free(d_tab); free(h_tab);
    return finish_dev(Aout, M);
out:
    lisd_free(d_tab); free(h_tab);
    lis_free2(4, perm, ptr, index, value);
    lisd_mirror_free(M);
    return err;
}

static LIS_INT dev_csr2bsr(LIS_MATRIX Ain, lisd_matrix *S, LIS_MATRIX Aout, int *done)
{
    const int n = Ain->n;
    const int bnr = Aout->conv_bnr, bnc = Aout->conv_bnc, bs = bnr * bnc;
    const int nr = 1 + (n - 1) / bnr;
    LIS_INT err, *bptr = NULL, *bindex = NULL;
    LIS_SCALAR *value = NULL;
    int *d_cnt = NULL, *d_over = NULL, *h_cnt = NULL;
    int over = 0;
    lisd_matrix *M = NULL;
    if (Ain->np != n || n <= 0) { *done = 0; return LIS_SUCCESS; }           /* host builder reports / handles these */
    M = (lisd_matrix *)calloc(1, sizeof(lisd_matrix));
    if (!M) { LIS_SETERR_MEM(sizeof(lisd_matrix)); return LIS_OUT_OF_MEMORY; }
    h_cnt = (int *)malloc(sizeof(int) * ((size_t)nr + 1));
    if (!h_cnt) { LIS_SETERR_MEM(nr * sizeof(int)); err = LIS_OUT_OF_MEMORY; goto out; }
    CKE(lisd_malloc((void **)&d_cnt, sizeof(int) * ((size_t)nr + 1)));
    CKE(lisd_malloc((void **)&d_over, 16));
    lisd_mark_busy();
    CK(lisb200_csr2bsr_count(n, nr, bnr, bnc, S->csr.ptr, S->csr.idx, d_cnt, d_over, lisd_stream()), "csr2bsr (block count)");
    CKE(lisd_download(&over, d_over, sizeof(int)));
    if (over) { *done = 0; err = LIS_SUCCESS; goto out; }
    CKE(lisd_download(h_cnt, d_cnt, sizeof(int) * (size_t)nr));
    {
        long long run = 0;
        for (int b = 0; b < nr; b++) { const int c = h_cnt[b]; h_cnt[b] = (int)run; run += c; }
        h_cnt[nr] = (int)run;
        if (run * bs > 0x7fffffffLL) { *done = 0; err = LIS_SUCCESS; goto out; }     /* 32-bit LIS_INT: let the host report it */
    }
    const int bnnz = h_cnt[nr];
    CKE(lis_host_malloc_array((void **)&bptr, ((size_t)nr + 1) * sizeof(int)));
    memcpy(bptr, h_cnt, sizeof(int) * ((size_t)nr + 1));
    CKE(dmalloc_pad((void **)&M->bptr, ((size_t)nr + 1) * sizeof(int), 16));
    CKE(out_alloc((void **)&M->bidx, (void **)&bindex, (size_t)bnnz * sizeof(int), 16));
    if (bindex) M->shared |= LISD_SH_BIDX;
    CKE(out_alloc((void **)&M->val, (void **)&value, (size_t)bnnz * (size_t)bs * sizeof(double), 16));
    if (value) M->shared |= LISD_SH_VAL;
    M->nr = nr; M->bnr = bnr; M->bnc = bnc; M->bnnz = bnnz;
    CKE(lisd_upload(M->bptr, bptr, ((size_t)nr + 1) * sizeof(int)));
    lisd_mark_busy();
    CK(lisb200_csr2bsr_fill(n, nr, bnr, bnc, S->csr.ptr, S->csr.idx, S->csr.val, M->bptr, M->bidx, M->val, lisd_stream()), "csr2bsr");
    if (!bindex) { CKE(lis_host_malloc_array((void **)&bindex, (size_t)bnnz * sizeof(int))); CKE(lisd_download(bindex, M->bidx, (size_t)bnnz * sizeof(int))); }
    if (!value) { CKE(lis_host_malloc_array((void **)&value, (size_t)bnnz * (size_t)bs * sizeof(double))); CKE(lisd_download(value, M->val, (size_t)bnnz * (size_t)bs * sizeof(double))); }
    CKE(lisd_sync());
    CKE(lis_matrix_set_bsr(bnr, bnc, bnnz, bptr, bindex, value, Aout));
    bptr = NULL; bindex = NULL; value = NULL;
    Aout->nnz = Ain->ptr[n];
    lisd_free(d_cnt); lisd_free(d_over); free(h_cnt);
    return finish_dev(Aout, M);
out:
    lisd_free(d_cnt); lisd_free(d_over); free(h_cnt);
    lis_free2(3, bptr, bindex, value);
    lisd_mirror_free(M);
    return err;
}

/* Acsr: assembled, unsplit CSR.  *done = 1 when Aout was built here (or an error is returned),
 * 0 when the host builder has to run. */
LIS_INT lisd_convert_from_csr(LIS_MATRIX Acsr, LIS_MATRIX Aout, int *done)
{
    lisd_matrix *S;
    LIS_INT err;
    *done = 0;
    if (!lisd_convert_on_device() || Acsr->is_splited || Acsr->n <= 0) return LIS_SUCCESS;
    switch (Aout->matrix_type) {
    case LIS_MATRIX_ELL: case LIS_MATRIX_DIA: case LIS_MATRIX_JAD: case LIS_MATRIX_BSR: break;
    default: return LIS_SUCCESS;
    }
    err = lisd_matrix_get(Acsr, &S);
    if (err) return err;
    if (Aout->matrix_type == LIS_MATRIX_DIA && !Acsr->is_sorted) {
        /* lis_matrix_dia.c:1217 sorts Ain's rows (and so do we, on the host arrays the caller sees); an input
         * whose rows are ascending already -- checked on the mirror, 1 ms instead of seconds of host work at
         * 512^3 -- only gets the flag */
        int unsorted = 1, *d_flag = NULL;
        if (!lisd_malloc((void **)&d_flag, 16)) {
            lisd_mark_busy();
            if (lisd_check(lisb200_csr_rows_unsorted(Acsr->n, S->csr.ptr, S->csr.idx, d_flag, lisd_stream()), "sortedness check") ||
                lisd_download(&unsorted, d_flag, sizeof(int))) unsorted = 1;
            lisd_free(d_flag);
        }
        if (unsorted) {
            lis_matrix_sort_csr(Acsr);                     /* drops the mirror */
            err = lisd_matrix_get(Acsr, &S);
            if (err) return err;
        } else Acsr->is_sorted = LIS_TRUE;
    }
    *done = 1;
    switch (Aout->matrix_type) {
    case LIS_MATRIX_ELL: return dev_csr2ell(Acsr, S, Aout);
    case LIS_MATRIX_DIA: return dev_csr2dia(Acsr, S, Aout);
    case LIS_MATRIX_JAD: return dev_csr2jad(Acsr, S, Aout, done);
    default:             return dev_csr2bsr(Acsr, S, Aout, done);
    }
}
