/*
 * lis_vector.c -- LIS_VECTOR objects and the lis_vector_* BLAS-1 API.
 *
 * Host C; every arithmetic loop of the reference (src/vector/lis_vector_opv.c,
 * lis_vector_ops.c) is a CUDA kernel behind include/lis_b200_kernels.h.  The public entry
 * points are host-synchronous like the reference; the lisd_* variants used by the Krylov
 * loops only enqueue work on the library's stream.
 *
 * Object management follows src/vector/lis_vector.c:115-420 of the reference (argument
 * checks, partition, zero initialisation, duplicate-from-matrix).
 */
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lis_device.h"
#include "lis_host.h"
#include "lis_b200_kernels.h"

/* ------------------------------------------------------------------ object management */
static void vector_init(LIS_VECTOR v)
{
    memset(v, 0, sizeof(struct LIS_VECTOR_STRUCT));
    v->label = LIS_LABEL_VECTOR;
    v->status = LIS_VECTOR_NULL;
    v->is_destroy = LIS_TRUE;
    v->is_copy = LIS_FALSE;
    v->origin = LIS_ORIGIN_0;
    v->nprocs = lisd_nranks();
    v->my_rank = lisd_rank();
}

static LIS_INT vector_alloc_zero(LIS_VECTOR v, size_t count)
{
    LIS_INT managed = 0;
    LIS_INT err = lisd_alloc_vector(count, &v->value, &managed);
    if (err) return err;
    v->b200_managed = managed;
    v->b200_capacity = count > 0 ? count : 1;
    if (managed) {
        /* zero on the device: the pages are born in HBM (managed == 2: device-only fallback) */
        err = lisd_memset(v->value, 0, v->b200_capacity * sizeof(LIS_SCALAR));
        if (err) return err;
        v->b200_resident = 1;
    } else {
        memset(v->value, 0, v->b200_capacity * sizeof(LIS_SCALAR));
        v->b200_resident = 0;
    }
    v->is_copy = LIS_TRUE;
    return LIS_SUCCESS;
}

LIS_INT lis_vector_create(LIS_Comm comm, LIS_VECTOR *vec)
{
    *vec = (LIS_VECTOR)lis_malloc(sizeof(struct LIS_VECTOR_STRUCT), "lis_vector_create::vec");
    if (*vec == NULL) { LIS_SETERR_MEM(sizeof(struct LIS_VECTOR_STRUCT)); return LIS_OUT_OF_MEMORY; }
    vector_init(*vec);
    (*vec)->comm = comm;
    return LIS_SUCCESS;
}

LIS_INT lis_vector_set_size(LIS_VECTOR vec, LIS_INT local_n, LIS_INT global_n)
{
    LIS_INT nprocs, my_rank, is, ie, err;
    LIS_INT *ranges;

    if (global_n > 0 && local_n > global_n) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "local n(=%D) is larger than global n(=%D)\n", local_n, global_n);
        return LIS_ERR_ILL_ARG;
    }
    if (local_n < 0 || global_n < 0) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "local n(=%D) or global n(=%D) are less than 0\n", local_n, global_n);
        return LIS_ERR_ILL_ARG;
    }
    err = lis_ranges_create(vec->comm, &local_n, &global_n, &ranges, &is, &ie, &nprocs, &my_rank);
    if (err) return err;
    vec->ranges = ranges;
    err = vector_alloc_zero(vec, (size_t)local_n);
    if (err) return err;
    vec->status = LIS_VECTOR_ASSEMBLED;
    vec->n = local_n; vec->gn = global_n; vec->np = local_n;
    vec->my_rank = my_rank; vec->nprocs = nprocs;
    vec->is = is; vec->ie = ie;
    return LIS_SUCCESS;
}

LIS_INT lis_vector_destroy(LIS_VECTOR vec)
{
    if (lis_is_malloc(vec)) {
        if (vec->value && vec->is_destroy) lisd_free_vector_bytes(vec->value, vec->b200_managed, vec->b200_capacity);
        if (vec->work) lis_free(vec->work);
        if (vec->ranges) lis_free(vec->ranges);
        lis_free(vec);
    }
    return LIS_SUCCESS;
}

/* accepts a vector or a matrix (common object header) */
LIS_INT lis_vector_duplicate(void *vin, LIS_VECTOR *vout)
{
    const LIS_VECTOR src = (LIS_VECTOR)vin;
    if (src->label != LIS_LABEL_VECTOR && src->label != LIS_LABEL_MATRIX) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "First argument is not LIS_VECTOR or LIS_MATRIX\n");
        return LIS_ERR_ILL_ARG;
    }
    LIS_VECTOR v = (LIS_VECTOR)lis_malloc(sizeof(struct LIS_VECTOR_STRUCT), "lis_vector_duplicate::vout");
    *vout = NULL;
    if (v == NULL) { LIS_SETERR_MEM(sizeof(struct LIS_VECTOR_STRUCT)); return LIS_OUT_OF_MEMORY; }
    vector_init(v);
    LIS_INT err = vector_alloc_zero(v, (size_t)src->np + (size_t)src->pad);
    if (err) { lis_free(v); return err; }
    if (src->nprocs > 1 && src->ranges) {
        v->ranges = (LIS_INT *)lis_malloc((size_t)(src->nprocs + 1) * sizeof(LIS_INT), "lis_vector_duplicate::ranges");
        if (v->ranges == NULL) { lis_vector_destroy(v); return LIS_OUT_OF_MEMORY; }
        memcpy(v->ranges, src->ranges, (size_t)(src->nprocs + 1) * sizeof(LIS_INT));
    }
    v->status = LIS_VECTOR_ASSEMBLED;
    v->precision = LIS_PRECISION_DEFAULT;
    v->n = src->n; v->gn = src->gn; v->np = src->np; v->pad = src->pad;
    v->comm = src->comm; v->my_rank = src->my_rank; v->nprocs = src->nprocs;
    v->is = src->is; v->ie = src->ie; v->origin = src->origin;
    *vout = v;
    return LIS_SUCCESS;
}

LIS_INT lis_vector_get_size(LIS_VECTOR v, LIS_INT *local_n, LIS_INT *global_n)
{
    *local_n = v->n; *global_n = v->gn;
    return LIS_SUCCESS;
}

LIS_INT lis_vector_get_range(LIS_VECTOR v, LIS_INT *is, LIS_INT *ie)
{
    *is = v->is; *ie = v->ie;
    return LIS_SUCCESS;
}

LIS_INT lis_vector_psd_reset_scale(LIS_VECTOR vec) { vec->is_scaled = LIS_FALSE; return LIS_SUCCESS; }

LIS_INT lis_vector_is_null(LIS_VECTOR v)
{
    return (v == NULL || !lis_is_malloc(v) || v->status == LIS_VECTOR_NULL) ? LIS_TRUE : LIS_FALSE;
}

/* ------------------------------------------------------------------ host element access */
LIS_INT lis_vector_get_value(LIS_VECTOR v, LIS_INT i, LIS_SCALAR *value)
{
    if (v->origin) i--;
    if (i < v->is || i >= v->ie) {
        if (v->origin) i++;
        LIS_SETERR3(LIS_ERR_ILL_ARG, "i(=%D) is less than %D or larger than %D\n", i, v->is + v->origin, v->ie - 1 + v->origin);
        return LIS_ERR_ILL_ARG;
    }
    if (v->b200_managed && v->b200_resident) {      /* one element: read it out of HBM */
        LIS_INT err = lisd_sync();
        if (err) return err;
        return lisd_download(value, v->value + (i - v->is), sizeof(LIS_SCALAR));
    }
    lisd_vec_host(v);
    *value = v->value[i - v->is];
    return LIS_SUCCESS;
}

LIS_INT lis_vector_get_values(LIS_VECTOR v, LIS_INT start, LIS_INT count, LIS_SCALAR value[])
{
    if (v->origin) start--;
    if (start < v->is || start >= v->ie) {
        if (v->origin) start++;
        LIS_SETERR3(LIS_ERR_ILL_ARG, "start(=%D) is less than %D or larger than %D\n", start, v->is + v->origin, v->ie - 1 + v->origin);
        return LIS_ERR_ILL_ARG;
    }
    if (start - v->is + count > v->n) {
        LIS_SETERR3(LIS_ERR_ILL_ARG, "start(=%D) + count(=%D) exceeds the range of vector v(=%D)\n", start, count, v->n);
        return LIS_ERR_ILL_ARG;
    }
    if (v->b200_managed && v->b200_resident) {
        /* bulk read straight out of HBM; residency is kept */
        LIS_INT err = lisd_sync();
        if (err) return err;
        return lisd_download(value, v->value + (start - v->is), (size_t)count * sizeof(LIS_SCALAR));
    }
    lisd_vec_host(v);
    memcpy(value, v->value + (start - v->is), (size_t)count * sizeof(LIS_SCALAR));
    return LIS_SUCCESS;
}

LIS_INT lis_vector_set_value(LIS_INT flag, LIS_INT i, LIS_SCALAR value, LIS_VECTOR v)
{
    if (v->origin) i--;
    if (i < v->is || i >= v->ie) {
        if (v->origin) i++;
        LIS_SETERR3(LIS_ERR_ILL_ARG, "i(=%D) is less than %D or larger than %D\n", i, v->is + v->origin, v->ie - 1 + v->origin);
        return LIS_ERR_ILL_ARG;
    }
    if (v->status == LIS_VECTOR_NULL) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "vector v is undefined\n");
        return LIS_ERR_ILL_ARG;
    }
    if (v->b200_managed == 2) {                 /* device-only storage: one element through the copy engine */
        LIS_SCALAR cur = 0.0;
        LIS_INT err = lisd_sync();
        if (!err && flag != LIS_INS_VALUE) err = lisd_download(&cur, v->value + (i - v->is), sizeof(LIS_SCALAR));
        if (err) return err;
        cur = (flag == LIS_INS_VALUE) ? value : cur + value;
        return lisd_upload(v->value + (i - v->is), &cur, sizeof(LIS_SCALAR));
    }
    if (v->b200_resident) lisd_vec_host(v);
    if (flag == LIS_INS_VALUE) v->value[i - v->is] = value;
    else v->value[i - v->is] += value;
    return LIS_SUCCESS;
}

LIS_INT lis_vector_set_values(LIS_INT flag, LIS_INT count, LIS_INT index[], LIS_SCALAR value[], LIS_VECTOR v)
{
    for (LIS_INT k = 0; k < count; k++) {
        LIS_INT err = lis_vector_set_value(flag, index[k], value[k], v);
        if (err) return err;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_vector_set_values2(LIS_INT flag, LIS_INT start, LIS_INT count, LIS_SCALAR value[], LIS_VECTOR v)
{
    if (v->origin) start--;
    if (start < v->is || start - v->is + count > v->n) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "start(=%D), count(=%D) is outside the local range\n", start, count);
        return LIS_ERR_ILL_ARG;
    }
    if (flag == LIS_INS_VALUE && v->b200_managed && v->b200_resident) {
        LIS_INT err = lisd_sync();
        if (err) return err;
        return lisd_upload(v->value + (start - v->is), value, (size_t)count * sizeof(LIS_SCALAR));
    }
    if (v->b200_managed == 2) {
        LIS_SCALAR *tmp = (LIS_SCALAR *)malloc((size_t)(count > 0 ? count : 1) * sizeof(LIS_SCALAR));
        if (!tmp) { LIS_SETERR_MEM(count * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
        LIS_INT err = lisd_sync();
        if (!err) err = lisd_download(tmp, v->value + (start - v->is), (size_t)count * sizeof(LIS_SCALAR));
        for (LIS_INT k = 0; !err && k < count; k++) tmp[k] += value[k];
        if (!err) err = lisd_upload(v->value + (start - v->is), tmp, (size_t)count * sizeof(LIS_SCALAR));
        free(tmp);
        return err;
    }
    lisd_vec_host(v);
    if (flag == LIS_INS_VALUE) memcpy(v->value + (start - v->is), value, (size_t)count * sizeof(LIS_SCALAR));
    else for (LIS_INT k = 0; k < count; k++) v->value[start - v->is + k] += value[k];
    return LIS_SUCCESS;
}

/* whole-vector host <-> vector transfer (reference: src/vector/lis_vector.c:952-1067; there a
 * global array is scattered/gathered over MPI, here each rank moves its own slice) */
LIS_INT lis_vector_scatter(LIS_SCALAR value[], LIS_VECTOR v)
{
    if (v->n == 0) return LIS_SUCCESS;
    return lis_vector_set_values2(LIS_INS_VALUE, v->is + v->origin, v->n, value + v->is, v);
}

LIS_INT lis_vector_gather(LIS_VECTOR v, LIS_SCALAR value[])
{
    if (v->n == 0) return LIS_SUCCESS;
    if (v->nprocs == 1) return lis_vector_get_values(v, v->is + v->origin, v->n, value);
    /* every rank receives the whole vector, like MPI_Allgatherv in the reference */
    LIS_INT err = lis_vector_get_values(v, v->is + v->origin, v->n, value + v->is);
    if (err) return err;
    return lisd_allgatherv_host(value, v->ranges, v->nprocs);
}

LIS_INT lis_vector_print(LIS_VECTOR x)
{
    LIS_SCALAR *hv = lisd_vec_host_view(x, 1);
    if (hv == NULL) return LIS_ERR_OUT_OF_MEMORY;
    for (LIS_INT i = 0; i < x->n; i++)
        printf("%6d  %e\n", (int)(i + x->is + x->origin), (double)hv[i]);
    return lisd_vec_host_done(x, hv, 0);
}

/* ------------------------------------------------------------------ BLAS-1 on the device */
LIS_INT lis_vector_check_same(LIS_VECTOR x, LIS_VECTOR y)
{
    if (x->n != y->n) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "length of vector x and y is not equal\n");
        return LIS_ERR_ILL_ARG;
    }
    return LIS_SUCCESS;
}

#define LISD_PREP1(x, what)                                   \
    do {                                                      \
        LIS_INT e_ = lisd_require(what);                      \
        if (e_) return e_;                                    \
        e_ = lisd_vec_device(x);                              \
        if (e_) return e_;                                    \
    } while (0)
#define LISD_PREP2(x, y, what)                                \
    do {                                                      \
        LIS_INT e_ = lis_vector_check_same(x, y);             \
        if (e_) return e_;                                    \
        LISD_PREP1(x, what);                                  \
        e_ = lisd_vec_device(y);                              \
        if (e_) return e_;                                    \
    } while (0)
#define LISD_LAUNCH(call, what) do { lisd_mark_busy(); return lisd_check((call), what); } while (0)

LIS_INT lisd_copy(LIS_VECTOR x, LIS_VECTOR y)
{
    LISD_PREP2(x, y, "lis_vector_copy");
    LISD_LAUNCH(lisb200_copy(x->n, x->value, y->value, lisd_stream()), "lis_vector_copy");
}
LIS_INT lisd_axpy(LIS_SCALAR alpha, LIS_VECTOR x, LIS_VECTOR y)
{
    LISD_PREP2(x, y, "lis_vector_axpy");
    LISD_LAUNCH(lisb200_axpy(x->n, alpha, x->value, y->value, lisd_stream()), "lis_vector_axpy");
}
LIS_INT lisd_xpay(LIS_VECTOR x, LIS_SCALAR alpha, LIS_VECTOR y)
{
    LISD_PREP2(x, y, "lis_vector_xpay");
    LISD_LAUNCH(lisb200_xpay(x->n, x->value, alpha, y->value, lisd_stream()), "lis_vector_xpay");
}
LIS_INT lisd_axpyz(LIS_SCALAR alpha, LIS_VECTOR x, LIS_VECTOR y, LIS_VECTOR z)
{
    if (x->n != y->n || x->n != z->n) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "length of vector x and y and z is not equal\n");
        return LIS_ERR_ILL_ARG;
    }
    LISD_PREP2(x, y, "lis_vector_axpyz");
    { LIS_INT e_ = lisd_vec_device(z); if (e_) return e_; }
    LISD_LAUNCH(lisb200_axpyz(x->n, alpha, x->value, y->value, z->value, lisd_stream()), "lis_vector_axpyz");
}
LIS_INT lisd_scale(LIS_SCALAR alpha, LIS_VECTOR x)
{
    LISD_PREP1(x, "lis_vector_scale");
    LISD_LAUNCH(lisb200_scale(x->n, alpha, x->value, lisd_stream()), "lis_vector_scale");
}
LIS_INT lisd_set_all(LIS_SCALAR alpha, LIS_VECTOR x)
{
    LISD_PREP1(x, "lis_vector_set_all");
    LISD_LAUNCH(lisb200_set_all(x->n, alpha, x->value, lisd_stream()), "lis_vector_set_all");
}
LIS_INT lisd_pmul(LIS_VECTOR x, LIS_VECTOR y, LIS_VECTOR z)
{
    LISD_PREP2(x, y, "lis_vector_pmul");
    { LIS_INT e_ = lis_vector_check_same(x, z); if (e_) return e_; e_ = lisd_vec_device(z); if (e_) return e_; }
    LISD_LAUNCH(lisb200_pmul(x->n, x->value, y->value, z->value, lisd_stream()), "lis_vector_pmul");
}

/* wait for the enqueued reduction kernel, read `count` mapped scalars, combine across ranks */
LIS_INT lisd_reduce_finish(double *vals, int count, int is_max)
{
    if (lisd_reduce_uses_nccl()) return lisd_reduce_nccl_finish(vals, count, is_max);
    LIS_INT err = lisd_sync();
    if (err) return err;
    for (int k = 0; k < count; k++) vals[k] = lisd_scalar_get(k);
    if (lisd_nranks() > 1) return is_max ? lisd_allreduce_max(vals, count) : lisd_allreduce_sum(vals, count);
    return LIS_SUCCESS;
}

/* ---- fused Krylov steps ---- */
LIS_INT lisd_jacobi_dot(LIS_VECTOR r, LIS_VECTOR dinv, LIS_VECTOR z, LIS_SCALAR *rho)
{
    LISD_PREP2(r, dinv, "jacobi+dot");
    { LIS_INT e_ = lis_vector_check_same(r, z); if (e_) return e_; e_ = lisd_vec_device(z); if (e_) return e_; }
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    LIS_INT err = lisd_check(lisb200_jacobi_dot(r->n, r->value, dinv->value, z->value, partial, lisd_counter(),
                                                lisd_scalar_dev(0), lisd_stream()), "jacobi+dot");
    if (err) return err;
    return lisd_reduce_finish(rho, 1, 0);
}

/* cg update + the next iteration's Jacobi psolve and <r,z>; *fused_out = 0 (nothing done) when the
 * pointers' alignment is mixed and the caller has to take the two separate steps */
LIS_INT lisd_cg_update_jacobi(LIS_SCALAR alpha, LIS_VECTOR p, LIS_VECTOR q, LIS_VECTOR x, LIS_VECTOR r, LIS_VECTOR dinv, LIS_VECTOR z,
                              LIS_REAL *nrm2_r, LIS_SCALAR *rho, int *fused_out)
{
    LISD_PREP2(p, q, "cg update");
    { LIS_INT e_ = lis_vector_check_same(p, x); if (e_) return e_; e_ = lis_vector_check_same(p, r); if (e_) return e_;
      e_ = lis_vector_check_same(p, dinv); if (e_) return e_; e_ = lis_vector_check_same(p, z); if (e_) return e_;
      e_ = lisd_vec_device(x); if (e_) return e_; e_ = lisd_vec_device(r); if (e_) return e_;
      e_ = lisd_vec_device(dinv); if (e_) return e_; e_ = lisd_vec_device(z); if (e_) return e_; }
    *fused_out = 0;
    if ((((uintptr_t)p->value | (uintptr_t)q->value | (uintptr_t)x->value | (uintptr_t)r->value | (uintptr_t)dinv->value | (uintptr_t)z->value) & 15) != 0)
        return LIS_SUCCESS;
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    LIS_INT err = lisd_check(lisb200_cg_update_jacobi(p->n, alpha, p->value, q->value, x->value, r->value, dinv->value, z->value, partial,
                                                      lisd_counter(), lisd_scalar_dev(0), lisd_stream()), "cg update + jacobi");
    if (err) return err;
    double v[2];
    err = lisd_reduce_finish(v, 2, 0);
    if (err) return err;
    *nrm2_r = sqrt(v[0]);
    *rho = v[1];
    *fused_out = 1;
    return LIS_SUCCESS;
}

LIS_INT lisd_cg_update(LIS_SCALAR alpha, LIS_VECTOR p, LIS_VECTOR q, LIS_VECTOR x, LIS_VECTOR r, LIS_REAL *nrm2_r)
{
    LISD_PREP2(p, q, "cg update");
    { LIS_INT e_ = lis_vector_check_same(p, x); if (e_) return e_; e_ = lis_vector_check_same(p, r); if (e_) return e_;
      e_ = lisd_vec_device(x); if (e_) return e_; e_ = lisd_vec_device(r); if (e_) return e_; }
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    LIS_INT err = lisd_check(lisb200_cg_update(p->n, alpha, p->value, q->value, x->value, r->value, partial, lisd_counter(),
                                               lisd_scalar_dev(0), lisd_stream()), "cg update");
    if (err) return err;
    double rr;
    err = lisd_reduce_finish(&rr, 1, 0);
    if (err) return err;
    *nrm2_r = sqrt(rr);
    return LIS_SUCCESS;
}

LIS_INT lisd_dot2(LIS_VECTOR a, LIS_VECTOR b, LIS_SCALAR out[2])
{
    LISD_PREP2(a, b, "dot2");
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    LIS_INT err = lisd_check(lisb200_dot2(a->n, a->value, b->value, partial, lisd_counter(), lisd_scalar_dev(0), lisd_stream()), "dot2");
    if (err) return err;
    return lisd_reduce_finish(out, 2, 0);
}

/* ---- reductions that stay on the device (single rank): GMRES' modified Gram-Schmidt ---- */
LIS_INT lisd_dot_to_slot(LIS_VECTOR x, LIS_VECTOR y, int slot)
{
    LISD_PREP2(x, y, "lis_vector_dot");
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    return lisd_check(lisb200_reduce(0, x->n, x->value, y->value, partial, lisd_counter(), lisd_dev_scalar(slot), lisd_stream()),
                      "lis_vector_dot");
}

LIS_INT lisd_axpy_from_slot(int slot, double scale, LIS_VECTOR x, LIS_VECTOR y)
{
    LISD_PREP2(x, y, "lis_vector_axpy");
    LISD_LAUNCH(lisb200_axpy_dev(x->n, lisd_dev_scalar(slot), scale, x->value, y->value, lisd_stream()), "lis_vector_axpy");
}

/* y += (scale * slot_in) * x, then <y,u> -> slot_out (u != NULL) or ||y||_2 -> *nrm2 (u == NULL, waits) */
LIS_INT lisd_mgs_step(int slot_in, double scale, LIS_VECTOR x, LIS_VECTOR y, LIS_VECTOR u, int slot_out, LIS_REAL *nrm2)
{
    LISD_PREP2(x, y, "gram-schmidt step");
    if (u) { LIS_INT e_ = lis_vector_check_same(x, u); if (e_) return e_; e_ = lisd_vec_device(u); if (e_) return e_; }
    if ((((size_t)x->value | (size_t)y->value | (size_t)(u ? u->value : NULL)) & 15) != 0) {
        /* storage handed out by this library is 256-byte aligned; anything else takes the two launches,
         * whose packed / scalar paths the fused kernel could not both mirror */
        LIS_INT e_ = lisd_axpy_from_slot(slot_in, scale, x, y);
        if (e_) return e_;
        if (u) return lisd_dot_to_slot(y, u, slot_out);
        return lis_vector_nrm2(y, nrm2);
    }
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    LIS_INT err = lisd_check(lisb200_mgs_step(u == NULL, x->n, lisd_dev_scalar(slot_in), scale, x->value, y->value, u ? u->value : NULL,
                                              partial, lisd_counter(), u ? lisd_dev_scalar(slot_out) : lisd_scalar_dev(0), lisd_stream()),
                             "gram-schmidt step");
    if (err || u) return err;
    double rr;
    err = lisd_reduce_finish(&rr, 1, 0);
    if (err) return err;
    *nrm2 = sqrt(rr);
    return LIS_SUCCESS;
}

static int all_aligned16(const void *a, const void *b, const void *c, const void *d, const void *e)
{
    return ((((size_t)a) | ((size_t)b) | ((size_t)c) | ((size_t)d) | ((size_t)e)) & 15) == 0;
}

/* y += alpha*x ; *nrm2 = ||y||_2  (one pass; waits) */
LIS_INT lisd_axpy_nrm2(LIS_SCALAR alpha, LIS_VECTOR x, LIS_VECTOR y, LIS_REAL *nrm2)
{
    LISD_PREP2(x, y, "axpy+nrm2");
    if (!all_aligned16(x->value, y->value, NULL, NULL, NULL)) {
        LIS_INT e_ = lisd_axpy(alpha, x, y);
        return e_ ? e_ : lis_vector_nrm2(y, nrm2);
    }
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    LIS_INT err = lisd_check(lisb200_mgs_step(1, x->n, NULL, alpha, x->value, y->value, NULL, partial, lisd_counter(),
                                              lisd_scalar_dev(0), lisd_stream()), "axpy+nrm2");
    if (err) return err;
    double rr;
    err = lisd_reduce_finish(&rr, 1, 0);
    if (err) return err;
    *nrm2 = sqrt(rr);
    return LIS_SUCCESS;
}

/* y += alpha*x ; *dot = <y,u> over all ranks  (one pass; waits) -- a Gram-Schmidt link of a row-partitioned GMRES,
 * where the coefficient is a host scalar that was combined across the ranks */
LIS_INT lisd_axpy_dot(LIS_SCALAR alpha, LIS_VECTOR x, LIS_VECTOR y, LIS_VECTOR u, LIS_SCALAR *dot)
{
    LISD_PREP2(x, y, "axpy+dot");
    { LIS_INT e_ = lis_vector_check_same(x, u); if (e_) return e_; e_ = lisd_vec_device(u); if (e_) return e_; }
    if (!all_aligned16(x->value, y->value, u->value, NULL, NULL)) {
        LIS_INT e_ = lisd_axpy(alpha, x, y);
        return e_ ? e_ : lis_vector_dot(y, u, dot);
    }
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    LIS_INT err = lisd_check(lisb200_mgs_step(0, x->n, NULL, alpha, x->value, y->value, u->value, partial, lisd_counter(),
                                              lisd_scalar_dev(0), lisd_stream()), "axpy+dot");
    if (err) return err;
    return lisd_reduce_finish(dot, 1, 0);
}

/* p = r + beta*(p - omega*v) */
LIS_INT lisd_bicgstab_p(LIS_SCALAR omega, LIS_SCALAR beta, LIS_VECTOR v, LIS_VECTOR r, LIS_VECTOR p)
{
    LISD_PREP2(v, p, "bicgstab p update");
    { LIS_INT e_ = lis_vector_check_same(v, r); if (e_) return e_; e_ = lisd_vec_device(r); if (e_) return e_; }
    if (!all_aligned16(v->value, r->value, p->value, NULL, NULL)) {
        LIS_INT e_ = lisd_axpy(-omega, v, p);
        return e_ ? e_ : lisd_xpay(r, beta, p);
    }
    LISD_LAUNCH(lisb200_bicgstab_p(v->n, omega, beta, v->value, r->value, p->value, lisd_stream()), "bicgstab p update");
}

/* x += alpha*phat ; x += omega*shat ; r -= omega*t ; *nrm2 = ||r||_2  (one pass; waits) */
LIS_INT lisd_bicgstab_update(LIS_SCALAR alpha, LIS_SCALAR omega, LIS_VECTOR phat, LIS_VECTOR shat, LIS_VECTOR t,
                             LIS_VECTOR x, LIS_VECTOR r, LIS_REAL *nrm2)
{
    LISD_PREP2(phat, shat, "bicgstab update");
    { LIS_INT e_ = lis_vector_check_same(phat, t); if (e_) return e_; e_ = lis_vector_check_same(phat, x); if (e_) return e_;
      e_ = lis_vector_check_same(phat, r); if (e_) return e_;
      e_ = lisd_vec_device(t); if (e_) return e_; e_ = lisd_vec_device(x); if (e_) return e_; e_ = lisd_vec_device(r); if (e_) return e_; }
    if (!all_aligned16(phat->value, shat->value, t->value, x->value, r->value)) {
        LIS_INT e_ = lisd_axpy(alpha, phat, x);
        if (!e_) e_ = lisd_axpy(omega, shat, x);
        if (!e_) e_ = lisd_axpy(-omega, t, r);
        return e_ ? e_ : lis_vector_nrm2(r, nrm2);
    }
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    LIS_INT err = lisd_check(lisb200_bicgstab_update(phat->n, alpha, omega, phat->value, shat->value, t->value, x->value, r->value,
                                                     partial, lisd_counter(), lisd_scalar_dev(0), lisd_stream()), "bicgstab update");
    if (err) return err;
    double rr;
    err = lisd_reduce_finish(&rr, 1, 0);
    if (err) return err;
    *nrm2 = sqrt(rr);
    return LIS_SUCCESS;
}

/* reductions: kernel -> mapped host scalar; ranks combined in rank order on the host */
LIS_INT lisd_reduce(int kind, LIS_VECTOR x, LIS_VECTOR y, LIS_SCALAR *value)
{
    LIS_INT err = lisd_require("vector reduction");
    if (err) return err;
    err = lisd_vec_device(x);
    if (err) return err;
    if (y) { err = lis_vector_check_same(x, y); if (err) return err; err = lisd_vec_device(y); if (err) return err; }
    double *partial = lisd_partial(0);
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    err = lisd_check(lisb200_reduce(kind, x->n, x->value, y ? y->value : x->value, partial,
                                    lisd_counter(), lisd_scalar_dev(0), lisd_stream()), "vector reduction");
    if (err) return err;
    return lisd_reduce_finish(value, 1, kind == 3);
}

/* ------------------------------------------------------------------ public, host-synchronous */
#define LIS_SYNC_RETURN(expr) do { LIS_INT e_ = (expr); if (e_) return e_; return lisd_sync(); } while (0)

LIS_INT lis_vector_copy(LIS_VECTOR vsrc, LIS_VECTOR vdst) { LIS_SYNC_RETURN(lisd_copy(vsrc, vdst)); }
LIS_INT lis_vector_axpy(LIS_SCALAR alpha, LIS_VECTOR vx, LIS_VECTOR vy) { LIS_SYNC_RETURN(lisd_axpy(alpha, vx, vy)); }
LIS_INT lis_vector_xpay(LIS_VECTOR vx, LIS_SCALAR alpha, LIS_VECTOR vy) { LIS_SYNC_RETURN(lisd_xpay(vx, alpha, vy)); }
LIS_INT lis_vector_axpyz(LIS_SCALAR alpha, LIS_VECTOR vx, LIS_VECTOR vy, LIS_VECTOR vz) { LIS_SYNC_RETURN(lisd_axpyz(alpha, vx, vy, vz)); }
LIS_INT lis_vector_scale(LIS_SCALAR alpha, LIS_VECTOR vx) { LIS_SYNC_RETURN(lisd_scale(alpha, vx)); }
LIS_INT lis_vector_set_all(LIS_SCALAR alpha, LIS_VECTOR vx) { LIS_SYNC_RETURN(lisd_set_all(alpha, vx)); }
LIS_INT lis_vector_pmul(LIS_VECTOR vx, LIS_VECTOR vy, LIS_VECTOR vz) { LIS_SYNC_RETURN(lisd_pmul(vx, vy, vz)); }

LIS_INT lis_vector_pdiv(LIS_VECTOR vx, LIS_VECTOR vy, LIS_VECTOR vz)
{
    LISD_PREP2(vx, vy, "lis_vector_pdiv");
    { LIS_INT e_ = lis_vector_check_same(vx, vz); if (e_) return e_; e_ = lisd_vec_device(vz); if (e_) return e_; }
    lisd_mark_busy();
    LIS_SYNC_RETURN(lisd_check(lisb200_pdiv(vx->n, vx->value, vy->value, vz->value, lisd_stream()), "lis_vector_pdiv"));
}
LIS_INT lis_vector_swap(LIS_VECTOR vsrc, LIS_VECTOR vdst)
{
    LISD_PREP2(vsrc, vdst, "lis_vector_swap");
    lisd_mark_busy();
    LIS_SYNC_RETURN(lisd_check(lisb200_swap(vsrc->n, vsrc->value, vdst->value, lisd_stream()), "lis_vector_swap"));
}
LIS_INT lis_vector_abs(LIS_VECTOR vx)
{
    LISD_PREP1(vx, "lis_vector_abs");
    lisd_mark_busy();
    LIS_SYNC_RETURN(lisd_check(lisb200_abs(vx->n, vx->value, lisd_stream()), "lis_vector_abs"));
}
LIS_INT lis_vector_reciprocal(LIS_VECTOR vx)
{
    LISD_PREP1(vx, "lis_vector_reciprocal");
    lisd_mark_busy();
    LIS_SYNC_RETURN(lisd_check(lisb200_reciprocal(vx->n, vx->value, lisd_stream()), "lis_vector_reciprocal"));
}
LIS_INT lis_vector_conjugate(LIS_VECTOR vx) { (void)vx; return LIS_SUCCESS; }   /* real scalars */
LIS_INT lis_vector_shift(LIS_SCALAR sigma, LIS_VECTOR vx)
{
    LISD_PREP1(vx, "lis_vector_shift");
    lisd_mark_busy();
    LIS_SYNC_RETURN(lisd_check(lisb200_shift(vx->n, sigma, vx->value, lisd_stream()), "lis_vector_shift"));
}

LIS_INT lis_vector_dot(LIS_VECTOR vx, LIS_VECTOR vy, LIS_SCALAR *value) { return lisd_reduce(0, vx, vy, value); }
LIS_INT lis_vector_nhdot(LIS_VECTOR vx, LIS_VECTOR vy, LIS_SCALAR *value) { return lisd_reduce(0, vx, vy, value); }
LIS_INT lis_vector_nrm2(LIS_VECTOR vx, LIS_REAL *value)
{
    LIS_SCALAR s;
    LIS_INT err = lisd_reduce(1, vx, NULL, &s);
    if (err) return err;
    *value = sqrt(s);
    return LIS_SUCCESS;
}
LIS_INT lis_vector_nrm1(LIS_VECTOR vx, LIS_REAL *value) { return lisd_reduce(2, vx, NULL, value); }
LIS_INT lis_vector_nrmi(LIS_VECTOR vx, LIS_REAL *value) { return lisd_reduce(3, vx, NULL, value); }
LIS_INT lis_vector_sum(LIS_VECTOR vx, LIS_SCALAR *value) { return lisd_reduce(4, vx, NULL, value); }
