// Exchange steps of the row/column-sharded sweep over NCCL.
// NCCL is loaded lazily with dlopen so the single-GPU path has no NCCL dependency.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace btf {

struct NcclShard;

int nccl_shard_unique_id(char* id128);
const char* nccl_shard_error();
NcclShard* nccl_shard_create(const char* id128, int world, int rank, int N, int M, int row_begin, int row_end,
                             int col_begin, int col_end);
void nccl_shard_destroy(NcclShard* s);

// in-place all-gather of the row blocks of W [N][K]
int nccl_allgather_rows(NcclShard* s, double* W, int K, cudaStream_t st);
// in-place all-gather of the column blocks of V [M][n]
int nccl_allgather_cols(NcclShard* s, double* V, int n, cudaStream_t st);
// in-place all-gather of a per-column vector [M]
int nccl_allgather_doubles(NcclShard* s, double* v, cudaStream_t st);
// x[i] += x[s * stride + i], s = 1..nsplit-1, in split order (also used on one GPU when split-K is deep)
void launch_collapse_splits(double* x, int nsplit, size_t stride, cudaStream_t st);
// collapse the split partials into split 0, then sum across ranks so that every rank
// holds the totals of ITS column block (a reduce-scatter with uneven blocks)
int nccl_reduce_col_stats(NcclShard* s, double* col_stats, int nsplit, size_t split_stride, int per_col_elems,
                          cudaStream_t st);
int nccl_allreduce_sum(NcclShard* s, double* p, int n, cudaStream_t st);
// in-place sum across ranks of a per-(j,t) array [M * T][per_col_elems / T ...]: rank r ends up with the totals of its column block
int nccl_reduce_scatter_cols(NcclShard* s, double* buf, size_t per_col_elems, cudaStream_t st);
// in-place all-gather of the column blocks of `narr` arrays [M][RD] (the Tau2 chain)
int nccl_allgather_tau(NcclShard* s, double* const* arrays, int narr, int RD, cudaStream_t st);
int nccl_shard_max_rows(const NcclShard* s);
int nccl_shard_max_cols(const NcclShard* s);
// column-sharded copy of a row-sharded FP64 matrix (see nccl_shard.cu)
int nccl_exchange_rows_f64(NcclShard* s, const double* src, long long src_ld, int T, double* dst, long long dst_ld,
                           double* tmp_send, double* tmp_recv, cudaStream_t st);
// column-sharded copy of the uint8 counts (see nccl_shard.cu)
int nccl_exchange_counts(NcclShard* s, const uint8_t* srcT, long long src_ld, int T, uint8_t* dst, long long dst_ld,
                         uint8_t* tmp, cudaStream_t st);
// scal->resid currently holds ss_total + (local partial); make it ss_total + sum of all partials
int nccl_allreduce_resid(NcclShard* s, Scalars* scal, cudaStream_t st);

}  // namespace btf
