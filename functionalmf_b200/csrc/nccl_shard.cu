#include "nccl_shard.h"
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace btf {

// minimal NCCL ABI (nccl.h): opaque comm, 128-byte unique id, enums
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint8 = 1, ncclInt32 = 2, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_api;
static char g_nccl_err[256] = "";
const char* nccl_shard_error() { return g_nccl_err; }

static bool load_api() {
    if (g_api.h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.h) break;
    }
    if (!g_api.h) { snprintf(g_nccl_err, sizeof(g_nccl_err), "cannot dlopen libnccl.so.2: %s", dlerror()); return false; }
#define SYM(field, name) *(void**)(&g_api.field) = dlsym(g_api.h, name); if (!g_api.field) { snprintf(g_nccl_err, sizeof(g_nccl_err), "missing symbol %s", name); return false; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Reduce, "ncclReduce")
    SYM(ReduceScatter, "ncclReduceScatter")
    SYM(Broadcast, "ncclBroadcast")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
}

#define NC(call)                                                                               \
    do {                                                                                       \
        int _r = (call);                                                                       \
        if (_r != ncclSuccess) {                                                               \
            snprintf(g_nccl_err, sizeof(g_nccl_err), "%s: %s", #call, g_api.GetErrorString(_r)); \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)

struct NcclShard {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, N = 0, M = 0;
    std::vector<int> rb, re, cb, ce;   // per-rank [begin, end) of rows and columns
    bool uniform_rows = false, uniform_cols = false;   // equal contiguous blocks: single collectives apply
    double* scratch = nullptr;
};

int nccl_shard_unique_id(char* id128) {
    if (!load_api()) return -1;
    ncclUniqueId id;
    NC(g_api.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return 0;
}

NcclShard* nccl_shard_create(const char* id128, int world, int rank, int N, int M, int row_begin, int row_end,
                             int col_begin, int col_end) {
    if (!load_api()) return nullptr;
    NcclShard* s = new NcclShard();
    s->world = world; s->rank = rank; s->N = N; s->M = M;
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    int r = g_api.CommInitRank(&s->comm, world, id, rank);
    if (r != ncclSuccess) {
        snprintf(g_nccl_err, sizeof(g_nccl_err), "ncclCommInitRank: %s", g_api.GetErrorString(r));
        delete s; return nullptr;
    }
    // exchange the shard boundaries
    int* dev = nullptr;
    cudaMalloc((void**)&dev, sizeof(int) * 4 * (world + 1));
    int mine[4] = {row_begin, row_end, col_begin, col_end};
    cudaMemcpy(dev + 4 * world, mine, sizeof(mine), cudaMemcpyHostToDevice);
    r = g_api.AllGather(dev + 4 * world, dev, 4, ncclInt32, s->comm, 0);
    cudaError_t ce = cudaDeviceSynchronize();
    std::vector<int> all(4 * world);
    cudaMemcpy(all.data(), dev, sizeof(int) * 4 * world, cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (r != ncclSuccess || ce != cudaSuccess) {
        snprintf(g_nccl_err, sizeof(g_nccl_err), "boundary all-gather failed");
        g_api.CommDestroy(s->comm); delete s; return nullptr;
    }
    for (int k = 0; k < world; ++k) {
        s->rb.push_back(all[4 * k]); s->re.push_back(all[4 * k + 1]);
        s->cb.push_back(all[4 * k + 2]); s->ce.push_back(all[4 * k + 3]);
    }
    cudaMalloc((void**)&s->scratch, 8 * sizeof(double));
    auto uniform = [&](const std::vector<int>& b, const std::vector<int>& e) {
        const int cnt = e[0] - b[0];
        for (int k = 0; k < world; ++k)
            if (e[k] - b[k] != cnt || b[k] != k * cnt) return false;
        return cnt > 0;
    };
    s->uniform_rows = uniform(s->rb, s->re);
    s->uniform_cols = uniform(s->cb, s->ce);
    return s;
}

void nccl_shard_destroy(NcclShard* s) {
    if (!s) return;
    if (s->comm) g_api.CommDestroy(s->comm);
    if (s->scratch) cudaFree(s->scratch);
    delete s;
}

static int gather_blocks(NcclShard* s, double* base, const std::vector<int>& b, const std::vector<int>& e, size_t unit,
                         cudaStream_t st, bool uniform) {
    if (uniform) {      // equal blocks: one in-place all-gather
        const size_t cnt = (size_t)(e[0] - b[0]) * unit;
        NC(g_api.AllGather(base + (size_t)b[s->rank] * unit, base, cnt, ncclFloat64, s->comm, st));
        return 0;
    }
    NC(g_api.GroupStart());
    for (int r = 0; r < s->world; ++r) {
        size_t cnt = (size_t)(e[r] - b[r]) * unit;
        if (cnt == 0) continue;
        double* p = base + (size_t)b[r] * unit;
        NC(g_api.Broadcast(p, p, cnt, ncclFloat64, r, s->comm, st));
    }
    NC(g_api.GroupEnd());
    return 0;
}

int nccl_allgather_rows(NcclShard* s, double* W, int K, cudaStream_t st) { return gather_blocks(s, W, s->rb, s->re, K, st, s->uniform_rows); }
int nccl_allgather_cols(NcclShard* s, double* V, int n, cudaStream_t st) { return gather_blocks(s, V, s->cb, s->ce, n, st, s->uniform_cols); }
int nccl_allgather_doubles(NcclShard* s, double* v, cudaStream_t st) { return gather_blocks(s, v, s->cb, s->ce, 1, st, s->uniform_cols); }

__global__ void collapse_splits_kernel(double* x, int nsplit, size_t stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < stride; i += (size_t)gridDim.x * blockDim.x) {
        double v = x[i];
        for (int s = 1; s < nsplit; ++s) v += x[(size_t)s * stride + i];
        x[i] = v;
    }
}

void launch_collapse_splits(double* x, int nsplit, size_t stride, cudaStream_t st) {
    if (nsplit > 1) collapse_splits_kernel<<<148 * 8, 256, 0, st>>>(x, nsplit, stride);
}

int nccl_reduce_col_stats(NcclShard* s, double* col_stats, int nsplit, size_t split_stride, int per_col_elems,
                          cudaStream_t st) {
    if (nsplit > 1) collapse_splits_kernel<<<148 * 8, 256, 0, st>>>(col_stats, nsplit, split_stride);
    if (s->uniform_cols) {   // equal column blocks: one in-place reduce-scatter
        const size_t cnt = (size_t)(s->ce[0] - s->cb[0]) * per_col_elems;
        NC(g_api.ReduceScatter(col_stats, col_stats + (size_t)s->rank * cnt, cnt, ncclFloat64, ncclSum, s->comm, st));
        return 0;
    }
    NC(g_api.GroupStart());
    for (int r = 0; r < s->world; ++r) {
        size_t cnt = (size_t)(s->ce[r] - s->cb[r]) * per_col_elems;
        if (cnt == 0) continue;
        double* p = col_stats + (size_t)s->cb[r] * per_col_elems;
        NC(g_api.Reduce(p, p, cnt, ncclFloat64, ncclSum, r, s->comm, st));
    }
    NC(g_api.GroupEnd());
    return 0;
}

// in-place sum of a per-(j,t) array [M * T][unit] across ranks; afterwards every rank holds the totals of ITS column block
int nccl_reduce_scatter_cols(NcclShard* s, double* buf, size_t per_col_elems, cudaStream_t st) {
    if (s->uniform_cols) {
        const size_t cnt = (size_t)(s->ce[0] - s->cb[0]) * per_col_elems;
        NC(g_api.ReduceScatter(buf, buf + (size_t)s->rank * cnt, cnt, ncclFloat64, ncclSum, s->comm, st));
        return 0;
    }
    NC(g_api.GroupStart());
    for (int r = 0; r < s->world; ++r) {
        size_t cnt = (size_t)(s->ce[r] - s->cb[r]) * per_col_elems;
        if (cnt == 0) continue;
        double* p = buf + (size_t)s->cb[r] * per_col_elems;
        NC(g_api.Reduce(p, p, cnt, ncclFloat64, ncclSum, r, s->comm, st));
    }
    NC(g_api.GroupEnd());
    return 0;
}

// the four Tau2 arrays [M][RD], column blocks gathered in place
int nccl_allgather_tau(NcclShard* s, double* const* arrays, int narr, int RD, cudaStream_t st) {
    for (int a = 0; a < narr; ++a)
        if (gather_blocks(s, arrays[a], s->cb, s->ce, (size_t)RD, st, s->uniform_cols)) return -1;
    return 0;
}

int nccl_shard_max_rows(const NcclShard* s) {
    int m = 0;
    for (int r = 0; r < s->world; ++r) m = std::max(m, s->re[r] - s->rb[r]);
    return m;
}

// Column-sharded copy of the counts: dst[p - p0][i] = cnt[i][p] for this rank's columns p in [p0, p1) = [cb T, ce T) and
// ALL rows i.  Every rank holds srcT = transpose of its own row block, [P][src_ld]; the (rank, peer) blocks move in
// world rounds of one send + one receive (contiguous row ranges of srcT), then a strided copy puts the block in place.
// tmp: >= (p1 - p0) * round_up(max rows per rank, 128) bytes.
int nccl_exchange_counts(NcclShard* s, const uint8_t* srcT, long long src_ld, int T, uint8_t* dst, long long dst_ld,
                         uint8_t* tmp, cudaStream_t st) {
    const int me = s->rank, W = s->world;
    const size_t ploc = (size_t)(s->ce[me] - s->cb[me]) * T;
    auto pad128 = [](int x) { return (long long)((x + 127) / 128 * 128); };
    for (int r = 0; r < W; ++r) {
        const int to = (me + r) % W, from = (me - r + W) % W;
        const long long from_ld = pad128(s->re[from] - s->rb[from]);
        const size_t send_bytes = (size_t)(s->ce[to] - s->cb[to]) * T * (size_t)src_ld;
        const size_t recv_bytes = ploc * (size_t)from_ld;
        const uint8_t* block = tmp;
        if (r == 0) {
            block = srcT + (size_t)s->cb[me] * T * (size_t)src_ld;
        } else {
            NC(g_api.GroupStart());
            if (send_bytes) NC(g_api.Send(srcT + (size_t)s->cb[to] * T * (size_t)src_ld, send_bytes, ncclUint8, to, s->comm, st));
            if (recv_bytes) NC(g_api.Recv(tmp, recv_bytes, ncclUint8, from, s->comm, st));
            NC(g_api.GroupEnd());
        }
        const size_t width = (size_t)(s->re[from] - s->rb[from]);
        if (ploc && width) {
            if (cudaMemcpy2DAsync(dst + s->rb[from], (size_t)dst_ld, block, (size_t)from_ld, width, ploc,
                                  cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
                snprintf(g_nccl_err, sizeof(g_nccl_err), "strided copy of a count block failed");
                return -1;
            }
        }
    }
    return 0;
}

// Column-sharded copy of an FP64 row-sharded matrix: dst[i][p - p0] = src_rank(i)[i - rb][p] for ALL rows i and this rank's
// columns p in [p0, p1).  src: this rank's rows [nloc][src_ld]; dst: [N_pad][dst_ld]; tmp: >= max_rows * max_ploc doubles for
// the packed send block and as many for the receive block.  Same ring of `world` rounds as the count exchange.
int nccl_exchange_rows_f64(NcclShard* s, const double* src, long long src_ld, int T, double* dst, long long dst_ld,
                           double* tmp_send, double* tmp_recv, cudaStream_t st) {
    const int me = s->rank, W = s->world;
    const size_t ploc = (size_t)(s->ce[me] - s->cb[me]) * T;
    const size_t nloc = (size_t)(s->re[me] - s->rb[me]);
    for (int r = 0; r < W; ++r) {
        const int to = (me + r) % W, from = (me - r + W) % W;
        const size_t to_w = (size_t)(s->ce[to] - s->cb[to]) * T;          // columns the receiver owns
        const size_t from_rows = (size_t)(s->re[from] - s->rb[from]);
        if (r == 0) {
            if (ploc && nloc &&
                cudaMemcpy2DAsync(dst + (size_t)s->rb[me] * dst_ld, (size_t)dst_ld * 8, src + (size_t)s->cb[me] * T, (size_t)src_ld * 8,
                                  ploc * 8, nloc, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
                snprintf(g_nccl_err, sizeof(g_nccl_err), "strided copy of a data block failed");
                return -1;
            }
            continue;
        }
        // pack my rows x the receiver's columns, exchange, unpack the sender's rows x my columns
        if (to_w && nloc &&
            cudaMemcpy2DAsync(tmp_send, to_w * 8, src + (size_t)s->cb[to] * T, (size_t)src_ld * 8, to_w * 8, nloc,
                              cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            snprintf(g_nccl_err, sizeof(g_nccl_err), "packing a data block failed");
            return -1;
        }
        NC(g_api.GroupStart());
        if (to_w * nloc) NC(g_api.Send(tmp_send, to_w * nloc, ncclFloat64, to, s->comm, st));
        if (ploc * from_rows) NC(g_api.Recv(tmp_recv, ploc * from_rows, ncclFloat64, from, s->comm, st));
        NC(g_api.GroupEnd());
        if (ploc && from_rows &&
            cudaMemcpy2DAsync(dst + (size_t)s->rb[from] * dst_ld, (size_t)dst_ld * 8, tmp_recv, ploc * 8, ploc * 8, from_rows,
                              cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            snprintf(g_nccl_err, sizeof(g_nccl_err), "unpacking a data block failed");
            return -1;
        }
    }
    return 0;
}

int nccl_shard_max_cols(const NcclShard* s) {
    int m = 0;
    for (int r = 0; r < s->world; ++r) m = std::max(m, s->ce[r] - s->cb[r]);
    return m;
}

int nccl_allreduce_sum(NcclShard* s, double* p, int n, cudaStream_t st) {
    NC(g_api.AllReduce(p, p, (size_t)n, ncclFloat64, ncclSum, s->comm, st));
    return 0;
}

__global__ void resid_local_kernel(const Scalars* sc, double* tmp) { tmp[0] = sc->resid - sc->ss_total; }
__global__ void resid_global_kernel(Scalars* sc, const double* tmp) { sc->resid = sc->ss_total + tmp[0]; }

int nccl_allreduce_resid(NcclShard* s, Scalars* scal, cudaStream_t st) {
    resid_local_kernel<<<1, 1, 0, st>>>(scal, s->scratch);
    NC(g_api.AllReduce(s->scratch, s->scratch, 1, ncclFloat64, ncclSum, s->comm, st));
    resid_global_kernel<<<1, 1, 0, st>>>(scal, s->scratch);
    return 0;
}

}  // namespace btf
