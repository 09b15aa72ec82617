// Shared pieces of the statistics kernels (stats_kernels.cu, stats_zpre.cu).
#pragma once
#include "kernels.h"

namespace btf {

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct StatsKArgs {
    const void* wt;
    const double* sv;
    const double* F;
    double* out;
    long long ld;
    int K, L, nct_z, nct_f, zw;
    int nchunks, chunks_per_split;
    int m_valid;
    long long out_split_stride;
};

template <bool TRANS, typename WT, int BM, int KC>
struct TileGeom {
    // Shared-memory row strides (elements).  A 64-bit LDS is served per half-warp, so
    // the four k-rows (or four m-rows) a half-warp touches must fall into disjoint
    // 8-word bank groups: double strides are = 4 (mod 16); byte tiles use strides whose
    // word offsets are distinct.  Every cp.async destination stays 16-byte aligned.
    static constexpr int WROWS = TRANS ? KC : BM;
    static constexpr int WSTR = sizeof(WT) == 1 ? (TRANS ? BM + 16 : 48) : (TRANS ? BM + 4 : KC + 4);
    static constexpr int SSTR = TRANS ? BM + 4 : KC + 4;
    static constexpr int WBYTES = WROWS * WSTR * (int)sizeof(WT);
    static constexpr int SBYTES = WROWS * SSTR * 8;
};

__host__ __device__ constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }


}  // namespace btf
