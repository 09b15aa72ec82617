// FP64 and HBM micro-benchmarks: the roofline denominators of the statistics kernels.
// MEASURED_PEAKS.json (driver-written) has HBM copy bandwidth and bf16 GEMM only; the
// FP64 pipes are measured here: independent DFMA chains and mma.sync.m8n8k4.f64 chains.
#include "../../include/btf_b200.h"
#include <cuda_runtime.h>
#include <stdio.h>

namespace {

template <int CH>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
    double c0[CH], c1[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { c0[c] = threadIdx.x * 1e-3; c1[c] = c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += c0[c] + c1[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double4* __restrict__ src, double4* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace

extern "C" double btf_fp64_peak(int32_t device, int32_t mode, int32_t iters) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 4;
    double* out = nullptr;
    cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (iters < 1) iters = 20000;
    constexpr int CH = 16;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) dfma_kernel<CH><<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
        else dmma_kernel<CH><<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops;
        if (mode == 0) flops = 2.0 * CH * (double)iters * blocks * 256;
        else flops = 2.0 * 256.0 * CH * (double)iters * blocks * 8;   // 8 warps/block, 8x8x4 FMAs per DMMA
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return err == cudaSuccess ? best : -1.0;
}

extern "C" double btf_hbm_copy_gbs(int32_t device, size_t bytes, int32_t iters) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    double4 *a = nullptr, *b = nullptr;
    bytes = bytes / 32 * 32;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) { cudaFree(a); return -1.0; }
    cudaMemset(a, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int it = 0; it < iters + 1; ++it) {
        cudaEventRecord(e0);
        copy_kernel<<<148 * 16, 512>>>(a, b, bytes / 32);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double gbs = 2.0 * bytes / (ms * 1e-3) / 1e9;
        if (it > 0 && gbs > best) best = gbs;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(a); cudaFree(b);
    return best;
}
