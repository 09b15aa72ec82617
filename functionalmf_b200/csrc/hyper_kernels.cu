// K5: hyper-parameter conditionals with on-device Philox RNG
//   Tau2 / c / b / a horseshoe+ chain   (factor.py:134-141)
//   lam2, lam2_a                        (factor.py:143-153)
//   sigma2                              (factor.py:130-132, genlasso.py:149-168)
//   nu2                                 (factor.py:411-416, genlasso.py:149-168)
#include "kernels.h"

namespace btf {

// One block per column j: V[j] (T x K) is staged in shared memory (coalesced), then one thread
// per penalty row r evaluates delta = Delta[r,:] . V[j,:,k] through the (<= p+2)-point stencil
// and runs the four-level inverse-gamma chain.
__global__ void __launch_bounds__(128) tau2_kernel(HyperArgs a) {
    extern __shared__ double vs[];            // [T][K]
    const int j = a.col_begin + blockIdx.x;
    const int K = a.K, T = a.T;
    const double* Vj = a.V + (size_t)j * T * K;
    for (int e = threadIdx.x; e < T * K; e += blockDim.x) vs[e] = Vj[e];
    __syncthreads();
    const double lo = a.stability, hi = 1.0 / a.stability;
    const double lam2 = a.scal->lam2;
    for (int r = threadIdx.x; r < a.RD; r += blockDim.x) {
        const int s0 = a.d_start[r], w = a.d_width[r];
        const double* coef = a.d_coef + (size_t)r * a.d_maxw;
        double ssq = 0.0;
        for (int k = 0; k < K; ++k) {
            double d = 0.0;
            for (int x = 0; x < w; ++x) d += coef[x] * vs[(s0 + x) * K + k];
            ssq += d * d;
        }
        const size_t o = (size_t)j * a.RD + r;
        double g0, g1, g2, g3;
        if (a.g_inject) {
            const double* g = a.g_inject + (size_t)j * 4 * a.RD + r;
            g0 = g[0]; g1 = g[a.RD]; g2 = g[2 * a.RD]; g3 = g[3 * a.RD];
        } else {
            Rng rng(a.seed, STREAM_TAU, a.scal->sweep, (uint64_t)o);
            g0 = rng.gamma(0.5 * (K + 1));
            g1 = rng.exponential(); g2 = rng.exponential(); g3 = rng.exponential();
        }
        double rate = ssq / (2.0 * lam2) + 1.0 / clampd(a.Tau2_c[o], lo, hi);
        double tau2 = 1.0 / (g0 * (1.0 / clampd(rate, lo, hi)));
        double c = 1.0 / (g1 * (1.0 / clampd(1.0 / tau2 + 1.0 / a.Tau2_b[o], lo, hi)));
        double b = 1.0 / (g2 * (1.0 / clampd(1.0 / c + 1.0 / a.Tau2_a[o], lo, hi)));
        double aa = 1.0 / (g3 * (1.0 / clampd(1.0 / b + 1.0, lo, hi)));
        a.Tau2[o] = tau2; a.Tau2_c[o] = c; a.Tau2_b[o] = b; a.Tau2_a[o] = aa;
    }
}

// lam_partials[j] = 0.5 * sum_r ssq(j,r) / tau2[j,r]   (recomputed from the current Tau2;
// also used when Tau2 is held fixed)
__global__ void lam_partial_kernel(HyperArgs a) {
    __shared__ double sh[32];
    const int j = a.col_begin + blockIdx.x;
    const int K = a.K;
    double acc = 0.0;
    for (int r = threadIdx.x; r < a.RD; r += blockDim.x) {
        const int s0 = a.d_start[r], w = a.d_width[r];
        const double* coef = a.d_coef + (size_t)r * a.d_maxw;
        const double* Vj = a.V + ((size_t)j * a.T + s0) * K;
        double ssq = 0.0;
        for (int k = 0; k < K; ++k) {
            double d = 0.0;
            for (int x = 0; x < w; ++x) d += coef[x] * Vj[(size_t)x * K + k];
            ssq += d * d;
        }
        acc += ssq / a.Tau2[(size_t)j * a.RD + r];
    }
    double tot = block_sum(acc, sh);
    if (threadIdx.x == 0) a.lam_partials[j] = 0.5 * tot;
}

void launch_tau2(const HyperArgs& a, cudaStream_t st) {
    const int ncol = a.col_end - a.col_begin;
    if (ncol <= 0) return;
    if (a.Tau2_a) {   // Tau2_a == nullptr: only the lam2 partials are wanted (Tau2 held fixed)
        size_t smem = (size_t)a.T * a.K * sizeof(double);
        static PerDeviceMax max_set;
        if (smem > 48 * 1024 && max_set.raise(smem)) cudaFuncSetAttribute(tau2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tau2_kernel<<<ncol, 128, smem, st>>>(a);
    }
    if (a.lam_partials) lam_partial_kernel<<<ncol, 128, 0, st>>>(a);
}

__global__ void nu2_kernel(ScalarStepArgs a) {
    Scalars* s = a.scal;
    const double a_post = a.prior_a + 0.5 * s->n_obs;
    const double b_post = a.prior_b + 0.5 * s->resid;
    double g;
    if (a.g_inject) g = a.g_inject[0];
    else { Rng rng(a.seed, STREAM_NU, s->sweep, 0); g = rng.gamma(a_post); }
    s->nu2 = 1.0 / (g * (1.0 / b_post));
    s->nu2_a_post = a_post;
    s->nu2_b_post = b_post;
}
void launch_nu2(const ScalarStepArgs& a, cudaStream_t st) { nu2_kernel<<<1, 1, 0, st>>>(a); }

__global__ void sigma2_kernel(ScalarStepArgs a, double n_free) {
    Scalars* s = a.scal;
    const double a_post = a.prior_a + 0.5 * n_free;
    const double b_post = a.prior_b + 0.5 * s->w_sumsq;
    double g;
    if (a.g_inject) g = a.g_inject[0];
    else { Rng rng(a.seed, STREAM_SIGMA, s->sweep, 0); g = rng.gamma(a_post); }
    s->sigma2 = 1.0 / (g * (1.0 / b_post));
}
void launch_sigma2(const ScalarStepArgs& a, double n_free, cudaStream_t st) {
    sigma2_kernel<<<1, 1, 0, st>>>(a, n_free);
}

__global__ void lam2_kernel(ScalarStepArgs a, const double* lam_partials, int M, int ref_compat, double shape) {
    __shared__ double sh[32];
    Scalars* s = a.scal;
    double rate;
    if (ref_compat) {
        rate = lam_partials[M - 1];          // factor.py:150 overwrites the rate per column
    } else {
        double v = 0.0;
        for (int j = threadIdx.x; j < M; j += blockDim.x) v += lam_partials[j];
        v = block_sum(v, sh);
        rate = v + 1.0 / s->lam2_a;
    }
    if (threadIdx.x == 0) {
        double g0, g1;
        if (a.g_inject) { g0 = a.g_inject[0]; g1 = a.g_inject[1]; }
        else { Rng rng(a.seed, STREAM_LAM, s->sweep, 0); g0 = rng.gamma(shape); g1 = rng.exponential(); }
        double lam2 = fmax(1e-5, 1.0 / (g0 * (1.0 / rate)));
        s->lam2 = lam2;
        s->lam2_a = 1.0 / (g1 * (1.0 / (1.0 / lam2 + 1.0)));
        s->lam2_rate = rate;
        s->lam2_shape = shape;
    }
}
void launch_lam2(const ScalarStepArgs& a, const double* lam_partials, int M, int ref_compat, double shape,
                 cudaStream_t st) {
    lam2_kernel<<<1, 256, 0, st>>>(a, lam_partials, M, ref_compat, shape);
}

// sum of squares of the free (lower-triangular) entries of W (factor.py:155-174): per-block
// partials, combined in a fixed order by w_sumsq_final_kernel (deterministic)
__global__ void __launch_bounds__(256) w_sumsq_kernel(const double* __restrict__ W, int N, int K, double* partials) {
    __shared__ double sh[32];
    double v = 0.0;
    const long long total = (long long)N * K;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int i = (int)(e / K), k = (int)(e - (long long)i * K);
        if (k <= i) { double w = W[e]; v += w * w; }
    }
    v = block_sum(v, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
}
__global__ void w_sumsq_final_kernel(const double* partials, int n, Scalars* scal) {
    __shared__ double sh[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
    v = block_sum(v, sh);
    if (threadIdx.x == 0) scal->w_sumsq = v;
}
void launch_w_sumsq(const double* W, int N, int K, Scalars* scal, double* partials, cudaStream_t st) {
    long long total = (long long)N * K;
    int nb = (int)((total + 2047) / 2048);
    if (nb > 296) nb = 296;
    if (nb < 1) nb = 1;
    w_sumsq_kernel<<<nb, 256, 0, st>>>(W, N, K, partials);
    w_sumsq_final_kernel<<<1, 256, 0, st>>>(partials, nb, scal);
}

__global__ void bump_sweep_kernel(Scalars* s) { s->sweep += 1ull; }
void launch_bump_sweep(Scalars* scal, cudaStream_t st) { bump_sweep_kernel<<<1, 1, 0, st>>>(scal); }
__global__ void clear_info_kernel(Scalars* s) { s->info_w = 0; s->info_v = 0; s->retries_v = 0; }
void launch_clear_info(Scalars* scal, cudaStream_t st) { clear_info_kernel<<<1, 1, 0, st>>>(scal); }

__global__ void set_resid_kernel(Scalars* s, const double* partials, int n) {
    __shared__ double sh[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
    v = block_sum(v, sh);
    if (threadIdx.x == 0) s->resid = s->ss_total + v;
}
void launch_set_resid(Scalars* scal, const double* partials, int n, cudaStream_t st) {
    set_resid_kernel<<<1, 1024, 0, st>>>(scal, partials, n);
}


// ---------------------------------------------------------------- prior initialisation
// Mirrors the constructor draws of the reference (factor.py:230-253, 293-304, 560-563;
// utils.py:115-124) with Philox instead of numpy's global MT19937 stream.
__global__ void init_scalars_kernel(Scalars* s, uint64_t seed, int mask, double sigma2_a, double sigma2_b,
                                    double nu2_a, double nu2_b) {
    Rng rng(seed, 100u, 0ull, 0);
    if (mask & 1) s->sigma2 = 1.0 / (rng.gamma(sigma2_a) * (1.0 / sigma2_b));        // factor.py:252-253
    if (mask & 2) {                                                                  // utils.py:122-124, factor.py:248-250
        double a = 1.0 / rng.gamma(0.5);
        double lam2 = 1.0 / (rng.gamma(0.5) * a);
        s->lam2 = fmin(fmax(lam2, 0.0), 4.0);
        s->lam2_a = a;
    }
    if (mask & 4) s->nu2 = 1.0 / (rng.gamma(nu2_a) * (1.0 / nu2_b));                 // factor.py:418-419
}
void launch_init_scalars(Scalars* s, uint64_t seed, int mask, double sigma2_a, double sigma2_b, double nu2_a,
                         double nu2_b, cudaStream_t st) {
    init_scalars_kernel<<<1, 1, 0, st>>>(s, seed, mask, sigma2_a, sigma2_b, nu2_a, nu2_b);
}

__global__ void init_tau2_kernel(double* Tau2, double* Ta, double* Tb, double* Tc, size_t n, uint64_t seed) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    Rng rng(seed, 101u, 0ull, e);
    double a = 1.0 / rng.gamma(0.5);              // utils.py:115-120
    double b = 1.0 / (rng.gamma(0.5) * a);
    double c = 1.0 / (rng.gamma(0.5) * b);
    double d = 1.0 / (rng.gamma(0.5) * c);
    Tau2[e] = fmin(fmax(d, 0.0), 9.0);            // factor.py:246
    Tc[e] = c; Tb[e] = b; Ta[e] = a;
}
void launch_init_tau2(double* Tau2, double* Ta, double* Tb, double* Tc, size_t n, uint64_t seed, cudaStream_t st) {
    init_tau2_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(Tau2, Ta, Tb, Tc, n, seed);
}

__global__ void init_W_kernel(double* W, int N, int K, const Scalars* s, uint64_t seed) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)N * K) return;
    int i = (int)(e / K), k = (int)(e % K);
    Rng rng(seed, 102u, 0ull, e);
    double w = sqrt(s->sigma2) * rng.normal();    // factor.py:230-233
    W[e] = (N > 1 && k > i) ? 0.0 : w;
}
void launch_init_W(double* W, int N, int K, const Scalars* s, uint64_t seed, cudaStream_t st) {
    size_t n = (size_t)N * K;
    init_W_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(W, N, K, s, seed);
}

__global__ void clip_kernel(double* x, size_t n, double lo, double hi) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) x[e] = fmin(fmax(x[e], lo), hi);
}
void launch_clip(double* x, size_t n, double lo, double hi, cudaStream_t st) {
    clip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, lo, hi);
}

}  // namespace btf
