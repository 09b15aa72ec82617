// K3 (blocked): batched block-banded FP64 Cholesky + MVN draw for the V columns, K and the
// block half-bandwidth Q = tf_order + 1 as compile-time constants.
//
// Same mathematics as band_solve_kernel (solve_kernels.cu; replaces
// sample_mvn_from_precision + CHOLMOD, fast_mvn.py:33-74, and the kron/SpGEMM assembly of
// factor.py:396-408) but organised by K x K blocks so that the O(n kd^2) trailing update
// runs on the FP64 tensor pipe:
//   per block column t (right-looking):
//     P1  potrf of the diagonal block by ONE warp, lane = row, row held in registers,
//         pivots exchanged with warp shuffles                       (|| P4 of the previous step)
//     P2  L_ut = A_ut L_tt^-T for the Q blocks below (one thread per row, register row),
//         y_t = L_tt^-1 b_t rides along as one more row
//     P3  A_uv -= L_ut L_vt^T for the Q(Q+1)/2 trailing block pairs with mma.sync.m8n8k4.f64,
//         b_u -= L_ut y_t, block column t -> global memory
//     P4  the block row entering the window is assembled from the per-(j,t) statistics and
//         the trend-filtering band Delta^T diag(1/(lam2 tau2)) Delta
//   backward substitution by block columns (mean and draw together), factor blocks streamed
//   back through a cp.async double buffer.
// Any K <= KB runs on the KB-wide instantiation: the KB - K extra unknowns per depth are
// decoupled unit-variance dummies (identity in the diagonal blocks, zero elsewhere).
// The live window is the lower triangle of a (Q+1) x (Q+1) block grid; diagonal d of that
// grid is a circular buffer of Q+1-d blocks (slot = base_d + row mod (Q+1-d)), so the window
// advances without copying: the entering row overwrites exactly the retired column blocks.
#include <cstdio>
#include <cstdlib>
#include "kernels.h"

namespace btf {

namespace {

__device__ __forceinline__ void cp_async16b(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int KB, int Q>
struct BlkGeom {
    static constexpr int KS = KB + 4;                      // padded row stride of a block in shared memory
    static constexpr int BLK = KB * KS;
    static constexpr int NBLK = (Q + 1) * (Q + 2) / 2;
    static constexpr int ROWS = Q * KB + 1;                // trsm rows (+1: right-hand side)
    static constexpr int NTMIN = KB > 16 ? 128 : 64;       // small CTAs: all columns resident in one wave
    static constexpr int NT = (ROWS > NTMIN ? (ROWS + 31) / 32 * 32 : NTMIN);
    static constexpr int MINB = KB > 16 ? 2 : 65536 / (128 * NT);   // residency target: <= 128 registers per thread for KB <= 16
    static constexpr int KK = KB * KB;                     // unpadded block (global layout)
    // the backward double buffer [2][(Q+1) KK] aliases the window region
    static constexpr int WREG = NBLK * BLK > 2 * (Q + 1) * KK ? NBLK * BLK : 2 * (Q + 1) * KK;
    __host__ __device__ static constexpr int base(int d) { return d * (Q + 1) - d * (d - 1) / 2; }
    __device__ static __forceinline__ int slot(int a, int d) { return base(d) + a % (Q + 1 - d); }
};

}  // namespace

template <int KB, int Q>
__global__ void __launch_bounds__(BlkGeom<KB, Q>::NT, BlkGeom<KB, Q>::MINB) band_blocked_kernel(BandSolveArgs a) {
    using G = BlkGeom<KB, Q>;
    constexpr int KS = G::KS, BLK = G::BLK, NT = G::NT, KK = G::KK;
    extern __shared__ __align__(16) double sm[];
    const int Kr = a.K;                                  // true embedding size (<= KB)
    const int L = Kr * (Kr + 1) / 2, nco = L + Kr, kd = Q * Kr, LS = kd + 1;
    const int T = a.T, n = T * Kr;                       // true system size (outputs, noise indices)
    const int ni = T * KB;                               // padded system size (internal workspaces)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jl = blockIdx.x, jg = a.col_begin + jl;
    const unsigned full = 0xffffffffu;

    double* Wb = sm;                                   // [NBLK][BLK]  (backward: 2 x (Q+1) x KK block columns)
    double* bw = Wb + G::WREG;                         // [Q+1][KB]    right-hand-side window
    double* dinv = bw + (Q + 1) * KB;                  // [KB]
    double* ycur = dinv + KB;                          // [KB]
    double* xw = ycur + KB;                            // [2][Q+1][KB] backward solution window
    double* Pband = xw + 2 * (Q + 1) * KB;             // [T][Q+1]
    double* linv = Pband + (size_t)T * (Q + 1);        // [RD]
    double* red = linv + a.RD;                         // [40]
    unsigned short* pairtab = reinterpret_cast<unsigned short*>(red + 40);   // [L] packed index -> (i << 8 | c)
    __shared__ int fail_flag;

    const double scale = a.homoskedastic ? 1.0 / a.scal->nu2 : 1.0;
    const double lam2 = a.scal->lam2;
    for (int r = tid; r < a.RD; r += NT) {
        double pv = 1.0 / (lam2 * a.Tau2[(size_t)jg * a.RD + r]);
        if (a.prior_clip > 0.0) pv = fmin(fmax(pv, a.prior_clip), 1.0 / a.prior_clip);
        linv[r] = pv;
    }
    if (tid == 0) fail_flag = 0;
    for (int e = tid; e < Kr * Kr; e += NT) {
        const int i = e / Kr, c = e % Kr;
        if (c <= i) pairtab[tri(i, c)] = (unsigned short)((i << 8) | c);
    }
    __syncthreads();
    for (int e = tid; e < T * (Q + 1); e += NT) {
        double s = 0.0;
        for (int x = a.pm_ptr[e]; x < a.pm_ptr[e + 1]; ++x) s += a.pm_coef[x] * linv[a.pm_row[x]];
        Pband[e] = s;
    }
    __syncthreads();

    // global workspace of this column: block columns [T][Q+1][KB][KB], then y [n], 1/diag [n]
    double* Lg = a.work_L + (size_t)jl * a.work_L_stride;
    double* yg = a.work_y + (size_t)jl * a.work_y_stride;     // [y (ni) | 1 / diag(L) (ni)]
    const bool have_stats = a.stats != nullptr;
    const double* stats0 = have_stats ? a.stats + (size_t)jg * T * nco : nullptr;

#ifdef BTF_BAND_PROFILE
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long c0 = clock64(), c1;
#define PROF(i) do { c1 = clock64(); pc[i] += c1 - c0; c0 = c1; } while (0)
#else
#define PROF(i)
#endif
    double jitter = 0.0, eps = a.eps;
    int attempt = 0;
    bool failed = false;

    // assemble block row `arow` of the precision into the window (all diagonals d = 0..Q)
    auto init_row = [&](int arow, int t0, int nthreads) {
        // diagonal block: statistics + prior diagonal (+ jitter); strictly-upper part unused
        const double* sb = have_stats ? stats0 + (size_t)arow * nco : nullptr;
        double* D = Wb + G::slot(arow, 0) * BLK;
        const double pd = Pband[arow * (Q + 1)] + jitter;
        // only the lower triangle is ever read: stream the packed statistics contiguously
#pragma unroll 4
        for (int e = t0; e < L; e += nthreads) {
            const int i = pairtab[e] >> 8, c = pairtab[e] & 0xff;
            double v = 0.0;
            if (have_stats) {
                double sv = sb[e];
                for (int s = 1; s < a.nsplit; ++s) sv += sb[s * a.split_stride + e];
                v = sv * scale;
            }
            if (c == i) v += pd;
            if (a.diag_band) a.diag_band[((size_t)jl * n + arow * Kr + i) * LS + kd - (i - c)] = v;
            D[i * KS + c] = v;
        }
        // padding rows Kr..KB-1: identity
        for (int e = t0; e < (KB - Kr) * KB; e += nthreads) {
            const int i = Kr + e / KB, c = e % KB;
            if (c <= i) D[i * KS + c] = (c == i) ? 1.0 : 0.0;
        }
        for (int e = t0; e < KB; e += nthreads) {
            double sv = 0.0;
            if (have_stats && e < Kr) {
                sv = sb[L + e];
                for (int s = 1; s < a.nsplit; ++s) sv += sb[s * a.split_stride + L + e];
            }
            bw[(arow % (Q + 1)) * KB + e] = sv * scale;
        }
        // off-diagonal blocks (arow, arow - d) = P[arow][arow-d] * I
#pragma unroll
        for (int d = 1; d <= Q; ++d) {
            if (arow - d >= 0) {
                double* B = Wb + G::slot(arow, d) * BLK;
                const double pv = Pband[(arow - d) * (Q + 1) + d];
                for (int e = t0; e < KB * KB; e += nthreads) {
                    const int i = e / KB, c = e % KB;
                    B[i * KS + c] = (i == c && i < Kr) ? pv : 0.0;
                }
                if (a.diag_band)
                    for (int e = t0; e < Kr; e += nthreads)
                        a.diag_band[((size_t)jl * n + arow * Kr + e) * LS + kd - d * Kr] = pv;
            }
        }
    };

    while (true) {
        // ---- initial window: block rows 0..Q
        for (int r = 0; r <= Q && r < T; ++r) init_row(r, tid, NT);
        __syncthreads();
        PROF(0);
        bool broke = false;
        for (int t = 0; t < T; ++t) {
            // ================= P1: potrf of the diagonal block (warp 0)  ||  P4 of step t-1
            if (warp == 0) {
                double* D = Wb + G::slot(t, 0) * BLK;
                double ar[KB];
#pragma unroll
                for (int c = 0; c < KB; ++c) ar[c] = (lane < KB) ? D[lane * KS + c] : 0.0;
                bool ok = true;
#pragma unroll
                for (int j = 0; j < KB; ++j) {
                    const double d = __shfl_sync(full, ar[j], j);
                    if (!(d > 0.0) || isinf(d)) ok = false;
                    const double rinv = rsqrt(ok ? d : 1.0);      // shortens the serial pivot chain
                    const double ljj = d * rinv;
                    const double lij = (lane == j) ? ljj : ar[j] * rinv;
                    ar[j] = lij;
#pragma unroll
                    for (int k = 0; k < KB; ++k) {
                        if (k > j) {            // rectangular loop + constant predicate: fully unrollable
                            const double lkj = __shfl_sync(full, lij, k);
                            if (lane >= k) ar[k] -= lij * lkj;
                        }
                    }
                    if (lane == j) { dinv[j] = rinv; yg[ni + t * KB + j] = rinv; }
                }
                if (!ok && lane == 0) fail_flag = 1;
                if (lane < KB) {
#pragma unroll
                    for (int c = 0; c < KB; ++c) D[lane * KS + c] = (c <= lane) ? ar[c] : 0.0;
                }
            } else if (t > 0 && t + Q < T) {
                init_row(t + Q, tid - 32, NT - 32);
            }
            __syncthreads();
            PROF(1);
            if (fail_flag) { broke = true; break; }

            // ================= P2: triangular solves against L_tt (one thread per row)
            double x[KB];
            int u = 0, irow = 0;
            bool active = false, is_rhs = false;
            if (tid < Q * KB) {
                u = 1 + tid / KB; irow = tid % KB;
                active = (t + u < T);
            } else if (tid == Q * KB) {
                active = true; is_rhs = true;
            }
            if (active) {
                const double* src = is_rhs ? bw + (t % (Q + 1)) * KB : Wb + G::slot(t + u, u) * BLK + irow * KS;
#pragma unroll
                for (int c = 0; c < KB; ++c) x[c] = src[c];
                const double* Ltt = Wb + G::slot(t, 0) * BLK;
#pragma unroll
                for (int j = 0; j < KB; ++j) {
                    x[j] *= dinv[j];
#pragma unroll
                    for (int k = 0; k < KB; ++k)
                        if (k > j) x[k] -= x[j] * Ltt[k * KS + j];
                }
                if (is_rhs) {
#pragma unroll
                    for (int c = 0; c < KB; ++c) { ycur[c] = x[c]; yg[t * KB + c] = x[c]; }
                } else {
                    double* dst = Wb + G::slot(t + u, u) * BLK + irow * KS;
#pragma unroll
                    for (int c = 0; c < KB; ++c) dst[c] = x[c];
                }
            }
            __syncthreads();
            PROF(2);

            // ================= P3: trailing update, right-hand side update, spill block column t
            if (active && !is_rhs) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < KB; ++c) acc += x[c] * ycur[c];
                bw[((t + u) % (Q + 1)) * KB + irow] -= acc;
            }
            {   // block column t -> global (unpadded blocks): 16-byte copies, slots resolved per block
                double* dstg = Lg + (size_t)t * (Q + 1) * KK;
#pragma unroll
                for (int ub = 0; ub <= Q; ++ub) {
                    const bool live = t + ub < T;
                    const double* srcb = Wb + G::slot(t + ub, ub) * BLK;
                    double2* dst2 = reinterpret_cast<double2*>(dstg + ub * KK);
                    for (int e = tid; e < KK / 2; e += NT) {
                        const int i = e / (KB / 2), c2 = e % (KB / 2);
                        double2 v = make_double2(0.0, 0.0);
                        if (live) v = *reinterpret_cast<const double2*>(srcb + i * KS + 2 * c2);
                        dst2[e] = v;
                    }
                    if (a.diag_chol && live) {
                        for (int e = tid; e < KK; e += NT) {
                            const int i = e / KB, c = e % KB, dist = ub * Kr + i - c;
                            if (i < Kr && c < Kr && dist >= 0 && dist <= kd)
                                a.diag_chol[((size_t)jl * n + (t + ub) * Kr + i) * LS + kd - dist] = srcb[i * KS + c];
                        }
                    }
                }
            }
            {   // A_uv -= L_ut L_vt^T on the tensor pipe; pairs unrolled statically (window slots
                // are then modulo-by-constant), the 8x8 tiles of a pair are spread over the warps
                constexpr int TB = KB / 8, TPB = TB * TB;
#pragma unroll
                for (int uu = 1; uu <= Q; ++uu) {
                    if (t + uu >= T) continue;
                    const double* A = Wb + G::slot(t + uu, uu) * BLK;
#pragma unroll
                    for (int vv = 1; vv <= uu; ++vv) {
                        const double* B = Wb + G::slot(t + vv, vv) * BLK;
                        double* C = Wb + G::slot(t + uu, uu - vv) * BLK;
                        for (int tl = warp; tl < TPB; tl += NT / 32) {
                            const int tm = tl / TB, tn = tl % TB;
                            double* cp = C + (tm * 8 + (lane >> 2)) * KS + tn * 8 + (lane & 3) * 2;
                            double c0 = cp[0], c1 = cp[1];
                            const double* ap = A + (tm * 8 + (lane >> 2)) * KS + (lane & 3);
                            const double* bp = B + (tn * 8 + (lane >> 2)) * KS + (lane & 3);
#pragma unroll
                            for (int ks = 0; ks < KB / 4; ++ks) dmma884(c0, c1, -ap[ks * 4], bp[ks * 4]);
                            cp[0] = c0; cp[1] = c1;
                        }
                    }
                }
            }
            __syncthreads();
            PROF(3);
        }
        if (!broke) break;
        __syncthreads();
        if (tid == 0) fail_flag = 0;
        if (a.force_psd && attempt < a.attempts) {
            jitter += eps; eps *= 10.0; ++attempt;
            __syncthreads();
            continue;
        }
        failed = true;
        break;
    }
    if (tid == 0) {
        if (a.diag_retries) a.diag_retries[jl] = attempt;
        if (attempt) atomicAdd(&a.scal->retries_v, attempt);
        if (failed) atomicAdd(&a.scal->info_v, 1);
    }
    if (failed) {
        if (a.resid_partials && tid == 0) a.resid_partials[jl] = 0.0;
        return;
    }
    __syncthreads();

    // ---- backward substitution by block columns: x_t = L_tt^-T (w_t - sum_u L_ut^T x_{t+u})
    // warp 0 carries the conditional mean (w = y), warp 1 the draw (w = y + z)
    double* Lc = Wb;                                   // [2][(Q+1)*KK] double buffer
    constexpr int COLE = (Q + 1) * KK;
    const unsigned long long sweep = a.scal->sweep;
    double* Vout = a.V + (size_t)jg * n;
    auto fetch = [&](int t, int buf) {
        const double* src = Lg + (size_t)t * COLE;
        double* dst = Lc + buf * COLE;
        for (int e = tid; e < COLE / 2; e += NT) cp_async16b(dst + 2 * e, src + 2 * e);
    };
    fetch(T - 1, (T - 1) & 1);
    cp_commit();
    for (int t = T - 1; t >= 0; --t) {
        if (t > 0) fetch(t - 1, (t - 1) & 1);
        cp_commit();
        cp_wait<1>();
        __syncthreads();
        const double* Lt = Lc + (t & 1) * COLE;
        if (warp < 2 && lane < KB) {
            const int rhs = warp, k = lane;
            double r = yg[t * KB + k];
            if (rhs && k < Kr) {
                double z;
                if (a.z_inject) z = a.z_inject[(size_t)jg * n + t * Kr + k];
                else { Rng rng(a.seed, STREAM_V, sweep, (uint64_t)jg * n + t * Kr + k); z = rng.normal(); }
                r += z;
            }
#pragma unroll
            for (int ub = 1; ub <= Q; ++ub) {
                if (t + ub < T) {
                    const double* Lu = Lt + ub * KK;
                    const double* xv = xw + (rhs * (Q + 1) + (t + ub) % (Q + 1)) * KB;
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int i = 0; i < KB; i += 2) { s0 += Lu[i * KB + k] * xv[i]; s1 += Lu[(i + 1) * KB + k] * xv[i + 1]; }
                    r -= s0 + s1;
                }
            }
            // triangular solve with L_tt^T inside the warp
            double xk = 0.0;
            const double dk = yg[ni + t * KB + k];
#pragma unroll
            for (int j = KB - 1; j >= 0; --j) {
                const double xj = __shfl_sync(0xffffffffu >> (32 - KB), r * dk, j);
                if (k == j) xk = xj;
                if (k < j) r -= Lt[j * KB + k] * xj;
            }
            xw[(rhs * (Q + 1) + t % (Q + 1)) * KB + k] = xk;
            if (k < Kr) {
                if (rhs) Vout[t * Kr + k] = xk;
                else if (a.diag_mean) a.diag_mean[(size_t)jg * n + t * Kr + k] = xk;
            }
        }
        __syncthreads();
    }
    cp_wait<0>();
    __syncthreads();
    PROF(4);

    // ---- nu2 by-product: sum_t v_t^T A_t v_t - 2 v_t . b_t with the UNSCALED statistics
    if (a.resid_partials && have_stats) {
        double accum = 0.0;
        for (int c = tid; c < nco; c += NT) {
            int k1 = 0, k2 = 0;
            double wgt = -2.0;
            if (c < L) {
                k1 = (int)((sqrt(8.0 * c + 1.0) - 1.0) * 0.5);
                while (k1 * (k1 + 1) / 2 > c) --k1;
                while ((k1 + 1) * (k1 + 2) / 2 <= c) ++k1;
                k2 = c - k1 * (k1 + 1) / 2;
                wgt = k1 == k2 ? 1.0 : 2.0;
            } else {
                k1 = c - L;
            }
            const double* sb = stats0 + c;
            const double* v = Vout;
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
            int t = 0;
            // four independent chains of loads per thread (the statistics come from L2 / HBM)
            for (; t + 4 <= T; t += 4) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                for (int s = 0; s < a.nsplit; ++s) {
                    const double* sp = sb + s * a.split_stride + (size_t)t * nco;
                    s0 += sp[0]; s1 += sp[nco]; s2 += sp[2 * nco]; s3 += sp[3 * nco];
                }
                const double* vt = v + (size_t)t * Kr;
                p0 += s0 * vt[k1] * (c < L ? vt[k2] : 1.0);
                p1 += s1 * vt[Kr + k1] * (c < L ? vt[Kr + k2] : 1.0);
                p2 += s2 * vt[2 * Kr + k1] * (c < L ? vt[2 * Kr + k2] : 1.0);
                p3 += s3 * vt[3 * Kr + k1] * (c < L ? vt[3 * Kr + k2] : 1.0);
            }
            for (; t < T; ++t) {
                double sv = 0.0;
                for (int s = 0; s < a.nsplit; ++s) sv += sb[s * a.split_stride + (size_t)t * nco];
                p0 += sv * v[(size_t)t * Kr + k1] * (c < L ? v[(size_t)t * Kr + k2] : 1.0);
            }
            accum += wgt * ((p0 + p1) + (p2 + p3));
        }
        double tot = block_sum(accum, red);
        if (tid == 0) a.resid_partials[jl] = tot;
    }
    PROF(5);
#ifdef BTF_BAND_PROFILE
    if (tid == 0 && (jl == 0 || jl == a.ncols_loc - 1))
        printf("band col %d: init %lld | potrf+assemble %lld | trsm %lld | trailing %lld | backward %lld | resid %lld cycles\n",
               jl, pc[0], pc[1], pc[2], pc[3], pc[4], pc[5]);
#endif
}

template <int KB, int Q>
static void launch_blocked_t(const BandSolveArgs& a, cudaStream_t st) {
    using G = BlkGeom<KB, Q>;
    size_t smem = ((size_t)G::WREG + (size_t)(Q + 1) * KB + 2 * KB + 2 * (Q + 1) * KB + (size_t)a.T * (Q + 1) + a.RD + 48 + (KB * (KB + 1) / 2 + 3) / 4) *
                  sizeof(double);
    auto kern = band_blocked_kernel<KB, Q>;
    static size_t max_set = 0;
    if (smem > 48 * 1024 && smem > max_set) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        max_set = smem;
    }
    kern<<<a.ncols_loc, G::NT, smem, st>>>(a);
}

bool launch_band_solve_blocked(const BandSolveArgs& a, cudaStream_t st) {
    // default: the look-ahead kernel (band_lookahead.cu); BTF_BAND_V1=1 keeps this file's kernel for A/B runs
    static const bool v1 = getenv("BTF_BAND_V1") != nullptr;
    if (!v1 && launch_band_solve_lookahead(a, st)) return true;
    const int Q = a.order + 1;
#define BTF_BLK(KB_)                                                  \
    do {                                                              \
        switch (Q) {                                                  \
            case 1: launch_blocked_t<KB_, 1>(a, st); return true;     \
            case 2: launch_blocked_t<KB_, 2>(a, st); return true;     \
            case 3: launch_blocked_t<KB_, 3>(a, st); return true;     \
            case 4: launch_blocked_t<KB_, 4>(a, st); return true;     \
            default: return false;                                    \
        }                                                             \
    } while (0)
    if (a.K <= 8) BTF_BLK(8);
    if (a.K <= 16) BTF_BLK(16);
    if (a.K <= 32) BTF_BLK(32);
#undef BTF_BLK
    return false;
}

}  // namespace btf
