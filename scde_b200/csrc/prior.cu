// prior.cu -- scde.expression.prior (R/functions.R:225-254) and scde.failure.probability (:725-750) on the device: the
// step immediately before the hot path.  The reference does this in interpreted R with
// stats::density: expression magnitudes and drop-out weights of every (gene, cell), a weighted, mirrored Gaussian kernel
// density estimate on a grid of 2 length.out + 1 points, of which the non-negative half is the prior.  O(G C) work:
//   pass 1   v = log10(exp((log(count) - corr.b) / corr.a) + 1),  w = 1 - drop-out probability; stored (16 B per element);
//            sum of w (deterministic: per-CTA partials added in CTA order), largest finite v
//   select   (max.quantile < 1 only) the two order statistics of R's type-7 quantile by an 8-pass radix select on the
//            order-preserving integer image of v
//   binning  density.default's BinDist: every point (and its mirror image) is spread linearly over its two neighbouring
//            grid points of the 2n-point working grid.  The bins are 64-bit FIXED-POINT sums (2^62 = the total mass 1):
//            integer atomics make the result independent of the order of the additions, so the prior is reproducible
//            bit for bit (relative bin error <= 1e-10, against 1e-12 of the reference's own sequential double sum)
// The 2n-point convolution with the Gaussian kernel and the interpolation onto the output grid are host work on 2048
// values (api.cu).  Nothing here is on the differential-expression path itself.
#include "common.cuh"
#include <cfloat>
#include <cmath>

namespace scde {
namespace {

constexpr int PR_THREADS = 256;

__device__ __forceinline__ unsigned long long ordered_key(double v) {  // order-preserving double -> uint64
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// models: n_cells x 12 column-major.  fail = 1 / (exp(conc.a m [+ conc.a2 m^2] + conc.b) + 1), NaN -> 0 (:748)
__device__ __forceinline__ double failure_prob(double m, double ca, double cb, double ca2, bool sq) {
    double eta = m * ca;
    if (sq) eta += m * m * ca2;
    eta += cb;
    double v = 1.0 / (exp(eta) + 1.0);
    return v != v ? 0.0 : v;
}

__global__ void __launch_bounds__(PR_THREADS)
failure_probability_kernel(const int32_t *__restrict__ counts, const double *__restrict__ mag_in, int64_t n, int G,
                           const double *__restrict__ models, int C, int sq, double *__restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e / G);
        const double m = mag_in ? mag_in[e] : (log((double)counts[e]) - models[(size_t)3 * C + c]) / models[(size_t)4 * C + c];
        out[e] = failure_prob(m, models[(size_t)1 * C + c], models[(size_t)0 * C + c], sq ? models[(size_t)11 * C + c] : 0.0, sq != 0);
    }
}

// pass 1: one CTA owns a contiguous slice of the elements; part[blockIdx] = its sum of w; vmax_key = max over finite v
__global__ void __launch_bounds__(PR_THREADS)
prior_pass1_kernel(const int32_t *__restrict__ counts, int64_t n, int G, const double *__restrict__ models, int C, int sq,
                   double *__restrict__ v_out, double *__restrict__ w_out, double *__restrict__ part,
                   unsigned long long *__restrict__ vmax_key, unsigned long long *__restrict__ n_finite) {
    __shared__ double s_sum[PR_THREADS];
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t e0 = (int64_t)blockIdx.x * per, e1 = min(n, e0 + per);
    double acc = 0.0;
    unsigned long long kmax = 0ull, nf = 0ull;
    for (int64_t e = e0 + threadIdx.x; e < e1; e += PR_THREADS) {
        const int c = (int)(e / G);
        const double m = (log((double)counts[e]) - models[(size_t)3 * C + c]) / models[(size_t)4 * C + c];  // :226, :694-697
        const double f = failure_prob(m, models[(size_t)1 * C + c], models[(size_t)0 * C + c],
                                      sq ? models[(size_t)11 * C + c] : 0.0, sq != 0);
        const double v = log10(exp(m) + 1.0);  // :228
        const double w = 1.0 - f;              // :229
        v_out[e] = v;
        w_out[e] = w;
        acc += w;
        if (v < INFINITY) {  // x[x < Inf] (:235); NaN compares false
            ++nf;
            const unsigned long long k = ordered_key(v);
            kmax = k > kmax ? k : kmax;
        }
    }
    s_sum[threadIdx.x] = acc;
    __syncthreads();
    for (int o = PR_THREADS / 2; o > 0; o >>= 1) {  // fixed tree: deterministic
        if (threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = s_sum[0];
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmax = other > kmax ? other : kmax;
        nf += __shfl_xor_sync(0xffffffffu, nf, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(vmax_key, kmax);
        atomicAdd(n_finite, nf);
    }
}

// one pass of the radix select: histogram of byte `shift/8` of the keys that match `prefix` in the bits above it
__global__ void __launch_bounds__(PR_THREADS)
select_hist_kernel(const double *__restrict__ v, int64_t n, unsigned long long prefix, int shift,
                   unsigned long long *__restrict__ hist) {
    __shared__ unsigned int s_h[256];
    s_h[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned long long himask = shift >= 56 ? 0ull : (~0ull << (shift + 8));
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const double x = v[e];
        if (!(x < INFINITY)) continue;
        const unsigned long long k = ordered_key(x);
        if ((k & himask) == (prefix & himask)) atomicAdd(&s_h[(unsigned)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_h[threadIdx.x]);
}

// BinDist of the mirrored points.  bins[2n]: fixed point, 2^62 = total mass.  inv_sum = 1 / sum(w).
__global__ void __launch_bounds__(PR_THREADS)
prior_bin_kernel(const double *__restrict__ v, const double *__restrict__ w, int64_t n, double inv_sum, double lo, double xdelta,
                 int nbin, unsigned long long *__restrict__ bins) {
    extern __shared__ unsigned long long s_bins[];  // [nbin] (the upper half of the 2n working grid stays zero)
    for (int i = threadIdx.x; i < nbin; i += PR_THREADS) s_bins[i] = 0ull;
    __syncthreads();
    const int ixmax = nbin - 2;
    const double SCALE = 4611686018427387904.0;  // 2^62
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const double x = v[e];
        if (!isfinite(x)) continue;           // BinDist skips non-finite points
        const double wi = (w[e] * inv_sum) / 2;  // wts / sum(wts), then c(wts/2, wts/2) (:230, :237)
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const double xs = side ? x : -1 * x;
            const double xpos = (xs - lo) / xdelta;
            const int ix = (int)floor(xpos);
            const double fx = xpos - ix;
            if (0 <= ix && ix <= ixmax) {
                atomicAdd(&s_bins[ix], (unsigned long long)__double2ll_rn(wi * (1 - fx) * SCALE));
                atomicAdd(&s_bins[ix + 1], (unsigned long long)__double2ll_rn(wi * fx * SCALE));
            } else if (ix == -1) {
                atomicAdd(&s_bins[0], (unsigned long long)__double2ll_rn(wi * fx * SCALE));
            } else if (ix == ixmax + 1) {
                atomicAdd(&s_bins[ix], (unsigned long long)__double2ll_rn(wi * (1 - fx) * SCALE));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbin; i += PR_THREADS)
        if (s_bins[i]) atomicAdd(&bins[i], s_bins[i]);
}

}  // namespace

static int prior_grid(int64_t n) {
    int64_t b = (n + PR_THREADS - 1) / PR_THREADS;
    if (b > 148 * 8) b = 148 * 8;
    return (int)(b < 1 ? 1 : b);
}

int prior_pass1_blocks(int64_t n) { return prior_grid(n); }

cudaError_t launch_failure_probability(const int32_t *counts, const double *mag, int64_t n, int G, const double *models, int C,
                                       int sq, double *out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    failure_probability_kernel<<<prior_grid(n), PR_THREADS, 0, st>>>(counts, mag, n, G, models, C, sq, out);
    return cudaGetLastError();
}

cudaError_t launch_prior_pass1(const int32_t *counts, int64_t n, int G, const double *models, int C, int sq, double *v, double *w,
                               double *part, unsigned long long *vmax_key, unsigned long long *n_finite, cudaStream_t st) {
    prior_pass1_kernel<<<prior_grid(n), PR_THREADS, 0, st>>>(counts, n, G, models, C, sq, v, w, part, vmax_key, n_finite);
    return cudaGetLastError();
}

cudaError_t launch_select_hist(const double *v, int64_t n, unsigned long long prefix, int shift, unsigned long long *hist,
                               cudaStream_t st) {
    select_hist_kernel<<<prior_grid(n), PR_THREADS, 0, st>>>(v, n, prefix, shift, hist);
    return cudaGetLastError();
}

cudaError_t launch_prior_bins(const double *v, const double *w, int64_t n, double inv_sum, double lo, double xdelta, int nbin,
                              unsigned long long *bins, cudaStream_t st) {
    const size_t smem = sizeof(unsigned long long) * (size_t)nbin;
    cudaError_t e = cudaFuncSetAttribute(prior_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    prior_bin_kernel<<<prior_grid(n), PR_THREADS, smem, st>>>(v, w, n, inv_sum, lo, xdelta, nbin, bins);
    return cudaGetLastError();
}

double prior_key_to_double(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double d;
    memcpy(&d, &b, sizeof(d));
    return d;
}

}  // namespace scde
