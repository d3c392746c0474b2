// Per-cycle classical features of (augmented) cycles, computed where the cycles already are (sm_100a).
//
// What it replaces (reference = PCGmix-EXTENDED): with `args.classical_space` the training loop takes every
// augmented batch back to the host and runs `classical.feature_vector_seg` on one channel of every cycle in a
// Python loop, concatenating pandas columns (train_model.py:519-532).  The per-cycle arithmetic of that
// function's first blocks is a handful of segmented reductions over the four heart states:
//
//   amplitude block (classical.py:284-303)   np.max per state, six ratios round(a/b, 4)
//   envelope block  (classical.py:305-360)   |hilbert(segment)| for S1, systole, S2, diastole and the whole
//                                            beat; np.trapz(dx=5) integrals, eight rounded integral ratios,
//                                            five means, eight mean ratios
//
//   moments block   (classical.py:893-905)   scipy.stats.skew / kurtosis of the whole beat and the four states
//                                            (cycle_moments_kernel, at the end of this file)
//
// (the duration block, :248-283, is pcgmix_duration_features in segment_kernels.cu; the Welch-PSD block, :358-643,
// is pcgmix_cycle_psd_features in psd_kernels.cu; the librosa / PyWavelets / antropy blocks are not provided).
//
// Numerics.  The reference works on float32 cycles, so every quantity above is float32 there.  Amplitude block:
// exact — a maximum is a selection, NaN propagates like np.max, and NumPy's round(x, 4) on a float32 scalar is
// rint(x * 1e4f) / 1e4f in float32, reproduced with __fmul_rn / rintf / __fdiv_rn.  Envelope block: SciPy takes a
// single-precision FFT of each segment (arbitrary length), zeroes the negative frequencies and transforms back;
// the imaginary part of that analytic signal is the circular convolution of the segment with the discrete
// Hilbert kernel
//     N even:  h[m] = (2/N) cot(pi m / N) for odd m, 0 for even m
//     N odd :  h[m] = (1/N) (cos(pi m / N) - (-1)^m) / sin(pi m / N)
// which is evaluated here directly (O(N^2) FMAs per segment out of shared memory: 1.7 M per cycle, 0.55 ms for a
// 4096-cycle batch; the reference's SciPy calls take 0.4 ms per cycle on one host core).  Both are float32 computations of the same quantity with different rounding
// orders, so parity is a tolerance (tests: 2e-5 relative on integrals and means, one unit of the fourth
// decimal on the rounded ratios), not bit equality.
//
// Segments follow the reference's slices exactly: S1 = data[:f1] (from column 0, not from f0), systole =
// data[f1:f2], S2 = data[f2:f3], diastole = data[f3:f4], RR = data[:f4], every bound clamped to the row length
// like a Python slice.  An empty segment makes the reference raise (np.max of nothing); here the cycle's
// features are NaN and PCGMIX_ERR_EMPTY_STATE is raised.

#include "common.cuh"

namespace pcgmix {

namespace {

constexpr int kFeatThreads = 256;
constexpr int kTile = 9;               // envelope outputs per thread and pass: ADJACENT samples, so that a thread's taps slide
                                       // over one register window; odd, so that the threads of a warp (9 floats apart) hit
                                       // 32 different banks
constexpr int kGuard = 2 * kTile;      // floats of zeroed slack in front of / behind the staged segment and behind the kernel

__device__ __forceinline__ float round4(float v) {      // NumPy's round(float32, 4)
    return __fdiv_rn(rintf(__fmul_rn(v, 10000.0f)), 10000.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// sum over the CTA; every thread gets the result (two barriers)
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = threadIdx.x < kFeatThreads / 32 ? scratch[threadIdx.x] : 0.0f;
    if (threadIdx.x < 32) {
        t = warp_sum(t);
        if (threadIdx.x == 0) scratch[kFeatThreads / 32] = t;
    }
    __syncthreads();
    const float r = scratch[kFeatThreads / 32];
    __syncthreads();
    return r;
}

struct FeatArgs {
    const float* x;
    const int32_t* frames;
    int32_t frame_stride;
    int32_t B, C, L, channel;
    float* features;
    int32_t* err;
    int32_t what;              // bit 0: amplitude block, bit 1: envelope block
};

__global__ void __launch_bounds__(kFeatThreads) cycle_features_kernel(const __grid_constant__ FeatArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_red[kFeatThreads / 32 + 1];
    __shared__ float s_max[4][kFeatThreads / 32];
    __shared__ int s_nan[4];
    const int b = blockIdx.x;
    const float* __restrict__ row = a.x + (static_cast<size_t>(b) * a.C + a.channel) * a.L;
    float* __restrict__ out = a.features + static_cast<size_t>(b) * PCGMIX_CYCLE_FEATURES;
    int c[5];
    bool sane = true;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int f = __ldg(a.frames + static_cast<size_t>(b) * a.frame_stride + s);
        sane = sane && f >= 0;
        c[s] = min(max(f, 0), a.L);
    }
    // the reference's slices: S1 starts at column 0; a bound below its predecessor gives an empty slice
    const int beg[5] = {0, c[1], c[2], c[3], 0};
    const int end[5] = {c[1], c[2], c[3], c[4], c[4]};
    bool empty = !sane;
#pragma unroll
    for (int s = 0; s < 5; ++s) empty = empty || end[s] <= beg[s];
    if (empty) {
        for (int i = threadIdx.x; i < PCGMIX_CYCLE_FEATURES; i += kFeatThreads) out[i] = __int_as_float(0x7fc00000);
        if (threadIdx.x == 0 && a.err != nullptr) atomicOr(a.err, static_cast<int>(PCGMIX_ERR_EMPTY_STATE));
        return;
    }

    if (a.what & 1) {
        // ---- amplitude block: per-state maximum (NaN propagates), six rounded ratios ------------------------
        if (threadIdx.x < 4) s_nan[threadIdx.x] = 0;
        __syncthreads();
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        bool nan[4] = {false, false, false, false};
        for (int t = threadIdx.x; t < c[4]; t += kFeatThreads) {
            const float v = __ldg(row + t);
            const int s = (t >= c[1]) + (t >= c[2]) + (t >= c[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (s == k) {
                    mx[k] = fmaxf(mx[k], v);
                    nan[k] = nan[k] || (v != v);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float m = mx[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFullMask, m, o));
            if ((threadIdx.x & 31) == 0) s_max[k][threadIdx.x >> 5] = m;
            if (__any_sync(kFullMask, nan[k]) && (threadIdx.x & 31) == 0) atomicOr(&s_nan[k], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                m[k] = s_max[k][0];
                for (int w = 1; w < kFeatThreads / 32; ++w) m[k] = fmaxf(m[k], s_max[k][w]);
                if (s_nan[k]) m[k] = __int_as_float(0x7fc00000);
                out[k] = m[k];
            }
            out[4] = round4(__fdiv_rn(m[0], m[2]));      // S1 / S2
            out[5] = round4(__fdiv_rn(m[1], m[3]));      // systole / diastole
            out[6] = round4(__fdiv_rn(m[1], m[0]));      // systole / S1
            out[7] = round4(__fdiv_rn(m[1], m[2]));      // systole / S2
            out[8] = round4(__fdiv_rn(m[3], m[0]));      // diastole / S1
            out[9] = round4(__fdiv_rn(m[3], m[2]));      // diastole / S2
        }
        __syncthreads();
    }

    if (a.what & 2) {
        // ---- envelope block ---------------------------------------------------------------------------------
        // shared memory: [guard] xs[2N] [guard] (the segment twice, so that (n - m) mod N is a plain offset), hs[N] [guard]
        // (Hilbert kernel, zero behind tap N - 1), es[N] (envelope); N <= c[4] <= L.  The guards are zero: taps past
        // N - 1 multiply them with a zero kernel value, and 0 * garbage could be NaN.
        const int cap = c[4];
        float* xs = smem + kGuard;
        float* hs = xs + 2 * cap + kGuard;
        float* es = hs + cap + kGuard;
        float integral[5], mean[5];
#pragma unroll 1
        for (int s = 0; s < 5; ++s) {
            const int n0 = beg[s];
            const int N = end[s] - n0;
            for (int t = threadIdx.x; t < N; t += kFeatThreads) {
                const float v = __ldg(row + n0 + t);
                xs[t] = v;
                xs[t + N] = v;
                float h = 0.0f;
                if (t > 0) {
                    float sn, cs;
                    sincospif(static_cast<float>(t) / static_cast<float>(N), &sn, &cs);
                    if (N & 1) {
                        h = (cs - ((t & 1) ? -1.0f : 1.0f)) / (sn * static_cast<float>(N));
                    } else {
                        h = (t & 1) ? 2.0f * cs / (sn * static_cast<float>(N)) : 0.0f;
                    }
                }
                hs[t] = h;
            }
            if (threadIdx.x < kGuard) {
                xs[-1 - static_cast<int>(threadIdx.x)] = 0.0f;
                xs[2 * N + threadIdx.x] = 0.0f;
                hs[N + threadIdx.x] = 0.0f;
            }
            __syncthreads();
            // Hx[n] = sum_m h[m] x[(n - m) mod N], taps in ascending m.  A thread owns kTile adjacent outputs; a chunk of
            // kTile taps (every tap for odd N, the odd ones for even N — the even ones are zero) needs one contiguous
            // window of the segment in registers: 2 kTile - 1 (3 kTile - 2) shared loads and kTile broadcast loads of the
            // kernel for kTile^2 FMAs, instead of five loads per four.
            for (int first = kTile * threadIdx.x; first < N; first += kTile * kFeatThreads) {
                float acc[kTile];
#pragma unroll
                for (int j = 0; j < kTile; ++j) acc[j] = 0.0f;
                if (N & 1) {
                    for (int m = 1; m < N; m += kTile) {
                        const float* src = xs + (first + N - m - (kTile - 1));
                        float w[2 * kTile - 1];
#pragma unroll
                        for (int i = 0; i < 2 * kTile - 1; ++i) w[i] = src[i];
#pragma unroll
                        for (int u = 0; u < kTile; ++u) {
                            const float h = hs[m + u];
#pragma unroll
                            for (int j = 0; j < kTile; ++j) acc[j] = fmaf(h, w[(kTile - 1) + j - u], acc[j]);
                        }
                    }
                } else {
                    for (int m = 1; m < N; m += 2 * kTile) {
                        const float* src = xs + (first + N - m - 2 * (kTile - 1));
                        float w[3 * kTile - 2];
#pragma unroll
                        for (int i = 0; i < 3 * kTile - 2; ++i) w[i] = src[i];
#pragma unroll
                        for (int u = 0; u < kTile; ++u) {
                            const float h = hs[m + 2 * u];
#pragma unroll
                            for (int j = 0; j < kTile; ++j) acc[j] = fmaf(h, w[2 * (kTile - 1) + j - 2 * u], acc[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < kTile; ++j) {
                    const int n = first + j;
                    if (n < N) {
                        const float re = xs[n];
                        es[n] = sqrtf(fmaf(re, re, acc[j] * acc[j]));
                    }
                }
            }
            __syncthreads();
            float part_int = 0.0f, part_sum = 0.0f;
            for (int t = threadIdx.x; t < N; t += kFeatThreads) {
                part_sum += es[t];
                if (t + 1 < N) part_int += __fmul_rn(5.0f, es[t + 1] + es[t]) * 0.5f;      // np.trapz term, dx = 5
            }
            integral[s] = block_sum(part_int, s_red);
            mean[s] = block_sum(part_sum, s_red) / static_cast<float>(N);
        }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                out[10 + s] = integral[s];
                out[23 + s] = mean[s];
            }
            out[15] = round4(__fdiv_rn(integral[0], integral[2]));      // S1 / S2
            out[16] = round4(__fdiv_rn(integral[1], integral[3]));      // systole / diastole
            out[17] = round4(__fdiv_rn(integral[0], integral[4]));      // S1 / RR
            out[18] = round4(__fdiv_rn(integral[1], integral[4]));      // systole / RR
            out[19] = round4(__fdiv_rn(integral[2], integral[4]));      // S2 / RR
            out[20] = round4(__fdiv_rn(integral[3], integral[4]));      // diastole / RR
            out[21] = round4(__fdiv_rn(integral[1], integral[0]));      // systole / S1
            out[22] = round4(__fdiv_rn(integral[3], integral[2]));      // diastole / S2
            out[28] = __fdiv_rn(mean[0], mean[4]);                       // S1 / RR
            out[29] = __fdiv_rn(mean[1], mean[4]);                       // systole / RR
            out[30] = __fdiv_rn(mean[2], mean[4]);                       // S2 / RR
            out[31] = __fdiv_rn(mean[3], mean[4]);                       // diastole / RR
            out[32] = __fdiv_rn(mean[1], mean[3]);                       // systole / diastole
            out[33] = __fdiv_rn(mean[1], mean[0]);                       // systole / S1
            out[34] = __fdiv_rn(mean[3], mean[2]);                       // diastole / S2
            out[35] = __fdiv_rn(mean[0], mean[2]);                       // S1 / S2
        }
    }
}

// Skewness / kurtosis block (classical.py:893-905): scipy.stats.skew and scipy.stats.kurtosis (biased estimators, Fisher's
// definition) of the whole beat and the four states.  One warp per segment: the mean, then the second, third and fourth
// central moments.  The reference's float32 arithmetic is followed where it matters — the mean and the deviations
// x - mean are float32, the powers are float32 products ((d*d)*d, (d*d)*(d*d), as SciPy's exponentiation by squaring forms
// them) — and only the sums are carried in float64 instead of NumPy's pairwise float32; a segment whose variance is not
// above (eps * mean)^2 gives NaN like there.  Tests: 2e-5 relative + 2e-6 absolute (a skewness near zero is a difference
// of large terms in both implementations).
struct MomentArgs {
    const float* x;
    const int32_t* frames;
    int32_t frame_stride;
    int32_t B, C, L, channel;
    float* features;
    int32_t* err;
};

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

__global__ void __launch_bounds__(160) cycle_moments_kernel(const __grid_constant__ MomentArgs a) {
    const int b = blockIdx.x;
    const int s = threadIdx.x >> 5;                      // segment: RR, S1, systole, S2, diastole
    const int lane = threadIdx.x & 31;
    const float* __restrict__ row = a.x + (static_cast<size_t>(b) * a.C + a.channel) * a.L;
    float* __restrict__ out = a.features + static_cast<size_t>(b) * PCGMIX_CYCLE_MOMENT_FEATURES;
    int c[5];
    bool sane = true;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int f = __ldg(a.frames + static_cast<size_t>(b) * a.frame_stride + k);
        sane = sane && f >= 0;
        c[k] = min(max(f, 0), a.L);
    }
    const int beg = s == 0 || s == 1 ? 0 : c[s - 1];
    const int end = s == 0 ? c[4] : c[s];
    const int n = end - beg;
    if (!sane || n <= 0) {
        if (lane == 0) {
            out[s] = out[5 + s] = __int_as_float(0x7fc00000);
            if (a.err != nullptr) atomicOr(a.err, static_cast<int>(PCGMIX_ERR_EMPTY_STATE));
        }
        return;
    }
    double sum = 0.0;
    for (int t = lane; t < n; t += 32) sum += static_cast<double>(__ldg(row + beg + t));
    const float mean = static_cast<float>(warp_sum_f64(sum) / static_cast<double>(n));
    double s2 = 0.0, s3 = 0.0, s4 = 0.0;
    for (int t = lane; t < n; t += 32) {
        const float d = __fsub_rn(__ldg(row + beg + t), mean);
        const float d2 = __fmul_rn(d, d);
        s2 += static_cast<double>(d2);
        s3 += static_cast<double>(__fmul_rn(d2, d));
        s4 += static_cast<double>(__fmul_rn(d2, d2));
    }
    const float m2 = static_cast<float>(warp_sum_f64(s2) / static_cast<double>(n));
    const float m3 = static_cast<float>(warp_sum_f64(s3) / static_cast<double>(n));
    const float m4 = static_cast<float>(warp_sum_f64(s4) / static_cast<double>(n));
    if (lane == 0) {
        const float tiny = __fmul_rn(1.1920928955078125e-07f, mean);            // float32 eps * mean
        const bool flat = m2 <= __fmul_rn(tiny, tiny);
        const float not_a_number = __int_as_float(0x7fc00000);
        out[s] = flat ? not_a_number : __fdiv_rn(m3, static_cast<float>(pow(static_cast<double>(m2), 1.5)));
        out[5 + s] = flat ? not_a_number : __fsub_rn(__fdiv_rn(m4, __fmul_rn(m2, m2)), 3.0f);
    }
}

}  // namespace

cudaError_t launch_cycle_moment_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C,
                                         int32_t L, int32_t channel, float* features, int32_t* err, cudaStream_t stream) {
    if (B == 0) return cudaSuccess;
    MomentArgs a{x, frames, frame_stride, B, C, L, channel, features, err};
    cycle_moments_kernel<<<static_cast<unsigned>(B), 160, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_cycle_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C,
                                  int32_t L, int32_t channel, int32_t what, float* features, int32_t* err,
                                  cudaStream_t stream) {
    if (B == 0) return cudaSuccess;
    FeatArgs a{x, frames, frame_stride, B, C, L, channel, features, err, what};
    const size_t smem = (what & 2) ? (static_cast<size_t>(L) * 4 + 3 * kGuard) * sizeof(float) : 0;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    if (smem > 48 * 1024) {
        // (per device and cheap: set every time rather than remembered per device)
        const cudaError_t e = cudaFuncSetAttribute(cycle_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    cycle_features_kernel<<<static_cast<unsigned>(B), kFeatThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace pcgmix
