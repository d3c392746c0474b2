// Power-spectral-density features of (augmented) cycles, computed where the cycles already are (sm_100a).
//
// What it replaces (reference = PCGmix-EXTENDED): the third block of `classical.feature_vector_seg`
// (classical.py:358-643), run per augmented cycle on the host by the loop at train_model.py:519-532.  For the
// whole beat (RR = data[:f4]), the systole (data[f1:f2]) and the diastole (data[f3:f4]):
//
//   freqs, psd = scipy.signal.welch(segment, 1000)      Hann window of min(256, n) samples (periodic), half
//                                                       overlap, mean removed per window, |rfft|^2 / (fs * sum w^2),
//                                                       one-sided (inner bins doubled), mean over the windows
//   integral   = np.trapz(|hilbert(psd)|, dx=5)         Hilbert envelope OF THE SPECTRUM, trapezoid rule
//   normalized = psd / integral
//   26 means: psd and normalized over all bins, and over the bins with lo <= freqs <= hi for twelve bands
//   (25-40, 40-60, ..., 180-200, 200-250, 250-300, 300-400 Hz; an empty band gives NaN like np.mean([])),
//
// then round(mean normalized systole / RR, 4) and the same for the diastole: 80 values per cycle.
//
// One CTA per cycle, float64 transforms.  Full windows (256 samples: all windows of a segment that is at least that
// long) go through a radix-2 FFT in shared memory.  A shorter segment is a single window of its own length (systoles
// nearly always: every cycle has its own transform length), and its spectrum (at most 128 bins) is evaluated
// directly: thread k accumulates bin k over the window, rotating its twiddle by one complex multiplication per
// sample (re-seeded from a table of the nper twiddles every 64 samples) — no FFT plan per length.  The Hilbert envelope of the spectrum is the circular
// convolution with the discrete Hilbert kernel (see feature_kernels.cu), again direct.
//
// Numerics.  SciPy computes all of this in float32 for float32 cycles (single-precision pocketfft).  Here the
// windowed samples are rounded to float32 like there, the transform and the convolution are accumulated in
// float64 and rounded to float32 where the reference holds float32 arrays (psd, envelope, normalized), and the
// bin frequencies are the same float64 products k * (1 / (nper * (1 / fs))) that np.fft.rfftfreq forms, so that a
// bin sitting exactly on a band edge (250 Hz = bin 64 of a 256-sample window) is classified alike.  Parity is a
// tolerance (tests: 2e-5 relative; a NumPy emulation of this file's arithmetic sits within 2e-6 of the reference's
// own statements on tests/golden/cycle_psd_features.npz), NaN patterns are identical.
//
// An empty segment makes the reference raise (hilbert of nothing); here the cycle's features are NaN and
// PCGMIX_ERR_EMPTY_STATE is raised.

#include "common.cuh"

namespace pcgmix {

namespace {

constexpr int kPsdThreads = 256;
constexpr int kMaxWindow = 256;                 // scipy.signal.welch's default nperseg
constexpr int kMaxBins = kMaxWindow / 2 + 1;
constexpr int kBands = 12;
__constant__ int c_band_lo[kBands] = {25, 40, 60, 80, 100, 120, 140, 160, 180, 200, 250, 300};
__constant__ int c_band_hi[kBands] = {40, 60, 80, 100, 120, 140, 160, 180, 200, 250, 300, 400};

struct PsdArgs {
    const float* x;
    const int32_t* frames;
    int32_t frame_stride;
    int32_t B, C, L, channel, fs;
    float* features;
    int32_t* err;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// sum over the CTA; every thread gets the result
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = threadIdx.x < kPsdThreads / 32 ? scratch[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) {
        t = warp_sum(t);
        if (threadIdx.x == 0) scratch[kPsdThreads / 32] = t;
    }
    __syncthreads();
    const double r = scratch[kPsdThreads / 32];
    __syncthreads();
    return r;
}

__device__ __forceinline__ float round4(float v) {      // NumPy's round(float32, 4)
    return __fdiv_rn(rintf(__fmul_rn(v, 10000.0f)), 10000.0f);
}

// periodic Hann window of n samples, float64 value rounded to float32 (scipy.signal.get_window('hann', n))
__device__ __forceinline__ float hann(int t, int n) {
    return n == 1 ? 1.0f : static_cast<float>(0.5 - 0.5 * cospi(2.0 * static_cast<double>(t) / static_cast<double>(n)));
}

__global__ void __launch_bounds__(kPsdThreads) cycle_psd_kernel(const __grid_constant__ PsdArgs a) {
    __shared__ double2 s_tw[kMaxWindow];         // (cos, sin)(2 pi j / nper)
    __shared__ double s_y[kMaxWindow];           // the window: (x - mean) * w, rounded to float32 like the reference's
    __shared__ double2 s_f[kMaxWindow];          // a full window's FFT, in place
    __shared__ double s_acc[kMaxBins];           // sum of the windows' spectra
    __shared__ float s_psd[kMaxBins];
    __shared__ float s_norm[kMaxBins];
    __shared__ float s_env[kMaxBins];
    __shared__ double s_h[kMaxBins];             // discrete Hilbert kernel of M points
    __shared__ double s_red[kPsdThreads / 32 + 1];
    __shared__ float s_mean_norm[3];
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const float* __restrict__ row = a.x + (static_cast<size_t>(b) * a.C + a.channel) * a.L;
    float* __restrict__ out = a.features + static_cast<size_t>(b) * PCGMIX_CYCLE_PSD_FEATURES;
    int c[5];
    bool sane = true;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int f = __ldg(a.frames + static_cast<size_t>(b) * a.frame_stride + s);
        sane = sane && f >= 0;
        c[s] = min(max(f, 0), a.L);
    }
    // the reference's slices: RR = data[:f4], systole = data[f1:f2], diastole = data[f3:f4]
    const int beg[3] = {0, c[1], c[3]};
    const int end[3] = {c[4], c[2], c[4]};
    bool empty = !sane;
#pragma unroll
    for (int s = 0; s < 3; ++s) empty = empty || end[s] <= beg[s];
    if (empty) {
        for (int i = tid; i < PCGMIX_CYCLE_PSD_FEATURES; i += kPsdThreads) out[i] = __int_as_float(0x7fc00000);
        if (tid == 0 && a.err != nullptr) atomicOr(a.err, static_cast<int>(PCGMIX_ERR_EMPTY_STATE));
        return;
    }
    const double fs = static_cast<double>(a.fs);

#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
        const int n = end[s] - beg[s];
        const int nper = min(kMaxWindow, n);
        const int step = nper - nper / 2;
        const int nseg = (n - nper / 2) / step;
        const int M = nper / 2 + 1;
        // twiddles, window energy
        double wsq = 0.0;
        float w_mine = 0.0f;                                 // this thread's sample of the window
        if (tid < nper) {
            double sn, cs;
            sincospi(2.0 * static_cast<double>(tid) / static_cast<double>(nper), &sn, &cs);
            s_tw[tid] = make_double2(cs, sn);
            w_mine = hann(tid, nper);
            wsq = static_cast<double>(w_mine) * static_cast<double>(w_mine);
        }
        if (tid < kMaxBins) s_acc[tid] = 0.0;
        const double scale = 1.0 / (fs * block_sum(wsq, s_red));
        // Welch: the windows' one-sided density spectra, summed
        for (int wdx = 0; wdx < nseg; ++wdx) {
            const float v = tid < nper ? __ldg(row + beg[s] + wdx * step + tid) : 0.0f;
            const float mean = static_cast<float>(block_sum(static_cast<double>(v), s_red) / static_cast<double>(nper));
            const double y_mine = static_cast<double>(__fmul_rn(__fsub_rn(v, mean), w_mine));
            if (nper == kMaxWindow) {
                // a full window (all but the short segments' single one): radix-2 FFT of the 256 real samples in shared
                // memory, float64 — 8 passes of 128 butterflies against 129 bins x 256 samples of the direct sum below
                s_f[__brev(static_cast<unsigned>(tid)) >> 24] = make_double2(y_mine, 0.0);
                __syncthreads();
#pragma unroll 1
                for (int half = 1; half < kMaxWindow; half <<= 1) {
                    if (tid < kMaxWindow / 2) {
                        const int j = tid & (half - 1);
                        const int i0 = ((tid - j) << 1) + j;
                        const double2 tw = s_tw[j * (kMaxWindow / 2 / half)];      // e^(-2 pi i j / (2 half)) = (tw.x, -tw.y)
                        const double2 lo = s_f[i0], hi = s_f[i0 + half];
                        const double tr = hi.x * tw.x + hi.y * tw.y;
                        const double ti = hi.y * tw.x - hi.x * tw.y;
                        s_f[i0] = make_double2(lo.x + tr, lo.y + ti);
                        s_f[i0 + half] = make_double2(lo.x - tr, lo.y - ti);
                    }
                    __syncthreads();
                }
                if (tid < M) {
                    const double2 X = s_f[tid];
                    double p = (X.x * X.x + X.y * X.y) * scale;
                    if (tid > 0 && tid < M - 1) p *= 2.0;          // not DC, not the Nyquist bin
                    s_acc[tid] += p;
                }
                __syncthreads();
                continue;
            }
            if (tid < nper) s_y[tid] = y_mine;
            __syncthreads();
            if (tid < M) {
                // bin `tid`: sum_t y[t] e^(-2 pi i tid t / nper); the twiddle advances by one complex multiplication per
                // sample (error ~ nper * 1e-16) and is re-read from the table every 64 samples
                const double2 w = s_tw[tid];                  // tid < M <= nper except nper = 1 (M = 1, tid = 0)
                double re = 0.0, im = 0.0;
                for (int t0 = 0; t0 < nper; t0 += 64) {
                    double2 tw = s_tw[(tid * t0) % nper];
                    const int t1 = min(t0 + 64, nper);
                    for (int t = t0; t < t1; ++t) {
                        const double y = s_y[t];
                        re = fma(y, tw.x, re);
                        im = fma(y, tw.y, im);
                        const double c = tw.x * w.x - tw.y * w.y;
                        tw.y = tw.y * w.x + tw.x * w.y;
                        tw.x = c;
                    }
                }
                double p = (re * re + im * im) * scale;
                const bool inner = tid > 0 && ((nper & 1) || tid < M - 1);      // not DC, not the Nyquist bin of an even length
                if (inner) p *= 2.0;
                s_acc[tid] += p;
            }
            __syncthreads();
        }
        // mean over the windows (float32 array from here on), Hilbert kernel of M points
        if (tid < M) {
            s_psd[tid] = static_cast<float>(s_acc[tid] / static_cast<double>(nseg));
            double h = 0.0;
            if (tid > 0) {
                double sn, cs;
                sincospi(static_cast<double>(tid) / static_cast<double>(M), &sn, &cs);
                if (M & 1) h = (cs - ((tid & 1) ? -1.0 : 1.0)) / (sn * static_cast<double>(M));
                else h = (tid & 1) ? 2.0 * cs / (sn * static_cast<double>(M)) : 0.0;
            }
            s_h[tid] = h;
        }
        __syncthreads();
        if (tid < M) {
            double hx = 0.0;
            int idx = tid;                                   // (tid - m) mod M for m = 1, 2, ...
            for (int m = 1; m < M; ++m) {
                idx = idx == 0 ? M - 1 : idx - 1;
                hx = fma(s_h[m], static_cast<double>(s_psd[idx]), hx);
            }
            const double p = static_cast<double>(s_psd[tid]);
            s_env[tid] = static_cast<float>(sqrt(p * p + hx * hx));
        }
        __syncthreads();
        // np.trapz(envelope, dx=5)
        double part = 0.0;
        if (tid + 1 < M) part = 5.0 * (static_cast<double>(s_env[tid + 1]) + static_cast<double>(s_env[tid])) * 0.5;
        const float integral = static_cast<float>(block_sum(part, s_red));
        if (tid < M) s_norm[tid] = __fdiv_rn(s_psd[tid], integral);
        __syncthreads();
        // the 26 means of this segment: thread j computes feature j (even: psd, odd: normalized; 0/1: all bins)
        if (tid < 26) {
            const float* src = (tid & 1) ? s_norm : s_psd;
            const int band = tid / 2 - 1;
            const double lo = band >= 0 ? static_cast<double>(c_band_lo[band]) : 0.0;
            const double hi = band >= 0 ? static_cast<double>(c_band_hi[band]) : 0.0;
            // np.fft.rfftfreq(nper, 1 / fs): k * (1 / (nper * d)) with d = 1 / fs, all in float64
            const double val = __ddiv_rn(1.0, __dmul_rn(static_cast<double>(nper), __ddiv_rn(1.0, fs)));
            double sum = 0.0;
            int count = 0;
            for (int k = 0; k < M; ++k) {
                const double f = __dmul_rn(static_cast<double>(k), val);
                if (band < 0 || (lo <= f && f <= hi)) {
                    sum += static_cast<double>(src[k]);
                    ++count;
                }
            }
            const float mean = count > 0 ? static_cast<float>(sum / static_cast<double>(count)) : __int_as_float(0x7fc00000);
            out[26 * s + tid] = mean;
            if (tid == 1) s_mean_norm[s] = mean;
        }
        __syncthreads();
    }
    if (tid == 0) {
        out[78] = round4(__fdiv_rn(s_mean_norm[1], s_mean_norm[0]));      // systole / RR
        out[79] = round4(__fdiv_rn(s_mean_norm[2], s_mean_norm[0]));      // diastole / RR
    }
}

}  // namespace

cudaError_t launch_cycle_psd_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C,
                                      int32_t L, int32_t channel, int32_t fs, float* features, int32_t* err,
                                      cudaStream_t stream) {
    if (B == 0) return cudaSuccess;
    PsdArgs a{x, frames, frame_stride, B, C, L, channel, fs, features, err};
    cycle_psd_kernel<<<static_cast<unsigned>(B), kPsdThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace pcgmix
