// First convolution block of the reference's ResNet9-1D, forward pass, on the (augmented) cycles where they
// already are (sm_100a).
//
// What it replaces (reference = PCGmix-EXTENDED): the consumer of `augment`'s output in the training loop is
// `model(data)` (train_model.py:536), whose first layer is `conv_block(in_channels, filters[0])` =
// nn.Conv1d(C, F, kernel_size=3, padding=1) + nn.BatchNorm1d(F) + nn.ReLU (models.py:468-473, used as `conv1` at
// models.py:523, called at :538).  SURVEY section 8(f)4 names it as the second downstream consumer of the path.
//
// Shape of the work.  Input (B, C, L) float32 (C = 1 or 4 band-split channels, L = 2500), output (B, F, L) with
// F = 64: the output is 16x the input, so the block is bound by WRITING 4*B*F*L bytes to HBM (2.6 GB for a
// 4096-cycle batch); the arithmetic is 3*C = 12 FMAs per output value.  Nothing here is a contraction worth a
// tensor core (K = 12); the design question is how to touch the 2.6 GB exactly once even in TRAINING mode, where
// batch normalisation needs the mean and variance of the convolution's output over (B, L) before any output
// value can be written.
//
// Training-mode statistics without the output.  z[b,f,l] = bias_f + sum_i w[f,i] * v_i(b,l), where v(b,l) is the
// 3*C-vector of the zero-padded input patch (x[b,c,l-1], x[b,c,l], x[b,c,l+1]).  So over the batch
//     mean_f = bias_f + w_f . E[v]            var_f = w_f^T (E[v v^T] - E[v] E[v]^T) w_f
// and the 3C + 3C(3C+1)/2 = 90 (C = 4) patch moments are a reduction over the INPUT (164 MB), not over the
// output: `patch_moments_kernel` (float32 products, float64 sums), then `fold_kernel` (one CTA, float64) turns
// moments + parameters into one record per filter, {a_f * w[f, :], a_f * (bias_f - mean_f) + beta_f} with
// a_f = gamma_f / sqrt(var_f + eps), updates the running statistics like torch does (momentum, unbiased
// variance) and saves mean / inverse standard deviation, and `apply_kernel` writes relu(record . patch) — one
// pass over the output in either mode (evaluation mode skips the moments and folds the running statistics).
//
// Issue rate.  655 M output values x 12 FMAs: a three-register FFMA issues every second cycle per scheduler on
// this part, which would take as long as the 2.6 GB of stores.  The records of two adjacent filters are therefore
// interleaved and the kernel works on (filter 2p, filter 2p+1) pairs with packed `fma.rn.f32x2` (FFMA2 in SASS):
// the input window is duplicated into (x, x) register pairs once per thread, and one packed instruction advances
// both filters — half the fma-pipe time, the same shared-memory loads.
//
// Numerics: float32 FMAs in a fixed order; torch's own result depends on the backend's summation order, so parity
// is a tolerance (tests: 2e-5 relative + 2e-5 absolute against the reference's modules run on the CPU, running
// statistics 1e-5 relative).  NaN propagates through the ReLU like torch's (max.NaN).

#include "common.cuh"

namespace pcgmix {

namespace {

constexpr int kThreads = 256;
constexpr int kMomentThreads = 128;     // patch_moments_kernel

__host__ __device__ constexpr int patch_len(int C) { return 3 * C; }
__host__ __device__ constexpr int moment_count(int C) { return patch_len(C) + patch_len(C) * (patch_len(C) + 1) / 2; }
// a record holds TWO filters, interleaved: {w0[0], w1[0], w0[1], w1[1], ..., shift0, shift1}, padded to 128-bit units
__host__ __device__ constexpr int pair_stride(int C) { return (2 * patch_len(C) + 2 + 3) & ~3; }  // floats

__device__ __forceinline__ float relu_nan(float v) {    // torch's ReLU keeps NaN; fmaxf would drop it
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));     // two IEEE float32 FMAs, one issue slot
    return d;
}

// moments[i] = sum v_i, moments[P + tri(i, j)] = sum v_i v_j (i <= j), over all (b, l).  `moments` zeroed by the caller.
// Persistent: the grid is the number of CTAs the GPU holds at once.  A thread takes four adjacent positions at a time
// (one 128-bit load and two border samples per channel, like apply_kernel; the next window is fetched before the
// current one's arithmetic), so a batch of 3C loads feeds 4 x 48 packed FMAs.  Row a of the product triangle is
// accumulated as packed pairs (v[a], v[a]) * (v[2k], v[2k+1]), k >= a / 2 — 42 FFMA2 instead of 78 FFMA at C = 4 (one
// product per odd row is computed twice and dropped).  A thread sums its own positions in float32 (a hundred-odd
// terms), a warp's 32 partial sums are added by a float32 shuffle tree, everything above that is float64.
template <int C, bool VEC>
__global__ void __launch_bounds__(kMomentThreads, 2) patch_moments_kernel(const float* __restrict__ x, uint32_t total_quads,
                                                                         uint32_t quads_per_row, uint32_t L,
                                                                         double* __restrict__ moments) {
    constexpr int P = patch_len(C);
    constexpr int HP = (P + 1) / 2;                         // packed pairs per patch (the last one padded with 0 when P is odd)
    constexpr int NM = moment_count(C);
    __shared__ double s_acc[NM];
    for (int i = threadIdx.x; i < NM; i += kMomentThreads) s_acc[i] = 0.0;
    __syncthreads();

    unsigned long long s1[HP];
    unsigned long long s2[P][HP];                           // row a uses k >= a / 2 only; the rest is never touched
#pragma unroll
    for (int k = 0; k < HP; ++k) s1[k] = 0ull;
#pragma unroll
    for (int a = 0; a < P; ++a)
#pragma unroll
        for (int k = 0; k < HP; ++k) s2[a][k] = 0ull;
    const unsigned long long ones = pack2(1.0f, 1.0f);

    // samples l0-1 .. l0+4 of every channel, zero outside the row
    auto fetch = [&](uint32_t t, float (&w)[C][6], uint32_t& l0) {
        const uint32_t b = t / quads_per_row;
        l0 = (t - b * quads_per_row) * 4u;
        const float* row = x + static_cast<size_t>(b) * C * L;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float* r = row + static_cast<size_t>(c) * L;
            if (VEC) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(r + l0));
                w[c][1] = m.x; w[c][2] = m.y; w[c][3] = m.z; w[c][4] = m.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) w[c][1 + j] = l0 + j < L ? __ldg(r + l0 + j) : 0.0f;
            }
            w[c][0] = l0 > 0 ? __ldg(r + l0 - 1) : 0.0f;
            w[c][5] = l0 + 4 < L ? __ldg(r + l0 + 4) : 0.0f;
        }
    };
    const uint32_t step = gridDim.x * kMomentThreads;
    uint32_t t = blockIdx.x * kMomentThreads + threadIdx.x;
    float nw[C][6];
    uint32_t nl0 = 0;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < 6; ++j) nw[c][j] = 0.0f;
    if (t < total_quads) fetch(t, nw, nl0);
    while (t < total_quads) {
        float w[C][6];
        unsigned long long wd[C][6];                        // (x, x)
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int j = 0; j < 6; ++j) { w[c][j] = nw[c][j]; wd[c][j] = pack2(nw[c][j], nw[c][j]); }
        const uint32_t l0 = nl0;
        const uint32_t nt = t + step;
        if (nt < total_quads && nt > t) fetch(nt, nw, nl0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!VEC && l0 + j >= L) break;                 // ragged last quad of a row: no such position
            unsigned long long vp[HP];
#pragma unroll
            for (int k = 0; k < HP; ++k) {
                const int i0 = 2 * k, i1 = 2 * k + 1;
                const float lo = w[i0 / 3][j + i0 % 3];
                const float hi = i1 < P ? w[i1 / 3][j + i1 % 3] : 0.0f;
                vp[k] = pack2(lo, hi);
            }
#pragma unroll
            for (int k = 0; k < HP; ++k) s1[k] = fma2(vp[k], ones, s1[k]);
#pragma unroll
            for (int a = 0; a < P; ++a) {
                const unsigned long long va = wd[a / 3][j + a % 3];
#pragma unroll
                for (int k = a / 2; k < HP; ++k) s2[a][k] = fma2(va, vp[k], s2[a][k]);
            }
        }
        if (nt <= t) break;                                 // 32-bit wrap-around guard
        t = nt;
    }

    const int lane = threadIdx.x & 31;
    auto warp_add = [&](float v, int slot) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
        if (lane == 0) atomicAdd(&s_acc[slot], static_cast<double>(v));
    };
#pragma unroll
    for (int k = 0; k < HP; ++k) {
        float lo, hi;
        unpack2(s1[k], lo, hi);
        warp_add(lo, 2 * k);
        if (2 * k + 1 < P) warp_add(hi, 2 * k + 1);
    }
    int idx = 0;
#pragma unroll
    for (int a = 0; a < P; ++a) {
#pragma unroll
        for (int q = a; q < P; ++q, ++idx) {
            float lo, hi;
            unpack2(s2[a][q / 2], lo, hi);
            warp_add((q & 1) ? hi : lo, P + idx);
        }
    }
    __syncthreads();
    for (int i2 = threadIdx.x; i2 < NM; i2 += kMomentThreads) atomicAdd(&moments[i2], s_acc[i2]);
}

struct FoldArgs {
    const float* weight;      // [F][C][3]
    const float* bias;        // [F] or nullptr
    const float* gamma;       // [F] or nullptr (1)
    const float* beta;        // [F] or nullptr (0)
    float* running_mean;      // [F] or nullptr
    float* running_var;       // [F] or nullptr
    const double* moments;    // batch statistics from these (nullptr: use the running statistics)
    float* records;           // [(F + 1) / 2][pair_stride]
    float* save_mean;         // [F] or nullptr
    float* save_invstd;       // [F] or nullptr
    double n;                 // B*L
    double eps;
    double momentum;
    int F;
    int C;
};

// One CTA.  Means and (doubled off-diagonal) covariances of the patch first, one thread per entry, then one thread per
// filter, float64.  Writes the folded record of every filter; with batch statistics also the running-statistics update
// (running = (1 - m) running + m batch, variance unbiased) torch performs in training mode.
constexpr int kFoldThreads = 128;
__global__ void __launch_bounds__(kFoldThreads) fold_kernel(FoldArgs a) {
    constexpr int kMaxP = patch_len(4);
    __shared__ double s_mu[kMaxP];
    __shared__ double s_cov[kMaxP * (kMaxP + 1) / 2];
    const int P = 3 * a.C;
    const int PS = (2 * P + 2 + 3) & ~3;
    if (a.moments != nullptr) {
        const double inv_n = 1.0 / a.n;
        for (int i = threadIdx.x; i < P; i += kFoldThreads) s_mu[i] = a.moments[i] * inv_n;
        __syncthreads();
        for (int t = threadIdx.x; t < P * (P + 1) / 2; t += kFoldThreads) {
            int i = 0, rem = t;
            while (rem >= P - i) { rem -= P - i; ++i; }
            const int j = i + rem;
            s_cov[t] = (a.moments[P + t] * inv_n - s_mu[i] * s_mu[j]) * (j == i ? 1.0 : 2.0);
        }
        __syncthreads();
    }
    for (int f = threadIdx.x; f < a.F; f += kFoldThreads) {
        const float* w = a.weight + static_cast<size_t>(f) * P;
        const double bias = a.bias ? static_cast<double>(a.bias[f]) : 0.0;
        double mean, var;
        if (a.moments != nullptr) {
            double m = 0.0, q = 0.0;
            int t = 0;
            for (int i = 0; i < P; ++i) {
                const double wi = static_cast<double>(w[i]);
                m += wi * s_mu[i];
                for (int j = i; j < P; ++j, ++t) q += wi * static_cast<double>(w[j]) * s_cov[t];
            }
            mean = bias + m;
            var = q < 0.0 ? 0.0 : q;                       // rounding only; NaN stays NaN like torch's statistics
            if (a.running_mean != nullptr)
                a.running_mean[f] = static_cast<float>((1.0 - a.momentum) * static_cast<double>(a.running_mean[f]) + a.momentum * mean);
            if (a.running_var != nullptr) {
                const double unbiased = a.n > 1.0 ? var * (a.n / (a.n - 1.0)) : var;
                a.running_var[f] = static_cast<float>((1.0 - a.momentum) * static_cast<double>(a.running_var[f]) + a.momentum * unbiased);
            }
        } else {
            mean = static_cast<double>(a.running_mean[f]);
            var = static_cast<double>(a.running_var[f]);
        }
        const double invstd = 1.0 / sqrt(var + a.eps);
        const double scale = (a.gamma ? static_cast<double>(a.gamma[f]) : 1.0) * invstd;
        const double shift = (bias - mean) * scale + (a.beta ? static_cast<double>(a.beta[f]) : 0.0);
        const int h = f & 1;
        float* rec = a.records + static_cast<size_t>(f >> 1) * PS;
        for (int i = 0; i < P; ++i) rec[2 * i + h] = static_cast<float>(static_cast<double>(w[i]) * scale);
        rec[2 * P + h] = static_cast<float>(shift);
        if (h == 0) {
            for (int i = 2 * P + 2; i < PS; ++i) rec[i] = 0.0f;
            if (f + 1 == a.F)                                   // odd filter count: the pair's second half is empty
                for (int i = 0; i <= P; ++i) rec[2 * i + 1] = 0.0f;
        }
        if (a.save_mean != nullptr) a.save_mean[f] = static_cast<float>(mean);
        if (a.save_invstd != nullptr) a.save_invstd[f] = static_cast<float>(invstd);
    }
}

// One thread = four adjacent output positions of one cycle, all F filters, two filters at a time: the 6-sample
// input window of every channel stays in registers as (x, x) pairs, a filter pair's record is a few broadcast
// 128-bit shared loads, and a warp's store of one filter is 512 contiguous bytes.  VEC: rows are 16-byte aligned
// multiples of four samples (128-bit loads / stores).
template <int C, bool VEC>
__global__ void __launch_bounds__(kThreads, 2) apply_kernel(const float* __restrict__ x, const float* __restrict__ records,
                                                             float* __restrict__ out, uint32_t total_quads, uint32_t quads_per_row,
                                                             uint32_t L, int F, int pairs_per_group) {
    constexpr int P = patch_len(C);
    constexpr int PS = pair_stride(C);
    extern __shared__ float4 s_rec[];                       // [pairs_per_group][PS / 4]
    // blockIdx.y: a group of filter pairs (one group for large batches; small batches are split so that the GPU is filled)
    const int p_begin = blockIdx.y * pairs_per_group;
    const int p_end = min((F + 1) >> 1, p_begin + pairs_per_group);
    {
        const float4* src = reinterpret_cast<const float4*>(records) + p_begin * (PS / 4);
        for (int i = threadIdx.x; i < (p_end - p_begin) * (PS / 4); i += kThreads) s_rec[i] = __ldg(src + i);
    }
    __syncthreads();

    const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total_quads) return;
    const uint32_t b = t / quads_per_row;
    const uint32_t l0 = (t - b * quads_per_row) * 4u;
    const float* row = x + static_cast<size_t>(b) * C * L;

    unsigned long long win[C][6];                           // samples l0-1 .. l0+4 as (x, x), zero outside the row
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float* r = row + static_cast<size_t>(c) * L;
        float w[6];
        if (VEC) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(r + l0));
            w[1] = m.x; w[2] = m.y; w[3] = m.z; w[4] = m.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) w[1 + j] = l0 + j < L ? __ldg(r + l0 + j) : 0.0f;
        }
        w[0] = l0 > 0 ? __ldg(r + l0 - 1) : 0.0f;
        w[5] = l0 + 4 < L ? __ldg(r + l0 + 4) : 0.0f;
#pragma unroll
        for (int j = 0; j < 6; ++j) win[c][j] = pack2(w[j], w[j]);
    }

    float* o = out + (static_cast<size_t>(b) * F + 2 * static_cast<size_t>(p_begin)) * L + l0;
#pragma unroll 2
    for (int p = p_begin; p < p_end; ++p) {
        unsigned long long rec[PS / 2];                     // rec[i] = (w0[i], w1[i]); rec[P] = (shift0, shift1)
#pragma unroll
        for (int i = 0; i < PS / 4; ++i) {
            const float4 q = s_rec[(p - p_begin) * (PS / 4) + i];
            rec[2 * i + 0] = pack2(q.x, q.y);
            rec[2 * i + 1] = pack2(q.z, q.w);
        }
        unsigned long long a0 = rec[P], a1 = rec[P], a2 = rec[P], a3 = rec[P];
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const unsigned long long w = rec[3 * c + k];
                a0 = fma2(w, win[c][k + 0], a0);
                a1 = fma2(w, win[c][k + 1], a1);
                a2 = fma2(w, win[c][k + 2], a2);
                a3 = fma2(w, win[c][k + 3], a3);
            }
        }
        float e0, e1, e2, e3, g0, g1, g2, g3;               // e: filter 2p, g: filter 2p + 1
        unpack2(a0, e0, g0); unpack2(a1, e1, g1); unpack2(a2, e2, g2); unpack2(a3, e3, g3);
        e0 = relu_nan(e0); e1 = relu_nan(e1); e2 = relu_nan(e2); e3 = relu_nan(e3);
        g0 = relu_nan(g0); g1 = relu_nan(g1); g2 = relu_nan(g2); g3 = relu_nan(g3);
        const bool second = 2 * p + 1 < F;
        if (VEC) {
            __stcs(reinterpret_cast<float4*>(o), make_float4(e0, e1, e2, e3));
            if (second) __stcs(reinterpret_cast<float4*>(o + L), make_float4(g0, g1, g2, g3));
        } else {
            if (l0 + 0 < L) o[0] = e0;
            if (l0 + 1 < L) o[1] = e1;
            if (l0 + 2 < L) o[2] = e2;
            if (l0 + 3 < L) o[3] = e3;
            if (second) {
                float* o1 = o + L;
                if (l0 + 0 < L) o1[0] = g0;
                if (l0 + 1 < L) o1[1] = g1;
                if (l0 + 2 < L) o1[2] = g2;
                if (l0 + 3 < L) o1[3] = g3;
            }
        }
        o += 2 * static_cast<size_t>(L);
    }
}

int sm_count_of_current_device() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
    return sms;
}

template <int C>
cudaError_t launch_for_channels(const float* x, const FoldArgs& fold_in, float* out, double* moments, int32_t B, int32_t L,
                                int32_t F, bool batch_stats, cudaStream_t stream) {
    FoldArgs fold = fold_in;
    if (batch_stats) {
        cudaError_t e = cudaMemsetAsync(moments, 0, sizeof(double) * moment_count(C), stream);
        if (e != cudaSuccess) return e;
        // persistent: as many CTAs as the GPU holds at once (a multiple of the SM count), each ends with one reduction
        const uint32_t qpr = (static_cast<uint32_t>(L) + 3u) / 4u;
        const uint32_t quads = static_cast<uint32_t>(B) * qpr;
        const bool vec_in = (L % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
        auto kernel = vec_in ? patch_moments_kernel<C, true> : patch_moments_kernel<C, false>;
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kMomentThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        const uint32_t want = (quads + kMomentThreads - 1) / kMomentThreads;
        const uint32_t cap = static_cast<uint32_t>(sm_count_of_current_device()) * static_cast<uint32_t>(per_sm);
        kernel<<<want < cap ? want : cap, kMomentThreads, 0, stream>>>(x, quads, qpr, static_cast<uint32_t>(L), moments);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        fold.moments = moments;
    } else {
        fold.moments = nullptr;
    }
    fold_kernel<<<1, kFoldThreads, 0, stream>>>(fold);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    const uint32_t quads_per_row = (static_cast<uint32_t>(L) + 3u) / 4u;
    const uint32_t total_quads = static_cast<uint32_t>(B) * quads_per_row;
    const bool vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const uint32_t blocks = (total_quads + kThreads - 1) / kThreads;
    // a batch too small to fill the GPU with one CTA per 1024 positions (the reference's batch of 64: 157 CTAs) is also
    // split along the filters, down to one filter pair per CTA
    const int pairs = (F + 1) / 2;
    const uint32_t fill = static_cast<uint32_t>(sm_count_of_current_device()) * 4u;
    int groups = 1;
    if (blocks < fill) {
        const uint32_t want = (fill + blocks - 1) / blocks;
        groups = want < static_cast<uint32_t>(pairs) ? static_cast<int>(want) : pairs;
    }
    const int per_group = (pairs + groups - 1) / groups;
    groups = (pairs + per_group - 1) / per_group;
    const size_t smem = static_cast<size_t>(per_group) * pair_stride(C) * sizeof(float);
    const dim3 grid(blocks, static_cast<unsigned>(groups));
    if (vec)
        apply_kernel<C, true><<<grid, kThreads, smem, stream>>>(x, fold.records, out, total_quads, quads_per_row, static_cast<uint32_t>(L), F, per_group);
    else
        apply_kernel<C, false><<<grid, kThreads, smem, stream>>>(x, fold.records, out, total_quads, quads_per_row, static_cast<uint32_t>(L), F, per_group);
    return cudaGetLastError();
}

}  // namespace

size_t first_conv_block_workspace_bytes(int32_t C, int32_t F) {
    const size_t moments = (sizeof(double) * static_cast<size_t>(3 * C + 3 * C * (3 * C + 1) / 2) + 15u) & ~static_cast<size_t>(15u);
    const size_t records = sizeof(float) * static_cast<size_t>((F + 1) / 2) * static_cast<size_t>((6 * C + 2 + 3) & ~3);
    return moments + records;
}

cudaError_t launch_first_conv_block(const FirstBlockArgs& p, cudaStream_t stream) {
    if (p.B == 0) return cudaSuccess;
    double* moments = static_cast<double*>(p.workspace);
    const size_t moments_bytes = (sizeof(double) * static_cast<size_t>(moment_count(p.C)) + 15u) & ~static_cast<size_t>(15u);
    FoldArgs fold;
    fold.weight = p.weight; fold.bias = p.bias; fold.gamma = p.gamma; fold.beta = p.beta;
    fold.running_mean = p.running_mean; fold.running_var = p.running_var;
    fold.moments = nullptr;
    fold.records = reinterpret_cast<float*>(static_cast<char*>(p.workspace) + moments_bytes);
    fold.save_mean = p.save_mean; fold.save_invstd = p.save_invstd;
    fold.n = static_cast<double>(p.B) * static_cast<double>(p.L);
    fold.eps = p.eps; fold.momentum = p.momentum; fold.F = p.F; fold.C = p.C;
    const bool batch_stats = p.batch_stats != 0;
    switch (p.C) {
        case 1: return launch_for_channels<1>(p.x, fold, p.out, moments, p.B, p.L, p.F, batch_stats, stream);
        case 2: return launch_for_channels<2>(p.x, fold, p.out, moments, p.B, p.L, p.F, batch_stats, stream);
        case 3: return launch_for_channels<3>(p.x, fold, p.out, moments, p.B, p.L, p.F, batch_stats, stream);
        case 4: return launch_for_channels<4>(p.x, fold, p.out, moments, p.B, p.L, p.F, batch_stats, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace pcgmix
