// Segmentation-index kernels: Springer 4-state annotations -> per-cycle frames, cut + zero-pad,
// duration features.  Integer work, bit-exact against the reference's notebook logic.
//
// What they replace (reference = PCGmix-EXTENDED, line numbers of the raw .ipynb JSON):
//   databuilder.ipynb:593-606  (cell 14) dense per-sample states -> transitions -> cycles
//   databuilder.ipynb:928-948  (cell 25) (position, state) table, //2 downsample, noise skip
//   databuilder.ipynb:370,:399 (cell 6)  spectrogram positions via round-half-even
//   databuilder.ipynb:627-632, :973-978, :403-411  cut + zero-pad
//   classical.py:245-283       durations, BPM, duration ratios
//
// Shape of the computation: one CTA per recording.  Transitions are found with a block-wide
// scan (warp shuffles + one shared array of warp totals) and compacted into shared memory; the
// cycle rule ("a transition into S1 that has a later S1; the four states must read
// S1,systole,S2,diastole") is a second scan over the transition list.  Cycle rows of all
// recordings are packed in (recording, time) order, which needs each recording's cycle count
// first: a count pass, a one-CTA scan, then the write pass (the int8 states are read twice;
// they are 1/16 of the bytes of the signals the cycles index, so this is noise).

#include "common.cuh"

namespace pcgmix {

namespace {

constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kBytesPerThread = 16;
constexpr int kTile = kSegThreads * kBytesPerThread;   // samples per sweep of the CTA
constexpr int kMaxTransitions = 8192;                  // per recording, held in shared memory

// Exclusive scan of one int per thread across the CTA; returns the thread's offset and the total.
__device__ __forceinline__ int block_exclusive_scan(int value, int* warp_totals, int& total) {
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    int incl = value;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_totals[wid] = incl;
    __syncthreads();
    int warp_base = 0;
    int sum = 0;
#pragma unroll
    for (int i = 0; i < kSegWarps; ++i) {
        const int wt = warp_totals[i];
        if (i < wid) warp_base += wt;
        sum += wt;
    }
    __syncthreads();   // warp_totals may be reused by the caller right away
    total = sum;
    return warp_base + incl - value;
}

__device__ __forceinline__ int block_max(int value, int* warp_totals) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) value = max(value, __shfl_xor_sync(kFullMask, value, d));
    if ((threadIdx.x & 31) == 0) warp_totals[threadIdx.x >> 5] = value;
    __syncthreads();
    int m = warp_totals[0];
#pragma unroll
    for (int i = 1; i < kSegWarps; ++i) m = max(m, warp_totals[i]);
    __syncthreads();
    return m;
}

// Position as the reference rescales it before forming offsets.
__device__ __forceinline__ int rescale(int pos, int downsample, int spec_cols, int rec_len) {
    if (spec_cols > 0) {
        // round(f * T_spec / len(y)) with Python's round-half-even on the float64 quotient
        const double q = static_cast<double>(static_cast<long long>(pos) * spec_cols) / static_cast<double>(rec_len);
        return static_cast<int>(rint(q));
    }
    return pos / downsample;
}

// Cycle rule over a transition list (generic pointers: shared or global memory).
//   write == false: only count.  write == true: emit rows at cycles + 8*(base + local index).
__device__ int emit_cycles(const int32_t* tpos, const int8_t* tcode, int n_trans, int rec, int downsample,
                           int spec_cols, int rec_len, bool noise_skip, bool write, int32_t* cycles, int base,
                           int max_cycles, int32_t* err, int* warp_totals) {
    // index of the last transition into S1
    int last = -1;
    for (int i = threadIdx.x; i < n_trans; i += blockDim.x)
        if (tcode[i] == PCGMIX_STATE_S1) last = max(last, i);
    last = block_max(last, warp_totals);

    int running = 0;
    unsigned flags = 0u;
    for (int start = 0; start < n_trans; start += blockDim.x) {
        const int i = start + threadIdx.x;
        int is_cycle = 0;
        if (i < n_trans && tcode[i] == PCGMIX_STATE_S1 && i < last) {
            // the reference looks at states[i:i+4]; i+1..i+3 exist because a later S1 exists only
            // if the pattern is intact, otherwise the comparison fails and it raises
            const int c1 = (i + 1 < n_trans) ? tcode[i + 1] : 0;
            const int c2 = (i + 2 < n_trans) ? tcode[i + 2] : 0;
            const int c3 = (i + 3 < n_trans) ? tcode[i + 3] : 0;
            const bool noisy = noise_skip && (c1 == PCGMIX_STATE_NOISE || c2 == PCGMIX_STATE_NOISE || c3 == PCGMIX_STATE_NOISE);
            if (!noisy) {
                if (c1 == PCGMIX_STATE_SYSTOLE && c2 == PCGMIX_STATE_S2 && c3 == PCGMIX_STATE_DIASTOLE && i + 4 < n_trans) {
                    is_cycle = 1;
                } else {
                    flags |= PCGMIX_ERR_BAD_PATTERN;
                }
            }
        }
        int total;
        const int off = block_exclusive_scan(is_cycle, warp_totals, total);
        if (write && is_cycle) {
            const int dst = base + running + off;
            if (dst < max_cycles) {
                int p[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) p[j] = rescale(tpos[i + j], downsample, spec_cols, rec_len);
                int4* row = reinterpret_cast<int4*>(cycles + static_cast<size_t>(dst) * 8);
                row[0] = make_int4(rec, p[0], p[4], 0);
                row[1] = make_int4(p[1] - p[0], p[2] - p[0], p[3] - p[0], p[4] - p[0]);
            } else {
                flags |= PCGMIX_ERR_OVERFLOW;
            }
        }
        running += total;
    }
    if (flags != 0u && err != nullptr) atomicOr(err, static_cast<int>(flags));
    return running;
}

// Dense states of one recording -> transition list in shared memory.
__device__ int find_transitions(const int8_t* st, int T, int32_t* tpos, int8_t* tcode, int32_t* err, int* warp_totals) {
    int running = 0;
    for (int tile = 0; tile < T; tile += kTile) {
        const int t_first = tile + threadIdx.x * kBytesPerThread;
        int8_t loc[kBytesPerThread + 1];
        // sample before this thread's slice (state[-1] := state[0] so that t = 0 is no transition)
        loc[0] = (t_first > 0 && t_first - 1 < T) ? st[t_first - 1] : ((T > 0) ? st[0] : 0);
        const bool vec_ok = (t_first + kBytesPerThread <= T) && ((reinterpret_cast<uintptr_t>(st + t_first) & 15u) == 0);
        if (vec_ok) {
            const int4 raw = *reinterpret_cast<const int4*>(st + t_first);
            const int8_t* rb = reinterpret_cast<const int8_t*>(&raw);
#pragma unroll
            for (int j = 0; j < kBytesPerThread; ++j) loc[j + 1] = rb[j];
        } else {
#pragma unroll
            for (int j = 0; j < kBytesPerThread; ++j) loc[j + 1] = (t_first + j < T) ? st[t_first + j] : loc[j];
        }
        unsigned changed = 0u;
#pragma unroll
        for (int j = 0; j < kBytesPerThread; ++j)
            if (t_first + j < T && t_first + j > 0 && loc[j + 1] != loc[j]) changed |= (1u << j);
        int total;
        int off = running + block_exclusive_scan(__popc(changed), warp_totals, total);
#pragma unroll
        for (int j = 0; j < kBytesPerThread; ++j) {
            if (changed & (1u << j)) {
                if (off < kMaxTransitions) {
                    tpos[off] = t_first + j;
                    tcode[off] = loc[j + 1];
                }
                ++off;
            }
        }
        running += total;
    }
    __syncthreads();
    if (running > kMaxTransitions) {
        if (threadIdx.x == 0 && err != nullptr) atomicOr(err, PCGMIX_ERR_OVERFLOW);
        running = kMaxTransitions;
    }
    return running;
}

template <bool WRITE>
__global__ void __launch_bounds__(kSegThreads)
segment_dense_kernel(const int8_t* __restrict__ states, int T, int downsample, int32_t* cycles, int max_cycles,
                     int32_t* cycle_count, int32_t* err) {
    __shared__ int32_t s_pos[kMaxTransitions];
    __shared__ int8_t s_code[kMaxTransitions];
    __shared__ int s_warp[kSegWarps];
    const int rec = blockIdx.x;
    const int8_t* st = states + static_cast<size_t>(rec) * T;
    const int n_trans = find_transitions(st, T, s_pos, s_code, WRITE ? nullptr : err, s_warp);
    const int base = WRITE ? cycle_count[rec] : 0;
    const int n = emit_cycles(s_pos, s_code, n_trans, rec, downsample, 0, 1, /*noise_skip=*/false, WRITE, cycles,
                              base, max_cycles, WRITE ? err : nullptr, s_warp);
    if (!WRITE && threadIdx.x == 0) cycle_count[rec] = n;
}

template <bool WRITE>
__global__ void __launch_bounds__(kSegThreads)
segment_table_kernel(const int32_t* __restrict__ positions, const int8_t* __restrict__ codes,
                     const int32_t* __restrict__ rec_offsets, int downsample, int spec_cols,
                     const int32_t* __restrict__ rec_len, int32_t* cycles, int max_cycles, int32_t* cycle_count,
                     int32_t* err) {
    __shared__ int s_warp[kSegWarps];
    const int rec = blockIdx.x;
    const int beg = rec_offsets[rec];
    const int n_trans = rec_offsets[rec + 1] - beg;
    const int len = (spec_cols > 0) ? rec_len[rec] : 1;
    const int base = WRITE ? cycle_count[rec] : 0;
    const int n = emit_cycles(positions + beg, codes + beg, n_trans, rec, downsample, spec_cols, len,
                              /*noise_skip=*/true, WRITE, cycles, base, max_cycles, WRITE ? err : nullptr, s_warp);
    if (!WRITE && threadIdx.x == 0) cycle_count[rec] = n;
}

// counts[0..R) -> exclusive offsets, total at [R].  One CTA; R is small next to the data.
__global__ void __launch_bounds__(kSegThreads) offsets_kernel(int32_t* cycle_count, int R) {
    __shared__ int s_warp[kSegWarps];
    int running = 0;
    for (int start = 0; start < R; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const int c = (i < R) ? cycle_count[i] : 0;
        int total;
        const int off = block_exclusive_scan(c, s_warp, total);
        if (i < R) cycle_count[i] = running + off;
        running += total;
    }
    if (threadIdx.x == 0) cycle_count[R] = running;
}

// out[i][c][0:L] = signal[rec][c][start:stop] cut (Python slice clipping), zero-filled to L.
// Source rows start at arbitrary sample offsets, so loads are scalar (coalesced); stores are
// float4 when the output row is 16-byte aligned.
template <int VEC>
__global__ void __launch_bounds__(256)
cut_cycles_kernel(const float* __restrict__ signal, int R, int C, int T, const int32_t* __restrict__ cycles, int n_cycles,
                  const int32_t* __restrict__ n_cycles_dev, float* __restrict__ out, int L) {
    const int n_live = n_cycles_dev ? min(n_cycles, *n_cycles_dev) : n_cycles;
    const long long n_rows = static_cast<long long>(n_live) * C;
    const int vec_per_row = L / VEC;
    for (long long r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int i = static_cast<int>(r / C);
        const int c = static_cast<int>(r - static_cast<long long>(i) * C);
        const int4 head = *reinterpret_cast<const int4*>(cycles + static_cast<size_t>(i) * 8);
        const int rec = head.x;
        const int start = min(max(head.y, 0), T);
        const int stop = min(max(head.z, start), T);
        // a row naming a recording that does not exist yields an all-zero cycle instead of a wild read
        const int n = static_cast<unsigned>(rec) < static_cast<unsigned>(R) ? stop - start : 0;
        const float* src = signal + (static_cast<size_t>(n > 0 ? rec : 0) * C + c) * T + start;
        float* dst = out + static_cast<size_t>(r) * L;
        for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) {
            float val[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int t = v * VEC + e;
                val[e] = (t < n) ? __ldg(src + t) : 0.0f;
            }
            if constexpr (VEC == 4) {
                __stcs(reinterpret_cast<float4*>(dst) + v, make_float4(val[0], val[1], val[2], val[3]));
            } else {
                __stcs(dst + v, val[0]);
            }
        }
    }
}

// Python's round(x, 4): the double nearest to x rounded (half-even, on the EXACT value of x)
// to four decimals.  x*1e4 is formed with its exact residual (one FMA) so ties are decided on
// the true product, not on its rounded image.
__device__ __forceinline__ double round4(double x) {
    if (!isfinite(x)) return x;
    const double p = x * 1.0e4;
    const double resid = fma(x, 1.0e4, -p);
    double r = rint(p);
    const double d = (p - r) + resid;
    if (d > 0.5) r += 1.0;
    else if (d < -0.5) r -= 1.0;
    return r / 1.0e4;
}

__global__ void __launch_bounds__(256)
duration_features_kernel(const int32_t* __restrict__ frames, int stride, int n, int fs, double* __restrict__ feat,
                         int32_t* err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* f = frames + static_cast<size_t>(i) * stride;
    const double dfs = static_cast<double>(fs);
    auto ms = [dfs](int len) { return trunc(static_cast<double>(static_cast<long long>(len) * 1000) / dfs); };
    const double rr = ms(f[4]);
    const double s1 = ms(f[1]);
    const double sy = ms(f[2] - f[1]);
    const double s2 = ms(f[3] - f[2]);
    const double di = ms(f[4] - f[3]);
    const bool zero = (rr == 0.0) || (s1 == 0.0) || (s2 == 0.0) || (di == 0.0);
    if (zero && err != nullptr) atomicOr(err, PCGMIX_ERR_ZERO_DIVISION);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    auto ratio = [nan](double a, double b) { return (b == 0.0) ? nan : round4(a / b); };
    double* o = feat + static_cast<size_t>(i) * 14;
    o[0] = rr;
    o[1] = ratio(60000.0, rr);
    o[2] = s1;
    o[3] = sy;
    o[4] = s2;
    o[5] = di;
    o[6] = ratio(s1, s2);
    o[7] = ratio(sy, di);
    o[8] = ratio(s1, rr);
    o[9] = ratio(sy, rr);
    o[10] = ratio(s2, rr);
    o[11] = ratio(di, rr);
    o[12] = ratio(sy, s1);
    o[13] = ratio(di, s2);
}

}  // namespace

cudaError_t launch_segment_dense(const int8_t* states, int32_t R, int32_t T, int32_t downsample, int32_t* cycles,
                                 int32_t max_cycles, int32_t* cycle_count, int32_t* err, cudaStream_t stream) {
    if (R == 0) {
        return cudaMemsetAsync(cycle_count, 0, sizeof(int32_t), stream);
    }
    segment_dense_kernel<false><<<R, kSegThreads, 0, stream>>>(states, T, downsample, cycles, max_cycles, cycle_count, err);
    offsets_kernel<<<1, kSegThreads, 0, stream>>>(cycle_count, R);
    segment_dense_kernel<true><<<R, kSegThreads, 0, stream>>>(states, T, downsample, cycles, max_cycles, cycle_count, err);
    return cudaGetLastError();
}

cudaError_t launch_segment_table(const int32_t* positions, const int8_t* codes, const int32_t* rec_offsets, int32_t R,
                                 int32_t downsample, int32_t spec_cols, const int32_t* rec_len, int32_t* cycles,
                                 int32_t max_cycles, int32_t* cycle_count, int32_t* err, cudaStream_t stream) {
    if (R == 0) {
        return cudaMemsetAsync(cycle_count, 0, sizeof(int32_t), stream);
    }
    segment_table_kernel<false><<<R, kSegThreads, 0, stream>>>(positions, codes, rec_offsets, downsample, spec_cols,
                                                               rec_len, cycles, max_cycles, cycle_count, err);
    offsets_kernel<<<1, kSegThreads, 0, stream>>>(cycle_count, R);
    segment_table_kernel<true><<<R, kSegThreads, 0, stream>>>(positions, codes, rec_offsets, downsample, spec_cols,
                                                              rec_len, cycles, max_cycles, cycle_count, err);
    return cudaGetLastError();
}

cudaError_t launch_cut_cycles(const float* signal, int32_t R, int32_t C, int32_t T, const int32_t* cycles,
                              int32_t n_cycles, const int32_t* n_cycles_dev, float* out, int32_t L,
                              cudaStream_t stream) {
    if (n_cycles == 0 || C == 0 || L == 0) return cudaSuccess;
    const long long rows = static_cast<long long>(n_cycles) * C;
    const unsigned grid = static_cast<unsigned>(rows < (1LL << 20) ? rows : (1LL << 20));
    const bool vec4 = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    if (vec4) {
        cut_cycles_kernel<4><<<grid, 256, 0, stream>>>(signal, R, C, T, cycles, n_cycles, n_cycles_dev, out, L);
    } else {
        cut_cycles_kernel<1><<<grid, 256, 0, stream>>>(signal, R, C, T, cycles, n_cycles, n_cycles_dev, out, L);
    }
    return cudaGetLastError();
}

cudaError_t launch_duration_features(const int32_t* frames, int32_t frame_stride, int32_t n, int32_t fs,
                                     double* features, int32_t* err, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    duration_features_kernel<<<(n + 255) / 256, 256, 0, stream>>>(frames, frame_stride, n, fs, features, err);
    return cudaGetLastError();
}

}  // namespace pcgmix
