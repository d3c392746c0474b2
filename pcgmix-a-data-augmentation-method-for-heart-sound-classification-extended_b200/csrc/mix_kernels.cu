// Fused gather-mix(-magnitude-warp) kernel for PCGmix / PCGmix+ on B200 (sm_100a).
//
// What it replaces (reference = PCGmix-EXTENDED):
//   augmentations.py:969-977 / :902-914  per-cycle Python loop
//   augmentations.py:289-304             mixup_keepdur_multidim_tensors (4 slice-assign chains)
//   augmentations.py:674-683, :924-928   magnitude_warp on the host, with D2H + H2D of the batch
//   augmentations2d.py:206-221, :419-426 the spectrogram variant; :286-395 the zero boxes
//
// One pass over HBM: every output element is produced from one read of the cycle itself
// (vector loads, streaming), an optional read of the partner's sample (only where the state
// windows overlap) and one vector store.  Nothing is a contraction, so there is no tensor-core
// work here; the bound is HBM bandwidth (DESIGN.md section "Roofline").
//
// Layout: a cycle is R rows of pitch P floats, contiguous.  The partner sample of (row, t) in
// state s sits at the SAME row, column t + (f2[s]-f1[s]), i.e. at a constant flat shift per
// state, so the whole cycle is handled as one flat array of R*P floats and rows only matter
// for "which column is this" (t = flat mod P).
//
// Numerics: the blend is __fadd_rn(__fmul_rn(a,lam), __fmul_rn(b,1-lam)) — three separately
// rounded fp32 operations like the reference's tensor expression; an FMA here would break
// bit parity.  The warp factor is evaluated in fp64 and the product
// fp64(mixed)*w is rounded once to fp32, like the reference's float64 product stored to fp32.

#include "common.cuh"

namespace pcgmix {

namespace {

constexpr int kUnroll = 4;          // vectors per thread, all loads issued before first use
constexpr int kMaxThreads = 256;

template <int VEC> struct Vec;
template <> struct Vec<4> {
    float v[4];
    // cycle's own samples: read exactly once by this CTA -> do not allocate in L1
    static __device__ __forceinline__ Vec load_stream(const float* p) {
        Vec r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
        return r;
    }
    // output: written once, never re-read here -> streaming store
    __device__ __forceinline__ void store_stream(float* p) const {
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
};
template <> struct Vec<1> {
    float v[1];
    static __device__ __forceinline__ Vec load_stream(const float* p) {
        Vec r;
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r.v[0]) : "l"(p));
        return r;
    }
    __device__ __forceinline__ void store_stream(float* p) const {
        asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(v[0]) : "memory");
    }
};

// Per-cycle state windows, broadcast into registers of every thread.
struct Windows {
    int lo[4];   // start of state s in this cycle
    int n[4];    // number of blended samples in state s (min of the two durations)
    int d[4];    // flat shift to the partner's sample: f2[s] - f1[s]
};

__device__ __forceinline__ int pick4(int s, int a0, int a1, int a2, int a3) {
    int r = a0;
    r = (s == 1) ? a1 : r;
    r = (s == 2) ? a2 : r;
    r = (s == 3) ? a3 : r;
    return r;
}

// Lanes 0..4 of every warp fetch the five offsets of the cycle and of its partner; durations,
// window lengths and shifts come from warp shuffles.  No shared memory, no block barrier: each
// warp runs ahead on its own.  Invalid input (partner out of range, offsets not monotone or
// outside [0,P]) degrades to "copy the cycle" and raises a bit in *err.
__device__ __forceinline__ Windows load_windows(const MixArgs& a, int b, int& partner, unsigned& bad) {
    const int lane = threadIdx.x & 31;
    int p = __ldg(a.mix + b);
    const bool bad_partner = static_cast<unsigned>(p) >= static_cast<unsigned>(a.B);
    if (bad_partner) p = b;
    int f1 = 0, f2 = 0;
    if (lane < 5) {
        f1 = __ldg(a.frames + static_cast<size_t>(b) * a.frame_stride + lane);
        f2 = __ldg(a.frames + static_cast<size_t>(p) * a.frame_stride + lane);
    }
    const int f1n = __shfl_down_sync(kFullMask, f1, 1);
    const int f2n = __shfl_down_sync(kFullMask, f2, 1);
    const int len1 = f1n - f1;
    const int len2 = f2n - f2;
    const bool ok = (f1 >= 0) & (f2 >= 0) & (len1 >= 0) & (len2 >= 0) & (f1n <= a.P) & (f2n <= a.P);
    const unsigned bad_frames = __ballot_sync(kFullMask, (lane < 4) && !ok);
    int n = min(len1, len2);
    if (bad_frames != 0u || bad_partner) n = 0;
    const int shift = f2 - f1;
    Windows w;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        w.lo[s] = __shfl_sync(kFullMask, f1, s);
        w.n[s] = __shfl_sync(kFullMask, n, s);
        w.d[s] = __shfl_sync(kFullMask, shift, s);
    }
    partner = p;
    bad = (bad_partner ? PCGMIX_ERR_BAD_PARTNER : 0u) | (bad_frames ? PCGMIX_ERR_BAD_FRAMES : 0u);
    return w;
}

// Which piece of the spline holds integer sample t: the same bracket SciPy's PPoly uses
// (knot_pos[k] <= t < knot_pos[k+1], last piece closed on the right).
__device__ __forceinline__ int spline_piece(double td, double inv_h, int K, const double* kpos) {
    int k = min(static_cast<int>(td * inv_h), K);
    if (td < kpos[k]) {
        --k;
    } else if (k < K && td >= kpos[k + 1]) {
        ++k;
    }
    return max(k, 0);
}

__device__ __forceinline__ double spline_eval(const double* c, double dt) {
    // c[0..3] multiply dt^3, dt^2, dt^1, dt^0
    return fma(fma(fma(c[0], dt, c[1]), dt, c[2]), dt, c[3]);
}

template <int VEC, bool MAGWARP, bool BOX>
__global__ void __launch_bounds__(kMaxThreads)
mix_kernel(const __grid_constant__ MixArgs a) {
    __shared__ __align__(16) double s_coef[MAGWARP ? 2 : 1][MAGWARP ? kMaxPieces * 4 : 1];
    __shared__ double s_kpos[MAGWARP ? kMaxPieces + 1 : 1];

    const int slot = blockIdx.x / a.chunks_per_cycle;
    const int chunk = blockIdx.x - slot * a.chunks_per_cycle;
    const int b = a.order ? __ldg(a.order + slot) : slot;
    const int vbeg = chunk * a.chunk_len;
    const int vend = min(vbeg + a.chunk_len, a.nvec);
    const float* __restrict__ xb = a.x + static_cast<size_t>(b) * a.n_per_cycle;
    float* __restrict__ ob = a.out + static_cast<size_t>(b) * a.n_per_cycle;

    // 1. The cycle's own samples do not depend on the windows: get them in flight first.
    Vec<VEC> own[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
        const int v = vbeg + threadIdx.x + k * blockDim.x;
        if (v < vend) own[k] = Vec<VEC>::load_stream(xb + static_cast<size_t>(v) * VEC);
    }

    // 2. State windows of this cycle against its partner.
    int partner;
    unsigned bad;
    const Windows w = load_windows(a, b, partner, bad);
    if (bad != 0u && chunk == 0 && threadIdx.x == 0 && a.err != nullptr) atomicOr(a.err, static_cast<int>(bad));
    const float* __restrict__ xp = a.x + static_cast<size_t>(partner) * a.n_per_cycle;

    // 3. PCGmix+: cubic coefficients of the (at most two) rows this chunk touches.
    int row_first = 0;
    if constexpr (MAGWARP) {
        row_first = (vbeg * VEC) / a.P;
        const int n_knots = a.K + 2;
        const int n_coef = (a.K + 1) * 4;
        for (int i = threadIdx.x; i < 2 * n_coef; i += blockDim.x) {
            const int rr = (i >= n_coef) ? 1 : 0;
            const int ci = i - rr * n_coef;
            const int row = row_first + rr;
            if (row < a.R) {
                const double* m = a.coefmat + static_cast<size_t>(ci) * n_knots;
                const double* y = a.knots + static_cast<size_t>(b) * n_knots * a.R + row;
                double acc = 0.0;
                for (int j = 0; j < n_knots; ++j) acc = fma(__ldg(m + j), __ldg(y + static_cast<size_t>(j) * a.R), acc);
                s_coef[rr][ci] = acc;
            }
        }
        for (int i = threadIdx.x; i < n_knots; i += blockDim.x) s_kpos[i] = __ldg(a.knot_pos + i);
        __syncthreads();
    }

    int tb0 = 0, tb1 = 0;
    if constexpr (BOX) {
        tb0 = 0;
        tb1 = a.P;
        if (a.tbox != nullptr) {
            tb0 = __ldg(a.tbox + static_cast<size_t>(b) * 2);
            tb1 = __ldg(a.tbox + static_cast<size_t>(b) * 2 + 1);
        }
    }

    // 4. (row, column) of this thread's first vector; later vectors advance by a fixed step.
    int e0 = (vbeg + threadIdx.x) * VEC;
    int row = e0 / a.P;
    int t0 = e0 - row * a.P;

    bool mixed[kUnroll][VEC];
    float other[kUnroll][VEC];
    int rows[kUnroll], cols[kUnroll];

#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
        const int v = vbeg + threadIdx.x + k * blockDim.x;
        rows[k] = row;
        cols[k] = t0;
        if (v < vend) {
            const int s0 = (t0 >= w.lo[1]) + (t0 >= w.lo[2]) + (t0 >= w.lo[3]);
            const int bound = pick4(s0, w.lo[1], w.lo[2], w.lo[3], a.P);
            int src[VEC];
            if (t0 + (VEC - 1) < bound) {
                // whole vector inside one state of one row (the common case)
                const int lo = pick4(s0, w.lo[0], w.lo[1], w.lo[2], w.lo[3]);
                const int n = pick4(s0, w.n[0], w.n[1], w.n[2], w.n[3]);
                const int d = pick4(s0, w.d[0], w.d[1], w.d[2], w.d[3]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    mixed[k][e] = static_cast<unsigned>(t0 + e - lo) < static_cast<unsigned>(n);
                    src[e] = e0 + e + d;
                }
            } else {
                // vector straddles a state start or a row end: decide per sample
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    int t = t0 + e;
                    if (t >= a.P) t -= a.P;
                    const int s = (t >= w.lo[1]) + (t >= w.lo[2]) + (t >= w.lo[3]);
                    const int lo = pick4(s, w.lo[0], w.lo[1], w.lo[2], w.lo[3]);
                    const int n = pick4(s, w.n[0], w.n[1], w.n[2], w.n[3]);
                    const int d = pick4(s, w.d[0], w.d[1], w.d[2], w.d[3]);
                    mixed[k][e] = static_cast<unsigned>(t - lo) < static_cast<unsigned>(n);
                    src[e] = e0 + e + d;
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) other[k][e] = mixed[k][e] ? __ldg(xp + src[e]) : 0.0f;
        }
        e0 += blockDim.x * VEC;
        t0 += a.rstep;
        row += a.qstep;
        if (t0 >= a.P) {
            t0 -= a.P;
            ++row;
        }
    }

    // 5. Blend, warp, box, store.
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
        const int v = vbeg + threadIdx.x + k * blockDim.x;
        if (v >= vend) continue;
        Vec<VEC> res;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float mine = own[k].v[e];
            const float blended = __fadd_rn(__fmul_rn(mine, a.lam), __fmul_rn(other[k][e], a.one_minus_lam));
            res.v[e] = mixed[k][e] ? blended : mine;
        }
        if constexpr (MAGWARP) {
            const int t = cols[k];
            const double td = static_cast<double>(t);
            const int piece = spline_piece(td, a.inv_h, a.K, s_kpos);
            const bool one_piece = (t + (VEC - 1) < a.P) &&
                                   (piece == a.K || static_cast<double>(t + (VEC - 1)) < s_kpos[piece + 1]);
            if (one_piece) {
                const double* c = &s_coef[rows[k] - row_first][piece * 4];
                const double c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
                const double dt = td - s_kpos[piece];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const double de = dt + static_cast<double>(e);
                    const double wv = fma(fma(fma(c0, de, c1), de, c2), de, c3);
                    res.v[e] = static_cast<float>(static_cast<double>(res.v[e]) * wv);
                }
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    int te = t + e;
                    int re = rows[k];
                    if (te >= a.P) {
                        te -= a.P;
                        ++re;
                    }
                    const double ted = static_cast<double>(te);
                    const int pe = spline_piece(ted, a.inv_h, a.K, s_kpos);
                    const double wv = spline_eval(&s_coef[re - row_first][pe * 4], ted - s_kpos[pe]);
                    res.v[e] = static_cast<float>(static_cast<double>(res.v[e]) * wv);
                }
            }
        }
        if constexpr (BOX) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                int te = cols[k] + e;
                int re = rows[k];
                if (te >= a.P) {
                    te -= a.P;
                    ++re;
                }
                const int f = re % a.F;
                if (f >= a.h1 && f < a.h2 && te >= tb0 && te < tb1) res.v[e] = 0.0f;
            }
        }
        res.store_stream(ob + static_cast<size_t>(v) * VEC);
    }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

template <int VEC>
cudaError_t launch_vec(MixArgs a, bool magwarp, bool box, cudaStream_t stream) {
    a.nvec = a.n_per_cycle / VEC;
    int target = kMaxThreads * kUnroll;                    // vector units per CTA
    if (magwarp) target = max(1, min(target, a.P / VEC));  // a chunk may touch at most two rows
    a.chunks_per_cycle = ceil_div(a.nvec, target);
    a.chunk_len = ceil_div(a.nvec, a.chunks_per_cycle);
    int threads = ceil_div(ceil_div(a.chunk_len, kUnroll), 32) * 32;
    threads = max(32, min(threads, kMaxThreads));
    a.qstep = (threads * VEC) / a.P;
    a.rstep = (threads * VEC) % a.P;
    const long long grid = static_cast<long long>(a.B) * a.chunks_per_cycle;
    if (grid <= 0 || grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    const dim3 g(static_cast<unsigned>(grid)), t(static_cast<unsigned>(threads));
    if (magwarp) {
        mix_kernel<VEC, true, false><<<g, t, 0, stream>>>(a);
    } else if (box) {
        mix_kernel<VEC, false, true><<<g, t, 0, stream>>>(a);
    } else {
        mix_kernel<VEC, false, false><<<g, t, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_mix(const MixArgs& base, bool magwarp, bool box, cudaStream_t stream) {
    MixArgs a = base;
    a.n_per_cycle = a.R * a.P;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out)) & 15u) == 0;
    const bool wide_rows = !magwarp || a.P >= 4;   // a PCGmix+ chunk may span at most two rows
    if (aligned16 && (a.n_per_cycle % 4) == 0 && wide_rows) return launch_vec<4>(a, magwarp, box, stream);
    return launch_vec<1>(a, magwarp, box, stream);
}

}  // namespace pcgmix
