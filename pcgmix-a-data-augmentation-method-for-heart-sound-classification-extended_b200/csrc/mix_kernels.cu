// Fused gather-mix(-magnitude-warp) kernel for PCGmix / PCGmix+ on B200 (sm_100a).
//
// What it replaces (reference = PCGmix-EXTENDED):
//   augmentations.py:969-977 / :902-914  per-cycle Python loop
//   augmentations.py:289-304             mixup_keepdur_multidim_tensors (4 slice-assign chains)
//   augmentations.py:674-683, :924-928   magnitude_warp on the host, with D2H + H2D of the batch
//   augmentations2d.py:206-221, :419-426 the spectrogram variant; :286-395 the zero boxes
//
// One pass over HBM: every output element is produced from one read of the cycle itself
// (128-bit streaming loads), an optional read of the partner's sample (only where the state
// windows overlap) and one 128-bit streaming store.  Nothing is a contraction, so there is no
// tensor-core work here; the bound is HBM bandwidth ("Roofline").
//
// Layout: a cycle is R rows of pitch P floats, contiguous.  The partner sample of (row, t) in
// state s sits in the SAME row at column t + (f2[s]-f1[s]), i.e. at a constant flat shift per
// state.  Two work decompositions share one kernel body:
//   ROWS  a CTA owns a slice of ONE row (P % VEC == 0): column = slice start + vector index,
//         no division anywhere; grid = (cycle slot, slice, row)
//   FLAT  a CTA owns a slice of the cycle's flat R*P array (rows not vector-aligned, e.g.
//         T = 250 spectrogram columns): (row, column) advance incrementally per vector
//
// Instruction budget matters as much as bytes here (at 6.5 TB/s an SM has ~55 issue slots per
// 128-bit vector per warp... the first version of this kernel spent 290 on it): the per-cycle state
// table lives in shared memory as four int4 {start, blended length, partner shift, next start},
// one LDS.128 per vector after three compares; the block size is a template parameter so every
// global address is "base + immediate".
//
// Numerics: the blend is __fadd_rn(__fmul_rn(a,lam), __fmul_rn(b,1-lam)) — three separately
// rounded fp32 operations like the reference's tensor expression; an FMA here would break
// bit parity.  The warp factor is a float64 Horner evaluation and the
// product fp64(mixed)*w is rounded once to fp32, like the reference's float64 product stored
// into a float32 array.

#include "common.cuh"

namespace pcgmix {

namespace {

constexpr int kUnroll = 4;          // vectors per thread, all loads issued before first use
constexpr int kFlatMaxPitch = 2048; // FLAT slices are used for rows shorter than this
constexpr int kNoBlend = -2147483647 - 1;

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

// Exact int -> double for 0 <= i < 2^31 without the (slow) I2F.F64 conversion pipe:
// 2^52 + i is exactly representable; subtracting 2^52 leaves i.
__device__ __forceinline__ double int_to_double(int i) {
    return __hiloint2double(0x43300000, i) - 4503599627370496.0;
}

// Warp 0 builds the cycle's state table: lanes 0..4 fetch the five offsets of the cycle and of
// its partner, durations / blended lengths / shifts come from warp shuffles (the per-cycle
// prefix sums are already in the offsets).  Invalid input (partner out of range, offsets not
// monotone, or clamped windows of unequal width: see pair_window) degrades to "copy the cycle" and raises a bit in *err.
__device__ __forceinline__ void build_windows(const MixArgs& a, int b, int4* s_win, int* s_partner, bool report) {
    const int lane = threadIdx.x;
    int p = __ldg(a.mix + b);
    const bool bad_partner = static_cast<unsigned>(p) >= static_cast<unsigned>(a.B);
    if (bad_partner) p = b;
    int f1 = 0, f2 = 0, n = 0, start = 0, shift = 0;
    bool ok;
    if (a.windows != nullptr) {
        // explicit windows (the reference's '(rand)' displacement): {start, length, shift} per state
        if (lane < 4) {
            const int32_t* w = a.windows + (static_cast<size_t>(b) * 4 + lane) * 3;
            f1 = __ldg(w);
            n = __ldg(w + 1);
            f2 = f1 + __ldg(w + 2);
        }
        const int next_lo = __shfl_down_sync(kFullMask, f1, 1);
        const int limit = (lane < 3) ? next_lo : a.P;
        ok = (f1 >= 0) & (n >= 0) & (f1 + n <= limit) & (f2 >= 0) & (f2 + n <= a.P);
        start = f1;
        shift = f2 - f1;
    } else {
        if (lane < 5) {
            f1 = __ldg(a.frames + static_cast<size_t>(b) * a.frame_stride + lane);
            f2 = __ldg(a.frames + static_cast<size_t>(p) * a.frame_stride + lane);
        }
        const int f1n_ = __shfl_down_sync(kFullMask, f1, 1);
        const int f2n_ = __shfl_down_sync(kFullMask, f2, 1);
        ok = pair_window(f1, f1n_, f2, f2n_, a.P, start, n, shift);
    }
    const int f1n = __shfl_down_sync(kFullMask, f1, 1);
    const unsigned bad_frames = __ballot_sync(kFullMask, (lane < 4) && !ok);
    if (bad_frames != 0u || bad_partner) n = 0;
    // "next start": the column where the state after s begins; P closes the last one
    const int next = (lane < 3) ? min(f1n, a.P) : a.P;
    if (lane < 4) s_win[lane] = make_int4(start, n, shift, next);
    if (lane == 0) {
        *s_partner = p;
        const unsigned bad = (bad_partner ? PCGMIX_ERR_BAD_PARTNER : 0u) | (bad_frames ? PCGMIX_ERR_BAD_FRAMES : 0u);
        if (bad != 0u && report && a.err != nullptr) atomicOr(a.err, static_cast<int>(bad));
    }
}

template <int VEC, int T, bool ROWS, bool MAGWARP, bool BOX>
__global__ void __launch_bounds__(T, (MAGWARP ? 1024 : 1280) / T)
mix_kernel(const __grid_constant__ MixArgs a) {
    static_assert(!MAGWARP || ROWS, "the fused warp needs row-aligned slices");
    __shared__ int4 s_win[4];
    __shared__ int s_partner;
    __shared__ __align__(16) double s_coef[MAGWARP ? kMaxPieces * 4 : 2];
    __shared__ double s_kpos[MAGWARP ? kMaxPieces + 1 : 1];
    __shared__ int s_kint[MAGWARP ? kMaxPieces + 1 : 1];
    // FLAT slices cover many short rows (spectrograms), and nearly every warp holds a vector that
    // straddles a state start or a row end.  One column table per CTA (shift to the partner's sample,
    // or kNoBlend) replaces the per-vector state logic: every column is classified once, used R times.
    __shared__ int s_lut[ROWS ? 1 : kFlatMaxPitch];

    const int slot = blockIdx.x;
    const int b = cycle_of_slot(a, slot);
    // first vector of this CTA, in vector units from the start of the cycle, and its column
    int vbeg, vend, row = 0, col0;
    if constexpr (ROWS) {
        row = blockIdx.z;
        const int seg_beg = blockIdx.y * a.chunk_len;                 // within the row
        const int seg_end = min(seg_beg + a.chunk_len, a.P / VEC);
        const int row_base = row * (a.P / VEC);
        vbeg = row_base + seg_beg;
        vend = row_base + seg_end;
        col0 = (seg_beg + threadIdx.x) * VEC;
    } else {
        vbeg = blockIdx.y * a.chunk_len;
        vend = min(vbeg + a.chunk_len, a.nvec);
        const int e = (vbeg + threadIdx.x) * VEC;
        row = e / a.P;
        col0 = e - row * a.P;
    }
    const size_t cyc = static_cast<size_t>(b) * a.n_per_cycle;
    const int v0 = vbeg + threadIdx.x;
    const float* __restrict__ own_ptr = a.x + cyc + static_cast<size_t>(v0) * VEC;
    float* __restrict__ out_ptr = a.out + cyc + static_cast<size_t>(v0) * VEC;
    const int left = vend - v0;                                        // vector k is live iff k*T < left

    // 1. The cycle's own samples do not depend on the windows: get them in flight first.
    Vec<VEC> own[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k)
        if (k * T < left) own[k] = Vec<VEC>::load_stream(own_ptr + k * T * VEC);

    // 2. State table of this cycle against its partner; spline coefficients of this row.
    if (threadIdx.x < 32) build_windows(a, b, s_win, &s_partner, ROWS ? (blockIdx.y == 0 && blockIdx.z == 0) : blockIdx.y == 0);
    if constexpr (MAGWARP) {
        const int n_knots = a.K + 2;
        const int n_coef = (a.K + 1) * 4;
        // the last warps of the CTA build the coefficients while warp 0 chases partner -> offsets
        for (int i = (T - 1) - threadIdx.x; i < n_coef; i += T) {
            const double* m = a.coefmat + static_cast<size_t>(i) * n_knots;
            const double* y = a.knots + static_cast<size_t>(b) * n_knots * a.R + row;
            double acc = 0.0;
            for (int j = 0; j < n_knots; ++j) acc = fma(__ldg(m + j), __ldg(y + static_cast<size_t>(j) * a.R), acc);
            s_coef[i] = acc;
        }
        for (int i = (T - 33) - static_cast<int>(threadIdx.x); i >= 0 && i < n_knots; i += T) {
            const double kp = __ldg(a.knot_pos + i);
            s_kpos[i] = kp;
            // first integer sample that belongs to the piece starting at this knot; the bracket
            // kpos[k] <= t < kpos[k+1] of SciPy's PPoly becomes kint[k] <= t < kint[k+1]
            s_kint[i] = (i == n_knots - 1) ? 0x7fffffff : static_cast<int>(ceil(kp));
        }
    }
    __syncthreads();
    const int lo1 = s_win[1].x, lo2 = s_win[2].x, lo3 = s_win[3].x;
    if constexpr (!ROWS) {
        for (int t = threadIdx.x; t < a.P; t += T) {
            const int s = (t >= lo1) + (t >= lo2) + (t >= lo3);
            const int4 w = s_win[s];
            s_lut[t] = (static_cast<unsigned>(t - w.x) < static_cast<unsigned>(w.y)) ? w.z : kNoBlend;
        }
        __syncthreads();
    }
    const float* __restrict__ par_ptr = a.x + static_cast<size_t>(s_partner) * a.n_per_cycle + static_cast<size_t>(v0) * VEC;

    int tb0 = 0, tb1 = 0;
    if constexpr (BOX) {
        tb1 = a.P;
        if (a.tbox != nullptr) {
            tb0 = __ldg(a.tbox + static_cast<size_t>(b) * 2);
            tb1 = __ldg(a.tbox + static_cast<size_t>(b) * 2 + 1);
        }
    }

    // 3. Per vector: which state, how many leading samples blend, where the partner's are.
    int live[kUnroll];          // bit e set: sample e of vector k is blended
    float other[kUnroll][VEC];
    int cols[kUnroll], rows[kUnroll];
    int col = col0;
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
        cols[k] = col;
        rows[k] = row;
        live[k] = 0;
        if constexpr (!ROWS) {
            if (k * T < left) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    int t = col + e;
                    if (t >= a.P) t -= a.P;
                    const int d = s_lut[t];
                    other[k][e] = 0.0f;
                    if (d != kNoBlend) {
                        other[k][e] = __ldg(par_ptr + k * T * VEC + e + d);
                        live[k] |= 1 << e;
                    }
                }
            }
        } else if (k * T < left) {
            const int s = (col >= lo1) + (col >= lo2) + (col >= lo3);
            const int4 w = s_win[s];                                   // {start, blended, shift, next start}
            const int ahead = col - w.x;
            if (__builtin_expect(ahead >= 0 && col + (VEC - 1) < w.w, 1)) {
                // whole vector inside one state of one row: samples e < m blend, partner is contiguous
                const int m = w.y - ahead;
                const float* src = par_ptr + k * T * VEC + w.z;
#pragma unroll
                for (int e = 0; e < VEC; ++e) other[k][e] = (e < m) ? __ldg(src + e) : 0.0f;
                live[k] = (1 << min(max(m, 0), VEC)) - 1;
            } else {
                // vector straddles a state start or a row end (or precedes f[0] > 0): per sample
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int t = col + e;
                    const int se = (t >= lo1) + (t >= lo2) + (t >= lo3);
                    const int4 we = s_win[se];
                    other[k][e] = 0.0f;
                    if (static_cast<unsigned>(t - we.x) < static_cast<unsigned>(we.y)) {
                        other[k][e] = __ldg(par_ptr + k * T * VEC + e + we.z);
                        live[k] |= 1 << e;
                    }
                }
            }
        }
        if constexpr (ROWS) {
            col += T * VEC;
        } else {
            col += a.rstep;
            row += a.qstep;
            if (col >= a.P) {
                col -= a.P;
                ++row;
            }
        }
    }

    // 4. Blend, warp, box, store.
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
        if (!(k * T < left)) continue;
        Vec<VEC> res = own[k];
        if (live[k] != 0) {                                            // most vectors lie outside every window
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float blended = __fadd_rn(__fmul_rn(res.v[e], a.lam), __fmul_rn(other[k][e], a.one_minus_lam));
                res.v[e] = (live[k] >> e & 1) ? blended : res.v[e];
            }
        }
        if constexpr (MAGWARP) {
            const int t = cols[k];
            // piece from an integer reciprocal guess (never above the true piece), then one fix-up
            int piece = min(static_cast<int>(__umulhi(static_cast<unsigned>(t), a.piece_magic)), a.K);
            while (t >= s_kint[piece + 1]) ++piece;                    // at most one step unless rows are tiny
            if (__builtin_expect(t + (VEC - 1) < s_kint[piece + 1], 1)) {
                const double2 c01 = *reinterpret_cast<const double2*>(&s_coef[piece * 4]);
                const double2 c23 = *reinterpret_cast<const double2*>(&s_coef[piece * 4 + 2]);
                const double dt = int_to_double(t) - s_kpos[piece];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const double de = dt + static_cast<double>(e);
                    const double wv = fma(fma(fma(c01.x, de, c01.y), de, c23.x), de, c23.y);
                    res.v[e] = static_cast<float>(static_cast<double>(res.v[e]) * wv);
                }
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int te = t + e;
                    int pe = min(static_cast<int>(__umulhi(static_cast<unsigned>(te), a.piece_magic)), a.K);
                    while (te >= s_kint[pe + 1]) ++pe;
                    const double de = int_to_double(te) - s_kpos[pe];
                    const double* c = &s_coef[pe * 4];
                    const double wv = fma(fma(fma(c[0], de, c[1]), de, c[2]), de, c[3]);
                    res.v[e] = static_cast<float>(static_cast<double>(res.v[e]) * wv);
                }
            }
        }
        if constexpr (BOX) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                int te = cols[k] + e;
                int re = rows[k];
                if (!ROWS && te >= a.P) {
                    te -= a.P;
                    ++re;
                }
                const int f = re % a.F;
                if (f >= a.h1 && f < a.h2 && te >= tb0 && te < tb1) res.v[e] = 0.0f;
            }
        }
        res.store_stream(out_ptr + k * T * VEC);
    }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

template <int VEC, int T, bool ROWS>
cudaError_t launch_t(const MixArgs& a, dim3 grid, bool magwarp, bool box, cudaStream_t stream) {
    if constexpr (ROWS) {
        if (magwarp) {
            mix_kernel<VEC, T, true, true, false><<<grid, T, 0, stream>>>(a);
            return cudaGetLastError();
        }
    }
    if (box) {
        mix_kernel<VEC, T, ROWS, false, true><<<grid, T, 0, stream>>>(a);
    } else {
        mix_kernel<VEC, T, ROWS, false, false><<<grid, T, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

// Pick the block size (a template parameter) that wastes the fewest thread slots on a slice of
// `len` vectors; a slice holds at most T*kUnroll vectors.
template <int VEC, bool ROWS>
cudaError_t launch_pick(MixArgs a, int units, bool magwarp, bool box, cudaStream_t stream) {
    // units = vectors per row (ROWS) or per cycle (FLAT)
    const int cap = 256 * kUnroll;
    const int slices = ceil_div(units, cap);
    a.chunk_len = ceil_div(units, slices);
    a.chunks_per_cycle = slices;
    const int need = ceil_div(a.chunk_len, kUnroll);            // threads that have work
    int T = 256;
    if (VEC == 4) T = need <= 128 ? 128 : need <= 160 ? 160 : need <= 192 ? 192 : need <= 224 ? 224 : 256;
    a.qstep = (T * VEC) / a.P;
    a.rstep = (T * VEC) % a.P;
    if (a.B > 2147483647 || slices > 65535 || (ROWS && a.R > 65535)) return cudaErrorInvalidConfiguration;
    const dim3 grid(static_cast<unsigned>(a.B), static_cast<unsigned>(slices), ROWS ? static_cast<unsigned>(a.R) : 1u);
    if constexpr (VEC == 4) {
        switch (T) {
            case 128: return launch_t<4, 128, ROWS>(a, grid, magwarp, box, stream);
            case 160: return launch_t<4, 160, ROWS>(a, grid, magwarp, box, stream);
            case 192: return launch_t<4, 192, ROWS>(a, grid, magwarp, box, stream);
            case 224: return launch_t<4, 224, ROWS>(a, grid, magwarp, box, stream);
            default: return launch_t<4, 256, ROWS>(a, grid, magwarp, box, stream);
        }
    } else {
        return launch_t<1, 256, ROWS>(a, grid, magwarp, box, stream);
    }
}

}  // namespace

cudaError_t launch_mix(const MixArgs& base, bool magwarp, bool box, cudaStream_t stream) {
    MixArgs a = base;
    a.n_per_cycle = a.R * a.P;
    if (magwarp) {
        // floor(2^32 * (K+1)/(P-1)) rounded down: umulhi(t, magic) never exceeds the true piece
        const double ratio = static_cast<double>(a.K + 1) / static_cast<double>(a.P - 1);
        const double scaled = ratio * 4294967296.0 * (1.0 - 1e-9);
        a.piece_magic = scaled >= 4294967295.0 ? 4294967295u : static_cast<unsigned>(scaled);
    }
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.out)) & 15u) == 0;
    const bool rows_fit_grid = a.R <= 65535;
    const bool vec_rows = aligned16 && (a.P % 4) == 0 && rows_fit_grid;
    // (a 4-wide vector of a FLAT slice may touch two rows, not more: rows shorter than a vector go scalar)
    const bool vec_flat = aligned16 && (a.n_per_cycle % 4) == 0 && !magwarp && a.P <= kFlatMaxPitch && a.P >= 4;
    // One CTA per row slice only pays off for long rows; spectrogram rows (128..250 columns) are
    // handled as slices of the cycle's flat plane, 1024 vectors per CTA.
    const bool long_rows = a.P > kFlatMaxPitch;
    if (vec_rows && (long_rows || !vec_flat)) {
        a.nvec = a.n_per_cycle / 4;
        return launch_pick<4, true>(a, a.P / 4, magwarp, box, stream);
    }
    if (vec_flat) {
        a.nvec = a.n_per_cycle / 4;
        return launch_pick<4, false>(a, a.nvec, magwarp, box, stream);
    }
    a.nvec = a.n_per_cycle;
    if (rows_fit_grid && (long_rows || magwarp)) return launch_pick<1, true>(a, a.P, magwarp, box, stream);
    if (magwarp) return cudaErrorInvalidConfiguration;
    if (a.P > kFlatMaxPitch) return cudaErrorInvalidConfiguration;
    return launch_pick<1, false>(a, a.nvec, magwarp, box, stream);
}

}  // namespace pcgmix
