// PCGmix / PCGmix+ straight from recordings that stay on the device (sm_100a).
//
// What it replaces (reference = PCGmix-EXTENDED), fused into ONE pass:
//   databuilder.ipynb:627-632, :973-978   seg_y = y_hat[start:stop]; seg_y.resize(L)   (cut + zero-pad)
//   dataloader_physionet.py:43-48         stacking the padded cycles into the (n, C, L) array a batch is drawn from
//   train_model.py:499                    data.to(device) of the padded batch
//   augmentations.py:969-977 / :902-914   the per-cycle loop around mixup_keepdur_multidim_tensors (:289-304)
//   augmentations.py:674-683, :924-928    magnitude_warp with its D2H + H2D round trip
//
// The padded (n, C, L) array never exists: batch slot i is row sel[i] of the cycle table
// {recording, abs_start, abs_stop, f0..f4} (segment kernels), its samples are read from
// signal[recording][c][abs_start + t] for t < min(abs_stop-abs_start, L) and are zero beyond —
// exactly what cut + resize would have stored — and the partner's samples are read from the
// partner's recording the same way.  Result: bit-identical to pcgmix_cut_cycles followed by
// pcgmix_mix1d / pcgmix_mix1d_magwarp (tests/test_resident_gpu.py), at
//   4*C*(len1 + M + L) bytes per cycle   instead of   4*C*(len + L)  +  4*C*(2L + M)
// (M = samples blended with the partner): with PhysioNet-shaped cycles (mean 1.1 k samples in
// L = 2500) that is 4.5 k instead of 9.5 k floats per row.
//
// Two kernels serve the entry point (capi.cu picks):
//   * with caller-provided scratch, L % 4 == 0, L >= 1024 and aligned tensors: resolve_kernel (below) turns
//     every batch slot into an 8-int record, then the RESIDENT variant of the persistent TMA-pipelined
//     kernel (mix_pipeline.cu) does the work — 0.75 / 0.59 of the measured HBM peak for PCGmix / PCGmix+;
//   * otherwise the direct-load kernel in this file (any L, any alignment of `out`).
//
// Direct-load kernel: a recording row starts anywhere, so the cycle's own samples are not 16-byte aligned: every
// thread loads the ALIGNED 128-bit vector that covers its columns and takes the missing head of
// the next vector from its neighbour lane with warp shuffles (lane 31 loads it itself); the
// misalignment (0..3 floats) is uniform per CTA.  Output rows are aligned: 128-bit streaming stores.
//
// Numerics as in mix_kernels.cu: three separately rounded fp32 operations for the blend, float64
// Horner for the warp factor, one rounding of fp64(sample)*w to fp32.  Padding is multiplied by the
// warp factor like every other sample (0*w keeps the sign/NaN behaviour of the reference product).

#include "common.cuh"

namespace pcgmix {

namespace {

constexpr int kUnroll = 4;   // vectors per thread, loads issued before first use

__device__ __forceinline__ double int_to_double(int i) {
    return __hiloint2double(0x43300000, i) - 4503599627370496.0;
}

__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void st_stream4(float* p, const float* v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// 128-bit load of signal[idx .. idx+3] (idx a multiple of 4, signal 16-byte aligned); the last,
// partial vector of the whole tensor is read element by element.
__device__ __forceinline__ float4 load_aligned(const float* signal, long long idx, long long n_sig) {
    if (idx + 4 <= n_sig) return ld_stream4(signal + idx);
    float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (idx < n_sig) r.x = __ldg(signal + idx);
    if (idx + 1 < n_sig) r.y = __ldg(signal + idx + 1);
    if (idx + 2 < n_sig) r.z = __ldg(signal + idx + 2);
    return r;
}

// Where a table row's samples of channel `row` live: element index of column 0 inside `signal`,
// and how many columns hold samples (the rest of the L columns is padding).  Same clamping as
// cut_cycles_kernel (segment_kernels.cu).
struct Source {
    long long first;
    int n;
};

__device__ __forceinline__ Source source_of(const MixArgs& a, int table_row, int row) {
    Source s{0, 0};
    if (static_cast<unsigned>(table_row) >= static_cast<unsigned>(a.n_table)) return s;
    const int4 head = __ldg(reinterpret_cast<const int4*>(a.cycles) + static_cast<size_t>(table_row) * 2);
    if (static_cast<unsigned>(head.x) >= static_cast<unsigned>(a.n_rec)) return s;
    const int start = min(max(head.y, 0), a.T_sig);
    const int stop = min(max(head.z, start), a.T_sig);
    s.first = (static_cast<long long>(head.x) * a.R + row) * a.T_sig + start;
    s.n = min(stop - start, a.P);
    return s;
}

// Warp 0: state table of batch slot b against its partner, offsets taken from the cycle table.
// Anything out of range degrades to "copy the cycle" and raises a bit in *err.
__device__ __forceinline__ void build_windows_resident(const MixArgs& a, int b, int my_row, int4* s_win, int* s_partner_row,
                                                       bool report) {
    const int lane = threadIdx.x;
    int p = __ldg(a.mix + b);
    bool bad_index = static_cast<unsigned>(p) >= static_cast<unsigned>(a.B);
    if (bad_index) p = b;
    int partner_row = a.sel ? __ldg(a.sel + p) : p;
    if (static_cast<unsigned>(my_row) >= static_cast<unsigned>(a.n_table) ||
        static_cast<unsigned>(partner_row) >= static_cast<unsigned>(a.n_table)) {
        bad_index = true;
        partner_row = -1;
    }
    int f1 = 0, f2 = 0;
    bool rec_ok = true;
    if (!bad_index) {
        if (lane < 5) {
            f1 = __ldg(a.cycles + static_cast<size_t>(my_row) * 8 + 3 + lane);
            f2 = __ldg(a.cycles + static_cast<size_t>(partner_row) * 8 + 3 + lane);
        }
        if (lane == 5) rec_ok = static_cast<unsigned>(__ldg(a.cycles + static_cast<size_t>(my_row) * 8)) < static_cast<unsigned>(a.n_rec);
        if (lane == 6) rec_ok = static_cast<unsigned>(__ldg(a.cycles + static_cast<size_t>(partner_row) * 8)) < static_cast<unsigned>(a.n_rec);
    }
    if (__ballot_sync(kFullMask, !rec_ok) != 0u) bad_index = true;
    const int f1n = __shfl_down_sync(kFullMask, f1, 1);
    const int f2n = __shfl_down_sync(kFullMask, f2, 1);
    int start = 0, n = 0, shift = 0;
    const bool ok = pair_window(f1, f1n, f2, f2n, a.P, start, n, shift);
    const unsigned bad_frames = __ballot_sync(kFullMask, (lane < 4) && !ok);
    if (bad_frames != 0u || bad_index) n = 0;
    const int next = (lane < 3) ? min(f1n, a.P) : a.P;
    if (lane < 4) s_win[lane] = make_int4(start, n, shift, next);
    if (lane == 0) {
        *s_partner_row = partner_row;
        const unsigned bad = (bad_index ? PCGMIX_ERR_BAD_PARTNER : 0u) | (bad_frames ? PCGMIX_ERR_BAD_FRAMES : 0u);
        if (bad != 0u && report && a.err != nullptr) atomicOr(a.err, static_cast<int>(bad));
    }
}

template <int VEC, int T, bool MAGWARP>
__global__ void __launch_bounds__(T, (MAGWARP ? 1024 : 1280) / T)
mix_resident_kernel(const __grid_constant__ MixArgs a) {
    __shared__ int4 s_win[4];
    __shared__ int s_partner_row;
    __shared__ __align__(16) double s_coef[MAGWARP ? kMaxPieces * 4 : 2];
    __shared__ double s_kpos[MAGWARP ? kMaxPieces + 1 : 1];
    __shared__ int s_kint[MAGWARP ? kMaxPieces + 1 : 1];

    const int slot = blockIdx.x;
    const int b = cycle_of_slot(a, slot);
    const int row = blockIdx.z;
    const int seg_beg = blockIdx.y * a.chunk_len;                     // vector units within the row
    const int seg_end = min(seg_beg + a.chunk_len, a.P / VEC);
    const int v0 = seg_beg + threadIdx.x;
    const int left = seg_end - v0;                                     // vector k is live iff k*T < left
    const int col0 = v0 * VEC;
    const int lane = threadIdx.x & 31;

    // 1. The cycle's own samples: resolve slot -> table row -> recording and get the loads in flight.
    const int my_row = a.sel ? __ldg(a.sel + b) : b;
    const Source own = source_of(a, my_row, row);
    const int shift = VEC == 4 ? static_cast<int>(own.first & 3) : 0;  // floats between the aligned vector and column 0
    const long long own_al = own.first - shift;
    float4 raw[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
        const int col = col0 + k * T * VEC;
        raw[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (k * T < left && col < own.n + shift) {
            if constexpr (VEC == 4) {
                raw[k] = load_aligned(a.signal, own_al + col, a.n_sig);
            } else {
                raw[k].x = __ldg(a.signal + own.first + col);
            }
        }
    }

    // 2. State table against the partner; spline coefficients of this row.
    if (threadIdx.x < 32) build_windows_resident(a, b, my_row, s_win, &s_partner_row, blockIdx.y == 0 && blockIdx.z == 0);
    if constexpr (MAGWARP) {
        const int n_knots = a.K + 2;
        const int n_coef = (a.K + 1) * 4;
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
            s_kint[i] = (i == n_knots - 1) ? 0x7fffffff : static_cast<int>(ceil(kp));
        }
    }
    __syncthreads();
    const int lo1 = s_win[1].x, lo2 = s_win[2].x, lo3 = s_win[3].x;
    const Source par = source_of(a, s_partner_row, row);
    const float* __restrict__ par_ptr = a.signal + par.first;

    // 3. Own vectors: shift the aligned stream into place (head of the next vector from lane+1).
    float own_v[kUnroll][VEC];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
        const int col = col0 + k * T * VEC;
        if constexpr (VEC == 4) {
            float s[7] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w, 0.0f, 0.0f, 0.0f};
            if (shift != 0) {                                          // uniform over the CTA
                s[4] = __shfl_down_sync(kFullMask, raw[k].x, 1);
                s[5] = __shfl_down_sync(kFullMask, raw[k].y, 1);
                s[6] = __shfl_down_sync(kFullMask, raw[k].z, 1);
                // no neighbour with this vector's successor: last lane of the warp, last live thread of the slice
                const bool need = (k * T < left) && (col + 4 - shift < own.n);
                if (need && (lane == 31 || !(k * T + 1 < left))) {
                    const float4 nx = load_aligned(a.signal, own_al + col + 4, a.n_sig);
                    s[4] = nx.x;
                    s[5] = nx.y;
                    s[6] = nx.z;
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float v = shift == 0 ? s[e] : shift == 1 ? s[e + 1] : shift == 2 ? s[e + 2] : s[e + 3];
                own_v[k][e] = (col + e < own.n) ? v : 0.0f;
            }
        } else {
            own_v[k][0] = (col < own.n) ? raw[k].x : 0.0f;
        }
    }

    // 4. Partner samples: column t of state s pairs with partner column t + shift_s, read from the
    //    partner's recording (zero where the partner's cut would have been padding).
    int live[kUnroll];
    float other[kUnroll][VEC];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
        const int col = col0 + k * T * VEC;
        live[k] = 0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) other[k][e] = 0.0f;
        if (k * T < left) {
            const int st = (col >= lo1) + (col >= lo2) + (col >= lo3);
            const int4 w = s_win[st];                                  // {start, blended, shift, next start}
            const int ahead = col - w.x;
            if (__builtin_expect(ahead >= 0 && col + (VEC - 1) < w.w, 1)) {
                const int m = w.y - ahead;
                const int pc = col + w.z;
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    if (e < m && pc + e < par.n) other[k][e] = __ldg(par_ptr + pc + e);
                live[k] = (1 << min(max(m, 0), VEC)) - 1;
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int t = col + e;
                    const int se = (t >= lo1) + (t >= lo2) + (t >= lo3);
                    const int4 we = s_win[se];
                    if (static_cast<unsigned>(t - we.x) < static_cast<unsigned>(we.y)) {
                        if (t + we.z < par.n) other[k][e] = __ldg(par_ptr + t + we.z);
                        live[k] |= 1 << e;
                    }
                }
            }
        }
    }

    // 5. Blend, warp, store.
    float* __restrict__ out_row = a.out + static_cast<size_t>(b) * a.n_per_cycle + static_cast<size_t>(row) * a.P;
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
        if (!(k * T < left)) continue;
        const int col = col0 + k * T * VEC;
        float res[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float blended = __fadd_rn(__fmul_rn(own_v[k][e], a.lam), __fmul_rn(other[k][e], a.one_minus_lam));
            res[e] = (live[k] >> e & 1) ? blended : own_v[k][e];
        }
        if constexpr (MAGWARP) {
            int piece = min(static_cast<int>(__umulhi(static_cast<unsigned>(col), a.piece_magic)), a.K);
            while (col >= s_kint[piece + 1]) ++piece;
            if (__builtin_expect(col + (VEC - 1) < s_kint[piece + 1], 1)) {
                const double2 c01 = *reinterpret_cast<const double2*>(&s_coef[piece * 4]);
                const double2 c23 = *reinterpret_cast<const double2*>(&s_coef[piece * 4 + 2]);
                const double dt = int_to_double(col) - s_kpos[piece];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const double de = dt + static_cast<double>(e);
                    const double wv = fma(fma(fma(c01.x, de, c01.y), de, c23.x), de, c23.y);
                    res[e] = static_cast<float>(static_cast<double>(res[e]) * wv);
                }
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int te = col + e;
                    int pe = min(static_cast<int>(__umulhi(static_cast<unsigned>(te), a.piece_magic)), a.K);
                    while (te >= s_kint[pe + 1]) ++pe;
                    const double de = int_to_double(te) - s_kpos[pe];
                    const double* c = &s_coef[pe * 4];
                    const double wv = fma(fma(fma(c[0], de, c[1]), de, c[2]), de, c[3]);
                    res[e] = static_cast<float>(static_cast<double>(res[e]) * wv);
                }
            }
        }
        if constexpr (VEC == 4) {
            st_stream4(out_row + col, res);
        } else {
            __stcs(out_row + col, res[0]);
        }
    }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

template <int VEC, int T>
cudaError_t launch_t(const MixArgs& a, dim3 grid, bool magwarp, cudaStream_t stream) {
    if (magwarp) {
        mix_resident_kernel<VEC, T, true><<<grid, T, 0, stream>>>(a);
    } else {
        mix_resident_kernel<VEC, T, false><<<grid, T, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

// One thread per batch slot: the slot's table row resolved into the 8-int record the pipelined
// kernel reads instead of `frames`: {f0..f4, first sample inside the recording, row of the recording's channel 0
// in `signal` viewed as [n_rec * R][T_sig], samples available}.  A slot whose table row or recording does not exist gets an empty record (all
// padding, nothing blended) and raises PCGMIX_ERR_BAD_PARTNER.
__global__ void __launch_bounds__(256) resolve_kernel(const __grid_constant__ MixArgs a, int32_t* __restrict__ records) {
    // the pipelined kernel behind us may start its prologue now; it waits (griddepcontrol.wait) for this
    // grid to finish before it reads a record
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int table_row = a.sel ? __ldg(a.sel + b) : b;
    int4 lo = make_int4(0, 0, 0, 0), hi = make_int4(0, 0, 0, 0);
    bool ok = static_cast<unsigned>(table_row) < static_cast<unsigned>(a.n_table);
    if (ok) {
        const int4 head = __ldg(reinterpret_cast<const int4*>(a.cycles) + static_cast<size_t>(table_row) * 2);
        const int4 tail = __ldg(reinterpret_cast<const int4*>(a.cycles) + static_cast<size_t>(table_row) * 2 + 1);
        ok = static_cast<unsigned>(head.x) < static_cast<unsigned>(a.n_rec);
        if (ok) {
            const int start = min(max(head.y, 0), a.T_sig);
            const int stop = min(max(head.z, start), a.T_sig);
            lo = make_int4(head.w, tail.x, tail.y, tail.z);
            hi = make_int4(tail.w, start, head.x * a.R, min(stop - start, a.P));
        }
    }
    reinterpret_cast<int4*>(records)[static_cast<size_t>(b) * 2] = lo;
    reinterpret_cast<int4*>(records)[static_cast<size_t>(b) * 2 + 1] = hi;
    if (!ok && a.err != nullptr) atomicOr(a.err, static_cast<int>(PCGMIX_ERR_BAD_PARTNER));
}

}  // namespace

cudaError_t launch_resolve_resident(const MixArgs& a, int32_t* records, cudaStream_t stream) {
    if (a.B == 0) return cudaSuccess;
    resolve_kernel<<<(a.B + 255) / 256, 256, 0, stream>>>(a, records);
    return cudaGetLastError();
}

cudaError_t launch_mix_resident(const MixArgs& base, bool magwarp, cudaStream_t stream) {
    MixArgs a = base;
    a.n_per_cycle = a.R * a.P;
    a.n_sig = static_cast<long long>(a.n_rec) * a.R * a.T_sig;
    if (magwarp) {
        const double ratio = static_cast<double>(a.K + 1) / static_cast<double>(a.P - 1);
        const double scaled = ratio * 4294967296.0 * (1.0 - 1e-9);
        a.piece_magic = scaled >= 4294967295.0 ? 4294967295u : static_cast<unsigned>(scaled);
    }
    if (a.R > 65535) return cudaErrorInvalidConfiguration;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(a.signal) | reinterpret_cast<uintptr_t>(a.out)) & 15u) == 0;
    const bool vec4 = aligned16 && (a.P % 4) == 0;
    const int units = vec4 ? a.P / 4 : a.P;                           // vectors per row
    const int slices = ceil_div(units, 256 * kUnroll);
    if (slices > 65535) return cudaErrorInvalidConfiguration;
    a.chunk_len = ceil_div(units, slices);
    a.chunks_per_cycle = slices;
    const dim3 grid(static_cast<unsigned>(a.B), static_cast<unsigned>(slices), static_cast<unsigned>(a.R));
    if (!vec4) return launch_t<1, 256>(a, grid, magwarp, stream);
    const int need = ceil_div(a.chunk_len, kUnroll);                  // threads that have work
    if (need <= 128) return launch_t<4, 128>(a, grid, magwarp, stream);
    if (need <= 160) return launch_t<4, 160>(a, grid, magwarp, stream);
    if (need <= 192) return launch_t<4, 192>(a, grid, magwarp, stream);
    if (need <= 224) return launch_t<4, 224>(a, grid, magwarp, stream);
    return launch_t<4, 256>(a, grid, magwarp, stream);
}

}  // namespace pcgmix
