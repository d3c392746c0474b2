// Shared declarations for the PCGmix B200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pcgmix_b200.h"

namespace pcgmix {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kMaxPieces = PCGMIX_MAX_KNOT + 1;   // cubic pieces of the magnitude-warp spline

// Arguments of the fused gather-mix(-warp) kernel.  A "cycle" is R rows of pitch P floats,
// contiguous (1D: R = channels, P = L samples; 2D: R = Ch*F, P = T time frames).
struct MixArgs {
    const float* x;
    float* out;
    const int32_t* frames;     // row b at frames + b*frame_stride, 5 offsets
    int32_t frame_stride;
    const int32_t* mix;        // [B]
    const int32_t* windows;    // [B][4][3] {start, blended length, partner shift} or nullptr (derive from frames)
    const int32_t* order;      // [B] or nullptr
    int32_t* err;              // device flag word or nullptr
    float lam;
    float one_minus_lam;
    int32_t B;
    int32_t R;
    int32_t P;
    int32_t n_per_cycle;       // R*P
    int32_t nvec;              // vector units per cycle (n_per_cycle / VEC)
    int32_t chunk_len;         // vector units per CTA
    int32_t chunks_per_cycle;
    int32_t qstep;             // blockDim*VEC = qstep*P + rstep
    int32_t rstep;
    // magnitude warp (PCGmix+)
    const double* knots;       // [B][K+2][R]
    const double* coefmat;     // [(K+1)*4][K+2]
    const double* knot_pos;    // [K+2]
    double inv_h;              // (K+1)/(P-1)
    uint32_t piece_magic;      // floor(2^32*(K+1)/(P-1)), rounded down: umulhi(t, magic) <= piece(t)
    int32_t K;
    // zero box (2D composites)
    const int32_t* tbox;       // [B][2] or nullptr
    int32_t F;
    int32_t h1;
    int32_t h2;
    // resident recordings (mix_resident.cu): cycles are cut out of `signal` on the fly
    const float* signal;       // [n_rec][R][T_sig]
    const int32_t* cycles;     // cycle table [n_table][8] {recording, abs_start, abs_stop, f0..f4}
    const int32_t* sel;        // [B] table row of every batch slot, or nullptr (slot i = row i)
    int32_t n_table;
    int32_t n_rec;
    int32_t T_sig;
    long long n_sig;           // n_rec*R*T_sig
};

// One heart state of a cycle against the same state of its partner, with the slice clamping of the
// reference's tensor expression (augmentations.py:289-304, augmentations2d.py:206-221):
//     d_new[:, f1:f1+n] = d_new[:, f1:f1+n]*lam + d2[:, f2:f2+n]*(1-lam),   n = min(len1, len2)
// Python slices are clamped to the row length P, so offsets beyond P are legal there (the data builder
// keeps cycles longer than the padded length: databuilder.ipynb cells 14/25 print a warning and truncate
// the samples, the offsets stay).  The assignment works when the clamped destination and source
// slices have equal width (blend that many samples), or when an empty destination meets a one-sample
// source (broadcast to nothing); every other combination makes the reference raise a shape error —
// except a one-sample source against a wider destination, which it broadcasts over the window; that
// case is refused here as well (PCGMIX_ERR_BAD_FRAMES).  Returns false for refused / non-monotone input.
__device__ __forceinline__ bool pair_window(int f1, int f1n, int f2, int f2n, int P, int& start, int& n, int& shift) {
    const int len1 = f1n - f1;
    const int len2 = f2n - f2;
    const int n0 = min(len1, len2);
    const int a1 = min(f1, P), e1 = min(f1 + n0, P);
    const int a2 = min(f2, P), e2 = min(f2 + n0, P);
    const int wd = e1 - a1, ws = e2 - a2;
    start = a1;
    shift = a2 - a1;
    n = wd == ws ? wd : 0;
    const bool sane = (f1 >= 0) & (f2 >= 0) & (len1 >= 0) & (len2 >= 0);
    return sane & ((wd == ws) | ((wd == 0) & (ws == 1)));
}

// Cycle handled in processing slot `slot`: order[slot], or the slot itself when no order was given.  An entry
// outside [0, B) is not followed (it would index out of the batch): the slot's own cycle is processed instead
// and PCGMIX_ERR_BAD_PARTNER is raised.
__device__ __forceinline__ int cycle_of_slot(const MixArgs& a, int slot) {
    if (a.order == nullptr) return slot;
    const int b = __ldg(a.order + slot);
    if (static_cast<unsigned>(b) < static_cast<unsigned>(a.B)) return b;
    if (a.err != nullptr) atomicOr(a.err, static_cast<int>(PCGMIX_ERR_BAD_PARTNER));
    return slot;
}

// mix_kernels.cu — direct-load kernel (any shape)
cudaError_t launch_mix(const MixArgs& base, bool magwarp, bool box, cudaStream_t stream);

// mix_resident.cu — cut + zero-pad + mix(+warp) in one pass over recordings that stay on the device
cudaError_t launch_mix_resident(const MixArgs& base, bool magwarp, cudaStream_t stream);
// slot -> {f0..f4, first sample in the recording, the recording's first row, samples available} records for the
// pipelined kernel's RESIDENT variant
cudaError_t launch_resolve_resident(const MixArgs& a, int32_t* records, cudaStream_t stream);

// mix_pipeline.cu — persistent TMA-pipelined kernel (rows of >= 1024 floats, P % 4 == 0, aligned)
struct PipelineTuning {
    int enabled;        // 0: always use the direct-load kernel
    int stages;         // ring depth (1..4), 0 = default
    int max_slice;      // elements per slice (multiple of 4), 0 = default
    int ctas_per_sm;    // cap on resident CTAs per SM, 0 = as many as fit
    int pbuf_pct;       // shared-memory budget of the packed partner windows, % of a slice, 0 = default
    int consumer_threads;  // lower bound on consumer threads per CTA, 0 = just enough for a slice
    int vec_per_thread; // (retired knob: the kernel always takes two 128-bit vectors per consumer thread and slice)
    int debug;          // profiling only (results become wrong): 1 skip stores, 2 skip arithmetic, 4 skip partner copies
    int spline_f32;     // PCGmix+: 1 = warp factor in float32 (<= 1e-5 relative, default), 0 = float64 (bit-faithful)
};
bool pipeline_applicable(const MixArgs& a, bool box);
// overlap_previous: launch with the programmatic-stream-serialization attribute, i.e. this kernel may start
// while the previous kernel on the stream (if it is one of ours, which trigger early) is still draining.
cudaError_t launch_mix_pipeline(const MixArgs& base, bool magwarp, const PipelineTuning& tune, bool overlap_previous,
                                unsigned long long previous_signature, cudaStream_t stream,
                                unsigned long long* full_grid_signature);

#ifdef PCGMIX_PROFILING
cudaError_t read_timeline(unsigned long long* host, int n_ctas);   // mix_pipeline.cu, profiling build only
#endif

// segment_kernels.cu
cudaError_t launch_segment_dense(const int8_t* states, int32_t R, int32_t T, int32_t downsample,
                                 int32_t* cycles, int32_t max_cycles, int32_t* cycle_count,
                                 int32_t* err, cudaStream_t stream);
cudaError_t launch_segment_table(const int32_t* positions, const int8_t* codes, const int32_t* rec_offsets,
                                 int32_t R, int32_t downsample, int32_t spec_cols, const int32_t* rec_len,
                                 int32_t* cycles, int32_t max_cycles, int32_t* cycle_count,
                                 int32_t* err, cudaStream_t stream);
cudaError_t launch_cut_cycles(const float* signal, int32_t R, int32_t C, int32_t T, const int32_t* cycles,
                              int32_t n_cycles, const int32_t* n_cycles_dev, float* out, int32_t L,
                              cudaStream_t stream);
// feature_kernels.cu
cudaError_t launch_cycle_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C,
                                  int32_t L, int32_t channel, int32_t what, float* features, int32_t* err,
                                  cudaStream_t stream);
// psd_kernels.cu — Welch PSD / band-mean block of feature_vector_seg
cudaError_t launch_cycle_psd_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C,
                                      int32_t L, int32_t channel, int32_t fs, float* features, int32_t* err,
                                      cudaStream_t stream);
cudaError_t launch_cycle_moment_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C,
                                         int32_t L, int32_t channel, float* features, int32_t* err, cudaStream_t stream);
cudaError_t launch_duration_features(const int32_t* frames, int32_t frame_stride, int32_t n, int32_t fs,
                                     double* features, int32_t* err, cudaStream_t stream);

// first_block_kernels.cu — Conv1d(C -> F, k=3, pad=1) + BatchNorm1d + ReLU forward of the reference's first model block
struct FirstBlockArgs {
    const float* x;            // [B][C][L]
    const float* weight;       // [F][C][3]
    const float* bias;         // [F] or nullptr
    const float* gamma;        // [F] or nullptr
    const float* beta;         // [F] or nullptr
    float* running_mean;       // [F] or nullptr; updated in place when batch statistics are used
    float* running_var;        // [F] or nullptr
    float* out;                // [B][F][L]
    void* workspace;           // first_conv_block_workspace_bytes(C, F), 16-byte aligned
    float* save_mean;          // [F] or nullptr: the mean / inverse standard deviation the block normalised with
    float* save_invstd;
    double eps;
    double momentum;
    int32_t B, C, L, F;
    int32_t batch_stats;       // 1: statistics of this batch (training mode), 0: the running statistics
};
size_t first_conv_block_workspace_bytes(int32_t C, int32_t F);
cudaError_t launch_first_conv_block(const FirstBlockArgs& p, cudaStream_t stream);

}  // namespace pcgmix
