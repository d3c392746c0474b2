/*
 * pcgmix_b200.h — C ABI of the B200-native PCGmix augmentation hot path.
 *
 * Everything here is `extern "C"`, takes plain DEVICE pointers + sizes + an explicit CUDA
 * stream, never allocates, never synchronises and never calls back into the host.  Return
 * value: 0 = launched, non-zero = argument or CUDA error (text via pcgmix_last_error()).
 *
 * Each entry point names the reference code it replaces (paths are into the upstream
 * PCGmix-EXTENDED tree):
 *
 *   pcgmix_mix1d ............ the per-cycle loop of `augment` for 'durratiomixup'
 *                             (augmentations.py:969-977) calling
 *                             mixup_keepdur_multidim_tensors (augmentations.py:289-304)
 *   pcgmix_mix1d_magwarp .... the same loop for 'durmixmagwarp' (augmentations.py:902-914)
 *                             fused with magnitude_warp (augmentations.py:674-683, called
 *                             at :924-928 through a device->host->device round trip)
 *   pcgmix_mix1d_windows .... the same two loops for the '(rand)' displacement variant
 *                             (augmentations.py:305-337)
 *   pcgmix_mix2d ............ the spectrogram loop (augmentations2d.py:419-426) calling
 *                             mixup_keepdur_multidim_tensors (augmentations2d.py:206-221);
 *                             the optional zero box covers durmixtimemask / durmixfreqmask /
 *                             durmixcutout (augmentations2d.py:286-395)
 *   pcgmix_segment_dense .... dense per-sample Springer states -> transitions -> complete
 *                             cycles -> frames[5] (databuilder.ipynb:593-606, cell 14)
 *   pcgmix_segment_table .... (position, state) transition table -> cycles
 *                             (databuilder.ipynb:928-948 cell 25; :370,:399 cell 6 when
 *                             spec_cols > 0)
 *   pcgmix_cut_cycles ....... cut each cycle out of its recording and zero-pad to L
 *                             (databuilder.ipynb:627-632, :973-978; :403-411)
 *   pcgmix_duration_features  per-cycle durations, BPM and duration ratios
 *                             (classical.py:245-283)
 *   pcgmix_mix1d_resident ... cut + zero-pad + PCGmix(+) in one pass over recordings that stay
 *                             on the device: databuilder.ipynb:627-632, :973-978 (cut, resize),
 *                             dataloader_physionet.py:43-48 (stacking the padded cycles),
 *                             train_model.py:499 (batch upload) and the augmentation loops
 *                             above, without ever materialising the padded (n, C, L) array
 *
 * Host draws (probability gate, pairing, lambda, spline knots) are NOT part of this ABI:
 * they are made on the host from the reference's seed rule and passed in, so both
 * implementations consume identical randomness.
 */
#ifndef PCGMIX_B200_H_
#define PCGMIX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCGMIX_B200_VERSION 104

/* bits OR-ed into *err_flag (device int32, may be NULL) by the kernels */
#define PCGMIX_ERR_BAD_PARTNER   1   /* mix[b] outside [0,B): cycle copied unmixed          */
#define PCGMIX_ERR_BAD_FRAMES    2   /* offsets negative / decreasing, or a pair's clamped windows
                                        differ in width (the reference raises): cycle copied  */
#define PCGMIX_ERR_BAD_PATTERN   4   /* S1 followed by something other than sys,S2,dia        */
#define PCGMIX_ERR_OVERFLOW      8   /* more cycles/transitions than the output can hold      */
#define PCGMIX_ERR_ZERO_DIVISION 16  /* duration ratio with a zero denominator                */
#define PCGMIX_ERR_EMPTY_STATE   32  /* a heart state without samples: features are NaN       */

#define PCGMIX_CYCLE_FEATURES 36     /* floats per cycle written by pcgmix_cycle_features       */
#define PCGMIX_CYCLE_PSD_FEATURES 80 /* floats per cycle written by pcgmix_cycle_psd_features   */
#define PCGMIX_MAX_FIRST_BLOCK_FILTERS 512 /* output channels pcgmix_first_conv_block accepts (records live in shared memory) */
#define PCGMIX_CYCLE_MOMENT_FEATURES 10 /* floats per cycle written by pcgmix_cycle_moment_features */

#define PCGMIX_MAX_KNOT 30           /* largest `knot` of durmixmagwarp(sigma,knot) supported */

/* state codes used by the segmentation entry points */
#define PCGMIX_STATE_S1       1
#define PCGMIX_STATE_SYSTOLE  2
#define PCGMIX_STATE_S2       3
#define PCGMIX_STATE_DIASTOLE 4
#define PCGMIX_STATE_NOISE    5

typedef void* pcgmix_stream_t;       /* a cudaStream_t / CUstream */

int pcgmix_version(void);
const char* pcgmix_last_error(void); /* thread-local, valid until the next failing call */
/* SM count and compute capability of the current device (for grid sizing by callers). */
int pcgmix_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/*
 * Kernel selection knobs (process-wide; meant for benchmarking, the defaults are the tuned ones).
 *   use_pipeline  1: rows of >= 1024 floats with P % 4 == 0 and 16-byte-aligned tensors go through
 *                 the persistent TMA-pipelined kernel; 0: always the direct-load kernel
 *   stages        shared-memory ring depth of the pipelined kernel (1..8; 0 = default)
 *   max_slice     elements of a row handled per pipeline step (multiple of 4; 0 = default)
 *   ctas_per_sm   cap on resident CTAs per SM (0 = as many as fit)
 *   pbuf_pct      shared-memory budget for staged partner windows, in % of a slice (0 = default);
 *                 slices whose windows do not fit read the partner straight from global memory
 *   consumer_threads  lower bound on consumer threads per CTA (0 = just enough for one slice)
 *   debug         must be 0 in production.  Bits 0-4 switch parts of the pipelined kernel off for
 *                 profiling and give WRONG output (1 no stores, 2 no arithmetic, 4 no partner
 *                 staging, 8 no coefficient set-up, 16 no knot loads, 64 no float32 safety check).  They
 *                 exist only in a library built with -DPCGMIX_PROFILING; the default build refuses
 *                 them.  Bit 5 (32: no coefficient table, every item takes the producers' per-item
 *                 path) changes no result and is always honoured
 * Results do not depend on any of these.
 */
int pcgmix_set_tuning(int32_t use_pipeline, int32_t stages, int32_t max_slice, int32_t ctas_per_sm,
                      int32_t pbuf_pct, int32_t consumer_threads, int32_t debug);

/*
 * How the pipelined kernel evaluates the magnitude-warp factor of PCGmix+ (process-wide).
 *   1 (default)  float32: normalised Horner on coefficients rounded from float64, fp32 product.  Within
 *                1e-5 relative of the reference (BASELINE's tolerance; measured ~3e-7, factors below 1/16 in
 *                magnitude are re-evaluated in float64), at the speed of plain PCGmix.
 *   0            float64 Horner and one rounding of fp64(sample)*w to fp32, like the reference's float64
 *                product stored into a float32 array: > 99.9 % of samples bit-equal to it, ~15 % slower.
 * The direct-load kernels (short or unaligned rows) always evaluate in float64.
 */
int pcgmix_set_spline_precision(int32_t float32_evaluation);
int pcgmix_get_spline_precision(void);

/*
 * Launch overlap between consecutive PCGmix launches on one stream (default: off).
 * When enabled, a pipelined launch whose buffers are disjoint from those of the previous two PCGmix
 * launches on the same stream (checked here: no read-after-write, write-after-read or
 * write-after-write overlap) is issued with the programmatic-stream-serialization attribute, and
 * every pipelined kernel signals `griddepcontrol.launch_dependents` as soon as its CTAs are running:
 * the next launch's CTAs take over SMs as this one's retire, so its pipeline fills while the previous
 * one drains (measured: 12-14 us per launch at 4096 cycles).  Only grids that fill the GPU exactly (CTA
 * residency as counted by the runtime's occupancy calculator) with identical geometry are overlapped,
 * so at most two are in flight.  Every OTHER entry point of this library that launches on the stream
 * (pcgmix_copy_small, the segmentation / cut / feature kernels, pcgmix_mix1d_resident) resets the
 * bookkeeping, so a mix launch is never made a programmatic dependent of the kernel that produced its
 * tables.  The library cannot see kernels of OTHER libraries: by enabling overlap the caller asserts
 * that nothing it enqueues itself between two PCGmix launches on that stream writes their inputs or
 * reads their outputs.  Results are unchanged.
 */
int pcgmix_set_launch_overlap(int32_t enable);
/* Number of launches issued so far with the overlap attribute (diagnostics). */
long long pcgmix_overlap_launches(void);

/*
 * PCGmix on time series.
 *   x, out ....... [B][C][L] fp32, contiguous, must not alias
 *   frames ....... int32, row b at frames + b*frame_stride: the five cumulative state offsets
 *                  (S1, systole, S2, diastole starts, next-S1) of cycle b.  frame_stride is 5
 *                  for a packed [B][5] array, 8 (with frames = cycles + 3) for the cycle table
 *                  written by pcgmix_segment_dense/_table
 *   mix .......... [B] int32: partner cycle of every cycle
 *   order ........ [B] int32 or NULL: processing order of the cycles (a permutation of 0..B-1);
 *                  changes nothing in the result, only which cycles are in flight together
 *   lam, one_minus_lam  fp32 lambda and fp32(1)-lambda as rounded by the caller
 * out[b,c,t] = x[b,c,t]*lam + x[mix[b],c,f2[s]+(t-f1[s])]*one_minus_lam for t inside the
 * first min(len1_s,len2_s) samples of state s, else x[b,c,t]; mul, mul, add each rounded
 * to fp32 (no FMA contraction), exactly like the reference's tensor expression.
 * Offsets may exceed L (the reference keeps cycles longer than the padded row): both windows are
 * clamped to the row like the reference's Python slices and blended when the clamped widths agree;
 * otherwise the reference raises a shape mismatch and this kernel copies the cycle and raises
 * PCGMIX_ERR_BAD_FRAMES.  Entries of `order` and `mix` outside [0,B) are not followed
 * (PCGMIX_ERR_BAD_PARTNER).
 */
int pcgmix_mix1d(const float* x, float* out, const int32_t* frames, int32_t frame_stride,
                 const int32_t* mix, const int32_t* order, float lam, float one_minus_lam,
                 int32_t B, int32_t C, int32_t L, int32_t* err_flag, pcgmix_stream_t stream);

/*
 * PCGmix+ : the mix above, multiplied in the same pass by a per-(cycle,channel) cubic spline.
 *   knots ........ [B][K+2][C] fp64, the reference's N(1,sigma) draw, in the drawn layout
 *   coefmat ...... [(K+1)*4][K+2] fp64: linear map knots -> piecewise-cubic coefficients of
 *                  the not-a-knot spline, row (k*4+i) = coefficient of dt^(3-i) on piece k
 *   knot_pos ..... [K+3] fp64: np.linspace(0, L-1, K+2), followed by ONE more value d >= 0: a bound
 *                  such that max_j |knot_j - 1| < d guarantees spline(t) > 0 at every sample
 *                  (0.999 / Lebesgue constant of the knot -> curve map; 0 = not known).  Used by
 *                  pcgmix_mix1d_resident, which knows which samples are zero padding: for rows
 *                  inside the bound padding is written as +0.0f without evaluating the spline
 *                  ((+0.0f) * w is +0.0f for every positive finite w)
 * out = fp32( fp64(mixed) * spline(t) ), like the reference's float64 product stored to fp32.
 */
int pcgmix_mix1d_magwarp(const float* x, float* out, const int32_t* frames, int32_t frame_stride,
                         const int32_t* mix, const int32_t* order, float lam, float one_minus_lam,
                         const double* knots, const double* coefmat, const double* knot_pos,
                         int32_t K, int32_t B, int32_t C, int32_t L,
                         int32_t* err_flag, pcgmix_stream_t stream);

/*
 * PCGmix / PCGmix+ with the blended windows given explicitly instead of derived from the offsets:
 * windows[b][s] = {start in cycle b, blended length, shift to the partner's sample} for the four
 * states (int32 [B][4][3]).  This is what the reference's '(rand)' displacement variant needs
 * (augmentations.py:305-337: the shorter state is placed at a seeded random offset inside the
 * longer one).  Windows must be ordered and disjoint.  knots == NULL selects plain PCGmix.
 */
int pcgmix_mix1d_windows(const float* x, float* out, const int32_t* windows, const int32_t* mix,
                         const int32_t* order, float lam, float one_minus_lam,
                         const double* knots, const double* coefmat, const double* knot_pos,
                         int32_t K, int32_t B, int32_t C, int32_t L,
                         int32_t* err_flag, pcgmix_stream_t stream);

/*
 * PCGmix on spectrograms: x, out are [B][Ch][F][T]; frames are in time-frame (column) units.
 * Optional zero box applied after the mix (composites): rows f in [h1,h2) x columns
 * t in [tbox[b][0], tbox[b][1]) are set to 0.  tbox == NULL means "all columns";
 * h1 >= h2 disables the box.
 */
int pcgmix_mix2d(const float* x, float* out, const int32_t* frames, int32_t frame_stride,
                 const int32_t* mix, const int32_t* order, float lam, float one_minus_lam,
                 int32_t B, int32_t Ch, int32_t F, int32_t T,
                 const int32_t* tbox, int32_t h1, int32_t h2,
                 int32_t* err_flag, pcgmix_stream_t stream);

/*
 * Dense states -> cycles.  states [R][T] int8 in {1,2,3,4}; positions are floor-divided by
 * `downsample` on the absolute index before the offsets are formed.  One cycle per transition
 * into S1 that has a later S1 in the same recording.  Output rows, in (recording, time)
 * order: cycles[i] = {recording, abs_start, abs_stop, f0..f4} (8 x int32, 16-byte aligned),
 * abs_* in rescaled units, f0 = 0.
 *   cycle_count .. [R+1] int32 out: row pointers — the cycles of recording r are rows
 *                  [cycle_count[r], cycle_count[r+1]); cycle_count[R] is the total
 *   max_cycles ... capacity of `cycles` (rows); cycles beyond it are dropped and flagged
 * At most 8192 transitions per recording are examined (more raises PCGMIX_ERR_OVERFLOW).
 * A transition into S1 (with a later S1) not followed by systole, S2, diastole raises
 * PCGMIX_ERR_BAD_PATTERN where the reference raises an exception.
 */
int pcgmix_segment_dense(const int8_t* states, int32_t R, int32_t T, int32_t downsample,
                         int32_t* cycles, int32_t max_cycles, int32_t* cycle_count,
                         int32_t* err_flag, pcgmix_stream_t stream);

/*
 * Transition table -> cycles.  positions/codes are the concatenated (position, state-code)
 * rows of R recordings; rec_offsets [R+1] delimits them.  If spec_cols > 0 the offsets are
 * formed from round_half_even(position*spec_cols/rec_len[r]) instead of position/downsample.
 */
int pcgmix_segment_table(const int32_t* positions, const int8_t* codes, const int32_t* rec_offsets,
                         int32_t R, int32_t downsample, int32_t spec_cols, const int32_t* rec_len,
                         int32_t* cycles, int32_t max_cycles, int32_t* cycle_count,
                         int32_t* err_flag, pcgmix_stream_t stream);

/*
 * Cut + zero-pad.  signal [R][C][T] fp32 (C bands/rows of every recording); for cycle i
 * out[i][c][0:L] = signal[rec][c][abs_start : abs_stop] truncated/zero-filled to L.
 * n_cycles may be read from the device (cycle_count[R]) by passing n_cycles_dev != NULL, in
 * which case `n_cycles` is only the grid bound.
 */
int pcgmix_cut_cycles(const float* signal, int32_t R, int32_t C, int32_t T,
                      const int32_t* cycles, int32_t n_cycles, const int32_t* n_cycles_dev,
                      float* out, int32_t L, pcgmix_stream_t stream);

/*
 * PCGmix / PCGmix+ straight from resident recordings.  Batch slot i is row sel[i] of the cycle
 * table `cycles` [n_table][8] = {recording, abs_start, abs_stop, f0..f4} (sel == NULL: slot i is
 * row i, B <= n_table); its samples are signal[recording][c][abs_start + t] for
 * t < min(abs_stop - abs_start, L) and zero beyond (what pcgmix_cut_cycles would store), its
 * state offsets are f0..f4 of that row, its partner is slot mix[i].
 *   out[i] == pcgmix_mix1d(_magwarp)(pcgmix_cut_cycles(signal, rows sel), frames of rows sel, mix)
 * bit for bit, with out [B][C][L].  knots == NULL: no magnitude warp (PCGmix); otherwise knots
 * [B][K+2][C], coefmat, knot_pos as for pcgmix_mix1d_magwarp.  `signal` and `cycles` must be
 * 16-byte aligned.  Rows / recordings / partners out of range and offsets outside [0, L] copy the
 * cycle unmixed and raise PCGMIX_ERR_BAD_PARTNER / PCGMIX_ERR_BAD_FRAMES in *err_flag.
 * scratch: [B][8] int32 of caller-owned device memory, 16-byte aligned, or NULL.  With scratch (and
 * L % 4 == 0, L >= 1024, 16-byte-aligned out) a small kernel resolves the slots into records there
 * and the persistent TMA-pipelined kernel does the work; without it a direct-load kernel is used.
 * The result is the same.
 */
int pcgmix_mix1d_resident(const float* signal, int32_t n_rec, int32_t C, int32_t T,
                          const int32_t* cycles, int32_t n_table, const int32_t* sel,
                          const int32_t* mix, const int32_t* order, float lam, float one_minus_lam,
                          const double* knots, const double* coefmat, const double* knot_pos, int32_t K,
                          float* out, int32_t B, int32_t L, int32_t* scratch, int32_t* err_flag,
                          pcgmix_stream_t stream);

/*
 * Classical per-cycle features of one channel of a batch of (augmented) cycles: what
 * classical.feature_vector_seg computes in its amplitude block (classical.py:284-303) and its Hilbert
 * envelope block (:305-360), for the loop at train_model.py:519-532.  x [B][C][L] fp32, frames as for
 * pcgmix_mix1d, `channel` the row of every cycle to use (the reference passes d[4]), `what` a mask:
 * 1 = amplitude block, 2 = envelope block.  features [B][PCGMIX_CYCLE_FEATURES] fp32:
 *    0..3   max amplitude of S1, systole, S2, diastole
 *    4..9   round(.,4) of max ratios S1/S2, sys/dia, sys/S1, sys/S2, dia/S1, dia/S2
 *   10..14  envelope integral (np.trapz, dx=5) of S1, systole, S2, diastole, RR
 *   15..22  round(.,4) of integral ratios S1/S2, sys/dia, S1/RR, sys/RR, S2/RR, dia/RR, sys/S1, dia/S2
 *   23..27  mean envelope of S1, systole, S2, diastole, RR
 *   28..35  mean-envelope ratios S1/RR, sys/RR, S2/RR, dia/RR, sys/dia, sys/S1, dia/S2, S1/S2
 * Segments are the reference's slices (S1 = data[:f1], RR = data[:f4], bounds clamped to L).  The amplitude
 * block is exact (float32, NumPy's round); the envelope block agrees with SciPy's single-precision FFT
 * to float32 rounding.  Blocks not requested are left untouched.  A cycle with an empty state gets NaN
 * features and raises PCGMIX_ERR_EMPTY_STATE (the reference raises).  L <= 14000 for the envelope block.
 */
int pcgmix_cycle_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C,
                          int32_t L, int32_t channel, int32_t what, float* features, int32_t* err_flag,
                          pcgmix_stream_t stream);

/*
 * Power-spectral-density block of classical.feature_vector_seg (classical.py:358-643) for one channel of a
 * batch of (augmented) cycles.  For the whole beat (data[:f4]), the systole (data[f1:f2]) and the diastole
 * (data[f3:f4]), in this order, 26 values each:
 *    +0      mean of scipy.signal.welch(segment, fs)'s PSD (Hann window of min(256, n) samples, half overlap,
 *            mean removed per window, density scaling, one-sided, mean over the windows)
 *    +1      mean of PSD / I, I = np.trapz(|hilbert(PSD)|, dx=5)
 *    +2+2j   mean of the PSD over the bins with lo_j <= f <= hi_j   } bands 25-40, 40-60, 60-80, 80-100, 100-120,
 *    +3+2j   mean of PSD / I over the same bins                     } 120-140, 140-160, 160-180, 180-200, 200-250,
 *                                                                     250-300, 300-400 Hz; NaN for a band without bins
 * then  78: round(mean(PSD/I) systole / RR, 4),  79: the same diastole / RR.  features [B][PCGMIX_CYCLE_PSD_FEATURES]
 * fp32.  Agrees with SciPy's single-precision arithmetic to float32 rounding (tests: 2e-5 relative, NaN patterns
 * identical).  A cycle with an empty beat, systole or diastole gets NaN features and raises PCGMIX_ERR_EMPTY_STATE
 * (the reference raises).  fs > 0.
 */
int pcgmix_cycle_psd_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C,
                              int32_t L, int32_t channel, int32_t fs, float* features, int32_t* err_flag,
                              pcgmix_stream_t stream);

/*
 * Skewness / kurtosis block of classical.feature_vector_seg (classical.py:893-905): scipy.stats.skew (0..4) and
 * scipy.stats.kurtosis (5..9; biased, Fisher) of RR = data[:f4], S1 = data[:f1], systole, S2, diastole of one channel of
 * every cycle.  features [B][PCGMIX_CYCLE_MOMENT_FEATURES] fp32; float32 arithmetic like SciPy's for float32 rows (tests:
 * 2e-5 relative + 2e-6 absolute); NaN for a segment without variance like there.  An empty segment gets NaN and raises
 * PCGMIX_ERR_EMPTY_STATE.
 */
int pcgmix_cycle_moment_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C,
                                 int32_t L, int32_t channel, float* features, int32_t* err_flag,
                                 pcgmix_stream_t stream);

/*
 * Forward pass of the first block of the reference's ResNet9-1D on a batch of (augmented) cycles: nn.Conv1d(C, F,
 * kernel_size=3, padding=1) + nn.BatchNorm1d(F) + nn.ReLU — models.py:468-473 (conv_block), instantiated as conv1 at
 * models.py:523 and applied to augment()'s output at models.py:538 / train_model.py:536.  SURVEY 8(f)4.
 *   x [B][C][L] fp32 (C = 1..4), weight [F][C][3], bias / gamma / beta [F] or NULL (0 / 1 / 0), out [B][F][L] fp32,
 *   disjoint from x.  F <= PCGMIX_MAX_FIRST_BLOCK_FILTERS.
 *   batch_stats = 1 (a module in training mode, or one without running statistics): normalises with the mean and
 *     biased variance of the convolution's output over (B, L) — obtained from second moments of the INPUT patches, so
 *     the output is written exactly once — and, when running_mean / running_var are given, updates them in place as
 *     torch does: running = (1 - momentum) * running + momentum * batch (variance unbiased, n / (n - 1)).
 *   batch_stats = 0 (evaluation mode): normalises with running_mean / running_var (required), which are not changed.
 *   save_mean / save_invstd [F] or NULL: the statistics used (what torch's batch_norm saves for its backward).
 *   workspace: pcgmix_first_conv_block_workspace(C, F) bytes of device memory, 16-byte aligned, private to the call
 *     until the stream has passed it.
 * Forward only (no gradient kernels).  float32 FMAs; agrees with torch's float32 modules to rounding (tests: 2e-5
 * relative + 2e-5 absolute on the output, 1e-5 relative on the statistics).  NaN propagates like torch's ReLU.
 * The batch variance is formed from float32 products of raw samples summed in float64, so its relative error is about
 * 1e-6 * (1 + mean^2 / variance) of the input: exact enough for band-passed, zero-centred heart-sound cycles (what the
 * reference feeds it), not for inputs riding on a large constant offset.
 * Enqueues at most one memset and three kernels on `stream`; no allocation, no synchronisation.
 */
long long pcgmix_first_conv_block_workspace(int32_t C, int32_t F);   /* bytes; -1 for unsupported sizes */
int pcgmix_first_conv_block(const float* x, const float* weight, const float* bias, const float* gamma,
                            const float* beta, float* running_mean, float* running_var, float* out, void* workspace,
                            int32_t B, int32_t C, int32_t L, int32_t F, int32_t batch_stats, double eps,
                            double momentum, float* save_mean, float* save_invstd, pcgmix_stream_t stream);

/* 14 fp64 features per cycle from frames [n][5] int32 (stride `frame_stride` int32 between rows). */
int pcgmix_duration_features(const int32_t* frames, int32_t frame_stride, int32_t n, int32_t fs,
                             double* features, int32_t* err_flag, pcgmix_stream_t stream);

/*
 * Copy `bytes` (multiple of 4, 4-byte-aligned pointers) between device memory and PINNED host memory
 * (either direction) with a small kernel that reaches the host buffer through unified addressing,
 * instead of with a copy engine.  Used for the per-step tables (offsets, pairing, knots; ~1 MB up)
 * and the class ids (down) so that they do not queue behind the large batch copies a prefetching
 * loader keeps in flight on the copy engines.  Enqueued on `stream`.
 */
int pcgmix_copy_small(void* dst, const void* src, int64_t bytes, pcgmix_stream_t stream);

/*
 * HOST helper (no CUDA): the pairing draw.  For every group g (group[i] in [0, n_groups)), with
 * members idx in ascending order, mix[idx] = random.Random(seed).sample(idx, len(idx)) exactly as
 * CPython computes it (MT19937, init_by_array seeding, _randbelow, pool algorithm) — the reference's
 * get_same_label_mix_indices / get_same_wav_mix_indices / get_same_dataset_mix_indices
 * (augmentations.py:500-556).  seed must be >= 0.
 */
int pcgmix_host_group_permutation(const int64_t* group, int64_t n, int64_t n_groups, uint64_t seed,
                                  int64_t* mix);

/*
 * Host-side replay of the reference's lambda and magnitude-warp knot draws from NumPy's legacy global stream
 * (augmentations.py:659-666 get_lambda, :677 magnitude_warp):
 *     np.random.seed(seed); lam = np.random.beta(alpha, alpha); knots = np.random.normal(1.0, sigma, n_knots)
 * *lam_out and knots_out[n_knots] receive bit-identical values.  state_out[624] / *pos_out / *has_gauss_out /
 * *gauss_out (all optional) receive the generator state NumPy would be left in (np.random.get_state()), so a
 * caller can mirror the reference's side effect on the global stream.  max_threads bounds the worker threads used
 * for the log/sqrt pass of large draws.  Returns non-zero (and draws nothing) when seed > 2^32-1 (NumPy refuses
 * such seeds) or alpha <= 0 (the reference then neither re-seeds nor draws lambda: use NumPy's stream as it is).
 */
int pcgmix_host_lambda_knots(uint64_t seed, double alpha, double sigma, int64_t n_knots, int32_t max_threads,
                             double* lam_out, double* knots_out, uint32_t* state_out, int32_t* pos_out,
                             int32_t* has_gauss_out, double* gauss_out);

/*
 * order[k] = cycles in pairing-chain order (b, mix[b], mix[mix[b]], ... for b = 0, 1, ... not yet visited): the
 * processing order that keeps a cycle read as "partner" in L2 until it is read as "itself".  Host code.
 */
int pcgmix_host_processing_order(const int64_t* mix, int64_t n, int32_t* order);

/*
 * The integer half of the host's share of one plain PCGmix / PCGmix+ step in a single call (no Python in between):
 * the reference's get_same_label_mix_indices (augmentations.py:500-514) plus the offset checks of this
 * implementation, packed into a staging buffer that pcgmix_copy_small uploads verbatim.  labels[B]: class id per
 * cycle; frames: CPU offsets, row b at frames + b*frame_stride (int64), 5 per row; L: row length; seed: the step
 * count; K >= 0: reserve a section for (B, K+2, C) float64 knots (the caller fills it); want_order: also write the
 * pairing-chain processing order.  info[0..3]: byte offsets in `packed` of {frames int32 [B][5], mix int32 [B],
 * order int32 [B], knots} (-1 = absent), info[4]: bytes used; mix_out[B]: the pairing as int64.  info holds 16 entries.
 * Returns 0, or: 1 bad arguments / buffer too small (info[4] = bytes needed); 2 offsets of cycle info[5] negative,
 * decreasing or beyond int32; 3 cycle info[5], state info[6]: clamped destination / source windows of
 * info[7] / info[8] samples (the reference raises a shape mismatch there).
 */
int pcgmix_host_prepare_step(const int64_t* labels, int64_t B, const int64_t* frames, int64_t frame_stride, int64_t L,
                             uint64_t seed, int32_t K, int32_t C, int32_t want_order, uint8_t* packed, int64_t capacity,
                             int64_t* info, int64_t* mix_out);

#ifdef __cplusplus
}
#endif
#endif /* PCGMIX_B200_H_ */
