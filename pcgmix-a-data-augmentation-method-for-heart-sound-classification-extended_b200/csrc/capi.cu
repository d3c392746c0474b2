// extern "C" boundary of libpcgmix_b200.so — see include/pcgmix_b200.h for the contract.
// Argument checks happen here; nothing below this file allocates or synchronises.

#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

thread_local char g_error[512] = "";

// Kernel-selection knobs (pcgmix_set_tuning).  Written and read under g_tuning_mutex: launches take a copy.
pcgmix::PipelineTuning g_tuning = {1, 0, 0, 0, 0, 0, 0, 0, 1};
std::mutex g_tuning_mutex;

pcgmix::PipelineTuning tuning_now() {
    std::lock_guard<std::mutex> lock(g_tuning_mutex);
    return g_tuning;
}

int fail(const char* what) {
    std::snprintf(g_error, sizeof(g_error), "%s", what);
    return 1;
}

int fail_cuda(const char* where, cudaError_t e) {
    std::snprintf(g_error, sizeof(g_error), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    return 2;
}

bool mul_fits_int32(long long a, long long b) { return a * b <= 2147483647LL; }

int check_mix_common(const float* x, float* out, const int32_t* frames, int32_t frame_stride, const int32_t* mix,
                     int32_t B, int32_t R, int32_t P) {
    if (x == nullptr || out == nullptr || frames == nullptr || mix == nullptr) return fail("null pointer argument");
    if (B < 0 || R <= 0 || P <= 0) return fail("B must be >= 0 and the cycle shape positive");
    if (frame_stride < 5) return fail("frame_stride must be >= 5");
    if (!mul_fits_int32(R, P)) return fail("a cycle must hold fewer than 2^31 samples");
    // out of place: partners read the original samples, so the two batches must not overlap anywhere
    const uintptr_t bytes = static_cast<uintptr_t>(B) * static_cast<uintptr_t>(R) * static_cast<uintptr_t>(P) * sizeof(float);
    const uintptr_t xa = reinterpret_cast<uintptr_t>(x), oa = reinterpret_cast<uintptr_t>(out);
    if (x == out || (xa < oa + bytes && oa < xa + bytes)) return fail("x and out must not overlap (partners read the original samples)");
    return 0;
}

// Small host->device upload done by SMs reading pinned host memory through UVA instead of by a
// copy engine: the per-step tables (~1 MB) must not queue behind the 164 MB batch copies that a
// prefetching loader keeps in flight on the same host->device engine.
template <typename T>
__global__ void upload_kernel(const T* __restrict__ src, T* __restrict__ dst, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        dst[i] = src[i];
}


// ---- launch overlap (programmatic dependent launch) ------------------------------------------------
// Two consecutive pipelined launches on one stream may overlap (the second fills its pipeline while
// the first drains) when the caller opted in AND the second launch neither reads what the previous
// ones write nor writes what they read or write.  The address ranges of the last two launches per
// stream are kept here for that check; anything else launches with ordinary stream serialisation.
struct Range { uintptr_t lo, hi; };
struct LaunchRecord {
    cudaStream_t stream;
    int device;
    bool valid;
    unsigned long long signature;   // launch geometry of a GPU-filling pipelined grid, 0 otherwise
    Range reads[6];
    int n_reads;
    Range write;
};
bool g_overlap_enabled = false;
long long g_overlap_launches = 0;       // launches issued with the overlap attribute (diagnostics)
std::mutex g_overlap_mutex;
constexpr int kHistorySlots = 16;
LaunchRecord g_history[kHistorySlots][2];   // up to 16 (device, stream) pairs, the last two launches of each
cudaStream_t g_history_stream[kHistorySlots];
int g_history_device[kHistorySlots];
int g_history_used = 0;

int current_device() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

bool intersects(const Range& a, const Range& b) { return a.lo < b.hi && b.lo < a.hi; }

Range range_of(const void* p, size_t bytes) {
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p);
    return Range{lo, p == nullptr ? lo : lo + bytes};
}

LaunchRecord record_of(const pcgmix::MixArgs& a, bool magwarp, cudaStream_t stream) {
    LaunchRecord r{};
    r.stream = stream;
    r.device = current_device();
    r.valid = true;
    const size_t cyc = static_cast<size_t>(a.B) * a.R * a.P * sizeof(float);
    r.write = range_of(a.out, cyc);
    int n = 0;
    r.reads[n++] = range_of(a.x, cyc);
    r.reads[n++] = range_of(a.frames, static_cast<size_t>(a.B) * a.frame_stride * 4);
    r.reads[n++] = range_of(a.mix, static_cast<size_t>(a.B) * 4);
    r.reads[n++] = range_of(a.order, static_cast<size_t>(a.B) * 4);
    if (magwarp) r.reads[n++] = range_of(a.knots, static_cast<size_t>(a.B) * (a.K + 2) * a.R * 8);
    r.n_reads = n;
    return r;
}

bool conflicts(const LaunchRecord& prev, const LaunchRecord& next) {
    if (intersects(prev.write, next.write)) return true;
    for (int i = 0; i < next.n_reads; ++i)
        if (intersects(prev.write, next.reads[i])) return true;
    for (int i = 0; i < prev.n_reads; ++i)
        if (intersects(next.write, prev.reads[i])) return true;
    return false;
}

// Decide whether this launch may overlap the previous one on its stream (buffers disjoint from the last
// two launches, caller opted in) and remember it; returns the previous launch's geometry signature.
bool overlap_decision(const LaunchRecord& now, bool pipelined, unsigned long long* previous_signature) {
    std::lock_guard<std::mutex> lock(g_overlap_mutex);
    int slot = -1;
    for (int i = 0; i < g_history_used; ++i)
        if (g_history_stream[i] == now.stream && g_history_device[i] == now.device) slot = i;
    if (slot < 0) {
        slot = g_history_used < kHistorySlots ? g_history_used++ : 0;
        g_history_stream[slot] = now.stream;
        g_history_device[slot] = now.device;
        g_history[slot][0].valid = g_history[slot][1].valid = false;
    }
    LaunchRecord* h = g_history[slot];
    bool ok = g_overlap_enabled && pipelined && h[0].valid;          // h[0]: previous launch, h[1]: the one before
    if (ok && conflicts(h[0], now)) ok = false;
    if (ok && h[1].valid && conflicts(h[1], now)) ok = false;
    *previous_signature = h[0].valid ? h[0].signature : 0ull;
    h[1] = h[0];
    h[0] = now;
    h[0].valid = false;                                               // until its geometry is known (set_signature)
    h[0].signature = 0ull;
    return ok;
}

void set_signature(cudaStream_t stream, unsigned long long signature, bool overlapped) {
    const int device = current_device();
    std::lock_guard<std::mutex> lock(g_overlap_mutex);
    for (int i = 0; i < g_history_used; ++i) {
        if (g_history_stream[i] != stream || g_history_device[i] != device) continue;
        g_history[i][0].signature = signature;
        g_history[i][0].valid = signature != 0ull;                    // only GPU-filling pipelined grids can be overlapped
        if (overlapped) ++g_overlap_launches;
    }
}

// A launch outside the mix bookkeeping went onto `stream` (an upload of per-step tables, a cut, a segmentation
// pass, ...): whatever it wrote may be an input of the next mix launch, so that launch must be ordered
// normally behind it.  Every entry point that launches anything but a bookkept mix kernel calls this.
void forget_stream(cudaStream_t stream) {
    const int device = current_device();
    std::lock_guard<std::mutex> lock(g_overlap_mutex);
    for (int i = 0; i < g_history_used; ++i)
        if (g_history_stream[i] == stream && g_history_device[i] == device) g_history[i][0].valid = g_history[i][1].valid = false;
}

cudaError_t dispatch_mix(const pcgmix::MixArgs& a, bool magwarp, bool box, cudaStream_t stream) {
    const pcgmix::PipelineTuning g_tuning = tuning_now();
    const bool pipelined = g_tuning.enabled && pcgmix::pipeline_applicable(a, box);
    unsigned long long previous = 0ull;
    const bool allowed = overlap_decision(record_of(a, magwarp, stream), pipelined, &previous);
    if (pipelined) {
        unsigned long long signature = 0ull;
        const cudaError_t e = pcgmix::launch_mix_pipeline(a, magwarp, g_tuning, allowed, previous, stream, &signature);
        set_signature(stream, e == cudaSuccess ? signature : 0ull, allowed && signature != 0ull && signature == previous);
        return e;
    }
    return pcgmix::launch_mix(a, magwarp, box, stream);
}

}  // namespace

extern "C" {

int pcgmix_version(void) { return PCGMIX_B200_VERSION; }

const char* pcgmix_last_error(void) { return g_error; }

int pcgmix_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail_cuda("cudaGetDevice", e);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return fail_cuda("cudaGetDeviceProperties", e);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

int pcgmix_set_tuning(int32_t use_pipeline, int32_t stages, int32_t max_slice, int32_t ctas_per_sm, int32_t pbuf_pct,
                      int32_t consumer_threads, int32_t debug) {
    if (stages < 0 || stages > 8 || max_slice < 0 || (max_slice % 4) != 0 || ctas_per_sm < 0) return fail("bad tuning value");
    std::lock_guard<std::mutex> lock(g_tuning_mutex);
    g_tuning.enabled = use_pipeline ? 1 : 0;
    g_tuning.stages = stages;
    g_tuning.max_slice = max_slice;
    g_tuning.ctas_per_sm = ctas_per_sm;
    g_tuning.pbuf_pct = pbuf_pct;
    g_tuning.consumer_threads = consumer_threads;
#ifdef PCGMIX_PROFILING
    g_tuning.debug = debug & 0xffff;
#else
    // bit 5 (no coefficient table: every item through the producers' per-item path) changes no result and stays
    if (debug & 0xffdf) return fail("the skip switches exist only in a library built with -DPCGMIX_PROFILING");
    g_tuning.debug = debug & 32;
#endif

    return 0;
}

int pcgmix_set_spline_precision(int32_t float32_evaluation) {
    std::lock_guard<std::mutex> lock(g_tuning_mutex);
    g_tuning.spline_f32 = float32_evaluation ? 1 : 0;
    return 0;
}

int pcgmix_get_spline_precision(void) {
    std::lock_guard<std::mutex> lock(g_tuning_mutex);
    return g_tuning.spline_f32;
}

long long pcgmix_overlap_launches(void) { return g_overlap_launches; }

#ifdef PCGMIX_PROFILING
// profiling build only: per-CTA {start, first item consumed, last store done} (globaltimer ns) of the last pipelined launch
int pcgmix_debug_timeline(unsigned long long* host, int32_t n_ctas) {
    const cudaError_t e = pcgmix::read_timeline(host, n_ctas);
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_debug_timeline", e);
}
#endif

int pcgmix_set_launch_overlap(int32_t enable) {
    std::lock_guard<std::mutex> lock(g_overlap_mutex);
    g_overlap_enabled = enable != 0;
    for (int i = 0; i < kHistorySlots; ++i) g_history[i][0].valid = g_history[i][1].valid = false;
    return 0;
}

int pcgmix_copy_small(void* dst, const void* src, int64_t bytes, pcgmix_stream_t stream) {
    if (bytes < 0 || (bytes > 0 && (dst == nullptr || src == nullptr))) return fail("bad copy argument");
    const uintptr_t bits = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | static_cast<uintptr_t>(bytes);
    if (bits & 3u) return fail("copy needs 4-byte aligned pointers and size");
    if (bytes == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    forget_stream(st);
    if ((bits & 15u) == 0) {
        const long long n = bytes / 16;
        const long long cap = bytes > (4ll << 20) ? 1184 : 64;          // large copies: enough loads in flight to fill the link
        const int blocks = static_cast<int>(n / 256 + 1 > cap ? cap : n / 256 + 1);
        upload_kernel<int4><<<blocks, 256, 0, st>>>(static_cast<const int4*>(src), static_cast<int4*>(dst), n);
    } else {
        const long long n = bytes / 4;
        const int blocks = static_cast<int>(n / 256 + 1 > 64 ? 64 : n / 256 + 1);
        upload_kernel<int><<<blocks, 256, 0, st>>>(static_cast<const int*>(src), static_cast<int*>(dst), n);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_copy_small", e);
}

int pcgmix_mix1d(const float* x, float* out, const int32_t* frames, int32_t frame_stride, const int32_t* mix,
                 const int32_t* order, float lam, float one_minus_lam, int32_t B, int32_t C, int32_t L,
                 int32_t* err_flag, pcgmix_stream_t stream) {
    if (int rc = check_mix_common(x, out, frames, frame_stride, mix, B, C, L)) return rc;
    if (B == 0) return 0;
    pcgmix::MixArgs a{};
    a.x = x; a.out = out; a.frames = frames; a.frame_stride = frame_stride; a.mix = mix; a.order = order;
    a.err = err_flag; a.lam = lam; a.one_minus_lam = one_minus_lam; a.B = B; a.R = C; a.P = L; a.F = 1;
    const cudaError_t e = dispatch_mix(a, false, false, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_mix1d", e);
}

int pcgmix_mix1d_magwarp(const float* x, float* out, const int32_t* frames, int32_t frame_stride,
                         const int32_t* mix, const int32_t* order, float lam, float one_minus_lam,
                         const double* knots, const double* coefmat, const double* knot_pos, int32_t K, int32_t B,
                         int32_t C, int32_t L, int32_t* err_flag, pcgmix_stream_t stream) {
    if (int rc = check_mix_common(x, out, frames, frame_stride, mix, B, C, L)) return rc;
    if (knots == nullptr || coefmat == nullptr || knot_pos == nullptr) return fail("null spline argument");
    if (K < 0 || K > PCGMIX_MAX_KNOT) return fail("knot count outside [0, PCGMIX_MAX_KNOT]");
    if (L < 2) return fail("magnitude warp needs at least two samples per row");
    if (B == 0) return 0;
    pcgmix::MixArgs a{};
    a.x = x; a.out = out; a.frames = frames; a.frame_stride = frame_stride; a.mix = mix; a.order = order;
    a.err = err_flag; a.lam = lam; a.one_minus_lam = one_minus_lam; a.B = B; a.R = C; a.P = L; a.F = 1;
    a.knots = knots; a.coefmat = coefmat; a.knot_pos = knot_pos; a.K = K;
    a.inv_h = static_cast<double>(K + 1) / static_cast<double>(L - 1);
    const cudaError_t e = dispatch_mix(a, true, false, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_mix1d_magwarp", e);
}

int pcgmix_mix1d_windows(const float* x, float* out, const int32_t* windows, const int32_t* mix, const int32_t* order,
                         float lam, float one_minus_lam, const double* knots, const double* coefmat,
                         const double* knot_pos, int32_t K, int32_t B, int32_t C, int32_t L, int32_t* err_flag,
                         pcgmix_stream_t stream) {
    if (windows == nullptr) return fail("null pointer argument");
    if (int rc = check_mix_common(x, out, windows, 5, mix, B, C, L)) return rc;
    const bool magwarp = knots != nullptr;
    if (magwarp) {
        if (coefmat == nullptr || knot_pos == nullptr) return fail("null spline argument");
        if (K < 0 || K > PCGMIX_MAX_KNOT) return fail("knot count outside [0, PCGMIX_MAX_KNOT]");
        if (L < 2) return fail("magnitude warp needs at least two samples per row");
    }
    if (B == 0) return 0;
    pcgmix::MixArgs a{};
    a.x = x; a.out = out; a.frames = windows; a.frame_stride = 12; a.windows = windows; a.mix = mix; a.order = order;
    a.err = err_flag; a.lam = lam; a.one_minus_lam = one_minus_lam; a.B = B; a.R = C; a.P = L; a.F = 1;
    if (magwarp) {
        a.knots = knots; a.coefmat = coefmat; a.knot_pos = knot_pos; a.K = K;
        a.inv_h = static_cast<double>(K + 1) / static_cast<double>(L - 1);
    }
    const cudaError_t e = dispatch_mix(a, magwarp, false, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_mix1d_windows", e);
}

int pcgmix_mix2d(const float* x, float* out, const int32_t* frames, int32_t frame_stride, const int32_t* mix,
                 const int32_t* order, float lam, float one_minus_lam, int32_t B, int32_t Ch, int32_t F, int32_t T,
                 const int32_t* tbox, int32_t h1, int32_t h2, int32_t* err_flag, pcgmix_stream_t stream) {
    if (Ch <= 0 || F <= 0) return fail("Ch and F must be positive");
    if (!mul_fits_int32(Ch, F)) return fail("Ch*F too large");
    if (int rc = check_mix_common(x, out, frames, frame_stride, mix, B, Ch * F, T)) return rc;
    if (B == 0) return 0;
    pcgmix::MixArgs a{};
    a.x = x; a.out = out; a.frames = frames; a.frame_stride = frame_stride; a.mix = mix; a.order = order;
    a.err = err_flag; a.lam = lam; a.one_minus_lam = one_minus_lam; a.B = B; a.R = Ch * F; a.P = T;
    a.tbox = tbox; a.F = F; a.h1 = h1 < 0 ? 0 : h1; a.h2 = h2 > F ? F : h2;
    const bool box = a.h1 < a.h2;
    const cudaError_t e = dispatch_mix(a, false, box, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_mix2d", e);
}

int pcgmix_segment_dense(const int8_t* states, int32_t R, int32_t T, int32_t downsample, int32_t* cycles,
                         int32_t max_cycles, int32_t* cycle_count, int32_t* err_flag, pcgmix_stream_t stream) {
    if (cycle_count == nullptr || (R > 0 && T > 0 && states == nullptr)) return fail("null pointer argument");
    if (R < 0 || T < 0 || downsample < 1 || max_cycles < 0) return fail("bad size argument");
    if (max_cycles > 0 && cycles == nullptr) return fail("null cycles with max_cycles > 0");
    if ((reinterpret_cast<uintptr_t>(cycles) & 15u) != 0) return fail("cycles must be 16-byte aligned");
    forget_stream(static_cast<cudaStream_t>(stream));
    const cudaError_t e = pcgmix::launch_segment_dense(states, R, T, downsample, cycles, max_cycles, cycle_count,
                                                       err_flag, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_segment_dense", e);
}

int pcgmix_segment_table(const int32_t* positions, const int8_t* codes, const int32_t* rec_offsets, int32_t R,
                         int32_t downsample, int32_t spec_cols, const int32_t* rec_len, int32_t* cycles,
                         int32_t max_cycles, int32_t* cycle_count, int32_t* err_flag, pcgmix_stream_t stream) {
    if (cycle_count == nullptr || rec_offsets == nullptr) return fail("null pointer argument");
    if (R < 0 || downsample < 1 || max_cycles < 0 || spec_cols < 0) return fail("bad size argument");
    if (spec_cols > 0 && rec_len == nullptr) return fail("rec_len required when spec_cols > 0");
    if (max_cycles > 0 && cycles == nullptr) return fail("null cycles with max_cycles > 0");
    if ((reinterpret_cast<uintptr_t>(cycles) & 15u) != 0) return fail("cycles must be 16-byte aligned");
    forget_stream(static_cast<cudaStream_t>(stream));
    const cudaError_t e = pcgmix::launch_segment_table(positions, codes, rec_offsets, R, downsample, spec_cols,
                                                       rec_len, cycles, max_cycles, cycle_count, err_flag,
                                                       static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_segment_table", e);
}

int pcgmix_cut_cycles(const float* signal, int32_t R, int32_t C, int32_t T, const int32_t* cycles, int32_t n_cycles,
                      const int32_t* n_cycles_dev, float* out, int32_t L, pcgmix_stream_t stream) {
    if (n_cycles < 0 || R < 0 || C < 0 || T < 0 || L < 0) return fail("bad size argument");
    if (n_cycles > 0 && (signal == nullptr || cycles == nullptr || out == nullptr)) return fail("null pointer argument");
    if ((reinterpret_cast<uintptr_t>(cycles) & 15u) != 0) return fail("cycles must be 16-byte aligned");
    forget_stream(static_cast<cudaStream_t>(stream));
    const cudaError_t e = pcgmix::launch_cut_cycles(signal, R, C, T, cycles, n_cycles, n_cycles_dev, out, L,
                                                    static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_cut_cycles", e);
}

int pcgmix_mix1d_resident(const float* signal, int32_t n_rec, int32_t C, int32_t T, const int32_t* cycles,
                          int32_t n_table, const int32_t* sel, const int32_t* mix, const int32_t* order, float lam,
                          float one_minus_lam, const double* knots, const double* coefmat, const double* knot_pos,
                          int32_t K, float* out, int32_t B, int32_t L, int32_t* scratch, int32_t* err_flag,
                          pcgmix_stream_t stream) {
    if (B < 0 || n_rec < 0 || n_table < 0 || C <= 0 || T < 0 || L <= 0) return fail("bad size argument");
    if (B == 0) return 0;
    if (signal == nullptr || cycles == nullptr || mix == nullptr || out == nullptr) return fail("null pointer argument");
    if (((reinterpret_cast<uintptr_t>(cycles) | reinterpret_cast<uintptr_t>(signal)) & 15u) != 0)
        return fail("signal and cycles must be 16-byte aligned");
    if (sel == nullptr && B > n_table) return fail("B exceeds the cycle table and no selection was given");
    if (!mul_fits_int32(C, L)) return fail("a cycle must hold fewer than 2^31 samples");
    const bool magwarp = knots != nullptr;
    if (magwarp) {
        if (coefmat == nullptr || knot_pos == nullptr) return fail("null spline argument");
        if (K < 0 || K > PCGMIX_MAX_KNOT) return fail("knot count outside [0, PCGMIX_MAX_KNOT]");
        if (L < 2) return fail("magnitude warp needs at least two samples per row");
    }
    pcgmix::MixArgs a{};
    a.signal = signal; a.n_rec = n_rec; a.T_sig = T; a.cycles = cycles; a.n_table = n_table; a.sel = sel;
    a.out = out; a.mix = mix; a.order = order; a.err = err_flag; a.lam = lam; a.one_minus_lam = one_minus_lam;
    a.B = B; a.R = C; a.P = L; a.F = 1;
    if (magwarp) {
        a.knots = knots; a.coefmat = coefmat; a.knot_pos = knot_pos; a.K = K;
        a.inv_h = static_cast<double>(K + 1) / static_cast<double>(L - 1);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    const pcgmix::PipelineTuning g_tuning = tuning_now();
    forget_stream(st);                                         // not part of the overlap bookkeeping of the mix launches
    if (scratch != nullptr && (reinterpret_cast<uintptr_t>(scratch) & 15u) == 0 && g_tuning.enabled &&
        pcgmix::pipeline_applicable(a, false)) {
        // slot records first, then the persistent TMA-pipelined kernel reading them in place of `frames`
        e = pcgmix::launch_resolve_resident(a, scratch, st);
        if (e == cudaSuccess) {
            a.frames = scratch;
            a.frame_stride = 8;
            unsigned long long signature = 0ull;
            e = pcgmix::launch_mix_pipeline(a, magwarp, g_tuning, false, 0ull, st, &signature);
        }
    } else {
        e = pcgmix::launch_mix_resident(a, magwarp, st);
    }
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_mix1d_resident", e);
}

int pcgmix_cycle_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C, int32_t L,
                          int32_t channel, int32_t what, float* features, int32_t* err_flag, pcgmix_stream_t stream) {
    if (B < 0 || C <= 0 || L <= 0 || frame_stride < 5 || channel < 0 || channel >= C) return fail("bad size argument");
    if ((what & ~3) != 0 || what == 0) return fail("`what` must be 1 (amplitude), 2 (envelope) or 3 (both)");
    if (B > 0 && (x == nullptr || frames == nullptr || features == nullptr)) return fail("null pointer argument");
    if ((what & 2) && L > 14000) return fail("the envelope block keeps four row-sized arrays in shared memory: L <= 14000");
    forget_stream(static_cast<cudaStream_t>(stream));
    const cudaError_t e = pcgmix::launch_cycle_features(x, frames, frame_stride, B, C, L, channel, what, features, err_flag,
                                                        static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_cycle_features", e);
}

int pcgmix_cycle_psd_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C, int32_t L,
                              int32_t channel, int32_t fs, float* features, int32_t* err_flag, pcgmix_stream_t stream) {
    if (B < 0 || C <= 0 || L <= 0 || frame_stride < 5 || channel < 0 || channel >= C || fs <= 0) return fail("bad size argument");
    if (B > 0 && (x == nullptr || frames == nullptr || features == nullptr)) return fail("null pointer argument");
    forget_stream(static_cast<cudaStream_t>(stream));
    const cudaError_t e = pcgmix::launch_cycle_psd_features(x, frames, frame_stride, B, C, L, channel, fs, features, err_flag,
                                                            static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_cycle_psd_features", e);
}

int pcgmix_cycle_moment_features(const float* x, const int32_t* frames, int32_t frame_stride, int32_t B, int32_t C, int32_t L,
                                 int32_t channel, float* features, int32_t* err_flag, pcgmix_stream_t stream) {
    if (B < 0 || C <= 0 || L <= 0 || frame_stride < 5 || channel < 0 || channel >= C) return fail("bad size argument");
    if (B > 0 && (x == nullptr || frames == nullptr || features == nullptr)) return fail("null pointer argument");
    forget_stream(static_cast<cudaStream_t>(stream));
    const cudaError_t e = pcgmix::launch_cycle_moment_features(x, frames, frame_stride, B, C, L, channel, features, err_flag,
                                                               static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_cycle_moment_features", e);
}

long long pcgmix_first_conv_block_workspace(int32_t C, int32_t F) {
    if (C < 1 || C > 4 || F < 1 || F > PCGMIX_MAX_FIRST_BLOCK_FILTERS) return -1;
    return static_cast<long long>(pcgmix::first_conv_block_workspace_bytes(C, F));
}

int pcgmix_first_conv_block(const float* x, const float* weight, const float* bias, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, float* out, void* workspace, int32_t B, int32_t C,
                            int32_t L, int32_t F, int32_t batch_stats, double eps, double momentum, float* save_mean,
                            float* save_invstd, pcgmix_stream_t stream) {
    if (B < 0 || L <= 0 || F <= 0) return fail("bad size argument");
    if (C < 1 || C > 4) return fail("the first block takes 1 to 4 input channels");
    if (F > PCGMIX_MAX_FIRST_BLOCK_FILTERS) return fail("too many filters (PCGMIX_MAX_FIRST_BLOCK_FILTERS)");
    if (!mul_fits_int32(B, L) || !mul_fits_int32(static_cast<long long>(B) * ((L + 3) / 4), 1)) return fail("B*L must be below 2^31");
    if (!(eps >= 0.0) || !(momentum >= 0.0 && momentum <= 1.0)) return fail("eps must be >= 0 and momentum in [0, 1]");
    if (B > 0 && (x == nullptr || weight == nullptr || out == nullptr || workspace == nullptr)) return fail("null pointer argument");
    if ((reinterpret_cast<uintptr_t>(workspace) & 15u) != 0) return fail("workspace must be 16-byte aligned");
    if (batch_stats == 0 && (running_mean == nullptr || running_var == nullptr))
        return fail("without batch statistics the running statistics are needed");
    if (batch_stats != 0 && static_cast<long long>(B) * L < 2 && B > 0)
        return fail("batch statistics need more than one value per filter");     // torch raises here too
    {
        const uintptr_t xb = reinterpret_cast<uintptr_t>(x), ob = reinterpret_cast<uintptr_t>(out);
        const uintptr_t xn = static_cast<uintptr_t>(B) * C * L * sizeof(float), on = static_cast<uintptr_t>(B) * F * L * sizeof(float);
        if (B > 0 && xb < ob + on && ob < xb + xn) return fail("x and out must not overlap");
    }
    forget_stream(static_cast<cudaStream_t>(stream));
    pcgmix::FirstBlockArgs p;
    p.x = x; p.weight = weight; p.bias = bias; p.gamma = gamma; p.beta = beta;
    p.running_mean = running_mean; p.running_var = running_var; p.out = out; p.workspace = workspace;
    p.save_mean = save_mean; p.save_invstd = save_invstd; p.eps = eps; p.momentum = momentum;
    p.B = B; p.C = C; p.L = L; p.F = F; p.batch_stats = batch_stats != 0 ? 1 : 0;
    const cudaError_t e = pcgmix::launch_first_conv_block(p, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_first_conv_block", e);
}

int pcgmix_duration_features(const int32_t* frames, int32_t frame_stride, int32_t n, int32_t fs, double* features,
                             int32_t* err_flag, pcgmix_stream_t stream) {
    if (n < 0 || fs <= 0 || frame_stride < 5) return fail("bad size argument");
    if (n > 0 && (frames == nullptr || features == nullptr)) return fail("null pointer argument");
    forget_stream(static_cast<cudaStream_t>(stream));
    const cudaError_t e = pcgmix::launch_duration_features(frames, frame_stride, n, fs, features, err_flag,
                                                           static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("pcgmix_duration_features", e);
}

}  // extern "C"
