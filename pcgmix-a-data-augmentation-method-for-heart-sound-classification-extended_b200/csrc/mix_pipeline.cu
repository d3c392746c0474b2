// Persistent, TMA-pipelined PCGmix / PCGmix+ kernel for B200 (sm_100a).
//
// Same arithmetic as mix_kernels.cu (see there for the reference lines it replaces and for the
// numerics rules), different data movement.  The direct-load kernel is bound by latency, not by
// bytes: every CTA first chases order -> partner -> offsets, then waits for its own loads, and
// only a few CTAs fit on an SM.  Here the HBM traffic is decoupled from the arithmetic:
//
//   producer warp   walks the work list (one item = one slice of one row of one cycle), resolves
//                   the item's state windows, and issues cp.async.bulk (TMA bulk) copies into a
//                   ring of shared-memory stages: the row slice itself, plus, for each of the
//                   four heart states, the 16-byte-aligned superset of the partner's window.
//                   Completion is signalled on the stage's "full" mbarrier (expect_tx/complete_tx).
//   consumer warps  wait on "full", blend + warp the slice in place in shared memory (128-bit
//                   shared loads/stores), fence to the async proxy and arrive on the stage's
//                   "computed" mbarrier.  Consumer warps never wait for one another.
//   store warp      waits on "computed", hands the slice to cp.async.bulk for the store to global
//                   memory, waits until that store has finished READING shared memory and releases
//                   the stage to the producer through the "empty" mbarrier.
//
// No register is tied up by a load in flight, so the number of bytes in flight per SM is set by
// the ring depth (stages x CTAs per SM), not by occupancy.  Work is distributed round-robin
// (item i -> CTA i mod grid), so CTAs that run together handle neighbouring slots of the
// processing order and a cycle read as "partner" is still in L2 when it is read as "itself".
//
// Shared-memory layout of a stage (slice_cap floats per slice):
//   xbuf [slice_cap]        the cycle's own samples, overwritten in place with the result
//   pbuf [slice_cap + 32]   partner windows, packed; window s starts at a 16-byte boundary
//   meta                    int4 win[4] = {local start, blended length, pbuf shift, local next
//                           start}, slice geometry, the row's spline coefficients
//
// Requirements (checked by the launcher, which otherwise uses the direct-load kernel): P % 4 == 0,
// 16-byte-aligned tensors, P >= 1024.
//
// RESIDENT variant (pcgmix_mix1d_resident): the cycle's own samples are not a row of a padded
// (B, C, L) tensor but lie somewhere inside a recording (mix_resident.cu explains the layout), at
// an element offset that is in general NOT a multiple of four.  A small kernel first resolves every
// batch slot into an 8-int record {f0..f4, first sample inside the recording, first row of the
// recording, samples available}; the producer reads records instead of `frames`.  Neither kind of
// TMA copy can move such a cycle to column 0 of xbuf (bulk copies and tensor tiles both need a
// 16-byte-aligned global start: benchmarks/tma_probe.cu), so the cycle's own samples travel as 4-byte
// cp.async copies issued by the producer warp — 128 contiguous bytes per warp instruction, no
// register or scoreboard held while they fly — whose completion arrives on the stage's `full`
// barrier (cp.async.mbarrier.arrive.noinc) next to the partner windows' transaction bytes.
// Consumers zero the columns behind the cycle's last sample (padding), blend and warp in place like the padded variant; the slice leaves through the same
// bulk store.  Partner windows are bulk copies of 16-byte-aligned supersets out of the partner's
// recording; a slice whose windows cannot be staged (a window reaches into the partner's padding, a
// superset would run past the end of the tensor, the windows exceed pbuf) reads its partner samples
// one by one from global memory — same result.

#include <mutex>
#include <type_traits>

#include "common.cuh"

namespace pcgmix {

namespace {

// Skip switches (pcgmix_set_tuning's `debug` bits) exist only in a library built with -DPCGMIX_PROFILING: they
// make the kernel produce WRONG output fast and have no business in the shipped build.
#ifdef PCGMIX_PROFILING
#define PCGMIX_SKIP(bit) (pa.debug & (bit))
#else
#define PCGMIX_SKIP(bit) false
#endif

#ifdef PCGMIX_PROFILING
// per-CTA {first instruction, first item consumed, last store done} in globaltimer nanoseconds and the SM it ran on, last launch only
__device__ unsigned long long g_timeline[4 * 2048];
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

constexpr int kMaxStages = 8;
constexpr int kResidentStages = 6;      // RESIDENT default: two CTAs per SM (registers); up to six stages each, fewer when the
                                        // slices are long (the launcher keeps two CTAs per SM)
constexpr int kHeaderBytes = 1024;
constexpr int kProducerWarps = 2;       // producer warps take alternate items (one warp's instruction stream per
                                        // item, ~600 dependent instructions with the spline set-up, was the bound)
constexpr int kResidentProducerWarps = 3;   // RESIDENT: a producer also issues the cycle's own samples as 4-byte copies
// the LAST warps of the CTA are the store warp, then the producers
                                       // (the SM's issue arbiter favours high warp ids; a producer in warp 0
                                       // is starved by consumer warps polling their barriers)
// PCGmix+ on padded batches replaces the second producer warp by the coefficient warp (see the kernel): its
// producers no longer touch the knots, and a 14th warp would cost every thread 8 registers at three CTAs per SM.
// The RESIDENT variant (two CTAs per SM; a producer also issues the cycle's own samples as 4-byte copies) has three.
__host__ __device__ constexpr int producer_warps(int warp_variant, bool resident) {
    return resident ? kResidentProducerWarps : (warp_variant != 0 ? 1 : kProducerWarps);
}
// Coefficient warps (PCGmix+ only) take alternate items like the producers.  One suffices at three CTAs per SM (37
// items per CTA per launch at the benchmark size); the RESIDENT variant's two CTAs per SM serve 55 items each, and one
// warp's dependent chain per item (~1.2 us among the polling consumer warps) was its bound.
__host__ __device__ constexpr int coefficient_warps(int warp_variant, bool resident) {
    return warp_variant == 0 ? 0 : (resident ? 2 : 1);
}
__host__ __device__ constexpr int helper_threads(int warp_variant, bool resident) {
    return 32 + 32 * producer_warps(warp_variant, resident) + 32 * coefficient_warps(warp_variant, resident);
}

struct StageMeta {
    int4 win[4];               // {local start, blended length, shift into pbase, local next start}
    const float* pbase;        // where partner samples are read from: the stage's pbuf, or (window set too
                               // large for pbuf) the partner's row in global memory
    long long out_offset;      // element offset of the slice in x / out
    int nvec;                  // 128-bit vectors in this slice
    int t_beg;                 // first column of the slice
    // RESIDENT only
    int own_n;                 // columns of the slice that hold samples (the rest is padding)
    int par_global;            // 1: partner windows not staged — pbase points into `signal`, guard reads with par_n
    int par_n;                 // not staged: partner columns (slice-local, before the window shift) that hold samples
    int positive;              // 1: the row's warp factor is certainly > 0 and finite at every sample
    int exact;                 // float32 variant: 1 = coef holds float64 coefficients in powers of dt (the row's factor
                               // may come close to zero); 0 = float32 coefficients in powers of u = dt/h
    alignas(16) double coef[kMaxPieces * 4];    // written by the coefficient warp
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// Blocking wait.  try_wait parks the thread in hardware for up to `suspend_ns` before it has to
// be re-issued, so waiting warps do not compete with the producer for issue slots.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t suspend_ns = 2000) {
    uint32_t done = 0;
    const uint32_t addr = smem_addr(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity), "r"(suspend_ns) : "memory");
    }
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// global -> shared asynchronous copy of one float (any 4-byte-aligned pair of addresses)
__device__ __forceinline__ void async_copy_f32(uint32_t dst_smem, const float* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
// one (pre-counted) arrival on `bar` once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void async_copies_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// shared -> global bulk copy; one bulk group per call
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// block until at most `pending` of this thread's bulk stores are still reading shared memory
__device__ __forceinline__ void bulk_store_wait_read(int pending) {
    switch (pending) {
        case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void bulk_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_proxy() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ double int_to_double(int i) {
    return __hiloint2double(0x43300000, i) - 4503599627370496.0;
}

// Float32 variant, rows flagged `exact` by the coefficient warp (a factor that may come closer to zero than 1/16,
// where the float32 evaluation's absolute error of ~3e-7 is no longer negligible relative to it): the consumers
// blend such a row as usual, skip the float32 factor, and this pass then multiplies the thread's vectors in place
// in float64, exactly like variant 1.  At the reference's sigma = 0.2 it runs for about one row in 10^5; it is
// kept out of line so that the hot loop carries none of it.
__device__ __noinline__ void exact_warp_pass(const MixArgs& a, const double* s_kpos, const int* s_kint, const double* coef,
                                             float* xbuf, int nvec, int t_beg, int ct, int nct, int vpt) {
    for (int k = 0; k < vpt; ++k) {
        const int v = ct + k * nct;
        if (v >= nvec) break;
        float4 val = *reinterpret_cast<float4*>(xbuf + v * 4);
        float r[4] = {val.x, val.y, val.z, val.w};
        for (int e = 0; e < 4; ++e) {
            const int t = t_beg + v * 4 + e;
            int pe = min(static_cast<int>(__umulhi(static_cast<unsigned>(t), a.piece_magic)), a.K);
            while (t >= s_kint[pe + 1]) ++pe;
            const double de = __hiloint2double(0x43300000, t) - 4503599627370496.0 - s_kpos[pe];
            const double* c = coef + pe * 4;
            const double wv = fma(fma(fma(c[0], de, c[1]), de, c[2]), de, c[3]);
            r[e] = static_cast<float>(static_cast<double>(r[e]) * wv);
        }
        *reinterpret_cast<float4*>(xbuf + v * 4) = make_float4(r[0], r[1], r[2], r[3]);
    }
}

struct PipeArgs {
    int n_items;               // B * R * slices_per_row (checked < 2^31 by the launcher)
    int step_rest;             // gridDim.x / B  } item -> (slot, rest) advances by these per item,
    int step_slot;             // gridDim.x % B  } so the device never divides
    int stepn_rest;            // the same for kProducerWarps items at once
    int stepn_slot;
    int slices_per_row;
    int slice_len;             // elements per slice (multiple of 4); the last slice of a row may be shorter
    int slice_cap;             // floats reserved for a slice in shared memory (>= slice_len)
    int pbuf_cap;              // floats reserved for the packed partner windows of a slice
    int stages;
    int stage_bytes;
    int header_bytes;          // barriers, knot tables and (PCGmix+) the coefficient matrix and table
    int debug;                 // profiling only: 1 = skip stores, 2 = skip arithmetic, 4 = skip partner copies
};

// WARP: 3 and 4 are 1 and 2 for rows with more than 32 coefficients (knot > 7), a separate instance because of the
// coefficient warp (see there).  0 = PCGmix (no magnitude warp); 1 = PCGmix+ with the warp factor evaluated in float64 and the product
// fp64(sample) * w rounded once to fp32, like the reference's float64 product stored into a float32 array
// (> 99.9 % of samples bit-equal to it); 2 = PCGmix+ with the factor evaluated in float32 (normalised Horner,
// coefficients rounded from float64) and an fp32 product: within 1e-5 relative of the reference (typically
// 3e-7), at the cost of plain PCGmix — no float64 pipe, no conversions (see "PCGmix+ arithmetic" in DESIGN.md
// for the ncu numbers that motivated it).  The absolute error of the float32 factor is ~3e-7, so the relative
// bound needs |w| >= 1/16: the producer checks every row's knots against the bound that guarantees it
// (|w - 1| <= Lambda * max_j |y_j - 1|, Lambda = Lebesgue constant of the knot -> curve map, from knot_pos[K+2]);
// rows outside it (9 % at sigma = 0.2) and the K+1 vectors per row that contain a knot take variant 1's
// float64 path, which is compiled into this variant too.
template <int NCT, int WARP, int VPT, bool RESIDENT>
__global__ void __launch_bounds__(NCT + helper_threads(WARP, RESIDENT), RESIDENT ? 2 : (NCT <= 192 ? 4 : NCT <= 320 ? 3 : 2))
mix_pipeline_kernel(const __grid_constant__ MixArgs a, const __grid_constant__ PipeArgs pa) {
    constexpr bool MAGWARP = WARP != 0;
    constexpr bool F32 = WARP == 2 || WARP == 4;
    constexpr bool WIDE = WARP >= 3;          // knot > 7: more than 32 coefficients per row (lane l owns l, l+32, l+64, l+96)
    constexpr int kThreads = NCT + helper_threads(WARP, RESIDENT);
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);                  // loads of the stage have landed
    uint64_t* computed = full + kMaxStages;                               // consumers are done with the stage
    uint64_t* empty = computed + kMaxStages;                              // the stage's store has left shared memory
    double* s_kpos = reinterpret_cast<double*>(smem + 256);               // kMaxPieces+1 doubles (256 B)
    int* s_kint = reinterpret_cast<int*>(smem + 256 + 256);               // kMaxPieces+1 ints (128 B)
    double* s_mat = reinterpret_cast<double*>(smem + kHeaderBytes);      // [(K+1)*4][K+2] coefficient matrix
    unsigned char* stages = smem + pa.header_bytes;
    const int S = pa.stages;

    // Cold start: the per-step tables (order, partner, offsets, knots) were evicted from L2 by the
    // previous step's 330 MB of traffic.  The whole grid pulls them back with one L2 prefetch per
    // 128-byte line, so the producers' dependent chain order -> partner -> offsets runs on L2 hits.
    {
        const long long gtid = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
        auto warm = [&](const void* base, long long bytes, long long first_line) {
            const long long line = gtid - first_line;
            if (base != nullptr && line >= 0 && line * 128 < bytes)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const char*>(base) + line * 128));
            return first_line + (bytes + 127) / 128;
        };
        long long next = 0;
        next = warm(a.order, static_cast<long long>(a.B) * 4, next);
        next = warm(a.mix, static_cast<long long>(a.B) * 4, next);
        next = warm(a.frames, static_cast<long long>(a.B) * a.frame_stride * 4, next);
        next = warm(a.windows, static_cast<long long>(a.B) * 48, next);
        if constexpr (MAGWARP) next = warm(a.knots, static_cast<long long>(a.B) * (a.K + 2) * a.R * 8, next);
    }
#ifdef PCGMIX_PROFILING
    if (threadIdx.x == 0 && blockIdx.x < 2048) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_timeline[4 * blockIdx.x] = global_ns();
        g_timeline[4 * blockIdx.x + 3] = smid;
    }
#endif
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            // producer (+ the row's coefficients from the coefficient warp) (+ RESIDENT: the producer warp's 32 lanes,
            // each when its share of the cycle's own samples has landed)
            mbar_init(&full[s], (MAGWARP ? 2 : 1) + (RESIDENT ? 32 : 0));
            mbar_init(&computed[s], NCT);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (MAGWARP) {
        const int n_knots = a.K + 2;
        for (int i = threadIdx.x; i < n_knots; i += kThreads) {
            const double kp = __ldg(a.knot_pos + i);
            s_kpos[i] = kp;
            s_kint[i] = (i == n_knots - 1) ? 0x7fffffff : static_cast<int>(ceil(kp));
        }
        // The matrix must sit in shared memory: with ~216 KB of the SM carved out for the stage
        // rings, L1 is a few KB and a per-item walk over the matrix in global memory costs a chain
        // of L2 round trips in the producer (measured: 1.3 us per item, the whole kernel's bound).
        for (int i = threadIdx.x; i < (a.K + 1) * 4 * n_knots; i += kThreads) s_mat[i] = __ldg(a.coefmat + i);
    }
    __syncthreads();

    auto stage_x = [&](int s) { return reinterpret_cast<float*>(stages + static_cast<size_t>(s) * pa.stage_bytes); };
    auto stage_p = [&](int s) { return stage_x(s) + pa.slice_cap; };
    auto stage_meta = [&](int s) { return reinterpret_cast<StageMeta*>(stage_p(s) + pa.pbuf_cap); };
    // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int n_it = (pa.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    // Programmatic dependent launch: as soon as every CTA of this grid is running, the next launch on
    // the stream (if the launcher allowed it to overlap) is made ready; its CTAs take over SMs as ours
    // retire, so its pipeline fills while ours drains.  The launcher only allows this between grids
    // that fill the whole GPU (so at most two of them are ever in flight) and whose buffers are
    // disjoint.  Harmless when nothing depends on us.
    if (threadIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // piece width h = (P-1)/(K+1) in samples; coefficient i of a piece (of dt^(3-i)) times h^(3-i) is the
    // coefficient of u^(3-i), u = dt/h in [0, 1)
    const double h_piece = 1.0 / a.inv_h;

    // item = rest * B + slot; this CTA's items are blockIdx.x, blockIdx.x + gridDim.x, ... (no division on the device)
    struct Cursor { int slot, rest; };
    auto advance1 = [&](Cursor c) {
        c.slot += pa.step_slot;
        c.rest += pa.step_rest;
        if (c.slot >= a.B) {
            c.slot -= a.B;
            ++c.rest;
        }
        return c;
    };
    auto row_of = [&](Cursor c) { return pa.slices_per_row == 1 ? c.rest : c.rest / pa.slices_per_row; };

    if (MAGWARP && threadIdx.x >= NCT + 32 + 32 * producer_warps(WARP, RESIDENT)) {
        // =================================== coefficient warps ===================================
        // Turns every item's K+2 knots into the 4(K+1) cubic coefficients its consumers need (coefficient i =
        // sum_j M[i][j] * knot_j, matrix in shared memory, knots broadcast by shuffle) and writes them into the
        // item's stage, then arrives on the stage's `full` barrier next to the producer.  Nothing else depends
        // on this arithmetic, so it runs beside the copies instead of in front of them: in the producers it sat
        // in the dependent chain ahead of every load (6 us of 62 per launch), as a table built up front by the
        // consumers it delayed the first items of every launch (3 us) — see profiles/README.md.
        const int lane = threadIdx.x & 31;
        const int n_knots = a.K + 2;
        const int n_coef = (a.K + 1) * 4;
        constexpr int NC = coefficient_warps(WARP, RESIDENT);
        const int cw = (threadIdx.x - (NCT + 32 + 32 * producer_warps(WARP, RESIDENT))) >> 5;   // this warp: items cw, cw + NC, ...
        // order[slot] -> knots[cycle] is a chain of two dependent loads: the cycle id is fetched two items ahead and
        // the knots one item ahead, so in steady state this warp never waits for memory.  (An id outside [0, B) is
        // replaced by the slot like everywhere else; the producer warp is the one that raises BAD_PARTNER for it.)
        auto raw_cycle = [&](Cursor c) { return a.order == nullptr ? c.slot : __ldg(a.order + c.slot); };
        auto knot_of = [&](int raw, Cursor c) {             // lane j: knot j of the item's (cycle, row)
            const int b = static_cast<unsigned>(raw) < static_cast<unsigned>(a.B) ? raw : c.slot;
            double y = 0.0;
            if (lane < n_knots && !PCGMIX_SKIP(16)) y = __ldg(a.knots + (static_cast<size_t>(b) * n_knots + lane) * a.R + row_of(c));
            return y;
        };
        // knot_pos[K+2] = 0.999 / Lambda (Lambda: Lebesgue constant of the knot -> curve map): knots within that of 1
        // keep the factor positive.  Rounded DOWN to fp32 and compared as bit patterns (non-negative floats order
        // like unsigned integers; NaN orders above every bound).
        unsigned safe_dev_bits = 0u;
        if constexpr (RESIDENT) safe_dev_bits = __float_as_uint(__double2float_rd(fmax(__ldg(a.knot_pos + a.K + 2), 0.0)));
        // float32 variant: coefficient i of a piece is scaled by h^(3 - (i & 3))
        const int power = 3 - (lane & 3);
        const double scale = power == 3 ? h_piece * h_piece * h_piece : power == 2 ? h_piece * h_piece : power == 1 ? h_piece : 1.0;
        // One item's coefficients -> its stage.  The narrow and the WIDE form are different kernel instances: this warp's
        // dependent chain per item is on the critical path (it shares its scheduler with polling consumer warps), and
        // with both forms in one kernel the narrow one ran 10 % slower (profiles/README.md).
        auto coefficients_of_item = [&](const double y_cur, const int stage, const uint32_t phase) {
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
            const double* mrow = s_mat + ((WIDE || lane < n_coef) ? lane : 0) * n_knots;   // (idle lanes redo row 0; never stored)
            for (int j = 0; j < (PCGMIX_SKIP(8) ? 0 : n_knots); ++j) {
                const double yj = __shfl_sync(kFullMask, y_cur, j);
                if constexpr (WIDE) {
                    if (lane < n_coef) acc0 = fma(mrow[j], yj, acc0);
                    if (lane + 32 < n_coef) acc1 = fma(mrow[32 * n_knots + j], yj, acc1);
                    if (lane + 64 < n_coef) acc2 = fma(mrow[64 * n_knots + j], yj, acc2);
                    if (lane + 96 < n_coef) acc3 = fma(mrow[96 * n_knots + j], yj, acc3);
                } else {
                    acc0 = fma(mrow[j], yj, acc0);
                }
            }
            // float32 variant: coefficients of u = dt/h (lane i: power 3 - (i & 3) of h), and is the factor certainly
            // >= 1/16 on every piece?  A cubic on [0, 1] lies inside the hull of its Bernstein coefficients
            // {d, d + c/3, d + 2c/3 + b/3, a + b + c + d}; lane i evaluates number (i & 3) of its piece.
            int exact = 0;
            float c32[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if constexpr (F32) {
                c32[0] = static_cast<float>(acc0 * scale);
                if constexpr (WIDE) {
                    c32[1] = static_cast<float>(acc1 * scale);
                    c32[2] = static_cast<float>(acc2 * scale);
                    c32[3] = static_cast<float>(acc3 * scale);
                }
                bool low = false;                               // a Bernstein coefficient below 1/16 (or NaN) on this lane
#pragma unroll
                for (int g = 0; g < (WIDE ? 4 : 1); ++g) {
                    const int base = lane & ~3;
                    const float ca = __shfl_sync(kFullMask, c32[g], base), cb = __shfl_sync(kFullMask, c32[g], base + 1);
                    const float cc = __shfl_sync(kFullMask, c32[g], base + 2), cd = __shfl_sync(kFullMask, c32[g], base + 3);
                    const int k = lane & 3;
                    const float bern = k == 0 ? cd : k == 1 ? cd + cc * (1.0f / 3.0f)
                                     : k == 2 ? cd + cc * (2.0f / 3.0f) + cb * (1.0f / 3.0f) : ca + cb + cc + cd;
                    if (lane + 32 * g < n_coef) low = low || !(bern >= 0.0625f);
                }
                exact = __any_sync(kFullMask, low) ? 1 : 0;
                if (PCGMIX_SKIP(64)) exact = 0;                 // (profiling: never take the float64 path)
            }
            // RESIDENT: the curve reproduces constants and is linear in the knots, |w(t) - 1| <= Lambda * max_j |y_j - 1|.
            // Rows inside the bound have a positive, finite factor everywhere; consumers then write padding
            // (exact +0.0f) without evaluating the spline.
            int row_positive = 0;
            if constexpr (RESIDENT) {
                const float dev = lane < n_knots ? fabsf(static_cast<float>(y_cur) - 1.0f) : 0.0f;
                row_positive = __all_sync(kFullMask, __float_as_uint(dev) < safe_dev_bits) ? 1 : 0;
            }
            mbar_wait(&empty[stage], phase ^ 1);            // the stage's previous slice has left shared memory
            StageMeta* meta = stage_meta(stage);
            if (F32 && !exact) {
                float* out32 = reinterpret_cast<float*>(meta->coef);
                if (lane < n_coef) out32[lane] = c32[0];
                if constexpr (WIDE) {
                    if (lane + 32 < n_coef) out32[lane + 32] = c32[1];
                    if (lane + 64 < n_coef) out32[lane + 64] = c32[2];
                    if (lane + 96 < n_coef) out32[lane + 96] = c32[3];
                }
            } else {
                if (lane < n_coef) meta->coef[lane] = acc0;
                if constexpr (WIDE) {
                    if (lane + 32 < n_coef) meta->coef[lane + 32] = acc1;
                    if (lane + 64 < n_coef) meta->coef[lane + 64] = acc2;
                    if (lane + 96 < n_coef) meta->coef[lane + 96] = acc3;
                }
            }
            if (lane == 0) {
                meta->exact = exact;
                meta->positive = row_positive;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        };
        auto advance_c = [&](Cursor c) {                    // to this warp's next item
#pragma unroll
            for (int k = 0; k < NC; ++k) c = advance1(c);
            return c;
        };
        Cursor c0{static_cast<int>(blockIdx.x % a.B), static_cast<int>(blockIdx.x / a.B)};
        for (int k = 0; k < cw; ++k) c0 = advance1(c0);
        Cursor c1 = advance_c(c0);
        int raw1 = n_it > cw + NC ? raw_cycle(c1) : 0;
        double y_next = n_it > cw ? knot_of(raw_cycle(c0), c0) : 0.0;
        int stage = cw % S;
        uint32_t phase = (cw / S) & 1;
        for (int it = cw; it < n_it; it += NC) {
            const double y_cur = y_next;
            const Cursor c2 = advance_c(c1);
            const int raw2 = it + 2 * NC < n_it ? raw_cycle(c2) : 0;
            if (it + NC < n_it) y_next = knot_of(raw1, c1); // both latencies hide behind this item's arithmetic
            raw1 = raw2;
            c1 = c2;
            coefficients_of_item(y_cur, stage, phase);
            stage += NC;                                   // NC <= S is guaranteed by the launcher
            if (stage >= S) {
                stage -= S;
                phase ^= 1;
            }
        }
    } else if (threadIdx.x >= NCT + 32) {
        // ===================================== producer warp =====================================
        const int lane = threadIdx.x & 31;
        const int pw = (threadIdx.x - (NCT + 32)) >> 5;      // this warp handles items pw, pw + kProducerWarps, ...
        // order[slot] -> mix[b] -> frames[partner] is a chain of dependent loads.  It is software-pipelined
        // across items: cycle ids are fetched three items ahead, partner ids two, offsets one, so none of
        // their latency sits in front of the copies.
        auto advance = [&](Cursor c) {                      // to this warp's next item
            c.slot += pa.stepn_slot;
            c.rest += pa.stepn_rest;
            if (c.slot >= a.B) {
                c.slot -= a.B;
                ++c.rest;
            }
            return c;
        };
        auto cycle_of = [&](Cursor c) { return cycle_of_slot(a, c.slot); };
        // per lane s: start of state s in the cycle (f1) and in the partner (f2); with explicit
        // windows ('(rand)' displacement) f1 = window start, f2 = f1 + shift, wn = blended length
        auto load_offsets = [&](int b, int p, int& f1, int& f2, int& wn) {
            if (a.windows != nullptr) {
                if (lane < 4) {
                    const int32_t* w = a.windows + (static_cast<size_t>(b) * 4 + lane) * 3;
                    f1 = __ldg(w);
                    wn = __ldg(w + 1);
                    f2 = f1 + __ldg(w + 2);
                }
            } else if (lane < (RESIDENT ? 8 : 5)) {
                // RESIDENT: lanes 5..7 carry {first sample in the recording, the recording's first row, samples available}
                f1 = __ldg(a.frames + static_cast<size_t>(b) * a.frame_stride + lane);
                f2 = __ldg(a.frames + static_cast<size_t>(p) * a.frame_stride + lane);
            }
        };
        int b0 = 0, b1 = 0, b2 = 0, p0 = 0, p1 = 0, f1 = 0, f2 = 0, wn = 0;
        bool bad0 = false;
        constexpr int NP = producer_warps(WARP, RESIDENT);
        // RESIDENT: the slot records are written by the kernel launched just before this one; everything
        // above (barriers, knot tables, coefficient matrix) did not need them
        if constexpr (RESIDENT) asm volatile("griddepcontrol.wait;" ::: "memory");
        Cursor c0{static_cast<int>(blockIdx.x % a.B), static_cast<int>(blockIdx.x / a.B)};   // item it
        for (int k = 0; k < pw; ++k) c0 = advance1(c0);
        Cursor c1 = advance(c0), c2 = advance(c1), c3 = advance(c2);                          // this warp's next three
        if (n_it > pw) {
            b0 = cycle_of(c0);
            p0 = __ldg(a.mix + b0);
            bad0 = static_cast<unsigned>(p0) >= static_cast<unsigned>(a.B);
            if (bad0) p0 = b0;
            load_offsets(b0, p0, f1, f2, wn);
        }
        if (n_it > pw + NP) {
            b1 = cycle_of(c1);
            p1 = __ldg(a.mix + b1);
        }
        if (n_it > pw + 2 * NP) b2 = cycle_of(c2);
        int stage = pw % S;
        uint32_t phase = (pw / S) & 1;
        for (int it = pw; it < n_it; it += NP) {
            const int rest = c0.rest;
            const int row = row_of(c0);
            const int slice = rest - row * pa.slices_per_row;
            const int b = b0;
            const int p = p0;
            const bool bad_partner = bad0;
            const int f1_cur = f1, f2_cur = f2, wn_cur = wn;
            // prefetch for the items behind this one
            bool bad1 = false;
            if (it + NP < n_it) {
                bad1 = static_cast<unsigned>(p1) >= static_cast<unsigned>(a.B);
                if (bad1) p1 = b1;
                load_offsets(b1, p1, f1, f2, wn);
            }
            int p2 = 0, b3 = 0;
            if (it + 2 * NP < n_it) {
                p2 = __ldg(a.mix + b2);
            }
            if (it + 3 * NP < n_it) b3 = cycle_of(c3);
            b0 = b1; p0 = p1; bad0 = bad1;
            b1 = b2; p1 = p2;
            b2 = b3;
            c0 = c1; c1 = c2; c2 = c3; c3 = advance(c3);

            const int f1n = __shfl_down_sync(kFullMask, f1_cur, 1);
            const int f2n = __shfl_down_sync(kFullMask, f2_cur, 1);
            // window of state `lane`: start column, blended samples, shift to the partner's column (pair_window
            // applies the reference's slice clamping; explicit windows arrive resolved from the host)
            int wstart = f1_cur, n = wn_cur, d = f2_cur - f1_cur;
            const bool ok = (a.windows != nullptr)
                ? ((f1_cur >= 0) & (wn_cur >= 0) & (f1_cur + wn_cur <= ((lane < 3) ? f1n : a.P)) & (f2_cur >= 0) & (f2_cur + wn_cur <= a.P))
                : pair_window(f1_cur, f1n, f2_cur, f2n, a.P, wstart, n, d);
            const unsigned bad_frames = __ballot_sync(kFullMask, (lane < 4) && !ok);
            if (bad_frames != 0u || bad_partner) n = 0;
            const int t_beg = slice * pa.slice_len;
            const int t_end = min(t_beg + pa.slice_len, a.P);
            // the part of state `lane`'s blended window that falls into this slice, as a
            // 16-byte-aligned range [src_lo, src_hi) of the partner's row
            const int w_beg = max(wstart, t_beg);
            const int w_end = min(wstart + n, t_end);
            bool have = (lane < 4) && (w_end > w_beg);
            // RESIDENT: where this slot's and the partner's samples of this channel start inside `signal`
            long long own_first = 0, par_first = 0;
            int own_n = 0, par_n = 0, par_mis = 0;
            bool stageable = true;
            if constexpr (RESIDENT) {
                own_first = static_cast<long long>(__shfl_sync(kFullMask, f1_cur, 6) + row) * a.T_sig + __shfl_sync(kFullMask, f1_cur, 5);
                par_first = static_cast<long long>(__shfl_sync(kFullMask, f2_cur, 6) + row) * a.T_sig + __shfl_sync(kFullMask, f2_cur, 5);
                own_n = __shfl_sync(kFullMask, f1_cur, 7);
                par_n = __shfl_sync(kFullMask, f2_cur, 7);
                par_mis = static_cast<int>((par_first + w_beg + d) & 3);      // misalignment of the window's first partner sample
                // a window that reaches into the partner's padding cannot be staged (the padding is not in memory)
                stageable = __all_sync(kFullMask, !have || w_end + d <= par_n);
            }
            const int src_lo = RESIDENT ? (w_beg + d - par_mis) : ((w_beg + d) & ~3);
            const int src_hi = RESIDENT ? (src_lo + ((w_end + d - src_lo + 3) & ~3)) : ((w_end + d + 3) & ~3);
            if constexpr (RESIDENT) {       // ... nor one whose aligned superset would run past the end of the tensor
                stageable = __all_sync(kFullMask, stageable && !(have && par_first + src_hi > a.n_sig));
            }
            const int cnt = have ? (src_hi - src_lo) : 0;
            int incl = cnt;
#pragma unroll
            for (int s = 1; s < 4; s <<= 1) {
                const int up = __shfl_up_sync(kFullMask, incl, s);
                if (lane >= s) incl += up;
            }
            const int off = incl - cnt;
            int total = __shfl_sync(kFullMask, incl, 3);
            // the slice's own samples (RESIDENT): own_local of them, the first one goes to column 0 of xbuf
            const int own_local = RESIDENT ? min(max(own_n - t_beg, 0), t_end - t_beg) : 0;
            const bool staged = total <= pa.pbuf_cap && !PCGMIX_SKIP(4) && stageable;   // else: consumers read from global memory
            if (!staged) {
                have = false;
                total = 0;
            }

            // the stage's previous slice (item it-S) must have left shared memory
            mbar_wait(&empty[stage], phase ^ 1);

            float* xbuf = stage_x(stage);
            float* pbuf = stage_p(stage);
            StageMeta* meta = stage_meta(stage);
            const long long row_off = (static_cast<long long>(b) * a.R + row) * a.P;
            const long long prow = RESIDENT ? par_first : (static_cast<long long>(p) * a.R + row) * a.P;
            const float* src_base = RESIDENT ? a.signal : a.x;
            if (lane < 4) {
                const int next = (lane < 3) ? min(f1n, a.P) : a.P;
                const int shift = staged ? (t_beg + d - src_lo + off) : d;
                meta->win[lane] = make_int4(wstart - t_beg, n, shift, next - t_beg);
            }
            if (lane == 0) {
                meta->pbase = staged ? pbuf : (src_base + prow + t_beg);
                if constexpr (RESIDENT) {
                    meta->own_n = own_local;
                    meta->par_global = staged ? 0 : 1;
                    meta->par_n = par_n - t_beg;
                }
                meta->out_offset = row_off + t_beg;
                meta->nvec = (t_end - t_beg) >> 2;
                meta->t_beg = t_beg;
                const unsigned bad = (bad_partner ? PCGMIX_ERR_BAD_PARTNER : 0u) | (bad_frames ? PCGMIX_ERR_BAD_FRAMES : 0u);
                if (bad != 0u && rest == 0 && a.err != nullptr) atomicOr(a.err, static_cast<int>(bad));
            }
            __syncwarp();
            if constexpr (RESIDENT) {
                if (lane == 0) mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(total) * 4u);
                {
                    const float* own_g = a.signal + (own_first + t_beg) + lane;
                    uint32_t dst = smem_addr(xbuf) + 4u * lane;
                    int i = lane;
                    for (; i + 96 < own_local; i += 128, own_g += 128, dst += 512u) {      // four copies per round of address arithmetic
                        async_copy_f32(dst, own_g);
                        async_copy_f32(dst + 128u, own_g + 32);
                        async_copy_f32(dst + 256u, own_g + 64);
                        async_copy_f32(dst + 384u, own_g + 96);
                    }
                    for (; i < own_local; i += 32, own_g += 32, dst += 128u) async_copy_f32(dst, own_g);
                    async_copies_arrive(&full[stage]);
                }
                if (have) bulk_load(pbuf + off, a.signal + prow + src_lo, static_cast<uint32_t>(cnt) * 4u, &full[stage]);
            } else {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>((t_end - t_beg) + total) * 4u);
                    bulk_load(xbuf, a.x + row_off + t_beg, static_cast<uint32_t>(t_end - t_beg) * 4u, &full[stage]);
                }
                if (have) bulk_load(pbuf + off, a.x + prow + src_lo, static_cast<uint32_t>(cnt) * 4u, &full[stage]);
            }
            stage += NP;                                   // NP <= S is guaranteed by the launcher
            if (stage >= S) {
                stage -= S;
                phase ^= 1;
            }
        }
    } else if (threadIdx.x >= NCT) {
        // ======================================= store warp ======================================
        if (threadIdx.x == NCT) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < n_it; ++it) {
                mbar_wait(&computed[stage], phase);
                const StageMeta* m = stage_meta(stage);
                if (!PCGMIX_SKIP(1)) bulk_store(a.out + m->out_offset, stage_x(stage), static_cast<uint32_t>(m->nvec) * 16u);
                bulk_store_wait_read(0);                   // the engine has read the slice out of shared memory
                mbar_arrive(&empty[stage]);
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            bulk_store_wait_all();                         // every result is in global memory before exit
#ifdef PCGMIX_PROFILING
            if (blockIdx.x < 2048) g_timeline[4 * blockIdx.x + 2] = global_ns();
#endif
        }
    } else {
        // ===================================== consumer warps ====================================
        const int ct = threadIdx.x;
        // With one slice per row a thread sees the same columns in every item, so which spline piece
        // they fall into (and the offset from the piece's knot) is found once, not once per vector.
        const bool fixed_cols = pa.slices_per_row == 1;
        int fixed_piece[VPT];
        double fixed_dt[F32 ? 1 : VPT];
        float fixed_u[F32 ? VPT : 1];
        const float du = static_cast<float>(a.inv_h);                           // one sample in units of u
        if constexpr (MAGWARP) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int t = min((ct + k * NCT) * 4, a.P - 4);
                int piece = min(static_cast<int>(__umulhi(static_cast<unsigned>(t), a.piece_magic)), a.K);
                while (t >= s_kint[piece + 1]) ++piece;
                const double dt = int_to_double(t) - s_kpos[piece];
                if constexpr (F32) fixed_u[k] = static_cast<float>(dt * a.inv_h); else fixed_dt[k] = dt;
                fixed_piece[k] = (t + 3 < s_kint[piece + 1]) ? piece : -1;      // -1: vector straddles a knot
            }
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < n_it; ++it) {
            float* xbuf = stage_x(stage);
            const StageMeta* meta = stage_meta(stage);
            // Every consumer warp waits on the stage's barrier itself and never on another consumer warp.  (One
            // elected polling warp + a hardware barrier for the rest was tried in round 2 — barrier polling is a
            // third of the executed instructions — and changed nothing on padded batches while costing the
            // consumer-bound RESIDENT variant 7 %: skip switch 128 of the profiling build.)
            if (PCGMIX_SKIP(128)) {
                if (ct < 32) mbar_wait(&full[stage], phase);
                asm volatile("bar.sync 1, %0;" ::"r"(NCT) : "memory");
            } else {
                mbar_wait(&full[stage], phase);
            }
#ifdef PCGMIX_PROFILING
            if (it == 0 && ct == 0 && blockIdx.x < 2048) g_timeline[4 * blockIdx.x + 1] = global_ns();
#endif

            const int lo1 = meta->win[1].x, lo2 = meta->win[2].x, lo3 = meta->win[3].x;
            const int nvec = meta->nvec;
            const int t_beg = meta->t_beg;
            const float* pbase = meta->pbase;
            // float32 variant: a row whose factor may come close to zero (practically never: see the coefficient
            // warp) is blended here and multiplied afterwards, in float64, by exact_warp_pass
            const bool exact = F32 && meta->exact != 0;
            const double* coef = meta->coef;
            const float* coef32 = reinterpret_cast<const float*>(meta->coef);
            const int own_n = RESIDENT ? meta->own_n : 0;
            const bool par_global = RESIDENT && meta->par_global != 0;
            const int par_n = RESIDENT ? meta->par_n : 0;
            const bool positive = RESIDENT && MAGWARP && meta->positive != 0;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int v = ct + k * NCT;
                if (v < nvec && !PCGMIX_SKIP(2)) {
                    const int col = v * 4;                                   // local column
                    float r[4];
                    if constexpr (RESIDENT) {
                        // the cycle's samples start at column 0 (the tensor copy put them there); columns from own_n on
                        // are padding, whatever the last box or the stage's previous slice left there
                        r[0] = r[1] = r[2] = r[3] = 0.0f;
                        if (col < own_n) {
                            const float4 mine = *reinterpret_cast<const float4*>(xbuf + col);
                            r[0] = mine.x;
                            r[1] = col + 1 < own_n ? mine.y : 0.0f;
                            r[2] = col + 2 < own_n ? mine.z : 0.0f;
                            r[3] = col + 3 < own_n ? mine.w : 0.0f;
                        }
                    } else {
                        const float4 mine = *reinterpret_cast<const float4*>(xbuf + col);
                        r[0] = mine.x; r[1] = mine.y; r[2] = mine.z; r[3] = mine.w;
                    }
                    const int s = (col >= lo1) + (col >= lo2) + (col >= lo3);
                    const int4 w = meta->win[s];
                    const int ahead = col - w.x;
                    // padding that no window touches: +0.0f, and +0.0f times a positive finite factor is +0.0f
                    bool padding = RESIDENT && MAGWARP && positive && col >= own_n;
                    if (RESIDENT && __builtin_expect(par_global, 0)) {
                        // partner windows not staged: sample by sample out of the partner's recording, zeros behind its end
                        padding = false;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int t = col + e;
                            const int se = (t >= lo1) + (t >= lo2) + (t >= lo3);
                            const int4 we = meta->win[se];
                            if (static_cast<unsigned>(t - we.x) < static_cast<unsigned>(we.y)) {
                                const float o = t + we.z < par_n ? __ldg(pbase + t + we.z) : 0.0f;
                                r[e] = __fadd_rn(__fmul_rn(r[e], a.lam), __fmul_rn(o, a.one_minus_lam));
                            }
                        }
                    } else if (__builtin_expect(ahead >= 0 && col + 3 < w.w, 1)) {
                        const int m = w.y - ahead;                           // leading samples that blend
                        if (m > 0) {
                            padding = false;
                            const float* src = pbase + col + w.z;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (e < m) r[e] = __fadd_rn(__fmul_rn(r[e], a.lam), __fmul_rn(src[e], a.one_minus_lam));
                            }
                        }
                    } else {
                        padding = false;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int t = col + e;
                            const int se = (t >= lo1) + (t >= lo2) + (t >= lo3);
                            const int4 we = meta->win[se];
                            if (static_cast<unsigned>(t - we.x) < static_cast<unsigned>(we.y))
                                r[e] = __fadd_rn(__fmul_rn(r[e], a.lam), __fmul_rn(pbase[t + we.z], a.one_minus_lam));
                        }
                    }
                    if (RESIDENT && MAGWARP && padding) {
                        *reinterpret_cast<float4*>(xbuf + col) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        continue;
                    }
                    if constexpr (F32) {
                        if (!exact) {
                            const int t = t_beg + col;
                            int piece;
                            float u0;
                            if (fixed_cols) {
                                piece = fixed_piece[k];
                                u0 = fixed_u[k];
                            } else {
                                piece = min(static_cast<int>(__umulhi(static_cast<unsigned>(t), a.piece_magic)), a.K);
                                while (t >= s_kint[piece + 1]) ++piece;
                                u0 = static_cast<float>((int_to_double(t) - s_kpos[piece]) * a.inv_h);
                                if (!(t + 3 < s_kint[piece + 1])) piece = -1;
                            }
                            if (__builtin_expect(piece >= 0, 1)) {
                                const float4 c = *reinterpret_cast<const float4*>(coef32 + piece * 4);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float u = fmaf(static_cast<float>(e), du, u0);
                                    r[e] = __fmul_rn(r[e], fmaf(fmaf(fmaf(c.x, u, c.y), u, c.z), u, c.w));
                                }
                            } else {                        // the vector contains a knot: piece per sample
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int te = t + e;
                                    int pe = min(static_cast<int>(__umulhi(static_cast<unsigned>(te), a.piece_magic)), a.K);
                                    while (te >= s_kint[pe + 1]) ++pe;
                                    const float u = static_cast<float>((int_to_double(te) - s_kpos[pe]) * a.inv_h);
                                    const float4 c = *reinterpret_cast<const float4*>(coef32 + pe * 4);
                                    r[e] = __fmul_rn(r[e], fmaf(fmaf(fmaf(c.x, u, c.y), u, c.z), u, c.w));
                                }
                            }
                        }
                    } else if constexpr (MAGWARP) {
                        const int t = t_beg + col;
                        int piece;
                        double dt;
                        if (fixed_cols) {
                            piece = fixed_piece[k];
                            dt = fixed_dt[k];
                        } else {
                            piece = min(static_cast<int>(__umulhi(static_cast<unsigned>(t), a.piece_magic)), a.K);
                            while (t >= s_kint[piece + 1]) ++piece;
                            dt = int_to_double(t) - s_kpos[piece];
                            if (!(t + 3 < s_kint[piece + 1])) piece = -1;
                        }
                        if (__builtin_expect(piece >= 0, 1)) {
                            const double2 c01 = *reinterpret_cast<const double2*>(&coef[piece * 4]);
                            const double2 c23 = *reinterpret_cast<const double2*>(&coef[piece * 4 + 2]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const double de = dt + static_cast<double>(e);
                                const double wv = fma(fma(fma(c01.x, de, c01.y), de, c23.x), de, c23.y);
                                r[e] = static_cast<float>(static_cast<double>(r[e]) * wv);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int te = t + e;
                                int pe = min(static_cast<int>(__umulhi(static_cast<unsigned>(te), a.piece_magic)), a.K);
                                while (te >= s_kint[pe + 1]) ++pe;
                                const double de = int_to_double(te) - s_kpos[pe];
                                const double* c = &coef[pe * 4];
                                const double wv = fma(fma(fma(c[0], de, c[1]), de, c[2]), de, c[3]);
                                r[e] = static_cast<float>(static_cast<double>(r[e]) * wv);
                            }
                        }
                    }
                    *reinterpret_cast<float4*>(xbuf + col) = make_float4(r[0], r[1], r[2], r[3]);
                }
            }
            if constexpr (F32) {
                if (__builtin_expect(exact, 0)) exact_warp_pass(a, s_kpos, s_kint, coef, xbuf, nvec, t_beg, ct, NCT, VPT);
            }
            fence_async_proxy();                           // my shared-memory writes -> visible to the TMA engine
            mbar_arrive(&computed[stage]);                 // (one arrival per warp instead of per thread: measured, no gain)
            if (++stage == S) {
                stage = 0;
                phase ^= 1;
            }
        }
    }
}

constexpr int kMaxDevices = 64;

// Per-device facts, filled on first use (a process may drive several GPUs: the reference trains with
// nn.DataParallel, train_model.py:385).
struct DeviceFacts {
    int sm_count;
};
std::mutex g_device_mutex;
DeviceFacts g_device[kMaxDevices] = {};

cudaError_t device_facts(int* device, DeviceFacts* facts) {
    cudaError_t e = cudaGetDevice(device);
    if (e != cudaSuccess) return e;
    if (*device < 0 || *device >= kMaxDevices) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_device_mutex);
    if (g_device[*device].sm_count == 0) {
        e = cudaDeviceGetAttribute(&g_device[*device].sm_count, cudaDevAttrMultiProcessorCount, *device);
        if (e != cudaSuccess) return e;
    }
    *facts = g_device[*device];
    return cudaSuccess;
}

// What a launch still has to decide once the kernel instance (and with it the true CTA residency) is known.
struct GridPlan {
    int device;
    int sm_count;
    int per_sm_wanted;          // residency the shared-memory / thread budget allows (and the tuning cap)
    bool overlap_previous;
    unsigned long long previous_signature;
    unsigned long long* signature_out;
};

template <int NCT, int MAGWARP, int VPT, bool RESIDENT>
cudaError_t launch_instance(const MixArgs& a, PipeArgs pa, size_t smem, GridPlan plan, cudaStream_t stream) {
    auto kernel = mix_pipeline_kernel<NCT, MAGWARP, VPT, RESIDENT>;
    // Per device and kernel instance: opt in to the large dynamic shared-memory carve-out (the attribute is
    // per device) and ask the runtime how many CTAs of this instance really fit on an SM at this size.
    struct Cached { bool opted_in; size_t smem; int ctas; };
    static Cached cache[kMaxDevices] = {};
    static std::mutex cache_mutex;
    int resident_ctas = 0;
    {
        std::lock_guard<std::mutex> lock(cache_mutex);
        Cached& c = cache[plan.device];
        if (!c.opted_in) {
            const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return e;
            c.opted_in = true;
        }
        if (c.smem != smem || c.ctas == 0) {
            int n = 0;
            const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, NCT + helper_threads(MAGWARP, RESIDENT), smem);
            if (e != cudaSuccess) return e;
            if (n < 1) return cudaErrorInvalidConfiguration;
            c.smem = smem;
            c.ctas = n;
        }
        resident_ctas = c.ctas;
    }
    const int per_sm = plan.per_sm_wanted < resident_ctas ? plan.per_sm_wanted : resident_ctas;
    long long grid = static_cast<long long>(plan.sm_count) * per_sm;
    // Geometry signature: non-zero only for a grid that fills the GPU exactly (as many CTAs as can be resident,
    // counted by the runtime from registers, threads and shared memory).  Two launches with EQUAL signatures
    // have the same CTA footprint and count, so the second can only become fully resident after the first has
    // fully retired: at most two such grids are ever in flight.  Overlap is only allowed then.
    const unsigned long long signature =
        (grid <= pa.n_items && per_sm == resident_ctas)
            ? ((static_cast<unsigned long long>(smem) << 32) ^ (static_cast<unsigned long long>(NCT) << 20) ^
               (static_cast<unsigned long long>(MAGWARP) << 30) ^ static_cast<unsigned long long>(grid)) : 0ull;
    if (plan.signature_out != nullptr) *plan.signature_out = signature;
    bool overlap = plan.overlap_previous;
    if (!RESIDENT && (signature == 0ull || signature != plan.previous_signature)) overlap = false;
    if (grid > pa.n_items) grid = pa.n_items;
    pa.step_rest = static_cast<int>(grid / a.B);
    pa.step_slot = static_cast<int>(grid % a.B);
    pa.stepn_rest = static_cast<int>((grid * producer_warps(MAGWARP, RESIDENT)) / a.B);
    pa.stepn_slot = static_cast<int>((grid * producer_warps(MAGWARP, RESIDENT)) % a.B);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(NCT + helper_threads(MAGWARP, RESIDENT));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, a, pa);
}

template <int NCT, int VPT, bool RESIDENT>
cudaError_t launch_nct(const MixArgs& a, const PipeArgs& pa, size_t smem, int warp, const GridPlan& plan, cudaStream_t stream) {
    return warp == 4 ? launch_instance<NCT, 4, VPT, RESIDENT>(a, pa, smem, plan, stream)
         : warp == 3 ? launch_instance<NCT, 3, VPT, RESIDENT>(a, pa, smem, plan, stream)
         : warp == 2 ? launch_instance<NCT, 2, VPT, RESIDENT>(a, pa, smem, plan, stream)
         : warp == 1 ? launch_instance<NCT, 1, VPT, RESIDENT>(a, pa, smem, plan, stream)
                     : launch_instance<NCT, 0, VPT, RESIDENT>(a, pa, smem, plan, stream);
}

}  // namespace

#ifdef PCGMIX_PROFILING
cudaError_t read_timeline(unsigned long long* host, int n_ctas) {
    return cudaMemcpyFromSymbol(host, g_timeline, sizeof(unsigned long long) * 4 * static_cast<size_t>(n_ctas > 2048 ? 2048 : n_ctas));
}
#endif

bool pipeline_applicable(const MixArgs& a, bool box) {
    const void* src = a.signal != nullptr ? static_cast<const void*>(a.signal) : static_cast<const void*>(a.x);
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(a.out)) & 15u) == 0;
    return !box && aligned16 && (a.P % 4) == 0 && a.P >= 1024;
}

cudaError_t launch_mix_pipeline(const MixArgs& base, bool magwarp, const PipelineTuning& tune, bool overlap_previous,
                                unsigned long long previous_signature, cudaStream_t stream,
                                unsigned long long* full_grid_signature) {
    MixArgs a = base;
    a.n_per_cycle = a.R * a.P;
    const bool resident = a.signal != nullptr;               // records in a.frames (stride 8), samples in a.signal
    if (resident) a.n_sig = static_cast<long long>(a.n_rec) * a.R * a.T_sig;
    if (magwarp) {
        const double ratio = static_cast<double>(a.K + 1) / static_cast<double>(a.P - 1);
        const double scaled = ratio * 4294967296.0 * (1.0 - 1e-9);
        a.piece_magic = scaled >= 4294967295.0 ? 4294967295u : static_cast<unsigned>(scaled);
    }
    int device = 0;
    DeviceFacts facts{};
    if (const cudaError_t e = device_facts(&device, &facts)) return e;
    const int g_sm_count = facts.sm_count;
    // kernel variant, see mix_pipeline_kernel: +2 when a row has more than 32 coefficients
    const int warp = magwarp ? (tune.spline_f32 ? 2 : 1) + ((a.K + 1) * 4 > 32 ? 2 : 0) : 0;
    PipeArgs pa{};
    const int max_slice = tune.max_slice > 0 ? (tune.max_slice > 3584 ? 3584 : tune.max_slice) : 3584;   // 448 threads x 2 vectors
    pa.slices_per_row = (a.P + max_slice - 1) / max_slice;
    int slice_len = (a.P + pa.slices_per_row - 1) / pa.slices_per_row;
    slice_len = (slice_len + 3) & ~3;
    pa.slice_len = slice_len;
    pa.slice_cap = (slice_len + 31) & ~31;
    // packed partner windows: physiological cycles blend < 2/3 of a padded row; anything larger
    // falls back to reading the partner from global memory (still exact, just not staged)
    const int pbuf_pct = tune.pbuf_pct > 0 ? tune.pbuf_pct : 62;
    pa.pbuf_cap = ((slice_len * pbuf_pct / 100 + 32) + 31) & ~31;
    pa.stages = tune.stages > 0 ? (tune.stages > kMaxStages ? kMaxStages : tune.stages) : (resident ? kResidentStages : 4);
    if (pa.stages < kResidentProducerWarps) pa.stages = kResidentProducerWarps;
    // only the (K+1)*4 coefficients in use are reserved at the end of the metadata block
    const size_t meta_bytes = sizeof(StageMeta) - sizeof(double) * (kMaxPieces * 4 - (magwarp ? (a.K + 1) * 4 : 0));
    const size_t stage_bytes = (static_cast<size_t>(pa.slice_cap) + pa.pbuf_cap) * sizeof(float) + meta_bytes;
    pa.stage_bytes = static_cast<int>((stage_bytes + 127) & ~static_cast<size_t>(127));
    const size_t mat_bytes = magwarp ? static_cast<size_t>(a.K + 1) * 4 * (a.K + 2) * sizeof(double) : 0;
    const long long n_items = static_cast<long long>(a.B) * a.R * pa.slices_per_row;
    if (n_items > 2147483647LL) return cudaErrorInvalidConfiguration;
    pa.n_items = static_cast<int>(n_items);
    pa.header_bytes = static_cast<int>((kHeaderBytes + mat_bytes + 127) & ~static_cast<size_t>(127));
    if (resident && tune.stages <= 0) {
        // RESIDENT default: as many stages (at most kResidentStages) as still let two CTAs share an SM
        const size_t per_cta = (228 * 1024) / 2 - 1024 - pa.header_bytes;
        const int fit = static_cast<int>(per_cta / static_cast<size_t>(pa.stage_bytes));
        pa.stages = fit < kResidentProducerWarps ? kResidentProducerWarps : (fit < kResidentStages ? fit : kResidentStages);
    }
    const size_t smem = pa.header_bytes + static_cast<size_t>(pa.stages) * pa.stage_bytes;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    pa.debug = tune.debug;                                  // (the kernel reads it only in a PCGMIX_PROFILING build)
    const int vpt = 2;                                                        // 128-bit vectors per consumer thread and slice
    const int need = ((slice_len / 4) + vpt - 1) / vpt;                       // consumer threads with work
    if (need > 448) return cudaErrorInvalidConfiguration;
    int nct = need <= 128 ? 128 : need <= 192 ? 192 : need <= 256 ? 256 : need <= 320 ? 320 : need <= 384 ? 384 : 448;
    if (tune.consumer_threads > nct) {                                   // more (lighter) consumer threads per slice
        const int want = tune.consumer_threads;
        nct = want <= 192 ? 192 : want <= 256 ? 256 : want <= 320 ? 320 : want <= 384 ? 384 : 448;
    }
    int per_sm = static_cast<int>((228 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : per_sm;
    const int by_threads = 2048 / (nct + helper_threads(warp, resident));
    per_sm = per_sm < by_threads ? per_sm : by_threads;
    if (tune.ctas_per_sm > 0 && tune.ctas_per_sm < per_sm) per_sm = tune.ctas_per_sm;
    if (pa.stages < kResidentProducerWarps) return cudaErrorInvalidConfiguration;
    GridPlan plan{device, g_sm_count, per_sm, overlap_previous, previous_signature, full_grid_signature};
    if (resident) {
        if (static_cast<long long>(a.n_rec) * a.R > 2147483647LL) return cudaErrorInvalidConfiguration;   // the records hold rows as int32
        // launched with the programmatic attribute: the kernel right before it on the stream is the slot-record
        // kernel, which lets it start early; the producers wait for the records (griddepcontrol.wait)
        plan.overlap_previous = true;
        switch (nct) {
            case 128: return launch_nct<128, 2, true>(a, pa, smem, warp, plan, stream);
            case 192: return launch_nct<192, 2, true>(a, pa, smem, warp, plan, stream);
            case 256: return launch_nct<256, 2, true>(a, pa, smem, warp, plan, stream);
            case 320: return launch_nct<320, 2, true>(a, pa, smem, warp, plan, stream);
            case 384: return launch_nct<384, 2, true>(a, pa, smem, warp, plan, stream);
            default: return launch_nct<448, 2, true>(a, pa, smem, warp, plan, stream);
        }
    }
    switch (nct) {
        case 128: return launch_nct<128, 2, false>(a, pa, smem, warp, plan, stream);
        case 192: return launch_nct<192, 2, false>(a, pa, smem, warp, plan, stream);
        case 256: return launch_nct<256, 2, false>(a, pa, smem, warp, plan, stream);
        case 320: return launch_nct<320, 2, false>(a, pa, smem, warp, plan, stream);
        case 384: return launch_nct<384, 2, false>(a, pa, smem, warp, plan, stream);
        default: return launch_nct<448, 2, false>(a, pa, smem, warp, plan, stream);
    }
}

}  // namespace pcgmix
