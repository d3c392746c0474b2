// Host-side replay of CPython's random.Random(seed).sample(pool, len(pool)) for the pairing draw.
//
// The reference pairs cycles with `random.Random(step).sample(idx, len(idx))` per group
// (augmentations.py:500-514, 528-556).  Doing that in Python costs ~2.3 ms per 4096 cycles and is
// the largest host cost of a step once the samples themselves no longer leave the GPU; this is the
// same algorithm in C++ (tests hold it equal to CPython's `random` for thousands of seeds/sizes):
//   * seeding: MT19937 init_by_array with the 32-bit words of |seed| (CPython random_seed, version 2)
//   * getrandbits(k), k <= 32: genrand_uint32() >> (32 - k)
//   * _randbelow(n): k = n.bit_length(); draw until r < n
//   * sample(population, k = n): pool algorithm — result[i] = pool[j]; pool[j] = pool[n-i-1]
//     (the set-based branch is never taken for k == n: setsize = 21 + 4^ceil(log4(3k)) > n)
//
// Second replay in this file: NumPy's LEGACY global stream, which the reference uses for lambda and for the
// magnitude-warp knots (augmentations.py:659-666, :677):
//     np.random.seed(step); lam = np.random.beta(alpha, alpha); knots = np.random.normal(1.0, sigma, size)
// i.e. RandomState's MT19937 seeded with init_genrand(step), 53-bit doubles (a>>5, b>>6), Johnk's / the
// Marsaglia-Tsang gamma beta, and the polar Box-Muller gauss with its cached second variate — the
// algorithms NumPy froze for RandomState (numpy/random/src/legacy/legacy-distributions.c).  `log`, `pow`
// and `sqrt` are the C library's, as in NumPy, and this file is compiled without FMA contraction, so
// the values are bit-equal to NumPy's (tests/test_host_logic.py holds them equal for thousands of seeds).
// 98 304 knots per 4096-cycle step took 1.0-3.7 ms in NumPy; here the raw words, the rejection scan and the
// log/sqrt arithmetic are separate passes and the last one is spread over a few threads.
// No CUDA here; plain host code in the same shared library.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "pcgmix_b200.h"

namespace {

struct Mt19937 {
    uint32_t mt[624];
    int index;

    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + static_cast<uint32_t>(i);
        index = 624;
    }
    void init_by_array(const uint32_t* key, int len) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = (624 > len ? 624 : len); k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + static_cast<uint32_t>(j);
            ++i; ++j;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
            if (j >= len) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - static_cast<uint32_t>(i);
            ++i;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    void seed(uint64_t s) {
        uint32_t key[2] = {static_cast<uint32_t>(s & 0xffffffffu), static_cast<uint32_t>(s >> 32)};
        init_by_array(key, key[1] != 0 ? 2 : 1);
    }
    void twist() {                                         // next block of 624 state words
        int k = 0;
        for (; k < 624 - 397; ++k) {
            const uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
            mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; k < 623; ++k) {
            const uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
            mt[k] = mt[k - 227] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    static uint32_t temper(uint32_t y) {
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    uint32_t next() {
        if (index >= 624) {
            twist();
            index = 0;
        }
        return temper(mt[index++]);
    }
    uint32_t randbelow(uint32_t n) {                       // 1 <= n < 2^31
        int bits = 0;
        for (uint32_t v = n; v; v >>= 1) ++bits;
        uint32_t r = next() >> (32 - bits);
        while (r >= n) r = next() >> (32 - bits);
        return r;
    }
};

// NumPy's RandomState on top of the same generator.
struct LegacyStream {
    Mt19937 mt;
    bool has_gauss = false;
    double gauss = 0.0;

    void seed(uint32_t s) {                                // RandomState.seed(int): mt19937_seed == init_genrand
        mt.init_genrand(s);
        has_gauss = false;
        gauss = 0.0;
    }
    double next_double() {
        const uint32_t a = mt.next() >> 5, b = mt.next() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    double next_gauss() {
        if (has_gauss) {
            const double t = gauss;
            has_gauss = false;
            gauss = 0.0;
            return t;
        }
        double x1, x2, r2;
        do {
            x1 = 2.0 * next_double() - 1.0;
            x2 = 2.0 * next_double() - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        const double f = std::sqrt(-2.0 * std::log(r2) / r2);
        gauss = f * x1;
        has_gauss = true;
        return f * x2;
    }
    double standard_exponential() { return -std::log(1.0 - next_double()); }
    double standard_gamma(double shape) {
        if (shape == 1.0) return standard_exponential();
        if (shape == 0.0) return 0.0;
        if (shape < 1.0) {
            for (;;) {
                const double u = next_double();
                const double v = standard_exponential();
                if (u <= 1.0 - shape) {
                    const double x = std::pow(u, 1.0 / shape);
                    if (x <= v) return x;
                } else {
                    const double y = -std::log((1.0 - u) / shape);
                    const double x = std::pow(1.0 - shape + shape * y, 1.0 / shape);
                    if (x <= v + y) return x;
                }
            }
        }
        const double b = shape - 1.0 / 3.0;
        const double c = 1.0 / std::sqrt(9.0 * b);
        for (;;) {
            double x, v;
            do {
                x = next_gauss();
                v = 1.0 + c * x;
            } while (v <= 0.0);
            v = v * v * v;
            const double u = next_double();
            if (u < 1.0 - 0.0331 * (x * x) * (x * x)) return b * v;
            if (std::log(u) < 0.5 * x * x + b * (1.0 - v + std::log(v))) return b * v;
        }
    }
    double beta(double a, double b) {
        if (a <= 1.0 && b <= 1.0) {                        // Johnk's algorithm
            for (;;) {
                const double u = next_double();
                const double v = next_double();
                const double x = std::pow(u, 1.0 / a);
                const double y = std::pow(v, 1.0 / b);
                const double xpy = x + y;
                if (xpy <= 1.0 && u + v > 0.0) {
                    if (xpy > 0.0) return x / xpy;
                    double log_x = std::log(u) / a;
                    double log_y = std::log(v) / b;
                    const double log_m = log_x > log_y ? log_x : log_y;
                    log_x -= log_m;
                    log_y -= log_m;
                    return std::exp(log_x - std::log(std::exp(log_x) + std::exp(log_y)));
                }
            }
        }
        const double ga = standard_gamma(a);
        const double gb = standard_gamma(b);
        return ga / (ga + gb);
    }

    // n normal(loc, scale) variates, same values and same final state as n calls of next_gauss(), in three
    // passes.  (1) The rejection scan: every attempt of the polar method consumes exactly four words, so
    // attempt j of a block sits at a fixed position whatever was accepted before it — x1, x2 and r2 of all
    // attempts of a 624-word block are computed in branch-free loops and the accepted ones compacted.
    // (2) f = sqrt(-2 log r2 / r2) for the accepted attempts (independent: spread over threads).  (3) The
    // variates in NumPy's order (f*x2 first, f*x1 is the cached second one).
    void fill_normal(double loc, double scale, int64_t n, double* out, int max_threads) {
        int64_t done = 0;
        if (n > 0 && has_gauss) out[done++] = loc + scale * next_gauss();
        const int64_t pairs = (n - done + 1) / 2;
        if (pairs > 0) {
            std::vector<double> x1(static_cast<size_t>(pairs) + 160), x2(static_cast<size_t>(pairs) + 160),
                f(static_cast<size_t>(pairs) + 160);
            uint32_t buf[640];                             // words generated and not yet consumed (tail of the current block)
            double xa[160], xb[160], rr[160];
            int have = 0;
            for (int k = mt.index; k < 624; ++k) buf[have++] = Mt19937::temper(mt.mt[k]);
            int64_t c = 0;
            while (c < pairs) {
                if (have < 4) {                            // fewer than one attempt left: next block behind the leftover
                    mt.twist();
                    for (int k = 0; k < 624; ++k) buf[have + k] = Mt19937::temper(mt.mt[k]);
                    have += 624;
                }
                const int na = have / 4;
                for (int j = 0; j < na; ++j) {
                    const double d1 = ((buf[4 * j] >> 5) * 67108864.0 + (buf[4 * j + 1] >> 6)) / 9007199254740992.0;
                    const double d2 = ((buf[4 * j + 2] >> 5) * 67108864.0 + (buf[4 * j + 3] >> 6)) / 9007199254740992.0;
                    const double a = 2.0 * d1 - 1.0, b = 2.0 * d2 - 1.0;
                    xa[j] = a;
                    xb[j] = b;
                    rr[j] = a * a + b * b;
                }
                int accepted = 0;
                for (int j = 0; j < na; ++j) accepted += (rr[j] < 1.0) & (rr[j] != 0.0);
                int used;
                if (c + accepted < pairs) {                // the whole block is needed: branch-free compaction
                    for (int j = 0; j < na; ++j) {
                        x1[static_cast<size_t>(c)] = xa[j];
                        x2[static_cast<size_t>(c)] = xb[j];
                        f[static_cast<size_t>(c)] = rr[j];
                        c += (rr[j] < 1.0) & (rr[j] != 0.0);
                    }
                    used = na;
                } else {                                   // the draw ends inside this block: stop at the exact attempt
                    int j = 0;
                    for (; c < pairs; ++j) {
                        if ((rr[j] < 1.0) & (rr[j] != 0.0)) {
                            x1[static_cast<size_t>(c)] = xa[j];
                            x2[static_cast<size_t>(c)] = xb[j];
                            f[static_cast<size_t>(c)] = rr[j];
                            ++c;
                        }
                    }
                    used = j;
                }
                have -= used * 4;
                std::memmove(buf, buf + used * 4, static_cast<size_t>(have) * sizeof(uint32_t));
            }
            mt.index = 624 - have;                         // the unconsumed words are the tail of the current block
            auto work = [&](int64_t lo, int64_t hi) {
                for (int64_t i = lo; i < hi; ++i) {
                    const double r2 = f[static_cast<size_t>(i)];
                    f[static_cast<size_t>(i)] = std::sqrt(-2.0 * std::log(r2) / r2);
                }
            };
            int threads = static_cast<int>(pairs / 8192);              // one thread per ~8k pairs, at least the caller's
            if (threads > max_threads) threads = max_threads;
            if (threads <= 1) {
                work(0, pairs);
            } else {
                std::vector<std::thread> pool;
                const int64_t chunk = (pairs + threads - 1) / threads;
                for (int t = 1; t < threads; ++t) pool.emplace_back(work, t * chunk, std::min<int64_t>(pairs, (t + 1) * chunk));
                work(0, std::min<int64_t>(pairs, chunk));
                for (auto& th : pool) th.join();
            }
            for (int64_t i = 0; i < pairs; ++i) {
                const double fi = f[static_cast<size_t>(i)];
                out[done++] = loc + scale * (fi * x2[static_cast<size_t>(i)]);
                if (done < n) {
                    out[done++] = loc + scale * (fi * x1[static_cast<size_t>(i)]);
                } else {                                               // odd count: the partner variate stays cached
                    gauss = fi * x1[static_cast<size_t>(i)];
                    has_gauss = true;
                }
            }
        }
    }
};

}  // namespace

extern "C" int pcgmix_host_lambda_knots(uint64_t seed, double alpha, double sigma, int64_t n_knots, int32_t max_threads,
                                        double* lam_out, double* knots_out, uint32_t* state_out, int32_t* pos_out,
                                        int32_t* has_gauss_out, double* gauss_out) {
    if (seed > 0xffffffffull || !(alpha > 0.0) || n_knots < 0 || lam_out == nullptr) return 1;
    if (n_knots > 0 && knots_out == nullptr) return 1;
    LegacyStream rs;
    rs.seed(static_cast<uint32_t>(seed));
    *lam_out = rs.beta(alpha, alpha);
    rs.fill_normal(1.0, sigma, n_knots, knots_out, max_threads < 1 ? 1 : max_threads);
    if (state_out != nullptr) {                            // what np.random.get_state() would show afterwards
        std::memcpy(state_out, rs.mt.mt, sizeof(rs.mt.mt));
        if (pos_out) *pos_out = rs.mt.index;
        if (has_gauss_out) *has_gauss_out = rs.has_gauss ? 1 : 0;
        if (gauss_out) *gauss_out = rs.gauss;
    }
    return 0;
}

// Pairing-chain processing order (b, mix[b], mix[mix[b]], ...): see draws.processing_order.
extern "C" int pcgmix_host_processing_order(const int64_t* mix, int64_t n, int32_t* order) {
    if (n < 0 || n >= (1ll << 31) || (n > 0 && (mix == nullptr || order == nullptr))) return 1;
    std::vector<char> seen(static_cast<size_t>(n), 0);
    int64_t k = 0;
    for (int64_t start = 0; start < n; ++start) {
        int64_t b = start;
        while (!seen[static_cast<size_t>(b)]) {
            seen[static_cast<size_t>(b)] = 1;
            order[k++] = static_cast<int32_t>(b);
            b = mix[b];
            if (b < 0 || b >= n) break;
        }
    }
    return 0;
}

extern "C" int pcgmix_host_group_permutation(const int64_t* group, int64_t n, int64_t n_groups, uint64_t seed,
                                             int64_t* mix) {
    if (n < 0 || n_groups < 0 || (n > 0 && (group == nullptr || mix == nullptr))) return 1;
    if (n >= (1ll << 31)) return 1;
    std::vector<std::vector<int64_t>> members(static_cast<size_t>(n_groups));
    for (int64_t i = 0; i < n; ++i) {
        if (group[i] < 0 || group[i] >= n_groups) return 1;
        members[static_cast<size_t>(group[i])].push_back(i);
    }
    Mt19937 rng;
    std::vector<int64_t> pool;
    for (auto& m : members) {
        const int64_t k = static_cast<int64_t>(m.size());
        if (k == 0) continue;
        rng.seed(seed);                                    // a FRESH generator per group, like the reference
        pool = m;
        for (int64_t i = 0; i < k; ++i) {
            const uint32_t j = rng.randbelow(static_cast<uint32_t>(k - i));
            mix[m[static_cast<size_t>(i)]] = pool[j];
            pool[j] = pool[static_cast<size_t>(k - i - 1)];
        }
    }
    return 0;
}


// The integer half of what the host contributes to one plain PCGmix / PCGmix+ step, in ONE call that never
// touches Python: offsets checked and narrowed to int32, same-label pairing, the clamped-window check,
// optional processing order — packed into the caller's (pinned) staging buffer in the layout the upload kernel
// copies verbatim, with a section reserved for the knots (the caller fills it from pcgmix_host_lambda_knots
// or from NumPy).  The drop-in augment() spent ~0.3 ms of interpreter time on these pieces at the reference's
// batch of 64 cycles, for a kernel that runs ~10 us (profiles/README.md).
//
//   labels[B]            class id per cycle (arg-max of the one-hot target)
//   frames               CPU offsets, row b at frames + b*frame_stride (int64 elements), 5 per row
//   packed / capacity    staging buffer; sections start at 16-byte boundaries
//   info[0..3]           byte offsets of {frames int32 [B][5], mix int32 [B], order int32 [B], knots f64 [B][K+2][C]}
//                        (-1: section absent), info[4] total bytes, info[5..8] on status 3: cycle, state,
//                        clamped destination width, clamped source width; on status 2: info[5] = cycle
//   mix_out[B]           the pairing as int64 (augment() returns it)
// Status: 0 ok; 1 bad arguments / buffer too small; 2 offsets negative, decreasing or beyond int32;
// 3 a pair's clamped windows differ in width (the reference raises a shape mismatch; see pair_window in
// common.cuh, same rule).
extern "C" int pcgmix_host_prepare_step(const int64_t* labels, int64_t B, const int64_t* frames, int64_t frame_stride,
                                        int64_t L, uint64_t seed, int32_t K, int32_t C, int32_t want_order,
                                        uint8_t* packed, int64_t capacity, int64_t* info, int64_t* mix_out) {
    if (B < 0 || B >= (1ll << 31) || L <= 0 || frame_stride < 5 || info == nullptr) return 1;
    if (B > 0 && (labels == nullptr || frames == nullptr || packed == nullptr || mix_out == nullptr)) return 1;
    if (K >= 0 && C <= 0) return 1;
    auto align16 = [](int64_t v) { return (v + 15) & ~static_cast<int64_t>(15); };
    const int64_t n_knots = K >= 0 ? B * (K + 2) * C : 0;
    const int64_t off_frames = 0;
    const int64_t off_mix = align16(off_frames + B * 20);
    const int64_t off_order = want_order ? align16(off_mix + B * 4) : -1;
    const int64_t off_knots = K >= 0 ? align16((want_order ? off_order : off_mix) + B * 4) : -1;
    const int64_t total = align16(K >= 0 ? off_knots + n_knots * 8 : (want_order ? off_order : off_mix) + B * 4);
    info[0] = off_frames; info[1] = off_mix; info[2] = off_order; info[3] = off_knots; info[4] = total;
    if (total > capacity) return 1;

    // offsets: non-negative, non-decreasing, int32
    int32_t* f32 = reinterpret_cast<int32_t*>(packed + off_frames);
    for (int64_t b = 0; b < B; ++b) {
        const int64_t* f = frames + b * frame_stride;
        bool ok = f[0] >= 0 && f[4] <= 2147483647ll;
        for (int s = 0; s < 4; ++s) ok = ok && f[s + 1] >= f[s];
        if (!ok) {
            info[5] = b;
            return 2;
        }
        for (int s = 0; s < 5; ++s) f32[b * 5 + s] = static_cast<int32_t>(f[s]);
    }

    // pairing: a seeded permutation inside every class, each class from a fresh Random(seed)
    {
        std::vector<int64_t> keys;                          // distinct class ids in order of first appearance
        std::vector<std::vector<int64_t>> members;
        for (int64_t i = 0; i < B; ++i) {
            size_t g = 0;
            while (g < keys.size() && keys[g] != labels[i]) ++g;
            if (g == keys.size()) {
                if (keys.size() >= 4096) return 1;          // (linear search: meant for a handful of classes)
                keys.push_back(labels[i]);
                members.emplace_back();
            }
            members[g].push_back(i);
        }
        Mt19937 seeded, rng;
        seeded.seed(seed);
        std::vector<int64_t> pool;
        for (auto& m : members) {
            const int64_t k = static_cast<int64_t>(m.size());
            rng = seeded;
            pool = m;
            for (int64_t i = 0; i < k; ++i) {
                const uint32_t j = rng.randbelow(static_cast<uint32_t>(k - i));
                mix_out[m[static_cast<size_t>(i)]] = pool[j];
                pool[j] = pool[static_cast<size_t>(k - i - 1)];
            }
        }
    }
    int32_t* mix32 = reinterpret_cast<int32_t*>(packed + off_mix);
    for (int64_t b = 0; b < B; ++b) mix32[b] = static_cast<int32_t>(mix_out[b]);

    // cycles running past the row end: blendable only if the clamped windows agree (pair_window's rule)
    for (int64_t b = 0; b < B; ++b) {
        const int32_t* f1 = f32 + b * 5;
        const int32_t* f2 = f32 + mix_out[b] * 5;
        if (f1[4] <= L && f2[4] <= L) continue;
        for (int s = 0; s < 4; ++s) {
            const int64_t n0 = std::min<int64_t>(f1[s + 1] - f1[s], f2[s + 1] - f2[s]);
            const int64_t wd = std::min<int64_t>(f1[s] + n0, L) - std::min<int64_t>(f1[s], L);
            const int64_t ws = std::min<int64_t>(f2[s] + n0, L) - std::min<int64_t>(f2[s], L);
            if (wd != ws && !(wd == 0 && ws == 1)) {
                info[5] = b; info[6] = s; info[7] = wd; info[8] = ws;
                return 3;
            }
        }
    }

    if (want_order) {
        int32_t* order = reinterpret_cast<int32_t*>(packed + off_order);
        if (pcgmix_host_processing_order(mix_out, B, order) != 0) return 1;
    }
    return 0;
}
