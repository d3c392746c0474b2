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
// No CUDA here; plain host code in the same shared library.

#include <cstdint>
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
    uint32_t next() {
        if (index >= 624) {
            for (int k = 0; k < 624; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            index = 0;
        }
        uint32_t y = mt[index++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    uint32_t randbelow(uint32_t n) {                       // 1 <= n < 2^31
        int bits = 0;
        for (uint32_t v = n; v; v >>= 1) ++bits;
        uint32_t r = next() >> (32 - bits);
        while (r >= n) r = next() >> (32 - bits);
        return r;
    }
};

}  // namespace

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
