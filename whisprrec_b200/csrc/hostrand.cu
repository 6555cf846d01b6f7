// Host-side replica of CPython's `random.sample(range(n), k)` on the Mersenne Twister state of the `random` module.
// SGL's edge dropout (reference src/utils/augmentor.py:77-111) keeps `random.sample(range(nnz), int(nnz * (1 - p)))`
// of the adjacency's non-zeros every epoch; in pure Python that is a second per million edges.  This is the same
// algorithm (Lib/random.py: sample, _randbelow_with_getrandbits; Modules/_randommodule.c: genrand_uint32, getrandbits),
// so the selected indices AND the generator state afterwards are identical to Python's.  No GPU involved.
#include <cmath>
#include <cstring>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace {

struct PyMT {
    uint32_t mt[624];
    int pos;
    uint32_t next() {
        if (pos >= 624) {
            static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
            int kk;
            uint32_t y;
            for (kk = 0; kk < 624 - 397; ++kk) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < 623; ++kk) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
            pos = 0;
        }
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    // random.getrandbits(k), 1 <= k <= 64
    uint64_t getrandbits(int k) {
        if (k <= 32) return next() >> (32 - k);
        const uint64_t lo = next();                      // low word first, the last word is the one that is shifted
        const uint64_t hi = next() >> (64 - k);
        return lo | (hi << 32);
    }
    // random._randbelow_with_getrandbits(n), n >= 1
    uint64_t randbelow(uint64_t n) {
        int k = 0;
        for (uint64_t t = n; t; t >>= 1) ++k;            // n.bit_length()
        uint64_t r = getrandbits(k);
        while (r >= n) r = getrandbits(k);
        return r;
    }
};

}  // namespace

extern "C" int wr_pyrandom_sample(uint32_t *host_state /* 624 words */, int *host_pos, int64_t n, int64_t k,
                                  int64_t *host_out /* k */) {
    if (!host_state || !host_pos || (k > 0 && !host_out)) return WR_E_NULL;
    if (n < 0 || k < 0 || k > n || *host_pos < 0 || *host_pos > 624) return WR_E_SIZE;
    PyMT g;
    memcpy(g.mt, host_state, sizeof(g.mt));
    g.pos = *host_pos;
    // Lib/random.py sample(): below `setsize` the population is copied and partially shuffled, above it picks are
    // tracked in a set
    double setsize = 21;
    if (k > 5) setsize += std::pow(4.0, std::ceil(std::log((double)k * 3.0) / std::log(4.0)));
    if ((double)n <= setsize) {
        std::vector<int64_t> pool((size_t)n);
        for (int64_t i = 0; i < n; ++i) pool[(size_t)i] = i;
        for (int64_t i = 0; i < k; ++i) {
            const uint64_t j = g.randbelow((uint64_t)(n - i));
            host_out[i] = pool[j];
            pool[j] = pool[(size_t)(n - i - 1)];
        }
    } else if (n <= ((int64_t)1 << 28)) {
        std::vector<bool> seen((size_t)n, false);
        for (int64_t i = 0; i < k; ++i) {
            uint64_t j = g.randbelow((uint64_t)n);
            while (seen[j]) j = g.randbelow((uint64_t)n);
            seen[j] = true;
            host_out[i] = (int64_t)j;
        }
    } else {                                            // few picks from a huge range: a hash set, not an n-bit map
        std::unordered_set<uint64_t> seen;
        seen.reserve((size_t)k * 2);
        for (int64_t i = 0; i < k; ++i) {
            uint64_t j = g.randbelow((uint64_t)n);
            while (seen.count(j)) j = g.randbelow((uint64_t)n);
            seen.insert(j);
            host_out[i] = (int64_t)j;
        }
    }
    memcpy(host_state, g.mt, sizeof(g.mt));
    *host_pos = g.pos;
    return WR_OK;
}
