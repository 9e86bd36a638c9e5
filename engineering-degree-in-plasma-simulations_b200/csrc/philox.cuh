// Counter-based RNG for every stochastic device kernel: Philox4x32-10 (Salmon et al., SC'11).
// Replaces the reference's global std::mt19937 `rnd` (ch4/v3/src/Rnd.cpp:4-18), which is sequential and
// cannot be shared by thousands of threads.  A stream is addressed, not advanced:
//   key     = (lo32(seed) ^ stream, hi32(seed))          stream = purpose + 16*species_id + 4096*rank
//   counter = (lo32(index), hi32(index), step, block)     index = particle / cell / new-particle number
// so results do not depend on the launch geometry.  No bit parity with mt19937 is possible or required;
// stochastic kernels are compared with the reference through ensemble statistics, and
// draw by draw with host restatements driven by the same streams (kept with the tests, together with a host copy of this generator),
// which are themselves pinned bit for bit against the compiled reference with its own draws.
#pragma once
#include <stdint.h>

enum RngPurpose { RNG_LOADER = 1, RNG_SOURCE = 2, RNG_HEAVY = 3, RNG_MCC = 4, RNG_MERGE = 5, RNG_DSMC = 6 };

#if defined(__CUDACC__)
#define PHILOX_HD __host__ __device__ __forceinline__
#else
#define PHILOX_HD static inline
#endif

PHILOX_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct PhiloxStream {
    uint32_t k0, k1, i0, i1, step, block;
    uint32_t out[4];
    int have;                                   // doubles left in out[]
    PHILOX_HD void init(uint64_t seed, uint32_t stream, uint64_t index, uint32_t step_) {
        k0 = (uint32_t)seed ^ stream; k1 = (uint32_t)(seed >> 32);
        i0 = (uint32_t)index; i1 = (uint32_t)(index >> 32); step = step_; block = 0; have = 0;
    }
    // uniform double in [0,1) with 53 random bits  (the reference's rnd(), Rnd.cpp:11-13)
    PHILOX_HD double next() {
        if (have == 0) {
            out[0] = i0; out[1] = i1; out[2] = step; out[3] = block++;
            philox4x32_10(out, k0, k1);
            have = 2;
        }
        have--;
        uint64_t bits = ((uint64_t)out[2 * have + 1] << 32) | out[2 * have];
        return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
    }
};

PHILOX_HD uint32_t rng_stream_id(int purpose, uint32_t species_id, int rank) { return (uint32_t)purpose + 16u * species_id + 4096u * (uint32_t)rank; }
