// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
// 1, 2, 3", SC'11), written from the published algorithm.  Host + device.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define PTB_HD __host__ __device__ __forceinline__
#else
#define PTB_HD inline
#endif

namespace ptb200 {

PTB_HD void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; round++) {
        const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
        const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = static_cast<uint32_t>(p1);
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = static_cast<uint32_t>(p0);
        c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// The ten round keys of a 64-bit seed (k += Weyl constants each round), so that a kernel whose seed sits in the constant
// bank spends no instructions on the schedule.
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};

PTB_HD PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys k;
    uint32_t a = static_cast<uint32_t>(seed), b = static_cast<uint32_t>(seed >> 32);
    for (int r = 0; r < 10; r++) {
        k.k0[r] = a, k.k1[r] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return k;
}

// philox4x32_10 with a precomputed key schedule: identical output words.
PTB_HD void philox4x32_10_keyed(uint32_t (&c)[4], const PhiloxKeys &k) {
#pragma unroll
    for (int round = 0; round < 10; round++) {
        const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
        const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k.k0[round];
        const uint32_t n1 = static_cast<uint32_t>(p1);
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k.k1[round];
        const uint32_t n3 = static_cast<uint32_t>(p0);
        c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
    }
}

// Two uniforms in [0,1) for path `index` under `seed`, each built from two 32-bit words the way NumPy's
// genrand_res53 does ((a>>5)*2^26 + (b>>6)) / 2^53 -- so the counter-based stream feeds the very same
// camera code as the replayed MT19937 stream.
PTB_HD void philox_uniform2(uint64_t seed, uint64_t index, double &u1, double &u2) {
    uint32_t c[4] = {static_cast<uint32_t>(index), static_cast<uint32_t>(index >> 32), 0u, 0u};
    philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    u1 = (static_cast<double>(c[0] >> 5) * 67108864.0 + static_cast<double>(c[1] >> 6)) / 9007199254740992.0;
    u2 = (static_cast<double>(c[2] >> 5) * 67108864.0 + static_cast<double>(c[3] >> 6)) / 9007199254740992.0;
}

}  // namespace ptb200
