// lkb_rng.h -- counter-based random fill keyed on (seed, GLOBAL row, component).
//
// The reference's `rand` TBP (src/AbstractTypes/AbstractVectors.fypp:310-321) is an unseeded
// Fortran generator; to make results independent of how rows are sharded across GPUs the device
// uses a stateless hash of the global row index.  h = splitmix64 finaliser; uniform in (0,1) with
// 53 bits (bit-exact on host and device); normal by Box-Muller on two streams.
#pragma once
#include <stdint.h>
#include <math.h>
#ifndef __CUDACC__
#define LKB_RNG_HD inline
#else
#define LKB_RNG_HD __host__ __device__ inline
#endif

namespace lkb {

LKB_RNG_HD uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
LKB_RNG_HD double u01(uint64_t seed_mixed, uint64_t row, uint64_t stream) {
    uint64_t h = mix64(seed_mixed ^ (row * 4ULL + stream));
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
LKB_RNG_HD double rng_uniform(uint64_t seed_mixed, uint64_t row, int comp) {
    return u01(seed_mixed, row, 2ULL * (uint64_t)comp);
}
LKB_RNG_HD double rng_normal(uint64_t seed_mixed, uint64_t row, int comp) {
    double u1 = u01(seed_mixed, row, 2ULL * (uint64_t)comp);
    double u2 = u01(seed_mixed, row, 2ULL * (uint64_t)comp + 1ULL);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

}  // namespace lkb
