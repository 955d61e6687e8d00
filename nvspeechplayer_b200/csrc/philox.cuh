// Philox4x32-10 (Salmon et al., SC'11) for the engine's counter-based noise mode.
// Independent of oracle/philox.h; both are pinned to the Random123 known-answer vectors.
//
// Noise contract (DESIGN.md "Noise"): generated sample i of a stream consumes draws 2i (aspiration,
// reference speechWaveGenerator.cpp:75) and 2i+1 (frication, :205).  One Philox block (4 words) therefore
// serves two consecutive samples: block index = i>>1, words {0,1} for even i, {2,3} for odd i.
// Counter = {lo(block), hi(block), lo(streamId), hi(streamId)}, key = {lo(seed), hi(seed)};
// the reference's rand() value is word>>1 (0..2^31-1 = glibc RAND_MAX).
#pragma once
#include <stdint.h>
#include "klatt_common.h"

namespace klatt {

struct Philox4 {
	uint32_t w[4];
};

KLATT_HD void philoxMulHiLo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#ifdef __CUDA_ARCH__
	lo = a * b;
	hi = __umulhi(a, b);
#else
	uint64_t p = (uint64_t)a * b;
	lo = (uint32_t)p;
	hi = (uint32_t)(p >> 32);
#endif
}

KLATT_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
	for (int round = 0; round < 10; ++round) {
		uint32_t hi0, lo0, hi1, lo1;
		philoxMulHiLo(0xD2511F53u, c0, hi0, lo0);
		philoxMulHiLo(0xCD9E8D57u, c2, hi1, lo1);
		uint32_t n0 = hi1 ^ c1 ^ k0;
		uint32_t n2 = hi0 ^ c3 ^ k1;
		c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	Philox4 r;
	r.w[0] = c0; r.w[1] = c1; r.w[2] = c2; r.w[3] = c3;
	return r;
}

KLATT_HD Philox4 noiseBlock(uint64_t seed, uint64_t streamId, uint64_t block) {
	return philox4x32_10((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)streamId, (uint32_t)(streamId >> 32),
	                     (uint32_t)seed, (uint32_t)(seed >> 32));
}

}  // namespace klatt
