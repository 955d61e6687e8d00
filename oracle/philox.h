/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Philox4x32-10 counter-based generator
 * (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11; the Random123 reference constants), restated for the oracle side of
 * the noise-parity contract.  The CUDA engine has its own independent
 * implementation (nvspeechplayer_b200/csrc/philox.cuh); both are pinned to the
 * published Random123 known-answer vectors in tests/test_philox.py.
 *
 * Noise contract shared with the engine (DESIGN.md "Noise"):
 *   draw d (0-based, in the order the reference calls rand(): aspiration then
 *   frication for every GENERATED sample, src/speechWaveGenerator.cpp:75 via
 *   :203, then :205) of stream `stream` under seed `seed` is
 *       word = philox4x32_10(ctr = {lo(d>>2), hi(d>>2), lo(stream), hi(stream)},
 *                            key = {lo(seed), hi(seed)})[d & 3]
 *       rand() value = word >> 1            (0 .. 2^31-1 == glibc RAND_MAX)
 */
#ifndef NVSP_ORACLE_PHILOX_H
#define NVSP_ORACLE_PHILOX_H
#include <stdint.h>

static inline void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int round = 0; round < 10; ++round) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline int oracle_noise_draw(uint64_t seed, uint64_t stream, uint64_t d) {
	uint64_t blk = d >> 2;
	uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
	uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
	uint32_t out[4];
	oracle_philox4x32_10(ctr, key, out);
	return (int)(out[d & 3] >> 1);
}
#endif
