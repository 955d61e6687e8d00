/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  A rand() that the compiled reference
 * binds to instead of libc's (linked into _ref/libspeechPlayer_ref_philox.so
 * with -Wl,-Bsymbolic).  The reference calls rand() twice per generated
 * sample (src/speechWaveGenerator.cpp:40 through :75 and :205); here those
 * calls return the per-stream Philox sequence defined in philox.h, so the
 * unmodified reference and the CUDA engine consume the same noise without a
 * replay buffer.  State is thread-local: one oracle stream at a time per thread.
 */
#include "philox.h"

static __thread uint64_t g_seed, g_stream, g_ndraws;
static __thread uint32_t g_cache[4];
static __thread uint64_t g_cache_blk = ~(uint64_t)0;

void oracle_noise_seed(uint64_t seed, uint64_t stream) {
	g_seed = seed; g_stream = stream; g_ndraws = 0; g_cache_blk = ~(uint64_t)0;
}
uint64_t oracle_noise_ndraws(void) { return g_ndraws; }

int rand(void) {
	uint64_t d = g_ndraws++;
	uint64_t blk = d >> 2;
	if (blk != g_cache_blk) {
		uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)g_stream, (uint32_t)(g_stream >> 32)};
		uint32_t key[2] = {(uint32_t)g_seed, (uint32_t)(g_seed >> 32)};
		oracle_philox4x32_10(ctr, key, g_cache);
		g_cache_blk = blk;
	}
	return (int)(g_cache[d & 3] >> 1);
}
