// Shared by the paired-warp FP32 kernels (klatt_f32.cu: rounds; klatt_f32_sched.cu: persistent stream scheduler):
// the shared-memory hand-over between the cascade warp and the parallel warp of a stream batch, and the role map.
#pragma once
#include <stdint.h>

namespace klatt {

__device__ __forceinline__ void zeroRow(int16_t *row, uint32_t from, uint32_t to) {
	for (uint32_t i = from; i < to; ++i) row[i] = 0;
}

// ---- the two sides of a stream in different warps of one block ------------------------------------------------
// block = 4 warps: warps 0,1 run the cascade side of streams [64b, 64b+32) and [64b+32, 64b+64) of the list,
// warps 2,3 the parallel side of the same streams.  Warp w and warp w+2 share one named barrier and a
// double-buffered shared-memory hand-over of (aspiration noise word, parallel-bank output, sawtooth value) x 8 ticks x 32 lanes.
struct XchgSmem {
	uint32_t base;  // shared-space address of this lane's column of the pair's [2 buffers][8 ticks][32 lanes] x 16 bytes
	uint32_t barId;
	__device__ __forceinline__ void put(uint32_t t, uint32_t wA, float par, float voice) {
		asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %3};" ::"r"(base + (t & 15u) * 512u), "r"(wA), "r"(__float_as_uint(par)),
		             "r"(__float_as_uint(voice)) : "memory");
	}
	__device__ __forceinline__ void get(uint32_t t, uint32_t &wA, float &par, float &voice) const {
		uint32_t p, v, pad;
		asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(wA), "=r"(p), "=r"(v), "=r"(pad) : "r"(base + (t & 15u) * 512u) : "memory");
		par = __uint_as_float(p);
		voice = __uint_as_float(v);
	}
	// immediate barrier ids: with a register operand ptxas reserves all 16 named barriers for the block, and the SM's
	// barrier pool then caps the resident blocks at 4 whatever the register count
	__device__ __forceinline__ void sync() {
		if (barId == 1u) asm volatile("bar.sync 1, 64;" ::: "memory");
		else asm volatile("bar.sync 2, 64;" ::: "memory");
	}
};
struct NullOut {
	__device__ __forceinline__ void push(int) {}
};

// Which side a warp runs.  A warp's scheduler is warp-id % 4, so "warps 0,1 cascade / 2,3 parallel" in every block would
// give two schedulers of an SM nothing but the (lighter, latency-bound) cascade warps and the other two nothing but the
// (heavier) parallel warps; flipping the assignment on a hash of the block index mixes both kinds on every scheduler.
__device__ __forceinline__ bool cascadeRole(uint32_t warp) {
#ifdef KLATT_NO_ROLE_FLIP  // A/B: cascade warps on schedulers 0,1 and parallel warps on 2,3 of every SM (one loop per instruction cache)
	const uint32_t flip = 0;
#else
	const uint32_t flip = (blockIdx.x * 0x9E3779B9u) >> 31;
#endif
	return ((warp >> 1) ^ flip) == 0;
}

constexpr int kPairBlock = 128;        // threads
constexpr int kPairStreams = 64;       // streams per block

}  // namespace klatt
