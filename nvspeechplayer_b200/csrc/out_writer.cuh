// int16 output staging shared by the render kernels: each thread owns one stream (one output row) and
// collects 8 consecutive samples in four 32-bit registers with a funnel-shift chain (no dynamic register
// indexing, no unrolling of the huge tick body), then issues ONE 16-byte store.  A warp store therefore
// touches 32 rows x 16 B: every 32-byte sector is completed by the same thread's next store while the line
// is still in L2, so HBM sees ~2 B/sample (see DESIGN.md "Output path").
#pragma once
#include <stdint.h>

namespace klatt {

struct OutWriter {
	int16_t *row;      // start of this stream's row
	uint32_t w0, w1, w2, w3;
	uint32_t count;    // samples pushed so far
	bool vec;          // 16-byte stores allowed (row is 16-byte aligned)

	__device__ __forceinline__ void init(int16_t *rowPtr, bool vecOk) {
		row = rowPtr; w0 = w1 = w2 = w3 = 0; count = 0; vec = vecOk;
	}
	__device__ __forceinline__ void push(int s) {
		// shift the 128-bit window right by one sample and insert the new one at the top
		w0 = __funnelshift_r(w0, w1, 16);
		w1 = __funnelshift_r(w1, w2, 16);
		w2 = __funnelshift_r(w2, w3, 16);
		w3 = (w3 >> 16) | ((uint32_t)s << 16);
		++count;
		if ((count & 7u) == 0) {
			int16_t *p = row + (count - 8);
			if (vec) {
#ifndef KLATT_OUT_PLAIN_STORES
				// st.global.cs: evict-first in L2 -- the output is written once and never read by the kernel, and must not push the
				// stream records out of the L2 (measured, ring scheduler: DRAM traffic of a step 64.7 -> 57.4 GB, same time)
				__stcs(reinterpret_cast<uint4 *>(p), make_uint4(w0, w1, w2, w3));
#else
				*reinterpret_cast<uint4 *>(p) = make_uint4(w0, w1, w2, w3);
#endif
			} else {
				storeScalar(p, 8);
			}
		}
	}
	// write the 1..7 samples of an incomplete last group
	__device__ __forceinline__ void flush() {
		uint32_t rem = count & 7u;
		if (rem == 0) return;
		// the window holds the last 8 pushes; the rem newest sit at the top: align them to the bottom
		for (uint32_t i = rem; i < 8; ++i) {
			w0 = __funnelshift_r(w0, w1, 16);
			w1 = __funnelshift_r(w1, w2, 16);
			w2 = __funnelshift_r(w2, w3, 16);
			w3 >>= 16;
		}
		storeScalar(row + (count - rem), rem);
	}
	__device__ __forceinline__ void storeScalar(int16_t *p, uint32_t n) {
		uint32_t a = w0, b = w1, c = w2, d = w3;
		for (uint32_t i = 0; i < n; ++i) {
			p[i] = (int16_t)(a & 0xffffu);
			a = __funnelshift_r(a, b, 16);
			b = __funnelshift_r(b, c, 16);
			c = __funnelshift_r(c, d, 16);
			d >>= 16;
		}
	}
};

}  // namespace klatt
