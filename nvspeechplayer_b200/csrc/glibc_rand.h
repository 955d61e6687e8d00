// Replica of glibc's rand()/srand() (random_r TYPE_3: additive feedback x[i] = x[i-3] + x[i-31], degree 31,
// seeded by a Lehmer LCG, first 310 outputs discarded, result = word >> 1).  The reference draws its noise from
// the process-global libc rand() (reference src/speechWaveGenerator.cpp:40), never seeds it, and shares it
// between all players; the per-handle drop-in API reproduces exactly that sequence so that a Linux build of the
// reference and this engine render identical int16 in FP64 mode.  Unlike libc's, this generator can be
// checkpointed and rewound, which the engine needs because the number of draws a synthesize() call consumes
// (two per GENERATED sample) is only known after the kernel has run.
#pragma once
#include <stdint.h>

namespace klatt {

class GlibcRand {
public:
	GlibcRand() { seed(1); }  // an unseeded program behaves as srand(1)
	void seed(unsigned int s) {
		if (s == 0) s = 1;
		int32_t r[34];
		r[0] = (int32_t)s;
		for (int i = 1; i < 31; ++i) {
			// 16807 * r[i-1] mod (2^31 - 1) without overflow (Schrage)
			int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
			int64_t word = 16807 * lo - 2836 * hi;
			if (word < 0) word += 2147483647;
			r[i] = (int32_t)word;
		}
		for (int i = 0; i < 31; ++i) ring_[i] = (uint32_t)r[i];
		// state as after glibc's initialisation: front pointer 3 ahead of rear
		f_ = 3; b_ = 0;
		for (int i = 0; i < 310; ++i) next();
	}
	int next() {
		uint32_t v = ring_[f_] + ring_[b_];
		ring_[f_] = v;
		f_ = (f_ + 1 == 31) ? 0 : f_ + 1;
		b_ = (b_ + 1 == 31) ? 0 : b_ + 1;
		return (int)(v >> 1);
	}
	void fill(int32_t *out, uint64_t n) {
		for (uint64_t i = 0; i < n; ++i) out[i] = next();
	}
	void skip(uint64_t n) {
		for (uint64_t i = 0; i < n; ++i) next();
	}

private:
	uint32_t ring_[31];
	int f_, b_;
};

}  // namespace klatt
