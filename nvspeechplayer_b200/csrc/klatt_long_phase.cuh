// The glottal phase of the long-utterance path, EXACT and parallel in time.
//
// The reference accumulates  pos = fmod(pos + quot, 1)  in double, one tick after the other (src/speechWaveGenerator.cpp:
// 54-58, :74); with a pitch period of a whole number of samples the tick on which the sawtooth wraps is decided by the last
// bit of that running sum, so an associative surrogate (a 2^-64 fixed-point prefix sum: what this path used in round 1) puts
// a few wraps per stream one sample off -- a full-scale error each.  A chain of FP64 roundings is not associative, but it is
// TRANSLATION-EQUIVARIANT on the coarsest grid it visits:
//
//   * every phase value lies in [0, 1): its binade's ulp divides u = 2^-53, the ulp of [0.5, 1);
//   * if P' = P + m*u (m an integer, |m*u| tiny) then RN(quot + P') = RN(quot + P) + m*u on every tick where the sum stays
//     in a binade below 1 -- the shift is a whole number of grid steps, and an EVEN number wherever the grid is finer than u,
//     so ties round the same way;
//   * the places where the residue of m matters: a tie in [0.5, 1) (parity of m), the wrap tick, whose sum lies in [1, 2)
//     with ulp 2u (an odd m is rounded to an even one), and a tie ON the wrap tick (parity of m/2).  A shift by a multiple of
//     4u is an even number of steps on every grid involved and passes through all of them unchanged.
//
// So a chunk of ticks that starts from a value s0 on the u-grid maps a start offset m to an end offset  m + a[m & 3]  with
// four integers a[0..3], read off FOUR speculative runs of the plain recurrence (from s0 + r*u, r = 0..3).  Maps of this form
// compose associatively (phaseMapCompose): a scan over the chunks gives every chunk the exact offset of its true start value,
// and a second run of the plain recurrence from that value IS the reference's sequence.  Chunks are anchored where the phase
// is in [0.52, 0.98] (u-grid, away from any binade edge), located with the exact fixed-point prefix sum.  The whole
// construction is verified, not trusted: the true run of a chunk must END on exactly the next chunk's true start; a mismatch
// (or a stretch without anchors: zero or negative pitch) sends the call through the serial fallback.
//
// Host + device code (tests/hostsim runs the same functions on adversarial increment sequences against the plain loop).
#pragma once
#include <stdint.h>
#include "klatt_common.h"
#include "klatt_f32_core.cuh"

namespace klatt {

constexpr double kPhaseUlp = 1.1102230246251565e-16;  // 2^-53

// phi: phase ENTERING a tick as 2^-64-cycle fixed point
KLATT_HD bool phaseAnchorOk(uint64_t phi) {
	const uint32_t top = (uint32_t)(phi >> 32);
	return top >= 0x851EB852u /* 0.52 */ && top <= 0xFAE147AEu /* 0.98 */;
}
KLATT_HD double phaseAnchorValue(uint64_t phi) {  // phi rounded down to the u-grid: a double in [0.5, 1)
	return (double)(int64_t)(phi >> 11) * kPhaseUlp;
}
KLATT_HD double phaseStep(double pos, double quot) { return fracRef(quot + pos); }  // one tick of src/speechWaveGenerator.cpp:55

constexpr int kPhaseRuns = 4;
struct PhaseMap {  // start offset m (units of u) -> end offset m + a[m & 3]
	int64_t a[kPhaseRuns];
};
KLATT_HD PhaseMap phaseMapIdentity() { PhaseMap r; for (int i = 0; i < kPhaseRuns; ++i) r.a[i] = 0; return r; }
KLATT_HD PhaseMap phaseMapCompose(const PhaseMap &second, const PhaseMap &first) {
	PhaseMap r;
#pragma unroll
	for (int i = 0; i < kPhaseRuns; ++i) r.a[i] = first.a[i] + second.a[(uint64_t)(i + first.a[i]) & 3u];
	return r;
}
KLATT_HD int64_t phaseMapApply(const PhaseMap &f, int64_t m) { return m + f.a[(uint64_t)m & 3u]; }
// e[r]: where the run from s0 + r*u arrives at the next anchor, whose assumed value is s0Next
KLATT_HD PhaseMap phaseMapOf(const double *e, double s0Next) {
	PhaseMap r;
#pragma unroll
	for (int i = 0; i < kPhaseRuns; ++i) r.a[i] = (int64_t)((e[i] - s0Next) * 9007199254740992.0) - i;  // exact: a multiple of u
	return r;
}
KLATT_HD double phaseFromOffset(double s0, int64_t m) { return s0 + (double)m * kPhaseUlp; }

// What the speculative pass leaves per chunk (chunk c nominally starts at tick c * L).
struct PhaseChunk {
	uint64_t anchor;   // first tick t >= c*L (t < (c+1)*L) whose entering phase passes phaseAnchorOk; chunk 0: tick 0; ~0: none
	uint64_t next;     // the next anchor after `anchor` (any later chunk's), or the stream's length
	double s0;         // assumed phase entering tick `anchor`
	PhaseMap map;      // offset at `anchor` -> offset at `next` (identity for anchorless chunks and for the last anchor)
};
constexpr uint64_t kNoAnchor = ~0ull;

// One chunk of the speculative pass.  `src` is positioned on tick t0 = c*L and yields, per call of src.tick(quot, fixedInc),
// the increment of the current tick (the double the recurrence adds, and the same value as 2^-64 fixed point) and moves on.
// phi: fixed-point phase entering tick t0.  Returns false when the walk ran into `limit` ticks without meeting the next anchor.
template <class Src>
KLATT_HD bool phaseSpeculateChunk(Src &src, uint64_t c, uint64_t L, uint64_t total, uint64_t phi, uint64_t limit, PhaseChunk &out) {
	const uint64_t t0 = c * L;
	out.anchor = kNoAnchor; out.next = total; out.s0 = 0.0; out.map = phaseMapIdentity();
	uint64_t t = t0;
	double quot;
	uint64_t fi;
	// 1. this chunk's anchor
	if (c == 0) {
		out.anchor = 0; out.s0 = 0.0;  // the stream starts from phase 0 exactly (src/speechWaveGenerator.cpp:49)
	} else {
		const uint64_t tEnd = (t0 + L < total) ? t0 + L : total;
		while (t < tEnd && !phaseAnchorOk(phi)) { src.tick(quot, fi); phi += fi; ++t; }
		if (t >= tEnd) return true;  // anchorless: an earlier chunk's runs cover these ticks
		out.anchor = t; out.s0 = phaseAnchorValue(phi);
	}
	// 2. the four runs up to the next anchor
	double p[kPhaseRuns];
#pragma unroll
	for (int i = 0; i < kPhaseRuns; ++i) p[i] = out.s0 + (double)i * kPhaseUlp;
	const uint64_t nextChunkStart = t0 + L;
	const uint64_t stop = (t0 + limit < total) ? t0 + limit : total;
	for (;;) {
		if (t >= total) { out.next = total; return true; }  // last anchor of the stream
		if (t >= nextChunkStart && phaseAnchorOk(phi)) break;
		if (t >= stop) return false;
		src.tick(quot, fi);
#pragma unroll
		for (int i = 0; i < kPhaseRuns; ++i) p[i] = phaseStep(p[i], quot);
		phi += fi;
		++t;
	}
	out.next = t;
	out.map = phaseMapOf(p, phaseAnchorValue(phi));
	return true;
}

}  // namespace klatt
