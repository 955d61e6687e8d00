// TEST INFRASTRUCTURE ONLY.  Compiles the FP32 production arithmetic (nvspeechplayer_b200/csrc/klatt_f32_core.cuh,
// the bodies of the CUDA kernels) for the HOST so that its numerics against the reference can be studied and
// regression-tested in the CPU-only container.  Nothing here is linked into libspeechPlayer.so; the product has
// no CPU path.  Differences from the device build: glibc libm instead of libdevice (<= 1-2 ulp), IEEE division
// instead of MUFU.RCP in fastRcp().
//
// hostsim_render_f32 mirrors how the engine drives the kernels:
//   mode 0  one general-kernel pass per `chunk` ticks, plans made inline at the pop tick (per-handle API)
//   mode 1  plans precomputed for the whole queue (klatt_plan_kernel), then ROUNDS: a stream with at least holdTicks
//           pure hold ticks ahead runs renderHoldF32(holdTicks), otherwise renderGeneralF32(genTicks)
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../nvspeechplayer_b200/csrc/klatt_f32_core.cuh"

using namespace klatt;

namespace {
struct ArrayOut {
	int16_t *p;
	uint32_t n;
	void push(int s) { p[n++] = (int16_t)s; }
};
}  // namespace

extern "C" int hostsim_render_f32(int sampleRate, const double *frames, const uint32_t *minDur, const uint32_t *fadeDur,
                                  const int32_t *userIndex, const uint8_t *isNull, uint32_t nFrames, uint64_t seed,
                                  uint64_t streamId, uint32_t maxSamples, int16_t *out, uint32_t chunk, int32_t *lastIndexOut,
                                  int mode, uint32_t holdTicks, uint32_t genTicks, uint32_t *holdTicksUsedOut) {
	StreamState *st = (StreamState *)calloc(1, sizeof(StreamState));
	st->fm.lastUserIndex = -1;
	st->fm.curIsNull = 1;
	st->fm.oldIsNull = 1;
	StreamDesc d;
	memset(&d, 0, sizeof d);
	d.state = st; d.frames = frames; d.minDur = minDur; d.fadeDur = fadeDur; d.userIndex = userIndex; d.isNull = isNull;
	d.qCount = nFrames; d.qBase = 0; d.streamId = streamId;
	std::vector<FadePlanF32> plans;
	if (mode == 1) {
		plans.resize(nFrames);
		int prevReal = -1;
		for (uint32_t j = 0; j < nFrames; ++j) {
			bool prevNull = (j == 0) || (isNull && isNull[j - 1]);
			bool curNull = isNull && isNull[j];
			double o[kNumParams], n[kNumParams];
			plannedFrames(prevReal >= 0 ? frames + (size_t)prevReal * kNumParams : nullptr, prevNull,
			              frames + (size_t)j * kNumParams, curNull, o, n);
			uint32_t F = fadeDur[j] > 1u ? fadeDur[j] : 1u;
			planFade(o, n, F, sampleRate, plans[j]);
			if (!curNull) prevReal = (int)j;
		}
		d.plans = plans.data();
	}
	NoiseConfig nc;
	nc.mode = kNoisePhilox; nc.seed = seed;
	uint32_t total = 0, holdUsed = 0;
	int32_t lui = -1;
	uint32_t qh = 0;
	if (chunk == 0) chunk = maxSamples;
	while (total < maxSamples) {
		uint32_t left = maxSamples - total;
		if (mode == 1 && holdTicks && left >= holdTicks && canHoldF32(*st, holdTicks)) {
			ArrayOut ao{out + total, 0};
			XchgSelf xc;
			renderHoldF32<kRoleBoth>(d, sampleRate, holdTicks, ao, nc, xc);
			total += holdTicks;
			holdUsed += holdTicks;
			lui = st->fm.lastUserIndex;
			continue;
		}
		uint32_t step = (mode == 1 && genTicks) ? genTicks : chunk;
		uint32_t want = left < step ? left : step;
		ArrayOut ao{out + total, 0};
		XchgSelf xc;
		uint32_t got = renderGeneralF32<kRoleBoth>(d, sampleRate, want, want, ao, nc, xc, &lui, &qh);
		total += got;
		if (got < want) break;
	}
	if (lastIndexOut) *lastIndexOut = lui;
	if (holdTicksUsedOut) *holdTicksUsedOut = holdUsed;
	free(st);
	return (int)total;
}
