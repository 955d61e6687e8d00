// TEST INFRASTRUCTURE ONLY.  Compiles the FP32 production arithmetic (nvspeechplayer_b200/csrc/klatt_f32_core.cuh,
// the body of the CUDA kernel) for the HOST so that its numerics against the reference can be studied and
// regression-tested in the CPU-only container.  Nothing here is linked into libspeechPlayer.so; the product has
// no CPU path.  Differences from the device build: glibc libm instead of libdevice (<= 1-2 ulp), IEEE division
// instead of MUFU.RCP in fastRcp().
#include <cstdlib>
#include <cstring>
#include <vector>
namespace klatt { double *g_dbgPhase = nullptr; }
#include "../../nvspeechplayer_b200/csrc/klatt_f32_core.cuh"

using namespace klatt;

namespace {
struct CoarseArray {
	float w[kCoarseWords];
	float &at(int i) { return w[i]; }
};
struct ArrayOut {
	int16_t *p;
	uint32_t n;
	void push(int s) { p[n++] = (int16_t)s; }
};
}  // namespace

extern "C" void hostsim_debug_phase(double *buf) { klatt::g_dbgPhase = buf; }

extern "C" int hostsim_render_f32(int sampleRate, const double *frames, const uint32_t *minDur, const uint32_t *fadeDur,
                                  const int32_t *userIndex, const uint8_t *isNull, uint32_t nFrames, uint64_t seed,
                                  uint64_t streamId, uint32_t maxSamples, int16_t *out, uint32_t chunk, int32_t *lastIndexOut) {
	StreamState *st = (StreamState *)calloc(1, sizeof(StreamState));
	st->fm.lastUserIndex = -1;
	st->fm.curIsNull = 1;
	st->fm.oldIsNull = 1;
	StreamDesc d;
	memset(&d, 0, sizeof d);
	d.state = st; d.frames = frames; d.minDur = minDur; d.fadeDur = fadeDur; d.userIndex = userIndex; d.isNull = isNull;
	d.qCount = nFrames; d.qBase = 0; d.streamId = streamId;
	NoiseConfig nc;
	nc.mode = kNoisePhilox; nc.seed = seed;
	uint32_t total = 0;
	int32_t lui = -1;
	uint32_t qh = 0;
	if (chunk == 0) chunk = maxSamples;
	while (total < maxSamples) {
		uint32_t want = maxSamples - total < chunk ? maxSamples - total : chunk;
		ArrayOut ao{out + total, 0};
		CoarseArray cs;
		uint32_t got = renderStreamF32(d, sampleRate, want, ao, cs, nc, &lui, &qh);
		total += got;
		if (got < want) break;
	}
	if (lastIndexOut) *lastIndexOut = lui;
	free(st);
	return (int)total;
}
