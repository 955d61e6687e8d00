/*
 * TEST INFRASTRUCTURE ONLY (oracle/) -- see klatt_oracle.h for the rules.
 *
 * Double-precision restatement of the reference hot path.  Every routine names
 * the reference lines whose behaviour it restates (paths relative to
 * /root/reference).  Evaluation ORDER of every floating-point expression is the
 * reference's (left-to-right C semantics of the original expressions); the
 * file is compiled with -ffp-contract=off so no FMA contraction sneaks in.
 */
#include "klatt_oracle.h"
#include "philox.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* parameter slots, ABI order of speechPlayer_frame_t (src/frame.h:22-46) */
enum {
	P_voicePitch = 0, P_vibratoPitchOffset, P_vibratoSpeed, P_voiceTurbulenceAmplitude, P_glottalOpenQuotient,
	P_voiceAmplitude, P_aspirationAmplitude,
	P_cf1, P_cf2, P_cf3, P_cf4, P_cf5, P_cf6, P_cfN0, P_cfNP,
	P_cb1, P_cb2, P_cb3, P_cb4, P_cb5, P_cb6, P_cbN0, P_cbNP,
	P_caNP, P_fricationAmplitude,
	P_pf1, P_pf2, P_pf3, P_pf4, P_pf5, P_pf6,
	P_pb1, P_pb2, P_pb3, P_pb4, P_pb5, P_pb6,
	P_pa1, P_pa2, P_pa3, P_pa4, P_pa5, P_pa6,
	P_parallelBypass, P_preFormantGain, P_outputGain, P_endVoicePitch,
	P_COUNT
};
typedef char klatt_param_count_check[(P_COUNT == KLATT_ORACLE_NUM_PARAMS) ? 1 : -1];

#define NP KLATT_ORACLE_NUM_PARAMS

/* one queued request: src/frame.cpp:21-28 */
typedef struct {
	unsigned minNumSamples, numFadeSamples;
	int isNull;
	double p[NP];
	double voicePitchInc;
	int userIndex;
} request_t;

/* one 2-pole section: src/speechWaveGenerator.cpp:90-110 */
typedef struct {
	int anti, setOnce;
	double frequency, bandwidth;
	double a, b, c;
	double p1, p2;
} resonator_t;

enum { NOISE_LIBC = 0, NOISE_PHILOX = 1, NOISE_REPLAY = 2 };

struct klatt_oracle {
	int sampleRate;
	/* frame manager state: src/frame.cpp:30-39 */
	request_t *queue; size_t qHead, qTail, qCap;
	request_t oldReq, newReq; int hasNew;
	double cur[NP]; int curIsNull;
	unsigned sampleCounter; int lastUserIndex;
	/* generator state: src/speechWaveGenerator.cpp:184-194 */
	double pitchPos, vibratoPos, aspLast, fricLast;
	resonator_t casc[8]; /* order of use: N0, NP, 6,5,4,3,2,1 */
	resonator_t par[6];  /* 1..6 */
	/* noise source */
	int noiseMode; uint64_t seed, stream; const int32_t *replay; size_t nReplay;
	uint64_t nDraws, nTicks, nFadeTicks;
	int fadedThisTick;
	double *dbgPhase; /* optional: pitchPos after every generated sample */
};

/* ---- noise draw: stands in for rand() (src/speechWaveGenerator.cpp:40) ---- */
static int next_draw(klatt_oracle_t *o) {
	uint64_t d = o->nDraws++;
	switch (o->noiseMode) {
	case NOISE_PHILOX: return oracle_noise_draw(o->seed, o->stream, d);
	case NOISE_REPLAY: return (d < o->nReplay) ? o->replay[d] : 0;
	default: return rand();
	}
}

/* src/utils.h:20-23 -- NaN target keeps the old value */
static double fade_value(double oldVal, double newVal, double ratio) {
	if (isnan(newVal)) return oldVal;
	return oldVal + ((newVal - oldVal) * ratio);
}

/* src/speechWaveGenerator.cpp:39-42, RAND_MAX of glibc = 2^31-1 */
static double noise_next(klatt_oracle_t *o, double *last) {
	*last = ((double)next_draw(o) / 2147483647) + 0.75 * (*last);
	return *last;
}

/* src/speechWaveGenerator.cpp:54-58 */
static double phase_next(int sampleRate, double *pos, double frequency) {
	double cyclePos = fmod((frequency / sampleRate) + *pos, 1);
	*pos = cyclePos;
	return cyclePos;
}

/* src/speechWaveGenerator.cpp:112-127 */
static void resonator_set(resonator_t *r, int sampleRate, double frequency, double bandwidth) {
	if (!r->setOnce || (frequency != r->frequency) || (bandwidth != r->bandwidth)) {
		const double PITWO = M_PI * 2;
		r->frequency = frequency;
		r->bandwidth = bandwidth;
		double rad = exp(-M_PI / sampleRate * bandwidth);
		r->c = -(rad * rad);
		r->b = rad * cos(PITWO / sampleRate * -frequency) * 2.0;
		r->a = 1.0 - r->b - r->c;
		if (r->anti && frequency != 0) {
			r->a = 1.0 / r->a;
			r->c *= -r->a;
			r->b *= -r->a;
		}
	}
	r->setOnce = 1;
}

/* src/speechWaveGenerator.cpp:129-135 */
static double resonate(resonator_t *r, int sampleRate, double in, double frequency, double bandwidth) {
	resonator_set(r, sampleRate, frequency, bandwidth);
	double out = r->a * in + r->b * r->p1 + r->c * r->p2;
	r->p2 = r->p1;
	r->p1 = r->anti ? in : out;
	return out;
}

/* src/frame.cpp:41-80: one tick of the frame manager */
static void frame_tick(klatt_oracle_t *o) {
	o->fadedThisTick = 0;
	o->sampleCounter++;
	if (o->hasNew) {
		if (o->sampleCounter > o->newReq.numFadeSamples) { /* :44-47 fade finished: new becomes old */
			o->oldReq = o->newReq;
			o->hasNew = 0;
		} else { /* :49-52 interpolate all 47 */
			double ratio = (double)o->sampleCounter / o->newReq.numFadeSamples;
			for (int i = 0; i < NP; ++i) o->cur[i] = fade_value(o->oldReq.p[i], o->newReq.p[i], ratio);
			o->fadedThisTick = 1;
		}
	} else if (o->sampleCounter > o->oldReq.minNumSamples) { /* :54 hold expired */
		if (o->qHead != o->qTail) { /* :55-72 */
			o->curIsNull = 0;
			o->newReq = o->queue[o->qHead++];
			o->hasNew = 1;
			if (o->newReq.isNull) { /* :59-63 fade to silence keeping the formants */
				memcpy(o->newReq.p, o->oldReq.p, sizeof o->newReq.p);
				o->newReq.p[P_preFormantGain] = 0;
				o->newReq.p[P_voicePitch] = o->cur[P_voicePitch];
				o->newReq.voicePitchInc = 0;
			} else if (o->oldReq.isNull) { /* :64-67 fade in from silence with the new formants */
				memcpy(o->oldReq.p, o->newReq.p, sizeof o->oldReq.p);
				o->oldReq.p[P_preFormantGain] = 0;
			}
			if (o->newReq.userIndex != -1) o->lastUserIndex = o->newReq.userIndex; /* :69 */
			o->sampleCounter = 0;                                                     /* :70 */
			o->newReq.p[P_voicePitch] += (o->newReq.voicePitchInc * o->newReq.numFadeSamples); /* :71 */
		} else {
			o->curIsNull = 1; /* :73-75 */
		}
	} else { /* :76-79 hold: only the pitch glides */
		o->cur[P_voicePitch] += o->oldReq.voicePitchInc;
		o->oldReq.p[P_voicePitch] = o->cur[P_voicePitch];
	}
}

/* src/speechWaveGenerator.cpp:72-86 */
static double voice_next(klatt_oracle_t *o, int *glottisOpen) {
	const double PITWO = M_PI * 2;
	const double *f = o->cur;
	double vibrato = (sin(phase_next(o->sampleRate, &o->vibratoPos, f[P_vibratoSpeed]) * PITWO) * 0.06 * f[P_vibratoPitchOffset]) + 1;
	double voice = phase_next(o->sampleRate, &o->pitchPos, f[P_voicePitch] * vibrato);
	double aspiration = noise_next(o, &o->aspLast) * 0.2;
	double turbulence = aspiration * f[P_voiceTurbulenceAmplitude];
	*glottisOpen = voice >= f[P_glottalOpenQuotient];
	if (!*glottisOpen) turbulence *= 0.01;
	voice = (voice * 2) - 1;
	voice += turbulence;
	voice *= f[P_voiceAmplitude];
	aspiration *= f[P_aspirationAmplitude];
	return aspiration + voice;
}

/* src/speechWaveGenerator.cpp:147-158 */
static double cascade_next(klatt_oracle_t *o, double input) {
	const double *f = o->cur;
	const int sr = o->sampleRate;
	input /= 2.0;
	double n0Output = resonate(&o->casc[0], sr, input, f[P_cfN0], f[P_cbN0]);
	double output = fade_value(input, resonate(&o->casc[1], sr, n0Output, f[P_cfNP], f[P_cbNP]), f[P_caNP]);
	for (int k = 0; k < 6; ++k) /* r6, r5, ... r1 */
		output = resonate(&o->casc[2 + k], sr, output, f[P_cf6 - k], f[P_cb6 - k]);
	return output;
}

/* src/speechWaveGenerator.cpp:170-180 */
static double parallel_next(klatt_oracle_t *o, double input) {
	const double *f = o->cur;
	const int sr = o->sampleRate;
	input /= 2.0;
	double output = 0;
	for (int k = 0; k < 6; ++k)
		output += (resonate(&o->par[k], sr, input, f[P_pf1 + k], f[P_pb1 + k]) - input) * f[P_pa1 + k];
	return fade_value(output, input, f[P_parallelBypass]);
}

klatt_oracle_t *klatt_oracle_create(int sampleRate) {
	/* src/speechPlayer.cpp:25-32, src/frame.cpp:85-88, src/speechWaveGenerator.cpp:104-110,145,194 */
	klatt_oracle_t *o = (klatt_oracle_t *)calloc(1, sizeof *o);
	if (!o) return NULL;
	o->sampleRate = sampleRate;
	o->curIsNull = 1;
	o->lastUserIndex = -1;
	o->oldReq.isNull = 1;
	o->casc[0].anti = 1; /* rN0 is the anti-resonator */
	o->noiseMode = NOISE_LIBC;
	return o;
}

void klatt_oracle_destroy(klatt_oracle_t *o) {
	if (!o) return;
	free(o->queue);
	free(o);
}

void klatt_oracle_noise_libc(klatt_oracle_t *o) { o->noiseMode = NOISE_LIBC; }
void klatt_oracle_noise_philox(klatt_oracle_t *o, uint64_t seed, uint64_t stream) {
	o->noiseMode = NOISE_PHILOX; o->seed = seed; o->stream = stream;
}
void klatt_oracle_noise_replay(klatt_oracle_t *o, const int32_t *draws, size_t n) {
	o->noiseMode = NOISE_REPLAY; o->replay = draws; o->nReplay = n;
}

/* src/speechPlayer.cpp:34-37 + src/frame.cpp:90-115 */
void klatt_oracle_queue_frame(klatt_oracle_t *o, const double *frame, unsigned minDuration, unsigned fadeDuration,
                              int userIndex, int purgeQueue) {
	request_t r;
	memset(&r, 0, sizeof r);
	r.minNumSamples = minDuration;
	r.numFadeSamples = fadeDuration > 1 ? fadeDuration : 1; /* speechPlayer.cpp:36 max(fadeDuration,1) */
	if (frame) {
		r.isNull = 0;
		memcpy(r.p, frame, sizeof r.p);
		r.voicePitchInc = (frame[P_endVoicePitch] - frame[P_voicePitch]) / r.minNumSamples; /* frame.cpp:98 */
	} else {
		r.isNull = 1;
	}
	r.userIndex = userIndex;
	if (purgeQueue) { /* frame.cpp:103-112 */
		o->qHead = o->qTail = 0;
		o->sampleCounter = o->oldReq.minNumSamples;
		if (o->hasNew) {
			o->oldReq.isNull = o->newReq.isNull;
			memcpy(o->oldReq.p, o->cur, sizeof o->oldReq.p);
			o->hasNew = 0;
		}
	}
	if (o->qTail == o->qCap) {
		if (o->qHead > 0) { /* compact */
			memmove(o->queue, o->queue + o->qHead, (o->qTail - o->qHead) * sizeof(request_t));
			o->qTail -= o->qHead; o->qHead = 0;
		}
		if (o->qTail == o->qCap) {
			size_t cap = o->qCap ? o->qCap * 2 : 64;
			o->queue = (request_t *)realloc(o->queue, cap * sizeof(request_t));
			o->qCap = cap;
		}
	}
	o->queue[o->qTail++] = r;
}

/* src/speechWaveGenerator.cpp:197-214 */
int klatt_oracle_synthesize(klatt_oracle_t *o, unsigned sampleCount, int16_t *out) {
	for (unsigned i = 0; i < sampleCount; ++i) {
		frame_tick(o); /* frame.cpp:121-126 */
		if (o->curIsNull) return (int)i;
		const double *f = o->cur;
		int glottisOpen;
		double voice = voice_next(o, &glottisOpen);
		double cascadeOut = cascade_next(o, voice * f[P_preFormantGain]);
		double fric = noise_next(o, &o->fricLast) * 0.3 * f[P_fricationAmplitude];
		double parallelOut = parallel_next(o, fric * f[P_preFormantGain]);
		double v = (cascadeOut + parallelOut) * f[P_outputGain];
		/* :208 with the Win32 min/max MACROS: (x<32000?x:32000) then (y>-32000?y:-32000); NaN -> +32000 */
		double scaled = v * 4000;
		double lo = (scaled < 32000) ? scaled : 32000;
		double cl = (lo > -32000) ? lo : -32000;
		out[i] = (int16_t)(int)cl; /* truncation toward zero */
		if (o->dbgPhase) o->dbgPhase[o->nTicks] = o->pitchPos;
		o->nTicks++;
		o->nFadeTicks += (uint64_t)o->fadedThisTick;
	}
	return (int)sampleCount;
}

void klatt_oracle_debug_phase(klatt_oracle_t *o, double *buf) { o->dbgPhase = buf; }
int klatt_oracle_get_last_index(const klatt_oracle_t *o) { return o->lastUserIndex; } /* frame.cpp:117-119 */
uint64_t klatt_oracle_ticks(const klatt_oracle_t *o) { return o->nTicks; }
uint64_t klatt_oracle_fade_ticks(const klatt_oracle_t *o) { return o->nFadeTicks; }
uint64_t klatt_oracle_draws(const klatt_oracle_t *o) { return o->nDraws; }

int klatt_oracle_render(int sampleRate, const double *frames, const uint32_t *minDur, const uint32_t *fadeDur,
                        const int32_t *userIndex, const uint8_t *isNull, unsigned nFrames,
                        int noiseMode, uint64_t seed, uint64_t stream, const int32_t *draws, size_t nDraws,
                        unsigned maxSamples, int16_t *out, uint64_t *fadeTicksOut) {
	klatt_oracle_t *o = klatt_oracle_create(sampleRate);
	if (!o) return -1;
	if (noiseMode == NOISE_PHILOX) klatt_oracle_noise_philox(o, seed, stream);
	else if (noiseMode == NOISE_REPLAY) klatt_oracle_noise_replay(o, draws, nDraws);
	for (unsigned j = 0; j < nFrames; ++j) {
		const double *fr = (isNull && isNull[j]) ? NULL : frames + (size_t)j * NP;
		klatt_oracle_queue_frame(o, fr, minDur[j], fadeDur[j], userIndex ? userIndex[j] : -1, 0);
	}
	int n = klatt_oracle_synthesize(o, maxSamples, out);
	if (fadeTicksOut) *fadeTicksOut = o->nFadeTicks;
	klatt_oracle_destroy(o);
	return n;
}
