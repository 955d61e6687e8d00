// FP32 production formulation of the Klatt hot path (one stream, serial in time).
//
// The function renderStreamF32() below IS the body of the batch kernel (klatt_f32.cu calls it once per
// thread); it is written against plain pointers so that the very same arithmetic can also be compiled for the
// host by the numerics study under tests/hostsim/ (test infrastructure; the product never runs it on a CPU).
//
// What it computes is the reference's per-sample loop (reference src/speechWaveGenerator.cpp:197-214 driven by
// the frame manager of src/frame.cpp:41-80), reformulated so that FP32 is accurate enough (DESIGN.md
// "FP32 formulation" has the derivations and the measured SNR):
//
//  * frame manager: same tick state machine (pop / fade / swap / hold, NULL-frame rewrites, userIndex, drain),
//    but a fade is walked with per-tick INCREMENTS prepared once per request by planFade() in double.
//  * pitch and glottal phase stay FP64 (src/speechWaveGenerator.cpp:55,74): the sawtooth turns phase error into
//    full-scale error (a wrap that lands one sample early rings through every resonator), FP32 phase gives
//    ~11 dB SNR.  Vibrato phase and its increment are 64-bit fixed point: exact accumulation, no drift.
//  * two-pole sections run in DELTA FORM:  d' = (1-rho)*d + a*(x - y);  y' = y + d'
//    (a = |1-pole|^2, rho = 1-|pole|^2), algebraically the reference's  y' = a*x + b*y + c*y1  (b = 2 - rho - a,
//    c = rho - 1; src :116-131) but only the SMALL quantities a, rho, d are represented, instead of cancelling
//    b ~ 2 against c ~ -1.
//  * coefficients during a fade: a linear fade of (f, bw) moves the pole along a complex geometric sequence,
//    so zeta = 1 - pole is advanced by  zeta' = zeta + omega - zeta*omega  (6 FMAs) and
//    rho' = rho + kappa*(1 - rho), instead of 14 x (exp + cos) per tick (src :113-125).
//  * noise, source shaping, mixes, gain, clamp and truncation follow src :40, :72-86, :150, :172-179, :207-208.
#pragma once
#include <math.h>
#include <stdint.h>
#include "klatt_common.h"
#include "philox.cuh"

namespace klatt {

// slots of GenStateF32::dir
enum Direct : int {
	dVibratoPitchOffset = 0, dVoiceTurbulenceAmplitude, dGlottalOpenQuotient, dVoiceAmplitude,
	dAspirationAmplitude, dCaNP, dFricationAmplitude, dPa1, dPa2, dPa3, dPa4, dPa5, dPa6, dParallelBypass,
	dPreFormantGain, dOutputGain
};
static_assert(dOutputGain + 1 == kNumDirect, "direct slots");

KLATT_HD constexpr int directParam(int i) {
	return i == dVibratoPitchOffset ? kVibratoPitchOffset
	     : i == dVoiceTurbulenceAmplitude ? kVoiceTurbulenceAmplitude : i == dGlottalOpenQuotient ? kGlottalOpenQuotient
	     : i == dVoiceAmplitude ? kVoiceAmplitude : i == dAspirationAmplitude ? kAspirationAmplitude
	     : i == dCaNP ? kCaNP : i == dFricationAmplitude ? kFricationAmplitude
	     : (i >= dPa1 && i <= dPa6) ? (kPa1 + (i - dPa1))
	     : i == dParallelBypass ? kParallelBypass : i == dPreFormantGain ? kPreFormantGain : kOutputGain;
}

// ---------------------------------------------------------------------------------------------------
// planFade: everything a fade from frame `o` to frame `n` over F ticks needs (double precision, once per request).
// `o`/`n` are what the reference holds in oldFrameRequest->frame / newFrameRequest->frame after the pop-tick
// rewrites (src/frame.cpp:59-71).  NaN in a target keeps the old value (src/utils.h:21).
// ---------------------------------------------------------------------------------------------------
KLATT_HD void poleTerms(double f, double bw, double srInv, float &zre, float &zim, float &rho) {
	const double PI = 3.14159265358979323846;
	double x = -PI * bw * srInv;        // log of the pole radius
	double th = 2.0 * PI * f * srInv;   // pole angle
	double em1 = expm1(x);              // r - 1
	double r = 1.0 + em1;
	double sh = sin(0.5 * th);
	zre = (float)(-em1 + 2.0 * r * sh * sh);  // 1 - r*cos(th), cancellation-free
	zim = (float)(-r * sin(th));
	rho = (float)(-expm1(2.0 * x));  // 1 - r^2
}

// vibratoSpeed (Hz) -> phase increment per tick in 2^-64 cycles
KLATT_HD int64_t vibratoIncrement(double speed, double srInv) {
	double t = speed * srInv;  // cycles per tick
	if (t != t) t = 0.0;
	if (!(t > -0.49 && t < 0.49)) t = fmod(t, 1.0);  // absurd speeds: only the fraction of a cycle matters
	if (t >= 0.5) t -= 1.0;
	if (t < -0.5) t += 1.0;
	return (int64_t)(t * 18446744073709551616.0);
}

KLATT_HD void planFade(const double *o, const double *n, uint32_t F, int sampleRate, FadePlanF32 &p) {
	const double srInv = 1.0 / (double)sampleRate;
	const double invF = 1.0 / (double)F;
	{
		double v0 = o[kVibratoSpeed], v1 = n[kVibratoSpeed];
		if (v1 != v1) v1 = v0;
		p.vibInc0 = vibratoIncrement(v0, srInv);
		p.vibIncFinal = vibratoIncrement(v0 + ((v1 - v0) * 1.0), srInv);
		p.vibIncStep = (p.vibIncFinal - p.vibInc0) / (int64_t)F;
	}
	for (int i = 0; i < kNumDirect; ++i) {
		double a = o[directParam(i)], b = n[directParam(i)];
		if (b != b) b = a;  // NaN target: keep
		p.dir0[i] = (float)a;
		p.dirStep[i] = (float)((b - a) * invF);
		p.dirFinal[i] = (float)(a + ((b - a) * 1.0));
	}
	for (int r = 0; r < kNumResonators; ++r) {
		double f0 = o[resFreqParam(r)], f1 = n[resFreqParam(r)];
		double b0 = o[resBwParam(r)], b1 = n[resBwParam(r)];
		if (f1 != f1) f1 = f0;
		if (b1 != b1) b1 = b0;
		poleTerms(f0, b0, srInv, p.z0re[r], p.z0im[r], p.rho0[r]);
		poleTerms(f0 + ((f1 - f0) * 1.0), b0 + ((b1 - b0) * 1.0), srInv, p.zFre[r], p.zFim[r], p.rhoF[r]);
		// per-tick ratio of the pole: q*exp(i*dth) = 1 - omega ; |ratio|^2 = 1 - kappa
		const double PI = 3.14159265358979323846;
		double xs = -PI * (b1 - b0) * invF * srInv;
		double dth = 2.0 * PI * (f1 - f0) * invF * srInv;
		double qm1 = expm1(xs);
		double q = 1.0 + qm1;
		double sh = sin(0.5 * dth);
		p.wre[r] = (float)(-qm1 + 2.0 * q * sh * sh);
		p.wim[r] = (float)(-q * sin(dth));
		p.kap[r] = (float)(-expm1(2.0 * xs));
		{  // the same pole ratio over kCoarseTicks ticks
			double xsA = xs * kCoarseTicks, dthA = dth * kCoarseTicks;
			double QAm1 = expm1(xsA), QA = 1.0 + QAm1, shA = sin(0.5 * dthA);
			p.Wre[r] = (float)(-QAm1 + 2.0 * QA * shA * shA);
			p.Wim[r] = (float)(-QA * sin(dthA));
			p.Kap[r] = (float)(-expm1(2.0 * xsA));
		}
		if (r == kResN0) {
			p.n0InvFade = !(f0 == 0 && f1 == 0);
			p.n0InvFinal = (f0 + ((f1 - f0) * 1.0)) != 0;
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// per-stream working set (registers on the device)
// ---------------------------------------------------------------------------------------------------
struct LiveF32 {
	float y[kNumResonators], d[kNumResonators];             // filter memories
	float zre[kNumResonators], zim[kNumResonators];         // zeta = 1 - pole
	float rho[kNumResonators], a[kNumResonators];           // 1-|pole|^2, |zeta|^2
	float dir[kNumDirect];
	float wre[kNumResonators], wim[kNumResonators], kap[kNumResonators];  // fade increments (resonators)
	float dir0[kNumDirect], dstep[kNumDirect];  // direct params during a fade: dir = dir0 + k*dstep (stateless: an
	                                            // accumulated "+= dstep" rounds the same way every tick and drifts)
	float invA0;     // 1/a of the anti-resonator
	float aspLast, fricLast;
	uint64_t vibratoPos;
	int64_t vibInc, vibIncStep;
	double pitchPos, pitch, pitchStep;
	bool n0Inv;
};

KLATT_HD float fastRcp(float x) {
#ifdef __CUDA_ARCH__
	return __fdividef(1.0f, x);
#else
	return 1.0f / x;
#endif
}

KLATT_HD void refreshA(LiveF32 &L) {
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) L.a[r] = fmaf(L.zre[r], L.zre[r], L.zim[r] * L.zim[r]);
	L.invA0 = fastRcp(L.a[kResN0]);
}

// one fade tick of every interpolated quantity (reference: 47 lerps + 14 setParams per tick, src/frame.cpp:50-52,
// src/speechWaveGenerator.cpp:113-125)
KLATT_HD void stepFade(LiveF32 &L, float k) {
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) {
		float zr = L.zre[r], zi = L.zim[r], wr = L.wre[r], wi = L.wim[r];
		float tr = fmaf(-zr, wr, wr);   // omega - zeta*omega, real
		tr = fmaf(zi, wi, tr);
		float ti = fmaf(-zr, wi, wi);   // imaginary
		ti = fmaf(-zi, wr, ti);
		L.zre[r] = zr + tr;
		L.zim[r] = zi + ti;
		L.rho[r] = fmaf(L.kap[r], 1.0f - L.rho[r], L.rho[r]);
	}
#pragma unroll
	for (int i = 0; i < kNumDirect; ++i) L.dir[i] = fmaf(k, L.dstep[i], L.dir0[i]);
	refreshA(L);
	L.pitch += L.pitchStep;
	L.vibInc += L.vibIncStep;
}

KLATT_HD void loadFadeStart(LiveF32 &L, const FadePlanF32 &p) {
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) { L.zre[r] = p.z0re[r]; L.zim[r] = p.z0im[r]; L.rho[r] = p.rho0[r]; }
#pragma unroll
	for (int i = 0; i < kNumDirect; ++i) L.dir[i] = p.dir0[i];
	L.vibInc = p.vibInc0;
}
KLATT_HD void loadFadeFinal(LiveF32 &L, const FadePlanF32 &p) {
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) { L.zre[r] = p.zFre[r]; L.zim[r] = p.zFim[r]; L.rho[r] = p.rhoF[r]; }
#pragma unroll
	for (int i = 0; i < kNumDirect; ++i) L.dir[i] = p.dirFinal[i];
	L.vibInc = p.vibIncFinal;
	refreshA(L);
}
KLATT_HD void loadFadeSteps(LiveF32 &L, const FadePlanF32 &p) {
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) { L.wre[r] = p.wre[r]; L.wim[r] = p.wim[r]; L.kap[r] = p.kap[r]; }
#pragma unroll
	for (int i = 0; i < kNumDirect; ++i) { L.dstep[i] = p.dirStep[i]; L.dir0[i] = p.dir0[i]; }
	L.vibIncStep = p.vibIncStep;
}

// Coarse (kCoarseTicks-tick) pole recurrence: 6 words per resonator that are touched once every kCoarseTicks
// samples, so they live outside the register file (shared memory on the device, [word][thread]).
//   Store::at(i) -> float&   i in [0, kCoarseWords)
constexpr int kCoarseWords = 6 * kNumResonators;  // zc_re, zc_im, rhoc, W_re, W_im, Kap

template <class Store>
KLATT_HD void coarseLoadSteps(Store &cs, const FadePlanF32 &p) {
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) {
		cs.at(3 * kNumResonators + r) = p.Wre[r];
		cs.at(4 * kNumResonators + r) = p.Wim[r];
		cs.at(5 * kNumResonators + r) = p.Kap[r];
	}
}
template <class Store>
KLATT_HD void coarseCapture(Store &cs, const LiveF32 &L) {
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) {
		cs.at(r) = L.zre[r];
		cs.at(kNumResonators + r) = L.zim[r];
		cs.at(2 * kNumResonators + r) = L.rho[r];
	}
}
template <class Store>
KLATT_HD void coarseAdvance(Store &cs, LiveF32 &L) {
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) {
		float zr = cs.at(r), zi = cs.at(kNumResonators + r), rh = cs.at(2 * kNumResonators + r);
		float wr = cs.at(3 * kNumResonators + r), wi = cs.at(4 * kNumResonators + r), kp = cs.at(5 * kNumResonators + r);
		float tr = fmaf(-zr, wr, wr);
		tr = fmaf(zi, wi, tr);
		float ti = fmaf(-zr, wi, wi);
		ti = fmaf(-zi, wr, ti);
		zr += tr;
		zi += ti;
		rh = fmaf(kp, 1.0f - rh, rh);
		cs.at(r) = zr; cs.at(kNumResonators + r) = zi; cs.at(2 * kNumResonators + r) = rh;
		L.zre[r] = zr; L.zim[r] = zi; L.rho[r] = rh;
	}
	refreshA(L);
}

// delta-form two-pole section; returns the new output
KLATT_HD float resonate(LiveF32 &L, int r, float x) {
	float w = fmaf(-L.rho[r], L.d[r], L.d[r]);  // (1-rho)*d
	w = fmaf(-L.a[r], L.y[r], w);
	float dn = fmaf(L.a[r], x, w);
	L.d[r] = dn;
	L.y[r] += dn;
	return L.y[r];
}

KLATT_HD float sinTurns(float t) {  // sin(2*pi*t), |t| <= 0.5
#ifdef __CUDA_ARCH__
	return sinpif(2.0f * t);
#else
	return sinf(6.283185307179586f * t);
#endif
}

// one generated sample (reference src/speechWaveGenerator.cpp:203-208); uA/uF are the two uniform draws in [0,1]
KLATT_HD int dspTick(LiveF32 &L, float uA, float uF, double srInv) {
	// ---- vibrato (:73) and glottal phase (:74, :55) ----
	L.vibratoPos += (uint64_t)L.vibInc;
	float vph = (float)(int32_t)(uint32_t)(L.vibratoPos >> 32) * 2.3283064365386963e-10f;  // cycles in [-0.5, 0.5)
	float vib = sinTurns(vph) * 0.06f * L.dir[dVibratoPitchOffset];
	double base = L.pitch * srInv;
#ifdef KLATT_EXPERIMENT_VIB64
	double vibd = sin((double)(int64_t)L.vibratoPos * (6.283185307179586 / 18446744073709551616.0)) * 0.06 * (double)L.dir[dVibratoPitchOffset];
	double pos = L.pitchPos + fma(base, vibd, base);
#else
	double pos = L.pitchPos + fma(base, (double)vib, base);
#endif
	if (pos >= 1.0) pos -= 1.0;
	if (!(pos >= 0.0 && pos < 1.0)) pos = fmod(pos, 1.0);  // negative or absurd pitch: the reference's fmod semantics
	L.pitchPos = pos;
	float voice = (float)pos;
	// ---- aspiration noise + turbulence (:40, :75-80) ----
	L.aspLast = fmaf(0.75f, L.aspLast, uA);
	float asp = L.aspLast * 0.2f;
	float turb = asp * L.dir[dVoiceTurbulenceAmplitude];
	if (voice < L.dir[dGlottalOpenQuotient]) turb *= 0.01f;
	float v = (fmaf(voice, 2.0f, -1.0f) + turb) * L.dir[dVoiceAmplitude];
	float src = fmaf(asp, L.dir[dAspirationAmplitude], v);
	const float halfGain = L.dir[dPreFormantGain] * 0.5f;
	// ---- cascade (:147-158) ----
	float ci = src * halfGain;
	float dx = ci - L.y[kResN0];  // anti-resonator: memories hold INPUTS (:133)
	float dx1 = fmaf(-L.rho[kResN0], L.d[kResN0], L.d[kResN0]);  // (1-rho) * previous input difference
	float n0 = L.n0Inv ? fmaf(dx - dx1, L.invA0, L.y[kResN0])
	                   : fmaf(L.a[kResN0], dx, dx1 + L.y[kResN0]);
	L.d[kResN0] = dx;
	L.y[kResN0] = ci;
	float np = resonate(L, kResNP, n0);
	float x = fmaf(np - ci, L.dir[dCaNP], ci);
#pragma unroll
	for (int r = kResCascade; r < kResParallel; ++r) x = resonate(L, r, x);
	// ---- frication noise + parallel bank (:205-206, :170-180) ----
	L.fricLast = fmaf(0.75f, L.fricLast, uF);
	float pin = (L.fricLast * 0.3f) * L.dir[dFricationAmplitude] * halfGain;
	float par = 0.0f;
#pragma unroll
	for (int k = 0; k < 6; ++k) par = fmaf(resonate(L, kResParallel + k, pin) - pin, L.dir[dPa1 + k], par);
	par = fmaf(pin - par, L.dir[dParallelBypass], par);
	// ---- mix, gain, clamp with the Win32 macro NaN behaviour (NaN -> +32000), truncate (:207-208) ----
	float s = (x + par) * L.dir[dOutputGain] * 4000.0f;
	s = fminf(s, 32000.0f);  // fminf(NaN, 32000) == 32000
	s = fmaxf(s, -32000.0f);
	return (int)s;
}

// ---------------------------------------------------------------------------------------------------
// renderStreamF32: advance one stream by up to sampleCount ticks.  Returns the number of samples produced.
// `plans` (may be null) holds precomputed FadePlanF32 for requests [qBase, qBase+qCount) whose predecessor was
// known at plan time; planValidFrom is the first absolute request index for which they may be used.
// ---------------------------------------------------------------------------------------------------
template <class Out, class Store>
KLATT_HD uint32_t renderStreamF32(const StreamDesc &desc, int sampleRate, uint32_t sampleCount, Out &out, Store &cs,
                                  const NoiseConfig &noise, int32_t *lastUserIndexOut, uint32_t *qHeadOut) {
	StreamState *st = desc.state;
	FrameMgrState &fm = st->fm;
	GenStateF32 &gs = st->gen.f32;
	const double srInv = 1.0 / (double)sampleRate;

	uint32_t counter = fm.counter, qHead = fm.qHead;
	uint32_t oldM = fm.oldM, newM = fm.newM, newF = fm.newF;
	int32_t lastUserIndex = fm.lastUserIndex;
	bool hasNew = fm.hasNew, curIsNull = fm.curIsNull, oldIsNull = fm.oldIsNull, newIsNull = fm.newIsNull;
	double oldInc = fm.oldInc, newInc = fm.newInc;

	LiveF32 L;
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) {
		L.y[r] = gs.y[r]; L.d[r] = gs.d[r]; L.zre[r] = gs.zre[r]; L.zim[r] = gs.zim[r]; L.rho[r] = gs.rho[r];
	}
#pragma unroll
	for (int i = 0; i < kNumDirect; ++i) L.dir[i] = gs.dir[i];
	L.aspLast = gs.aspLast; L.fricLast = gs.fricLast; L.vibratoPos = gs.vibratoPos; L.vibInc = gs.vibInc; L.vibIncStep = 0;
	L.pitchPos = gs.pitchPos; L.pitch = fm.curFrame[kVoicePitch]; L.pitchStep = gs.pitchStep;
	L.n0Inv = gs.n0Inv != 0;
	refreshA(L);
	uint64_t gen = gs.samplesGenerated;

	// purge prologue, src/frame.cpp:103-112 (the dropped requests were removed on the host).  fm.curFrame is kept
	// current at every exit, so the snapshot is available here.
	if (fm.purgePending) {
		fm.purgePending = 0;
		counter = oldM;
		if (hasNew) {
			oldIsNull = newIsNull;
			for (int i = 0; i < kNumParams; ++i) fm.oldFrame[i] = fm.curFrame[i];
			hasNew = false;
		}
	}
	const FadePlanF32 *plan = &gs.plan;
	if (hasNew) {
		loadFadeSteps(L, *plan);
		coarseLoadSteps(cs, *plan);
		for (int i = 0; i < 3 * kNumResonators; ++i) cs.at(i) = gs.coarse[i];
	}
	uint32_t coarseAt = hasNew ? gs.coarseAt : 0x80000000u;  // fade tick of the last coarse sync (0x80000000: none)

	Philox4 blk;
	blk.w[0] = blk.w[1] = blk.w[2] = blk.w[3] = 0;
	uint64_t blkIndex = ~0ull;
	const float kInvRandMax = 1.0f / 2147483647.0f;

	uint32_t produced = 0;
	for (; produced < sampleCount; ++produced) {
		// ================= frame manager tick, src/frame.cpp:41-80 =================
		counter++;
		if (hasNew) {
			if (counter > newF) {  // :44-47 the fade is over: new becomes old; cur keeps its ratio-1 value
				for (int i = 0; i < kNumParams; ++i) fm.oldFrame[i] = fm.newFrame[i];
				oldM = newM; oldInc = newInc; oldIsNull = newIsNull;
				hasNew = false;
			} else {  // :49-52
				if (counter == 1) {  // the fade starts from the (possibly rewritten) old frame, not from the stale cur
					loadFadeStart(L, *plan);
					L.pitch = fm.oldFrame[kVoicePitch];
				}
				stepFade(L, (float)counter);
#ifdef KLATT_EXPERIMENT_EXACT_COEF
				{
					double ratio = (double)counter / (double)newF;
					for (int r = 0; r < kNumResonators; ++r) {
						double f0 = fm.oldFrame[resFreqParam(r)], f1 = fm.newFrame[resFreqParam(r)];
						double b0 = fm.oldFrame[resBwParam(r)], b1 = fm.newFrame[resBwParam(r)];
						poleTerms(f0 + (f1 - f0) * ratio, b0 + (b1 - b0) * ratio, srInv, L.zre[r], L.zim[r], L.rho[r]);
					}
					for (int i = 0; i < kNumDirect; ++i) {
						double a = fm.oldFrame[directParam(i)], b = fm.newFrame[directParam(i)];
						L.dir[i] = (float)(a + (b - a) * ratio);
					}
					refreshA(L);
				}
#endif
				if (counter < newF && (gen & (uint64_t)(kCoarseTicks - 1)) == 0) {
					// drift control, on a grid of ABSOLUTE sample indices (identical for every chunking of the
					// render, and the same loop iteration for every lane of a batch that started together)
					if (counter - coarseAt == (uint32_t)kCoarseTicks) coarseAdvance(cs, L);
					else coarseCapture(cs, L);
					coarseAt = counter;
				}
				if (counter == newF) {  // ratio == 1: land exactly on the planned end values
					loadFadeFinal(L, *plan);
					L.pitch = fm.oldFrame[kVoicePitch] + ((fm.newFrame[kVoicePitch] - fm.oldFrame[kVoicePitch]) * 1.0);
					L.n0Inv = plan->n0InvFinal != 0;
				} else {
					L.n0Inv = plan->n0InvFade != 0;
				}
			}
		} else if (counter > oldM) {  // :54
			uint32_t rel = qHead - desc.qBase;
			if (rel < desc.qCount) {  // :55-72
				curIsNull = false;
				// keep the snapshot invariants: cur == what the hold rendered with, old pitch follows the glide (:78)
				fm.oldFrame[kVoicePitch] = L.pitch;
				for (int i = 0; i < kNumParams; ++i) fm.curFrame[i] = fm.oldFrame[i];
				newM = desc.minDur[rel];
				uint32_t fd = desc.fadeDur[rel];
				newF = fd > 1u ? fd : 1u;  // src/speechPlayer.cpp:36
				newIsNull = desc.isNull ? (desc.isNull[rel] != 0) : false;
				int32_t ux = desc.userIndex ? desc.userIndex[rel] : -1;
				qHead++;
				hasNew = true;
				if (newIsNull) {  // :59-63
					for (int i = 0; i < kNumParams; ++i) fm.newFrame[i] = fm.oldFrame[i];
					fm.newFrame[kPreFormantGain] = 0;
					fm.newFrame[kVoicePitch] = L.pitch;
					newInc = 0;
				} else {
					const double *fr = desc.frames + (size_t)rel * kNumParams;
					for (int i = 0; i < kNumParams; ++i) fm.newFrame[i] = fr[i];
					newInc = (fr[kEndVoicePitch] - fr[kVoicePitch]) / (double)newM;  // src/frame.cpp:98
					if (oldIsNull) {  // :64-67
						for (int i = 0; i < kNumParams; ++i) fm.oldFrame[i] = fm.newFrame[i];
						fm.oldFrame[kPreFormantGain] = 0;
					}
				}
				if (ux != -1) lastUserIndex = ux;  // :69
				counter = 0;                       // :70
				fm.newFrame[kVoicePitch] += (newInc * (double)newF);  // :71
				// plan the fade (double precision, once per request)
				planFade(fm.oldFrame, fm.newFrame, newF, sampleRate, gs.plan);
				plan = &gs.plan;
				loadFadeSteps(L, *plan);
				coarseLoadSteps(cs, *plan);
				coarseAt = 0x80000000u;
				double tgt = fm.newFrame[kVoicePitch], from = fm.oldFrame[kVoicePitch];
				L.pitchStep = (tgt != tgt) ? 0.0 : (tgt - from) / (double)newF;
				// NOTE: this tick still renders with the stale working set (cur is untouched on the pop tick)
			} else {
				curIsNull = true;  // :73-75
			}
		} else {  // :76-79 hold: only the pitch glides
			L.pitch += oldInc;
		}
		if (curIsNull) break;  // src/speechWaveGenerator.cpp:210

		// ================= noise draws (two per generated sample) =================
		uint32_t drawA, drawF;
		if (noise.mode == kNoisePhilox) {
			uint64_t b = gen >> 1;
			if (b != blkIndex) { blk = noiseBlock(noise.seed, desc.streamId, b); blkIndex = b; }
			bool odd = (gen & 1ull) != 0;
			drawA = (odd ? blk.w[2] : blk.w[0]) >> 1;
			drawF = (odd ? blk.w[3] : blk.w[1]) >> 1;
		} else {
			uint64_t d0 = 2 * gen - desc.replayBase;
			drawA = (d0 < desc.replayLen) ? (uint32_t)desc.replay[d0] : 0u;
			drawF = (d0 + 1 < desc.replayLen) ? (uint32_t)desc.replay[d0 + 1] : 0u;
		}
		gen++;
#ifdef KLATT_HOSTSIM_DEBUG
		if (g_dbgPhase) g_dbgPhase[gen - 1] = L.pitchPos;
#endif
		out.push(dspTick(L, (float)drawA * kInvRandMax, (float)drawF * kInvRandMax, srInv));
	}

	// ---- store the stream back; keep fm.curFrame / fm.oldFrame meaningful for purge and for the next plan ----
	if (hasNew) {
		if (counter >= 1) {
			double ratio = (double)counter / (double)newF;
			for (int i = 0; i < kNumParams; ++i) {
				double o = fm.oldFrame[i], n = fm.newFrame[i];
				fm.curFrame[i] = (n != n) ? o : o + ((n - o) * ratio);
			}
		}
	} else {
		for (int i = 0; i < kNumParams; ++i) fm.curFrame[i] = fm.oldFrame[i];
		fm.oldFrame[kVoicePitch] = L.pitch;
	}
	fm.curFrame[kVoicePitch] = L.pitch;
	fm.counter = counter; fm.qHead = qHead; fm.oldM = oldM; fm.newM = newM; fm.newF = newF;
	fm.lastUserIndex = lastUserIndex;
	fm.hasNew = hasNew; fm.curIsNull = curIsNull; fm.oldIsNull = oldIsNull; fm.newIsNull = newIsNull;
	fm.oldInc = oldInc; fm.newInc = newInc;
#pragma unroll
	for (int r = 0; r < kNumResonators; ++r) {
		gs.y[r] = L.y[r]; gs.d[r] = L.d[r]; gs.zre[r] = L.zre[r]; gs.zim[r] = L.zim[r]; gs.rho[r] = L.rho[r];
	}
#pragma unroll
	for (int i = 0; i < kNumDirect; ++i) gs.dir[i] = L.dir[i];
	gs.aspLast = L.aspLast; gs.fricLast = L.fricLast; gs.vibratoPos = L.vibratoPos; gs.vibInc = L.vibInc;
	gs.pitchPos = L.pitchPos; gs.pitchStep = L.pitchStep; gs.n0Inv = L.n0Inv ? 1u : 0u;
	gs.samplesGenerated = gen;
	gs.coarseAt = coarseAt;
	if (hasNew)
		for (int i = 0; i < 3 * kNumResonators; ++i) gs.coarse[i] = cs.at(i);
	*lastUserIndexOut = lastUserIndex;
	*qHeadOut = qHead;
	return produced;
}

}  // namespace klatt
