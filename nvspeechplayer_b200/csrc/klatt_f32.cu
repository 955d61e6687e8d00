// klatt_batch_f32: the production kernel.  One stream per thread; the whole per-stream working set (28 filter
// memories, 56 coefficient-state words, 17 direct parameters, fade increments, FP64 pitch/phase) lives in
// registers, the frame queue and the per-request fade plans are read from HBM only on pop ticks, and the int16
// output leaves as 16-byte stores (out_writer.cuh).  The arithmetic is renderStreamF32() in klatt_f32_core.cuh.
#include <cuda_runtime.h>
#include "klatt_common.h"
#include "klatt_f32_core.cuh"
#include "out_writer.cuh"

namespace klatt {

constexpr int kF32Block = 128;

namespace {
struct SmemCoarse {
	float *base;  // this thread's column of the [kCoarseWords][kF32Block] block
	__device__ __forceinline__ float &at(int i) { return base[i * kF32Block]; }
};
}  // namespace

__global__ void __launch_bounds__(kF32Block)
klatt_batch_f32_kernel(const StreamDesc *__restrict__ descs, uint32_t numStreams, int sampleRate, uint32_t sampleCount,
                       int16_t *__restrict__ out, size_t rowStride, uint32_t *__restrict__ samplesWritten,
                       StreamResult *__restrict__ results, NoiseConfig noise) {
	const uint32_t s = blockIdx.x * kF32Block + threadIdx.x;
	if (s >= numStreams) return;
	const StreamDesc desc = descs[s];
	__shared__ float coarse[kCoarseWords * kF32Block];
	SmemCoarse cs;
	cs.base = coarse + threadIdx.x;
	OutWriter ow;
	int16_t *row = out + (size_t)s * rowStride;
	ow.init(row, ((reinterpret_cast<uintptr_t>(row) & 15u) == 0));
	int32_t lastUserIndex;
	uint32_t qHead;
	uint32_t produced = renderStreamF32(desc, sampleRate, sampleCount, ow, cs, noise, &lastUserIndex, &qHead);
	ow.flush();
	for (uint32_t i = produced; i < sampleCount; ++i) row[i] = 0;  // drained: the rest of the row is silence
	if (samplesWritten) samplesWritten[s] = produced;
	if (results) {
		StreamResult res;
		res.written = produced; res.lastUserIndex = lastUserIndex; res.qHead = qHead; res.pad = 0;
		results[s] = res;
	}
}

cudaError_t launchKlattF32(const StreamDesc *descs, uint32_t numStreams, int sampleRate, uint32_t sampleCount,
                           int16_t *out, size_t rowStride, uint32_t *samplesWritten, StreamResult *results,
                           NoiseConfig noise, cudaStream_t stream) {
	if (numStreams == 0 || sampleCount == 0) return cudaSuccess;
	dim3 grid((numStreams + kF32Block - 1) / kF32Block);
	klatt_batch_f32_kernel<<<grid, kF32Block, 0, stream>>>(descs, numStreams, sampleRate, sampleCount, out, rowStride,
	                                                       samplesWritten, results, noise);
	return cudaGetLastError();
}

}  // namespace klatt
