#include <cuda_runtime.h>
#include "klatt_common.h"
namespace klatt {
cudaError_t launchKlattF32(const StreamDesc *, uint32_t, int, uint32_t, int16_t *, size_t, uint32_t *, StreamResult *,
                           NoiseConfig, cudaStream_t) {
	return cudaErrorNotSupported;
}
}
