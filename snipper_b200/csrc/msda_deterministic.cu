// Deterministic (atomics-free, bit-reproducible) backward -- placeholder until the two-pass
// implementation lands; the C ABI reports MSDA_ERR_CUDA(cudaErrorNotSupported) meanwhile.
#include "msda_common.cuh"
#include "msda_internal.h"

namespace msda {

size_t deterministic_workspace_bytes(const OpDims &) { return 256; }

cudaError_t launch_backward_deterministic_f32(const float *, const int64_t *, const int64_t *,
                                              const float *, const float *, const float *, float *,
                                              float *, float *, const OpDims &, void *, cudaStream_t)
{
    return cudaErrorNotSupported;
}

}  // namespace msda
