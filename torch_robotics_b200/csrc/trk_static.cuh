// Register-resident kernel family for compiled-in robots (filled in below the generic family).
#pragma once
#include <cuda_runtime.h>

#include "trk_device.cuh"

namespace trk {

inline int match_static_family(const float* /*tree_words*/, int /*nwords*/) { return 0; }
inline const char* static_family_name(int family) { return family == 0 ? "generic" : "?"; }
inline bool static_objective_supported(int /*family*/) { return false; }
inline cudaError_t launch_static_fk_forward(int, int, const float*, long long, const SlotMap&, int, float*, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t launch_static_fk_backward(int, int, const float*, long long, const SlotMap&, int, const float*, float*, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t launch_static_objective(int, int, const float*, int, const float*, long long, const float*, float, int, float*, float*, float*, uint8_t*, int*, cudaStream_t) { return cudaErrorNotSupported; }

}  // namespace trk
