// cuda_utils.h -- device bring-up with the reference's entry point name
// (src/utils/cuda_utils.h:49, cuda_utils.cpp:21-57).  The library uses the CUDA runtime, so the
// "context" handed back is the device's primary context; callers that go on to call cuCtxDestroy on
// it (src/main.cpp:236, result ignored there) may simply drop that call.
#ifndef FLOW3D_CUDA_UTILS_H_
#define FLOW3D_CUDA_UTILS_H_

typedef struct CUctx_st* CUcontext;  // same opaque type as <cuda.h>

bool InitCudaContextWithFirstAvailableDevice(CUcontext* cu_context);

#endif  // FLOW3D_CUDA_UTILS_H_
