// oracle/ref_cudart_shim.cpp -- TEST INFRASTRUCTURE.  The reference's preprocess_weights_for_mixed_gemm asks
// the CUDA runtime for the SM version (T/cpp/tensorrt_llm/common/cudaUtils.h:230-239) and throws
// "Unsupported Arch" for anything above sm_90 (cutlass_preprocessors.cpp:130-150).  This shim answers
// "sm_80" so the reference's host code can run on a GPU-less container and on a B200 box alike.
#include <cuda_runtime_api.h>

extern "C"
{
cudaError_t cudaGetDevice(int* d)
{
    *d = 0;
    return cudaSuccess;
}

cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr attr, int)
{
    *v = (attr == cudaDevAttrComputeCapabilityMajor) ? 8 : 0;
    return cudaSuccess;
}

const char* cudaGetErrorString(cudaError_t)
{
    return "shim";
}

cudaError_t cudaGetLastError()
{
    return cudaSuccess;
}

cudaError_t cudaDeviceSynchronize()
{
    return cudaSuccess;
}

cudaError_t cudaMemcpy(void*, const void*, size_t, cudaMemcpyKind)
{
    return cudaSuccess;
}
}
