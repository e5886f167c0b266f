// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE.  extern "C" entry points around the REFERENCE's own
// cutlass_preprocessors.cpp, which oracle/Makefile compiles in place from /root/reference (no reference
// source is copied into this repo).  Output: oracle/_ref/libref_quant.so (git-ignored, travels with gpurun).
// Used to (a) pin oracle/woq_oracle.c bit-for-bit and (b) generate tests/golden/*.npz.
#include "tensorrt_llm/kernels/cutlass_kernels/cutlass_preprocessors.h"
#include <cuda_fp16.h>
#include <cstdint>
#include <vector>

using namespace tensorrt_llm::kernels::cutlass_kernels;

extern "C"
{

// T/cpp/tensorrt_llm/kernels/cutlass_kernels/cutlass_preprocessors.cpp:615-721, <half, half>
int ref_symmetric_quantize_f16(const void* w, int K, int N, int8_t* raw, int8_t* proc, void* scales)
{
    try
    {
        symmetric_quantize<half, half>(proc, raw, reinterpret_cast<half*>(scales), reinterpret_cast<const half*>(w),
            std::vector<size_t>{size_t(K), size_t(N)}, QuantType::INT8_WEIGHT_ONLY);
    }
    catch (...)
    {
        return -1;
    }
    return 0;
}

// same, <half, float>: fp32 weights, fp16 scales
int ref_symmetric_quantize_f32w_f16s(const float* w, int K, int N, int8_t* raw, int8_t* proc, void* scales)
{
    try
    {
        symmetric_quantize<half, float>(proc, raw, reinterpret_cast<half*>(scales), w,
            std::vector<size_t>{size_t(K), size_t(N)}, QuantType::INT8_WEIGHT_ONLY);
    }
    catch (...)
    {
        return -1;
    }
    return 0;
}

// :537-578
int ref_preprocess_weights(const int8_t* raw, int K, int N, int8_t* proc)
{
    try
    {
        preprocess_weights_for_mixed_gemm(proc, raw, std::vector<size_t>{size_t(K), size_t(N)},
            QuantType::INT8_WEIGHT_ONLY);
    }
    catch (...)
    {
        return -1;
    }
    return 0;
}

// individual steps :158-219, :364-381, :383-405 (arch forced to 80 like the shim)
int ref_permute_rows(const int8_t* in, int K, int N, int8_t* out)
{
    try
    {
        permute_B_rows_for_mixed_gemm(out, in, std::vector<size_t>{size_t(K), size_t(N)}, QuantType::INT8_WEIGHT_ONLY, 80);
    }
    catch (...)
    {
        return -1;
    }
    return 0;
}

int ref_transpose(const int8_t* in, int K, int N, int8_t* out)
{
    try
    {
        subbyte_transpose(out, in, std::vector<size_t>{size_t(K), size_t(N)}, QuantType::INT8_WEIGHT_ONLY);
    }
    catch (...)
    {
        return -1;
    }
    return 0;
}

int ref_add_bias_interleave(int8_t* inout, size_t n)
{
    try
    {
        add_bias_and_interleave_quantized_tensor_inplace(inout, n, QuantType::INT8_WEIGHT_ONLY);
    }
    catch (...)
    {
        return -1;
    }
    return 0;
}
}
