// oracle/ref_gpu_harness.cu -- TEST INFRASTRUCTURE, not product code.
// C entry points over the REFERENCE's own CUDA kernels, compiled for sm_100a from the sources where they lie under
// /root/reference (oracle/Makefile target `refgpu`; output oracle/_ref/libref_gpu.so, git-ignored):
//   * weight_only_gemv_launcher<int8_t, half>   T/cpp/tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.cu:371-378
//   * mmha::mmha_launch_kernel<uint16_t, KVLinearBuffer, ..., 64>
//                                                T/cpp/tensorrt_llm/kernels/decoderMaskedMultiheadAttention/
//                                                decoderMaskedMultiheadAttention64_half.cu (Launch.h:177-188)
// Used by tests/test_reference_kernels_gpu.py to pin (a) the C restatement of the GEMV arithmetic and (b) this repo's
// GEMV / MMHA kernels against the reference kernels running on the same B200, and by bench.py's optional
// reference-kernel timings.  The parameter block is filled the way the plugin does it
// (T/cpp/tensorrt_llm/plugins/gptAttentionCommon/gptAttentionCommon.cpp:108-206, 650-780).
#include "tensorrt_llm/kernels/decoderMaskedMultiheadAttention.h"
#include "tensorrt_llm/kernels/kvCacheUtils.h"
#include "tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.h"

#include <cmath>
#include <cstring>

namespace tensorrt_llm
{
namespace kernels
{
namespace mmha
{
template <typename T, typename KVCacheBuffer, typename KernelParamsType, int Dh>
void mmha_launch_kernel(const KernelParamsType& params, const KVCacheBuffer& kv_cache_buffer, const cudaStream_t& stream);
extern template void mmha_launch_kernel<uint16_t, KVLinearBuffer, Masked_multihead_attention_params<uint16_t>, 64>(
    const Masked_multihead_attention_params<uint16_t>&, const KVLinearBuffer&, const cudaStream_t&);
} // namespace mmha
} // namespace kernels
} // namespace tensorrt_llm

using namespace tensorrt_llm::kernels;

extern "C" int ref_gpu_gemv(const void* x, const void* w_processed, const void* scales, const void* bias, void* out, int k,
    int n, void* stream)
{
    weight_only_gemv_launcher<int8_t, half>(static_cast<const half*>(x), static_cast<const int8_t*>(w_processed),
        static_cast<const half*>(scales), static_cast<const half*>(bias), static_cast<half*>(out), k, n,
        ActivationType::Identity, QuantType::INT8_WEIGHT_ONLY, static_cast<cudaStream_t>(stream));
    return (int) cudaGetLastError();
}

// qkv [B, 3*H*64] fp16; out [B, H*64]; kv_cache [B, 2, H, Smax, 64] (int8 or fp16); sequence_lengths [B] (device) = keys
// already cached per sequence; past_len = the host copy of the same value (the plugin's past_key_value_length[0]);
// total_padding [B] device ints (zeros when the prompts were not padded).
extern "C" int ref_gpu_mmha(const void* qkv, void* out, void* kv_cache, const int* sequence_lengths, const int* masked_tokens,
    const int* total_padding, const float* kv_scale_orig_quant, const float* kv_scale_quant_orig, int batch, int heads,
    int max_seq_len, int past_len, int max_input_len, int int8_kv, float q_scaling, void* stream)
{
    const int Dh = 64, hidden = heads * Dh;
    Masked_multihead_attention_params<uint16_t> p;
    memset(&p, 0, sizeof(p));
    p.out = static_cast<uint16_t*>(out);
    p.q = static_cast<const uint16_t*>(qkv);
    p.k = p.q + hidden;
    p.v = p.q + 2 * hidden;
    p.stride = 3 * hidden;
    if (int8_kv)
    {
        p.kv_scale_orig_quant = kv_scale_orig_quant;
        p.kv_scale_quant_orig = kv_scale_quant_orig;
    }
    p.int8_kv_cache = int8_kv != 0;
    p.batch_size = batch;
    p.beam_width = 1;
    p.memory_max_len = max_seq_len;
    p.length_per_sample = sequence_lengths;
    p.timestep = past_len; // step + max_prefix_prompt_length - 1 with step = past_kv_length + 1
    p.num_heads = heads;
    p.hidden_size_per_head = Dh;
    p.inv_sqrt_dh = 1.f / (sqrtf((float) Dh) * q_scaling);
    p.total_padding_tokens = total_padding;
    p.masked_tokens = masked_tokens;
    p.max_input_length = max_input_len;
    const int elem = int8_kv ? 1 : 2;
    KVLinearBuffer kv(batch, 1, max_seq_len, heads * Dh * elem);
    kv.data = static_cast<int8_t*>(kv_cache);
    mmha::mmha_launch_kernel<uint16_t, KVLinearBuffer, Masked_multihead_attention_params<uint16_t>, 64>(
        p, kv, static_cast<cudaStream_t>(stream));
    return (int) cudaGetLastError();
}
