// GPTAttentionPlugin -- TensorRT IPluginV2DynamicExt with the reference's identity and contract
// (T/cpp/tensorrt_llm/plugins/gptAttentionPlugin/gptAttentionPlugin.h:39-194 and
//  T/cpp/tensorrt_llm/plugins/gptAttentionCommon/gptAttentionCommon.h:36-147):
//   type "GPTAttention", version "1", namespace "tensorrt_llm";
//   16 creator fields (T/tensorrt_llm/functional.py:2876-2934): num_heads head_size unidirectional q_scaling
//     rotary_embedding_dim neox_rotary_style context_fmha_type multi_block_mode multi_query_mode int8_kv_cache
//     fp8_kv_cache remove_input_padding mask_type paged_kv_cache type_id in_flight_batching;
//   inputs by index (gptAttentionPlugin.h:105-168): 0 qkv, 1 past_key_value, 2 sequence_length, 3 past_key_value_length
//     (HOST [past_len, is_context]), 4 masked_tokens, 5 input_lengths, 6 max_input_length (shape only),
//     7 cache_indirection (Smax = dims[2]), 8 kv_orig_quant_scale, 9 kv_quant_orig_scale (int8/fp8 KV only), then
//     kv_cache_block_pointers [B, beam, 2, 2 * max_blocks_per_seq] int32 pairs when paged_kv_cache (input 1 is then the
//     block pool [blocks, 2, H, tokens_per_block, Dh]);
//   outputs: 0 context [B, S, H*Dh], 1 present_key_value (same buffer as input 1);
//   serialization: the 37-byte common block (gptAttentionCommon.cpp:862-890) || bool inFlightBatching = 38 bytes.
// Only the Whisper hot-path configuration executes on B200 (fp16 I/O, contiguous or paged int8 / fp16 KV cache, beam 1, no
// rotary / multi-query / in-flight batching); other configurations are constructible and serializable (so
// engines round-trip) but enqueue reports an error instead of computing something else.
#pragma once

#include "pluginCommon.h"

#include <string>
#include <vector>

namespace nvinfer1
{
namespace plugin
{

class GPTAttentionPlugin : public IPluginV2DynamicExt
{
public:
    GPTAttentionPlugin() = delete;
    GPTAttentionPlugin(int num_heads, int head_size, int unidirectional, float q_scaling, int rotary_embedding_dim,
        bool neox_rotary_style, int context_fmha_type, bool multi_block_mode, bool multi_query_mode, bool int8_kv_cache,
        bool fp8_kv_cache, bool remove_input_padding, int mask_type, bool paged_kv_cache, nvinfer1::DataType type,
        bool in_flight_batching);
    GPTAttentionPlugin(const void* data, size_t length);
    ~GPTAttentionPlugin() override = default;

    nvinfer1::IPluginV2DynamicExt* clone() const noexcept override;
    nvinfer1::DimsExprs getOutputDimensions(int outputIndex, const nvinfer1::DimsExprs* inputs, int nbInputs,
        nvinfer1::IExprBuilder& exprBuilder) noexcept override;
    bool supportsFormatCombination(
        int pos, const nvinfer1::PluginTensorDesc* inOut, int nbInputs, int nbOutputs) noexcept override;
    void configurePlugin(const nvinfer1::DynamicPluginTensorDesc* in, int nbInputs,
        const nvinfer1::DynamicPluginTensorDesc* out, int nbOutputs) noexcept override;
    size_t getWorkspaceSize(const nvinfer1::PluginTensorDesc* inputs, int nbInputs,
        const nvinfer1::PluginTensorDesc* outputs, int nbOutputs) const noexcept override;
    int enqueue(const nvinfer1::PluginTensorDesc* inputDesc, const nvinfer1::PluginTensorDesc* outputDesc,
        const void* const* inputs, void* const* outputs, void* workspace, cudaStream_t stream) noexcept override;
    nvinfer1::DataType getOutputDataType(
        int index, const nvinfer1::DataType* inputTypes, int nbInputs) const noexcept override;
    const char* getPluginType() const noexcept override;
    const char* getPluginVersion() const noexcept override;
    int getNbOutputs() const noexcept override;
    int initialize() noexcept override;
    void terminate() noexcept override;
    size_t getSerializationSize() const noexcept override;
    void serialize(void* buffer) const noexcept override;
    void destroy() noexcept override;
    void setPluginNamespace(const char* pluginNamespace) noexcept override;
    const char* getPluginNamespace() const noexcept override;

    static constexpr size_t kSerializedSize = 38;

private:
    // input slots (gptAttentionPlugin.h:105-168)
    enum : int
    {
        kQKV = 0,
        kPAST_KV = 1,
        kSEQUENCE_LENGTH = 2,
        kPAST_KV_LENGTH = 3,
        kMASKED_TOKENS = 4,
        kINPUT_LENGTHS = 5,
        kMAX_INPUT_LENGTH = 6,
        kCACHE_INDIR = 7,
        kKV_QUANT_SCALE = 8,
        kKV_DEQUANT_SCALE = 9
    };

    const char* unsupportedReason() const;
    int blockPointersIdx() const // getKVCacheBlockPointersIdx, gptAttentionPlugin.h:150-153
    {
        return mInt8KVCache ? 10 : 8;
    }

    std::string mNamespace;
    int mNumHeads, mHeadSize, mUnidirectional;
    float mQScaling;
    int mRotaryEmbeddingDim;
    bool mNeoxRotaryStyle, mEnableContextFMHA, mFMHAForceFP32Acc, mMultiBlockMode, mMultiQueryMode, mInt8KVCache,
        mFp8KVCache, mRemovePadding;
    int mMaskType; // AttentionMaskType: 0 padding, 1 causal, 2 bidirectional (T/cpp/tensorrt_llm/kernels/gptKernels.h:26-34)
    bool mPagedKVCache;
    nvinfer1::DataType mType;
    bool mInFlightBatching;
};

class GPTAttentionPluginCreator : public IPluginCreator
{
public:
    GPTAttentionPluginCreator();
    const char* getPluginName() const noexcept override;
    const char* getPluginVersion() const noexcept override;
    const nvinfer1::PluginFieldCollection* getFieldNames() noexcept override;
    nvinfer1::IPluginV2* createPlugin(const char* name, const nvinfer1::PluginFieldCollection* fc) noexcept override;
    nvinfer1::IPluginV2* deserializePlugin(
        const char* name, const void* serialData, size_t serialLength) noexcept override;
    void setPluginNamespace(const char* pluginNamespace) noexcept override;
    const char* getPluginNamespace() const noexcept override;

private:
    nvinfer1::PluginFieldCollection mFC{};
    std::vector<nvinfer1::PluginField> mPluginAttributes;
    std::string mNamespace;
};

} // namespace plugin
} // namespace nvinfer1
