// GPTAttentionPlugin -- see the header.  Behaviour follows gptAttentionPlugin.cpp / gptAttentionCommon.cpp of the
// reference (T/cpp/tensorrt_llm/plugins/...), cited per method.
#include "gptAttentionPlugin.h"

#include "b200_whisper.h"

#include <cstring>
#include <optional>

using namespace nvinfer1;
using nvinfer1::plugin::GPTAttentionPlugin;
using nvinfer1::plugin::GPTAttentionPluginCreator;
using b200::plugin::read;
using b200::plugin::write;

namespace
{
constexpr const char* kName = "GPTAttention"; // gptAttentionPlugin.cpp:35-36
constexpr const char* kVersion = "1";
} // namespace

GPTAttentionPlugin::GPTAttentionPlugin(int num_heads, int head_size, int unidirectional, float q_scaling,
    int rotary_embedding_dim, bool neox_rotary_style, int context_fmha_type, bool multi_block_mode, bool multi_query_mode,
    bool int8_kv_cache, bool fp8_kv_cache, bool remove_input_padding, int mask_type, bool paged_kv_cache,
    nvinfer1::DataType type, bool in_flight_batching)
    : mNumHeads(num_heads)
    , mHeadSize(head_size)
    , mUnidirectional(unidirectional)
    , mQScaling(q_scaling)
    , mRotaryEmbeddingDim(rotary_embedding_dim)
    , mNeoxRotaryStyle(neox_rotary_style)
    , mEnableContextFMHA(context_fmha_type != 0)  // ContextFMHAType::disabled == 0
    , mFMHAForceFP32Acc(context_fmha_type == 2)   // ContextFMHAType::enabled_with_fp32_acc == 2
    , mMultiBlockMode(multi_block_mode)
    , mMultiQueryMode(multi_query_mode)
    , mInt8KVCache(int8_kv_cache)
    , mFp8KVCache(fp8_kv_cache)
    , mRemovePadding(remove_input_padding)
    , mMaskType(mask_type)
    , mPagedKVCache(paged_kv_cache)
    , mType(type)
    , mInFlightBatching(in_flight_batching)
{
    B200_PLUGIN_ASSERT(mNumHeads > 0 && mHeadSize > 0);
    B200_PLUGIN_ASSERT(!mInFlightBatching || mRemovePadding); // gptAttentionPlugin.cpp:49
}

// common block then inFlightBatching, exact length (gptAttentionCommon.cpp:304-330, gptAttentionPlugin.cpp:52-61)
GPTAttentionPlugin::GPTAttentionPlugin(const void* data, size_t length)
{
    B200_PLUGIN_ASSERT(data != nullptr && length == kSerializedSize);
    const char* d = static_cast<const char*>(data);
    read(d, mNumHeads);
    read(d, mHeadSize);
    read(d, mUnidirectional);
    read(d, mQScaling);
    read(d, mRotaryEmbeddingDim);
    read(d, mNeoxRotaryStyle);
    read(d, mEnableContextFMHA);
    read(d, mFMHAForceFP32Acc);
    read(d, mMultiBlockMode);
    read(d, mMultiQueryMode);
    read(d, mInt8KVCache);
    read(d, mFp8KVCache);
    read(d, mRemovePadding);
    read(d, mMaskType);
    read(d, mPagedKVCache);
    read(d, mType);
    read(d, mInFlightBatching);
    B200_PLUGIN_ASSERT(!mInFlightBatching || mRemovePadding);
}

IPluginV2DynamicExt* GPTAttentionPlugin::clone() const noexcept
{
    try
    {
        char buf[kSerializedSize];
        serialize(buf);
        auto* p = new GPTAttentionPlugin(buf, kSerializedSize);
        p->setPluginNamespace(mNamespace.c_str());
        return p;
    }
    catch (const std::exception& e)
    {
        b200::plugin::logError(e.what());
    }
    return nullptr;
}

// output 0: qkv dims with last = H * Dh; output 1: same dims as input 1 (gptAttentionPlugin.cpp:78-92)
DimsExprs GPTAttentionPlugin::getOutputDimensions(
    int outputIndex, const DimsExprs* inputs, int /*nbInputs*/, IExprBuilder& exprBuilder) noexcept
{
    if (outputIndex == 0)
    {
        DimsExprs ret = inputs[kQKV];
        ret.d[2] = exprBuilder.constant(mNumHeads * mHeadSize);
        return ret;
    }
    return inputs[outputIndex == 1 ? kPAST_KV : 0];
}

// int32 for the length/mask/indirection tensors, kFLOAT for the KV scales, kINT8 for the cache when int8 KV is on,
// mType otherwise; all LINEAR (gptAttentionPlugin.cpp:94-129)
bool GPTAttentionPlugin::supportsFormatCombination(
    int pos, const PluginTensorDesc* inOut, int nbInputs, int /*nbOutputs*/) noexcept
{
    if (inOut[pos].format != TensorFormat::kLINEAR)
        return false;
    if (pos == kSEQUENCE_LENGTH || pos == kPAST_KV_LENGTH || pos == kMASKED_TOKENS || pos == kINPUT_LENGTHS
        || pos == kMAX_INPUT_LENGTH || pos == kCACHE_INDIR)
        return inOut[pos].type == nvinfer1::DataType::kINT32;
    if ((mInt8KVCache || mFp8KVCache) && (pos == kKV_QUANT_SCALE || pos == kKV_DEQUANT_SCALE))
        return inOut[pos].type == nvinfer1::DataType::kFLOAT;
    if (mPagedKVCache && pos == blockPointersIdx())
        return inOut[pos].type == nvinfer1::DataType::kINT32; // pointers as pairs of int32 (gptAttentionPlugin.cpp:106-110)
    if (mInt8KVCache && (pos == kPAST_KV || pos == nbInputs + 1))
        return inOut[pos].type == nvinfer1::DataType::kINT8;
    return inOut[pos].type == mType;
}

void GPTAttentionPlugin::configurePlugin(const DynamicPluginTensorDesc*, int, const DynamicPluginTensorDesc*, int) noexcept {}

// The reference carves nine scratch buffers for the context phase (gptAttentionCommon.cpp:452-464); the single-kernel
// B200 context phase and the generation kernel need none.
size_t GPTAttentionPlugin::getWorkspaceSize(const PluginTensorDesc*, int, const PluginTensorDesc*, int) const noexcept
{
    return 0;
}

const char* GPTAttentionPlugin::unsupportedReason() const
{
    if (mType != nvinfer1::DataType::kHALF)
        return "GPTAttention on B200: only float16 activations are implemented";
    if (mHeadSize != 64)
        return "GPTAttention on B200: only head_size 64 is implemented";
    if (mRotaryEmbeddingDim != 0)
        return "GPTAttention on B200: rotary embeddings are not on the Whisper hot path";
    if (mMultiQueryMode || mFp8KVCache || mInFlightBatching || mRemovePadding)
        return "GPTAttention on B200: multi-query / fp8 KV / in-flight batching / packed input are not on the "
               "Whisper hot path";
    if (mMaskType != 1 || !mUnidirectional)
        return "GPTAttention on B200: only the causal mask is implemented";
    return nullptr;
}

// gptAttentionPlugin.cpp:203-379 (enqueueSome): host scalars [past_len, is_context] pick the phase.
int GPTAttentionPlugin::enqueue(const PluginTensorDesc* inputDesc, const PluginTensorDesc* outputDesc,
    const void* const* inputs, void* const* outputs, void* /*workspace*/, cudaStream_t stream) noexcept
{
    if (const char* why = unsupportedReason())
    {
        b200::plugin::logError(why);
        return B200_ERR_UNSUPPORTED;
    }
    const int* host_scalars = static_cast<const int*>(inputs[kPAST_KV_LENGTH]); // HOST memory (:261-262)
    const int past_kv_len = host_scalars[0];
    const bool is_context = host_scalars[1] != 0;
    const int nbSeq = inputDesc[kINPUT_LENGTHS].dims.d[0];
    const int max_input_len = inputDesc[kMAX_INPUT_LENGTH].dims.d[0]; // the shape carries the value (:284)
    const int max_seq_len = inputDesc[kCACHE_INDIR].dims.d[2];         // (:335)
    const int beam_width = inputDesc[kCACHE_INDIR].dims.d[1];
    const float* kv_scale_orig_quant = nullptr;
    const float* kv_scale_quant_orig = nullptr;
    if (mInt8KVCache)
    {
        kv_scale_orig_quant = static_cast<const float*>(inputs[kKV_QUANT_SCALE]);
        kv_scale_quant_orig = static_cast<const float*>(inputs[kKV_DEQUANT_SCALE]);
    }
    void* key_value_cache = outputs[1]; // present == past buffer by convention (test_gpt_attention.py:245-248)
    (void) outputDesc;
    // paged KV cache (gptAttentionPlugin.cpp:314-326): input 1 is the block pool [blocks, 2, H, tokens_per_block, Dh], the
    // block-pointer input [B, beam, 2, 2 * max_blocks_per_seq] holds 64-bit device addresses as pairs of int32
    const void* const* block_pointers = nullptr;
    int max_blocks_per_seq = 0, tokens_per_block = 0;
    if (mPagedKVCache)
    {
        max_blocks_per_seq = inputDesc[blockPointersIdx()].dims.d[3] / 2;
        tokens_per_block = inputDesc[kPAST_KV].dims.d[3];
        block_pointers = static_cast<const void* const*>(inputs[blockPointersIdx()]);
        if (beam_width != 1)
        {
            b200::plugin::logError("GPTAttention on B200: beam search is not on the Whisper hot path (beam width 1)");
            return B200_ERR_UNSUPPORTED;
        }
    }
    int rc;
    if (is_context && mPagedKVCache)
    {
        rc = b200_attention_context_paged(inputs[kQKV], static_cast<const int32_t*>(inputs[kINPUT_LENGTHS]), outputs[0],
            block_pointers, max_blocks_per_seq, tokens_per_block, kv_scale_orig_quant, nbSeq, max_input_len, mNumHeads,
            mHeadSize, max_seq_len, mInt8KVCache ? 1 : 0, mQScaling, reinterpret_cast<b200_stream_t>(stream));
    }
    else if (is_context)
    {
        rc = b200_attention_context(inputs[kQKV], static_cast<const int32_t*>(inputs[kINPUT_LENGTHS]), outputs[0],
            key_value_cache, kv_scale_orig_quant, nbSeq, max_input_len, mNumHeads, mHeadSize, max_seq_len,
            mInt8KVCache ? 1 : 0, mQScaling, reinterpret_cast<b200_stream_t>(stream));
    }
    else
    {
        if (beam_width != 1)
        {
            b200::plugin::logError("GPTAttention on B200: beam search is not on the Whisper hot path (beam width 1)");
            return B200_ERR_UNSUPPORTED;
        }
        b200_mmha_params p{};
        p.qkv = inputs[kQKV];
        p.qkv_bias = nullptr; // gptAttentionCommon.cpp:723
        p.out = outputs[0];
        p.kv_cache = key_value_cache;
        p.sequence_lengths = static_cast<const int32_t*>(inputs[kSEQUENCE_LENGTH]);
        p.masked_tokens = static_cast<const int32_t*>(inputs[kMASKED_TOKENS]);
        p.kv_scale_orig_quant = kv_scale_orig_quant;
        p.kv_scale_quant_orig = kv_scale_quant_orig;
        p.batch_size = nbSeq;
        p.num_heads = mNumHeads;
        p.head_size = mHeadSize;
        p.max_seq_len = max_seq_len;
        p.past_kv_length = past_kv_len;
        p.int8_kv_cache = mInt8KVCache ? 1 : 0;
        p.q_scaling = mQScaling;
        rc = mPagedKVCache ? b200_mmha_generation_paged(&p, block_pointers, max_blocks_per_seq, tokens_per_block,
                                 reinterpret_cast<b200_stream_t>(stream))
                           : b200_mmha_generation(&p, reinterpret_cast<b200_stream_t>(stream));
    }
    if (rc != B200_OK)
        b200::plugin::logError(b200_last_error());
    return rc;
}

nvinfer1::DataType GPTAttentionPlugin::getOutputDataType(
    int index, const nvinfer1::DataType* inputTypes, int /*nbInputs*/) const noexcept
{
    return inputTypes[index]; // gptAttentionPlugin.cpp:419-424
}

const char* GPTAttentionPlugin::getPluginType() const noexcept
{
    return kName;
}

const char* GPTAttentionPlugin::getPluginVersion() const noexcept
{
    return kVersion;
}

int GPTAttentionPlugin::getNbOutputs() const noexcept
{
    return 2;
}

int GPTAttentionPlugin::initialize() noexcept
{
    return 0; // no cuBLAS handles to create: the context phase is one kernel here (reference: gptAttentionCommon.cpp:810-842)
}

void GPTAttentionPlugin::terminate() noexcept {}

size_t GPTAttentionPlugin::getSerializationSize() const noexcept
{
    return kSerializedSize;
}

void GPTAttentionPlugin::serialize(void* buffer) const noexcept
{
    char* d = static_cast<char*>(buffer);
    write(d, mNumHeads);
    write(d, mHeadSize);
    write(d, mUnidirectional);
    write(d, mQScaling);
    write(d, mRotaryEmbeddingDim);
    write(d, mNeoxRotaryStyle);
    write(d, mEnableContextFMHA);
    write(d, mFMHAForceFP32Acc);
    write(d, mMultiBlockMode);
    write(d, mMultiQueryMode);
    write(d, mInt8KVCache);
    write(d, mFp8KVCache);
    write(d, mRemovePadding);
    write(d, mMaskType);
    write(d, mPagedKVCache);
    write(d, mType);
    write(d, mInFlightBatching);
}

void GPTAttentionPlugin::destroy() noexcept
{
    delete this;
}

void GPTAttentionPlugin::setPluginNamespace(const char* libNamespace) noexcept
{
    mNamespace = libNamespace ? libNamespace : "";
}

const char* GPTAttentionPlugin::getPluginNamespace() const noexcept
{
    return mNamespace.c_str();
}

// ---- creator (gptAttentionCommon.cpp:918-950, gptAttentionPlugin.cpp:459-533) ----

GPTAttentionPluginCreator::GPTAttentionPluginCreator()
{
    const struct
    {
        const char* name;
        PluginFieldType type;
    } fields[] = {{"num_heads", PluginFieldType::kINT32}, {"head_size", PluginFieldType::kINT32},
        {"unidirectional", PluginFieldType::kINT32}, {"q_scaling", PluginFieldType::kFLOAT32},
        {"rotary_embedding_dim", PluginFieldType::kINT32}, {"neox_rotary_style", PluginFieldType::kINT8},
        {"context_fmha_type", PluginFieldType::kINT8}, {"multi_block_mode", PluginFieldType::kINT8},
        {"multi_query_mode", PluginFieldType::kINT8}, {"int8_kv_cache", PluginFieldType::kINT32},
        {"fp8_kv_cache", PluginFieldType::kINT32}, {"remove_input_padding", PluginFieldType::kINT8},
        {"mask_type", PluginFieldType::kINT32}, {"paged_kv_cache", PluginFieldType::kINT32},
        {"type_id", PluginFieldType::kINT32}, {"in_flight_batching", PluginFieldType::kINT32}};
    for (const auto& f : fields)
        mPluginAttributes.emplace_back(PluginField(f.name, nullptr, f.type, 1));
    mFC.nbFields = static_cast<int32_t>(mPluginAttributes.size());
    mFC.fields = mPluginAttributes.data();
}

const char* GPTAttentionPluginCreator::getPluginName() const noexcept
{
    return kName;
}

const char* GPTAttentionPluginCreator::getPluginVersion() const noexcept
{
    return kVersion;
}

const PluginFieldCollection* GPTAttentionPluginCreator::getFieldNames() noexcept
{
    return &mFC;
}

namespace
{
// typed lookup of one scalar field; throws when the field is missing or has the wrong type
// (the reference's PluginFieldParser::getScalar<T>().value(), T/cpp/tensorrt_llm/plugins/common/plugin.cpp)
template <typename T>
T scalarField(const PluginFieldCollection* fc, const char* name, PluginFieldType type)
{
    for (int i = 0; i < fc->nbFields; ++i)
    {
        const PluginField& f = fc->fields[i];
        if (f.name != nullptr && !std::strcmp(f.name, name))
        {
            B200_PLUGIN_ASSERT(f.type == type && f.data != nullptr);
            T v;
            std::memcpy(&v, f.data, sizeof(T));
            return v;
        }
    }
    throw b200::plugin::PluginError(std::string("GPTAttention: missing plugin field ") + name);
}
} // namespace

IPluginV2* GPTAttentionPluginCreator::createPlugin(const char* /*name*/, const PluginFieldCollection* fc) noexcept
{
    try
    {
        B200_PLUGIN_ASSERT(fc != nullptr);
        auto* obj = new GPTAttentionPlugin(scalarField<int32_t>(fc, "num_heads", PluginFieldType::kINT32),
            scalarField<int32_t>(fc, "head_size", PluginFieldType::kINT32),
            scalarField<int32_t>(fc, "unidirectional", PluginFieldType::kINT32),
            scalarField<float>(fc, "q_scaling", PluginFieldType::kFLOAT32),
            scalarField<int32_t>(fc, "rotary_embedding_dim", PluginFieldType::kINT32),
            scalarField<int8_t>(fc, "neox_rotary_style", PluginFieldType::kINT8) != 0,
            scalarField<int8_t>(fc, "context_fmha_type", PluginFieldType::kINT8),
            scalarField<int8_t>(fc, "multi_block_mode", PluginFieldType::kINT8) != 0,
            scalarField<int8_t>(fc, "multi_query_mode", PluginFieldType::kINT8) != 0,
            scalarField<int32_t>(fc, "int8_kv_cache", PluginFieldType::kINT32) != 0,
            scalarField<int32_t>(fc, "fp8_kv_cache", PluginFieldType::kINT32) != 0,
            scalarField<int8_t>(fc, "remove_input_padding", PluginFieldType::kINT8) != 0,
            scalarField<int32_t>(fc, "mask_type", PluginFieldType::kINT32),
            scalarField<int32_t>(fc, "paged_kv_cache", PluginFieldType::kINT32) != 0,
            static_cast<nvinfer1::DataType>(scalarField<int32_t>(fc, "type_id", PluginFieldType::kINT32)),
            scalarField<int32_t>(fc, "in_flight_batching", PluginFieldType::kINT32) != 0);
        obj->setPluginNamespace(mNamespace.c_str());
        return obj;
    }
    catch (const std::exception& e)
    {
        b200::plugin::logError(e.what());
    }
    return nullptr;
}

IPluginV2* GPTAttentionPluginCreator::deserializePlugin(const char* /*name*/, const void* serialData, size_t serialLength) noexcept
{
    try
    {
        auto* obj = new GPTAttentionPlugin(serialData, serialLength);
        obj->setPluginNamespace(mNamespace.c_str());
        return obj;
    }
    catch (const std::exception& e)
    {
        b200::plugin::logError(e.what());
    }
    return nullptr;
}

void GPTAttentionPluginCreator::setPluginNamespace(const char* libNamespace) noexcept
{
    mNamespace = libNamespace ? libNamespace : "";
}

const char* GPTAttentionPluginCreator::getPluginNamespace() const noexcept
{
    return mNamespace.c_str();
}
