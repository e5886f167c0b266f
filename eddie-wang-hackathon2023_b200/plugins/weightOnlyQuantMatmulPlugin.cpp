// WeightOnlyQuantMatmulPlugin -- see the header.  Behaviour follows
// T/cpp/tensorrt_llm/plugins/weightOnlyQuantMatmulPlugin/weightOnlyQuantMatmulPlugin.cpp (cited per method).
#include "weightOnlyQuantMatmulPlugin.h"

#include "b200_whisper.h"

#include <cstring>

using namespace nvinfer1;
using nvinfer1::plugin::WeightOnlyQuantMatmulPlugin;
using nvinfer1::plugin::WeightOnlyQuantMatmulPluginCreator;
using b200::plugin::read;
using b200::plugin::write;

namespace
{
constexpr const char* kName = "WeightOnlyQuantMatmul"; // reference .cpp:25-26
constexpr const char* kVersion = "1";
constexpr int kInt8WeightOnly = 1; // weightTypeId: 1 = int8, 2 = int4 (reference .h:47-48)
} // namespace

WeightOnlyQuantMatmulPlugin::WeightOnlyQuantMatmulPlugin(nvinfer1::DataType type, int weightTypeId)
{
    init(type, weightTypeId);
}

// deserialization: DataType || int, exact length (reference .cpp:37-47)
WeightOnlyQuantMatmulPlugin::WeightOnlyQuantMatmulPlugin(const void* data, size_t length)
{
    B200_PLUGIN_ASSERT(data != nullptr && length == sizeof(nvinfer1::DataType) + sizeof(int));
    const char* d = static_cast<const char*>(data);
    nvinfer1::DataType type;
    int weightTypeId = 0;
    read(d, type);
    read(d, weightTypeId);
    init(type, weightTypeId);
}

void WeightOnlyQuantMatmulPlugin::init(nvinfer1::DataType type, int weightTypeId)
{
    mType = type;
    mWeightTypeId = weightTypeId;
    // reference .cpp:49-65 accepts (kHALF, 1) and (kHALF, 2); int4 is outside the B200 hot path
    B200_PLUGIN_ASSERT(mType == nvinfer1::DataType::kHALF && mWeightTypeId == kInt8WeightOnly);
}

IPluginV2DynamicExt* WeightOnlyQuantMatmulPlugin::clone() const noexcept
{
    try
    {
        auto* p = new WeightOnlyQuantMatmulPlugin(mType, mWeightTypeId);
        p->setPluginNamespace(mNamespace.c_str());
        p->mWorkspaceMaxSize = mWorkspaceMaxSize;
        return p;
    }
    catch (const std::exception& e)
    {
        b200::plugin::logError(e.what());
    }
    return nullptr;
}

// out = [m1, ..., N] with N = 4 * weight.d[1] (int8 bytes viewed as float32), reference .cpp:73-110
DimsExprs WeightOnlyQuantMatmulPlugin::getOutputDimensions(
    int outputIndex, const DimsExprs* inputs, int nbInputs, IExprBuilder& exprBuilder) noexcept
{
    try
    {
        B200_PLUGIN_ASSERT(nbInputs == 3);
        B200_PLUGIN_ASSERT(outputIndex == 0);
        const int nbDimsA = inputs[0].nbDims;
        B200_PLUGIN_ASSERT(nbDimsA >= 2);
        B200_PLUGIN_ASSERT(inputs[1].nbDims == 2);
        DimsExprs ret;
        ret.nbDims = nbDimsA;
        for (int i = 0; i < nbDimsA - 1; ++i)
            ret.d[i] = inputs[0].d[i];
        ret.d[nbDimsA - 1] = exprBuilder.constant(inputs[1].d[1]->getConstantValue() * 4);
        return ret;
    }
    catch (const std::exception& e)
    {
        b200::plugin::logError(e.what());
    }
    return DimsExprs{};
}

// activation / scales / output: mType LINEAR; weight: kFLOAT LINEAR (the int8-as-float hack), reference .cpp:112-141
bool WeightOnlyQuantMatmulPlugin::supportsFormatCombination(
    int pos, const PluginTensorDesc* inOut, int /*nbInputs*/, int /*nbOutputs*/) noexcept
{
    if (pos < 0 || pos > 3 || inOut[pos].format != TensorFormat::kLINEAR)
        return false;
    return inOut[pos].type == (pos == 1 ? nvinfer1::DataType::kFLOAT : mType);
}

// workspace sized for the largest profile, reference .cpp:143-160
void WeightOnlyQuantMatmulPlugin::configurePlugin(
    const DynamicPluginTensorDesc* in, int /*nbInputs*/, const DynamicPluginTensorDesc* /*out*/, int /*nbOutputs*/) noexcept
{
    int maxM = 1;
    for (int i = 0; i < in[0].max.nbDims - 1; ++i)
        maxM *= in[0].max.d[i];
    const int maxK = in[0].max.d[in[0].max.nbDims - 1];
    const int maxN = in[1].max.d[1] * 4;
    mWorkspaceMaxSize = b200_woq_workspace_bytes(maxM, maxN, maxK);
}

size_t WeightOnlyQuantMatmulPlugin::getWorkspaceSize(
    const PluginTensorDesc* inputs, int /*nbInputs*/, const PluginTensorDesc* /*outputs*/, int /*nbOutputs*/) const noexcept
{
    if (mWorkspaceMaxSize != 0)
        return mWorkspaceMaxSize;
    int m = 1;
    for (int i = 0; i < inputs[0].dims.nbDims - 1; ++i)
        m *= inputs[0].dims.d[i];
    return b200_woq_workspace_bytes(m, inputs[1].dims.d[1] * 4, inputs[0].dims.d[inputs[0].dims.nbDims - 1]);
}

// m = prod(dims[:-1]), n = 4 * weightDims[1], k = actDims[-1]; m == 1 -> GEMV, else GEMM: the split lives behind
// b200_woq_int8_gemm (reference .cpp:162-222)
int WeightOnlyQuantMatmulPlugin::enqueue(const PluginTensorDesc* inputDesc, const PluginTensorDesc* /*outputDesc*/,
    const void* const* inputs, void* const* outputs, void* workspace, cudaStream_t stream) noexcept
{
    int m = 1;
    for (int i = 0; i < inputDesc[0].dims.nbDims - 1; ++i)
        m *= inputDesc[0].dims.d[i];
    const int n = inputDesc[1].dims.d[1] * 4;
    const int k = inputDesc[0].dims.d[inputDesc[0].dims.nbDims - 1];
    const size_t ws = workspace ? (mWorkspaceMaxSize ? mWorkspaceMaxSize : b200_woq_workspace_bytes(m, n, k)) : 0;
    const int rc = b200_woq_int8_gemm(inputs[0], m, k, static_cast<const int8_t*>(inputs[1]), inputs[2], n, outputs[0],
        workspace, ws, reinterpret_cast<b200_stream_t>(stream));
    if (rc != B200_OK)
        b200::plugin::logError(b200_last_error());
    return rc;
}

nvinfer1::DataType WeightOnlyQuantMatmulPlugin::getOutputDataType(
    int /*index*/, const nvinfer1::DataType* /*inputTypes*/, int /*nbInputs*/) const noexcept
{
    return mType;
}

const char* WeightOnlyQuantMatmulPlugin::getPluginType() const noexcept
{
    return kName;
}

const char* WeightOnlyQuantMatmulPlugin::getPluginVersion() const noexcept
{
    return kVersion;
}

int WeightOnlyQuantMatmulPlugin::getNbOutputs() const noexcept
{
    return 1;
}

int WeightOnlyQuantMatmulPlugin::initialize() noexcept
{
    return 0;
}

void WeightOnlyQuantMatmulPlugin::terminate() noexcept {}

size_t WeightOnlyQuantMatmulPlugin::getSerializationSize() const noexcept
{
    return sizeof(nvinfer1::DataType) + sizeof(int); // reference .cpp:251-254
}

void WeightOnlyQuantMatmulPlugin::serialize(void* buffer) const noexcept
{
    char* d = static_cast<char*>(buffer);
    write(d, mType);
    write(d, mWeightTypeId);
}

void WeightOnlyQuantMatmulPlugin::destroy() noexcept
{
    delete this;
}

void WeightOnlyQuantMatmulPlugin::setPluginNamespace(const char* libNamespace) noexcept
{
    mNamespace = libNamespace ? libNamespace : "";
}

const char* WeightOnlyQuantMatmulPlugin::getPluginNamespace() const noexcept
{
    return mNamespace.c_str();
}

// ---- creator (reference .cpp:281-347) ----

WeightOnlyQuantMatmulPluginCreator::WeightOnlyQuantMatmulPluginCreator()
{
    mPluginAttributes.emplace_back(PluginField("type_id", nullptr, PluginFieldType::kINT32, 1));
    mPluginAttributes.emplace_back(PluginField("weight_type_id", nullptr, PluginFieldType::kINT32, 1));
    mFC.nbFields = static_cast<int32_t>(mPluginAttributes.size());
    mFC.fields = mPluginAttributes.data();
}

const char* WeightOnlyQuantMatmulPluginCreator::getPluginName() const noexcept
{
    return kName;
}

const char* WeightOnlyQuantMatmulPluginCreator::getPluginVersion() const noexcept
{
    return kVersion;
}

const PluginFieldCollection* WeightOnlyQuantMatmulPluginCreator::getFieldNames() noexcept
{
    return &mFC;
}

IPluginV2* WeightOnlyQuantMatmulPluginCreator::createPlugin(const char* /*name*/, const PluginFieldCollection* fc) noexcept
{
    try
    {
        B200_PLUGIN_ASSERT(fc != nullptr);
        nvinfer1::DataType type = nvinfer1::DataType::kHALF;
        int weightTypeId = 0;
        bool haveType = false, haveWeightType = false;
        for (int i = 0; i < fc->nbFields; ++i)
        {
            const PluginField& f = fc->fields[i];
            if (!std::strcmp(f.name, "weight_type_id"))
            {
                B200_PLUGIN_ASSERT(f.type == PluginFieldType::kINT32);
                weightTypeId = *static_cast<const int*>(f.data);
                haveWeightType = true;
            }
            else if (!std::strcmp(f.name, "type_id"))
            {
                B200_PLUGIN_ASSERT(f.type == PluginFieldType::kINT32);
                type = static_cast<nvinfer1::DataType>(*static_cast<const int32_t*>(f.data));
                haveType = true;
            }
        }
        B200_PLUGIN_ASSERT(haveType && haveWeightType);
        auto* obj = new WeightOnlyQuantMatmulPlugin(type, weightTypeId);
        obj->setPluginNamespace(mNamespace.c_str());
        return obj;
    }
    catch (const std::exception& e)
    {
        b200::plugin::logError(e.what());
    }
    return nullptr;
}

IPluginV2* WeightOnlyQuantMatmulPluginCreator::deserializePlugin(
    const char* /*name*/, const void* serialData, size_t serialLength) noexcept
{
    try
    {
        auto* obj = new WeightOnlyQuantMatmulPlugin(serialData, serialLength);
        obj->setPluginNamespace(mNamespace.c_str());
        return obj;
    }
    catch (const std::exception& e)
    {
        b200::plugin::logError(e.what());
    }
    return nullptr;
}

void WeightOnlyQuantMatmulPluginCreator::setPluginNamespace(const char* libNamespace) noexcept
{
    mNamespace = libNamespace ? libNamespace : "";
}

const char* WeightOnlyQuantMatmulPluginCreator::getPluginNamespace() const noexcept
{
    return mNamespace.c_str();
}
