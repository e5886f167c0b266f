// WeightOnlyQuantMatmulPlugin -- TensorRT IPluginV2DynamicExt with the reference's identity and contract
// (T/cpp/tensorrt_llm/plugins/weightOnlyQuantMatmulPlugin/weightOnlyQuantMatmulPlugin.h:41-122):
//   type "WeightOnlyQuantMatmul", version "1", namespace "tensorrt_llm";
//   fields  type_id:int32 (nvinfer1::DataType), weight_type_id:int32 (1 = int8, 2 = int4);
//   inputs  activation [.., K] fp16 | weight [K, N/4] kFLOAT (int8 bytes) | scales [N] fp16;  output [.., N] fp16;
//   serialization  DataType mType (4 B) || int mWeightTypeId (4 B).
// The class only marshals into the C ABI (include/b200_whisper.h); it owns nothing between calls.
#pragma once

#include "pluginCommon.h"

#include <string>
#include <vector>

namespace nvinfer1
{
namespace plugin
{

class WeightOnlyQuantMatmulPlugin : public IPluginV2DynamicExt
{
public:
    WeightOnlyQuantMatmulPlugin() = delete;
    WeightOnlyQuantMatmulPlugin(nvinfer1::DataType type, int weightTypeId);
    WeightOnlyQuantMatmulPlugin(const void* data, size_t length);
    ~WeightOnlyQuantMatmulPlugin() override = default;

    // IPluginV2DynamicExt
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

    // IPluginV2Ext
    nvinfer1::DataType getOutputDataType(
        int index, const nvinfer1::DataType* inputTypes, int nbInputs) const noexcept override;

    // IPluginV2
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

private:
    void init(nvinfer1::DataType type, int weightTypeId);

    std::string mNamespace;
    size_t mWorkspaceMaxSize{0};
    nvinfer1::DataType mType;
    int mWeightTypeId;
};

class WeightOnlyQuantMatmulPluginCreator : public IPluginCreator
{
public:
    WeightOnlyQuantMatmulPluginCreator();
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
