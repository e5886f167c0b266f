/*
 * NvInferRuntime.h (shim) -- the subset of TensorRT's public plugin API that the two plugins of this repo use, with
 * the same type, enumerator and virtual-method names as TensorRT 9's NvInferRuntime.h / NvInferRuntimeCommon.h /
 * NvInferRuntimePlugin.h (interface names are the public, Apache-2.0-licensed API surface).
 *
 * TensorRT is not installed in this image, so this header exists to (a) compile-check the plugin
 * classes and (b) drive them from tests through include/b200_plugin_harness.h with a tiny in-process registry.
 * A production build against a real TensorRT passes -DB200_WITH_TENSORRT and the TensorRT include directory; the
 * plugin sources include <NvInferRuntime.h> either way and use nothing outside this subset.
 * This is NOT ABI-compatible with libnvinfer: only source-compatible for the plugin sources.
 */
#ifndef B200_NVINFER_RUNTIME_SHIM_H
#define B200_NVINFER_RUNTIME_SHIM_H

#include <cstddef>
#include <cstdint>
#include <cuda_runtime_api.h>

#define NV_TENSORRT_VERSION 9000
#define B200_TRT_SHIM 1

namespace nvinfer1
{

using AsciiChar = char;

enum class DataType : int32_t
{
    kFLOAT = 0,
    kHALF = 1,
    kINT8 = 2,
    kINT32 = 3,
    kBOOL = 4,
    kUINT8 = 5,
    kFP8 = 6,
    kBF16 = 7,
    kINT64 = 8
};

class Dims32
{
public:
    static constexpr int32_t MAX_DIMS{8};
    int32_t nbDims;
    int32_t d[MAX_DIMS];
};

using Dims = Dims32;

enum class TensorFormat : int32_t
{
    kLINEAR = 0,
    kCHW2 = 1,
    kHWC8 = 2,
    kCHW4 = 3,
    kCHW16 = 4,
    kCHW32 = 5
};

using PluginFormat = TensorFormat;

struct PluginTensorDesc
{
    Dims dims;
    DataType type;
    TensorFormat format;
    float scale;
};

struct DynamicPluginTensorDesc
{
    PluginTensorDesc desc;
    Dims min;
    Dims max;
};

enum class DimensionOperation : int32_t
{
    kSUM = 0,
    kPROD = 1,
    kMAX = 2,
    kMIN = 3,
    kSUB = 4,
    kEQUAL = 5,
    kLESS = 6,
    kFLOOR_DIV = 7,
    kCEIL_DIV = 8
};

// In TensorRT these two forward to an opaque implementation; the shim implements them inline for constants,
// which is all the harness needs (runtime shapes are concrete there).
class IDimensionExpr
{
public:
    bool isConstant() const noexcept
    {
        return mConstant;
    }

    int32_t getConstantValue() const noexcept
    {
        return mValue;
    }

    bool mConstant{true};
    int32_t mValue{0};
};

class IExprBuilder
{
public:
    const IDimensionExpr* constant(int32_t value) noexcept;
    const IDimensionExpr* operation(
        DimensionOperation op, const IDimensionExpr& first, const IDimensionExpr& second) noexcept;
    ~IExprBuilder();

private:
    struct Node
    {
        IDimensionExpr e;
        Node* next;
    };

    Node* mHead{nullptr};
};

class DimsExprs
{
public:
    int32_t nbDims;
    const IDimensionExpr* d[Dims::MAX_DIMS];
};

enum class PluginFieldType : int32_t
{
    kFLOAT16 = 0,
    kFLOAT32 = 1,
    kFLOAT64 = 2,
    kINT8 = 3,
    kINT16 = 4,
    kINT32 = 5,
    kCHAR = 6,
    kDIMS = 7,
    kUNKNOWN = 8
};

class PluginField
{
public:
    const AsciiChar* name;
    const void* data;
    PluginFieldType type;
    int32_t length;

    PluginField(const AsciiChar* const name_ = nullptr, const void* const data_ = nullptr,
        const PluginFieldType type_ = PluginFieldType::kUNKNOWN, const int32_t length_ = 0) noexcept
        : name(name_)
        , data(data_)
        , type(type_)
        , length(length_)
    {
    }
};

struct PluginFieldCollection
{
    int32_t nbFields;
    const PluginField* fields;
};

class IGpuAllocator;

class IPluginV2
{
public:
    virtual int32_t getTensorRTVersion() const noexcept
    {
        return NV_TENSORRT_VERSION;
    }

    virtual const AsciiChar* getPluginType() const noexcept = 0;
    virtual const AsciiChar* getPluginVersion() const noexcept = 0;
    virtual int32_t getNbOutputs() const noexcept = 0;
    virtual Dims getOutputDimensions(int32_t index, const Dims* inputs, int32_t nbInputDims) noexcept = 0;
    virtual bool supportsFormat(DataType type, PluginFormat format) const noexcept = 0;
    virtual void configureWithFormat(const Dims* inputDims, int32_t nbInputs, const Dims* outputDims, int32_t nbOutputs,
        DataType type, PluginFormat format, int32_t maxBatchSize) noexcept
        = 0;
    virtual int32_t initialize() noexcept = 0;
    virtual void terminate() noexcept = 0;
    virtual size_t getWorkspaceSize(int32_t maxBatchSize) const noexcept = 0;
    virtual int32_t enqueue(int32_t batchSize, const void* const* inputs, void* const* outputs, void* workspace,
        cudaStream_t stream) noexcept
        = 0;
    virtual size_t getSerializationSize() const noexcept = 0;
    virtual void serialize(void* buffer) const noexcept = 0;
    virtual void destroy() noexcept = 0;
    virtual IPluginV2* clone() const noexcept = 0;
    virtual void setPluginNamespace(const AsciiChar* pluginNamespace) noexcept = 0;
    virtual const AsciiChar* getPluginNamespace() const noexcept = 0;

protected:
    IPluginV2() = default;
    virtual ~IPluginV2() noexcept = default;
};

class IPluginV2Ext : public IPluginV2
{
public:
    virtual DataType getOutputDataType(int32_t index, const DataType* inputTypes, int32_t nbInputs) const noexcept = 0;
    virtual bool isOutputBroadcastAcrossBatch(
        int32_t outputIndex, const bool* inputIsBroadcasted, int32_t nbInputs) const noexcept
        = 0;
    virtual bool canBroadcastInputAcrossBatch(int32_t inputIndex) const noexcept = 0;
    virtual void configurePlugin(const Dims* inputDims, int32_t nbInputs, const Dims* outputDims, int32_t nbOutputs,
        const DataType* inputTypes, const DataType* outputTypes, const bool* inputIsBroadcast,
        const bool* outputIsBroadcast, PluginFormat floatFormat, int32_t maxBatchSize) noexcept
        = 0;

    virtual void attachToContext(void* /*cudnn*/, void* /*cublas*/, IGpuAllocator* /*allocator*/) noexcept {}

    virtual void detachFromContext() noexcept {}

    IPluginV2Ext* clone() const noexcept override = 0;

protected:
    IPluginV2Ext() = default;
    ~IPluginV2Ext() noexcept override = default;

    void configureWithFormat(const Dims*, int32_t, const Dims*, int32_t, DataType, PluginFormat, int32_t) noexcept override
    {
    }
};

class IPluginV2DynamicExt : public IPluginV2Ext
{
public:
    IPluginV2DynamicExt* clone() const noexcept override = 0;
    virtual DimsExprs getOutputDimensions(
        int32_t outputIndex, const DimsExprs* inputs, int32_t nbInputs, IExprBuilder& exprBuilder) noexcept
        = 0;
    static constexpr int32_t kFORMAT_COMBINATION_LIMIT = 100;
    virtual bool supportsFormatCombination(
        int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs, int32_t nbOutputs) noexcept
        = 0;
    virtual void configurePlugin(const DynamicPluginTensorDesc* in, int32_t nbInputs, const DynamicPluginTensorDesc* out,
        int32_t nbOutputs) noexcept
        = 0;
    virtual size_t getWorkspaceSize(const PluginTensorDesc* inputs, int32_t nbInputs, const PluginTensorDesc* outputs,
        int32_t nbOutputs) const noexcept
        = 0;
    virtual int32_t enqueue(const PluginTensorDesc* inputDesc, const PluginTensorDesc* outputDesc,
        const void* const* inputs, void* const* outputs, void* workspace, cudaStream_t stream) noexcept
        = 0;

protected:
    IPluginV2DynamicExt() = default;
    ~IPluginV2DynamicExt() noexcept override = default;

private:
    // the static-shape entry points of the base classes are not used by dynamic plugins
    Dims getOutputDimensions(int32_t, const Dims*, int32_t) noexcept final
    {
        return Dims{-1, {}};
    }

    bool isOutputBroadcastAcrossBatch(int32_t, const bool*, int32_t) const noexcept final
    {
        return false;
    }

    bool canBroadcastInputAcrossBatch(int32_t) const noexcept final
    {
        return true;
    }

    bool supportsFormat(DataType, PluginFormat) const noexcept final
    {
        return false;
    }

    void configurePlugin(const Dims*, int32_t, const Dims*, int32_t, const DataType*, const DataType*, const bool*,
        const bool*, PluginFormat, int32_t) noexcept final
    {
    }

    size_t getWorkspaceSize(int32_t) const noexcept final
    {
        return 0;
    }

    int32_t enqueue(int32_t, const void* const*, void* const*, void*, cudaStream_t) noexcept final
    {
        return 1;
    }
};

class IPluginCreator
{
public:
    virtual int32_t getTensorRTVersion() const noexcept
    {
        return NV_TENSORRT_VERSION;
    }

    virtual const AsciiChar* getPluginName() const noexcept = 0;
    virtual const AsciiChar* getPluginVersion() const noexcept = 0;
    virtual const PluginFieldCollection* getFieldNames() noexcept = 0;
    virtual IPluginV2* createPlugin(const AsciiChar* name, const PluginFieldCollection* fc) noexcept = 0;
    virtual IPluginV2* deserializePlugin(const AsciiChar* name, const void* serialData, size_t serialLength) noexcept = 0;
    virtual void setPluginNamespace(const AsciiChar* pluginNamespace) noexcept = 0;
    virtual const AsciiChar* getPluginNamespace() const noexcept = 0;

    IPluginCreator() = default;
    virtual ~IPluginCreator() = default;
};

class IPluginRegistry
{
public:
    virtual bool registerCreator(IPluginCreator& creator, const AsciiChar* const pluginNamespace) noexcept = 0;
    virtual IPluginCreator* const* getPluginCreatorList(int32_t* const numCreators) const noexcept = 0;
    virtual IPluginCreator* getPluginCreator(const AsciiChar* const pluginName, const AsciiChar* const pluginVersion,
        const AsciiChar* const pluginNamespace = "") noexcept
        = 0;
    virtual bool deregisterCreator(const IPluginCreator& creator) noexcept = 0;

protected:
    virtual ~IPluginRegistry() noexcept = default;
};

class ILogger
{
public:
    enum class Severity : int32_t
    {
        kINTERNAL_ERROR = 0,
        kERROR = 1,
        kWARNING = 2,
        kINFO = 3,
        kVERBOSE = 4
    };
    virtual void log(Severity severity, const AsciiChar* msg) noexcept = 0;
    virtual ~ILogger() = default;
};

} // namespace nvinfer1

extern "C" nvinfer1::IPluginRegistry* getPluginRegistry() noexcept;

#endif // B200_NVINFER_RUNTIME_SHIM_H
