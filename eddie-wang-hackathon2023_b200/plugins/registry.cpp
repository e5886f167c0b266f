// registry.cpp -- library entry point of the plugin layer.
//
//   extern "C" bool initLibNvInferPlugins(void* logger, const char* libNamespace)
// is the one symbol the reference's Python loads the plugin library for (T/tensorrt_llm/plugin/plugin.py:10-22,
// T/cpp/tensorrt_llm/plugins/api/InferPlugin.cpp:147-170, export list T/cpp/tensorrt_llm/plugins/exports.map:19-32):
// it registers the creators with TensorRT's global plugin registry under the given namespace; idempotent and
// mutex-protected like the reference's PluginCreatorRegistry (InferPlugin.cpp:52-140).
//
// Without TensorRT (this image) the registry is the small in-process one below, reached through the same
// getPluginRegistry() call the reference's Python side makes through trt.get_plugin_registry().
#include "gptAttentionPlugin.h"
#include "pluginCommon.h"
#include "weightOnlyQuantMatmulPlugin.h"

#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace
{
std::mutex g_mutex;
nvinfer1::ILogger* g_logger = nullptr;
std::vector<std::unique_ptr<nvinfer1::IPluginCreator>> g_creators; // owned for the life of the library
} // namespace

namespace b200
{
namespace plugin
{
void logError(const char* msg) noexcept
{
    nvinfer1::ILogger* lg = g_logger;
    if (lg != nullptr)
        lg->log(nvinfer1::ILogger::Severity::kERROR, msg);
    else
        std::fprintf(stderr, "[b200 plugin] %s\n", msg);
}
} // namespace plugin
} // namespace b200

#ifndef B200_WITH_TENSORRT
// ---- shim-only pieces: expression builder for constants and an in-process registry ------------------------
namespace nvinfer1
{

const IDimensionExpr* IExprBuilder::constant(int32_t value) noexcept
{
    Node* n = new Node{IDimensionExpr{}, mHead};
    n->e.mConstant = true;
    n->e.mValue = value;
    mHead = n;
    return &n->e;
}

const IDimensionExpr* IExprBuilder::operation(
    DimensionOperation op, const IDimensionExpr& a, const IDimensionExpr& b) noexcept
{
    const int32_t x = a.getConstantValue(), y = b.getConstantValue();
    int32_t r = 0;
    switch (op)
    {
    case DimensionOperation::kSUM: r = x + y; break;
    case DimensionOperation::kPROD: r = x * y; break;
    case DimensionOperation::kMAX: r = x > y ? x : y; break;
    case DimensionOperation::kMIN: r = x < y ? x : y; break;
    case DimensionOperation::kSUB: r = x - y; break;
    case DimensionOperation::kEQUAL: r = x == y; break;
    case DimensionOperation::kLESS: r = x < y; break;
    case DimensionOperation::kFLOOR_DIV: r = y ? x / y : 0; break;
    case DimensionOperation::kCEIL_DIV: r = y ? (x + y - 1) / y : 0; break;
    }
    return constant(r);
}

IExprBuilder::~IExprBuilder()
{
    while (mHead != nullptr)
    {
        Node* n = mHead->next;
        delete mHead;
        mHead = n;
    }
}

} // namespace nvinfer1

namespace
{
class ShimRegistry : public nvinfer1::IPluginRegistry
{
public:
    bool registerCreator(nvinfer1::IPluginCreator& creator, const char* const pluginNamespace) noexcept override
    {
        std::lock_guard<std::mutex> lk(mMu);
        const std::string ns = pluginNamespace ? pluginNamespace : "";
        for (auto* c : mList)
            if (!std::strcmp(c->getPluginName(), creator.getPluginName())
                && !std::strcmp(c->getPluginVersion(), creator.getPluginVersion()) && ns == c->getPluginNamespace())
                return false; // duplicate (TensorRT logs an error and returns false)
        creator.setPluginNamespace(ns.c_str());
        mList.push_back(&creator);
        return true;
    }

    nvinfer1::IPluginCreator* const* getPluginCreatorList(int32_t* const numCreators) const noexcept override
    {
        *numCreators = static_cast<int32_t>(mList.size());
        return mList.data();
    }

    nvinfer1::IPluginCreator* getPluginCreator(
        const char* const name, const char* const version, const char* const ns) noexcept override
    {
        std::lock_guard<std::mutex> lk(mMu);
        for (auto* c : mList)
            if (!std::strcmp(c->getPluginName(), name) && !std::strcmp(c->getPluginVersion(), version)
                && !std::strcmp(c->getPluginNamespace(), ns ? ns : ""))
                return c;
        return nullptr;
    }

    bool deregisterCreator(const nvinfer1::IPluginCreator& creator) noexcept override
    {
        std::lock_guard<std::mutex> lk(mMu);
        for (size_t i = 0; i < mList.size(); ++i)
            if (mList[i] == &creator)
            {
                mList.erase(mList.begin() + static_cast<long>(i));
                return true;
            }
        return false;
    }

private:
    mutable std::mutex mMu;
    std::vector<nvinfer1::IPluginCreator*> mList;
};
} // namespace

extern "C" nvinfer1::IPluginRegistry* getPluginRegistry() noexcept
{
    static ShimRegistry registry;
    return &registry;
}
#endif // !B200_WITH_TENSORRT

namespace
{
template <typename Creator>
void addCreator(const char* libNamespace)
{
    auto creator = std::make_unique<Creator>();
    creator->setPluginNamespace(libNamespace);
    if (getPluginRegistry()->registerCreator(*creator, libNamespace))
        g_creators.push_back(std::move(creator));
    // already registered (second call): drop the new instance -- initLibNvInferPlugins is idempotent
}
} // namespace

extern "C" bool initLibNvInferPlugins(void* logger, const char* libNamespace)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (logger != nullptr)
        g_logger = static_cast<nvinfer1::ILogger*>(logger);
    const char* ns = libNamespace ? libNamespace : "";
    addCreator<nvinfer1::plugin::GPTAttentionPluginCreator>(ns);
    addCreator<nvinfer1::plugin::WeightOnlyQuantMatmulPluginCreator>(ns);
    return true;
}
