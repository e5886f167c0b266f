// pluginCommon.h -- plumbing shared by the two TensorRT plugins of the hot path.
// Counterpart of T/cpp/tensorrt_llm/plugins/common/plugin.h:86-100 (little-endian memcpy read/write serializers) and
// checkMacrosPlugin.h (PLUGIN_ASSERT throws; creators catch and return nullptr).
#pragma once

#ifdef B200_WITH_TENSORRT
#include <NvInferRuntime.h>
#else
#include "trt_stub/NvInferRuntime.h"
#endif

#include <cstring>
#include <stdexcept>
#include <string>

namespace b200
{
namespace plugin
{

constexpr const char* kPluginNamespace = "tensorrt_llm"; // T/tensorrt_llm/plugin/plugin.py:7

struct PluginError : public std::runtime_error
{
    using std::runtime_error::runtime_error;
};

void logError(const char* msg) noexcept; // forwards to the ILogger given to initLibNvInferPlugins (stderr otherwise)

#define B200_PLUGIN_ASSERT(cond)                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
            throw ::b200::plugin::PluginError(std::string("Assertion failed: ") + #cond + " (" + __FILE__ + ":"        \
                + std::to_string(__LINE__) + ")");                                                                     \
    } while (0)

// Serialization helpers: same byte layout as the reference's write()/read() (memcpy of the object representation).
template <typename T>
inline void write(char*& buffer, const T& val)
{
    std::memcpy(buffer, &val, sizeof(T));
    buffer += sizeof(T);
}

template <typename T>
inline void read(const char*& buffer, T& val)
{
    std::memcpy(&val, buffer, sizeof(T));
    buffer += sizeof(T);
}

} // namespace plugin
} // namespace b200
