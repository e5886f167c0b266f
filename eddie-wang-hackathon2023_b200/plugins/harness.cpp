// harness.cpp -- implementation of include/b200_plugin_harness.h (test driver for the plugin classes).
#include "b200_plugin_harness.h"

#include "pluginCommon.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace nvinfer1;

namespace
{
PluginTensorDesc toDesc(const b200_tensor_desc& t)
{
    PluginTensorDesc d{};
    d.dims.nbDims = t.nb_dims;
    for (int i = 0; i < Dims::MAX_DIMS; ++i)
        d.dims.d[i] = i < t.nb_dims ? t.d[i] : 0;
    d.type = static_cast<DataType>(t.dtype);
    d.format = static_cast<TensorFormat>(t.format);
    d.scale = 1.f;
    return d;
}

std::vector<PluginTensorDesc> toDescs(const b200_tensor_desc* t, int n)
{
    std::vector<PluginTensorDesc> v;
    for (int i = 0; i < n; ++i)
        v.push_back(toDesc(t[i]));
    return v;
}

IPluginV2DynamicExt* P(void* p)
{
    return static_cast<IPluginV2DynamicExt*>(p);
}
} // namespace

extern "C"
{

void* b200_plugin_get_creator(const char* name, const char* version, const char* plugin_namespace)
{
    return getPluginRegistry()->getPluginCreator(name, version, plugin_namespace);
}

int b200_plugin_creator_field_names(void* creator, char* buf, size_t buf_len)
{
    const PluginFieldCollection* fc = static_cast<IPluginCreator*>(creator)->getFieldNames();
    std::string s;
    for (int i = 0; i < fc->nbFields; ++i)
    {
        s += fc->fields[i].name;
        s += '\n';
    }
    if (buf != nullptr && buf_len > 0)
    {
        std::strncpy(buf, s.c_str(), buf_len - 1);
        buf[buf_len - 1] = 0;
    }
    return fc->nbFields;
}

void* b200_plugin_create(void* creator, const char* layer_name, const b200_plugin_field* fields, int nb_fields)
{
    std::vector<PluginField> pf;
    for (int i = 0; i < nb_fields; ++i)
        pf.emplace_back(fields[i].name, fields[i].data, static_cast<PluginFieldType>(fields[i].type), fields[i].length);
    PluginFieldCollection fc{nb_fields, pf.data()};
    // createPlugin returns IPluginV2*; both plugins of this library are IPluginV2DynamicExt
    return dynamic_cast<IPluginV2DynamicExt*>(static_cast<IPluginCreator*>(creator)->createPlugin(layer_name, &fc));
}

void* b200_plugin_deserialize(void* creator, const char* layer_name, const void* data, size_t length)
{
    return dynamic_cast<IPluginV2DynamicExt*>(
        static_cast<IPluginCreator*>(creator)->deserializePlugin(layer_name, data, length));
}

void* b200_plugin_clone(void* plugin)
{
    return P(plugin)->clone();
}

void b200_plugin_destroy(void* plugin)
{
    P(plugin)->terminate();
    P(plugin)->destroy();
}

const char* b200_plugin_type(void* plugin)
{
    return P(plugin)->getPluginType();
}

const char* b200_plugin_version(void* plugin)
{
    return P(plugin)->getPluginVersion();
}

const char* b200_plugin_namespace(void* plugin)
{
    return P(plugin)->getPluginNamespace();
}

int b200_plugin_nb_outputs(void* plugin)
{
    return P(plugin)->getNbOutputs();
}

size_t b200_plugin_serialization_size(void* plugin)
{
    return P(plugin)->getSerializationSize();
}

void b200_plugin_serialize(void* plugin, void* buffer)
{
    P(plugin)->serialize(buffer);
}

int b200_plugin_output_dims(
    void* plugin, int output_index, const b200_tensor_desc* inputs, int nb_inputs, b200_tensor_desc* out)
{
    IExprBuilder builder;
    std::vector<DimsExprs> in(nb_inputs);
    for (int i = 0; i < nb_inputs; ++i)
    {
        in[i].nbDims = inputs[i].nb_dims;
        for (int j = 0; j < inputs[i].nb_dims; ++j)
            in[i].d[j] = builder.constant(inputs[i].d[j]);
    }
    const DimsExprs r = P(plugin)->getOutputDimensions(output_index, in.data(), nb_inputs, builder);
    if (r.nbDims <= 0)
        return 1;
    out->nb_dims = r.nbDims;
    for (int j = 0; j < r.nbDims; ++j)
        out->d[j] = r.d[j]->getConstantValue();
    return 0;
}

int b200_plugin_output_dtype(void* plugin, int output_index, const int32_t* input_types, int nb_inputs)
{
    std::vector<DataType> t;
    for (int i = 0; i < nb_inputs; ++i)
        t.push_back(static_cast<DataType>(input_types[i]));
    return static_cast<int>(P(plugin)->getOutputDataType(output_index, t.data(), nb_inputs));
}

int b200_plugin_supports_format(void* plugin, int pos, const b200_tensor_desc* in_out, int nb_inputs, int nb_outputs)
{
    auto d = toDescs(in_out, nb_inputs + nb_outputs);
    return P(plugin)->supportsFormatCombination(pos, d.data(), nb_inputs, nb_outputs) ? 1 : 0;
}

void b200_plugin_configure(
    void* plugin, const b200_tensor_desc* inputs, int nb_inputs, const b200_tensor_desc* outputs, int nb_outputs)
{
    std::vector<DynamicPluginTensorDesc> in, out;
    for (int i = 0; i < nb_inputs; ++i)
    {
        const PluginTensorDesc d = toDesc(inputs[i]);
        in.push_back(DynamicPluginTensorDesc{d, d.dims, d.dims});
    }
    for (int i = 0; i < nb_outputs; ++i)
    {
        const PluginTensorDesc d = toDesc(outputs[i]);
        out.push_back(DynamicPluginTensorDesc{d, d.dims, d.dims});
    }
    P(plugin)->configurePlugin(in.data(), nb_inputs, out.data(), nb_outputs);
    P(plugin)->initialize();
}

size_t b200_plugin_workspace_size(
    void* plugin, const b200_tensor_desc* inputs, int nb_inputs, const b200_tensor_desc* outputs, int nb_outputs)
{
    auto in = toDescs(inputs, nb_inputs);
    auto out = toDescs(outputs, nb_outputs);
    return P(plugin)->getWorkspaceSize(in.data(), nb_inputs, out.data(), nb_outputs);
}

int b200_plugin_enqueue(void* plugin, const b200_tensor_desc* inputs, int nb_inputs, const b200_tensor_desc* outputs,
    int nb_outputs, const void* const* input_ptrs, void* const* output_ptrs, void* workspace, void* stream)
{
    auto in = toDescs(inputs, nb_inputs);
    auto out = toDescs(outputs, nb_outputs);
    return P(plugin)->enqueue(in.data(), out.data(), input_ptrs, output_ptrs, workspace, static_cast<cudaStream_t>(stream));
}
}
