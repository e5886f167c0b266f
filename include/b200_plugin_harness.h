/*
 * b200_plugin_harness.h -- C entry points that drive the TensorRT plugin CLASSES (creator lookup in the registry,
 * createPlugin from a PluginFieldCollection, serialize / deserialize / clone, shape inference, format checks,
 * enqueue) without a TensorRT engine.  TensorRT is not installed in this image, so this is how the tests exercise
 * exactly the code TensorRT would call: plugin registry -> IPluginCreator -> IPluginV2DynamicExt virtuals.
 * It plays the role polygraphy's TrtRunner / tensorrt_llm.runtime.Session play in the reference's tests
 * (T/tests/quantization/test_weight_only_quant_matmul.py:32-82, T/tests/attention/test_gpt_attention.py:225-262).
 */
#ifndef B200_PLUGIN_HARNESS_H
#define B200_PLUGIN_HARNESS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

/* mirrors nvinfer1::PluginField: `type` is nvinfer1::PluginFieldType (0 f16, 1 f32, 2 f64, 3 i8, 4 i16, 5 i32) */
typedef struct b200_plugin_field
{
    const char* name;
    const void* data;
    int32_t type;
    int32_t length;
} b200_plugin_field;

/* mirrors nvinfer1::PluginTensorDesc (dims, DataType, TensorFormat::kLINEAR = 0) */
typedef struct b200_tensor_desc
{
    int32_t nb_dims;
    int32_t d[8];
    int32_t dtype;  /* nvinfer1::DataType: 0 f32, 1 f16, 2 i8, 3 i32 */
    int32_t format; /* nvinfer1::TensorFormat */
} b200_tensor_desc;

/* Looks the creator up in the plugin registry (call initLibNvInferPlugins first, as the reference's Python does). */
void* b200_plugin_get_creator(const char* name, const char* version, const char* plugin_namespace);
/* Writes the creator's field names, '\n'-separated, into buf; returns the number of fields. */
int b200_plugin_creator_field_names(void* creator, char* buf, size_t buf_len);
void* b200_plugin_create(void* creator, const char* layer_name, const b200_plugin_field* fields, int nb_fields);
void* b200_plugin_deserialize(void* creator, const char* layer_name, const void* data, size_t length);
void* b200_plugin_clone(void* plugin);
void b200_plugin_destroy(void* plugin);
const char* b200_plugin_type(void* plugin);
const char* b200_plugin_version(void* plugin);
const char* b200_plugin_namespace(void* plugin);
int b200_plugin_nb_outputs(void* plugin);
size_t b200_plugin_serialization_size(void* plugin);
void b200_plugin_serialize(void* plugin, void* buffer);
/* getOutputDimensions with constant input dims; returns 0 on success */
int b200_plugin_output_dims(void* plugin, int output_index, const b200_tensor_desc* inputs, int nb_inputs,
    b200_tensor_desc* out);
int b200_plugin_output_dtype(void* plugin, int output_index, const int32_t* input_types, int nb_inputs);
int b200_plugin_supports_format(void* plugin, int pos, const b200_tensor_desc* in_out, int nb_inputs, int nb_outputs);
/* configurePlugin(min = max = the given descriptors) */
void b200_plugin_configure(void* plugin, const b200_tensor_desc* inputs, int nb_inputs, const b200_tensor_desc* outputs,
    int nb_outputs);
size_t b200_plugin_workspace_size(void* plugin, const b200_tensor_desc* inputs, int nb_inputs,
    const b200_tensor_desc* outputs, int nb_outputs);
int b200_plugin_enqueue(void* plugin, const b200_tensor_desc* inputs, int nb_inputs, const b200_tensor_desc* outputs,
    int nb_outputs, const void* const* input_ptrs, void* const* output_ptrs, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_PLUGIN_HARNESS_H */
