"""Plugin library loading and a Python handle on the TensorRT plugin classes.

Mirror of T/tensorrt_llm/plugin/plugin.py:7-22 (`_load_plugin_lib`: CDLL(..., RTLD_GLOBAL) then
`initLibNvInferPlugins(None, b"tensorrt_llm")`) and of the PluginConfig attribute bag (:33-140) for the flags that
concern this path.  In the reference the creators are then fetched with
trt.get_plugin_registry().get_plugin_creator(name, '1', 'tensorrt_llm'); TensorRT is absent here, so TrtPlugin reaches
the same registry / creator / plugin virtuals through include/b200_plugin_harness.h.
"""
import ctypes

import numpy as np

from .. import _lib

TRT_LLM_PLUGIN_NAMESPACE = 'tensorrt_llm'

# nvinfer1::PluginFieldType / nvinfer1::DataType codes
_FIELD_TYPE = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int8): 3,
               np.dtype(np.int16): 4, np.dtype(np.int32): 5}
TRT_DTYPE = {"float32": 0, "float16": 1, "int8": 2, "int32": 3}


class _Field(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("data", ctypes.c_void_p), ("type", ctypes.c_int32), ("length", ctypes.c_int32)]


class _Desc(ctypes.Structure):
    _fields_ = [("nb_dims", ctypes.c_int32), ("d", ctypes.c_int32 * 8), ("dtype", ctypes.c_int32), ("format", ctypes.c_int32)]


_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return
    vp, i, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    sigs = {
        "initLibNvInferPlugins": (ctypes.c_bool, [vp, ctypes.c_char_p]),
        "b200_plugin_get_creator": (vp, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]),
        "b200_plugin_creator_field_names": (i, [vp, ctypes.c_char_p, sz]),
        "b200_plugin_create": (vp, [vp, ctypes.c_char_p, ctypes.POINTER(_Field), i]),
        "b200_plugin_deserialize": (vp, [vp, ctypes.c_char_p, vp, sz]),
        "b200_plugin_clone": (vp, [vp]),
        "b200_plugin_destroy": (None, [vp]),
        "b200_plugin_type": (ctypes.c_char_p, [vp]),
        "b200_plugin_version": (ctypes.c_char_p, [vp]),
        "b200_plugin_namespace": (ctypes.c_char_p, [vp]),
        "b200_plugin_nb_outputs": (i, [vp]),
        "b200_plugin_serialization_size": (sz, [vp]),
        "b200_plugin_serialize": (None, [vp, vp]),
        "b200_plugin_output_dims": (i, [vp, i, ctypes.POINTER(_Desc), i, ctypes.POINTER(_Desc)]),
        "b200_plugin_output_dtype": (i, [vp, i, ctypes.POINTER(ctypes.c_int32), i]),
        "b200_plugin_supports_format": (i, [vp, i, ctypes.POINTER(_Desc), i, i]),
        "b200_plugin_configure": (None, [vp, ctypes.POINTER(_Desc), i, ctypes.POINTER(_Desc), i]),
        "b200_plugin_workspace_size": (sz, [vp, ctypes.POINTER(_Desc), i, ctypes.POINTER(_Desc), i]),
        "b200_plugin_enqueue": (i, [vp, ctypes.POINTER(_Desc), i, ctypes.POINTER(_Desc), i, ctypes.POINTER(vp),
                                    ctypes.POINTER(vp), vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _bound = True


def _load_plugin_lib():
    """Same two steps as the reference: load with RTLD_GLOBAL, call initLibNvInferPlugins(None, namespace)."""
    lib = _lib.load()
    _bind(lib)
    assert lib.initLibNvInferPlugins(None, TRT_LLM_PLUGIN_NAMESPACE.encode('utf-8'))
    return lib


def get_plugin_creator(name, version='1', namespace=TRT_LLM_PLUGIN_NAMESPACE):
    """trt.get_plugin_registry().get_plugin_creator(name, version, namespace) -> opaque creator handle or None."""
    lib = _load_plugin_lib()
    return lib.b200_plugin_get_creator(name.encode(), version.encode(), namespace.encode())


def _desc(shape, dtype):
    d = _Desc()
    d.nb_dims = len(shape)
    for k, v in enumerate(shape):
        d.d[k] = int(v)
    d.dtype = TRT_DTYPE[dtype] if isinstance(dtype, str) else int(dtype)
    d.format = 0
    return d


def _descs(items):
    arr = (_Desc * max(len(items), 1))()
    for k, (shape, dtype) in enumerate(items):
        arr[k] = _desc(shape, dtype)
    return arr


class TrtPlugin:
    """Handle on one IPluginV2DynamicExt instance created through its registered creator."""

    def __init__(self, handle, lib):
        if not handle:
            raise RuntimeError("plugin creation failed (creator returned nullptr; see stderr)")
        self._h, self._lib = handle, lib

    @classmethod
    def create(cls, name, fields, version='1', namespace=TRT_LLM_PLUGIN_NAMESPACE, layer_name="layer"):
        """fields: list of (field_name, numpy array) exactly as the reference builds its trt.PluginField list."""
        lib = _load_plugin_lib()
        creator = lib.b200_plugin_get_creator(name.encode(), version.encode(), namespace.encode())
        assert creator, f"no creator registered for {name} v{version} in namespace {namespace}"
        keep = [np.ascontiguousarray(a) for _, a in fields]
        arr = (_Field * max(len(fields), 1))()
        for k, ((fname, _), a) in enumerate(zip(fields, keep)):
            arr[k] = _Field(fname.encode(), a.ctypes.data, _FIELD_TYPE[a.dtype], max(int(a.size), 1))
        return cls(lib.b200_plugin_create(creator, layer_name.encode(), arr, len(fields)), lib)

    @classmethod
    def deserialize(cls, name, data: bytes, version='1', namespace=TRT_LLM_PLUGIN_NAMESPACE):
        lib = _load_plugin_lib()
        creator = lib.b200_plugin_get_creator(name.encode(), version.encode(), namespace.encode())
        buf = ctypes.create_string_buffer(data, len(data))
        return cls(lib.b200_plugin_deserialize(creator, b"layer", ctypes.cast(buf, ctypes.c_void_p), len(data)), lib)

    @staticmethod
    def field_names(name, version='1', namespace=TRT_LLM_PLUGIN_NAMESPACE):
        lib = _load_plugin_lib()
        creator = lib.b200_plugin_get_creator(name.encode(), version.encode(), namespace.encode())
        buf = ctypes.create_string_buffer(4096)
        n = lib.b200_plugin_creator_field_names(creator, buf, 4096)
        names = buf.value.decode().split("\n")[:n]
        return names

    def clone(self):
        return TrtPlugin(self._lib.b200_plugin_clone(self._h), self._lib)

    def destroy(self):
        if self._h:
            self._lib.b200_plugin_destroy(self._h)
            self._h = None

    plugin_type = property(lambda self: self._lib.b200_plugin_type(self._h).decode())
    plugin_version = property(lambda self: self._lib.b200_plugin_version(self._h).decode())
    plugin_namespace = property(lambda self: self._lib.b200_plugin_namespace(self._h).decode())
    num_outputs = property(lambda self: self._lib.b200_plugin_nb_outputs(self._h))

    def serialize(self) -> bytes:
        n = self._lib.b200_plugin_serialization_size(self._h)
        buf = ctypes.create_string_buffer(n)
        self._lib.b200_plugin_serialize(self._h, ctypes.cast(buf, ctypes.c_void_p))
        return buf.raw

    def output_dims(self, index, inputs):
        """inputs: list of (shape, dtype)."""
        out = _Desc()
        rc = self._lib.b200_plugin_output_dims(self._h, index, _descs(inputs), len(inputs), ctypes.byref(out))
        if rc != 0:
            raise RuntimeError("getOutputDimensions failed")
        return tuple(out.d[k] for k in range(out.nb_dims))

    def output_dtype(self, index, input_dtypes):
        arr = (ctypes.c_int32 * len(input_dtypes))(*[TRT_DTYPE[d] for d in input_dtypes])
        return self._lib.b200_plugin_output_dtype(self._h, index, arr, len(input_dtypes))

    def supports_format(self, pos, in_out, nb_inputs):
        return bool(self._lib.b200_plugin_supports_format(self._h, pos, _descs(in_out), nb_inputs, len(in_out) - nb_inputs))

    def configure(self, inputs, outputs):
        self._lib.b200_plugin_configure(self._h, _descs(inputs), len(inputs), _descs(outputs), len(outputs))

    def workspace_size(self, inputs, outputs):
        return self._lib.b200_plugin_workspace_size(self._h, _descs(inputs), len(inputs), _descs(outputs), len(outputs))

    def enqueue(self, inputs, outputs, input_ptrs, output_ptrs, workspace_ptr, stream_ptr):
        ip = (ctypes.c_void_p * len(input_ptrs))(*input_ptrs)
        op = (ctypes.c_void_p * len(output_ptrs))(*output_ptrs)
        return self._lib.b200_plugin_enqueue(self._h, _descs(inputs), len(inputs), _descs(outputs), len(outputs), ip, op,
                                             workspace_ptr, stream_ptr)


class PluginConfig:
    """The plugin switches of T/tensorrt_llm/plugin/plugin.py:33-140 that concern the Whisper path."""

    def __init__(self):
        self.gpt_attention_plugin = False
        self.weight_only_quant_matmul_plugin = False
        self.context_fmha_type = 0
        self.remove_input_padding = False
        self.paged_kv_cache = False
        self.in_flight_batching = False

    def set_gpt_attention_plugin(self, dtype='float16'):
        self.gpt_attention_plugin = dtype
        return self

    def set_weight_only_quant_matmul_plugin(self, dtype='float16'):
        self.weight_only_quant_matmul_plugin = dtype
        return self
