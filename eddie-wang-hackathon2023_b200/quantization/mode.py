"""QuantMode flag set -- same public names, bit values and predicates as the reference's
T/tensorrt_llm/quantization/mode.py (spec: T/tests/quantization/test_mode.py), so configs serialized by the
reference's build.py (--use_weight_only --weight_only_precision int8 --int8_kv_cache) mean the same thing here."""
import enum

_FLAG_NAMES = ("INT4_WEIGHTS", "INT8_WEIGHTS", "ACTIVATIONS", "PER_CHANNEL", "PER_TOKEN", "INT8_KV_CACHE", "FP8_KV_CACHE")


class QuantMode(enum.IntFlag):
    INT4_WEIGHTS = 1 << 0
    INT8_WEIGHTS = 1 << 1
    ACTIVATIONS = 1 << 2
    PER_CHANNEL = 1 << 3
    PER_TOKEN = 1 << 4
    INT8_KV_CACHE = 1 << 5
    FP8_KV_CACHE = 1 << 6
    COUNT = 1 << 7
    WEIGHTS_AND_ACTIVATIONS = (1 << 0) | (1 << 1) | (1 << 2)
    VALID_FLAGS = (1 << 7) - 1

    def _all(self, bits, mask=None):
        mask = QuantMode.VALID_FLAGS if mask is None else mask
        return (int(self) & int(mask)) == int(bits)

    def _any(self, bits):
        return (int(self) & int(bits)) != 0

    def is_int8_weight_only(self):
        return self._all(QuantMode.INT8_WEIGHTS, QuantMode.WEIGHTS_AND_ACTIVATIONS)

    def is_int4_weight_only(self):
        return self._all(QuantMode.INT4_WEIGHTS, QuantMode.WEIGHTS_AND_ACTIVATIONS)

    def is_weight_only(self):
        return self.is_int8_weight_only() or self.is_int4_weight_only()

    def has_act_and_weight_quant(self):
        return self._all(QuantMode.INT8_WEIGHTS | QuantMode.ACTIVATIONS, QuantMode.WEIGHTS_AND_ACTIVATIONS)

    def has_per_token_dynamic_scaling(self):
        return self._any(QuantMode.PER_TOKEN)

    def has_act_static_scaling(self):
        return not self.has_per_token_dynamic_scaling()

    def has_per_channel_scaling(self):
        return self._any(QuantMode.PER_CHANNEL)

    def has_int8_kv_cache(self):
        return self._any(QuantMode.INT8_KV_CACHE)

    def has_fp8_kv_cache(self):
        return self._any(QuantMode.FP8_KV_CACHE)

    def has_any_quant(self):
        return self._any(QuantMode.INT8_WEIGHTS | QuantMode.ACTIVATIONS | QuantMode.INT8_KV_CACHE | QuantMode.FP8_KV_CACHE)

    def set_int8_kv_cache(self):
        return self | QuantMode.INT8_KV_CACHE

    def set_fp8_kv_cache(self):
        return self | QuantMode.FP8_KV_CACHE

    @staticmethod
    def from_description(quantize_weights=False, quantize_activations=False, per_token=False, per_channel=False,
                         use_int4_weights=False, use_int8_kv_cache=False, use_fp8_kv_cache=False):
        bad = (quantize_activations and not quantize_weights) or (
            (per_token or per_channel) and not (quantize_weights and quantize_activations))
        if bad:
            raise ValueError(
                "Unsupported combination of QuantMode args: "
                f"{quantize_weights=}, {quantize_activations=}, {per_token=}, {per_channel=}, {use_int4_weights=}, "
                f"{use_int8_kv_cache=}, {use_fp8_kv_cache=}")
        picks = {
            "INT4_WEIGHTS": quantize_weights and use_int4_weights,
            "INT8_WEIGHTS": quantize_weights and not use_int4_weights,
            "ACTIVATIONS": quantize_activations,
            "PER_CHANNEL": per_channel,
            "PER_TOKEN": per_token,
            "INT8_KV_CACHE": use_int8_kv_cache,
            "FP8_KV_CACHE": use_fp8_kv_cache,
        }
        mode = QuantMode(0)
        for name in _FLAG_NAMES:
            if picks[name]:
                mode = mode | QuantMode[name]
        return mode

    @staticmethod
    def use_smooth_quant(per_token=False, per_channel=False):
        return QuantMode.from_description(True, True, per_token, per_channel)

    @staticmethod
    def use_weight_only(use_int4_weights=False):
        return QuantMode.from_description(True, False, False, False, use_int4_weights)
