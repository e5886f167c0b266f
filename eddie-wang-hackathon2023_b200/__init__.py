"""B200-native (sm_100a) quantized Whisper decoder hot path, behind the reference's operator API.

Host-side mirror of the reference's Python interface for this path (same names and argument meaning), bound with
ctypes to the C ABI in include/b200_whisper.h.  Import as ``b200_whisper`` (see ../b200_whisper.py).

    ops.symmetric_quantize_last_axis_of_batched_matrix   <- torch.ops.fastertransformer.* (thop/weightOnlyQuantOp.cpp)
    quantization.functional.weight_only_quant_matmul     <- tensorrt_llm/quantization/functional.py:51-74
    quantization.layer.WeightOnlyQuantLinear/RowLinear   <- tensorrt_llm/quantization/layer.py:268-382
    quantization.mode.QuantMode                          <- tensorrt_llm/quantization/mode.py
    functional.gpt_attention / conv1d / ...              <- tensorrt_llm/functional.py:2202-2244,2738-2971
    runtime.WhisperDecoding                              <- examples/whisper/decoding.py (greedy loop, CUDA graph)
    whisper_utils.log_mel_spectrogram / pad_or_trim      <- examples/whisper/whisper_utils.py:56-145 (GPU log-Mel front end)
    tokenizer.get_tokenizer / Tokenizer                  <- examples/whisper/tokenizer.py:125-265, decoding.py:423-486
    runtime.WhisperPipeline, load_checkpoint, *_kv_scales <- examples/whisper/run.py:33-66, build.py:146-154, weight.py:236-243
    summarize.evaluate / word_error_rate / load_dataset  <- examples/whisper/summarize.py:56-185 (WER harness)
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401
from . import functional  # noqa: F401
from . import quantization  # noqa: F401
from .quantization import QuantMode  # noqa: F401
from . import runtime  # noqa: F401
from . import whisper_utils  # noqa: F401
from . import tokenizer  # noqa: F401
from . import summarize  # noqa: F401

__all__ = ["ops", "functional", "quantization", "QuantMode", "whisper_utils", "tokenizer", "summarize", "load", "launch_count"]


def load():
    """Loads libb200_whisper.so (raises if it is not built; there is no fallback)."""
    return _lib.load()


def launch_count():
    """Number of CUDA kernels this library has launched in this process."""
    return int(_lib.load().b200_launch_count())
