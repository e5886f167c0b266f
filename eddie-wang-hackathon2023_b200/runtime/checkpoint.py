"""Data formats either side of the hot path.

* Checkpoint: the OpenAI-style dict ``{"dims": {...}, "model_state_dict": {...}}`` that the reference's build flow
  loads with torch.load (T/examples/whisper/build.py:146-154,215-223; field names T/examples/whisper/torch_model.py:12-22).
* KV-cache calibration scales: one little-endian fp32 per decoder layer in
  ``<dir>/model.decoder.blocks.{i}.attn.query_key_value.scale_y_quant_orig.bin`` (= max|qkv| / 127), read by
  T/examples/whisper/weight.py:236-243 into kv_quant_orig_scale = t and kv_orig_quant_scale = 1 / t.
  The int8 CROSS-attention cache is a B200 extension (the reference keeps cross K/V in fp16), so its scales use a
  sibling name: ``model.decoder.blocks.{i}.cross_attn.key_value.scale_y_quant_orig.bin``.
"""
import os
from dataclasses import asdict, dataclass

import numpy as np
import torch

SELF_KV_SCALE_FILE = "model.decoder.blocks.{i}.attn.query_key_value.scale_y_quant_orig.bin"
CROSS_KV_SCALE_FILE = "model.decoder.blocks.{i}.cross_attn.key_value.scale_y_quant_orig.bin"


@dataclass
class ModelDimensions:  # field order of torch_model.py:12-22
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


_REQUIRED_DECODER_KEYS = ("decoder.token_embedding.weight", "decoder.positional_embedding", "decoder.ln.weight")


def check_state_dict(dims, state_dict):
    """Shape checks the runtime relies on, reported with the offending key (the reference fails inside TensorRT weight
    assignment instead)."""
    for k in _REQUIRED_DECODER_KEYS:
        if k not in state_dict:
            raise KeyError(f"checkpoint has no '{k}'")
    d = dims.n_text_state
    if tuple(state_dict["decoder.token_embedding.weight"].shape) != (dims.n_vocab, d):
        raise ValueError("decoder.token_embedding.weight does not match dims (n_vocab, n_text_state)")
    if tuple(state_dict["decoder.positional_embedding"].shape) != (dims.n_text_ctx, d):
        raise ValueError("decoder.positional_embedding does not match dims (n_text_ctx, n_text_state)")
    if d % dims.n_text_head != 0 or d // dims.n_text_head != 64:
        raise ValueError("the attention kernels are built for head size 64 (every released Whisper size)")
    for i in range(dims.n_text_layer):
        for name, shape in ((f"decoder.blocks.{i}.attn.query.weight", (d, d)),
                            (f"decoder.blocks.{i}.cross_attn.key.weight", (d, dims.n_audio_state)),
                            (f"decoder.blocks.{i}.mlp.0.weight", (4 * d, d)),
                            (f"decoder.blocks.{i}.mlp.2.weight", (d, 4 * d))):
            if name not in state_dict:
                raise KeyError(f"checkpoint has no '{name}'")
            if tuple(state_dict[name].shape) != shape:
                raise ValueError(f"{name}: shape {tuple(state_dict[name].shape)}, expected {shape}")


def load_checkpoint(path, map_location="cpu"):
    """-> (ModelDimensions, model_state_dict).  `path`: a torch-saved ``{"dims", "model_state_dict"}`` file."""
    ckpt = torch.load(path, map_location=map_location, weights_only=True)
    if not isinstance(ckpt, dict) or "dims" not in ckpt or "model_state_dict" not in ckpt:
        raise ValueError(f"{path}: not a Whisper checkpoint (expected the keys 'dims' and 'model_state_dict')")
    dims = ModelDimensions(**{k: int(v) for k, v in dict(ckpt["dims"]).items()})
    check_state_dict(dims, ckpt["model_state_dict"])
    return dims, ckpt["model_state_dict"]


def save_checkpoint(path, dims, state_dict):
    torch.save({"dims": asdict(dims) if not isinstance(dims, dict) else dict(dims),
                "model_state_dict": {k: v.detach().cpu() for k, v in state_dict.items()}}, path)


def write_kv_scales(quantize_dir, kv_scales, cross_kv_scales=None):
    """One fp32 per layer and file, the format weight.py:236-243 reads."""
    os.makedirs(quantize_dir, exist_ok=True)
    for i, s in enumerate(kv_scales):
        np.asarray([s], dtype="<f4").tofile(os.path.join(quantize_dir, SELF_KV_SCALE_FILE.format(i=i)))
    for i, s in enumerate(cross_kv_scales or ()):
        np.asarray([s], dtype="<f4").tofile(os.path.join(quantize_dir, CROSS_KV_SCALE_FILE.format(i=i)))


def read_kv_scales(quantize_dir, n_layer, cross=False):
    """-> list of n_layer floats (scale_y_quant_orig).  A missing or malformed file is an error, not a silent None
    (the reference's `fromfile` returns None and fails later on `1.0 / t`)."""
    pattern = CROSS_KV_SCALE_FILE if cross else SELF_KV_SCALE_FILE
    out = []
    for i in range(n_layer):
        p = os.path.join(quantize_dir, pattern.format(i=i))
        if not os.path.exists(p):
            raise FileNotFoundError(p)
        t = np.fromfile(p, dtype="<f4")
        if t.shape != (1,) or not np.isfinite(t[0]) or t[0] <= 0:
            raise ValueError(f"{p}: expected one positive fp32, got {t!r}")
        out.append(float(t[0]))
    return out
