"""CPU tests of the data formats either side of the hot path: the checkpoint dict the
reference's build flow loads (T/examples/whisper/build.py:146-154) and the per-layer KV-cache calibration files
T/examples/whisper/weight.py:236-243 reads."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import whisper_oracle as wo

REF_WEIGHT_PY = "/root/reference/tensorrt_llm_july-release-v1/examples/whisper/weight.py"


def test_checkpoint_round_trip(tmp_path):
    from b200_whisper.runtime import ModelDimensions, load_checkpoint, save_checkpoint
    dims = wo.MICRO
    sd = wo.synthetic_state_dict(dims, seed=1)
    p = str(tmp_path / "micro.pt")
    save_checkpoint(p, dims, sd)
    raw = torch.load(p, weights_only=True)
    assert set(raw) == {"dims", "model_state_dict"} and raw["dims"]["n_text_layer"] == dims.n_text_layer
    assert list(raw["dims"]) == ["n_mels", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer", "n_vocab",
                                 "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer"]  # torch_model.py:12-22
    d2, sd2 = load_checkpoint(p)
    assert isinstance(d2, ModelDimensions) and d2.__dict__ == dims.__dict__
    assert set(sd2) == set(sd) and all(torch.equal(sd2[k], sd[k]) for k in sd)


def test_checkpoint_errors(tmp_path):
    from b200_whisper.runtime import load_checkpoint, save_checkpoint
    dims = wo.MICRO
    sd = wo.synthetic_state_dict(dims, seed=1)
    p = str(tmp_path / "bad.pt")
    torch.save({"weights": 1}, p)
    with pytest.raises(ValueError, match="not a Whisper checkpoint"):
        load_checkpoint(p)
    broken = dict(sd)
    del broken["decoder.blocks.1.mlp.2.weight"]
    save_checkpoint(p, dims, broken)
    with pytest.raises(KeyError, match="decoder.blocks.1.mlp.2.weight"):
        load_checkpoint(p)
    broken = dict(sd)
    broken["decoder.blocks.0.mlp.0.weight"] = torch.zeros(3, 3)
    save_checkpoint(p, dims, broken)
    with pytest.raises(ValueError, match="decoder.blocks.0.mlp.0.weight"):
        load_checkpoint(p)
    wide = wo.ModelDimensions(80, 96, 128, 4, 2, 1024, 64, 128, 4, 2)     # head size 32
    save_checkpoint(p, wide, sd)
    with pytest.raises(ValueError, match="head size 64"):
        load_checkpoint(p)


def test_kv_scale_files(tmp_path):
    from b200_whisper.runtime import read_kv_scales, write_kv_scales
    from b200_whisper.runtime import checkpoint as ck
    q = str(tmp_path / "quantize")
    kv = [0.03125, 0.0421, 1.5]
    ckv = [0.011, 0.022, 0.033]
    write_kv_scales(q, kv, ckv)
    for i, s in enumerate(kv):
        name = "model.decoder.blocks." + str(i) + ".attn.query_key_value.scale_y_quant_orig.bin"   # weight.py:239
        raw = open(os.path.join(q, name), "rb").read()
        assert raw == struct.pack("<f", s)                                                           # fp32 [1]
        # what the reference's fromfile(dir, name, [1], np.float32) returns (weight.py:15-22)
        assert np.fromfile(os.path.join(q, name), dtype=np.float32).reshape([1])[0] == np.float32(s)
    assert read_kv_scales(q, 3) == [float(np.float32(s)) for s in kv]
    assert read_kv_scales(q, 3, cross=True) == [float(np.float32(s)) for s in ckv]
    with pytest.raises(FileNotFoundError):
        read_kv_scales(q, 4)
    with open(os.path.join(q, ck.SELF_KV_SCALE_FILE.format(i=1)), "wb") as f:
        f.write(struct.pack("<ff", 1.0, 2.0))
    with pytest.raises(ValueError, match="one positive fp32"):
        read_kv_scales(q, 3)
    if os.path.exists(REF_WEIGHT_PY):  # the literal the reference concatenates
        src = open(REF_WEIGHT_PY).read()
        assert "'.attn.query_key_value.scale_y_quant_orig.bin'" in src and "'model.decoder.blocks.'+str(i)+" in src
