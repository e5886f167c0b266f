"""CPU tests: oracle/whisper_oracle.py (restatement of T/examples/whisper/torch_model.py) against golden vectors
produced by the reference model itself (tests/golden/make_whisper_golden.py), and the internal consistency of the
quantized-semantics helpers."""
import os

import numpy as np
import pytest
import torch

from oracle import whisper_oracle as wo
from oracle import woq

GOLD = os.path.join(os.path.dirname(__file__), "golden", "whisper_micro_golden.npz")
REF_DIR = "/root/reference/tensorrt_llm_july-release-v1/examples/whisper"


def test_matches_reference_golden():
    g = np.load(GOLD)
    dims = wo.MICRO
    sd = wo.synthetic_state_dict(dims, seed=1)
    mel = torch.from_numpy(g["mel"].astype(np.float32))
    with torch.no_grad():
        xa = wo.encoder_forward(sd, dims, mel)
        assert np.allclose(xa[:, ::16, ::16].numpy(), g["xa_sample"], atol=1e-5)
        assert abs(xa.double().sum().item() - float(g["xa_sum"])) < 1e-2
        prompt = g["prompt"].tolist()
        tokens = torch.tensor(prompt).repeat(mel.shape[0], 1)
        logits, _ = wo.decoder_forward(sd, dims, tokens, xa)
        assert np.allclose(logits[:, :, ::8].numpy(), g["prompt_logits_sample"], atol=1e-4)
        toks, step_logits = wo.greedy_decode(sd, dims, xa, prompt, g["tokens"].shape[1])
    assert np.array_equal(toks.numpy(), g["tokens"])
    assert np.allclose(torch.stack(step_logits, 1)[:, :, ::8].numpy(), g["step_logits_sample"], atol=1e-4)


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="/root/reference absent")
def test_matches_reference_live():
    import sys
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_DIR)
    try:
        from torch_model import ModelDimensions, Whisper
    finally:
        sys.path.remove(REF_DIR)
    dims = wo.MICRO
    sd = wo.synthetic_state_dict(dims, seed=5)
    model = Whisper(ModelDimensions(**dims.__dict__)).float().eval()
    model.load_state_dict(sd, strict=False)
    torch.manual_seed(3)
    mel = torch.randn(1, dims.n_mels, 2 * dims.n_audio_ctx).clamp(-1, 1)
    tokens = torch.tensor([[1, 2, 3, 4]])
    with torch.no_grad():
        xa_r = model.encoder(mel)
        xa_o = wo.encoder_forward(sd, dims, mel)
        assert torch.allclose(xa_r, xa_o, atol=1e-5)
        assert torch.allclose(model.decoder(tokens, xa_r), wo.decoder_forward(sd, dims, tokens, xa_o)[0], atol=1e-4)


def test_incremental_equals_full_context():
    # the explicit KV cache must reproduce a full-context forward (what the hook-based cache guarantees upstream)
    dims = wo.MICRO
    sd = wo.synthetic_state_dict(dims, seed=2)
    torch.manual_seed(0)
    xa = torch.randn(2, dims.n_audio_ctx, dims.n_audio_state)
    tokens = torch.randint(0, dims.n_vocab, (2, 6))
    with torch.no_grad():
        full, _ = wo.decoder_forward(sd, dims, tokens, xa)
        lg, st = wo.decoder_forward(sd, dims, tokens[:, :3], xa)
        outs = [lg]
        for t in range(3, 6):
            lg, st = wo.decoder_forward(sd, dims, tokens[:, t:t + 1], xa, st)
            outs.append(lg)
    assert torch.allclose(torch.cat(outs, 1), full, atol=1e-4)


def test_dequantized_weight_semantics():
    torch.manual_seed(0)
    w = torch.randn(128, 64) * 0.05  # [out, in]
    wq = wo.dequantized_linear_weight(w)
    # same thing from first principles: fp16(fp16(q) * s16), with q, s from the (reference-pinned) quantizer
    raw, _, scales = woq.symmetric_quantize_int8(w.half().t().contiguous().numpy(), np.float16)
    expect = (raw.astype(np.float16) * scales[None, :]).astype(np.float32).T
    assert np.array_equal(wq.numpy(), expect)
    assert (wq - w).abs().max() <= w.abs().max() / 128 * 0.51 + 1e-3


def test_kv_roundtrip_error_bound():
    torch.manual_seed(0)
    x = (torch.randn(4, 7, 64) * 2).half().float()
    s = float(x.abs().max()) / 127.0
    y = wo.kv_int8_roundtrip(x, s)
    assert (y - x).abs().max() <= 0.5 * s * 1.01 + 2e-3
