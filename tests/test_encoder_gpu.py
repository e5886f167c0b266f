"""WhisperEncoder (conv stem on tcgen05, bidirectional attention, int8 weight-only GEMMs at M = B * frames) against the
oracle encoder (oracle/whisper_oracle.py, restating W/torch_model.py:152-171) with identically dequantized weights, and
the whole pipeline mel -> encoder -> int8 cross-KV -> greedy decoder against the oracle's tokens."""
import pytest
import torch

from oracle import whisper_oracle as wo

pytestmark = pytest.mark.gpu


def _mel(B, dims, seed):
    torch.manual_seed(seed)
    return torch.randn(B, dims.n_mels, 2 * dims.n_audio_ctx).clamp(-1, 1).half().float()


@pytest.mark.parametrize("dims_name,B", [("MICRO", 2), ("TINY_SHORT", 1)])
def test_encoder_matches_oracle(dims_name, B):
    from b200_whisper.runtime import WhisperEncoder
    if dims_name == "MICRO":
        dims = wo.MICRO
    else:  # tiny widths (384, 6 heads, 4 layers) over 200 frames: exercises partial key tiles and several m-tiles
        dims = wo.ModelDimensions(80, 200, 384, 6, 4, 1024, 64, 384, 6, 4)
    sd = wo.synthetic_state_dict(dims, seed=3)
    sdq = wo.quantize_state_dict(sd, dims, decoder_only=False)
    mel = _mel(B, dims, 11)
    with torch.no_grad():
        ref = wo.encoder_forward(sdq, dims, mel)
    enc = WhisperEncoder(dims, sd)
    out = enc(mel.cuda())
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs().max().item()
    # fp16 activations between kernels vs the fp32 oracle: a few 1e-3 of the (LayerNorm-normalised, O(1)) output
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), f"encoder output err {err}"
    assert (out.float().cpu() - ref).abs().mean().item() <= 3e-3


def test_mel_to_tokens_pipeline_matches_oracle():
    from b200_whisper.runtime import WhisperDecoding, WhisperEncoder
    dims = wo.MICRO
    B, n_new, prompt = 2, 6, [3, 7, 11]
    sd = wo.synthetic_state_dict(dims, seed=1)
    sdq = wo.quantize_state_dict(sd, dims, decoder_only=False)
    mel = _mel(B, dims, 21)
    enc = WhisperEncoder(dims, sd)
    xa = enc(mel.cuda())
    torch.cuda.synchronize()
    # the oracle decodes from the SAME encoder output (fp16 values), so token identity tests the decoder path while the
    # encoder parity is covered above
    xa_ref = xa.float().cpu()
    with torch.no_grad():
        kv_s, ckv_s = wo.calibrate_kv_scales(sdq, dims, xa_ref, prompt, n_steps=4)
        ref_tokens, _ = wo.greedy_decode(sdq, dims, xa_ref, prompt, n_new, kv_s, ckv_s, act_fp16=True)
    dec = WhisperDecoding(dims, sd, B, kv_s, ckv_s)
    dec.set_encoder_output(xa)
    got = dec.decode([prompt] * B, n_new)
    assert got.cpu().tolist() == ref_tokens.tolist()


def test_encoder_large_v2_width_full_length_matches_oracle():
    """The headline encoder shape end to end, not kernel by kernel: large-v2 width (1280, 20 heads), all 1500 frames,
    4 layers, batch 2 -- the conv stem at full size, the M = 3000 tcgen05 GEMMs with their in-place residual epilogues,
    the tcgen05 attention over 24 key tiles with a partial last one -- against the oracle encoder
    (T/tensorrt_llm/models/whisper/model.py:124-172; oracle W/torch_model.py:152-171) with identically dequantized
    weights."""
    from b200_whisper.runtime import WhisperEncoder
    dims = wo.ModelDimensions(80, 1500, 1280, 20, 4, 51865, 448, 1280, 20, 1)
    B = 2
    sd = wo.synthetic_state_dict(dims, seed=4)
    sd = {k: v for k, v in sd.items() if k.startswith("encoder.")}
    sdq = wo.quantize_state_dict(sd, dims, decoder_only=False)
    mel = _mel(B, dims, 31)
    with torch.no_grad():
        ref = wo.encoder_forward(sdq, dims, mel)
    enc = WhisperEncoder(dims, sd)
    out = enc(mel.cuda())
    torch.cuda.synchronize()
    assert tuple(out.shape) == (B, 1500, 1280) and torch.isfinite(out.float()).all()
    diff = (out.float().cpu() - ref).abs()
    scale = max(1.0, ref.abs().max().item())
    print(f"\n[encoder 1280 x 1500 x 4 layers, B=2] max |diff| {diff.max().item():.4f}, mean {diff.mean().item():.5f}, scale {scale:.2f}")
    assert diff.max().item() <= 2e-2 * scale, f"encoder output err {diff.max().item()}"
    assert diff.mean().item() <= 3e-3
