"""GPU parity of the whole quantized decoder (WhisperDecoding) against the oracle model with identically dequantized
weights and int8-round-tripped KV: logits within tolerance, greedy token ids identical (BASELINE.json gate 3), on
seeded synthetic weights/inputs (no checkpoint offline)."""
import pytest
import torch

from oracle import whisper_oracle as wo

pytestmark = pytest.mark.gpu

PROMPT = [3, 7, 11]


def build(dims, seed, B, S_enc=None):
    from b200_whisper.runtime import WhisperDecoding
    sd = wo.synthetic_state_dict(dims, seed=seed, decoder_only=True)
    sdq = wo.quantize_state_dict(sd, dims)
    torch.manual_seed(100 + seed)
    S_enc = S_enc or dims.n_audio_ctx
    xa = torch.randn(B, S_enc, dims.n_text_state).half().float()
    with torch.no_grad():
        kv_s, ckv_s = wo.calibrate_kv_scales(sdq, dims, xa, PROMPT, n_steps=6)
    dec = WhisperDecoding(dims, sd, B, kv_s, ckv_s, n_audio_ctx=S_enc)
    dec.set_encoder_output(xa.cuda().half())
    return sdq, xa, kv_s, ckv_s, dec


@pytest.mark.parametrize("use_graph", [False, True])
def test_micro_greedy_tokens_and_logits(use_graph):
    dims = wo.MICRO
    B, n_new = 2, 16
    sdq, xa, kv_s, ckv_s, dec = build(dims, seed=1, B=B)
    with torch.no_grad():
        ref_tokens, ref_logits = wo.greedy_decode(sdq, dims, xa, PROMPT, n_new, kv_s, ckv_s, act_fp16=True)
    margins = torch.stack([(l.topk(2).values[:, 0] - l.topk(2).values[:, 1]) for l in ref_logits], 1)
    got = dec.decode([PROMPT] * B, n_new, use_graph=use_graph).cpu().long()
    # token identity wherever the oracle's own top-1 margin exceeds the fp16 noise floor; a flipped near-tie would
    # change every later token, so compare up to the first near-tie per sequence
    for b in range(B):
        upto = n_new
        weak = (margins[b] < 0.02).nonzero()
        if len(weak):
            upto = int(weak[0]) + 1
        assert upto >= 4, "pick another seed: oracle margins too small"
        assert got[b, :upto].tolist() == ref_tokens[b, :upto].tolist(), (b, got[b].tolist(), ref_tokens[b].tolist())


def test_micro_first_step_logits_tolerance():
    dims = wo.MICRO
    B = 2
    sdq, xa, kv_s, ckv_s, dec = build(dims, seed=2, B=B)
    with torch.no_grad():
        logits, _ = wo.decoder_forward(sdq, dims, torch.tensor([PROMPT] * B), xa, None, kv_s, ckv_s, act_fp16=True)
    dec.reset()
    dec.prefill([PROMPT] * B)
    torch.cuda.synchronize()
    got = dec.logits.cpu()
    ref = logits[:, -1]
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item(), f"prefill logits err {err} (ref max {ref.abs().max().item()})"


def test_tiny_dims_two_steps():
    """Whisper-tiny width (d=384, 6 heads, vocab 51865, 1500 frames) with 2 layers: context + 3 generation steps."""
    dims = wo.ModelDimensions(80, 1500, 384, 6, 2, 51865, 448, 384, 6, 2)
    B = 2
    sdq, xa, kv_s, ckv_s, dec = build(dims, seed=3, B=B)
    with torch.no_grad():
        ref_tokens, ref_logits = wo.greedy_decode(sdq, dims, xa, PROMPT, 4, kv_s, ckv_s, act_fp16=True)
    dec.reset()
    dec.prefill([PROMPT] * B)
    torch.cuda.synchronize()
    err = (dec.logits.cpu() - ref_logits[0]).abs().max().item()
    assert err <= 2e-2 * ref_logits[0].abs().max().item(), f"logits err {err}"
    for t in range(1, 4):
        # teacher-force the oracle's tokens so later steps stay comparable even if a near-tie flips
        dec.tokens.copy_(ref_tokens[:, t - 1].to(torch.int32))
        dec.step()
        torch.cuda.synchronize()
        err = (dec.logits.cpu() - ref_logits[t]).abs().max().item()
        assert err <= 2e-2 * ref_logits[t].abs().max().item(), f"step {t} logits err {err}"


def test_large_v2_full_size_determinism_and_utterance_independence():
    """BASELINE full size (large-v2 decoder, 32 layers, 1500 encoder frames), size-independent properties:
      * run-to-run determinism: two decoders built from the same seeds produce bit-identical logits and tokens
        (every split-K / split-KV reduction is in fixed order);
      * utterance independence -- the multi-GPU sharding property: the logits of an utterance do not depend,
        bit for bit, on which other utterances share its batch (here: rows 0-1 decoded in a batch of 4 and in a batch of 2),
        CUDA graph and eager."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from b200_whisper.runtime import WhisperDecoding
    dev = torch.device("cuda")
    dims = bench.Dims()
    L = dims.n_text_layer
    sd = bench.gpu_state_dict(dims, dev, seed=0)
    g = torch.Generator(device=dev).manual_seed(7)
    caches = [torch.randint(-127, 128, (4, 2, dims.n_text_head, dims.n_audio_ctx, 64), generator=g, device=dev,
                            dtype=torch.int8) for _ in range(L)]

    def run(batch, use_graph):
        dec = WhisperDecoding(dims, sd, batch, kv_scales=[0.05] * L, cross_kv_scales=[0.03] * L, device=dev)
        dec.set_cross_kv([c[:batch].contiguous() for c in caches])
        dec.reset()
        toks = [dec.prefill([bench.PROMPT] * batch).clone()]
        logits = [dec.logits.clone()]
        if use_graph:
            dec.capture()
        for _ in range(3):
            toks.append((dec.step() if use_graph else (dec._step_body() or dec.next_tokens)).clone())
            logits.append(dec.logits.clone())
        torch.cuda.synchronize()
        return torch.stack(toks), torch.stack(logits)

    t4, l4 = run(4, True)
    t4b, l4b = run(4, True)
    assert torch.equal(t4, t4b) and torch.equal(l4, l4b), "two identical runs differ"
    assert torch.isfinite(l4).all()
    t2, l2 = run(2, True)
    assert torch.equal(l2, l4[:, :2]) and torch.equal(t2, t4[:, :2]), "an utterance's logits depend on its batch mates"
    t2e, l2e = run(2, False)
    assert torch.equal(l2e, l2) and torch.equal(t2e, t2), "CUDA-graph replay and eager launches differ"


def test_large_v2_width_two_layers_against_oracle():
    """Whisper large-v2 WIDTH (d = 1280, 20 heads, fc 5120, vocab 51865, 1500 encoder frames) with 2 layers, batch 3:
    every decode-shape kernel of the headline configuration (qkv 1280->3840, fc1 1280->5120 with folded LayerNorm, fc2
    5120->1280 with its recycled TMEM stages, cross-attention over 1500 frames, the 256-row cross-K/V projections)
    against the oracle with identically dequantized weights: logits within tolerance, greedy tokens identical while
    the oracle's own top-1 margin is above the fp16 noise floor."""
    dims = wo.ModelDimensions(80, 1500, 1280, 20, 2, 51865, 448, 1280, 20, 2)
    B, n_new = 3, 5
    sdq, xa, kv_s, ckv_s, dec = build(dims, seed=5, B=B)  # seed 5: the greedy tokens vary (36316, 20855, ..., 26186)
    with torch.no_grad():
        ref_tokens, ref_logits = wo.greedy_decode(sdq, dims, xa, PROMPT, n_new, kv_s, ckv_s, act_fp16=True)
    dec.reset()
    got = [dec.prefill([PROMPT] * B).clone()]
    torch.cuda.synchronize()
    err = (dec.logits.cpu() - ref_logits[0]).abs().max().item()
    assert err <= 2e-2 * ref_logits[0].abs().max().item(), f"prefill logits err {err}"
    dec.capture()
    for t in range(1, n_new):
        got.append(dec.step().clone())
        torch.cuda.synchronize()
        same_history = all(int(got[s][b]) == int(ref_tokens[b, s]) for s in range(t) for b in range(B))
        if same_history:
            err = (dec.logits.cpu() - ref_logits[t]).abs().max().item()
            assert err <= 3e-2 * ref_logits[t].abs().max().item(), f"step {t} logits err {err}"
    got = torch.stack(got, 1).cpu().long()
    margins = torch.stack([(l.topk(2).values[:, 0] - l.topk(2).values[:, 1]) for l in ref_logits], 1)
    for b in range(B):
        # tokens must be identical up to the first place where they differ -- and that place must be a near-tie of the
        # oracle itself (a flipped near-tie changes every later token); seed 5 leaves every row a clear start
        diff = (got[b] != ref_tokens[b]).nonzero()
        first = int(diff[0]) if len(diff) else n_new
        assert first >= 3, (b, got[b].tolist(), ref_tokens[b].tolist())
        if first < n_new:
            assert margins[b, first].item() < 0.05, (b, first, margins[b].tolist(), got[b].tolist(), ref_tokens[b].tolist())
