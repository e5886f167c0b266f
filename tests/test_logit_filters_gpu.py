"""Device logit filters + greedy update (b200_whisper_filtered_argmax) against the reference roll-outs
(tests/golden/logit_filter_golden.npz) and against the oracle on longer random roll-outs, incl. inside the decoder."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from logit_cases import BLANK, CASES, EOT, NO_TS, PROMPT, SUPPRESS, TS_BEGIN, V, synth_logits  # noqa: E402

from oracle.logit_filters import FilterConfig, greedy_step  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "logit_filter_golden.npz"))


@pytest.mark.parametrize("case", range(len(CASES)))
def test_kernel_reproduces_reference_rollouts(case):
    from b200_whisper.functional import WhisperLogitFilter
    seed, steps, batch, mi, boost = CASES[case]
    filt = WhisperLogitFilter(batch, V, EOT, NO_TS, TS_BEGIN, BLANK, SUPPRESS, None if mi < 0 else mi)
    nxt = torch.zeros((batch,), dtype=torch.int32, device="cuda")
    for t in range(steps):
        lg = torch.from_numpy(synth_logits(seed, t, batch, boost)).cuda()
        keep = lg.clone()
        filt(lg, nxt)
        torch.cuda.synchronize()
        assert torch.equal(lg, keep)  # the logits are not modified
        assert nxt.cpu().tolist() == GOLD[f"c{case}_tokens"][t].tolist(), (case, t)
        np.testing.assert_allclose(filt.sum_logprobs.cpu().numpy(), GOLD[f"c{case}_sumlp"][t], rtol=2e-5, atol=3e-4)


def test_kernel_matches_oracle_on_long_random_rollout():
    from b200_whisper.functional import WhisperLogitFilter
    batch, steps = 5, 40
    cfg = FilterConfig(EOT, None, TS_BEGIN, BLANK, [], 30)  # no <|notimestamps|>, no suppress list
    filt = WhisperLogitFilter(batch, V, EOT, None, TS_BEGIN, BLANK, (), 30)
    nxt = torch.zeros((batch,), dtype=torch.int32, device="cuda")
    sampled = [[] for _ in range(batch)]
    sums = np.zeros(batch, np.float32)
    boost = ["ts", "text", "spread", "text", "ts", "ts", "text", "text", "text", "eot", "text"]
    for t in range(steps):
        lg = synth_logits(77, t, batch, boost)
        filt(torch.from_numpy(lg).cuda(), nxt)
        got = nxt.cpu().tolist()
        for b in range(batch):
            last = sampled[b][-1] if sampled[b] else PROMPT[-1]
            tok, sums[b] = greedy_step(lg[b], sampled[b], last, sums[b], cfg)
            sampled[b].append(tok)
        assert got == [s[-1] for s in sampled], t
    np.testing.assert_allclose(filt.sum_logprobs.cpu().numpy(), sums, rtol=1e-4, atol=1e-3)


def test_decoder_with_filters_matches_oracle_filters_on_its_own_logits():
    """WhisperDecoding with the filters enabled (inside the captured CUDA graph): every chosen token equals the oracle
    filter applied to the logits the decoder itself produced at that step."""
    from b200_whisper.runtime import WhisperDecoding
    from oracle import whisper_oracle as wo
    dims = wo.MICRO
    B, n_new, prompt = 2, 8, [3, 7, 11]
    sd = wo.synthetic_state_dict(dims, seed=1, decoder_only=True)
    torch.manual_seed(5)
    xa = torch.randn(B, dims.n_audio_ctx, dims.n_text_state).half()
    eot, ts_begin, blank = dims.n_vocab - 40, dims.n_vocab - 30, 5  # a miniature vocabulary layout
    cfg = FilterConfig(eot, None, ts_begin, blank, [1, 2, 9], 10)
    dec = WhisperDecoding(dims, sd, B, [0.05] * dims.n_text_layer, [0.05] * dims.n_text_layer)
    dec.enable_logit_filters(eot, None, ts_begin, blank, [1, 2, 9], 10)
    dec.set_encoder_output(xa.cuda())
    dec.reset()
    sampled = [[] for _ in range(B)]
    sums = np.zeros(B, np.float32)

    def check(tokens):
        lg = dec.logits.float().cpu().numpy()
        got = tokens.cpu().tolist()
        for b in range(B):
            last = sampled[b][-1] if sampled[b] else prompt[-1]
            tok, sums[b] = greedy_step(lg[b], sampled[b], last, sums[b], cfg)
            sampled[b].append(tok)
        assert got == [s[-1] for s in sampled]

    check(dec.prefill([prompt] * B))
    dec.capture()
    for _ in range(n_new - 1):
        check(dec.step())
    np.testing.assert_allclose(dec.logit_filter.sum_logprobs.cpu().numpy(), sums, rtol=1e-4, atol=1e-3)


def test_filter_kernel_leaves_the_shared_scratch_pool_clean():
    """The filter kernel borrows a slot of the library's pool of self-resetting arrival counters (64 slots handed out
    round-robin).  Regression: it used to leave its partial results behind, so after 64 more allocations another kernel
    (here the split cross-attention, which merges its parts when an arrival counter reaches the part count) started from
    non-zero counters and never merged.  Run the filter, cycle through the whole pool with cross-attention calls, and
    require every one of them to match the first."""
    from b200_whisper.functional import WhisperLogitFilter, cross_attention
    torch.manual_seed(0)
    B, H, S = 2, 2, 96            # 4 (row, head) pairs: the split kernel with global partials + counters
    q = torch.randn(B, H * 64, device="cuda").half()
    kv = torch.randint(-127, 128, (B, 2, H, S, 64), dtype=torch.int8, device="cuda")
    scale = torch.tensor([0.02], device="cuda")
    want = cross_attention(q, kv, scale, H, 64)
    V = 4096
    filt = WhisperLogitFilter(8, V, V - 40, None, V - 30, 5, [1, 2, 9], 10)
    logits = torch.randn(8, V, device="cuda")
    nxt = torch.empty(8, dtype=torch.int32, device="cuda")
    for _ in range(70):           # dirty (formerly) every slot of the pool
        filt(logits, nxt)
    torch.cuda.synchronize()
    for i in range(70):
        got = torch.full_like(want, float("nan"))   # a merge that never happens leaves the NaNs in place
        cross_attention(q, kv, scale, H, 64, out=got)
        assert torch.equal(got, want), i
