"""Debug aid for the persistent step kernel: after ONE step, decode the flagged-word buffers of its scratch area (last
layer's qkv, ctx, x[1], q, x[2], u, x[0]) and compare them with the operator chain's intermediates on the same state."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import whisper_oracle as wo  # noqa: E402
from b200_whisper.runtime import WhisperDecoding  # noqa: E402

L = int(os.environ.get("DBG_L", "1"))
B = int(os.environ.get("DBG_B", "2"))
NSTEP = int(os.environ.get("DBG_STEPS", "2"))
dims = wo.ModelDimensions(80, 96, 128, 2, 2, 1024, 64, 128, 2, L) if os.environ.get("DBG_DIMS", "micro") == "micro" else \
    wo.ModelDimensions(80, 1500, 1280, 20, 1, 51865, 448, 1280, 20, L)
sd = wo.synthetic_state_dict(dims, seed=1, decoder_only=True)
torch.manual_seed(101)
xa = torch.randn(B, dims.n_audio_ctx, dims.n_text_state).half().cuda()
d, dff = dims.n_text_state, 4 * dims.n_text_state


def frag_index(row, k):
    kb, kk = k >> 6, k & 63
    T, r = kk >> 4, kk & 15
    hi, w, e = r >> 3, (r & 7) >> 1, r & 1
    g, up = row & 7, row >> 3
    return ((((kb * 4 + w) * 32 + 4 * g + T) * 4 + 2 * hi + up) << 1) + e


def decode(words, n, frag):
    """words: int64 tensor of flagged words -> (values [B, n] fp16 as float, flags [B, n // 2])"""
    w = words.cpu().numpy().astype(np.uint64)
    pay = (w & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    flg = (w >> np.uint64(32)).astype(np.uint32)
    halves = pay.view(np.float16).astype(np.float32)  # 2 halves per word
    vals = np.zeros((B, n), np.float32)
    flags = np.zeros((B, n // 2), np.uint32)
    for r in range(B):
        for k in range(0, n, 2):
            wi = (frag_index(r, k) >> 1) if frag else ((r * n + k) >> 1)
            vals[r, k], vals[r, k + 1] = halves[2 * wi], halves[2 * wi + 1]
            flags[r, k // 2] = flg[wi]
    return vals, flags


decs = []
for use in (True, False):
    dec = WhisperDecoding(dims, sd, B, kv_scales=[0.04] * L, cross_kv_scales=[0.03] * L, n_audio_ctx=dims.n_audio_ctx)
    dec.step_kernel = use
    dec.set_encoder_output(xa)
    dec.reset()
    dec.prefill([[3, 7, 11]] * B)
    decs.append(dec)
new, old = decs
for step in range(NSTEP):
    for dec in decs:
        dec._step_body()
    torch.cuda.synchronize()
    print(f"--- step {step}: status {new.step_kernel_status():#x}; logits max diff "
          f"{(new.logits - old.logits).abs().max().item():.4g} (scale {old.logits.abs().max().item():.3g})")
    top = old.logits.topk(2, -1).values
    print("tokens new", new.next_tokens.cpu().tolist(), "old", old.next_tokens.cpu().tolist(), "old top-2 margins",
          [round(v, 4) for v in (top[:, 0] - top[:, 1]).cpu().tolist()])
    sc = new._step_scratch
    ctl = sc[:16].view(torch.int32).cpu().tolist()
    print("control words [arrivals, exits, status, step counter]:", ctl)
    off = 256
    bufs = {}
    for name, nbytes in (("x0", 64 * d), ("x1", 64 * d), ("x2", 64 * d), ("ctx", 64 * d), ("q", 64 * d), ("qkv", 192 * d),
                         ("u", 64 * dff)):
        bufs[name] = sc[off:off + nbytes].view(torch.int64)
        off += nbytes
    ref = {"qkv": old._bufs[("qkv", B, 3 * d, 0)], "q": old._bufs[("q", B, d, 0)], "ctx": old._bufs[("ctx", B, d, 0)],
           "u": old._bufs[("u", B, 4 * d, 0)], "x0": old._bufs[("x", B, d, 0)]}
    for name, n, frag in (("qkv", 3 * d, False), ("q", d, False), ("ctx", d, True), ("u", dff, True), ("x0", d, True)):
        vals, flags = decode(bufs[name], n, frag)
        r = ref[name].float().cpu().numpy()
        print(f"{name:4s} max |diff| {np.abs(vals - r).max():.4g} (scale {np.abs(r).max():.3g}); flags {sorted(set(flags.ravel().tolist()))}")
    xo = new._bufs[("x", B, d, 0)].float().cpu().numpy()
    print(f"x_out max |diff| {np.abs(xo - ref['x0'].float().cpu().numpy()).max():.4g}")
