"""tools/step_phases.py -- per-phase timeline of ONE launch of the persistent decoder-step kernel (large-v2, batch B):
%globaltimer stamps taken by every CTA when a grid barrier lets it through and when its work of the phase is done
(b200_debug_decoder_step_timeline).  Prints, per phase type, the time from the LAST CTA's "work done" of the previous
phase to the FIRST / LAST CTA passing the barrier (barrier cost), and the work time of the phase (mean / max over CTAs).

    python tools/step_phases.py [--batch 16] [--layers 32]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import b200_whisper as bw  # noqa: E402
from b200_whisper.runtime import WhisperDecoding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--layers", type=int, default=32)
args = ap.parse_args()
dev = torch.device("cuda")
dims = bench.Dims()
dims.n_text_layer = args.layers
L, B = dims.n_text_layer, args.batch
lib = bw.load()
sd = bench.gpu_state_dict(dims, dev, seed=0)
dec = WhisperDecoding(dims, sd, B, kv_scales=[0.05] * L, cross_kv_scales=[0.03] * L, device=dev)
g = torch.Generator(device=dev).manual_seed(7)
dec.set_cross_kv([torch.randint(-127, 128, (B, 2, dims.n_text_head, dims.n_audio_ctx, 64), generator=g, device=dev,
                                dtype=torch.int8) for _ in range(L)])
dec.reset()
dec.prefill([bench.PROMPT] * B)
for _ in range(3):
    dec.step()
torch.cuda.synchronize()
G = torch.cuda.get_device_properties(0).multi_processor_count
NP = 512
buf = torch.zeros((G * NP * 2 + NP * 16,), dtype=torch.int64, device=dev)
lib.b200_debug_decoder_step_timeline(buf.data_ptr())
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
dec.step()
ev1.record()
torch.cuda.synchronize()
lib.b200_debug_decoder_step_timeline(None)
print(f"step (eager launch incl. head) {ev0.elapsed_time(ev1) * 1e3:.1f} us; status {dec.step_kernel_status():#x}")
raw = buf.cpu().numpy().astype(np.int64)
t = raw[:G * NP * 2].reshape(G, NP, 2)
fine = raw[G * NP * 2:].reshape(NP, 16)
nph = 1 + 8 * L
t0 = t[:, 0, 0].min()
passed = t[:, 1:nph + 1, 0] - t0   # [G, phase]: barrier passed (phase index p = 1..nph waits for arrival p)
done = t[:, 0:nph, 1] - t0         # [G, phase]: work done of phase p (0 = embed)
names = ["embed", "qkv", "mmha", "attn_out", "cross_q", "xattn", "cross_out", "fc1", "fc2"]
print(f"kernel span {(done[:, nph - 1].max()) / 1e3:.1f} us over {nph} phases, {G} CTAs")
rows = {}
for p in range(1, nph):
    kind = names[1 + (p - 1) % 8]
    prev_last = done[:, p - 1].max()
    first_pass, last_pass = passed[:, p - 1].min(), passed[:, p - 1].max()
    work = done[:, p] - passed[:, p - 1]
    rows.setdefault(kind, []).append((first_pass - prev_last, last_pass - prev_last, work.mean(), work.max(),
                                      done[:, p].max() - prev_last))
print(f"{'phase':10s} {'barrier first':>14s} {'barrier last':>13s} {'work mean':>10s} {'work max':>9s} {'phase total':>12s}   (ns, mean over layers)")
tot = 0.0
for kind in names[1:]:
    a = np.array(rows[kind], dtype=np.float64)
    m = a.mean(0)
    tot += m[4]
    print(f"{kind:10s} {m[0]:14.0f} {m[1]:13.0f} {m[2]:10.0f} {m[3]:9.0f} {m[4]:12.0f}")
print(f"sum of phase totals per layer: {tot / 1e3:.2f} us")

# CTA 0 (thread 0), matmul phases, SM cycles after the barrier pass [0]: all weight tiles of the first quarter ready [4]
# (the A fragments were requested before), MMAs done [5], k-partials in shared memory [1], reduced / slots released [2],
# epilogue stored [3]
print("CTA 0 inside the matmul phases (SM cycles after the barrier pass; mean over layers):")
print(f"{'phase':10s} {'weights ready':>14s} {'MMAs done':>10s} {'partials':>9s} {'reduced':>8s} {'stored':>8s}")
for off, kind in ((1, "qkv"), (3, "attn_out"), (4, "cross_q"), (6, "cross_out"), (7, "fc1"), (8, "fc2")):
    v = np.array([[fine[off + 8 * l][k] - fine[off + 8 * l][0] for k in (4, 5, 1, 2, 3)] for l in range(L)], dtype=np.float64)
    m = v.mean(0)
    print(f"{kind:10s} {m[0]:14.0f} {m[1]:10.0f} {m[2]:9.0f} {m[3]:8.0f} {m[4]:8.0f}")
print("CTA 0 barrier path (SM cycles; mean over the matmul -> matmul barriers): arrive called -> release issued, "
      "release issued -> poll satisfied (includes waiting for slower CTAs), poll satisfied -> consumer warps released")
v = np.array([[fine[p][9] - fine[p][8], fine[p + 1][7] - fine[p][9], fine[p + 1][0] - fine[p + 1][7]]
              for p in range(1, nph - 1) if (p % 8) not in (2, 5) and ((p + 1) % 8) not in (2, 5)], dtype=np.float64)
print("  ", " ".join(f"{x:8.0f}" for x in v.mean(0)))
v = np.array([[fine[5 + 8 * l][k] for k in (10, 11, 12)] for l in range(L)], dtype=np.float64).mean(0)
print(f"CTA 30 (2 pairs), warp 0 in the cross-attention phase: {v[1]:.1f} chunks, {v[0]:.0f} cycles waiting for ring data, "
      f"{v[2]:.0f} cycles in the phase -> {(v[2] - v[0]) / max(v[1], 1):.0f} cycles of work per chunk")
