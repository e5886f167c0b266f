"""Kernel timeline of ONE CUDA-graph replay of the batch-16 large-v2 decoder step (torch.profiler / CUPTI):
per-kernel start, duration, gap to the previous kernel's end, plus per-kernel-type totals.  Shows what ncu's
serialised launch list cannot: overlap from programmatic dependent launch and the real critical path."""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import b200_whisper as bw
from b200_whisper.runtime import WhisperDecoding


def main():
    layers = int(os.environ.get("LAYERS", "32"))
    dev = torch.device("cuda", 0)
    dims = bench.Dims()
    dims.n_text_layer = layers
    B = 16
    sd = bench.gpu_state_dict(dims, dev, seed=0)
    dec = WhisperDecoding(dims, sd, B, kv_scales=[0.05] * layers, cross_kv_scales=[0.03] * layers, device=dev)
    del sd
    g = torch.Generator(device=dev).manual_seed(1)
    dec.set_cross_kv([torch.randint(-127, 128, (B, 2, 20, 1500, 64), generator=g, device=dev, dtype=torch.int8)
                      for _ in range(layers)])
    dec.reset()
    dec.prefill([bench.PROMPT] * B)
    dec.capture()
    for _ in range(5):
        dec.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            dec.step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    # split into replays by the embed kernel
    starts = [i for i, e in enumerate(evs) if "embed_kernel" in e.name]
    if len(starts) < 2:
        print("could not find step boundaries; events:", len(evs))
        return
    step = evs[starts[1]:starts[2]] if len(starts) > 2 else evs[starts[1]:]
    t0 = step[0].time_range.start
    agg = collections.defaultdict(lambda: [0, 0.0])
    prev_end = t0
    lines = []
    for e in step:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        m = re.match(r"(?:void )?(?:b200::)?(\w+)", e.name)
        nm = m.group(1) if m else e.name[:40]
        gap = e.time_range.start - prev_end
        prev_end = max(prev_end, e.time_range.end)
        agg[nm][0] += 1
        agg[nm][1] += d
        lines.append(f"{s:9.1f} us  dur {d:7.2f}  gap {gap:7.2f}  {nm}")
    total = step[-1].time_range.end - t0
    print(f"step wall time {total:.1f} us, {len(step)} kernels, sum of durations {sum(v[1] for v in agg.values()):.1f} us")
    for nm, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {nm:40s} n={n:4d} total={t:8.1f} us avg={t / n:7.2f} us")
    print("first layers:")
    for ln in lines[:40]:
        print(ln)


if __name__ == "__main__":
    main()
