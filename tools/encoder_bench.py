"""Times the large-v2 encoder (random-init int8 weight-only weights) for a batch of utterances and its main kernels."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import bench
from b200_whisper.runtime import WhisperEncoder

B = int(os.environ.get("BATCH", "16"))
dev = torch.device("cuda")
dims = bench.Dims()
g = torch.Generator(device=dev).manual_seed(0)
d, L = dims.n_audio_state, dims.n_audio_layer


def rn(*shape, std):
    return torch.randn(*shape, generator=g, device=dev) * std


sd = {"encoder.conv1.weight": rn(d, 80, 3, std=240 ** -0.5), "encoder.conv1.bias": rn(d, std=0.02),
      "encoder.conv2.weight": rn(d, d, 3, std=(3 * d) ** -0.5), "encoder.conv2.bias": rn(d, std=0.02),
      "encoder.positional_embedding": rn(dims.n_audio_ctx, d, std=0.1),
      "encoder.ln_post.weight": torch.ones(d, device=dev), "encoder.ln_post.bias": torch.zeros(d, device=dev)}
for i in range(L):
    p = f"encoder.blocks.{i}"
    for nm, (o, k) in {"attn.query": (d, d), "attn.key": (d, d), "attn.value": (d, d), "attn.out": (d, d),
                       "mlp.0": (4 * d, d), "mlp.2": (d, 4 * d)}.items():
        sd[f"{p}.{nm}.weight"] = rn(o, k, std=k ** -0.5)
        if nm != "attn.key":
            sd[f"{p}.{nm}.bias"] = rn(o, std=0.02)
    for nm in ("attn_ln", "mlp_ln"):
        sd[f"{p}.{nm}.weight"] = torch.ones(d, device=dev)
        sd[f"{p}.{nm}.bias"] = torch.zeros(d, device=dev)
enc = WhisperEncoder(dims, sd)
del sd
mel = torch.randn((B, 80, 3000), generator=g, device=dev).clamp(-1, 1).half()
for _ in range(2):
    out = enc(mel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = enc(mel)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
T = dims.n_audio_ctx
flop = B * (L * (2 * T * d * (12 * d) + 4 * T * T * d) + 2 * 3000 * d * 240 + 2 * T * d * 3 * d)
print(f"encoder large-v2, batch {B}: {ms:.1f} ms ({ms / B:.2f} ms per utterance), {flop / ms / 1e9:.0f} TFLOP/s "
      f"(GEMM + attention + stem FLOPs {flop / 1e12:.1f} T); output finite: {bool(torch.isfinite(out.float()).all())}")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    enc(mel)
    torch.cuda.synchronize()
import collections
agg = collections.defaultdict(float)
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name.split("<")[0].split("(")[0][-60:]] += e.time_range.end - e.time_range.start
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"   {v / 1e3:8.2f} ms  {k}")
# launch-by-launch view of the second layer (start relative to the layer's first kernel, duration)
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
             key=lambda e: e.time_range.start)
names = [e.name.split("<")[0].split("(")[0][-40:] for e in evs]
ln_idx = [i for i, n in enumerate(names) if "layernorm" in n]
if len(ln_idx) >= 5:
    i0, i1 = ln_idx[2], ln_idx[4]
    t0 = evs[i0].time_range.start
    print("second layer, launch by launch (us):")
    for e, n in zip(evs[i0:i1], names[i0:i1]):
        print(f"   +{e.time_range.start - t0:8.1f}  dur {e.time_range.end - e.time_range.start:8.1f}  {n}")
    print(f"   layer total {evs[i1].time_range.start - t0:8.1f}")
