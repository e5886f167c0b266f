"""Per-CTA timeline of the whole-pair cross-attention kernel INSIDE the captured decoder step (library built with
B200_EXTRA_DEFS=-DB200_XA_DEBUG=1): %globaltimer stamps of every CTA at entry, dependency return, after warp 0's chunks
of pair 0 / 1 / 2 / ..., and exit, of the LAST cross-attention launch of a step.  Shows whether the kernel is bound by its
stream (all CTAs advance together) or by the CTAs that hold the most pairs."""
import os
import statistics
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import bench
from b200_whisper import _lib
from b200_whisper.runtime import WhisperDecoding


def main():
    layers = int(os.environ.get("LAYERS", "8"))
    B = int(os.environ.get("BATCH", "16"))
    dev = torch.device("cuda", 0)
    dims = bench.Dims()
    dims.n_text_layer = layers
    sd = bench.gpu_state_dict(dims, dev, seed=0)
    dec = WhisperDecoding(dims, sd, B, kv_scales=[0.05] * layers, cross_kv_scales=[0.03] * layers, device=dev)
    del sd
    g = torch.Generator(device=dev).manual_seed(1)
    dec.set_cross_kv([torch.randint(-127, 128, (B, 2, 20, 1500, 64), generator=g, device=dev, dtype=torch.int8)
                      for _ in range(layers)])
    lib = dec.lib
    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
    buf = torch.zeros((nsm, 8), dtype=torch.int64, device=dev)
    _lib.check(lib.b200_debug_xa_timeline(buf.data_ptr()))
    dec.reset()
    dec.prefill([bench.PROMPT] * B)
    dec.capture()
    for _ in range(6):
        dec.step()
    torch.cuda.synchronize()
    t = buf.cpu().numpy().astype("int64")
    used = t[:, 0] > 0
    if not used.any():
        print("no stamps: build with B200_EXTRA_DEFS=-DB200_XA_DEBUG=1")
        return
    t = t[used]
    t0 = t[:, 0].min()
    names = ["entry", "dependency return", "pair 0 done (warp 0)", "pair 1 done", "pair 2 done", "pair 3 done", "later", "exit"]
    print(f"# last cross-attention launch of a {layers}-layer batch-{B} step graph, {len(t)} CTAs; ns after the first CTA's entry")
    for k, nm in enumerate(names):
        col = t[:, k]
        col = col[col > 0] - t0
        if len(col) == 0:
            continue
        qs = statistics.quantiles(col.tolist(), n=10) if len(col) >= 10 else [float(col.min())] * 9
        print(f"{nm:24s} n={len(col):4d}  min {col.min():7d}  p10 {qs[0]:9.0f}  median {statistics.median(col.tolist()):9.0f}  "
              f"p90 {qs[8]:9.0f}  max {col.max():7d}")
    # CTAs by number of pairs
    has3 = t[:, 4] > 0
    for lab, sel in (("CTAs with a third pair stamp", has3), ("CTAs without", ~has3)):
        if sel.any():
            ex = t[sel, 7] - t0
            print(f"{lab:30s} n={int(sel.sum()):4d}  exit median {statistics.median(ex.tolist()):9.0f}  max {ex.max():7d}")
    dur_pairs = []
    for k in (3, 4):
        sel = (t[:, k] > 0) & (t[:, k - 1] > 0)
        if sel.any():
            d = t[sel, k] - t[sel, k - 1]
            dur_pairs.append(f"pair {k - 2}: median {statistics.median(d.tolist()):.0f} ns (n={int(sel.sum())})")
    sel = t[:, 2] > 0
    d0 = t[sel, 2] - t[sel, 1]
    print("per-pair time of warp 0 -- pair 0 (from the dependency return): "
          f"median {statistics.median(d0.tolist()):.0f} ns; " + "; ".join(dur_pairs))


if __name__ == "__main__":
    main()
