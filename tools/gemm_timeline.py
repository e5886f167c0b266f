"""Global-timer timeline of the 192 GEMM launches of one captured decoder step: for every launch the earliest CTA
entry, the earliest return of griddepcontrol.wait and the latest CTA exit (ns, relative to the first entry)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
os.environ["B200_TC_DEBUG"] = "1"  # the stamps only exist in a debug build of the library
import importlib
importlib.import_module("eddie-wang-hackathon2023_b200._build").build()
import torch

import bench
from b200_whisper import _lib
from b200_whisper.runtime import WhisperDecoding


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    dims = bench.Dims()
    L, B = dims.n_text_layer, 16
    sd = bench.gpu_state_dict(dims, dev, seed=0)
    dec = WhisperDecoding(dims, sd, B, kv_scales=[0.05] * L, cross_kv_scales=[0.03] * L, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    dec.set_cross_kv([torch.randint(-127, 128, (B, 2, 20, 1500, 64), generator=g, device=dev, dtype=torch.int8)
                      for _ in range(L)])
    dec.reset()
    dec.prefill([bench.PROMPT] * B)
    n = 6 * L
    big = torch.iinfo(torch.int64).max
    init = torch.tensor([[big, big, 0, 0]] * (2 * n), dtype=torch.int64, device=dev)
    buf = init.clone()
    lib.b200_debug_tc_timeline(buf.data_ptr(), 2 * n)  # capture() runs the body twice: warm-up then captured
    dec.capture()
    lib.b200_debug_tc_timeline(None, 0)
    for _ in range(3):
        dec.step()
    torch.cuda.synchronize()
    buf.copy_(init)
    dec.step()
    torch.cuda.synchronize()
    t = buf[n:].cpu().tolist()  # the captured launches are the second half
    t0 = t[0][0]
    names = ["qkv(fold)", "attn_out", "cross_q(fold)", "cross_out", "fc1(fold)", "fc2"]
    print("launch            entry  last-entry   dep-return     exit | exit-dep  dep-prev_gemm_exit")
    prev_exit = None
    for i, (a, b, c, e) in enumerate(t[:18] + t[-6:]):
        nm = names[i % 6]
        gap = (b - prev_exit) if prev_exit is not None else 0
        print(f"{nm:14s} {a - t0:8d} {e - t0:10d} {b - t0:10d} {c - t0:10d} | {c - b:8d} {gap:8d}")
        prev_exit = c
    # averages over all layers
    import statistics
    for j, nm in enumerate(names):
        post = [t[i][2] - t[i][1] for i in range(j, n, 6)]
        pre = [t[i][1] - t[i][0] for i in range(j, n, 6)]
        print(f"{nm:14s} mean post-dependency {statistics.mean(post):7.0f} ns, resident before dependency {statistics.mean(pre):7.0f} ns")
    for j, nm in ((1, "attn_out->cross_q"), (3, "cross_out->fc1"), (4, "fc1->fc2")):
        gaps = [t[i + 1][1] - t[i][2] for i in range(j, n - 1, 6)]
        print(f"{nm:20s} consumer dependency return - producer last exit: mean {statistics.mean(gaps):7.0f} ns")
    gaps = [t[i + 1][1] - t[i][2] for i in range(5, n - 1, 6)]
    print(f"{'fc2->qkv(next)':20s} consumer dependency return - producer last exit: mean {statistics.mean(gaps):7.0f} ns")


if __name__ == "__main__":
    main()
