"""Phase timeline (SM cycles, CTA (0,0,0)) of one GEMM of the CAPTURED decoder step: the stamps of the last layer's
launch with the given N / folded-LN flag survive a graph replay.  usage: tc_timing_insitu.py [N fold] ..."""
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

names = {0: "entry", 1: "prologue done", 2: "weight TMA issued", 15: "dependency wait returned (producer)",
         3: "first weight tile landed", 4: "first A tile in TMEM", 13: "dequant done, stats begin", 14: "stats pushed",
         6: "last MMA committed", 7: "accumulator ready", 8: "partials pushed", 9: "inbox complete", 11: "stats merged",
         10: "slice reduced+stored", 12: "TMEM freed"}


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    dims = bench.Dims()
    L, B = dims.n_text_layer, 16
    sd = bench.gpu_state_dict(dims, dev, seed=0)
    cases = [(1280, 1), (1280, 0), (3840, 1), (5120, 1)]
    dbg = torch.zeros(80, dtype=torch.int64, device=dev)
    for n, fold in cases:
        dec = WhisperDecoding(dims, sd, B, kv_scales=[0.05] * L, cross_kv_scales=[0.03] * L, device=dev)
        g = torch.Generator(device=dev).manual_seed(1)
        dec.set_cross_kv([torch.randint(-127, 128, (B, 2, 20, 1500, 64), generator=g, device=dev, dtype=torch.int8)
                          for _ in range(L)])
        dec.reset()
        dec.prefill([bench.PROMPT] * B)
        lib.b200_debug_tc_timing_filter(n, fold)
        lib.b200_debug_tc_timing(dbg.data_ptr())
        dec.capture()
        lib.b200_debug_tc_timing(None)
        for _ in range(3):
            dec.step()
        torch.cuda.synchronize()
        dbg.zero_()
        dec.step()
        torch.cuda.synchronize()
        t = dbg.cpu().tolist()
        print(f"--- last launch with N={n} folded_ln={fold} inside the captured step")
        for sl in sorted(names, key=lambda k: t[k] if t[k] else 1 << 62):
            if t[sl]:
                print(f"   {names[sl]:38s} +{t[sl] - t[0]:7d} cycles")
        for i in range(6):
            q = t[16 + 4 * i: 20 + 4 * i]
            if any(q):
                print("     kb %2d: dequant [next loaded %7d, A arrived %7d]   mma [woke %7d, committed %7d]"
                      % (i, *[v - t[0] if v else 0 for v in q]))
        del dec
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
