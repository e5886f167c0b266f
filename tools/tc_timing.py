"""Prints the phase timeline (SM cycles) of CTA (0,0,0) of the tcgen05 weight-only GEMM for the decode shapes."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
os.environ["B200_TC_DEBUG"] = "1"  # the stamps only exist in a debug build of the library
import importlib
importlib.import_module("eddie-wang-hackathon2023_b200._build").build()
import torch
import b200_whisper as bw
from b200_whisper import _lib

lib = _lib.load()
names = {0: "entry", 1: "prologue done", 2: "TMA issued", 13: "dequant ready to wait", 3: "first tile landed", 4: "first A tile in TMEM",
         5: "first MMA committed", 6: "last MMA committed", 7: "accumulator ready", 8: "partial stored", 9: "cluster sync 1",
         10: "slice reduced+stored", 12: "TMEM freed"}
dbg = torch.zeros(80, dtype=torch.int64, device="cuda")
for (m, k, n, ln, res) in [(256, 1280, 3840, False, False), (128, 1280, 3840, False, False), (16, 1280, 3840, False, False),
                           ]:
    torch.manual_seed(0)
    x = torch.randn((m, k), device="cuda").half()
    w = (torch.randn((k, n), device="cuda") * 0.05).half()
    proc, scales = bw.ops.symmetric_quantize_last_axis_of_batched_matrix(w, torch.int8)
    g = torch.ones(k, device="cuda").half(); b = torch.zeros(k, device="cuda").half()
    out = torch.zeros((m, n), device="cuda").half()
    ws = torch.empty(lib.b200_woq_workspace_bytes(m, n, k) + 1024, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def run():
        if ln:
            return lib.b200_woq_int8_gemm_ln_fused(x.data_ptr(), g.data_ptr(), b.data_ptr(), 1e-5, m, k, proc.data_ptr(), scales.data_ptr(), n,
                                                   None, 0, out.data_ptr() if res else None, out.data_ptr(), ws.data_ptr(), ws.numel(), st)
        return lib.b200_woq_int8_gemm_fused(x.data_ptr(), m, k, proc.data_ptr(), scales.data_ptr(), n, None, 0,
                                            out.data_ptr() if res else None, out.data_ptr(), ws.data_ptr(), ws.numel(), st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    lib.b200_debug_tc_timing(dbg.data_ptr())
    # cold-ish: flush L2 with a big write first
    junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); junk.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    lib.b200_debug_tc_timing(None)
    t = dbg.cpu().tolist()
    print(f"--- M={m} K={k} N={n} ln={ln} residual={res}: event time {e0.elapsed_time(e1)*1e3:.1f} us")
    for sl in [0, 1, 2, 13, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]:
        if t[sl]:
            print(f"   {names[sl]:26s} +{t[sl]-t[0]:7d} cycles")
    print("   per k-block (cycles since entry): dequant thread0 [loaded next, arrived]  mma thread [woke, committed]")
    for i in range(12):
        q = t[16 + 4 * i: 20 + 4 * i]
        if any(q):
            print("     kb %2d: deq %7d %7d   mma %7d %7d" % (i, *[v - t[0] if v else 0 for v in q]))
    dbg.zero_()
