#!/usr/bin/env python
"""bench.py -- decoder tokens/s of the quantized Whisper large-v2 decoder hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]): Whisper large-v2 decoder, int8 weight-only + int8 self/cross KV cache
(1500 encoder frames), batch 16 utterances per GPU, random-init weights and synthetic inputs.  A "step" is one
greedy decoder step for the whole batch (16 tokens).  `value` = tokens/s with everything resident in HBM (CUDA-graph
replays, token feedback on the device); `e2e` = the same through WhisperDecoding.step_host(): host token ids in
(pinned, H2D), host token ids out (D2H) every step.  Each step streams ~2.9 GB (weights + cross-KV) >> 126 MB L2, so
no explicit L2 flush is needed between steps.

Multi-GPU: utterances are independent, so ranks run the identical per-GPU workload (weak scaling, no collective on the
hot path); token ids are all-gathered once after the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decoder_tokens_per_s_large_v2_int8_bs16"
UNIT = "tokens/s"
BATCH = 16
PROMPT = [50258, 50259, 50359]  # sot, <|en|>, <|transcribe|>  (T/examples/whisper/decoding.py:314-319)


class Dims:
    n_mels, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer = 80, 1500, 1280, 20, 32
    n_vocab, n_text_ctx, n_text_state, n_text_head, n_text_layer = 51865, 448, 1280, 20, 32


def gpu_state_dict(dims, device, seed=0):
    """Random-init decoder weights with the checkpoint's key names, drawn directly on the GPU."""
    import math

    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    d = dims.n_text_state
    sd = {}

    def rn(*shape, std):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std).half()

    sd["decoder.token_embedding.weight"] = rn(dims.n_vocab, d, std=0.1)
    sd["decoder.positional_embedding"] = rn(dims.n_text_ctx, d, std=0.5)
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}"
        for a in ("attn", "cross_attn"):
            for nm, bias in (("query", True), ("key", False), ("value", True), ("out", True)):
                sd[f"{p}.{a}.{nm}.weight"] = rn(d, d, std=1 / math.sqrt(d))
                if bias:
                    sd[f"{p}.{a}.{nm}.bias"] = rn(d, std=0.02)
            sd[f"{p}.{a}_ln.weight"] = 1 + rn(d, std=0.05)
            sd[f"{p}.{a}_ln.bias"] = rn(d, std=0.02)
        sd[f"{p}.mlp.0.weight"] = rn(4 * d, d, std=1 / math.sqrt(d))
        sd[f"{p}.mlp.0.bias"] = rn(4 * d, std=0.02)
        sd[f"{p}.mlp.2.weight"] = rn(d, 4 * d, std=1 / math.sqrt(4 * d))
        sd[f"{p}.mlp.2.bias"] = rn(d, std=0.02)
        sd[f"{p}.mlp_ln.weight"] = 1 + rn(d, std=0.05)
        sd[f"{p}.mlp_ln.bias"] = rn(d, std=0.02)
    sd["decoder.ln.weight"] = 1 + rn(d, std=0.05)
    sd["decoder.ln.bias"] = rn(d, std=0.02)
    return sd


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Number of samples taken so far (brackets the timed region: the sampler itself starts before the warm-up
        because nvidia-smi needs ~0.1 s to come up, longer than a 64-step timed region)."""
        return len(self.lines)

    def stop(self, first=0, last=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = self.lines[first:last] if last is not None and last > first else self.lines[max(0, first - 2):]
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(steps, seconds_budget=25.0):
    """The reference's PyTorch CPU path (summarize.py --test_torch -> torch_model + greedy loop), as restated by
    oracle/whisper_oracle.py, fp32 on the host cores, large-v2 decoder, batch 16, cross K/V given (random), bounded
    to `steps` generation steps.  Returns the cpu_baseline dict."""
    import torch

    from oracle import whisper_oracle as wo
    torch.set_num_threads(os.cpu_count() or 1)
    dims = wo.LARGE_V2
    t0 = time.perf_counter()
    sd = wo.synthetic_state_dict(dims, seed=0, decoder_only=True)
    build_s = time.perf_counter() - t0
    B = BATCH
    state = wo.DecoderState(dims.n_text_layer)
    g = torch.Generator().manual_seed(1)
    for i in range(dims.n_text_layer):
        state.ck[i] = torch.randn(B, dims.n_audio_ctx, dims.n_text_state, generator=g)
        state.cv[i] = torch.randn(B, dims.n_audio_ctx, dims.n_text_state, generator=g)
    xa = torch.zeros(B, 1, dims.n_text_state)  # unused: cross K/V are pre-filled
    tokens = torch.tensor([PROMPT] * B)
    with torch.no_grad():
        logits, state = wo.decoder_forward(sd, dims, tokens, xa, state)  # context step, untimed
        cur = logits[:, -1].argmax(-1)[:, None]
        done = 0
        t0 = time.perf_counter()
        while done < steps and (time.perf_counter() - t0) < seconds_budget:
            logits, state = wo.decoder_forward(sd, dims, cur, xa, state)
            cur = logits[:, -1].argmax(-1)[:, None]
            done += 1
        dt = time.perf_counter() - t0
    return {"value": B * done / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} greedy decoder steps, batch {B}, large-v2 decoder fp32, cross K/V given; "
                      f"oracle/whisper_oracle.py (restates T/examples/whisper/torch_model.py); weights built in {build_s:.1f}s",
            "ms_per_step": 1e3 * dt / max(done, 1)}


REF_WHISPER = "/root/reference/tensorrt_llm_july-release-v1/examples/whisper"


def reference_cpu_baseline(steps, seconds_budget=25.0):
    """The UNMODIFIED reference PyTorch model (examples/whisper/torch_model.py: Whisper, TextDecoder.forward :196-218,
    install_kv_cache_hooks :270-301 -- what `summarize.py --test_torch` drives through decoding.py:743-783) on the host
    cores, fp32, large-v2 decoder, batch 16, same synthetic weights as the oracle port.  Only available where
    /root/reference exists (the build container); returns None elsewhere."""
    if not os.path.exists(os.path.join(REF_WHISPER, "torch_model.py")):
        return None
    import torch
    sys.dont_write_bytecode = True  # the reference tree is read-only
    sys.path.insert(0, REF_WHISPER)
    try:
        from torch_model import ModelDimensions, Whisper
    except Exception:
        return None
    finally:
        sys.path.pop(0)
    from oracle import whisper_oracle as wo
    torch.set_num_threads(os.cpu_count() or 1)
    d = wo.LARGE_V2
    # the encoder is not on the timed path: keep it at one layer so the model fits comfortably in host memory
    dims = ModelDimensions(d.n_mels, d.n_audio_ctx, d.n_audio_state, d.n_audio_head, 1, d.n_vocab, d.n_text_ctx,
                           d.n_text_state, d.n_text_head, d.n_text_layer)
    t0 = time.perf_counter()
    model = Whisper(dims).float().eval()
    sd = wo.synthetic_state_dict(d, seed=0, decoder_only=True)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing), (missing[:3], unexpected[:3])
    build_s = time.perf_counter() - t0
    B = BATCH
    g = torch.Generator().manual_seed(1)
    xa = torch.randn(B, d.n_audio_ctx, d.n_text_state, generator=g)
    cache, hooks = model.install_kv_cache_hooks()
    tokens = torch.tensor([PROMPT] * B)
    with torch.no_grad():
        logits = model.decoder(tokens, xa, kv_cache=cache)  # context step + cross K/V projection, untimed
        cur = logits[:, -1].argmax(-1)[:, None]
        done = 0
        t0 = time.perf_counter()
        while done < steps and (time.perf_counter() - t0) < seconds_budget:
            logits = model.decoder(cur, xa, kv_cache=cache)
            cur = logits[:, -1].argmax(-1)[:, None]
            done += 1
        dt = time.perf_counter() - t0
    for h in hooks:
        h.remove()
    return {"value": B * done / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"{done} greedy decoder steps, batch {B}, large-v2 decoder fp32: the reference's own "
                      f"examples/whisper/torch_model.py (TextDecoder + install_kv_cache_hooks) imported from "
                      f"/root/reference; model built in {build_s:.1f}s",
            "ms_per_step": 1e3 * dt / max(done, 1)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    # the reference's own module when its tree is present (this container), the oracle port of it otherwise (GPU box)
    cb = reference_cpu_baseline(max(args.steps, 1), seconds_budget=60.0) or cpu_baseline(max(args.steps, 1), seconds_budget=60.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "whisper-large-v2 decoder step, batch 16, 1500 encoder frames, CPU fp32 (reference PyTorch path)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def time_decoder(dims, sd, batch, dev, steps, warmup, seed):
    """Device-resident CUDA-graph steps of a decoder with `batch` utterances: returns ms per step (CUDA events)."""
    import torch
    from b200_whisper.runtime import WhisperDecoding
    L = dims.n_text_layer
    dec = WhisperDecoding(dims, sd, batch, kv_scales=[0.05] * L, cross_kv_scales=[0.03] * L, device=dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    dec.set_cross_kv([torch.randint(-127, 128, (batch, 2, dims.n_text_head, dims.n_audio_ctx, 64), generator=g, device=dev,
                                    dtype=torch.int8) for _ in range(L)])
    dec.reset()
    dec.prefill([PROMPT] * batch)
    dec.capture()
    for _ in range(max(warmup, 3)):
        dec.step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        dec.step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    del dec
    torch.cuda.empty_cache()
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="utterances per GPU (weak scaling; the headline is 16)")
    ap.add_argument("--total-batch", type=int, default=0,
                    help="BASELINE.json configs[3]: this many utterances sharded over the GPUs (strong scaling); "
                         "per-GPU batch = total / world.  The default run reports this at 64 as the `strong_scaling` key.")
    ap.add_argument("--no-extras", action="store_true", help="skip the batch-1 and strong-scaling extra measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--layers", type=int, default=None, help="debug only: fewer decoder layers (result is INVALID)")
    ap.add_argument("--profile", action="store_true",
                    help="profiling aid: run ONE eager decoder step between cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints no bench line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import b200_whisper as bw
    from b200_whisper.runtime import WhisperDecoding

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    dims = Dims()
    if args.layers:
        dims.n_text_layer = args.layers
    B = args.batch
    strong = args.total_batch > 0
    if strong:
        assert args.total_batch % world == 0, "--total-batch must be divisible by the number of GPUs"
        B = args.total_batch // world
    lib = bw.load()
    sd = gpu_state_dict(dims, dev, seed=rank)
    L = dims.n_text_layer
    dec = WhisperDecoding(dims, sd, B, kv_scales=[0.05] * L, cross_kv_scales=[0.03] * L, device=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    # synthetic int8 cross-KV cache (the encoder and CrossAttn_KV run once per utterance, outside the decoder step)
    dec.set_cross_kv([torch.randint(-127, 128, (B, 2, dims.n_text_head, dims.n_audio_ctx, 64), generator=g, device=dev,
                                    dtype=torch.int8) for _ in range(L)])
    dec.reset()
    dec.prefill([PROMPT] * B)
    if args.profile:
        for _ in range(2):
            dec.step()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        dec.step()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        print("profiled one eager decoder step", flush=True)
        return
    launches_before = lib.b200_launch_count()
    if not args.no_graph:
        dec.capture()
    # capture() runs the step body three times (warm-up, device-resident graph, host-I/O graph)
    launches_per_step = (lib.b200_launch_count() - launches_before) // 3 if not args.no_graph else None
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: device-resident steps -------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        dec.step()
    if rank == 0:
        # nvidia-smi needs ~0.1 s to deliver its first sample; keep the GPU busy with filler work meanwhile (NOT decoder
        # steps: they would advance the sequence length and change the measured workload)
        filler = torch.randn((4096, 4096), device=dev, dtype=torch.float16)
        t_w = time.perf_counter()
        while sampler.mark() == 0 and time.perf_counter() - t_w < 1.5:
            for _ in range(8):
                filler @ filler
            torch.cuda.synchronize(dev)
        del filler
    barrier()
    s_first = sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.b200_launch_count()
    ev0.record()
    for _ in range(args.steps):
        if dec._host_len >= dims.n_text_ctx:  # only with --steps beyond the text context: rewind the lengths (one fill)
            dec.rewind(len(PROMPT))
        dec.step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    eager_launches = lib.b200_launch_count() - l0
    clocks = sampler.stop(s_first, sampler.mark()) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1e3)

    # ---- e2e: host buffers in/out every step --------------------------------------------------------
    import numpy as np
    host_tokens = np.array(dec.next_tokens.cpu().numpy(), dtype=np.int32)
    # same self-attention lengths as the device-resident measurement above (the workload must not drift between the two)
    dec.rewind(len(PROMPT) + max(args.warmup, 3) - 3)
    for _ in range(3):
        host_tokens = dec.step_host(host_tokens).numpy().copy()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if dec._host_len >= dims.n_text_ctx:
            dec.rewind(len(PROMPT))
        host_tokens = dec.step_host(host_tokens).numpy().copy()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(te.item())

    # ---- e2e through the reference's own boundary: the same step, one plugin enqueue per operator ----------------------
    # (WeightOnlyQuantMatmulPlugin::enqueue x 6 + GPTAttentionPlugin::enqueue per layer via include/b200_plugin_harness.h,
    # unfused LayerNorm / bias / GELU / residual layers in between, eager launches, host token buffers in and out)
    e2e_plugin = None
    if not args.no_extras and args.layers is None:
        from b200_whisper.runtime import PluginDecoderStep
        stepper = PluginDecoderStep(dec)
        dec.rewind(len(PROMPT) + max(args.warmup, 3) - 3)
        n_p = max(4, min(args.steps, 16))
        for _ in range(3):
            host_tokens = stepper.step_host(host_tokens).numpy().copy()
        barrier()
        enq0 = stepper.enqueues
        t0 = time.perf_counter()
        for _ in range(n_p):
            host_tokens = stepper.step_host(host_tokens).numpy().copy()
        barrier()
        tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        e2e_plugin = {"value": world * B * n_p / float(tp.item()), "unit": UNIT, "steps": n_p,
                      "ms_per_step": 1e3 * float(tp.item()) / n_p,
                      "plugin_enqueues_per_step": (stepper.enqueues - enq0) // n_p,
                      "path": "IPluginV2DynamicExt::enqueue per operator (eager, unfused glue), host token buffers"}
        # the same enqueue sequence captured once in a CUDA graph (as TensorRT captures an execution context) and replayed
        try:
            stepper.capture()
            n_g = max(8, min(args.steps, 32))
            for _ in range(3):
                host_tokens = stepper.step_host_graph(host_tokens).numpy().copy()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_g):
                host_tokens = stepper.step_host_graph(host_tokens).numpy().copy()
            barrier()
            tg = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            e2e_plugin["cuda_graph"] = {"value": world * B * n_g / float(tg.item()), "unit": UNIT, "steps": n_g,
                                        "ms_per_step": 1e3 * float(tg.item()) / n_g,
                                        "path": "the same plugin enqueues and glue layers captured in one CUDA graph, host "
                                                "token buffers in and out per step"}
        except Exception as e:  # the eager figure stands on its own
            e2e_plugin["cuda_graph"] = {"error": str(e)[:200]}
        stepper.close()

    # ---- extras: BASELINE.json configs[1] (batch 1) and configs[3] (64 utterances sharded over the GPUs) -----------
    extras = {}
    if not args.no_extras and not args.no_graph and args.layers is None:
        steps_x = min(args.steps, 32)
        if world == 1:
            ms1 = time_decoder(dims, sd, 1, dev, steps_x, args.warmup, 300)
            b1_bytes = L * (12 * dims.n_text_state ** 2) + dims.n_vocab * dims.n_text_state * 2 \
                + L * 2 * dims.n_text_state * dims.n_audio_ctx + L * 2 * dims.n_text_state * (len(PROMPT) + steps_x)
            extras["batch1"] = {"workload": "configs[1]: batch 1 decoder step", "ms_per_step": ms1, "tokens_per_s": 1e3 / ms1,
                                "algorithmic_bytes_per_step": b1_bytes,
                                "frac_of_hbm_peak": b1_bytes / (ms1 / 1e3) / 1e9 / measured_peak()[0]}
        if not strong and 64 % world == 0:
            bs = 64 // world
            ms_s = time_decoder(dims, sd, bs, dev, steps_x, args.warmup, 400 + rank)
            ts = torch.tensor([ms_s], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            ms_s = float(ts.item())
            extras["strong_scaling"] = {"workload": "configs[3]: 64 utterances sharded over the GPUs", "total_batch": 64,
                                        "batch_per_gpu": bs, "n_gpus": world, "ms_per_step": ms_s,
                                        "tokens_per_s": 64 / (ms_s / 1e3), "scaling": "strong"}
    del sd

    # ---- rooflines, measured live with CUDA events around CUDA-graph replays (no host launch cost inside) ---------
    # The dominant kernel BY TIME is the tcgen05 weight-only GEMM (192 launches per step, ~60 % of the step in
    # profiles/r01_launches_*.txt); the dominant kernel BY BYTES is the cross-attention (68 % of the step's bytes).
    # Both get a roofline entry; `roofline` is the GEMM family, `roofline_cross_attention` the attention kernel.
    peak, peak_src = measured_peak()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")  # dram bytes per launch from the round's ncu --set full captures
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f)
    cur = torch.cuda.current_stream(dev)

    def graph_ms(body, reps=5):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            body()  # warm-up outside capture
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            body()
        gr.replay()
        torch.cuda.synchronize(dev)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(reps):
            gr.replay()
        k1.record()
        torch.cuda.synchronize(dev)
        return k0.elapsed_time(k1) / reps

    d = dims.n_text_state
    xs = torch.randn((B, d), device=dev).half()
    x2 = xs.clone()
    qkv_o = torch.empty((B, 3 * d), device=dev).half()
    q_o = torch.empty((B, d), device=dev).half()
    u_o = torch.empty((B, 4 * d), device=dev).half()
    ctx_i = torch.randn((B, d), device=dev).half()
    lib.b200_set_static_kv_hint(1)

    def gemm_family():
        # the six GEMM launches of every layer exactly as the decoder step issues them (folded LayerNorm, bias,
        # GELU, residual); 32 layers of distinct weights = 629 MB per sweep, far beyond L2
        for lay in dec.layers:
            if dec.fuse_ln:
                dec._gemm_ln(xs, lay["attn_ln"], B, lay["qkv"], qkv_o)
            else:
                dec._gemm(xs, B, lay["qkv"], qkv_o)
            dec._gemm(ctx_i, B, lay["attn_out"], x2, residual=x2)
            if dec.fuse_ln:
                dec._gemm_ln(x2, lay["cross_ln"], B, lay["cross_q"], q_o)
            else:
                dec._gemm(x2, B, lay["cross_q"], q_o)
            dec._gemm(ctx_i, B, lay["cross_out"], x2, residual=x2)
            if dec.fuse_ln:
                dec._gemm_ln(x2, lay["mlp_ln"], B, lay["fc1"], u_o, act=L_.ACT_GELU_ERF)
            else:
                dec._gemm(x2, B, lay["fc1"], u_o, act=L_.ACT_GELU_ERF)
            dec._gemm(u_o, B, lay["fc2"], x2, residual=x2)

    from b200_whisper import _lib as L_
    n_gemm = 6 * L
    gemm_ms = graph_ms(gemm_family) / n_gemm

    def gemm_bytes(lin):  # int8 weights + fp16 scales, activations in and out
        return lin.k * lin.n + 2 * lin.n + 2 * B * lin.k + 2 * B * lin.n
    lay0 = dec.layers[0]
    gemm_b = sum(gemm_bytes(lay0[nm]) for nm in ("qkv", "attn_out", "cross_q", "cross_out", "fc1", "fc2")) / 6.0
    gemm_ach = gemm_b / (gemm_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "woq_gemm_tc_kernel<16,10,7,cluster> (6 launches per layer: qkv, attn_out, cross_q, "
                "cross_out, fc1, fc2; LayerNorm folded in 3 of them)", "achieved": gemm_ach, "peak": peak,
                "unit": "GB/s", "frac": gemm_ach / peak, "traffic": traffic.get("woq_gemm_tc_kernel"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": gemm_b, "avg_launch_ms": gemm_ms,
                "launches_timed": n_gemm, "timing": "CUDA events around graph replays of the 192 launches, PDL on"}

    q = torch.randn((B, d), device=dev).half()
    out = torch.empty_like(q)

    def xattn_family():
        for i in range(L):  # 32 different caches (1.97 GB) per sweep: every launch misses L2
            lay = dec.layers[i]
            L_.check(lib.b200_cross_attention(q.data_ptr(), dec.cross_kv[i].data_ptr(), lay["ckv_qo"].data_ptr(),
                                              out.data_ptr(), B, 1, dims.n_text_head, 64, dims.n_audio_ctx, 1,
                                              dec.ws.data_ptr(), dec.ws.numel(),
                                              torch.cuda.current_stream(dev).cuda_stream))
    xa_ms = graph_ms(xattn_family) / L
    xa_bytes = 2 * dims.n_text_head * 64 * dims.n_audio_ctx * B  # 3.84 MB per sequence per call
    xa_ach = xa_bytes / (xa_ms / 1e3) / 1e9
    roofline_xa = {"bound": "hbm", "kernel": "cross_attention_rowhead_kernel<int8>", "achieved": xa_ach, "peak": peak,
                   "unit": "GB/s", "frac": xa_ach / peak, "traffic": traffic.get("cross_attention_rowhead_kernel"),
                   "algorithmic_bytes_per_launch": xa_bytes, "avg_launch_ms": xa_ms}

    # The same kernel IN the captured step, by difference: the whole step's graph with and without its 32
    # cross-attention launches (the isolated replays above have the GPU to themselves; inside the step the kernel shares
    # the SMs with the early-launched prologue of the next GEMM and starts from that GEMM's dependency wait)
    if not args.no_graph and not dec.step_kernel:
        def step_ms(n=24):
            for _ in range(3):
                dec.step()
            torch.cuda.synchronize(dev)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for _ in range(n):
                dec.step()
            k1.record()
            torch.cuda.synchronize(dev)
            return k0.elapsed_time(k1) / n
        dec.rewind(len(PROMPT) + max(args.warmup, 3))
        t_with = step_ms()
        dec._measure_without_cross_attention = True
        dec.graph = None
        dec.capture()
        dec.rewind(len(PROMPT) + max(args.warmup, 3))
        t_without = step_ms()
        dec._measure_without_cross_attention = False
        dec.graph = None
        dec.capture()
        xa_in_ms = max(t_with - t_without, 1e-6) / L
        roofline_xa["in_graph"] = {"avg_launch_ms": xa_in_ms, "achieved": xa_bytes / (xa_in_ms / 1e3) / 1e9,
                                   "frac": xa_bytes / (xa_in_ms / 1e3) / 1e9 / peak,
                                   "method": "(step graph with - step graph without its 32 cross-attention launches) / 32"}
        roofline_xa["isolated_note"] = "achieved / frac above: graph of 32 back-to-back launches over 32 distinct caches"

    # ---- per-shape GEMM numbers for the "GEMV HBM GB/s" half of the metric ------------------------------------
    gemm_stats = {}
    x16 = torch.randn((B, 5120), device=dev).half()
    o16 = torch.empty((B, 5120), device=dev).half()
    for name in ("qkv", "attn_out", "fc1", "fc2"):
        lins = [lay[name] for lay in dec.layers]

        def sweep():
            for lin in lins:
                L_.check(lib.b200_woq_int8_gemm(x16.data_ptr(), B, lin.k, lin.weight.data_ptr(), lin.scales.data_ptr(),
                                                lin.n, o16.data_ptr(), dec.ws.data_ptr(), dec.ws.numel(),
                                                torch.cuda.current_stream(dev).cuda_stream))
        gms = graph_ms(sweep) / L
        lin = lins[0]
        gbytes = gemm_bytes(lin)
        gemm_stats[f"{name}_{lin.k}x{lin.n}_m{B}"] = {"avg_launch_us": 1e3 * gms, "GB/s": gbytes / (gms / 1e3) / 1e9,
                                                     "frac_of_peak": gbytes / (gms / 1e3) / 1e9 / peak}
    lib.b200_set_static_kv_hint(0)

    # algorithmic bytes of one whole step, for the step-level roofline fraction
    t_mid = len(PROMPT) + args.warmup + args.steps // 2
    step_bytes = L * (12 * d * d) + L * 2 * (3 * d + 3 * d + 4 * d + 2 * d) + dims.n_vocab * d * 2 \
        + B * L * 2 * d * dims.n_audio_ctx + B * L * 2 * d * t_mid
    step_frac = step_bytes / (ms_max / args.steps / 1e3) / 1e9 / peak

    if world > 1:
        gathered = [torch.empty_like(dec.next_tokens) for _ in range(world)]
        dist.all_gather(gathered, dec.next_tokens)  # the only collective: token ids, after the timed region

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(steps=8)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f16 activations x int8 weights, int8 KV (fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": "whisper-large-v2 decoder greedy step: int8 weight-only + int8 self/cross KV, "
                                   f"batch {B} per GPU, 1500 encoder frames, {L} layers, fp16 logits over 51865 tokens",
                       "batch_per_gpu": B, "l2": "inputs larger than L2 (about 2.9 GB streamed per step)",
                       "cuda_graph": not args.no_graph, "valid": args.layers is None},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": 4 * B},
            "gpu_launches": int((launches_per_step or 0) * args.steps if launches_per_step else eager_launches),
            "launches_per_step": launches_per_step,
            "roofline": roofline,
            "roofline_cross_attention": roofline_xa,
            "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "frac_of_peak": step_frac},
            "kernels": gemm_stats,
        }
        line.update(extras)
        if e2e_plugin is not None:
            line["e2e_plugin"] = e2e_plugin
        if cb is not None:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
