"""Whole pipeline at full size on one B200: 16 waveforms of 30 s -> log-Mel -> large-v2 encoder -> int8 cross-KV caches ->
64 greedy decoder steps (CUDA graph, device logit filters) -> token ids on the host.  Random-init weights (no checkpoint
offline), synthetic audio; CUDA events per stage after one warm-up pass."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import bench
from b200_whisper.runtime import WhisperPipeline
from b200_whisper.tokenizer import get_tokenizer

B = int(os.environ.get("BATCH", "16"))
N_NEW = int(os.environ.get("TOKENS", "64"))
dev = torch.device("cuda")
dims = bench.Dims()
sd = bench.gpu_state_dict(dims, dev, seed=0)
g = torch.Generator(device=dev).manual_seed(1)
d, L = dims.n_audio_state, dims.n_audio_layer


def rn(*shape, std):
    return (torch.randn(*shape, generator=g, device=dev) * std).half()


sd.update({"encoder.conv1.weight": rn(d, 80, 3, std=240 ** -0.5), "encoder.conv1.bias": rn(d, std=0.02),
           "encoder.conv2.weight": rn(d, d, 3, std=(3 * d) ** -0.5), "encoder.conv2.bias": rn(d, std=0.02),
           "encoder.positional_embedding": rn(dims.n_audio_ctx, d, std=0.1),
           "encoder.ln_post.weight": torch.ones(d, device=dev).half(), "encoder.ln_post.bias": torch.zeros(d, device=dev).half()})
for i in range(L):
    p = f"encoder.blocks.{i}"
    for nm, (o, k) in {"attn.query": (d, d), "attn.key": (d, d), "attn.value": (d, d), "attn.out": (d, d),
                       "mlp.0": (4 * d, d), "mlp.2": (d, 4 * d)}.items():
        sd[f"{p}.{nm}.weight"] = rn(o, k, std=k ** -0.5)
        if nm != "attn.key":
            sd[f"{p}.{nm}.bias"] = rn(o, std=0.02)
    for nm in ("attn_ln", "mlp_ln"):
        sd[f"{p}.{nm}.weight"] = torch.ones(d, device=dev).half()
        sd[f"{p}.{nm}.bias"] = torch.zeros(d, device=dev).half()
scales = [0.05] * dims.n_text_layer
pipe = WhisperPipeline(dims, sd, B, scales, scales)
del sd
tk = get_tokenizer(True)  # ids only: sot / en / transcribe prompt, timestamp rules, control tokens suppressed
pipe.enable_filters(tk)
audio = (0.1 * torch.randn(B, 480000, generator=g, device=dev)).float()
audio_host = audio.cpu().pin_memory()
prompt = list(tk.sot_sequence)


def one_pass(ev):
    ev[0].record()
    a = audio_host.to(dev, non_blocking=True)
    mel = pipe.log_mel(a)
    ev[1].record()
    xa = pipe.get_audio_features(mel)
    ev[2].record()
    pipe.decoder.set_encoder_output(xa)
    ev[3].record()
    tok = pipe.decoder.decode([prompt] * B, N_NEW)
    host = tok.cpu()
    ev[4].record()
    return host


ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
one_pass(ev)
torch.cuda.synchronize()
import subprocess, threading, time
def clocks(tag):
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
                         capture_output=True, text=True).stdout.strip()
    print(tag, out, flush=True)
samples = []
stop = False
def sampler():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip()
        samples.append((time.perf_counter(), out))
th = threading.Thread(target=sampler); th.start()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(6):
    one_pass(ev)
    e[0].record()
    tok = pipe.decoder.decode([prompt] * B, N_NEW)
    e[1].record()
    tok = pipe.decoder.decode([prompt] * B, N_NEW)
    e[2].record()
    torch.cuda.synchronize()
    print(f"rep {rep}: decode in pass {ev[3].elapsed_time(ev[4]):.2f} ms, decode again {e[0].elapsed_time(e[1]):.2f} ms, and again {e[1].elapsed_time(e[2]):.2f} ms", flush=True)
stop = True; th.join()
import collections
print(collections.Counter(s for _, s in samples).most_common(12))
