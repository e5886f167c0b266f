"""Generates tests/golden/whisper_micro_golden.npz from the REFERENCE PyTorch Whisper
(T/examples/whisper/torch_model.py, imported from /root/reference -- build container only).

The model is the MICRO configuration of oracle/whisper_oracle.py with oracle.synthetic_state_dict(seed=1); inputs are
a seeded synthetic log-mel.  Stored: encoder output checksum + sample, decoder logits for the prompt, and the greedy
token ids of a 12-token decode with the reference's own hook-based KV cache (torch_model.py:270-301), logit filters off.
Run:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_whisper_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/tensorrt_llm_july-release-v1/examples/whisper")

from oracle import whisper_oracle as wo  # noqa: E402


def main():
    from torch_model import ModelDimensions, Whisper  # the reference
    dims = wo.MICRO
    sd = wo.synthetic_state_dict(dims, seed=1)
    model = Whisper(ModelDimensions(**dims.__dict__)).float().eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("mask" in m or "alignment" in m for m in missing), (missing, unexpected)
    torch.manual_seed(1)
    B = 2
    mel = torch.randn(B, dims.n_mels, 2 * dims.n_audio_ctx).clamp(-1, 1).half().float()
    prompt = [3, 7, 11]
    n_new = 12
    with torch.no_grad():
        xa = model.encoder(mel)
        tokens = torch.tensor(prompt).repeat(B, 1)
        prompt_logits = model.decoder(tokens, xa)
        cache, hooks = model.install_kv_cache_hooks()
        cur = tokens
        out = []
        step_logits = []
        for _ in range(n_new):
            lg = model.decoder(cur, xa, kv_cache=cache)[:, -1]
            nxt = lg.argmax(-1)
            out.append(nxt)
            step_logits.append(lg)
            cur = nxt[:, None]
        for h in hooks:
            h.remove()
    np.savez_compressed(
        os.path.join(HERE, "whisper_micro_golden.npz"),
        mel=mel.numpy().astype(np.float16),  # values are stored in fp16; the test feeds exactly these
        xa_sample=xa[:, ::16, ::16].numpy(), xa_sum=np.float64(xa.double().sum().item()),
        prompt=np.array(prompt), prompt_logits_sample=prompt_logits[:, :, ::8].numpy(),
        tokens=torch.stack(out, 1).numpy(), step_logits_sample=torch.stack(step_logits, 1)[:, :, ::8].numpy())
    print("tokens", torch.stack(out, 1).tolist())


if __name__ == "__main__":
    main()
