"""Golden vectors for the Whisper logit filters + greedy update, produced by the REFERENCE's own
classes: ApplyTimestampRules, SuppressBlank, SuppressTokens, GreedyDecoder of T/examples/whisper/decoding.py:134-300
(imported from /root/reference with its TensorRT imports stubbed).  Each case is a short greedy roll-out over the
deterministic synthetic logits of logit_cases.py; stored: the chosen tokens and the running sum_logprobs of every step.

    python tests/golden/make_logit_filter_golden.py      (needs /root/reference; writes logit_filter_golden.npz)
"""
import os
import sys
import types

import numpy as np
import torch

W = "/root/reference/tensorrt_llm_july-release-v1/examples/whisper"
for name in ("tensorrt_llm", "tensorrt_llm.runtime", "tensorrt_llm.runtime.session", "tensorrt_llm.logger",
             "tensorrt_llm._utils", "build"):
    m = types.ModuleType(name)
    for attr in ("ModelConfig", "SamplingConfig", "Session", "TensorInfo", "str_dtype_to_torch", "str_dtype_to_trt",
                 "trt_dtype_to_torch", "get_engine_name"):
        setattr(m, attr, object)
    sys.modules[name] = m
sys.modules["tensorrt_llm"].runtime = sys.modules["tensorrt_llm.runtime"]
sys.modules["tensorrt_llm"].logger = sys.modules["tensorrt_llm.logger"]
sys.path.insert(0, W)
sys.dont_write_bytecode = True
import decoding as ref  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from logit_cases import BLANK, CASES, EOT, NO_TS, PROMPT, SUPPRESS, TS_BEGIN, V, synth_logits  # noqa: E402


class Tok:
    eot, no_timestamps, timestamp_begin = EOT, NO_TS, TS_BEGIN

    @staticmethod
    def encode(s):
        assert s == " "
        return [BLANK]


def main():
    out = {}
    for i, (seed, steps, batch, mi, boost) in enumerate(CASES):
        tokens = torch.tensor([PROMPT] * batch)
        sb = len(PROMPT)
        filters = [ref.SuppressBlank(Tok, sb), ref.SuppressTokens(SUPPRESS),
                   ref.ApplyTimestampRules(Tok, sb, None if mi < 0 else mi)]
        dec = ref.GreedyDecoder(0.0, EOT)
        sum_lp = torch.zeros(batch)
        toks, lps = [], []
        for t in range(steps):
            work = torch.from_numpy(synth_logits(seed, t, batch, boost)).clone()
            for f in filters:
                f.apply(work, tokens)
            tokens, _ = dec.update(tokens, work, sum_lp)
            toks.append(tokens[:, -1].clone())
            lps.append(sum_lp.clone())
        out[f"c{i}_tokens"] = torch.stack(toks).numpy().astype(np.int32)
        out[f"c{i}_sumlp"] = torch.stack(lps).numpy().astype(np.float32)
        print(i, out[f"c{i}_tokens"].T.tolist())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "logit_filter_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
