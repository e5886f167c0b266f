"""Generates tests/golden/quant_golden.npz and quant_digests.json from the REFERENCE's own
cutlass_preprocessors.cpp (compiled in place into oracle/_ref/libref_quant.so by oracle/Makefile).

Run in the build container (needs /root/reference):  python tests/golden/make_quant_golden.py

Inputs follow the reference tests' generator, T/tests/quantization/_utils.py:15-22
(torch.manual_seed(0); torch.rand(shape, fp16) * 2 - 1), so they can be re-created anywhere.
Small shapes are stored in full; the reference tests' own shapes
(T/tests/quantization/test_weight_only_quant_matmul.py:112-130) and the Whisper large-v2 shapes are
stored as sha256 digests of (raw, processed, scales).
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import woq  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def gen_weight(k, n, dtype=torch.float16, seed=0):
    torch.manual_seed(seed)
    return (torch.rand((k, n), dtype=dtype) * 2 - 1.0).numpy()


def digest(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    full = {}
    for (k, n) in [(64, 64), (128, 64), (64, 128), (192, 256)]:
        for dt, tag in [(torch.float16, "f16"), (torch.float32, "f32")]:
            w = gen_weight(k, n, dt)
            raw, proc, scales = woq.ref_symmetric_quantize_int8(w)
            key = f"k{k}_n{n}_{tag}"
            full[key + "_w"] = w
            full[key + "_raw"] = raw
            full[key + "_proc"] = proc
            full[key + "_scales"] = scales
    np.savez_compressed(os.path.join(HERE, "quant_golden.npz"), **full)

    digests = {}
    # (k, n): reference test shapes are quoted (n, k) there; W is [k, n] here.
    shapes = [(4096, 1024), (512, 4096), (12288, 6144), (1280, 3840), (1280, 1280), (1280, 5120), (5120, 1280),
              (384, 1152), (384, 1536), (1536, 384)]
    for (k, n) in shapes:
        w = gen_weight(k, n)
        raw, proc, scales = woq.ref_symmetric_quantize_int8(w)
        digests[f"k{k}_n{n}_f16"] = {"raw": digest(raw), "proc": digest(proc), "scales": digest(scales),
                                     "input": "torch.manual_seed(0); torch.rand((k,n),fp16)*2-1"}
    with open(os.path.join(HERE, "quant_digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)
    print("wrote", len(full) // 4, "full vectors and", len(digests), "digests")


if __name__ == "__main__":
    main()
