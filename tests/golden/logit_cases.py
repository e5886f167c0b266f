"""Deterministic synthetic logits shared by make_logit_filter_golden.py and the tests (pure integer hashing, so the
values are identical on every platform) and the roll-out cases."""
import numpy as np

# multilingual Whisper vocabulary layout (W/tokenizer.py): eot 50257, sot 50258, ..., no_timestamps 50363, <|0.00|> 50364
V, EOT, NO_TS, TS_BEGIN, BLANK = 51865, 50257, 50363, 50364, 220
PROMPT = [50258, 50259, 50359]
SUPPRESS = sorted({1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522,
                   542, 873, 893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961,
                   4183, 4667, 6585, 6647, 7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157,
                   14635, 15265, 15618, 16553, 16604, 18362, 18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279,
                   29464, 31650, 32302, 32470, 36865, 42863, 47425, 49870, 50254, 50258, 50358, 50359, 50360, 50361,
                   50362})
# (seed, steps, batch, max_initial_timestamp_index or -1, boost pattern)
CASES = [(0, 10, 3, 50, ["ts", "text", "text", "ts", "ts", "text", "eot"]),
         (1, 8, 2, -1, ["spread", "text", "ts", "ts"]),
         (2, 8, 2, 50, ["text", "ts", "eot", "text"]),
         (3, 12, 4, 50, ["ts", "ts", "text", "spread", "text", "ts", "text", "text", "eot"])]


def _mix(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(33)
    x = (x * np.uint64(0xff51afd7ed558ccd)) & np.uint64(0xffffffffffffffff)
    x ^= x >> np.uint64(33)
    x = (x * np.uint64(0xc4ceb9fe1a85ec53)) & np.uint64(0xffffffffffffffff)
    x ^= x >> np.uint64(33)
    return x


def synth_logits(seed, step, batch, boost):
    """float32 [batch, V], every value fp16-representable."""
    v = np.arange(V, dtype=np.uint64)[None, :]
    b = np.arange(batch, dtype=np.uint64)[:, None]
    base = v * np.uint64(2654435761) + b * np.uint64(40503) + np.uint64(step * 977 + seed * 1000003 + 12345)
    acc = np.zeros((batch, V), np.float64)
    for r in range(4):
        h = _mix(base + np.uint64(r * 0x9e3779b9))
        acc += (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    x = (acc - 2.0) * 3.4  # roughly N(0, 2)
    for bb in range(batch):
        mode = boost[(step + bb) % len(boost)]
        if mode == "ts":
            x[bb, TS_BEGIN:TS_BEGIN + 400] += 6.0
        elif mode == "eot":
            x[bb, EOT] += 14.0
        elif mode == "spread":
            x[bb, TS_BEGIN:] += 2.5  # no single timestamp wins but their mass does
    return x.astype(np.float16).astype(np.float32)
