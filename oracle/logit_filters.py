"""TEST INFRASTRUCTURE: CPU restatement (numpy, float32 like the reference) of the Whisper logit filters and the greedy
token update, following T/examples/whisper/decoding.py

    SuppressBlank        :202-209     logits[:, encode(" ") + [eot]] = -inf at the first sampled position
    SuppressTokens       :212-217     logits[:, suppress] = -inf
    ApplyTimestampRules  :134-199     <|notimestamps|> off; timestamps come in pairs; timestamps never decrease; first
                                      token is a timestamp <= max_initial; if the timestamp probability mass beats every
                                      text token, text is masked
    GreedyDecoder.update :274-293     argmax, sum_logprobs += logprob(next) unless the sequence already ended, and a
                                      sequence that produced eot keeps producing eot

Pinned by tests/golden/logit_filter_golden.npz (roll-outs through the reference's own classes,
tests/golden/make_logit_filter_golden.py).  State per sequence is what the device kernel keeps: number of sampled
tokens, last and penultimate sampled token, last timestamp token."""
import numpy as np


class FilterConfig:
    def __init__(self, eot, no_timestamps, timestamp_begin, blank, suppress, max_initial_timestamp_index=None):
        self.eot, self.no_timestamps, self.timestamp_begin, self.blank = eot, no_timestamps, timestamp_begin, blank
        self.suppress = np.asarray(sorted(set(suppress)), np.int64)
        self.max_initial = max_initial_timestamp_index


def _logsumexp(x):
    m = np.max(x)
    if not np.isfinite(m):
        return np.float32(-np.inf)
    return np.float32(m + np.log(np.sum(np.exp((x - m).astype(np.float32)), dtype=np.float32)))


def filter_logits(logits, sampled, cfg):
    """logits [V] float32 (modified copy returned); sampled: list of tokens generated so far for this sequence."""
    lg = logits.astype(np.float32).copy()
    ts, eot = cfg.timestamp_begin, cfg.eot
    n = len(sampled)
    if n == 0:  # SuppressBlank
        lg[cfg.blank] = -np.inf
        lg[eot] = -np.inf
    lg[cfg.suppress] = -np.inf  # SuppressTokens
    if cfg.no_timestamps is not None:
        lg[cfg.no_timestamps] = -np.inf
    last_was_ts = n >= 1 and sampled[-1] >= ts
    penult_was_ts = n < 2 or sampled[-2] >= ts
    if last_was_ts:
        if penult_was_ts:
            lg[ts:] = -np.inf
        else:
            lg[:eot] = -np.inf
    stamps = [t for t in sampled if t >= ts]
    if stamps:
        last = stamps[-1] if (last_was_ts and not penult_was_ts) else stamps[-1] + 1
        lg[ts:last] = -np.inf
    if n == 0:
        lg[:ts] = -np.inf
        if cfg.max_initial is not None:
            lg[ts + cfg.max_initial + 1:] = -np.inf
    # probability mass rule (the shared normaliser of log_softmax cancels)
    if _logsumexp(lg[ts:]) > np.max(lg[:ts]):
        lg[:ts] = -np.inf
    return lg


def greedy_step(logits, sampled, last_context_token, sum_logprob, cfg):
    """One GreedyDecoder.update for one sequence.  Returns (next_token, new_sum_logprob)."""
    lg = filter_logits(logits, sampled, cfg)
    nxt = int(np.argmax(lg))
    logprob = np.float32(lg[nxt] - _logsumexp(lg))
    ended = last_context_token == cfg.eot
    if not ended:
        sum_logprob = np.float32(sum_logprob + logprob)
    else:
        nxt = cfg.eot
    return nxt, sum_logprob
