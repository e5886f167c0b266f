"""CPU checks of the persistent decoder-step entry point's host side: the A-fragment activation layout is a bijection
that puts every mma.sync.m16n8k16 A operand of a warp in one contiguous 512-byte line, the scratch size formula, and
argument validation before any device work (no GPU needed)."""
import ctypes

import numpy as np


def frag_index(row, k):
    """Mirror of frag_index() in csrc/decoder_step.cu."""
    kb, kk = k >> 6, k & 63
    T, r = kk >> 4, kk & 15
    hi, w, e = r >> 3, (r & 7) >> 1, r & 1
    g, up = row & 7, row >> 3
    return ((((kb * 4 + w) * 32 + 4 * g + T) * 4 + 2 * hi + up) << 1) + e


def test_fragment_layout_is_a_bijection_and_matches_mma_operands():
    K = 1280
    idx = np.array([[frag_index(r, k) for k in range(K)] for r in range(16)])
    assert sorted(idx.ravel().tolist()) == list(range(16 * K))
    # lane (g, T), MMA w of k-block kb reads 8 consecutive halves = registers a0..a3 of mma.sync.m16n8k16:
    # a0 = (row g, k 2T', 2T'+1), a1 = (row g+8, same), a2 = (row g, +8), a3 = (row g+8, +8) with the physical
    # k = 64 kb + 16 T + 2 w (+1) (+8): the k's the weight word w of 16-byte chunk T dequantizes to (common.cuh)
    for kb in (0, 7, 19):
        for w in range(4):
            for g in range(8):
                for T in range(4):
                    base = ((kb * 4 + w) * 32 + 4 * g + T) * 8
                    k0 = 64 * kb + 16 * T + 2 * w
                    want = [(g, k0), (g, k0 + 1), (g + 8, k0), (g + 8, k0 + 1),
                            (g, k0 + 8), (g, k0 + 9), (g + 8, k0 + 8), (g + 8, k0 + 9)]
                    assert [int(idx[r, k]) for r, k in want] == list(range(base, base + 8))


def test_scratch_size_and_argument_validation():
    import b200_whisper
    from b200_whisper import _lib
    lib = b200_whisper.load()
    assert lib.b200_decoder_step_scratch_bytes(20, 5120) == 256 + 192 * 1280 + 32 * 5120
    assert lib.b200_decoder_step_scratch_bytes(0, 5120) == 0
    assert lib.b200_decoder_step(None, None) == 1
    p = _lib.DecoderStepParams()
    assert lib.b200_decoder_step(ctypes.byref(p), None) == 1  # null pointers
    buf = np.zeros(4096, dtype=np.uint8)
    a = buf.ctypes.data + (-buf.ctypes.data) % 256
    for f in ("layers", "tokens", "sequence_lengths", "tok_emb", "pos_emb", "x_out", "scratch"):
        setattr(p, f, a)
    p.n_layers, p.batch_size, p.num_heads, p.d_ff = 2, 17, 20, 5120
    p.max_seq_len, p.enc_len, p.vocab, p.n_ctx = 448, 1500, 51865, 448
    assert lib.b200_decoder_step(ctypes.byref(p), None) == 2 and b"batch_size" in lib.b200_last_error()
    p.batch_size, p.num_heads = 16, 24  # d = 1536 > 1280
    assert lib.b200_decoder_step(ctypes.byref(p), None) == 2
    assert ctypes.sizeof(_lib.DecoderLayer) == 32 * 8
