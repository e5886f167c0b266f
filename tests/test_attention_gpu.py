"""GPU parity of the attention kernels through the operator API (gpt_attention / cross_attention).

The generation/context procedure follows T/tests/attention/test_gpt_attention.py (:52-56 int8-KV cases, :325-331 scale
ranges, :445-449 masked_tokens, :580-831 step loop): a context step over `in_len` tokens (half of them padding) and 7
generation steps, batch 2, 4 heads x 64, max_seq_len = in_len + 24, inputs ~1e-3, tolerances ctx 5e-3 / gen 2e-3.
The reference compares with HF GPT2Attention; here the fp32 torch attention below is that reference.
int8 cache writes are checked bit-exactly like T/cpp/tests/runtime/transposeKVKernelTest.cpp:79-147.
"""
import math

import numpy as np
import pytest
import torch

from oracle import woq

pytestmark = pytest.mark.gpu


def ref_attention(q, k, v, key_mask=None):
    """q [B,H,Sq,D], k/v [B,H,Sk,D] fp32; key_mask [B,Sq,Sk] bool (True = visible)."""
    s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None], float("-inf"))
    return torch.softmax(s, dim=-1) @ v


def split_heads(x, H):
    B, S, hidden = x.shape
    return x.view(B, S, H, hidden // H).permute(0, 2, 1, 3).float()


@pytest.mark.parametrize("int8", [True, False])
@pytest.mark.parametrize("in_len", [128, 4])
def test_context_then_generation(int8, in_len):
    from b200_whisper.functional import gpt_attention
    torch.manual_seed(42)
    B, H, D = 2, 4, 64
    hidden = H * D
    max_seq = in_len + 24
    dev = "cuda"
    scale_mag = 1e-3 if in_len == 128 else 1.0
    kv_dequant = torch.tensor([5e-4 if in_len == 128 else 0.05], dtype=torch.float32, device=dev)
    kv_quant = 1.0 / kv_dequant
    cache = torch.zeros((B, 2, H, max_seq, D), dtype=torch.int8 if int8 else torch.float16, device=dev)
    input_lengths = torch.full((B,), max(in_len // 2, 1), dtype=torch.int32, device=dev)
    if in_len == 4:
        input_lengths[:] = in_len  # Whisper prompts are not padded
    masked_tokens = torch.zeros((B, max_seq), dtype=torch.int32, device=dev)
    for i in range(B):
        masked_tokens[i, int(input_lengths[i]):in_len] = 1
    cache_ind = torch.zeros((B, 1, max_seq), dtype=torch.int32, device=dev)
    max_in = torch.zeros((in_len,), dtype=torch.int32, device=dev)
    L = int(input_lengths[0])

    # ---- context ----
    qkv = (torch.randn((B, in_len, 3 * hidden), device=dev) * scale_mag).half()
    seq_len = torch.full((B,), in_len, dtype=torch.int32, device=dev)
    out, cache = gpt_attention(qkv, cache, seq_len, torch.tensor([0, 1], dtype=torch.int32), masked_tokens,
                               input_lengths, max_in, cache_ind, H, D, 1.0, 0, False, False, False, kv_quant, kv_dequant,
                               int8)
    torch.cuda.synchronize()
    q, k, v = [split_heads(t, H) for t in qkv.float().split(hidden, dim=-1)]
    causal = torch.tril(torch.ones(in_len, in_len, dtype=torch.bool, device=dev))
    vis = causal[None] & (torch.arange(in_len, device=dev)[None, None, :] < input_lengths[:, None, None])
    ref = ref_attention(q, k, v, vis).permute(0, 2, 1, 3).reshape(B, in_len, hidden)
    err = (out.float()[:, :L] - ref[:, :L]).abs().max().item()
    assert err <= 5e-3 * max(scale_mag, 1e-3) / 1e-3 * 1e-3 + 5e-3 * scale_mag, f"context err {err}"
    # cache contents: bit-exact int8 quantization of the valid rows, zeros in the padded rows
    kc = cache[:, 0, :, :in_len].cpu()
    vc = cache[:, 1, :, :in_len].cpu()
    if int8:
        k16 = qkv[..., hidden:2 * hidden].view(B, in_len, H, D).permute(0, 2, 1, 3).cpu().numpy()
        v16 = qkv[..., 2 * hidden:].view(B, in_len, H, D).permute(0, 2, 1, 3).cpu().numpy()
        ek = woq.kv_quantize_int8(np.ascontiguousarray(k16), float(kv_quant))
        ev = woq.kv_quantize_int8(np.ascontiguousarray(v16), float(kv_quant))
        assert np.array_equal(kc.numpy()[:, :, :L], ek[:, :, :L])
        assert np.array_equal(vc.numpy()[:, :, :L], ev[:, :, :L])
        assert (kc.numpy()[:, :, L:] == 0).all()
    past_k = k.clone()
    past_v = v.clone()
    if int8:
        past_k = cache[:, 0, :, :in_len].float() * kv_dequant
        past_v = cache[:, 1, :, :in_len].float() * kv_dequant
        past_k = past_k.half().float()
        past_v = past_v.half().float()

    # ---- generation ----
    for step in range(1, 8):
        past_len = in_len + step - 1
        qkv1 = (torch.randn((B, 1, 3 * hidden), device=dev) * scale_mag).half()
        seq_len = torch.full((B,), past_len, dtype=torch.int32, device=dev)
        out, cache = gpt_attention(qkv1, cache, seq_len, torch.tensor([past_len, 0], dtype=torch.int32), masked_tokens,
                                   input_lengths, max_in, cache_ind, H, D, 1.0, 0, False, False, False, kv_quant,
                                   kv_dequant, int8)
        torch.cuda.synchronize()
        q1, k1, v1 = [split_heads(t, H) for t in qkv1.float().split(hidden, dim=-1)]
        k_all = torch.cat([past_k, k1], dim=2)
        v_all = torch.cat([past_v, v1], dim=2)
        vis = torch.ones((B, 1, past_len + 1), dtype=torch.bool, device=dev)
        vis[:, 0, :past_len] = masked_tokens[:, :past_len] == 0
        ref = ref_attention(q1, k_all, v_all, vis).permute(0, 2, 1, 3).reshape(B, 1, hidden)
        err = (out.float() - ref).abs().max().item()
        assert err <= 2e-3 * max(scale_mag / 1e-3 * 1e-3, 1e-3) + 2e-3 * scale_mag, f"generation step {step} err {err}"
        # the new token must now be in the cache (quantized), and is what later steps attend to
        if int8:
            newk = cache[:, 0, :, past_len].float() * kv_dequant
            newv = cache[:, 1, :, past_len].float() * kv_dequant
            ek = woq.kv_quantize_int8(np.ascontiguousarray(k1[:, :, 0].half().cpu().numpy()), float(kv_quant))
            assert np.array_equal(cache[:, 0, :, past_len].cpu().numpy(), ek)
            past_k = torch.cat([past_k, newk.half().float()[:, :, None]], dim=2)
            past_v = torch.cat([past_v, newv.half().float()[:, :, None]], dim=2)
        else:
            past_k, past_v = k_all, v_all


# the last three shapes have enough (row, head) pairs for the one-CTA-per-pair kernel, with fewer chunks than warps
# (16, 20, *), (10, 16, 700), (40, 4, 100), (38, 4, 333), (8, 20, 1500): a few pairs are left after the whole rounds of the grid and
# are shared by the two CTAs of a cluster (split row-head kernel); (37, 4, 1500): exactly one pair per SM, nothing to share
@pytest.mark.parametrize("B,H,S", [(1, 20, 1500), (16, 20, 1500), (3, 6, 1500), (2, 2, 96), (5, 4, 333),
                                   (10, 16, 700), (40, 4, 100), (9, 20, 1), (16, 20, 130), (38, 4, 333), (8, 20, 1500),
                                   (37, 4, 1500), (16, 20, 65)])
@pytest.mark.parametrize("int8", [True, False])
@pytest.mark.parametrize("split", [0, 3])
def test_cross_attention(B, H, S, int8, split):
    """split = 3: both cluster-sharing modes on (bit 0: the pairs left over after whole rounds of the grid are shared by the two
    CTAs of a cluster; bit 1: with fewer pairs than half the SMs EVERY pair is shared by a cluster of 2 or 4 CTAs); 0: off."""
    import b200_whisper
    from b200_whisper.functional import cross_attention, cross_kv_pack
    if split and not ((B * H >= 148 and S > 64) or (2 * B * H <= 148 and S > 192)):
        pytest.skip("neither sharing mode applies to this shape")
    prev = b200_whisper.load().b200_set_cross_attention_split(split)
    try:
        _cross_attention_case(B, H, S, int8)
    finally:
        b200_whisper.load().b200_set_cross_attention_split(prev)


def _cross_attention_case(B, H, S, int8):
    from b200_whisper.functional import cross_attention, cross_kv_pack
    torch.manual_seed(B * 131 + S)
    D = 64
    dev = "cuda"
    k = torch.randn((B, S, H * D), device=dev).half()
    v = torch.randn((B, S, H * D), device=dev).half()
    q = (torch.randn((B, H * D), device=dev) * 1.5).half()
    t = float(max(k.abs().max(), v.abs().max())) / 127.0
    oq = torch.tensor([1.0 / t], dtype=torch.float32, device=dev)
    qo = torch.tensor([t], dtype=torch.float32, device=dev)
    cache = cross_kv_pack(k, v, oq, H, D, int8)
    if int8:
        ek = woq.kv_quantize_int8(k.view(B, S, H, D).permute(0, 2, 1, 3).contiguous().cpu().numpy(), float(oq))
        # the int8 cross cache is kept in offset-binary form (stored byte = q + 128)
        vals = (cache.view(torch.uint8) ^ 0x80).view(torch.int8)
        assert np.array_equal(vals[:, 0].cpu().numpy(), ek)
        kd = (vals[:, 0].float() * qo).half().float()
        vd = (vals[:, 1].float() * qo).half().float()
    else:
        kd, vd = cache[:, 0].float(), cache[:, 1].float()
    out = cross_attention(q, cache, qo, H, D, int8)
    torch.cuda.synchronize()
    ref = ref_attention(q.float().view(B, H, 1, D), kd, vd).reshape(B, H * D)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"cross attention err {err}"


def test_cross_attention_multi_query_rows():
    """Context phase: S_q prompt rows per sequence attend to their sequence's cache."""
    from b200_whisper.functional import cross_attention, cross_kv_pack
    torch.manual_seed(5)
    _multi_query_rows(3, 4, 200, 4)
    _multi_query_rows(6, 8, 130, 5)  # 240 (row, head) pairs: one-CTA-per-pair kernel
    import b200_whisper
    prev = b200_whisper.load().b200_set_cross_attention_split(3)
    try:
        _multi_query_rows(3, 10, 200, 5)  # 150 pairs: two of them shared by the CTAs of a cluster
        _multi_query_rows(2, 4, 1500, 3)  # 24 pairs: every pair shared by a cluster of four CTAs
    finally:
        b200_whisper.load().b200_set_cross_attention_split(prev)


def _multi_query_rows(B, H, S, Sq):
    from b200_whisper.functional import cross_attention, cross_kv_pack
    D = 64
    dev = "cuda"
    k = torch.randn((B, S, H * D), device=dev).half()
    v = torch.randn((B, S, H * D), device=dev).half()
    q = torch.randn((B * Sq, H * D), device=dev).half()
    one = torch.ones((1,), dtype=torch.float32, device=dev)
    cache = cross_kv_pack(k, v, one, H, D, False)
    out = cross_attention(q, cache, one, H, D, False)
    torch.cuda.synchronize()
    qq = q.float().view(B, Sq, H, D).permute(0, 2, 1, 3)
    ref = ref_attention(qq, cache[:, 0].float(), cache[:, 1].float()).permute(0, 2, 1, 3).reshape(B * Sq, H * D)
    assert (out.float() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


def test_attention_argument_errors():
    from b200_whisper.functional import gpt_attention
    dev = "cuda"
    qkv = torch.zeros((1, 1, 3 * 64), dtype=torch.float16, device=dev)
    cache = torch.zeros((1, 2, 1, 8, 64), dtype=torch.float16, device=dev)
    z = torch.zeros((1,), dtype=torch.int32, device=dev)
    ci = torch.zeros((1, 1, 8), dtype=torch.int32, device=dev)
    with pytest.raises(ValueError):  # past_key_value_length on the GPU: the reference's example bug
        gpt_attention(qkv, cache, z, torch.tensor([0, 1], dtype=torch.int32, device=dev), None, z, z, ci, 1, 64, 1.0, 0,
                      False, False, False, None, None, False)
    with pytest.raises(RuntimeError):  # past length beyond the cache capacity
        gpt_attention(qkv, cache, z, torch.tensor([8, 0], dtype=torch.int32), None, z, z, ci, 1, 64, 1.0, 0, False,
                      False, False, None, None, False)
    with pytest.raises(NotImplementedError):
        gpt_attention(qkv, cache, z, torch.tensor([0, 1], dtype=torch.int32), None, z, z, ci, 1, 64, 1.0, 32, False,
                      False, False, None, None, False)


@pytest.mark.parametrize("B,S,H", [(2, 1500, 20), (1, 100, 3), (2, 64, 1), (1, 1, 2), (3, 77, 6), (1, 129, 4)])
def test_bidirectional_attention(B, S, H):
    """Encoder self-attention (no mask, no cache) against W/torch_model.py:88-103 evaluated in fp32."""
    from b200_whisper.functional import bidirectional_attention
    torch.manual_seed(B * 7 + S)
    D = 64
    qkv = (torch.randn((B, S, 3 * H * D), device="cuda") * 1.2).half()
    out = bidirectional_attention(qkv, H, D)
    torch.cuda.synchronize()
    q, k, v = [t.float().view(B, S, H, D).permute(0, 2, 1, 3) for t in qkv.split(H * D, dim=-1)]
    w = torch.softmax((q * D ** -0.25) @ (k * D ** -0.25).transpose(-1, -2), dim=-1)
    ref = (w @ v).permute(0, 2, 1, 3).reshape(B, S, H * D)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"bidirectional attention err {err}"


@pytest.mark.parametrize("int8", [True, False])
@pytest.mark.parametrize("in_len,tokens_per_block", [(4, 16), (100, 64), (37, 8)])
def test_paged_kv_cache_equals_linear(int8, in_len, tokens_per_block):
    """Paged KV cache (KVBlockArray, K/kvCacheUtils.h:34-112; kv_cache_block_pointers of gpt_attention): a context step and 9
    generation steps over a block pool in shuffled order give bit-identical outputs and, gathered back through the block
    table, a bit-identical cache compared with the contiguous KVLinearBuffer path.  Blocks a sequence has not reached
    yet all point at one poisoned guard block that must stay untouched."""
    from b200_whisper.functional import gpt_attention
    torch.manual_seed(7 + in_len)
    B, H, D = 3, 4, 64
    hidden = H * D
    n_gen = 9
    max_seq = in_len + 24
    dev = "cuda"
    dt = torch.int8 if int8 else torch.float16
    kv_dequant = torch.tensor([0.05], dtype=torch.float32, device=dev)
    kv_quant = 1.0 / kv_dequant
    max_blocks = (max_seq + tokens_per_block - 1) // tokens_per_block
    used_blocks = (in_len + n_gen + tokens_per_block - 1) // tokens_per_block      # blocks a sequence ever writes
    n_pool = B * 2 * used_blocks + 1
    pool = torch.zeros((n_pool, 2, H, tokens_per_block, D), dtype=dt, device=dev)   # [blocks, 2, H, tpb, Dh]
    guard = n_pool - 1
    pool[guard] = 77
    esz = pool.element_size()
    block_bytes = 2 * H * tokens_per_block * D * esz
    perm = torch.randperm(n_pool - 1).tolist()
    table = torch.full((B, 1, 2, max_blocks), guard, dtype=torch.int64)
    it = iter(perm)
    for b in range(B):
        for kv in range(2):
            for j in range(used_blocks):
                table[b, 0, kv, j] = next(it)
    block_id = table.clone()
    pointers = (pool.data_ptr() + block_id * block_bytes).to(dev)
    linear = torch.zeros((B, 2, H, max_seq, D), dtype=dt, device=dev)
    input_lengths = torch.full((B,), in_len, dtype=torch.int32, device=dev)
    masked_tokens = torch.zeros((B, max_seq), dtype=torch.int32, device=dev)
    cache_ind = torch.zeros((B, 1, max_seq), dtype=torch.int32, device=dev)
    max_in = torch.zeros((in_len,), dtype=torch.int32, device=dev)

    def both(qkv, seq_len, host):
        a, _ = gpt_attention(qkv, linear, seq_len, host, masked_tokens, input_lengths, max_in, cache_ind, H, D, 1.0, 0,
                             False, False, False, kv_quant, kv_dequant, int8)
        b, _ = gpt_attention(qkv, pool, seq_len, host, masked_tokens, input_lengths, max_in, cache_ind, H, D, 1.0, 0,
                             False, False, False, kv_quant, kv_dequant, int8, kv_cache_block_pointers=pointers)
        torch.cuda.synchronize()
        assert torch.equal(a, b)

    def gathered(n_tokens):
        out = torch.zeros((B, 2, H, n_tokens, D), dtype=dt, device=dev)
        for b in range(B):
            for kv in range(2):
                for t in range(n_tokens):
                    blk = int(block_id[b, 0, kv, t // tokens_per_block])
                    out[b, kv, :, t] = pool[blk, 0, :, t % tokens_per_block]
        return out

    qkv = torch.randn((B, in_len, 3 * hidden), device=dev).half()
    both(qkv, torch.full((B,), in_len, dtype=torch.int32, device=dev), torch.tensor([0, 1], dtype=torch.int32))
    assert torch.equal(gathered(in_len), linear[:, :, :, :in_len])
    for step in range(n_gen):
        past_len = in_len + step
        qkv1 = torch.randn((B, 1, 3 * hidden), device=dev).half()
        both(qkv1, torch.full((B,), past_len, dtype=torch.int32, device=dev), torch.tensor([past_len, 0], dtype=torch.int32))
    n = in_len + n_gen
    assert torch.equal(gathered(n), linear[:, :, :, :n])
    assert (pool[guard] == 77).all()                     # nothing was written through the not-yet-allocated entries
    assert (pool[:guard, 1] == 0).all()                  # the tables point at the first half of every pool slot only
