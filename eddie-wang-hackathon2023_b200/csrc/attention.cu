// attention.cu -- masked multi-head attention with int8 (or fp16) KV cache: generation step, context (prompt)
// phase with cache fill, and the cached cross-attention over the encoder frames.
//
// Reference semantics followed (T/ = /root/reference/tensorrt_llm_july-release-v1/):
//   generation  T/cpp/tensorrt_llm/kernels/decoderMaskedMultiheadAttention/decoderMaskedMultiheadAttentionTemplate.h:1195-2017
//     - the new K/V are quantized and stored, but THIS step's q.k and p.v use the unquantized values (:1503,1517,1920,1933)
//     - cached values dequantize as half(scale_quant_orig * float(int8))  (decoderMaskedMultiheadAttentionUtils.h:2357-2365)
//     - stores quantize as cvt.rni.sat.s8.f32(scale_orig_quant * float(x))  (Utils.h:2276-2286,2382-2390)
//     - logits = q.k * inv_sqrt_dh in fp32; masked keys get probability 0 and do not enter the max (:1678-1680,1730)
//     - probabilities are exp(x - max) / (sum + 1e-6)  (:1756)
//   context     T/cpp/tensorrt_llm/plugins/gptAttentionCommon/gptAttentionCommon.cpp:361-620,
//               softmax T/cpp/tensorrt_llm/kernels/unfusedAttentionKernels.cu:179-257 (mask adds -10000, 1e-6 in the sum),
//               cache fill unfusedAttentionKernels.cu:1552-1646
//   cross       T/tensorrt_llm/layers/attention.py:308-323,385-406; oracle T/examples/whisper/torch_model.py:88-103
// Like the reference (Template.h:1765) the probabilities of cached keys are rounded to fp16 before P.V (HFMA2 chains of
// at most 8 keys, flushed to fp32); the current token's term and all sums stay fp32.  Pinned against the reference
// kernel itself running on B200 (tests/test_reference_kernels_gpu.py): identical cache bytes, outputs within 2e-3.
//
// Cache layout: [B, 2, H, Smax, Dh] (KVLinearBuffer, T/cpp/tensorrt_llm/kernels/kvCacheUtils.h:114-170). Dh = 64.
#include <float.h>
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "attn_device.cuh"

namespace b200
{


// 16 consecutive cache elements -> float[16], in the reference's dequant arithmetic.
template <bool INT8>
__device__ __forceinline__ void load16(const void* base, size_t elem_off, float scale_quant_orig, float (&out)[16])
{
    if constexpr (INT8)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(static_cast<const int8_t*>(base) + elem_off);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
#pragma unroll
            for (int b = 0; b < 4; ++b)
            {
                const int q = static_cast<int8_t>((w[i] >> (8 * b)) & 0xff);
                out[4 * i + b] = __half2float(__float2half_rn(scale_quant_orig * static_cast<float>(q)));
            }
        }
    }
    else
    {
        const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __half*>(base) + elem_off);
        const uint4 v0 = p[0], v1 = p[1];
        const __half2* h0 = reinterpret_cast<const __half2*>(&v0);
        const __half2* h1 = reinterpret_cast<const __half2*>(&v1);
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const float2 a = __half22float2(h0[i]);
            const float2 b = __half22float2(h1[i]);
            out[2 * i] = a.x;
            out[2 * i + 1] = a.y;
            out[8 + 2 * i] = b.x;
            out[8 + 2 * i + 1] = b.y;
        }
    }
}

// 16 consecutive cache elements -> float[16] WITHOUT the dequant scale (int8: the integer values; fp16: the values).
// int8 path: xor 0x80 turns two's complement into biased bytes, then the 0x6400|b trick gives exact fp16 integers.
// Output order is a fixed permutation of the 16 dims (the caller permutes q / o the same way): out[4i..4i+3] =
// dims 4i+{0, 2, 1, 3}.
template <bool INT8>
__device__ __forceinline__ void load16_raw_perm(const void* base, size_t elem_off, float (&out)[16])
{
    if constexpr (INT8)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(static_cast<const int8_t*>(base) + elem_off);
        const uint32_t w[4] = {v.x ^ 0x80808080u, v.y ^ 0x80808080u, v.z ^ 0x80808080u, v.w ^ 0x80808080u};
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            __half2 lo, hi;
            dequant_word(w[i], lo, hi); // lo = (b0, b2), hi = (b1, b3)
            const float2 a = __half22float2(lo);
            const float2 b = __half22float2(hi);
            out[4 * i + 0] = a.x;
            out[4 * i + 1] = a.y;
            out[4 * i + 2] = b.x;
            out[4 * i + 3] = b.y;
        }
    }
    else
    {
        float t[16];
        load16<false>(base, elem_off, 1.0f, t);
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            out[4 * i + 0] = t[4 * i + 0];
            out[4 * i + 1] = t[4 * i + 2];
            out[4 * i + 2] = t[4 * i + 1];
            out[4 * i + 3] = t[4 * i + 3];
        }
    }
}


__device__ __forceinline__ float block_reduce_max(float v, float* red, int nwarps)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0)
        red[warp] = v;
    __syncthreads();
    v = (lane < nwarps) ? red[lane] : -FLT_MAX;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    return v;
}

__device__ __forceinline__ float block_reduce_sum(float v, float* red, int nwarps)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0)
        red[warp] = v;
    __syncthreads();
    v = (lane < nwarps) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    return v;
}


// =====================================================================================================
// Generation step.  A PAIR of warps per (batch, head), no block-wide barrier except the final merge.  Lane geometry:
// 4 lanes x 16 dims per key, 8 keys per warp instruction (512 contiguous cache bytes for int8); the two warps of a
// pair take alternating groups of 8 keys, so a pass of the pair covers 16 * NIT keys.  Under programmatic dependent
// launch everything that does not depend on this step's projection happens BEFORE griddepcontrol.wait when the
// caller promised a static cache (b200_set_static_kv_hint): the first pass of K and V is fetched and converted to
// fp16 registers, the length and the scales are read.  After the wait only q.k, the softmax, p.v and the merge of
// the two halves (shared memory) remain.  64 fp16 registers of K/V per thread keep the CTA small enough to share an
// SM with a GEMM CTA of the preceding projection.
// =====================================================================================================
// Paged KV cache (KVBlockArray, K/kvCacheUtils.h:34-112): a table [B, 2, max_blocks_per_seq] of pointers to blocks laid
// out [H, tokens_per_block, Dh] (getKVLocalIdx :104-112); tokens_per_block is a power of two.  PAGED = false keeps the
// contiguous KVLinearBuffer addressing and compiles to the same code as before the paged variant existed.
struct KvPaged
{
    const void* const* table;
    int max_blocks_per_seq;
    int tokens_per_block_log2;
};

// start of the Dh-element row of token `key` of (sequence b, K or V, head h)
template <bool PAGED>
__device__ __forceinline__ char* kv_row(char* linear_base, const KvPaged& pg, int b, int kv, int h, int key, size_t esz)
{
    if constexpr (!PAGED)
    {
        return linear_base + (size_t) key * kDh * esz;
    }
    else
    {
        const unsigned long long blk = __ldg(reinterpret_cast<const unsigned long long*>(pg.table)
            + ((size_t) (b * 2 + kv)) * pg.max_blocks_per_seq + (key >> pg.tokens_per_block_log2));
        const int local = key & ((1 << pg.tokens_per_block_log2) - 1);
        return reinterpret_cast<char*>(blk) + ((size_t) ((h << pg.tokens_per_block_log2) + local) * kDh) * esz;
    }
}

constexpr int kMmhaWarps = 4; // 2 (batch, head) pairs per CTA

template <bool INT8, bool PAGED>
__global__ void __launch_bounds__(kMmhaWarps * 32, 2) mmha_generation_kernel(const b200_mmha_params p, const int early_kv,
    const KvPaged pg)
{
    constexpr int NIT = INT8 ? 4 : 2; // key groups of 8 per warp and pass
    constexpr int kPart = kDh + 4;    // m, l, 2 pad, o[64]
    __shared__ __align__(16) float parts[kMmhaWarps / 2][kPart];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = lane & 3, kl = lane >> 2;
    const int half = warp & 1, pair = warp >> 1;
    const int H = p.num_heads, Smax = p.max_seq_len;
    const int hidden = H * kDh;
    const int gp = blockIdx.x * (kMmhaWarps / 2) + pair;
    const bool active = gp < p.batch_size * H;
    const int b = active ? gp / H : 0, h = active ? gp - b * H : 0;
    const size_t esz = INT8 ? 1 : 2;
    char* kc = static_cast<char*>(p.kv_cache) + ((size_t) (b * 2 + 0) * H + h) * Smax * kDh * esz; // unused when PAGED
    char* vc = static_cast<char*>(p.kv_cache) + ((size_t) (b * 2 + 1) * H + h) * Smax * kDh * esz;

    // keys past the end are fetched (and ignored) up to a cap: the whole linear buffer exists, but only the blocks up
    // to the token being appended are guaranteed to be allocated in a paged cache
    int paged_cap = 0;
    __half2 kw[NIT][8], vw[NIT][8];
    auto fetch = [&](int k0)
    {
        KvChunk<INT8> kreg[NIT], vreg[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            const int key = min(k0 + (2 * it + half) * 8 + kl, PAGED ? paged_cap : Smax - 1);
            if constexpr (PAGED)
            {
                kreg[it].load(kv_row<true>(kc, pg, b, 0, h, key, esz), (size_t) chunk * 16);
                vreg[it].load(kv_row<true>(vc, pg, b, 1, h, key, esz), (size_t) chunk * 16);
            }
            else
            {
                kreg[it].load(kc, (size_t) key * kDh + chunk * 16);
                vreg[it].load(vc, (size_t) key * kDh + chunk * 16);
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            kreg[it].unpack(kw[it]);
            vreg[it].unpack(vw[it]);
        }
    };
    grid_dep_launch_dependents();
    if (!early_kv)
        grid_dep_wait();
    if constexpr (PAGED)
        paged_cap = min(p.sequence_lengths ? p.sequence_lengths[b] : p.past_kv_length, Smax - 1);
    fetch(0);
    int tlen = p.sequence_lengths ? p.sequence_lengths[b] : p.past_kv_length;
    tlen = min(tlen, Smax - 1);
    const float inv_sqrt_dh = 1.f / (sqrtf((float) kDh) * p.q_scaling); // gptAttentionCommon.cpp:163
    const float s_qo = INT8 ? p.kv_scale_quant_orig[0] : 1.f;
    const float s_oq = INT8 ? p.kv_scale_orig_quant[0] : 1.f;
    const float sscale = s_qo * inv_sqrt_dh;
    if (early_kv)
        grid_dep_wait(); // qkv comes from the previous kernel
    const int* mask = p.masked_tokens ? p.masked_tokens + (size_t) b * Smax : nullptr;

    const __half* qkv = static_cast<const __half*>(p.qkv) + (size_t) b * 3 * hidden + h * kDh + chunk * 16;
    const __half* bias = p.qkv_bias ? static_cast<const __half*>(p.qkv_bias) + h * kDh + chunk * 16 : nullptr;
    __half qh[16], kh[16], vh[16];
    load16_half(qkv, bias, qh);
    if (half == 0)
    {
        load16_half(qkv + hidden, bias ? bias + hidden : nullptr, kh);
        load16_half(qkv + 2 * hidden, bias ? bias + 2 * hidden : nullptr, vh);
    }
    __half2 q2[8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        q2[2 * i] = __halves2half2(qh[4 * i], qh[4 * i + 2]);
        q2[2 * i + 1] = __halves2half2(qh[4 * i + 1], qh[4 * i + 3]);
    }
    // warp 0 of the pair: append this step's K and V (lane group 0 writes K, group 1 writes V; 16 dims per lane) and
    // score the current token with its unquantized k (Template.h:1503,1517,1920,1933)
    float s_cur = -FLT_MAX;
    if (half == 0)
    {
        if (active)
        {
            if constexpr (PAGED)
            {
                if (kl == 0)
                    store16<INT8>(kv_row<true>(kc, pg, b, 0, h, tlen, esz), (size_t) chunk * 16, s_oq, kh);
                else if (kl == 1)
                    store16<INT8>(kv_row<true>(vc, pg, b, 1, h, tlen, esz), (size_t) chunk * 16, s_oq, vh);
            }
            else
            {
                if (kl == 0)
                    store16<INT8>(kc, (size_t) tlen * kDh + chunk * 16, s_oq, kh);
                else if (kl == 1)
                    store16<INT8>(vc, (size_t) tlen * kDh + chunk * 16, s_oq, vh);
            }
        }
        float sc0 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            sc0 = fmaf(__half2float(qh[i]), __half2float(kh[i]), sc0);
        sc0 += __shfl_xor_sync(0xffffffffu, sc0, 1);
        sc0 += __shfl_xor_sync(0xffffffffu, sc0, 2);
        s_cur = sc0 * inv_sqrt_dh;
    }

    float m_run = s_cur, l_run = 0.f; // l_run: this lane group's share of the cached keys' denominator
    float o[16];                      // in units of the dequant scale (integer V values)
#pragma unroll
    for (int i = 0; i < 16; ++i)
        o[i] = 0.f;

    for (int k0 = 0; k0 < tlen; k0 += 16 * NIT)
    {
        if (k0 > 0)
            fetch(k0);
        float sc[NIT];
        float m_new = m_run;
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            sc[it] = -FLT_MAX;
            const int kg = k0 + (2 * it + half) * 8;
            if (kg < tlen) // warp-uniform: groups of 8 keys beyond the length cost nothing
            {
                const int key = kg + kl;
                __half2 h0 = __hmul2(q2[0], kw[it][0]);
                __half2 h1 = __hmul2(q2[4], kw[it][4]);
                h0 = __hfma2(q2[1], kw[it][1], h0);
                h1 = __hfma2(q2[5], kw[it][5], h1);
                h0 = __hfma2(q2[2], kw[it][2], h0);
                h1 = __hfma2(q2[6], kw[it][6], h1);
                h0 = __hfma2(q2[3], kw[it][3], h0);
                h1 = __hfma2(q2[7], kw[it][7], h1);
                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                float sv = (f0.x + f0.y) + (f1.x + f1.y);
                sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                sv += __shfl_xor_sync(0xffffffffu, sv, 2);
                const bool valid = key < tlen && !(mask && mask[key]); // masked keys: probability 0, not in the max
                sv = valid ? sv * sscale : -FLT_MAX;
                sc[it] = sv;
                m_new = fmaxf(m_new, sv);
            }
        }
        m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, 4));
        m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, 8));
        m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, 16));
        const float corr = m_new == -FLT_MAX ? 1.f : __expf(m_run - m_new);
        m_run = m_new;
        l_run *= corr;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            o[i] *= corr;
        __half2 o2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            o2[i] = __float2half2_rn(0.f);
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            if (k0 + (2 * it + half) * 8 >= tlen || sc[it] == -FLT_MAX)
                continue; // beyond the length / masked key (also keeps stale cache bits out of the fp16 path)
            const float e = __expf(sc[it] - m_new);
            l_run += e;
            const __half2 p2 = __float2half2_rn(e);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                o2[i] = __hfma2(p2, vw[it][i], o2[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            const float2 f = __half22float2(o2[i]);
            o[2 * i] += f.x;
            o[2 * i + 1] += f.y;
        }
    }
    // reduce over the 8 key groups
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 4);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 8);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 16);
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        float v = o[i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        o[i] = v;
    }
    // warp 1 of the pair hands its state to warp 0 through shared memory
    float* pr = parts[pair];
    if (half == 1 && kl == 0)
    {
        // o[2i], o[2i+1] hold the pair of w[i]: w[2j] = dims (4j, 4j+2), w[2j+1] = dims (4j+1, 4j+3)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(pr + 4 + chunk * 16 + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        if (chunk == 0)
        {
            pr[0] = m_run;
            pr[1] = l_run;
        }
    }
    __syncthreads();
    if (half == 0 && active)
    {
        const float m1 = pr[0], l1 = pr[1];
        const float m = fmaxf(m_run, m1); // m_run >= s_cur > -FLT_MAX
        const float w0 = __expf(m_run - m), w1 = m1 == -FLT_MAX ? 0.f : __expf(m1 - m);
        const float e_cur = __expf(s_cur - m);
        const float inv_sum = __fdividef(1.f, l_run * w0 + l1 * w1 + e_cur + 1.e-6f); // Template.h:1756
        if (kl == 0)
        {
            __half* dst = static_cast<__half*>(p.out) + (size_t) b * hidden + h * kDh + chunk * 16;
            const float a0 = w0 * s_qo, a1 = w1 * s_qo;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const float4 q4 = *reinterpret_cast<const float4*>(pr + 4 + chunk * 16 + 4 * j);
                dst[4 * j + 0] = __float2half_rn((o[4 * j + 0] * a0 + q4.x * a1 + e_cur * __half2float(vh[4 * j + 0])) * inv_sum);
                dst[4 * j + 2] = __float2half_rn((o[4 * j + 1] * a0 + q4.y * a1 + e_cur * __half2float(vh[4 * j + 2])) * inv_sum);
                dst[4 * j + 1] = __float2half_rn((o[4 * j + 2] * a0 + q4.z * a1 + e_cur * __half2float(vh[4 * j + 1])) * inv_sum);
                dst[4 * j + 3] = __float2half_rn((o[4 * j + 3] * a0 + q4.w * a1 + e_cur * __half2float(vh[4 * j + 3])) * inv_sum);
            }
        }
    }
}

// =====================================================================================================
// Context phase.  grid (H, B), 128 threads; K and V of the whole prompt for this (b, h) are staged in shared
// memory as fp16 (unquantized values are used for the attention itself, as the reference does), the cache rows
// [0, S) are written (int8-quantized when requested), and each warp handles query rows i = warp, warp+4, ...
// causal: key j is visible to query i iff j <= i and j < input_length[b].
// =====================================================================================================
template <bool INT8, bool PAGED>
__global__ void __launch_bounds__(128) attention_context_kernel(const __half* __restrict__ qkv,
    const int* __restrict__ input_lengths, __half* __restrict__ out, void* __restrict__ kv_cache,
    const float* __restrict__ kv_scale_orig_quant, int S, int H, int Smax, float q_scaling, const KvPaged pg)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __half* sK = reinterpret_cast<__half*>(s_raw);          // [S][64]
    __half* sV = sK + (size_t) S * kDh;                     // [S][64]
    float* sP = reinterpret_cast<float*>(sV + (size_t) S * kDh); // [4 warps][S]

    const int h = blockIdx.x, b = blockIdx.y;
    const int hidden = H * kDh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    grid_dep_wait();
    grid_dep_launch_dependents();

    const int len = input_lengths ? min(input_lengths[b], S) : S;
    const float s_oq = INT8 ? kv_scale_orig_quant[0] : 1.f;
    const float inv_sqrt_dh = 1.f / (sqrtf((float) kDh) * q_scaling);
    const size_t esz = INT8 ? 1 : 2;
    char* kc = PAGED ? nullptr : static_cast<char*>(kv_cache) + ((size_t) (b * 2 + 0) * H + h) * Smax * kDh * esz;
    char* vc = PAGED ? nullptr : static_cast<char*>(kv_cache) + ((size_t) (b * 2 + 1) * H + h) * Smax * kDh * esz;

    // stage K, V; fill the cache.  The reference zeroes the padded rows of its K/V scratch before the transpose
    // (gptAttentionCommon.cpp:481), so padded cache rows hold quantized zeros.
    for (int idx = tid; idx < S * 4; idx += 128)
    {
        const int t = idx >> 2, c = idx & 3;
        __half kh[16], vh[16];
        if (t < len)
        {
            const __half* src = qkv + ((size_t) b * S + t) * 3 * hidden + h * kDh + c * 16;
            load16_half(src + hidden, nullptr, kh);
            load16_half(src + 2 * hidden, nullptr, vh);
        }
        else
        {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                kh[i] = vh[i] = __float2half(0.f);
        }
        *reinterpret_cast<uint4*>(&sK[t * kDh + c * 16]) = *reinterpret_cast<const uint4*>(&kh[0]);
        *reinterpret_cast<uint4*>(&sK[t * kDh + c * 16 + 8]) = *reinterpret_cast<const uint4*>(&kh[8]);
        *reinterpret_cast<uint4*>(&sV[t * kDh + c * 16]) = *reinterpret_cast<const uint4*>(&vh[0]);
        *reinterpret_cast<uint4*>(&sV[t * kDh + c * 16 + 8]) = *reinterpret_cast<const uint4*>(&vh[8]);
        store16<INT8>(kv_row<PAGED>(kc, pg, b, 0, h, t, esz), (size_t) c * 16, s_oq, kh);
        store16<INT8>(kv_row<PAGED>(vc, pg, b, 1, h, t, esz), (size_t) c * 16, s_oq, vh);
    }
    __syncthreads();

    float* myP = sP + (size_t) warp * S;
    for (int i = warp; i < S; i += 4)
    {
        // lane owns dims 2*lane, 2*lane+1 of q and of the output
        const __half2 q2 = *reinterpret_cast<const __half2*>(qkv + ((size_t) b * S + i) * 3 * hidden + h * kDh + 2 * lane);
        const float2 qf = __half22float2(q2);
        const int nvis = min(i + 1, len); // visible keys
        float lmax = -FLT_MAX;
        for (int j = 0; j < nvis; ++j)
        {
            const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(&sK[j * kDh + 2 * lane]));
            float s = qf.x * kf.x + qf.y * kf.y;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1)
                s += __shfl_xor_sync(0xffffffffu, s, o);
            s *= inv_sqrt_dh;
            if (lane == 0)
                myP[j] = s;
            lmax = fmaxf(lmax, s);
        }
        __syncwarp();
        float lsum = 0.f;
        for (int j = lane; j < nvis; j += 32)
        {
            const float e = __expf(myP[j] - lmax);
            myP[j] = e;
            lsum += e;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1)
            lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        __syncwarp();
        const float inv = __fdividef(1.f, lsum + 1.e-6f);
        float ox = 0.f, oy = 0.f;
        for (int j = 0; j < nvis; ++j)
        {
            const float pj = myP[j];
            const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(&sV[j * kDh + 2 * lane]));
            ox = fmaf(pj, vf.x, ox);
            oy = fmaf(pj, vf.y, oy);
        }
        __half2 o2 = __floats2half2_rn(ox * inv, oy * inv);
        if (nvis == 0)
            o2 = __floats2half2_rn(0.f, 0.f);
        *reinterpret_cast<__half2*>(out + ((size_t) b * S + i) * hidden + h * kDh + 2 * lane) = o2;
        __syncwarp();
    }
}

// =====================================================================================================
// Cross-attention over a (typically int8) cross-KV cache [B, 2, H, S, 64]: the dominant byte stream of the
// decoder step at batch >= 6 (3.84 MB per sequence per layer).  HBM-bound streaming design, fourth iteration:
//   v1 CTA-wide items + shared scores: 6.7 thread-instructions/byte, issue-bound at 2.1 TB/s;
//   v2 finer items + in-kernel merge: slower (five CTA barriers and a __threadfence per 16 KB item);
//   v3 warp-private items, register-staged LDG: 2.6 instructions/byte but latency-bound (long-scoreboard stalls,
//      only ~13 warps/SM with 8 KB in flight each, loads not overlapped with the math of the same warp): 2.6 TB/s.
//   v4 (this): warp-private work AND a warp-private shared-memory ring filled by 1-D TMA bulk copies (UBLKCP):
//      * a chunk = 64 keys (int8: 4 KB of K + 4 KB of V, both contiguous) -> two bulk copies onto one mbarrier;
//        each warp keeps 3 chunks in flight (24 KB), 8 warps per CTA, one CTA per SM: 192 KB in flight per SM, and
//        the copies of chunk i+3 are issued before the math of chunk i+1 starts -- no load latency on the warp's
//        critical path, no CTA barrier, no producer/consumer handshake across warps;
//      * every warp owns a CONTIGUOUS range of chunks of the flattened (row, head, chunk) space and carries the
//        online-softmax state (m, l, o) across consecutive chunks of the same (row, head); the cross-lane reduction
//        of o (48 shuffles) happens once per (row, head) range, not per chunk;
//      * lane geometry 4 lanes x 16 dims per key, 8 keys per warp instruction; the int8 cross cache is stored in
//        offset-binary form, so PRMT + HSUB2 alone give exact fp16 integers; the scores of 16 keys (two warp
//        iterations) are ONE mma.sync.m16n8k16 chain (the converted registers are already the A fragment, q is
//        column 0 of B), p.v is chained 8 keys at a time with HFMA2 and flushed to fp32, the dequant scale is hoisted
//        out of both products, the softmax runs in the log2 domain;
//      * the ranges of a (row, head) are merged by the last warp to arrive (self-resetting counter).
//   v5: cross_attention_rowhead_kernel below (whole (row, head) pairs per CTA, merge in shared memory) is what runs
//      from 8 pairs up; this split kernel remains for a handful of pairs.
// =====================================================================================================
// (warps per CTA, ring stages per warp, keys per chunk) -- one CTA per SM; WARPS*STAGES*CK*128 B of shared memory (int8)
// OCC = CTAs per SM the row-head kernel is sized for (F: half-size CTAs, so a GEMM CTA of another stream can share the SM)
struct XaCfgA { static constexpr int W = 8, ST = 3, CK = 64, OCC = 1; };
struct XaCfgB { static constexpr int W = 16, ST = 3, CK = 32, OCC = 1; };
struct XaCfgC { static constexpr int W = 12, ST = 2, CK = 64, OCC = 1; };
struct XaCfgD { static constexpr int W = 13, ST = 2, CK = 64, OCC = 1; };
struct XaCfgE { static constexpr int W = 9, ST = 3, CK = 64, OCC = 1; };
struct XaCfgF { static constexpr int W = 6, ST = 2, CK = 64, OCC = 2; };


struct XAttnParams
{
    const __half* q;   // [R, H*64]
    const void* kv;    // [B, 2, H, S, 64]
    const float* scale_quant_orig;
    __half* out;       // [R, H*64]
    float* partials;   // [R*H][max_parts][66]
    int* counters;     // [R*H] arrival counters (library owned, self-resetting)
    int B, H, S;       // B = number of query rows R
    int q_per_seq;     // query rows per cache sequence (1 in the generation phase, S_prompt in the context phase)
    int nch;           // chunks per (row, head) = ceil(S / keys per chunk)
    int chunks_per_warp;
    int max_parts;
    int early_kv; // the cache is not written by the kernel right before this one: stream it before the PDL wait
    float inv_sqrt_dh;
    long long* dbg; // B200_XA_DEBUG builds: 8 %globaltimer stamps per CTA of the row-head kernel (b200_debug_xa_timeline)
};

#if defined(B200_XA_DEBUG)
#define XA_STAMP(slot)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        if (p.dbg != nullptr && threadIdx.x == 0)                                                                      \
        {                                                                                                              \
            long long t_;                                                                                              \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                     \
            p.dbg[(size_t) blockIdx.x * 8 + (slot)] = t_;                                                              \
        }                                                                                                              \
    } while (0)
#else
#define XA_STAMP(slot)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
    } while (0)
#endif



template <bool INT8, typename CFG>
__global__ void __launch_bounds__(CFG::W * 32, 1) cross_attention_kernel(const XAttnParams p)
{
    constexpr int kXaWarps = CFG::W, kXaStages = CFG::ST;
    constexpr int ESZ = INT8 ? 1 : 2;
    constexpr int CK = INT8 ? CFG::CK : CFG::CK / 2; // keys per chunk (same bytes per chunk for both cache types)
    constexpr int NIT = CK / 8;             // warp iterations per chunk
    constexpr int kHalfBytes = CK * kDh * ESZ; // bytes of K (or V) per chunk: 4096
    constexpr int kStageBytes = 2 * kHalfBytes;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = lane & 3, kl = lane >> 2;
    uint8_t* ring = smem + (size_t) warp * kXaStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t) kXaWarps * kXaStages * kStageBytes) + warp * kXaStages;

    const int total = p.B * p.H * p.nch;
    const int gw = blockIdx.x * kXaWarps + warp;
    const int c_begin = (int) min((long long) gw * p.chunks_per_warp, (long long) total);
    const int c_end = min(c_begin + p.chunks_per_warp, total);

    if (lane == 0)
    {
        for (int s = 0; s < kXaStages; ++s)
            mbar_init(&bars[s], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncwarp();
    grid_dep_launch_dependents();

    // bulk copies of chunk c (K half then V half) into stage s; the cache is static data: no dependency wait needed
    const uint64_t pol = policy_evict_first();
    auto issue = [&](int c, int s)
    {
        const int bh = c / p.nch, ch = c - bh * p.nch;
        const int b = (bh / p.H) / p.q_per_seq, h = bh % p.H;
        const int key0 = ch * CK;
        const int nk = min(CK, p.S - key0);
        const uint32_t bytes = (uint32_t) nk * kDh * ESZ;
        const uint8_t* kb = static_cast<const uint8_t*>(p.kv) + (((size_t) (b * 2 + 0) * p.H + h) * p.S + key0) * (size_t) (kDh * ESZ);
        const uint8_t* vb = static_cast<const uint8_t*>(p.kv) + (((size_t) (b * 2 + 1) * p.H + h) * p.S + key0) * (size_t) (kDh * ESZ);
        mbar_arrive_expect_tx(&bars[s], 2 * bytes);
        bulk_g2s_hint(ring + s * kStageBytes, kb, bytes, &bars[s], pol);
        bulk_g2s_hint(ring + s * kStageBytes + kHalfBytes, vb, bytes, &bars[s], pol);
    };
    if (c_begin >= c_end)
        return;
    if (!p.early_kv)
        grid_dep_wait();
    if (lane == 0)
    {
        for (int j = 0; j < kXaStages && c_begin + j < c_end; ++j)
            issue(c_begin + j, j);
    }
    if (p.early_kv)
        grid_dep_wait(); // q comes from the previous kernel

    const float s_qo = INT8 ? p.scale_quant_orig[0] : 1.f;
    const float sscale = s_qo * p.inv_sqrt_dh * 1.4426950408889634f;

    int cur_bh = -1;
    uint32_t bq[8];
    float m_run = -FLT_MAX, l_run = 0.f; // l_run: this lane group's share of the denominator; m_run in log2 units
    float o[16];

    auto flush = [&](int bh)
    {
        // reduce this warp's running state over the 8 key groups and publish / merge it
        float l = l_run;
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
        for (int i = 0; i < 16; ++i)
        {
            float v = o[i];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            o[i] = v * s_qo; // hoisted V dequant scale
        }
        // which warps hold a part of this (row, head)?
        const int first = (bh * p.nch) / p.chunks_per_warp;
        const int last = ((bh + 1) * p.nch - 1) / p.chunks_per_warp;
        const int nparts = last - first + 1;
        // o[2i], o[2i+1] hold the pair of w[i]: w[2j] = dims (4j, 4j+2), w[2j+1] = dims (4j+1, 4j+3)
        if (nparts == 1)
        {
            if (kl == 0)
            {
                const float inv = 1.f / l;
                __half* dst = p.out + (size_t) bh * kDh + chunk * 16;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    dst[4 * j + 0] = __float2half_rn(o[4 * j + 0] * inv);
                    dst[4 * j + 2] = __float2half_rn(o[4 * j + 1] * inv);
                    dst[4 * j + 1] = __float2half_rn(o[4 * j + 2] * inv);
                    dst[4 * j + 3] = __float2half_rn(o[4 * j + 3] * inv);
                }
            }
            return;
        }
        float* pr = p.partials + ((size_t) bh * p.max_parts + (gw - first)) * (kDh + 2);
        if (kl == 0)
        {
            float* dst = pr + 2 + chunk * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                __stcg(dst + 4 * j + 0, o[4 * j + 0]);
                __stcg(dst + 4 * j + 2, o[4 * j + 1]);
                __stcg(dst + 4 * j + 1, o[4 * j + 2]);
                __stcg(dst + 4 * j + 3, o[4 * j + 3]);
            }
            if (chunk == 0)
            {
                __stcg(pr, m_run);
                __stcg(pr + 1, l);
            }
        }
        __threadfence();
        __syncwarp();
        int is_last = 0;
        if (lane == 0)
            is_last = (atomicAdd(&p.counters[bh], 1) == nparts - 1) ? 1 : 0;
        is_last = __shfl_sync(0xffffffffu, is_last, 0);
        if (is_last)
        {
            __threadfence();
            const float* pb = p.partials + (size_t) bh * p.max_parts * (kDh + 2);
            float gm = -FLT_MAX;
            for (int s2 = 0; s2 < nparts; ++s2)
                gm = fmaxf(gm, __ldcg(pb + s2 * (kDh + 2)));
            float gl = 0.f, a0 = 0.f, a1 = 0.f;
            for (int s2 = 0; s2 < nparts; ++s2)
            {
                const float* ps = pb + s2 * (kDh + 2);
                const float w = fast_exp2(__ldcg(ps) - gm);
                gl += w * __ldcg(ps + 1);
                a0 += w * __ldcg(ps + 2 + lane);
                a1 += w * __ldcg(ps + 2 + 32 + lane);
            }
            const float inv = 1.f / gl;
            p.out[(size_t) bh * kDh + lane] = __float2half_rn(a0 * inv);
            p.out[(size_t) bh * kDh + 32 + lane] = __float2half_rn(a1 * inv);
            if (lane == 0)
                p.counters[bh] = 0;
        }
    };


    for (int c = c_begin; c < c_end; ++c)
    {
        const int it_local = c - c_begin;
        const int s = it_local % kXaStages;
        const int bh = c / p.nch, ch = c - bh * p.nch;
        const int nk = min(CK, p.S - ch * CK);
        if (bh != cur_bh)
        {
            if (cur_bh >= 0)
                flush(cur_bh);
            cur_bh = bh;
            m_run = -FLT_MAX;
            l_run = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                o[i] = 0.f;
            __half qh[16];
            load16_half(p.q + (size_t) bh * kDh + chunk * 16, nullptr, qh);
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                // B fragment of the score MMA: q is column 0, i.e. only the lanes of key group 0 carry it
                bq[2 * i] = kl == 0 ? h2u(__halves2half2(qh[4 * i], qh[4 * i + 2])) : 0u;
                bq[2 * i + 1] = kl == 0 ? h2u(__halves2half2(qh[4 * i + 1], qh[4 * i + 3])) : 0u;
            }
        }
        mbar_wait(&bars[s], (it_local / kXaStages) & 1);
        const uint8_t* kst = ring + s * kStageBytes + (size_t) (kl * kDh + chunk * 16) * ESZ;
        const uint8_t* vst = kst + kHalfBytes;

        if (nk == CK)
            xa_chunk<INT8, NIT, true>(kst, vst, nk, kl, lane, sscale, bq, m_run, l_run, o);
        else
            xa_chunk<INT8, NIT, false>(kst, vst, nk, kl, lane, sscale, bq, m_run, l_run, o);
        // stage s is drained: refill it with chunk c + kXaStages
        __syncwarp();
        if (lane == 0 && c + kXaStages < c_end)
            issue(c + kXaStages, s);
    }
    flush(cur_bh);
}

// ---- cross attention, one CTA per (row, head) at a time ("row-head" kernel) --------------------------------------
// Used when there are enough (row, head) pairs to fill the GPU.  CTA i walks the pairs i, i + grid, ...; the W warps
// of the CTA split the pair's chunks into contiguous runs, stream them through their private TMA rings exactly like
// the split kernel above, and merge their partial softmax states through shared memory -- no global partials, no
// fences, no counters.  q of the next pair is fetched while the current one is being processed.
//
// SPLIT (launched as clusters of two CTAs): the pairs that do not fill a whole round of the grid -- 320 pairs on 148 CTAs
// leave 24 -- are not dealt whole to the first CTAs (a third pair for 24 of them while 124 wait: the kernel is issue-bound
// on exactly those CTAs) but split by keys over the two CTAs of a cluster.  Rank 1 pushes its warps' partial softmax
// states into rank 0's shared memory (st.async, complete_tx on an mbarrier there) and rank 0 merges all 2 W of them:
// critical path 2.5 instead of 3 pairs.
__device__ __forceinline__ uint32_t xa_cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t xa_cluster_id()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void xa_push_f32(uint32_t remote_addr, float v, uint32_t remote_mbar)
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
                 "r"(__float_as_uint(v)), "r"(remote_mbar)
                 : "memory");
}
__device__ __forceinline__ uint32_t xa_mapa(uint32_t local_smem_addr, uint32_t cta_rank)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(cta_rank));
    return remote;
}

// CS = 0: whole pairs only.  CS = 2 / 4: launched as clusters of CS CTAs that share ONE pair per cluster by keys -- the pair left
// over after the whole rounds (CS = 2, many pairs), or, with fewer pairs than SMs / CS (batch 1-3: 20-60 pairs), every pair: the
// grid then has pairs x CS CTAs instead of one CTA per pair on a mostly idle GPU.
template <bool INT8, typename CFG, int CS = 0>
__global__ void __launch_bounds__(CFG::W * 32, CFG::OCC) cross_attention_rowhead_kernel(const XAttnParams p)
{
    constexpr bool SPLIT = CS > 0;
    constexpr int kRanks = CS > 0 ? CS : 1;
    constexpr int W = CFG::W, ST = CFG::ST;
    constexpr int ESZ = INT8 ? 1 : 2;
    constexpr int CK = INT8 ? CFG::CK : CFG::CK / 2;
    constexpr int NIT = CK / 8;
    constexpr int kHalfBytes = CK * kDh * ESZ;
    constexpr int kStageBytes = 2 * kHalfBytes;
    constexpr int kPart = kDh + 4; // m, l, 2 pad floats (keeps the o rows 16-byte aligned), o[64]
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = lane & 3, kl = lane >> 2;
    uint8_t* ring = smem + (size_t) warp * ST * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t) W * ST * kStageBytes) + warp * ST;
    // (barrier block rounded up to 16 bytes: the partials are read as float4 -- an odd W * ST, config E, misaligned them)
    float* parts = reinterpret_cast<float*>(smem + (size_t) W * ST * kStageBytes + ((sizeof(uint64_t) * W * ST + 15) & ~size_t(15))); // [2][W][68]

    XA_STAMP(0);
    const int RH = p.B * p.H;
    const int wc0 = warp * p.nch / W, wc1 = (warp + 1) * p.nch / W;                   // this warp's chunks of every pair
    const int n_w = wc1 - wc0;
    // SPLIT: whole rounds of the grid are dealt pair by pair; the remaining pairs (host: at most one per cluster) are
    // shared by the two CTAs of a cluster: rank k takes chunks [k nch / 2, (k + 1) nch / 2), split over its warps
    const int rounds = RH / (int) gridDim.x;
    const int nbh = SPLIT ? rounds : (RH - (int) blockIdx.x + (int) gridDim.x - 1) / (int) gridDim.x; // whole pairs of this CTA
    int left_bh = -1, lc0 = 0, n_l = 0; // the shared pair, this warp's first chunk of it and chunk count
    uint32_t crank = 0;
    float* lparts = nullptr;   // [2 W][68] partial states of the shared pair (rank 0 merges)
    uint64_t* lbar = nullptr;  // rank 0: the partials of rank 1 have landed
    if constexpr (SPLIT)
    {
        lparts = parts + 2 * W * kPart;
        lbar = reinterpret_cast<uint64_t*>(lparts + kRanks * W * kPart);
        crank = xa_cluster_rank();
        const int cand = rounds * (int) gridDim.x + (int) xa_cluster_id();
        if (cand < RH)
        {
            left_bh = cand;
            const int h0 = (int) crank * p.nch / kRanks, h1 = ((int) crank + 1) * p.nch / kRanks;
            lc0 = h0 + warp * (h1 - h0) / W;
            n_l = h0 + (warp + 1) * (h1 - h0) / W - lc0;
        }
    }
    const int tot = nbh * n_w + n_l;

    if (lane == 0)
    {
        for (int s = 0; s < ST; ++s)
            mbar_init(&bars[s], 1);
        if (SPLIT && warp == 0 && left_bh >= 0 && crank == 0)
        {
            mbar_init(lbar, 1);
            mbar_arrive_expect_tx(lbar, (uint32_t) ((kRanks - 1) * W * (kDh + 2) * sizeof(float)));
        }
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncwarp();
    if constexpr (SPLIT)
    {
        // rank 0's inbox barrier is armed before rank 1 may push into it: arrive now, rank 1 waits right before its push
        if (left_bh >= 0)
        {
            __syncthreads();
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        }
    }
    grid_dep_launch_dependents();

    const uint64_t pol = policy_evict_first();
    auto issue = [&](int i, int s)
    {
        int bh, ch;
        if (SPLIT && i >= nbh * n_w)
        {
            bh = left_bh;
            ch = lc0 + (i - nbh * n_w);
        }
        else
        {
            const int r = i / n_w;
            ch = wc0 + (i - r * n_w);
            bh = blockIdx.x + r * gridDim.x;
        }
        const int b = (bh / p.H) / p.q_per_seq, h = bh % p.H;
        const int key0 = ch * CK;
        const int nk = min(CK, p.S - key0);
        const uint32_t bytes = (uint32_t) nk * kDh * ESZ;
        const uint8_t* kb = static_cast<const uint8_t*>(p.kv) + (((size_t) (b * 2 + 0) * p.H + h) * p.S + key0) * (size_t) (kDh * ESZ);
        const uint8_t* vb = static_cast<const uint8_t*>(p.kv) + (((size_t) (b * 2 + 1) * p.H + h) * p.S + key0) * (size_t) (kDh * ESZ);
        mbar_arrive_expect_tx(&bars[s], 2 * bytes);
        bulk_g2s_hint(ring + s * kStageBytes, kb, bytes, &bars[s], pol);
        bulk_g2s_hint(ring + s * kStageBytes + kHalfBytes, vb, bytes, &bars[s], pol);
    };
    if (!p.early_kv)
        grid_dep_wait();
    if (lane == 0)
    {
        for (int j = 0; j < ST && j < tot; ++j)
            issue(j, j);
    }
    // the dequant scale is static data like the cache: requested ahead of the dependency wait.  (Behind the wait, next to
    // the q load, it was the longer of the two stalls -- 13.9 % of all warp samples against 6.5 % for q: a 4-byte tensor
    // touched once per step does not survive 2.8 GB of streaming in L2, q was just written.)
    const float s_qo = INT8 ? __ldg(p.scale_quant_orig) : 1.f;
    if (p.early_kv)
        grid_dep_wait(); // q comes from the previous kernel

    XA_STAMP(1);
    // q of the first pair
    const int first_bh = (!SPLIT || nbh > 0) ? (int) blockIdx.x : left_bh;
    const uint4* qsrc = reinterpret_cast<const uint4*>(p.q + (size_t) first_bh * kDh + chunk * 16);
    uint4 qn0 = __ldg(qsrc), qn1 = __ldg(qsrc + 1);
    const float sscale = s_qo * p.inv_sqrt_dh * 1.4426950408889634f;

    const int npairs = nbh + ((SPLIT && left_bh >= 0) ? 1 : 0);
    int i = 0;
    for (int r = 0; r < npairs; ++r)
    {
        const bool shared = SPLIT && r == nbh; // the pair this CTA shares with the other CTA of its cluster
        const int bh = shared ? left_bh : (int) blockIdx.x + r * (int) gridDim.x;
        const int c_first = shared ? lc0 : wc0, c_cnt = shared ? n_l : n_w;
        uint32_t bq[8];
        float koff = 0.f; // int8 cache: 1152 * sum of q over the 64 dims, the bias of the 1024 + byte key values (xa_chunk KOFF)
        {
            const uint32_t u[8] = {qn0.x, qn0.y, qn0.z, qn0.w, qn1.x, qn1.y, qn1.z, qn1.w}; // u[j] = (d2j, d2j+1)
            float qs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                // B fragment of the score MMA: q is column 0, i.e. only the lanes of key group 0 carry it
                bq[2 * j] = kl == 0 ? __byte_perm(u[2 * j], u[2 * j + 1], 0x5410) : 0u;     // (d4j, d4j+2)
                bq[2 * j + 1] = kl == 0 ? __byte_perm(u[2 * j], u[2 * j + 1], 0x7632) : 0u; // (d4j+1, d4j+3)
                if constexpr (INT8)
                {
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u[2 * j]));
                    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u[2 * j + 1]));
                    qs += (f0.x + f0.y) + (f1.x + f1.y);
                }
            }
            if constexpr (INT8)
            {
                qs += __shfl_xor_sync(0xffffffffu, qs, 1);
                qs += __shfl_xor_sync(0xffffffffu, qs, 2);
                koff = 1152.f * qs;
            }
        }
        if (r + 1 < npairs)
        {
            const int nbh_next = (SPLIT && r + 1 == nbh) ? left_bh : bh + (int) gridDim.x;
            const uint4* qs = reinterpret_cast<const uint4*>(p.q + (size_t) nbh_next * kDh + chunk * 16);
            qn0 = __ldg(qs);
            qn1 = __ldg(qs + 1);
        }
        float m_run = -FLT_MAX, l_run = 0.f;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            o[j] = 0.f;
        for (int j = 0; j < c_cnt; ++j, ++i)
        {
            const int s = i % ST;
            const int nk = min(CK, p.S - (c_first + j) * CK);
            mbar_wait(&bars[s], (i / ST) & 1);
            const uint8_t* kst = ring + s * kStageBytes + (size_t) (kl * kDh + chunk * 16) * ESZ;
            const uint8_t* vst = kst + kHalfBytes;
            // int8 cache: the keys enter the score MMA as 1024 + byte (PRMT only), the constant part is removed once per
            // (row, head) through koff -- 8 HSUB2 fewer per 8 keys in a loop whose issue slots are its limit (ncu source
            // view of the round-1 kernel: 128 PRMT + 128 HADD2 of 471 instructions per 64-key chunk)
            if (nk == CK)
                xa_chunk<INT8, NIT, true, INT8>(kst, vst, nk, kl, lane, sscale, bq, m_run, l_run, o, koff);
            else
                xa_chunk<INT8, NIT, false, INT8>(kst, vst, nk, kl, lane, sscale, bq, m_run, l_run, o, koff);
            __syncwarp();
            if (lane == 0 && i + ST < tot)
                issue(i + ST, s);
        }
        // this warp's state, reduced over its 8 key groups -> shared memory
        float l = l_run;
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
        for (int j = 0; j < 16; ++j)
        {
            float v = o[j];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            o[j] = v;
        }
        XA_STAMP(2 + (r < 4 ? r : 4)); // warp 0 finished its chunks of pair r
        if (SPLIT && shared && crank != 0)
        {
            // rank 1: push this warp's partial into slot W + warp of rank 0's inbox and leave
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); // rank 0 has armed its inbox barrier
            if (kl == 0)
            {
                const uint32_t dst = xa_mapa(smem_u32(lparts + (size_t) ((int) crank * W + warp) * kPart), 0u);
                const uint32_t bar = xa_mapa(smem_u32(lbar), 0u);
                const uint32_t od = dst + (uint32_t) (4 + chunk * 16) * 4u;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    // o[2i], o[2i+1] hold the pair of w[i]: w[2j] = dims (4j, 4j+2), w[2j+1] = dims (4j+1, 4j+3)
                    xa_push_f32(od + (uint32_t) (4 * j + 0) * 4u, o[4 * j + 0], bar);
                    xa_push_f32(od + (uint32_t) (4 * j + 1) * 4u, o[4 * j + 2], bar);
                    xa_push_f32(od + (uint32_t) (4 * j + 2) * 4u, o[4 * j + 1], bar);
                    xa_push_f32(od + (uint32_t) (4 * j + 3) * 4u, o[4 * j + 3], bar);
                }
                if (chunk == 0)
                {
                    xa_push_f32(dst, m_run, bar);
                    xa_push_f32(dst + 4u, l, bar);
                }
            }
            XA_STAMP(7);
            return;
        }
        float* pr = (SPLIT && shared) ? lparts + (size_t) warp * kPart : parts + ((r & 1) * W + warp) * kPart;
        if (kl == 0)
        {
            // o[2i], o[2i+1] hold the pair of w[i]: w[2j] = dims (4j, 4j+2), w[2j+1] = dims (4j+1, 4j+3)
            float* dst = pr + 4 + chunk * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(o[4 * j + 0], o[4 * j + 2], o[4 * j + 1], o[4 * j + 3]);
            if (chunk == 0)
            {
                pr[0] = m_run;
                pr[1] = l;
            }
        }
        __syncthreads();
        if (warp == r % W)
        {
            const float* pb = (SPLIT && shared) ? lparts : parts + (r & 1) * W * kPart;
            const int np = (SPLIT && shared) ? kRanks * W : W;
            if (SPLIT && shared)
                mbar_wait(lbar, 0); // the W partials of rank 1 have landed
            float gm = -FLT_MAX;
#pragma unroll
            for (int w2 = 0; w2 < kRanks * W; ++w2)
                if (w2 < np)
                    gm = fmaxf(gm, pb[w2 * kPart]);
            float gl = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < kRanks * W; ++w2)
            {
                if (w2 < np)
                {
                    const float* ps = pb + w2 * kPart;
                    const float wt = fast_exp2(ps[0] - gm);
                    gl += wt * ps[1];
                    a0 += wt * ps[4 + lane];
                    a1 += wt * ps[4 + 32 + lane];
                }
            }
            const float inv = s_qo / gl; // hoisted V dequant scale
            p.out[(size_t) bh * kDh + lane] = __float2half_rn(a0 * inv);
            p.out[(size_t) bh * kDh + 32 + lane] = __float2half_rn(a1 * inv);
        }
    }
    XA_STAMP(7);
}

// ---- cross attention with its q projection inside ("q-fused row-head" kernel, generation phase, int8 cache) -------
// One launch per layer instead of two: the LayerNorm-folded q projection (x -> LN -> Wq, the `cross_q` GEMM of the chain)
// is computed by the attention CTAs themselves.  What this buys is not the projection's arithmetic (82 k MACs per
// (row, head)) but a kernel boundary: the cross-KV stream -- 61 MB per layer at batch 16, the largest byte stream of the
// step -- starts when the self-attention output projection finishes instead of one 3 us projection kernel later, and the
// q projection runs under the cover of the ring's first fill (W x ST x 8 KB per SM).
//  * CTA -> (head h, a contiguous run of batch rows): CTAs c, c + H, c + 2H, ... serve head c % H and split the B rows
//    of that head between them (148 CTAs, 20 heads, 16 rows: 2 or 3 rows per CTA -- the same 3-pair critical path as the
//    round-robin dealing of the plain row-head kernel), so ONE 64-column slice of Wq (K x 64 int8) serves all pairs of
//    the CTA.
//  * q = LN(x) Wq on mma.sync m16n8k16: the CTA's rows are rows 0..7 of A, the warps split K in 64-wide blocks.  The
//    reference weight layout stores, per column and k-block, four 16-byte chunks whose 32-bit words are exactly the
//    B fragments of four consecutive k16 steps; a lane loads the whole 16-byte chunk `t = lane % 4` (coalesced: a quad
//    reads 64 contiguous bytes, a column pair 128) and the k order inside the block is permuted consistently for A
//    (lane (g, t) feeds x[row g][64 kb + 16 t + ...]), which a dot product does not notice.  Same arithmetic as the
//    folded GEMM: exact integers x gamma rounded to fp16 (HMUL2), fp32 accumulation, then
//    q = fp16(fp16(rstd * (scale * acc - mean * c1s) + c2) + bias)   (b200_woq_ln_fold_prepare supplies c1s, c2).
//  * weights, gamma, scales and the first ring stages are requested BEFORE griddepcontrol.wait (static data).
struct XqParams
{
    const __half* x;      // [B, K] raw residual rows, K = H * 64
    const uint8_t* W;     // processed int8 weight of the q Linear, [K/2][2K] bytes (N = K)
    const __half* scales; // [K]
    const __half* bias;   // [K] or null
    const __half* gamma;  // [K]
    const float* c1s;     // [K]
    const float* c2;      // [K]
    const void* kv;       // [B, 2, H, S, 64] offset-binary int8
    const float* scale_quant_orig;
    __half* out;          // [B, K]
    int B, H, S, nch;
    int maxr;             // rows per CTA the shared-memory layout is sized for (<= 8)
    float eps, inv_sqrt_dh;
};

constexpr int kXqMaxKb = 2; // 64-wide k-blocks per warp (H <= 2 * warps)

template <typename CFG>
__global__ void __launch_bounds__(CFG::W * 32, 1) cross_attention_qproj_kernel(const XqParams p)
{
    constexpr int W = CFG::W, ST = CFG::ST, CK = CFG::CK;
    constexpr int NIT = CK / 8;
    constexpr int kHalfBytes = CK * kDh;
    constexpr int kStageBytes = 2 * kHalfBytes;
    constexpr int kPart = kDh + 4;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = lane & 3, kl = lane >> 2; // attention: 16-dim slice / key group;  projection: t / g of the fragments
    uint8_t* ring = smem + (size_t) warp * ST * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t) W * ST * kStageBytes) + warp * ST;
    float* parts = reinterpret_cast<float*>(smem + (size_t) W * ST * kStageBytes + ((sizeof(uint64_t) * W * ST + 15) & ~size_t(15))); // [2][W][68]
    float* qpart = parts + 2 * W * kPart;                       // [W][maxr][64] partial projections (one per k range)
    float* stat = qpart + (size_t) W * p.maxr * kDh;            // [W][maxr][2] partial sums of (x - x0), (x - x0)^2
    float* x0s = stat + (size_t) W * p.maxr * 2;                // [maxr rounded up to 4] x0 of every row
    __half* qs = reinterpret_cast<__half*>(x0s + ((p.maxr + 3) & ~3)); // [maxr][64] finished q rows (16-byte aligned)

    const int K = p.H * kDh, nkb = p.H;
    // this CTA's head and rows
    const int h = (int) blockIdx.x % p.H, idx = (int) blockIdx.x / p.H;
    const int nc = ((int) gridDim.x - h + p.H - 1) / p.H; // CTAs serving head h
    const int row0 = idx * p.B / nc, nr = (idx + 1) * p.B / nc - row0; // 1 <= nr <= maxr (host guarantees)
    const int wc0 = warp * p.nch / W, wc1 = (warp + 1) * p.nch / W;
    const int n_w = wc1 - wc0;
    const int tot = nr * n_w;

    if (lane == 0)
    {
        for (int s = 0; s < ST; ++s)
            mbar_init(&bars[s], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncwarp();
    grid_dep_launch_dependents();

    const uint64_t pol = policy_evict_first();
    auto issue = [&](int i, int s)
    {
        const int r = i / n_w, ch = wc0 + (i - r * n_w);
        const int b = row0 + r;
        const int key0 = ch * CK;
        const int nk = min(CK, p.S - key0);
        const uint32_t bytes = (uint32_t) nk * kDh;
        const uint8_t* kb = static_cast<const uint8_t*>(p.kv) + (((size_t) (b * 2 + 0) * p.H + h) * p.S + key0) * (size_t) kDh;
        const uint8_t* vb = static_cast<const uint8_t*>(p.kv) + (((size_t) (b * 2 + 1) * p.H + h) * p.S + key0) * (size_t) kDh;
        mbar_arrive_expect_tx(&bars[s], 2 * bytes);
        bulk_g2s_hint(ring + s * kStageBytes, kb, bytes, &bars[s], pol);
        bulk_g2s_hint(ring + s * kStageBytes + kHalfBytes, vb, bytes, &bars[s], pol);
    };
    // the cache is static inside a decoder step: the stream starts before the dependency wait
    if (lane == 0)
    {
        for (int j = 0; j < ST && j < tot; ++j)
            issue(j, j);
    }

    // ---- static operands of the projection: this warp's k-blocks of the head's 64 weight columns, gamma, and the
    // per-column vectors of the outputs this thread will finish ----
    const int kb0 = warp * nkb / W, nkw = (warp + 1) * nkb / W - kb0; // 0 .. kXqMaxKb
    uint4 wv[kXqMaxKb][8], gv[kXqMaxKb][2];
#pragma unroll
    for (int j = 0; j < kXqMaxKb; ++j)
    {
        if (j < nkw)
        {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
            {
                const int n = h * kDh + 8 * nt + kl;
                wv[j][nt] = __ldg(reinterpret_cast<const uint4*>(
                    p.W + (size_t) (n >> 1) * 2 * K + (n & 1) * 64 + (size_t) (kb0 + j) * 128 + chunk * 16));
            }
            const uint4* g4 = reinterpret_cast<const uint4*>(p.gamma + (size_t) (kb0 + j) * 64 + chunk * 16);
            gv[j][0] = __ldg(g4);
            gv[j][1] = __ldg(g4 + 1);
        }
    }
    const int fr = (int) threadIdx.x >> 6, fc = (int) threadIdx.x & 63; // finishing thread: row slot, column of the head
    float f_scale = 0.f, f_c1 = 0.f, f_c2 = 0.f, f_bias = 0.f;
    if (fr < nr)
    {
        const int n = h * kDh + fc;
        f_scale = __half2float(__ldg(p.scales + n));
        f_c1 = __ldg(p.c1s + n);
        f_c2 = __ldg(p.c2 + n);
        if (p.bias != nullptr)
            f_bias = __half2float(__ldg(p.bias + n));
    }
    const float s_qo = __ldg(p.scale_quant_orig);
    const float sscale = s_qo * p.inv_sqrt_dh * 1.4426950408889634f;

    grid_dep_wait(); // x comes from the previous kernel

    // ---- q projection of the CTA's rows ----
    {
        const bool live = kl < nr; // fragment row g = kl carries batch row row0 + g
        const __half* xrow = p.x + (size_t) (row0 + (live ? kl : 0)) * K;
        uint4 xa[kXqMaxKb][2];
        float sh = 0.f;
        if (live)
            sh = __half2float(__ldcg(xrow));
#pragma unroll
        for (int j = 0; j < kXqMaxKb; ++j)
        {
            xa[j][0] = xa[j][1] = make_uint4(0u, 0u, 0u, 0u);
            if (j < nkw && live)
            {
                const uint4* x4 = reinterpret_cast<const uint4*>(xrow + (size_t) (kb0 + j) * 64 + chunk * 16);
                xa[j][0] = __ldcg(x4);
                xa[j][1] = __ldcg(x4 + 1);
            }
        }
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
            acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int j = 0; j < kXqMaxKb; ++j)
        {
            if (j < nkw)
            {
                const uint32_t a_lo[4] = {xa[j][0].x, xa[j][0].y, xa[j][0].z, xa[j][0].w};
                const uint32_t a_hi[4] = {xa[j][1].x, xa[j][1].y, xa[j][1].z, xa[j][1].w};
                const uint32_t g_lo[4] = {gv[j][0].x, gv[j][0].y, gv[j][0].z, gv[j][0].w};
                const uint32_t g_hi[4] = {gv[j][1].x, gv[j][1].y, gv[j][1].z, gv[j][1].w};
                if (live)
                {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                    {
                        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&a_lo[c]));
                        const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&a_hi[c]));
                        const float d0 = f0.x - sh, d1 = f0.y - sh, d2 = f1.x - sh, d3 = f1.y - sh;
                        sa += (d0 + d1) + (d2 + d3);
                        sb = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, sb))));
                    }
                }
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                {
                    const uint32_t ww[4] = {wv[j][nt].x, wv[j][nt].y, wv[j][nt].z, wv[j][nt].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                    {
                        __half2 lo, hi;
                        dequant_word(ww[c], lo, hi);
                        const __half2 b0 = __hmul2(lo, *reinterpret_cast<const __half2*>(&g_lo[c]));
                        const __half2 b1 = __hmul2(hi, *reinterpret_cast<const __half2*>(&g_hi[c]));
                        mma_m16n8k16(acc[nt][0], acc[nt][1], acc[nt][2], acc[nt][3], a_lo[c], 0u, a_hi[c], 0u, h2u(b0), h2u(b1));
                    }
                }
            }
        }
        sa += __shfl_xor_sync(0xffffffffu, sa, 1);
        sa += __shfl_xor_sync(0xffffffffu, sa, 2);
        sb += __shfl_xor_sync(0xffffffffu, sb, 1);
        sb += __shfl_xor_sync(0xffffffffu, sb, 2);
        if (live)
        {
            float* qp = qpart + ((size_t) warp * p.maxr + kl) * kDh + 2 * chunk;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
                *reinterpret_cast<float2*>(qp + 8 * nt) = make_float2(acc[nt][0], acc[nt][1]);
            if (chunk == 0)
            {
                *reinterpret_cast<float2*>(stat + ((size_t) warp * p.maxr + kl) * 2) = make_float2(sa, sb);
                if (warp == 0)
                    x0s[kl] = sh;
            }
        }
        __syncthreads();
        if (fr < nr)
        {
            float sum = 0.f, ta = 0.f, tb = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < W; ++w2)
            {
                sum += qpart[((size_t) w2 * p.maxr + fr) * kDh + fc];
                const float2 st2 = *reinterpret_cast<const float2*>(stat + ((size_t) w2 * p.maxr + fr) * 2);
                ta += st2.x;
                tb += st2.y;
            }
            const float x0 = x0s[fr];
            const float rk = 1.f / (float) K;
            const float mean = x0 + ta * rk;
            const float var = fmaxf((tb - ta * ta * rk) * rk, 0.f);
            const float rstd = rsqrtf(var + p.eps);
            const float y = rstd * (sum * f_scale - mean * f_c1) + f_c2;
            __half o = __float2half_rn(y);
            if (p.bias != nullptr)
                o = __float2half_rn(__half2float(o) + f_bias);
            qs[fr * kDh + fc] = o;
        }
        __syncthreads();
    }

    int i = 0;
    for (int r = 0; r < nr; ++r)
    {
        const int bh = (row0 + r) * p.H + h;
        uint32_t bq[8];
        float koff = 0.f; // 1152 * sum of q over the 64 dims, the bias of the 1024 + byte key values (xa_chunk KOFF)
        {
            const uint4* qsrc = reinterpret_cast<const uint4*>(qs + r * kDh + chunk * 16);
            const uint4 qn0 = qsrc[0], qn1 = qsrc[1];
            const uint32_t u[8] = {qn0.x, qn0.y, qn0.z, qn0.w, qn1.x, qn1.y, qn1.z, qn1.w}; // u[j] = (d2j, d2j+1)
            float qsum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                bq[2 * j] = kl == 0 ? __byte_perm(u[2 * j], u[2 * j + 1], 0x5410) : 0u;     // (d4j, d4j+2)
                bq[2 * j + 1] = kl == 0 ? __byte_perm(u[2 * j], u[2 * j + 1], 0x7632) : 0u; // (d4j+1, d4j+3)
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u[2 * j]));
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u[2 * j + 1]));
                qsum += (f0.x + f0.y) + (f1.x + f1.y);
            }
            qsum += __shfl_xor_sync(0xffffffffu, qsum, 1);
            qsum += __shfl_xor_sync(0xffffffffu, qsum, 2);
            koff = 1152.f * qsum;
        }
        float m_run = -FLT_MAX, l_run = 0.f;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            o[j] = 0.f;
        for (int j = 0; j < n_w; ++j, ++i)
        {
            const int s = i % ST;
            const int nk = min(CK, p.S - (wc0 + j) * CK);
            mbar_wait(&bars[s], (i / ST) & 1);
            const uint8_t* kst = ring + s * kStageBytes + (size_t) (kl * kDh + chunk * 16);
            const uint8_t* vst = kst + kHalfBytes;
            if (nk == CK)
                xa_chunk<true, NIT, true, true>(kst, vst, nk, kl, lane, sscale, bq, m_run, l_run, o, koff);
            else
                xa_chunk<true, NIT, false, true>(kst, vst, nk, kl, lane, sscale, bq, m_run, l_run, o, koff);
            __syncwarp();
            if (lane == 0 && i + ST < tot)
                issue(i + ST, s);
        }
        float l = l_run;
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
        for (int j = 0; j < 16; ++j)
        {
            float v = o[j];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            o[j] = v;
        }
        float* pr = parts + ((r & 1) * W + warp) * kPart;
        if (kl == 0)
        {
            float* dst = pr + 4 + chunk * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(o[4 * j + 0], o[4 * j + 2], o[4 * j + 1], o[4 * j + 3]);
            if (chunk == 0)
            {
                pr[0] = m_run;
                pr[1] = l;
            }
        }
        __syncthreads();
        if (warp == r % W)
        {
            const float* pb = parts + (r & 1) * W * kPart;
            float gm = -FLT_MAX;
#pragma unroll
            for (int w2 = 0; w2 < W; ++w2)
                gm = fmaxf(gm, pb[w2 * kPart]);
            float gl = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < W; ++w2)
            {
                const float* ps = pb + w2 * kPart;
                const float wt = fast_exp2(ps[0] - gm);
                gl += wt * ps[1];
                a0 += wt * ps[4 + lane];
                a1 += wt * ps[4 + 32 + lane];
            }
            const float inv = s_qo / gl;
            p.out[(size_t) bh * kDh + lane] = __float2half_rn(a0 * inv);
            p.out[(size_t) bh * kDh + 32 + lane] = __float2half_rn(a1 * inv);
        }
    }
}

// fp16 K, V [B, S, H*64] -> cache [B, 2, H, S, 64] (int8-quantized or fp16).  grid (S, B), 128 threads... one
// thread per 16-element chunk: H*4 chunks per token for K and again for V.
template <bool INT8>
__global__ void cross_kv_pack_kernel(const __half* __restrict__ k, const __half* __restrict__ v, void* __restrict__ cache,
    const float* __restrict__ scale_orig_quant, int S, int H)
{
    const int t = blockIdx.x, b = blockIdx.y;
    const float s_oq = INT8 ? scale_orig_quant[0] : 1.f;
    for (int idx = threadIdx.x; idx < 2 * H * 4; idx += blockDim.x)
    {
        const int kv = idx / (H * 4);
        const int r = idx % (H * 4);
        const int h = r >> 2, c = r & 3;
        const __half* src = (kv == 0 ? k : v) + ((size_t) b * S + t) * H * kDh + h * kDh + c * 16;
        __half x[16];
        load16_half(src, nullptr, x);
        const size_t off = (((size_t) (b * 2 + kv) * H + h) * S + t) * kDh + c * 16;
        store16<INT8, 0x80808080u>(cache, off, s_oq, x);
    }
}

} // namespace b200

using namespace b200;

static int paged_view(const void* const* block_pointers, int max_blocks_per_seq, int tokens_per_block, int max_seq_len,
    KvPaged* pg)
{
    B200_REQUIRE(block_pointers != nullptr, B200_ERR_INVALID_ARG, "null pointer (block_pointers)");
    B200_REQUIRE(tokens_per_block > 0 && (tokens_per_block & (tokens_per_block - 1)) == 0, B200_ERR_INVALID_ARG,
        "tokens_per_block %d must be a power of 2 (kvCacheUtils.h:45-46)", tokens_per_block);
    B200_REQUIRE(max_blocks_per_seq > 0 && (long long) max_blocks_per_seq * tokens_per_block >= max_seq_len,
        B200_ERR_INVALID_ARG, "%d blocks of %d tokens do not cover max_seq_len %d", max_blocks_per_seq, tokens_per_block,
        max_seq_len);
    pg->table = block_pointers;
    pg->max_blocks_per_seq = max_blocks_per_seq;
    pg->tokens_per_block_log2 = 0;
    while ((1 << pg->tokens_per_block_log2) < tokens_per_block)
        ++pg->tokens_per_block_log2;
    return B200_OK;
}

static int mmha_generation_impl(const b200_mmha_params* p, const KvPaged* paged, b200_stream_t stream)
{
    B200_REQUIRE(p != nullptr, B200_ERR_INVALID_ARG, "null params");
    B200_REQUIRE(p->qkv && p->out && (p->kv_cache || paged), B200_ERR_INVALID_ARG, "null pointer (qkv/out/kv_cache)");
    B200_REQUIRE(p->head_size == kDh, B200_ERR_UNSUPPORTED, "head_size %d unsupported (only 64)", p->head_size);
    B200_REQUIRE(p->batch_size >= 0 && p->num_heads > 0 && p->max_seq_len > 0, B200_ERR_INVALID_ARG, "bad sizes");
    B200_REQUIRE(p->past_kv_length >= 0 && p->past_kv_length < p->max_seq_len, B200_ERR_INVALID_ARG,
        "past_kv_length %d must be in [0, max_seq_len=%d)", p->past_kv_length, p->max_seq_len);
    B200_REQUIRE(!p->int8_kv_cache || (p->kv_scale_orig_quant && p->kv_scale_quant_orig), B200_ERR_INVALID_ARG,
        "int8 KV cache needs both scales");
    B200_REQUIRE(p->q_scaling != 0.f, B200_ERR_INVALID_ARG, "q_scaling must be non-zero");
    if (p->batch_size == 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    const dim3 grid((p->num_heads * p->batch_size + kMmhaWarps / 2 - 1) / (kMmhaWarps / 2));
    const int early = static_kv_hint() ? 1 : 0;
    const KvPaged none{nullptr, 0, 0};
    if (paged != nullptr)
    {
        if (p->int8_kv_cache)
            B200_LAUNCH((mmha_generation_kernel<true, true>), grid, dim3(kMmhaWarps * 32), 0, as_stream(stream), *p, early, *paged);
        else
            B200_LAUNCH((mmha_generation_kernel<false, true>), grid, dim3(kMmhaWarps * 32), 0, as_stream(stream), *p, early, *paged);
    }
    else if (p->int8_kv_cache)
        B200_LAUNCH((mmha_generation_kernel<true, false>), grid, dim3(kMmhaWarps * 32), 0, as_stream(stream), *p, early, none);
    else
        B200_LAUNCH((mmha_generation_kernel<false, false>), grid, dim3(kMmhaWarps * 32), 0, as_stream(stream), *p, early, none);
    return B200_OK;
}

extern "C" int b200_mmha_generation(const b200_mmha_params* p, b200_stream_t stream)
{
    return mmha_generation_impl(p, nullptr, stream);
}

extern "C" int b200_mmha_generation_paged(const b200_mmha_params* p, const void* const* block_pointers, int max_blocks_per_seq,
    int tokens_per_block, b200_stream_t stream)
{
    B200_REQUIRE(p != nullptr, B200_ERR_INVALID_ARG, "null params");
    KvPaged pg{};
    if (int rc = paged_view(block_pointers, max_blocks_per_seq, tokens_per_block, p->max_seq_len, &pg))
        return rc;
    return mmha_generation_impl(p, &pg, stream);
}

static int attention_context_impl(const void* qkv, const int32_t* input_lengths, void* out, void* kv_cache,
    const KvPaged* paged, const float* kv_scale_orig_quant, int batch_size, int seq_len, int num_heads, int head_size,
    int max_seq_len, int int8_kv_cache, float q_scaling, b200_stream_t stream)
{
    B200_REQUIRE(qkv && out && (kv_cache || paged), B200_ERR_INVALID_ARG, "null pointer (qkv/out/kv_cache)");
    B200_REQUIRE(head_size == kDh, B200_ERR_UNSUPPORTED, "head_size %d unsupported (only 64)", head_size);
    B200_REQUIRE(batch_size >= 0 && seq_len >= 0 && num_heads > 0, B200_ERR_INVALID_ARG, "bad sizes");
    B200_REQUIRE(seq_len <= max_seq_len, B200_ERR_INVALID_ARG, "seq_len %d exceeds max_seq_len %d", seq_len, max_seq_len);
    B200_REQUIRE(!int8_kv_cache || kv_scale_orig_quant, B200_ERR_INVALID_ARG, "int8 KV cache needs the quant scale");
    B200_REQUIRE(q_scaling != 0.f, B200_ERR_INVALID_ARG, "q_scaling must be non-zero");
    if (batch_size == 0 || seq_len == 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    const size_t smem = (size_t) seq_len * kDh * 2 * sizeof(__half) + sizeof(float) * 4 * seq_len;
    B200_REQUIRE(smem <= 200 * 1024, B200_ERR_UNSUPPORTED, "context length %d too large for the single-kernel path", seq_len);
    const dim3 grid(num_heads, batch_size);
    const KvPaged pg = paged != nullptr ? *paged : KvPaged{nullptr, 0, 0};
    const __half* q = static_cast<const __half*>(qkv);
    __half* o = static_cast<__half*>(out);
    cudaStream_t st = as_stream(stream);
#define B200_CTX_LAUNCH(I8, PG)                                                                                        \
    do                                                                                                                 \
    {                                                                                                                  \
        if (smem > 48 * 1024)                                                                                          \
            B200_CUDA(cudaFuncSetAttribute((attention_context_kernel<I8, PG>), cudaFuncAttributeMaxDynamicSharedMemorySize, \
                (int) smem));                                                                                          \
        B200_LAUNCH((attention_context_kernel<I8, PG>), grid, dim3(128), smem, st, q, input_lengths, o, kv_cache,       \
            kv_scale_orig_quant, seq_len, num_heads, max_seq_len, q_scaling, pg);                                      \
    } while (0)
    if (paged != nullptr)
    {
        if (int8_kv_cache)
            B200_CTX_LAUNCH(true, true);
        else
            B200_CTX_LAUNCH(false, true);
    }
    else if (int8_kv_cache)
        B200_CTX_LAUNCH(true, false);
    else
        B200_CTX_LAUNCH(false, false);
#undef B200_CTX_LAUNCH
    return B200_OK;
}

extern "C" int b200_attention_context(const void* qkv, const int32_t* input_lengths, void* out, void* kv_cache,
    const float* kv_scale_orig_quant, int batch_size, int seq_len, int num_heads, int head_size, int max_seq_len,
    int int8_kv_cache, float q_scaling, b200_stream_t stream)
{
    return attention_context_impl(qkv, input_lengths, out, kv_cache, nullptr, kv_scale_orig_quant, batch_size, seq_len,
        num_heads, head_size, max_seq_len, int8_kv_cache, q_scaling, stream);
}

extern "C" int b200_attention_context_paged(const void* qkv, const int32_t* input_lengths, void* out,
    const void* const* block_pointers, int max_blocks_per_seq, int tokens_per_block, const float* kv_scale_orig_quant,
    int batch_size, int seq_len, int num_heads, int head_size, int max_seq_len, int int8_kv_cache, float q_scaling,
    b200_stream_t stream)
{
    KvPaged pg{};
    if (int rc = paged_view(block_pointers, max_blocks_per_seq, tokens_per_block, max_seq_len, &pg))
        return rc;
    return attention_context_impl(qkv, input_lengths, out, nullptr, &pg, kv_scale_orig_quant, batch_size, seq_len, num_heads,
        head_size, max_seq_len, int8_kv_cache, q_scaling, stream);
}

namespace b200
{
struct XaPlan
{
    int nch, chunks_per_warp, max_parts, blocks, cfg, rowhead;
    int split; // row-head kernel as clusters of two CTAs that share the pairs left over after the whole rounds
    size_t smem, ws_bytes;
};

static long long* g_xa_dbg = nullptr; // b200_debug_xa_timeline
// b200_set_cross_attention_split / env B200_XA_SPLIT: bit 0 = left-over pairs shared inside 2-CTA clusters, bit 1 = with few
// pairs (batch 1-3) every pair shared by a cluster of 4 / 2 CTAs.  Both default OFF: each shortens the kernel and lengthens the
// step (batch 16: 1.322 vs 1.293 ms with bit 0; batch 1: 0.991 vs 0.932 ms with bit 1) -- CTAs that own whole SMs, launched
// as clusters, start later and leave later, and the next GEMM's prologue pays for it (profiles/r02_xattn_timeline.txt)
static int g_xa_split = -1;
static int xa_split_default()
{
    const char* e = getenv("B200_XA_SPLIT");
    return e != nullptr ? (atoi(e) & 3) : 0;
}
static int g_xa_cfg = -1;  // env B200_XA_CFG = A .. F
static int g_xa_mode = -1; // env B200_XA_MODE = split | rowhead | auto (default)

static XaPlan xattn_plan(int R, int H, int S, int int8)
{
    if (g_xa_cfg < 0)
    {
        const char* e = getenv("B200_XA_CFG");
        g_xa_cfg = (e != nullptr && e[0] >= 'A' && e[0] <= 'F') ? e[0] - 'A' : 2;
    }
    static const int kW[6] = {XaCfgA::W, XaCfgB::W, XaCfgC::W, XaCfgD::W, XaCfgE::W, XaCfgF::W};
    static const int kST[6] = {XaCfgA::ST, XaCfgB::ST, XaCfgC::ST, XaCfgD::ST, XaCfgE::ST, XaCfgF::ST};
    static const int kCK[6] = {XaCfgA::CK, XaCfgB::CK, XaCfgC::CK, XaCfgD::CK, XaCfgE::CK, XaCfgF::CK};
    static const int kOCC[6] = {1, 1, 1, 1, 1, XaCfgF::OCC};
    const int W = kW[g_xa_cfg], ST = kST[g_xa_cfg], CKi = kCK[g_xa_cfg];
    if (g_xa_mode < 0)
    {
        const char* e = getenv("B200_XA_MODE");
        g_xa_mode = (e != nullptr && e[0] == 's') ? 0 : (e != nullptr && e[0] == 'r') ? 1 : 2;
    }
    XaPlan pl{};
    pl.cfg = g_xa_cfg;
    const int ck = int8 ? CKi : CKi / 2;
    pl.nch = (S + ck - 1) / ck;
    // whole (row, head) pairs per CTA: merge inside the CTA, no workspace
    // measured at batch 1-3 (20-60 pairs): whole pairs per CTA beat the split-across-CTAs kernel by 15-20 % of the step,
    // so the split kernel is only kept for a handful of pairs
    pl.rowhead = g_xa_mode == 2 ? ((long long) R * H >= 8 ? 1 : 0) : g_xa_mode;
    if (pl.rowhead)
    {
        const int slots = num_sms() * kOCC[g_xa_cfg];
        pl.blocks = R * H < slots ? R * H : slots;
        pl.smem = (size_t) W * ST * 2 * ck * kDh * (int8 ? 1 : 2) + ((sizeof(uint64_t) * W * ST + 15) & ~size_t(15))
            + sizeof(float) * 2 * W * (kDh + 4);
        pl.ws_bytes = 0;
        // 2-CTA clusters share the pairs that do not fill a whole round (at most one per cluster).  Opt-in
        // (b200_set_cross_attention_split / env B200_XA_SPLIT=1): the kernel itself ends 1.0 us earlier at batch 16, but the
        // step gets 0.9 us per layer SLOWER -- the CTAs that used to leave after two pairs are where the next GEMM starts
        // its weight stream and dequant ahead of its dependency (tools/xa_timeline.py)
        if (g_xa_split < 0)
            g_xa_split = xa_split_default();
        const int left = R * H - (R * H / pl.blocks) * pl.blocks;
        if ((g_xa_split & 1) && kOCC[g_xa_cfg] == 1 && pl.blocks % 2 == 0 && R * H >= pl.blocks && left > 0 && 2 * left <= pl.blocks
            && pl.nch >= 2)
        {
            pl.split = 2;
            pl.smem += sizeof(float) * 2 * W * (kDh + 4) + 16;
        }
        // few pairs (batch 1-3): every pair shared by a cluster of 4 (or 2) CTAs -- pairs x CS CTAs instead of one CTA per pair
        else if ((g_xa_split & 2) && kOCC[g_xa_cfg] == 1 && 2 * R * H <= slots && pl.nch >= 4)
        {
            pl.split = (4 * R * H <= slots && pl.nch >= 8) ? 4 : 2;
            pl.blocks = R * H * pl.split;
            pl.smem += sizeof(float) * pl.split * W * (kDh + 4) + 16;
        }
        return pl;
    }
    const long long total = (long long) R * H * pl.nch;
    // one persistent CTA per SM, every warp a contiguous run of chunks_per_warp chunks.  (Dealing the chunks evenly
    // to all 148 x W warps measured slower than this rounding, which leaves a few SMs free for the next kernel's
    // programmatic early launch.)
    const long long warps = (long long) num_sms() * W;
    pl.chunks_per_warp = (int) ((total + warps - 1) / warps);
    if (pl.chunks_per_warp < 1)
        pl.chunks_per_warp = 1;
    const long long used_warps = (total + pl.chunks_per_warp - 1) / pl.chunks_per_warp;
    pl.blocks = (int) ((used_warps + W - 1) / W);
    pl.max_parts = (pl.nch + pl.chunks_per_warp - 1) / pl.chunks_per_warp + 1;
    pl.smem = (size_t) W * ST * 2 * ck * kDh * (int8 ? 1 : 2) + sizeof(uint64_t) * W * ST;
    pl.ws_bytes = (size_t) R * H * pl.max_parts * (kDh + 2) * sizeof(float);
    return pl;
}

template <bool INT8, typename CFG>
static int xattn_launch(const XAttnParams& p, const XaPlan& pl, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(cross_attention_kernel<INT8, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pl.smem));
        attr_set = true;
    }
    B200_LAUNCH((cross_attention_kernel<INT8, CFG>), dim3(pl.blocks), dim3(CFG::W * 32), pl.smem, st, p);
    return B200_OK;
}

template <bool INT8, typename CFG, int CS>
static int xattn_launch_rowhead_split(const XAttnParams& p, const XaPlan& pl, cudaStream_t st)
{
    auto kern = cross_attention_rowhead_kernel<INT8, CFG, CS>;
    static size_t attr_smem = 0;
    if (pl.smem > attr_smem)
    {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pl.smem));
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_smem = pl.smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pl.blocks);
    cfg.blockDim = dim3(CFG::W * 32);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl_enabled())
    {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CS;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    count_launch();
    B200_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return B200_OK;
}

template <bool INT8, typename CFG>
static int xattn_launch_rowhead(const XAttnParams& p, const XaPlan& pl, cudaStream_t st)
{
    if (pl.split)
    {
        if constexpr (CFG::OCC == 1)
            return pl.split == 4 ? xattn_launch_rowhead_split<INT8, CFG, 4>(p, pl, st)
                                 : xattn_launch_rowhead_split<INT8, CFG, 2>(p, pl, st);
    }
    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(cross_attention_rowhead_kernel<INT8, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pl.smem));
        if (getenv("B200_XA_NOCARVE") == nullptr)
            B200_CUDA(cudaFuncSetAttribute(cross_attention_rowhead_kernel<INT8, CFG>, cudaFuncAttributePreferredSharedMemoryCarveout,
                cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    B200_LAUNCH((cross_attention_rowhead_kernel<INT8, CFG>), dim3(pl.blocks), dim3(CFG::W * 32), pl.smem, st, p);
    return B200_OK;
}

int* tc_counter_slot(int needed, cudaStream_t stream);
} // namespace b200

extern "C" size_t b200_cross_attention_workspace_bytes(int batch_size, int num_heads, int head_size, int kv_len)
{
    if (batch_size <= 0 || num_heads <= 0 || kv_len <= 0 || head_size != kDh)
        return 0;
    const size_t a = xattn_plan(batch_size, num_heads, kv_len, 1).ws_bytes;
    const size_t b = xattn_plan(batch_size, num_heads, kv_len, 0).ws_bytes;
    return a > b ? a : b;
}

extern "C" int b200_cross_attention(const void* q, const void* cross_kv, const float* kv_scale_quant_orig, void* out,
    int batch_size, int q_rows_per_seq, int num_heads, int head_size, int kv_len, int int8_kv_cache, void* workspace,
    size_t workspace_bytes, b200_stream_t stream)
{
    B200_REQUIRE(q_rows_per_seq >= 1 && batch_size % q_rows_per_seq == 0, B200_ERR_INVALID_ARG,
        "q_rows_per_seq=%d must divide the number of query rows %d", q_rows_per_seq, batch_size);
    B200_REQUIRE(q && cross_kv && out, B200_ERR_INVALID_ARG, "null pointer (q/cross_kv/out)");
    B200_REQUIRE(head_size == kDh, B200_ERR_UNSUPPORTED, "head_size %d unsupported (only 64)", head_size);
    B200_REQUIRE(batch_size >= 0 && num_heads > 0 && kv_len > 0, B200_ERR_INVALID_ARG, "bad sizes");
    B200_REQUIRE(!int8_kv_cache || kv_scale_quant_orig, B200_ERR_INVALID_ARG, "int8 cross-KV needs the dequant scale");
    if (batch_size == 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    XAttnParams p{};
    p.q = static_cast<const __half*>(q);
    p.kv = cross_kv;
    p.scale_quant_orig = kv_scale_quant_orig;
    p.out = static_cast<__half*>(out);
    p.partials = static_cast<float*>(workspace);
    p.B = batch_size;
    p.H = num_heads;
    p.S = kv_len;
    p.q_per_seq = q_rows_per_seq;
    const XaPlan pl = xattn_plan(batch_size, num_heads, kv_len, int8_kv_cache);
    p.nch = pl.nch;
    p.chunks_per_warp = pl.chunks_per_warp;
    p.max_parts = pl.max_parts;
    p.early_kv = static_kv_hint() ? 1 : 0;
    p.inv_sqrt_dh = 1.f / sqrtf((float) kDh);
    p.dbg = g_xa_dbg;
    cudaStream_t st = as_stream(stream);
    if (pl.rowhead)
    {
        switch (pl.cfg * 2 + (int8_kv_cache ? 1 : 0))
        {
        case 0: return xattn_launch_rowhead<false, XaCfgA>(p, pl, st);
        case 1: return xattn_launch_rowhead<true, XaCfgA>(p, pl, st);
        case 2: return xattn_launch_rowhead<false, XaCfgB>(p, pl, st);
        case 3: return xattn_launch_rowhead<true, XaCfgB>(p, pl, st);
        case 4: return xattn_launch_rowhead<false, XaCfgC>(p, pl, st);
        case 5: return xattn_launch_rowhead<true, XaCfgC>(p, pl, st);
        case 6: return xattn_launch_rowhead<false, XaCfgD>(p, pl, st);
        case 7: return xattn_launch_rowhead<true, XaCfgD>(p, pl, st);
        case 8: return xattn_launch_rowhead<false, XaCfgE>(p, pl, st);
        case 9: return xattn_launch_rowhead<true, XaCfgE>(p, pl, st);
        case 10: return xattn_launch_rowhead<false, XaCfgF>(p, pl, st);
        default: return xattn_launch_rowhead<true, XaCfgF>(p, pl, st);
        }
    }
    B200_REQUIRE(workspace && workspace_bytes >= pl.ws_bytes, B200_ERR_WORKSPACE,
        "cross attention: workspace of %zu bytes needed, got %zu", pl.ws_bytes, workspace_bytes);
    p.counters = tc_counter_slot(batch_size * num_heads, st);
    B200_REQUIRE(p.counters != nullptr, B200_ERR_UNSUPPORTED, "cross attention: %d (row, head) pairs exceed the counter slot",
        batch_size * num_heads);
    switch (pl.cfg * 2 + (int8_kv_cache ? 1 : 0))
    {
    case 0: return xattn_launch<false, XaCfgA>(p, pl, st);
    case 1: return xattn_launch<true, XaCfgA>(p, pl, st);
    case 2: return xattn_launch<false, XaCfgB>(p, pl, st);
    case 3: return xattn_launch<true, XaCfgB>(p, pl, st);
    case 4: return xattn_launch<false, XaCfgC>(p, pl, st);
    case 5: return xattn_launch<true, XaCfgC>(p, pl, st);
    case 6: return xattn_launch<false, XaCfgD>(p, pl, st);
    case 7: return xattn_launch<true, XaCfgD>(p, pl, st);
    case 8: return xattn_launch<false, XaCfgE>(p, pl, st);
    case 9: return xattn_launch<true, XaCfgE>(p, pl, st);
    case 10: return xattn_launch<false, XaCfgF>(p, pl, st);
    default: return xattn_launch<true, XaCfgF>(p, pl, st);
    }
}

namespace b200
{
// launch plan of the q-fused kernel: grid, rows per CTA, shared memory; ok = 0 when the shape is not handled
struct XqPlan
{
    int ok, blocks, maxr, nch;
    size_t smem;
};

static XqPlan xq_plan(int B, int H, int S)
{
    using CFG = XaCfgC;
    XqPlan pl{};
    if (B < 1 || H < 1 || S < 1 || H > kXqMaxKb * CFG::W)
        return pl;
    const int sms = num_sms();
    if ((long long) B * H < sms) // fewer pairs than SMs: the plain kernels (whole pairs or split pairs per CTA) serve these
        return pl;
    pl.blocks = sms;
    const int nc_min = sms / H; // CTAs of the least-served head
    pl.maxr = (B + nc_min - 1) / nc_min;
    if (pl.maxr * kDh > CFG::W * 32) // one finishing thread per (row, column)
        return pl;
    pl.nch = (S + CFG::CK - 1) / CFG::CK;
    pl.smem = (size_t) CFG::W * CFG::ST * 2 * CFG::CK * kDh + ((sizeof(uint64_t) * CFG::W * CFG::ST + 15) & ~size_t(15))
        + sizeof(float) * 2 * CFG::W * (kDh + 4) + sizeof(float) * CFG::W * pl.maxr * (kDh + 2) + sizeof(float) * ((pl.maxr + 3) & ~3) + sizeof(__half) * pl.maxr * kDh;
    pl.ok = pl.smem <= 227 * 1024 ? 1 : 0;
    return pl;
}
} // namespace b200

extern "C" int b200_cross_attention_qproj_supported(int batch_size, int num_heads, int head_size, int kv_len)
{
    if (head_size != kDh)
        return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
        return 0;
    return xq_plan(batch_size, num_heads, kv_len).ok;
}

extern "C" int b200_cross_attention_qproj(const void* x, const void* ln_gamma, const float* c1s, const float* c2, float ln_eps,
    const int8_t* Wproc, const void* scales, const void* bias, const void* cross_kv, const float* kv_scale_quant_orig,
    void* out, int batch_size, int num_heads, int head_size, int kv_len, b200_stream_t stream)
{
    B200_REQUIRE(x && ln_gamma && c1s && c2 && Wproc && scales && cross_kv && kv_scale_quant_orig && out, B200_ERR_INVALID_ARG,
        "null pointer");
    B200_REQUIRE(head_size == kDh, B200_ERR_UNSUPPORTED, "head_size %d unsupported (only 64)", head_size);
    B200_REQUIRE_DEVICE();
    const XqPlan pl = xq_plan(batch_size, num_heads, kv_len);
    B200_REQUIRE(pl.ok, B200_ERR_UNSUPPORTED, "q-fused cross attention: batch %d x %d heads x %d keys not handled (use "
        "b200_woq_int8_gemm_ln_folded + b200_cross_attention)", batch_size, num_heads, kv_len);
    using CFG = XaCfgC;
    XqParams p{};
    p.x = static_cast<const __half*>(x);
    p.W = reinterpret_cast<const uint8_t*>(Wproc);
    p.scales = static_cast<const __half*>(scales);
    p.bias = static_cast<const __half*>(bias);
    p.gamma = static_cast<const __half*>(ln_gamma);
    p.c1s = c1s, p.c2 = c2;
    p.kv = cross_kv;
    p.scale_quant_orig = kv_scale_quant_orig;
    p.out = static_cast<__half*>(out);
    p.B = batch_size, p.H = num_heads, p.S = kv_len, p.nch = pl.nch, p.maxr = pl.maxr;
    p.eps = ln_eps;
    p.inv_sqrt_dh = 1.f / sqrtf((float) kDh);
    static size_t attr_smem = 0;
    if (pl.smem > attr_smem)
    {
        B200_CUDA(cudaFuncSetAttribute(cross_attention_qproj_kernel<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B200_CUDA(cudaFuncSetAttribute(cross_attention_qproj_kernel<CFG>, cudaFuncAttributePreferredSharedMemoryCarveout,
            cudaSharedmemCarveoutMaxShared));
        attr_smem = 227 * 1024;
    }
    B200_LAUNCH((cross_attention_qproj_kernel<CFG>), dim3(pl.blocks), dim3(CFG::W * 32), pl.smem, as_stream(stream), p);
    return B200_OK;
}

/* Tuning switch (process-wide, set before the launches are captured; returns the previous value): 1 = the whole-pair
 * cross-attention kernel runs as clusters of two CTAs that share the (row, head) pairs left over after the whole rounds
 * of the grid.  Default 0 (env B200_XA_SPLIT=1 turns it on). */
extern "C" int b200_set_cross_attention_split(int enabled)
{
    if (g_xa_split < 0)
        g_xa_split = xa_split_default();
    const int prev = g_xa_split;
    g_xa_split = enabled & 3;
    return prev;
}

/* Debug aid (B200_XA_DEBUG=1 builds only; a no-op otherwise): device buffer of >= 8 int64 per SM receiving %globaltimer
 * stamps of every CTA of the following row-head cross-attention launches: entry, dependency return, warp 0 done with
 * pair 0 / 1 / 2 / 3 / later, exit.  NULL switches it off. */
extern "C" int b200_debug_xa_timeline(void* device_buffer)
{
    g_xa_dbg = static_cast<long long*>(device_buffer);
    return B200_OK;
}

extern "C" int b200_cross_kv_pack(const void* k, const void* v, void* cross_kv, const float* kv_scale_orig_quant,
    int batch_size, int kv_len, int num_heads, int head_size, int int8_kv_cache, b200_stream_t stream)
{
    B200_REQUIRE(k && v && cross_kv, B200_ERR_INVALID_ARG, "null pointer (k/v/cross_kv)");
    B200_REQUIRE(head_size == kDh, B200_ERR_UNSUPPORTED, "head_size %d unsupported (only 64)", head_size);
    B200_REQUIRE(!int8_kv_cache || kv_scale_orig_quant, B200_ERR_INVALID_ARG, "int8 cross-KV needs the quant scale");
    if (batch_size <= 0 || kv_len <= 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    const dim3 grid(kv_len, batch_size);
    if (int8_kv_cache)
        cross_kv_pack_kernel<true><<<grid, 160, 0, as_stream(stream)>>>(static_cast<const __half*>(k),
            static_cast<const __half*>(v), cross_kv, kv_scale_orig_quant, kv_len, num_heads);
    else
        cross_kv_pack_kernel<false><<<grid, 160, 0, as_stream(stream)>>>(static_cast<const __half*>(k),
            static_cast<const __half*>(v), cross_kv, kv_scale_orig_quant, kv_len, num_heads);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
