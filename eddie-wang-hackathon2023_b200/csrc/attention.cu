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
// Deliberate deviation (more accurate, inside the reference tests' tolerances): probabilities stay fp32 instead of
// being rounded to fp16 before P.V (Template.h:1765).
//
// Cache layout: [B, 2, H, Smax, Dh] (KVLinearBuffer, T/cpp/tensorrt_llm/kernels/kvCacheUtils.h:114-170). Dh = 64.
#include <float.h>

#include "common.cuh"

namespace b200
{

constexpr int kDh = 64;

__device__ __forceinline__ int8_t quant_s8(float v)
{
    int32_t r;
    asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return static_cast<int8_t>(r);
}

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

template <bool INT8>
__device__ __forceinline__ void store16(void* base, size_t elem_off, float scale_orig_quant, const __half (&x)[16])
{
    if constexpr (INT8)
    {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                v |= (static_cast<uint32_t>(static_cast<uint8_t>(quant_s8(scale_orig_quant * __half2float(x[4 * i + b]))))
                    << (8 * b));
            w[i] = v;
        }
        *reinterpret_cast<uint4*>(static_cast<int8_t*>(base) + elem_off) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    else
    {
        uint4* p = reinterpret_cast<uint4*>(static_cast<__half*>(base) + elem_off);
        p[0] = *reinterpret_cast<const uint4*>(&x[0]);
        p[1] = *reinterpret_cast<const uint4*>(&x[8]);
    }
}

__device__ __forceinline__ void load16_half(const __half* src, const __half* bias, __half (&x)[16])
{
    const uint4 v0 = *reinterpret_cast<const uint4*>(src);
    const uint4 v1 = *reinterpret_cast<const uint4*>(src + 8);
    *reinterpret_cast<uint4*>(&x[0]) = v0;
    *reinterpret_cast<uint4*>(&x[8]) = v1;
    if (bias != nullptr)
    {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            x[i] = __hadd(x[i], bias[i]); // fp16 add, like add(q, q_bias) at Template.h:1406-1407
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
// Generation step.  grid (H, B), 128 threads.  Lane geometry: 4 lanes per key (16 dims each), 8 keys per warp
// instruction, so a warp reads 512 contiguous cache bytes (int8) per step.
// =====================================================================================================
template <bool INT8>
__global__ void __launch_bounds__(128) mmha_generation_kernel(const b200_mmha_params p)
{
    extern __shared__ float s_qk[]; // [Smax + 1]
    __shared__ float s_red[4];
    __shared__ float s_out[4][kDh];

    const int h = blockIdx.x, b = blockIdx.y;
    const int H = p.num_heads, Smax = p.max_seq_len;
    const int hidden = H * kDh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = lane & 3, kl = lane >> 2;

    grid_dep_wait();
    grid_dep_launch_dependents();

    int tlen = p.sequence_lengths ? p.sequence_lengths[b] : p.past_kv_length;
    tlen = min(tlen, Smax - 1);
    const float inv_sqrt_dh = 1.f / (sqrtf((float) kDh) * p.q_scaling); // gptAttentionCommon.cpp:163
    const float s_qo = INT8 ? p.kv_scale_quant_orig[0] : 1.f;
    const float s_oq = INT8 ? p.kv_scale_orig_quant[0] : 1.f;

    const __half* qkv = static_cast<const __half*>(p.qkv) + (size_t) b * 3 * hidden + h * kDh + chunk * 16;
    const __half* bias = p.qkv_bias ? static_cast<const __half*>(p.qkv_bias) + h * kDh + chunk * 16 : nullptr;
    __half qh[16], kh[16], vh[16];
    load16_half(qkv, bias, qh);
    load16_half(qkv + hidden, bias ? bias + hidden : nullptr, kh);
    load16_half(qkv + 2 * hidden, bias ? bias + 2 * hidden : nullptr, vh);
    float q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        q[i] = __half2float(qh[i]);

    const size_t esz = INT8 ? 1 : 2;
    char* kc = static_cast<char*>(p.kv_cache) + ((size_t) (b * 2 + 0) * H + h) * Smax * kDh * esz;
    char* vc = static_cast<char*>(p.kv_cache) + ((size_t) (b * 2 + 1) * H + h) * Smax * kDh * esz;

    // append this step's K and V (warp 0: lanes 0-3 write K chunks, lanes 4-7 write V chunks)
    if (warp == 0 && lane < 8)
    {
        if (lane < 4)
            store16<INT8>(kc, (size_t) tlen * kDh + chunk * 16, s_oq, kh);
        else
            store16<INT8>(vc, (size_t) tlen * kDh + chunk * 16, s_oq, vh);
    }

    // ---- q.k over the cache ----
    float lmax = -FLT_MAX;
    const int* mask = p.masked_tokens ? p.masked_tokens + (size_t) b * Smax : nullptr;
    for (int t0 = warp * 8; t0 < tlen; t0 += 32)
    {
        const int t = t0 + kl;
        float s = 0.f;
        if (t < tlen)
        {
            float kf[16];
            load16<INT8>(kc, (size_t) t * kDh + chunk * 16, s_qo, kf);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                s = fmaf(q[i], kf[i], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (t < tlen && chunk == 0)
        {
            s *= inv_sqrt_dh;
            s_qk[t] = s;
            if (!(mask && mask[t]))
                lmax = fmaxf(lmax, s);
        }
    }
    // current token (unquantized k)
    if (warp == 0)
    {
        float s = 0.f;
        if (lane < 4)
        {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                s = fmaf(q[i], __half2float(kh[i]), s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (lane == 0)
        {
            s *= inv_sqrt_dh;
            s_qk[tlen] = s;
            lmax = fmaxf(lmax, s);
        }
    }
    const float gmax = block_reduce_max(lmax, s_red, 4);

    // ---- softmax numerators ----
    float lsum = 0.f;
    for (int t = tid; t <= tlen; t += 128)
    {
        const bool masked = (t < tlen) && mask && mask[t];
        const float e = masked ? 0.f : __expf(s_qk[t] - gmax);
        s_qk[t] = e;
        lsum += e;
    }
    const float gsum = block_reduce_sum(lsum, s_red, 4); // contains a __syncthreads: s_qk is complete
    const float inv_sum = __fdividef(1.f, gsum + 1.e-6f);

    // ---- p.v ----
    float o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        o[i] = 0.f;
    for (int t0 = warp * 8; t0 < tlen; t0 += 32)
    {
        const int t = t0 + kl;
        if (t < tlen)
        {
            const float pt = s_qk[t];
            float vf[16];
            load16<INT8>(vc, (size_t) t * kDh + chunk * 16, s_qo, vf);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                o[i] = fmaf(pt, vf[i], o[i]);
        }
    }
    if (warp == 0 && lane < 4)
    {
        const float pt = s_qk[tlen];
#pragma unroll
        for (int i = 0; i < 16; ++i)
            o[i] = fmaf(pt, __half2float(vh[i]), o[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        float v = o[i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        o[i] = v;
    }
    if (lane < 4)
    {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            s_out[warp][chunk * 16 + i] = o[i];
    }
    __syncthreads();
    if (tid < kDh)
    {
        const float v = (s_out[0][tid] + s_out[1][tid]) + (s_out[2][tid] + s_out[3][tid]);
        static_cast<__half*>(p.out)[(size_t) b * hidden + h * kDh + tid] = __float2half_rn(v * inv_sum);
    }
}

// =====================================================================================================
// Context phase.  grid (H, B), 128 threads; K and V of the whole prompt for this (b, h) are staged in shared
// memory as fp16 (unquantized values are used for the attention itself, as the reference does), the cache rows
// [0, S) are written (int8-quantized when requested), and each warp handles query rows i = warp, warp+4, ...
// causal: key j is visible to query i iff j <= i and j < input_length[b].
// =====================================================================================================
template <bool INT8>
__global__ void __launch_bounds__(128) attention_context_kernel(const __half* __restrict__ qkv,
    const int* __restrict__ input_lengths, __half* __restrict__ out, void* __restrict__ kv_cache,
    const float* __restrict__ kv_scale_orig_quant, int S, int H, int Smax, float q_scaling)
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
    char* kc = static_cast<char*>(kv_cache) + ((size_t) (b * 2 + 0) * H + h) * Smax * kDh * esz;
    char* vc = static_cast<char*>(kv_cache) + ((size_t) (b * 2 + 1) * H + h) * Smax * kDh * esz;

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
        store16<INT8>(kc, (size_t) t * kDh + c * 16, s_oq, kh);
        store16<INT8>(vc, (size_t) t * kDh + c * 16, s_oq, vh);
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
// decoder step at batch >= 6 (3.84 MB per sequence per layer).  HBM-bound streaming design, third iteration
// (v1: CTA-wide items with shared-memory scores, 6.7 thread-instructions per byte, issue-bound at 2.1 TB/s;
//  v2: finer items + in-kernel merge, slower: five CTA barriers and a __threadfence per 16 KB item):
//   * the work item is WARP-private: (query row, head, range of <= 128 keys).  A warp reads its K range with sixteen
//     independent 128-bit loads per lane (512 contiguous bytes per warp instruction, 8 KB in flight per warp), keeps
//     the 16 scores per lane in registers, does the softmax with shuffles only, then streams V the same way.
//     No shared memory, no CTA barrier, no producer/consumer handshake: with 16 resident warps per SM there are
//     128 KB of loads in flight per SM, several times what Little's law needs for HBM3e.
//   * lane geometry: 4 lanes x 16 dims per key, 8 keys per warp instruction (the same lane owns the same key in the
//     K and V phases, so probabilities never leave registers).
//   * int8 -> fp16 by xor 0x80 + PRMT + HSUB2 (exact integers); products chained four at a time with HFMA2 and
//     flushed to fp32 (same scheme as the GEMV); the dequant scale is hoisted out of both dot products.
//   * the splits of one (row, head) are merged by the last warp to arrive (self-resetting counter).
// =====================================================================================================
constexpr int kXaWarps = 8;

struct XAttnParams
{
    const __half* q;   // [R, H*64]
    const void* kv;    // [B, 2, H, S, 64]
    const float* scale_quant_orig;
    __half* out;       // [R, H*64]
    float* partials;   // [R*H*nsplit][66]
    int* counters;     // [R*H] arrival counters (library owned, self-resetting)
    int B, H, S;       // B = number of query rows R
    int q_per_seq;     // query rows per cache sequence (1 in the generation phase, S_prompt in the context phase)
    int nsplit, keys_per_split;
    int items_per_warp;
    float inv_sqrt_dh;
};

// 16 cache bytes (or 16 fp16) of one key -> 8 half2 in the pair order (d0,d2) (d1,d3) (d4,d6) (d5,d7) ...
template <bool INT8>
struct XaChunk;

template <>
struct XaChunk<true>
{
    uint4 v;

    __device__ __forceinline__ void load(const uint8_t* p)
    {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(p));
    }

    __device__ __forceinline__ void zero()
    {
        v = make_uint4(0, 0, 0, 0); // int8 zeros
    }

    __device__ __forceinline__ void unpack(__half2 (&w)[8]) const
    {
        dequant_word(v.x ^ 0x80808080u, w[0], w[1]);
        dequant_word(v.y ^ 0x80808080u, w[2], w[3]);
        dequant_word(v.z ^ 0x80808080u, w[4], w[5]);
        dequant_word(v.w ^ 0x80808080u, w[6], w[7]);
    }
};

template <>
struct XaChunk<false>
{
    uint4 v0, v1;

    __device__ __forceinline__ void load(const uint8_t* p)
    {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v0.x), "=r"(v0.y), "=r"(v0.z), "=r"(v0.w)
                     : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v1.x), "=r"(v1.y), "=r"(v1.z), "=r"(v1.w)
                     : "l"(p + 16));
    }

    __device__ __forceinline__ void zero()
    {
        v0 = v1 = make_uint4(0, 0, 0, 0);
    }

    __device__ __forceinline__ void unpack(__half2 (&w)[8]) const
    {
        const uint32_t u[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w}; // u[j] = (d2j, d2j+1)
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const uint32_t lo = __byte_perm(u[2 * i], u[2 * i + 1], 0x5410); // (d4i, d4i+2)
            const uint32_t hi = __byte_perm(u[2 * i], u[2 * i + 1], 0x7632); // (d4i+1, d4i+3)
            w[2 * i] = *reinterpret_cast<const __half2*>(&lo);
            w[2 * i + 1] = *reinterpret_cast<const __half2*>(&hi);
        }
    }
};

template <bool INT8>
__global__ void __launch_bounds__(kXaWarps * 32) cross_attention_kernel(const XAttnParams p)
{
    constexpr int ESZ = INT8 ? 1 : 2;
    constexpr int NIT = INT8 ? 16 : 8; // 8 keys per iteration: <= 128 (int8) / 64 (fp16) keys per item
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = lane & 3, kl = lane >> 2;
    const int items = p.B * p.H * p.nsplit;
    const int gw = blockIdx.x * kXaWarps + warp;

    grid_dep_launch_dependents();
    const float s_qo = INT8 ? p.scale_quant_orig[0] : 1.f;
    const float sscale = s_qo * p.inv_sqrt_dh;
    bool waited = false;

    for (int it_w = 0; it_w < p.items_per_warp; ++it_w)
    {
        const int item = gw * p.items_per_warp + it_w;
        if (item >= items)
            break;
        const int bh = item / p.nsplit, sp = item % p.nsplit;
        const int b = (bh / p.H) / p.q_per_seq, h = bh % p.H;
        const int key0 = sp * p.keys_per_split;
        const int nkeys = min(p.keys_per_split, p.S - key0);
        const uint8_t* kbase = static_cast<const uint8_t*>(p.kv)
            + ((((size_t) (b * 2 + 0) * p.H + h) * p.S + key0) * kDh + chunk * 16) * ESZ;
        const uint8_t* vbase = static_cast<const uint8_t*>(p.kv)
            + ((((size_t) (b * 2 + 1) * p.H + h) * p.S + key0) * kDh + chunk * 16) * ESZ;

        // ---- K: all loads of the item in flight before anything depends on them (cache data: no PDL wait needed)
        XaChunk<INT8> kc[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            const int key = it * 8 + kl;
            if (key < nkeys)
                kc[it].load(kbase + (size_t) key * kDh * ESZ);
            else
                kc[it].zero();
        }
        if (!waited)
        {
            grid_dep_wait(); // q comes from the previous kernel
            waited = true;
        }
        __half2 q2[8];
        {
            __half qh[16];
            load16_half(p.q + (size_t) bh * kDh + chunk * 16, nullptr, qh);
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                q2[2 * i] = __halves2half2(qh[4 * i], qh[4 * i + 2]);
                q2[2 * i + 1] = __halves2half2(qh[4 * i + 1], qh[4 * i + 3]);
            }
        }
        float sc[NIT];
        float m = -FLT_MAX;
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            __half2 w[8];
            kc[it].unpack(w);
            __half2 h0 = __hmul2(q2[0], w[0]);
            __half2 h1 = __hmul2(q2[4], w[4]);
            h0 = __hfma2(q2[1], w[1], h0);
            h1 = __hfma2(q2[5], w[5], h1);
            h0 = __hfma2(q2[2], w[2], h0);
            h1 = __hfma2(q2[6], w[6], h1);
            h0 = __hfma2(q2[3], w[3], h0);
            h1 = __hfma2(q2[7], w[7], h1);
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
            float s = (f0.x + f0.y) + (f1.x + f1.y);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s = (it * 8 + kl < nkeys) ? s * sscale : -FLT_MAX;
            sc[it] = s;
            m = fmaxf(m, s);
        }
        // ---- V loads go out now; the softmax below overlaps their latency
        XaChunk<INT8> vc[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            const int key = it * 8 + kl;
            if (key < nkeys)
                vc[it].load(vbase + (size_t) key * kDh * ESZ);
            else
                vc[it].zero();
        }
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        float l = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            const float e = (it * 8 + kl < nkeys) ? __expf(sc[it] - m) : 0.f;
            sc[it] = e;
            l += e;
        }
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);

        // ---- P.V: four keys chained in fp16, then flushed to fp32
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
            o[i] = 0.f;
#pragma unroll
        for (int it4 = 0; it4 < NIT; it4 += 4)
        {
            __half2 o2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                o2[i] = __float2half2_rn(0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const __half2 p2 = __float2half2_rn(sc[it4 + j]);
                __half2 w[8];
                vc[it4 + j].unpack(w);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    o2[i] = __hfma2(p2, w[i], o2[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const float2 f = __half22float2(o2[i]);
                o[2 * i] += f.x;
                o[2 * i + 1] += f.y;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
        {
            float v = o[i];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            o[i] = v * s_qo; // hoisted V dequant scale
        }
        // o[2i], o[2i+1] hold the pair of w[i]: w[2j] = dims (4j, 4j+2), w[2j+1] = dims (4j+1, 4j+3)
        if (p.nsplit == 1)
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
            continue;
        }
        float* pr = p.partials + (size_t) item * (kDh + 2);
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
                __stcg(pr, m);
                __stcg(pr + 1, l);
            }
        }
        // last split of this (row, head) to arrive merges all partials
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0)
            last = (atomicAdd(&p.counters[bh], 1) == p.nsplit - 1) ? 1 : 0;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last)
        {
            __threadfence();
            const float* pb = p.partials + (size_t) bh * p.nsplit * (kDh + 2);
            float gm = -FLT_MAX;
            for (int s2 = 0; s2 < p.nsplit; ++s2)
                gm = fmaxf(gm, __ldcg(pb + s2 * (kDh + 2)));
            float gl = 0.f, a0 = 0.f, a1 = 0.f;
            for (int s2 = 0; s2 < p.nsplit; ++s2)
            {
                const float* ps = pb + s2 * (kDh + 2);
                const float w = __expf(__ldcg(ps) - gm);
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
        store16<INT8>(cache, off, s_oq, x);
    }
}

} // namespace b200

using namespace b200;

extern "C" int b200_mmha_generation(const b200_mmha_params* p, b200_stream_t stream)
{
    B200_REQUIRE(p != nullptr, B200_ERR_INVALID_ARG, "null params");
    B200_REQUIRE(p->qkv && p->out && p->kv_cache, B200_ERR_INVALID_ARG, "null pointer (qkv/out/kv_cache)");
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
    const dim3 grid(p->num_heads, p->batch_size);
    const size_t smem = sizeof(float) * (p->max_seq_len + 1);
    B200_REQUIRE(smem <= 200 * 1024, B200_ERR_UNSUPPORTED, "max_seq_len %d too large", p->max_seq_len);
    if (p->int8_kv_cache)
    {
        if (smem > 48 * 1024)
            B200_CUDA(cudaFuncSetAttribute(mmha_generation_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        B200_LAUNCH(mmha_generation_kernel<true>, grid, dim3(128), smem, as_stream(stream), *p);
    }
    else
    {
        if (smem > 48 * 1024)
            B200_CUDA(cudaFuncSetAttribute(mmha_generation_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        B200_LAUNCH(mmha_generation_kernel<false>, grid, dim3(128), smem, as_stream(stream), *p);
    }
    return B200_OK;
}

extern "C" int b200_attention_context(const void* qkv, const int32_t* input_lengths, void* out, void* kv_cache,
    const float* kv_scale_orig_quant, int batch_size, int seq_len, int num_heads, int head_size, int max_seq_len,
    int int8_kv_cache, float q_scaling, b200_stream_t stream)
{
    B200_REQUIRE(qkv && out && kv_cache, B200_ERR_INVALID_ARG, "null pointer (qkv/out/kv_cache)");
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
    if (int8_kv_cache)
    {
        if (smem > 48 * 1024)
            B200_CUDA(cudaFuncSetAttribute(attention_context_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        B200_LAUNCH(attention_context_kernel<true>, grid, dim3(128), smem, as_stream(stream), static_cast<const __half*>(qkv),
            input_lengths, static_cast<__half*>(out), kv_cache, kv_scale_orig_quant, seq_len, num_heads, max_seq_len, q_scaling);
    }
    else
    {
        if (smem > 48 * 1024)
            B200_CUDA(cudaFuncSetAttribute(attention_context_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        B200_LAUNCH(attention_context_kernel<false>, grid, dim3(128), smem, as_stream(stream), static_cast<const __half*>(qkv),
            input_lengths, static_cast<__half*>(out), kv_cache, kv_scale_orig_quant, seq_len, num_heads, max_seq_len, q_scaling);
    }
    return B200_OK;
}

namespace b200
{
static void xattn_plan(int B, int H, int S, int int8, int& nsplit, int& kps, int& items_per_warp, int& blocks)
{
    // warp-private items of <= 128 (int8) / 64 (fp16) keys; every warp gets the same number of items
    const int max_keys = int8 ? 128 : 64;
    const int pairs = B * H;
    const int slots = num_sms() * 2 * kXaWarps; // two 8-warp CTAs per SM are resident (register-limited)
    nsplit = (S + max_keys - 1) / max_keys;
    // when there are fewer items than warp slots, split finer (down to 32 keys) so more SMs pull bytes
    while (pairs * nsplit * 2 <= slots && (S + nsplit * 2 - 1) / (nsplit * 2) >= 32)
        nsplit *= 2;
    kps = (S + nsplit - 1) / nsplit;
    kps = (kps + 7) & ~7; // whole 8-key warp iterations
    nsplit = (S + kps - 1) / kps;
    const int items = pairs * nsplit;
    items_per_warp = (items + slots - 1) / slots;
    blocks = (items + items_per_warp * kXaWarps - 1) / (items_per_warp * kXaWarps);
}

int* tc_counter_slot(int needed);
} // namespace b200

extern "C" size_t b200_cross_attention_workspace_bytes(int batch_size, int num_heads, int head_size, int kv_len)
{
    if (batch_size <= 0 || num_heads <= 0 || kv_len <= 0 || head_size != kDh)
        return 0;
    int ns, kps, ipw, blocks, best = 0;
    for (int int8 = 0; int8 < 2; ++int8)
    {
        xattn_plan(batch_size, num_heads, kv_len, int8, ns, kps, ipw, blocks);
        best = ns > best ? ns : best;
    }
    return (size_t) batch_size * num_heads * best * (kDh + 2) * sizeof(float);
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
    int blocks = 1;
    xattn_plan(batch_size, num_heads, kv_len, int8_kv_cache, p.nsplit, p.keys_per_split, p.items_per_warp, blocks);
    p.inv_sqrt_dh = 1.f / sqrtf((float) kDh);
    const size_t need = p.nsplit > 1 ? (size_t) batch_size * num_heads * p.nsplit * (kDh + 2) * sizeof(float) : 0;
    B200_REQUIRE(need == 0 || (workspace && workspace_bytes >= need), B200_ERR_WORKSPACE,
        "cross attention: workspace of %zu bytes needed, got %zu", need, workspace_bytes);
    if (p.nsplit > 1)
    {
        p.counters = tc_counter_slot(batch_size * num_heads);
        B200_REQUIRE(p.counters != nullptr, B200_ERR_UNSUPPORTED, "cross attention: %d (row, head) pairs exceed the counter slot",
            batch_size * num_heads);
    }
    cudaStream_t st = as_stream(stream);
    if (int8_kv_cache)
        B200_LAUNCH(cross_attention_kernel<true>, dim3(blocks), dim3(kXaWarps * 32), 0, st, p);
    else
        B200_LAUNCH(cross_attention_kernel<false>, dim3(blocks), dim3(kXaWarps * 32), 0, st, p);
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
