// attn_device.cuh -- device helpers shared by the attention kernels (attention.cu) and the persistent decoder-step
// kernel (decoder_step.cu): int8 KV quantization / conversion, the mma.sync score helper and the streaming
// cross-attention chunk (scores on the tensor cores, online softmax in the log2 domain, p.v on the fp16 pipe).
// Reference semantics: see the header comment of attention.cu.
#pragma once

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

// XORV flips the sign bit of every stored byte (offset-binary form of the cross cache)
template <bool INT8, uint32_t XORV = 0u>
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
            w[i] = v ^ XORV;
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

// 16 cache elements of one key held in registers (int8: one 128-bit load, fp16: two)
template <bool INT8>
struct KvChunk;

template <>
struct KvChunk<true>
{
    uint4 v;

    __device__ __forceinline__ void load(const void* base, size_t elem_off)
    {
        v = __ldg(reinterpret_cast<const uint4*>(static_cast<const int8_t*>(base) + elem_off));
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
struct KvChunk<false>
{
    uint4 v0, v1;

    __device__ __forceinline__ void load(const void* base, size_t elem_off)
    {
        const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __half*>(base) + elem_off);
        v0 = __ldg(p);
        v1 = __ldg(p + 1);
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

__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t h2u(__half2 h)
{
    return *reinterpret_cast<uint32_t*>(&h);
}

// D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col)
__device__ __forceinline__ void mma_m16n8k16(float& c0, float& c1, float& c2, float& c3, uint32_t a0, uint32_t a1,
    uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 16 cache bytes (or 16 fp16) of one key, read from shared memory -> 8 half2 in the pair order
// (d0,d2) (d1,d3) (d4,d6) (d5,d7) ...
template <bool INT8>
__device__ __forceinline__ void xa_load16(const uint8_t* p, __half2 (&w)[8])
{
    if constexpr (INT8)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        // the cross cache stores offset-binary bytes (q + 128), so the magic-number conversion needs no sign fix-up
        dequant_word(v.x, w[0], w[1]);
        dequant_word(v.y, w[2], w[3]);
        dequant_word(v.z, w[4], w[5]);
        dequant_word(v.w, w[6], w[7]);
    }
    else
    {
        const uint4 v0 = *reinterpret_cast<const uint4*>(p);
        const uint4 v1 = *reinterpret_cast<const uint4*>(p + 16);
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
}

// 16 offset-binary cache bytes -> 8 half2 holding 1024 + byte (PRMT only, no bias subtraction): for the score MMA the
// constant 1152 = 1024 + 128 per element is removed once per (row, head) as 1152 * sum(q) instead of once per element
__device__ __forceinline__ void xa_load16_biased(const uint8_t* p, __half2 (&w)[8])
{
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        const uint32_t l = __byte_perm(u[i], 0x64646464u, 0x5250);
        const uint32_t h = __byte_perm(u[i], 0x64646464u, 0x5351);
        w[2 * i] = *reinterpret_cast<const __half2*>(&l);
        w[2 * i + 1] = *reinterpret_cast<const __half2*>(&h);
    }
}

// one chunk of keys: scores on the tensor cores (mma.sync m16n8k16, the 16 keys of two warp iterations are the
// rows of A, q is column 0 of B), online softmax, then p.v on the fp16 pipe.  FULL: no key of the chunk is masked.
// KOFF (int8 cache only): the keys enter the MMA as 1024 + byte and `koff` = 1152 * sum_d q[d] is subtracted from the
// fp32 score (exact products, fp32 accumulation: the cancellation costs ~1e-6 relative) -- 8 fewer instructions per lane
// and 8 keys in a loop that is issue-bound.  The running state is only rescaled when the maximum moved.
template <bool INT8, int NIT, bool FULL, bool KOFF = false>
__device__ __forceinline__ void xa_chunk(const uint8_t* kst, const uint8_t* vst, int nk, int kl, int lane, float sscale,
    const uint32_t (&bq)[8], float& m_run, float& l_run, float (&o)[16], float koff = 0.f)
{
    constexpr int ESZ = INT8 ? 1 : 2;
    float sc[NIT];
    float m_new = m_run;
#pragma unroll
    for (int it = 0; it < NIT; it += 2)
    {
        __half2 w0[8], w1[8];
        if constexpr (KOFF && INT8)
        {
            xa_load16_biased(kst + (size_t) it * 8 * kDh, w0);
            xa_load16_biased(kst + (size_t) (it + 1) * 8 * kDh, w1);
        }
        else
        {
            xa_load16<INT8>(kst + (size_t) it * 8 * kDh * ESZ, w0);
            xa_load16<INT8>(kst + (size_t) (it + 1) * 8 * kDh * ESZ, w1);
        }
        // four INDEPENDENT accumulators (one per 16-dim slice) summed afterwards: a chain of four dependent mma.sync
        // costs four tensor-pipe latencies per 16 keys, and this loop has only 2-3 warps per scheduler to hide them
        float ca[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            ca[j][0] = ca[j][1] = ca[j][2] = ca[j][3] = 0.f;
            mma_m16n8k16(ca[j][0], ca[j][1], ca[j][2], ca[j][3], h2u(w0[2 * j]), h2u(w1[2 * j]), h2u(w0[2 * j + 1]),
                h2u(w1[2 * j + 1]), bq[2 * j], bq[2 * j + 1]);
        }
        const float c0 = (ca[0][0] + ca[1][0]) + (ca[2][0] + ca[3][0]);
        const float c2 = (ca[0][2] + ca[1][2]) + (ca[2][2] + ca[3][2]);
        // column 0 of D lives in the lanes with chunk == 0: c0 = key it*8+kl, c2 = key (it+1)*8+kl
        float s0 = __shfl_sync(0xffffffffu, c0, lane & ~3), s1 = __shfl_sync(0xffffffffu, c2, lane & ~3);
        if constexpr (KOFF && INT8)
        {
            s0 = (s0 - koff) * sscale;
            s1 = (s1 - koff) * sscale;
        }
        else
        {
            s0 *= sscale;
            s1 *= sscale;
        }
        if (!FULL)
        {
            s0 = (it * 8 + kl < nk) ? s0 : -FLT_MAX;
            s1 = ((it + 1) * 8 + kl < nk) ? s1 : -FLT_MAX;
        }
        sc[it] = s0;
        sc[it + 1] = s1;
        m_new = fmaxf(m_new, fmaxf(s0, s1));
    }
    m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, 4));
    m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, 8));
    m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, 16));
    // online softmax in the log2 domain: rescale the running state to the new maximum
    if (m_new != m_run) // warp-uniform (m_new was reduced over the warp; every lane carries the same m_run)
    {
        const float corr = fast_exp2(m_run - m_new); // 0 on the first chunk (m_run = -FLT_MAX)
        m_run = m_new;
        l_run *= corr;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            o[i] *= corr;
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it)
    {
        const float e = (FULL || it * 8 + kl < nk) ? fast_exp2(sc[it] - m_new) : 0.f;
        sc[it] = e;
        l_run += e;
    }
    // ---- p.v: up to 8 keys chained in fp16, then flushed to fp32 ----
    __half2 o2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        o2[i] = __float2half2_rn(0.f);
#pragma unroll
    for (int it = 0; it < NIT; ++it)
    {
        if (!INT8 && !FULL && it * 8 + kl >= nk)
            continue; // fp16 cache: stale shared-memory bits beyond the last key could decode to NaN
        const __half2 p2 = __float2half2_rn(sc[it]);
        __half2 w[8];
        xa_load16<INT8>(vst + (size_t) it * 8 * kDh * ESZ, w);
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

} // namespace b200
