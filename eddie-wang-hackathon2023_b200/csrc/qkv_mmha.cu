// qkv_mmha.cu -- ONE kernel for the first two operators of a decoder layer in the generation phase:
//     LayerNorm -> qkv projection (int8 weight-only, 1280 -> 3840, + bias) -> masked self-attention over the int8 KV cache
// (b200_qkv_mmha_decode, include/b200_whisper.h).  The reference runs them as the weight-only matmul plugin, a bias
// layer and the GPTAttention plugin (weightOnlyQuantMatmulPlugin.cpp:162-222 + gptAttentionCommon.cpp:649-780); round 1
// already folded the LayerNorm and the bias into the matmul.  What is left between them is one kernel boundary and a
// round trip of q / k / v through L2 -- about 3.5 us of the 39 us a layer takes at batch 16 -- and both operators are
// HEAD-LOCAL: the attention of head h needs exactly the 192 output columns (64 of q, 64 of k, 64 of v) of head h.
//
//   * grid = (S, H): a thread-block cluster of S CTAs per head, CTA r of it multiplies k-range r of the 192 columns.
//     Before the dependency on the previous kernel resolves (programmatic dependent launch) every CTA has its weight
//     slice in shared memory (96 row-pair segments of the reference's preprocessed layout, cp.async.bulk + mbarrier),
//     the per-column vectors in registers, and -- for the (batch row, head) pairs it will finish -- the first pass of
//     the K / V cache converted to fp16 registers.
//   * after the wait: the raw residual rows of its k-range straight from L2 into mma.sync A fragments (the reference
//     layout's k permutation is applied by the load addresses), B fragments by PRMT + HSUB2 dequant (x gamma: folded
//     LayerNorm) from shared memory, 3 column tiles per warp, fp32 accumulators; LayerNorm statistics of the k-range
//     from the same fragments (sums of differences from the row's first element).
//   * split-K reduction + "transpose" in one DSMEM exchange: every CTA pushes, for each batch row, its partial sums to
//     the CTA that OWNS the row (row mod S) with st.async + mbarrier complete_tx; the owner adds the S partials in rank
//     order (deterministic), applies scale / folded LayerNorm / bias with the rounding of the per-operator kernels and
//     has q, k, v of its (row, head) pairs in shared memory.
//   * the attention of those pairs, two warps per pair, is mmha_generation_kernel's arithmetic line by line (attention.cu):
//     cached keys as exact fp16 integers, q.k in HFMA2 chains, fp32 softmax with 1 / (sum + 1e-6), the current token's
//     k / v unquantized, the new K / V quantized with cvt.rni.sat and appended.
// int8 KV cache, linear buffer, no padding mask, batch <= 16, hidden size = H * 64 with K / 64 divisible by the cluster
// size: the decoder runtime's generation step.  Anything else takes the two-kernel route.
#include <float.h>
#include <stdlib.h>

#include "attn_device.cuh"
#include "common.cuh"
#include "tcgen05.cuh"

namespace b200
{
namespace
{
constexpr int kQmWarps = 8;
constexpr int kQmThreads = kQmWarps * 32;
constexpr int kQmCols = 3 * kDh;  // 192 output columns per head
constexpr int kQmTiles = kQmCols / 8;
constexpr int kQmRows = 16;       // batch rows (MMA M)
constexpr int kQmPart = kDh + 4;  // attention partial: m, l, 2 pad, o[64]
} // namespace

struct QmParams
{
    const __half* x;       // [B, K] raw residual rows
    const int8_t* W;       // preprocessed int8 [K, 3K]
    const __half* scales;  // [3K]
    const __half* bias;    // [3K] or null
    const __half* gamma;   // [K]
    const float* c1s;      // [3K]
    const float* c2;       // [3K]
    int8_t* cache;         // [B, 2, H, Smax, 64]
    const int* seq_len;    // [B]
    const float* s_oq;     // kv_orig_quant_scale
    const float* s_qo;     // kv_quant_orig_scale
    __half* out;           // [B, K]
    int B, H, K, Smax, S;  // S = cluster size (k-splits)
    float eps;
};

// D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col)
__device__ __forceinline__ void qm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NKB, int S> // k-blocks (of 64) per CTA, cluster size
__global__ void __launch_bounds__(kQmThreads, 2) qkv_mmha_decode_kernel(const QmParams p)
{
    constexpr int KU = NKB * 64;
    extern __shared__ __align__(128) uint8_t smem[];
    // [ weights: 96 row pairs x 2 KU bytes | inbox: S x (16 / S) rows x 192 fp32 | stats inbox: S x (16 / S) x 2 | qkv of the owned
    //   rows as fp16 | attention partials | barriers ]
    uint8_t* sW = smem;
    uint4* xs = reinterpret_cast<uint4*>(sW + (size_t) 96 * 2 * KU);    // activations of the k-range, A-fragment order: [kb][w][lane] x 16 B
    __half* sgamma = reinterpret_cast<__half*>(xs + NKB * 4 * 32);      // gamma of the k-range
    float* inbox = reinterpret_cast<float*>(sgamma + KU);
    float* sinbox = inbox + kQmRows * kQmCols;            // [rank][local row][2]
    __half* sqkv = reinterpret_cast<__half*>(sinbox + kQmRows * 2); // [local row][192]
    float* parts = reinterpret_cast<float*>(sqkv + kQmRows * kQmCols); // [warp pair][kQmPart] (over-allocated for 16 rows: fine)
    float* srow_shift = parts + (kQmWarps / 2) * kQmPart + 4; // [local row]: x[row][0]
    uint64_t* w_full = reinterpret_cast<uint64_t*>(srow_shift + kQmRows);
    uint64_t* in_full = w_full + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r = (int) cluster_ctarank(); // == blockIdx.x
    const int h = blockIdx.y;
    const int K = p.K, N = 3 * K;
    constexpr int rows_per = kQmRows / S; // rows finished by a CTA: row = lr * S + rank

    if (threadIdx.x == 0)
    {
        mbar_init(w_full, 1);
        mbar_init(in_full, 1);
        fence_mbar_init();
        fence_proxy_async_smem();
        // every rank (this one included) pushes rows_per rows of 192 partial sums and 2 statistics each
        mbar_arrive_expect_tx(in_full, (uint32_t) (S * rows_per * (kQmCols + 2) * sizeof(float)));
        // weights never depend on the previous kernel: 96 row-pair segments of this CTA's k-range (issued below, one per thread)
        mbar_arrive_expect_tx(w_full, (uint32_t) (96 * 2 * KU));
    }
    __syncthreads();
    if ((threadIdx.x & 31) < 12)
    {
        // 12 bulk copies per warp (UBLKCP is issued lane by lane: a single lane issuing all 96 serialises ~100 cycles each)
        const int seg = (threadIdx.x >> 5) * 12 + (threadIdx.x & 31);
        const int part = seg >> 5, j = seg & 31;
        const size_t rp = (size_t) (part * K + h * kDh) / 2 + j;
        bulk_g2s_hint(sW + (size_t) seg * 2 * KU, reinterpret_cast<const uint8_t*>(p.W) + rp * 2 * K + (size_t) r * 2 * KU,
            (uint32_t) 2 * KU, w_full, policy_evict_first());
    }
    __syncthreads();
    cluster_sync_all(); // every rank's inbox barrier is armed before anybody can push into it
    grid_dep_launch_dependents();

    // ---- static operands, before the dependency wait ----
    // this warp's tiles: warp, warp + 8, warp + 16 -> tile T covers columns part * 64 + j * 8 .. + 7 of the head (part = q / k / v)
    for (int i = threadIdx.x; i < KU / 8; i += kQmThreads)
        reinterpret_cast<uint4*>(sgamma)[i] = __ldg(reinterpret_cast<const uint4*>(p.gamma + (size_t) r * KU) + i);
    // attention: warp pair wp finishes the local rows wp, wp + 4, ...; fetch the first pass of the first pair's cache now
    constexpr int NIT = 4;
    const int chunk = lane & 3, kl = lane >> 2;
    const int half = warp & 1, wp = warp >> 1;
    const float s_qo = __ldg(p.s_qo), s_oq = __ldg(p.s_oq);
    KvChunk<true> kreg[NIT], vreg[NIT]; // raw cache bytes of one pass; converted to fp16 at the point of use
    auto cache_k = [&](int b) { return reinterpret_cast<char*>(p.cache) + ((size_t) (b * 2 + 0) * p.H + h) * p.Smax * kDh; };
    auto cache_v = [&](int b) { return reinterpret_cast<char*>(p.cache) + ((size_t) (b * 2 + 1) * p.H + h) * p.Smax * kDh; };
    auto fetch = [&](int b, int k0)
    {
        const char* kc = cache_k(b);
        const char* vc = cache_v(b);
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            const int key = min(k0 + (2 * it + half) * 8 + kl, p.Smax - 1);
            kreg[it].load(kc, (size_t) key * kDh + chunk * 16);
            vreg[it].load(vc, (size_t) key * kDh + chunk * 16);
        }
    };
    // the finalising items of this thread: (local row, column) = idx / 192, idx % 192 for idx = tid, tid + 256, ...: their
    // per-column vectors are static
    constexpr int kMaxItems = (rows_per * kQmCols + kQmThreads - 1) / kQmThreads; // 3 at S = 4
    float f_sc[kMaxItems], f_c1[kMaxItems], f_c2[kMaxItems], f_bias[kMaxItems];
#pragma unroll
    for (int i = 0; i < kMaxItems; ++i)
    {
        const int idx = threadIdx.x + i * kQmThreads;
        f_sc[i] = f_c1[i] = f_c2[i] = f_bias[i] = 0.f;
        if (idx < rows_per * kQmCols)
        {
            const int col = idx % kQmCols;
            const int n = (col >> 6) * K + h * kDh + (col & 63); // global output column
            f_sc[i] = __half2float(__ldg(p.scales + n));
            f_c1[i] = __ldg(p.c1s + n);
            f_c2[i] = __ldg(p.c2 + n);
            f_bias[i] = p.bias != nullptr ? __half2float(__ldg(p.bias + n)) : 0.f;
        }
    }
    const int b_first = wp * S + r; // local row wp of this rank
    int tlen_first = 0;
    if (wp < rows_per && b_first < p.B)
    {
        tlen_first = min(__ldg(p.seq_len + b_first), p.Smax - 1); // written by the previous step: static here
        fetch(b_first, 0);
    }

    grid_dep_wait(); // x comes from the previous kernel

    // ---- the activations of this CTA's k-range: coalesced 16-byte loads of the row-major residual rows, scattered into shared
    // memory in A-fragment order (word ((kb * 4 + w) * 32 + 4 g + T) * 4 + 2 hi + up holds the k pair 64 kb + 16 T + 8 hi + 2 w of
    // row g + 8 up: the order the dequantized weight words come in, common.cuh dequant_word), so every MMA's A operand is one
    // conflict-free 128-bit read per lane.  (Loading the fragments straight from global memory -- 80 four-byte loads per
    // thread -- ran into the load / store unit's queue: ncu `lg throttle`, 14.9 us per layer.)
    const bool v0 = g < p.B, v1 = g + 8 < p.B;
    for (int c = threadIdx.x; c < kQmRows * (KU / 8); c += kQmThreads)
    {
        const int row = c / (KU / 8), ch = c - row * (KU / 8);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row < p.B)
            v = __ldcg(reinterpret_cast<const uint4*>(p.x + (size_t) row * K + (size_t) r * KU) + ch);
        const int kb = ch >> 3, T = (ch & 7) >> 1, hi = ch & 1;
        uint32_t* dst = reinterpret_cast<uint32_t*>(xs) + ((size_t) (kb * 4) * 32 + 4 * (row & 7) + T) * 4 + 2 * hi + (row >> 3);
        dst[0 * 128] = v.x; // w = 0 .. 3: 128 words apart
        dst[1 * 128] = v.y;
        dst[2 * 128] = v.z;
        dst[3 * 128] = v.w;
    }
    // the rows' first elements (global k = 0, 1): the common shift of the LayerNorm sums of every rank
    unsigned shw0 = 0u, shw1 = 0u;
    if (warp == 0)
    {
        shw0 = v0 ? __ldcg(reinterpret_cast<const unsigned*>(p.x + (size_t) g * K)) : 0u;
        shw1 = v1 ? __ldcg(reinterpret_cast<const unsigned*>(p.x + (size_t) (g + 8) * K)) : 0u;
    }
    __syncthreads();
    // the shifts of the rows this rank finishes, for the epilogue (row lr * S + r): held by the lanes with t == 0 of any warp
    if (warp == 0 && t == 0)
    {
        if (g % S == r)
            srow_shift[g / S] = __low2float(*reinterpret_cast<const __half2*>(&shw0));
        if ((g + 8) % S == r)
            srow_shift[(g + 8) / S] = __low2float(*reinterpret_cast<const __half2*>(&shw1));
    }
    mbar_wait(w_full, 0);

    // ---- 3 tiles x NKB k-blocks x 4 MMAs per warp ----
    float sd0 = 0.f, sq0 = 0.f, sd1 = 0.f, sq1 = 0.f; // warp 0: LayerNorm sums of rows g, g + 8 (reduced over the lane quad below)
    const float sh0 = __low2float(*reinterpret_cast<const __half2*>(&shw0));
    const float sh1 = __low2float(*reinterpret_cast<const __half2*>(&shw1));
    float acc[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    const uint32_t lane_off = (uint32_t) ((g >> 1) * 2 * KU + (g & 1) * 64 + t * 16);
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb)
    {
        uint4 wv[3], av[4];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            wv[i] = *reinterpret_cast<const uint4*>(sW + (size_t) (warp + 8 * i) * 4 * 2 * KU + lane_off + kb * 128);
#pragma unroll
        for (int w = 0; w < 4; ++w)
            av[w] = xs[(kb * 4 + w) * 32 + lane];
        const uint4 glo = *reinterpret_cast<const uint4*>(sgamma + kb * 64 + 16 * t);
        const uint4 ghi = *reinterpret_cast<const uint4*>(sgamma + kb * 64 + 16 * t + 8);
#pragma unroll
        for (int w = 0; w < 4; ++w)
        {
            const uint32_t gl = w == 0 ? glo.x : w == 1 ? glo.y : w == 2 ? glo.z : glo.w;
            const uint32_t gh = w == 0 ? ghi.x : w == 1 ? ghi.y : w == 2 ? ghi.z : ghi.w;
            const uint32_t a4[4] = {av[w].x, av[w].y, av[w].z, av[w].w};
#pragma unroll
            for (int i = 0; i < 3; ++i)
            {
                const uint32_t word = w == 0 ? wv[i].x : w == 1 ? wv[i].y : w == 2 ? wv[i].z : wv[i].w;
                __half2 lo, hi;
                dequant_word(word, lo, hi);
                lo = __hmul2(lo, *reinterpret_cast<const __half2*>(&gl));
                hi = __hmul2(hi, *reinterpret_cast<const __half2*>(&gh));
                qm_mma(acc[i], a4, h2u(lo), h2u(hi));
            }
        }
        if (warp == 0)
        {
            // LayerNorm statistics of rows g and g + 8 over this k-block: sums of (x - x[row][0]) and of its square
#pragma unroll
            for (int w = 0; w < 4; ++w)
            {
                const uint32_t r0[2] = {av[w].x, av[w].z};
                const uint32_t r1[2] = {av[w].y, av[w].w};
#pragma unroll
                for (int i = 0; i < 2; ++i)
                {
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&r0[i]));
                    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&r1[i]));
                    const float d0 = f0.x - sh0, d1 = f0.y - sh0, d2 = f1.x - sh1, d3 = f1.y - sh1;
                    sd0 += d0 + d1;
                    sq0 = fmaf(d0, d0, fmaf(d1, d1, sq0));
                    sd1 += d2 + d3;
                    sq1 = fmaf(d2, d2, fmaf(d3, d3, sq1));
                }
            }
        }
    }

    // ---- push the partial sums to the rows' owners: inbox[sender rank][local row][column] ----
    const uint32_t in_bar = smem_u32(in_full);
    {
        const int own0 = g % S, own1 = (g + 8) % S;   // owner ranks of rows g and g + 8
        const int lr0 = g / S, lr1 = (g + 8) / S;
        const uint32_t bar0 = mapa_u32(in_bar, (uint32_t) own0), bar1 = mapa_u32(in_bar, (uint32_t) own1);
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
            const int col = (warp + 8 * i) * 8 + 2 * t;
            const uint32_t a0 = mapa_u32(smem_u32(inbox + ((size_t) r * rows_per + lr0) * kQmCols + col), (uint32_t) own0);
            const uint32_t a1 = mapa_u32(smem_u32(inbox + ((size_t) r * rows_per + lr1) * kQmCols + col), (uint32_t) own1);
            st_async_f32(a0, acc[i][0], bar0);
            st_async_f32(a0 + 4, acc[i][1], bar0);
            st_async_f32(a1, acc[i][2], bar1);
            st_async_f32(a1 + 4, acc[i][3], bar1);
        }
        if (warp == 0)
        {
            // the statistics of this rank's k-range, reduced over the four lanes that share a row; plain sums add up across
            // ranks (common shift)
#pragma unroll
            for (int o = 1; o < 4; o <<= 1)
            {
                sd0 += __shfl_xor_sync(0xffffffffu, sd0, o);
                sq0 += __shfl_xor_sync(0xffffffffu, sq0, o);
                sd1 += __shfl_xor_sync(0xffffffffu, sd1, o);
                sq1 += __shfl_xor_sync(0xffffffffu, sq1, o);
            }
            if (t == 0)
            {
                const uint32_t s0 = mapa_u32(smem_u32(sinbox + ((size_t) r * rows_per + lr0) * 2), (uint32_t) own0);
                const uint32_t s1 = mapa_u32(smem_u32(sinbox + ((size_t) r * rows_per + lr1) * 2), (uint32_t) own1);
                st_async_f32(s0, sd0, bar0);
                st_async_f32(s0 + 4, sq0, bar0);
                st_async_f32(s1, sd1, bar1);
                st_async_f32(s1 + 4, sq1, bar1);
            }
        }
    }

    // ---- owner: add the S partials in rank order, folded LayerNorm, bias -> q | k | v of its rows (fp16, shared memory) ----
    mbar_wait(in_full, 0);
#pragma unroll
    for (int i = 0; i < kMaxItems; ++i)
    {
        const int idx = threadIdx.x + i * kQmThreads;
        if (idx >= rows_per * kQmCols)
            break;
        const int lr = idx / kQmCols, col = idx - lr * kQmCols;
        float s = 0.f, sd = 0.f, sq = 0.f;
        for (int q = 0; q < S; ++q)
        {
            s += inbox[((size_t) q * rows_per + lr) * kQmCols + col];
            sd += sinbox[((size_t) q * rows_per + lr) * 2];
            sq += sinbox[((size_t) q * rows_per + lr) * 2 + 1];
        }
        const float rk = 1.f / (float) K;
        const float md = sd * rk;
        const float mean = srow_shift[lr] + md;
        const float rstd = rsqrtf(fmaxf(sq * rk - md * md, 0.f) + p.eps);
        float v = s * f_sc[i];
        v = rstd * (v - mean * f_c1[i]) + f_c2[i];
        __half o = __float2half_rn(v);
        if (p.bias != nullptr)
            o = __float2half_rn(__half2float(o) + f_bias[i]);
        sqkv[lr * kQmCols + col] = o;
    }
    __syncthreads();

    // ---- masked self-attention of the owned (row, head) pairs: mmha_generation_kernel's arithmetic (attention.cu) ----
    const float inv_sqrt_dh = 0.125f; // 1 / sqrt(64), q_scaling = 1 (gptAttentionCommon.cpp:163)
    const float sscale = s_qo * inv_sqrt_dh;
    const int n_iter = (rows_per + kQmWarps / 2 - 1) / (kQmWarps / 2);
    for (int itp = 0; itp < n_iter; ++itp)
    {
        const int lr = wp + itp * (kQmWarps / 2);
        const int b = lr * S + r;
        const bool active = lr < rows_per && b < p.B;
        int tlen = 0;
        if (active)
        {
            if (itp == 0)
                tlen = tlen_first;
            else
            {
                tlen = min(__ldg(p.seq_len + b), p.Smax - 1);
                fetch(b, 0);
            }
        }
        char* kc = cache_k(active ? b : 0);
        char* vc = cache_v(active ? b : 0);
        __half qh[16], kh[16], vh[16];
        float s_cur = -FLT_MAX;
        float m_run = -FLT_MAX, l_run = 0.f;
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
            o[i] = 0.f;
        if (active)
        {
            const __half* qs = sqkv + lr * kQmCols + chunk * 16;
            load16_half(qs, nullptr, qh);
            if (half == 0)
            {
                load16_half(qs + kDh, nullptr, kh);
                load16_half(qs + 2 * kDh, nullptr, vh);
                // append this step's K and V (lane group 0 writes K, group 1 writes V; 16 dims per lane)
                if (kl == 0)
                    store16<true>(kc, (size_t) tlen * kDh + chunk * 16, s_oq, kh);
                else if (kl == 1)
                    store16<true>(vc, (size_t) tlen * kDh + chunk * 16, s_oq, vh);
                float sc0 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    sc0 = fmaf(__half2float(qh[i]), __half2float(kh[i]), sc0);
                sc0 += __shfl_xor_sync(0xffffffffu, sc0, 1);
                sc0 += __shfl_xor_sync(0xffffffffu, sc0, 2);
                s_cur = sc0 * inv_sqrt_dh;
            }
            __half2 q2[8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                q2[2 * i] = __halves2half2(qh[4 * i], qh[4 * i + 2]);
                q2[2 * i + 1] = __halves2half2(qh[4 * i + 1], qh[4 * i + 3]);
            }
            m_run = s_cur;
            for (int k0 = 0; k0 < tlen; k0 += 16 * NIT)
            {
                if (k0 > 0)
                    fetch(b, k0);
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
                        __half2 kw[8];
                        kreg[it].unpack(kw);
                        __half2 h0 = __hmul2(q2[0], kw[0]);
                        __half2 h1 = __hmul2(q2[4], kw[4]);
                        h0 = __hfma2(q2[1], kw[1], h0);
                        h1 = __hfma2(q2[5], kw[5], h1);
                        h0 = __hfma2(q2[2], kw[2], h0);
                        h1 = __hfma2(q2[6], kw[6], h1);
                        h0 = __hfma2(q2[3], kw[3], h0);
                        h1 = __hfma2(q2[7], kw[7], h1);
                        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                        float sv = (f0.x + f0.y) + (f1.x + f1.y);
                        sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                        sv += __shfl_xor_sync(0xffffffffu, sv, 2);
                        sv = key < tlen ? sv * sscale : -FLT_MAX;
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
                        continue;
                    const float e = __expf(sc[it] - m_new);
                    l_run += e;
                    const __half2 p2 = __float2half2_rn(e);
                    __half2 vw[8];
                    vreg[it].unpack(vw);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        o2[i] = __hfma2(p2, vw[i], o2[i]);
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
        }
        // warp 1 of the pair hands its state to warp 0 through shared memory
        float* pr = parts + (size_t) wp * kQmPart;
        if (active && half == 1 && kl == 0)
        {
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
                __half* dst = p.out + (size_t) b * K + h * kDh + chunk * 16;
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
        __syncthreads(); // the partial area is reused by the next pair of this warp pair
    }
    // a CTA must not exit while its peers may still push into its shared memory: all pushes have landed once every rank has
    // passed its own inbox wait, which this barrier implies
    cluster_sync_all();
}

template <int NKB>
static size_t qm_smem_bytes()
{
    return (size_t) 96 * 2 * NKB * 64 + (size_t) NKB * 4 * 32 * 16 + (size_t) NKB * 64 * 2
        + sizeof(float) * (kQmRows * kQmCols + kQmRows * 2) + sizeof(__half) * kQmRows * kQmCols
        + sizeof(float) * ((kQmWarps / 2) * kQmPart + 4 + kQmRows) + 2 * sizeof(uint64_t) + 64;
}

template <int NKB, int S>
static int qm_launch(const QmParams& p, cudaStream_t stream)
{
    auto kern = qkv_mmha_decode_kernel<NKB, S>;
    const size_t smem = qm_smem_bytes<NKB>();
    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned) S, (unsigned) p.H);
    cfg.blockDim = dim3(kQmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl_enabled())
    {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned) S;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    count_launch();
    B200_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return B200_OK;
}

} // namespace b200

using namespace b200;

/* 1 if b200_qkv_mmha_decode handles this shape (hidden size = num_heads * 64, batch <= 16, int8 linear KV cache) */
extern "C" int b200_qkv_mmha_decode_supported(int batch_size, int num_heads, int head_size)
{
    if (head_size != kDh || batch_size < 1 || batch_size > kQmRows || num_heads < 1)
        return 0;
    // cluster size 4 when the k-blocks (= heads) divide by it, else 2; per-CTA slices of 1 .. 5 k-blocks are instantiated
    const int nkb = num_heads; // K / 64
    if (nkb % 4 == 0)
        return nkb / 4 <= 5 ? 1 : 0;
    if (nkb % 2 == 0)
        return (nkb / 2 == 1 || nkb / 2 == 3 || nkb / 2 == 5) ? 1 : 0;
    return 0;
}

extern "C" int b200_qkv_mmha_decode(const void* x, const void* ln_gamma, const float* c1s, const float* c2, float ln_eps,
    const int8_t* Wproc, const void* scales, const void* bias, void* kv_cache, const int32_t* sequence_lengths,
    const float* kv_scale_orig_quant, const float* kv_scale_quant_orig, void* out, int batch_size, int num_heads,
    int head_size, int max_seq_len, b200_stream_t stream)
{
    B200_REQUIRE(x && ln_gamma && c1s && c2 && Wproc && scales && kv_cache && sequence_lengths && kv_scale_orig_quant
            && kv_scale_quant_orig && out,
        B200_ERR_INVALID_ARG, "null pointer");
    B200_REQUIRE(b200_qkv_mmha_decode_supported(batch_size, num_heads, head_size), B200_ERR_UNSUPPORTED,
        "qkv + attention fusion: batch %d, %d heads of %d", batch_size, num_heads, head_size);
    B200_REQUIRE(max_seq_len > 0, B200_ERR_INVALID_ARG, "bad sizes");
    B200_REQUIRE_DEVICE();
    QmParams p{};
    p.x = static_cast<const __half*>(x);
    p.W = Wproc;
    p.scales = static_cast<const __half*>(scales);
    p.bias = static_cast<const __half*>(bias);
    p.gamma = static_cast<const __half*>(ln_gamma);
    p.c1s = c1s, p.c2 = c2;
    p.cache = static_cast<int8_t*>(kv_cache);
    p.seq_len = sequence_lengths;
    p.s_oq = kv_scale_orig_quant, p.s_qo = kv_scale_quant_orig;
    p.out = static_cast<__half*>(out);
    p.B = batch_size, p.H = num_heads, p.K = num_heads * kDh, p.Smax = max_seq_len;
    p.eps = ln_eps;
    const int nkb = num_heads;
    const cudaStream_t st = as_stream(stream);
    if (nkb % 4 == 0)
    {
        p.S = 4;
        switch (nkb / 4)
        {
        case 1: return qm_launch<1, 4>(p, st);
        case 2: return qm_launch<2, 4>(p, st);
        case 3: return qm_launch<3, 4>(p, st);
        case 4: return qm_launch<4, 4>(p, st);
        default: return qm_launch<5, 4>(p, st);
        }
    }
    p.S = 2;
    switch (nkb / 2)
    {
    case 1: return qm_launch<1, 2>(p, st);
    case 3: return qm_launch<3, 2>(p, st);
    default: return qm_launch<5, 2>(p, st);
    }
}
