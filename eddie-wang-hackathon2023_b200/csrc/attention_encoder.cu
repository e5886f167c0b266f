// attention_encoder.cu -- bidirectional (encoder) multi-head attention over the 1500 audio frames, head size 64.
//
// Replaces the unfused attention of the reference's encoder blocks (T/tensorrt_llm/layers/attention.py:283-406 with
// no mask / no KV cache, used by T/tensorrt_llm/models/whisper/model.py:124-172; oracle W/torch_model.py:88-103):
//     softmax(q k^T / sqrt(64)) v        q, k, v [B, S, H, 64] taken from one fused projection [B, S, 3*H*64].
// Flash-attention style: one CTA per (64 query rows, head, batch), 4 warps x 16 rows; K and V tiles of 64 keys are
// double-buffered in shared memory with cp.async (16-byte chunks XOR-swizzled by the row so every ldmatrix is
// conflict-free); scores and P.V run on the tensor cores (mma.sync.m16n8k16, fp32 accumulate), the online softmax
// lives in registers (ex2 with the scale folded into one FFMA, lazily rescaled accumulators), the score accumulators are re-used in place as the fp16 A fragments of P.
// That warp-level mma.sync kernel (round 1) is kept for comparison (B200_ENC_ATTN=mma: 588 us per launch at batch 16).
// The default since round 2 is attention_bidir_tc_kernel below (277 us), the same flash-attention recurrence on the
// 5th-generation tensor cores:
//   * CTA = 128 query rows of one (batch, head); Q and the K / V tiles (64 rows x 64 dims, 128B-swizzled) arrive by 3-D
//     TMA boxes straight out of the fused [B, S, 3*H*64] projection (rows beyond S are zero-filled by the TMA unit);
//   * S_j = Q.K_j^T is one 128 x 64 x 64 UMMA chain (SS form, four tcgen05.mma of K = 16) into 64 TMEM columns; the
//     score buffer is double-buffered and the MMA warp runs two tiles ahead, so the softmax warps never wait for it;
//   * four softmax warps own one query row per thread = one TMEM lane: tcgen05.ld brings the row's 64 scores into
//     registers, row maximum / exp2 / row sum need no shuffles at all, P goes back to TMEM as packed fp16 (tcgen05.st,
//     double-buffered) and never touches shared memory;
//   * O += P_j.V_j is a 128 x 64 x 64 UMMA chain with A = P FROM TMEM and B = the V tile as it lies in shared memory
//     (keys x dims = MN-major B operand: the same TMA image as a K-major tile, only the instruction descriptor's
//     b_major bit differs);
//   * the accumulator O (64 TMEM columns) is rescaled lazily (only when a row maximum grew by more than 2^8 since the
//     maximum the accumulators are expressed in -- the rule of the mma.sync kernel) by the softmax warps between two
//     P.V chains; warp 4 is the TMA producer (four K / V stages), warp 5 issues the MMAs from one elected lane;
//   * 256 TMEM columns and 80 KB of shared memory per CTA: two CTAs share an SM.
// Measured on the way (profiles/r02_encoder_attention.txt): 128-key tiles with S and P sharing columns 426 us; Q.K^T of
// the next tile under the exponentials 374 us; scores read from TMEM in 32-column pieces instead of 128 live registers
// (no spills) 326 us; 64-key tiles with both buffers doubled 306 us; key masking hoisted out of the per-element path (ncu:
// ISETP + FSEL were 22 % of all instructions, on every tile) 284 us; one redundant barrier wait per tile removed 277 us.
// Moving part of the exponentials to an FMA-pipe polynomial (the FlashAttention-4 trick, ex2_fma / B200_ATTN_POLY_EVERY)
// is neutral to slower at every stage (348 / 373 / 407 us at a quarter / third / half on the 326 us version; 293 / 286 us
// at 1/4 and 1/8 on the 284 us version): the kernel is not MUFU-bound, its softmax warps (two per scheduler) are
// latency-bound on the barrier / TMEM round trips of each tile.
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace b200
{
namespace
{
constexpr int kD = 64;       // head size
constexpr int kBM = 64;      // query rows per CTA
constexpr int kBN = 64;      // keys per tile
constexpr int kTile = kBN * kD * 2; // bytes of one K or V tile (8 KB)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid)
{
    const uint32_t s = smem_u32(smem);
    const int sz = valid ? 16 : 0; // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA pipe (x <= ~8): round-to-nearest split x = n + f, |f| <= 0.5, degree-4 polynomial of 2^f (relative
// error 4e-5, a tenth of the fp16 rounding P gets anyway), n added to the exponent field (the FlashAttention-4 trick for
// kernels whose exponentials saturate the MUFU unit: 16 per clock per SM)
__device__ __forceinline__ float ex2_fma(float x)
{
    x = fmaxf(x, -126.f);
    const float t = x + 12582912.f; // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const float f = x - (t - 12582912.f);
    float p = fmaf(f, 0.0096181291f, 0.0555041087f);
    p = fmaf(p, f, 0.2402265070f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// byte offset of (row, 16-byte chunk c) inside a swizzled [rows][64 halfs] tile
__device__ __forceinline__ uint32_t swz(int row, int c)
{
    return (uint32_t) (row * 128 + ((c ^ (row & 7)) << 4));
}
} // namespace

__global__ void __launch_bounds__(128) attention_bidir_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int S, int H)
{
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t* sQ = sm;                 // [64][64] halfs
    uint8_t* sK = sm + kTile;         // 2 stages
    uint8_t* sV = sm + 3 * kTile;     // 2 stages
    const int q0 = blockIdx.x * kBM, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const size_t row_stride = (size_t) 3 * H * kD;
    const __half* base = qkv + (size_t) b * S * row_stride + (size_t) h * kD;
    const __half* qg = base;
    const __half* kg = base + (size_t) H * kD;
    const __half* vg = base + (size_t) 2 * H * kD;

    grid_dep_wait();
    grid_dep_launch_dependents();

    auto load_tile = [&](uint8_t* dst, const __half* src, int r0)
    {
        // 64 rows x 8 chunks of 16 bytes; 128 threads x 4
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const int idx = tid + 128 * i;
            const int r = idx >> 3, c = idx & 7;
            const bool ok = r0 + r < S;
            cp_async16(dst + swz(r, c), src + (size_t) (ok ? r0 + r : 0) * row_stride + c * 8, ok);
        }
    };
    load_tile(sQ, qg, q0);
    load_tile(sK, kg, 0);
    load_tile(sV, vg, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    const int n_tiles = (S + kBN - 1) / kBN;
    uint32_t qa[4][4]; // A fragments of this warp's 16 query rows, 4 k-steps of 16 dims
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            o[j][i] = 0.f;
    float m0 = -FLT_MAX, m1 = -FLT_MAX, l0 = 0.f, l1 = 0.f; // rows g and g + 8 of the warp's 16; m in raw score units
    const float sl2 = 0.125f * 1.4426950408889634f;       // 1/sqrt(64) * log2(e)

    for (int it = 0; it < n_tiles; ++it)
    {
        const int st = it & 1;
        if (it + 1 < n_tiles)
        {
            load_tile(sK + (st ^ 1) * kTile, kg, (it + 1) * kBN);
            load_tile(sV + (st ^ 1) * kTile, vg, (it + 1) * kBN);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        else
        {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (it == 0)
        {
            // matrices of one ldmatrix.x4: (rows 0-7 | 8-15) x (cols 0-7 | 8-15) of the 16 x 16 A tile
            const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                ldsm_x4(qa[kk], smem_u32(sQ) + swz(r, kk * 2 + (lane >> 4)));
        }
        const uint32_t kb = smem_u32(sK + st * kTile), vb = smem_u32(sV + st * kTile);

        // ---- scores: 16 rows x 64 keys ----
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                s[j][i] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
        {
#pragma unroll
            for (int np = 0; np < 4; ++np) // two key groups of 8 per ldmatrix.x4
            {
                // matrices: (keys np*16 + 0-7, dims kk*16 + 0-7), (same keys, dims +8), (keys +8, dims 0-7), (keys +8, dims +8)
                const int kr = np * 16 + (lane & 7) + (lane >> 4) * 8;
                uint32_t kf[4];
                ldsm_x4(kf, kb + swz(kr, kk * 2 + ((lane >> 3) & 1)));
                mma16816(s[2 * np], qa[kk], kf[0], kf[1]);
                mma16816(s[2 * np + 1], qa[kk], kf[2], kf[3]);
            }
        }
        // ---- online softmax; columns of s[j]: keys it*64 + j*8 + 2t, +1.  The running maxima m0 / m1 are kept in RAW
        // score units; the 1/sqrt(64) * log2(e) factor rides on the FFMA that feeds ex2.  ncu had this kernel at 53 % tensor
        // pipe with 6.6 other instructions per HMMA, so the loop is trimmed to what every tile needs:
        //   * keys beyond S exist only in the last tile: no per-key predicate elsewhere;
        //   * the accumulators are rescaled only when a row maximum grew by more than kLazy (in log2 units) since the
        //     maximum the accumulators are expressed in; until then P is formed against that older maximum (P <= 2^kLazy:
        //     no overflow in fp16 or fp32, and the final o / l cancels the common factor exactly).
        if (it == n_tiles - 1 && (S % kBN) != 0)
        {
            const int key0 = it * kBN + 2 * t;
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const int k = key0 + j * 8;
                s[j][0] = k < S ? s[j][0] : -FLT_MAX;
                s[j][1] = k + 1 < S ? s[j][1] : -FLT_MAX;
                s[j][2] = k < S ? s[j][2] : -FLT_MAX;
                s[j][3] = k + 1 < S ? s[j][3] : -FLT_MAX;
            }
        }
        float mx0 = s[0][0], mx1 = s[0][2];
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        constexpr float kLazy = 8.0f;
        const bool grow = (mx0 - m0) * sl2 > kLazy || (mx1 - m1) * sl2 > kLazy; // always true on the first tile
        if (__any_sync(0xffffffffu, grow))
        {
            const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
            const float c0 = ex2((m0 - n0) * sl2), c1 = ex2((m1 - n1) * sl2);
            m0 = n0;
            m1 = n1;
            l0 *= c0;
            l1 *= c1;
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                o[j][0] *= c0;
                o[j][1] *= c0;
                o[j][2] *= c1;
                o[j][3] *= c1;
            }
        }
        const float b0 = -m0 * sl2, b1 = -m1 * sl2;
        uint32_t pa[4][4]; // P as A fragments: k-step kk covers keys kk*16 .. +15 = score tiles 2kk, 2kk+1
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const float p0 = ex2(fmaf(s[j][0], sl2, b0)), p1 = ex2(fmaf(s[j][1], sl2, b0));
            const float p2 = ex2(fmaf(s[j][2], sl2, b1)), p3 = ex2(fmaf(s[j][3], sl2, b1));
            l0 += p0 + p1;
            l1 += p2 + p3;
            pa[j >> 1][(j & 1) * 2 + 0] = pack_h2(p0, p1);
            pa[j >> 1][(j & 1) * 2 + 1] = pack_h2(p2, p3);
        }
        // ---- o += P (16 x 64 keys) . V (64 keys x 64 dims) ----
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
        {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) // two dim groups of 8 per ldmatrix.x4.trans
            {
                // matrices: (keys kk*16 + 0-7, dims dp*16 + 0-7), (keys +8, same dims), (keys 0-7, dims +8), (keys +8, dims +8)
                const int vr = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                uint32_t vf[4];
                ldsm_x4_trans(vf, vb + swz(vr, dp * 2 + (lane >> 4)));
                mma16816(o[2 * dp], pa[kk], vf[0], vf[1]);
                mma16816(o[2 * dp + 1], pa[kk], vf[2], vf[3]);
            }
        }
        __syncthreads(); // everyone is done with stage st before the next iteration's loads overwrite it
    }
    // ---- finish: row sums over the quad, normalise, store ----
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __half* ob = out + (size_t) b * S * H * kD + (size_t) h * kD;
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
        const int d = j * 8 + 2 * t;
        if (r0 < S)
            *reinterpret_cast<uint32_t*>(ob + (size_t) r0 * H * kD + d) = pack_h2(o[j][0] * i0, o[j][1] * i0);
        if (r1 < S)
            *reinterpret_cast<uint32_t*>(ob + (size_t) r1 * H * kD + d) = pack_h2(o[j][2] * i1, o[j][3] * i1);
    }
}

// =====================================================================================================================
// tcgen05 / TMEM kernel
// =====================================================================================================================
namespace
{
constexpr int kTM = 128;                 // query rows per CTA (UMMA M)
constexpr int kTN = 64;                  // keys per tile
constexpr int kTcBox = 64 * kD * 2;      // bytes of one TMA box (64 rows x 64 dims): 8 KB; Q = two boxes
constexpr int kTcStages = 4;             // K / V tiles in flight
// TMEM columns: S double-buffered, P double-buffered, O -- 256 in all, so two CTAs share an SM's 512
constexpr uint32_t kTcTmemCols = 256;
constexpr uint32_t kTcColS = 0;          // [0, 64), [64, 128): fp32 scores of tile j in buffer j & 1
constexpr uint32_t kTcColP = 128;        // [128, 160), [160, 192): fp16 probabilities, two keys per column
constexpr uint32_t kTcColO = 192;        // [192, 256): fp32 output accumulator
// kind::f16, fp32 accumulate, A and B fp16; M = 128, N = 64; Q.K^T: both operands K-major; P.V: B (= V) MN-major
constexpr uint32_t kIdescQK = (1u << 4) | ((uint32_t) (kTN >> 3) << 17) | ((uint32_t) (kTM >> 4) << 24);
constexpr uint32_t kIdescPV = (1u << 4) | (1u << 16) | ((uint32_t) (kD >> 3) << 17) | ((uint32_t) (kTM >> 4) << 24);
#ifndef B200_ATTN_POLY_EVERY
#define B200_ATTN_POLY_EVERY 0
#endif
constexpr int kPolyEvery = B200_ATTN_POLY_EVERY; // every n-th pair of probabilities takes the FMA-pipe exponential (0: none)
constexpr size_t kTcSmem = (size_t) (2 + 2 * kTcStages) * kTcBox + 1024 /* alignment slack */ + 256 /* barriers */;
} // namespace

struct AttnTcParams
{
    __half* out; // [B, S, H * 64]
    int S, H;
};

__global__ void __launch_bounds__(192, 2) attention_bidir_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;                       // 128 rows: two boxes
    uint8_t* sK = sQ + 2 * kTcBox;
    uint8_t* sV = sK + kTcStages * kTcBox;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kTcStages * kTcBox);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                 // [stages]
    uint64_t* v_full = k_full + kTcStages;
    uint64_t* k_free = v_full + kTcStages;
    uint64_t* v_free = k_free + kTcStages;
    uint64_t* s_full = v_free + kTcStages;       // [2] S = Q.K^T of a tile is in TMEM
    uint64_t* s_free = s_full + 2;               // [2] the 128 softmax threads have read it for the last time
    uint64_t* p_full = s_free + 2;               // [2] the 128 softmax threads have written P
    uint64_t* p_free = p_full + 2;               // [2] the tile's P.V chain has finished
    uint64_t* o_done = p_free + 2;               // the last P.V chain has finished
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kTM, h = blockIdx.y, b = blockIdx.z;
    const int S = p.S, H = p.H;
    const int n_tiles = (S + kTN - 1) / kTN;

    if (threadIdx.x == 0)
    {
        mbar_init(q_full, 1);
        for (int s = 0; s < kTcStages; ++s)
        {
            mbar_init(&k_full[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&k_free[s], 1);
            mbar_init(&v_free[s], 1);
        }
        for (int s = 0; s < 2; ++s)
        {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_free[s], 128);
            mbar_init(&p_full[s], 128);
            mbar_init(&p_free[s], 1);
        }
        mbar_init(o_done, 1);
        fence_mbar_init();
    }
    if (warp == 4 && lane == 0)
        tma_prefetch_desc(&tmQKV);
    if (warp == 5)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTcTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 4)
    {
        // ---- TMA producer ----
        if (elect_one_sync())
        {
            grid_dep_wait(); // the projection kernel in front of this one wrote qkv
            mbar_arrive_expect_tx(q_full, 2 * kTcBox);
            tma_load_3d(sQ, &tmQKV, h * kD, q0, b, q_full);
            tma_load_3d(sQ + kTcBox, &tmQKV, h * kD, q0 + 64, b, q_full);
            for (int j = 0; j < n_tiles; ++j)
            {
                const int st = j % kTcStages;
                if (j >= kTcStages)
                    mbar_wait(&k_free[st], ((j / kTcStages) - 1) & 1);
                mbar_arrive_expect_tx(&k_full[st], kTcBox);
                tma_load_3d(sK + st * kTcBox, &tmQKV, (H + h) * kD, j * kTN, b, &k_full[st]);
                if (j >= kTcStages)
                    mbar_wait(&v_free[st], ((j / kTcStages) - 1) & 1);
                mbar_arrive_expect_tx(&v_full[st], kTcBox);
                tma_load_3d(sV + st * kTcBox, &tmQKV, (2 * H + h) * kD, j * kTN, b, &v_full[st]);
            }
        }
    }
    else if (warp == 5)
    {
        // ---- MMA issuer: the whole warp runs the warp-uniform loop, one elected lane issues ----
        grid_dep_launch_dependents();
        mbar_wait(q_full, 0);
        const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ));
        // S_j[128 x 64] = Q[128 x 64] . K_j[64 x 64]^T into score buffer j & 1: two tiles ahead of the softmax warps, so
        // they never wait for the tensor cores
        auto issue_qk = [&](int j)
        {
            const int st = j % kTcStages;
            mbar_wait(&k_full[st], (j / kTcStages) & 1);
            tc_fence_after();
            const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK + st * kTcBox));
            if (elect_one_sync())
            {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    tc_mma_ss(tmem_base + kTcColS + (uint32_t) (j & 1) * kTN, qdesc + 2 * k4, kdesc + 2 * k4, kIdescQK, k4 != 0 ? 1u : 0u);
                tc_commit(&k_free[st]);
                tc_commit(&s_full[j & 1]);
            }
            __syncwarp();
        };
        issue_qk(0);
        if (n_tiles > 1)
            issue_qk(1);
        for (int j = 0; j < n_tiles; ++j)
        {
            const int st = j % kTcStages, bf = j & 1;
            const uint32_t ph = (uint32_t) (j >> 1) & 1;
            mbar_wait(&p_full[bf], ph);
            mbar_wait(&v_full[st], (j / kTcStages) & 1);
            tc_fence_after();
            const uint64_t vdesc = umma_desc_k_sw128(smem_u32(sV + st * kTcBox));
            if (elect_one_sync())
            {
                // O[128 x 64] += P_j[128 x 64] (TMEM, 8 columns per 16 keys) . V_j[64 keys x 64 dims] (MN-major B: 16 keys
                // = two 8-row swizzle atoms of 1024 bytes)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    tc_mma_ts(tmem_base + kTcColO, tmem_base + kTcColP + (uint32_t) bf * 32 + 8 * k4, vdesc + 128 * k4, kIdescPV,
                        (j | k4) != 0 ? 1u : 0u);
                tc_commit(&v_free[st]);
                tc_commit(&p_free[bf]);
                if (j == n_tiles - 1)
                    tc_commit(o_done);
            }
            __syncwarp();
            if (j + 2 < n_tiles)
            {
                mbar_wait(&s_free[bf], ph); // the softmax threads are done with S_j
                issue_qk(j + 2);
            }
        }
    }
    else
    {
        // ---- softmax warps: thread = query row = TMEM lane ----
        const uint32_t trow = tmem_base + ((uint32_t) (warp * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f; // 1/sqrt(64) * log2(e)
        constexpr float kLazy = 8.0f;
        float m = -FLT_MAX, l = 0.f; // m: the maximum (raw score units) the accumulators are expressed in
        for (int j = 0; j < n_tiles; ++j)
        {
            const int bf = j & 1;
            const uint32_t ph = (uint32_t) (j >> 1) & 1;
            const uint32_t tS = trow + kTcColS + (uint32_t) bf * kTN, tP = trow + kTcColP + (uint32_t) bf * 32;
            mbar_wait(&s_full[bf], ph);
            tc_fence_after();
            const int valid = S - j * kTN; // keys of this tile that exist (>= 64 except in the last tile)
            // The row's 64 scores are read from TMEM twice, 32 columns at a time (TMEM reads are nowhere near a limit; a
            // whole row of scores in registers spilled): pass 1 = row maximum, pass 2 = exponentials.  Four independent
            // max / sum chains instead of one long dependent chain per row.
            uint32_t ra[32], rb[32];
            float mx4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
            tc_ld_x32(tS, ra);
            tc_ld_x32(tS + 32, rb);
            tc_wait_ld();
            if (valid < kTN)
            {
                // the last tile only: keys beyond S become -FLT_MAX once, here, so neither pass carries a per-element
                // predicate (ncu: as written before, ISETP + FSEL were 22 % of the kernel's instructions on EVERY tile)
                const uint32_t neg = __float_as_uint(-FLT_MAX);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                {
                    ra[i] = i < valid ? ra[i] : neg;
                    rb[i] = 32 + i < valid ? rb[i] : neg;
                }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i)
            {
                mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(ra[i]));
                mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(rb[i]));
            }
            // every score is in registers now: the Q.K^T of tile j + 2 may overwrite this buffer
            tc_fence_before();
            mbar_arrive(&s_free[bf]);
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            const bool grow = (mx - m) * sl2 > kLazy; // always true on the first tile
            if (__any_sync(0xffffffffu, grow))
            {
                const float nm = grow ? fmaxf(m, mx) : m;
                const float corr = ex2((m - nm) * sl2); // 0 on the first tile, 1 for rows that keep their maximum
                m = nm;
                l *= corr;
                if (j > 0)
                {
                    // O is quiet between the previous tile's P.V chain (p_free) and this tile's (it waits for p_full)
                    mbar_wait(&p_free[bf ^ 1], (uint32_t) ((j - 1) >> 1) & 1);
                    tc_fence_after();
                    uint32_t orr[32];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
                    {
                        tc_ld_x32(trow + kTcColO + 32 * hh, orr);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            orr[i] = __float_as_uint(__uint_as_float(orr[i]) * corr);
                        tc_st_x32p(trow + kTcColO + 32 * hh, orr);
                    }
                }
            }
            // (this P buffer was last read by the P.V chain of tile j - 2: the MMA warp issued it BEFORE the Q.K^T chain of
            // tile j, and s_full[j] -- a tcgen05.commit after that chain -- has been waited for above, so it is done)
            const float bias = -m * sl2;
            float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                const uint32_t* cur = c ? rb : ra;
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2)
                {
                    const float s0 = __uint_as_float(cur[i]), s1 = __uint_as_float(cur[i + 1]);
                    float p0, p1;
                    if (kPolyEvery > 0 && ((i >> 1) % (kPolyEvery > 0 ? kPolyEvery : 1)) == 0)
                    {
                        p0 = ex2_fma(fmaf(s0, sl2, bias));
                        p1 = ex2_fma(fmaf(s1, sl2, bias));
                    }
                    else
                    {
                        p0 = ex2(fmaf(s0, sl2, bias));
                        p1 = ex2(fmaf(s1, sl2, bias));
                    }
                    rs4[(i >> 1) & 3] += p0 + p1;
                    pk[i >> 1] = pack_h2(p0, p1); // keys (2c, 2c + 1) = one 32-bit TMEM column of the A operand
                }
                tc_st_x16(tP + 16 * c, pk);
            }
            l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(&p_full[bf]);
        }
        // ---- finish: O / l -> fp16 -> global (one 128-byte row per thread) ----
        mbar_wait(o_done, 0);
        tc_fence_after();
        const int r = q0 + warp * 32 + lane;
        const float inv = 1.f / l;
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t) b * S + min(r, S - 1)) * H * kD + (size_t) h * kD);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
        {
            uint32_t orr[32];
            tc_ld_x32(trow + kTcColO + 32 * hh, orr);
            tc_wait_ld();
            if (r < S)
            {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                {
                    uint4 v;
                    v.x = pack_h2(__uint_as_float(orr[8 * c + 0]) * inv, __uint_as_float(orr[8 * c + 1]) * inv);
                    v.y = pack_h2(__uint_as_float(orr[8 * c + 2]) * inv, __uint_as_float(orr[8 * c + 3]) * inv);
                    v.z = pack_h2(__uint_as_float(orr[8 * c + 4]) * inv, __uint_as_float(orr[8 * c + 5]) * inv);
                    v.w = pack_h2(__uint_as_float(orr[8 * c + 6]) * inv, __uint_as_float(orr[8 * c + 7]) * inv);
                    dst[4 * hh + c] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5)
    {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
    }
}
} // namespace b200

using namespace b200;

static int enc_attn_mode()
{
    static int mode = -1; // 0 = tcgen05 (default), 1 = mma.sync (B200_ENC_ATTN=mma)
    if (mode < 0)
    {
        const char* e = getenv("B200_ENC_ATTN");
        mode = (e != nullptr && e[0] == 'm') ? 1 : 0;
    }
    return mode;
}

extern "C" int b200_attention_bidirectional_fp16(const void* qkv, void* out, int batch_size, int seq_len, int num_heads,
    int head_size, b200_stream_t stream)
{
    B200_REQUIRE(qkv && out, B200_ERR_INVALID_ARG, "null pointer (qkv/out)");
    B200_REQUIRE(head_size == 64, B200_ERR_UNSUPPORTED, "head_size %d unsupported (only 64)", head_size);
    B200_REQUIRE(batch_size >= 0 && seq_len >= 0 && num_heads > 0, B200_ERR_INVALID_ARG, "bad sizes");
    if (batch_size == 0 || seq_len == 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    if (enc_attn_mode() == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0)
    {
        // 3-D view [B][S][3*H*64] of the fused projection; a box = 64 dims x 64 rows of one batch element, rows past S
        // zero-filled
        CUtensorMap tm;
        const uint64_t row_bytes = (uint64_t) 3 * num_heads * kD * 2;
        if (int rc = make_tmap_3d(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, qkv, (uint64_t) 3 * num_heads * kD, (uint64_t) seq_len,
                (uint64_t) batch_size, row_bytes, row_bytes * (uint64_t) seq_len, kD, 64, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B))
            return rc;
        static bool attr_tc = false;
        if (!attr_tc)
        {
            B200_CUDA(cudaFuncSetAttribute(attention_bidir_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTcSmem));
            B200_CUDA(cudaFuncSetAttribute(attention_bidir_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                cudaSharedmemCarveoutMaxShared));
            attr_tc = true;
        }
        AttnTcParams prm{static_cast<__half*>(out), seq_len, num_heads};
        B200_LAUNCH(attention_bidir_tc_kernel, dim3((seq_len + kTM - 1) / kTM, num_heads, batch_size), dim3(192), kTcSmem,
            as_stream(stream), tm, prm);
        return B200_OK;
    }
    const size_t smem = 5 * (size_t) kTile;
    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(attention_bidir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    B200_LAUNCH(attention_bidir_kernel, dim3((seq_len + kBM - 1) / kBM, num_heads, batch_size), dim3(128), smem,
        as_stream(stream), static_cast<const __half*>(qkv), static_cast<__half*>(out), seq_len, num_heads);
    return B200_OK;
}
