// attention_encoder.cu -- bidirectional (encoder) multi-head attention over the 1500 audio frames, head size 64.
//
// Replaces the unfused attention of the reference's encoder blocks (T/tensorrt_llm/layers/attention.py:283-406 with
// no mask / no KV cache, used by T/tensorrt_llm/models/whisper/model.py:124-172; oracle W/torch_model.py:88-103):
//     softmax(q k^T / sqrt(64)) v        q, k, v [B, S, H, 64] taken from one fused projection [B, S, 3*H*64].
// Flash-attention style: one CTA per (64 query rows, head, batch), 4 warps x 16 rows; K and V tiles of 64 keys are
// double-buffered in shared memory with cp.async (16-byte chunks XOR-swizzled by the row so every ldmatrix is
// conflict-free); scores and P.V run on the tensor cores (mma.sync.m16n8k16, fp32 accumulate), the online softmax
// lives in registers (ex2 with the scale folded into one FFMA, lazily rescaled accumulators), the score accumulators are re-used in place as the fp16 A fragments of P.
// Round 1 uses the warp-level mma.sync path (the encoder runs once per utterance, off the decoder-step metric); a
// tcgen05 / TMEM version of this kernel is the natural round-2 upgrade.
#include <float.h>

#include "common.cuh"

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
} // namespace b200

using namespace b200;

extern "C" int b200_attention_bidirectional_fp16(const void* qkv, void* out, int batch_size, int seq_len, int num_heads,
    int head_size, b200_stream_t stream)
{
    B200_REQUIRE(qkv && out, B200_ERR_INVALID_ARG, "null pointer (qkv/out)");
    B200_REQUIRE(head_size == 64, B200_ERR_UNSUPPORTED, "head_size %d unsupported (only 64)", head_size);
    B200_REQUIRE(batch_size >= 0 && seq_len >= 0 && num_heads > 0, B200_ERR_INVALID_ARG, "bad sizes");
    if (batch_size == 0 || seq_len == 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
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
