// woq_gemv.cu -- fp16 activations x per-channel int8 weights for small M (decode GEMV), SIMT path.
//
// Replaces weight_only_gemv_launcher / int8_weight_only_gemv_interleave
//   T/cpp/tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.cu:136-205,371-378
// on the reference's exact preprocessed layout ([N/2][2K] bytes, see quantize.cu).
//
// B200 design (HBM-bound streaming of K*N weight bytes, read exactly once):
//  * The weight matrix is one contiguous byte array and every CTA owns a contiguous range of row pairs, so a
//    CTA's slab is a single contiguous region: a producer warp streams it with 1-D TMA bulk copies
//    (cp.async.bulk, SASS UBLKCP) into a shared-memory ring guarded by full/empty mbarriers.  Weights do not
//    depend on the previous kernel, so the producer starts before griddepcontrol.wait (PDL): the slab is in
//    flight while the producer kernel of the activations is still finishing.
//  * K is split across the consumer warps: warp w owns the w-th 512-byte piece (256 k's of both interleaved
//    columns) of every row pair, so its 16 activations per lane per row of A live in registers for the whole
//    kernel and shared memory only carries the int8 stream (128-bit LDS per lane, conflict-free: consecutive
//    lanes read consecutive 16-byte chunks).
//  * In-register dequant: PRMT + HSUB2 (0x6400|b trick) then HMUL2 by the fp16 column scale, i.e. the same
//    effective weight fp16(q * s) as both reference kernels (.cu:44-53).  Products are chained four at a time
//    with HFMA2 and flushed to fp32 accumulators; lanes are reduced with xor-shuffles 16, 8, 2, 1 exactly as the
//    reference lane layout requires (.cu:190-193), warps through shared memory.
//  * Epilogue (bias / GELU / residual) is fused; the reference needs separate TensorRT layers for those.
#include "common.cuh"

namespace b200
{

struct GemvParams
{
    const __half* A;
    const uint8_t* W;
    const __half* scales;
    const __half* bias;
    const __half* residual;
    __half* C;
    int K, N;
    int activation;
    int row_pairs_total; // N / 2
    int rows_per_cta;    // row pairs per CTA
    int rows_per_stage;  // row pairs per ring stage
    int num_stages;
    int num_pieces;      // ceil(2K / 512)
    int lda, ldc;        // row strides of A and C (elements)
};

constexpr int kGemvMaxStages = 8;

template <int M, int PPW>
__global__ void __launch_bounds__(672) woq_gemv_kernel(const GemvParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int nw = (blockDim.x >> 5) - 1; // consumer warps
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int row_bytes = 2 * p.K;
    const int stage_bytes = p.rows_per_stage * row_bytes;

    uint8_t* ring = smem;
    float* partial = reinterpret_cast<float*>(smem + (size_t) p.num_stages * stage_bytes);
    // partial: [nw][rows_per_cta * 2][M]
    uint64_t* full = reinterpret_cast<uint64_t*>(partial + (size_t) nw * p.rows_per_cta * 2 * M);
    uint64_t* empty = full + kGemvMaxStages;

    const int row0 = blockIdx.x * p.rows_per_cta;
    const int rows = min(p.rows_per_cta, p.row_pairs_total - row0);
    const int num_chunks = (rows + p.rows_per_stage - 1) / p.rows_per_stage;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < p.num_stages; ++s)
        {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nw);
        }
        fence_mbar_init();
    }
    __syncthreads();
    grid_dep_launch_dependents();

    if (warp == nw)
    {
        // ===== producer: stream this CTA's contiguous weight slab =====
        if (lane == 0)
        {
            const uint64_t pol = policy_evict_first(); // weights are read once per step
            const uint8_t* src = p.W + (size_t) row0 * row_bytes;
            for (int it = 0; it < num_chunks; ++it)
            {
                const int s = it % p.num_stages;
                if (it >= p.num_stages)
                    mbar_wait(&empty[s], ((it / p.num_stages) - 1) & 1);
                const int r = min(p.rows_per_stage, rows - it * p.rows_per_stage);
                const uint32_t bytes = (uint32_t) r * row_bytes;
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s_hint(ring + (size_t) s * stage_bytes, src + (size_t) it * stage_bytes, bytes, &full[s], pol);
            }
        }
        return;
    }

    // ===== consumers =====
    // Activations come from the previous kernel: wait for it (no-op without PDL).
    grid_dep_wait();

    // lane geometry inside a 512-byte piece: t-block = lane/8, column parity = (lane%8)/4, 16-byte chunk = lane%4
    const int par = (lane >> 2) & 1;
    __half2 x[PPW][M][8];
    int off[PPW]; // byte offset of this lane's chunk inside a row pair, or -1
#pragma unroll
    for (int i = 0; i < PPW; ++i)
    {
        const int piece = warp + nw * i;
        const int o = piece * 512 + lane * 16;
        off[i] = (piece < p.num_pieces && o < row_bytes) ? o : -1;
        const int kbase = 64 * (o >> 7) + 16 * (lane & 3);
#pragma unroll
        for (int m = 0; m < M; ++m)
        {
            uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
            if (off[i] >= 0)
            {
                const uint4* ap = reinterpret_cast<const uint4*>(p.A + (size_t) m * p.lda + kbase);
                v0 = __ldg(ap);
                v1 = __ldg(ap + 1);
            }
            x[i][m][0] = *reinterpret_cast<__half2*>(&v0.x);
            x[i][m][1] = *reinterpret_cast<__half2*>(&v0.y);
            x[i][m][2] = *reinterpret_cast<__half2*>(&v0.z);
            x[i][m][3] = *reinterpret_cast<__half2*>(&v0.w);
            x[i][m][4] = *reinterpret_cast<__half2*>(&v1.x);
            x[i][m][5] = *reinterpret_cast<__half2*>(&v1.y);
            x[i][m][6] = *reinterpret_cast<__half2*>(&v1.z);
            x[i][m][7] = *reinterpret_cast<__half2*>(&v1.w);
        }
    }

    for (int it = 0; it < num_chunks; ++it)
    {
        const int s = it % p.num_stages;
        mbar_wait(&full[s], (it / p.num_stages) & 1);
        const uint8_t* stage = ring + (size_t) s * stage_bytes;
        const int r_in = min(p.rows_per_stage, rows - it * p.rows_per_stage);
        for (int r = 0; r < r_in; ++r)
        {
            const int rl = it * p.rows_per_stage + r; // row pair local to the CTA
            const __half sc = p.scales[2 * (row0 + rl) + par];
            const __half2 sc2 = __half2half2(sc);
            float acc[M];
#pragma unroll
            for (int m = 0; m < M; ++m)
                acc[m] = 0.f;
#pragma unroll
            for (int i = 0; i < PPW; ++i)
            {
                if (off[i] < 0)
                    continue;
                const uint4 wv = *reinterpret_cast<const uint4*>(stage + (size_t) r * row_bytes + off[i]);
                __half2 w2[8];
                dequant_word(wv.x, w2[0], w2[4]);
                dequant_word(wv.y, w2[1], w2[5]);
                dequant_word(wv.z, w2[2], w2[6]);
                dequant_word(wv.w, w2[3], w2[7]);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    w2[j] = __hmul2(w2[j], sc2); // effective weight fp16(q * s), as the reference
#pragma unroll
                for (int m = 0; m < M; ++m)
                {
                    __half2 h0 = __hmul2(x[i][m][0], w2[0]);
                    __half2 h1 = __hmul2(x[i][m][4], w2[4]);
                    h0 = __hfma2(x[i][m][1], w2[1], h0);
                    h1 = __hfma2(x[i][m][5], w2[5], h1);
                    h0 = __hfma2(x[i][m][2], w2[2], h0);
                    h1 = __hfma2(x[i][m][6], w2[6], h1);
                    h0 = __hfma2(x[i][m][3], w2[3], h0);
                    h1 = __hfma2(x[i][m][7], w2[7], h1);
                    const float2 f0 = __half22float2(h0);
                    const float2 f1 = __half22float2(h1);
                    acc[m] += (f0.x + f0.y) + (f1.x + f1.y);
                }
            }
#pragma unroll
            for (int m = 0; m < M; ++m)
            {
                float v = acc[m];
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                if ((lane & ~4) == 0) // lanes 0 and 4: columns 2j and 2j+1
                    partial[((size_t) warp * p.rows_per_cta * 2 + rl * 2 + par) * M + m] = v;
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(&empty[s]);
    }

    // cross-warp reduction + fused epilogue (consumer warps only: named barrier 1)
    asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
    const int outs = rows * 2 * M;
    for (int o = threadIdx.x; o < outs; o += nw * 32)
    {
        const int m = o % M;
        const int c = o / M; // column local to the CTA
        float v = 0.f;
        for (int w = 0; w < nw; ++w)
            v += partial[((size_t) w * p.rows_per_cta * 2 + c) * M + m];
        const int n = 2 * row0 + c;
        const size_t idx = (size_t) m * p.ldc + n;
        p.C[idx] = epilogue_apply(v, 1.0f, p.bias, p.activation, p.residual, n, idx);
    }
}

template <int M, int PPW>
static int launch_gemv(const GemvParams& p, int nw, int grid, size_t smem, cudaStream_t stream)
{
    auto kern = woq_gemv_kernel<M, PPW>;
    if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    B200_LAUNCH(kern, dim3(grid), dim3((nw + 1) * 32), smem, stream, p);
    return B200_OK;
}

template <int M>
static int dispatch_ppw(const GemvParams& p, int ppw, int nw, int grid, size_t smem, cudaStream_t stream)
{
    switch (ppw)
    {
    case 1: return launch_gemv<M, 1>(p, nw, grid, smem, stream);
    case 2: return launch_gemv<M, 2>(p, nw, grid, smem, stream);
    case 3: return launch_gemv<M, 3>(p, nw, grid, smem, stream);
    case 4: return launch_gemv<M, 4>(p, nw, grid, smem, stream);
    default: set_error("woq gemv: K=%d too large for the SIMT path", p.K); return B200_ERR_UNSUPPORTED;
    }
}

// SIMT path entry: any M (processed in chunks of <= 4 rows; intended for M <= 4).
int woq_gemv_simt(const __half* A, int M, int K, const uint8_t* W, const __half* scales, int N, const __half* bias,
    int activation, const __half* residual, __half* C, cudaStream_t stream)
{
    GemvParams p{};
    p.W = W;
    p.scales = scales;
    p.bias = bias;
    p.activation = activation;
    p.K = K;
    p.N = N;
    p.lda = K;
    p.ldc = N;
    p.row_pairs_total = N / 2;
    p.num_pieces = (2 * K + 511) / 512;

    // K split: one 512-byte piece per warp up to 20 warps, then 2..4 pieces per warp.
    int ppw = 1;
    while ((p.num_pieces + ppw - 1) / ppw > 20)
        ++ppw;
    const int nw = (p.num_pieces + ppw - 1) / ppw;
    // with M = 4 and 3+ pieces per warp the activations no longer fit in registers: fall back to 2-row chunks
    const int mchunk_max = (ppw >= 3) ? 2 : 4;

    const int sms = num_sms();
    const int row_bytes = 2 * K;
    // aim for ~2 co-resident CTAs per SM so the tail of one overlaps the ramp of the other
    int rows_per_cta = (p.row_pairs_total + 2 * sms - 1) / (2 * sms);
    if (rows_per_cta < 2)
        rows_per_cta = 2;
    // keep the whole slab in the ring when it is small; otherwise 8 stages of >= 4 KB
    int rows_per_stage = (4096 + row_bytes - 1) / row_bytes;
    if (rows_per_stage < 1)
        rows_per_stage = 1;
    if (rows_per_stage > rows_per_cta)
        rows_per_stage = rows_per_cta;
    int stages = (rows_per_cta + rows_per_stage - 1) / rows_per_stage;
    if (stages > kGemvMaxStages)
        stages = kGemvMaxStages;
    while ((size_t) stages * rows_per_stage * row_bytes > 96 * 1024 && stages > 2)
        --stages;
    B200_REQUIRE((size_t) stages * rows_per_stage * row_bytes <= 200 * 1024, B200_ERR_UNSUPPORTED,
        "woq gemv: K=%d too large for the shared-memory ring", K);
    p.rows_per_cta = rows_per_cta;
    p.rows_per_stage = rows_per_stage;
    p.num_stages = stages;
    const int grid = (p.row_pairs_total + rows_per_cta - 1) / rows_per_cta;

    for (int m0 = 0; m0 < M; m0 += mchunk_max)
    {
        const int mc = (M - m0 < mchunk_max) ? (M - m0) : mchunk_max;
        p.A = A + (size_t) m0 * K;
        p.C = C + (size_t) m0 * N;
        p.residual = residual ? residual + (size_t) m0 * N : nullptr;
        const size_t smem = (size_t) stages * rows_per_stage * row_bytes
            + sizeof(float) * (size_t) nw * rows_per_cta * 2 * mc + sizeof(uint64_t) * 2 * kGemvMaxStages;
        int rc;
        switch (mc)
        {
        case 1: rc = dispatch_ppw<1>(p, ppw, nw, grid, smem, stream); break;
        case 2: rc = dispatch_ppw<2>(p, ppw, nw, grid, smem, stream); break;
        case 3: rc = dispatch_ppw<3>(p, ppw, nw, grid, smem, stream); break;
        default: rc = dispatch_ppw<4>(p, ppw, nw, grid, smem, stream); break;
        }
        if (rc != B200_OK)
            return rc;
    }
    return B200_OK;
}

} // namespace b200
