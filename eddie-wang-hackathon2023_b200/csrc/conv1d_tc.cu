// conv1d_tc.cu -- Conv1d encoder stem as an implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces functional.conv1d -> TensorRT IConvolutionLayer (T/tensorrt_llm/functional.py:2202-2244,
// T/tensorrt_llm/layers/conv.py:52-94); Whisper: conv1 80->1280 k3 s1 p1 and conv2 1280->1280 k3 s2 p1, each followed
// by GELU (T/tensorrt_llm/models/whisper/model.py:135-157; oracle T/examples/whisper/torch_model.py:143-144,157-158).
//
//   y[b, co, t] = bias[co] + sum_{tap, ci} w[co, ci, tap] * x[b, ci, t*stride + tap - pad]
//
// "Swap-AB" like the other GEMMs of this library: the 128 output channels of a tile are the UMMA M dimension (lanes of
// the TMEM accumulator), NT = 256 consecutive output time steps are UMMA N, and K runs over (tap, input channel):
//   * A operand: the weights, re-laid once per call as Wt[tap][co][ci_pad] (K-major, ci padded to a multiple of 64
//     with zeros); one 2-D TMA box (64 ci x 128 co, 128B swizzle) per k-block;
//   * B operand: the input, transposed once per call to time-major Xt[b][t][ci_pad]; one 3-D TMA box per k-block whose
//     row coordinate starts at t0*stride + tap - pad and whose ROW TRAVERSAL STRIDE is the convolution stride, so the
//     im2col gather (every second time step for conv2) and the zero padding at both ends (out-of-bounds rows are
//     zero-filled) are done by the TMA unit -- no im2col buffer exists;
//   * warp 4 = TMA producer, warp 5 = MMA issuer (tcgen05.mma.kind::f16, SS form, fp32 accumulator in TMEM),
//     warps 0-3 = epilogue: tcgen05.ld, + bias, GELU, fp16, 8-byte stores along t (the [B, Cout, Tout] layout of the
//     reference is the natural one for lanes = channels).
// The round-1 SIMT kernel (conv1d.cu) stays as the fallback for calls without a workspace.
#include "tcgen05.cuh"

namespace b200
{

struct ConvTcParams
{
    const __half* bias;
    __half* y;
    int Cout, Tout, cin_blocks, ksize, stride, pad, activation;
};

constexpr int kConvATile = 128 * 128; // 128 output channels x 64 k (fp16)

template <int NT, int SS>
__global__ void __launch_bounds__(192, 1)
    conv1d_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const ConvTcParams p)
{
    constexpr int BTile = NT * 128;
    constexpr uint32_t kTmemCols = tmem_cols_pow2(NT);
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t) (NT >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smA = smem;
    uint8_t* smB = smem + SS * kConvATile;
    uint64_t* full = reinterpret_cast<uint64_t*>(smB + SS * BTile);
    uint64_t* smem_free = full + SS;
    uint64_t* acc_done = smem_free + SS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * NT, co0 = blockIdx.y * 128, b = blockIdx.z;
    const int nkb = p.ksize * p.cin_blocks;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < SS; ++s)
        {
            mbar_init(&full[s], 1);
            mbar_init(&smem_free[s], 1);
        }
        mbar_init(acc_done, 1);
        fence_mbar_init();
    }
    if (warp == 4 && lane == 0)
    {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 5)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 4)
    {
        if (elect_one_sync())
        {
            for (int i = 0; i < nkb; ++i)
            {
                const int ss = i % SS;
                if (i >= SS)
                    mbar_wait(&smem_free[ss], ((i / SS) - 1) & 1);
                const int tap = i / p.cin_blocks, cb = i - tap * p.cin_blocks;
                mbar_arrive_expect_tx(&full[ss], kConvATile + BTile);
                tma_load_2d(smA + ss * kConvATile, &tmW, cb * 64, tap * p.Cout + co0, &full[ss]);
                // a TMA box dimension is at most 256 tensor elements: with stride 2 one box yields 128 output rows, so
                // the NT rows of the tile arrive as NT / 128 boxes
#pragma unroll
                for (int hb = 0; hb < NT / 128; ++hb)
                    tma_load_3d(smB + ss * BTile + hb * 128 * 128, &tmX, cb * 64, (t0 + hb * 128) * p.stride + tap - p.pad, b,
                        &full[ss]);
            }
        }
    }
    else if (warp == 5)
    {
        for (int i = 0; i < nkb; ++i)
        {
            const int ss = i % SS;
            mbar_wait(&full[ss], (i / SS) & 1);
            tc_fence_after();
            const uint64_t adesc = umma_desc_k_sw128(smem_u32(smA + ss * kConvATile));
            const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smB + ss * BTile));
            if (elect_one_sync())
            {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    tc_mma_ss(tmem_base, adesc + 2 * k4, bdesc + 2 * k4, kIdesc, (i | k4) != 0 ? 1u : 0u);
                tc_commit(&smem_free[ss]);
                if (i == nkb - 1)
                    tc_commit(acc_done);
            }
            __syncwarp();
        }
    }
    else
    {
        // ---- epilogue: thread = output channel (TMEM lane), 16 time steps per tcgen05.ld ----
        const int co = co0 + threadIdx.x;
        const float bv = (p.bias != nullptr && co < p.Cout) ? __half2float(__ldg(p.bias + co)) : 0.f;
        mbar_wait(acc_done, 0);
        tc_fence_after();
        const uint32_t lane_field = (uint32_t) (warp * 32) << 16;
        __half* yrow = p.y + ((size_t) b * p.Cout + co) * p.Tout;
        const bool vec = (p.Tout & 3) == 0; // every row start is 8-byte aligned
#pragma unroll 1
        for (int c16 = 0; c16 < NT / 16; ++c16)
        {
            uint32_t acc[16];
            tc_ld_x16(tmem_base + lane_field + c16 * 16, acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int t = t0 + c16 * 16;
            if (co >= p.Cout || t >= p.Tout)
                continue;
            __align__(8) __half o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
            {
                __half h = __float2half_rn(__uint_as_float(acc[i]) + bv);
                if (p.activation == B200_ACT_GELU_ERF)
                    h = __float2half_rn(gelu_erf(__half2float(h)));
                else if (p.activation == B200_ACT_GELU_TANH)
                    h = __float2half_rn(gelu_tanh(__half2float(h)));
                o[i] = h;
            }
            if (vec && t + 16 <= p.Tout)
            {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint2*>(yrow + t + 4 * q) = *reinterpret_cast<const uint2*>(&o[4 * q]);
            }
            else
            {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (t + i < p.Tout)
                        yrow[t + i] = o[i];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5)
    {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// w [Cout][Cin][k] -> Wt [k][Cout][Cin_pad] (zero padded input channels)
__global__ void conv_relayout_w_kernel(const __half* __restrict__ w, __half* __restrict__ wt, int Cout, int Cin, int cin_pad,
    int ksize)
{
    const size_t total = (size_t) ksize * Cout * cin_pad;
    for (size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t) gridDim.x * blockDim.x)
    {
        const int ci = (int) (idx % cin_pad);
        const int co = (int) ((idx / cin_pad) % Cout);
        const int tap = (int) (idx / ((size_t) cin_pad * Cout));
        wt[idx] = ci < Cin ? w[((size_t) co * Cin + ci) * ksize + tap] : __float2half(0.f);
    }
}

// x [B][Cin][T] -> Xt [B][T][Cin_pad] (zero padded input channels); 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) conv_transpose_x_kernel(const __half* __restrict__ x, __half* __restrict__ xt, int Cin,
    int cin_pad, int T)
{
    __shared__ __half tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 8 rows per pass
    const __half* xb = x + (size_t) b * Cin * T;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int c = c0 + ty + 8 * r, t = t0 + tx;
        tile[ty + 8 * r][tx] = (c < Cin && t < T) ? xb[(size_t) c * T + t] : __float2half(0.f);
    }
    __syncthreads();
    __half* xo = xt + (size_t) b * T * cin_pad;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int t = t0 + ty + 8 * r, c = c0 + tx;
        if (t < T && c < cin_pad)
            xo[(size_t) t * cin_pad + c] = tile[tx][ty + 8 * r];
    }
}

} // namespace b200

using namespace b200;

static inline int conv_cin_pad(int c_in)
{
    return (c_in + 63) / 64 * 64;
}

extern "C" size_t b200_conv1d_workspace_bytes(int batch_size, int c_in, int c_out, int t_in, int ksize)
{
    if (batch_size <= 0 || c_in <= 0 || c_out <= 0 || t_in <= 0 || ksize <= 0)
        return 0;
    const size_t cp = (size_t) conv_cin_pad(c_in);
    const size_t wt = ((size_t) ksize * c_out * cp * sizeof(__half) + 1023) & ~size_t(1023);
    return wt + (size_t) batch_size * t_in * cp * sizeof(__half);
}

extern "C" int b200_conv1d_fp16_tc(const void* x, const void* w, const void* bias, void* y, int batch_size, int c_in,
    int c_out, int t_in, int ksize, int stride, int pad, int activation, void* workspace, size_t workspace_bytes,
    b200_stream_t stream)
{
    B200_REQUIRE(x && w && y, B200_ERR_INVALID_ARG, "null pointer (x/w/y)");
    B200_REQUIRE(c_in > 0 && c_out > 0 && t_in > 0, B200_ERR_INVALID_ARG, "bad sizes");
    B200_REQUIRE(ksize >= 1 && ksize <= 8, B200_ERR_UNSUPPORTED, "kernel size %d unsupported (1..8)", ksize);
    B200_REQUIRE(stride == 1 || stride == 2, B200_ERR_UNSUPPORTED, "stride %d unsupported (1 or 2)", stride);
    B200_REQUIRE(pad >= 0 && pad < ksize, B200_ERR_INVALID_ARG, "pad %d must be in [0, ksize)", pad);
    B200_REQUIRE(activation >= B200_ACT_NONE && activation <= B200_ACT_GELU_TANH, B200_ERR_INVALID_ARG,
        "unknown activation %d", activation);
    const int t_out = (t_in + 2 * pad - ksize) / stride + 1;
    B200_REQUIRE(t_out > 0, B200_ERR_INVALID_ARG, "empty output");
    if (batch_size <= 0)
        return B200_OK;
    const size_t need = b200_conv1d_workspace_bytes(batch_size, c_in, c_out, t_in, ksize);
    B200_REQUIRE(workspace != nullptr && workspace_bytes >= need, B200_ERR_WORKSPACE,
        "conv1d (tensor-core path): workspace of %zu bytes needed, got %zu", need, workspace_bytes);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, B200_ERR_INVALID_ARG,
        "conv1d: workspace must be 256-byte aligned");
    B200_REQUIRE_DEVICE();
    cudaStream_t st = as_stream(stream);
    const int cp = conv_cin_pad(c_in);
    __half* wt = static_cast<__half*>(workspace);
    __half* xt = reinterpret_cast<__half*>(
        static_cast<char*>(workspace) + (((size_t) ksize * c_out * cp * sizeof(__half) + 1023) & ~size_t(1023)));

    {
        const size_t total = (size_t) ksize * c_out * cp;
        const int blocks = (int) ((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        conv_relayout_w_kernel<<<blocks, 256, 0, st>>>(static_cast<const __half*>(w), wt, c_out, c_in, cp, ksize);
        B200_LAUNCH_CHECK();
        const dim3 grid((t_in + 31) / 32, cp / 32, batch_size);
        conv_transpose_x_kernel<<<grid, 256, 0, st>>>(static_cast<const __half*>(x), xt, c_in, cp, t_in);
        B200_LAUNCH_CHECK();
    }

    constexpr int NT = 256, SS = 4; // 256 output time steps per CTA: halves the weight re-reads from L2 per FLOP
    CUtensorMap tmW, tmX;
    if (int rc = make_tmap_2d(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, wt, (uint64_t) cp, (uint64_t) ksize * c_out,
            (uint64_t) cp * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    if (int rc = make_tmap_3d(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, xt, (uint64_t) cp, (uint64_t) t_in, (uint64_t) batch_size,
            (uint64_t) cp * 2, (uint64_t) t_in * cp * 2, 64, (uint32_t) (128 * stride), 1, (uint32_t) stride,
            CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    ConvTcParams p{};
    p.bias = static_cast<const __half*>(bias);
    p.y = static_cast<__half*>(y);
    p.Cout = c_out;
    p.Tout = t_out;
    p.cin_blocks = cp / 64;
    p.ksize = ksize;
    p.stride = stride;
    p.pad = pad;
    p.activation = activation;
    auto kern = conv1d_tc_kernel<NT, SS>;
    const size_t smem = 1024 + (size_t) SS * (kConvATile + NT * 128) + sizeof(uint64_t) * (2 * SS + 1) + 16;
    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    const dim3 grid((t_out + NT - 1) / NT, (c_out + 127) / 128, batch_size);
    kern<<<grid, 192, smem, st>>>(tmW, tmX, p);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
