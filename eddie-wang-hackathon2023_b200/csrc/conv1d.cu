// conv1d.cu -- Conv1d encoder stem (fp16 in/out, fp32 accumulate) with optional fused GELU.
//
// Replaces functional.conv1d -> TensorRT IConvolutionLayer (T/tensorrt_llm/functional.py:2202-2244,
// T/tensorrt_llm/layers/conv.py:52-94); Whisper uses conv1 80->d k3 s1 p1 and conv2 d->d k3 s2 p1, each followed
// by GELU (T/tensorrt_llm/models/whisper/model.py:135-157; oracle T/examples/whisper/torch_model.py:143-144,157-158).
//
// This file: shared-memory tiled direct convolution on CUDA cores (64 out-channels x 64 time steps per CTA, 4x4
// register tile per thread) -- the workspace-free fallback.  The production path is the tcgen05 implicit GEMM in
// conv1d_tc.cu (38x faster on conv2 at batch 16).
#include "common.cuh"

namespace b200
{

constexpr int kCoTile = 64, kTTile = 64, kCiChunk = 16, kMaxK = 3;

template <int STRIDE>
__global__ void __launch_bounds__(256) conv1d_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
    const __half* __restrict__ bias, __half* __restrict__ y, int Cin, int Cout, int Tin, int Tout, int ksize, int pad,
    int activation)
{
    constexpr int XW = kTTile * STRIDE + kMaxK; // staged input width (covers (kTTile-1)*STRIDE + ksize)
    __shared__ float ws[kCiChunk][kMaxK][kCoTile];
    __shared__ float xs[kCiChunk][XW];

    const int t0 = blockIdx.x * kTTile, co0 = blockIdx.y * kCoTile, b = blockIdx.z;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4; // tx -> time, ty -> out channel
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j] = 0.f;

    const __half* xb = x + (size_t) b * Cin * Tin;
    const int tin0 = t0 * STRIDE - pad;

    for (int ci0 = 0; ci0 < Cin; ci0 += kCiChunk)
    {
        for (int idx = threadIdx.x; idx < kCiChunk * kMaxK * kCoTile; idx += 256)
        {
            const int co = idx % kCoTile;
            const int kk = (idx / kCoTile) % kMaxK;
            const int ci = idx / (kCoTile * kMaxK);
            float v = 0.f;
            if (ci0 + ci < Cin && co0 + co < Cout && kk < ksize)
                v = __half2float(w[((size_t) (co0 + co) * Cin + ci0 + ci) * ksize + kk]);
            ws[ci][kk][co] = v;
        }
        for (int idx = threadIdx.x; idx < kCiChunk * XW; idx += 256)
        {
            const int j = idx % XW, ci = idx / XW;
            const int tin = tin0 + j;
            float v = 0.f;
            if (ci0 + ci < Cin && tin >= 0 && tin < Tin)
                v = __half2float(xb[(size_t) (ci0 + ci) * Tin + tin]);
            xs[ci][j] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int ci = 0; ci < kCiChunk; ++ci)
        {
#pragma unroll
            for (int kk = 0; kk < kMaxK; ++kk)
            {
                float wv[4], xv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    wv[i] = ws[ci][kk][ty + 16 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    xv[j] = xs[ci][(tx + 16 * j) * STRIDE + kk];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        const int co = co0 + ty + 16 * i;
        if (co >= Cout)
            continue;
        const float bv = bias ? __half2float(bias[co]) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const int t = t0 + tx + 16 * j;
            if (t >= Tout)
                continue;
            __half o = __float2half_rn(acc[i][j] + bv);
            if (activation == B200_ACT_GELU_ERF)
                o = __float2half_rn(gelu_erf(__half2float(o)));
            else if (activation == B200_ACT_GELU_TANH)
                o = __float2half_rn(gelu_tanh(__half2float(o)));
            y[((size_t) b * Cout + co) * Tout + t] = o;
        }
    }
}

} // namespace b200

using namespace b200;

extern "C" int b200_conv1d_fp16(const void* x, const void* w, const void* bias, void* y, int batch_size, int c_in,
    int c_out, int t_in, int ksize, int stride, int pad, int activation, b200_stream_t stream)
{
    B200_REQUIRE(x && w && y, B200_ERR_INVALID_ARG, "null pointer (x/w/y)");
    B200_REQUIRE(c_in > 0 && c_out > 0 && t_in > 0, B200_ERR_INVALID_ARG, "bad sizes");
    B200_REQUIRE(ksize >= 1 && ksize <= kMaxK, B200_ERR_UNSUPPORTED, "kernel size %d unsupported (1..3)", ksize);
    B200_REQUIRE(stride == 1 || stride == 2, B200_ERR_UNSUPPORTED, "stride %d unsupported (1 or 2)", stride);
    B200_REQUIRE(pad >= 0 && pad < ksize, B200_ERR_INVALID_ARG, "pad %d must be in [0, ksize)", pad);
    B200_REQUIRE(activation >= B200_ACT_NONE && activation <= B200_ACT_GELU_TANH, B200_ERR_INVALID_ARG,
        "unknown activation %d", activation);
    const int t_out = (t_in + 2 * pad - ksize) / stride + 1;
    B200_REQUIRE(t_out > 0, B200_ERR_INVALID_ARG, "empty output");
    if (batch_size <= 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    const dim3 grid((t_out + kTTile - 1) / kTTile, (c_out + kCoTile - 1) / kCoTile, batch_size);
    const __half *xh = static_cast<const __half*>(x), *wh = static_cast<const __half*>(w), *bh = static_cast<const __half*>(bias);
    if (stride == 1)
        conv1d_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(xh, wh, bh, static_cast<__half*>(y), c_in, c_out, t_in, t_out, ksize, pad, activation);
    else
        conv1d_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(xh, wh, bh, static_cast<__half*>(y), c_in, c_out, t_in, t_out, ksize, pad, activation);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
