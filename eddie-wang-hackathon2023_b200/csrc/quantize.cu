// quantize.cu -- int8 weight-only quantizer + "preprocess for mixed gemm" layout transform on the GPU.
//
// Bit-exact replacement for the reference's single-threaded host code:
//   symmetric_quantize (int8)           T/cpp/tensorrt_llm/kernels/cutlass_kernels/cutlass_preprocessors.cpp:615-721
//   preprocess_weights_for_mixed_gemm   same file :537-578 (Sm80 layout details)
// HBM-bound byte work: W[K,N] is read twice (amax pass, quantize pass) with 128-byte coalesced rows, the
// processed layout is written once as 128-byte segments.  One 64(k) x 64(n) tile per CTA is staged in shared
// memory so both the read and the permuted write are coalesced.
//
// Processed layout, closed form (checked against the reference binary in tests/):
//   proc viewed as [N/2][2K] bytes.  Row j, byte o:  t = o/128, w = o%128, n = 2j + w/64, kk = w%64,
//   k' = 64t + 4(kk/4) + {0,2,1,3}[kk%4],  k = 16(k'/16) + P[k'%16],  P = 0 1 8 9 2 3 10 11 4 5 12 13 6 7 14 15,
//   value = uint8(q[k][n] + 128).
#include "common.cuh"

namespace b200
{

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v)
{
    return v;
}
template <>
__device__ __forceinline__ float to_f32<__half>(__half v)
{
    return __half2float(v);
}

// amax[n] = max_k |w[k][n]| as the bit pattern of a non-negative float (ordered like uint32).
// grid (N/64, ceil(K/256)), block (64, 4): thread (x, y) walks rows y, y+4, ... of its 256-row slab.
template <typename T>
__global__ void __launch_bounds__(256) col_amax_kernel(const T* __restrict__ w, int K, int N, uint32_t* __restrict__ amax)
{
    const int n = blockIdx.x * 64 + threadIdx.x;
    const int k0 = blockIdx.y * 256;
    const int k1 = min(K, k0 + 256);
    float m = 0.f;
    for (int k = k0 + threadIdx.y; k < k1; k += 4)
        m = fmaxf(m, fabsf(to_f32(w[(size_t) k * N + n])));
    __shared__ float red[4][64];
    red[threadIdx.y][threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.y == 0)
    {
        m = fmaxf(fmaxf(red[0][threadIdx.x], red[1][threadIdx.x]), fmaxf(red[2][threadIdx.x], red[3][threadIdx.x]));
        atomicMax(&amax[n], __float_as_uint(m));
    }
}

template <typename S>
__global__ void write_scales_kernel(const uint32_t* __restrict__ amax, int N, S* __restrict__ scales)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N)
    {
        const float s = __uint_as_float(amax[n]) * (1.f / 128.f); // cutlass_preprocessors.cpp:641,669
        if constexpr (sizeof(S) == 2)
            scales[n] = __float2half_rn(s);
        else
            scales[n] = s;
    }
}

__device__ __constant__ int8_t c_perm16[16] = {0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15};

// One CTA = one 64(k) x 64(n) tile.  HAS_W: quantize from w (else the tile comes from raw_in).
template <typename T, bool HAS_W>
__global__ void __launch_bounds__(256) quantize_layout_kernel(const T* __restrict__ w, const int8_t* __restrict__ raw_in,
    const uint32_t* __restrict__ amax, int K, int N, int8_t* __restrict__ raw_out, uint8_t* __restrict__ proc)
{
    __shared__ int8_t tile[64][64 + 4]; // [k][n], +4 keeps rows word-aligned and de-conflicts column reads
    const int n0 = blockIdx.x * 64;
    const int t = blockIdx.y; // k-block of 64
    const int k0 = t * 64;
    const int tid = threadIdx.x;

    // phase 1: coalesced read of 64 rows x 64 columns, quantize, stage in smem (and emit raw, coalesced)
    {
        const int c = tid & 63;
        const int r0 = tid >> 6; // 0..3
        float scale = 0.f;
        if constexpr (HAS_W)
            scale = __uint_as_float(amax[n0 + c]) * (1.f / 128.f);
#pragma unroll 4
        for (int r = r0; r < 64; r += 4)
        {
            int8_t q;
            if constexpr (HAS_W)
            {
                const float v = to_f32(w[(size_t) (k0 + r) * N + n0 + c]);
                // fp32 IEEE division by the fp32 scale, C round() (half away from zero), clamp: :683-687
                const float s = roundf(__fdiv_rn(v, scale));
                q = static_cast<int8_t>(fmaxf(-128.f, fminf(127.f, s)));
            }
            else
            {
                q = raw_in[(size_t) (k0 + r) * N + n0 + c];
            }
            tile[r][c] = q;
            if (raw_out != nullptr)
                raw_out[(size_t) (k0 + r) * N + n0 + c] = q;
        }
    }
    __syncthreads();

    // phase 2: thread -> one 16-byte chunk of the processed layout.
    // 64 columns = 32 output rows j, each 128 bytes for this t: chunk c8 in 0..7 -> column parity c8/4, cc = c8%4.
    {
        const int jl = tid >> 3;  // 0..31
        const int c8 = tid & 7;   // 0..7
        const int nl = 2 * jl + (c8 >> 2);
        const int cc = c8 & 3;
        uint32_t out[4];
#pragma unroll
        for (int word = 0; word < 4; ++word)
        {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
            {
                const int sw = (b == 1) ? 2 : (b == 2 ? 1 : b); // {0,2,1,3}
                const int kp = 4 * word + sw;                    // k' % 16
                const int kl = 16 * cc + c_perm16[kp];
                const uint32_t u = static_cast<uint32_t>(static_cast<int>(tile[kl][nl]) + 128) & 0xffu;
                v |= u << (8 * b);
            }
            out[word] = v;
        }
        const size_t row = (size_t) (n0 / 2 + jl);
        uint4* dst = reinterpret_cast<uint4*>(proc + row * (size_t) (2 * K) + (size_t) t * 128 + (size_t) c8 * 16);
        *dst = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

static int check_shape(int K, int N)
{
    B200_REQUIRE(K > 0 && N > 0, B200_ERR_INVALID_ARG, "K and N must be positive (got K=%d N=%d)", K, N);
    // reference: rows % 16 (permute, :187), cols % 64 (interleave tile, :498); the 64-row interleave tile needs K % 64
    B200_REQUIRE(K % 64 == 0, B200_ERR_INVALID_ARG, "K=%d must be a multiple of 64 (reference layout tile)", K);
    B200_REQUIRE(N % 64 == 0, B200_ERR_INVALID_ARG, "N=%d must be a multiple of 64 (reference layout tile)", N);
    return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" int b200_symmetric_quantize_int8(const void* w, int w_dtype, int K, int N, int8_t* proc, int8_t* raw,
    void* scales, int scale_dtype, b200_stream_t stream_)
{
    B200_REQUIRE(w && proc && scales, B200_ERR_INVALID_ARG, "null pointer (w/proc/scales)");
    B200_REQUIRE(w_dtype == B200_DTYPE_F16 || w_dtype == B200_DTYPE_F32, B200_ERR_INVALID_ARG,
        "weight dtype must be fp16 or fp32");
    B200_REQUIRE(scale_dtype == B200_DTYPE_F16 || scale_dtype == B200_DTYPE_F32, B200_ERR_INVALID_ARG,
        "scale dtype must be fp16 or fp32");
    if (int rc = check_shape(K, N))
        return rc;
    B200_REQUIRE_DEVICE();
    cudaStream_t stream = as_stream(stream_);
    uint32_t* amax = nullptr;
    B200_CUDA(cudaMallocAsync(&amax, sizeof(uint32_t) * N, stream));
    B200_CUDA(cudaMemsetAsync(amax, 0, sizeof(uint32_t) * N, stream));
    dim3 g1(N / 64, (K + 255) / 256), b1(64, 4);
    dim3 g2(N / 64, K / 64);
    if (w_dtype == B200_DTYPE_F16)
        col_amax_kernel<__half><<<g1, b1, 0, stream>>>(static_cast<const __half*>(w), K, N, amax);
    else
        col_amax_kernel<float><<<g1, b1, 0, stream>>>(static_cast<const float*>(w), K, N, amax);
    B200_LAUNCH_CHECK();
    if (scale_dtype == B200_DTYPE_F16)
        write_scales_kernel<__half><<<(N + 255) / 256, 256, 0, stream>>>(amax, N, static_cast<__half*>(scales));
    else
        write_scales_kernel<float><<<(N + 255) / 256, 256, 0, stream>>>(amax, N, static_cast<float*>(scales));
    B200_LAUNCH_CHECK();
    if (w_dtype == B200_DTYPE_F16)
        quantize_layout_kernel<__half, true><<<g2, 256, 0, stream>>>(
            static_cast<const __half*>(w), nullptr, amax, K, N, raw, reinterpret_cast<uint8_t*>(proc));
    else
        quantize_layout_kernel<float, true><<<g2, 256, 0, stream>>>(
            static_cast<const float*>(w), nullptr, amax, K, N, raw, reinterpret_cast<uint8_t*>(proc));
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaFreeAsync(amax, stream));
    return B200_OK;
}

extern "C" int b200_preprocess_weights_int8(const int8_t* raw, int K, int N, int8_t* proc, b200_stream_t stream_)
{
    B200_REQUIRE(raw && proc, B200_ERR_INVALID_ARG, "null pointer (raw/proc)");
    if (int rc = check_shape(K, N))
        return rc;
    B200_REQUIRE_DEVICE();
    dim3 g2(N / 64, K / 64);
    quantize_layout_kernel<float, false><<<g2, 256, 0, as_stream(stream_)>>>(
        nullptr, raw, nullptr, K, N, nullptr, reinterpret_cast<uint8_t*>(proc));
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_symmetric_quantize_int8_host(
    const void* w, int w_dtype, int K, int N, int8_t* proc, int8_t* raw, void* scales, int scale_dtype)
{
    B200_REQUIRE(w && proc && scales, B200_ERR_INVALID_ARG, "null pointer (w/proc/scales)");
    B200_REQUIRE(w_dtype == B200_DTYPE_F16 || w_dtype == B200_DTYPE_F32, B200_ERR_INVALID_ARG,
        "weight dtype must be fp16 or fp32");
    B200_REQUIRE(scale_dtype == B200_DTYPE_F16 || scale_dtype == B200_DTYPE_F32, B200_ERR_INVALID_ARG,
        "scale dtype must be fp16 or fp32");
    if (int rc = check_shape(K, N))
        return rc;
    B200_REQUIRE_DEVICE();
    const size_t wb = (size_t) K * N * (w_dtype == B200_DTYPE_F16 ? 2 : 4);
    const size_t sb = (size_t) N * (scale_dtype == B200_DTYPE_F16 ? 2 : 4);
    const size_t qb = (size_t) K * N;
    char* d = nullptr;
    B200_CUDA(cudaMalloc(&d, wb + 2 * qb + sb + 256));
    char* dw = d;
    int8_t* dproc = reinterpret_cast<int8_t*>(d + ((wb + 127) / 128) * 128);
    int8_t* draw = dproc + qb;
    void* dsc = draw + qb;
    int rc = B200_OK;
    cudaError_t e = cudaMemcpy(dw, w, wb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        rc = b200_symmetric_quantize_int8(dw, w_dtype, K, N, dproc, raw ? draw : nullptr, dsc, scale_dtype, nullptr);
        if (rc == B200_OK)
        {
            e = cudaMemcpy(proc, dproc, qb, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess && raw)
                e = cudaMemcpy(raw, draw, qb, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess)
                e = cudaMemcpy(scales, dsc, sb, cudaMemcpyDeviceToHost);
        }
    }
    cudaFree(d);
    if (e != cudaSuccess)
    {
        set_error("cudaMemcpy failed: %s", cudaGetErrorString(e));
        return B200_ERR_CUDA;
    }
    return rc;
}

extern "C" int b200_preprocess_weights_int8_host(const int8_t* raw, int K, int N, int8_t* proc)
{
    B200_REQUIRE(raw && proc, B200_ERR_INVALID_ARG, "null pointer (raw/proc)");
    if (int rc = check_shape(K, N))
        return rc;
    B200_REQUIRE_DEVICE();
    const size_t qb = (size_t) K * N;
    int8_t* d = nullptr;
    B200_CUDA(cudaMalloc(&d, 2 * qb));
    cudaError_t e = cudaMemcpy(d, raw, qb, cudaMemcpyHostToDevice);
    int rc = B200_OK;
    if (e == cudaSuccess)
    {
        rc = b200_preprocess_weights_int8(d, K, N, d + qb, nullptr);
        if (rc == B200_OK)
            e = cudaMemcpy(proc, d + qb, qb, cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (e != cudaSuccess)
    {
        set_error("cudaMemcpy failed: %s", cudaGetErrorString(e));
        return B200_ERR_CUDA;
    }
    return rc;
}
