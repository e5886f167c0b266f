// glue.cu -- decoder-step glue: LayerNorm, token+position embedding, fp16 logits + argmax (SIMT path).
//
// These are TensorRT-native layers in the reference graph (T/tensorrt_llm/models/whisper/model.py:74-118,257-292)
// and torch ops in the oracle (T/examples/whisper/torch_model.py:25-27,205-218).  They are the "next" row 1 of
// SURVEY.md section 8f and are needed to measure whole decoder steps.
#include <float.h>

#include "common.cuh"

namespace b200
{
int* tc_counter_slot(int needed, cudaStream_t stream);

// y = (x - mean) * rsqrt(var + eps) * gamma + beta, statistics in fp32 (torch_model.py:25-27 casts to float).
// one CTA (128 threads) per row; cols % 8 == 0; row cached in registers (cols <= 8192).
template <int VEC_PER_THREAD>
__global__ void __launch_bounds__(128) layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
    const __half* __restrict__ beta, __half* __restrict__ y, int cols, float eps)
{
    grid_dep_wait();
    grid_dep_launch_dependents();
    const int row = blockIdx.x;
    const __half* xr = x + (size_t) row * cols;
    __half* yr = y + (size_t) row * cols;
    const int nvec = cols / 8;
    float v[VEC_PER_THREAD][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i)
    {
        const int idx = threadIdx.x + i * 128;
        if (idx < nvec)
        {
            const uint4 u = *reinterpret_cast<const uint4*>(xr + idx * 8);
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const float2 f = __half22float2(h[j]);
                v[i][2 * j] = f.x;
                v[i][2 * j + 1] = f.y;
                sum += f.x + f.y;
            }
        }
    }
    __shared__ float red[4];
    __shared__ float bc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0)
        red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0)
        bc = (red[0] + red[1] + red[2] + red[3]) / (float) cols;
    __syncthreads();
    const float mean = bc;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i)
    {
        const int idx = threadIdx.x + i * 128;
        if (idx < nvec)
        {
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const float d = v[i][j] - mean;
                sq += d * d;
            }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    __syncthreads();
    if (lane == 0)
        red[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0)
        bc = rsqrtf((red[0] + red[1] + red[2] + red[3]) / (float) cols + eps);
    __syncthreads();
    const float rstd = bc;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i)
    {
        const int idx = threadIdx.x + i * 128;
        if (idx < nvec)
        {
            const uint4 g = *reinterpret_cast<const uint4*>(gamma + idx * 8);
            const uint4 bt = *reinterpret_cast<const uint4*>(beta + idx * 8);
            const __half2* gh = reinterpret_cast<const __half2*>(&g);
            const __half2* bh = reinterpret_cast<const __half2*>(&bt);
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const float2 gf = __half22float2(gh[j]);
                const float2 bf = __half22float2(bh[j]);
                oh[j] = __floats2half2_rn((v[i][2 * j] - mean) * rstd * gf.x + bf.x, (v[i][2 * j + 1] - mean) * rstd * gf.y + bf.y);
            }
            *reinterpret_cast<uint4*>(yr + idx * 8) = o;
        }
    }
}

// The same LayerNorm for MANY rows (the encoder: 24000 rows of 1280): one WARP per row, 8 rows per CTA, the row cached in
// registers (VPL 16-byte vectors per lane), both reductions by shuffles -- no shared memory, no block barrier.  The
// CTA-per-row kernel above spends its time in four barriers per 2.5 KB row (105 us for 123 MB of traffic).
template <int VPL>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
    const __half* __restrict__ beta, __half* __restrict__ y, int rows, int cols, float eps)
{
    grid_dep_wait();
    grid_dep_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows)
        return;
    const __half* xr = x + (size_t) row * cols;
    __half* yr = y + (size_t) row * cols;
    const int nvec = cols / 8;
    float v[VPL][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
    {
        const int idx = lane + i * 32;
        if (idx < nvec)
        {
            const uint4 u = *reinterpret_cast<const uint4*>(xr + idx * 8);
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const float2 f = __half22float2(h[j]);
                v[i][2 * j] = f.x;
                v[i][2 * j + 1] = f.y;
                sum += f.x + f.y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float) cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
        if (lane + i * 32 < nvec)
        {
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const float d = v[i][j] - mean;
                sq += d * d;
            }
        }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float) cols + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i)
    {
        const int idx = lane + i * 32;
        if (idx < nvec)
        {
            const uint4 g = __ldg(reinterpret_cast<const uint4*>(gamma + idx * 8));
            const uint4 bt = __ldg(reinterpret_cast<const uint4*>(beta + idx * 8));
            const __half2* gh = reinterpret_cast<const __half2*>(&g);
            const __half2* bh = reinterpret_cast<const __half2*>(&bt);
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const float2 gf = __half22float2(gh[j]);
                const float2 bf = __half22float2(bh[j]);
                oh[j] = __floats2half2_rn((v[i][2 * j] - mean) * rstd * gf.x + bf.x, (v[i][2 * j + 1] - mean) * rstd * gf.y + bf.y);
            }
            *reinterpret_cast<uint4*>(yr + idx * 8) = o;
        }
    }
}

// out[r] = tok_emb[tokens[r]] + pos_emb[positions[r]]  (fp16 add, like torch_model.py:205-209 in fp16)
__global__ void embed_kernel(const int* __restrict__ tokens, const int* __restrict__ positions,
    const __half* __restrict__ tok_emb, const __half* __restrict__ pos_emb, __half* __restrict__ out, int cols, int vocab,
    int n_ctx)
{
    grid_dep_wait();
    grid_dep_launch_dependents();
    const int r = blockIdx.x;
    int tok = tokens[r];
    int pos = positions[r];
    tok = min(max(tok, 0), vocab - 1);
    pos = min(max(pos, 0), n_ctx - 1);
    const __half2* te = reinterpret_cast<const __half2*>(tok_emb + (size_t) tok * cols);
    const __half2* pe = reinterpret_cast<const __half2*>(pos_emb + (size_t) pos * cols);
    __half2* o = reinterpret_cast<__half2*>(out + (size_t) r * cols);
    for (int i = threadIdx.x; i < cols / 2; i += blockDim.x)
        o[i] = __hadd2(te[i], pe[i]);
}

// SIMT logits: one warp per vocabulary row, all `rows` activations (<= 16 per pass) staged in shared memory.
// logits[r][v] = sum_k x[r][k] * emb[v][k], fp32 accumulation.
constexpr int kLogitsRowsPerPass = 16;

__global__ void __launch_bounds__(256) logits_simt_kernel(const __half* __restrict__ x, const __half* __restrict__ emb,
    float* __restrict__ logits, int rows, int cols, int vocab, int row0)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __half* sx = reinterpret_cast<__half*>(s_raw); // [nr][cols]
    grid_dep_wait();
    grid_dep_launch_dependents();
    const int nr = min(kLogitsRowsPerPass, rows - row0);
    for (int i = threadIdx.x; i < nr * cols / 8; i += blockDim.x)
        reinterpret_cast<uint4*>(sx)[i] = reinterpret_cast<const uint4*>(x + (size_t) row0 * cols)[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = blockIdx.x * 8 + warp;
    if (v >= vocab)
        return;
    float acc[kLogitsRowsPerPass];
#pragma unroll
    for (int r = 0; r < kLogitsRowsPerPass; ++r)
        acc[r] = 0.f;
    const __half* er = emb + (size_t) v * cols;
    for (int k = lane * 8; k < cols; k += 256)
    {
        const uint4 e4 = __ldg(reinterpret_cast<const uint4*>(er + k));
        const __half2* eh = reinterpret_cast<const __half2*>(&e4);
        float ef[8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const float2 f = __half22float2(eh[j]);
            ef[2 * j] = f.x;
            ef[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int r = 0; r < kLogitsRowsPerPass; ++r)
        {
            if (r < nr)
            {
                const uint4 x4 = *reinterpret_cast<const uint4*>(sx + (size_t) r * cols + k);
                const __half2* xh = reinterpret_cast<const __half2*>(&x4);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    const float2 f = __half22float2(xh[j]);
                    acc[r] = fmaf(f.x, ef[2 * j], acc[r]);
                    acc[r] = fmaf(f.y, ef[2 * j + 1], acc[r]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kLogitsRowsPerPass; ++r)
    {
        if (r < nr)
        {
            float a = acc[r];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1)
                a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0)
                logits[(size_t) (row0 + r) * vocab + v] = a;
        }
    }
}

// [B, C, T] -> [B, T, C] + pos[T, C]: the encoder's `x.permute(0, 2, 1) + positional_embedding`
// (T/tensorrt_llm/models/whisper/model.py:158-162, oracle torch_model.py:159-162).  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) transpose_add_pos_kernel(const __half* __restrict__ x, const __half* __restrict__ pos,
    __half* __restrict__ y, int C, int T)
{
    __shared__ __half tile[32][33];
    grid_dep_wait();
    grid_dep_launch_dependents();
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const __half* xb = x + (size_t) b * C * T;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int c = c0 + ty + 8 * r, t = t0 + tx;
        tile[ty + 8 * r][tx] = (c < C && t < T) ? xb[(size_t) c * T + t] : __float2half(0.f);
    }
    __syncthreads();
    __half* yb = y + (size_t) b * T * C;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int t = t0 + ty + 8 * r, c = c0 + tx;
        if (t < T && c < C)
            yb[(size_t) t * C + c]
                = __float2half_rn(__half2float(tile[tx][ty + 8 * r]) + __half2float(pos[(size_t) t * C + c]));
    }
}

// argmax over fp32 logits, first index on ties (torch.argmax).  grid (parts, rows), 256 threads: every CTA scans a
// slice of the row with 8 independent loads in flight per thread, publishes (value, index) as one order-preserving
// 64-bit key with atomicMax, and the last CTA of a row to arrive writes the token and resets the scratch words.
__device__ __forceinline__ unsigned long long argmax_key(float x, int idx)
{
    const uint32_t b = __float_as_uint(x);
    const uint32_t u = (b & 0x80000000u) ? ~b : (b | 0x80000000u); // monotone in x
    return ((unsigned long long) u << 32) | (unsigned long long) (0xffffffffu - (uint32_t) idx); // ties: lowest index wins
}

__global__ void __launch_bounds__(256) argmax_kernel(const float* __restrict__ logits, int* __restrict__ next_token, int vocab,
    unsigned long long* __restrict__ packed, int* __restrict__ counters)
{
    grid_dep_wait();
    grid_dep_launch_dependents();
    const int r = blockIdx.y, parts = gridDim.x;
    const float* lr = logits + (size_t) r * vocab;
    const int per = (vocab + parts - 1) / parts;
    const int v0 = blockIdx.x * per, v1 = min(vocab, v0 + per);
    float best = -FLT_MAX;
    int bi = 0x7fffffff;
    for (int base = v0 + threadIdx.x; base < v1; base += 8 * 256)
    {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const int v = base + j * 256;
            x[j] = v < v1 ? __ldcs(lr + v) : -FLT_MAX;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const int v = base + j * 256;
            if (x[j] > best || (x[j] == best && v < bi))
            {
                best = x[j];
                bi = v;
            }
        }
    }
    unsigned long long key = argmax_key(best, bi);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
    {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
        key = ok > key ? ok : key;
    }
    __shared__ unsigned long long sk[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        sk[warp] = key;
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int w = 1; w < 8; ++w)
            key = sk[w] > key ? sk[w] : key;
        atomicMax(&packed[r], key);
        __threadfence();
        if (atomicAdd(&counters[r], 1) == parts - 1)
        {
            __threadfence();
            const unsigned long long win = atomicExch(&packed[r], 0ull); // read + reset for the next launch
            next_token[r] = (int) (0xffffffffu - (uint32_t) (win & 0xffffffffull));
            counters[r] = 0;
        }
    }
}

// Fire-and-forget L2 prefetch of a byte range (cp.async.bulk.prefetch.L2): used on a side stream to pull the next
// layer's cross-KV cache into the 126 MB L2 while the latency-bound small kernels of the current layer leave HBM idle.
__global__ void l2_prefetch_kernel(const uint8_t* __restrict__ p, size_t bytes, uint32_t chunk)
{
    const size_t n = (bytes + chunk - 1) / chunk;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    {
        const size_t off = i * chunk;
        const size_t left = bytes - off;
        const uint32_t sz = (uint32_t) (left < chunk ? left : chunk) & ~15u;
        if (sz != 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(sz) : "memory");
    }
}

int logits_tc(const __half* x, const __half* emb, float* logits, int rows, int cols, int vocab, cudaStream_t stream);
extern thread_local int g_logits_policy;
thread_local int g_logits_policy = 0; // 0 auto (tcgen05), 1 simt; per calling thread (test switch)

} // namespace b200

using namespace b200;

extern "C" int b200_layernorm_fp16(const void* x, const void* gamma, const void* beta, void* y, int rows, int cols,
    float eps, b200_stream_t stream)
{
    B200_REQUIRE(x && gamma && beta && y, B200_ERR_INVALID_ARG, "null pointer");
    B200_REQUIRE(cols > 0 && cols % 8 == 0 && cols <= 8192, B200_ERR_UNSUPPORTED, "cols=%d must be a multiple of 8 and <= 8192", cols);
    if (rows <= 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    const int vpt = (cols / 8 + 127) / 128;
    cudaStream_t st = as_stream(stream);
    const __half *xh = static_cast<const __half*>(x), *g = static_cast<const __half*>(gamma), *b = static_cast<const __half*>(beta);
    __half* yh = static_cast<__half*>(y);
    // many rows (the encoder): a warp per row; the decoder's handful of rows keeps the CTA-per-row kernel
    const int vpl = (cols / 8 + 31) / 32;
    if (rows >= 1024 && vpl <= 8)
    {
        const dim3 grid((rows + 7) / 8);
        if (vpl <= 2)
            B200_LAUNCH(layernorm_rows_kernel<2>, grid, dim3(256), 0, st, xh, g, b, yh, rows, cols, eps);
        else if (vpl <= 5)
            B200_LAUNCH(layernorm_rows_kernel<5>, grid, dim3(256), 0, st, xh, g, b, yh, rows, cols, eps);
        else
            B200_LAUNCH(layernorm_rows_kernel<8>, grid, dim3(256), 0, st, xh, g, b, yh, rows, cols, eps);
        return B200_OK;
    }
    if (vpt <= 1)
        B200_LAUNCH(layernorm_kernel<1>, dim3(rows), dim3(128), 0, st, xh, g, b, yh, cols, eps);
    else if (vpt <= 2)
        B200_LAUNCH(layernorm_kernel<2>, dim3(rows), dim3(128), 0, st, xh, g, b, yh, cols, eps);
    else if (vpt <= 4)
        B200_LAUNCH(layernorm_kernel<4>, dim3(rows), dim3(128), 0, st, xh, g, b, yh, cols, eps);
    else
        B200_LAUNCH(layernorm_kernel<8>, dim3(rows), dim3(128), 0, st, xh, g, b, yh, cols, eps);
    return B200_OK;
}

extern "C" int b200_embed_tokens_fp16(const int32_t* tokens, const int32_t* positions, const void* tok_emb,
    const void* pos_emb, void* out, int rows, int cols, int vocab, int n_ctx, b200_stream_t stream)
{
    B200_REQUIRE(tokens && positions && tok_emb && pos_emb && out, B200_ERR_INVALID_ARG, "null pointer");
    B200_REQUIRE(cols > 0 && cols % 2 == 0, B200_ERR_INVALID_ARG, "cols=%d must be even", cols);
    if (rows <= 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    B200_LAUNCH(embed_kernel, dim3(rows), dim3(128), 0, as_stream(stream), tokens, positions,
        static_cast<const __half*>(tok_emb), static_cast<const __half*>(pos_emb), static_cast<__half*>(out), cols, vocab,
        n_ctx);
    return B200_OK;
}

extern "C" int b200_l2_prefetch(const void* ptr, size_t bytes, b200_stream_t stream)
{
    B200_REQUIRE(ptr != nullptr || bytes == 0, B200_ERR_INVALID_ARG, "null pointer");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, B200_ERR_INVALID_ARG, "pointer must be 16-byte aligned");
    if (bytes < 16)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    const uint32_t chunk = 8192;
    const size_t n = (bytes + chunk - 1) / chunk;
    int blocks = (int) ((n + 63) / 64);
    if (blocks > num_sms())
        blocks = num_sms();
    l2_prefetch_kernel<<<blocks, 64, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(ptr), bytes, chunk);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" size_t b200_logits_workspace_bytes(int rows, int vocab)
{
    if (rows <= 0 || vocab <= 0)
        return 0;
    return (size_t) rows * vocab * sizeof(float);
}

extern "C" int b200_logits_set_kernel_policy(int policy)
{
    B200_REQUIRE(policy == 0 || policy == 1, B200_ERR_INVALID_ARG, "policy must be 0 (auto) or 1 (simt)");
    g_logits_policy = policy;
    return B200_OK;
}

extern "C" int b200_logits_argmax_fp16(const void* x, const void* emb, void* logits_fp32, int32_t* next_token, int rows,
    int cols, int vocab, void* workspace, size_t workspace_bytes, b200_stream_t stream)
{
    B200_REQUIRE(x && emb, B200_ERR_INVALID_ARG, "null pointer (x/emb)");
    B200_REQUIRE(logits_fp32 || next_token, B200_ERR_INVALID_ARG, "nothing to compute: logits and next_token are both NULL");
    B200_REQUIRE(cols > 0 && cols % 64 == 0, B200_ERR_UNSUPPORTED, "cols=%d must be a multiple of 64", cols);
    if (rows <= 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    float* lg = static_cast<float*>(logits_fp32);
    if (lg == nullptr)
    {
        B200_REQUIRE(workspace && workspace_bytes >= (size_t) rows * vocab * sizeof(float), B200_ERR_WORKSPACE,
            "logits: workspace of %zu bytes needed", (size_t) rows * vocab * sizeof(float));
        lg = static_cast<float*>(workspace);
    }
    cudaStream_t st = as_stream(stream);
    if (g_logits_policy == 0)
    {
        if (int rc = logits_tc(static_cast<const __half*>(x), static_cast<const __half*>(emb), lg, rows, cols, vocab, st))
            return rc;
    }
    else
    {
        const size_t smem = (size_t) kLogitsRowsPerPass * cols * sizeof(__half);
        B200_REQUIRE(smem <= 200 * 1024, B200_ERR_UNSUPPORTED, "cols=%d too large for the SIMT logits path", cols);
        if (smem > 48 * 1024)
            B200_CUDA(cudaFuncSetAttribute(logits_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        for (int r0 = 0; r0 < rows; r0 += kLogitsRowsPerPass)
        {
            B200_LAUNCH(logits_simt_kernel, dim3((vocab + 7) / 8), dim3(256), smem, st, static_cast<const __half*>(x),
                static_cast<const __half*>(emb), lg, rows, cols, vocab, r0);
        }
    }
    if (next_token != nullptr)
    {
        // scratch: one 64-bit key and one arrival counter per row (library owned, self-resetting)
        B200_REQUIRE(rows <= 4096, B200_ERR_UNSUPPORTED, "argmax: %d rows exceed the scratch slot", rows);
        int* slot = tc_counter_slot(3 * rows + 2, st);
        B200_REQUIRE(slot != nullptr, B200_ERR_CUDA, "argmax: no scratch slot");
        unsigned long long* packed = reinterpret_cast<unsigned long long*>(slot);
        int* counters = slot + 2 * rows;
        const int parts = rows >= 64 ? 2 : 8;
        B200_LAUNCH(argmax_kernel, dim3(parts, rows), dim3(256), 0, st, static_cast<const float*>(lg), next_token, vocab, packed,
            counters);
    }
    return B200_OK;
}

extern "C" int b200_transpose_add_pos_fp16(const void* x, const void* pos, void* y, int batch_size, int channels, int t,
    b200_stream_t stream)
{
    B200_REQUIRE(x && pos && y, B200_ERR_INVALID_ARG, "null pointer (x/pos/y)");
    B200_REQUIRE(channels > 0 && t > 0, B200_ERR_INVALID_ARG, "bad sizes");
    if (batch_size <= 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    B200_LAUNCH(transpose_add_pos_kernel, dim3((t + 31) / 32, (channels + 31) / 32, batch_size), dim3(256), 0,
        as_stream(stream), static_cast<const __half*>(x), static_cast<const __half*>(pos), static_cast<__half*>(y), channels, t);
    return B200_OK;
}
