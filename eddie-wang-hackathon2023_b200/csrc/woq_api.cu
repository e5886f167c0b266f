// woq_api.cu -- C ABI entry points of the weight-only matmul and their dispatch.
// Mirrors WeightOnlyQuantMatmulPlugin::enqueue's m == 1 -> GEMV / else -> tensor-core GEMM split
// (T/cpp/tensorrt_llm/plugins/weightOnlyQuantMatmulPlugin/weightOnlyQuantMatmulPlugin.cpp:162-222), with the
// B200 crossover (measured): the tcgen05 kernel for every M (its weight stream and dequant run ahead of the
// dependency wait, which a register-resident SIMT GEMV cannot do), the SIMT GEMV only for a single row with deep K.
#include "common.cuh"

namespace b200
{
int woq_gemv_simt(const __half* A, int M, int K, const uint8_t* W, const __half* scales, int N, const __half* bias,
    int activation, const __half* residual, __half* C, cudaStream_t stream);
int woq_gemm_tc(const __half* A, int M, int K, const uint8_t* W, const __half* scales, int N, const __half* bias,
    int activation, const __half* residual, __half* C, void* workspace, size_t workspace_bytes, cudaStream_t stream,
    const __half* fold_gamma, const float* fold_c1s, const float* fold_c2, float ln_eps);
size_t woq_tc_workspace_bytes(int max_m, int N, int K);
int tc_init();
void tc_set_debug_buffer(long long* p);
void tc_set_debug_filter(int n, int fold);
void tc_set_timeline_buffer(long long* p, int max_launches);
bool woq_tc_can_fold_ln(int M, int N, int K);
void woq_tc_plan_query(int M, int N, int K, int* mt, int* m_tiles, int* n_tiles, int* splits, int* cluster);

// 0 auto, 1 simt, 2 tcgen05.  Per calling thread (a test / tuning switch, never set on the serving path): plugin
// instances enqueueing from other threads keep the automatic dispatch whatever a test thread forces.
static thread_local int g_policy = 0;

// One thread per weight column n: walks the column through the preprocessed layout exactly like the GEMM's dequant
// warps do (same 16-byte chunks, same PRMT/HSUB2 conversion, same HMUL2 by the gamma pair) and sums
//   c1[n] = sum_k fp16(Wint[k][n] * gamma[k])      c2[n] = sum_k beta[k] * Wint[k][n]
// both scaled by the column's dequant scale at the end.
__global__ void ln_fold_prepare_kernel(const uint8_t* __restrict__ W, const __half* __restrict__ scales,
    const __half* __restrict__ gamma, const __half* __restrict__ beta, int K, int N, float* __restrict__ c1s,
    float* __restrict__ c2)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N)
        return;
    const uint8_t* row = W + (size_t) (n >> 1) * 2 * K + (n & 1) * 64;
    const __half2* g2 = reinterpret_cast<const __half2*>(gamma);
    const __half2* b2 = reinterpret_cast<const __half2*>(beta);
    float a1 = 0.f, a2 = 0.f;
    for (int kb = 0; kb < K / 64; ++kb)
    {
        for (int ch = 0; ch < 4; ++ch)
        {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + (size_t) kb * 128 + ch * 16));
            const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int w = 0; w < 4; ++w)
            {
                __half2 lo, hi;
                dequant_word(words[w], lo, hi);
                const int plo = kb * 32 + ch * 8 + w, phi = plo + 4; // k pairs (2p, 2p+1)
                const float2 flo = __half22float2(__hmul2(lo, g2[plo]));
                const float2 fhi = __half22float2(__hmul2(hi, g2[phi]));
                a1 += (flo.x + flo.y) + (fhi.x + fhi.y);
                const float2 wl = __half22float2(lo), wh = __half22float2(hi);
                const float2 bl = __half22float2(b2[plo]), bh = __half22float2(b2[phi]);
                a2 += wl.x * bl.x + wl.y * bl.y + wh.x * bh.x + wh.y * bh.y;
            }
        }
    }
    const float sc = __half2float(scales[n]);
    c1s[n] = sc * a1;
    c2[n] = sc * a2;
}
} // namespace b200

using namespace b200;

extern "C" int b200_woq_set_kernel_policy(int policy)
{
    B200_REQUIRE(policy >= 0 && policy <= 2, B200_ERR_INVALID_ARG, "policy must be 0 (auto), 1 (simt) or 2 (tcgen05)");
    g_policy = policy;
    return B200_OK;
}

extern "C" size_t b200_woq_workspace_bytes(int max_m, int n, int k)
{
    if (max_m < 1 || n < 1 || k < 1)
        return 0;
    return woq_tc_workspace_bytes(max_m, n, k);
}

static int woq_dispatch(const void* A, int M, int K, const int8_t* Wproc, const void* scales, int N, const void* bias,
    int activation, const void* residual, void* C, void* workspace, size_t workspace_bytes, b200_stream_t stream,
    const void* ln_gamma, const void* ln_beta, float ln_eps, const float* fold_c1s = nullptr, const float* fold_c2 = nullptr)
{
    B200_REQUIRE(A && Wproc && scales && C, B200_ERR_INVALID_ARG, "null pointer (A/W/scales/C)");
    B200_REQUIRE(M >= 0, B200_ERR_INVALID_ARG, "M=%d must be >= 0", M);
    B200_REQUIRE(K > 0 && K % 64 == 0, B200_ERR_INVALID_ARG, "K=%d must be a positive multiple of 64", K);
    B200_REQUIRE(N > 0 && N % 64 == 0, B200_ERR_INVALID_ARG, "N=%d must be a positive multiple of 64", N);
    B200_REQUIRE(activation >= B200_ACT_NONE && activation <= B200_ACT_GELU_TANH, B200_ERR_INVALID_ARG,
        "unknown activation %d", activation);
    if (M == 0)
        return B200_OK; // empty batch: nothing to do (the reference would launch an empty grid)
    B200_REQUIRE_DEVICE();
    // measured crossover (tools/gemv_crossover.py, graph replays over 32 weight sets): the tcgen05 kernel wins from
    // M = 1 on every decoder shape but the deep-K single row (K = 5120, M = 1: 4.3 vs 5.1 us)
    const bool simt = (g_policy == 1) || (g_policy == 0 && M == 1 && K >= 4096);
    const bool fold = ln_gamma != nullptr && fold_c1s != nullptr && fold_c2 != nullptr && !simt && woq_tc_can_fold_ln(M, N, K);
    if (!fold)
    {
        const __half* a = static_cast<const __half*>(A);
        if (ln_gamma != nullptr)
        {
            // the SIMT GEMV keeps its activations in registers: normalise into the workspace first
            const size_t need = (size_t) M * K * sizeof(__half);
            B200_REQUIRE(workspace != nullptr && workspace_bytes >= need, B200_ERR_WORKSPACE,
                "woq gemm (separate LayerNorm): workspace of %zu bytes needed", need);
            if (int rc = b200_layernorm_fp16(A, ln_gamma, ln_beta, workspace, M, K, ln_eps, stream))
                return rc;
            a = static_cast<const __half*>(workspace);
            // the rest of the workspace stays available for split-K slabs
            const size_t used = (need + 255) & ~size_t(255);
            workspace = static_cast<char*>(workspace) + used;
            workspace_bytes = workspace_bytes > used ? workspace_bytes - used : 0;
        }
        if (simt)
            return woq_gemv_simt(a, M, K, reinterpret_cast<const uint8_t*>(Wproc), static_cast<const __half*>(scales), N,
                static_cast<const __half*>(bias), activation, static_cast<const __half*>(residual),
                static_cast<__half*>(C), as_stream(stream));
        return woq_gemm_tc(a, M, K, reinterpret_cast<const uint8_t*>(Wproc), static_cast<const __half*>(scales), N,
            static_cast<const __half*>(bias), activation, static_cast<const __half*>(residual), static_cast<__half*>(C),
            workspace, workspace_bytes, as_stream(stream), nullptr, nullptr, nullptr, 0.f);
    }
    return woq_gemm_tc(static_cast<const __half*>(A), M, K, reinterpret_cast<const uint8_t*>(Wproc),
        static_cast<const __half*>(scales), N, static_cast<const __half*>(bias), activation,
        static_cast<const __half*>(residual), static_cast<__half*>(C), workspace, workspace_bytes, as_stream(stream),
        static_cast<const __half*>(ln_gamma), fold_c1s, fold_c2, ln_eps);
}

extern "C" int b200_woq_ln_fold_prepare(const int8_t* Wproc, const void* scales, const void* ln_gamma, const void* ln_beta,
    int K, int N, float* c1s, float* c2, b200_stream_t stream)
{
    B200_REQUIRE(Wproc && scales && ln_gamma && ln_beta && c1s && c2, B200_ERR_INVALID_ARG, "null pointer");
    B200_REQUIRE(K > 0 && K % 64 == 0 && N > 0 && N % 64 == 0, B200_ERR_INVALID_ARG,
        "K=%d and N=%d must be positive multiples of 64", K, N);
    B200_REQUIRE_DEVICE();
    // plain (fully serialised) launch: a one-off preparation step, no programmatic overlap with its producers
    ln_fold_prepare_kernel<<<(N + 127) / 128, 128, 0, as_stream(stream)>>>(reinterpret_cast<const uint8_t*>(Wproc),
        static_cast<const __half*>(scales), static_cast<const __half*>(ln_gamma), static_cast<const __half*>(ln_beta), K,
        N, c1s, c2);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_woq_int8_gemm_ln_folded(const void* X, const void* ln_gamma, const void* ln_beta, const float* c1s,
    const float* c2, float ln_eps, int M, int K, const int8_t* Wproc, const void* scales, int N, const void* bias,
    int activation, const void* residual, void* C, void* workspace, size_t workspace_bytes, b200_stream_t stream)
{
    B200_REQUIRE(ln_gamma && ln_beta && c1s && c2, B200_ERR_INVALID_ARG, "null pointer (ln_gamma/ln_beta/c1s/c2)");
    B200_REQUIRE(workspace == nullptr || workspace != C, B200_ERR_INVALID_ARG, "workspace must not alias C");
    return woq_dispatch(X, M, K, Wproc, scales, N, bias, activation, residual, C, workspace, workspace_bytes, stream,
        ln_gamma, ln_beta, ln_eps, c1s, c2);
}

extern "C" int b200_woq_int8_gemm_fused(const void* A, int M, int K, const int8_t* Wproc, const void* scales, int N,
    const void* bias, int activation, const void* residual, void* C, void* workspace, size_t workspace_bytes,
    b200_stream_t stream)
{
    return woq_dispatch(A, M, K, Wproc, scales, N, bias, activation, residual, C, workspace, workspace_bytes, stream,
        nullptr, nullptr, 0.f);
}

extern "C" int b200_woq_int8_gemm_ln_fused(const void* X, const void* ln_gamma, const void* ln_beta, float ln_eps, int M,
    int K, const int8_t* Wproc, const void* scales, int N, const void* bias, int activation, const void* residual, void* C,
    void* workspace, size_t workspace_bytes, b200_stream_t stream)
{
    B200_REQUIRE(ln_gamma && ln_beta, B200_ERR_INVALID_ARG, "null pointer (ln_gamma/ln_beta)");
    B200_REQUIRE(workspace == nullptr || workspace != C, B200_ERR_INVALID_ARG, "workspace must not alias C");
    return woq_dispatch(X, M, K, Wproc, scales, N, bias, activation, residual, C, workspace, workspace_bytes, stream,
        ln_gamma, ln_beta, ln_eps);
}

extern "C" int b200_woq_int8_gemm(const void* A, int M, int K, const int8_t* Wproc, const void* scales, int N, void* C,
    void* workspace, size_t workspace_bytes, b200_stream_t stream)
{
    return b200_woq_int8_gemm_fused(
        A, M, K, Wproc, scales, N, nullptr, B200_ACT_NONE, nullptr, C, workspace, workspace_bytes, stream);
}

extern "C" int b200_debug_woq_plan(int M, int N, int K, int* plan5)
{
    B200_REQUIRE(plan5 != nullptr && M >= 1 && N >= 64 && K >= 64 && K % 64 == 0 && N % 64 == 0, B200_ERR_INVALID_ARG,
        "bad arguments");
    woq_tc_plan_query(M, N, K, &plan5[0], &plan5[1], &plan5[2], &plan5[3], &plan5[4]);
    return B200_OK;
}

extern "C" int b200_init(void)
{
    B200_REQUIRE_DEVICE();
    return tc_init();
}

// Debug aid: device buffer of >= 16 int64 that receives clock64() stamps of CTA (0,0,0) of every following tcgen05
// GEMM launch (phase boundaries, see TC_STAMP in woq_gemm_tc.cu); NULL switches it off.
// Debug aid: every following tcgen05 GEMM launch (up to max_launches) records, in launch order, 4 int64 global-timer
// values: [0] earliest CTA entry, [1] earliest return of the dependency wait, [2] latest CTA exit.  The caller
// initialises [0] and [1] to INT64_MAX and [2] to 0.  NULL switches it off.
extern "C" int b200_debug_tc_timeline(void* device_buffer, int max_launches)
{
    tc_set_timeline_buffer(static_cast<long long*>(device_buffer), max_launches);
    return B200_OK;
}

extern "C" int b200_debug_tc_timing_filter(int n, int folded_ln)
{
    tc_set_debug_filter(n, folded_ln);
    return B200_OK;
}

extern "C" int b200_debug_tc_timing(void* device_buffer)
{
    tc_set_debug_buffer(static_cast<long long*>(device_buffer));
    return B200_OK;
}
