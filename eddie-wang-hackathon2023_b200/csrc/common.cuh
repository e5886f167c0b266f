// common.cuh -- shared host/device helpers for the sm_100a kernels (error plumbing, mbarrier / TMA / tcgen05 PTX).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "b200_whisper.h"

namespace b200
{

// ---- host-side error plumbing -----------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();
bool device_ok(); // true iff the current device is sm_100 (checked once)

#define B200_REQUIRE(cond, code, ...)                                                                                  \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            ::b200::set_error(__VA_ARGS__);                                                                            \
            return (code);                                                                                             \
        }                                                                                                              \
    } while (0)

#define B200_CUDA(call)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
        {                                                                                                              \
            ::b200::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);            \
            return B200_ERR_CUDA;                                                                                      \
        }                                                                                                              \
    } while (0)

#define B200_REQUIRE_DEVICE()                                                                                          \
    B200_REQUIRE(::b200::device_ok(), B200_ERR_CUDA, "no sm_100 CUDA device available (there is no CPU fallback)")

// launch bookkeeping
#define B200_LAUNCH_CHECK()                                                                                            \
    do                                                                                                                 \
    {                                                                                                                  \
        ::b200::count_launch();                                                                                        \
        B200_CUDA(cudaGetLastError());                                                                                 \
    } while (0)

static inline cudaStream_t as_stream(b200_stream_t s)
{
    return reinterpret_cast<cudaStream_t>(s);
}

bool static_kv_hint(); // attention kernels may read KV caches before griddepcontrol.wait (b200_set_static_kv_hint)
bool pdl_enabled(); // programmatic dependent launch for the hot-path kernels (b200_set_pdl / env B200_PDL)

#ifdef __CUDACC__
// Launches `kernel` with cudaLaunchKernelEx; when PDL is enabled the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization, so the grid may start while its predecessor on the stream is
// still draining.  Every kernel launched through this helper calls griddepcontrol.wait before it touches memory
// written by the predecessor, and only reads weights / cache data before that point.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(
    void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                                                            \
    do                                                                                                                 \
    {                                                                                                                  \
        ::b200::count_launch();                                                                                        \
        B200_CUDA(::b200::launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__));                                 \
    } while (0)

#ifdef __CUDACC__
// ---- device helpers ------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// generic-proxy writes -> visible to the async proxy (TMA / tcgen05 reading shared memory)
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity))
    {
    }
}

// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Same with an L2 eviction-priority hint (createpolicy result).
__device__ __forceinline__ void bulk_g2s_hint(
    void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// 2-D tiled TMA load (SASS: UTMALDG).  c0 = innermost coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// Programmatic dependent launch: wait for the producer grid's memory to be visible / allow dependents to start.
__device__ __forceinline__ void grid_dep_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ void grid_dep_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// biased uint8 x4 (one 32-bit word of the processed weight layout) -> two half2 of signed values.
// Word bytes [b0 b1 b2 b3] hold k-offsets {2i, 8+2i, 2i+1, 8+2i+1} (row permutation + byte swizzle of the
// reference layout, cutlass_preprocessors.cpp:154-157,392-398), so lo = (b0, b2) and hi = (b1, b3) are
// k-adjacent pairs.  0x6400 | b is the fp16 value 1024 + b; subtracting 1152 = 1024 + 128 removes the bias.
__device__ __forceinline__ void dequant_word(uint32_t w, __half2& lo, __half2& hi)
{
    const uint32_t l = __byte_perm(w, 0x64646464u, 0x5250);
    const uint32_t h = __byte_perm(w, 0x64646464u, 0x5351);
    const uint32_t magic = 0x64806480u;
    lo = __hsub2(*reinterpret_cast<const __half2*>(&l), *reinterpret_cast<const __half2*>(&magic));
    hi = __hsub2(*reinterpret_cast<const __half2*>(&h), *reinterpret_cast<const __half2*>(&magic));
}

__device__ __forceinline__ float gelu_erf(float x)
{
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float gelu_tanh(float x)
{
    return 0.5f * x * (1.0f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
}

// Epilogue shared by the matmul kernels: rounding after every step mirrors the reference's separate fp16
// elementwise layers (quantization/layer.py:311-312).
__device__ __forceinline__ __half epilogue_apply(
    float acc, float scale, const __half* bias, int activation, const __half* residual, int n, size_t idx)
{
    __half o = __float2half_rn(acc * scale);
    if (bias != nullptr)
        o = __float2half_rn(__half2float(o) + __half2float(bias[n]));
    if (activation == B200_ACT_GELU_ERF)
        o = __float2half_rn(gelu_erf(__half2float(o)));
    else if (activation == B200_ACT_GELU_TANH)
        o = __float2half_rn(gelu_tanh(__half2float(o)));
    if (residual != nullptr)
        o = __float2half_rn(__half2float(o) + __half2float(residual[idx]));
    return o;
}
#endif // __CUDACC__

} // namespace b200
