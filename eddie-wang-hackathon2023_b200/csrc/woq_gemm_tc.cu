// woq_gemm_tc.cu -- fp16 activations x per-channel int8 weights on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Replaces CutlassFpAIntBGemmRunner<half,uint8_t>::gemm
//   T/cpp/tensorrt_llm/kernels/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:358-435
//   (kernel T/cpp/tensorrt_llm/cutlass_extensions/include/cutlass_extensions/gemm/kernel/fpA_intB_gemm.h:57-490),
// which does not build for sm_100 (static_assert at fpA_intB_gemm.h:483-485), on the reference's exact
// preprocessed weight layout ([N/2][2K] bytes, see quantize.cu).
//
// B200 design ("swap-AB", weights are the UMMA A operand and live in TMEM):
//   D[n, m] = sum_k W16[n, k] * X[m, k]      UMMA M = 128 weight columns n, UMMA N = MT activation rows m.
// 320 threads per CTA:
//  * warp 8 (one elected lane): TMA producer.  Per 64-wide k-block it loads the int8 weight tile (64 row pairs x 128 B,
//    one 2-D box of the [N/2][2K] byte matrix, 128B-swizzled so the dequant warps' 128-bit reads are conflict-free) into
//    an SS-deep ring BEFORE griddepcontrol.wait (weights never depend on the previous kernel), and after the wait the
//    fp16 activation tiles (MT rows x 64 k, K-major, 128B swizzle = the canonical UMMA B layout) -- for decode launches
//    as ONE 3-D box covering the CTA's whole k range.  Weights and activations complete on separate mbarriers.
//  * warps 0-7: dequant.  Thread = weight column (TMEM lane) x k-half: it reads its 32 bytes of the k-block, converts
//    them with PRMT/HSUB2 (the 0x6400|b trick: exact fp16 integers; the fp16 column scale is applied to the fp32
//    accumulator in the epilogue) and writes 16 packed-half2 registers straight into TMEM with tcgen05.st.32x32b.x16.
//    The reference layout's row permutation + byte swizzle make every converted register a k-adjacent pair, i.e.
//    exactly one 32-bit TMEM column of the K-major A operand -- no shuffles, no second shared-memory round trip.
//    With a folded LayerNorm the pairs are multiplied by gamma here and the row statistics are reduced from the
//    activation tiles in shared memory (see TcParams).
//  * warp 9: issues tcgen05.mma.kind::f16 (elect.sync lane) with A from TMEM and B from shared memory, fp32
//    accumulator in TMEM, tcgen05.commit to release stages; allocates / frees TMEM.
//  * warps 0-7 again: epilogue.  Split-K CTAs of a tile form a thread-block cluster: partial accumulators are pushed
//    with st.async into the inbox of the CTA that owns the column slice (complete_tx on its mbarrier), summed in rank
//    order (deterministic) and finished with bias / GELU / residual.  Unsplit tiles (large M) finish their own columns;
//    B200_SPLITK=global selects fp32 slabs in the workspace + last-CTA reduction.
// Decode tiles (MT <= 32) use 96 registers, 115 KB of shared memory and 256 TMEM columns so that two CTAs -- this GEMM's
// and the next one's, launched programmatically -- share an SM.
// HBM traffic: every weight byte is read once per m-tile; activations are re-read per n-tile from L2.
#include <cuda.h>

#include <mutex>
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "tcgen05.cuh"

namespace b200
{

struct TcParams
{
    const __half* scales;
    const __half* bias;
    const __half* residual;
    __half* C;
    float* slabs;  // split-K partial tiles (workspace), global-memory reduction mode
    int* counters; // per-tile arrival counters (library owned, self-resetting)
    // folded LayerNorm (fold_gamma != nullptr): the activation operand is the RAW residual stream x; gamma is
    // multiplied into the dequantized weights, the row statistics are computed next to the main loop and applied to
    // the accumulator:  y = rstd * (acc - mean * c1s[n]) + c2[n]   (see b200_woq_ln_fold_prepare)
    const __half* fold_x; // [M, K] raw rows (same memory the activation tensor map points at)
    const __half* fold_gamma;
    const float* fold_c1s;
    const float* fold_c2;
    float ln_eps;
    int M, N, K;
    int ldc;
    int activation;
    int kb_total; // K / 64
    int splits;
    int cluster;  // 1: the `splits` CTAs of a tile form a thread-block cluster and reduce through DSMEM
    int x3_depth; // > 0: tmX is a 3-D map (64 k, rows, k-blocks) whose box holds this many k-blocks: ONE activation load
    int mma_burst; // 1: decode launches issue all their MMAs after one wait (env B200_TC_BURST=0 restores the per-block loop)
    long long* gt;  // optional: 4 global-timer values of this launch (min entry, min dependency return, max store, -)
    long long* dbg; // optional: clock64() stamps of CTA (0,0,0) at the phase boundaries (b200_debug_tc_timing)
};

__device__ __forceinline__ bool tc_burst_enabled(const TcParams& p)
{
    return p.mma_burst != 0;
}

__device__ __forceinline__ long long global_timer_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#if defined(B200_TC_DEBUG)
#define TC_STAMP(slot)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)                                 \
            p.dbg[slot] = clock64();                                                                                   \
    } while (0)
#define TC_GT(op, slot)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        if (p.gt != nullptr)                                                                                           \
            op(reinterpret_cast<unsigned long long*>(p.gt + (slot)), (unsigned long long) global_timer_ns());          \
    } while (0)
#else // release build: no stamps in the kernel (python __graft_entry__.py build with B200_TC_DEBUG=1 enables them)
#define TC_STAMP(slot)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
    } while (0)
#define TC_GT(op, slot)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
    } while (0)
#endif


// bias / activation / residual with the reference's per-layer fp16 rounding (see epilogue_apply in common.cuh).
// The activation is a template parameter: with a run-time switch the compiler predicates the erf / tanh polynomials
// into every output's instruction stream (they cost issue slots even when masked off; the 256-row epilogue spent 70 %
// of the kernel there).
template <int ACT>
__device__ __forceinline__ __half finish_output(float acc, bool has_bias, float biasv, bool has_res, float res)
{
    __half o = __float2half_rn(acc);
    if (has_bias)
        o = __float2half_rn(__half2float(o) + biasv);
    if constexpr (ACT == B200_ACT_GELU_ERF)
        o = __float2half_rn(gelu_erf(__half2float(o)));
    else if constexpr (ACT == B200_ACT_GELU_TANH)
        o = __float2half_rn(gelu_tanh(__half2float(o)));
    if (has_res)
        o = __float2half_rn(__half2float(o) + res);
    return o;
}

// Large-tile (M >= 64 rows per tile, unsplit) epilogue arithmetic.  ncu (source counters, M = 24000, 1280 -> 5120 with
// GELU) showed that kernel issue-bound on its EPILOGUE: 79 warp instructions per 32 outputs, led by FSEL (erff evaluates
// both of its argument ranges and selects), per-row bounds checks and 64-bit index arithmetic.  This variant
//   * adds bias and residual with HADD2: the sum of two fp16 values is exact in fp32 whenever it can influence the fp16
//     rounding, so __hadd(a, b) == fp16(float(a) + float(b)) bit for bit (what finish_output computes);
//   * evaluates erf branch-free (Abramowitz & Stegun 7.1.26 with MUFU.RCP / MUFU.EX2, |error| <= 2e-7): over all 63488
//     finite fp16 inputs the rounded GELU differs from the erff-based one in 0.7 % of them, by at most 2 fp16 ulp
//     (3.1e-5 absolute).  The decode-sized epilogues keep erff (finish_output_rt), so the decoder's bits do not change.
#ifndef B200_TC_FAST_ERF
#define B200_TC_FAST_ERF 1
#endif
__device__ __forceinline__ float erf_branch_free(float x)
{
    const float ax = fabsf(x);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
    float pl = fmaf(1.061405429f, t, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
    return copysignf(fmaf(-pl * t, e, 1.0f), x);
}

template <int ACT>
__device__ __forceinline__ __half finish_output_tile(float acc, bool has_bias, __half bias_h, bool has_res, __half res_h)
{
    __half o = __float2half_rn(acc);
    if (has_bias)
        o = __hadd(o, bias_h);
    if constexpr (ACT == B200_ACT_GELU_ERF)
    {
        const float xf = __half2float(o);
#if B200_TC_FAST_ERF
        o = __float2half_rn(0.5f * xf * (1.0f + erf_branch_free(xf * 0.70710678118654752440f)));
#else
        o = __float2half_rn(gelu_erf(xf));
#endif
    }
    else if constexpr (ACT == B200_ACT_GELU_TANH)
        o = __float2half_rn(gelu_tanh(__half2float(o)));
    if (has_res)
        o = __hadd(o, res_h);
    return o;
}

// Run-time activation: used by the decode-sized cluster reduction, where ONE compact instruction stream matters more than
// the masked-off instructions (six specialised copies of that epilogue measured 8 % slower on the decoder step:
// instruction-cache misses on the critical path of a 3 us kernel).
// (out of line on purpose: the inlined erff / tanhf bodies sat, mostly unexecuted, in every copy of the finishing code)
__device__ __noinline__ float activation_rt(float x, int activation)
{
    return activation == B200_ACT_GELU_ERF ? gelu_erf(x) : gelu_tanh(x);
}

__device__ __forceinline__ __half finish_output_rt(float acc, bool has_bias, float biasv, int activation, bool has_res, float res)
{
    __half o = __float2half_rn(acc);
    if (has_bias)
        o = __float2half_rn(__half2float(o) + biasv);
    if (activation != B200_ACT_NONE)
        o = __float2half_rn(activation_rt(__half2float(o), activation));
    if (has_res)
        o = __float2half_rn(__half2float(o) + res);
    return o;
}

// calls f(activation tag, folded-LayerNorm tag) with compile-time constants; the branch is warp-uniform
template <typename F>
__device__ __forceinline__ void tc_dispatch(int activation, bool fold, F&& f)
{
    if (activation == B200_ACT_GELU_ERF)
    {
        if (fold)
            f(std::integral_constant<int, B200_ACT_GELU_ERF>{}, std::true_type{});
        else
            f(std::integral_constant<int, B200_ACT_GELU_ERF>{}, std::false_type{});
    }
    else if (activation == B200_ACT_GELU_TANH)
    {
        if (fold)
            f(std::integral_constant<int, B200_ACT_GELU_TANH>{}, std::true_type{});
        else
            f(std::integral_constant<int, B200_ACT_GELU_TANH>{}, std::false_type{});
    }
    else
    {
        if (fold)
            f(std::integral_constant<int, B200_ACT_NONE>{}, std::true_type{});
        else
            f(std::integral_constant<int, B200_ACT_NONE>{}, std::false_type{});
    }
}

__device__ __forceinline__ void tc_ld_x8p(uint32_t taddr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}


constexpr int kWTileBytes = 64 * 128; // 64 row pairs x 128 B = 128 columns x 64 k

constexpr int kTcDequantWarps = 8;
constexpr int kTcThreads = (kTcDequantWarps + 2) * 32; // + TMA producer warp + MMA/TMEM warp

// Shared-memory carve-up, shared by the kernel and the host-side size computation.
template <int MT, int SS, int AS>
struct TcSmem
{
    static constexpr int XTileBytes = MT * 128;
    static constexpr size_t ring = (size_t) SS * (kWTileBytes + XTileBytes);
    static constexpr size_t rbuf = (size_t) 128 * MT * sizeof(float); // cluster reduction inbox [S][MT][128/S]
    static_assert(AS <= SS, "the TMEM A ring is never deeper than the shared-memory ring");
    static constexpr size_t bars = (sizeof(uint64_t) * (3 * SS + AS + 3) + 16 + 15) & ~size_t(15); // keeps what follows 16-byte aligned

    static constexpr size_t ln_bytes(int K) // folded LayerNorm: gamma [K] fp16 + (mean, M2) per sender rank and row
    {
        return (size_t) K * 2 + (size_t) 9 * MT * 2 * sizeof(float);
    }

    static constexpr size_t total(bool cluster, bool ln, int K)
    {
        return 1024 + ring + bars + (cluster ? rbuf : 0) + (ln ? ln_bytes(K) : 0);
    }
};

// CL = true: the instantiation launched for cluster split-K contains ONLY that epilogue (no unsplit, no global-slab
// code): a smaller instruction footprint is measurably faster for the decode-sized kernels.
template <int MT, int SS, int AS, bool CL>
__global__ void __launch_bounds__(kTcThreads, MT <= 32 ? 2 : 1)
    woq_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const TcParams p)
{
    using L = TcSmem<MT, SS, AS>;
    const bool is_cluster = CL || p.cluster != 0;
    constexpr int XTileBytes = L::XTileBytes;
    constexpr uint32_t kTmemCols = tmem_cols_pow2(32 * AS + MT);
    constexpr uint32_t kDCol = 32 * AS;
    constexpr int kHalfCols = MT / 2; // accumulator columns handled by one k-half warp group in the epilogue
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 (1<<4), a=b=f16 (0), K-major both,
    // N>>3 at bit 17, M>>4 at bit 24
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t) (MT >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
    constexpr int kProducerWarp = kTcDequantWarps, kMmaWarp = kTcDequantWarps + 1;
    constexpr int kDq = kTcDequantWarps * 32;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms (TMA and UMMA agree on address bits 7..9)
    // (pointer arithmetic on the __shared__ array itself, so the compiler keeps the shared address space: LDS/STS, not
    // generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smW = smem;
    uint8_t* smX = smem + SS * kWTileBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smX + SS * XTileBytes); // weight tile of block i has landed
    uint64_t* xfull = full + SS;       // activation tile of block i has landed (separate: the weights never wait for x)
    uint64_t* stage_free = xfull + SS; // MMA of block i done (index i % SS): its smem stage and TMEM A stage are reusable
    uint64_t* a_ready = stage_free + SS; // dequantized A tile of block i is in TMEM stage i % AS
    uint64_t* acc_done = a_ready + AS;
    uint64_t* red_bar = acc_done + 1;  // cluster inbox: partial accumulators of all ranks have landed
    uint64_t* stat_bar = red_bar + 1;  // cluster inbox: LayerNorm partials of all ranks have landed (they arrive earlier)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stat_bar + 1);
    int* last_flag = reinterpret_cast<int*>(tmem_slot + 1);
    uint8_t* extra = reinterpret_cast<uint8_t*>(full) + L::bars;
    float* rbuf = reinterpret_cast<float*>(extra);                       // cluster mode: [S][MT][128/S]
    uint8_t* ln_base = extra + (is_cluster ? L::rbuf : 0);
    __half* ln_g = reinterpret_cast<__half*>(ln_base);                   // folded LN: gamma of this split's k range
    float* ln_part = reinterpret_cast<float*>(ln_g + p.K);               // [sender rank <= 8][MT][2] partial mean, M2
    float* ln_fin = ln_part + 8 * MT * 2;                                // [MT][2] merged mean, rstd

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
    // (32-bit: split <= 16 and kb_total = K / 64; the 64-bit form cost two calls into the division routine per CTA prologue)
    const int kb_begin = (int) (((unsigned) split * (unsigned) p.kb_total) / (unsigned) p.splits);
    const int kb_end = (int) (((unsigned) (split + 1) * (unsigned) p.kb_total) / (unsigned) p.splits);
    const int nkb = kb_end - kb_begin;
    const bool fold = p.fold_gamma != nullptr;
    if (threadIdx.x == 0)
    {
        TC_STAMP(0);
        TC_GT(atomicMin, 0);
        TC_GT(atomicMax, 3);
    }

    // ---- prologue.  The producer thread owns the `full` barriers: it initialises them and immediately starts the
    // weight stream (weights never depend on the previous kernel), while the other warps set up the rest. ----
    const int pre = nkb < SS ? nkb : SS;
    // the whole k range of this CTA fits the ring (always true for decode shapes): all activation tiles complete on
    // ONE barrier, so the MMA loop and the statistics wait once instead of once per k-block
    const bool x_single = nkb <= SS;
    // "A from shared memory" for the k-blocks beyond the TMEM ring (fc2 at batch <= 16: ten blocks, seven TMEM stages): instead
    // of recycling TMEM stages AFTER the dependency (wait for the MMAs of block i - AS, then dequantize block i), the dequant
    // warps write those blocks as fp16 K-major / 128B-swizzled UMMA A tiles over the int8 stages they have already
    // consumed (two 8 KB stages per 16 KB tile) and the MMA warp issues them in SS form -- all of the CTA's dequantisation
    // happens ahead of the dependency again.  Needs the tiles to fit into consumed stages: 2 (nkb - AS) <= AS.
    const bool a_smem = p.mma_burst != 0 && x_single && nkb > AS && 2 * (nkb - AS) <= AS;
    if (warp == kProducerWarp && elect_one_sync())
    {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmX); // first used right after the dependency wait: keep its fetch off the critical path
        for (int s = 0; s < SS; ++s)
        {
            mbar_init(&full[s], 1);
            mbar_init(&xfull[s], 1);
        }
        fence_mbar_init();
        fence_proxy_async_smem();
        for (int i = 0; i < pre; ++i)
        {
            mbar_arrive_expect_tx(&full[i], kWTileBytes);
            tma_load_2d(smW + i * kWTileBytes, &tmW, (kb_begin + i) * 128, n_tile * 64, &full[i]);
        }
        TC_STAMP(2);
    }
    {
        const int t = threadIdx.x;
        if (t < SS)
            mbar_init(&stage_free[t], 1);
        else if (t < SS + AS)
            mbar_init(&a_ready[t - SS], kTcDequantWarps);
        else if (t == SS + AS)
            mbar_init(acc_done, 1);
        else if (t == SS + AS + 1)
        {
            mbar_init(red_bar, 1);
            mbar_init(stat_bar, 1);
            if (is_cluster)
            {
                mbar_arrive_expect_tx(red_bar, (uint32_t) (128 * MT * sizeof(float))); // inbox bytes from all ranks
                if (fold)
                    mbar_arrive_expect_tx(stat_bar, (uint32_t) (p.splits * MT * 2 * sizeof(float)));
            }
        }
        if (t <= SS + AS + 1)
            fence_mbar_init();
    }
    if (warp == kMmaWarp)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // dequant threads: column scale requested early so its latency hides behind the barrier
    const int T = ((warp & 3) << 5) | lane; // weight column inside the tile == TMEM lane (warps 0-3 and 4-7 alias)
    const int kh = (warp >> 2) & 1;         // which half of the 64-wide k-block / of the accumulator columns
    const int n = n_tile * 128 + T;
    __half sc = __float2half(0.f);
    if (warp < kTcDequantWarps && n < p.N)
        sc = __ldg(p.scales + n);
    // gamma never depends on the previous kernel: this split's slice is REQUESTED here and stored to shared memory after
    // the CTA / cluster barriers below (storing it here put the load's latency into every CTA's prologue)
    uint4 g_reg = make_uint4(0u, 0u, 0u, 0u);
    const bool g_one = nkb * 8 <= kDq; // one 16-byte piece per dequant thread at most (always true for decode launches)
    if (fold && warp < kTcDequantWarps)
    {
        const uint4* g4 = reinterpret_cast<const uint4*>(p.fold_gamma) + (size_t) kb_begin * 8;
        if (g_one)
        {
            if ((int) threadIdx.x < nkb * 8)
                g_reg = __ldg(g4 + threadIdx.x);
        }
        else
        {
            for (int idx = threadIdx.x; idx < nkb * 8; idx += kDq)
                reinterpret_cast<uint4*>(ln_g)[idx] = __ldg(g4 + idx);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (is_cluster)
    {
        // every rank's inbox barrier is armed before anybody can push into it.  RELAXED arrival: the barrier initialisation
        // is already published cluster-wide by fence.mbarrier_init.release.cluster; arrive.release would add a full
        // MEMBAR.ALL.GPU + ERRBAR to every CTA's prologue (seen as such in the SASS and as 1-2 us in "prologue done")
        asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    grid_dep_launch_dependents(); // PDL: the next kernel may start its own prologue / weight prefetch now
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    if (threadIdx.x == 0)
        TC_STAMP(1);

    if (warp == kProducerWarp)
    {
        // ===== TMA producer (continued) =====
        if (elect_one_sync())
        {
            grid_dep_wait();
            TC_STAMP(15);
            TC_GT(atomicMin, 1);
            if (p.x3_depth > 0)
            {
                // the whole k range of this CTA as one 3-D box (a rank with one block less also receives its
                // neighbour's first block into an unused stage: the byte count is the full box either way)
                mbar_arrive_expect_tx(&xfull[0], (uint32_t) (p.x3_depth * XTileBytes));
                tma_load_3d(smX, &tmX, 0, m_tile * MT, kb_begin, &xfull[0]);
            }
            else if (x_single)
            {
                mbar_arrive_expect_tx(&xfull[0], (uint32_t) (nkb * XTileBytes));
                for (int i = 0; i < nkb; ++i)
                    tma_load_2d(smX + i * XTileBytes, &tmX, (kb_begin + i) * 64, m_tile * MT, &xfull[0]);
            }
            else
            {
                for (int i = 0; i < pre; ++i)
                {
                    mbar_arrive_expect_tx(&xfull[i], XTileBytes);
                    tma_load_2d(smX + i * XTileBytes, &tmX, (kb_begin + i) * 64, m_tile * MT, &xfull[i]);
                }
            }
            for (int i = pre; i < nkb; ++i)
            {
                const int ss = i % SS;
                mbar_wait(&stage_free[ss], ((i / SS) - 1) & 1);
                mbar_arrive_expect_tx(&full[ss], kWTileBytes);
                tma_load_2d(smW + ss * kWTileBytes, &tmW, (kb_begin + i) * 128, n_tile * 64, &full[ss]);
                mbar_arrive_expect_tx(&xfull[ss], XTileBytes);
                tma_load_2d(smX + ss * XTileBytes, &tmX, (kb_begin + i) * 64, m_tile * MT, &xfull[ss]);
            }
        }
    }
    else if (warp == kMmaWarp)
    {
        // ===== MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues =====
        const uint32_t d_tmem = tmem_base + kDCol;
        // Decode launches (the CTA's whole k range fits the TMEM A ring and arrives on ONE activation barrier): the per
        // k-block loop below costs ~490 cycles per block on the critical path (barrier try-wait, fence, election,
        // descriptor set-up in front of every four MMAs: 2200 of the 5300 cycles a 5-block projection spends after its
        // dependency, tools/tc_timing_insitu.py).  Here every A stage is awaited BEFORE the activation barrier (they complete
        // ahead of the dependency), and once the activations land all 4 nkb MMAs are issued back to back.
        // A CTA with more k-blocks than TMEM A stages (fc2: ten blocks, six stages) bursts the first AS blocks and walks the
        // recycled stages with the per-block loop.
        const bool burst = x_single && nkb > 0 && tc_burst_enabled(p);
        int i0 = 0;
        if (burst)
        {
            const int nb = (nkb < AS || a_smem) ? nkb : AS;
            for (int i = 0; i < nb; ++i)
                mbar_wait(&a_ready[i % AS], (i / AS) & 1);
            mbar_wait(&xfull[0], 0);
            tc_fence_after();
            if (lane == 0)
                TC_STAMP(16 + 2);
            const uint64_t bdesc0 = umma_desc_k_sw128(smem_u32(smX));
            if (elect_one_sync())
            {
#pragma unroll 1
                for (int i = 0; i < nb; ++i)
                {
                    // stage i: activation tile i (XTileBytes apart: the descriptor's 14-bit start address counts 16-byte
                    // units) against TMEM A stage i
                    const uint64_t bdesc = bdesc0 + (uint64_t) (i * (XTileBytes >> 4));
                    if (i < AS)
                    {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc_mma_ts(d_tmem, tmem_base + i * 32 + k4 * 8, bdesc + 2 * k4, kIdesc, (i | k4) != 0 ? 1u : 0u);
                        if (i + AS < nkb && !a_smem)
                            tc_commit(&stage_free[i]); // (i < AS <= SS: stage i) the dequant warps recycle TMEM stage i for block i + AS
                    }
                    else
                    {
                        // block i lives in shared memory as a K-major A tile (a_smem): SS form
                        const uint64_t adesc = umma_desc_k_sw128(smem_u32(smW + (size_t) (i - AS) * 2 * kWTileBytes));
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc_mma_ss(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, kIdesc, 1u);
                    }
                }
                if (nb == nkb)
                    tc_commit(acc_done);
            }
            __syncwarp();
            if (lane == 0)
                TC_STAMP(16 + 4 * (nb - 1 < 11 ? nb - 1 : 11) + 3);
            i0 = nb;
        }
        for (int i = i0; i < nkb; ++i)
        {
            const int ss = i % SS, as = i % AS;
            mbar_wait(&a_ready[as], (i / AS) & 1);
            if (!x_single)
                mbar_wait(&xfull[ss], (i / SS) & 1);
            else if (i == 0)
                mbar_wait(&xfull[0], 0);
            tc_fence_after();
            if (lane == 0 && i < 12)
                TC_STAMP(16 + 4 * i + 2);
            const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smX + ss * XTileBytes));
            if (elect_one_sync())
            {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                {
                    // K advance inside the 128-byte swizzle atom: 16 halves = 32 bytes = 2 descriptor units
                    tc_mma_ts(d_tmem, tmem_base + as * 32 + k4 * 8, bdesc + 2 * k4, kIdesc, (i | k4) != 0 ? 1u : 0u);
                }
                // stage_free is only consumed when a later block recycles this shared-memory or TMEM stage
                if (i + AS < nkb)
                    tc_commit(&stage_free[ss]);
                if (i == nkb - 1)
                    tc_commit(acc_done);
            }
            __syncwarp();
            if (lane == 0 && i < 12)
                TC_STAMP(16 + 4 * i + 3);
        }
        if (lane == 0)
            TC_STAMP(6);
    }
    else
    {
        // ===== warps 0..7: (LayerNorm of the activation tile,) dequant, then epilogue =====
        const int tq = threadIdx.x; // 0..255
        if (fold && g_one)
        {
            if (tq < nkb * 8)
                reinterpret_cast<uint4*>(ln_g)[tq] = g_reg;
            asm volatile("bar.sync 1, %0;" ::"n"(kDq) : "memory"); // gamma staged before any dequant thread multiplies by it
        }
        const int jl = T >> 1, hf = T & 1, sw = jl & 7;
        const float scf = __half2float(sc); // per-column dequant scale, applied in the epilogue
        const uint32_t lane_field = (uint32_t) ((warp & 3) * 32) << 16;

        // this thread's 32 bytes of k-block i: chunks 2*kh and 2*kh+1 of its 64-byte column slice
        uint4 v[2];
        auto load_w = [&](int i)
        {
            const int ss = i % SS;
            mbar_wait(&full[ss], (i / SS) & 1);
            const uint8_t* rowp = smW + ss * kWTileBytes + jl * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c)
                v[c] = *reinterpret_cast<const uint4*>(rowp + (((4 * hf + 2 * kh + c) ^ sw) << 4));
        };

        if (nkb > 0)
        {
            load_w(0);
            if (tq == 0)
                TC_STAMP(3);
        }
        for (int i = 0; i < nkb; ++i)
        {
            const int as = i % AS;
            uint32_t r[16];
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                const uint32_t words[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
                for (int w = 0; w < 4; ++w)
                {
                    __half2 lo, hi;
                    dequant_word(words[w], lo, hi); // exact integers; the column scale is applied to the fp32 accumulator
                    // chunk = k-slice of 16: column 8c+w holds k = 2w, 2w+1; column 8c+4+w holds 8+2w, 8+2w+1
                    r[8 * c + w] = *reinterpret_cast<uint32_t*>(&lo);
                    r[8 * c + 4 + w] = *reinterpret_cast<uint32_t*>(&hi);
                }
            }
            if (fold)
            {
                // r[idx] is the k pair kh*16 + idx of the block: scale it by the matching gamma pair (rounded to fp16
                // exactly like b200_woq_ln_fold_prepare does when it sums the column)
                const uint4* g4 = reinterpret_cast<const uint4*>(ln_g + i * 64 + kh * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                {
                    const uint4 g = g4[q];
                    const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                    {
                        const __half2 pr = __hmul2(*reinterpret_cast<const __half2*>(&r[4 * q + w]),
                            *reinterpret_cast<const __half2*>(&gw[w]));
                        r[4 * q + w] = *reinterpret_cast<const uint32_t*>(&pr);
                    }
                }
            }
            if (i >= AS && a_smem)
            {
                if (i == AS) // every thread has read the int8 tiles of the stages that are overwritten from here on
                    asm volatile("bar.sync 1, %0;" ::"n"(kDq) : "memory");
                // row T of the A tile (128 B = 64 k), this thread's k-half = 16-byte chunks 4 kh .. 4 kh + 3, XOR-swizzled by row
                uint8_t* arow = smW + (size_t) (i - AS) * 2 * kWTileBytes + T * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(arow + (((4 * kh + q) ^ (T & 7)) << 4))
                        = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                if (i + 1 < nkb)
                    load_w(i + 1);
                fence_proxy_async_smem(); // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&a_ready[as]);
                continue;
            }
            if (i >= AS)
            {
                // TMEM stage `as` was last read by the MMAs of block i - AS
                const int j = i - AS;
                mbar_wait(&stage_free[j % SS], (j / SS) & 1);
                tc_fence_after();
            }
            tc_st_x16(tmem_base + lane_field + as * 32 + kh * 16, r);
            // overlap the TMEM store with fetching the next block
            if (i + 1 < nkb)
                load_w(i + 1);
            if (tq == 0 && i < 12)
                TC_STAMP(16 + 4 * i + 0);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&a_ready[as]);
            if (tq == 0 && i < 12)
                TC_STAMP(16 + 4 * i + 1);
            if (i == 0 && tq == 0)
                TC_STAMP(4);
        }

        const int m_valid = min(MT, p.M - m_tile * MT);
        // cluster geometry of the split-K reduction (also used by the LayerNorm statistics exchange)
        const uint32_t S = (uint32_t) p.splits;
        const uint32_t my_rank = is_cluster ? cluster_ctarank() : 0u;
        // cluster mode: element (ml, n) goes to the rank that owns column slice n / nslice
        const int nslice = 128 / (int) S;
        uint32_t push_addr = 0, push_bar = 0;
        const bool has_bias = p.bias != nullptr;
        // static operands (weights-side vectors) of the final outputs this thread will write, requested before the
        // accumulator is waited for
        float own_bias = 0.f, own_c1 = 0.f, own_c2 = 0.f;
        // raw bits of the residual values, converted where they are used (a half -> float conversion right after the load
        // made every warp wait for the load BEFORE it pushed its partial sums: 4 % of all warp samples on that one line)
        unsigned short own_res[4] = {0, 0, 0, 0};
        // cluster mode: this thread reduces elements (ml_first + j*ml_step, own_nn), j = 0, 1, ... -- one fixed column
        // (256 threads are a multiple of the slice width), so bias and the fold vectors are per-thread constants
        const int ns_shift = __ffs(nslice) - 1; // nslice is a power of two
        const int ml_first = tq >> ns_shift, ml_step = kDq >> ns_shift;
        int own_nn = 0;
        bool own_valid = false;
        if (is_cluster)
        {
            const uint32_t owner = (uint32_t) (T >> ns_shift);
            // inbox layout [sender rank][ml][nl]
            push_addr = mapa_u32(smem_u32(rbuf + ((size_t) my_rank * MT) * nslice + (T & (nslice - 1))), owner);
            push_bar = mapa_u32(smem_u32(red_bar), owner);
            own_nn = n_tile * 128 + (int) my_rank * nslice + (tq & (nslice - 1));
            own_valid = ml_first < m_valid && own_nn < p.N;
        }
        else
        {
            own_nn = n;
            own_valid = n < p.N;
        }
        if (own_valid)
        {
            if (has_bias)
                own_bias = __half2float(__ldg(p.bias + own_nn));
            if (fold)
            {
                own_c1 = __ldg(p.fold_c1s + own_nn);
                own_c2 = __ldg(p.fold_c2 + own_nn);
            }
        }
        float fin_cn = 0.f; // element count of the partial this thread will merge (rank tq & 7 of row tq >> 3)
        if (fold && (tq >> 3) < MT && (tq & 7) < (int) S)
        {
            const int r = tq & 7;
            fin_cn = 64.f * (float) (((r + 1) * p.kb_total) / p.splits - (r * p.kb_total) / p.splits);
        }
        if constexpr (MT <= 32)
        {
            if (fold)
            {
                // ---- LayerNorm statistics without touching global memory: every CTA of the split-K cluster reduces
                // the rows of ITS k range straight from the activation tiles the MMAs consume (shared memory, TMA
                // already brought them), as (count, mean, M2); the partials travel to all ranks next to the partial
                // accumulators (same inbox barrier) and are merged with Chan's formula -- two-pass accuracy, no
                // extra round trip.  TPR threads share a row: 8 chunks of 8 halves x G k-block phases. ----
                constexpr int TPR = kDq / MT, G = TPR / 8;
                const int row = tq / TPR, c = tq & 7, g = (tq & (TPR - 1)) >> 3;
                if (tq == 0)
                    TC_STAMP(13);
                // per thread: sums of (v - sh) and (v - sh)^2 with sh = the first value it sees (a sample of the row, so
                // the cancellation in M2 = b - a^2/n is benign); no divisions in the loop
                float sh = 0.f, sa = 0.f, sb = 0.f;
                int cnt = 0;
                for (int i = g; i < nkb; i += G)
                {
                    const int ss = i % SS;
                    if (!x_single)
                        mbar_wait(&xfull[ss], (i / SS) & 1);
                    else if (i == g)
                        mbar_wait(&xfull[0], 0);
                    const uint4 u = *reinterpret_cast<const uint4*>(smX + ss * XTileBytes + row * 128 + ((c ^ (row & 7)) << 4));
                    const __half2* h = reinterpret_cast<const __half2*>(&u);
                    if (i == g)
                        sh = __low2float(h[0]);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                    {
                        const float2 f = __half22float2(h[q]);
                        const float d0 = f.x - sh, d1 = f.y - sh;
                        sa += d0 + d1;
                        sb = fmaf(d0, d0, fmaf(d1, d1, sb));
                    }
                    cnt += 8;
                }
                float cn = (float) cnt;
                const float rn = cnt > 0 ? __fdividef(1.f, cn) : 0.f;
                float cm = sh + sa * rn, cM2 = sb - sa * sa * rn;
                // the 8 chunk threads of one k-block phase hold equal counts: exact halving merges
#pragma unroll
                for (int o = 1; o < 8; o <<= 1)
                {
                    const float om = __shfl_xor_sync(0xffffffffu, cm, o);
                    const float oM = __shfl_xor_sync(0xffffffffu, cM2, o);
                    const float dl = om - cm;
                    cm += 0.5f * dl;
                    cM2 += oM + dl * dl * (0.5f * cn);
                    cn *= 2.f;
                }
                if constexpr (G == 2)
                {
                    // the two phases may differ by one k-block: general Chan merge
                    const float on = __shfl_xor_sync(0xffffffffu, cn, 8);
                    const float om = __shfl_xor_sync(0xffffffffu, cm, 8);
                    const float oM = __shfl_xor_sync(0xffffffffu, cM2, 8);
                    const float nn = cn + on, dl = om - cm;
                    const float inv = nn > 0.f ? __fdividef(1.f, nn) : 0.f;
                    cm += dl * on * inv;
                    cM2 += oM + dl * dl * cn * on * inv;
                    cn = nn;
                }
                // every thread of the row now holds the CTA's partial: thread r of the row pushes it to rank r
                {
                    const uint32_t r = (uint32_t) (tq & (TPR - 1));
                    float* mine = ln_part + ((size_t) my_rank * MT + row) * 2; // [sender rank][row][mean, M2]
                    if (is_cluster)
                    {
                        if (r < S)
                        {
                            const uint32_t dst = mapa_u32(smem_u32(mine), r), bar = mapa_u32(smem_u32(stat_bar), r);
                            st_async_f32(dst, cm, bar);
                            st_async_f32(dst + 4u, cM2, bar);
                        }
                    }
                    else if (r == 0)
                    {
                        mine[0] = cm;
                        mine[1] = cM2;
                    }
                }
                if (tq == 0)
                    TC_STAMP(14);
            }
        }
        // merges the per-rank partials into (mean, rstd) per row: thread (row, r) takes rank r's partial (it covers
        // k-blocks [r*kb_total/S, (r+1)*kb_total/S), 64 elements each), 3 shuffle rounds of Chan merges
        auto ln_finish = [&]()
        {
            const int row = tq >> 3, r = tq & 7;
            float cn = fin_cn, cm = 0.f, cM2 = 0.f;
            if (row < MT && r < (int) S)
            {
                cm = ln_part[(r * MT + row) * 2];
                cM2 = ln_part[(r * MT + row) * 2 + 1];
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1)
            {
                const float on = __shfl_xor_sync(0xffffffffu, cn, o);
                const float om = __shfl_xor_sync(0xffffffffu, cm, o);
                const float oM = __shfl_xor_sync(0xffffffffu, cM2, o);
                const float nn = cn + on, dl = om - cm;
                const float inv = nn > 0.f ? __fdividef(1.f, nn) : 0.f;
                cm += dl * on * inv;
                cM2 += oM + dl * dl * cn * on * inv;
                cn = nn;
            }
            if (row < MT && r == 0)
            {
                ln_fin[2 * row] = cm;
                ln_fin[2 * row + 1] = rsqrtf(__fdividef(cM2, cn) + p.ln_eps);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kDq) : "memory");
        };
        // (folded LayerNorm: y = rstd * (acc - mean * c1s) + c2, applied in the epilogue variants below)

        // ---- epilogue: thread (T, kh) owns accumulator row T, columns [kh*MT/2, (kh+1)*MT/2) ----
        const int n_tiles = gridDim.x, m_tiles = gridDim.y;
        const int tile_id = m_tile * n_tiles + n_tile;
        const bool direct = !CL && (p.splits == 1);
        const bool has_res = p.residual != nullptr;
        const size_t slab_elems = (size_t) 128 * MT;
        float* slab = (direct || is_cluster)
            ? nullptr
            : p.slabs + ((size_t) split * m_tiles * n_tiles + tile_id) * slab_elems + (size_t) T * MT + kh * kHalfCols;
        mbar_wait(acc_done, 0);
        tc_fence_after();
        // the residual is an earlier kernel's output: it may only be read after the dependency wait, which acc_done
        // implies (activation TMA -> MMA -> commit); its latency hides behind the cluster exchange below
        if (own_valid && is_cluster && has_res)
        {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ml_first + j * ml_step < m_valid)
                    own_res[j] = __ldcg(reinterpret_cast<const unsigned short*>(p.residual)
                        + (size_t) (m_tile * MT + ml_first + j * ml_step) * p.ldc + own_nn);
        }
        if (tq == 0)
            TC_STAMP(7);
        if (fold && !is_cluster)
        {
            asm volatile("bar.sync 1, %0;" ::"n"(kDq) : "memory"); // this CTA's own partial is complete
            ln_finish();
        }
        if constexpr (!CL)
        if (direct)
        {
            if constexpr (MT >= 64)
            {
            // ---- unsplit tile (large M): every thread finishes its column for its half of the rows; TMEM is read in
            // batches of up to 32 columns per wait ----
            tc_dispatch(p.activation, fold, [&](auto act_c, auto fold_c)
            {
                constexpr int ACT = decltype(act_c)::value;
                constexpr bool FOLD = decltype(fold_c)::value;
                constexpr int CB = kHalfCols >= 32 ? 32 : kHalfCols;
#pragma unroll 1
                const __half bias_h = __float2half_rn(own_bias); // own_bias came from an fp16 value: exact
                for (int cb = 0; cb < kHalfCols / CB; ++cb)
                {
                    const int ml0 = kh * kHalfCols + cb * CB;
                    const int rows_here = m_valid - ml0; // rows of this batch inside the matrix (may be <= 0)
                    const size_t base = (size_t) (m_tile * MT + ml0) * p.ldc + n;
                    const bool full = rows_here >= CB;
                    // The residual usually IS the output buffer (x += ...), so the compiler must keep every residual
                    // load behind the store of the row before it: one exposed load latency per row (it made the
                    // M = 24000 GEMMs with a residual 2-5x slower than the plain ones).  Fetch the batch's residuals
                    // up front, ahead of the TMEM read, and keep them in registers.
                    __half resv[CB];
                    if (has_res && n < p.N)
                    {
                        const __half* rp = p.residual + base;
#pragma unroll
                        for (int i = 0; i < CB; ++i)
                            if (full || i < rows_here)
                                resv[i] = rp[(size_t) i * p.ldc];
                    }
                    uint32_t acc[CB];
#pragma unroll
                    for (int q = 0; q < CB / 8; ++q)
                        tc_ld_x8p(tmem_base + lane_field + kDCol + kh * kHalfCols + cb * CB + q * 8, &acc[q * 8]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (n < p.N)
                    {
                        __half* cp = p.C + base;
                        if (full)
                        {
                            // whole batch inside the matrix (all but the last m-tile): no per-row predicate
#pragma unroll
                            for (int i = 0; i < CB; ++i)
                            {
                                float v = __uint_as_float(acc[i]) * scf;
                                if constexpr (FOLD)
                                    v = ln_fin[2 * (ml0 + i) + 1] * (v - ln_fin[2 * (ml0 + i)] * own_c1) + own_c2;
                                cp[(size_t) i * p.ldc] = finish_output_tile<ACT>(v, has_bias, bias_h, has_res, resv[i]);
                            }
                        }
                        else
                        {
#pragma unroll
                            for (int i = 0; i < CB; ++i)
                                if (i < rows_here)
                                {
                                    float v = __uint_as_float(acc[i]) * scf;
                                    if constexpr (FOLD)
                                        v = ln_fin[2 * (ml0 + i) + 1] * (v - ln_fin[2 * (ml0 + i)] * own_c1) + own_c2;
                                    cp[(size_t) i * p.ldc] = finish_output_tile<ACT>(v, has_bias, bias_h, has_res, resv[i]);
                                }
                        }
                    }
                }
            });
                    }
            else
            {
                // decode-sized tiles: one compact instruction stream (code size is latency in a 3 us kernel)
#pragma unroll 1
                for (int c8 = 0; c8 < kHalfCols / 8; ++c8)
                {
                    uint32_t acc[8];
                    tc_ld_x8(tmem_base + lane_field + kDCol + kh * kHalfCols + c8 * 8, acc);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const int ml0 = kh * kHalfCols + c8 * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (ml0 + i < m_valid && n < p.N)
                        {
                            const int ml = ml0 + i;
                            float v = __uint_as_float(acc[i]) * scf;
                            if (fold)
                                v = ln_fin[2 * ml + 1] * (v - ln_fin[2 * ml] * own_c1) + own_c2;
                            const size_t idx = (size_t) (m_tile * MT + ml) * p.ldc + n;
                            const float res = has_res ? __half2float(p.residual[idx]) : 0.f;
                            p.C[idx] = finish_output_rt(v, has_bias, own_bias, p.activation, has_res, res);
                        }
                }
            }
        }
#pragma unroll 1
        for (int c8 = 0; (CL || !direct) && c8 < kHalfCols / 8; ++c8)
        {
            uint32_t acc[8];
            tc_ld_x8(tmem_base + lane_field + kDCol + kh * kHalfCols + c8 * 8, acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int ml0 = kh * kHalfCols + c8 * 8;
            if (is_cluster)
            {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    st_async_f32(push_addr + (uint32_t) ((ml0 + i) * nslice) * 4u, __uint_as_float(acc[i]) * scf, push_bar);
            }
            else if constexpr (!CL)
            {
                float4* dst = reinterpret_cast<float4*>(slab + c8 * 8);
                __stcg(dst, make_float4(__uint_as_float(acc[0]) * scf, __uint_as_float(acc[1]) * scf,
                                __uint_as_float(acc[2]) * scf, __uint_as_float(acc[3]) * scf));
                __stcg(dst + 1, make_float4(__uint_as_float(acc[4]) * scf, __uint_as_float(acc[5]) * scf,
                                    __uint_as_float(acc[6]) * scf, __uint_as_float(acc[7]) * scf));
            }
        }
        if (tq == 0)
            TC_STAMP(8);
        // the accumulator has been read: let the MMA warp free the TMEM columns while the reduction runs
        tc_fence_before();
        asm volatile("bar.arrive 2, %0;" ::"n"(kDq + 32) : "memory");
        if (is_cluster)
        {
            // ---- split-K reduction through distributed shared memory: every rank received the partial values of its
            // column slice from all ranks (st.async + complete_tx on its inbox barrier) and sums them in rank order ----
            if (fold)
            {
                // the LayerNorm partials were pushed before the accumulators: merge them while those are in flight
                mbar_wait(stat_bar, 0);
                ln_finish();
            }
            if (tq == 0)
                TC_STAMP(11);
            mbar_wait(red_bar, 0);
            if (tq == 0)
                TC_STAMP(9);
            if (own_valid)
            {
                const int nl = tq & (nslice - 1);
                auto reduce_one = [&](int ml, float res)
                {
                    const size_t idx = (size_t) (m_tile * MT + ml) * p.ldc + own_nn;
                    const float* src = rbuf + (size_t) ml * nslice + nl;
                    float sum = 0.f;
#pragma unroll
                    for (uint32_t q = 0; q < 8; ++q)
                        if (q < S)
                            sum += src[(size_t) q * MT * nslice];
#pragma unroll 1
                    for (uint32_t q = 8; q < S; ++q) // 16-CTA clusters (opt-in): rolled, off the usual path
                        sum += src[(size_t) q * MT * nslice];
                    if (fold)
                        sum = ln_fin[2 * ml + 1] * (sum - ln_fin[2 * ml] * own_c1) + own_c2;
                    p.C[idx] = finish_output_rt(sum, has_bias, own_bias, p.activation, has_res, res);
                };
                // rows of this thread that get an unrolled copy of the finishing code (with the prefetched residual): 16-row
                // tiles split eight ways -- every projection with a residual -- have ONE row per thread, split four ways (qkv,
                // fc1: no residual) two; every copy less is shorter code on the critical path (four copies 1.198 ms, two
                // 1.185, one + the rolled loop below 1.168 with the other trims)
                constexpr int kUnrolledRows = MT <= 16 ? 1 : 4;
#pragma unroll
                for (int j = 0; j < kUnrolledRows; ++j)
                    if (ml_first + j * ml_step < m_valid)
                        reduce_one(ml_first + j * ml_step, __half2float(__ushort_as_half(own_res[j])));
                for (int ml = ml_first + kUnrolledRows * ml_step; ml < m_valid; ml += ml_step) // further rows of this thread
                    reduce_one(ml, has_res ? __half2float(p.residual[(size_t) (m_tile * MT + ml) * p.ldc + own_nn]) : 0.f);
            }
            if (tq == 0)
                TC_STAMP(10);
        }
        else if constexpr (!CL)
        if (!direct)
        {
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(kDq) : "memory");
            if (tq == 0)
            {
                const int old = atomicAdd(&p.counters[tile_id], 1);
                *last_flag = (old == p.splits - 1) ? 1 : 0;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kDq) : "memory");
            if (*last_flag)
            {
                __threadfence();
                // deterministic reduction in split order; all loads of an 8-column group are independent and issued
                // before the adds (SPLIT_UNROLL splits x 2 x 128-bit loads in flight per thread)
                const float* base = p.slabs + (size_t) tile_id * slab_elems + (size_t) T * MT + kh * kHalfCols;
                const size_t split_stride = (size_t) m_tiles * n_tiles * slab_elems;
#pragma unroll 1
                for (int c8 = 0; c8 < kHalfCols / 8; ++c8)
                {
                    const int ml0 = kh * kHalfCols + c8 * 8;
                    if (ml0 >= m_valid)
                        break;
                    float sum[8], res[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        sum[i] = 0.f;
                        res[i] = (has_res && ml0 + i < m_valid && n < p.N)
                            ? __half2float(p.residual[(size_t) (m_tile * MT + ml0 + i) * p.ldc + n])
                            : 0.f;
                    }
                    constexpr int SPLIT_UNROLL = 4;
                    for (int s0 = 0; s0 < p.splits; s0 += SPLIT_UNROLL)
                    {
                        uint4 w4[SPLIT_UNROLL][2];
#pragma unroll
                        for (int u = 0; u < SPLIT_UNROLL; ++u)
                        {
                            const int sidx = min(s0 + u, p.splits - 1);
                            const uint4* src = reinterpret_cast<const uint4*>(base + (size_t) sidx * split_stride + c8 * 8);
                            w4[u][0] = __ldcg(src);
                            w4[u][1] = __ldcg(src + 1);
                        }
#pragma unroll
                        for (int u = 0; u < SPLIT_UNROLL; ++u)
                        {
                            if (s0 + u < p.splits)
                            {
#pragma unroll
                                for (int q4 = 0; q4 < 2; ++q4)
                                {
                                    sum[4 * q4 + 0] += __uint_as_float(w4[u][q4].x);
                                    sum[4 * q4 + 1] += __uint_as_float(w4[u][q4].y);
                                    sum[4 * q4 + 2] += __uint_as_float(w4[u][q4].z);
                                    sum[4 * q4 + 3] += __uint_as_float(w4[u][q4].w);
                                }
                            }
                        }
                    }
                    if (n < p.N)
                    {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (ml0 + i < m_valid)
                                p.C[(size_t) (m_tile * MT + ml0 + i) * p.ldc + n]
                                    = finish_output_rt(sum[i], has_bias, own_bias, p.activation, has_res, res[i]); // (no fold here)
                    }
                }
                if (tq == 0)
                    p.counters[tile_id] = 0; // self-reset for the next launch using this slot
            }
        }
    }

    if (threadIdx.x == 0)
        TC_GT(atomicMax, 2);
    // ---- teardown: no CTA-wide barrier; the MMA warp waits for the epilogue warps' last TMEM read only ----
    if (warp == kMmaWarp)
    {
        asm volatile("bar.sync 2, %0;" ::"n"(kDq + 32) : "memory");
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        if (lane == 0)
            TC_STAMP(12);
    }
}


// =====================================================================================================
// fp16 x fp16 -> fp32 "swap-AB" GEMM for the logits projection: out[m, v] = sum_k X[m, k] * E[v, k].
// E (token embedding, [vocab, K] fp16, not quantized in the reference: model.py:231,290) is the UMMA A operand
// straight from TMA-staged shared memory (128 vocabulary rows per CTA), X the B operand (MT <= 256 rows).
// =====================================================================================================
struct LogitsParams
{
    float* out; // [M, vocab] fp32
    int M, vocab, kb_total;
};

template <int MT, int SS>
__global__ void __launch_bounds__(192, 1)
    fp16_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmX, const LogitsParams p)
{
    constexpr int ETileBytes = 128 * 128;
    constexpr int XTileBytes = MT * 128;
    constexpr uint32_t kTmemCols = tmem_cols_pow2(MT);
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t) (MT >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // (pointer arithmetic on the __shared__ array itself, so the compiler keeps the shared address space: LDS/STS, not
    // generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smE = smem;
    uint8_t* smX = smem + SS * ETileBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smX + SS * XTileBytes);
    uint64_t* smem_free = full + SS;
    uint64_t* acc_done = smem_free + SS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y;
    const int nkb = p.kb_total;

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
        tma_prefetch_desc(&tmE);
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
    grid_dep_launch_dependents(); // PDL: the next kernel may start its own prologue / weight prefetch now
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 4)
    {
        if (elect_one_sync())
        {
            const int pre = nkb < SS ? nkb : SS;
            for (int i = 0; i < pre; ++i)
            {
                mbar_arrive_expect_tx(&full[i], ETileBytes + XTileBytes);
                tma_load_2d(smE + i * ETileBytes, &tmE, i * 64, n_tile * 128, &full[i]);
            }
            grid_dep_wait();
            for (int i = 0; i < pre; ++i)
                tma_load_2d(smX + i * XTileBytes, &tmX, i * 64, m_tile * MT, &full[i]);
            for (int i = pre; i < nkb; ++i)
            {
                const int ss = i % SS;
                mbar_wait(&smem_free[ss], ((i / SS) - 1) & 1);
                mbar_arrive_expect_tx(&full[ss], ETileBytes + XTileBytes);
                tma_load_2d(smE + ss * ETileBytes, &tmE, i * 64, n_tile * 128, &full[ss]);
                tma_load_2d(smX + ss * XTileBytes, &tmX, i * 64, m_tile * MT, &full[ss]);
            }
        }
    }
    else if (warp == 5)
    {
        // the whole warp runs the warp-uniform loop; one elected lane issues the tensor-core instructions
        for (int i = 0; i < nkb; ++i)
        {
            const int ss = i % SS;
            mbar_wait(&full[ss], (i / SS) & 1);
            tc_fence_after();
            const uint64_t adesc = umma_desc_k_sw128(smem_u32(smE + ss * ETileBytes));
            const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smX + ss * XTileBytes));
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
        mbar_wait(acc_done, 0);
        tc_fence_after();
        const int v = n_tile * 128 + threadIdx.x;
        const uint32_t lane_field = (uint32_t) (warp * 32) << 16;
#pragma unroll 1
        for (int c16 = 0; c16 < MT / 16; ++c16)
        {
            uint32_t acc[16];
            tc_ld_x16(tmem_base + lane_field + c16 * 16, acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i)
            {
                const int m = m_tile * MT + c16 * 16 + i;
                if (m < p.M && v < p.vocab)
                    p.out[(size_t) m * p.vocab + v] = __uint_as_float(acc[i]);
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

// =====================================================================================================
// Large-M path (encoder, M in the thousands): PERSISTENT tcgen05 GEMM with two TMEM accumulators.
//
// The kernel above dequantizes its weight tile once per m-tile (94 times per weight tile at M = 24000) with the warps
// that also run the epilogue, so epilogue and main loop of consecutive tiles cannot overlap (M = 24000: 1.1-1.25 PFLOP/s
// plain, 0.6-0.85 with a residual / GELU epilogue; tensor pipe 40-55 % active).  Here the int8 weights are expanded ONCE
// per call to exact fp16 integers ([N][K], K-major; 6.5-13 MB, about 1-3 % of the GEMM's time at M = 24000) and the GEMM
// is a plain fp16 x fp16 SS-UMMA pipeline:
//   D[m, n] = sum_k X[m, k] * W16[n, k]     UMMA M = 128 activation rows, UMMA N = 256 weight columns, fp32 in TMEM
//  * one CTA per SM walks tiles t = blockIdx.x, + gridDim.x, ... (n fastest: the CTAs running together share a few
//    128-row activation panels in L2; the fp16 weights stay L2-resident);
//  * warp 0 (one lane): TMA producer, 4 stages of (16 KB activations + 32 KB weights), 128B swizzle;
//  * warp 1 (one lane): tcgen05.mma issuer; the 512 TMEM columns hold TWO 128 x 256 accumulators, so the MMAs of tile
//    i + 1 run while
//  * warps 2-9 drain tile i: tcgen05.ld 16 columns at a time, column scale in fp32, bias / GELU / residual with the
//    same rounding sequence as finish_output_tile, 32-byte vector stores (a thread owns a row: 16 consecutive columns).
// =====================================================================================================
struct LgParams
{
    const __half* scales;
    const __half* bias;
    const __half* residual;
    __half* C;
    int M, N, ldc, kb_total, m_tiles, n_tiles;
    int m_pairs; // multicast variant: pairs of m-tiles (the two CTAs of a cluster take m-tiles 2 i and 2 i + 1 of one n-tile)
};

// 2-D tiled TMA load delivered to the same shared-memory offset (and signalled on the same barrier offset) of every CTA
// in cta_mask of the cluster
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar, uint16_t cta_mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
// tcgen05.commit arriving on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

constexpr int kLgBM = 128, kLgBN = 256, kLgStages = 4, kLgThreads = 320; // producer, MMA issuer, 8 epilogue warps
constexpr int kLgATile = kLgBM * 128, kLgBTile = kLgBN * 128; // bytes per stage (64 halves = 128 B per row)

// int8 (reference layout, [N/2][2K] bytes) -> fp16 exact integers [N][K]; one thread per 16-byte chunk = 16 k values
__global__ void __launch_bounds__(256) woq_expand_fp16_kernel(const uint8_t* __restrict__ W, __half* __restrict__ W16, int N, int K)
{
    grid_dep_launch_dependents();
    const int cpr = K / 16; // chunks per column
    const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long) N * cpr)
        return;
    const int n = (int) (idx / cpr), c = (int) (idx - (long long) n * cpr);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(W + (size_t) (n >> 1) * 2 * K + (size_t) (c >> 2) * 128 + (n & 1) * 64 + (c & 3) * 16));
    const uint32_t words[4] = {v.x, v.y, v.z, v.w};
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int w = 0; w < 4; ++w)
    {
        __half2 l, h;
        dequant_word(words[w], l, h); // l = k (2w, 2w+1), h = k (8+2w, 9+2w) of the chunk
        lo[w] = *reinterpret_cast<uint32_t*>(&l);
        hi[w] = *reinterpret_cast<uint32_t*>(&h);
    }
    grid_dep_wait(); // the scratch buffer may still be read by the GEMM two launches back
    uint4* dst = reinterpret_cast<uint4*>(W16 + (size_t) n * K + (size_t) c * 16);
    dst[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    dst[1] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

// Epilogue of one 128 x 256 tile for one thread (= output row m): this warp's half of the columns (chalf) from the TMEM
// accumulator at tmem_acc (lane field and accumulator column offset included), scale / bias from shared memory.
template <int ACT>
__device__ __forceinline__ void lg_drain_tile(const LgParams& p, uint32_t tmem_acc, const float* sc, const __half* bs, int m, int n0,
    int chalf)
{
    const bool has_bias = p.bias != nullptr, has_res = p.residual != nullptr;
    const bool row_ok = m < p.M;
    const size_t rbase = (size_t) m * p.ldc + n0;
    // 32 columns per iteration; the residual of the NEXT 32 columns is requested before this iteration's TMEM
    // read (a dependent global load per 16 columns made the epilogue, not the main loop, the limit of the K = 1280
    // GEMMs with a residual: 118 us against 69 us for the plain product at M = 24000)
    const bool res_row = has_res && row_ok;
    uint4 rn[4];
    auto fetch_res = [&](int c32)
    {
        const bool ok = res_row && n0 + c32 * 32 < p.N; // N is a multiple of 64: a 32-column group is in or out as a whole
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + rbase + c32 * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            rn[q] = ok ? rp[q] : make_uint4(0u, 0u, 0u, 0u);
    };
    const int c32_0 = chalf * (kLgBN / 64), c32_1 = c32_0 + kLgBN / 64;
    fetch_res(c32_0);
#pragma unroll 1
    for (int c32 = c32_0; c32 < c32_1; ++c32)
    {
        uint4 rc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            rc[q] = rn[q];
        if (c32 + 1 < c32_1)
            fetch_res(c32 + 1);
        const bool cols_ok = n0 + c32 * 32 < p.N;
        uint32_t acc[32];
        tc_ld_x32(tmem_acc + c32 * 32, acc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const __half* rh = reinterpret_cast<const __half*>(rc);
        __align__(16) __half o[32];
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
            const float4 s4 = *reinterpret_cast<const float4*>(sc + c32 * 32 + q * 4);
            const uint2 b4 = *reinterpret_cast<const uint2*>(bs + c32 * 32 + q * 4);
            const __half* bh = reinterpret_cast<const __half*>(&b4);
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                o[q * 4 + i] = finish_output_tile<ACT>(__uint_as_float(acc[q * 4 + i]) * sv[i], has_bias, bh[i], has_res,
                    rh[q * 4 + i]);
        }
        if (row_ok && cols_ok)
        {
            uint4* cp = reinterpret_cast<uint4*>(p.C + rbase + c32 * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                cp[q] = *reinterpret_cast<const uint4*>(&o[q * 8]);
        }
    }
}

// MC: clusters of two CTAs work on two m-tiles of the same n-tile; each CTA fetches half of the 256-column weight tile
// and multicasts it to both (L2 -> SM traffic per k-block 32 KB instead of 48 KB per CTA: the kernel without it sits at
// the L2 feed rate, 148 SMs x 48 KB per 2.1 M MACs).  A stage is refilled only after BOTH CTAs' MMAs have released it.
template <int ACT, bool MC>
__global__ void __launch_bounds__(kLgThreads, 1)
    woq_gemm_large_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const LgParams p)
{
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t) (kLgBN >> 3) << 17) | ((uint32_t) (kLgBM >> 4) << 24);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smA = smem;
    uint8_t* smB = smem + kLgStages * kLgATile;
    uint64_t* full = reinterpret_cast<uint64_t*>(smB + kLgStages * kLgBTile);
    uint64_t* empty = full + kLgStages;
    uint64_t* acc_full = empty + kLgStages; // [2] MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;     // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* sm_scale = reinterpret_cast<float*>(tmem_slot + 4); // [2][256]
    __half* sm_bias = reinterpret_cast<__half*>(sm_scale + 2 * kLgBN); // [2][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.kb_total;
    const uint32_t crank = MC ? cluster_ctarank() : 0u;
    // work items: tiles (plain) or pairs of m-tiles (multicast); every role walks the same sequence
    const int t_first = MC ? (int) (blockIdx.x >> 1) : (int) blockIdx.x;
    const int t_stride = MC ? (int) (gridDim.x >> 1) : (int) gridDim.x;
    const int num_tiles = (MC ? p.m_pairs : p.m_tiles) * p.n_tiles;
    auto tile_of = [&](int t, int& m_tile, int& n_tile)
    {
        const int mq = t / p.n_tiles;
        n_tile = t - mq * p.n_tiles;
        m_tile = MC ? 2 * mq + (int) crank : mq;
    };

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < kLgStages; ++s)
        {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MC ? 2 : 1);
        }
        for (int b = 0; b < 2; ++b)
        {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 8);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if constexpr (MC)
        cluster_sync_all(); // both CTAs' barriers exist before either multicasts into the other
    grid_dep_launch_dependents();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0)
    {
        if (elect_one_sync())
        {
            grid_dep_wait(); // activations AND the expanded weights come from earlier kernels on the stream
            int g = 0;
            for (int t = t_first; t < num_tiles; t += t_stride)
            {
                int m_tile, n_tile;
                tile_of(t, m_tile, n_tile);
                const int m_load = m_tile < p.m_tiles ? m_tile : p.m_tiles - 1; // odd tail: load valid rows, store nothing
                for (int kb = 0; kb < nkb; ++kb, ++g)
                {
                    const int s = g % kLgStages;
                    if (g >= kLgStages)
                        mbar_wait(&empty[s], ((g / kLgStages) - 1) & 1);
                    mbar_arrive_expect_tx(&full[s], kLgATile + kLgBTile);
                    tma_load_2d(smA + s * kLgATile, &tmX, kb * 64, m_load * kLgBM, &full[s]);
                    if constexpr (MC)
                        tma_load_2d_mc(smB + s * kLgBTile + crank * (kLgBTile / 2), &tmW, kb * 64,
                            n_tile * kLgBN + (int) crank * (kLgBN / 2), &full[s], (uint16_t) 3);
                    else
                        tma_load_2d(smB + s * kLgBTile, &tmW, kb * 64, n_tile * kLgBN, &full[s]);
                }
            }
        }
    }
    else if (warp == 1)
    {
        int g = 0, it = 0;
        for (int t = t_first; t < num_tiles; t += t_stride, ++it)
        {
            const int buf = it & 1, use = it >> 1;
            if (use > 0)
            {
                mbar_wait(&acc_empty[buf], (use - 1) & 1); // the epilogue has drained this accumulator
                tc_fence_after();
            }
            const uint32_t d_tmem = tmem_base + (uint32_t) buf * kLgBN;
            for (int kb = 0; kb < nkb; ++kb, ++g)
            {
                const int s = g % kLgStages;
                mbar_wait(&full[s], (g / kLgStages) & 1);
                tc_fence_after();
                const uint64_t adesc = umma_desc_k_sw128(smem_u32(smA + s * kLgATile));
                const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smB + s * kLgBTile));
                if (elect_one_sync())
                {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        tc_mma_ss(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, kIdesc, (kb | k4) != 0 ? 1u : 0u);
                    if constexpr (MC)
                        tc_commit_mc(&empty[s], (uint16_t) 3); // the stage is shared: both producers wait for both MMAs
                    else
                        tc_commit(&empty[s]);
                    if (kb == nkb - 1)
                        tc_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    }
    else
    {
        // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4, thread = one output row; warps 2-5 take the first 128
        // columns of the tile, warps 6-9 the other 128 (a GELU epilogue on four warps alone did not keep up with the main loop)
        const int quarter = warp & 3;
        const int et = (int) threadIdx.x - 64; // 0..255
        const int chalf = et >> 7;
        const uint32_t lane_field = (uint32_t) (quarter * 32) << 16;
        const bool has_bias = p.bias != nullptr, has_res = p.residual != nullptr;
        int it = 0;
        for (int t = t_first; t < num_tiles; t += t_stride, ++it)
        {
            const int buf = it & 1, use = it >> 1;
            int m_tile, n_tile;
            tile_of(t, m_tile, n_tile);
            const int n0 = n_tile * kLgBN;
            // column vectors of this tile -> shared memory (this buffer's previous reader finished two tiles ago)
            {
                const int c = et;
                const int n = n0 + c;
                sm_scale[buf * kLgBN + c] = n < p.N ? __half2float(__ldg(p.scales + n)) : 0.f;
                sm_bias[buf * kLgBN + c] = (has_bias && n < p.N) ? __ldg(p.bias + n) : __float2half(0.f);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
            lg_drain_tile<ACT>(p, tmem_base + lane_field + (uint32_t) buf * kLgBN, sm_scale + buf * kLgBN, sm_bias + buf * kLgBN,
                m_tile * kLgBM + quarter * 32 + lane, n0, chalf);
            // accumulator drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (MC)
        cluster_sync_all(); // no CTA leaves while its mate may still multicast into it or arrive on its barriers
    if (warp == 1)
    {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- CTA-pair variant (tcgen05 cta_group::2) -------------------------------------------------------------------------
// The two CTAs of a cluster compute ONE 256 x 256 tile: CTA r stages activation rows [128 r, 128 r + 128) and weight
// columns [128 r, 128 r + 128) of every k-block (32 KB per stage instead of 48: SIX stages in flight), the leader (rank 0)
// issues tcgen05.mma.cta_group::2 (UMMA M = 256, N = 256), which reads both CTAs' shared memory and writes 128 rows x 256
// columns into each CTA's TMEM.  Barriers: the TMA loads of both CTAs signal the LEADER's `full` barrier (2 producer arrivals
// + 64 KB of transactions per stage); the leader's tcgen05.commit multicasts its arrival to both CTAs' `empty` and `acc_full`
// barriers; the epilogue warps of both CTAs arrive on the LEADER's `acc_empty` (16 arrivals).
constexpr int kL2Stages = 6, kL2BTile = 128 * 128;

__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes)
{
    // (no .release.cluster: that form compiles to MEMBAR.ALL.GPU + ERRBAR per stage and starved the pipeline -- tensor pipe
    // 33 % active; arming a barrier orders nothing)
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data to THIS CTA's shared memory, completion on the barrier at cluster address mbar_cluster
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, int c0, int c1, uint32_t mbar_cluster)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(mbar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_mma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

template <int ACT>
__global__ void __launch_bounds__(kLgThreads, 1)
    woq_gemm_large2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const LgParams p)
{
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t) (kLgBN >> 3) << 17) | ((uint32_t) ((2 * kLgBM) >> 4) << 24);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smA = smem;
    uint8_t* smB = smem + kL2Stages * kLgATile;
    uint64_t* full = reinterpret_cast<uint64_t*>(smB + kL2Stages * kL2BTile);
    uint64_t* empty = full + kL2Stages;
    uint64_t* acc_full = empty + kL2Stages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* sm_scale = reinterpret_cast<float*>(tmem_slot + 4);          // [2][256]
    __half* sm_bias = reinterpret_cast<__half*>(sm_scale + 2 * kLgBN);  // [2][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.kb_total;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;
    const int t_first = (int) (blockIdx.x >> 1), t_stride = (int) (gridDim.x >> 1);
    const int num_tiles = p.m_pairs * p.n_tiles;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < kL2Stages; ++s)
        {
            mbar_init(&full[s], 2);  // (leader's copy is the one in use) one arrive.expect_tx per CTA of the pair
            mbar_init(&empty[s], 1); // the leader's commit, multicast to both CTAs
        }
        for (int b = 0; b < 2; ++b)
        {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 16); // (leader's copy) the eight epilogue warps of both CTAs
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1)
    {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all(); // both CTAs' barriers and TMEM exist before any cross-CTA traffic
    grid_dep_launch_dependents();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0)
    {
        if (elect_one_sync())
        {
            grid_dep_wait();
            int g = 0;
            for (int t = t_first; t < num_tiles; t += t_stride)
            {
                const int mq = t / p.n_tiles, n_tile = t - mq * p.n_tiles;
                const int m_tile = 2 * mq + (int) crank;
                const int m_load = m_tile < p.m_tiles ? m_tile : p.m_tiles - 1;
                for (int kb = 0; kb < nkb; ++kb, ++g)
                {
                    const int s = g % kL2Stages;
                    if (g >= kL2Stages)
                        mbar_wait(&empty[s], ((g / kL2Stages) - 1) & 1);
                    const uint32_t lfull = mapa_u32(smem_u32(&full[s]), 0u); // the leader's barrier
                    mbar_arrive_expect_tx_cluster(lfull, kLgATile + kL2BTile);
                    tma_load_2d_pair(smA + s * kLgATile, &tmX, kb * 64, m_load * kLgBM, lfull);
                    tma_load_2d_pair(smB + s * kL2BTile, &tmW, kb * 64, n_tile * kLgBN + (int) crank * (kLgBN / 2), lfull);
                }
            }
        }
    }
    else if (warp == 1)
    {
        // one lane runs the whole issue loop (no election, reconvergence or descriptor set-up per k-block: the loop has to
        // stay well under the 512 cycles of tensor work a k-block holds)
        if (leader && elect_one_sync())
        {
            const uint64_t adesc0 = umma_desc_k_sw128(smem_u32(smA)), bdesc0 = umma_desc_k_sw128(smem_u32(smB));
            int g = 0, it = 0, s = 0;
            uint32_t ph = 0;
            for (int t = t_first; t < num_tiles; t += t_stride, ++it)
            {
                const int buf = it & 1, use = it >> 1;
                if (use > 0)
                {
                    mbar_wait(&acc_empty[buf], (use - 1) & 1); // both CTAs' epilogues have drained this accumulator
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + (uint32_t) buf * kLgBN;
                for (int kb = 0; kb < nkb; ++kb, ++g)
                {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    // the descriptors' 14-bit start-address field counts 16-byte units: stage s is s tiles further
                    const uint64_t adesc = adesc0 + (uint64_t) (s * (kLgATile >> 4));
                    const uint64_t bdesc = bdesc0 + (uint64_t) (s * (kL2BTile >> 4));
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        tc_mma_ss_pair(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, kIdesc, (kb | k4) != 0 ? 1u : 0u);
                    tc_commit_pair(&empty[s], (uint16_t) 3);
                    if (kb == nkb - 1)
                        tc_commit_pair(&acc_full[buf], (uint16_t) 3);
                    if (++s == kL2Stages)
                    {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    }
    else
    {
        const int quarter = warp & 3;
        const int et = (int) threadIdx.x - 64; // 0..255
        const int chalf = et >> 7;
        const uint32_t lane_field = (uint32_t) (quarter * 32) << 16;
        const bool has_bias = p.bias != nullptr;
        int it = 0;
        for (int t = t_first; t < num_tiles; t += t_stride, ++it)
        {
            const int buf = it & 1, use = it >> 1;
            const int mq = t / p.n_tiles, n_tile = t - mq * p.n_tiles;
            const int m_tile = 2 * mq + (int) crank;
            const int n0 = n_tile * kLgBN;
            {
                const int n = n0 + et;
                sm_scale[buf * kLgBN + et] = n < p.N ? __half2float(__ldg(p.scales + n)) : 0.f;
                sm_bias[buf * kLgBN + et] = (has_bias && n < p.N) ? __ldg(p.bias + n) : __float2half(0.f);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
            lg_drain_tile<ACT>(p, tmem_base + lane_field + (uint32_t) buf * kLgBN, sm_scale + buf * kLgBN, sm_bias + buf * kLgBN,
                m_tile * kLgBM + quarter * 32 + lane, n0, chalf);
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0u));
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1)
    {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once,
        []()
        {
            void* sym = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess
                && qres == cudaDriverEntryPointSuccess)
                fn = reinterpret_cast<PFN_encodeTiled>(sym);
        });
    return fn;
}

// The driver entry point needs a CURRENT context.  A host thread that has made no runtime call yet (a second TensorRT
// execution context enqueueing from its own thread) has none: bind the device's primary context once per thread with a
// no-op runtime call.  (Found by tests/test_plugin_gpu.py::test_two_threads_enqueue_concurrently: CUresult 201.)
static void bind_context_once()
{
    static thread_local bool bound = false;
    if (!bound)
    {
        cudaFree(nullptr);
        bound = true;
    }
}

int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dt, const void* base, uint64_t dim0, uint64_t dim1,
    uint64_t stride1_bytes, uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz)
{
    bind_context_once();
    PFN_encodeTiled enc = get_encode();
    B200_REQUIRE(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int) r);
    return B200_OK;
}

// 3-D variant with a traversal (element) stride on dimension 1: box1 counts tensor elements, so box1 / estride1 rows land
// in shared memory (used by the strided Conv1d: every second time step).
int make_tmap_3d(CUtensorMap* out, CUtensorMapDataType dt, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
    uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2, uint32_t estride1,
    CUtensorMapSwizzle swz)
{
    bind_context_once();
    PFN_encodeTiled enc = get_encode();
    B200_REQUIRE(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[3] = {dim0, dim1, dim2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, box2};
    cuuint32_t estr[3] = {1, estride1, 1};
    CUresult r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int) r);
    return B200_OK;
}

struct TcPlan
{
    int MT, m_tiles, n_tiles, splits;
    int cluster; // 1: splits CTAs per tile form a cluster (DSMEM reduction); 0: global fp32 slabs
    size_t slab_bytes;
};

static long long* g_tc_gt = nullptr; // per-launch global-timer records [launch][4]
static int g_tc_gt_cap = 0, g_tc_gt_seq = 0;
void tc_set_timeline_buffer(long long* p, int max_launches)
{
    g_tc_gt = p;
    g_tc_gt_cap = max_launches;
    g_tc_gt_seq = 0;
}
static long long* g_tc_dbg = nullptr;
static int g_tc_dbg_n = 0, g_tc_dbg_fold = 0; // optional filter: only launches with this N (and folded-LN flag) stamp
void tc_set_debug_filter(int n, int fold)
{
    g_tc_dbg_n = n;
    g_tc_dbg_fold = fold;
}
void tc_set_debug_buffer(long long* p)
{
    g_tc_dbg = p;
}

static int g_splitk_mode = -1; // -1 auto (env B200_SPLITK: "cluster" | "global"), 0 global slabs, 1 cluster

static int g_cluster16 = -1; // env B200_CLUSTER16=1 enables 16-CTA (non-portable) clusters for deep-K decode GEMMs

#ifndef B200_TC_SMALL_GRID_128
#define B200_TC_SMALL_GRID_128 1
#endif

TcPlan plan_tc(int M, int N, int K)
{
    if (g_cluster16 < 0)
    {
        const char* e = getenv("B200_CLUSTER16");
        g_cluster16 = (e != nullptr && e[0] == '1') ? 1 : 0; // opt-in: measured neutral on the decoder step
    }
    if (g_splitk_mode < 0)
    {
        const char* e = getenv("B200_SPLITK");
        g_splitk_mode = (e != nullptr && e[0] == 'g') ? 0 : 1;
    }
    TcPlan pl{};
    pl.MT = M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256;
    // 256-row tiles are never split, so a problem with fewer of them than SMs leaves the GPU under-filled (M = 256,
    // 5120 -> 1280: 10 CTAs walking all of K, 38.8 us against 11.1 us at M = 128): take 128-row tiles there, which
    // doubles the tile count and keeps the cluster split-K available
    if (pl.MT == 256 && B200_TC_SMALL_GRID_128 && (long long) ((M + 255) / 256) * ((N + 127) / 128) < num_sms())
        pl.MT = 128;
    pl.m_tiles = (M + pl.MT - 1) / pl.MT;
    pl.n_tiles = (N + 127) / 128;
    const int kb_total = K / 64;
    const int tiles = pl.m_tiles * pl.n_tiles;
    const int sms = num_sms();
    int splits = 1;
    int cluster = 0;
    if (tiles < sms && tiles <= 4096)
    {
        if (g_splitk_mode == 1 && pl.MT <= 128)
        {
            // cluster split-K: power-of-two cluster (<= 8, portable) along z, at least two k-blocks per CTA.
            // Decode tiles (MT <= 32) fit two CTAs per SM, so a grid may exceed the SM count a little; it must still
            // leave most second slots free for the NEXT kernel's early (programmatic) launch.
            const int cap = pl.MT <= 32 ? sms + sms / 8 : sms;
            int s2 = 8;
            while (s2 > 1 && (tiles * s2 > cap || kb_total < 2 * s2))
                s2 >>= 1;
            // deep-K decode GEMMs (fc2): an 8-way split leaves more k-blocks per CTA than the TMEM ring holds, so part
            // of the dequant would run after the dependency resolves; a 16-CTA (non-portable) cluster halves that
            if (g_cluster16 && pl.MT <= 32 && s2 == 8 && kb_total > 8 * 6 && tiles * 16 <= cap && kb_total >= 32)
                s2 = 16;
            // 128-row tiles with a shallow K (1280: 20 k-blocks): the 64 KB DSMEM exchange of a split costs more than
            // the halved main loop saves (1280 -> 5120 at M = 128: 13.4 us split 2-way, slower than the 9.4 us of M = 256
            // whose 80 tiles run unsplit); keep the split for deep K only
            static int min_kb128 = -1;
            if (min_kb128 < 0)
            {
                const char* e = getenv("B200_TC_SPLIT128_MIN_KB");
                min_kb128 = e != nullptr ? atoi(e) : 40;
            }
            if (pl.MT == 128 && kb_total < min_kb128)
                s2 = 1;
            splits = s2;
            cluster = s2 > 1 ? 1 : 0;
        }
        else
        {
            splits = sms / tiles;
            // keep at least 2 k-blocks per split so the per-CTA fixed cost is amortised
            if (splits > kb_total / 2)
                splits = kb_total / 2;
            if (splits < 1)
                splits = 1;
            // 256-row tiles: the fp32 slab round trip of a 128 KB tile costs more than the split saves (measured:
            // 53 us split 4-way vs ~12 us unsplit at M = 256, 1280 -> 3840)
            if (pl.MT == 256 && g_splitk_mode == 1)
                splits = 1;
        }
    }
    pl.splits = splits;
    pl.cluster = cluster;
    pl.slab_bytes = (splits > 1 && !cluster) ? (size_t) splits * tiles * 128 * pl.MT * sizeof(float) : 0;
    return pl;
}

static int* g_counters = nullptr;
static std::mutex g_counter_mu;
// Counter / scratch slots are handed out PER STREAM: a slot is only safe to reuse once the launch that used it has
// finished, and stream order is the one ordering the library can rely on (launch N + kSlotsPerStream on a stream runs
// after launch N even under programmatic dependent launch, whose overlap reaches one or two kernels back).  Every
// stream the library sees (including a capture stream: the addresses are baked into the captured graph, and the
// graph's nodes keep the stream's order) owns a private range of the pool, so kernels of concurrent streams -- forked
// decoder chains, an encoder overlapping a decoder graph replay, two plugin contexts enqueueing from two threads --
// never share a slot.  More than kSlotStreams distinct streams recycle the least recently used range.
constexpr int kSlotStreams = 16, kSlotsPerStream = 16, kCounterSlots = kSlotStreams * kSlotsPerStream, kCounterSlotInts = 16384;

struct SlotRange
{
    cudaStream_t stream;
    unsigned next;
    unsigned long long last_use;
    bool used;
};

static SlotRange g_slot_ranges[kSlotStreams];
static unsigned long long g_slot_clock = 0;

int tc_init()
{
    std::lock_guard<std::mutex> lk(g_counter_mu);
    if (g_counters == nullptr)
    {
        B200_CUDA(cudaMalloc(&g_counters, sizeof(int) * kCounterSlots * kCounterSlotInts));
        B200_CUDA(cudaMemset(g_counters, 0, sizeof(int) * kCounterSlots * kCounterSlotInts));
    }
    return B200_OK;
}

// Hands out one self-resetting counter slot (kCounterSlotInts ints, all zero between launches) from `stream`'s range.
int* tc_counter_slot(int needed, cudaStream_t stream)
{
    if (needed > kCounterSlotInts)
        return nullptr;
    if (g_counters == nullptr && tc_init() != B200_OK)
        return nullptr;
    std::lock_guard<std::mutex> lk(g_counter_mu);
    int idx = -1;
    for (int i = 0; i < kSlotStreams && idx < 0; ++i)
        if (g_slot_ranges[i].used && g_slot_ranges[i].stream == stream)
            idx = i;
    if (idx < 0)
    {
        for (int i = 0; i < kSlotStreams && idx < 0; ++i)
            if (!g_slot_ranges[i].used)
                idx = i;
        if (idx < 0)
        {
            idx = 0;
            for (int i = 1; i < kSlotStreams; ++i)
                if (g_slot_ranges[i].last_use < g_slot_ranges[idx].last_use)
                    idx = i;
        }
        g_slot_ranges[idx] = SlotRange{stream, 0u, 0ull, true};
    }
    SlotRange& r = g_slot_ranges[idx];
    r.last_use = ++g_slot_clock;
    return g_counters + ((size_t) idx * kSlotsPerStream + (r.next++ % kSlotsPerStream)) * kCounterSlotInts;
}

template <int MT, int SS, int AS, bool CL>
static int launch_tc_impl(const CUtensorMap& tmW, const CUtensorMap& tmX, const TcParams& p, dim3 grid, cudaStream_t stream)
{
    auto kern = woq_gemm_tc_kernel<MT, SS, AS, CL>;
    const size_t smem = TcSmem<MT, SS, AS>::total(p.cluster != 0, p.fold_gamma != nullptr, p.K);
    B200_REQUIRE(smem <= 227 * 1024, B200_ERR_UNSUPPORTED, "woq gemm: %zu bytes of shared memory needed", smem);
    static size_t attr_smem = 0;
    if (smem > attr_smem)
    {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        // decode tiles are sized so that two CTAs (this GEMM's and the next one's, launched early) share an SM
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        if (g_cluster16 > 0)
            B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_smem = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kTcThreads);
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
    if (p.cluster)
    {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = (unsigned) p.splits;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    count_launch();
    B200_CUDA(cudaLaunchKernelEx(&cfg, kern, tmW, tmX, p));
    return B200_OK;
}

template <int MT, int SS, int AS>
static int launch_tc(const CUtensorMap& tmW, const CUtensorMap& tmX, const TcParams& p, dim3 grid, cudaStream_t stream)
{
    // cluster split-K (every decode GEMM) runs the instantiation that carries only that epilogue
    if constexpr (MT <= 128)
    {
        if (p.cluster)
            return launch_tc_impl<MT, SS, AS, true>(tmW, tmX, p, grid, stream);
    }
    return launch_tc_impl<MT, SS, AS, false>(tmW, tmX, p, grid, stream);
}

// Can LayerNorm be folded into the GEMM for this shape?  (decode-sized m-tiles, cluster or unsplit reduction, the
// whole k range of a CTA resident in the activation ring)
bool woq_tc_can_fold_ln(int M, int N, int K)
{
    if (M > 32)
        return false;
    const TcPlan pl = plan_tc(M, N, K);
    if (!(pl.cluster || pl.splits == 1) || pl.splits > 8)
        return false; // the statistics travel through the cluster inbox (8 sender slots)
    const int nkb = (K / 64 + pl.splits - 1) / pl.splits;
    return nkb <= (pl.MT == 16 ? 10 : 7); // the activation stages must not be recycled before they are summed
}

// tcgen05 path entry: any M >= 1.  fold_gamma != nullptr: A is the raw residual stream and the LayerNorm is folded
// into the kernel (see TcParams).
bool woq_large_applies(int M, int N, int K, size_t workspace_bytes);
int woq_gemm_large(const __half* A, int M, int K, const uint8_t* W, const __half* scales, int N, const __half* bias,
    int activation, const __half* residual, __half* C, void* workspace, cudaStream_t stream);

int woq_gemm_tc(const __half* A, int M, int K, const uint8_t* W, const __half* scales, int N, const __half* bias,
    int activation, const __half* residual, __half* C, void* workspace, size_t workspace_bytes, cudaStream_t stream,
    const __half* fold_gamma, const float* fold_c1s, const float* fold_c2, float ln_eps)
{
    // (a workspace, output or residual that is not aligned for the vector / TMA accesses of the large-M kernel simply keeps
    // the per-tile kernel)
    if (fold_gamma == nullptr && workspace != nullptr && woq_large_applies(M, N, K, workspace_bytes)
        && (reinterpret_cast<uintptr_t>(workspace) & 127) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0
        && (residual == nullptr || (reinterpret_cast<uintptr_t>(residual) & 15) == 0))
        return woq_gemm_large(A, M, K, W, scales, N, bias, activation, residual, C, workspace, stream);
    const TcPlan pl = plan_tc(M, N, K);
    B200_REQUIRE(pl.slab_bytes == 0 || (workspace != nullptr && workspace_bytes >= pl.slab_bytes), B200_ERR_WORKSPACE,
        "woq gemm: workspace of %zu bytes needed for split-K, got %zu", pl.slab_bytes, workspace_bytes);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
        B200_ERR_INVALID_ARG, "woq gemm: A and W must be 16-byte aligned for TMA");
    if (g_counters == nullptr)
    {
        if (int rc = tc_init())
            return rc;
    }
    CUtensorMap tmW, tmX;
    if (int rc = make_tmap_2d(&tmW, CU_TENSOR_MAP_DATA_TYPE_UINT8, W, (uint64_t) 2 * K, (uint64_t) N / 2,
            (uint64_t) 2 * K, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    // decode-sized launches whose per-CTA k range fits the ring fetch their activations with ONE 3-D TMA box
    const int nkb_max = (K / 64 + pl.splits - 1) / pl.splits;
    const int ring = pl.MT == 16 ? 10 : pl.MT == 32 ? 7 : pl.MT == 64 ? 6 : 4;
    static const bool x3_enabled = getenv("B200_X3") == nullptr || getenv("B200_X3")[0] != '0';
    const int x3_depth = (x3_enabled && pl.MT <= 32 && pl.cluster && nkb_max <= ring) ? nkb_max : 0;
    if (x3_depth > 0)
    {
        if (int rc = make_tmap_3d(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, A, 64, (uint64_t) M, (uint64_t) K / 64,
                (uint64_t) K * 2, 128, 64, (uint32_t) pl.MT, (uint32_t) x3_depth, 1, CU_TENSOR_MAP_SWIZZLE_128B))
            return rc;
    }
    else if (int rc = make_tmap_2d(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, A, (uint64_t) K, (uint64_t) M, (uint64_t) K * 2, 64,
                 (uint32_t) pl.MT, CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    TcParams p{};
    p.x3_depth = x3_depth;
    static const bool burst_on = getenv("B200_TC_BURST") == nullptr || getenv("B200_TC_BURST")[0] != '0';
    p.mma_burst = burst_on ? 1 : 0;
    p.scales = scales;
    p.bias = bias;
    p.residual = residual;
    p.C = C;
    p.slabs = static_cast<float*>(workspace);
    p.counters = tc_counter_slot(pl.m_tiles * pl.n_tiles <= kCounterSlotInts ? pl.m_tiles * pl.n_tiles : 1, stream);
    if (fold_gamma != nullptr)
    {
        p.fold_x = A;
        p.fold_gamma = fold_gamma;
        p.fold_c1s = fold_c1s;
        p.fold_c2 = fold_c2;
        p.ln_eps = ln_eps;
    }
    p.M = M;
    p.N = N;
    p.K = K;
    p.ldc = N;
    p.activation = activation;
    p.kb_total = K / 64;
    p.splits = pl.splits;
    p.cluster = pl.cluster;
    p.gt = (g_tc_gt != nullptr && g_tc_gt_seq < g_tc_gt_cap) ? g_tc_gt + 4 * (size_t) (g_tc_gt_seq++) : nullptr;
    p.dbg = (g_tc_dbg_n == 0 || (g_tc_dbg_n == N && g_tc_dbg_fold == (fold_gamma != nullptr ? 1 : 0))) ? g_tc_dbg : nullptr;
    dim3 grid(pl.n_tiles, pl.m_tiles, pl.splits);
    switch (pl.MT)
    {
    case 16: return launch_tc<16, 10, 7>(tmW, tmX, p, grid, stream); // 7 x 32 A columns + 16 accumulator columns = 240 of the 256 a CTA may hold
    case 32: return launch_tc<32, 7, 7>(tmW, tmX, p, grid, stream); // 7 x 32 + 32 = all 256 columns
    case 64: return launch_tc<64, 6, 6>(tmW, tmX, p, grid, stream);
    case 128: return launch_tc<128, 4, 4>(tmW, tmX, p, grid, stream);
    default: return launch_tc<256, 4, 4>(tmW, tmX, p, grid, stream);
    }
}

// ---- large-M path: host side ----
static int g_large_m = -1; // rows from which the persistent large-M kernel is used (env B200_LARGE_M; 0 = never)
static int large_m_threshold()
{
    if (g_large_m < 0)
    {
        const char* e = getenv("B200_LARGE_M");
        g_large_m = e != nullptr ? atoi(e) : 4096;
    }
    return g_large_m;
}

bool woq_large_applies(int M, int N, int K, size_t workspace_bytes)
{
    const int thr = large_m_threshold();
    return thr > 0 && M >= thr && N % 64 == 0 && K % 64 == 0 && workspace_bytes >= (size_t) N * K * sizeof(__half);
}

template <int ACT, bool MC>
static int launch_large(const CUtensorMap& tmX, const CUtensorMap& tmW, const LgParams& p, int grid, cudaStream_t stream)
{
    const size_t smem = 1024 + (size_t) kLgStages * (kLgATile + kLgBTile) + 128 + 2 * kLgBN * (sizeof(float) + sizeof(__half));
    auto kern = woq_gemm_large_kernel<ACT, MC>;
    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    if constexpr (!MC)
    {
        B200_LAUNCH(kern, dim3(grid), dim3(kLgThreads), smem, stream, tmX, tmW, p);
        return B200_OK;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kLgThreads);
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
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    count_launch();
    B200_CUDA(cudaLaunchKernelEx(&cfg, kern, tmX, tmW, p));
    return B200_OK;
}

template <int ACT>
static int launch_large2(const CUtensorMap& tmX, const CUtensorMap& tmW, const LgParams& p, int grid, cudaStream_t stream)
{
    const size_t smem = 1024 + (size_t) kL2Stages * (kLgATile + kL2BTile) + 256 + 2 * kLgBN * (sizeof(float) + sizeof(__half));
    auto kern = woq_gemm_large2_kernel<ACT>;
    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kLgThreads);
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
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    count_launch();
    B200_CUDA(cudaLaunchKernelEx(&cfg, kern, tmX, tmW, p));
    return B200_OK;
}

static int launch_large2_act(int activation, const CUtensorMap& tmX, const CUtensorMap& tmW, const LgParams& p, int grid,
    cudaStream_t stream)
{
    switch (activation)
    {
    case B200_ACT_GELU_ERF: return launch_large2<B200_ACT_GELU_ERF>(tmX, tmW, p, grid, stream);
    case B200_ACT_GELU_TANH: return launch_large2<B200_ACT_GELU_TANH>(tmX, tmW, p, grid, stream);
    default: return launch_large2<B200_ACT_NONE>(tmX, tmW, p, grid, stream);
    }
}

template <bool MC>
static int launch_large_act(int activation, const CUtensorMap& tmX, const CUtensorMap& tmW, const LgParams& p, int grid,
    cudaStream_t stream)
{
    switch (activation)
    {
    case B200_ACT_GELU_ERF: return launch_large<B200_ACT_GELU_ERF, MC>(tmX, tmW, p, grid, stream);
    case B200_ACT_GELU_TANH: return launch_large<B200_ACT_GELU_TANH, MC>(tmX, tmW, p, grid, stream);
    default: return launch_large<B200_ACT_NONE, MC>(tmX, tmW, p, grid, stream);
    }
}

int woq_gemm_large(const __half* A, int M, int K, const uint8_t* W, const __half* scales, int N, const __half* bias,
    int activation, const __half* residual, __half* C, void* workspace, cudaStream_t stream)
{
    __half* W16 = static_cast<__half*>(workspace);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W16) & 127) == 0
            && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (residual == nullptr || (reinterpret_cast<uintptr_t>(residual) & 15) == 0),
        B200_ERR_INVALID_ARG, "woq gemm (large M): A / C / residual must be 16-byte and the workspace 128-byte aligned");
    {
        const long long chunks = (long long) N * (K / 16);
        B200_LAUNCH(woq_expand_fp16_kernel, dim3((unsigned) ((chunks + 255) / 256)), dim3(256), 0, stream, W, W16, N, K);
    }
    CUtensorMap tmX, tmW;
    if (int rc = make_tmap_2d(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, A, (uint64_t) K, (uint64_t) M, (uint64_t) K * 2, 64, kLgBM,
            CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    // B200_LARGE_MC: 2 (default) = CTA-pair MMA (tcgen05 cta_group::2), 1 = two independent MMAs with the weight tile
    // multicast, 0 = every CTA fetches its whole weight tile.  In modes 1 and 2 a TMA box is half a weight tile.
    static const int mode = [] { const char* e = getenv("B200_LARGE_MC"); return e == nullptr ? 2 : atoi(e); }();
    const bool mc = mode != 0;
    if (int rc = make_tmap_2d(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, W16, (uint64_t) K, (uint64_t) N, (uint64_t) K * 2, 64,
            mc ? kLgBN / 2 : kLgBN, CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    LgParams p{};
    p.scales = scales;
    p.bias = bias;
    p.residual = residual;
    p.C = C;
    p.M = M;
    p.N = N;
    p.ldc = N;
    p.kb_total = K / 64;
    p.m_tiles = (M + kLgBM - 1) / kLgBM;
    p.n_tiles = (N + kLgBN - 1) / kLgBN;
    p.m_pairs = (p.m_tiles + 1) / 2;
    if (mc)
    {
        const int items = p.m_pairs * p.n_tiles, clusters = num_sms() / 2;
        const int grid = 2 * (items < clusters ? items : clusters);
        if (mode == 2)
            return launch_large2_act(activation, tmX, tmW, p, grid, stream);
        return launch_large_act<true>(activation, tmX, tmW, p, grid, stream);
    }
    const int tiles = p.m_tiles * p.n_tiles;
    return launch_large_act<false>(activation, tmX, tmW, p, tiles < num_sms() ? tiles : num_sms(), stream);
}

// Host-side launch plan of the tcgen05 path, exported for tests (no device work; without a GPU the SM count is 148).
void woq_tc_plan_query(int M, int N, int K, int* mt, int* m_tiles, int* n_tiles, int* splits, int* cluster)
{
    const TcPlan pl = plan_tc(M, N, K);
    *mt = pl.MT;
    *m_tiles = pl.m_tiles;
    *n_tiles = pl.n_tiles;
    *splits = pl.splits;
    *cluster = pl.cluster;
}

size_t woq_tc_workspace_bytes(int max_m, int N, int K)
{
    // plan_tc never uses more than num_sms() slabs of one 128 x MT fp32 tile (splits * tiles <= num_sms), and MT
    // grows with M, so the class of max_m bounds every smaller M.
    (void) N;
    (void) K;
    const int MT = max_m <= 16 ? 16 : max_m <= 32 ? 32 : max_m <= 64 ? 64 : max_m <= 128 ? 128 : 256;
    size_t bytes = (size_t) num_sms() * 128 * MT * sizeof(float);
    // large-M path: the weights expanded to fp16 live in the workspace for the duration of the call
    const int thr = large_m_threshold();
    if (thr > 0 && max_m >= thr && (size_t) N * K * sizeof(__half) > bytes)
        bytes = (size_t) N * K * sizeof(__half);
    return bytes;
}

template <int MT, int SS>
static int launch_logits(const CUtensorMap& tmE, const CUtensorMap& tmX, const LogitsParams& p, dim3 grid, cudaStream_t stream)
{
    auto kern = fp16_gemm_tc_kernel<MT, SS>;
    const size_t smem = 1024 + (size_t) SS * (128 * 128 + MT * 128) + sizeof(uint64_t) * (2 * SS + 1) + 16;
    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    B200_LAUNCH(kern, grid, dim3(192), smem, stream, tmE, tmX, p);
    return B200_OK;
}

int logits_tc(const __half* x, const __half* emb, float* logits, int rows, int cols, int vocab, cudaStream_t stream)
{
    B200_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(emb) & 15) == 0,
        B200_ERR_INVALID_ARG, "logits: x and emb must be 16-byte aligned for TMA");
    const int MT = rows <= 16 ? 16 : rows <= 32 ? 32 : rows <= 64 ? 64 : rows <= 128 ? 128 : 256;
    CUtensorMap tmE, tmX;
    if (int rc = make_tmap_2d(&tmE, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, emb, (uint64_t) cols, (uint64_t) vocab,
            (uint64_t) cols * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    if (int rc = make_tmap_2d(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, x, (uint64_t) cols, (uint64_t) rows,
            (uint64_t) cols * 2, 64, (uint32_t) MT, CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    LogitsParams p{};
    p.out = logits;
    p.M = rows;
    p.vocab = vocab;
    p.kb_total = cols / 64;
    dim3 grid((vocab + 127) / 128, (rows + MT - 1) / MT);
    switch (MT)
    {
    case 16: return launch_logits<16, 6>(tmE, tmX, p, grid, stream);
    case 32: return launch_logits<32, 6>(tmE, tmX, p, grid, stream);
    case 64: return launch_logits<64, 6>(tmE, tmX, p, grid, stream);
    case 128: return launch_logits<128, 5>(tmE, tmX, p, grid, stream);
    default: return launch_logits<256, 4>(tmE, tmX, p, grid, stream);
    }
}

} // namespace b200
