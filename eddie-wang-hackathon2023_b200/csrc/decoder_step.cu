// decoder_step.cu -- ONE persistent kernel for a whole generation step of the quantized Whisper decoder stack
// (b200_decoder_step, include/b200_whisper.h).
//
// Why: with one kernel per operator the batch-16 step is 8 dependent launches per layer x 32 layers, each costing
// 3-5 us of dependency latency for 0.25-1 us of HBM time (profiles/r01_step_timeline_v7.txt): the step ran at a third
// of the HBM roofline although every kernel moved exactly its algorithmic bytes.  The reference has the same shape
// (one TensorRT layer / plugin enqueue per operator: weightOnlyQuantMatmulPlugin.cpp:162-222,
// gptAttentionCommon.cpp:649-780).  Here the dependency chain stays -- it is the model -- but everything that does
// not depend on it is taken off it:
//
//   * one CTA per SM, resident for the whole step; warp 10 (one lane) is a TMA producer that walks a STATIC schedule
//     of 10 KB items -- int8 weight tiles (8 output columns x 1280 k in the reference's preprocessed layout = four
//     contiguous 2560-byte row-pair segments) and cross-KV chunks (80 keys: 5 KB of K + 5 KB of V) -- through a
//     19-slot shared-memory ring with cp.async.bulk + mbarrier complete_tx.  Weights and the cross-KV cache never
//     depend on activations, so the producer runs up to 190 KB per SM (28 MB per chip, more than a layer's weights)
//     ahead of the consumers: when a phase's activations arrive its weights are already in shared memory, and while
//     the consumers wait on a grid barrier HBM keeps streaming the next phases' bytes.
//   * warps 0-9 consume.  Matmuls (16 batch rows): every warp owns two 64-wide k-blocks of each tile, converts the
//     biased int8 bytes to fp16 in registers (PRMT + HSUB2, exact integers; the reference layout's row permutation
//     and byte swizzle make each converted word the B fragment of an mma.sync.m16n8k16), multiplies by the LayerNorm
//     gamma pair when the LayerNorm is folded in, and accumulates in fp32 with the batch rows as the MMA's M.  The
//     activations are kept in global memory in "A-fragment order", so the A operands are coalesced 128-bit loads
//     straight from L2 into registers -- no shared-memory staging of activations at all.  The ten k-partials are
//     reduced through shared memory in fixed warp order (deterministic); the epilogue applies the column scale, the
//     folded-LayerNorm correction (statistics from the A fragments), bias / GELU / residual with the same per-layer
//     fp16 rounding as the per-operator kernels (common.cuh epilogue_apply).
//   * self-attention: (batch, head) pairs dealt to CTAs, 3-10 warps per pair over alternating groups of 8 keys, first
//     pass of the int8 cache fetched before the grid barrier; same arithmetic as mmha_generation_kernel.
//   * cross-attention: whole (batch, head) pairs per CTA, chunks dealt round-robin to the ten warps straight from the
//     ring, merge through shared memory; same inner loop as cross_attention_rowhead_kernel (attn_device.cuh).
//   * phases are separated by a grid barrier: bar.sync, one red.release.gpu per CTA, one ld.acquire.gpu poller per CTA.
//     Every wait is bounded (a stuck barrier sets the status word and the kernel drains instead of hanging the GPU).
#include <float.h>
#include <stdlib.h>

#include "attn_device.cuh"
#include "common.cuh"

namespace b200
{

constexpr int kDsCW = 10;                 // consumer warps
constexpr int kDsConsumers = kDsCW * 32;  // 320 threads
constexpr int kDsThreads = kDsConsumers + 32; // + the ring producer warp
constexpr int kDsSlots = 19;
constexpr int kDsSlotBytes = 10240;
constexpr int kDsUnitK = 1280;            // k extent of a weight unit (8 columns x 1280 k = one slot)
constexpr int kDsRoundTiles = 5;          // tiles (8 columns each) per reduction round
constexpr int kDsChunkKeys = 80;          // keys per cross-KV chunk: 5120 B of K + 5120 B of V
constexpr int kDsScratchFloats = kDsCW * kDsRoundTiles * 16 * 8; // 25600 B: k-partials / attention merge area
constexpr int kDsStatFloats = kDsCW * 16 * 3;                    // per (warp, row): count, mean, M2
constexpr int kDsPart = kDh + 4;          // attention partial: m, l, 2 pad, o[64]
constexpr size_t kDsSmemBytes = (size_t) kDsSlots * kDsSlotBytes + sizeof(float) * (kDsScratchFloats + kDsStatFloats)
    + sizeof(uint64_t) * 2 * kDsSlots + 64 + 2 * 32 * sizeof(uint64_t);
constexpr long long kDsWaitCycles = 3000000000ll; // SM cycles before a wait gives up (~1.5 s; a step takes ~1 ms)
constexpr int kDsMmhaWarpsPerPair = 3;
constexpr int kDsMaxPairsPerCta = kDsCW / kDsMmhaWarpsPerPair;

enum
{
    DS_ERR_GRID_BARRIER = 1,
    DS_ERR_RING_FULL_WAIT = 2,
    DS_ERR_RING_EMPTY_WAIT = 3
};

// element (row, k) of a 16-row activation matrix in A-fragment order (index in halves); see the header comment
__host__ __device__ __forceinline__ int frag_index(int row, int k)
{
    const int kb = k >> 6, kk = k & 63;
    const int T = kk >> 4, r = kk & 15;
    const int hi = r >> 3, w = (r & 7) >> 1, e = r & 1;
    const int g = row & 7, up = row >> 3;
    return ((((kb * 4 + w) * 32 + 4 * g + T) * 4 + 2 * hi + up) << 1) + e;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint4 ldcg_u4(const void* p)
{
    return __ldcg(reinterpret_cast<const uint4*>(p));
}

__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(kDsConsumers) : "memory");
}

struct DsShared
{
    uint8_t* ring;
    float* scratch;
    float* stats;
    uint64_t* full;
    uint64_t* empty;
    uint32_t dead;     // shared-space address of the CTA-local "a wait timed out, stop waiting" flag
    uint32_t progress; // shared-space address of the ring producer's item count (read by the L2 prefetch lane)
};

struct DsCtx
{
    DsShared sm;
    unsigned* sync;    // [0] arrivals, [1] exits, [2] status
    long long* dbg;    // optional %globaltimer stamps [CTA][phase][2] (wait returned, work done); tools/step_phases.py
    unsigned phase;    // grid barriers passed so far
    unsigned item;     // ring items consumed so far by this CTA
    int c, G;          // CTA index, number of CTAs
    int tid, warp, lane;
};

__device__ __forceinline__ int ds_dead(const DsShared& sm)
{
    int v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(sm.dead) : "memory");
    return v;
}

__device__ __forceinline__ void ds_set_dead(const DsShared& sm)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(sm.dead), "r"(1) : "memory");
}

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// bounded mbarrier wait (a lost TMA completion or a schedule mismatch must not hang the GPU)
__device__ __forceinline__ void ds_mbar_wait(DsCtx& cx, uint64_t* bar, uint32_t parity, int code)
{
    if (mbar_try_wait(bar, parity))
        return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
    {
        if (ds_dead(cx.sm) || clock64() - t0 > kDsWaitCycles)
        {
            if (!ds_dead(cx.sm))
            {
                ds_set_dead(cx.sm);
                atomicCAS(cx.sync + 2, 0u, (unsigned) code | ((unsigned) cx.c << 8) | (cx.phase << 16));
            }
            return;
        }
    }
}

// ---- grid barrier --------------------------------------------------------------------------------------------
// arrive: every consumer thread's global stores of the phase are ordered before the CTA's release increment
__device__ __forceinline__ long long ds_globaltimer()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

[[maybe_unused]] constexpr int kDsDbgPhases = 512; // stamps per CTA in the debug buffer

// fine-grained stamps of CTA 0 inside a phase (debug buffer region after the per-CTA barrier stamps)
__device__ __forceinline__ void ds_stamp(DsCtx& cx, int slot)
{
#if defined(B200_DS_DEBUG)
    if (cx.dbg != nullptr && cx.c == 0 && cx.tid == 0 && cx.phase < kDsDbgPhases)
        cx.dbg[(size_t) cx.G * kDsDbgPhases * 2 + (size_t) cx.phase * 16 + slot] = clock64(); // SM cycles: fine resolution
#endif
}

__device__ __forceinline__ void ds_grid_arrive(DsCtx& cx)
{
    ds_stamp(cx, 8);
#if defined(B200_DS_DEBUG)
    if (cx.dbg != nullptr && cx.tid == 0 && cx.phase < kDsDbgPhases)
        cx.dbg[((size_t) cx.c * kDsDbgPhases + cx.phase) * 2 + 1] = ds_globaltimer();
#endif
    // bar.sync orders every consumer thread's stores before thread 0's release (cumulativity); the release itself is
    // the only GPU-scope fence on this side (MEMBAR.ALL.GPU + RED, no sequentially-consistent fence)
    consumer_sync();
    if (cx.tid == 0)
        red_release_add_u32(cx.sync, 1u);
    ds_stamp(cx, 9);
    ++cx.phase;
}

// wait until all G CTAs have arrived `phase` times: one thread polls with relaxed GPU-scope loads, then bar.sync.
// The writers release (MEMBAR.GPU + RED) after their data stores, so the data is in L2 before the count moves; every
// load of data written by another CTA is an L1-bypassing GPU-scope load (ld.global.cg) issued after the bar.sync,
// i.e. served by L2 after the count was observed there.  An acquire on this side (ld.acquire.gpu / fence) costs a
// CCTL.IVALL per poll -- the L1 invalidation turns the kernel's few register spills into L2 round trips on the critical
// path (measured: +1.1 us per barrier) -- and protects nothing that is read through L1.
__device__ __forceinline__ void ds_grid_wait(DsCtx& cx)
{
    if (cx.tid == 0 && !ds_dead(cx.sm))
    {
        const unsigned target = cx.phase * (unsigned) cx.G;
        unsigned spins = 0;
        long long t0 = 0;
        while (ld_relaxed_u32(cx.sync) < target)
        {
            if ((++spins & 255u) == 0u)
            {
                if (t0 == 0)
                    t0 = clock64();
                if (clock64() - t0 > kDsWaitCycles || ld_relaxed_u32(cx.sync + 2) != 0u)
                {
                    ds_set_dead(cx.sm);
                    atomicCAS(cx.sync + 2, 0u, (unsigned) DS_ERR_GRID_BARRIER | ((unsigned) cx.c << 8) | (cx.phase << 16));
                    break;
                }
            }
        }
        ds_stamp(cx, 7);
    }
    consumer_sync();
#if defined(B200_DS_DEBUG)
    if (cx.dbg != nullptr && cx.tid == 0 && cx.phase < kDsDbgPhases)
        cx.dbg[((size_t) cx.c * kDsDbgPhases + cx.phase) * 2] = ds_globaltimer();
#endif
}

// ---- static schedule helpers (shared by the producer and the consumers) ------------------------------------------
struct DsGemmShape
{
    int K, N, rot; // rot: tile t belongs to CTA (t + rot) % G -- moves the CTAs that get an extra tile around
    __device__ __forceinline__ int nq() const { return (K + kDsUnitK - 1) / kDsUnitK; }
    __device__ __forceinline__ int first_tile(int c, int G) const { return ((c - rot) % G + G) % G; }
    __device__ __forceinline__ int ntiles(int c, int G) const
    {
        const int f = first_tile(c, G), NT = N >> 3;
        return f < NT ? (NT - f + G - 1) / G : 0;
    }
};

__device__ __forceinline__ int ds_rot(int which, int G)
{
    // qkv, attn_out, cross_q, cross_out, fc1, fc2: spread the CTAs that receive an extra tile over the grid, away from
    // the low CTA indices, which hold an extra (batch, head) pair in the attention phases
    const int r[6] = {72, 24, 36, 48, 108, 60};
    return r[which] % G;
}

struct DsModel
{
    const b200_decoder_layer* layers;
    int L, B, H, d, dff, Smax, S, nch;
};

// ---- producers: one lane each, both walk the whole step's static schedule -------------------------------------------
//   PF = false  warp 10: fills the shared-memory ring (cp.async.bulk + mbarrier), at most kDsSlots items ahead
//   PF = true   warp 11: the same walk, `l2_ahead` items ahead of the ring producer, issuing cp.async.bulk.prefetch.L2
//               only: HBM latency is paid into L2 (tens of MB in flight chip-wide), the ring then fills at L2 latency
__device__ __forceinline__ void ds_l2_prefetch(const void* p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <bool PF>
__device__ __forceinline__ void ds_producer(const DsModel& m, DsCtx& cx, unsigned l2_ahead)
{
    const uint64_t pol = policy_evict_first();
    unsigned it = 0;
    // returns false when the kernel is draining after a timed-out wait
    auto begin_item = [&](uint32_t bytes, uint8_t*& dst, uint64_t*& bar) -> bool
    {
        if constexpr (PF)
        {
            unsigned prog;
            for (;;)
            {
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(prog) : "r"(cx.sm.progress) : "memory");
                if (it < prog + l2_ahead)
                    break;
                if (ds_dead(cx.sm))
                    return false;
                __nanosleep(64);
            }
            dst = nullptr, bar = nullptr;
            ++it;
            return true;
        }
        else
        {
            const unsigned s = it % kDsSlots, use = it / kDsSlots;
            if (use > 0)
                ds_mbar_wait(cx, &cx.sm.empty[s], (use - 1) & 1, DS_ERR_RING_EMPTY_WAIT);
            if (ds_dead(cx.sm))
                return false;
            mbar_arrive_expect_tx(&cx.sm.full[s], bytes);
            ++it;
            asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(cx.sm.progress), "r"(it) : "memory");
            dst = cx.sm.ring + (size_t) s * kDsSlotBytes, bar = &cx.sm.full[s];
            return true;
        }
    };
    auto gemm = [&](const int8_t* W, int K, int N, int which) -> bool
    {
        const DsGemmShape g{K, N, ds_rot(which, cx.G)};
        const int nq = g.nq(), KU = K / nq, first = g.first_tile(cx.c, cx.G), nt = g.ntiles(cx.c, cx.G);
        const uint8_t* base = reinterpret_cast<const uint8_t*>(W);
        for (int j = 0; j < nt; ++j)
        {
            const int tile = first + j * cx.G;
            for (int q = 0; q < nq; ++q)
            {
                uint8_t* dst;
                uint64_t* bar;
                if (!begin_item((uint32_t) 8 * KU, dst, bar))
                    return false;
#pragma unroll
                for (int rp = 0; rp < 4; ++rp)
                {
                    const uint8_t* src = base + (size_t) (4 * tile + rp) * 2 * K + (size_t) q * 2 * KU;
                    if constexpr (PF)
                        ds_l2_prefetch(src, (uint32_t) 2 * KU);
                    else
                        bulk_g2s_hint(dst + (size_t) rp * 2 * KU, src, (uint32_t) 2 * KU, bar, pol);
                }
            }
        }
        return true;
    };
    const int pairs = m.B * m.H;
    for (int l = 0; l < m.L; ++l)
    {
        const b200_decoder_layer& ly = m.layers[l];
        if (!gemm(ly.qkv_w, m.d, 3 * m.d, 0) || !gemm(ly.attn_out_w, m.d, m.d, 1) || !gemm(ly.cross_q_w, m.d, m.d, 2))
            return;
        for (int p = cx.c; p < pairs; p += cx.G)
        {
            const int b = p / m.H, h = p - b * m.H;
            const uint8_t* kb = static_cast<const uint8_t*>(ly.cross_kv) + ((size_t) (b * 2 + 0) * m.H + h) * m.S * kDh;
            const uint8_t* vb = static_cast<const uint8_t*>(ly.cross_kv) + ((size_t) (b * 2 + 1) * m.H + h) * m.S * kDh;
            for (int ch = 0; ch < m.nch; ++ch)
            {
                const int key0 = ch * kDsChunkKeys;
                const uint32_t bytes = (uint32_t) min(kDsChunkKeys, m.S - key0) * kDh;
                uint8_t* dst;
                uint64_t* bar;
                if (!begin_item(2 * bytes, dst, bar))
                    return;
                if constexpr (PF)
                {
                    ds_l2_prefetch(kb + (size_t) key0 * kDh, bytes);
                    ds_l2_prefetch(vb + (size_t) key0 * kDh, bytes);
                }
                else
                {
                    bulk_g2s_hint(dst, kb + (size_t) key0 * kDh, bytes, bar, pol);
                    bulk_g2s_hint(dst + kDsSlotBytes / 2, vb + (size_t) key0 * kDh, bytes, bar, pol);
                }
            }
        }
        if (!gemm(ly.cross_out_w, m.d, m.d, 3) || !gemm(ly.fc1_w, m.d, m.dff, 4) || !gemm(ly.fc2_w, m.dff, m.d, 5))
            return;
    }
}

// ---- matmul phase ------------------------------------------------------------------------------------------------
struct DsGemm
{
    const int8_t* W;
    const __half* scales;
    const __half* bias;
    const __half* gamma; // folded LayerNorm iff non-null (then K <= 1280)
    const float* c1s;
    const float* c2;
    const __half* A;        // A-fragment order, 16 rows x K
    const __half* resid;    // A-fragment order, 16 rows x N, or null
    __half* out_frag;       // A-fragment order, 16 rows x N, or null
    __half* out_rm;         // row-major [rows][N], or null
    int K, N, act, which;
    float eps;
};

__device__ __forceinline__ __half ds_finish(float acc, bool has_bias, float biasv, int activation, bool has_res, float res)
{
    __half o = __float2half_rn(acc);
    if (has_bias)
        o = __float2half_rn(__half2float(o) + biasv);
    if (activation == B200_ACT_GELU_ERF)
        o = __float2half_rn(gelu_erf(__half2float(o)));
    if (has_res)
        o = __float2half_rn(__half2float(o) + res);
    return o;
}

// A fragments (4 MMAs' worth: 16 registers) of k-block `kb` of the activation matrix, this lane's share
__device__ __forceinline__ void ds_load_a(const __half* A, int kb, int lane, uint4 (&dst)[4])
{
    const __half* ap = A + ((size_t) (kb * 4) * 32 + lane) * 8;
#pragma unroll
    for (int w = 0; w < 4; ++w)
        dst[w] = ldcg_u4(ap + (size_t) w * 32 * 8);
}

__device__ __forceinline__ uint32_t ds_sel4(const uint4& v, int w) // w is a compile-time constant after unrolling
{
    return w == 0 ? v.x : (w == 1 ? v.y : (w == 2 ? v.z : v.w));
}

__device__ __forceinline__ __half2 ds_u2h2(uint32_t u)
{
    return *reinterpret_cast<__half2*>(&u);
}

// One k-block (64 k) of up to RT weight tiles: 128-bit shared-memory reads of this lane's column / 16-k chunk, PRMT +
// HSUB2 dequant, optional gamma pair, and the MMAs with w outer / tile inner so the tiles' accumulator chains interleave.
// D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col); not volatile: a pure function of its operands, so the
// compiler may interleave the independent chains of different tiles and hoist the dequant of the next operand
__device__ __forceinline__ void ds_mma(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// One k-block (64 k) of exactly NT weight tiles (compile-time: no per-tile branches around the warp-synchronous MMAs):
// 128-bit shared-memory reads of this lane's column / 16-k chunk, PRMT + HSUB2 dequant, optional gamma pair, and the
// MMAs with w outer / tile inner so the tiles' accumulator chains interleave.
template <int NT, bool FOLD>
__device__ __forceinline__ void ds_mma_kblock(float (&acc)[NT][4], const uint4 (&af)[4], const uint8_t* ring, unsigned item0,
    int item_stride, uint32_t lane_off, const uint4& glo, const uint4& ghi)
{
    uint4 wv[NT];
#pragma unroll
    for (int jj = 0; jj < NT; ++jj)
        wv[jj] = *reinterpret_cast<const uint4*>(ring + (size_t) ((item0 + (unsigned) (jj * item_stride)) % kDsSlots) * kDsSlotBytes + lane_off);
#pragma unroll
    for (int w = 0; w < 4; ++w)
    {
        const __half2 g_lo = ds_u2h2(ds_sel4(glo, w)), g_hi = ds_u2h2(ds_sel4(ghi, w));
#pragma unroll
        for (int jj = 0; jj < NT; ++jj)
        {
            __half2 lo, hi;
            dequant_word(ds_sel4(wv[jj], w), lo, hi);
            if constexpr (FOLD)
            {
                lo = __hmul2(lo, g_lo);
                hi = __hmul2(hi, g_hi);
            }
            ds_mma(acc[jj], af[w], h2u(lo), h2u(hi));
        }
    }
}

// this warp's two k-blocks of NT tiles starting at ring item `item0` (unit stride `item_stride`), partial sums to `scr`
template <int NT, bool FOLD>
__device__ __forceinline__ void ds_mma_tiles(const uint4 (&af0)[4], const uint4 (&af1)[4], bool kv0, bool kv1, const uint8_t* ring,
    unsigned item0, int item_stride, uint32_t off0, uint32_t off1, const uint4 (&glo)[2], const uint4 (&ghi)[2], float* scr,
    int g, int t, bool accumulate)
{
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        if (accumulate)
        {
            const float2 lo = *reinterpret_cast<const float2*>(scr + (size_t) j * 128 + g * 8 + 2 * t);
            const float2 hi = *reinterpret_cast<const float2*>(scr + (size_t) j * 128 + (g + 8) * 8 + 2 * t);
            acc[j][0] = lo.x, acc[j][1] = lo.y, acc[j][2] = hi.x, acc[j][3] = hi.y;
        }
        else
            acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    }
    if (kv0)
        ds_mma_kblock<NT, FOLD>(acc, af0, ring, item0, item_stride, off0, glo[0], ghi[0]);
    if (kv1)
        ds_mma_kblock<NT, FOLD>(acc, af1, ring, item0, item_stride, off1, glo[1], ghi[1]);
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        *reinterpret_cast<float2*>(scr + (size_t) j * 128 + g * 8 + 2 * t) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2*>(scr + (size_t) j * 128 + (g + 8) * 8 + 2 * t) = make_float2(acc[j][2], acc[j][3]);
    }
}

// rt (1..5) tiles as groups of at most three (register budget), each group a branch-free instantiation
template <bool FOLD>
__device__ __forceinline__ void ds_mma_round(int rt, const uint4 (&af0)[4], const uint4 (&af1)[4], bool kv0, bool kv1,
    const uint8_t* ring, unsigned item0, int item_stride, uint32_t off0, uint32_t off1, const uint4 (&glo)[2],
    const uint4 (&ghi)[2], float* scr, int g, int t, bool accumulate)
{
#define DS_TILES(NT, FIRST)                                                                                            \
    ds_mma_tiles<NT, FOLD>(af0, af1, kv0, kv1, ring, item0 + (unsigned) ((FIRST) * item_stride), item_stride, off0, off1, glo, \
        ghi, scr + (size_t) (FIRST) * 128, g, t, accumulate)
    switch (rt)
    {
    case 1: DS_TILES(1, 0); break;
    case 2: DS_TILES(2, 0); break;
    case 3: DS_TILES(3, 0); break;
    case 4: DS_TILES(2, 0); DS_TILES(2, 2); break;
    case 5: DS_TILES(3, 0); DS_TILES(2, 3); break;
    default: break;
    }
#undef DS_TILES
}

__device__ __forceinline__ void ds_gemm_phase(const DsGemm& a, const DsModel& m, DsCtx& cx)
{
    const int lane = cx.lane, warp = cx.warp, tid = cx.tid;
    const int g = lane >> 2, t = lane & 3;
    const DsGemmShape shp{a.K, a.N, ds_rot(a.which, cx.G)};
    const int nq = shp.nq(), KU = a.K / nq, nkbu = KU >> 6;
    const int first = shp.first_tile(cx.c, cx.G), nt = shp.ntiles(cx.c, cx.G);
    const bool fold = a.gamma != nullptr; // only with nq == 1 (hidden size <= 1280, checked on the host)
    // tiles per reduction round.  Deep-K matmuls (fc2: four 1280-k units per tile) take 2 tiles per round so a round
    // never holds more than 8 ring slots, and run on the second, register-lean code path below.
    constexpr int RB = 2;
    const int R = nq == 1 ? kDsRoundTiles : RB;
    // this warp's k-blocks inside a unit: warp, warp + 10
    const int kbl0 = warp, kbl1 = warp + kDsCW;
    const bool kv0 = kbl0 < nkbu, kv1 = kbl1 < nkbu;
    // byte offset of this lane's 16 bytes inside a k-block of a weight unit: column g of the tile (row pair g/2, parity
    // g%2), 16-k chunk t
    const uint32_t lane_off = (uint32_t) ((g >> 1) * 2 * KU + (g & 1) * 64 + t * 16);

    // ---- static operands, requested before the grid barrier ----
    // epilogue item of this thread in a round: (tile jj, row, column pair cp)
    const int e_jj = tid >> 6, e_row = (tid & 63) >> 2, e_cp = tid & 3;
    float e_sc[2] = {0.f, 0.f}, e_bias[2] = {0.f, 0.f}, e_c1[2] = {0.f, 0.f}, e_c2[2] = {0.f, 0.f}, e_res[2] = {0.f, 0.f};
    // The per-column vectors are static data in HBM: requested before the grid barrier (a DRAM round trip hidden behind
    // the wait).
    auto load_epi_static = [&](int j0)
    {
        if (e_jj < min(R, nt - j0))
        {
            const int n0 = 8 * (first + (j0 + e_jj) * cx.G) + 2 * e_cp;
            const float2 s2 = __half22float2(__ldg(reinterpret_cast<const __half2*>(a.scales + n0)));
            e_sc[0] = s2.x, e_sc[1] = s2.y;
            if (a.bias != nullptr)
            {
                const float2 b2 = __half22float2(__ldg(reinterpret_cast<const __half2*>(a.bias + n0)));
                e_bias[0] = b2.x, e_bias[1] = b2.y;
            }
            if (fold)
            {
                const float2 c1 = __ldg(reinterpret_cast<const float2*>(a.c1s + n0));
                const float2 c2 = __ldg(reinterpret_cast<const float2*>(a.c2 + n0));
                e_c1[0] = c1.x, e_c1[1] = c1.y, e_c2[0] = c2.x, e_c2[1] = c2.y;
            }
            if (a.resid != nullptr && e_row < m.B)
            {
                // written at least two phases ago: safe to fetch ahead of this phase's barrier
                const float2 r2 = __half22float2(__ldcg(reinterpret_cast<const __half2*>(a.resid + frag_index(e_row, n0))));
                e_res[0] = r2.x, e_res[1] = r2.y;
            }
        }
    };
    load_epi_static(0);
    uint4 glo[2], ghi[2]; // gamma pairs of this thread's B fragments: lo = k 16t+2w.., hi = +8, per k-block
    glo[0] = glo[1] = ghi[0] = ghi[1] = make_uint4(0u, 0u, 0u, 0u);
    if (fold)
    {
        if (kv0)
        {
            const uint4* gp = reinterpret_cast<const uint4*>(a.gamma + 64 * kbl0 + 16 * t);
            glo[0] = __ldg(gp), ghi[0] = __ldg(gp + 1);
        }
        if (kv1)
        {
            const uint4* gp = reinterpret_cast<const uint4*>(a.gamma + 64 * kbl1 + 16 * t);
            glo[1] = __ldg(gp), ghi[1] = __ldg(gp + 1);
        }
    }

    ds_grid_wait(cx);
    ds_stamp(cx, 0);

    for (int j0 = 0; j0 < nt || j0 == 0; j0 += R)
    {
        const int rt = max(0, min(R, nt - j0));
        float* scr = cx.sm.scratch + ((size_t) (warp * kDsRoundTiles) * 16) * 8;
        if (nq == 1)
        {
            // ---- path A: one 1280-k unit per tile, up to 5 tiles, optional folded LayerNorm ----
            uint4 af0[4], af1[4];
            if (kv0)
                ds_load_a(a.A, kbl0, lane, af0);
            if (kv1)
                ds_load_a(a.A, kbl1, lane, af1);
#pragma unroll
            for (int jj = 0; jj < kDsRoundTiles; ++jj)
            {
                if (jj < rt)
                {
                    const unsigned it = cx.item + (unsigned) jj;
                    ds_mbar_wait(cx, &cx.sm.full[it % kDsSlots], (it / kDsSlots) & 1, DS_ERR_RING_FULL_WAIT);
                }
            }
            ds_stamp(cx, 4);
            if (fold)
                ds_mma_round<true>(rt, af0, af1, kv0, kv1, cx.sm.ring, cx.item, 1, lane_off + kbl0 * 128, lane_off + kbl1 * 128, glo,
                    ghi, scr, g, t, false);
            else
                ds_mma_round<false>(rt, af0, af1, kv0, kv1, cx.sm.ring, cx.item, 1, lane_off + kbl0 * 128, lane_off + kbl1 * 128, glo,
                    ghi, scr, g, t, false);
            ds_stamp(cx, 5);
            if (fold && j0 == 0)
            {
                // LayerNorm statistics of rows g and g+8 over this warp's k-blocks (shifted sums, then exact halving
                // merges over the 4 lanes of a row: equal counts); after the MMAs were issued: off their critical path
                float st_n = 0.f, st_mean[2] = {0.f, 0.f}, st_m2[2] = {0.f, 0.f};
                if (kv0)
                {
                    const float sh0 = __low2float(*reinterpret_cast<const __half2*>(&af0[0].x));
                    const float sh1 = __low2float(*reinterpret_cast<const __half2*>(&af0[0].y));
                    float sa0 = 0.f, sb0 = 0.f, sa1 = 0.f, sb1 = 0.f;
                    auto accum = [&](const uint4 (&af)[4])
                    {
#pragma unroll
                        for (int w = 0; w < 4; ++w)
                        {
                            const uint32_t r0[2] = {af[w].x, af[w].z};
                            const uint32_t r1[2] = {af[w].y, af[w].w};
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                            {
                                const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&r0[i]));
                                const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&r1[i]));
                                const float d0 = f0.x - sh0, d1 = f0.y - sh0, d2 = f1.x - sh1, d3 = f1.y - sh1;
                                sa0 += d0 + d1;
                                sb0 = fmaf(d0, d0, fmaf(d1, d1, sb0));
                                sa1 += d2 + d3;
                                sb1 = fmaf(d2, d2, fmaf(d3, d3, sb1));
                            }
                        }
                    };
                    accum(af0);
                    if (kv1)
                        accum(af1);
                    float cn = kv1 ? 32.f : 16.f;
                    const float rn = 1.f / cn;
                    float cm0 = sh0 + sa0 * rn, cM0 = sb0 - sa0 * sa0 * rn;
                    float cm1 = sh1 + sa1 * rn, cM1 = sb1 - sa1 * sa1 * rn;
#pragma unroll
                    for (int o = 1; o < 4; o <<= 1)
                    {
                        const float om0 = __shfl_xor_sync(0xffffffffu, cm0, o), oM0 = __shfl_xor_sync(0xffffffffu, cM0, o);
                        const float om1 = __shfl_xor_sync(0xffffffffu, cm1, o), oM1 = __shfl_xor_sync(0xffffffffu, cM1, o);
                        const float dl0 = om0 - cm0, dl1 = om1 - cm1;
                        cm0 += 0.5f * dl0;
                        cM0 += oM0 + dl0 * dl0 * (0.5f * cn);
                        cm1 += 0.5f * dl1;
                        cM1 += oM1 + dl1 * dl1 * (0.5f * cn);
                        cn *= 2.f;
                    }
                    st_n = cn, st_mean[0] = cm0, st_m2[0] = cM0, st_mean[1] = cm1, st_m2[1] = cM1;
                }
                if (t == 0)
                {
                    float* sp = cx.sm.stats + (size_t) warp * 16 * 3;
                    sp[g * 3 + 0] = st_n, sp[g * 3 + 1] = st_mean[0], sp[g * 3 + 2] = st_m2[0];
                    sp[(g + 8) * 3 + 0] = st_n, sp[(g + 8) * 3 + 1] = st_mean[1], sp[(g + 8) * 3 + 2] = st_m2[1];
                }
            }
        }
        else
        {
            // ---- path B: deep K (nq units per tile), <= 2 tiles per round.  The A fragments of two quarters are
            // requested together, so the nq L2 round trips overlap pairwise instead of queueing behind each other. ----
            // The partial sums of a warp live in its own scratch rows between quarters (re-read by the same thread: no
            // barrier needed), so the register footprint is one quarter pair's A fragments + <= 2 tiles of accumulators.
            uint4 afA[2][4], afB[2][4]; // even / odd quarter of a pair: [k-block][w]
#pragma unroll 1
            for (int q = 0; q < nq; q += 2)
            {
                const bool two = q + 1 < nq;
                if (kv0)
                    ds_load_a(a.A, q * nkbu + kbl0, lane, afA[0]);
                if (kv1)
                    ds_load_a(a.A, q * nkbu + kbl1, lane, afA[1]);
                if (two && kv0)
                    ds_load_a(a.A, (q + 1) * nkbu + kbl0, lane, afB[0]);
                if (two && kv1)
                    ds_load_a(a.A, (q + 1) * nkbu + kbl1, lane, afB[1]);
#pragma unroll
                for (int jj = 0; jj < RB; ++jj)
                {
                    if (jj < rt)
                    {
                        const unsigned it = cx.item + (unsigned) (jj * nq + q);
                        ds_mbar_wait(cx, &cx.sm.full[it % kDsSlots], (it / kDsSlots) & 1, DS_ERR_RING_FULL_WAIT);
                        if (two)
                            ds_mbar_wait(cx, &cx.sm.full[(it + 1) % kDsSlots], ((it + 1) / kDsSlots) & 1, DS_ERR_RING_FULL_WAIT);
                    }
                }
                if (q == 0)
                    ds_stamp(cx, 4);
                ds_mma_round<false>(rt, afA[0], afA[1], kv0, kv1, cx.sm.ring, cx.item + (unsigned) q, nq, lane_off + kbl0 * 128,
                    lane_off + kbl1 * 128, glo, ghi, scr, g, t, q > 0);
                if (two)
                    ds_mma_round<false>(rt, afB[0], afB[1], kv0, kv1, cx.sm.ring, cx.item + (unsigned) (q + 1), nq,
                        lane_off + kbl0 * 128, lane_off + kbl1 * 128, glo, ghi, scr, g, t, true);
            }
            ds_stamp(cx, 5);
        }
        ds_stamp(cx, 1);
        if (j0 > 0)
            load_epi_static(j0); // later rounds (small grids only): in flight under the reduction
        consumer_sync();
        ds_stamp(cx, 2);
        if (tid == 0)
        {
            // the round's weight slots are drained: hand them back to the producer
            for (int u = 0; u < rt * nq; ++u)
                mbar_arrive(&cx.sm.empty[(cx.item + (unsigned) u) % kDsSlots]);
        }
        cx.item += (unsigned) (rt * nq);
        // ---- reduce the ten k-partials in warp order, epilogue ----
        if (e_jj < rt)
        {
            float s0 = 0.f, s1 = 0.f;
            const float* src = cx.sm.scratch + ((size_t) e_jj * 16 + e_row) * 8 + 2 * e_cp;
#pragma unroll
            for (int w = 0; w < kDsCW; ++w)
            {
                const float2 v = *reinterpret_cast<const float2*>(src + (size_t) w * kDsRoundTiles * 16 * 8);
                s0 += v.x;
                s1 += v.y;
            }
            float v0 = s0 * e_sc[0], v1 = s1 * e_sc[1];
            if (fold)
            {
                // Chan merge of the ten k-range partials of this row, in warp order
                float cn = 0.f, cm = 0.f, cM = 0.f;
#pragma unroll
                for (int w = 0; w < kDsCW; ++w)
                {
                    const float* sp = cx.sm.stats + ((size_t) w * 16 + e_row) * 3;
                    const float on = sp[0], om = sp[1], oM = sp[2];
                    const float nn = cn + on, dl = om - cm;
                    const float inv = nn > 0.f ? __fdividef(1.f, nn) : 0.f;
                    cm += dl * on * inv;
                    cM += oM + dl * dl * cn * on * inv;
                    cn = nn;
                }
                const float rstd = rsqrtf(__fdividef(cM, cn) + a.eps);
                v0 = rstd * (v0 - cm * e_c1[0]) + e_c2[0];
                v1 = rstd * (v1 - cm * e_c1[1]) + e_c2[1];
            }
            if (e_row < m.B)
            {
                const bool hb = a.bias != nullptr, hr = a.resid != nullptr;
                const __half o0 = ds_finish(v0, hb, e_bias[0], a.act, hr, e_res[0]);
                const __half o1 = ds_finish(v1, hb, e_bias[1], a.act, hr, e_res[1]);
                const __half2 o2 = __halves2half2(o0, o1);
                const int n0 = 8 * (first + (j0 + e_jj) * cx.G) + 2 * e_cp;
                if (a.out_frag != nullptr)
                    *reinterpret_cast<__half2*>(a.out_frag + frag_index(e_row, n0)) = o2;
                if (a.out_rm != nullptr)
                    *reinterpret_cast<__half2*>(a.out_rm + (size_t) e_row * a.N + n0) = o2;
            }
        }
        if (j0 + R < nt)
            consumer_sync(); // the scratch area is reused by the next round
    }
    ds_stamp(cx, 3);
    ds_grid_arrive(cx);
}

// ---- masked self-attention (generation), int8 KV cache ------------------------------------------------------------
// Same arithmetic as mmha_generation_kernel (attention.cu): cached keys as exact fp16 integers, q.k in HFMA2 chains of 4,
// fp32 softmax with 1/(sum + 1e-6), p.v in HFMA2 chains of <= 4 keys flushed to fp32, the current token's k / v
// unquantized, the new K / V quantized with cvt.rni.sat and appended.
struct DsMmha
{
    const __half* qkv; // row-major [B][3d]
    int8_t* cache;
    const float* s_oq;
    const float* s_qo;
    const int* seq_len;
    __half* ctx_frag;
};

__device__ __forceinline__ void ds_mmha_phase(const DsMmha& a, const DsModel& m, DsCtx& cx)
{
    constexpr int NIT = 4;
    const int lane = cx.lane, warp = cx.warp;
    const int chunk = lane & 3, kl = lane >> 2;
    const int pairs = m.B * m.H;
    const int nbh = cx.c < pairs ? (pairs - cx.c + cx.G - 1) / cx.G : 0; // <= 3 (checked on the host)
    // three warps per pair WHATEVER the batch size: the split of the keys over warps fixes the summation order, and an
    // utterance's bits must not depend on its batch mates (the multi-GPU sharding property)
    constexpr int nw = kDsMmhaWarpsPerPair;
    const int r = warp / nw, wi = warp - r * nw;
    const bool active = nbh > 0 && r < nbh;
    const int p = cx.c + r * cx.G;
    const int b = active ? p / m.H : 0, h = active ? p - b * m.H : 0;
    const int hidden = m.d;
    char* kc = reinterpret_cast<char*>(a.cache) + ((size_t) (b * 2 + 0) * m.H + h) * m.Smax * kDh;
    char* vc = reinterpret_cast<char*>(a.cache) + ((size_t) (b * 2 + 1) * m.H + h) * m.Smax * kDh;
    float* parts = cx.sm.scratch; // [warp][kDsPart]

    int tlen = 0;
    float s_qo = 1.f, s_oq = 1.f;
    // raw cache bytes of one pass (16 dims of NIT keys per lane for K and for V); converted to fp16 at the point of use:
    // keeping the converted values live across the barrier spilled registers, and a spill is an L2 round trip here
    KvChunk<true> kreg[NIT], vreg[NIT];
    auto fetch = [&](int pass)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            const int key = min((wi + nw * (NIT * pass + it)) * 8 + kl, m.Smax - 1);
            kreg[it].load(kc, (size_t) key * kDh + chunk * 16);
            vreg[it].load(vc, (size_t) key * kDh + chunk * 16);
        }
    };
    if (active)
    {
        // the cache rows below the length, the length and the scales were written by earlier launches: fetch them
        // ahead of the barrier
        tlen = min(a.seq_len[b], m.Smax - 1);
        s_qo = __ldg(a.s_qo);
        s_oq = __ldg(a.s_oq);
        fetch(0);
    }
    ds_grid_wait(cx);

    float m_run = -FLT_MAX, l_run = 0.f, s_cur = -FLT_MAX;
    __half2 vcur_h = __float2half2_rn(0.f); // the finalising warp's two dims of this step's v
    float o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        o[i] = 0.f;
    const float inv_sqrt_dh = 0.125f; // 1 / sqrt(64), q_scaling = 1 (gptAttentionCommon.cpp:163)
    if (active)
    {
        const float sscale = s_qo * inv_sqrt_dh;
        const __half* qp = a.qkv + (size_t) b * 3 * hidden + h * kDh + chunk * 16;
        __half qh[16], kh[16], vh[16];
        {
            const uint4 v0 = ldcg_u4(qp), v1 = ldcg_u4(qp + 8);
            *reinterpret_cast<uint4*>(&qh[0]) = v0;
            *reinterpret_cast<uint4*>(&qh[8]) = v1;
        }
        if (wi == 0)
        {
            vcur_h = __ldcg(reinterpret_cast<const __half2*>(a.qkv + (size_t) b * 3 * hidden + 2 * hidden + h * kDh + 2 * lane));
            const uint4 k0 = ldcg_u4(qp + hidden), k1 = ldcg_u4(qp + hidden + 8);
            const uint4 v0 = ldcg_u4(qp + 2 * hidden), v1 = ldcg_u4(qp + 2 * hidden + 8);
            *reinterpret_cast<uint4*>(&kh[0]) = k0;
            *reinterpret_cast<uint4*>(&kh[8]) = k1;
            *reinterpret_cast<uint4*>(&vh[0]) = v0;
            *reinterpret_cast<uint4*>(&vh[8]) = v1;
            // append this step's K and V (lane group 0 writes K, group 1 writes V; 16 dims per lane)
            if (kl == 0)
                store16<true>(kc, (size_t) tlen * kDh + chunk * 16, s_oq, kh);
            else if (kl == 1)
                store16<true>(vc, (size_t) tlen * kDh + chunk * 16, s_oq, vh);
            float sc0 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                sc0 = fmaf(__half2float(qh[i]), __half2float(kh[i]), sc0);
            sc0 += __shfl_xor_sync(0xffffffffu, sc0, 1);
            sc0 += __shfl_xor_sync(0xffffffffu, sc0, 2);
            s_cur = sc0 * inv_sqrt_dh;
        }
        __half2 q2[8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            q2[2 * i] = __halves2half2(qh[4 * i], qh[4 * i + 2]);
            q2[2 * i + 1] = __halves2half2(qh[4 * i + 1], qh[4 * i + 3]);
        }
        const int ngroups = (tlen + 7) >> 3;
        for (int pass = 0; (wi + nw * NIT * pass) < ngroups; ++pass)
        {
            if (pass > 0)
                fetch(pass);
            float sc[NIT];
            float m_new = m_run;
#pragma unroll
            for (int it = 0; it < NIT; ++it)
            {
                sc[it] = -FLT_MAX;
                const int kg = (wi + nw * (NIT * pass + it)) * 8;
                if (kg < tlen)
                {
                    __half2 kw[8];
                    kreg[it].unpack(kw);
                    __half2 h0 = __hmul2(q2[0], kw[0]);
                    __half2 h1 = __hmul2(q2[4], kw[4]);
                    h0 = __hfma2(q2[1], kw[1], h0);
                    h1 = __hfma2(q2[5], kw[5], h1);
                    h0 = __hfma2(q2[2], kw[2], h0);
                    h1 = __hfma2(q2[6], kw[6], h1);
                    h0 = __hfma2(q2[3], kw[3], h0);
                    h1 = __hfma2(q2[7], kw[7], h1);
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                    float sv = (f0.x + f0.y) + (f1.x + f1.y);
                    sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                    sv += __shfl_xor_sync(0xffffffffu, sv, 2);
                    sv = (kg + kl < tlen) ? sv * sscale : -FLT_MAX;
                    sc[it] = sv;
                    m_new = fmaxf(m_new, sv);
                }
            }
            m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, 4));
            m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, 8));
            m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, 16));
            const float corr = m_new == -FLT_MAX ? 1.f : __expf(m_run - m_new);
            m_run = m_new;
            l_run *= corr;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                o[i] *= corr;
            __half2 o2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                o2[i] = __float2half2_rn(0.f);
#pragma unroll
            for (int it = 0; it < NIT; ++it)
            {
                if ((wi + nw * (NIT * pass + it)) * 8 >= tlen || sc[it] == -FLT_MAX)
                    continue;
                const float e = __expf(sc[it] - m_new);
                l_run += e;
                const __half2 p2 = __float2half2_rn(e);
                __half2 vw[8];
                vreg[it].unpack(vw);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    o2[i] = __hfma2(p2, vw[i], o2[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const float2 f = __half22float2(o2[i]);
                o[2 * i] += f.x;
                o[2 * i + 1] += f.y;
            }
        }
        // reduce over the 8 key groups
        l_run += __shfl_xor_sync(0xffffffffu, l_run, 4);
        l_run += __shfl_xor_sync(0xffffffffu, l_run, 8);
        l_run += __shfl_xor_sync(0xffffffffu, l_run, 16);
#pragma unroll
        for (int i = 0; i < 16; ++i)
        {
            float v = o[i];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            o[i] = v;
        }
        float* pr = parts + (size_t) warp * kDsPart;
        if (kl == 0)
        {
            // o[2i], o[2i+1] hold the pair of w[i]: w[2j] = dims (4j, 4j+2), w[2j+1] = dims (4j+1, 4j+3) -> natural order
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(pr + 4 + chunk * 16 + 4 * j) = make_float4(o[4 * j], o[4 * j + 2], o[4 * j + 1], o[4 * j + 3]);
            if (chunk == 0)
            {
                pr[0] = m_run;
                pr[1] = l_run;
                pr[2] = s_cur; // meaningful for the pair's first warp only
            }
        }
    }
    consumer_sync();
    if (active && wi == 0)
    {
        const float* pb = parts + (size_t) (r * nw) * kDsPart;
        const float sc_cur = pb[2];
        float gm = sc_cur;
        for (int w2 = 0; w2 < nw; ++w2)
            gm = fmaxf(gm, pb[w2 * kDsPart]);
        const float e_cur = __expf(sc_cur - gm);
        float gl = e_cur, a0 = 0.f, a1 = 0.f;
        for (int w2 = 0; w2 < nw; ++w2)
        {
            const float* ps = pb + w2 * kDsPart;
            const float wt = ps[0] == -FLT_MAX ? 0.f : __expf(ps[0] - gm);
            gl += wt * ps[1];
            const float2 ov = *reinterpret_cast<const float2*>(ps + 4 + 2 * lane);
            a0 += wt * ov.x;
            a1 += wt * ov.y;
        }
        const float inv_sum = __fdividef(1.f, gl + 1.e-6f); // Template.h:1756
        const float2 vcur = __half22float2(vcur_h);
        const __half2 o2 = __floats2half2_rn((a0 * s_qo + e_cur * vcur.x) * inv_sum, (a1 * s_qo + e_cur * vcur.y) * inv_sum);
        *reinterpret_cast<__half2*>(a.ctx_frag + frag_index(b, h * kDh + 2 * lane)) = o2;
    }
    ds_grid_arrive(cx);
}

// ---- cross-attention over the int8 cross-KV cache, chunks from the ring -----------------------------------------
struct DsXattn
{
    const __half* q;  // row-major [B][d]
    const float* s_qo;
    __half* ctx_frag;
};

__device__ __forceinline__ void ds_xattn_phase(const DsXattn& a, const DsModel& m, DsCtx& cx)
{
    constexpr int NIT = kDsChunkKeys / 8;
    const int lane = cx.lane, warp = cx.warp;
    const int chunk = lane & 3, kl = lane >> 2;
    const int pairs = m.B * m.H;
    const int nbh = cx.c < pairs ? (pairs - cx.c + cx.G - 1) / cx.G : 0;
    float* parts = cx.sm.scratch; // [2][kDsCW][kDsPart]
    const float s_qo = __ldg(a.s_qo);
    const float sscale = s_qo * 0.125f * 1.4426950408889634f;

    ds_grid_wait(cx);
#if defined(B200_DS_DEBUG)
    long long dbg_wait = 0, dbg_chunks = 0;
    const long long dbg_t0 = clock64();
#endif

    uint4 qn0 = make_uint4(0, 0, 0, 0), qn1 = qn0;
    if (nbh > 0)
    {
        const int b = cx.c / m.H, h = cx.c - b * m.H;
        const __half* qs = a.q + (size_t) b * m.d + h * kDh + chunk * 16;
        qn0 = ldcg_u4(qs);
        qn1 = ldcg_u4(qs + 8);
    }
    for (int r = 0; r < nbh; ++r)
    {
        const int p = cx.c + r * cx.G;
        const int b = p / m.H, h = p - b * m.H;
        uint32_t bq[8];
        float koff; // 1152 * sum of q over the 64 dims: the bias of the 1024 + byte key values (xa_chunk KOFF)
        {
            const uint32_t u[8] = {qn0.x, qn0.y, qn0.z, qn0.w, qn1.x, qn1.y, qn1.z, qn1.w}; // u[j] = (d2j, d2j+1)
            float qs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                bq[2 * j] = kl == 0 ? __byte_perm(u[2 * j], u[2 * j + 1], 0x5410) : 0u;
                bq[2 * j + 1] = kl == 0 ? __byte_perm(u[2 * j], u[2 * j + 1], 0x7632) : 0u;
                const float2 f0 = __half22float2(ds_u2h2(u[2 * j])), f1 = __half22float2(ds_u2h2(u[2 * j + 1]));
                qs += (f0.x + f0.y) + (f1.x + f1.y);
            }
            qs += __shfl_xor_sync(0xffffffffu, qs, 1);
            qs += __shfl_xor_sync(0xffffffffu, qs, 2);
            koff = 1152.f * qs;
        }
        if (r + 1 < nbh)
        {
            const int pn = p + cx.G;
            const int bn = pn / m.H, hn = pn - bn * m.H;
            const __half* qs = a.q + (size_t) bn * m.d + hn * kDh + chunk * 16;
            qn0 = ldcg_u4(qs);
            qn1 = ldcg_u4(qs + 8);
        }
        float m_run = -FLT_MAX, l_run = 0.f;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            o[j] = 0.f;
        for (int ch = 0; ch < m.nch; ++ch)
        {
            const unsigned i = (unsigned) (r * m.nch + ch);
            if ((int) (i % kDsCW) != warp)
                continue;
            const unsigned it = cx.item + i;
            const unsigned s = it % kDsSlots;
            const int nk = min(kDsChunkKeys, m.S - ch * kDsChunkKeys);
#if defined(B200_DS_DEBUG)
            const long long tw0 = clock64();
#endif
            ds_mbar_wait(cx, &cx.sm.full[s], (it / kDsSlots) & 1, DS_ERR_RING_FULL_WAIT);
#if defined(B200_DS_DEBUG)
            dbg_wait += clock64() - tw0;
            ++dbg_chunks;
#endif
            const uint8_t* kst = cx.sm.ring + (size_t) s * kDsSlotBytes + (size_t) (kl * kDh + chunk * 16);
            const uint8_t* vst = kst + kDsSlotBytes / 2;
            // (Splitting the 80-key chunk into 48 + 32 keys -- two online-softmax updates per slot, 6 instead of 10 score
            // registers -- removed the spills of this loop but produced wrong sums for full chunks on the B200; the cause
            // was not found in the time available, so the chunk stays one 10-iteration update.)
            if (nk == kDsChunkKeys)
                xa_chunk<true, NIT, true, true>(kst, vst, nk, kl, lane, sscale, bq, m_run, l_run, o, koff);
            else
                xa_chunk<true, NIT, false, true>(kst, vst, nk, kl, lane, sscale, bq, m_run, l_run, o, koff);
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&cx.sm.empty[s]);
        }
        // this warp's state, reduced over its 8 key groups -> shared memory
        float l = l_run;
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
        for (int j = 0; j < 16; ++j)
        {
            float v = o[j];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            o[j] = v;
        }
        float* pr = parts + (size_t) ((r & 1) * kDsCW + warp) * kDsPart;
        if (kl == 0)
        {
            float* dst = pr + 4 + chunk * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(o[4 * j + 0], o[4 * j + 2], o[4 * j + 1], o[4 * j + 3]);
            if (chunk == 0)
            {
                pr[0] = m_run;
                pr[1] = l;
            }
        }
        consumer_sync();
        if (warp == r % kDsCW)
        {
            const float* pb = parts + (size_t) (r & 1) * kDsCW * kDsPart;
            float gm = -FLT_MAX;
#pragma unroll
            for (int w2 = 0; w2 < kDsCW; ++w2)
                gm = fmaxf(gm, pb[w2 * kDsPart]);
            float gl = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < kDsCW; ++w2)
            {
                const float* ps = pb + w2 * kDsPart;
                const float wt = fast_exp2(ps[0] - gm); // 0 for a warp that saw no chunk of this pair (m = -FLT_MAX)
                gl += wt * ps[1];
                const float2 ov = *reinterpret_cast<const float2*>(ps + 4 + 2 * lane);
                a0 += wt * ov.x;
                a1 += wt * ov.y;
            }
            const float inv = s_qo / gl; // hoisted V dequant scale
            *reinterpret_cast<__half2*>(a.ctx_frag + frag_index(b, h * kDh + 2 * lane)) = __floats2half2_rn(a0 * inv, a1 * inv);
        }
    }
    cx.item += (unsigned) (nbh * m.nch);
#if defined(B200_DS_DEBUG)
    if (cx.dbg != nullptr && cx.c == 30 && cx.tid == 0 && cx.phase < kDsDbgPhases)
    {
        long long* f = cx.dbg + (size_t) cx.G * kDsDbgPhases * 2 + (size_t) cx.phase * 16;
        f[10] = dbg_wait, f[11] = dbg_chunks, f[12] = clock64() - dbg_t0;
    }
#endif
    ds_grid_arrive(cx);
}

struct DsParams
{
    DsModel m;
    const int* tokens;
    const int* seq_len;
    const __half* tok_emb;
    const __half* pos_emb;
    __half* x_out;
    unsigned* sync;
    long long* dbg;
    __half* x;   // A-fragment order [16 x d]
    __half* ctx; // A-fragment order [16 x d]
    __half* u;   // A-fragment order [16 x dff]
    __half* qkv; // row-major [16][3d]
    __half* q;   // row-major [16][d]
    int vocab, n_ctx;
    int l2_ahead; // items the L2 prefetch lane runs ahead of the ring producer (0: off)
    float eps;
};

__global__ void __launch_bounds__(kDsThreads, 1) decoder_step_kernel(const DsParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    DsCtx cx;
    cx.sm.ring = smem;
    cx.sm.scratch = reinterpret_cast<float*>(smem + (size_t) kDsSlots * kDsSlotBytes);
    cx.sm.stats = cx.sm.scratch + kDsScratchFloats;
    cx.sm.full = reinterpret_cast<uint64_t*>(cx.sm.stats + kDsStatFloats);
    cx.sm.empty = cx.sm.full + kDsSlots;
    cx.sm.dead = smem_u32(cx.sm.empty + kDsSlots);
    cx.sm.progress = cx.sm.dead + 4;
    cx.sync = p.sync;
    cx.dbg = p.dbg;
    cx.phase = 0;
    cx.item = 0;
    cx.c = blockIdx.x;
    cx.G = gridDim.x;
    cx.tid = threadIdx.x;
    cx.warp = threadIdx.x >> 5;
    cx.lane = threadIdx.x & 31;
    const DsModel& m = p.m;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < kDsSlots; ++s)
        {
            mbar_init(&cx.sm.full[s], 1);
            mbar_init(&cx.sm.empty[s], 1);
        }
        reinterpret_cast<volatile int*>(cx.sm.empty + kDsSlots)[0] = 0;
        reinterpret_cast<volatile int*>(cx.sm.empty + kDsSlots)[1] = 0;
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncthreads();

    if (cx.warp >= kDsCW)
    {
        // ---- producer warps: weights and the cross-KV cache are static data, no dependency on the previous kernel ----
        // (an L2 prefetch lane running 12-48 items ahead of the ring producer was measured: 1-10 % slower, the
        // prefetched lines compete with the ring's own fills; ds_producer<true> is kept for experiments only)
        if (cx.lane == 0 && cx.warp == kDsCW)
            ds_producer<false>(m, cx, 0u);
        return;
    }

    // ---- consumers ----
    grid_dep_wait(); // tokens / lengths / the previous step's cache rows come from earlier kernels on the stream
#if defined(B200_DS_DEBUG)
    if (cx.dbg != nullptr && cx.tid == 0)
        cx.dbg[(size_t) cx.c * kDsDbgPhases * 2] = ds_globaltimer();
#endif
    {
        // token + positional embedding -> x (A-fragment order); fp16 add like embed_kernel (glue.cu)
        const int nkb = m.d >> 6;
        for (int kb = cx.c; kb < nkb; kb += cx.G)
        {
            for (int i = cx.tid; i < 16 * 32; i += kDsConsumers)
            {
                const int row = i >> 5, k = 64 * kb + 2 * (i & 31);
                if (row < m.B)
                {
                    const int tok = min(max(p.tokens[row], 0), p.vocab - 1);
                    const int pos = min(max(p.seq_len[row], 0), p.n_ctx - 1);
                    const __half2 te = *reinterpret_cast<const __half2*>(p.tok_emb + (size_t) tok * m.d + k);
                    const __half2 pe = *reinterpret_cast<const __half2*>(p.pos_emb + (size_t) pos * m.d + k);
                    *reinterpret_cast<__half2*>(p.x + frag_index(row, k)) = __hadd2(te, pe);
                }
            }
        }
        ds_grid_arrive(cx);
    }
    // One call site per phase function (all inlined: no stack frame, no parameter structs in local memory -- the
    // acquire loads of the grid barrier invalidate L1, which would turn every spilled field into an L2 round trip on
    // the critical path).  The layer's 32 pointers are staged in shared memory one layer ahead.
    unsigned long long* ltab = reinterpret_cast<unsigned long long*>(cx.sm.stats + kDsStatFloats) + 2 * kDsSlots + 8;
    if (cx.tid < 32)
        ltab[cx.tid] = __ldg(reinterpret_cast<const unsigned long long*>(m.layers) + cx.tid);
    consumer_sync();
#pragma unroll 1
    for (int l = 0; l < m.L; ++l)
    {
        const unsigned long long* lp = ltab + (l & 1) * 32;
        const bool last = l + 1 == m.L;
        if (!last && cx.tid < 32) // next layer's pointers (consumed after several more consumer_sync()s)
            ltab[((l + 1) & 1) * 32 + cx.tid] = __ldg(reinterpret_cast<const unsigned long long*>(m.layers + l + 1) + cx.tid);
#pragma unroll 1
        for (int sidx = 0; sidx < 8; ++sidx)
        {
            if (sidx == 1)
            {
                DsMmha a{p.qkv, reinterpret_cast<int8_t*>(lp[27]), reinterpret_cast<const float*>(lp[28]),
                    reinterpret_cast<const float*>(lp[29]), p.seq_len, p.ctx};
                ds_mmha_phase(a, m, cx);
            }
            else if (sidx == 4)
            {
                DsXattn a{p.q, reinterpret_cast<const float*>(lp[31]), p.ctx};
                ds_xattn_phase(a, m, cx);
            }
            else
            {
                // matmul g: 0 qkv, 1 attn_out, 2 cross_q, 3 cross_out, 4 fc1, 5 fc2; even g carry a folded LayerNorm
                const int g = sidx == 0 ? 0 : (sidx < 4 ? sidx - 1 : sidx - 2);
                const bool fold = (g & 1) == 0;
                const int wb = 9 * (g >> 1) + (fold ? 1 : 6); // index of the weight pointer in b200_decoder_layer
                DsGemm a;
                a.gamma = fold ? reinterpret_cast<const __half*>(lp[wb - 1]) : nullptr;
                a.W = reinterpret_cast<const int8_t*>(lp[wb]);
                a.scales = reinterpret_cast<const __half*>(lp[wb + 1]);
                a.bias = reinterpret_cast<const __half*>(lp[wb + 2]);
                a.c1s = fold ? reinterpret_cast<const float*>(lp[wb + 3]) : nullptr;
                a.c2 = fold ? reinterpret_cast<const float*>(lp[wb + 4]) : nullptr;
                a.K = g == 5 ? m.dff : m.d;
                a.N = g == 0 ? 3 * m.d : (g == 4 ? m.dff : m.d);
                a.A = (g == 1 || g == 3) ? p.ctx : (g == 5 ? p.u : p.x);
                a.resid = (g & 1) ? p.x : nullptr;
                a.out_frag = (g & 1) ? p.x : (g == 4 ? p.u : nullptr);
                a.out_rm = g == 0 ? p.qkv : (g == 2 ? p.q : ((g == 5 && last) ? p.x_out : nullptr));
                a.act = g == 4 ? B200_ACT_GELU_ERF : B200_ACT_NONE;
                a.which = g;
                a.eps = p.eps;
                ds_gemm_phase(a, m, cx);
            }
        }
    }
    // ---- leave the barrier words zero for the next launch: the last CTA out resets them ----
    if (cx.tid == 0)
    {
        __threadfence(); // this CTA's last arrival on sync[0] is performed before its exit is counted
        const unsigned old = atomicAdd(cx.sync + 1, 1u);
        if (old == (unsigned) cx.G - 1u)
        {
            cx.sync[0] = 0u;
            cx.sync[1] = 0u;
            __threadfence();
        }
    }
}

} // namespace b200

using namespace b200;

static void* g_ds_debug = nullptr;

/* Debug aid: device buffer of n_ctas * 512 * 2 int64 receiving %globaltimer stamps of every following step launch
 * (per CTA and phase: [0] grid wait returned, [1] phase work done); NULL switches it off. */
extern "C" int b200_debug_decoder_step_timeline(void* device_buffer)
{
    g_ds_debug = device_buffer;
    return B200_OK;
}

extern "C" size_t b200_decoder_step_scratch_bytes(int num_heads, int d_ff)
{
    if (num_heads <= 0 || d_ff <= 0)
        return 0;
    const size_t d = (size_t) num_heads * kDh;
    return 256 + 32 * d * 6 + 32 * (size_t) d_ff;
}

extern "C" int b200_decoder_step(const b200_decoder_step_params* p, b200_stream_t stream)
{
    B200_REQUIRE(p != nullptr, B200_ERR_INVALID_ARG, "null params");
    B200_REQUIRE(p->layers && p->tokens && p->sequence_lengths && p->tok_emb && p->pos_emb && p->x_out && p->scratch,
        B200_ERR_INVALID_ARG, "null pointer (layers/tokens/sequence_lengths/tok_emb/pos_emb/x_out/scratch)");
    B200_REQUIRE(p->n_layers > 0 && p->num_heads > 0 && p->batch_size >= 0 && p->max_seq_len > 0 && p->enc_len > 0
            && p->vocab > 0 && p->n_ctx > 0,
        B200_ERR_INVALID_ARG, "bad sizes");
    B200_REQUIRE(p->batch_size <= 16, B200_ERR_UNSUPPORTED, "batch_size %d > 16 rows per step kernel", p->batch_size);
    const int d = p->num_heads * kDh;
    B200_REQUIRE(d <= kDsUnitK, B200_ERR_UNSUPPORTED, "hidden size %d > %d", d, kDsUnitK);
    const int nq = (p->d_ff + kDsUnitK - 1) / kDsUnitK;
    B200_REQUIRE(p->d_ff % 64 == 0 && p->d_ff % (64 * nq) == 0, B200_ERR_UNSUPPORTED,
        "d_ff %d must split into %d equal multiples of 64", p->d_ff, nq);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(p->scratch) & 255) == 0, B200_ERR_INVALID_ARG, "scratch must be 256-byte aligned");
    if (p->batch_size == 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    int G = num_sms();
    if (p->max_ctas > 0 && p->max_ctas < G)
        G = p->max_ctas;
    B200_REQUIRE(p->batch_size * p->num_heads <= kDsMaxPairsPerCta * G, B200_ERR_UNSUPPORTED,
        "%d (batch, head) pairs exceed %d per CTA on %d CTAs", p->batch_size * p->num_heads, kDsMaxPairsPerCta, G);
    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(decoder_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kDsSmemBytes));
        attr_set = true;
    }
    DsParams k{};
    k.m.layers = p->layers;
    k.m.L = p->n_layers, k.m.B = p->batch_size, k.m.H = p->num_heads, k.m.d = d, k.m.dff = p->d_ff;
    k.m.Smax = p->max_seq_len, k.m.S = p->enc_len, k.m.nch = (p->enc_len + kDsChunkKeys - 1) / kDsChunkKeys;
    k.tokens = p->tokens, k.seq_len = p->sequence_lengths;
    k.tok_emb = static_cast<const __half*>(p->tok_emb), k.pos_emb = static_cast<const __half*>(p->pos_emb);
    k.x_out = static_cast<__half*>(p->x_out);
    char* s = static_cast<char*>(p->scratch);
    k.sync = reinterpret_cast<unsigned*>(s);
    s += 256;
    k.x = reinterpret_cast<__half*>(s), s += 32 * (size_t) d;
    k.ctx = reinterpret_cast<__half*>(s), s += 32 * (size_t) d;
    k.u = reinterpret_cast<__half*>(s), s += 32 * (size_t) p->d_ff;
    k.qkv = reinterpret_cast<__half*>(s), s += 96 * (size_t) d;
    k.q = reinterpret_cast<__half*>(s);
    k.vocab = p->vocab, k.n_ctx = p->n_ctx, k.eps = p->ln_eps;
    k.dbg = static_cast<long long*>(g_ds_debug);
    k.l2_ahead = 0;
    B200_LAUNCH(decoder_step_kernel, dim3(G), dim3(kDsThreads), kDsSmemBytes, as_stream(stream), k);
    return B200_OK;
}

extern "C" int b200_decoder_step_status(const void* scratch, int32_t* status_host)
{
    B200_REQUIRE(scratch && status_host, B200_ERR_INVALID_ARG, "null pointer");
    B200_REQUIRE_DEVICE();
    unsigned w[4] = {0, 0, 0, 0};
    B200_CUDA(cudaDeviceSynchronize());
    B200_CUDA(cudaMemcpy(w, scratch, sizeof(w), cudaMemcpyDeviceToHost));
    *status_host = (int32_t) w[2];
    return B200_OK;
}
