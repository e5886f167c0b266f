// decoder_step.cu -- ONE persistent kernel for a whole generation step of the quantized Whisper decoder stack
// (b200_decoder_step, include/b200_whisper.h).
//
// Why: with one kernel per operator the batch-16 step is 8 dependent launches per layer x 32 layers, each costing
// 3-5 us of dependency latency for 0.25-1 us of HBM time (profiles/r01_step_timeline_v7.txt): the step ran at a third
// of the HBM roofline although every kernel moved exactly its algorithmic bytes.  The reference has the same shape
// (one TensorRT layer / plugin enqueue per operator: weightOnlyQuantMatmulPlugin.cpp:162-222,
// gptAttentionCommon.cpp:649-780).  Here the dependency chain stays -- it is the model -- but everything that does
// not depend on it is taken off it, and the hops of the chain themselves are made as short as the memory system allows:
//
//   * one CTA per SM, resident for the whole step; warp 10 (one lane) is a TMA producer that walks a STATIC schedule
//     of 10 KB items -- int8 weight tiles (8 output columns x 1280 k in the reference's preprocessed layout = four
//     contiguous 2560-byte row-pair segments) and cross-KV chunks (80 keys: 5 KB of K + 5 KB of V) -- through a
//     19-slot shared-memory ring with cp.async.bulk + mbarrier complete_tx.  Weights and the cross-KV cache never
//     depend on activations, so the producer runs up to 190 KB per SM (28 MB per chip, more than a layer's weights)
//     ahead of the consumers: when a phase's activations arrive its weights are already in shared memory.
//   * NO grid barrier between the phases.  Every activation that crosses CTAs lives in global memory as "flagged
//     words": 8 bytes = one half2 (or one fp32 partial sum) + the 32-bit GENERATION of the phase that wrote it, stored
//     and loaded as ONE 64-bit access (single-copy atomic), relaxed, GPU scope.  A consumer simply loads the words it
//     needs and retries until every flag carries the generation it expects: no fence, no release / acquire, no
//     counter, no waiting for CTAs whose data it does not read.  One hop of the dependency chain = one store reaching
//     L2 + one load from L2 (the grid barrier this replaces cost a bar.sync + MEMBAR.GPU + RED + poll = ~1.9 us per
//     phase, measured; profiles/r02_step_phases_barrier.txt).  Generations are unique per (step, layer, phase): the
//     step counter lives in the scratch area and is bumped by the last CTA to leave; buffers are never zeroed again.
//     Progress: all CTAs are co-resident and walk the phases in the same order, a CTA only waits for data of EARLIER
//     phases, so the wait graph is acyclic.  Every wait is bounded (a stuck wait sets the status word and the kernel
//     drains instead of hanging the GPU).
//   * warps 0-9 consume.  Matmuls (16 batch rows): a unit of work is one tile of 8 output columns over <= 1280 k; every
//     warp owns two 64-wide k-blocks of the unit, converts the biased int8 bytes to fp16 in registers (PRMT + HSUB2,
//     exact integers; the reference layout's row permutation and byte swizzle make each converted word the B fragment
//     of an mma.sync.m16n8k16), multiplies by the LayerNorm gamma pair when the LayerNorm is folded in, and accumulates
//     in fp32 with the batch rows as the MMA's M.  Activations are kept in "A-fragment order", so a warp's A operands
//     are contiguous flagged words straight from L2 -- no shared-memory staging of activations.  The ten k-partials
//     are reduced through shared memory in fixed warp order (deterministic); the epilogue applies the column scale, the
//     folded-LayerNorm correction (row statistics from the A fragments: sums of differences from the row's first
//     element), bias / GELU / residual with the per-layer fp16 rounding of the per-operator kernels.  Deep-K matmuls
//     (fc2, K = 5120) split K over groups of 4 CTAs: partial sums travel as flagged fp32 words and the tile's finisher
//     adds the four quarters in fixed order -- a CTA never ingests more than 16 x 1280 activations per phase.
//   * self-attention: (batch, head) pairs dealt to CTAs, 3 warps per pair over alternating groups of 8 keys, first
//     pass of the int8 cache fetched before q arrives; same arithmetic as mmha_generation_kernel.
//   * cross-attention: whole (batch, head) pairs per CTA, chunks dealt round-robin to the ten warps straight from the
//     ring, merge through shared memory; same inner loop as cross_attention_rowhead_kernel (attn_device.cuh).
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
constexpr int kDsMaxSplit = 4;            // deepest K = 4 units (d_ff = 4 d)
constexpr int kDsRoundTiles = 5;          // tiles (8 columns each) per reduction round
constexpr int kDsChunkKeys = 80;          // keys per cross-KV chunk: 5120 B of K + 5120 B of V
constexpr int kDsScratchFloats = kDsCW * kDsRoundTiles * 16 * 8; // 25600 B: k-partials / attention merge area
constexpr int kDsStatFloats = kDsCW * 16 * 2 + 16;               // per (warp, row): sum d, sum d^2; + the 16 row shifts
constexpr int kDsPart = kDh + 4;          // attention partial: m, l, 2 pad, o[64]
constexpr size_t kDsSmemBytes = (size_t) kDsSlots * kDsSlotBytes + sizeof(float) * (kDsScratchFloats + kDsStatFloats)
    + sizeof(uint64_t) * 2 * kDsSlots + 64 + 2 * 32 * sizeof(uint64_t);
constexpr long long kDsWaitCycles = 3000000000ll; // SM cycles before a wait gives up (~1.5 s; a step takes ~1 ms)
constexpr int kDsMmhaWarpsPerPair = 3;
constexpr int kDsMaxPairsPerCta = kDsCW / kDsMmhaWarpsPerPair;
constexpr unsigned kDsGenPerStep = 512;   // generations reserved per step: 1 + 8 per layer (n_layers <= 63)

enum
{
    DS_ERR_DATA_WAIT = 1,
    DS_ERR_RING_FULL_WAIT = 2,
    DS_ERR_RING_EMPTY_WAIT = 3
};

// element (row, k) of a 16-row activation matrix in A-fragment order (index in halves); the flagged word of the pair
// (k, k + 1), k even, is word frag_index(row, k) / 2
__host__ __device__ __forceinline__ int frag_index(int row, int k)
{
    const int kb = k >> 6, kk = k & 63;
    const int T = kk >> 4, r = kk & 15;
    const int hi = r >> 3, w = (r & 7) >> 1, e = r & 1;
    const int g = row & 7, up = row >> 3;
    return ((((kb * 4 + w) * 32 + 4 * g + T) * 4 + 2 * hi + up) << 1) + e;
}

// ---- flagged words ------------------------------------------------------------------------------------------------
// low 32 bits: payload (half2 or fp32), high 32 bits: generation.  64-bit scalar accesses: single-copy atomic.
__device__ __forceinline__ uint2 ll_ld1(const uint2* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return make_uint2((uint32_t) v, (uint32_t) (v >> 32));
}

__device__ __forceinline__ void ll_ld2(const uint2* p, uint2& a, uint2& b) // p 16-byte aligned; two independent words
{
    unsigned long long x, y;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
    a = make_uint2((uint32_t) x, (uint32_t) (x >> 32));
    b = make_uint2((uint32_t) y, (uint32_t) (y >> 32));
}

__device__ __forceinline__ void ll_st1(uint2* p, uint32_t payload, uint32_t gen)
{
    const unsigned long long v = (unsigned long long) payload | ((unsigned long long) gen << 32);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void ll_st2(uint2* p, uint32_t pay0, uint32_t pay1, uint32_t gen)
{
    const unsigned long long x = (unsigned long long) pay0 | ((unsigned long long) gen << 32);
    const unsigned long long y = (unsigned long long) pay1 | ((unsigned long long) gen << 32);
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
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
    uint32_t dead; // shared-space address of the CTA-local "a wait timed out, stop waiting" flag
};

struct DsCtx
{
    DsShared sm;
    unsigned* sync;    // [1] exits, [2] status, [3] step counter
    long long* dbg;    // optional %globaltimer stamps [CTA][phase][2] (inputs arrived, work done); tools/step_phases.py
    unsigned phase;    // phases finished so far (debug stamps only)
    unsigned item;     // ring items consumed so far by this CTA
    unsigned base;     // first generation of this step
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

// Bounded spinning: returns true when the caller should stop waiting (this CTA or another one timed out).
struct DsSpin
{
    long long t0 = 0;
    unsigned n = 0;
};

__device__ __noinline__ bool ds_spin_check(DsCtx& cx, DsSpin& s, int code)
{
    if (ds_dead(cx.sm))
        return true;
    if (s.t0 == 0)
        s.t0 = clock64();
    if (clock64() - s.t0 > kDsWaitCycles || ld_relaxed_u32(cx.sync + 2) != 0u)
    {
        ds_set_dead(cx.sm);
        atomicCAS(cx.sync + 2, 0u, (unsigned) code | ((unsigned) cx.c << 8) | (cx.phase << 16));
        return true;
    }
    return false;
}

__device__ __forceinline__ bool ds_spin(DsCtx& cx, DsSpin& s, int code)
{
    return ((++s.n & 1023u) == 0u) && ds_spin_check(cx, s, code);
}

// bounded mbarrier wait (a lost TMA completion or a schedule mismatch must not hang the GPU)
__device__ __forceinline__ void ds_mbar_wait(DsCtx& cx, uint64_t* bar, uint32_t parity, int code)
{
    if (mbar_try_wait(bar, parity))
        return;
    DsSpin s;
    while (!mbar_try_wait(bar, parity))
        if (ds_spin(cx, s, code))
            return;
}

// ---- debug stamps ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long ds_globaltimer()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

[[maybe_unused]] constexpr int kDsDbgPhases = 512; // stamps per CTA in the debug buffer

// fine-grained stamps of CTA 0 inside a phase (debug buffer region after the per-CTA stamps)
__device__ __forceinline__ void ds_stamp(DsCtx& cx, int slot)
{
#if defined(B200_DS_DEBUG)
    if (cx.dbg != nullptr && cx.c == 0 && cx.tid == 0 && cx.phase < kDsDbgPhases)
        cx.dbg[(size_t) cx.G * kDsDbgPhases * 2 + (size_t) cx.phase * 16 + slot] = clock64(); // SM cycles: fine resolution
#endif
}

// per-CTA, per-phase: [0] the phase's inputs have arrived, [1] its work is done
__device__ __forceinline__ void ds_phase_stamp(DsCtx& cx, int which)
{
#if defined(B200_DS_DEBUG)
    if (cx.dbg != nullptr && cx.tid == 0 && cx.phase < kDsDbgPhases)
        cx.dbg[((size_t) cx.c * kDsDbgPhases + cx.phase) * 2 + which] = ds_globaltimer();
#endif
}

// ---- static schedule (shared by the producer and the consumers) ------------------------------------------------------
// A matmul's units: tile t (8 output columns) x k-unit q (KU <= 1280 k).  K <= 1280: every CTA takes the tiles
// grp, grp + G, ... of the rotated CTA index grp.  Deeper K: CTAs work in groups of nq; the group takes the tiles
// grp, grp + NG, ... and member q multiplies k-unit q of each of them.
struct DsSched
{
    int nq, KU, NG, grp, q, nt;
};

__device__ __forceinline__ int ds_rot(int which, int G)
{
    // qkv, attn_out, cross_q, cross_out, fc1, fc2: spread the CTAs that receive an extra tile over the grid, away from
    // the low CTA indices, which hold an extra (batch, head) pair in the attention phases
    const int r[6] = {72, 24, 36, 48, 108, 60};
    return r[which] % G;
}

__device__ __forceinline__ DsSched ds_sched(int K, int N, int which, int c, int G)
{
    DsSched s;
    s.nq = (K + kDsUnitK - 1) / kDsUnitK;
    s.KU = K / s.nq;
    const int NT = N >> 3;
    if (s.nq == 1)
    {
        s.NG = G;
        s.grp = ((c - ds_rot(which, G)) % G + G) % G;
        s.q = 0;
    }
    else
    {
        s.NG = G / s.nq;
        s.grp = c / s.nq;
        s.q = c - s.grp * s.nq;
    }
    s.nt = (s.grp < s.NG && s.grp < NT) ? (NT - s.grp + s.NG - 1) / s.NG : 0;
    return s;
}

struct DsModel
{
    const b200_decoder_layer* layers;
    int L, B, H, d, dff, Smax, S, nch;
};

// ---- producer: one lane walks the whole step's static schedule, at most kDsSlots items ahead of the consumers --------
__device__ __forceinline__ void ds_producer(const DsModel& m, DsCtx& cx)
{
    const uint64_t pol = policy_evict_first();
    unsigned it = 0;
    // returns false when the kernel is draining after a timed-out wait
    auto begin_item = [&](uint32_t bytes, uint8_t*& dst, uint64_t*& bar) -> bool
    {
        const unsigned s = it % kDsSlots, use = it / kDsSlots;
        if (use > 0)
            ds_mbar_wait(cx, &cx.sm.empty[s], (use - 1) & 1, DS_ERR_RING_EMPTY_WAIT);
        if (ds_dead(cx.sm))
            return false;
        mbar_arrive_expect_tx(&cx.sm.full[s], bytes);
        ++it;
        dst = cx.sm.ring + (size_t) s * kDsSlotBytes, bar = &cx.sm.full[s];
        return true;
    };
    auto gemm = [&](const int8_t* W, int K, int N, int which) -> bool
    {
        const DsSched g = ds_sched(K, N, which, cx.c, cx.G);
        const uint8_t* base = reinterpret_cast<const uint8_t*>(W);
        for (int j = 0; j < g.nt; ++j)
        {
            const int tile = g.grp + j * g.NG;
            uint8_t* dst;
            uint64_t* bar;
            if (!begin_item((uint32_t) 8 * g.KU, dst, bar))
                return false;
#pragma unroll
            for (int rp = 0; rp < 4; ++rp)
                bulk_g2s_hint(dst + (size_t) rp * 2 * g.KU, base + (size_t) (4 * tile + rp) * 2 * K + (size_t) g.q * 2 * g.KU,
                    (uint32_t) 2 * g.KU, bar, pol);
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
                bulk_g2s_hint(dst, kb + (size_t) key0 * kDh, bytes, bar, pol);
                bulk_g2s_hint(dst + kDsSlotBytes / 2, vb + (size_t) key0 * kDh, bytes, bar, pol);
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
    const uint2* A;      // flagged words, A-fragment order, 16 rows x K
    const uint2* resid;  // flagged words, A-fragment order, 16 rows x N, or null
    uint2* out_frag;     // flagged words, A-fragment order, 16 rows x N, or null
    uint2* out_rm;       // flagged words, row-major [rows][N / 2], or null
    __half* out_plain;   // plain fp16 row-major [rows][N] (the step's result), or null
    uint2* part;         // split-K partial sums: flagged fp32 words [nq][N / 8][16 x 8]
    uint32_t genA, genR, genO;
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

// A fragments (4 MMAs' worth) of k-block `kb`: this lane's 16 flagged words.  Rows >= B are never written by anybody:
// their flags are ignored and their payload replaced by zeros.  Returns whether every needed word carries `gen`.
__device__ __forceinline__ bool ds_load_a(const uint2* A, int kb, int lane, uint32_t gen, bool v0, bool v1, uint4 (&dst)[4])
{
    const uint2* ap = A + ((size_t) (kb * 4) * 32 + lane) * 4;
    bool ok = true;
#pragma unroll
    for (int w = 0; w < 4; ++w)
    {
        uint2 a, b, c, d;
        ll_ld2(ap + (size_t) w * 128, a, b);
        ll_ld2(ap + (size_t) w * 128 + 2, c, d);
        ok = ok && (a.y == gen || !v0) && (b.y == gen || !v1) && (c.y == gen || !v0) && (d.y == gen || !v1);
        dst[w] = make_uint4(v0 ? a.x : 0u, v1 ? b.x : 0u, v0 ? c.x : 0u, v1 ? d.x : 0u);
    }
    return ok;
}

__device__ __forceinline__ uint32_t ds_sel4(const uint4& v, int w) // w is a compile-time constant after unrolling
{
    return w == 0 ? v.x : (w == 1 ? v.y : (w == 2 ? v.z : v.w));
}

__device__ __forceinline__ __half2 ds_u2h2(uint32_t u)
{
    return *reinterpret_cast<__half2*>(&u);
}

// D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col); not volatile: a pure function of its operands, so the
// compiler may interleave the independent chains of different tiles and hoist the dequant of the next operand
__device__ __forceinline__ void ds_mma(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// This warp's two k-blocks of NT weight tiles (ring slots wp[0..NT)): 128-bit shared-memory reads of this lane's column
// / 16-k chunk, PRMT + HSUB2 dequant, optional gamma pair, MMAs with the tiles' accumulator chains interleaved; the
// k-partials go to the warp's rows of the scratch area.
template <int NT, bool FOLD>
__device__ __forceinline__ void ds_mma_tiles(const uint4 (&af)[2][4], bool kv0, bool kv1, const uint8_t* const (&wp)[2], uint32_t off0,
    uint32_t off1, const uint4 (&glo)[2], const uint4 (&ghi)[2], float* scr, int g, int t)
{
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int kbi = 0; kbi < 2; ++kbi)
    {
        if (kbi == 0 ? kv0 : kv1)
        {
            uint4 wv[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j)
                wv[j] = *reinterpret_cast<const uint4*>(wp[j] + (kbi == 0 ? off0 : off1));
#pragma unroll
            for (int w = 0; w < 4; ++w)
            {
#pragma unroll
                for (int j = 0; j < NT; ++j)
                {
                    __half2 lo, hi;
                    dequant_word(ds_sel4(wv[j], w), lo, hi);
                    if constexpr (FOLD)
                    {
                        lo = __hmul2(lo, ds_u2h2(ds_sel4(glo[kbi], w)));
                        hi = __hmul2(hi, ds_u2h2(ds_sel4(ghi[kbi], w)));
                    }
                    ds_mma(acc[j], af[kbi][w], h2u(lo), h2u(hi));
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        *reinterpret_cast<float2*>(scr + (size_t) j * 128 + g * 8 + 2 * t) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2*>(scr + (size_t) j * 128 + (g + 8) * 8 + 2 * t) = make_float2(acc[j][2], acc[j][3]);
    }
}

__device__ __forceinline__ void ds_gemm_phase(const DsGemm& a, const DsModel& m, DsCtx& cx)
{
    const int lane = cx.lane, warp = cx.warp, tid = cx.tid;
    const int g = lane >> 2, t = lane & 3;
    const DsSched sc = ds_sched(a.K, a.N, a.which, cx.c, cx.G);
    const int nkbu = sc.KU >> 6, NT = a.N >> 3;
    const bool fold = a.gamma != nullptr; // only with nq == 1 (hidden size <= 1280, checked on the host)
    // this warp's k-blocks inside the unit: warp, warp + 10
    const int kbl0 = warp, kbl1 = warp + kDsCW;
    const bool kv0 = kbl0 < nkbu, kv1 = kbl1 < nkbu;
    const bool v0 = g < m.B, v1 = g + 8 < m.B;
    // byte offset of this lane's 16 bytes inside a k-block of a weight unit: column g of the tile (row pair g/2, parity
    // g%2), 16-k chunk t
    const uint32_t lane_off = (uint32_t) ((g >> 1) * 2 * sc.KU + (g & 1) * 64 + t * 16);
    const uint32_t off0 = lane_off + kbl0 * 128, off1 = lane_off + kbl1 * 128;

    // ---- static operands (HBM round trips hidden behind the wait for the activations) ----
    // epilogue item of this thread in a round: (tile jj, row, column pair cp)
    const int e_jj = tid >> 6, e_row = (tid & 63) >> 2, e_cp = tid & 3;
    float e_sc[2] = {0.f, 0.f}, e_bias[2] = {0.f, 0.f}, e_c1[2] = {0.f, 0.f}, e_c2[2] = {0.f, 0.f};
    auto load_epi_static = [&](int j0)
    {
        if (e_jj < min(kDsRoundTiles, sc.nt - j0))
        {
            const int n0 = 8 * (sc.grp + (j0 + e_jj) * sc.NG) + 2 * e_cp;
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

    // the first round's weight tiles have been in shared memory for a while: take their barrier waits (~90 cycles each)
    // off the critical path that starts when the activations arrive
    if (kv0)
        for (int jj = 0; jj < min(kDsRoundTiles, sc.nt); ++jj)
        {
            const unsigned it = cx.item + (unsigned) jj;
            ds_mbar_wait(cx, &cx.sm.full[it % kDsSlots], (it / kDsSlots) & 1, DS_ERR_RING_FULL_WAIT);
        }
    // the residual words of this thread's first epilogue item: written two phases ago and consumed in full by every CTA
    // since, so they are there -- one L2 round trip hidden behind the wait for A (re-read in the epilogue if not)
    uint2 e_res = make_uint2(0u, 0u);
    if (a.resid != nullptr && e_jj < min(kDsRoundTiles, sc.nt) && e_row < m.B)
        e_res = ll_ld1(a.resid + (frag_index(e_row, 8 * (sc.grp + e_jj * sc.NG) + 2 * e_cp) >> 1));

    // ---- the activations: this warp's two k-blocks of A (and, for a folded LayerNorm, the rows' first elements) ----
    uint4 af[2][4];
    uint32_t shw0 = 0u, shw1 = 0u; // half2 words (row g, k 0..1), (row g + 8, k 0..1)
    if (sc.nt > 0 && kv0)
    {
        const int kbA0 = sc.q * nkbu + kbl0, kbA1 = sc.q * nkbu + kbl1;
        if (lane == 0)
        {
            // cheap gate: one word per warp until the first of its producers has delivered, then the full set
            DsSpin sp;
            const uint2* gate = a.A + (size_t) (kbA0 * 4) * 32 * 4;
            while (ll_ld1(gate).y != a.genA)
                if (ds_spin(cx, sp, DS_ERR_DATA_WAIT))
                    break;
        }
        __syncwarp();
        DsSpin sp;
        for (;;)
        {
            bool ok = ds_load_a(a.A, kbA0, lane, a.genA, v0, v1, af[0]);
            if (kv1)
                ok = ds_load_a(a.A, kbA1, lane, a.genA, v0, v1, af[1]) && ok;
            if (fold)
            {
                uint2 s0, s1;
                ll_ld2(a.A + (size_t) (4 * g) * 4, s0, s1);
                ok = ok && (s0.y == a.genA || !v0) && (s1.y == a.genA || !v1);
                shw0 = v0 ? s0.x : 0u, shw1 = v1 ? s1.x : 0u;
            }
            const unsigned bad = __ballot_sync(0xffffffffu, !ok);
            if (bad == 0u)
                break;
            // Not all there yet.  ONE lane (the first with a missing word) keeps polling its own 16 words -- 512 bytes
            // per round trip instead of the warp's 16 KB: a thousand waiting warps must not saturate L2 with polls
            // while the producers' stores and the weight streams need it -- then the warp loads the whole set again.
            bool stop = false;
            if (lane == __ffs(bad) - 1)
            {
                uint4 tmp[4];
                for (;;)
                {
                    bool k = ds_load_a(a.A, kbA0, lane, a.genA, v0, v1, tmp);
                    if (kv1)
                        k = ds_load_a(a.A, kbA1, lane, a.genA, v0, v1, tmp) && k;
                    if (k)
                        break;
                    if (ds_spin(cx, sp, DS_ERR_DATA_WAIT))
                    {
                        stop = true;
                        break;
                    }
                }
            }
            if (__any_sync(0xffffffffu, stop))
                break;
        }
    }
    ds_phase_stamp(cx, 0);
    ds_stamp(cx, 0);

    for (int j0 = 0; j0 < sc.nt; j0 += kDsRoundTiles)
    {
        const int rt = min(kDsRoundTiles, sc.nt - j0);
        float* scr = cx.sm.scratch + (size_t) (warp * kDsRoundTiles) * 128;
        if (kv0)
        {
#pragma unroll 1
            for (int jj = 0; jj < rt; jj += 2)
            {
                const bool two = jj + 1 < rt;
                const unsigned it0 = cx.item + (unsigned) jj, it1 = it0 + (two ? 1u : 0u);
                if (j0 > 0) // (round 0 was waited for before the activations)
                {
                    ds_mbar_wait(cx, &cx.sm.full[it0 % kDsSlots], (it0 / kDsSlots) & 1, DS_ERR_RING_FULL_WAIT);
                    if (two)
                        ds_mbar_wait(cx, &cx.sm.full[it1 % kDsSlots], (it1 / kDsSlots) & 1, DS_ERR_RING_FULL_WAIT);
                }
                const uint8_t* const wp[2] = {cx.sm.ring + (size_t) (it0 % kDsSlots) * kDsSlotBytes,
                    cx.sm.ring + (size_t) (it1 % kDsSlots) * kDsSlotBytes};
                if (jj == 0)
                    ds_stamp(cx, 4);
                if (two)
                {
                    if (fold)
                        ds_mma_tiles<2, true>(af, kv0, kv1, wp, off0, off1, glo, ghi, scr + (size_t) jj * 128, g, t);
                    else
                        ds_mma_tiles<2, false>(af, kv0, kv1, wp, off0, off1, glo, ghi, scr + (size_t) jj * 128, g, t);
                }
                else
                {
                    if (fold)
                        ds_mma_tiles<1, true>(af, kv0, kv1, wp, off0, off1, glo, ghi, scr + (size_t) jj * 128, g, t);
                    else
                        ds_mma_tiles<1, false>(af, kv0, kv1, wp, off0, off1, glo, ghi, scr + (size_t) jj * 128, g, t);
                }
            }
        }
        else
        {
            // a warp without a k-block of this (short) unit contributes zeros
            for (int i = lane; i < rt * 128; i += 32)
                scr[i] = 0.f;
        }
        ds_stamp(cx, 5);
        if (fold && j0 == 0)
        {
            // LayerNorm statistics of rows g and g + 8 over this warp's k-blocks: sums of (x - x[row][0]) and of its
            // square in fp32 (the shift keeps the cancellation in var = E[d^2] - E[d]^2 harmless), reduced over the four
            // lanes that share a row; merged over the warps by plain addition in the epilogue
            float sd0 = 0.f, sq0 = 0.f, sd1 = 0.f, sq1 = 0.f;
            const float sh0 = __low2float(ds_u2h2(shw0)), sh1 = __low2float(ds_u2h2(shw1));
            if (kv0)
            {
#pragma unroll
                for (int kbi = 0; kbi < 2; ++kbi)
                {
                    if (kbi == 1 && !kv1)
                        break;
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                    {
                        const uint32_t r0[2] = {af[kbi][w].x, af[kbi][w].z};
                        const uint32_t r1[2] = {af[kbi][w].y, af[kbi][w].w};
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                        {
                            const float2 f0 = __half22float2(ds_u2h2(r0[i])), f1 = __half22float2(ds_u2h2(r1[i]));
                            const float d0 = f0.x - sh0, d1 = f0.y - sh0, d2 = f1.x - sh1, d3 = f1.y - sh1;
                            sd0 += d0 + d1;
                            sq0 = fmaf(d0, d0, fmaf(d1, d1, sq0));
                            sd1 += d2 + d3;
                            sq1 = fmaf(d2, d2, fmaf(d3, d3, sq1));
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 1; o < 4; o <<= 1)
            {
                sd0 += __shfl_xor_sync(0xffffffffu, sd0, o);
                sq0 += __shfl_xor_sync(0xffffffffu, sq0, o);
                sd1 += __shfl_xor_sync(0xffffffffu, sd1, o);
                sq1 += __shfl_xor_sync(0xffffffffu, sq1, o);
            }
            if (t == 0)
            {
                float* sp = cx.sm.stats + (size_t) warp * 32;
                *reinterpret_cast<float2*>(sp + 2 * g) = make_float2(sd0, sq0);
                *reinterpret_cast<float2*>(sp + 2 * (g + 8)) = make_float2(sd1, sq1);
                if (warp == 0)
                {
                    cx.sm.stats[kDsCW * 32 + g] = sh0;
                    cx.sm.stats[kDsCW * 32 + g + 8] = sh1;
                }
            }
        }
        ds_stamp(cx, 1);
        if (j0 > 0)
            load_epi_static(j0); // later rounds (small grids only): in flight under the reduction
        consumer_sync();
        ds_stamp(cx, 2);
        if (tid == 0)
        {
            // the round's weight slots are drained: hand them back to the producer
            for (int u = 0; u < rt; ++u)
                mbar_arrive(&cx.sm.empty[(cx.item + (unsigned) u) % kDsSlots]);
        }
        cx.item += (unsigned) rt;
        // ---- reduce the ten k-partials in warp order, epilogue ----
        if (e_jj < rt && e_row < m.B)
        {
            const int j = j0 + e_jj;
            const int tile = sc.grp + j * sc.NG;
            const int n0 = 8 * tile + 2 * e_cp;
            float s0 = 0.f, s1 = 0.f;
            const float* src = cx.sm.scratch + ((size_t) e_jj * 16 + e_row) * 8 + 2 * e_cp;
#pragma unroll
            for (int w = 0; w < kDsCW; ++w)
            {
                const float2 v = *reinterpret_cast<const float2*>(src + (size_t) w * kDsRoundTiles * 128);
                s0 += v.x;
                s1 += v.y;
            }
            bool finish = true;
            if (sc.nq > 1)
            {
                // split K: the quarters of a tile meet in global memory as flagged fp32 words; member (j mod nq) of
                // the group adds them in the order 0, 1, .. nq-1 (whoever it is: bit-reproducible) and finishes the tile
                const size_t widx = (size_t) (e_row * 4 + e_cp) * 2;
                if (j % sc.nq != sc.q)
                {
                    ll_st2(a.part + ((size_t) sc.q * NT + tile) * 128 + widx, __float_as_uint(s0), __float_as_uint(s1), a.genO);
                    finish = false;
                }
                else
                {
                    float a0 = 0.f, a1 = 0.f;
                    for (int qq = 0; qq < sc.nq; ++qq)
                    {
                        if (qq == sc.q)
                        {
                            a0 += s0, a1 += s1;
                            continue;
                        }
                        const uint2* pp = a.part + ((size_t) qq * NT + tile) * 128 + widx;
                        uint2 p0, p1;
                        DsSpin sp;
                        for (;;)
                        {
                            ll_ld2(pp, p0, p1);
                            if ((p0.y == a.genO && p1.y == a.genO) || ds_spin(cx, sp, DS_ERR_DATA_WAIT))
                                break;
                        }
                        a0 += __uint_as_float(p0.x), a1 += __uint_as_float(p1.x);
                    }
                    s0 = a0, s1 = a1;
                }
            }
            if (finish)
            {
                float v0f = s0 * e_sc[0], v1f = s1 * e_sc[1];
                if (fold)
                {
                    float sd = 0.f, sq = 0.f;
#pragma unroll
                    for (int w = 0; w < kDsCW; ++w)
                    {
                        const float2 p = *reinterpret_cast<const float2*>(cx.sm.stats + (size_t) w * 32 + 2 * e_row);
                        sd += p.x;
                        sq += p.y;
                    }
                    const float rk = 1.f / (float) a.K;
                    const float md = sd * rk;
                    const float mean = cx.sm.stats[kDsCW * 32 + e_row] + md;
                    const float var = fmaxf(sq * rk - md * md, 0.f);
                    const float rstd = rsqrtf(var + a.eps);
                    v0f = rstd * (v0f - mean * e_c1[0]) + e_c2[0];
                    v1f = rstd * (v1f - mean * e_c1[1]) + e_c2[1];
                }
                const int fw = frag_index(e_row, n0) >> 1;
                float r0 = 0.f, r1 = 0.f;
                const bool hr = a.resid != nullptr;
                if (hr)
                {
                    // written two phases ago and consumed in full by every CTA since: there already (the check is free)
                    uint2 rw = e_res;
                    DsSpin sp;
                    while ((j0 > 0 || rw.y != a.genR) && !ds_spin(cx, sp, DS_ERR_DATA_WAIT))
                    {
                        rw = ll_ld1(a.resid + fw);
                        if (rw.y == a.genR)
                            break;
                    }
                    const float2 r2 = __half22float2(ds_u2h2(rw.x));
                    r0 = r2.x, r1 = r2.y;
                }
                const bool hb = a.bias != nullptr;
                const __half2 o2 = __halves2half2(ds_finish(v0f, hb, e_bias[0], a.act, hr, r0), ds_finish(v1f, hb, e_bias[1], a.act, hr, r1));
                if (a.out_frag != nullptr)
                    ll_st1(a.out_frag + fw, h2u(o2), a.genO);
                if (a.out_rm != nullptr)
                    ll_st1(a.out_rm + (((size_t) e_row * a.N + n0) >> 1), h2u(o2), a.genO);
                if (a.out_plain != nullptr)
                    *reinterpret_cast<__half2*>(a.out_plain + (size_t) e_row * a.N + n0) = o2;
            }
        }
        consumer_sync(); // the scratch / statistics areas are reused by the next round and by the next phase
    }
    ds_stamp(cx, 3);
    ds_phase_stamp(cx, 1);
    ++cx.phase;
}

// 16 consecutive halves (8 flagged words, 64-byte aligned) -> x; returns whether all carry `gen`
__device__ __forceinline__ bool ds_load16h(const uint2* p, uint32_t gen, __half (&x)[16])
{
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        uint2 a, b;
        ll_ld2(p + 2 * i, a, b);
        ok = ok && a.y == gen && b.y == gen;
        *reinterpret_cast<uint32_t*>(&x[4 * i]) = a.x;
        *reinterpret_cast<uint32_t*>(&x[4 * i + 2]) = b.x;
    }
    return ok;
}

// ---- masked self-attention (generation), int8 KV cache ------------------------------------------------------------
// Same arithmetic as mmha_generation_kernel (attention.cu): cached keys as exact fp16 integers, q.k in HFMA2 chains of 4,
// fp32 softmax with 1/(sum + 1e-6), p.v in HFMA2 chains of <= 4 keys flushed to fp32, the current token's k / v
// unquantized, the new K / V quantized with cvt.rni.sat and appended.
struct DsMmha
{
    const uint2* qkv; // flagged words, row-major [B][3d / 2]
    int8_t* cache;
    const float* s_oq;
    const float* s_qo;
    const int* seq_len;
    uint2* ctx_frag;
    uint32_t genI, genO;
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
    // raw cache bytes of one pass (16 dims of NIT keys per lane for K and for V); converted to fp16 at the point of use
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
    __half qh[16], kh[16], vh[16];
    __half2 vcur_h = __float2half2_rn(0.f); // the finalising warp's two dims of this step's v
    if (active)
    {
        // the cache rows below the length, the length and the scales were written by earlier launches: fetch them
        // while q is still on its way
        tlen = min(a.seq_len[b], m.Smax - 1);
        s_qo = __ldg(a.s_qo);
        s_oq = __ldg(a.s_oq);
        fetch(0);
        const uint2* qp = a.qkv + (((size_t) b * 3 * hidden + h * kDh + chunk * 16) >> 1);
        if (lane == 0)
        {
            DsSpin sp;
            while (ll_ld1(qp).y != a.genI)
                if (ds_spin(cx, sp, DS_ERR_DATA_WAIT))
                    break;
        }
        __syncwarp();
        DsSpin sp;
        for (;;)
        {
            bool ok = ds_load16h(qp, a.genI, qh);
            if (wi == 0)
            {
                ok = ds_load16h(qp + (hidden >> 1), a.genI, kh) && ok;
                ok = ds_load16h(qp + hidden, a.genI, vh) && ok;
                const uint2 vw = ll_ld1(a.qkv + (((size_t) b * 3 * hidden + 2 * hidden + h * kDh + 2 * lane) >> 1));
                ok = ok && vw.y == a.genI;
                vcur_h = ds_u2h2(vw.x);
            }
            if (__all_sync(0xffffffffu, ok) || ds_spin(cx, sp, DS_ERR_DATA_WAIT))
                break;
        }
    }
    ds_phase_stamp(cx, 0);

    float m_run = -FLT_MAX, l_run = 0.f, s_cur = -FLT_MAX;
    float o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        o[i] = 0.f;
    const float inv_sqrt_dh = 0.125f; // 1 / sqrt(64), q_scaling = 1 (gptAttentionCommon.cpp:163)
    if (active)
    {
        const float sscale = s_qo * inv_sqrt_dh;
        if (wi == 0)
        {
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
        ll_st1(a.ctx_frag + (frag_index(b, h * kDh + 2 * lane) >> 1), h2u(o2), a.genO);
    }
    consumer_sync(); // the merge area is reused by the next phase
    ds_phase_stamp(cx, 1);
    ++cx.phase;
}

// ---- cross-attention over the int8 cross-KV cache, chunks from the ring -----------------------------------------
struct DsXattn
{
    const uint2* q; // flagged words, row-major [B][d / 2]
    const float* s_qo;
    uint2* ctx_frag;
    uint32_t genI, genO;
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
#if defined(B200_DS_DEBUG)
    long long dbg_wait = 0, dbg_chunks = 0;
    const long long dbg_t0 = clock64();
#endif
    for (int r = 0; r < nbh; ++r)
    {
        const int p = cx.c + r * cx.G;
        const int b = p / m.H, h = p - b * m.H;
        __half qh[16];
        {
            const uint2* qp = a.q + (((size_t) b * m.d + h * kDh + chunk * 16) >> 1);
            if (r == 0 && lane == 0)
            {
                DsSpin sp;
                while (ll_ld1(qp).y != a.genI)
                    if (ds_spin(cx, sp, DS_ERR_DATA_WAIT))
                        break;
            }
            __syncwarp();
            DsSpin sp;
            for (;;)
            {
                const bool ok = ds_load16h(qp, a.genI, qh);
                if (__all_sync(0xffffffffu, ok) || ds_spin(cx, sp, DS_ERR_DATA_WAIT))
                    break;
            }
        }
        if (r == 0)
            ds_phase_stamp(cx, 0);
        uint32_t bq[8];
        float koff; // 1152 * sum of q over the 64 dims: the bias of the 1024 + byte key values (xa_chunk KOFF)
        {
            const uint32_t* u = reinterpret_cast<const uint32_t*>(qh); // u[j] = (d2j, d2j+1)
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
            ll_st1(a.ctx_frag + (frag_index(b, h * kDh + 2 * lane) >> 1), h2u(__floats2half2_rn(a0 * inv, a1 * inv)), a.genO);
        }
    }
    if (nbh == 0)
        ds_phase_stamp(cx, 0);
    cx.item += (unsigned) (nbh * m.nch);
    consumer_sync(); // the merge area is reused by the next phase
#if defined(B200_DS_DEBUG)
    if (cx.dbg != nullptr && cx.c == 30 && cx.tid == 0 && cx.phase < kDsDbgPhases)
    {
        long long* f = cx.dbg + (size_t) cx.G * kDsDbgPhases * 2 + (size_t) cx.phase * 16;
        f[10] = dbg_wait, f[11] = dbg_chunks, f[12] = clock64() - dbg_t0;
    }
#endif
    ds_phase_stamp(cx, 1);
    ++cx.phase;
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
    uint2* x[3]; // residual stream, rotating: flagged words, A-fragment order [16 x d]
    uint2* ctx;  // A-fragment order [16 x d]
    uint2* u;    // A-fragment order [16 x dff]
    uint2* qkv;  // row-major [16][3d]
    uint2* q;    // row-major [16][d]
    uint2* part; // split-K partial sums [4][d / 8][128]
    int vocab, n_ctx;
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
        // ---- producer warp: weights and the cross-KV cache are static data, no dependency on the previous kernel ----
        cx.base = 0;
        if (cx.lane == 0)
            ds_producer(m, cx);
        return;
    }

    // ---- consumers ----
    grid_dep_wait(); // tokens / lengths / the previous step's cache rows and step counter come from earlier kernels
    // the step's first generation: every CTA reads the counter here, the LAST CTA to leave bumps it (see the end)
    cx.base = (ld_relaxed_u32(cx.sync + 3) + 1u) * kDsGenPerStep;
    ds_phase_stamp(cx, 0);
    {
        // token + positional embedding -> x[0] (A-fragment order); fp16 add like embed_kernel (glue.cu)
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
                    ll_st1(p.x[0] + (frag_index(row, k) >> 1), h2u(__hadd2(te, pe)), cx.base);
                }
            }
        }
        ds_phase_stamp(cx, 1);
        ++cx.phase;
    }
    // The layer's 32 pointers are staged in shared memory one layer ahead.
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
        // generations: x[0] of this layer carries base + 8 l (the embedding or the previous layer's fc2); phase s of the
        // layer (qkv, mmha, attn_out, cross_q, xattn, cross_out, fc1, fc2) writes base + 8 l + s + 1
        const unsigned g0 = cx.base + 8u * (unsigned) l;
#pragma unroll 1
        for (int sidx = 0; sidx < 8; ++sidx)
        {
            if (sidx == 1)
            {
                DsMmha a{p.qkv, reinterpret_cast<int8_t*>(lp[27]), reinterpret_cast<const float*>(lp[28]),
                    reinterpret_cast<const float*>(lp[29]), p.seq_len, p.ctx, g0 + 1u, g0 + 2u};
                ds_mmha_phase(a, m, cx);
            }
            else if (sidx == 4)
            {
                DsXattn a{p.q, reinterpret_cast<const float*>(lp[31]), p.ctx, g0 + 4u, g0 + 5u};
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
                // residual stream: x[0] -(attn_out)-> x[1] -(cross_out)-> x[2] -(fc2)-> x[0]
                a.A = g == 0 ? p.x[0] : (g == 1 || g == 3) ? p.ctx : g == 2 ? p.x[1] : g == 4 ? p.x[2] : p.u;
                a.genA = g == 0 ? g0 : g0 + (unsigned) sidx; // the phase before this one produced A (x[0]: see above)
                a.resid = g == 1 ? p.x[0] : g == 3 ? p.x[1] : g == 5 ? p.x[2] : nullptr;
                a.genR = g == 1 ? g0 : g == 3 ? g0 + 3u : g0 + 6u;
                a.out_frag = g == 1 ? p.x[1] : g == 3 ? p.x[2] : g == 5 ? p.x[0] : g == 4 ? p.u : nullptr;
                a.out_rm = g == 0 ? p.qkv : (g == 2 ? p.q : nullptr);
                a.out_plain = (g == 5 && last) ? p.x_out : nullptr;
                a.genO = g0 + (unsigned) sidx + 1u;
                a.part = p.part;
                a.act = g == 4 ? B200_ACT_GELU_ERF : B200_ACT_NONE;
                a.which = g;
                a.eps = p.eps;
                ds_gemm_phase(a, m, cx);
            }
        }
    }
    // ---- the last CTA out bumps the step counter (the next launch's generations) and clears the exit count ----
    consumer_sync();
    if (cx.tid == 0)
    {
        const unsigned old = atomicAdd(cx.sync + 1, 1u);
        if (old == (unsigned) cx.G - 1u)
        {
            cx.sync[1] = 0u;
            cx.sync[3] = cx.sync[3] + 1u;
        }
    }
}

} // namespace b200

using namespace b200;

static void* g_ds_debug = nullptr;

/* Debug aid: device buffer of n_ctas * 512 * 2 int64 receiving %globaltimer stamps of every following step launch
 * (per CTA and phase: [0] the phase's inputs arrived, [1] phase work done); NULL switches it off. */
extern "C" int b200_debug_decoder_step_timeline(void* device_buffer)
{
    g_ds_debug = device_buffer;
    return B200_OK;
}

/* 256 bytes of control words, then flagged-word buffers (8 bytes per half2 / fp32): three residual streams, the attention
 * context, q, qkv, the MLP's hidden activations and the split-K partial sums. */
extern "C" size_t b200_decoder_step_scratch_bytes(int num_heads, int d_ff)
{
    if (num_heads <= 0 || d_ff <= 0)
        return 0;
    const size_t d = (size_t) num_heads * kDh;
    return 256 + 64 * d * (3 + 1 + 1 + 3) + 64 * (size_t) d_ff + (size_t) kDsMaxSplit * 128 * d;
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
    B200_REQUIRE(1 + 8 * p->n_layers <= (int) kDsGenPerStep, B200_ERR_UNSUPPORTED, "%d layers > %d", p->n_layers,
        ((int) kDsGenPerStep - 1) / 8);
    const int d = p->num_heads * kDh;
    B200_REQUIRE(d <= kDsUnitK, B200_ERR_UNSUPPORTED, "hidden size %d > %d", d, kDsUnitK);
    const int nq = (p->d_ff + kDsUnitK - 1) / kDsUnitK;
    B200_REQUIRE(nq <= kDsMaxSplit && p->d_ff % 64 == 0 && p->d_ff % (64 * nq) == 0, B200_ERR_UNSUPPORTED,
        "d_ff %d must split into at most %d equal multiples of 64", p->d_ff, kDsMaxSplit);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(p->scratch) & 255) == 0, B200_ERR_INVALID_ARG, "scratch must be 256-byte aligned");
    if (p->batch_size == 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    int G = num_sms();
    if (p->max_ctas > 0 && p->max_ctas < G)
        G = p->max_ctas;
    B200_REQUIRE(G >= nq, B200_ERR_UNSUPPORTED, "%d CTAs cannot form a split-K group of %d", G, nq);
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
    for (int i = 0; i < 3; ++i)
        k.x[i] = reinterpret_cast<uint2*>(s), s += 64 * (size_t) d;
    k.ctx = reinterpret_cast<uint2*>(s), s += 64 * (size_t) d;
    k.q = reinterpret_cast<uint2*>(s), s += 64 * (size_t) d;
    k.qkv = reinterpret_cast<uint2*>(s), s += 192 * (size_t) d;
    k.u = reinterpret_cast<uint2*>(s), s += 64 * (size_t) p->d_ff;
    k.part = reinterpret_cast<uint2*>(s);
    k.vocab = p->vocab, k.n_ctx = p->n_ctx, k.eps = p->ln_eps;
    k.dbg = static_cast<long long*>(g_ds_debug);
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
