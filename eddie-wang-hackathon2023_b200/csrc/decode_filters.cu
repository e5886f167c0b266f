// decode_filters.cu -- Whisper logit filters + greedy token update on the device.
//
// Replaces the per-sequence Python loops of T/examples/whisper/decoding.py
//   SuppressBlank :202-209, SuppressTokens :212-217, ApplyTimestampRules :134-199, GreedyDecoder.update :274-293
// (they call .tolist() on the token history of every sequence at every step, main_loop :785-821) by ONE kernel per
// step: every CTA scans a slice of a sequence's logits, evaluates the masks arithmetically from four integers of
// per-sequence state (sampled count, last / penultimate sampled token, last timestamp), keeps an online
// (max, argmax, sum-of-exp) for the text range and for the timestamp range, and the last CTA of a row to arrive applies
// the "timestamp probability mass beats every text token" rule, picks the token, adds its log-probability to
// sum_logprobs, forces eot after eot, and advances the state.  Stays inside the captured CUDA graph.
#include <math_constants.h>

#include "common.cuh"

namespace b200
{
int* tc_counter_slot(int needed, cudaStream_t stream);

struct RangeAcc
{
    float m; // running maximum (-inf when empty)
    float s; // sum of exp(x - m)
    int i;   // argmax (lowest index on ties)
};

__device__ __forceinline__ void acc_add(RangeAcc& a, float x, int idx)
{
    if (x > a.m)
    {
        a.s = a.s * __expf(a.m - x) + 1.f; // exp(-inf) = 0 on the first element
        a.m = x;
        a.i = idx;
    }
    else
    {
        a.s += __expf(x - a.m);
        if (x == a.m && idx < a.i)
            a.i = idx;
    }
}

__device__ __forceinline__ RangeAcc acc_merge(const RangeAcc& a, const RangeAcc& b)
{
    if (b.s == 0.f)
        return a;
    if (a.s == 0.f)
        return b;
    RangeAcc r;
    r.m = fmaxf(a.m, b.m);
    r.s = a.s * __expf(a.m - r.m) + b.s * __expf(b.m - r.m);
    r.i = (a.m > b.m || (a.m == b.m && a.i < b.i)) ? a.i : b.i;
    return r;
}

__device__ __forceinline__ RangeAcc acc_shfl_xor(const RangeAcc& a, int o)
{
    RangeAcc r;
    r.m = __shfl_xor_sync(0xffffffffu, a.m, o);
    r.s = __shfl_xor_sync(0xffffffffu, a.s, o);
    r.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    return r;
}

struct FilterParams
{
    int eot, no_timestamps, timestamp_begin, blank, max_initial;
};

// state[row] = {n_sampled, last, penultimate, last_timestamp + 1 (0: none yet)}
__global__ void __launch_bounds__(256) whisper_filtered_argmax_kernel(const float* __restrict__ logits, int vocab,
    const uint32_t* __restrict__ suppress_bitmap, const FilterParams prm, int4* __restrict__ state, int* __restrict__ next_token,
    float* __restrict__ sum_logprobs, float* __restrict__ scratch, int* __restrict__ counters)
{
    grid_dep_wait();
    grid_dep_launch_dependents();
    const int r = blockIdx.y, parts = gridDim.x;
    const int4 st = state[r];
    const int n = st.x, last = st.y, penult = st.z, last_ts = st.w - 1;
    const int ts = prm.timestamp_begin, eot = prm.eot;
    const bool last_was_ts = n >= 1 && last >= ts;
    const bool penult_was_ts = n < 2 || penult >= ts;
    const int ts_floor = last_ts >= 0 ? ((last_was_ts && !penult_was_ts) ? last_ts : last_ts + 1) : ts; // [ts, ts_floor) masked
    const int ts_cap = (n == 0 && prm.max_initial >= 0) ? ts + prm.max_initial : vocab;                 // (ts_cap, V) masked

    const float* lr = logits + (size_t) r * vocab;
    const int per = (vocab + parts - 1) / parts;
    const int v0 = blockIdx.x * per, v1 = min(vocab, v0 + per);
    RangeAcc text{-CUDART_INF_F, 0.f, 0x7fffffff}, stamp{-CUDART_INF_F, 0.f, 0x7fffffff};
    for (int base = v0 + threadIdx.x; base < v1; base += 8 * 256)
    {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const int v = base + j * 256;
            x[j] = v < v1 ? __ldcs(lr + v) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const int v = base + j * 256;
            if (v >= v1)
                continue;
            bool masked = (suppress_bitmap != nullptr && ((suppress_bitmap[v >> 5] >> (v & 31)) & 1u)) || v == prm.no_timestamps;
            masked |= n == 0 && (v == prm.blank || v == eot || v < ts);
            if (last_was_ts)
                masked |= penult_was_ts ? (v >= ts) : (v < eot);
            masked |= v >= ts && (v < ts_floor || v > ts_cap);
            if (masked)
                continue;
            if (v < ts)
                acc_add(text, x[j], v);
            else
                acc_add(stamp, x[j], v);
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
    {
        text = acc_merge(text, acc_shfl_xor(text, o));
        stamp = acc_merge(stamp, acc_shfl_xor(stamp, o));
    }
    __shared__ RangeAcc sh[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
    {
        sh[0][warp] = text;
        sh[1][warp] = stamp;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < 8; ++w)
        {
            text = acc_merge(text, sh[0][w]);
            stamp = acc_merge(stamp, sh[1][w]);
        }
        float* mine = scratch + ((size_t) r * parts + blockIdx.x) * 6;
        mine[0] = text.m, mine[1] = text.s, mine[2] = __int_as_float(text.i);
        mine[3] = stamp.m, mine[4] = stamp.s, mine[5] = __int_as_float(stamp.i);
        __threadfence();
        if (atomicAdd(&counters[r], 1) == parts - 1)
        {
            __threadfence();
            RangeAcc T{-CUDART_INF_F, 0.f, 0x7fffffff}, S{-CUDART_INF_F, 0.f, 0x7fffffff};
            for (int q = 0; q < parts; ++q) // fixed order: bit-reproducible
            {
                volatile float* pq = scratch + ((size_t) r * parts + q) * 6;
                T = acc_merge(T, RangeAcc{pq[0], pq[1], __float_as_int(pq[2])});
                S = acc_merge(S, RangeAcc{pq[3], pq[4], __float_as_int(pq[5])});
                // the scratch lives in the library's shared pool of self-resetting counter slots: every word must be
                // zero again when the launch ends, or the next user of the slot (split-K / split-KV arrival counters)
                // starts from garbage
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    pq[j] = 0.f;
            }
            const float lse_t = T.s > 0.f ? T.m + logf(T.s) : -CUDART_INF_F;
            const float lse_s = S.s > 0.f ? S.m + logf(S.s) : -CUDART_INF_F;
            int nxt;
            float top, lse;
            if (lse_s > T.m) // the timestamp mass beats every text token: text is masked (decoding.py:190-199)
            {
                nxt = S.i, top = S.m, lse = lse_s;
            }
            else
            {
                const bool pick_text = T.m >= S.m; // ties: the lower index, and text indices are below timestamp_begin
                nxt = pick_text ? T.i : S.i;
                top = pick_text ? T.m : S.m;
                const float hi = fmaxf(lse_t, lse_s), lo = fminf(lse_t, lse_s);
                lse = (lo == -CUDART_INF_F) ? hi : hi + log1pf(expf(lo - hi));
            }
            const bool ended = n >= 1 && last == eot; // GreedyDecoder.update: eot stays eot, sum_logprobs frozen
            if (ended)
                nxt = eot;
            else if (sum_logprobs != nullptr)
                sum_logprobs[r] += top - lse;
            next_token[r] = nxt;
            state[r] = make_int4(n + 1, nxt, last, nxt >= ts ? nxt + 1 : st.w);
            counters[r] = 0;
        }
    }
}
} // namespace b200

using namespace b200;

extern "C" int b200_whisper_filtered_argmax(const float* logits, int rows, int vocab, const uint32_t* suppress_bitmap,
    int eot, int no_timestamps, int timestamp_begin, int blank_token, int max_initial_timestamp_index, int32_t* decode_state,
    int32_t* next_token, float* sum_logprobs, b200_stream_t stream)
{
    B200_REQUIRE(logits && decode_state && next_token, B200_ERR_INVALID_ARG, "null pointer (logits/decode_state/next_token)");
    B200_REQUIRE(vocab > 0 && eot >= 0 && eot < vocab && timestamp_begin > eot && timestamp_begin <= vocab,
        B200_ERR_INVALID_ARG, "bad vocabulary layout (vocab %d, eot %d, timestamp_begin %d)", vocab, eot, timestamp_begin);
    if (rows <= 0)
        return B200_OK;
    B200_REQUIRE(rows <= 256, B200_ERR_UNSUPPORTED, "filtered argmax: %d rows exceed the scratch slot", rows);
    B200_REQUIRE_DEVICE();
    constexpr int parts = 8;
    int* slot = tc_counter_slot(rows * (parts * 6 + 1), as_stream(stream));
    B200_REQUIRE(slot != nullptr, B200_ERR_CUDA, "filtered argmax: no scratch slot");
    // counters first (they must be zero between launches and reset themselves), partial results after them
    int* counters = slot;
    float* scratch = reinterpret_cast<float*>(slot + ((rows + 3) & ~3));
    FilterParams prm{eot, no_timestamps, timestamp_begin, blank_token, max_initial_timestamp_index};
    B200_LAUNCH(whisper_filtered_argmax_kernel, dim3(parts, rows), dim3(256), 0, as_stream(stream), logits, vocab,
        suppress_bitmap, prm, reinterpret_cast<int4*>(decode_state), next_token, sum_logprobs, scratch, counters);
    return B200_OK;
}

// =====================================================================================================
// Language detection and the no-speech probability (T/examples/whisper/decoding.py:703-741 and :762-766): both read
// the logits of the start-of-transcript position.
//   * range softmax: the logits outside [range_lo, range_hi) (the 99 language tokens are contiguous) count as -inf;
//     range_argmax[row] = the most probable token of the range, range_probs[row][j] = softmax over the range;
//   * probe_prob[row] = softmax over the WHOLE vocabulary evaluated at probe_token (the nospeech token).
// One CTA per row: two passes over the row (max, then sum of exponentials, like torch's softmax), fp32 throughout.
// =====================================================================================================
namespace b200
{
__global__ void __launch_bounds__(256) range_softmax_kernel(const float* __restrict__ logits, int vocab, int range_lo,
    int range_hi, int probe_token, int32_t* __restrict__ range_argmax, float* __restrict__ range_probs,
    float* __restrict__ probe_prob)
{
    __shared__ float red_f[8];
    __shared__ int red_i[8];
    __shared__ float bc[2];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* x = logits + (size_t) row * vocab;
    grid_dep_wait();
    grid_dep_launch_dependents();

    auto block_max = [&](float v) -> float
    {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1)
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0)
            red_f[warp] = v;
        __syncthreads();
        float r = red_f[0];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            r = fmaxf(r, red_f[w]);
        __syncthreads();
        return r;
    };
    auto block_sum = [&](float v) -> float
    {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0)
            red_f[warp] = v;
        __syncthreads();
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w)
            r += red_f[w];
        __syncthreads();
        return r;
    };

    if (probe_prob != nullptr)
    {
        float m = -CUDART_INF_F;
        for (int i = tid; i < vocab; i += 256)
            m = fmaxf(m, x[i]);
        m = block_max(m);
        float s = 0.f;
        for (int i = tid; i < vocab; i += 256)
            s += expf(x[i] - m);
        s = block_sum(s);
        if (tid == 0)
            probe_prob[row] = expf(x[probe_token] - m) / s;
    }
    if (range_argmax != nullptr || range_probs != nullptr)
    {
        float m = -CUDART_INF_F;
        int arg = range_hi;
        for (int i = range_lo + tid; i < range_hi; i += 256)
            if (x[i] > m)
            {
                m = x[i];
                arg = i;
            }
        // (max, lowest index) across the block: torch.argmax returns the first maximum
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1)
        {
            const float om = __shfl_xor_sync(0xffffffffu, m, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > m || (om == m && oa < arg))
            {
                m = om;
                arg = oa;
            }
        }
        if (lane == 0)
        {
            red_f[warp] = m;
            red_i[warp] = arg;
        }
        __syncthreads();
        if (tid == 0)
        {
            for (int w = 1; w < 8; ++w)
                if (red_f[w] > m || (red_f[w] == m && red_i[w] < arg))
                {
                    m = red_f[w];
                    arg = red_i[w];
                }
            bc[0] = m;
            if (range_argmax != nullptr)
                range_argmax[row] = arg;
        }
        __syncthreads();
        m = bc[0];
        float s = 0.f;
        for (int i = range_lo + tid; i < range_hi; i += 256)
            s += expf(x[i] - m);
        s = block_sum(s);
        if (range_probs != nullptr)
            for (int i = range_lo + tid; i < range_hi; i += 256)
                range_probs[(size_t) row * (range_hi - range_lo) + (i - range_lo)] = expf(x[i] - m) / s;
    }
}
} // namespace b200

extern "C" int b200_logits_range_softmax(const float* logits, int rows, int vocab, int range_lo, int range_hi,
    int probe_token, int32_t* range_argmax, float* range_probs, float* probe_prob, b200_stream_t stream)
{
    B200_REQUIRE(logits != nullptr, B200_ERR_INVALID_ARG, "null pointer (logits)");
    B200_REQUIRE(vocab > 0 && range_lo >= 0 && range_lo < range_hi && range_hi <= vocab, B200_ERR_INVALID_ARG,
        "bad token range [%d, %d) for a vocabulary of %d", range_lo, range_hi, vocab);
    B200_REQUIRE(probe_prob == nullptr || (probe_token >= 0 && probe_token < vocab), B200_ERR_INVALID_ARG,
        "probe token %d outside the vocabulary", probe_token);
    B200_REQUIRE(range_argmax || range_probs || probe_prob, B200_ERR_INVALID_ARG, "no output requested");
    if (rows <= 0)
        return B200_OK;
    B200_REQUIRE_DEVICE();
    B200_LAUNCH(range_softmax_kernel, dim3(rows), dim3(256), 0, as_stream(stream), logits, vocab, range_lo, range_hi,
        probe_token, range_argmax, range_probs, probe_prob);
    return B200_OK;
}
