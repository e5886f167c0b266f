// log_mel.cu -- Whisper log-Mel front end on the device (the data format in front of the
// encoder stem).
//
// Replaces log_mel_spectrogram of T/examples/whisper/whisper_utils.py:99-145 (torch.stft on the host in the reference:
// run.py:44-46, summarize.py:120-122):
//   frames of 400 samples every 160, centre = True (200 samples of reflect padding), periodic Hann window, |DFT|^2 over
//   201 bins, last frame dropped, mel = filters[n_mels, 201] @ power, log10(max(mel, 1e-10)),
//   max(., max over the utterance - 8), (. + 4) / 4.
//
// 0.96 GFLOP per 30 s utterance as a direct real DFT (0.48 after the radix-2 step below) -- small next to the
// 1.9 TFLOP encoder, so this stays on the fp32 CUDA cores: the spectrum spans 8 decades (the clamp sits 80 dB below the utterance maximum) and fp16/tf32 tensor-core
// products would put their rounding noise at about -66 dB of every frame's strongest bin.  A sequentially accumulated
// fp32 DFT lands within 1e-5 of the exact value in output units (the reference's own fp32 FFT: 3.6e-5).
//
// log_mel_power_kernel: one CTA = 32 consecutive frames of one utterance.
//   1. the 5360 samples the frames cover go to shared memory once (coalesced, reflect-indexed), each 160-sample row
//      shifted by one word so that the next step is free of bank conflicts;
//   2. one radix-2 step (sum and difference of the two halves of every windowed frame, laid out sample-major
//      [n < 200][32 frames]) halves the arithmetic; a thread owns one DFT bin for all 32 frames (warps 0-3 the even
//      bins, 4-7 the odd ones): per sample n one twiddle (cos, sin)[k n mod 400] from a shared table (computed in fp64
//      per CTA) and eight 128-bit broadcast loads feed 64 FMAs on 64 register accumulators;
//   3. power[k][f] overwrites xs; the mel projection walks the filterbank 32 bins at a time and skips zero weights with
//      a ballot (391 of the 16080 weights are non-zero), lanes = frames so the stores along t are 128-byte rows;
//   4. log10 and the utterance maximum (atomicMax on an order-preserving integer encoding).
// log_mel_normalize_kernel applies the max - 8 floor and (x + 4) / 4 and writes fp32 (the reference's dtype) or fp16 (what
// the encoder stem consumes: `.type(torch.float16)`, run.py:45).
#include "common.cuh"

namespace b200
{

constexpr int kMelNfft = 400, kMelHop = 160, kMelBins = 201, kMelFT = 32;
constexpr int kMelSeg = (kMelFT - 1) * kMelHop + kMelNfft;         // samples covered by the frames of a CTA: 5360
constexpr int kMelSegPad = ((kMelSeg + kMelSeg / kMelHop + 1) + 3) & ~3; // row-shifted copy, rounded to 16 bytes
constexpr int kMelThreads = 256;
constexpr size_t kMelSmemBytes = sizeof(float) * (kMelNfft * kMelFT + kMelSegPad + 2 * kMelNfft + kMelNfft);

__device__ __forceinline__ uint32_t float_order_encode(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float float_order_decode(uint32_t e)
{
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

__global__ void __launch_bounds__(kMelThreads, 2) log_mel_power_kernel(const float* __restrict__ audio, int n_samples,
    int n_valid, const float* __restrict__ filters, int n_mels, int n_frames, float* __restrict__ log_spec,
    uint32_t* __restrict__ max_enc)
{
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                                   // [2][200][32] half-frame sums / differences, later power [201][32]
    float* seg = xs + kMelNfft * kMelFT;              // row-shifted raw samples
    float2* tw = reinterpret_cast<float2*>(seg + kMelSegPad); // (cos, sin)(2 pi j / 400)
    float* win = reinterpret_cast<float*>(tw + kMelNfft);
    __shared__ float red[kMelThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, f0 = blockIdx.x * kMelFT;
    const float* a = audio + (size_t) b * n_valid;

    // ---- 1. samples f0*160 - 200 ... of the (virtually) right-padded, reflect-extended utterance ----
    for (int i = tid; i < kMelSeg; i += kMelThreads)
    {
        int j = f0 * kMelHop + i - kMelNfft / 2;
        j = j < 0 ? -j : j;
        j = j >= n_samples ? 2 * (n_samples - 1) - j : j;
        j = min(max(j, 0), n_samples - 1);            // frames past the end of a partial tile: any finite value
        seg[i + i / kMelHop] = j < n_valid ? __ldg(a + j) : 0.f; // [n_valid, n_samples) is the `padding` argument: zeros
    }
    for (int j = tid; j < kMelNfft; j += kMelThreads)
    {
        double s, c;
        sincospi((double) j / (kMelNfft / 2), &s, &c);
        tw[j] = make_float2((float) c, (float) s);
        win[j] = (float) (0.5 - 0.5 * c);             // torch.hann_window(400): periodic
    }
    __syncthreads();

    // ---- 2. one radix-2 step: exp(-2 pi i k (n + 200) / 400) = (-1)^k exp(-2 pi i k n / 400), so
    //      X[k] = sum_{n < 200} (xw[n] + (-1)^k xw[n + 200]) tw[k n mod 400]   with xw = window * samples.
    //      xs[0][n][f] = xw[n] + xw[n + 200] feeds the even bins, xs[1][n][f] = xw[n] - xw[n + 200] the odd ones: half the
    //      multiply-adds of the plain DFT.  lanes = frames: conflict-free on both sides ----
    constexpr int kHalf = kMelNfft / 2;
    for (int idx = tid; idx < kHalf * kMelFT; idx += kMelThreads)
    {
        const int f = idx & (kMelFT - 1), n = idx >> 5;
        const int i0 = f * kMelHop + n, i1 = i0 + kHalf;
        const float lo = win[n] * seg[i0 + i0 / kMelHop], hi = win[n + kHalf] * seg[i1 + i1 / kMelHop];
        xs[idx] = lo + hi;
        xs[kHalf * kMelFT + idx] = lo - hi;
    }
    __syncthreads();

    float re[kMelFT], im[kMelFT];
#pragma unroll
    for (int f = 0; f < kMelFT; ++f)
        re[f] = im[f] = 0.f;
    // warps 0-3: even bins 0, 2, ..., 200; warps 4-7: odd bins 1, 3, ..., 199 (warp-uniform source array)
    const int odd = tid >> 7;
    const int k = 2 * (tid & 127) + odd;
    if (k < kMelBins)
    {
        const float* src = xs + odd * (kHalf * kMelFT);
        int j = 0;
#pragma unroll 2
        for (int n = 0; n < kHalf; ++n)
        {
            const float2 t = tw[j];
            j += k;
            j = j >= kMelNfft ? j - kMelNfft : j;
            const float4* xr = reinterpret_cast<const float4*>(src + n * kMelFT);
#pragma unroll
            for (int q = 0; q < kMelFT / 4; ++q)
            {
                const float4 v = xr[q];
                re[4 * q + 0] = fmaf(v.x, t.x, re[4 * q + 0]);
                im[4 * q + 0] = fmaf(v.x, t.y, im[4 * q + 0]);
                re[4 * q + 1] = fmaf(v.y, t.x, re[4 * q + 1]);
                im[4 * q + 1] = fmaf(v.y, t.y, im[4 * q + 1]);
                re[4 * q + 2] = fmaf(v.z, t.x, re[4 * q + 2]);
                im[4 * q + 2] = fmaf(v.z, t.y, im[4 * q + 2]);
                re[4 * q + 3] = fmaf(v.w, t.x, re[4 * q + 3]);
                im[4 * q + 3] = fmaf(v.w, t.y, im[4 * q + 3]);
            }
        }
    }
    __syncthreads(); // every thread is done reading xs
    // ---- 3. power[k][f] over xs ----
    if (k < kMelBins)
    {
        float4* pr = reinterpret_cast<float4*>(xs + k * kMelFT);
#pragma unroll
        for (int q = 0; q < kMelFT / 4; ++q)
            pr[q] = make_float4(re[4 * q] * re[4 * q] + im[4 * q] * im[4 * q],
                re[4 * q + 1] * re[4 * q + 1] + im[4 * q + 1] * im[4 * q + 1],
                re[4 * q + 2] * re[4 * q + 2] + im[4 * q + 2] * im[4 * q + 2],
                re[4 * q + 3] * re[4 * q + 3] + im[4 * q + 3] * im[4 * q + 3]);
    }
    __syncthreads();

    // ---- mel projection + log10: warp = group of mel bands, lane = frame ----
    const int f = f0 + lane;
    float vmax = -3.0e38f;
    for (int m = warp; m < n_mels; m += kMelThreads / 32)
    {
        const float* fr = filters + (size_t) m * kMelBins;
        float acc = 0.f;
        for (int k0 = 0; k0 < kMelBins; k0 += 32)
        {
            const float wv = (k0 + lane < kMelBins) ? __ldg(fr + k0 + lane) : 0.f;
            uint32_t mask = __ballot_sync(0xffffffffu, wv != 0.f);
            while (mask)
            {
                const int bit = __ffs(mask) - 1;
                mask &= mask - 1;
                acc = fmaf(__shfl_sync(0xffffffffu, wv, bit), xs[(k0 + bit) * kMelFT + lane], acc);
            }
        }
        const float v = log10f(fmaxf(acc, 1e-10f));
        if (f < n_frames)
        {
            log_spec[((size_t) b * n_mels + m) * n_frames + f] = v;
            vmax = fmaxf(vmax, v);
        }
    }
    // ---- 4. utterance maximum ----
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0)
        red[warp] = vmax;
    __syncthreads();
    if (tid == 0)
    {
        float v = red[0];
#pragma unroll
        for (int w = 1; w < kMelThreads / 32; ++w)
            v = fmaxf(v, red[w]);
        atomicMax(max_enc + b, float_order_encode(v));
    }
}

template <typename OutT>
__global__ void __launch_bounds__(256) log_mel_normalize_kernel(const float* __restrict__ log_spec,
    const uint32_t* __restrict__ max_enc, OutT* __restrict__ out, size_t per_utt, size_t total)
{
    for (size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t) gridDim.x * blockDim.x)
    {
        const float floor_v = float_order_decode(__ldg(max_enc + idx / per_utt)) - 8.0f;
        const float v = (fmaxf(log_spec[idx], floor_v) + 4.0f) / 4.0f;
        if constexpr (sizeof(OutT) == 2)
            out[idx] = __float2half_rn(v);
        else
            out[idx] = v;
    }
}

} // namespace b200

using namespace b200;

static inline size_t log_mel_header_bytes(int batch_size)
{
    return ((size_t) batch_size * sizeof(uint32_t) + 255) & ~size_t(255);
}

extern "C" int b200_log_mel_frames(int n_samples, int padding)
{
    if (n_samples < 0 || padding < 0)
        return 0;
    return (n_samples + padding) / kMelHop;
}

extern "C" size_t b200_log_mel_workspace_bytes(int batch_size, int n_samples, int padding, int n_mels)
{
    if (batch_size <= 0 || n_mels <= 0)
        return 0;
    const int n_frames = b200_log_mel_frames(n_samples, padding);
    return log_mel_header_bytes(batch_size) + (size_t) batch_size * n_mels * n_frames * sizeof(float);
}

extern "C" int b200_log_mel_spectrogram(const float* audio, int batch_size, int n_samples, int padding,
    const float* mel_filters, int n_mels, void* out, int out_dtype, void* workspace, size_t workspace_bytes,
    b200_stream_t stream)
{
    B200_REQUIRE(audio && mel_filters && out, B200_ERR_INVALID_ARG, "null pointer (audio/mel_filters/out)");
    B200_REQUIRE(n_samples > 0 && padding >= 0 && n_mels > 0, B200_ERR_INVALID_ARG, "bad sizes");
    B200_REQUIRE(out_dtype == B200_DTYPE_F32 || out_dtype == B200_DTYPE_F16, B200_ERR_INVALID_ARG,
        "out_dtype must be B200_DTYPE_F32 or B200_DTYPE_F16");
    const long long n_total = (long long) n_samples + padding;
    // torch.stft's reflect padding needs more samples than the pad width (n_fft / 2)
    B200_REQUIRE(n_total > kMelNfft / 2, B200_ERR_INVALID_ARG, "log-mel: %lld samples, reflect padding needs more than %d",
        n_total, kMelNfft / 2);
    B200_REQUIRE(n_total < (1ll << 30), B200_ERR_UNSUPPORTED, "log-mel: utterance too long");
    const int n_frames = (int) (n_total / kMelHop);
    B200_REQUIRE(n_frames > 0, B200_ERR_INVALID_ARG, "log-mel: fewer than %d samples give no frame", kMelHop);
    if (batch_size <= 0)
        return B200_OK;
    B200_REQUIRE(batch_size <= 65535, B200_ERR_UNSUPPORTED, "log-mel: batch %d > 65535", batch_size);
    const size_t need = b200_log_mel_workspace_bytes(batch_size, n_samples, padding, n_mels);
    B200_REQUIRE(workspace != nullptr && workspace_bytes >= need, B200_ERR_WORKSPACE,
        "log-mel: workspace of %zu bytes needed, got %zu", need, workspace_bytes);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, B200_ERR_INVALID_ARG,
        "log-mel: workspace must be 16-byte aligned");
    B200_REQUIRE_DEVICE();
    cudaStream_t st = as_stream(stream);
    uint32_t* max_enc = static_cast<uint32_t*>(workspace);
    float* log_spec = reinterpret_cast<float*>(static_cast<char*>(workspace) + log_mel_header_bytes(batch_size));

    static bool attr_set = false;
    if (!attr_set)
    {
        B200_CUDA(cudaFuncSetAttribute(log_mel_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) kMelSmemBytes));
        attr_set = true;
    }
    B200_CUDA(cudaMemsetAsync(max_enc, 0, (size_t) batch_size * sizeof(uint32_t), st)); // 0 orders below every float
    const dim3 grid((n_frames + kMelFT - 1) / kMelFT, batch_size);
    log_mel_power_kernel<<<grid, kMelThreads, kMelSmemBytes, st>>>(audio, (int) n_total, n_samples, mel_filters, n_mels,
        n_frames, log_spec, max_enc);
    B200_LAUNCH_CHECK();

    const size_t per_utt = (size_t) n_mels * n_frames, total = per_utt * batch_size;
    const int blocks = (int) ((total + 255) / 256 < (size_t) (num_sms() * 8) ? (total + 255) / 256 : num_sms() * 8);
    if (out_dtype == B200_DTYPE_F16)
        log_mel_normalize_kernel<__half><<<blocks, 256, 0, st>>>(log_spec, max_enc, static_cast<__half*>(out), per_utt, total);
    else
        log_mel_normalize_kernel<float><<<blocks, 256, 0, st>>>(log_spec, max_enc, static_cast<float*>(out), per_utt, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
