/*
 * b200_whisper.h -- C ABI of the B200-native (sm_100a) quantized Whisper decoder hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, `extern "C"`, CUDA stream last, int status.
 * The C++ TensorRT plugin classes under eddie-wang-hackathon2023_b200/plugins/ (same names, fields and
 * serialization as the reference's) only marshal into these entry points, and so does the Python mirror of
 * the reference's operator API (ctypes).  Every entry point cites the reference interface it replaces;
 * paths are relative to /root/reference/tensorrt_llm_july-release-v1/ (T/).
 *
 * All `const void*` tensors are DEVICE pointers unless the function name ends in `_host`.
 * fp16 tensors are IEEE binary16 (`half`), row-major, innermost dimension contiguous.
 * Functions are asynchronous on `stream` and return B200_OK (0) or an error code; b200_last_error()
 * returns a thread-local message for the last failure.  There is NO CPU fallback anywhere: on a machine
 * without an sm_100 device every compute entry point returns B200_ERR_CUDA.
 */
#ifndef B200_WHISPER_H
#define B200_WHISPER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

typedef struct CUstream_st* b200_stream_t; /* == cudaStream_t */

enum
{
    B200_OK = 0,
    B200_ERR_INVALID_ARG = 1,
    B200_ERR_UNSUPPORTED = 2,
    B200_ERR_CUDA = 3,
    B200_ERR_WORKSPACE = 4
};

/* dtype codes (same numbering as nvinfer1::DataType: kFLOAT=0, kHALF=1, kINT8=2, kINT32=3) */
enum
{
    B200_DTYPE_F32 = 0,
    B200_DTYPE_F16 = 1,
    B200_DTYPE_I8 = 2,
    B200_DTYPE_I32 = 3
};

/* epilogue activation codes for the fused matmul / conv entry points */
enum
{
    B200_ACT_NONE = 0,
    B200_ACT_GELU_ERF = 1, /* torch nn.GELU(), T/examples/whisper/torch_model.py:120,143-144 */
    B200_ACT_GELU_TANH = 2 /* TRT-LLM gelu(), T/tensorrt_llm/functional.py:2044-2056 */
};

const char* b200_last_error(void);
/* Programmatic dependent launch for the hot-path kernels (default on; env B200_PDL=0 disables): each kernel may start
 * while its predecessor on the stream drains and waits (griddepcontrol.wait) before touching the predecessor's output. */
int b200_set_pdl(int enabled);
/* Promise that the KV caches, sequence_lengths and KV scales read by b200_mmha_generation / b200_cross_attention are
 * NOT written by the kernel launched right before them on the stream (true inside a decoder step: a cache row is
 * written one step before it is read, the lengths are bumped after the step's last kernel).  With the hint the
 * attention kernels fetch (and, for self-attention, convert) the cache before griddepcontrol.wait, so only the dot
 * products remain once the preceding projection has finished.  Default off (always safe).  The switch belongs to the
 * CALLING THREAD (thread-local): it qualifies the launches that thread makes next and is invisible to other threads, so
 * plugin instances enqueueing concurrently never see each other's hint. */
int b200_set_static_kv_hint(int enabled);
/* Fire-and-forget prefetch of [ptr, ptr+bytes) into L2 (cp.async.bulk.prefetch.L2); ptr 16-byte aligned. */
int b200_l2_prefetch(const void* ptr, size_t bytes, b200_stream_t stream);
int b200_abi_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
unsigned long long b200_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Weight preparation.
 * Replaces  symmetric_quantize<half,half|float>  T/cpp/tensorrt_llm/kernels/cutlass_kernels/cutlass_preprocessors.cpp:615-721
 *           preprocess_weights_for_mixed_gemm    same file :537-578 (Sm80..Sm90 layout: row permute, column major,
 *                                                2-column interleave in 64-row tiles, +128 bias and byte swizzle)
 * as reached from torch.ops.fastertransformer.symmetric_quantize_last_axis_of_batched_matrix
 *           T/cpp/tensorrt_llm/thop/weightOnlyQuantOp.cpp:143-236.
 * The reference runs these single-threaded on the host; here they are CUDA kernels, bit-exact with it.
 *
 * w: [K, N] row-major (the transpose of torch's [out, in] Linear weight, as T/examples/whisper/weight.py:76-77
 * passes it), dtype B200_DTYPE_F16 or B200_DTYPE_F32.  K % 64 == 0 and N % 64 == 0 (reference checks :187-192,
 * :498-500).  proc: K*N bytes in the processed layout; raw (nullable): [K, N] int8; scales: [N] of scale_dtype
 * (F16 or F32).  Stored scale = amax/128 rounded to scale_dtype; quantization divides by the fp32 scale.
 * ---------------------------------------------------------------------------------------------- */
int b200_symmetric_quantize_int8(const void* w, int w_dtype, int K, int N, int8_t* proc, int8_t* raw, void* scales,
    int scale_dtype, b200_stream_t stream);
int b200_preprocess_weights_int8(const int8_t* raw, int K, int N, int8_t* proc, b200_stream_t stream);
/* Host-pointer conveniences (H2D copy, kernels, D2H copy, stream synchronised on return) -- these mirror the
 * reference torch op, which takes and returns CPU tensors. */
int b200_symmetric_quantize_int8_host(const void* w, int w_dtype, int K, int N, int8_t* proc, int8_t* raw,
    void* scales, int scale_dtype);
int b200_preprocess_weights_int8_host(const int8_t* raw, int K, int N, int8_t* proc);

/* ------------------------------------------------------------------------------------------------
 * fp16 activations x per-channel int8 weights (weightOnlyQuantMatmul / fpA_intB).
 * Replaces  weight_only_gemv_launcher                 T/cpp/tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.cu:371-378 (m == 1)
 *           CutlassFpAIntBGemmRunner<half,uint8_t>::gemm  T/cpp/tensorrt_llm/kernels/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:358-435 (m > 1)
 *           ...::getWorkspaceSize                     same file :426-435
 * as dispatched by WeightOnlyQuantMatmulPlugin::enqueue  T/cpp/tensorrt_llm/plugins/weightOnlyQuantMatmulPlugin/weightOnlyQuantMatmulPlugin.cpp:162-222.
 *
 * A [M, K] fp16; Wproc: K*N bytes in the processed layout above; scales [N] fp16; C [M, N] fp16.
 * C = fp16( (sum_k A[m,k] * q[k,n]) * scales[n] ), fp32 accumulation.  K % 64 == 0, N % 64 == 0, M >= 1.
 * The _fused variant additionally applies, in this order and each rounded to fp16 like the reference's separate
 * elementwise layers (T/tensorrt_llm/quantization/layer.py:311-312): + bias[n], activation, + residual[m, n].
 * bias / residual may be NULL.  residual may alias C.
 * ---------------------------------------------------------------------------------------------- */
size_t b200_woq_workspace_bytes(int max_m, int n, int k);
int b200_woq_int8_gemm(const void* A, int M, int K, const int8_t* Wproc, const void* scales, int N, void* C,
    void* workspace, size_t workspace_bytes, b200_stream_t stream);
int b200_woq_int8_gemm_fused(const void* A, int M, int K, const int8_t* Wproc, const void* scales, int N,
    const void* bias, int activation, const void* residual, void* C, void* workspace, size_t workspace_bytes,
    b200_stream_t stream);
/* LayerNorm + matmul in one call: C = epilogue( LN(X; gamma, beta, eps) x W ), X [M, K] fp16 the raw residual stream
 * (the LayerNorm -> Linear pairs of T/tensorrt_llm/models/whisper/model.py:86-118).  This entry normalises into the
 * workspace (which must hold M*K fp16 on top of the split-K needs) and then runs the matmul: two launches. */
int b200_woq_int8_gemm_ln_fused(const void* X, const void* ln_gamma, const void* ln_beta, float ln_eps, int M, int K,
    const int8_t* Wproc, const void* scales, int N, const void* bias, int activation, const void* residual, void* C,
    void* workspace, size_t workspace_bytes, b200_stream_t stream);
/* The same operation with the LayerNorm FOLDED into the tcgen05 kernel (decode-sized M <= 32, K <= 1536; other shapes
 * take the two-launch route above).  Algebra:
 *     LN(x)[k] = (x[k] - mean) * rstd * gamma[k] + beta[k]
 *     y[n]     = rstd * ( sum_k x[k] * W'[k][n]  -  mean * c1[n] ) * scale[n]  +  c2[n]
 * with W'[k][n] = fp16(Wint[k][n] * gamma[k]) formed by the dequant warps (one extra HMUL2 per k pair),
 * c1[n] = sum_k W'[k][n] and c2[n] = scale[n] * sum_k beta[k] * Wint[k][n].  The tensor cores therefore consume the
 * RAW rows of x straight from TMA, and (mean, rstd) are only needed in the epilogue: they are computed by the idle
 * dequant warps while the activation tiles and MMAs are in flight, so no LayerNorm work sits on the critical path.
 * b200_woq_ln_fold_prepare fills c1s[n] = scale[n] * c1[n] and c2[n] (fp32 [N] each) once per (layer norm, weight)
 * pair, walking the preprocessed layout with the same arithmetic as the kernel so the mean term cancels exactly. */
int b200_woq_ln_fold_prepare(const int8_t* Wproc, const void* scales, const void* ln_gamma, const void* ln_beta, int K,
    int N, float* c1s, float* c2, b200_stream_t stream);
int b200_woq_int8_gemm_ln_folded(const void* X, const void* ln_gamma, const void* ln_beta, const float* c1s,
    const float* c2, float ln_eps, int M, int K, const int8_t* Wproc, const void* scales, int N, const void* bias,
    int activation, const void* residual, void* C, void* workspace, size_t workspace_bytes, b200_stream_t stream);
/* Forces a kernel family for tests/benchmarks: 0 = auto, 1 = SIMT GEMV, 2 = tcgen05 GEMM.  Thread-local: it affects the
 * calling thread's launches only (other threads keep the automatic dispatch). */
int b200_woq_set_kernel_policy(int policy);
/* Debug aid: device buffer of >= 16 int64 receiving clock64() stamps of CTA (0,0,0) of each following tcgen05 GEMM
 * launch at its phase boundaries; NULL switches it off. */
/* Debug / test aid (host only): the launch plan of the tcgen05 path for a shape: plan5 = {m-tile rows, m tiles, n tiles,
 * k splits, 1 if the splits form a thread-block cluster}. */
int b200_debug_woq_plan(int M, int N, int K, int* plan5);
/* Debug aid: the next max_launches tcgen05 GEMM launches record (launch order) 4 int64 %globaltimer values each:
 * earliest CTA entry, earliest dependency-wait return, latest CTA exit, unused.  Caller presets INT64_MAX/INT64_MAX/0. */
int b200_debug_tc_timeline(void* device_buffer, int max_launches);
int b200_debug_tc_timing_filter(int n, int folded_ln); /* only launches with this N and folded-LN flag stamp (0: all) */
int b200_debug_tc_timing(void* device_buffer);
/* One-time allocation of library-owned device state (split-K tile counters).  Call before CUDA-graph capture. */
int b200_init(void);

/* ------------------------------------------------------------------------------------------------
 * Masked multi-head attention, generation phase, contiguous KV cache [B, 2, H, Smax, Dh], fp16 I/O,
 * int8 or fp16 cache.
 * Replaces  masked_multihead_attention (Dh=64 half)   T/cpp/tensorrt_llm/kernels/decoderMaskedMultiheadAttention/decoderMaskedMultiheadAttentionTemplate.h:1195-2017
 *           Masked_multihead_attention_params         T/cpp/tensorrt_llm/kernels/decoderMaskedMultiheadAttention.h:70-188
 *           KVLinearBuffer                            T/cpp/tensorrt_llm/kernels/kvCacheUtils.h:114-170
 * as filled by GPTAttentionPluginCommon::enqueueGeneration  T/cpp/tensorrt_llm/plugins/gptAttentionCommon/gptAttentionCommon.cpp:649-780.
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200_mmha_params
{
    const void* qkv;              /* [B, 3*H*Dh] fp16: q | k | v of the current token (params.q/k/v + stride) */
    const void* qkv_bias;         /* [3*H*Dh] fp16 or NULL (plugin passes NULL, gptAttentionCommon.cpp:723) */
    void* out;                    /* [B, H*Dh] fp16 */
    void* kv_cache;               /* [B, 2, H, Smax, Dh] int8 or fp16; read for t < length, written at t = length */
    const int32_t* sequence_lengths; /* [B] device: tokens already in the cache per sequence (length_per_sample);
                                        NULL => past_kv_length for every sequence */
    const int32_t* masked_tokens; /* [B, Smax] device, nonzero = key excluded (padding); may be NULL */
    const float* kv_scale_orig_quant; /* [1] device fp32, used when cache is int8 */
    const float* kv_scale_quant_orig; /* [1] device fp32 */
    int32_t batch_size;           /* B (beam width 1) */
    int32_t num_heads;            /* H */
    int32_t head_size;            /* Dh: 64 */
    int32_t max_seq_len;          /* Smax (memory_max_len) */
    int32_t past_kv_length;       /* host scalar, timestep */
    int32_t int8_kv_cache;        /* 1: int8 cache, 0: fp16 cache */
    float q_scaling;              /* inv_sqrt_dh = 1 / (sqrt(Dh) * q_scaling) */
} b200_mmha_params;

int b200_mmha_generation(const b200_mmha_params* params, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Generation phase, first two operators of a decoder layer as ONE kernel:
 *     LayerNorm -> fused qkv projection (int8 weight-only + bias) -> masked self-attention over the int8 KV cache.
 * Replaces, for the decoder step, the sequence  weight-only matmul plugin enqueue (weightOnlyQuantMatmulPlugin.cpp:162-222)
 * -> bias layer -> GPTAttention plugin enqueueGeneration (gptAttentionCommon.cpp:649-780)  of  T/tensorrt_llm/models/whisper/
 * model.py:74-118, i.e. b200_woq_int8_gemm_ln_folded + b200_mmha_generation here: both operators are head-local, so a
 * thread-block cluster per head computes the head's 192 projection columns (split over k), exchanges the partial sums
 * through distributed shared memory and runs the head's attention on the result -- one kernel boundary and one round
 * trip of q / k / v through L2 fewer per layer (csrc/qkv_mmha.cu).
 * x [B, d] fp16 raw residual rows (d = num_heads * 64); ln_gamma [d] fp16; c1s / c2 [3d] fp32 from b200_woq_ln_fold_prepare;
 * Wproc / scales / bias: the qkv Linear (N = 3d) as for b200_woq_int8_gemm_ln_folded; kv_cache [B, 2, H, max_seq_len, 64]
 * int8 (row sequence_lengths[b] is written); out [B, d] fp16.  batch_size <= 16.  b200_qkv_mmha_decode_supported tells
 * whether a shape is handled (otherwise call the two operators).
 * ---------------------------------------------------------------------------------------------- */
int b200_qkv_mmha_decode_supported(int batch_size, int num_heads, int head_size);
int b200_qkv_mmha_decode(const void* x, const void* ln_gamma, const float* c1s, const float* c2, float ln_eps,
    const int8_t* Wproc, const void* scales, const void* bias, void* kv_cache, const int32_t* sequence_lengths,
    const float* kv_scale_orig_quant, const float* kv_scale_quant_orig, void* out, int batch_size, int num_heads,
    int head_size, int max_seq_len, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Context (prompt) phase: causal attention over S tokens per sequence + KV cache fill (optionally int8).
 * Replaces  GPTAttentionPluginCommon::enqueueContext  T/cpp/tensorrt_llm/plugins/gptAttentionCommon/gptAttentionCommon.cpp:361-620
 *           (add_fusedQKV_bias_transpose, transpose4dBatchMajorKVCache T/cpp/tensorrt_llm/kernels/unfusedAttentionKernels.cu:1106-1490,1552-1646,
 *            cuBLAS QK^T, softmax_kernel :179-257, cuBLAS PV, transpose) -- nine launches there, one here.
 * qkv [B, S, 3*H*Dh] fp16; input_lengths [B] device (tokens >= length are padding: their outputs are
 * unspecified and they are not attended to); out [B, S, H*Dh]; kv_cache as above, rows [0, S) written.
 * ---------------------------------------------------------------------------------------------- */
int b200_attention_context(const void* qkv, const int32_t* input_lengths, void* out, void* kv_cache,
    const float* kv_scale_orig_quant, int batch_size, int seq_len, int num_heads, int head_size, int max_seq_len,
    int int8_kv_cache, float q_scaling, b200_stream_t stream);

/* The same two operators over a PAGED KV cache (KVBlockArray, T/cpp/tensorrt_llm/kernels/kvCacheUtils.h:34-112; plugin
 * inputs: gptAttentionPlugin.cpp:314-326): block_pointers is a device array [B, 2, max_blocks_per_seq] of device pointers
 * (K table then V table per sequence, beam width 1), every block laid out [num_heads, tokens_per_block, head_size] in the
 * cache dtype; tokens_per_block is a power of 2 and max_blocks_per_seq * tokens_per_block >= max_seq_len.  Only the
 * blocks up to the token being written have to be allocated.  p->kv_cache is ignored.  Results and cache bytes are
 * identical with the linear-buffer entry points. */
int b200_mmha_generation_paged(const b200_mmha_params* p, const void* const* block_pointers, int max_blocks_per_seq,
    int tokens_per_block, b200_stream_t stream);
int b200_attention_context_paged(const void* qkv, const int32_t* input_lengths, void* out,
    const void* const* block_pointers, int max_blocks_per_seq, int tokens_per_block, const float* kv_scale_orig_quant,
    int batch_size, int seq_len, int num_heads, int head_size, int max_seq_len, int int8_kv_cache, float q_scaling,
    b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Cached cross-attention over the encoder frames with an int8 (or fp16) cross-KV cache
 * [B, 2, H, S_enc, Dh] produced once per utterance by the cross_kv_cache_warping model.
 * Reference: unfused path  T/tensorrt_llm/layers/attention.py:308-323,385-406 and CrossAttn_KV
 * T/tensorrt_llm/models/whisper/model.py:469-555 (fp16 there; the int8 cache follows the MMHA int8 KV
 * convention, decoderMaskedMultiheadAttentionUtils.h:2276-2286,2357-2390).  No reference CUDA kernel exists
 * (Template.h:1283); the oracle is torch_model.py:88-103 with int8 round-tripped K/V.
 * q [R, H*Dh] fp16; out [R, H*Dh] fp16; softmax(q K^T / sqrt(Dh)) V over all kv_len keys, no mask.
 * R = num_q_rows query rows; row r attends to cache sequence r / q_rows_per_seq (1 per sequence in the
 * generation phase, S_prompt per sequence in the context phase).  Workspace holds split-KV partials
 * (size it with batch_size = R).
 * ---------------------------------------------------------------------------------------------- */
size_t b200_cross_attention_workspace_bytes(int batch_size, int num_heads, int head_size, int kv_len);
int b200_cross_attention(const void* q, const void* cross_kv, const float* kv_scale_quant_orig, void* out,
    int num_q_rows, int q_rows_per_seq, int num_heads, int head_size, int kv_len, int int8_kv_cache, void* workspace,
    size_t workspace_bytes, b200_stream_t stream);
/* Cross attention of the generation phase WITH its q projection (one launch instead of two per layer): the attention
 * CTAs compute q = Linear_q(LayerNorm(x)) for their own (row, head) pairs -- the `cross_attn.q` weight-only Linear of
 * T/tensorrt_llm/layers/attention.py:308-313 behind the decoder layer's cross_attention_layernorm
 * (T/tensorrt_llm/models/whisper/model.py:257-292) -- so the cross-KV stream starts one kernel boundary earlier and the
 * projection runs while the first ring stages are in flight (csrc/attention.cu, cross_attention_qproj_kernel).
 * x [B, d] fp16 raw residual rows (d = num_heads * 64); ln_gamma [d]; c1s / c2 [d] fp32 from b200_woq_ln_fold_prepare;
 * Wproc / scales / bias: the q Linear (N = K = d) as for b200_woq_int8_gemm_ln_folded; cross_kv: the int8 cache written
 * by b200_cross_kv_pack, which must NOT be written by the kernel launched right before (it is streamed before the
 * dependency wait); out [B, d] fp16.  b200_cross_attention_qproj_supported tells whether a shape is handled (at least
 * one (row, head) pair per SM, at most 6 rows per CTA); otherwise call the two operators. */
int b200_cross_attention_qproj_supported(int batch_size, int num_heads, int head_size, int kv_len);
int b200_cross_attention_qproj(const void* x, const void* ln_gamma, const float* c1s, const float* c2, float ln_eps,
    const int8_t* Wproc, const void* scales, const void* bias, const void* cross_kv, const float* kv_scale_quant_orig,
    void* out, int batch_size, int num_heads, int head_size, int kv_len, b200_stream_t stream);
/* Tuning switch of b200_cross_attention (process-wide bit mask; set it before launches are captured in a graph; returns
 * the previous value).  Bit 0: when the (row, head) pairs do not fill whole rounds of the grid (320 pairs on 148 CTAs leave
 * 24), the kernel runs as clusters of two CTAs that share one left-over pair each, split by keys and merged through
 * distributed shared memory, instead of dealing those pairs whole to the first CTAs (default off: measured faster as a
 * kernel and slower in the decoder step).  Bit 1: with fewer pairs than half the SMs (batch 1-3) EVERY pair is shared by a
 * cluster of 4 or 2 CTAs instead of one CTA per pair (default off as well: 0.991 vs 0.932 ms per batch-1 step).  Same
 * results up to the order of the softmax merge.  Default 0; env B200_XA_SPLIT=<mask>. */
int b200_set_cross_attention_split(int enabled);
/* Debug aid (library built with B200_XA_DEBUG=1; a no-op otherwise): device buffer of >= 8 int64 per SM receiving
 * %globaltimer stamps of every CTA of the following whole-pair cross-attention launches (entry, dependency return,
 * first warp done with pair 0 / 1 / 2 / 3 / later, exit); NULL switches it off (tools/xa_timeline.py). */
int b200_debug_xa_timeline(void* device_buffer);
/* Packs fp16 K and V projections [B, S, H*Dh] into the cross-KV cache layout, quantizing with
 * cvt.rni.sat.s8.f32(scale * x) when int8_kv_cache (same rule as the self-attention cache).
 * The int8 CROSS cache is stored in offset-binary form: byte = (uint8)(q + 128), i.e. the two's-complement byte with
 * its sign bit flipped (the same +128 bias the reference applies to int8 weights, cutlass_preprocessors.cpp:470-506).
 * The values are identical; the form lets the attention kernel turn bytes into fp16 without a sign fix-up.  The
 * reference has no int8 cross cache (its cross K/V are fp16 engine tensors), so this layout is private to the
 * b200_cross_kv_pack -> b200_cross_attention pair; the self-attention cache keeps plain int8. */
int b200_cross_kv_pack(const void* k, const void* v, void* cross_kv, const float* kv_scale_orig_quant, int batch_size,
    int kv_len, int num_heads, int head_size, int int8_kv_cache, b200_stream_t stream);

/* x [B, C, T] fp16 -> y [B, T, C] = x^T + pos [T, C]: the encoder's permute + positional embedding
 * (T/tensorrt_llm/models/whisper/model.py:158-162). */
int b200_transpose_add_pos_fp16(const void* x, const void* pos, void* y, int batch_size, int channels, int t,
    b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Bidirectional (encoder) attention, head size 64: out = softmax(q k^T / 8) v per (batch, head), no mask, no cache.
 * Replaces the unfused encoder attention of T/tensorrt_llm/layers/attention.py:283-406 as used by
 * T/tensorrt_llm/models/whisper/model.py:124-172 (oracle W/torch_model.py:88-103).
 * qkv [B, S, 3*H*64] fp16 (q | k | v, the output of one fused projection); out [B, S, H*64] fp16.
 * ---------------------------------------------------------------------------------------------- */
int b200_attention_bidirectional_fp16(const void* qkv, void* out, int batch_size, int seq_len, int num_heads,
    int head_size, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Whisper logit filters + greedy token update on the device (one launch per step, graph-capturable).
 * Replaces the per-sequence Python loops of T/examples/whisper/decoding.py: SuppressBlank :202-209, SuppressTokens
 * :212-217, ApplyTimestampRules :134-199, GreedyDecoder.update :274-293 (temperature 0).
 * logits [rows, vocab] fp32 (not modified); suppress_bitmap: bit v of word v/32 set = token v suppressed (may be NULL);
 * no_timestamps / max_initial_timestamp_index: -1 = absent; decode_state [rows][4] int32, zero-initialised when a
 * sequence starts (sampled count, last, penultimate, last timestamp + 1), advanced by the call;
 * next_token [rows]; sum_logprobs [rows] fp32 accumulated in place (may be NULL).
 * ---------------------------------------------------------------------------------------------- */
int b200_whisper_filtered_argmax(const float* logits, int rows, int vocab, const uint32_t* suppress_bitmap, int eot,
    int no_timestamps, int timestamp_begin, int blank_token, int max_initial_timestamp_index, int32_t* decode_state,
    int32_t* next_token, float* sum_logprobs, b200_stream_t stream);

/* Language detection and the no-speech probability, both from the logits of the start-of-transcript position
 * (T/examples/whisper/decoding.py:703-741: logits outside the language tokens set to -inf, argmax, softmax; :762-766:
 * softmax over the whole vocabulary at the nospeech token).  logits [rows, vocab] fp32; [range_lo, range_hi) = the
 * language tokens (contiguous ids); range_argmax [rows] int32, range_probs [rows, range_hi - range_lo] fp32,
 * probe_prob [rows] fp32 -- each output may be NULL. */
int b200_logits_range_softmax(const float* logits, int rows, int vocab, int range_lo, int range_hi, int probe_token,
    int32_t* range_argmax, float* range_probs, float* probe_prob, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Conv1d encoder stem (fp16, fp32 accumulate), optional fused GELU.
 * Replaces  functional.conv1d -> TensorRT IConvolutionLayer  T/tensorrt_llm/functional.py:2202-2244,
 *           layers.Conv1d T/tensorrt_llm/layers/conv.py:52-94 (use: T/tensorrt_llm/models/whisper/model.py:135-157).
 * x [B, Cin, T]; w [Cout, Cin, ksize] (the reference's [out, in, k, 1] with the trailing 1 dropped);
 * bias [Cout] or NULL; y [B, Cout, Tout], Tout = (T + 2*pad - ksize) / stride + 1.
 * ---------------------------------------------------------------------------------------------- */
int b200_conv1d_fp16(const void* x, const void* w, const void* bias, void* y, int batch_size, int c_in, int c_out,
    int t_in, int ksize, int stride, int pad, int activation, b200_stream_t stream);
/* The same operator as an implicit GEMM on the tensor cores (tcgen05, TMEM accumulator): the weights are re-laid to
 * [tap][Cout][Cin_pad] and the input is transposed to time-major [B][T][Cin_pad] inside the workspace (Cin_pad = Cin
 * rounded up to 64), then every k-block's im2col tile is ONE 3-D TMA box whose row traversal stride is the convolution
 * stride and whose out-of-bounds rows are the zero padding.  workspace: b200_conv1d_workspace_bytes(), 256-byte
 * aligned.  ksize 1..8, stride 1 or 2. */
size_t b200_conv1d_workspace_bytes(int batch_size, int c_in, int c_out, int t_in, int ksize);
int b200_conv1d_fp16_tc(const void* x, const void* w, const void* bias, void* y, int batch_size, int c_in, int c_out,
    int t_in, int ksize, int stride, int pad, int activation, void* workspace, size_t workspace_bytes,
    b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Whisper log-Mel front end (the data format in front of the encoder stem).
 * Replaces  log_mel_spectrogram  T/examples/whisper/whisper_utils.py:99-145 (host torch.stft in the reference,
 *           callers run.py:44-46, summarize.py:120-122).
 * audio [B, n_samples] fp32 on the device, 16 kHz; `padding` zero samples are appended to every utterance (:129-130);
 * mel_filters [n_mels, 201] fp32 on the device (the reference's assets/mel_filters.npz, whisper_utils.py:81-97);
 * out [B, n_mels, n_frames], n_frames = (n_samples + padding) / 160, dtype B200_DTYPE_F32 (the reference's result) or
 * B200_DTYPE_F16 (what the encoder consumes, run.py:45).  The max - 8 floor uses the maximum of each utterance: the
 * reference is called with one utterance at a time.  n_samples + padding > 200 (reflect padding of torch.stft).
 * workspace: b200_log_mel_workspace_bytes(), 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
int b200_log_mel_frames(int n_samples, int padding);
size_t b200_log_mel_workspace_bytes(int batch_size, int n_samples, int padding, int n_mels);
int b200_log_mel_spectrogram(const float* audio, int batch_size, int n_samples, int padding, const float* mel_filters,
    int n_mels, void* out, int out_dtype, void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder-step glue (TensorRT-native layers in the reference:
 * T/tensorrt_llm/models/whisper/model.py:74-118,257-292).
 * ---------------------------------------------------------------------------------------------- */
/* y = LayerNorm(x) * gamma + beta over the last dim; fp32 statistics like torch_model.py:25-27. */
int b200_layernorm_fp16(const void* x, const void* gamma, const void* beta, void* y, int rows, int cols, float eps,
    b200_stream_t stream);
/* out[r, :] = tok_emb[tokens[r], :] + pos_emb[positions[r], :]   (torch_model.py:205-209) */
int b200_embed_tokens_fp16(const int32_t* tokens, const int32_t* positions, const void* tok_emb, const void* pos_emb,
    void* out, int rows, int cols, int vocab, int n_ctx, b200_stream_t stream);
/* logits[r, v] = sum_k x[r, k] * emb[v, k]  (fp16 weights, not quantized: model.py:231,290), optionally also
 * next_token[r] = argmax_v logits[r, v] (first index on ties, like torch.argmax on the fp32 logits).
 * logits may be NULL when only the argmax is wanted.  workspace: b200_logits_workspace_bytes(). */
size_t b200_logits_workspace_bytes(int rows, int vocab);
int b200_logits_argmax_fp16(const void* x, const void* emb, void* logits_fp32, int32_t* next_token, int rows,
    int cols, int vocab, void* workspace, size_t workspace_bytes, b200_stream_t stream);
/* 0 = auto (tcgen05 fp16 GEMM), 1 = SIMT (tests / comparison). */
int b200_logits_set_kernel_policy(int policy);

/* ------------------------------------------------------------------------------------------------
 * Persistent decoder step: the WHOLE stack of one generation step (token + position embedding, then per layer
 * LN->qkv, masked self-attention with int8 KV append, out-projection, LN->cross-q, cached cross-attention over the
 * int8 cross-KV, out-projection, LN->fc1+GELU, fc2; rows = batch <= 16) as ONE kernel launch.
 * Replaces, for the generation phase, the per-operator chain the reference engine executes for a decoder step
 *   WeightOnlyQuantMatmulPlugin::enqueue   T/cpp/tensorrt_llm/plugins/weightOnlyQuantMatmulPlugin/weightOnlyQuantMatmulPlugin.cpp:162-222   (x6 per layer)
 *   GPTAttentionPluginCommon::enqueueGeneration  T/cpp/tensorrt_llm/plugins/gptAttentionCommon/gptAttentionCommon.cpp:649-780
 *   cross attention                        T/tensorrt_llm/layers/attention.py:308-323,385-406
 *   graph of WhisperDecoder                T/tensorrt_llm/models/whisper/model.py:74-118,257-292
 * with the same arithmetic as the per-operator entry points above (same weight layout, folded LayerNorm vectors from
 * b200_woq_ln_fold_prepare, same KV-cache layouts and quantization rules, same epilogue rounding).
 *
 * One CTA per SM stays resident for the whole step.  A producer warp streams every int8 weight tile and every
 * cross-KV chunk of the step through ONE shared-memory ring with 1-D TMA bulk copies, in a static order, as far
 * ahead as the ring allows (weights and cross-KV never depend on the step's activations), so HBM keeps streaming
 * while the consumer warps sit in the dependency chain; the ten consumer warps run the matmuls (in-register
 * int8->fp16 dequant, mma.sync m16n8k16 with the 16 batch rows as M, K split over warps), the attention kernels'
 * inner loops and the epilogues.  Phases are separated by grid-wide release/acquire barriers on `sync`.
 *
 * Activations that feed a matmul are kept in "A-fragment order" ([K/64][4][32 lanes][8 halves], 16 rows): the
 * consumer warps load their mma.sync A operands with coalesced 128-bit loads, no shared-memory staging.
 * All scratch buffers are caller-owned (b200_decoder_step_scratch_bytes); `sync` must be zero before the first call
 * and is left zero by every call.
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200_decoder_layer
{
    /* LayerNorm -> qkv [d -> 3d] (bias: q | 0 | v), folded-LN vectors from b200_woq_ln_fold_prepare */
    const void* attn_ln_gamma; const int8_t* qkv_w; const void* qkv_scales; const void* qkv_bias;
    const float* qkv_c1s; const float* qkv_c2;
    const int8_t* attn_out_w; const void* attn_out_scales; const void* attn_out_bias;
    const void* cross_ln_gamma; const int8_t* cross_q_w; const void* cross_q_scales; const void* cross_q_bias;
    const float* cross_q_c1s; const float* cross_q_c2;
    const int8_t* cross_out_w; const void* cross_out_scales; const void* cross_out_bias;
    const void* mlp_ln_gamma; const int8_t* fc1_w; const void* fc1_scales; const void* fc1_bias;
    const float* fc1_c1s; const float* fc1_c2;
    const int8_t* fc2_w; const void* fc2_scales; const void* fc2_bias;
    void* self_kv;                  /* [B, 2, H, Smax, 64] int8 (KVLinearBuffer) */
    const float* kv_scale_orig_quant; const float* kv_scale_quant_orig;
    const void* cross_kv;           /* [B, 2, H, S_enc, 64] int8, offset-binary (b200_cross_kv_pack) */
    const float* cross_kv_scale_quant_orig;
} b200_decoder_layer;

typedef struct b200_decoder_step_params
{
    const b200_decoder_layer* layers; /* DEVICE array [n_layers] */
    int32_t n_layers, batch_size, num_heads, d_ff, max_seq_len, enc_len, vocab, n_ctx;
    const int32_t* tokens;           /* [B] device: input token ids of this step */
    const int32_t* sequence_lengths; /* [B] device: tokens already in the self KV cache (= position of this token) */
    const void* tok_emb;             /* [vocab, d] fp16 */
    const void* pos_emb;             /* [n_ctx, d] fp16 */
    void* x_out;                     /* [B, d] fp16 row-major: the residual stream after the last layer */
    void* scratch;                   /* b200_decoder_step_scratch_bytes(), zero-initialised once */
    float ln_eps;
    int32_t max_ctas;                /* 0 = one per SM; tests use smaller grids */
} b200_decoder_step_params;

size_t b200_decoder_step_scratch_bytes(int num_heads, int d_ff);
int b200_decoder_step(const b200_decoder_step_params* params, b200_stream_t stream);
/* Debug: error word of the last steps run on this scratch (0 = no barrier / ring wait timed out); synchronises. */
int b200_decoder_step_status(const void* scratch, int32_t* status_host);
/* Debug aid: device buffer of n_ctas * 512 * 2 int64 that receives %globaltimer stamps of every following step launch
 * (per CTA and phase: [0] grid barrier passed, [1] phase work done); NULL switches it off (tools/step_phases.py). */
int b200_debug_decoder_step_timeline(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* B200_WHISPER_H */
