/*
 * oracle/woq_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, gcc) of the reference's int8 weight-only quantizer, its
 * "preprocess for mixed gemm" layout transform, and the arithmetic contract of the
 * fp16 x int8 matmul.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may link or call this file.  The product path (csrc/) never does.
 *
 * Paths below are relative to /root/reference/tensorrt_llm_july-release-v1/ (T/).
 *
 *   symmetric_quantize (int8 branch)        T/cpp/tensorrt_llm/kernels/cutlass_kernels/cutlass_preprocessors.cpp:615-721
 *   permute_B_rows_for_mixed_gemm           same file :158-219
 *   subbyte_transpose (int8)                same file :225-362
 *   interleave_column_major_tensor          same file :478-535   (Sm80 traits: ColumnMajorTileInterleave<64,2>,
 *                                           T/cpp/tensorrt_llm/cutlass_extensions/include/cutlass_extensions/gemm/kernel/mixed_gemm_B_layout.h:66-79)
 *   add_bias_and_interleave_int8s_inplace   same file :383-405
 *   preprocess_weights_for_mixed_gemm       same file :537-578
 *   GEMV arithmetic                         T/cpp/tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.cu:44-53,165-203
 *   CUTLASS GEMM arithmetic                 T/cpp/tensorrt_llm/cutlass_extensions/include/cutlass_extensions/gemm/warp/mma_tensorop_dequantizer.h:253-270
 *
 * Pinning: validated bit-for-bit against the reference's own cutlass_preprocessors.cpp compiled
 * from /root/reference (oracle/_ref, see oracle/Makefile) and against the fixtures under
 * tests/golden/ generated from that binary (tests/golden/make_quant_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef _Float16 f16;

static inline f16 bits_to_f16(uint16_t b)
{
    f16 h;
    memcpy(&h, &b, 2);
    return h;
}

static inline uint16_t f16_to_bits(f16 h)
{
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}

/* ---------------------------------------------------------------------------------------------
 * Step 1 of the layout transform: within every group of 16 rows (k), row r of the output takes
 * row map[r] of the input, map = 0 1 8 9 2 3 10 11 4 5 12 13 6 7 14 15.
 * Follows cutlass_preprocessors.cpp:158-219 (int8: B_ROWS_PER_MMA = 16, ELTS_PER_REG = 4).
 * in/out: [K][N] row-major int8.
 * ------------------------------------------------------------------------------------------- */
void oracle_permute_B_rows_int8(int8_t* out, const int8_t* in, int K, int N)
{
    for (int base = 0; base < K; base += 16)
    {
        for (int r = 0; r < 16; ++r)
        {
            const int src = 8 * ((r % 4) / 2) + (r % 2) + 2 * (r / 4);
            memcpy(out + (size_t) (base + r) * N, in + (size_t) (base + src) * N, (size_t) N);
        }
    }
}

/* Step 2: [K][N] -> [N][K].  Follows cutlass_preprocessors.cpp:225-362 (int8 branch is a plain
 * byte transpose done in 64x64 cache tiles; the tiling does not change the result). */
void oracle_transpose_int8(int8_t* out, const int8_t* in, int K, int N)
{
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n)
            out[(size_t) n * K + k] = in[(size_t) k * N + n];
}

/* Step 3: interleave 2 columns in tiles of 64 rows.  Input is the column-major tensor, i.e. the
 * [N][K] byte array of step 2 (the reference still calls K "rows" and N "cols").
 * Follows cutlass_preprocessors.cpp:478-535 with rows_per_tile = 64, interleave = 2, operating on
 * 32-bit words (4 int8) exactly as the reference does. */
void oracle_interleave_column_major_int8(int8_t* out, const int8_t* in, int K, int N)
{
    const uint32_t* src = (const uint32_t*) in;
    uint32_t* dst = (uint32_t*) out;
    const int num_vec_rows = K / 4;
    const int vec_rows_per_tile = 64 / 4;
    const int interleave = 2;
    for (int read_col = 0; read_col < N; ++read_col)
    {
        const int64_t write_col = read_col / interleave;
        for (int base_vec_row = 0; base_vec_row < num_vec_rows; base_vec_row += vec_rows_per_tile)
        {
            const int lim = base_vec_row + vec_rows_per_tile < num_vec_rows ? base_vec_row + vec_rows_per_tile
                                                                            : num_vec_rows;
            for (int vec_read_row = base_vec_row; vec_read_row < lim; ++vec_read_row)
            {
                const int64_t vec_write_row = (int64_t) interleave * base_vec_row
                    + (int64_t) vec_rows_per_tile * (read_col % interleave) + vec_read_row % vec_rows_per_tile;
                const int64_t read_offset = (int64_t) read_col * num_vec_rows + vec_read_row;
                const int64_t write_offset = write_col * num_vec_rows * interleave + vec_write_row;
                dst[write_offset] = src[read_offset];
            }
        }
    }
}

/* Step 4: add 128 to every int8 (-> biased uint8) and swap bytes 1 and 2 of every 32-bit word.
 * Follows cutlass_preprocessors.cpp:383-405. */
void oracle_add_bias_and_interleave_int8_inplace(int8_t* buf, size_t num_elts)
{
    for (size_t i = 0; i < num_elts; ++i)
        buf[i] = (int8_t) ((int) buf[i] + 128);
    for (size_t base = 0; base + 3 < num_elts; base += 4)
    {
        int8_t t = buf[base + 1];
        buf[base + 1] = buf[base + 2];
        buf[base + 2] = t;
    }
}

/* The whole chain for Sm80..Sm90 layout details (row permute yes, column major yes, interleave 2).
 * Follows cutlass_preprocessors.cpp:537-578.  raw: [K][N] int8; proc: K*N bytes.
 * Returns 0, or -1 if the shape violates the reference's checks (K % 16, N % 64 -> :187-192, :498-500). */
int oracle_preprocess_weights_int8(int8_t* proc, const int8_t* raw, int K, int N)
{
    if (K % 64 != 0 || N % 64 != 0)
        return -1;
    const size_t bytes = (size_t) K * N;
    int8_t* a = (int8_t*) malloc(bytes);
    int8_t* b = (int8_t*) malloc(bytes);
    if (!a || !b)
    {
        free(a);
        free(b);
        return -2;
    }
    oracle_permute_B_rows_int8(a, raw, K, N);
    oracle_transpose_int8(b, a, K, N);
    oracle_interleave_column_major_int8(a, b, K, N);
    oracle_add_bias_and_interleave_int8_inplace(a, bytes);
    memcpy(proc, a, bytes);
    free(a);
    free(b);
    return 0;
}

/* Closed form of the same transform: for byte offset o in row j of the
 * [N/2][2K] output, which (k, n) of the raw matrix it holds.  Used to cross-check the chain. */
void oracle_preprocess_closed_form_int8(int8_t* proc, const int8_t* raw, int K, int N)
{
    static const int P[16] = {0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15};
    static const int SW[4] = {0, 2, 1, 3};
    for (int j = 0; j < N / 2; ++j)
    {
        for (int o = 0; o < 2 * K; ++o)
        {
            const int t = o / 128, w = o % 128;
            const int n = 2 * j + w / 64;
            const int kk = w % 64;
            const int kp = 64 * t + 4 * (kk / 4) + SW[kk % 4];
            const int k = 16 * (kp / 16) + P[kp % 16];
            proc[(size_t) j * 2 * K + o] = (int8_t) ((int) raw[(size_t) k * N + n] + 128);
        }
    }
}

/* Inverse: processed bytes -> raw [K][N] int8 (what test_conversion's identity activation recovers,
 * T/tests/quantization/test_weight_only_quant_matmul.py:121-130). */
void oracle_unprocess_int8(int8_t* raw, const int8_t* proc, int K, int N)
{
    static const int P[16] = {0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15};
    static const int SW[4] = {0, 2, 1, 3};
    for (int j = 0; j < N / 2; ++j)
    {
        for (int o = 0; o < 2 * K; ++o)
        {
            const int t = o / 128, w = o % 128;
            const int n = 2 * j + w / 64;
            const int kk = w % 64;
            const int kp = 64 * t + 4 * (kk / 4) + SW[kk % 4];
            const int k = 16 * (kp / 16) + P[kp % 16];
            raw[(size_t) k * N + n] = (int8_t) ((int) (uint8_t) proc[(size_t) j * 2 * K + o] - 128);
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * symmetric_quantize, int8 branch, per column of W[K][N].
 * Follows cutlass_preprocessors.cpp:641-687: amax in fp32, scale_f32 = amax * (1/128) (fp32 multiply),
 * stored scale = ComputeType(scale_f32); q = int8(clamp(round(w / scale_f32), -128, 127)) where
 * round() is C round (half away from zero) and the division is fp32.
 * w_is_f16: weights are fp16 bit patterns (uint16_t) else fp32.
 * scale_is_f16: scales written as fp16 bit patterns (uint16_t) else fp32.
 * proc may be NULL (raw only).
 * ------------------------------------------------------------------------------------------- */
int oracle_symmetric_quantize_int8(const void* w, int w_is_f16, int K, int N, int8_t* raw, int8_t* proc,
    void* scales, int scale_is_f16)
{
    float* col_scale = (float*) malloc(sizeof(float) * (size_t) N);
    if (!col_scale)
        return -2;
    for (int n = 0; n < N; ++n)
        col_scale[n] = 0.f;
    const uint16_t* w16 = (const uint16_t*) w;
    const float* w32 = (const float*) w;
    for (int k = 0; k < K; ++k)
    {
        for (int n = 0; n < N; ++n)
        {
            const float v = w_is_f16 ? (float) bits_to_f16(w16[(size_t) k * N + n]) : w32[(size_t) k * N + n];
            const float a = fabsf(v);
            if (a > col_scale[n])
                col_scale[n] = a;
        }
    }
    const float quant_range_scale = 1.f / 128.f;
    for (int n = 0; n < N; ++n)
    {
        col_scale[n] *= quant_range_scale;
        if (scale_is_f16)
            ((uint16_t*) scales)[n] = f16_to_bits((f16) col_scale[n]);
        else
            ((float*) scales)[n] = col_scale[n];
    }
    for (int k = 0; k < K; ++k)
    {
        for (int n = 0; n < N; ++n)
        {
            const float v = w_is_f16 ? (float) bits_to_f16(w16[(size_t) k * N + n]) : w32[(size_t) k * N + n];
            const float scaled = roundf(v / col_scale[n]);
            const float clipped = fmaxf(-128.f, fminf(127.f, scaled));
            raw[(size_t) k * N + n] = (int8_t) clipped;
        }
    }
    free(col_scale);
    if (proc)
        return oracle_preprocess_weights_int8(proc, raw, K, N);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Arithmetic contract of the fp16 x int8 matmul.
 *   effective weight  w16[k][n] = fp16( fp16(q[k][n]) * s16[n] )      (both reference kernels)
 *   mode 0 ("cutlass"): C = fp16( sum_k fp32(a16) * fp32(w16) )         exact products, fp32 accumulate
 *   mode 1 ("gemv")   : C = fp16( sum_k fp32( fp16(a16 * w16) ) )       product rounded to fp16 first
 *   mode 2 ("ideal")  : C = fp16( double sum_k a16 * w16 )              rounding-order independent centre
 *   mode 3 ("gemv_exact"): mode 1 with the reference GEMV kernel's exact summation order
 *                       (weightOnlyMatrixVectorMultiplication.cu:165-203): 16 lanes per weight column, lane (o, q)
 *                       (o = lane/8, q = lane%4) sums k = 64*(o + 4j) + 16q + p serially over j, p in fp32; then the
 *                       butterfly xor 16, 8, 2, 1 = ((T0+T2)+(T1+T3)) over o, then ((V0+V2)+(V1+V3)) over q.
 *                       Pinned bit-exactly against the reference kernel running on B200
 *                       (tests/test_reference_kernels_gpu.py).
 * Accumulation order of modes 0-2 is plain k = 0..K-1 (the GPU kernels use other orders; tests compare with a
 * tolerance, see tests/test_woq_matmul.py).
 * A: [M][K] fp16 bits, raw: [K][N] int8, scales: [N] fp16 bits, C: [M][N] fp16 bits.
 * ------------------------------------------------------------------------------------------- */
void oracle_woq_matmul(const uint16_t* A, int M, int K, const int8_t* raw, const uint16_t* scales, int N,
    uint16_t* C, int mode)
{
    f16* wcol = (f16*) malloc(sizeof(f16) * (size_t) K);
    for (int n = 0; n < N; ++n)
    {
        const f16 s = bits_to_f16(scales[n]);
        for (int k = 0; k < K; ++k)
        {
            const f16 q = (f16) (float) raw[(size_t) k * N + n];
            wcol[k] = (f16) ((float) q * (float) s);
        }
        for (int m = 0; m < M; ++m)
        {
            float acc = 0.f;
            double accd = 0.0;
            if (mode == 3)
            {
                float T[4][4];
                for (int o = 0; o < 4; ++o)
                    for (int q = 0; q < 4; ++q)
                    {
                        float v = 0.f;
                        for (int blk = o; blk * 64 < K; blk += 4)
                            for (int p16 = 0; p16 < 16; ++p16)
                            {
                                const int k = 64 * blk + 16 * q + p16;
                                const f16 a = bits_to_f16(A[(size_t) m * K + k]);
                                v += (float) (f16) ((float) a * (float) wcol[k]);
                            }
                        T[o][q] = v;
                    }
                float V[4];
                for (int q = 0; q < 4; ++q)
                    V[q] = (T[0][q] + T[2][q]) + (T[1][q] + T[3][q]);
                acc = (V[0] + V[2]) + (V[1] + V[3]);
                C[(size_t) m * N + n] = f16_to_bits((f16) acc);
                continue;
            }
            for (int k = 0; k < K; ++k)
            {
                const f16 a = bits_to_f16(A[(size_t) m * K + k]);
                if (mode == 0)
                    acc += (float) a * (float) wcol[k];
                else if (mode == 1)
                    acc += (float) (f16) ((float) a * (float) wcol[k]);
                else
                    accd += (double) (float) a * (double) (float) wcol[k];
            }
            C[(size_t) m * N + n] = f16_to_bits(mode == 2 ? (f16) accd : (f16) acc);
        }
    }
    free(wcol);
}

/* int8 KV cache quantisation of one value: cvt.rni.sat.s8.f32(scale * float(x16)).
 * Follows T/cpp/tensorrt_llm/kernels/decoderMaskedMultiheadAttentionUtils.h:2276-2286,2382-2390 and the
 * expected value of T/cpp/tests/runtime/transposeKVKernelTest.cpp:79-84. */
void oracle_kv_quantize_int8(const uint16_t* x, size_t n, float scale_orig_quant, int8_t* out)
{
    for (size_t i = 0; i < n; ++i)
    {
        float v = scale_orig_quant * (float) bits_to_f16(x[i]);
        float r = nearbyintf(v); /* default rounding mode = round-half-even, as cvt.rni */
        if (r > 127.f)
            r = 127.f;
        if (r < -128.f)
            r = -128.f;
        if (v != v)
            r = 0.f; /* cvt.*.sat of NaN gives 0 */
        out[i] = (int8_t) r;
    }
}

/* int8 KV cache dequantisation: half(scale_quant_orig * float(q)).  Follows ...Utils.h:2357-2365. */
void oracle_kv_dequantize_int8(const int8_t* q, size_t n, float scale_quant_orig, uint16_t* out)
{
    for (size_t i = 0; i < n; ++i)
        out[i] = f16_to_bits((f16) (scale_quant_orig * (float) q[i]));
}
