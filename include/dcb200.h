/*
 * dcb200 -- C ABI of the B200-native DynamiCrafter denoising hot path.
 *
 * Nothing like this exists in the reference (pure PyTorch): each entry point
 * replaces a class of ATen / cuDNN / cuBLAS call sites of the reference's hot path, cited
 * below as file:line under /root/reference.  The reference-side binding is a ctypes stub
 * (INTEGRATION.md); the Python mirror classes in dynamicrafter_b200/ are that binding.
 *
 * Conventions: plain pointers and sizes only; every pointer is DEVICE memory unless named
 * h_*; the library never allocates, never synchronises and never throws.  All functions
 * return 0 on success or a negative dcb_status; dcb_last_error() gives a thread-local
 * message.  `stream` is a cudaStream_t passed as void*.  Kernels are stateless, so calls are
 * CUDA-graph capturable and thread-safe when callers use distinct buffers.
 *
 * Activation layout everywhere: channels-last token matrix [rows, C] with rows ordered
 * (batch, frame, y, x); bf16 operands feed the tensor cores, the residual stream is fp32.
 */
#ifndef DCB200_H
#define DCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DCB_OK = 0,
    DCB_ERR_INVALID = -1,   /* bad argument (shape / alignment / null pointer) */
    DCB_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed */
    DCB_ERR_UNSUPPORTED = -3,
    DCB_ERR_DEVICE = -4     /* not an sm_100 device */
} dcb_status;

#define DCB_MAX_TAPS 9

const char* dcb_last_error(void);
int dcb_abi_version(void);
/* 0 if the current device is compute capability 10.x, DCB_ERR_DEVICE otherwise (the reference has no counterpart: it
 * runs wherever torch does; dcb200 has no other path and says so instead of falling back). */
int dcb_check_device(void);

/* ------------------------------------------------------------------------------------------
 * dcb_tap_gemm: tcgen05/TMEM/TMA implicit GEMM,  out[r, n] = epi( sum_{tap,c} A[r + off(tap), c] * W[n, tap*C + c] )
 *
 * Replaces: nn.Linear / 1x1 conv  (attention.py:53-57,75-76,269,290,336,418,438; ae_modules.py:31-50)
 *           nn.Conv2d 3x3 s1 p1   (openaimodel3d.py:154,179,185,96,386,545; ae_modules.py:117,161,173,491,533)
 *           nn.Conv3d (3,1,1)     (openaimodel3d.py:255-266)
 *
 * A is a bf16 tensor viewed as [D3a, D2, D1, C] (C contiguous; strides s1,s2,s3 in elements,
 * multiples of 8).  The output domain is [D3, D2, D1]; an M-tile is a box (b1,b2,b3) of it with
 * b1*b2*b3 <= 128.  For each tap the box is shifted by (o1,o2,o3); reads outside the tensor are
 * zero (TMA out-of-bounds fill) which implements the convolution padding.  W is bf16 [N, ntaps*C]
 * row-major ("K-major").  C % 8 == 0, and C % 64 == 0 when ntaps > 1
 * (a ragged last K chunk is zero-filled by TMA).
 *
 * Epilogue (fp32): acc + bias[n] + group_bias[r / rows_per_group, n]; if geglu, tile columns
 * [0,BN/2) are values and [BN/2,BN) gates and out = v * gelu_erf(g) (N_out = N/2, attention.py:421-422);
 * then + residual[r, n]; written to out_f32 and/or out_bf16 (row stride ld_out elements).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const void* a;            /* bf16 */
    int64_t a_s1, a_s2, a_s3; /* element strides of dims 1..3 of A (dim 0 = channel, stride 1) */
    int32_t C;                /* channels per tap */
    int32_t D1, D2, D3;       /* output domain */
    int32_t A1, A2, A3;       /* extents of A's dims 1..3 (>= domain; larger for phase-plane tricks) */
    int32_t b1, b2, b3;       /* M-tile box */
    int32_t ntaps;
    int32_t taps[DCB_MAX_TAPS][3];
    const void* w;            /* bf16 [N, ntaps*C] */
    int64_t w_ld;             /* row stride of w in elements (0 = ntaps*C) */
    int32_t N;                /* GEMM N (rows of w) */
    int32_t block_n;          /* 128 or 160 (0 = choose) */
    const float* bias;        /* [N_out] or NULL (for geglu: [N], value half then gate half per tile) */
    const float* group_bias;  /* [groups, ld_group_bias] or NULL */
    int32_t rows_per_group;
    int32_t ld_group_bias;
    const float* residual;    /* [rows, ld_residual] fp32 or NULL */
    int32_t ld_residual;
    int32_t geglu;
    float* out_f32;           /* either may be NULL, not both */
    void* out_bf16;
    int32_t ld_out;
    float out_scale;          /* applied to acc before bias (1.0 = none) */
    /* Optional fused GroupNorm statistics of the OUTPUT (NULL = off).  gn_acc: int64 [units, 32, 2], zero-initialised by
     * the caller; the epilogue adds, per unit and per group of gn_cg consecutive output columns, (sum, sum of squares) of
     * the stored values as fixed-point integers (value * 2^20, rounded): integer atomics commute, so the totals are bitwise
     * reproducible.  gn_unit_mode 0: unit = D3 index (a frame of a 2-D conv / linear, a video of a temporal conv);
     * 1: unit = D3 index * D2 + D2 index (a frame of a temporal conv; needs b2 == 1).  Requires every 32-row warp segment
     * of a tile to lie inside one unit: b1*b2 in {8, 16} or a multiple of 32, or b3 == 1 with mode 0.
     * dcb_group_norm_stream_apply consumes the accumulator directly (no statistics pass, no finalize launch). */
    unsigned long long* gn_acc;
    int32_t gn_cg;
    int32_t gn_unit_mode;
    /* Optional LayerNorm folding (NULL = off): the LayerNorm between two GEMMs (attention.py:242-246: norm1 -> to_q/k/v,
     * norm2 -> to_q, norm3 -> GEGLU) never runs as a kernel of its own.
     *   producer (out_f32 given, ln_acc != NULL): besides the fp32 output the epilogue writes a bf16 copy of it to ln_x_bf16
     *     (row stride ld_ln_x) and adds, per output row, (sum * 2^24, sum of squares * 2^28) of the stored fp32 values to
     *     ln_acc (int64 [rows, 2], zero-initialised by the caller; integer atomics: bitwise reproducible).
     *   consumer (out_bf16 given, ln_acc != NULL): A is the producer's raw bf16 copy, W must hold gamma-scaled weights
     *     W'[n,k] = W[n,k] * gamma[k], bias the folded bias b[n] + sum_k W[n,k] beta[k], ln_colsum[n] = sum_k W'[n,k];
     *     the epilogue evaluates  rstd_r * acc[r,n] - rstd_r * mean_r * ln_colsum[n] + bias[n]  with (mean, rstd) of row r
     *     from ln_acc (ln_channels = row length of the normalised tensor, ln_eps as in nn.LayerNorm), then GEGLU if set.
     *     No residual / group_bias / out_scale in this mode. */
    unsigned long long* ln_acc;
    void* ln_x_bf16;
    int32_t ld_ln_x;
    int32_t ln_channels;
    const float* ln_colsum;
    float ln_eps;
    int32_t reserved0;
    /* Optional output row strides in elements for the D1 / D2 / D3 steps of the output domain (0, 0, 0 = dense rows:
     * ld_out, ld_out*D1, ld_out*D1*D2).  Lets a launch scatter its rows into a larger tensor -- the four phase GEMMs of
     * the nearest-x2 upsample + 3x3 conv (openaimodel3d.py:98-106, ae_modules.py:123-127) each write every second pixel
     * of every second row.  Each stride * element size must be a multiple of 16 bytes; no residual in this mode. */
    int64_t out_s1, out_s2, out_s3;
} dcb_tap_gemm_args;

int dcb_tap_gemm(const dcb_tap_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * GroupNorm(32 groups) over channels-last fp32 activations, split in statistics + apply.
 * Replaces ATen group_norm at basics.py:76-87 (eps 1e-5), attention.py:265,331 (1e-6),
 * openaimodel3d.py:256-262 (1e-5, 5-D: an instance spans all frames), ae_modules.py:15-16 (1e-6).
 *
 * x = channel concat of x1 [G*S, C1] and optional x2 [G*S, C2] (the UNet skip concat,
 * openaimodel3d.py:596, is never materialised in fp32).  G instances of S rows each.
 * stats: [G, 32, 2] (mean, rstd).  workspace: dcb_group_norm_workspace_bytes(G) bytes.
 * counters: G zero-initialised uint32 (left zero on return).
 * apply writes bf16 y = silu?(norm(x)*gamma+beta) and/or raw = bf16(x), both [G*S, C1+C2].
 * ------------------------------------------------------------------------------------------ */
size_t dcb_group_norm_workspace_bytes(int G);
int dcb_group_norm_stats(const float* x1, const float* x2, int C1, int C2, int G, int S, float eps,
                         float* stats, void* workspace, unsigned* counters, void* stream);
int dcb_group_norm_apply(const float* x1, const float* x2, int C1, int C2, int G, int S,
                         const float* stats, const float* gamma, const float* beta, int silu_act,
                         void* y_bf16, void* raw_bf16, void* stream);
/* (mean, rstd) [G, 32, 2] from the fixed-point accumulator a dcb_tap_gemm launch filled (gn_acc): instance g owns
 * `units_per_instance` consecutive units; count = elements per (instance, group).  Inspection / tests only -- the
 * streaming apply below reads the accumulator itself. */
int dcb_group_norm_stats_from_acc(const unsigned long long* gn_acc, int units_per_instance, int G, double count, float eps,
                                  float* stats, void* stream);
/* y = silu?((x - mean) * rstd * gamma + beta) with the statistics taken from gn_acc (see dcb_tap_gemm): ONE streaming pass,
 * x fp32 or bf16 [G*S, C], y bf16; instance g owns units [g * units_per_instance, (g+1) * units_per_instance).  The
 * consumer half of the producer-fused GroupNorm.  Replaces basics.py:76-87 / openaimodel3d.py:210-236,255-266.
 * stat_rows: rows per instance the statistics in gn_acc cover (0 = S); larger than S when the instance is spread over
 * several GPUs and gn_acc holds the all-reduced sums (frame-sharded mode). */
int dcb_group_norm_stream_apply(const void* x, int x_is_bf16, int C, int G, int S, long long stat_rows,
                                const unsigned long long* gn_acc, int units_per_instance, float eps, const float* gamma,
                                const float* beta, int silu_act, void* y_bf16, void* stream);
/* Two-source variant for the GroupNorm of a channel concatenation cat([x1, x2], 1) (openaimodel3d.py:596-597: the skip
 * connections of the UNet's output blocks): x1 [G*S, C1], x2 [G*S, C2] fp32, each with the accumulator its own producer
 * GEMM filled.  Requires the concatenation's groups ((C1 + C2) / 32 channels) to be unions of whole producer groups of one
 * source (C1 % cg == 0, cg % (C_i / 32) == 0), else DCB_ERR_UNSUPPORTED.  raw_bf16 (may be NULL): bf16 copy of the
 * concatenation (operand of the ResBlock's 1x1 skip conv, openaimodel3d.py:186-189). */
int dcb_group_norm_stream_apply2(const float* x1, const float* x2, int C1, int C2, int G, int S,
                                 const unsigned long long* gn_acc1, int units_per_instance1,
                                 const unsigned long long* gn_acc2, int units_per_instance2, float eps,
                                 const float* gamma, const float* beta, int silu_act, void* y_bf16, void* raw_bf16,
                                 void* stream);
/* Same statistics / apply for a bf16 source (block-internal activations the producer GEMM wrote as bf16):
 * single source, C % 8 == 0. */
int dcb_group_norm_stats_bf16(const void* x, int C, int G, int S, float eps, float* stats, void* workspace,
                              unsigned* counters, void* stream);
int dcb_group_norm_apply_bf16(const void* x, int C, int G, int S, const float* stats, const float* gamma,
                              const float* beta, int silu_act, void* y_bf16, void* stream);
/* Sharded instances (frames or pixels of one video spread over GPUs): raw per-(instance, group)
 * (sum, sum of squares) in double of the local shard -> all-reduce by the caller -> (mean, rstd) with the global
 * element count -> dcb_group_norm_apply.  sums: [G, 32, 2] double. */
int dcb_group_norm_sums(const float* x1, const float* x2, int C1, int C2, int G, int S, double* sums,
                        void* workspace, unsigned* counters, void* stream);
int dcb_group_norm_sums_bf16(const void* x, int C, int G, int S, double* sums, void* workspace,
                             unsigned* counters, void* stream);
int dcb_group_norm_stats_from_sums(const double* sums, int G, double count, float eps, float* stats, void* stream);
/* fp32-output apply for the fp32 parity mode on sharded instances (single fp32 source). */
int dcb_group_norm_apply_f32out(const float* x, int C, int G, int S, const float* stats, const float* gamma,
                                const float* beta, int silu_act, float* y, void* stream);
/* Statistics + apply in one cooperative launch (every block resident; blocks of an instance meet at a counter):
 * same results as stats + apply, one HBM pass less for tensors that stay in L2.  counters: 2 * G unsigned, zero
 * before the first launch (self-cleaning); workspace as for dcb_group_norm_stats.  C1, C2 multiples of 8. */
int dcb_group_norm_fused(const float* x1, const float* x2, int C1, int C2, int G, int S, float eps,
                         const float* gamma, const float* beta, int silu_act, void* y_bf16, void* raw_bf16,
                         void* workspace, unsigned* counters, void* stream);
int dcb_group_norm_fused_bf16(const void* x, int C, int G, int S, float eps, const float* gamma,
                              const float* beta, int silu_act, void* y_bf16, void* workspace,
                              unsigned* counters, void* stream);
/* LayerNorm over the last dim (attention.py:225-227), fp32 in, bf16 out. C % 4 == 0, C <= 2560. */
int dcb_layer_norm(const float* x, int rows, int C, const float* gamma, const float* beta, float eps,
                   void* y_bf16, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention (head_dim 64).  q/k/v/o are bf16 column slices of row-major buffers: head h lives at
 * columns [h*64, h*64+64) and rows advance by ld* elements.
 *
 * dcb_flash_attention: batch x heads independent problems, nq queries each; up to two key/value
 * segments with independent softmaxes:  o = sum_s seg_scale[s] * softmax(q k_s^T * scale) v_s.
 * Segment s of batch b reads kv batch b / kv_batch_div[s] (text keys are shared by the 16 frames of
 * a video).  Replaces attention.py:101-125 (self) and :89-94,129-142 (text + image cross-attention).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const void* q; int64_t ldq;
    const void* k[2]; const void* v[2]; int64_t ldk[2]; int64_t ldv[2];
    int32_t nk[2]; int32_t kv_batch_div[2]; float seg_scale[2];
    int32_t nseg, batch, heads, nq, head_dim;
    void* o; int64_t ldo;
    float softmax_scale;
} dcb_attention_args;
int dcb_flash_attention(const dcb_attention_args* args, void* stream);

/* Self-attention over L <= 16 frames for every (batch, location, head); token row of frame f is
 * ((b*L + f)*HW + pix).  Replaces attention.py:387-390 on '(b hw) t c' tensors (no transposes). */
int dcb_temporal_attention(const void* q, const void* k, const void* v, void* o, long long ldq,
                           long long ldk, long long ldv, long long ldo, int B, int L, int HW, int heads,
                           int head_dim, float softmax_scale, void* stream);
/* Single-head attention with a wide head (d = C <= 512, C % 64 == 0) for every frame in one launch: o = softmax(q k^T *
 * scale) v over the `tokens` rows of each frame; q/k/v/o are bf16 [frames*tokens, >= C] with row pitches ld*.  Flash
 * style on tcgen05 / TMEM (no N x N scores in HBM).  Replaces AttnBlock.forward's two bmm + softmax (ae_modules.py:61-74). */
int dcb_vae_attention(const void* q, const void* k, const void* v, void* o, long long ldq, long long ldk, long long ldv,
                      long long ldo, int frames, int tokens, int C, float softmax_scale, void* stream);
/* y = softmax(scale * x) row-wise, fp32 in, bf16 out (materialised-score variant of ae_modules.py:66-68; kept for tests). */
int dcb_softmax_rows(const float* x, long long rows, int n, long long ldx, float scale, void* y_bf16,
                     long long ldy, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout / resampling helpers (replace einops rearranges and F.interpolate, openaimodel3d.py:566,
 * 602,103; ae_modules.py:124) and the embedding MLPs (openaimodel3d.py:168-174,370-380,550-577).
 * ------------------------------------------------------------------------------------------ */
int dcb_nchw_to_rows(const float* src, int B, int C, int T, int HW, int Cpad, void* dst_bf16, void* stream);
int dcb_rows_to_nchw(const float* src, int ld, int B, int C, int T, int HW, float* dst, void* stream);
int dcb_upsample2x(const float* src, int nf, int H, int W, int C, void* dst_bf16, void* stream);
int dcb_space_to_depth(const float* src, int nf, int H, int W, int C, void* dst_bf16, void* stream);
int dcb_cast_bf16(const float* src, long long n, void* dst_bf16, void* stream);
/* SDS guidance gradient (guidance_pipeline.py:326-424), the second consumer of apply_model: from the cond / uncond
 * predictions at noisy = sqrt(a_t) latents + sqrt(1 - a_t) noise (dcb_q_sample):
 *   pred = e_u + s (e_c - e_u), optionally guidance-rescaled (:326-345; guidance_rescale = 0.7 for the 512 / 1024 models);
 *   x0 = (noisy - sqrt_1ma[b] pred) / sqrt_a[b] (:392);  diff = latents - x0;
 *   grad = w_t[b] diff (weight_type 0, "t": w = 1 - a_t) | diff / clamp(mean |diff|, 1e-4) (1, "ada") | diff (2, "uniform");
 *   nan_to_num (:410); block_sumsq[k] = partial sums of (latents - (latents - grad))^2 over 256-element blocks, from which
 *   loss = 0.5 * sum / (B * per_sample) / B (:417-419).
 * fp32 contiguous [B, per_sample] tensors, per_sample % 256 == 0; e_uncond NULL = no guidance.  Workspaces: stds_ws [2B],
 * mean_abs_ws [B], diff_ws [B * per_sample] floats; block_sumsq [B * per_sample / 256] doubles. */
int dcb_sds_grad(const float* latents, const float* noisy, const float* e_cond, const float* e_uncond, int B,
                 long long per_sample, const float* sqrt_a, const float* sqrt_1ma, const float* w_t, float cfg_scale,
                 float guidance_rescale, int weight_type, float* stds_ws, float* mean_abs_ws, float* diff_ws, float* grad,
                 double* block_sumsq, void* stream);
/* Batched block transpose src [Z, A, B, inner_bytes] -> dst [Z, B, A, inner_bytes] (16-byte multiples): the frame <-> pixel
 * re-ordering on either side of the all-to-all of the frame-sharded mode (parallel.py).  With
 * residual_f32 != NULL the source holds bf16 values and out_f32[dst order] = residual_f32[dst order] + src (dst unused):
 * the residual add of a temporal block (openaimodel3d.py:279, attention.py:412) where the fp32 stream lives. */
int dcb_block_transpose(const void* src, void* dst, int Z, int A, int B, long long inner_bytes,
                        const float* residual_f32, float* out_f32, void* stream);
/* y = gelu(x), exact erf (nn.GELU of the Resampler feed-forward, resampler.py:27-34); contiguous, n % 8 == 0. */
int dcb_gelu_bf16(const float* src, long long n, void* dst_bf16, void* stream);
/* Sinusoidal embedding out [B, dim] = [cos(t * f_i) | sin(t * f_i)], f_i = 10000^(-i / (dim / 2))
 * (utils_diffusion.py:8-27 timestep_embedding, called at openaimodel3d.py:550,574). */
int dcb_timestep_embedding(const long long* t, int B, int dim, float* out, void* stream);
/* out[m,n] (+)= silu_out?( sum_k silu_in?(a[m,k]) * w[n,k] + bias[n] ); fp32, tiny M.  The Linear / SiLU chains on the
 * embedding vector: time_embed and fps_embedding (openaimodel3d.py:370-381,551,575) and the emb_layers of ALL ResBlocks
 * (openaimodel3d.py:168-174,219) as one launch over their concatenated weights. */
int dcb_small_linear(const float* a, int lda, const float* w, const float* bias, int M, int N, int K,
                     int silu_in, int silu_out, int accumulate, float* out, int ldo, void* stream);

/* Output stage (scripts/evaluation/inference.py:153-159): decoded frames [B, C, T, H, W] fp32 ->
 * uint8 [B, T, H, W, C] = trunc((clamp(v, -1, 1) + 1) / 2 * 255), byte-exact with the reference's torch ops. */
int dcb_video_to_uint8(const float* src, int B, int C, int T, long long HW, unsigned char* dst, void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32 parity mode (north_star: per-step eps within rel-L2 <= 1e-3 of the fp32 reference; csrc/fp32_mode.cu).
 * Contractions stay on the tensor cores: an fp32 operand becomes three bf16 columns blocks [hi | lo | hi]
 * (dcb_split_bf16x3), the weights [hi | hi | lo], and dcb_tap_gemm runs over the 3x longer K.  Everything else
 * stays fp32: *_f32out variants of the operand producers, exact-erf GEGLU and CUDA-core attention.
 * dcb_attention_f32: same argument block as dcb_flash_attention with float q/k/v/o; inner > 1 selects
 * `inner` interleaved sequences (token stride = inner rows: temporal attention over '(b hw) t c').
 * ------------------------------------------------------------------------------------------ */
int dcb_split_bf16x3(const float* x, long long rows, int C, long long ldx, void* out_bf16, void* stream);
int dcb_geglu_f32(const float* x, long long rows, int inner, long long ldx, float* out, void* stream);
int dcb_attention_f32(const dcb_attention_args* args, int inner, void* stream);
int dcb_group_norm_fused_f32out(const float* x1, const float* x2, int C1, int C2, int G, int S, float eps,
                                const float* gamma, const float* beta, int silu_act, float* y, float* raw,
                                void* workspace, unsigned* counters, void* stream);
int dcb_layer_norm_f32out(const float* x, int rows, int C, const float* gamma, const float* beta, float eps,
                          float* y, void* stream);
int dcb_nchw_to_rows_f32out(const float* src, int B, int C, int T, int HW, int Cpad, float* dst, void* stream);
int dcb_upsample2x_f32out(const float* src, int nf, int H, int W, int C, float* dst, void* stream);
int dcb_space_to_depth_f32out(const float* src, int nf, int H, int W, int C, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused classifier-free guidance + guidance rescale + v/eps conversion + DDIM x_{t-1} update.
 * Replaces ~25 pointwise launches of DDIMSampler.p_sample_ddim (ddim.py:226-277),
 * rescale_noise_cfg (utils_diffusion.py:147-158), predict_*_from_z_and_v (ddpm3d.py:239-251).
 * Scalars are computed by the host in the reference's fp32 order (see DDIMSampler); e_uncond
 * NULL = no guidance.  stds_ws: 2*B floats (needed when guidance_rescale > 0).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t is_v;
    float cfg_scale, guidance_rescale, one_minus_guidance_rescale;
    float sqrt_acp_t, sqrt_1m_acp_t;   /* sqrt_alphas_cumprod[t], sqrt_one_minus_alphas_cumprod[t] (v-param) */
    float sqrt_a_t, sqrt_1m_a_t;       /* sqrt(ddim_alphas[index]), ddim_sqrt_one_minus_alphas[index] (eps-param) */
    float sqrt_a_prev, dir_coef, sigma_t, temperature, x0_rescale;
} dcb_ddim_coeffs;
int dcb_ddim_update(const float* x, const float* e_cond, const float* e_uncond, const float* noise,
                    int B, long long per_sample, const dcb_ddim_coeffs* c, float* stds_ws, float* x_prev,
                    float* pred_x0, void* stream);
/* Three-branch guidance of the multi-condition sampler (ddim_multiplecond.py:229-236):
 * model_output = e_uncond + cfg_img * (e_uncond_img - e_uncond) + cfg_scale * (e_cond - e_uncond_img); the rest as above. */
int dcb_ddim_update_multicond(const float* x, const float* e_cond, const float* e_uncond, const float* e_uncond_img,
                              float cfg_img, const float* noise, int B, long long per_sample, const dcb_ddim_coeffs* c,
                              float* stds_ws, float* x_prev, float* pred_x0, void* stream);

/* Forward diffusion of clean latents: out[b] = a[b] * x0[b] + c[b] * noise[b] (two rounded products, one rounded sum:
 * bit-exact with torch's eager expression).  Replaces LatentDiffusion.q_sample (ddpm3d.py:305-308; the sampler's
 * mask / x0 blend path, ddim.py:174-180) and DDIMSampler.stochastic_encode (ddim.py:304-319).  a, c: [B] device floats. */
int dcb_q_sample(const float* x0, const float* noise, const float* a, const float* c, int B, long long per_sample,
                 float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCB200_H */
