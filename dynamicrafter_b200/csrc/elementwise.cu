// Layout conversion, resampling gathers, the tiny embedding MLPs and the fused CFG + DDIM update.
// All HBM-/latency-bound: coalesced 128-bit accesses, no tensor cores.
#include <cuda_bf16.h>

#include "dcb_internal.h"

namespace dcb {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
    uint4 p;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    p.x = *reinterpret_cast<uint32_t*>(&a); p.y = *reinterpret_cast<uint32_t*>(&b);
    p.z = *reinterpret_cast<uint32_t*>(&c); p.w = *reinterpret_cast<uint32_t*>(&d);
    return p;
}

// [B, C, T, HW] fp32 (the reference's NCTHW, openaimodel3d.py:566) -> rows (b,t,p) x Cpad bf16, zero padded.
__device__ __forceinline__ void put8(__nv_bfloat16* p, const float* v) { *reinterpret_cast<uint4*>(p) = pack8_bf16(v); }
__device__ __forceinline__ void put8(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// TOut = bf16 (operand) or float (fp32 parity mode) for the three conv-input layout kernels below
template <typename TOut>
__global__ void nchw_to_rows_kernel(const float* __restrict__ src, int B, int C, int T, int HW, int Cpad,
                                    TOut* __restrict__ dst) {
    const long long row = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long rows = static_cast<long long>(B) * T * HW;
    if (row >= rows) return;
    const int p = row % HW, t = (row / HW) % T, b = row / (static_cast<long long>(HW) * T);
    for (int c0 = 0; c0 < Cpad; c0 += 8) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c0 + e;
            v[e] = c < C ? __ldg(src + ((static_cast<long long>(b) * C + c) * T + t) * HW + p) : 0.f;
        }
        put8(dst + row * Cpad + c0, v);
    }
}

// rows (b,t,p) x ld fp32 (first C columns) -> [B, C, T, HW] fp32 (openaimodel3d.py:602).
__global__ void rows_to_nchw_kernel(const float* __restrict__ src, int ld, int B, int C, int T, int HW,
                                    float* __restrict__ dst) {
    const long long row = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long rows = static_cast<long long>(B) * T * HW;
    if (row >= rows) return;
    const int p = row % HW, t = (row / HW) % T, b = row / (static_cast<long long>(HW) * T);
    for (int c = 0; c < C; ++c) dst[((static_cast<long long>(b) * C + c) * T + t) * HW + p] = src[row * ld + c];
}

// nearest x2 (openaimodel3d.py:103, ae_modules.py:124): fp32 [nf,H,W,C] -> bf16 [nf,2H,2W,C]
template <typename TOut>
__global__ void upsample2x_kernel(const float* __restrict__ src, int nf, int H, int W, int C,
                                  TOut* __restrict__ dst) {
    const int O = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(nf) * 4 * H * W * O;
    if (idx >= total) return;
    const int o = idx % O;
    const long long pix = idx / O;
    const int x = pix % (2 * W), y = (pix / (2 * W)) % (2 * H);
    const long long f = pix / (4LL * H * W);
    const float* s = src + ((f * H + (y >> 1)) * W + (x >> 1)) * C + o * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    put8(dst + pix * C + o * 8, v);
}

// 2x2 phase split for the stride-2 conv (openaimodel3d.py:68): fp32 [nf,H,W,C] -> bf16 [4,nf,Hp,Wp,C],
// plane (py*2+px)[f, i, j] = src[f, 2i+py, 2j+px] (zero where outside).
template <typename TOut>
__global__ void space_to_depth_kernel(const float* __restrict__ src, int nf, int H, int W, int C, int Hp, int Wp,
                                      TOut* __restrict__ dst) {
    const int O = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = 4LL * nf * Hp * Wp * O;
    if (idx >= total) return;
    const int o = idx % O;
    const long long pix = idx / O;
    const int j = pix % Wp, i = (pix / Wp) % Hp;
    const long long f = (pix / (static_cast<long long>(Wp) * Hp)) % nf;
    const int plane = pix / (static_cast<long long>(Wp) * Hp * nf);
    const int y = 2 * i + (plane >> 1), x = 2 * j + (plane & 1);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (y < H && x < W) {
        const float* s = src + ((f * H + y) * W + x) * C + o * 8;
        const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    put8(dst + pix * C + o * 8, v);
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, long long n8, __nv_bfloat16* __restrict__ dst) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n8) return;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    reinterpret_cast<uint4*>(dst)[i] = pack8_bf16(v);
}

// Batched block transpose: src [Z, A, B, inner] -> dst [Z, B, A, inner], moved as 16-byte vectors (inner16 per block).
// The frame <-> pixel re-ordering of the frame-sharded mode: (t_loc, dst rank, pixels) <-> (rank, t_loc,
// pixels) on either side of the all-to-all.  kAdd: src holds bf16 deltas (8 per vector) and the result is
// out_f32[dst index] = residual_f32[dst index] + delta -- the residual add of a temporal block, applied where the
// fp32 residual stream lives, so only bf16 crosses NVLink.
template <bool kAdd>
__global__ void block_transpose_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int A, int B, long long inner16,
                                       long long total16, const float* __restrict__ residual, float* __restrict__ out) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total16;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long blk = i / inner16, v = i - blk * inner16;        // dst block index (z, b, a)
        const long long zb = blk / A;
        const int a = static_cast<int>(blk - zb * A);
        const long long z = zb / B;
        const int b = static_cast<int>(zb - z * B);
        const uint4 u = __ldg(src + ((z * A + a) * B + b) * inner16 + v);
        if constexpr (kAdd) {
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(residual) + 2 * i);
            const float4 r1 = __ldg(reinterpret_cast<const float4*>(residual) + 2 * i + 1);
            float2 f[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) f[k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
            reinterpret_cast<float4*>(out)[2 * i] = make_float4(r0.x + f[0].x, r0.y + f[0].y, r0.z + f[1].x, r0.w + f[1].y);
            reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(r1.x + f[2].x, r1.y + f[2].y, r1.z + f[3].x, r1.w + f[3].y);
        } else {
            dst[i] = u;
        }
    }
}

// sinusoidal embedding, cos || sin (utils_diffusion.py:17-23), fp32.
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int B, int dim, float* __restrict__ out) {
    const int half = dim / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * half) return;
    const int b = idx / half, i = idx % half;
    const float freq = expf(__fdiv_rn(__fmul_rn(-9.210340371976184f, static_cast<float>(i)), static_cast<float>(half)));
    const float arg = __fmul_rn(static_cast<float>(t[b]), freq);
    out[b * dim + i] = cosf(arg);
    out[b * dim + half + i] = sinf(arg);
    if ((dim & 1) && i == 0) out[b * dim + dim - 1] = 0.f;
}

// out[m, n] (+)= act_out( sum_k act_in(a[m,k]) * W[n,k] + bias[n] ), M <= 8, fp32 weights; one warp per n.
// Used for time_embed / fps_embedding / all ResBlock emb_layers (openaimodel3d.py:168-174,370-380,551,575).
template <int kM>
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ a, int lda, const float* __restrict__ w, const float* __restrict__ bias,
                    int M, int N, int K, int silu_in, int silu_out, int accumulate, float* __restrict__ out, int ldo) {
    // the (activated) input rows are staged once per block in shared memory when they fit (use_smem): without it every
    // warp re-evaluates SiLU of the same M x K inputs for its own output column (the 20 160-column emb_layers launch spent
    // 64 us on that instead of the 16 us its 103 MB of weights need)
    extern __shared__ float s_a[];
    const bool use_smem = static_cast<size_t>(M) * K * sizeof(float) <= 40960;
    if (use_smem) {
        for (int i = threadIdx.x; i < M * (K >> 2); i += blockDim.x) {
            const int m = i / (K >> 2), q = i - m * (K >> 2);
            float4 av = __ldg(reinterpret_cast<const float4*>(a + static_cast<size_t>(m) * lda) + q);
            if (silu_in) { av.x = silu_f(av.x); av.y = silu_f(av.y); av.z = silu_f(av.z); av.w = silu_f(av.w); }
            reinterpret_cast<float4*>(s_a)[i] = av;
        }
        __syncthreads();
    }
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    float acc[kM];
#pragma unroll
    for (int m = 0; m < kM; ++m) acc[m] = 0.f;
    const float4* wr = reinterpret_cast<const float4*>(w + static_cast<size_t>(n) * K);
    for (int q = lane; q < (K >> 2); q += 32) {
        const float4 wv = __ldg(wr + q);
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            if (m < M) {
                float4 av;
                if (use_smem) {
                    av = reinterpret_cast<const float4*>(s_a)[m * (K >> 2) + q];
                } else {
                    av = __ldg(reinterpret_cast<const float4*>(a + static_cast<size_t>(m) * lda) + q);
                    if (silu_in) { av.x = silu_f(av.x); av.y = silu_f(av.y); av.z = silu_f(av.z); av.w = silu_f(av.w); }
                }
                acc[m] += (av.x * wv.x + av.y * wv.y) + (av.z * wv.z + av.w * wv.w);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < kM; ++m) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
    }
    if (lane == 0) {
        for (int m = 0; m < M && m < kM; ++m) {
            float v = acc[m] + (bias != nullptr ? bias[n] : 0.f);
            if (silu_out) v = silu_f(v);
            float* o = out + static_cast<size_t>(m) * ldo + n;
            *o = accumulate ? (*o + v) : v;
        }
    }
}

// ------------------------------------------------------------------------------- CFG + DDIM update
struct DdimParams {
    int use_cfg, is_v, use_rescale;
    float cfg_scale, guidance_rescale, one_minus_gr;
    float sqrt_acp_t, sqrt_1m_acp_t;        // 1000-entry tables at the DDPM timestep (v-param, ddpm3d.py:239-251)
    float sqrt_a_t, sqrt_1m_a_t;            // DDIM tables at index (eps-param, ddim.py:258)
    float sqrt_a_prev, dir_coef, sigma_t, temperature, x0_rescale;
};

// three-branch guidance of the multi-condition sampler (ddim_multiplecond.py:234), in the reference's order:
// (e_u + cfg_img * (e_ui - e_u)) + s * (e_c - e_ui)
__device__ __forceinline__ float cfg3(float c, float u, float ui, float s, float cfg_img) {
    return __fadd_rn(__fadd_rn(u, __fmul_rn(cfg_img, __fsub_rn(ui, u))), __fmul_rn(s, __fsub_rn(c, ui)));
}

// per-sample unbiased std of e_cond and of the CFG combination (utils_diffusion.py:152-153); one block per sample.
__global__ void __launch_bounds__(1024)
cfg_std_kernel(const float* __restrict__ e_c, const float* __restrict__ e_u, const float* __restrict__ e_ui,
               long long per_sample, float cfg_scale, float cfg_img, float* __restrict__ stds) {
    __shared__ double red[2][32];
    __shared__ double s_mean[2];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* pc = e_c + b * per_sample;
    const float* pu = e_u + b * per_sample;
    const float* pi = e_ui != nullptr ? e_ui + b * per_sample : nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        double a0 = 0.0, a1 = 0.0;
        const double m0 = pass ? s_mean[0] : 0.0, m1 = pass ? s_mean[1] : 0.0;
        if ((per_sample & 3) == 0 && pi == nullptr) {
            // 16-byte loads, four independent double chains per quantity: the loop is latency bound (one block per sample)
            double b0[4] = {0.0, 0.0, 0.0, 0.0}, b1[4] = {0.0, 0.0, 0.0, 0.0};
            for (long long i = 4LL * tid; i < per_sample; i += 4LL * blockDim.x) {
                const float4 c4 = *reinterpret_cast<const float4*>(pc + i), u4 = *reinterpret_cast<const float4*>(pu + i);
                const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, us[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float cfg = __fadd_rn(us[e], __fmul_rn(cfg_scale, __fsub_rn(cs[e], us[e])));
                    if (pass == 0) { b0[e] += cs[e]; b1[e] += cfg; }
                    else { b0[e] += (cs[e] - m0) * (cs[e] - m0); b1[e] += (cfg - m1) * (cfg - m1); }
                }
            }
            a0 = (b0[0] + b0[1]) + (b0[2] + b0[3]);
            a1 = (b1[0] + b1[1]) + (b1[2] + b1[3]);
        } else
        for (long long i = tid; i < per_sample; i += blockDim.x) {
            const float c = pc[i], u = pu[i];
            const float cfg = pi == nullptr ? __fadd_rn(u, __fmul_rn(cfg_scale, __fsub_rn(c, u)))
                                            : cfg3(c, u, pi[i], cfg_scale, cfg_img);
            if (pass == 0) { a0 += c; a1 += cfg; }
            else { a0 += (c - m0) * (c - m0); a1 += (cfg - m1) * (cfg - m1); }
        }
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        }
        if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; }
        __syncthreads();
        if (tid == 0) {
            double t0 = 0.0, t1 = 0.0;
            for (int w = 0; w < (blockDim.x >> 5); ++w) { t0 += red[0][w]; t1 += red[1][w]; }
            if (pass == 0) { s_mean[0] = t0 / per_sample; s_mean[1] = t1 / per_sample; }
            else {
                stds[2 * b] = static_cast<float>(sqrt(t0 / (per_sample - 1)));
                stds[2 * b + 1] = static_cast<float>(sqrt(t1 / (per_sample - 1)));
            }
        }
        __syncthreads();
    }
}

// One pass over x, e_cond, e_uncond, noise -> x_prev, pred_x0 (ddim.py:226-277).  Every arithmetic step is
// an explicitly rounded fp32 op in the reference's order (no FMA contraction), so eps/v results match
// torch bit for bit when guidance_rescale == 0.
__global__ void __launch_bounds__(256)
ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ e_c, const float* __restrict__ e_u,
                   const float* __restrict__ e_ui, float cfg_img,
                   const float* __restrict__ noise, const float* __restrict__ stds, long long per_sample,
                   long long total, DdimParams p, float* __restrict__ x_prev, float* __restrict__ pred_x0) {
    const long long i4 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
    if (i4 >= total) return;
    const int b = i4 / per_sample;
    float factor = 1.f;
    if (p.use_cfg && p.use_rescale) factor = __fdiv_rn(stds[2 * b], stds[2 * b + 1]);
    const float4 xv = *reinterpret_cast<const float4*>(x + i4);
    const float4 cv = *reinterpret_cast<const float4*>(e_c + i4);
    float4 uv = cv;
    if (p.use_cfg) uv = *reinterpret_cast<const float4*>(e_u + i4);
    const float4 nv = *reinterpret_cast<const float4*>(noise + i4);
    float4 iv = uv;
    if (p.use_cfg && e_ui != nullptr) iv = *reinterpret_cast<const float4*>(e_ui + i4);
    const float is[4] = {iv.x, iv.y, iv.z, iv.w};
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, cs[4] = {cv.x, cv.y, cv.z, cv.w};
    const float us[4] = {uv.x, uv.y, uv.z, uv.w}, ns[4] = {nv.x, nv.y, nv.z, nv.w};
    float xp[4], x0[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float mo = cs[e];
        if (p.use_cfg) {
            mo = e_ui == nullptr ? __fadd_rn(us[e], __fmul_rn(p.cfg_scale, __fsub_rn(cs[e], us[e])))
                                 : cfg3(cs[e], us[e], is[e], p.cfg_scale, cfg_img);
            if (p.use_rescale) {
                const float resc = __fmul_rn(mo, factor);
                mo = __fadd_rn(__fmul_rn(p.guidance_rescale, resc), __fmul_rn(p.one_minus_gr, mo));
            }
        }
        float e_t, px0;
        if (p.is_v) {
            e_t = __fadd_rn(__fmul_rn(p.sqrt_acp_t, mo), __fmul_rn(p.sqrt_1m_acp_t, xs[e]));
            px0 = __fsub_rn(__fmul_rn(p.sqrt_acp_t, xs[e]), __fmul_rn(p.sqrt_1m_acp_t, mo));
        } else {
            e_t = mo;
            px0 = __fdiv_rn(__fsub_rn(xs[e], __fmul_rn(p.sqrt_1m_a_t, e_t)), p.sqrt_a_t);
        }
        px0 = __fmul_rn(px0, p.x0_rescale);
        const float dir = __fmul_rn(p.dir_coef, e_t);
        const float nz = __fmul_rn(__fmul_rn(p.sigma_t, ns[e]), p.temperature);
        xp[e] = __fadd_rn(__fadd_rn(__fmul_rn(p.sqrt_a_prev, px0), dir), nz);
        x0[e] = px0;
    }
    *reinterpret_cast<float4*>(x_prev + i4) = make_float4(xp[0], xp[1], xp[2], xp[3]);
    *reinterpret_cast<float4*>(pred_x0 + i4) = make_float4(x0[0], x0[1], x0[2], x0[3]);
}

// ---------------------------------------------------------------------------------------------
// SDS guidance (guidance_pipeline.py:326-424): the second consumer of apply_model.  From the two predictions at the noisy
// latents:  pred = u + s (c - u)  [+ guidance rescale, :326-345]  ->  x0 = (noisy - sqrt(1 - a_t) pred) / sqrt(a_t) (:392)
// ->  diff = latents - x0  ->  grad = w diff with w = 1 - a_t ("t"), 1 ("uniform") or 1 / clamp(mean |diff|, 1e-4)
// ("ada") (:395-408)  ->  nan_to_num  ->  target = latents - grad, loss = 0.5 mean((latents - target)^2) / B (:417-419).
// Every step an explicitly rounded fp32 op in the reference's order.  Per-sample coefficients live on the device
// (timesteps differ per sample).  Two launches: differences (+ per-sample mean |diff| for "ada"), then the gradient and
// the per-sample sums of squares of the loss.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sds_diff_kernel(const float* __restrict__ latents, const float* __restrict__ noisy, const float* __restrict__ e_c,
                const float* __restrict__ e_u, const float* __restrict__ stds, const float* __restrict__ sqrt_a,
                const float* __restrict__ sqrt_1ma, long long per_sample, long long total, float cfg_scale, int use_cfg,
                float guidance_rescale, float* __restrict__ diff) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    const int b = static_cast<int>(i / per_sample);
    float mo = e_c[i];
    if (use_cfg) {
        const float u = e_u[i];
        mo = __fadd_rn(u, __fmul_rn(cfg_scale, __fsub_rn(mo, u)));
        if (guidance_rescale > 0.f) {
            const float resc = __fmul_rn(mo, __fdiv_rn(stds[2 * b], stds[2 * b + 1]));
            mo = __fadd_rn(__fmul_rn(guidance_rescale, resc), __fmul_rn(__fsub_rn(1.0f, guidance_rescale), mo));
        }
    }
    const float x0 = __fdiv_rn(__fsub_rn(noisy[i], __fmul_rn(sqrt_1ma[b], mo)), sqrt_a[b]);
    diff[i] = __fsub_rn(latents[i], x0);
}

// one block per sample: mean |diff| (double accumulation) -> 1 / clamp(mean, 1e-4) is applied by the caller as a divisor
__global__ void __launch_bounds__(1024)
sds_mean_abs_kernel(const float* __restrict__ diff, long long per_sample, float* __restrict__ mean_abs) {
    __shared__ double red[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* p = diff + b * per_sample;
    double a = 0.0;
    for (long long i = tid; i < per_sample; i += blockDim.x) a += fabs(static_cast<double>(p[i]));
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[warp] = a;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
        mean_abs[b] = static_cast<float>(t / per_sample);
    }
}

// grad = weight(diff), nan_to_num, loss partial: one block per (sample, slice); sumsq[b] += sum (latents - (latents - grad))^2
__global__ void __launch_bounds__(256)
sds_grad_kernel(const float* __restrict__ latents, const float* __restrict__ diff, const float* __restrict__ w_t,
                const float* __restrict__ mean_abs, int weight_type, long long per_sample, long long total,
                float* __restrict__ grad, double* __restrict__ sumsq) {
    __shared__ double red[8];
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    double sq = 0.0;
    int b = 0;
    if (i < total) {
        b = static_cast<int>(i / per_sample);
        float g = diff[i];
        if (weight_type == 0) g = __fmul_rn(w_t[b], g);
        else if (weight_type == 1) g = __fdiv_rn(g, fmaxf(mean_abs[b], 1e-4f));
        if (g != g) g = 0.f;                                               // torch.nan_to_num defaults
        else if (isinf(g)) g = g > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
        grad[i] = g;
        const float l = latents[i];
        const float d = __fsub_rn(l, __fsub_rn(l, g));
        sq = static_cast<double>(d) * static_cast<double>(d);
    }
    // a block never straddles two samples when per_sample % 256 == 0 (checked by the host); fixed-order reduction
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        sumsq[blockIdx.x] = t;                                             // per-block partial, summed by the host in order
    }
}

// y = gelu(x) (exact erf, nn.GELU default; resampler.py:32), fp32 in -> bf16 operand of the next GEMM.
__global__ void __launch_bounds__(256)
gelu_bf16_kernel(const float* __restrict__ src, long long n8, __nv_bfloat16* __restrict__ dst) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n8) return;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.5f * v[e] * (1.0f + erff(v[e] * 0.70710678118654752440f));
    reinterpret_cast<uint4*>(dst)[i] = pack8_bf16(v);
}

// Output stage (scripts/evaluation/inference.py:153-159, funcs.py:212-216): clamp to [-1, 1], (v + 1) / 2 * 255 with
// the reference's fp32 operation order, truncate to uint8, [B, C, T, H, W] -> [B, T, H, W, C].  One thread per pixel.
__global__ void __launch_bounds__(256)
video_to_uint8_kernel(const float* __restrict__ src, int B, int C, int T, long long HW, unsigned char* __restrict__ dst) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(B) * T * HW;
    if (idx >= total) return;
    const long long p = idx % HW;
    const int t = static_cast<int>((idx / HW) % T), b = static_cast<int>(idx / (HW * T));
    for (int c = 0; c < C; ++c) {
        float v = __ldg(src + ((static_cast<long long>(b) * C + c) * T + t) * HW + p);
        v = fminf(fmaxf(v, -1.0f), 1.0f);
        v = __fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), 255.0f);
        dst[idx * C + c] = static_cast<unsigned char>(static_cast<int>(v));      // trunc toward zero, like .to(torch.uint8)
    }
}

}  // namespace dcb

using namespace dcb;

static inline unsigned blocks_for(long long n, int bs) { return static_cast<unsigned>((n + bs - 1) / bs); }

extern "C" int dcb_nchw_to_rows(const float* src, int B, int C, int T, int HW, int Cpad, void* dst_bf16, void* stream) {
    if (!src || !dst_bf16 || B < 1 || C < 1 || T < 1 || HW < 1 || Cpad < C || (Cpad & 7))
        return set_error(DCB_ERR_INVALID, "nchw_to_rows: bad argument");
    const long long rows = static_cast<long long>(B) * T * HW;
    nchw_to_rows_kernel<__nv_bfloat16><<<blocks_for(rows, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        src, B, C, T, HW, Cpad, static_cast<__nv_bfloat16*>(dst_bf16));
    return check_launch("nchw_to_rows");
}

// fp32-output variants of the three conv-input layout kernels (fp32 parity mode)
extern "C" int dcb_nchw_to_rows_f32out(const float* src, int B, int C, int T, int HW, int Cpad, float* dst, void* stream) {
    if (!src || !dst || B < 1 || C < 1 || T < 1 || HW < 1 || Cpad < C || (Cpad & 7))
        return set_error(DCB_ERR_INVALID, "nchw_to_rows_f32out: bad argument");
    const long long rows = static_cast<long long>(B) * T * HW;
    nchw_to_rows_kernel<float><<<blocks_for(rows, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(src, B, C, T, HW, Cpad, dst);
    return check_launch("nchw_to_rows_f32out");
}

extern "C" int dcb_upsample2x_f32out(const float* src, int nf, int H, int W, int C, float* dst, void* stream) {
    if (!src || !dst || nf < 1 || H < 1 || W < 1 || C < 8 || (C & 7))
        return set_error(DCB_ERR_INVALID, "upsample2x_f32out: bad argument");
    const long long total = static_cast<long long>(nf) * 4 * H * W * (C >> 3);
    upsample2x_kernel<float><<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, nf, H, W, C, dst);
    return check_launch("upsample2x_f32out");
}

extern "C" int dcb_space_to_depth_f32out(const float* src, int nf, int H, int W, int C, float* dst, void* stream) {
    if (!src || !dst || nf < 1 || H < 1 || W < 1 || C < 8 || (C & 7))
        return set_error(DCB_ERR_INVALID, "space_to_depth_f32out: bad argument");
    const int Hp = (H + 1) / 2, Wp = (W + 1) / 2;
    const long long total = 4LL * nf * Hp * Wp * (C >> 3);
    space_to_depth_kernel<float><<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, nf, H, W, C, Hp, Wp, dst);
    return check_launch("space_to_depth_f32out");
}

extern "C" int dcb_rows_to_nchw(const float* src, int ld, int B, int C, int T, int HW, float* dst, void* stream) {
    if (!src || !dst || B < 1 || C < 1 || T < 1 || HW < 1 || ld < C)
        return set_error(DCB_ERR_INVALID, "rows_to_nchw: bad argument");
    const long long rows = static_cast<long long>(B) * T * HW;
    rows_to_nchw_kernel<<<blocks_for(rows, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(src, ld, B, C, T, HW, dst);
    return check_launch("rows_to_nchw");
}

extern "C" int dcb_upsample2x(const float* src, int nf, int H, int W, int C, void* dst_bf16, void* stream) {
    if (!src || !dst_bf16 || nf < 1 || H < 1 || W < 1 || C < 8 || (C & 7))
        return set_error(DCB_ERR_INVALID, "upsample2x: bad argument");
    const long long total = static_cast<long long>(nf) * 4 * H * W * (C >> 3);
    upsample2x_kernel<__nv_bfloat16><<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, nf, H, W, C, static_cast<__nv_bfloat16*>(dst_bf16));
    return check_launch("upsample2x");
}

extern "C" int dcb_space_to_depth(const float* src, int nf, int H, int W, int C, void* dst_bf16, void* stream) {
    if (!src || !dst_bf16 || nf < 1 || H < 1 || W < 1 || C < 8 || (C & 7))
        return set_error(DCB_ERR_INVALID, "space_to_depth: bad argument");
    const int Hp = (H + 1) / 2, Wp = (W + 1) / 2;
    const long long total = 4LL * nf * Hp * Wp * (C >> 3);
    space_to_depth_kernel<__nv_bfloat16><<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, nf, H, W, C, Hp, Wp, static_cast<__nv_bfloat16*>(dst_bf16));
    return check_launch("space_to_depth");
}

extern "C" int dcb_cast_bf16(const float* src, long long n, void* dst_bf16, void* stream) {
    if (!src || !dst_bf16 || n < 8 || (n & 7)) return set_error(DCB_ERR_INVALID, "cast_bf16: n must be a multiple of 8");
    cast_bf16_kernel<<<blocks_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, n / 8, static_cast<__nv_bfloat16*>(dst_bf16));
    return check_launch("cast_bf16");
}

extern "C" int dcb_sds_grad(const float* latents, const float* noisy, const float* e_cond, const float* e_uncond, int B,
                            long long per_sample, const float* sqrt_a, const float* sqrt_1ma, const float* w_t,
                            float cfg_scale, float guidance_rescale, int weight_type, float* stds_ws, float* mean_abs_ws,
                            float* diff_ws, float* grad, double* block_sumsq, void* stream) {
    if (!latents || !noisy || !e_cond || !sqrt_a || !sqrt_1ma || !w_t || !stds_ws || !mean_abs_ws || !diff_ws || !grad ||
        !block_sumsq || B < 1 || per_sample < 256 || (per_sample & 255) || weight_type < 0 || weight_type > 2)
        return set_error(DCB_ERR_INVALID, "sds_grad: bad argument (per_sample must be a multiple of 256)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int use_cfg = e_uncond != nullptr ? 1 : 0;
    if (use_cfg && guidance_rescale > 0.f)
        cfg_std_kernel<<<B, 1024, 0, s>>>(e_cond, e_uncond, nullptr, per_sample, cfg_scale, 0.f, stds_ws);
    const long long total = static_cast<long long>(B) * per_sample;
    sds_diff_kernel<<<blocks_for(total, 256), 256, 0, s>>>(latents, noisy, e_cond, e_uncond, stds_ws, sqrt_a, sqrt_1ma, per_sample,
                                                          total, cfg_scale, use_cfg, guidance_rescale, diff_ws);
    if (weight_type == 1) sds_mean_abs_kernel<<<B, 1024, 0, s>>>(diff_ws, per_sample, mean_abs_ws);
    sds_grad_kernel<<<blocks_for(total, 256), 256, 0, s>>>(latents, diff_ws, w_t, mean_abs_ws, weight_type, per_sample, total, grad,
                                                          block_sumsq);
    return check_launch("sds_grad");
}

extern "C" int dcb_block_transpose(const void* src, void* dst, int Z, int A, int B, long long inner_bytes,
                                   const float* residual_f32, float* out_f32, void* stream) {
    if (!src || Z < 1 || A < 1 || B < 1 || inner_bytes < 16 || (inner_bytes & 15) || (reinterpret_cast<uintptr_t>(src) & 15))
        return set_error(DCB_ERR_INVALID, "block_transpose: blocks must be 16-byte aligned multiples of 16 bytes");
    const long long inner16 = inner_bytes / 16, total16 = static_cast<long long>(Z) * A * B * inner16;
    const long long want = (total16 + 255) / 256;
    const unsigned blocks = static_cast<unsigned>(want < 148LL * 16 ? want : 148LL * 16);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (residual_f32 != nullptr) {
        if (!out_f32 || ((reinterpret_cast<uintptr_t>(residual_f32) | reinterpret_cast<uintptr_t>(out_f32)) & 15))
            return set_error(DCB_ERR_INVALID, "block_transpose: residual / out must be 16-byte aligned fp32");
        block_transpose_kernel<true><<<blocks, 256, 0, s>>>(static_cast<const uint4*>(src), nullptr, A, B, inner16, total16,
                                                            residual_f32, out_f32);
    } else {
        if (!dst || (reinterpret_cast<uintptr_t>(dst) & 15)) return set_error(DCB_ERR_INVALID, "block_transpose: dst");
        block_transpose_kernel<false><<<blocks, 256, 0, s>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), A, B,
                                                             inner16, total16, nullptr, nullptr);
    }
    return check_launch("block_transpose");
}

extern "C" int dcb_timestep_embedding(const long long* t, int B, int dim, float* out, void* stream) {
    if (!t || !out || B < 1 || dim < 2) return set_error(DCB_ERR_INVALID, "timestep_embedding: bad argument");
    timestep_embedding_kernel<<<blocks_for(static_cast<long long>(B) * (dim / 2), 128), 128, 0,
                                static_cast<cudaStream_t>(stream)>>>(t, B, dim, out);
    return check_launch("timestep_embedding");
}

extern "C" int dcb_small_linear(const float* a, int lda, const float* w, const float* bias, int M, int N, int K,
                                int silu_in, int silu_out, int accumulate, float* out, int ldo, void* stream) {
    if (!a || !w || !out || M < 1 || N < 1 || K < 4 || (K & 3) || (lda & 3))
        return set_error(DCB_ERR_INVALID, "small_linear: bad argument (K, lda multiples of 4)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int m0 = 0; m0 < M; m0 += 8) {
        const int m = (M - m0) < 8 ? (M - m0) : 8;
        const size_t stage = static_cast<size_t>(m) * K * sizeof(float);
        small_linear_kernel<8><<<blocks_for(static_cast<long long>(N) * 32, 256), 256, stage <= 40960 ? stage : 0, s>>>(
            a + static_cast<size_t>(m0) * lda, lda, w, bias, m, N, K, silu_in, silu_out, accumulate,
            out + static_cast<size_t>(m0) * ldo, ldo);
    }
    return check_launch("small_linear");
}

static int ddim_update_impl(const float* x, const float* e_cond, const float* e_uncond, const float* e_uncond_img,
                            float cfg_img, const float* noise, int B, long long per_sample, const dcb_ddim_coeffs* c,
                            float* stds_ws, float* x_prev, float* pred_x0, void* stream) {
    if (!x || !e_cond || !noise || !c || !x_prev || !pred_x0 || B < 1 || per_sample < 4 || (per_sample & 3))
        return set_error(DCB_ERR_INVALID, "ddim_update: bad argument (per_sample multiple of 4)");
    if (e_uncond_img != nullptr && e_uncond == nullptr)
        return set_error(DCB_ERR_INVALID, "ddim_update: the image-only branch needs the unconditional branch");
    const int use_cfg = (e_uncond != nullptr) ? 1 : 0;
    const int use_rescale = (use_cfg && c->guidance_rescale > 0.f) ? 1 : 0;
    if (use_rescale && stds_ws == nullptr) return set_error(DCB_ERR_INVALID, "ddim_update: stds workspace needed");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DdimParams p;
    p.use_cfg = use_cfg; p.is_v = c->is_v; p.use_rescale = use_rescale;
    p.cfg_scale = c->cfg_scale; p.guidance_rescale = c->guidance_rescale; p.one_minus_gr = c->one_minus_guidance_rescale;
    p.sqrt_acp_t = c->sqrt_acp_t; p.sqrt_1m_acp_t = c->sqrt_1m_acp_t;
    p.sqrt_a_t = c->sqrt_a_t; p.sqrt_1m_a_t = c->sqrt_1m_a_t;
    p.sqrt_a_prev = c->sqrt_a_prev; p.dir_coef = c->dir_coef; p.sigma_t = c->sigma_t;
    p.temperature = c->temperature; p.x0_rescale = c->x0_rescale;
    if (use_rescale)
        cfg_std_kernel<<<B, 1024, 0, s>>>(e_cond, e_uncond, e_uncond_img, per_sample, c->cfg_scale, cfg_img, stds_ws);
    const long long total = static_cast<long long>(B) * per_sample;
    ddim_update_kernel<<<blocks_for(total / 4, 256), 256, 0, s>>>(x, e_cond, e_uncond, e_uncond_img, cfg_img, noise, stds_ws,
                                                                  per_sample, total, p, x_prev, pred_x0);
    return check_launch("ddim_update");
}

extern "C" int dcb_ddim_update(const float* x, const float* e_cond, const float* e_uncond, const float* noise,
                               int B, long long per_sample, const dcb_ddim_coeffs* c, float* stds_ws, float* x_prev,
                               float* pred_x0, void* stream) {
    return ddim_update_impl(x, e_cond, e_uncond, nullptr, 0.f, noise, B, per_sample, c, stds_ws, x_prev, pred_x0, stream);
}

extern "C" int dcb_ddim_update_multicond(const float* x, const float* e_cond, const float* e_uncond,
                                         const float* e_uncond_img, float cfg_img, const float* noise, int B,
                                         long long per_sample, const dcb_ddim_coeffs* c, float* stds_ws, float* x_prev,
                                         float* pred_x0, void* stream) {
    if (e_uncond_img == nullptr) return set_error(DCB_ERR_INVALID, "ddim_update_multicond: null image-only branch");
    return ddim_update_impl(x, e_cond, e_uncond, e_uncond_img, cfg_img, noise, B, per_sample, c, stds_ws, x_prev, pred_x0,
                            stream);
}

// out[b, i] = a[b] * x0[b, i] + c[b] * noise[b, i] with two explicitly rounded products and one rounded sum (the order
// torch's eager `a * x0 + c * noise` evaluates, ddpm3d.py:305-308 q_sample / ddim.py:304-319 stochastic_encode).
namespace dcb {
__global__ void __launch_bounds__(256)
q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const float* __restrict__ a,
                const float* __restrict__ c, long long per, long long total, float* __restrict__ out) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    const long long b = i / per;
    out[i] = __fadd_rn(__fmul_rn(__ldg(a + b), x0[i]), __fmul_rn(__ldg(c + b), noise[i]));
}
}  // namespace dcb

extern "C" int dcb_q_sample(const float* x0, const float* noise, const float* a, const float* c, int B, long long per_sample,
                            float* out, void* stream) {
    if (!x0 || !noise || !a || !c || !out || B < 1 || per_sample < 1) return set_error(DCB_ERR_INVALID, "q_sample: bad argument");
    const long long total = static_cast<long long>(B) * per_sample;
    q_sample_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x0, noise, a, c, per_sample, total, out);
    return check_launch("q_sample");
}

extern "C" int dcb_video_to_uint8(const float* src, int B, int C, int T, long long HW, unsigned char* dst, void* stream) {
    if (!src || !dst || B < 1 || C < 1 || T < 1 || HW < 1) return set_error(DCB_ERR_INVALID, "video_to_uint8: bad argument");
    const long long total = static_cast<long long>(B) * T * HW;
    video_to_uint8_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, B, C, T, HW, dst);
    return check_launch("video_to_uint8");
}

extern "C" int dcb_gelu_bf16(const float* src, long long n, void* dst_bf16, void* stream) {
    if (!src || !dst_bf16 || n < 8 || (n & 7)) return set_error(DCB_ERR_INVALID, "gelu_bf16: n must be a positive multiple of 8");
    gelu_bf16_kernel<<<blocks_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, n / 8, static_cast<__nv_bfloat16*>(dst_bf16));
    return check_launch("gelu_bf16");
}
