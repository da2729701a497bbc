// fp32 parity mode (BASELINE.json north_star: per-step eps within rel-L2 <= 1e-3 of the fp32 reference).
//
// The tensor cores stay the contraction engine: every fp32 operand x is split into two bf16 terms
// x = hi + lo (+ O(2^-17 |x|)), and  a.w ~= a_hi w_hi + a_lo w_hi + a_hi w_lo  is ONE tap_gemm launch over a
// three times longer K:  A' = [a_hi | a_lo | a_hi],  W' = [w_hi | w_hi | w_lo]  (accumulated in fp32 in TMEM).
// Everything between the GEMMs (norm outputs, q/k/v, softmax, GEGLU, attention) is kept in fp32; attention runs on
// the CUDA cores (one thread per query row, keys tiled through shared memory, full-precision expf).
// This is a verification mode, several times slower than the bf16 mode, and is never selected implicitly.
#include <cuda_bf16.h>

#include "dcb_internal.h"

namespace dcb {

// ---------------------------------------------------------------- x -> [hi | lo | hi] bf16 rows of 3C
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ x, long long rows, int C, long long ldx, __nv_bfloat16* __restrict__ out) {
    const int Q = C >> 2;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= rows * Q) return;
    const long long r = idx / Q;
    const int q = static_cast<int>(idx - r * Q);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + q);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        hi[e] = __float2bfloat16_rn(f[e]);
        lo[e] = __float2bfloat16_rn(f[e] - __bfloat162float(hi[e]));
    }
    __nv_bfloat16* o = out + r * (3LL * C) + q * 4;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(o + C) = *reinterpret_cast<const uint2*>(lo);
    *reinterpret_cast<uint2*>(o + 2 * C) = *reinterpret_cast<const uint2*>(hi);
}

// ---------------------------------------------------------------- out = value * gelu_erf(gate), fp32 (attention.py:422)
__global__ void __launch_bounds__(256)
geglu_f32_kernel(const float* __restrict__ x, long long rows, int inner, long long ldx, float* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= rows * inner) return;
    const long long r = idx / inner;
    const int c = static_cast<int>(idx - r * inner);
    const float v = x[r * ldx + c], g = x[r * ldx + inner + c];
    out[idx] = v * (0.5f * g * (1.0f + erff(g * 0.70710678118654752440f)));
}

// ---------------------------------------------------------------- fp32 attention on the CUDA cores
struct AttnF32Params {
    const float* q; long long ldq;
    const float* k[2]; const float* v[2]; long long ldk[2], ldv[2];
    int nk[2], kv_div[2]; float seg_scale[2];
    int nseg, nq, heads, inner;          // inner > 1: `inner` interleaved sequences (temporal attention over '(b hw) t c')
    float* o; long long ldo;
    float scale;
};

constexpr int kAQ = 128, kAK = 32, kAD = 64;

// sequence `bi` of a tensor whose rows are (outer, token, inner): first row and row stride
__device__ __forceinline__ long long seq_row0(long long bi, int n, int inner) { return (bi / inner) * n * inner + bi % inner; }

__global__ void __launch_bounds__(kAQ)
attention_f32_kernel(const AttnF32Params p) {
    __shared__ float sK[kAK][kAD];
    __shared__ float sV[kAK][kAD];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * kAQ, h = blockIdx.y;
    const long long bi = blockIdx.z;
    const int qi = q0 + tid;
    const bool qok = qi < p.nq;
    const long long qrow = seq_row0(bi, p.nq, p.inner) + static_cast<long long>(qok ? qi : 0) * p.inner;
    float q[kAD], out[kAD];
    {
        const float4* src = reinterpret_cast<const float4*>(p.q + qrow * p.ldq + h * kAD);
#pragma unroll
        for (int d = 0; d < kAD / 4; ++d) {
            const float4 t = __ldg(src + d);
            q[4 * d] = t.x * p.scale; q[4 * d + 1] = t.y * p.scale; q[4 * d + 2] = t.z * p.scale; q[4 * d + 3] = t.w * p.scale;
        }
    }
#pragma unroll
    for (int d = 0; d < kAD; ++d) out[d] = 0.f;
    for (int seg = 0; seg < p.nseg; ++seg) {
        const int nk = p.nk[seg];
        const long long kvb = bi / p.kv_div[seg];
        const long long krow0 = seq_row0(kvb, nk, p.inner);
        float acc[kAD];
#pragma unroll
        for (int d = 0; d < kAD; ++d) acc[d] = 0.f;
        float m = -INFINITY, l = 0.f;
        for (int k0 = 0; k0 < nk; k0 += kAK) {
            __syncthreads();
            for (int i = tid; i < kAK * (kAD / 4); i += kAQ) {
                const int j = i / (kAD / 4), d4 = i % (kAD / 4);
                float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
                if (k0 + j < nk) {
                    const long long row = krow0 + static_cast<long long>(k0 + j) * p.inner;
                    kk = __ldg(reinterpret_cast<const float4*>(p.k[seg] + row * p.ldk[seg] + h * kAD) + d4);
                    vv = __ldg(reinterpret_cast<const float4*>(p.v[seg] + row * p.ldv[seg] + h * kAD) + d4);
                }
                *reinterpret_cast<float4*>(&sK[j][4 * d4]) = kk;
                *reinterpret_cast<float4*>(&sV[j][4 * d4]) = vv;
            }
            __syncthreads();
            const int jn = min(kAK, nk - k0);
            float s[kAK];
            float mt = m;
#pragma unroll
            for (int j = 0; j < kAK; ++j) {
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < kAD; ++d) a = fmaf(q[d], sK[j][d], a);
                s[j] = j < jn ? a : -INFINITY;
                mt = fmaxf(mt, s[j]);
            }
            const float corr = expf(m - mt);          // first tile: exp(-inf) = 0
            m = mt;
            l *= corr;
#pragma unroll
            for (int d = 0; d < kAD; ++d) acc[d] *= corr;
#pragma unroll
            for (int j = 0; j < kAK; ++j) {
                const float pj = expf(s[j] - m);      // masked keys: exp(-inf) = 0
                l += pj;
#pragma unroll
                for (int d = 0; d < kAD; ++d) acc[d] = fmaf(pj, sV[j][d], acc[d]);
            }
        }
        const float inv = p.seg_scale[seg] / l;
#pragma unroll
        for (int d = 0; d < kAD; ++d) out[d] = fmaf(acc[d], inv, out[d]);
    }
    if (qok) {
        float4* dst = reinterpret_cast<float4*>(p.o + qrow * p.ldo + h * kAD);
#pragma unroll
        for (int d = 0; d < kAD / 4; ++d) dst[d] = make_float4(out[4 * d], out[4 * d + 1], out[4 * d + 2], out[4 * d + 3]);
    }
}

}  // namespace dcb

using namespace dcb;

extern "C" int dcb_split_bf16x3(const float* x, long long rows, int C, long long ldx, void* out_bf16, void* stream) {
    if (x == nullptr || out_bf16 == nullptr || rows < 1 || C < 4 || (C & 3) || ldx < C || (ldx & 3) ||
        (reinterpret_cast<uintptr_t>(x) & 15))
        return set_error(DCB_ERR_INVALID, "split_bf16x3: rows=%lld C=%d ldx=%lld", rows, C, ldx);
    const long long n = rows * (C >> 2);
    split_bf16x3_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, rows, C, ldx, static_cast<__nv_bfloat16*>(out_bf16));
    return check_launch("split_bf16x3");
}

extern "C" int dcb_geglu_f32(const float* x, long long rows, int inner, long long ldx, float* out, void* stream) {
    if (x == nullptr || out == nullptr || rows < 1 || inner < 1 || ldx < 2LL * inner)
        return set_error(DCB_ERR_INVALID, "geglu_f32: bad argument");
    const long long n = rows * inner;
    geglu_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, inner, ldx, out);
    return check_launch("geglu_f32");
}

extern "C" int dcb_attention_f32(const dcb_attention_args* a, int inner, void* stream) {
    if (a == nullptr || a->q == nullptr || a->o == nullptr || a->nseg < 1 || a->nseg > 2 || inner < 1)
        return set_error(DCB_ERR_INVALID, "attention_f32: bad argument");
    if (a->head_dim != kAD) return set_error(DCB_ERR_UNSUPPORTED, "attention_f32: head_dim must be 64");
    if (a->batch < 1 || a->heads < 1 || a->nq < 1 || (a->ldq & 3) || (a->ldo & 3) || a->heads > 65535 || a->batch > 65535)
        return set_error(DCB_ERR_INVALID, "attention_f32: shape/alignment");
    AttnF32Params p;
    p.q = static_cast<const float*>(a->q); p.ldq = a->ldq;
    p.o = static_cast<float*>(a->o); p.ldo = a->ldo;
    p.nseg = a->nseg; p.nq = a->nq; p.heads = a->heads; p.inner = inner;
    for (int s = 0; s < 2; ++s) {
        const bool on = s < a->nseg;
        if (on && (a->k[s] == nullptr || a->v[s] == nullptr || a->nk[s] < 1 || a->kv_batch_div[s] < 1 ||
                   (a->ldk[s] & 3) || (a->ldv[s] & 3)))
            return set_error(DCB_ERR_INVALID, "attention_f32: segment %d", s);
        p.k[s] = on ? static_cast<const float*>(a->k[s]) : nullptr;
        p.v[s] = on ? static_cast<const float*>(a->v[s]) : nullptr;
        p.ldk[s] = a->ldk[s]; p.ldv[s] = a->ldv[s];
        p.nk[s] = on ? a->nk[s] : 0;
        p.kv_div[s] = on ? a->kv_batch_div[s] : 1;
        p.seg_scale[s] = a->seg_scale[s];
    }
    p.scale = a->softmax_scale;
    dim3 grid((a->nq + kAQ - 1) / kAQ, a->heads, a->batch);
    attention_f32_kernel<<<grid, kAQ, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("attention_f32");
}
