// GroupNorm (32 groups) statistics + apply(+SiLU) and LayerNorm for channels-last fp32 activations.
// HBM-bound: 128-bit loads, warp-shuffle / register accumulation, bf16 operand written once.
//
// Replaces ATen native_group_norm / layer_norm at basics.py:76-87 (eps 1e-5, fp32), attention.py:225-227,
// 265,331 (LN eps 1e-5, GN eps 1e-6), openaimodel3d.py:256-262, ae_modules.py:15-16 (eps 1e-6).
#include <cuda_bf16.h>

#include "dcb_internal.h"

namespace dcb {

constexpr int kGroups = 32;
constexpr int kMaxChunks = 1024;
constexpr int kMaxSplitChannels = 4096;   // smem slots of the stats kernel: row_split * C <= this

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// Stats.  Instance n covers S consecutive rows of C = C1 + C2 channels (x2 = channel-concatenated
// second source or null).  grid = (nchunks, G).  Each block reduces its rows into 32 (sum, sumsq)
// pairs; the last block of an instance (threadfence + counter) folds the partials in double and
// writes (mean, rstd).  Deterministic: partials are combined in a fixed order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2, int S,
                int rows_per_chunk, float* __restrict__ partial, unsigned* __restrict__ counters,
                float* __restrict__ stats, float eps) {
    // per-(row split, channel) partial sums; every slot has exactly one writer and is folded in a
    // fixed order -> bitwise reproducible (no floating-point atomics anywhere)
    __shared__ float s_part[kMaxSplitChannels * 2];
    __shared__ bool s_last;
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = C1 + C2, cg = C / kGroups;
    const int r0 = chunk * rows_per_chunk;
    const int r1 = min(S, r0 + rows_per_chunk);
    const int Q1 = C1 >> 2, Q2 = C2 >> 2;
    const int Qb1 = (Q1 + 31) >> 5, Qb2 = (Q2 + 31) >> 5, Qb = Qb1 + Qb2;
    const int row_split = Qb >= 8 ? 1 : 8 / Qb;          // warps sharing one 128-channel block split its rows

    for (int item = warp; item < Qb * row_split; item += 8) {
        const int cb = item / row_split, split = item - cb * row_split;
        const bool second = cb >= Qb1;
        const float* x = second ? x2 : x1;
        const int Cs = second ? C2 : C1;
        const int coff = second ? C1 : 0;
        const int q = (second ? cb - Qb1 : cb) * 32 + lane;
        if (q >= (Cs >> 2)) continue;
        const float* base = x + (static_cast<size_t>(n) * S) * Cs;
        float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
        int r = r0 + split;
        // two rows in flight per iteration for memory-level parallelism
        for (; r + row_split < r1; r += 2 * row_split) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(r) * Cs) + q);
            const float4 b = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(r + row_split) * Cs) + q);
            s[0] += a.x + b.x; s[1] += a.y + b.y; s[2] += a.z + b.z; s[3] += a.w + b.w;
            ss[0] += a.x * a.x + b.x * b.x; ss[1] += a.y * a.y + b.y * b.y;
            ss[2] += a.z * a.z + b.z * b.z; ss[3] += a.w * a.w + b.w * b.w;
        }
        for (; r < r1; r += row_split) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(r) * Cs) + q);
            s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
            ss[0] += a.x * a.x; ss[1] += a.y * a.y; ss[2] += a.z * a.z; ss[3] += a.w * a.w;
        }
        float* dst = s_part + (static_cast<size_t>(split) * C + coff + q * 4) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) { dst[2 * e] = s[e]; dst[2 * e + 1] = ss[e]; }
    }
    __syncthreads();
    float* pbase = partial + (static_cast<size_t>(n) * nchunks + chunk) * (kGroups * 2);
    if (tid < kGroups) {
        float a = 0.f, b = 0.f;
        for (int sp = 0; sp < row_split; ++sp)
            for (int c = tid * cg; c < (tid + 1) * cg; ++c) {
                a += s_part[(static_cast<size_t>(sp) * C + c) * 2];
                b += s_part[(static_cast<size_t>(sp) * C + c) * 2 + 1];
            }
        pbase[tid * 2] = a;
        pbase[tid * 2 + 1] = b;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&counters[n], 1u);
        s_last = (prev == static_cast<unsigned>(nchunks - 1));
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        // 8 threads per group, each a fixed strided subset of the chunks, then a fixed-order shuffle tree
        const int g = tid >> 3, sub = tid & 7;
        double sum = 0.0, sq = 0.0;
        const volatile float* pp = partial + static_cast<size_t>(n) * nchunks * (kGroups * 2);
        for (int c = sub; c < nchunks; c += 8) {
            sum += static_cast<double>(pp[c * kGroups * 2 + g * 2]);
            sq += static_cast<double>(pp[c * kGroups * 2 + g * 2 + 1]);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
        if (sub == 0) {
            const double cnt = static_cast<double>(S) * cg;
            const double mean = sum / cnt;
            double var = sq / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            stats[(static_cast<size_t>(n) * kGroups + g) * 2] = static_cast<float>(mean);
            stats[(static_cast<size_t>(n) * kGroups + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        }
        if (tid == 0) counters[n] = 0;   // self-cleaning: ready for the next launch
    }
}

// ---------------------------------------------------------------------------------------------
// Apply: y = (x - mean) * rstd * gamma + beta, optional SiLU, bf16 out [rows, C] (the tensor-core
// operand of the consumer GEMM/conv).  Optionally also emits bf16(x) (raw copy / channel concat)
// for the 1x1 skip convolution.  grid = (chunks, G).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 p;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    p.x = *reinterpret_cast<uint32_t*>(&a); p.y = *reinterpret_cast<uint32_t*>(&b);
    p.z = *reinterpret_cast<uint32_t*>(&c); p.w = *reinterpret_cast<uint32_t*>(&d);
    return p;
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2, int S,
                int rows_per_chunk, const float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, int do_silu, __nv_bfloat16* __restrict__ y,
                __nv_bfloat16* __restrict__ raw) {
    __shared__ float s_stat[kGroups * 2];
    const int n = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const int C = C1 + C2, cg = C / kGroups;
    if (tid < kGroups * 2) s_stat[tid] = stats[static_cast<size_t>(n) * kGroups * 2 + tid];
    __syncthreads();
    const int r0 = chunk * rows_per_chunk;
    const int r1 = min(S, r0 + rows_per_chunk);
    const int O = C >> 3, O1 = C1 >> 3;
    const int items = (r1 - r0) * O;
    for (int idx = tid; idx < items; idx += 256) {
        const int rr = idx / O, o = idx - rr * O;
        const size_t row = static_cast<size_t>(n) * S + r0 + rr;
        const float* src = (o < O1) ? (x1 + row * C1 + o * 8) : (x2 + row * C2 + (o - O1) * 8);
        const float4 a = __ldg(reinterpret_cast<const float4*>(src));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const int c0 = o * 8;
        if (raw != nullptr) *reinterpret_cast<uint4*>(raw + row * C + c0) = pack8(v);
        if (y != nullptr) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
            const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int g = (c0 + e) / cg;
                float t = (v[e] - s_stat[2 * g]) * s_stat[2 * g + 1] * gm[e] + bt[e];
                v[e] = do_silu ? silu(t) : t;
            }
            *reinterpret_cast<uint4*>(y + row * C + c0) = pack8(v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim (C <= 1280*... kMaxQuads*128), one warp per row, values kept in
// registers (two-pass variance), bf16 out.
// ---------------------------------------------------------------------------------------------
template <int kMaxQuads>
__global__ void __launch_bounds__(256)
layer_norm_kernel(const float* __restrict__ x, int rows, int C, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int Q = C >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * C);
    float4 v[kMaxQuads];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxQuads; ++i) {
        const int q = lane + i * 32;
        if (q < Q) {
            v[i] = __ldg(xr + q);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxQuads; ++i) {
        const int q = lane + i * 32;
        if (q < Q) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + b * b) + (c * c + d * d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / C + eps);
    __nv_bfloat16* yr = y + static_cast<size_t>(warp) * C;
#pragma unroll
    for (int i = 0; i < kMaxQuads; ++i) {
        const int q = lane + i * 32;
        if (q < Q) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + q);
            __nv_bfloat162 lo = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
            __nv_bfloat162 hi = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
            uint2 p;
            p.x = *reinterpret_cast<uint32_t*>(&lo);
            p.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(yr + q * 4) = p;
        }
    }
}

static int pick_chunks(int G, int S) {
    // enough blocks to fill the machine (~4 per SM) without making chunks smaller than 32 rows
    int want = (8 * num_sms() + G - 1) / G;   // ~8 resident 256-thread blocks per SM keeps enough bytes in flight
    int maxc = (S + 15) / 16;
    int c = want < maxc ? want : maxc;
    if (c > kMaxChunks) c = kMaxChunks;
    if (c < 1) c = 1;
    return c;
}

}  // namespace dcb

using namespace dcb;

extern "C" size_t dcb_group_norm_workspace_bytes(int G) {
    return static_cast<size_t>(G) * kMaxChunks * kGroups * 2 * sizeof(float);
}

extern "C" int dcb_group_norm_stats(const float* x1, const float* x2, int C1, int C2, int G, int S, float eps,
                                    float* stats, void* workspace, unsigned* counters, void* stream) {
    const int C = C1 + C2;
    if (x1 == nullptr || stats == nullptr || workspace == nullptr || counters == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_stats: null argument");
    if (C % kGroups != 0 || (C1 & 3) || (C2 & 3) || G < 1 || S < 1 || (C2 > 0 && x2 == nullptr) || C > kMaxSplitChannels)
        return set_error(DCB_ERR_INVALID, "group_norm_stats: C1=%d C2=%d G=%d S=%d", C1, C2, G, S);
    const int nchunks = pick_chunks(G, S);
    const int rpc = (S + nchunks - 1) / nchunks;
    const int nch = (S + rpc - 1) / rpc;
    gn_stats_kernel<<<dim3(nch, G), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x1, x2, C1, C2, S, rpc, static_cast<float*>(workspace), counters, stats, eps);
    return check_launch("group_norm_stats");
}

extern "C" int dcb_group_norm_apply(const float* x1, const float* x2, int C1, int C2, int G, int S,
                                    const float* stats, const float* gamma, const float* beta, int silu_act,
                                    void* y_bf16, void* raw_bf16, void* stream) {
    const int C = C1 + C2;
    if (x1 == nullptr || (y_bf16 == nullptr && raw_bf16 == nullptr))
        return set_error(DCB_ERR_INVALID, "group_norm_apply: null argument");
    if (y_bf16 != nullptr && (stats == nullptr || gamma == nullptr || beta == nullptr))
        return set_error(DCB_ERR_INVALID, "group_norm_apply: stats/gamma/beta required");
    if ((y_bf16 != nullptr && C % kGroups != 0) || (C1 & 7) || (C2 & 7) || G < 1 || S < 1 || (C2 > 0 && x2 == nullptr))
        return set_error(DCB_ERR_INVALID, "group_norm_apply: C1=%d C2=%d G=%d S=%d", C1, C2, G, S);
    const int nchunks = pick_chunks(G, S);
    const int rpc = (S + nchunks - 1) / nchunks;
    const int nch = (S + rpc - 1) / rpc;
    gn_apply_kernel<<<dim3(nch, G), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x1, x2, C1, C2, S, rpc, stats, gamma, beta, silu_act, static_cast<__nv_bfloat16*>(y_bf16),
        static_cast<__nv_bfloat16*>(raw_bf16));
    return check_launch("group_norm_apply");
}

extern "C" int dcb_layer_norm(const float* x, int rows, int C, const float* gamma, const float* beta, float eps,
                              void* y_bf16, void* stream) {
    if (x == nullptr || gamma == nullptr || beta == nullptr || y_bf16 == nullptr)
        return set_error(DCB_ERR_INVALID, "layer_norm: null argument");
    if (rows < 1 || C < 4 || (C & 3) || C > 20 * 128)
        return set_error(DCB_ERR_INVALID, "layer_norm: rows=%d C=%d (C %% 4 == 0, C <= 2560)", rows, C);
    const int blocks = (rows + 7) / 8;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(y_bf16);
    if (C <= 5 * 128) layer_norm_kernel<5><<<blocks, 256, 0, s>>>(x, rows, C, gamma, beta, eps, y);
    else if (C <= 10 * 128) layer_norm_kernel<10><<<blocks, 256, 0, s>>>(x, rows, C, gamma, beta, eps, y);
    else layer_norm_kernel<20><<<blocks, 256, 0, s>>>(x, rows, C, gamma, beta, eps, y);
    return check_launch("layer_norm");
}
