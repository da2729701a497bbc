// GroupNorm (32 groups) statistics + apply(+SiLU) and LayerNorm for channels-last fp32 activations.
// HBM-bound: 128-bit loads with 4 rows in flight per lane, register accumulation, bf16 operand written once.
//
// Replaces ATen native_group_norm / layer_norm at basics.py:76-87 (eps 1e-5, fp32), attention.py:225-227,
// 265,331 (LN eps 1e-5, GN eps 1e-6), openaimodel3d.py:256-262, ae_modules.py:15-16 (eps 1e-6).
#include <cstdlib>
#include <cuda_bf16.h>

#include "dcb_internal.h"
#include "ptx.cuh"

namespace dcb {

constexpr int kGroups = 32;
constexpr int kMaxChunks = 512;
constexpr int kMaxC = 4096;
constexpr int kMaxSplitC = 4096;          // row_split * C bound (dynamic smem of the stats kernel = 8 B each)

// silu(x) = x / (1 + 2^(-x log2 e)) with the raw MUFU approximations (ex2.approx / rcp.approx, ~1 ulp each and no
// range-fixup code: the apply kernel is close to issue-bound, not only HBM-bound)
__device__ __forceinline__ float silu_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}

struct GnPlan { int row_split, rows_per_chunk, nchunks; };

// 4 / 8 consecutive channels of a row, fp32 or bf16 source
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}

// 16 bytes -> 4 fp32 or 8 bf16 channels as floats (tag pointer selects the source type)
__device__ __forceinline__ void unpack16(const uint4& u, float (&v)[4], const float*) {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
}
__device__ __forceinline__ void unpack16(const uint4& u, float (&v)[8], const __nv_bfloat16*) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}

static int gcd_i(int a, int b) { return b == 0 ? a : gcd_i(b, a % b); }

// One work item = (128-channel block, row split); a block's 8 warps share Qb*row_split items evenly.
static GnPlan plan_stats(int C1, int C2, int S, int vec = 4) {
    const int Qb = (C1 / vec + 31) / 32 + (C2 / vec + 31) / 32;
    int s = 8 / gcd_i(Qb, 8);
    const int cap = kMaxSplitC / (C1 + C2) > 0 ? kMaxSplitC / (C1 + C2) : 1;
    if (s > cap) s = cap;
    GnPlan p;
    p.row_split = s;
    p.rows_per_chunk = 8 * s;                                   // 8 rows per item, all 8 loads in flight
    p.nchunks = (S + p.rows_per_chunk - 1) / p.rows_per_chunk;
    if (p.nchunks > kMaxChunks) {
        p.rows_per_chunk = (S + kMaxChunks - 1) / kMaxChunks;
        p.rows_per_chunk = (p.rows_per_chunk + s - 1) / s * s;
        p.nchunks = (S + p.rows_per_chunk - 1) / p.rows_per_chunk;
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// Stats.  Instance n covers S consecutive rows of C = C1 + C2 channels (x2 = channel-concatenated
// second source or null).  grid = (nchunks, G).  Per-(row split, channel) partial sums land in smem
// slots with exactly one writer each and are folded in a fixed order; the last block of an instance
// (threadfence + counter) folds the per-chunk partials in double and writes (mean, rstd).
// Bitwise reproducible: no floating-point atomics anywhere.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gn_stats_kernel(const T* __restrict__ x1, const T* __restrict__ x2, int C1, int C2, int S,
                int rows_per_chunk, int row_split, float* __restrict__ partial, unsigned* __restrict__ counters,
                float* __restrict__ stats, float eps, double* __restrict__ sums) {
    extern __shared__ float s_part[];       // [row_split][C][2]
    __shared__ bool s_last;
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = C1 + C2, cg = C / kGroups;
    const int r0 = chunk * rows_per_chunk;
    const int r1 = min(S, r0 + rows_per_chunk);
    const int Qb1 = ((C1 >> 2) + 31) >> 5, Qb2 = ((C2 >> 2) + 31) >> 5, Qb = Qb1 + Qb2;

    for (int item = warp; item < Qb * row_split; item += 8) {
        const int cb = item / row_split, split = item - cb * row_split;
        const bool second = cb >= Qb1;
        const T* x = second ? x2 : x1;
        const int Cs = second ? C2 : C1;
        const int coff = second ? C1 : 0;
        const int q = (second ? cb - Qb1 : cb) * 32 + lane;
        if (q >= (Cs >> 2)) continue;
        const T* base = x + (static_cast<size_t>(n) * S) * Cs + q * 4;
        const size_t rs = static_cast<size_t>(Cs);            // row stride in elements
        float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = r0 + split; r < r1; r += 8 * row_split) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = r + u * row_split;
                v[u] = rr < r1 ? ld4(base + rr * rs) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
                ss[0] += v[u].x * v[u].x; ss[1] += v[u].y * v[u].y; ss[2] += v[u].z * v[u].z; ss[3] += v[u].w * v[u].w;
            }
        }
        float* dst = s_part + (static_cast<size_t>(split) * C + coff + q * 4) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) { dst[2 * e] = s[e]; dst[2 * e + 1] = ss[e]; }
    }
    __syncthreads();
    float* pbase = partial + (static_cast<size_t>(n) * nchunks + chunk) * (kGroups * 2);
    {
        // 8 threads per group: fixed strided subset of the (split, channel) slots, then a fixed shuffle tree
        const int g = tid >> 3, sub = tid & 7;
        float a = 0.f, b = 0.f;
        const int slots = row_split * cg;
        for (int i = sub; i < slots; i += 8) {
            const int sp = i / cg, c = g * cg + (i - sp * cg);
            a += s_part[(static_cast<size_t>(sp) * C + c) * 2];
            b += s_part[(static_cast<size_t>(sp) * C + c) * 2 + 1];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (sub == 0) {
            pbase[g * 2] = a;
            pbase[g * 2 + 1] = b;
            __threadfence();
        }
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&counters[n], 1u);
        s_last = (prev == static_cast<unsigned>(nchunks - 1));
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        // coalesced fold of the [nchunks][64] partials: thread (crow = tid/64, col = tid%64) takes chunks crow, crow+4,
        // ... with 8 independent loads in flight; the 4 row partials are then added in a fixed order (deterministic).
        __shared__ double s_fold[4][kGroups * 2];
        const int col = tid & 63, crow = tid >> 6;
        const float* pp = partial + static_cast<size_t>(n) * nchunks * (kGroups * 2) + col;
        double acc = 0.0;
        for (int c = crow; c < nchunks; c += 32) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (c + 4 * u < nchunks) ? __ldcg(pp + static_cast<size_t>(c + 4 * u) * (kGroups * 2)) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += static_cast<double>(v[u]);
        }
        s_fold[crow][col] = acc;
        __syncthreads();
        if (tid < kGroups) {
            const int g = tid;
            const double sum = ((s_fold[0][2 * g] + s_fold[1][2 * g]) + s_fold[2][2 * g]) + s_fold[3][2 * g];
            const double sq = ((s_fold[0][2 * g + 1] + s_fold[1][2 * g + 1]) + s_fold[2][2 * g + 1]) + s_fold[3][2 * g + 1];
            const double cnt = static_cast<double>(S) * cg;
            const double mean = sum / cnt;
            double var = sq / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            if (stats != nullptr) {
                stats[(static_cast<size_t>(n) * kGroups + g) * 2] = static_cast<float>(mean);
                stats[(static_cast<size_t>(n) * kGroups + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
            }
            if (sums != nullptr) {           // raw (sum, sum of squares): partial statistics of a sharded instance
                sums[(static_cast<size_t>(n) * kGroups + g) * 2] = sum;
                sums[(static_cast<size_t>(n) * kGroups + g) * 2 + 1] = sq;
            }
        }
        if (tid == 0) counters[n] = 0;   // self-cleaning: ready for the next launch
    }
}

// ---------------------------------------------------------------------------------------------
// Apply: y = silu?(x * A[c] + B[c]) with A = rstd*gamma, B = beta - mean*rstd*gamma folded once per
// block into smem; bf16 out [rows, C] (the tensor-core operand of the consumer GEMM/conv).  Optionally
// also emits bf16(x) (raw copy / channel concat) for the 1x1 skip convolution.  grid = (chunks, G).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 p;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    p.x = *reinterpret_cast<uint32_t*>(&a); p.y = *reinterpret_cast<uint32_t*>(&b);
    p.z = *reinterpret_cast<uint32_t*>(&c); p.w = *reinterpret_cast<uint32_t*>(&d);
    return p;
}

__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* v) { *reinterpret_cast<uint4*>(p) = pack8(v); }
__device__ __forceinline__ void store8(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

template <typename T, typename TOut = __nv_bfloat16>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const T* __restrict__ x1, const T* __restrict__ x2, int C1, int C2, int S,
                int rows_per_chunk, const float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, int do_silu, TOut* __restrict__ y,
                TOut* __restrict__ raw) {
    extern __shared__ float s_ab[];          // A[C] | B[C]
    const int n = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const int C = C1 + C2, cg = C / kGroups;
    if (y != nullptr) {
        for (int c = tid; c < C; c += 256) {
            const int g = c / cg;
            const float mean = stats[(static_cast<size_t>(n) * kGroups + g) * 2];
            const float rstd = stats[(static_cast<size_t>(n) * kGroups + g) * 2 + 1];
            const float a = rstd * gamma[c];
            s_ab[c] = a;
            s_ab[C + c] = beta[c] - mean * a;
        }
        __syncthreads();
    }
    const int r0 = chunk * rows_per_chunk;
    const int r1 = min(S, r0 + rows_per_chunk);
    const int O = C >> 3, O1 = C1 >> 3;
    const int items = (r1 - r0) * O;
    int rr = tid / O, o = tid - rr * O;
    const int drr = 256 / O, dO = 256 - drr * O;          // idx += 256 without a division per item
    for (int idx = tid; idx < items; idx += 256) {
        const size_t row = static_cast<size_t>(n) * S + r0 + rr;
        const T* src = (o < O1) ? (x1 + row * C1 + o * 8) : (x2 + row * C2 + (o - O1) * 8);
        float v[8];
        ld8(src, v);
        const int c0 = o * 8;
        if (raw != nullptr) store8(raw + row * C + c0, v);
        if (y != nullptr) {
            const float4 A0 = *reinterpret_cast<const float4*>(s_ab + c0), A1 = *reinterpret_cast<const float4*>(s_ab + c0 + 4);
            const float4 B0 = *reinterpret_cast<const float4*>(s_ab + C + c0), B1 = *reinterpret_cast<const float4*>(s_ab + C + c0 + 4);
            const float am[8] = {A0.x, A0.y, A0.z, A0.w, A1.x, A1.y, A1.z, A1.w};
            const float bm[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float t = fmaf(v[e], am[e], bm[e]);
                v[e] = do_silu ? (sizeof(TOut) == 4 ? t / (1.0f + expf(-t)) : silu_fast(t)) : t;
            }
            store8(y + row * C + c0, v);
        }
        rr += drr;
        o += dO;
        if (o >= O) { o -= O; ++rr; }
    }
}

// ---------------------------------------------------------------------------------------------
// Fused statistics + apply in ONE cooperative launch: grid = (nchunks, G) with every block resident.  A block
// reduces its rows of instance n (same item scheme as gn_stats_kernel), publishes 64 partial sums, waits on the
// per-instance arrival counter (the only inter-block dependency), folds all partials of its instance in the same
// fixed order as every other block (bitwise identical statistics everywhere, no broadcast needed) and normalises
// the rows it has just read -- they come back from L1 / L2 instead of HBM.  Saves a launch, the serial last-block
// finalisation and one HBM pass for every tensor that fits in L2.  counters: [2 * G] (arrive, leave), self-cleaning.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

#ifndef DCB_GN_U
#define DCB_GN_U 4          // phase 2: rows in flight per thread
#endif
#ifndef DCB_GN_P1
#define DCB_GN_P1 8         // phase 1: rows in flight per lane
#endif
#ifndef DCB_GN_OCC
#define DCB_GN_OCC 3        // resident blocks per SM the register allocation must allow
#endif
// TOut = bf16 (tensor-core operand) or float (fp32 parity mode, fp32_mode.cu)
template <typename T, typename TOut = __nv_bfloat16>
__global__ void __launch_bounds__(256, DCB_GN_OCC)
gn_fused_kernel(const T* __restrict__ x1, const T* __restrict__ x2, int C1, int C2, int S, int rows_per_chunk,
                int row_split, int col_mode, float* __restrict__ partial, unsigned* __restrict__ counters, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, int do_silu,
                TOut* __restrict__ y, TOut* __restrict__ raw) {
    extern __shared__ float s_mem[];        // phase 1: [row_split][C][2] partial slots; phase 2: A[C] | B[C]
    __shared__ double s_fold[4][kGroups * 2];
    __shared__ float s_mr[kGroups * 2];
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x, G = gridDim.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = C1 + C2, cg = C / kGroups;
    const int r0 = chunk * rows_per_chunk;
    const int r1 = min(S, r0 + rows_per_chunk);
    constexpr int kV = 16 / sizeof(T);        // channels per 16-byte load: 4 (fp32) or 8 (bf16)
    const int Qb1 = (C1 / kV + 31) >> 5, Qb2 = (C2 / kV + 31) >> 5, Qb = Qb1 + Qb2;
    pdl_launch_dependents();
    pdl_wait();

    // ---- phase 1, column mode (C <= 2048: one pass of 8-channel columns covers C): the same thread -> column mapping
    // as phase 2 (a thread owns an 8-channel column and walks down the rows, 16 / 32-byte loads, 4 rows in flight);
    // row_split = rows the block covers per step = 256 / (C / 8), one smem slot per (row lane, channel)
    if (col_mode) {
        const int O = C >> 3, O1 = C1 >> 3, nrows = r1 - r0;
        const int rpp = row_split;
        const int my_r = tid / O, my_o = tid - my_r * O;
        if (my_r < rpp) {
            const int c0 = my_o * 8;
            const size_t row0 = static_cast<size_t>(n) * S + r0 + my_r;
            const bool first = my_o < O1;
            const T* src = first ? (x1 + row0 * C1 + c0) : (x2 + row0 * C2 + (c0 - C1));
            const size_t sstep = static_cast<size_t>(rpp) * (first ? C1 : C2);
            float s[8], ss[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] = 0.f; ss[e] = 0.f; }
            for (int r = my_r; r < nrows; r += rpp * 4) {
                float v[4][8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (r + u * rpp < nrows) ld8(src + u * sstep, v[u]);
                    else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int e = 0; e < 8; ++e) { s[e] += v[u][e]; ss[e] = fmaf(v[u][e], v[u][e], ss[e]); }
                src += 4 * sstep;
            }
            float* dst = s_mem + (static_cast<size_t>(my_r) * C + c0) * 2;
#pragma unroll
            for (int e = 0; e < 8; ++e) { dst[2 * e] = s[e]; dst[2 * e + 1] = ss[e]; }
        }
    } else
    // ---- phase 1, item mode: per-(row split, channel) sums of this block's rows; an item = (32 lanes x 16 bytes of
    // channels, row split), 8 rows in flight per lane
    for (int item = warp; item < Qb * row_split; item += 8) {
        const int cb = item / row_split, split = item - cb * row_split;
        const bool second = cb >= Qb1;
        const T* x = second ? x2 : x1;
        const int Cs = second ? C2 : C1;
        const int coff = second ? C1 : 0;
        const int q = (second ? cb - Qb1 : cb) * 32 + lane;
        if (q >= Cs / kV) continue;
        const T* base = x + (static_cast<size_t>(n) * S) * Cs + q * kV;
        const size_t rs = static_cast<size_t>(Cs);
        float s[kV], ss[kV];
#pragma unroll
        for (int e = 0; e < kV; ++e) { s[e] = 0.f; ss[e] = 0.f; }
        constexpr int kR = DCB_GN_P1;         // rows in flight per lane
        for (int r = r0 + split; r < r1; r += kR * row_split) {
            uint4 raw16[kR];
#pragma unroll
            for (int u = 0; u < kR; ++u) {
                const int rr = r + u * row_split;
                raw16[u] = rr < r1 ? __ldg(reinterpret_cast<const uint4*>(base + rr * rs)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kR; ++u) {
                float v[kV];
                unpack16(raw16[u], v, static_cast<const T*>(nullptr));
#pragma unroll
                for (int e = 0; e < kV; ++e) { s[e] += v[e]; ss[e] = fmaf(v[e], v[e], ss[e]); }
            }
        }
        float* dst = s_mem + (static_cast<size_t>(split) * C + coff + q * kV) * 2;
#pragma unroll
        for (int e = 0; e < kV; ++e) { dst[2 * e] = s[e]; dst[2 * e + 1] = ss[e]; }
    }
    __syncthreads();
    {
        float* pbase = partial + (static_cast<size_t>(n) * nchunks + chunk) * (kGroups * 2);
        const int g = tid >> 3, sub = tid & 7;
        float a = 0.f, b = 0.f;
        const int slots = row_split * cg;
        for (int i = sub; i < slots; i += 8) {
            const int sp = i / cg, c = g * cg + (i - sp * cg);
            a += s_mem[(static_cast<size_t>(sp) * C + c) * 2];
            b += s_mem[(static_cast<size_t>(sp) * C + c) * 2 + 1];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (sub == 0) {
            pbase[g * 2] = a;
            pbase[g * 2 + 1] = b;
            __threadfence();
        }
    }
    __syncthreads();
    // ---- instance barrier (all blocks are resident: cooperative launch)
    if (tid == 0) {
        atomicAdd(&counters[n], 1u);
        unsigned spins = 0;
        while (ld_acquire_u32(&counters[n]) < static_cast<unsigned>(nchunks)) {
            __nanosleep(40);
            if (++spins > (1u << 24)) __trap();          // watchdog: never hang the GPU on a protocol bug
        }
    }
    __syncthreads();
    // ---- every block folds the [nchunks][64] partials of its instance in the same fixed order
    {
        const int col = tid & 63, crow = tid >> 6;
        const float* pp = partial + static_cast<size_t>(n) * nchunks * (kGroups * 2) + col;
        double acc = 0.0;
        for (int c = crow; c < nchunks; c += 64) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (c + 4 * u < nchunks) ? __ldcg(pp + static_cast<size_t>(c + 4 * u) * (kGroups * 2)) : 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u) acc += static_cast<double>(v[u]);
        }
        s_fold[crow][col] = acc;
        __syncthreads();
        if (tid < kGroups) {
            const int g = tid;
            const double sum = ((s_fold[0][2 * g] + s_fold[1][2 * g]) + s_fold[2][2 * g]) + s_fold[3][2 * g];
            const double sq = ((s_fold[0][2 * g + 1] + s_fold[1][2 * g + 1]) + s_fold[2][2 * g + 1]) + s_fold[3][2 * g + 1];
            const double cnt = static_cast<double>(S) * cg;
            const double mean = sum / cnt;
            double var = sq / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            s_mr[2 * g] = static_cast<float>(mean);
            s_mr[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        }
        __syncthreads();
    }
    if (tid == 0) {          // leave counter: the last block out re-arms both counters for the next launch
        const unsigned prev = atomicAdd(&counters[G + n], 1u);
        if (prev == static_cast<unsigned>(nchunks - 1)) { counters[n] = 0; counters[G + n] = 0; }
    }
    // ---- phase 2: y = silu?(x * A + B) on the rows this block has just read
    for (int c = tid; c < C; c += 256) {
        const int g = c / cg;
        const float a = s_mr[2 * g + 1] * gamma[c];
        s_mem[c] = a;
        s_mem[C + c] = beta[c] - s_mr[2 * g] * a;
    }
    __syncthreads();
    // a thread owns one 8-channel column (its affine coefficients live in registers) and walks down the rows of the
    // chunk with a constant pointer stride, U rows in flight
    const int O = C >> 3, O1 = C1 >> 3, nrows = r1 - r0;
    constexpr int U = DCB_GN_U;
    for (int o0 = 0; o0 < O; o0 += 256) {
        const int ow = min(256, O - o0);               // columns of this pass
        const int rpp = 256 / ow;                      // rows covered by the block per step
        const int my_r = tid / ow, my_o = o0 + tid - my_r * ow;
        if (my_r >= rpp) continue;
        const int c0 = my_o * 8;
        float am[8], bm[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { am[e] = s_mem[c0 + e]; bm[e] = s_mem[C + c0 + e]; }
        const size_t row0 = static_cast<size_t>(n) * S + r0 + my_r;
        const bool first = my_o < O1;
        const T* src = first ? (x1 + row0 * C1 + c0) : (x2 + row0 * C2 + (c0 - C1));
        const size_t sstep = static_cast<size_t>(rpp) * (first ? C1 : C2);
        const size_t dstep = static_cast<size_t>(rpp) * C;
        TOut* dst = y + row0 * C + c0;
        TOut* rdst = raw != nullptr ? raw + row0 * C + c0 : nullptr;
        for (int r = my_r; r < nrows; r += rpp * U) {
            float v[U][8];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r + u * rpp < nrows) ld8(src + u * sstep, v[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (r + u * rpp < nrows) {
                    if (rdst != nullptr) store8(rdst + u * dstep, v[u]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float t = fmaf(v[u][e], am[e], bm[e]);
                        v[u][e] = do_silu ? (sizeof(TOut) == 4 ? t / (1.0f + expf(-t)) : silu_fast(t)) : t;
                    }
                    store8(dst + u * dstep, v[u]);
                }
            }
            src += U * sstep;
            dst += U * dstep;
            if (rdst != nullptr) rdst += U * dstep;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim (C <= 1280*... kMaxQuads*128), one warp per row, values kept in
// registers (two-pass variance), bf16 out.
// ---------------------------------------------------------------------------------------------
template <int kMaxQuads, typename TOut = __nv_bfloat16>
__global__ void __launch_bounds__(256)
layer_norm_kernel(const float* __restrict__ x, int rows, int C, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, TOut* __restrict__ y) {
    // persistent warps: row, row + W, row + 2W, ... with the next row's loads issued before the current row's
    // arithmetic (two rows in flight per warp)
    const int lane = threadIdx.x & 31;
    const int W = gridDim.x * (blockDim.x >> 5);
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_launch_dependents();
    pdl_wait();
    if (row >= rows) return;
    const int Q = C >> 2;
    const float inv_c = 1.0f / C;
    float4 v[kMaxQuads], nx[kMaxQuads];
    {
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * C);
#pragma unroll
        for (int i = 0; i < kMaxQuads; ++i)
            if (lane + i * 32 < Q) v[i] = __ldg(xr + lane + i * 32);
    }
    while (true) {
        const int next = row + W;
        if (next < rows) {
            const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(next) * C);
#pragma unroll
            for (int i = 0; i < kMaxQuads; ++i)
                if (lane + i * 32 < Q) nx[i] = __ldg(xr + lane + i * 32);
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxQuads; ++i)
            if (lane + i * 32 < Q) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * inv_c;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxQuads; ++i) {
            if (lane + i * 32 < Q) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                sq += (a * a + b * b) + (c * c + d * d);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq * inv_c + eps);
        TOut* yr = y + static_cast<size_t>(row) * C;
#pragma unroll
        for (int i = 0; i < kMaxQuads; ++i) {
            const int q = lane + i * 32;
            if (q < Q) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + q);
                const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
                const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
                if constexpr (sizeof(TOut) == 4) {
                    *reinterpret_cast<float4*>(yr + q * 4) = make_float4(o0, o1, o2, o3);
                } else {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(o2, o3);
                    uint2 p;
                    p.x = *reinterpret_cast<uint32_t*>(&lo);
                    p.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(yr + q * 4) = p;
                }
            }
        }
        if (next >= rows) break;
        row = next;
#pragma unroll
        for (int i = 0; i < kMaxQuads; ++i) v[i] = nx[i];
    }
}

static int apply_rows_per_chunk(int G, int S, int C, int* nchunks) {
    // ~64 KB of fp32 input per block, at least 8 resident blocks' worth of chunks
    int rpc = (16384 + C - 1) / C;
    if (rpc < 1) rpc = 1;
    int nch = (S + rpc - 1) / rpc;
    if (nch > 65535) { nch = 65535; rpc = (S + nch - 1) / nch; nch = (S + rpc - 1) / rpc; }
    *nchunks = nch;
    (void)G;
    return rpc;
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
    if (C % kGroups != 0 || (C1 & 3) || (C2 & 3) || G < 1 || S < 1 || (C2 > 0 && x2 == nullptr) || C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_stats: C1=%d C2=%d G=%d S=%d", C1, C2, G, S);
    const GnPlan pl = plan_stats(C1, C2, S);
    const size_t smem = static_cast<size_t>(pl.row_split) * C * 2 * sizeof(float);
    gn_stats_kernel<float><<<dim3(pl.nchunks, G), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        x1, x2, C1, C2, S, pl.rows_per_chunk, pl.row_split, static_cast<float*>(workspace), counters, stats, eps, nullptr);
    return check_launch("group_norm_stats");
}

extern "C" int dcb_group_norm_sums(const float* x1, const float* x2, int C1, int C2, int G, int S, double* sums,
                                   void* workspace, unsigned* counters, void* stream) {
    const int C = C1 + C2;
    if (x1 == nullptr || sums == nullptr || workspace == nullptr || counters == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_sums: null argument");
    if (C % kGroups != 0 || (C1 & 3) || (C2 & 3) || G < 1 || S < 1 || (C2 > 0 && x2 == nullptr) || C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_sums: C1=%d C2=%d G=%d S=%d", C1, C2, G, S);
    const GnPlan pl = plan_stats(C1, C2, S);
    const size_t smem = static_cast<size_t>(pl.row_split) * C * 2 * sizeof(float);
    gn_stats_kernel<float><<<dim3(pl.nchunks, G), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        x1, x2, C1, C2, S, pl.rows_per_chunk, pl.row_split, static_cast<float*>(workspace), counters, nullptr, 0.f, sums);
    return check_launch("group_norm_sums");
}

extern "C" int dcb_group_norm_sums_bf16(const void* x, int C, int G, int S, double* sums, void* workspace,
                                        unsigned* counters, void* stream) {
    if (x == nullptr || sums == nullptr || workspace == nullptr || counters == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_sums_bf16: null argument");
    if (C % kGroups != 0 || (C & 7) || G < 1 || S < 1 || C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_sums_bf16: C=%d G=%d S=%d", C, G, S);
    const GnPlan pl = plan_stats(C, 0, S);
    const size_t smem = static_cast<size_t>(pl.row_split) * C * 2 * sizeof(float);
    gn_stats_kernel<__nv_bfloat16><<<dim3(pl.nchunks, G), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), nullptr, C, 0, S, pl.rows_per_chunk, pl.row_split,
        static_cast<float*>(workspace), counters, nullptr, 0.f, sums);
    return check_launch("group_norm_sums_bf16");
}

// (mean, rstd) from (all-reduced) raw sums: count = elements per (instance, group) over ALL shards.
__global__ void gn_stats_from_sums_kernel(const double* __restrict__ sums, int n, double count, float eps,
                                          float* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mean = sums[2 * i] / count;
    double var = sums[2 * i + 1] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * i] = static_cast<float>(mean);
    stats[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

extern "C" int dcb_group_norm_stats_from_sums(const double* sums, int G, double count, float eps, float* stats,
                                              void* stream) {
    if (sums == nullptr || stats == nullptr || G < 1 || count <= 0.0)
        return set_error(DCB_ERR_INVALID, "group_norm_stats_from_sums: bad argument");
    const int n = G * kGroups;
    gn_stats_from_sums_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(sums, n, count, eps, stats);
    return check_launch("group_norm_stats_from_sums");
}

// (mean, rstd) of one (instance, group) from the int64 fixed-point accumulator a tap_gemm epilogue filled
// (gn_acc[unit][group][2] = round(sum * 2^20), round(sumsq * 2^20)); an instance = `upi` consecutive units.
constexpr double kGnFixInv = 1.0 / 1048576.0;
__device__ __forceinline__ void gn_stats_from_acc(const unsigned long long* __restrict__ acc, int n, int g, int upi,
                                                  double count, float eps, float& mean_out, float& rstd_out) {
    long long s = 0, q = 0;
    const unsigned long long* a = acc + (static_cast<size_t>(n) * upi * kGroups + g) * 2;
    for (int u = 0; u < upi; ++u) {
        s += static_cast<long long>(__ldcg(a + static_cast<size_t>(u) * kGroups * 2));
        q += static_cast<long long>(__ldcg(a + static_cast<size_t>(u) * kGroups * 2 + 1));
    }
    const double mean = static_cast<double>(s) * kGnFixInv / count;
    double var = static_cast<double>(q) * kGnFixInv / count - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_out = static_cast<float>(mean);
    rstd_out = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

__global__ void gn_acc_to_stats_kernel(const unsigned long long* __restrict__ acc, int upi, int n_total, double count,
                                       float eps, float* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    float m, r;
    gn_stats_from_acc(acc, i / kGroups, i % kGroups, upi, count, eps, m, r);
    stats[2 * i] = m;
    stats[2 * i + 1] = r;
}

extern "C" int dcb_group_norm_stats_from_acc(const unsigned long long* gn_acc, int units_per_instance, int G, double count,
                                             float eps, float* stats, void* stream) {
    if (gn_acc == nullptr || stats == nullptr || units_per_instance < 1 || G < 1 || count <= 0)
        return set_error(DCB_ERR_INVALID, "group_norm_stats_from_acc: bad argument");
    const int n = G * kGroups;
    gn_acc_to_stats_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(gn_acc, units_per_instance, n, count,
                                                                                        eps, stats);
    return check_launch("group_norm_stats_from_acc");
}

extern "C" int dcb_group_norm_stats_bf16(const void* x, int C, int G, int S, float eps, float* stats, void* workspace,
                                         unsigned* counters, void* stream) {
    if (x == nullptr || stats == nullptr || workspace == nullptr || counters == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_stats_bf16: null argument");
    if (C % kGroups != 0 || (C & 7) || G < 1 || S < 1 || C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_stats_bf16: C=%d G=%d S=%d", C, G, S);
    const GnPlan pl = plan_stats(C, 0, S);
    const size_t smem = static_cast<size_t>(pl.row_split) * C * 2 * sizeof(float);
    gn_stats_kernel<__nv_bfloat16><<<dim3(pl.nchunks, G), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), nullptr, C, 0, S, pl.rows_per_chunk, pl.row_split,
        static_cast<float*>(workspace), counters, stats, eps, nullptr);
    return check_launch("group_norm_stats_bf16");
}

extern "C" int dcb_group_norm_apply(const float* x1, const float* x2, int C1, int C2, int G, int S,
                                    const float* stats, const float* gamma, const float* beta, int silu_act,
                                    void* y_bf16, void* raw_bf16, void* stream) {
    const int C = C1 + C2;
    if (x1 == nullptr || (y_bf16 == nullptr && raw_bf16 == nullptr))
        return set_error(DCB_ERR_INVALID, "group_norm_apply: null argument");
    if (y_bf16 != nullptr && (stats == nullptr || gamma == nullptr || beta == nullptr))
        return set_error(DCB_ERR_INVALID, "group_norm_apply: stats/gamma/beta required");
    if ((y_bf16 != nullptr && C % kGroups != 0) || (C1 & 7) || (C2 & 7) || G < 1 || S < 1 || (C2 > 0 && x2 == nullptr) ||
        C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_apply: C1=%d C2=%d G=%d S=%d", C1, C2, G, S);
    int nch = 1;
    const int rpc = apply_rows_per_chunk(G, S, C, &nch);
    const size_t smem = static_cast<size_t>(C) * 2 * sizeof(float);
    gn_apply_kernel<float, __nv_bfloat16><<<dim3(nch, G), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        x1, x2, C1, C2, S, rpc, stats, gamma, beta, silu_act, static_cast<__nv_bfloat16*>(y_bf16),
        static_cast<__nv_bfloat16*>(raw_bf16));
    return check_launch("group_norm_apply");
}

// fp32 output (fp32 parity mode with sharded instances: sums -> all-reduce -> stats -> this)
extern "C" int dcb_group_norm_apply_f32out(const float* x, int C, int G, int S, const float* stats, const float* gamma,
                                           const float* beta, int silu_act, float* y, void* stream) {
    if (x == nullptr || y == nullptr || stats == nullptr || gamma == nullptr || beta == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_apply_f32out: null argument");
    if (C % kGroups != 0 || (C & 7) || G < 1 || S < 1 || C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_apply_f32out: C=%d G=%d S=%d", C, G, S);
    int nch = 1;
    const int rpc = apply_rows_per_chunk(G, S, C, &nch);
    const size_t smem = static_cast<size_t>(C) * 2 * sizeof(float);
    gn_apply_kernel<float, float><<<dim3(nch, G), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        x, nullptr, C, 0, S, rpc, stats, gamma, beta, silu_act, y, nullptr);
    return check_launch("group_norm_apply_f32out");
}

extern "C" int dcb_group_norm_apply_bf16(const void* x, int C, int G, int S, const float* stats, const float* gamma,
                                         const float* beta, int silu_act, void* y_bf16, void* stream) {
    if (x == nullptr || y_bf16 == nullptr || stats == nullptr || gamma == nullptr || beta == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_apply_bf16: null argument");
    if (C % kGroups != 0 || (C & 7) || G < 1 || S < 1 || C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_apply_bf16: C=%d G=%d S=%d", C, G, S);
    int nch = 1;
    const int rpc = apply_rows_per_chunk(G, S, C, &nch);
    const size_t smem = static_cast<size_t>(C) * 2 * sizeof(float);
    gn_apply_kernel<__nv_bfloat16, __nv_bfloat16><<<dim3(nch, G), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), nullptr, C, 0, S, rpc, stats, gamma, beta, silu_act,
        static_cast<__nv_bfloat16*>(y_bf16), nullptr);
    return check_launch("group_norm_apply_bf16");
}

// ---------------------------------------------------------------------------------------------
// Streaming apply for statistics the producer GEMM's epilogue accumulated (gn_acc, see tap_gemm): y = silu?(x * A[c] +
// B[c]), ONE pass over x (4 or 2 bytes in, 2 bytes out per element): no statistics pass, no finalize launch, no grid
// barrier, no cooperative launch.  Every block derives (mean, rstd) of its instance from the 32 x 2 x units int64 sums
// (a few hundred bytes from L2).  grid = (chunks, G); blockDim = (C / 8) * rpp threads: a thread owns an 8-channel column
// and walks down the rows of its chunk with U rows in flight (U = 4 fp32 / 8 bf16: >= 128 bytes per thread).
// ---------------------------------------------------------------------------------------------
#ifndef DCB_GN_STREAM_U
#define DCB_GN_STREAM_U 4
#endif
constexpr int kStreamU = DCB_GN_STREAM_U;
template <typename T>
__global__ void __launch_bounds__(512)
gn_stream_apply_kernel(const T* __restrict__ x, int C, int S, long long stat_rows, int rows_per_chunk,
                       const unsigned long long* __restrict__ acc, int upi, float eps, const float* __restrict__ gamma, const float* __restrict__ beta, int do_silu,
                       __nv_bfloat16* __restrict__ y) {
    __shared__ float s_mr[kGroups * 2];
    const int n = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const int cg = C / kGroups;
    const int O = C >> 3;                        // 8-channel columns per row
    const int rpp = blockDim.x / O;              // rows covered by the block per step
    const int my_r = tid / O, my_o = tid - my_r * O;
    const int c0 = my_o * 8;
    // affine coefficients that do not depend on the statistics
    float gm[8], bt[8];
    {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
        gm[0] = g0.x; gm[1] = g0.y; gm[2] = g0.z; gm[3] = g0.w; gm[4] = g1.x; gm[5] = g1.y; gm[6] = g1.z; gm[7] = g1.w;
        bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
    }
    pdl_launch_dependents();
    pdl_wait();
    if (tid < kGroups) {
        float m, r;
        gn_stats_from_acc(acc, n, tid, upi, static_cast<double>(stat_rows) * cg, eps, m, r);
        s_mr[2 * tid] = m;
        s_mr[2 * tid + 1] = r;
    }
    __syncthreads();
    float am[8], bm[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int g = (c0 + e) / cg;
        am[e] = s_mr[2 * g + 1] * gm[e];
        bm[e] = bt[e] - s_mr[2 * g] * am[e];
    }
    const int r0 = chunk * rows_per_chunk;
    const int nrows = min(S, r0 + rows_per_chunk) - r0;
    constexpr int U = kStreamU;                  // rows in flight per thread
    const size_t row0 = static_cast<size_t>(n) * S + r0 + my_r;
    const T* src = x + row0 * C + c0;
    const size_t step = static_cast<size_t>(rpp) * C;
    __nv_bfloat16* dst = y + row0 * C + c0;
    if constexpr (sizeof(T) == 2) {
        // bf16 rows are 16 bytes per thread: keep twice as many rows in flight (raw 16-byte registers, unpacked at use)
        // so that a thread still has >= 128 bytes outstanding
        constexpr int UB = 2 * U;
        for (int r = my_r; r < nrows; r += rpp * UB) {
            uint4 raw[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u)
                if (r + u * rpp < nrows) raw[u] = __ldg(reinterpret_cast<const uint4*>(src + u * step));
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                if (r + u * rpp < nrows) {
                    float v[8];
                    unpack16(raw[u], v, static_cast<const __nv_bfloat16*>(nullptr));
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float t = fmaf(v[e], am[e], bm[e]);
                        v[e] = do_silu ? silu_fast(t) : t;
                    }
                    store8(dst + u * step, v);
                }
            }
            src += UB * step;
            dst += UB * step;
        }
    } else {
        for (int r = my_r; r < nrows; r += rpp * U) {
            float v[U][8];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r + u * rpp < nrows) ld8(src + u * step, v[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (r + u * rpp < nrows) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float t = fmaf(v[u][e], am[e], bm[e]);
                        v[u][e] = do_silu ? silu_fast(t) : t;
                    }
                    store8(dst + u * step, v[u]);
                }
            }
            src += U * step;
            dst += U * step;
        }
    }
}

template <typename T>
static int launch_gn_stream_apply(const T* x, int C, int G, int S, long long stat_rows, const unsigned long long* acc, int upi, float eps,
                                  const float* gamma, const float* beta, int silu_act, void* y, cudaStream_t stream) {
    const int O = C / 8;
    if (O > 512) return set_error(DCB_ERR_UNSUPPORTED, "group_norm_stream_apply: C=%d > 4096", C);
    int rpp = 288 / O;                              // ~256-320 threads per block
    if (rpp < 1) rpp = 1;
    const int threads = O * rpp;
    constexpr int U = kStreamU;
    // ONE wave of co-resident blocks with equal work (no tail): chunks per instance = resident capacity / G, but at least
    // one full batch of U rows per thread
    static int occ_cache[16][3] = {};                  // [device][threads class] resident blocks per SM
    int dev = 0;
    cudaGetDevice(&dev);
    const int cls = threads <= 256 ? 0 : (threads <= 320 ? 1 : 2);
    int occ = (dev >= 0 && dev < 16) ? occ_cache[dev][cls] : 0;
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_stream_apply_kernel<T>, cls == 0 ? 256 : (cls == 1 ? 320 : 512), 0) !=
                cudaSuccess || occ < 1)
            occ = 2;
        if (dev >= 0 && dev < 16) occ_cache[dev][cls] = occ;
    }
    const int capacity = occ * num_sms();
    int nch = capacity / G;
    if (nch < 1) nch = 1;
    int rpc = (S + nch - 1) / nch;
    if (rpc < rpp * U) rpc = rpp * U;               // (bf16: up to 2U rows in flight, the tail slots are predicated)
    rpc = (rpc + rpp - 1) / rpp * rpp;
    nch = (S + rpc - 1) / rpc;
    launch_pdl(kPdlGnStream, gn_stream_apply_kernel<T>, dim3(nch, G), dim3(threads), 0, stream, x, C, S, stat_rows, rpc, acc, upi, eps, gamma,
               beta, silu_act, static_cast<__nv_bfloat16*>(y));
    return check_launch("group_norm_stream_apply");
}

// Two-source variant for the skip-concat GroupNorms of the UNet's output blocks (openaimodel3d.py:596: cat([h, hs.pop()], 1)
// then in_layers GroupNorm): x1 [G*S, C1] and x2 [G*S, C2] are fp32 and come from two different producer GEMMs, each with
// its own accumulator over C_i / 32-channel groups.  A group of the concatenated tensor ((C1 + C2) / 32 channels) is a
// union of WHOLE producer groups of ONE source when C1 % cg == 0 and cg % (C_i / 32) == 0 (true for 1280+1280, 640+640,
// 320+320; the host checks): its sums are the sums of those producer groups.  Optionally also writes the raw bf16 copy of
// the concat (operand of the ResBlock's 1x1 skip conv).  One streaming pass, no statistics pass, no grid barrier.
__global__ void __launch_bounds__(512)
gn_stream_apply2_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2, int S, int rows_per_chunk,
                        const unsigned long long* __restrict__ acc1, int upi1, const unsigned long long* __restrict__ acc2,
                        int upi2, float eps, const float* __restrict__ gamma, const float* __restrict__ beta, int do_silu,
                        __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ raw) {
    __shared__ float s_mr[kGroups * 2];
    const int n = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const int C = C1 + C2;
    const int cg = C / kGroups;
    const int O = C >> 3;
    const int rpp = blockDim.x / O;
    const int my_r = tid / O, my_o = tid - my_r * O;
    const int c0 = my_o * 8;
    float gm[8], bt[8];
    {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
        gm[0] = g0.x; gm[1] = g0.y; gm[2] = g0.z; gm[3] = g0.w; gm[4] = g1.x; gm[5] = g1.y; gm[6] = g1.z; gm[7] = g1.w;
        bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
    }
    pdl_launch_dependents();
    pdl_wait();
    if (tid < kGroups) {
        // concat group tid = producer groups [pg0, pg0 + k) of source 1 or 2
        const int ch0 = tid * cg;
        const bool first = ch0 < C1;
        const int cgs = (first ? C1 : C2) / kGroups;            // channels per producer group of that source
        const int pg0 = (first ? ch0 : ch0 - C1) / cgs, k = cg / cgs;
        const unsigned long long* acc = first ? acc1 : acc2;
        const int upi = first ? upi1 : upi2;
        long long sx = 0, sq = 0;
        const unsigned long long* a = acc + (static_cast<size_t>(n) * upi * kGroups + pg0) * 2;
        for (int u = 0; u < upi; ++u)
            for (int j = 0; j < k; ++j) {
                sx += static_cast<long long>(__ldcg(a + (static_cast<size_t>(u) * kGroups + j) * 2));
                sq += static_cast<long long>(__ldcg(a + (static_cast<size_t>(u) * kGroups + j) * 2 + 1));
            }
        const double count = static_cast<double>(S) * cg;
        const double mean = static_cast<double>(sx) * kGnFixInv / count;
        double var = static_cast<double>(sq) * kGnFixInv / count - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mr[2 * tid] = static_cast<float>(mean);
        s_mr[2 * tid + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
    float am[8], bm[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int g = (c0 + e) / cg;
        am[e] = s_mr[2 * g + 1] * gm[e];
        bm[e] = bt[e] - s_mr[2 * g] * am[e];
    }
    const int r0 = chunk * rows_per_chunk;
    const int nrows = min(S, r0 + rows_per_chunk) - r0;
    constexpr int U = kStreamU;
    const size_t row0 = static_cast<size_t>(n) * S + r0 + my_r;
    const bool from1 = c0 < C1;
    const int Cs = from1 ? C1 : C2;
    const float* src = (from1 ? x1 + c0 : x2 + (c0 - C1)) + row0 * Cs;
    const size_t sstep = static_cast<size_t>(rpp) * Cs, dstep = static_cast<size_t>(rpp) * C;
    __nv_bfloat16* dst = y + row0 * C + c0;
    __nv_bfloat16* rdst = raw != nullptr ? raw + row0 * C + c0 : nullptr;
    for (int r = my_r; r < nrows; r += rpp * U) {
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + u * rpp < nrows) ld8(src + u * sstep, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r + u * rpp < nrows) {
                if (rdst != nullptr) store8(rdst + u * dstep, v[u]);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float t = fmaf(v[u][e], am[e], bm[e]);
                    v[u][e] = do_silu ? silu_fast(t) : t;
                }
                store8(dst + u * dstep, v[u]);
            }
        }
        src += U * sstep;
        dst += U * dstep;
        if (rdst != nullptr) rdst += U * dstep;
    }
}

template <typename T, typename TOut = __nv_bfloat16>
static int launch_gn_fused(const T* x1, const T* x2, int C1, int C2, int G, int S, float eps, const float* gamma,
                           const float* beta, int silu_act, void* y, void* raw, void* workspace, unsigned* counters,
                           cudaStream_t stream) {
    const int C = C1 + C2;
    const GnPlan pl = plan_stats(C1, C2, S, static_cast<int>(16 / sizeof(T)));
    static int per_sm = -1;                                   // resident blocks per SM at the largest smem request
    if (per_sm < 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gn_fused_kernel<T, TOut>, 256, kMaxSplitC * 2 * sizeof(float)) !=
            cudaSuccess || n < 1)
            return set_error(DCB_ERR_CUDA, "group_norm_fused: occupancy query failed");
        per_sm = n > 6 ? 6 : n;
    }
    const int capacity = per_sm * num_sms();
    if (G > capacity) return set_error(DCB_ERR_UNSUPPORTED, "group_norm_fused: G=%d exceeds %d resident blocks", G, capacity);
    // fill the resident capacity: small tensors get few-row chunks (row_split 1: an item = one 128-channel block
    // over all rows of the chunk, every load in flight at once), large ones the streaming plan of the stats kernel
    int nch = capacity / G;
    if (nch > kMaxChunks) nch = kMaxChunks;
    int rpc = (S + nch - 1) / nch;
    int split = 1;
    static int col_env = -1;                        // DCB_GN_COLS=0: item-mode phase 1 everywhere (A/B testing)
    if (col_env < 0) { const char* e = getenv("DCB_GN_COLS"); col_env = (e == nullptr) ? 1 : atoi(e); }
    const int col_mode = (col_env && C <= 2048 && rpc > 8) ? 1 : 0;
    if (col_mode) {
        split = 256 / (C >> 3);                     // rows per step of the column mapping
    } else if (rpc > 8) {
        split = pl.row_split;
        rpc = (rpc + split - 1) / split * split;
    }
    nch = (S + rpc - 1) / rpc;
    size_t smem = static_cast<size_t>(split) * C * 2 * sizeof(float);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.gridDim = dim3(nch, G);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (pdl_enabled(kPdlGroupNorm)) {
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, gn_fused_kernel<T, TOut>, x1, x2, C1, C2, S, rpc, split, col_mode,
                                       static_cast<float*>(workspace), counters, eps, gamma, beta, silu_act,
                                       static_cast<TOut*>(y), static_cast<TOut*>(raw));
    if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "group_norm_fused launch: %s", cudaGetErrorString(e));
    return DCB_OK;
}

extern "C" int dcb_group_norm_fused(const float* x1, const float* x2, int C1, int C2, int G, int S, float eps,
                                    const float* gamma, const float* beta, int silu_act, void* y_bf16, void* raw_bf16,
                                    void* workspace, unsigned* counters, void* stream) {
    const int C = C1 + C2;
    if (x1 == nullptr || y_bf16 == nullptr || gamma == nullptr || beta == nullptr || workspace == nullptr || counters == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_fused: null argument");
    if (C % kGroups != 0 || (C1 & 7) || (C2 & 7) || G < 1 || S < 1 || (C2 > 0 && x2 == nullptr) || C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_fused: C1=%d C2=%d G=%d S=%d", C1, C2, G, S);
    return launch_gn_fused<float>(x1, x2, C1, C2, G, S, eps, gamma, beta, silu_act, y_bf16, raw_bf16, workspace, counters,
                                  static_cast<cudaStream_t>(stream));
}

extern "C" int dcb_group_norm_fused_f32out(const float* x1, const float* x2, int C1, int C2, int G, int S, float eps,
                                           const float* gamma, const float* beta, int silu_act, float* y, float* raw,
                                           void* workspace, unsigned* counters, void* stream) {
    const int C = C1 + C2;
    if (x1 == nullptr || y == nullptr || gamma == nullptr || beta == nullptr || workspace == nullptr || counters == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_fused_f32out: null argument");
    if (C % kGroups != 0 || (C1 & 7) || (C2 & 7) || G < 1 || S < 1 || (C2 > 0 && x2 == nullptr) || C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_fused_f32out: C1=%d C2=%d G=%d S=%d", C1, C2, G, S);
    return launch_gn_fused<float, float>(x1, x2, C1, C2, G, S, eps, gamma, beta, silu_act, y, raw, workspace, counters,
                                         static_cast<cudaStream_t>(stream));
}

extern "C" int dcb_group_norm_fused_bf16(const void* x, int C, int G, int S, float eps, const float* gamma,
                                         const float* beta, int silu_act, void* y_bf16, void* workspace,
                                         unsigned* counters, void* stream) {
    if (x == nullptr || y_bf16 == nullptr || gamma == nullptr || beta == nullptr || workspace == nullptr || counters == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_fused_bf16: null argument");
    if (C % kGroups != 0 || (C & 7) || G < 1 || S < 1 || C > kMaxC)
        return set_error(DCB_ERR_INVALID, "group_norm_fused_bf16: C=%d G=%d S=%d", C, G, S);
    return launch_gn_fused<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), nullptr, C, 0, G, S, eps, gamma, beta,
                                          silu_act, y_bf16, nullptr, workspace, counters, static_cast<cudaStream_t>(stream));
}

extern "C" int dcb_group_norm_stream_apply(const void* x, int x_is_bf16, int C, int G, int S, long long stat_rows,
                                           const unsigned long long* gn_acc, int units_per_instance, float eps,
                                           const float* gamma, const float* beta, int silu_act, void* y_bf16, void* stream) {
    if (x == nullptr || y_bf16 == nullptr || gn_acc == nullptr || gamma == nullptr || beta == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_stream_apply: null argument");
    if (C % kGroups != 0 || (C & 7) || G < 1 || S < 1 || C > kMaxC || G > 65535 || units_per_instance < 1)
        return set_error(DCB_ERR_INVALID, "group_norm_stream_apply: C=%d G=%d S=%d", C, G, S);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (stat_rows <= 0) stat_rows = S;
    if (x_is_bf16)
        return launch_gn_stream_apply<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), C, G, S, stat_rows, gn_acc,
                                                     units_per_instance, eps, gamma, beta, silu_act, y_bf16, s);
    return launch_gn_stream_apply<float>(static_cast<const float*>(x), C, G, S, stat_rows, gn_acc, units_per_instance, eps, gamma,
                                         beta, silu_act, y_bf16, s);
}

extern "C" int dcb_group_norm_stream_apply2(const float* x1, const float* x2, int C1, int C2, int G, int S,
                                            const unsigned long long* gn_acc1, int units_per_instance1,
                                            const unsigned long long* gn_acc2, int units_per_instance2, float eps,
                                            const float* gamma, const float* beta, int silu_act, void* y_bf16, void* raw_bf16,
                                            void* stream) {
    if (x1 == nullptr || x2 == nullptr || y_bf16 == nullptr || gn_acc1 == nullptr || gn_acc2 == nullptr || gamma == nullptr ||
        beta == nullptr)
        return set_error(DCB_ERR_INVALID, "group_norm_stream_apply2: null argument");
    const int C = C1 + C2;
    if (C1 < 1 || C2 < 1 || C % kGroups != 0 || (C1 & 7) || (C2 & 7) || C1 % kGroups != 0 || C2 % kGroups != 0 || G < 1 || S < 1 ||
        C > kMaxC || G > 65535 || units_per_instance1 < 1 || units_per_instance2 < 1)
        return set_error(DCB_ERR_INVALID, "group_norm_stream_apply2: C1=%d C2=%d G=%d S=%d", C1, C2, G, S);
    const int cg = C / kGroups;
    if (C1 % cg != 0 || cg % (C1 / kGroups) != 0 || cg % (C2 / kGroups) != 0)
        return set_error(DCB_ERR_UNSUPPORTED, "group_norm_stream_apply2: the groups of the concatenation (%d channels) are not unions "
                         "of whole producer groups (%d / %d channels)", cg, C1 / kGroups, C2 / kGroups);
    const int O = C / 8;
    if (O > 512) return set_error(DCB_ERR_UNSUPPORTED, "group_norm_stream_apply2: C=%d > 4096", C);
    int rpp = 288 / O;
    if (rpp < 1) rpp = 1;
    const int threads = O * rpp;
    const int capacity = 2 * num_sms();
    int nch = capacity / G;
    if (nch < 1) nch = 1;
    int rpc = (S + nch - 1) / nch;
    if (rpc < rpp * kStreamU) rpc = rpp * kStreamU;
    rpc = (rpc + rpp - 1) / rpp * rpp;
    nch = (S + rpc - 1) / rpc;
    launch_pdl(kPdlGnStream, gn_stream_apply2_kernel, dim3(nch, G), dim3(threads), 0, static_cast<cudaStream_t>(stream), x1, x2, C1, C2,
               S, rpc, gn_acc1, units_per_instance1, gn_acc2, units_per_instance2, eps, gamma, beta, silu_act,
               static_cast<__nv_bfloat16*>(y_bf16), static_cast<__nv_bfloat16*>(raw_bf16));
    return check_launch("group_norm_stream_apply2");
}

extern "C" int dcb_layer_norm(const float* x, int rows, int C, const float* gamma, const float* beta, float eps,
                              void* y_bf16, void* stream) {
    if (x == nullptr || gamma == nullptr || beta == nullptr || y_bf16 == nullptr)
        return set_error(DCB_ERR_INVALID, "layer_norm: null argument");
    if (rows < 1 || C < 4 || (C & 3) || C > 20 * 128)
        return set_error(DCB_ERR_INVALID, "layer_norm: rows=%d C=%d (C %% 4 == 0, C <= 2560)", rows, C);
    // persistent warps: as many blocks as stay resident (register-limited per variant)
    const int kq = C <= 3 * 128 ? 3 : (C <= 5 * 128 ? 5 : (C <= 10 * 128 ? 10 : 20));
    const int per_sm = kq == 3 ? 4 : (kq == 5 ? 3 : (kq == 10 ? 2 : 1));
    int blocks = (rows + 7) / 8;
    if (blocks > num_sms() * per_sm) blocks = num_sms() * per_sm;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(y_bf16);
    if (kq == 3) launch_pdl(kPdlLayerNorm, layer_norm_kernel<3, __nv_bfloat16>, dim3(blocks), dim3(256), 0, s, x, rows, C, gamma, beta, eps, y);
    else if (kq == 5) launch_pdl(kPdlLayerNorm, layer_norm_kernel<5, __nv_bfloat16>, dim3(blocks), dim3(256), 0, s, x, rows, C, gamma, beta, eps, y);
    else if (kq == 10) launch_pdl(kPdlLayerNorm, layer_norm_kernel<10, __nv_bfloat16>, dim3(blocks), dim3(256), 0, s, x, rows, C, gamma, beta, eps, y);
    else launch_pdl(kPdlLayerNorm, layer_norm_kernel<20, __nv_bfloat16>, dim3(blocks), dim3(256), 0, s, x, rows, C, gamma, beta, eps, y);
    return check_launch("layer_norm");
}

extern "C" int dcb_layer_norm_f32out(const float* x, int rows, int C, const float* gamma, const float* beta, float eps,
                                     float* y, void* stream) {
    if (x == nullptr || y == nullptr || gamma == nullptr || beta == nullptr)
        return set_error(DCB_ERR_INVALID, "layer_norm_f32out: null argument");
    if (rows < 1 || C < 4 || (C & 3) || C > 10 * 128)
        return set_error(DCB_ERR_INVALID, "layer_norm_f32out: rows=%d C=%d (C %% 4 == 0, C <= 1280)", rows, C);
    int blocks = (rows + 7) / 8;
    if (blocks > num_sms() * 2) blocks = num_sms() * 2;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (C <= 5 * 128) layer_norm_kernel<5, float><<<blocks, 256, 0, s>>>(x, rows, C, gamma, beta, eps, y);
    else layer_norm_kernel<10, float><<<blocks, 256, 0, s>>>(x, rows, C, gamma, beta, eps, y);
    return check_launch("layer_norm_f32out");
}
