// tcgen05 / TMEM / TMA flash attention for the spatial self-attention of the UNet (attention.py:101-125, head_dim 64).
//
// CTA = one 128-query tile of one (batch, head); 320 threads:
//   warp 0      TMA producer: Q tile once, then K_j / V_j tiles (128 keys x 64) into 2-deep rings (SWIZZLE_128B)
//   warp 1      TMEM owner + tcgen05.mma issuer:  S_j = Q K_j^T (128x128, fp32 in TMEM, double buffered) and
//               O_j = P_j V_j (128x64, written over the consumed S_j columns; V is the MN-major B operand)
//   warps 2..9  softmax: TWO threads per query row (warps w and w + 4 share a TMEM lane quarter): both find the row
//               max of S_j (tcgen05.ld, no shuffles), each exponentiates ITS 64 keys into the bf16 P_j tile in smem
//               (A operand of the PV MMA) and accumulates ITS 32 output columns of O += O_j in registers with the
//               usual (lazy) rescale; partial row sums are combined once at the end; O / l -> bf16 -> swizzled smem
//               -> TMA store.  Two threads per row double the warps the schedulers can pick from: the softmax is
//               latency bound (exp2 + TMEM round trips), not throughput bound.
// S_{j+1} is issued before P_j V_j, so the tensor core works on the next scores while the softmax warps are busy;
// ~113 KB smem and 256 TMEM columns per CTA let two CTAs share an SM for further overlap.
// The N x N score matrix never leaves the SM (the eager reference materialises 1 GB per layer at 320x512).
#include "dcb_internal.h"
#include "ptx.cuh"

namespace dcb {

constexpr int kFQ = 128;              // queries per CTA
constexpr int kFK = 128;              // keys per tile
constexpr int kFD = 64;               // head dim
constexpr int kTileBytes = 128 * 128; // 128 rows x 64 bf16
constexpr int kFThreads = 320;
constexpr int kFSmem = 7 * kTileBytes + 256;   // Q, K[2], V[2], P[2 halves] + barriers

// -DDCB_TRACE: CTA (0,0,0) logs (clock, event, tile) per role: 0 producer, 1 mma, 2 softmax warp 2, 3 softmax warp 6
#ifdef DCB_TRACE
static __device__ unsigned long long g_flash_trace[4 * 4096];
struct FlashTrace {
    unsigned n = 1;
    __device__ __forceinline__ void ev(int role, unsigned code, unsigned arg) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && n < 4096u)
            g_flash_trace[role * 4096 + n++] = (static_cast<unsigned long long>(clock64()) << 20) | (code << 12) | (arg & 0xfffu);
    }
    __device__ __forceinline__ void done(int role) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_flash_trace[role * 4096] = n;
    }
};
#define FTR_DECL FlashTrace ftr;
#define FTR(role, code, arg) ftr.ev(role, code, arg)
#define FTR_DONE(role) ftr.done(role)
#else
#define FTR_DECL
#define FTR(role, code, arg)
#define FTR_DONE(role)
#endif

struct FlashTcParams {
    int nk, ntiles;
    float scale_log2;
};

// MN-major (N contiguous), 128-byte swizzled B operand: 8-row (K) groups 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f via the 1.5 * 2^23 trick, degree-3 minimax
// polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P), exponent added
// as an integer.  Measured on B200: the XU pipe is the busiest pipe of this kernel (54 %) but not its limiter --
// offloading 25 % of the exponentials is neutral, 50 % costs 6 % (7 more issue slots per element), so the default
// mask is 0; kept for head dims / clocks where the balance differs.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.f);
    const float t = x + 12582912.f;
    const float n = t - 12582912.f;
    const float f = x - n;
    float p = fmaf(0.05517196f, f, 0.24261114f);
    p = fmaf(p, f, 0.69326091f);
    p = fmaf(p, f, 0.99992806f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

#ifndef DCB_FLASH_POLY_MASK
#define DCB_FLASH_POLY_MASK 0x0          // which of the four 8-key groups of a 32-key chunk use ex2_poly (bit g4)
#endif

__global__ void __launch_bounds__(kFThreads, 2)
flash_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FlashTcParams p) {
    extern __shared__ __align__(1024) uint8_t fsm[];
    uint8_t* sQ = fsm;
    uint8_t* sK = fsm + kTileBytes;            // [2]
    uint8_t* sV = fsm + 3 * kTileBytes;        // [2]
    uint8_t* sP = fsm + 5 * kTileBytes;        // [2 key halves of 64] ; half 0 doubles as the output staging tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(fsm + 7 * kTileBytes);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;               // [2]
    uint64_t* v_full = bars + 3;               // [2]
    uint64_t* kv_empty = bars + 5;             // [2]
    uint64_t* s_full = bars + 7;               // [2]
    uint64_t* p_full = bars + 9;               // [2]
    uint64_t* o_full = bars + 11;              // [2]
    uint64_t* s_free = bars + 13;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kFQ, head = blockIdx.y, batch = blockIdx.z;
    const int T = p.ntiles;
    FTR_DECL

    if ((smem_u32(fsm) & 1023) != 0) __trap();           // SWIZZLE_128B tiles need 1024-byte alignment
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); mbar_init(&o_full[i], 1); mbar_init(&s_free[i], 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();        // prologue done: the next kernel may start its own
    pdl_wait();                     // our predecessor's q / k / v are complete and visible

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, kTileBytes);
            tma_load_3d(sQ, &tmQ, q_full, head * kFD, q0, batch);
        }
        __syncwarp();
        for (int j = 0; j < T; ++j) {
            const int b = j & 1;
            mbar_wait(&kv_empty[b], ((j >> 1) & 1) ^ 1);
            if (elect_one()) {
                FTR(0, 1, j);
                mbar_arrive_expect_tx(&k_full[b], kTileBytes);
                tma_load_3d(sK + b * kTileBytes, &tmK, &k_full[b], head * kFD, j * kFK, batch);
                mbar_arrive_expect_tx(&v_full[b], kTileBytes);
                tma_load_3d(sV + b * kTileBytes, &tmV, &v_full[b], head * kFD, j * kFK, batch);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t kIdescS = make_idesc_bf16(128, 128);
        constexpr uint32_t kIdescPV = make_idesc_bf16(128, 64) | (1u << 16);       // B (= V) is MN-major
        const uint32_t aQ = smem_u32(sQ);
        auto issue_S = [&](int j) {
            const int b = j & 1;
            mbar_wait(&k_full[b], (j >> 1) & 1);
            if (lane == 0) FTR(1, 2, j);
            if (j >= 2) mbar_wait(&s_free[b], ((j - 2) >> 1) & 1);      // softmax done with O_{j-2} in this buffer
            tc_fence_after();
            if (lane == 0) FTR(1, 3, j);
            if (elect_one()) {
                const uint64_t ad = make_sw128_kmajor_desc(aQ);
                const uint64_t bd = make_sw128_kmajor_desc(smem_u32(sK + b * kTileBytes));
#pragma unroll
                for (int k = 0; k < kFD / 16; ++k) umma_bf16_ss(tmem_base + b * 128, ad + 2 * k, bd + 2 * k, kIdescS, k != 0);
                umma_commit(&s_full[b]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_S(0);
        for (int j = 0; j < T; ++j) {
            const int b = j & 1;
            if (j + 1 < T) issue_S(j + 1);                                // next scores while softmax_j runs
            mbar_wait(&p_full[b], (j >> 1) & 1);
            if (lane == 0) FTR(1, 4, j);
            mbar_wait(&v_full[b], (j >> 1) & 1);
            tc_fence_after();
            if (lane == 0) FTR(1, 5, j);
            if (elect_one()) {
                const uint32_t aP = smem_u32(sP), aV = smem_u32(sV + b * kTileBytes);
#pragma unroll
                for (int kk = 0; kk < kFK / 16; ++kk) {
                    const uint64_t ad = make_sw128_kmajor_desc(aP + (kk >> 2) * kTileBytes) + 2 * (kk & 3);
                    const uint64_t bd = make_sw128_mnmajor_desc(aV + kk * 2048);
                    umma_bf16_ss(tmem_base + b * 128, ad, bd, kIdescPV, kk != 0);
                }
                umma_commit(&o_full[b]);
                umma_commit(&kv_empty[b]);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ softmax / output: 2 threads per query row
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;          // which 64 keys of a tile / which 32 output columns are mine
        const int r = q * 32 + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const float sl2 = p.scale_log2;
        float m_run = -INFINITY, l_run = 0.f;      // l_run: sum over MY keys only
        float o[kFD / 2];
#pragma unroll
        for (int i = 0; i < kFD / 2; ++i) o[i] = 0.f;

        // pass 1 of a tile: wait for S_j, row max over all 128 keys (both threads of a row compute the same value)
        auto row_max = [&](int j) -> float {
            const int b = j & 1;
            const uint32_t t_s = t_row + b * 128;
            const int kmax = p.nk - j * kFK;
            mbar_wait(&s_full[b], (j >> 1) & 1);
            tc_fence_after();
            if (lane == 0 && q == 2) FTR(2 + half, 6, j);
            float mx = -INFINITY;
            // four round trips of 32 columns: two loads in flight per trip were measured slower (register spills).
            // Also measured (round 2): each thread of a row taking the max of its own 64 keys and swapping it with its
            // partner through shared memory (bf16, one 64-thread named barrier) halves the TMEM reads of this pass but is
            // no faster (595 vs 609 TFLOP/s at 2560 keys): the pass is not what the softmax warps wait on.  Starting every
            // second CTA of an SM 800 / 1500 / 2200 clk late (to put the MUFU-bound exponential phases of the two
            // resident CTAs in anti-phase) changes nothing either (612 / 612 / 610 / 609 TFLOP/s).
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t v0[32];
                tmem_ld32(t_s + c4 * 32, v0);
                tmem_ld_wait();
                if (kmax >= kFK) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v0[i + 1])));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c4 * 32 + i < kmax) mx = fmaxf(mx, __uint_as_float(v0[i]));
                }
            }
            return mx;
        };
        // software pipeline: the row max of tile j + 1 is computed while the tensor core runs P_j V_j
        float mx = row_max(0);
        for (int j = 0; j < T; ++j) {
            const int b = j & 1;
            const uint32_t t_s = t_row + b * 128;
            const int kmax = p.nk - j * kFK;                               // keys valid in this tile (>= 128 unless last)
            const bool ragged = kmax < kFK;                                // warp-uniform
            // ---- lazy rescale: keep exponentiating against the stale max until it is off by more than 2^8;
            // softmax is invariant to the reference value, p <= 256 is harmless in fp32 / bf16.
            const bool upd = (mx - m_run) * sl2 > 8.0f;                    // first tile: m_run = -inf -> true
            if (__any_sync(0xffffffffu, upd)) {
                const float m_new = upd ? mx : m_run;
                const float alpha = ex2((m_run - m_new) * sl2);            // 1 for rows that keep their max
                m_run = m_new;
                l_run *= alpha;
#pragma unroll
                for (int i = 0; i < kFD / 2; ++i) o[i] *= alpha;
            }
            const float msc = m_run * sl2;
            float rs = 0.f;
            // ---- pass 2: my 64 keys: P = exp2(S*scale - m) -> bf16 A operand in smem (P_{j-1} was consumed: o_full[j-1]
            // awaited by both threads of the row before they got here)
            uint8_t* prow = sP + half * kTileBytes + r * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(t_s + half * 64 + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float xx = fmaf(__uint_as_float(v[g4 * 8 + i]), sl2, -msc);
                        e[i] = ((DCB_FLASH_POLY_MASK >> g4) & 1) ? ex2_poly(xx) : ex2(xx);
                    }
                    if (ragged) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (half * 64 + c * 32 + g4 * 8 + i >= kmax) e[i] = 0.f;
                    }
                    rs += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
                    uint4 pk;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(e[0], e[1]), h1 = __floats2bfloat162_rn(e[2], e[3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(e[4], e[5]), h3 = __floats2bfloat162_rn(e[6], e[7]);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    const int chunk = c * 4 + g4;                          // 16-byte chunk inside the 128-byte row
                    *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = pk;
                }
            }
            l_run += rs;
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            if (lane == 0 && q == 2) FTR(2 + half, 7, j);
            // ---- row max of the next tile (its S is already in the other TMEM buffer) while P_j V_j runs
            float mx_next = -INFINITY;
            if (j + 1 < T) mx_next = row_max(j + 1);
            // ---- my 32 columns of O += P_j V_j
            if (lane == 0 && q == 2) FTR(2 + half, 8, j);
            mbar_wait(&o_full[b], (j >> 1) & 1);
            tc_fence_after();
            if (lane == 0 && q == 2) FTR(2 + half, 9, j);
            {
                uint32_t v0[32];
                tmem_ld32(t_s + half * 32, v0);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(v0[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[b]);
            mx = mx_next;
        }
        // ---- combine the two partial row sums (through the idle second P half), normalise, stage as bf16
        // (SWIZZLE_128B rows of 64) in sP half 0, TMA store (rows >= nq are clipped)
        float* s_l = reinterpret_cast<float*>(sP + kTileBytes);           // [2][128]; all PV MMAs have retired
        s_l[half * 128 + r] = l_run;
        named_bar_sync(1, 256);
        const float inv = 1.f / (s_l[r] + s_l[128 + r]);
        uint8_t* orow = sP + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 pk;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * c] * inv, o[8 * c + 1] * inv);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(orow + (((half * 4 + c) ^ (r & 7)) << 4)) = pk;
        }
        fence_proxy_async();
        named_bar_sync(1, 256);
        if (warp == 2 && lane == 0) {
            tma_store_3d(&tmO, sP, head * kFD, q0, batch);
            tma_store_commit();
            tma_store_wait_all<0>();
        }
    }
    if (warp == 0 && lane == 0) FTR_DONE(0);
    if (warp == 1 && lane == 0) FTR_DONE(1);
    if ((warp == 2 || warp == 6) && lane == 0) FTR_DONE(2 + ((warp - 2) >> 2));
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem_base);
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode3(CUtensorMap* m, const void* base, long long ld, int cols, int n, int batch) {
    static EncodeTiledFn2 fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return set_error(DCB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
        fn = reinterpret_cast<EncodeTiledFn2>(ptr);
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)n, (cuuint64_t)batch};
    const cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)n};
    const cuuint32_t box[3] = {64, 128, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return set_error(DCB_ERR_CUDA, "flash_tc: cuTensorMapEncodeTiled failed (rc=%d)", (int)rc);
    return DCB_OK;
}

// Single-segment self-attention with nq == nk on the tensor-core path; returns DCB_ERR_UNSUPPORTED if the
// arguments do not fit (the caller then uses the mma.sync kernel).
int flash_tc_launch(const dcb_attention_args* a, cudaStream_t stream) {
    if (a->nseg != 1 || a->nk[0] != a->nq || a->kv_batch_div[0] != 1 || a->seg_scale[0] != 1.0f || a->head_dim != 64)
        return DCB_ERR_UNSUPPORTED;
    const uintptr_t al = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k[0]) |
                         reinterpret_cast<uintptr_t>(a->v[0]) | reinterpret_cast<uintptr_t>(a->o);
    if ((al & 15) || (a->ldq & 7) || (a->ldk[0] & 7) || (a->ldv[0] & 7) || (a->ldo & 7)) return DCB_ERR_UNSUPPORTED;
    const int cols = a->heads * 64;
    CUtensorMap tq, tk, tv, to;
    int rc;
    if ((rc = encode3(&tq, a->q, a->ldq, cols, a->nq, a->batch))) return rc;
    if ((rc = encode3(&tk, a->k[0], a->ldk[0], cols, a->nq, a->batch))) return rc;
    if ((rc = encode3(&tv, a->v[0], a->ldv[0], cols, a->nq, a->batch))) return rc;
    if ((rc = encode3(&to, a->o, a->ldo, cols, a->nq, a->batch))) return rc;
    static bool cfg[kMaxDevices] = {};
    const int dv = current_device() & (kMaxDevices - 1);
    if (!cfg[dv]) {
        cudaError_t e = cudaFuncSetAttribute(flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem);
        if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "flash_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        cfg[dv] = true;
    }
    FlashTcParams p;
    p.nk = a->nq;
    p.ntiles = (a->nq + kFK - 1) / kFK;
    p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
    dim3 grid((a->nq + kFQ - 1) / kFQ, a->heads, a->batch);
    launch_pdl(kPdlAttention, flash_tc_kernel, grid, dim3(kFThreads), kFSmem, stream, tq, tk, tv, to, p);
    return check_launch("flash_tc");
}

}  // namespace dcb

#ifdef DCB_TRACE
extern "C" int dcb_debug_flash_trace(unsigned long long* dst, int cap) {
    cudaDeviceSynchronize();
    if (cap < 4 * 4096) return -1;
    cudaMemcpyFromSymbol(dst, dcb::g_flash_trace, 4 * 4096 * sizeof(unsigned long long));
    return 4 * 4096;
}
#endif
