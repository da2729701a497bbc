// tcgen05 / TMEM / TMA flash attention for the spatial self-attention of the UNet (attention.py:101-125, head_dim 64).
//
// CTA = one 128-query tile of one (batch, head); 192 threads:
//   warp 0      TMA producer: Q tile once, then K_j / V_j tiles (128 keys x 64) into 2-deep rings (SWIZZLE_128B)
//   warp 1      TMEM owner + tcgen05.mma issuer:  S_j = Q K_j^T (128x128, fp32 in TMEM, double buffered) and
//               O_j = P_j V_j (128x64, written over the consumed S_j columns; V is the MN-major B operand)
//   warps 2..5  softmax: one thread per query row (no shuffles): tcgen05.ld S_j, running max / sum in fp32,
//               P_j = exp2(...) as bf16 into smem (A operand of the PV MMA), O += O_j with the usual rescale
//               kept in registers, final O / l -> bf16 -> swizzled smem -> TMA store.
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
constexpr int kFThreads = 192;
constexpr int kFSmem = 7 * kTileBytes + 256;   // Q, K[2], V[2], P[2 halves] + barriers

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

    if ((smem_u32(fsm) & 1023) != 0) __trap();           // SWIZZLE_128B tiles need 1024-byte alignment
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&s_free[i], 4);
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
            if (j >= 2) mbar_wait(&s_free[b], ((j - 2) >> 1) & 1);      // softmax done with O_{j-2} in this buffer
            tc_fence_after();
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
            mbar_wait(&v_full[b], (j >> 1) & 1);
            tc_fence_after();
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
        // ------------------------------------------------------------------ softmax / output: thread = query row
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const float sl2 = p.scale_log2;
        float m_run = -INFINITY, l_run = 0.f;
        float o[kFD];
#pragma unroll
        for (int i = 0; i < kFD; ++i) o[i] = 0.f;

        for (int j = 0; j < T; ++j) {
            const int b = j & 1;
            const uint32_t t_s = t_row + b * 128;
            const int kmax = p.nk - j * kFK;                               // keys valid in this tile (>= 128 unless last)
            mbar_wait(&s_full[b], (j >> 1) & 1);
            tc_fence_after();
            const bool ragged = kmax < kFK;                                // warp-uniform
            // ---- pass 1: tile row max (two 64-column batches so only two TMEM round trips are exposed)
            float mx = -INFINITY;
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t v0[32], v1[32];
                tmem_ld32(t_s + c2 * 64, v0);
                tmem_ld32(t_s + c2 * 64 + 32, v1);
                tmem_ld_wait();
                if (!ragged) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (c2 * 64 + i < kmax) mx = fmaxf(mx, __uint_as_float(v0[i]));
                        if (c2 * 64 + 32 + i < kmax) mx = fmaxf(mx, __uint_as_float(v1[i]));
                    }
                }
            }
            // ---- lazy rescale: keep exponentiating against the stale max until it is off by more than 2^8;
            // softmax is invariant to the reference value, p <= 256 is harmless in fp32 / bf16.
            const bool upd = (mx - m_run) * sl2 > 8.0f;                    // first tile: m_run = -inf -> true
            if (__any_sync(0xffffffffu, upd)) {
                const float m_new = upd ? mx : m_run;
                const float alpha = ex2((m_run - m_new) * sl2);            // 1 for rows that keep their max
                m_run = m_new;
                l_run *= alpha;
#pragma unroll
                for (int i = 0; i < kFD; ++i) o[i] *= alpha;
            }
            const float msc = m_run * sl2;
            float rs = 0.f;
            // ---- pass 2: P = exp2(S*scale - m) -> bf16 A operand in smem (P_{j-1} was consumed: o_full[j-1] awaited)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld32(t_s + c * 32, v);
                tmem_ld_wait();
                uint8_t* prow = sP + (c >> 1) * kTileBytes + r * 128;
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) e[i] = ex2(fmaf(__uint_as_float(v[g4 * 8 + i]), sl2, -msc));
                    if (ragged) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (c * 32 + g4 * 8 + i >= kmax) e[i] = 0.f;
                    }
                    rs += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
                    uint4 pk;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(e[0], e[1]), h1 = __floats2bfloat162_rn(e[2], e[3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(e[4], e[5]), h3 = __floats2bfloat162_rn(e[6], e[7]);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    const int chunk = (c & 1) * 4 + g4;                    // 16-byte chunk inside the 128-byte row
                    *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = pk;
                }
            }
            l_run += rs;
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            // ---- O += P_j V_j
            mbar_wait(&o_full[b], (j >> 1) & 1);
            tc_fence_after();
            {
                uint32_t v0[32], v1[32];
                tmem_ld32(t_s, v0);
                tmem_ld32(t_s + 32, v1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) { o[i] += __uint_as_float(v0[i]); o[32 + i] += __uint_as_float(v1[i]); }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[b]);
        }
        // ---- normalise, stage as bf16 (SWIZZLE_128B rows of 64) in sP half 0, TMA store (rows >= nq are clipped)
        const float inv = 1.f / l_run;
        uint8_t* orow = sP + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint4 pk;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * c] * inv, o[8 * c + 1] * inv);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(orow + ((c ^ (r & 7)) << 4)) = pk;
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (warp == 2 && lane == 0) {
            tma_store_3d(&tmO, sP, head * kFD, q0, batch);
            tma_store_commit();
            tma_store_wait_all<0>();
        }
    }
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
    static bool cfg = false;
    if (!cfg) {
        cudaError_t e = cudaFuncSetAttribute(flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem);
        if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "flash_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        cfg = true;
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
