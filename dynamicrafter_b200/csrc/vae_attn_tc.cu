// Flash-style single-head attention with a WIDE head (d = C <= 512) on tcgen05 / TMEM / TMA: the mid-block attention of
// the VAE (ae_modules.py:53-78: softmax(q k^T * C^-1/2) v over the h*w tokens of one frame, C = 512 at full size).
// The N x N score matrix never exists in HBM (the eager reference materialises 9216^2 fp32 = 340 MB per frame at
// 576x1024) and all frames run in ONE launch.
//
// CTA = (128-query tile, 256-column slice of the output, frame); 192 threads:
//   warp 0      TMA producer: a 4-stage ring of {Q chunk, K chunk} (128 x 64 bf16 each) for the score phase and of
//               V chunks (128 keys x 64 output columns) for the P V phase, in the exact order the MMA warp consumes them
//   warp 1      TMEM owner + MMA issuer:  S_j = sum over the d/64 chunks of Q_c K_{j,c}^T (128x128 fp32, double buffered
//               in TMEM columns 0..255) and O[:, slice] += P_j V_j (TMEM columns 256..511, accumulated IN TMEM)
//   warps 2..5  softmax: one thread per query row; online softmax with lazy rescale (the running maximum is only
//               refreshed when it grew by more than 2^8, and only then O is read back, scaled and stored again with
//               tcgen05.ld / tcgen05.st); P_j -> bf16 A operand in shared memory; at the end O / l -> bf16 -> TMA store.
// d = 512 does not fit TMEM next to a double-buffered S, so the output is computed in 256-column slices (grid.y) and
// the scores are recomputed per slice (1.5x the FLOPs of the ideal; the whole VAE attention is < 0.1 % of a video).
#include "dcb_internal.h"
#include "ptx.cuh"

namespace dcb {

constexpr int kVTile = 128 * 128;             // 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kVStages = 4;
constexpr int kVStageBytes = 2 * kVTile;
constexpr int kVThreads = 192;
constexpr int kVSmem = kVStages * kVStageBytes + 2 * kVTile + 256;      // ring + P (2 key halves) + barriers
constexpr int kVSlice = 256;                  // output columns per CTA

struct VaeAttnParams {
    int nk, ntiles;          // tokens per frame, key tiles
    int dchunks;             // C / 64
    int slice_chunks;        // 64-column chunks of THIS launch's widest slice (<= 4)
    int C;
    float scale_log2;
};

__device__ __forceinline__ uint64_t vdesc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ void vtma_ld3(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void vtma_st3(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float vex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kVThreads, 1)
vae_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const VaeAttnParams p) {
    extern __shared__ __align__(1024) uint8_t vsm[];
    uint8_t* ring = vsm;                                        // [kVStages][2 tiles]
    uint8_t* sP = vsm + kVStages * kVStageBytes;                // [2 key halves of 64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kVTile);
    uint64_t* full = bars;                      // [kVStages]
    uint64_t* empty = bars + kVStages;          // [kVStages]
    uint64_t* s_full = bars + 2 * kVStages;     // [2]
    uint64_t* s_free = bars + 2 * kVStages + 2; // [2]
    uint64_t* p_full = bars + 2 * kVStages + 4;
    uint64_t* o_done = bars + 2 * kVStages + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kVStages + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, slice = blockIdx.y, frame = blockIdx.z;
    const int T = p.ntiles, DC = p.dchunks;
    const int col0 = slice * kVSlice;                           // first output column of this CTA
    const int na = min(kVSlice, p.C - col0) / 64;               // 64-column chunks of this slice (1..4)

    if ((smem_u32(vsm) & 1023) != 0) __trap();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
        for (int i = 0; i < kVStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); }
        mbar_init(p_full, 4);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (order = MMA consumption order)
        uint32_t it = 0;
        auto load_S = [&](int j) {
            for (int c = 0; c < DC; ++c, ++it) {
                const uint32_t s = it % kVStages, ph = (it / kVStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[s], 2 * kVTile);
                    vtma_ld3(ring + s * kVStageBytes, &tmQ, &full[s], c * 64, q0, frame);
                    vtma_ld3(ring + s * kVStageBytes + kVTile, &tmK, &full[s], c * 64, j * 128, frame);
                }
                __syncwarp();
            }
        };
        auto load_V = [&](int j) {
            for (int a = 0; a < na; ++a, ++it) {
                const uint32_t s = it % kVStages, ph = (it / kVStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[s], kVTile);
                    vtma_ld3(ring + s * kVStageBytes, &tmV, &full[s], col0 + a * 64, j * 128, frame);
                }
                __syncwarp();
            }
        };
        load_S(0);
        for (int j = 0; j < T; ++j) {
            if (j + 1 < T) load_S(j + 1);
            load_V(j);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t kIdescS = make_idesc_bf16(128, 128);
        constexpr uint32_t kIdescPV = make_idesc_bf16(128, 64) | (1u << 16);       // B (= V) is MN-major
        uint32_t it = 0;
        auto issue_S = [&](int j) {
            const int b = j & 1;
            if (j >= 2) mbar_wait(&s_free[b], ((j - 2) >> 1) & 1);
            tc_fence_after();
            for (int c = 0; c < DC; ++c, ++it) {
                const uint32_t s = it % kVStages, ph = (it / kVStages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t base = smem_u32(ring + s * kVStageBytes);
                    const uint64_t ad = make_sw128_kmajor_desc(base), bd = make_sw128_kmajor_desc(base + kVTile);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + b * 128, ad + 2 * k, bd + 2 * k, kIdescS, (c | k) != 0);
                    umma_commit(&empty[s]);
                    if (c == DC - 1) umma_commit(&s_full[b]);
                }
                __syncwarp();
            }
        };
        auto issue_PV = [&](int j) {
            mbar_wait(p_full, j & 1);
            tc_fence_after();
            for (int a = 0; a < na; ++a, ++it) {
                const uint32_t s = it % kVStages, ph = (it / kVStages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t aP = smem_u32(sP), aV = smem_u32(ring + s * kVStageBytes);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint64_t ad = make_sw128_kmajor_desc(aP + (kk >> 2) * kVTile) + 2 * (kk & 3);
                        const uint64_t bd = vdesc_mn(aV + kk * 2048);
                        umma_bf16_ss(tmem_base + 256 + a * 64, ad, bd, kIdescPV, (j | kk) != 0);
                    }
                    umma_commit(&empty[s]);
                    if (a == na - 1) umma_commit(o_done);
                }
                __syncwarp();
            }
        };
        issue_S(0);
        for (int j = 0; j < T; ++j) {
            if (j + 1 < T) issue_S(j + 1);
            issue_PV(j);
        }
    } else {
        // ------------------------------------------------------------------ softmax: one thread per query row
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const float sl2 = p.scale_log2;
        float m_run = -INFINITY, l_run = 0.f;
        uint8_t* prow = sP + r * 128;
        for (int j = 0; j < T; ++j) {
            const int b = j & 1;
            const uint32_t t_s = t_row + b * 128;
            const int kmax = p.nk - j * 128;
            const bool ragged = kmax < 128;
            mbar_wait(&s_full[b], (j >> 1) & 1);
            tc_fence_after();
            float mx = -INFINITY;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t v[32];
                tmem_ld32(t_s + c4 * 32, v);
                tmem_ld_wait();
                if (!ragged) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c4 * 32 + i < kmax) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            // P_{j-1} V_{j-1} has retired: the P tile is free and O is stable
            if (j > 0) {
                mbar_wait(o_done, (j - 1) & 1);
                tc_fence_after();
            }
            const bool upd = (mx - m_run) * sl2 > 8.0f;                      // first tile: m_run = -inf -> true
            if (__any_sync(0xffffffffu, upd)) {
                const float m_new = upd ? mx : m_run;
                const float alpha = vex2((m_run - m_new) * sl2);              // 1 for rows that keep their reference
                m_run = m_new;
                l_run *= alpha;
                if (j > 0) {
                    for (int c = 0; c < na * 2; ++c) {
                        uint32_t v[32];
                        tmem_ld32(t_row + 256 + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st32(t_row + 256 + c * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            const float msc = m_run * sl2;
            float rs = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t v[32];
                tmem_ld32(t_s + c4 * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        e[i] = vex2(fmaf(__uint_as_float(v[g4 * 8 + i]), sl2, -msc));
                        if (ragged && c4 * 32 + g4 * 8 + i >= kmax) e[i] = 0.f;
                    }
                    rs += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
                    uint4 pk;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(e[0], e[1]), h1 = __floats2bfloat162_rn(e[2], e[3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(e[4], e[5]), h3 = __floats2bfloat162_rn(e[6], e[7]);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    const int chunk = c4 * 4 + g4;                       // 16-byte chunk of the 256-byte P row (128 keys)
                    *reinterpret_cast<uint4*>(prow + (chunk >> 3) * kVTile + (((chunk & 7) ^ (r & 7)) << 4)) = pk;
                }
            }
            l_run += rs;
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(p_full);
                mbar_arrive(&s_free[b]);
            }
        }
        // ---- O / l -> bf16 -> staging (the ring: every load has been consumed) -> TMA store, 64 columns per box
        mbar_wait(o_done, (T - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l_run;
        for (int c = 0; c < na * 2; ++c) {
            uint32_t v[32];
            tmem_ld32(t_row + 256 + c * 32, v);
            tmem_ld_wait();
            uint8_t* orow = ring + (c >> 1) * kVTile + r * 128;
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[g4 * 8]) * inv, __uint_as_float(v[g4 * 8 + 1]) * inv);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[g4 * 8 + 2]) * inv, __uint_as_float(v[g4 * 8 + 3]) * inv);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[g4 * 8 + 4]) * inv, __uint_as_float(v[g4 * 8 + 5]) * inv);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(v[g4 * 8 + 6]) * inv, __uint_as_float(v[g4 * 8 + 7]) * inv);
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                const int chunk = (c & 1) * 4 + g4;
                *reinterpret_cast<uint4*>(orow + ((chunk ^ (r & 7)) << 4)) = pk;
            }
        }
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync(1, 128);
        if (warp == 2 && lane == 0) {
            for (int a = 0; a < na; ++a) vtma_st3(&tmO, ring + a * kVTile, col0 + a * 64, q0, frame);
            tma_store_commit();
            tma_store_wait_all<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*VEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int vencode(CUtensorMap* m, const void* base, long long ld, int cols, int n, int batch) {
    static VEncodeFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return set_error(DCB_ERR_CUDA, "vae_attention: cuTensorMapEncodeTiled entry point not found");
        fn = reinterpret_cast<VEncodeFn>(ptr);
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)n, (cuuint64_t)batch};
    const cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)n};
    const cuuint32_t box[3] = {64, 128, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return set_error(DCB_ERR_CUDA, "vae_attention: cuTensorMapEncodeTiled failed (rc=%d)", (int)rc);
    return DCB_OK;
}

}  // namespace dcb

using namespace dcb;

extern "C" int dcb_vae_attention(const void* q, const void* k, const void* v, void* o, long long ldq, long long ldk,
                                 long long ldv, long long ldo, int frames, int tokens, int C, float softmax_scale,
                                 void* stream) {
    if (!q || !k || !v || !o || frames < 1 || tokens < 1) return set_error(DCB_ERR_INVALID, "vae_attention: bad argument");
    if (C < 64 || C > 512 || (C & 63)) return set_error(DCB_ERR_UNSUPPORTED, "vae_attention: C=%d must be a multiple of 64, <= 512", C);
    const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                         reinterpret_cast<uintptr_t>(o);
    if ((al & 15) || (ldq & 7) || (ldk & 7) || (ldv & 7) || (ldo & 7))
        return set_error(DCB_ERR_INVALID, "vae_attention: 16-byte aligned bases and row pitches required");
    if (frames > 65535) return set_error(DCB_ERR_INVALID, "vae_attention: too many frames");
    CUtensorMap tq, tk, tv, to;
    int rc;
    if ((rc = vencode(&tq, q, ldq, C, tokens, frames))) return rc;
    if ((rc = vencode(&tk, k, ldk, C, tokens, frames))) return rc;
    if ((rc = vencode(&tv, v, ldv, C, tokens, frames))) return rc;
    if ((rc = vencode(&to, o, ldo, C, tokens, frames))) return rc;
    static bool cfg[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || !cfg[dev]) {
        cudaError_t e = cudaFuncSetAttribute(vae_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVSmem);
        if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "vae_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 16) cfg[dev] = true;
    }
    VaeAttnParams p;
    p.nk = tokens;
    p.ntiles = (tokens + 127) / 128;
    p.dchunks = C / 64;
    p.slice_chunks = (C < kVSlice ? C : kVSlice) / 64;
    p.C = C;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    dim3 grid((tokens + 127) / 128, (C + kVSlice - 1) / kVSlice, frames);
    launch_pdl(kPdlAttention, vae_attn_tc_kernel, grid, dim3(kVThreads), kVSmem, static_cast<cudaStream_t>(stream), tq, tk, tv,
               to, p);
    return check_launch("vae_attention");
}
