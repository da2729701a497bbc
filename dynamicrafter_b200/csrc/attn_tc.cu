// tcgen05 / TMEM / TMA kernels for the two SHORT-key attention roles of the UNet (head_dim 64):
//
//   kCross     text + image cross-attention (attention.py:89-94,129-142): 77 text keys shared by the frames of a video
//              + 16 image keys per frame, two independent softmaxes.  The 96-row K / V tile of one (frame, head) stays
//              resident in shared memory while the CTA streams 128-query tiles:  S = Q K^T is ONE 128x96 UMMA chain,
//              the softmax threads normalise both segments and write P (bf16, already divided by the segment sums and
//              multiplied by the image scale), O = P V is one 128x64 chain over the 96 keys.
//   kTemporal  self-attention over <= 16 frames per (batch, location, head) (attention.py:387-390).  8 locations x 16
//              frames are packed into one 128-row tile (8 TMA boxes of 16 frame rows each, read with the frame stride
//              of the channels-last layout, so no '(b t) hw c <-> (b hw) t c' transposes exist); S = Q K^T is a
//              128x128 UMMA chain of which only the 8 diagonal 16x16 blocks are read, P is block diagonal (the
//              off-diagonal zeros are written once), O = P V.
//
// CTA = 320 threads: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2..5 / 6..9 two softmax groups working
// on alternate tiles (ping-pong: group A's exponentials run under group B's TMEM round trips and MMAs).  Each group owns
// a "set": Q (+K, V) smem tiles, a P tile, 128 TMEM columns (S, then O over the consumed S columns).  Work = a
// contiguous range of the global tile list per CTA (grid = number of SMs), so K / V are reloaded only when the
// (frame, head) changes.  Everything that touches HBM is TMA; the score matrix never leaves the SM.
#include "dcb_internal.h"
#include "ptx.cuh"

namespace dcb {

constexpr int kTTile = 128 * 128;            // 128 rows x 64 bf16 (SWIZZLE_128B)
constexpr int kTThreads = 320;
constexpr int kCrossKeys = 96;               // 80 text slots (77 used) + 16 image slots
constexpr int kCrossText = 80;
constexpr int kCrossKVBytes = kCrossKeys * 128;
// load ring (decoupled from the two compute sets so that several tiles' worth of HBM requests are always in flight)
constexpr int kCrossStages = 5;              // Q tiles
constexpr int kTempStages = 4;               // {Q, K, V} tiles
constexpr int kCrossSmem = 4 * kCrossKVBytes + kCrossStages * kTTile + 2 * 2 * kTTile + 256;   // K/V x2, Q ring, P x2
constexpr int kTempSmem = kTempStages * 3 * kTTile + 2 * kTTile + 256;                         // {Q,K,V} ring, P x1

struct AttnTcParams {
    int mode_tiles;          // tiles per (batch, head): cross: ceil(nq / 128); temporal: ceil(HW / 8)
    int heads, batch;
    int total_tiles;         // batch * heads * mode_tiles
    int nk0, nk1;            // cross: valid text / image keys
    int kv_div0, kv_div1;    // cross: kv batch = batch / div
    int L;                   // temporal: frames (<= 16)
    float scale_log2, seg0, seg1;
};

__device__ __forceinline__ uint64_t desc_sw128_mn(uint32_t smem_addr) {     // MN-major B operand (V), 8-row groups 1024 B apart
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ void tma_ld3(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_st3(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <bool kTemporal>
__global__ void __launch_bounds__(kTThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
               const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
               const __grid_constant__ CUtensorMap tmV1, const __grid_constant__ CUtensorMap tmO, const AttnTcParams p) {
    constexpr int NS = kTemporal ? kTempStages : kCrossStages;
    constexpr int kStageBytes = kTemporal ? 3 * kTTile : kTTile;
    constexpr int NP = kTemporal ? 1 : 2;                    // P buffers (temporal: one, shared by both groups)
    extern __shared__ __align__(1024) uint8_t asm_[];
    // ---- shared memory carve-up
    uint8_t* kv_base = asm_;                                 // cross: 2 x {K, V} of 96 x 128 B (double buffered over pairs)
    uint8_t* ring = kTemporal ? asm_ : asm_ + 4 * kCrossKVBytes;
    uint8_t* p_base = ring + NS * kStageBytes;
    auto sQ = [&](int st) { return ring + st * kStageBytes; };
    auto sP = [&](int s) { return p_base + (NP == 2 ? s : 0) * 2 * kTTile; };
    uint64_t* bars = reinterpret_cast<uint64_t*>(p_base + NP * 2 * kTTile);
    uint64_t* q_full = bars;                  // [NS]
    uint64_t* stage_free = bars + NS;         // [NS]
    uint64_t* s_full = bars + 2 * NS;         // [2]
    uint64_t* p_full = bars + 2 * NS + 2;     // [2]
    uint64_t* o_full = bars + 2 * NS + 4;     // [2]
    uint64_t* s_free = bars + 2 * NS + 6;     // [2]
    uint64_t* kv_full = bars + 2 * NS + 8;    // [2] cross
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // contiguous tile range of this CTA
    const long long T = p.total_tiles;
    const int t_begin = static_cast<int>(T * blockIdx.x / gridDim.x);
    const int t_end = static_cast<int>(T * (blockIdx.x + 1) / gridDim.x);
    const int ntiles = t_end - t_begin;
    // cross: sequence number (within this CTA) of the (batch, head) pair a tile belongs to
    const int pair0 = t_begin / p.mode_tiles;
    auto pair_seq = [&](int i) { return (t_begin + i) / p.mode_tiles - pair0; };

    if ((smem_u32(asm_) & 1023) != 0) __trap();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK0); tma_prefetch_desc(&tmV0); tma_prefetch_desc(&tmO);
        if (!kTemporal) { tma_prefetch_desc(&tmK1); tma_prefetch_desc(&tmV1); }
        for (int i = 0; i < NS; ++i) { mbar_init(&q_full[i], 1); mbar_init(&stage_free[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&s_free[i], 4);
            mbar_init(&kv_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    // tile index -> (batch, head, tile within the pair)
    auto decode = [&](int t, int& b, int& h, int& j) {
        j = t % p.mode_tiles;
        const int bh = t / p.mode_tiles;
        h = bh % p.heads;
        b = bh / p.heads;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        int cur_seq = -1;
        int first_of_prev = 0;                     // first tile (CTA-local index) of the previous pair
        for (int i = 0; i < ntiles; ++i) {
            const int st = i % NS, u = i / NS;
            int b, h, j;
            decode(t_begin + i, b, h, j);
            mbar_wait(&stage_free[st], (u & 1) ^ 1);
            if (!kTemporal) {
                const int m = pair_seq(i);
                if (m != cur_seq) {
                    // K / V buffer m & 1 was last read by the tiles of pair m - 2, the last of which is tile
                    // first_of_prev - 1: it and its predecessor (the other softmax group) must have finished
                    if (m >= 2) {
                        for (int e = first_of_prev - 1; e >= first_of_prev - 2 && e >= 0; --e)
                            if (e + NS > i) mbar_wait(&stage_free[e % NS], (e / NS) & 1);   // (older: seen when its stage was reloaded)
                    }
                    first_of_prev = i;
                    cur_seq = m;
                    uint8_t* kb = kv_base + (m & 1) * 2 * kCrossKVBytes;
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&kv_full[m & 1], 2 * kCrossKVBytes);
                        tma_ld3(kb, &tmK0, &kv_full[m & 1], h * 64, 0, b / p.kv_div0);
                        tma_ld3(kb + kCrossKVBytes, &tmV0, &kv_full[m & 1], h * 64, 0, b / p.kv_div0);
                        tma_ld3(kb + kCrossText * 128, &tmK1, &kv_full[m & 1], h * 64, 0, b / p.kv_div1);
                        tma_ld3(kb + kCrossKVBytes + kCrossText * 128, &tmV1, &kv_full[m & 1], h * 64, 0, b / p.kv_div1);
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    mbar_arrive_expect_tx(&q_full[st], kTTile);
                    tma_ld3(sQ(st), &tmQ, &q_full[st], h * 64, j * 128, b);
                }
            } else {
                if (elect_one()) {
                    mbar_arrive_expect_tx(&q_full[st], 3 * kTTile);
#pragma unroll
                    for (int loc = 0; loc < 8; ++loc) {
                        tma_load_4d(sQ(st) + loc * 2048, &tmQ, &q_full[st], h * 64, j * 8 + loc, 0, b);
                        tma_load_4d(sQ(st) + kTTile + loc * 2048, &tmK0, &q_full[st], h * 64, j * 8 + loc, 0, b);
                        tma_load_4d(sQ(st) + 2 * kTTile + loc * 2048, &tmV0, &q_full[st], h * 64, j * 8 + loc, 0, b);
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr int kSN = kTemporal ? 128 : kCrossKeys;
        constexpr uint32_t kIdescS = make_idesc_bf16(128, kSN);
        constexpr uint32_t kIdescPV = make_idesc_bf16(128, 64) | (1u << 16);       // B (= V) is MN-major
        int cur_seq = -1;
        auto k_of = [&](int i) -> uint32_t {        // smem address of the K tile of CTA-local tile i (V follows)
            if (kTemporal) return smem_u32(sQ(i % NS) + kTTile);
            return smem_u32(kv_base + (pair_seq(i) & 1) * 2 * kCrossKVBytes);
        };
        auto issue_PV = [&](int j) {
            const int s = j & 1;
            mbar_wait(&p_full[s], (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t aP = smem_u32(sP(s)), aV = k_of(j) + (kTemporal ? kTTile : kCrossKVBytes);
#pragma unroll
                for (int kk = 0; kk < kSN / 16; ++kk) {
                    const uint64_t ad = make_sw128_kmajor_desc(aP + (kk >> 2) * kTTile) + 2 * (kk & 3);
                    const uint64_t bd = desc_sw128_mn(aV + kk * 2048);
                    umma_bf16_ss(tmem_base + s * 128, ad, bd, kIdescPV, kk != 0);
                }
                umma_commit(&o_full[s]);
            }
            __syncwarp();
        };
        for (int i = 0; i < ntiles; ++i) {
            const int s = i & 1, st = i % NS;
            if (!kTemporal) {
                const int m = pair_seq(i);
                if (m != cur_seq) {
                    cur_seq = m;
                    mbar_wait(&kv_full[m & 1], (m >> 1) & 1);
                }
            }
            mbar_wait(&q_full[st], (i / NS) & 1);
            mbar_wait(&s_free[s], ((i >> 1) & 1) ^ 1);            // softmax group done with O_{i-2} in this TMEM set
            tc_fence_after();
            if (elect_one()) {
                const uint64_t ad = make_sw128_kmajor_desc(smem_u32(sQ(st)));
                const uint64_t bd = make_sw128_kmajor_desc(k_of(i));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + s * 128, ad + 2 * k, bd + 2 * k, kIdescS, k != 0);
                umma_commit(&s_full[s]);
            }
            __syncwarp();
            if (i >= 1) issue_PV(i - 1);                          // S_i is in flight while the softmax of tile i-1 finishes
        }
        if (ntiles >= 1) issue_PV(ntiles - 1);
    } else {
        // ------------------------------------------------------------------ softmax groups
        const int q = warp & 3;                         // TMEM lane quarter
        const int g = (warp - 2) >> 2;                  // group = TMEM set = tile parity
        const int r = q * 32 + lane;                    // tile row of this thread
        const bool leader = (warp - 2) == g * 4 && lane == 0;
        const uint32_t t_row = tmem_base + g * 128 + (static_cast<uint32_t>(q * 32) << 16);
        const float sl2 = p.scale_log2;
        uint8_t* const prow0 = sP(g) + r * 128;          // row r of P tile 0 (keys 0..63); tile 1 at + kTTile
        if (kTemporal) {
            // zero the (shared) P tile once: only the diagonal blocks are rewritten per tile
            if (g == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    *reinterpret_cast<uint4*>(prow0 + (c << 4)) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(prow0 + kTTile + (c << 4)) = make_uint4(0, 0, 0, 0);
                }
            }
            named_bar_sync(3, 256);
        }
        for (int i = g, it = 0; i < ntiles; i += 2, ++it) {
            const uint32_t ph = it & 1;
            const int st = i % NS;
            int b, h, j;
            decode(t_begin + i, b, h, j);
            mbar_wait(&s_full[g], ph);
            tc_fence_after();
            if (!kTemporal) {
                // ---- two segments: text keys [0, nk0), image keys [80, 80 + nk1)
                uint32_t v[kCrossKeys];
                {
                    uint32_t a0[32], a1[32], a2[32];
                    tmem_ld32(t_row, a0);
                    tmem_ld32(t_row + 32, a1);
                    tmem_ld32(t_row + 64, a2);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 32; ++k) { v[k] = a0[k]; v[32 + k] = a1[k]; v[64 + k] = a2[k]; }
                }
                float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
                for (int k = 0; k < kCrossText; ++k)
                    if (k < p.nk0) m0 = fmaxf(m0, __uint_as_float(v[k]));
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k < p.nk1) m1 = fmaxf(m1, __uint_as_float(v[kCrossText + k]));
                const float ms0 = m0 * sl2, ms1 = m1 * sl2;
                float e[kCrossKeys];
                float l0 = 0.f, l1 = 0.f;
#pragma unroll
                for (int k = 0; k < kCrossText; ++k) {
                    e[k] = (k < p.nk0) ? ex2f(fmaf(__uint_as_float(v[k]), sl2, -ms0)) : 0.f;
                    l0 += e[k];
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    e[kCrossText + k] = (k < p.nk1) ? ex2f(fmaf(__uint_as_float(v[kCrossText + k]), sl2, -ms1)) : 0.f;
                    l1 += e[kCrossText + k];
                }
                const float i0 = p.seg0 / l0, i1 = (p.nk1 > 0) ? p.seg1 / l1 : 0.f;
#pragma unroll
                for (int c = 0; c < kCrossKeys / 8; ++c) {
                    const float sc = (c < kCrossText / 8) ? i0 : i1;
                    uint4 pk;
                    pk.x = pack2(e[8 * c] * sc, e[8 * c + 1] * sc);
                    pk.y = pack2(e[8 * c + 2] * sc, e[8 * c + 3] * sc);
                    pk.z = pack2(e[8 * c + 4] * sc, e[8 * c + 5] * sc);
                    pk.w = pack2(e[8 * c + 6] * sc, e[8 * c + 7] * sc);
                    *reinterpret_cast<uint4*>(prow0 + (c >> 3) * kTTile + (((c & 7) ^ (r & 7)) << 4)) = pk;
                }
            } else {
                // ---- my sequence = the 16 rows of my location: columns [16 * (r / 16), + 16) of S
                uint32_t a[32];
                tmem_ld32(t_row + q * 32, a);
                tmem_ld_wait();
                const bool hi = lane >= 16;
                float sv[16];
                float m = -INFINITY;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    sv[k] = __uint_as_float(hi ? a[16 + k] : a[k]);
                    if (k < p.L) m = fmaxf(m, sv[k]);
                }
                const float ms = m * sl2;
                float l = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    sv[k] = (k < p.L) ? ex2f(fmaf(sv[k], sl2, -ms)) : 0.f;
                    l += sv[k];
                }
                const float inv = 1.f / l;
                // the single P tile is shared by both groups: P_{i-1} V_{i-1} (the other group's tile) must have retired
                if (i >= 1) mbar_wait(&o_full[g ^ 1], ((i - 1) >> 1) & 1);
                const int c0 = 2 * (r >> 4);                 // first 16-byte chunk of my diagonal block (keys 16 * (r / 16) ..)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = c0 + cc;
                    uint4 pk;
                    pk.x = pack2(sv[8 * cc] * inv, sv[8 * cc + 1] * inv);
                    pk.y = pack2(sv[8 * cc + 2] * inv, sv[8 * cc + 3] * inv);
                    pk.z = pack2(sv[8 * cc + 4] * inv, sv[8 * cc + 5] * inv);
                    pk.w = pack2(sv[8 * cc + 6] * inv, sv[8 * cc + 7] * inv);
                    *reinterpret_cast<uint4*>(prow0 + (c >> 3) * kTTile + (((c & 7) ^ (r & 7)) << 4)) = pk;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[g]);
            // ---- O = P V -> bf16 -> staging (the Q tile of this stage: S has consumed it) -> TMA store
            mbar_wait(&o_full[g], ph);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld32(t_row, o0);
            tmem_ld32(t_row + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);
            uint8_t* orow = sQ(st) + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t* src = (c < 4) ? (o0 + 8 * c) : (o1 + 8 * (c - 4));
                uint4 pk;
                pk.x = pack2(__uint_as_float(src[0]), __uint_as_float(src[1]));
                pk.y = pack2(__uint_as_float(src[2]), __uint_as_float(src[3]));
                pk.z = pack2(__uint_as_float(src[4]), __uint_as_float(src[5]));
                pk.w = pack2(__uint_as_float(src[6]), __uint_as_float(src[7]));
                *reinterpret_cast<uint4*>(orow + ((c ^ (r & 7)) << 4)) = pk;
            }
            fence_proxy_async();
            named_bar_sync(1 + g, 128);
            if (leader) {
                if (!kTemporal) {
                    tma_st3(&tmO, sQ(st), h * 64, j * 128, b);
                } else {
#pragma unroll
                    for (int loc = 0; loc < 8; ++loc) tma_store_4d(&tmO, sQ(st) + loc * 2048, h * 64, j * 8 + loc, 0, b);
                }
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(&stage_free[st]);              // the stage (Q, K, V tiles + staging) is free again
            }
        }
        if (leader) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem_base);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
    static EncodeFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeFn>(ptr);
    }
    return fn;
}

// bf16 tensor map, SWIZZLE_128B, 64-column boxes; dims / strides innermost first (strides in elements, strides[0] = 1)
static int encode_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                       const cuuint32_t* box) {
    EncodeFn fn = encode_fn();
    if (fn == nullptr) return set_error(DCB_ERR_CUDA, "attn_tc: cuTensorMapEncodeTiled entry point not found");
    cuuint64_t gstr[4];
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    for (int i = 1; i < rank; ++i) gstr[i - 1] = strides[i] * 2;
    CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return set_error(DCB_ERR_CUDA, "attn_tc: cuTensorMapEncodeTiled failed (rc=%d)", (int)rc);
    return DCB_OK;
}

template <bool kTemporal>
static int configure(int smem) {
    static bool done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || !done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<kTemporal>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "attn_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 16) done[dev] = true;
    }
    return DCB_OK;
}

// Text + image cross-attention on the tensor-core path.  DCB_ERR_UNSUPPORTED if the arguments do not fit (two
// segments, <= 80 text and <= 16 image keys, 16-byte aligned rows); the caller then uses the mma.sync kernel.
int cross_tc_launch(const dcb_attention_args* a, cudaStream_t stream) {
    if (a->nseg != 2 || a->nk[0] < 1 || a->nk[0] > kCrossText || a->nk[1] < 1 || a->nk[1] > 16 || a->head_dim != 64)
        return DCB_ERR_UNSUPPORTED;
    const uintptr_t al = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->o) |
                         reinterpret_cast<uintptr_t>(a->k[0]) | reinterpret_cast<uintptr_t>(a->v[0]) |
                         reinterpret_cast<uintptr_t>(a->k[1]) | reinterpret_cast<uintptr_t>(a->v[1]);
    if ((al & 15) || (a->ldq & 7) || (a->ldo & 7) || (a->ldk[0] & 7) || (a->ldv[0] & 7) || (a->ldk[1] & 7) || (a->ldv[1] & 7))
        return DCB_ERR_UNSUPPORTED;
    if (a->batch % a->kv_batch_div[0] || a->batch % a->kv_batch_div[1]) return DCB_ERR_UNSUPPORTED;
    const cuuint64_t cols = static_cast<cuuint64_t>(a->heads) * 64;
    CUtensorMap tq, to, tk0, tv0, tk1, tv1;
    int rc;
    {
        const cuuint64_t dims[3] = {cols, (cuuint64_t)a->nq, (cuuint64_t)a->batch};
        const cuuint32_t box[3] = {64, 128, 1};
        const cuuint64_t sq[3] = {1, (cuuint64_t)a->ldq, (cuuint64_t)a->ldq * a->nq};
        const cuuint64_t so[3] = {1, (cuuint64_t)a->ldo, (cuuint64_t)a->ldo * a->nq};
        if ((rc = encode_bf16(&tq, a->q, 3, dims, sq, box))) return rc;
        if ((rc = encode_bf16(&to, a->o, 3, dims, so, box))) return rc;
    }
    for (int s = 0; s < 2; ++s) {
        const cuuint64_t nk = a->nk[s], kb = a->batch / a->kv_batch_div[s];
        const cuuint64_t dims[3] = {cols, nk, kb};
        const cuuint32_t box[3] = {64, s == 0 ? (cuuint32_t)kCrossText : 16u, 1};
        const cuuint64_t sk[3] = {1, (cuuint64_t)a->ldk[s], (cuuint64_t)a->ldk[s] * nk};
        const cuuint64_t sv[3] = {1, (cuuint64_t)a->ldv[s], (cuuint64_t)a->ldv[s] * nk};
        if ((rc = encode_bf16(s == 0 ? &tk0 : &tk1, a->k[s], 3, dims, sk, box))) return rc;
        if ((rc = encode_bf16(s == 0 ? &tv0 : &tv1, a->v[s], 3, dims, sv, box))) return rc;
    }
    if ((rc = configure<false>(kCrossSmem))) return rc;
    AttnTcParams p = {};
    p.mode_tiles = (a->nq + 127) / 128;
    p.heads = a->heads; p.batch = a->batch;
    p.total_tiles = a->batch * a->heads * p.mode_tiles;
    p.nk0 = a->nk[0]; p.nk1 = a->nk[1];
    p.kv_div0 = a->kv_batch_div[0]; p.kv_div1 = a->kv_batch_div[1];
    p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
    p.seg0 = a->seg_scale[0]; p.seg1 = a->seg_scale[1];
    const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    launch_pdl(kPdlAttention, attn_tc_kernel<false>, dim3(grid), dim3(kTThreads), kCrossSmem, stream, tq, tk0, tv0, tk1, tv1, to, p);
    return check_launch("cross_tc");
}

int temporal_tc_launch(const void* q, const void* k, const void* v, void* o, long long ldq, long long ldk, long long ldv,
                       long long ldo, int B, int L, int HW, int heads, float softmax_scale, cudaStream_t stream) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                         reinterpret_cast<uintptr_t>(o);
    if ((al & 15) || (ldq & 7) || (ldk & 7) || (ldv & 7) || (ldo & 7) || L < 1 || L > 16) return DCB_ERR_UNSUPPORTED;
    const cuuint64_t cols = static_cast<cuuint64_t>(heads) * 64;
    CUtensorMap tq, tk, tv, to;
    const void* ptrs[4] = {q, k, v, o};
    const long long lds[4] = {ldq, ldk, ldv, ldo};
    CUtensorMap* maps[4] = {&tq, &tk, &tv, &to};
    for (int i = 0; i < 4; ++i) {
        const cuuint64_t ld = lds[i];
        const cuuint64_t dims[4] = {cols, (cuuint64_t)HW, (cuuint64_t)L, (cuuint64_t)B};
        const cuuint64_t str[4] = {1, ld, ld * HW, ld * HW * L};
        const cuuint32_t box[4] = {64, 1, 16, 1};
        int rc = encode_bf16(maps[i], ptrs[i], 4, dims, str, box);
        if (rc) return rc;
    }
    int rc;
    if ((rc = configure<true>(kTempSmem))) return rc;
    AttnTcParams p = {};
    p.mode_tiles = (HW + 7) / 8;
    p.heads = heads; p.batch = B;
    const long long total = static_cast<long long>(B) * heads * p.mode_tiles;
    if (total > 0x7fffffffLL) return DCB_ERR_UNSUPPORTED;
    p.total_tiles = static_cast<int>(total);
    p.L = L;
    p.kv_div0 = p.kv_div1 = 1;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    p.seg0 = p.seg1 = 1.f;
    const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    launch_pdl(kPdlAttention, attn_tc_kernel<true>, dim3(grid), dim3(kTThreads), kTempSmem, stream, tq, tk, tv, tk, tv, to, p);
    return check_launch("temporal_tc");
}

}  // namespace dcb
