// tcgen05 / TMEM / TMA flash attention, ping-pong variant, for the spatial self-attention of the UNet
// (attention.py:101-125, head_dim 64).  EXPERIMENTAL (DCB_FLASH_PP=1; flash_tc.cu stays the default): parity green on all
// flash tests (2560 / 640 / 160 / 9216 / 2304 keys, ragged tiles), 536 TFLOP/s at 2560 keys and 588 at 9216 against
// flash_tc's 600 / 665.  What the variants measured (tools/trace_flash_pp.py, profiles/r02_flash_pp_timeline.txt,
// profiles/r02_summary.md):
//   * one MMA-issuing warp for both sets paces the CTA: a 128 x 64 x 16 tcgen05.mma takes ~90 clk to issue (32 to
//     execute), 16 per 64-key step -> ~2100 clk; one issuer per set halves that chain (kept); separate S / P V issuers
//     per set do not help further;
//   * a 64-key step of both sets then still takes ~2000 clk with one OR two threads per query row (8 or 16 softmax
//     warps; one read of every score or two), with the two groups started together or 400..1400 clk apart: the exponential
//     phases (MUFU: 1024 clk per step), the TMEM reads (64 KB per step at ~64 B/clk) and the shared-memory traffic of the
//     small MMAs (6 KB per 32-clk instruction) are each about half of that, none alone explains it.
// Next: per-pipe utilisation from an ncu --set full capture of this kernel, and P kept in TMEM (tcgen05.mma with A from
// TMEM) to take the P stores and half of the operand reads off shared memory.
// Design:
//
//   * ONE CTA per SM works on TWO 128-query tiles ("sets" A and B) of one (batch, head) and walks over the key tiles
//     once for both: every K_j / V_j tile is fetched once per 256 queries (half the L2 traffic of flash_tc);
//   * two softmax groups (4 warps each, ONE thread per query row) ping-pong: group A exponentiates tile j of set A while
//     the tensor core computes S_j of set B and P V of the previous step -- the MUFU-bound exponential phase of one group
//     runs under the TMEM round trips, row maxima and MMA waits of the other;
//   * keys are processed in 64-key sub-tiles with S double buffered per set (the next scores are in TMEM before the
//     softmax group asks for them) and P double buffered; a softmax thread reads its 64 scores from TMEM ONCE into
//     registers (flash_tc reads them twice and both threads of a row scan all 128 for the max);
//   * O is accumulated IN TMEM by the P V MMAs (use_acc), not read back every tile; the running maximum is refreshed
//     lazily (only when it grew by more than 2^8), and only then O is scaled in place (tcgen05.ld / st).
//
//   warp 0      TMA producer: Q_A, Q_B once, then {K_j, V_j} into a 3-stage ring (SWIZZLE_128B)
//   warp 1      TMEM owner + MMA issuer, order  S_0^A S_0^B S_1^A S_1^B | P V_t^A, S_{t+2}^A | P V_t^B, S_{t+2}^B | ...
//   warps 2..5  softmax group A, warps 6..9 group B (one thread per query row); warp 10: MMA issuer of set B
// TMEM (512 columns): S_A 0..63 | 64..127, O_A 128..191, S_B 256..319 | 320..383, O_B 384..447.
#include "dcb_internal.h"
#include "ptx.cuh"

namespace dcb {

constexpr int kPTile = 128 * 128;              // 128 rows x 64 bf16 (SWIZZLE_128B)
constexpr int kPThreads = 352;                 // TMA warp, MMA issuer A, 4 + 4 softmax warps, MMA issuer B
constexpr int kPKV = 3;                        // {K, V} ring depth
constexpr int kPSmem = 2 * kPTile + kPKV * 2 * kPTile + 2 * 2 * kPTile + 256;     // Q x2, {K,V} ring, P x2 (2 halves each)

// -DDCB_TRACE: CTA (0,0,0) logs (clock, event, tile) per role: 0 mma warp, 1 softmax group A (warp 2), 2 group B (warp 6)
#ifdef DCB_TRACE
static __device__ unsigned long long g_pp_trace[3 * 4096];
struct PpTrace {
    unsigned n = 1;
    __device__ __forceinline__ void ev(int role, unsigned code, unsigned arg) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && n < 4096u)
            g_pp_trace[role * 4096 + n++] = (static_cast<unsigned long long>(clock64()) << 20) | (code << 12) | (arg & 0xfffu);
    }
    __device__ __forceinline__ void done(int role) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_pp_trace[role * 4096] = n;
    }
};
#define PTR_DECL PpTrace ptr_;
#define PTR(role, code, arg) ptr_.ev(role, code, arg)
#define PTR_DONE(role) ptr_.done(role)
#else
#define PTR_DECL
#define PTR(role, code, arg)
#define PTR_DONE(role)
#endif

struct FlashPpParams {
    int nq, nk, ntiles;                        // queries = keys per (batch, head); 64-key sub-tiles
    float scale_log2;
    int stagger_clk;                           // start delay of softmax group B (DCB_FLASH_PP_STAGGER, clocks)
};

__device__ __forceinline__ uint64_t pp_desc_mn(uint32_t smem_addr) {         // MN-major B operand (V), 8-row groups 1024 B apart
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ void pp_ld3(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void pp_st3(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void pp_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
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
__device__ __forceinline__ void pp_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float pp_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pp_pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kPThreads, 1)
flash_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FlashPpParams p) {
    extern __shared__ __align__(1024) uint8_t psm[];
    uint8_t* sQ = psm;                                   // [2 sets]
    uint8_t* sKV = psm + 2 * kPTile;                     // [kPKV][K 128 keys, V 128 keys]
    uint8_t* sP = sKV + kPKV * 2 * kPTile;               // [2 sets][2 buffers] of 128 rows x 64 keys
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kPTile);
    uint64_t* q_full = bars;                             // [1]
    uint64_t* kv_full = bars + 1;                        // [kPKV]
    uint64_t* kv_empty = bars + 1 + kPKV;                // [kPKV]
    uint64_t* s_full = bars + 1 + 2 * kPKV;              // [set][buffer]
    uint64_t* p_full = s_full + 4;                       // [set][buffer]
    uint64_t* o_full = p_full + 4;                       // [set][buffer]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 256, head = blockIdx.y, batch = blockIdx.z;
    const int T2 = p.ntiles;                             // 64-key sub-tiles
    const int TS = (T2 + 1) >> 1;                        // 128-key K / V stages
    const int nsets = (q0 + 128 < p.nq) ? 2 : 1;         // the second query tile may not exist (odd tile count)
    PTR_DECL

    if ((smem_u32(psm) & 1023) != 0) __trap();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
        mbar_init(q_full, 1);
        for (int i = 0; i < kPKV; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], nsets); }
        for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); }
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
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, nsets * kPTile);
            pp_ld3(sQ, &tmQ, q_full, head * 64, q0, batch);
            if (nsets == 2) pp_ld3(sQ + kPTile, &tmQ, q_full, head * 64, q0 + 128, batch);
        }
        __syncwarp();
        for (int j = 0; j < TS; ++j) {
            const int st = j % kPKV;
            mbar_wait(&kv_empty[st], ((j / kPKV) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&kv_full[st], 2 * kPTile);
                pp_ld3(sKV + st * 2 * kPTile, &tmK, &kv_full[st], head * 64, j * 128, batch);
                pp_ld3(sKV + st * 2 * kPTile + kPTile, &tmV, &kv_full[st], head * 64, j * 128, batch);
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp == 10) {
        // ------------------------------------------------------------------ MMA issuers: warp 1 for set A, warp 10 for set B.
        // A tcgen05.mma of this size (128 x 64 x 16) takes ~90 clk to ISSUE and 32 clk to execute, so the issue rate of a
        // warp paces its set (measured: one issuer for both sets ~2100 clk per 64-key step, one per set ~1150; splitting
        // S and P V over two more warps did not help).  tcgen05.commit tracks the MMAs of its own thread.
        const int g = warp == 1 ? 0 : 1;
        constexpr uint32_t kIdescS = make_idesc_bf16(128, 64);
        constexpr uint32_t kIdescPV = make_idesc_bf16(128, 64) | (1u << 16);       // B (= V) is MN-major
        // every descriptor is a constant base plus a small offset in its address field (>> 4 units; all tiles lie below 256 KB)
        const uint64_t dQ0 = make_sw128_kmajor_desc(smem_u32(sQ));                 // + g * (kPTile >> 4)
        const uint64_t dK0 = make_sw128_kmajor_desc(smem_u32(sKV));                // + st * (2 kPTile >> 4) + b * (8192 >> 4)
        const uint64_t dV0 = pp_desc_mn(smem_u32(sKV + kPTile));                   // + st * (2 kPTile >> 4) + (4 b + kk) * (2048 >> 4)
        const uint64_t dP0 = make_sw128_kmajor_desc(smem_u32(sP));                 // + (2 g + b) * (kPTile >> 4) + 2 kk
        // S_t of set g: Q_g K_t^T (64 keys) -> TMEM columns g * 256 + (t & 1) * 64 .. + 63
        auto issue_S = [&](int t) {
            const int j = t >> 1, st = j % kPKV, b = t & 1;
            mbar_wait(&kv_full[st], (j / kPKV) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t ad = dQ0 + static_cast<uint64_t>(g * (kPTile >> 4));
                const uint64_t bd = dK0 + static_cast<uint64_t>(st * (2 * kPTile >> 4) + b * (8192 >> 4));
                const uint32_t d = tmem_base + g * 256 + b * 64;
                umma_bf16_ss(d, ad, bd, kIdescS, 0u);
                umma_bf16_ss(d, ad + 2, bd + 2, kIdescS, 1u);
                umma_bf16_ss(d, ad + 4, bd + 4, kIdescS, 1u);
                umma_bf16_ss(d, ad + 6, bd + 6, kIdescS, 1u);
                umma_commit(&s_full[g * 2 + b]);
            }
            __syncwarp();
        };
        if (g < nsets) {
            mbar_wait(q_full, 0);
            for (int t = 0; t < 2 && t < T2; ++t) issue_S(t);
            for (int t = 0; t < T2; ++t) {
                const int j = t >> 1, st = j % kPKV, b = t & 1;
                mbar_wait(&p_full[g * 2 + b], (t >> 1) & 1);      // P_t^g written, S_t^g read, O_g rescaled if needed
                tc_fence_after();
                if (lane == 0 && g == 0) PTR(0, 5 + g, t);
                if (elect_one()) {
                    const uint64_t ad = dP0 + static_cast<uint64_t>((g * 2 + b) * (kPTile >> 4));
                    const uint64_t bd = dV0 + static_cast<uint64_t>(st * (2 * kPTile >> 4) + b * 4 * (2048 >> 4));
                    const uint32_t d = tmem_base + g * 256 + 128;
                    umma_bf16_ss(d, ad, bd, kIdescPV, t != 0);
                    umma_bf16_ss(d, ad + 2, bd + (2048 >> 4), kIdescPV, 1u);
                    umma_bf16_ss(d, ad + 4, bd + 2 * (2048 >> 4), kIdescPV, 1u);
                    umma_bf16_ss(d, ad + 6, bd + 3 * (2048 >> 4), kIdescPV, 1u);
                    umma_commit(&o_full[g * 2 + b]);
                    if (b == 1 || t == T2 - 1) umma_commit(&kv_empty[st]);        // this set is done with the stage (count = sets)
                }
                __syncwarp();
                if (t + 2 < T2) issue_S(t + 2);                  // into the S buffer the softmax group has just drained
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax groups: ONE thread per query row.  Every score is read from TMEM exactly
        // once (TMEM reads run at ~64 B/clk per SM: 1024 clk for the two 128 x 64 fp32 tiles of a step, the same as the MUFU
        // time of their exponentials -- reading scores twice, or O back every tile as flash_tc does, is what bounds that kernel)
        const int q = warp & 3;                         // TMEM lane quarter
        const int g = (warp - 2) >> 2;                  // group = set
        const int r = q * 32 + lane;
        if (g < nsets) {
            const uint32_t t_set = tmem_base + g * 256 + (static_cast<uint32_t>(q * 32) << 16);
            const uint32_t t_o = t_set + 128;
            const float sl2 = p.scale_log2;
            float m_run = -INFINITY, l_run = 0.f;
            if (g == 1 && p.stagger_clk > 0) {
                // ping-pong needs the two groups in ANTI-phase (one exponentiating while the other reads TMEM / waits):
                // identical groups started together stay in lockstep, so group B starts half a step late
                const long long t0 = clock64();
                while (clock64() - t0 < p.stagger_clk) __nanosleep(32);
            }
            for (int t = 0; t < T2; ++t) {
                const int b = t & 1;
                const int kmax = p.nk - t * 64;             // valid keys in this sub-tile (>= 64 unless last)
                if (q == 2 && lane == 0) PTR(1 + g, 0, t);
                mbar_wait(&s_full[g * 2 + b], (t >> 1) & 1);
                tc_fence_after();
                if (q == 2 && lane == 0) PTR(1 + g, 1, t);
                uint32_t v[64];
                {
                    uint32_t a0[32], a1[32];
                    tmem_ld32(t_set + b * 64, a0);
                    tmem_ld32(t_set + b * 64 + 32, a1);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 32; ++k) { v[k] = a0[k]; v[32 + k] = a1[k]; }
                }
                if (q == 2 && lane == 0) PTR(1 + g, 2, t);
                float mx = -INFINITY;
                if (kmax >= 64) {
                    // four independent chains (a single dependent chain of 32 max3 ops costs ~200 clk of pure latency)
                    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int k = 0; k < 64; k += 8) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            m4[u] = fmaxf(m4[u], fmaxf(__uint_as_float(v[k + 2 * u]), __uint_as_float(v[k + 2 * u + 1])));
                    }
                    mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                } else {
#pragma unroll
                    for (int k = 0; k < 64; ++k)
                        if (k < kmax) mx = fmaxf(mx, __uint_as_float(v[k]));
                }
                // P buffer b was read by P V_{t-2}: that MMA must have retired before the buffer is rewritten
                if (t >= 2) mbar_wait(&o_full[g * 2 + b], ((t >> 1) - 1) & 1);
                // lazy rescale: keep the stale reference until the true max is more than 2^8 above it
                const bool upd = (mx - m_run) * sl2 > 8.0f;                    // first tile: m_run = -inf -> true
                if (__any_sync(0xffffffffu, upd)) {
                    const float m_new = upd ? mx : m_run;
                    const float alpha = (t > 0) ? pp_ex2((m_run - m_new) * sl2) : 0.f;   // 1 for rows that keep their max
                    m_run = m_new;
                    if (t > 0) {
                        // every P V up to t-1 must have retired before O is scaled in place
                        mbar_wait(&o_full[g * 2 + (b ^ 1)], ((t - 1) >> 1) & 1);
                        tc_fence_after();
                        l_run *= alpha;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t o[32];
                            tmem_ld32(t_o + c * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            pp_tmem_st32(t_o + c * 32, o);
                        }
                        pp_tmem_st_wait();
                    }
                }
                if (q == 2 && lane == 0) PTR(1 + g, 3, t);
                const float msc = m_run * sl2;
                float rs = 0.f;
                uint8_t* const prow = sP + (g * 2 + b) * kPTile + r * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {               // 8 keys = one 16-byte chunk of the bf16 row
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        e[i] = pp_ex2(fmaf(__uint_as_float(v[c * 8 + i]), sl2, -msc));
                        if (kmax < 64 && c * 8 + i >= kmax) e[i] = 0.f;
                    }
                    rs += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
                    uint4 pk;
                    pk.x = pp_pack2(e[0], e[1]); pk.y = pp_pack2(e[2], e[3]);
                    pk.z = pp_pack2(e[4], e[5]); pk.w = pp_pack2(e[6], e[7]);
                    *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = pk;
                }
                l_run += rs;
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[g * 2 + b]);
                if (q == 2 && lane == 0) PTR(1 + g, 4, t);
            }
            // ---- O / l -> bf16 -> staging (this set's Q tile: every S MMA has retired) -> TMA store
            if (T2 >= 2) mbar_wait(&o_full[g * 2 + ((T2 - 2) & 1)], ((T2 - 2) >> 1) & 1);
            mbar_wait(&o_full[g * 2 + ((T2 - 1) & 1)], ((T2 - 1) >> 1) & 1);
            tc_fence_after();
            const float inv = 1.f / l_run;
            uint8_t* orow = sQ + g * kPTile + r * 128;
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t o[32];
                tmem_ld32(t_o + c2 * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 pk;
                    pk.x = pp_pack2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
                    pk.y = pp_pack2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
                    pk.z = pp_pack2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
                    pk.w = pp_pack2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + (((c2 * 4 + c) ^ (r & 7)) << 4)) = pk;
                }
            }
            fence_proxy_async();
            named_bar_sync(1 + g, 128);
            if (q == 0 && lane == 0) {
                pp_st3(&tmO, sQ + g * kPTile, head * 64, q0 + g * 128, batch);
                tma_store_commit();
                tma_store_wait_all<0>();
            }
        }
    }
    if (warp == 1 && lane == 0) PTR_DONE(0);
    if ((warp == 2 || warp == 6) && lane == 0) PTR_DONE(1 + ((warp - 2) >> 2));
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*EncodeTiledFnP)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int pp_encode3(CUtensorMap* m, const void* base, long long ld, int cols, int n, int batch) {
    static EncodeTiledFnP fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return set_error(DCB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
        fn = reinterpret_cast<EncodeTiledFnP>(ptr);
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)n, (cuuint64_t)batch};
    const cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)n};
    const cuuint32_t box[3] = {64, 128, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return set_error(DCB_ERR_CUDA, "flash_pp: cuTensorMapEncodeTiled failed (rc=%d)", (int)rc);
    return DCB_OK;
}

// Single-segment self-attention with nq == nk; DCB_ERR_UNSUPPORTED if the arguments do not fit (the caller falls back).
int flash_pp_launch(const dcb_attention_args* a, cudaStream_t stream) {
    if (a->nseg != 1 || a->nk[0] != a->nq || a->kv_batch_div[0] != 1 || a->seg_scale[0] != 1.0f || a->head_dim != 64)
        return DCB_ERR_UNSUPPORTED;
    const uintptr_t al = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k[0]) |
                         reinterpret_cast<uintptr_t>(a->v[0]) | reinterpret_cast<uintptr_t>(a->o);
    if ((al & 15) || (a->ldq & 7) || (a->ldk[0] & 7) || (a->ldv[0] & 7) || (a->ldo & 7)) return DCB_ERR_UNSUPPORTED;
    const int cols = a->heads * 64;
    CUtensorMap tq, tk, tv, to;
    int rc;
    if ((rc = pp_encode3(&tq, a->q, a->ldq, cols, a->nq, a->batch))) return rc;
    if ((rc = pp_encode3(&tk, a->k[0], a->ldk[0], cols, a->nq, a->batch))) return rc;
    if ((rc = pp_encode3(&tv, a->v[0], a->ldv[0], cols, a->nq, a->batch))) return rc;
    if ((rc = pp_encode3(&to, a->o, a->ldo, cols, a->nq, a->batch))) return rc;
    static bool cfg[kMaxDevices] = {};
    const int dv = current_device() & (kMaxDevices - 1);
    if (!cfg[dv]) {
        cudaError_t e = cudaFuncSetAttribute(flash_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem);
        if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "flash_pp: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        cfg[dv] = true;
    }
    FlashPpParams p;
    p.nq = a->nq;
    p.nk = a->nq;
    p.ntiles = (a->nq + 63) / 64;                  // 64-key sub-tiles
    p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
    static int stagger = -1;
    if (stagger < 0) { const char* e = getenv("DCB_FLASH_PP_STAGGER"); stagger = e ? atoi(e) : 700; }
    p.stagger_clk = stagger;
    dim3 grid((a->nq + 255) / 256, a->heads, a->batch);
    launch_pdl(kPdlAttention, flash_pp_kernel, grid, dim3(kPThreads), kPSmem, stream, tq, tk, tv, to, p);
    return check_launch("flash_pp");
}

}  // namespace dcb

#ifdef DCB_TRACE
extern "C" int dcb_debug_pp_trace(unsigned long long* dst, int cap) {
    cudaDeviceSynchronize();
    if (cap < 3 * 4096) return -1;
    cudaMemcpyFromSymbol(dst, dcb::g_pp_trace, 3 * 4096 * sizeof(unsigned long long));
    return 3 * 4096;
}
#endif
