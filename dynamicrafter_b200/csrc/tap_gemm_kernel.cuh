// dcb_tap_gemm: persistent, warp-specialised tcgen05 implicit GEMM for sm_100a.
//
//   out[r, n] = epilogue( sum_{tap} sum_{c} A[r shifted by tap, c] * W[n, tap*C + c] )
//
// One kernel serves nn.Linear, 1x1 conv, 3x3 conv (9 taps over y/x) and the (3,1,1) temporal conv
// (3 taps over frames): the A operand is fetched by TMA as a box of a 4-D channels-last tensor and a
// tap is just a coordinate shift of that box; TMA's out-of-bounds zero fill is the conv padding.
//
// Two launch modes of the same template (k2):
//   single   one persistent CTA per SM, 320 threads, 128 x BN tile, cta_group::1.
//   pair     clusters of 2 CTAs (cta_group::2), 576 threads each, one 256 x BN tile per pair: every CTA stages its own
//            128 A rows and HALF of the B tile, the leader CTA issues the MMAs for both and multicasts the commit, so
//            shared-memory operand traffic per FLOP halves (a single CTA at BN=128 is smem-bandwidth bound at ~50 %
//            of the tensor pipe).  Persistent over tiles, N fastest so an A tile is re-read from L2.
// Warp roles:
//   warp 0      TMA producer: {A 128x64, B rows x 64} bf16 stages, SWIZZLE_128B, full/empty mbarrier ring
//   warp 1      TMEM owner + tcgen05.mma issuer (the whole warp runs the loop, elect.sync picks the issuing lane so
//               the loop stays warp-uniform), 2-deep TMEM accumulator ring
//   warps 2..   epilogue, 2 (single) or 4 (pair) groups of 4 warps; a group covers the 128 TMEM lanes and takes
//               every EpiGroups-th 16-column step: tcgen05.ld -> fused bias / per-video embedding / GEGLU /
//               residual -> swizzled smem slot (64-byte rows) -> TMA tile store.  The fp32 residual chunk is
//               PREFETCHED by TMA into the same slot (in place), so the epilogue issues no per-thread global loads
//               or stores and all HBM traffic of the kernel is asynchronous and fully coalesced.  Out-of-range rows
//               and columns (ragged tiles, N < 32) are clipped by the TMA unit.
//   kCS         optional fused GroupNorm statistics (per-tile column sums of the stored values); separate TU.
#pragma once
#include <cstdlib>

#include "dcb_internal.h"
#include "ptx.cuh"

namespace dcb {

// Division by a launch-constant divisor as multiply-high + shift (valid for dividends < 2^31): the tile loops of all
// 18 warps decompose tile / row indices every tile, and a hardware-less 32-bit division costs ~25 issue slots each.
struct FastDiv {
    uint32_t mul, shr, d;
    __host__ static FastDiv make(uint32_t d) {
        FastDiv f;
        f.d = d;
        if (d <= 1) { f.mul = 0; f.shr = 0; f.d = 1; return f; }
        uint32_t lg = 0;
        while ((1u << lg) < d) ++lg;
        const uint32_t p = 31 + lg;
        f.mul = static_cast<uint32_t>(((1ull << p) + d - 1) / d);
        f.shr = p - 32;
        return f;
    }
    __device__ __forceinline__ int div(int n) const {
        return d == 1 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul) >> shr);
    }
    __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
        q = div(n);
        r = n - q * static_cast<int>(d);
    }
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
// epilogue groups of 4 warps (one warp per TMEM lane quarter): 4 groups in CTA-pair mode (576 threads), else 2
template <bool k2> struct EpiGroups { static constexpr int value = k2 ? 4 : 2; };
template <bool k2> struct Threads { static constexpr int value = 64 + 128 * EpiGroups<k2>::value; };
constexpr int kABytes = kBlockM * kBlockK * 2;
// epilogue chunk = 64-byte rows (SWIZZLE_64B): 16 fp32 or 32 bf16 output columns x 128 rows = 8 KB per slot
constexpr int kSlotBytes = kBlockM * 64;
// slots per epilogue group: depth of the TMA store / residual-prefetch ring (pairs are the HBM-heavy big launches)
template <bool k2> struct EpiSlots { static constexpr int value = 2; };

// smem ring depth: {A 16 KB + B rows x 128 B} per stage, 32 KB of epilogue slots, <= 227 KB in total.
// In CTA-pair mode (k2) each CTA stages only half of the B tile.
template <int BN, bool k2> struct StageCount {
    static constexpr int stage_bytes = kABytes + (k2 ? BN / 2 : BN) * kBlockK * 2;
    static constexpr int epi_bytes = EpiGroups<k2>::value * EpiSlots<k2>::value * kSlotBytes;
    static constexpr int fit = (232448 - epi_bytes - 1024 - 256) / stage_bytes;
    static constexpr int value = fit > 8 ? 8 : fit;
};

// -DDCB_TRACE (DCB_NVCC_EXTRA=-DDCB_TRACE, tools/trace_gemm.py): six role lanes of CTA 0 (producer, mma, epilogue
// leaders 0..3) log (clock, event, argument) into their own 4096-entry region with plain stores; entry 0 = count.
#ifdef DCB_TRACE
static __device__ unsigned long long g_tap_trace[6 * 4096];
struct TraceLane {
    unsigned n = 1;
    __device__ __forceinline__ void ev(int role, unsigned code, unsigned arg) {
        if (blockIdx.x == 0 && n < 4096u)
            g_tap_trace[role * 4096 + n++] = (static_cast<unsigned long long>(clock64()) << 20) | (code << 12) | (arg & 0xfffu);
    }
    __device__ __forceinline__ void done(int role) { if (blockIdx.x == 0) g_tap_trace[role * 4096] = n; }
};
#define DCB_TR_DECL TraceLane trl;
#define DCB_TR_DONE(role) trl.done(role)
#define DCB_TR(role, code, arg) trl.ev(role, code, arg)
#else
#define DCB_TR_DECL
#define DCB_TR_DONE(role)
#define DCB_TR(role, code, arg)
#endif

struct TapGemmParams {
    int D1, D2, D3, b1, b2, b3, t1, t2, t3;
    int num_m_tiles, num_n_tiles, rows_per_tile;
    int kchunks, ntaps;
    int taps[DCB_MAX_TAPS][3];
    int N, N_out;
    const float* bias;
    const float* group_bias;
    int rows_per_group, ld_group_bias;
    int has_residual, geglu, out_bf16;
    const float* residual;
    int ld_residual;
    float out_scale;
    uint32_t stage_bytes, res_bytes;
    unsigned long long* gn_acc;    // fused GroupNorm statistics: int64 fixed-point [units][32][2] accumulator, or null
    int gn_cg, gn_unit_mode, seg_rows, seg_lw;
    // LayerNorm folding (see dcb200.h): producer = fp32-out kernel that also writes a bf16 copy + per-row (sum, sumsq);
    // consumer = bf16-out kernel whose epilogue applies rstd * acc - rstd * mean * colsum[n] + bias[n]
    unsigned long long* ln_acc;
    __nv_bfloat16* ln_x;
    int ld_ln_x;
    const float* ln_colsum;
    float ln_inv_c, ln_eps;
    // ragged last N tile: the UMMA runs at N = n_eff_last (the valid columns rounded up to 32) and each CTA stages only
    // b_rows_last rows of B (second tensor map), so a partial tile costs tensor time and fill bandwidth for what it holds
    uint32_t idesc_last, stage_bytes_last;
    int b_rows_last;
    // work order.  sched_wave = 0 (default): N fastest, tile = m item * NT + n tile -- neighbouring workers share an A
    // tile at the same time.  sched_wave = 1 (opt-in, DCB_GEMM_ORDER=1): groups of `sched_w` consecutive M items, N tile
    // major inside a group, so one worker walks over all N tiles of ONE M item and every worker of a wave runs the same
    // N tile (equal cost per wave with a narrow last tile); measured slower, kept for experiments
    int sched_w, sched_full_groups, sched_wave;
    FastDiv fd_gsz, fd_w, fd_r, fd_ntiles;
    FastDiv fd_t1, fd_t2, fd_b1, fd_b2, fd_rpg, fd_cg;
};

// Exact (erf-based) GELU, attention.py:422.  erfc(z) via Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, i.e. fp32
// round-off level), branch-free.  With a = |x|, z = a / sqrt2 and hp = 0.5 a erfc(z) >= 0:
//   x >= 0: gelu = x - 0.5 x erfc(z) = x - hp        x < 0: gelu = 0.5 x erfc(z) = -hp        => gelu = max(x, 0) - hp
// (no cancellation, no select).  15 instructions, two of them MUFU: the epilogue of the K=320 GEGLU GEMMs is
// issue-slot bound (ncu: 70 % issue utilisation, 36 instructions per output before this form), not tensor bound.
__device__ __forceinline__ float gelu_erf(float x) {
    const float a = fabsf(x);
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164190f, a, 1.0f)));     // 1 / (1 + p z), argument in [1, inf)
    float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);                    // 0.5 folded into the coefficients
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, 0.5f * -0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    poly *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * x) * -0.72134752044f));       // exp(-z^2) = 2^(-x^2 log2(e) / 2)
    return fmaxf(x, 0.f) - a * (e * poly);
}

// Fused GroupNorm statistics: (sum, sum of squares) of one staged epilogue chunk (128 rows x 64 bytes in the
// SWIZZLE_64B slot, fp32 x16 or bf16 x32 columns) over the 32 rows of the calling warp, read back from shared memory
// AFTER the group barrier (no shuffles on the accumulator path: 16 conflict-free 4-byte loads per lane).  Half-warp h
// takes rows 2i + h, lane w the 4-byte word w of the row; the halves are folded with one shuffle per value, the columns
// of one GroupNorm group (consecutive lanes) with a segmented shuffle reduction, and the first lane of every group adds
// the result to the per-(unit, group) int64 fixed-point accumulator (integer atomics commute: bitwise reproducible).
// Rows that share a unit: kSeg = 32 (whole warp), 16 or 8 consecutive rows -> 1, 2 or 4 units per warp.
constexpr double kGnFixScale = 1048576.0;          // 2^20
constexpr float kLnFixSum = 16777216.0f;           // 2^24: per-row sum
constexpr float kLnFixSq = 268435456.0f;           // 2^28: per-row sum of squares
template <bool kF32, int kSeg>
__device__ __forceinline__ void gn_acc_from_slot(const uint8_t* slot, int q, int lane, unsigned long long* __restrict__ acc,
                                                 const long long* unit, int col0, int n_out, const FastDiv cg) {
    constexpr int NSEG = 32 / kSeg;                   // units covered by this warp
    constexpr int IPS = 16 / NSEG;                    // loop iterations (row pairs) per segment
    constexpr int NV = kF32 ? 1 : 2;                  // columns per lane
    const int h = lane >> 4, w = lane & 15;
    const uint8_t* base = slot + (q * 32 + h) * 64 + ((w & 3) << 2);
    float sx[NSEG], sq[NSEG];
#pragma unroll
    for (int sg = 0; sg < NSEG; ++sg) { sx[sg] = 0.f; sq[sg] = 0.f; }
    const int col = col0 + w * NV;
    // a bf16 lane holds two adjacent columns: they are in the same group unless the boundary falls between them
    int gid0, pos0;
    cg.divmod(col, gid0, pos0);                       // group of my (first) column, position inside the group
    const bool split = !kF32 && (pos0 + 1 == static_cast<int>(cg.d));
    float sx2[NSEG], sq2[NSEG];                       // second column when it starts a new group
#pragma unroll
    for (int sg = 0; sg < NSEG; ++sg) { sx2[sg] = 0.f; sq2[sg] = 0.f; }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        // row q*32 + 2i + h: SWIZZLE_64B chunk index ^= (row / 2) % 4 = i % 4
        const uint32_t u = *reinterpret_cast<const uint32_t*>(base + i * 128 + ((((w >> 2) ^ (i & 3))) << 4));
        const int sg = i / IPS;
        if constexpr (kF32) {
            const float v = __uint_as_float(u);
            sx[sg] += v;
            sq[sg] = fmaf(v, v, sq[sg]);
        } else {
            const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
            sx[sg] += v.x;
            sq[sg] = fmaf(v.x, v.x, sq[sg]);
            if (split) { sx2[sg] += v.y; sq2[sg] = fmaf(v.y, v.y, sq2[sg]); }
            else { sx[sg] += v.y; sq[sg] = fmaf(v.y, v.y, sq[sg]); }
        }
    }
    // group of my (first) column; columns past n_out (ragged N tiles) contribute nothing and join no group
    const bool in0 = col < n_out, in1 = !kF32 && (col + 1 < n_out);
    const int gid = in0 ? gid0 : -1 - w;
#pragma unroll
    for (int sg = 0; sg < NSEG; ++sg) {
        float a = sx[sg] + __shfl_xor_sync(0xffffffffu, sx[sg], 16);
        float b = sq[sg] + __shfl_xor_sync(0xffffffffu, sq[sg], 16);
        float a2 = 0.f, b2 = 0.f;
        if constexpr (!kF32) {
            a2 = sx2[sg] + __shfl_xor_sync(0xffffffffu, sx2[sg], 16);
            b2 = sq2[sg] + __shfl_xor_sync(0xffffffffu, sq2[sg], 16);
            if (!in1 && !split) { /* second column out of range: its (zero-filled) values were added to a: harmless */ }
        }
        // segmented reduction over the 16 column lanes: lane w + off is in my group iff it stays inside the group's columns
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            const float ta = __shfl_down_sync(0xffffffffu, a, off, 16);
            const float tb = __shfl_down_sync(0xffffffffu, b, off, 16);
            if (w + off < 16 && pos0 + off * NV < static_cast<int>(cg.d) && col + off * NV < n_out) { a += ta; b += tb; }
        }
        const bool head = in0 && (w == 0 || pos0 < NV);       // first lane of a group inside this chunk
        if (h == 0 && unit[sg] >= 0) {
            if (head) {
                unsigned long long* d = acc + (unit[sg] * 32 + gid) * 2;
                atomicAdd(d, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(a) * kGnFixScale)));
                atomicAdd(d + 1, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(b) * kGnFixScale)));
            }
            if (!kF32 && split && in1) {
                unsigned long long* d = acc + (unit[sg] * 32 + gid0 + 1) * 2;
                atomicAdd(d, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(a2) * kGnFixScale)));
                atomicAdd(d + 1, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(b2) * kGnFixScale)));
            }
        }
    }
}

// kMode: 0 plain, 1 fused GroupNorm statistics of the output, 2 LayerNorm folding (producer if kF32, else consumer)
template <int BN, bool kF32, bool k2, int kMode>
__global__ void __launch_bounds__(Threads<k2>::value, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut,
                const __grid_constant__ CUtensorMap tmRes, const TapGemmParams p) {
    constexpr bool kCS = kMode == 1;
    constexpr bool kLNP = kMode == 2 && kF32;             // LayerNorm producer: bf16 copy + per-row statistics
    constexpr bool kLNC = kMode == 2 && !kF32;            // LayerNorm consumer: per-row scale / shift in the epilogue
    constexpr int kStages = StageCount<BN, k2>::value;
    constexpr int kSlots = EpiSlots<k2>::value;
    constexpr int kG = EpiGroups<k2>::value;
    constexpr int kEpiWarps = 4 * kG;
    constexpr int kEpiBytes = kG * kSlots * kSlotBytes;
    constexpr int kBRows = k2 ? BN / 2 : BN;               // B rows staged by this CTA
    constexpr int kBBytes = kBRows * kBlockK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr int kAccStride = (BN <= 128) ? 128 : 256;
    constexpr int kTmemCols = 2 * kAccStride;
    constexpr uint32_t kIdesc = make_idesc_bf16(k2 ? 2 * kBlockM : kBlockM, BN);
    // CTA pair: rank 0 (leader) issues tcgen05.mma.cta_group::2 for the 256 x BN tile; each CTA owns 128 of its rows
    const uint32_t rank = k2 ? cluster_ctarank() : 0u;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi = smem + kStages * kStageBytes;                      // 1024-aligned (stage sizes are multiples)
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kEpiBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tfull = bars + 2 * kStages;
    uint64_t* tempty = bars + 2 * kStages + 2;
    uint64_t* rfull = bars + 2 * kStages + 4;                         // [group][slot]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + kG * kSlots);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    DCB_TR_DECL
    // work item = (m tile [pair], n tile); workers = CTAs (or CTA pairs) striding over the items
    const int num_tiles = (k2 ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles) * p.num_n_tiles;
    const int worker = k2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int nworkers = k2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int kblocks = p.ntaps * p.kchunks;
    // tile -> (m tile, n tile) and m tile -> box origin (c1, c2, c3), all with multiply-high divisions
    auto split_tile = [&](int tile, int& m_tile, int& n_tile) {
        int mt;
        if (p.sched_wave) {
            int g, u, ml;
            p.fd_gsz.divmod(tile, g, u);
            if (g < p.sched_full_groups) p.fd_w.divmod(u, n_tile, ml); else p.fd_r.divmod(u, n_tile, ml);
            mt = g * p.sched_w + ml;
        } else {
            p.fd_ntiles.divmod(tile, mt, n_tile);
        }
        m_tile = k2 ? 2 * mt + static_cast<int>(rank) : mt;
    };
    auto origin = [&](int m_tile, int& c1, int& c2, int& c3) {
        int q1, i1, i2, i3;
        p.fd_t1.divmod(m_tile, q1, i1);
        p.fd_t2.divmod(q1, i3, i2);
        c1 = i1 * p.b1; c2 = i2 * p.b2; c3 = i3 * p.b3;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmB2);
        tma_prefetch_desc(&tmOut);
        if (p.has_residual) tma_prefetch_desc(&tmRes);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], k2 ? 2 : 1);              // pair: both producers arrive on the leader's barrier
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], k2 ? 2 * kEpiWarps : kEpiWarps);
        }
        for (int i = 0; i < kG * kSlots; ++i) mbar_init(&rfull[i], 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (k2) tmem_alloc_pair<kTmemCols>(tmem_slot); else tmem_alloc<kTmemCols>(tmem_slot);
    }
    tc_fence_before();
    if constexpr (k2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // prologue done: let the next kernel of the stream start its own prologue, then wait for our predecessor's data
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (whole warp runs the loop with
        // warp-uniform state; one elected lane issues, so operands stay in uniform registers)
        uint32_t it = 0;
        for (int tile = worker; tile < num_tiles; tile += nworkers) {
            int m_tile, n_tile, i1, i2, i3;
            split_tile(tile, m_tile, n_tile);
            origin(m_tile, i1, i2, i3);
            const bool last_n = n_tile == p.num_n_tiles - 1;
            const CUtensorMap* tb = last_n ? &tmB2 : &tmB;
            const uint32_t tx_bytes = last_n ? p.stage_bytes_last : p.stage_bytes;
            const int b_row0 = n_tile * BN + (k2 ? static_cast<int>(rank) * (last_n ? p.b_rows_last : kBRows) : 0);
            for (int tap = 0; tap < p.ntaps; ++tap) {
                const int c1 = i1 + p.taps[tap][0], c2 = i2 + p.taps[tap][1], c3 = i3 + p.taps[tap][2];
                for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (elect_one()) {
                        DCB_TR(0, 1, it);
                        uint8_t* sa = smem + s * kStageBytes;
                        if constexpr (k2) {
                            const uint32_t lfull = mapa_cluster(smem_u32(&full[s]), 0);      // leader's barrier
                            mbar_arrive_expect_tx_cluster(lfull, tx_bytes);
                            tma_load_4d_pair(sa, &tmA, lfull, kc * kBlockK, c1, c2, c3);
                            tma_load_2d_pair(sa + kABytes, tb, lfull, (tap * p.kchunks + kc) * kBlockK, b_row0);
                        } else {
                            mbar_arrive_expect_tx(&full[s], tx_bytes);
                            tma_load_4d(sa, &tmA, &full[s], kc * kBlockK, c1, c2, c3);
                            tma_load_2d(sa + kABytes, tb, &full[s], (tap * p.kchunks + kc) * kBlockK, b_row0);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA; one elected lane issues)
        if (rank == 0) {
            uint32_t it = 0, tc = 0;
            for (int tile = worker; tile < num_tiles; tile += nworkers, ++tc) {
                const uint32_t a = tc & 1, aph = (tc >> 1) & 1;
                mbar_wait(&tempty[a], aph ^ 1);
                tc_fence_after();
                if (lane == 0) DCB_TR(1, 2, tc);
                const uint32_t d_tmem = tmem_base + a * kAccStride;
                int m_tile_, n_tile_;
                split_tile(tile, m_tile_, n_tile_);
                const uint32_t idesc = (n_tile_ == p.num_n_tiles - 1) ? p.idesc_last : kIdesc;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        DCB_TR(1, 3, it);
                        const uint32_t sa = smem_u32(smem + s * kStageBytes);
                        const uint64_t adesc = make_sw128_kmajor_desc(sa);
                        const uint64_t bdesc = make_sw128_kmajor_desc(sa + kABytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) address field
                            if constexpr (k2) umma_bf16_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                            else umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        }
                        // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                        if constexpr (k2) umma_commit_pair(&empty[s]); else umma_commit(&empty[s]);
                        // accumulator complete -> epilogue (of both CTAs)
                        if (kb == kblocks - 1) {
                            if constexpr (k2) umma_commit_pair(&tfull[a]); else umma_commit(&tfull[a]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: 2 groups x 4 warps
        constexpr int CW = kF32 ? 16 : 32;         // output columns per chunk (64-byte smem rows)
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;           // chunk parity handled by this group
        const int r = q * 32 + lane;               // tile row owned by this thread
        const bool leader = (q == 0 && lane == 0);
        const int bar_id = 1 + grp;
        uint8_t* slots = epi + grp * kSlots * kSlotBytes;
        uint64_t* my_rfull = rfull + grp * kSlots;
        const int ncols = p.geglu ? BN / 2 : BN;   // output columns per tile
        const int nch = ncols / CW;
        // chunks of a tile that hold real columns (the last N tile may be narrow): this group takes grp, grp+kG, ...
        auto nch_of = [&](int n_tile) {
            const int vc = p.N_out - n_tile * ncols;
            return vc >= ncols ? nch : (vc + CW - 1) / CW;
        };
        const bool tma_res = kF32 && p.has_residual;          // fp32 out: residual chunk prefetched in place by TMA
        const int swz = (r >> 1) & 3;              // SWIZZLE_64B: 16-byte chunk index ^= (row / 2) % 4
        uint8_t* const row_base = nullptr;
        (void)row_base;
        uint32_t tc = 0, cnt = 0;

        // leader only: TMA-prefetch stream of residual chunks, kSlots - 2 work items ahead of the one being processed
        int pf_tile = worker, pf_ci = grp;
        uint32_t pf_cnt = 0;
        auto prefetch_upto = [&](uint32_t last_cnt) {
            while (pf_cnt <= last_cnt && pf_tile < num_tiles) {
                int m_tile, n_tile, c1, c2, c3;
                split_tile(pf_tile, m_tile, n_tile);
                if (pf_ci >= nch_of(n_tile)) { pf_tile += nworkers; pf_ci = grp; continue; }
                origin(m_tile, c1, c2, c3);
                const int s = pf_cnt & (kSlots - 1);
                mbar_arrive_expect_tx(&my_rfull[s], p.res_bytes);
                tma_load_4d(slots + s * kSlotBytes, &tmRes, &my_rfull[s], n_tile * ncols + pf_ci * CW, c1, c2, c3);
                ++pf_cnt;
                pf_ci += kG;
            }
        };
        const uint32_t tempty_addr0 = k2 ? mapa_cluster(smem_u32(&tempty[0]), 0) : 0u;   // leader's ring barriers
        // box coordinates of this thread's tile row (the same for every tile)
        int r_1, r_2, r_3;
        {
            int rq;
            p.fd_b1.divmod(r, rq, r_1);
            p.fd_b2.divmod(rq, r_3, r_2);
        }

        for (int tile = worker; tile < num_tiles; tile += nworkers, ++tc) {
            int m_tile, n_tile, o1, o2, o3;
            split_tile(tile, m_tile, n_tile);
            origin(m_tile, o1, o2, o3);
            const float* gb = nullptr;
            const float* res_row = nullptr;        // bf16-out path: residual row read straight from global
            bool row_ok = false;
            long long grow_ = 0;
            if (p.group_bias != nullptr || (!kF32 && p.has_residual) || kCS || kMode == 2) {
                const int g1 = o1 + r_1, g2 = o2 + r_2, g3 = o3 + r_3;
                if (r < p.rows_per_tile && g1 < p.D1 && g2 < p.D2 && g3 < p.D3) {
                    row_ok = true;
                    const long long grow = (static_cast<long long>(g3) * p.D2 + g2) * p.D1 + g1;
                    grow_ = grow;
                    if (p.group_bias != nullptr) gb = p.group_bias + p.fd_rpg.div(static_cast<int>(grow)) * p.ld_group_bias;
                    if (!kF32 && p.has_residual) res_row = p.residual + grow * p.ld_residual;
                }
            }
            // fused GroupNorm statistics: this warp's 32 rows are 32 / seg_lw segments, each inside one unit (-1: no rows)
            long long cs_unit[4] = {-1, -1, -1, -1};
            bool warp_ragged = false;
            if constexpr (kCS) {
                warp_ragged = __any_sync(0xffffffffu, !row_ok);
                const int rb = q * 32;                                    // first tile row of this warp
                const int nseg = p.seg_lw == 32 ? 1 : (p.seg_lw == 16 ? 2 : 4);
                for (int sg = 0; sg < nseg; ++sg) {
                    const int row = rb + sg * p.seg_lw;                   // first row of the segment
                    int rq, ru, rg2, rg3;
                    p.fd_b1.divmod(row, rq, ru);
                    p.fd_b2.divmod(rq, rg3, rg2);
                    const int g2 = o2 + rg2, g3 = o3 + rg3;
                    if (row < p.rows_per_tile && g2 < p.D2 && g3 < p.D3)
                        cs_unit[sg] = p.gn_unit_mode ? static_cast<long long>(g3) * p.D2 + g2 : g3;
                }
            }
            const uint32_t a = tc & 1, aph = (tc >> 1) & 1;
            bool acc_ready = false;
            const uint32_t t_row = tmem_base + a * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
            float ln_ra = 0.f, ln_rb = 0.f;            // consumer: rstd and -rstd * mean of this thread's row
            float ln_s = 0.f, ln_q = 0.f;              // producer: sum / sum of squares of the row's columns seen here
            if constexpr (kLNC) {
                if (row_ok) {
                    const longlong2 sq = __ldg(reinterpret_cast<const longlong2*>(p.ln_acc + 2 * grow_));
                    const float mean = static_cast<float>(sq.x) * (1.0f / kLnFixSum) * p.ln_inv_c;
                    const float ex2 = static_cast<float>(sq.y) * (1.0f / kLnFixSq) * p.ln_inv_c;
                    ln_ra = rsqrtf(fmaxf(ex2 - mean * mean, 0.f) + p.ln_eps);
                    ln_rb = -ln_ra * mean;
                }
            }

            const int nch_t = nch_of(n_tile);
            for (int ci = grp; ci < nch_t; ci += kG, ++cnt) {
                const int s = cnt & (kSlots - 1);
                uint8_t* slot = slots + s * kSlotBytes;
                const int col0 = n_tile * ncols + ci * CW;              // first output column of this chunk
                if (!acc_ready) {
                    mbar_wait(&tfull[a], aph);
                    tc_fence_after();
                    acc_ready = true;
                    if (leader) DCB_TR(2 + grp, 4 + grp, tc);
                }
                if (tma_res) {
                    if (leader) {
                        tma_store_wait_read<0>();                // slot of item cnt + 1 was read by the latest store
                        prefetch_upto(cnt + 1);
                    }
                    mbar_wait(&my_rfull[s], (cnt / kSlots) & 1);
                    if (leader) DCB_TR(2 + grp, 8 + grp, cnt);
                }
                // (no TMA residual: the slot was proven free before the previous chunk's barrier, see below)
                uint8_t* orow = slot + r * 64;                          // 64-byte rows, SWIZZLE_64B
                // 16 columns at a time (keeps the 576-thread CTA under 113 registers per thread)
#pragma unroll
                for (int hf = 0; hf < CW / 16; ++hf) {
                    const int cb = col0 + hf * 16;
                    float o[16];
                    if (!kF32 && p.has_residual) {
                        if (res_row != nullptr && cb + 16 <= p.N_out) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 t = __ldg(reinterpret_cast<const float4*>(res_row + cb + j));
                                o[j] = t.x; o[j + 1] = t.y; o[j + 2] = t.z; o[j + 3] = t.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                o[j] = (res_row != nullptr && cb + j < p.N_out) ? __ldg(res_row + cb + j) : 0.f;
                        }
                    }                                          // (otherwise o[] is assigned from the accumulator below)
                    const bool o_has_res = !kF32 && p.has_residual;
                    uint32_t v[16];
                    tmem_ld16(t_row + ci * CW + hf * 16, v);
                    if (!kF32 && p.geglu) {
                        uint32_t g[16];
                        tmem_ld16(t_row + BN / 2 + ci * CW + hf * 16, g);
                        tmem_ld_wait();
                        const int wn = n_tile * BN + ci * CW + hf * 16;   // packed weight column of the value half
                        const float s_ = p.out_scale;
                        if constexpr (kLNC) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + wn + j));
                                const float4 bg = __ldg(reinterpret_cast<const float4*>(p.bias + wn + BN / 2 + j));
                                const float4 sv = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + wn + j));
                                const float4 sg = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + wn + BN / 2 + j));
                                o[j] = fmaf(__uint_as_float(v[j]), ln_ra, fmaf(ln_rb, sv.x, bv.x)) *
                                       gelu_erf(fmaf(__uint_as_float(g[j]), ln_ra, fmaf(ln_rb, sg.x, bg.x)));
                                o[j + 1] = fmaf(__uint_as_float(v[j + 1]), ln_ra, fmaf(ln_rb, sv.y, bv.y)) *
                                           gelu_erf(fmaf(__uint_as_float(g[j + 1]), ln_ra, fmaf(ln_rb, sg.y, bg.y)));
                                o[j + 2] = fmaf(__uint_as_float(v[j + 2]), ln_ra, fmaf(ln_rb, sv.z, bv.z)) *
                                           gelu_erf(fmaf(__uint_as_float(g[j + 2]), ln_ra, fmaf(ln_rb, sg.z, bg.z)));
                                o[j + 3] = fmaf(__uint_as_float(v[j + 3]), ln_ra, fmaf(ln_rb, sv.w, bv.w)) *
                                           gelu_erf(fmaf(__uint_as_float(g[j + 3]), ln_ra, fmaf(ln_rb, sg.w, bg.w)));
                            }
                        } else
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
                            if (p.bias != nullptr) {
                                bv = __ldg(reinterpret_cast<const float4*>(p.bias + wn + j));
                                bg = __ldg(reinterpret_cast<const float4*>(p.bias + wn + BN / 2 + j));
                            }
                            o[j] = (__uint_as_float(v[j]) * s_ + bv.x) * gelu_erf(__uint_as_float(g[j]) * s_ + bg.x);
                            o[j + 1] = (__uint_as_float(v[j + 1]) * s_ + bv.y) * gelu_erf(__uint_as_float(g[j + 1]) * s_ + bg.y);
                            o[j + 2] = (__uint_as_float(v[j + 2]) * s_ + bv.z) * gelu_erf(__uint_as_float(g[j + 2]) * s_ + bg.z);
                            o[j + 3] = (__uint_as_float(v[j + 3]) * s_ + bv.w) * gelu_erf(__uint_as_float(g[j + 3]) * s_ + bg.w);
                        }
                    } else if constexpr (kLNC) {
                        tmem_ld_wait();
                        if (cb + 16 <= p.N_out) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + cb + j));
                                const float4 ss = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + cb + j));
                                o[j] = fmaf(__uint_as_float(v[j]), ln_ra, fmaf(ln_rb, ss.x, bb.x));
                                o[j + 1] = fmaf(__uint_as_float(v[j + 1]), ln_ra, fmaf(ln_rb, ss.y, bb.y));
                                o[j + 2] = fmaf(__uint_as_float(v[j + 2]), ln_ra, fmaf(ln_rb, ss.z, bb.z));
                                o[j + 3] = fmaf(__uint_as_float(v[j + 3]), ln_ra, fmaf(ln_rb, ss.w, bb.w));
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                o[j] = cb + j < p.N_out ? fmaf(__uint_as_float(v[j]), ln_ra, fmaf(ln_rb, __ldg(p.ln_colsum + cb + j),
                                                                                                    __ldg(p.bias + cb + j))) : 0.f;
                        }
                    } else {
                        tmem_ld_wait();
                        if (o_has_res) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = fmaf(__uint_as_float(v[j]), p.out_scale, o[j]);
                        } else if (p.out_scale != 1.0f) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) * p.out_scale;
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
                        }
                        if (p.bias != nullptr) {
                            if (cb + 16 <= p.N_out) {
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + cb + j));
                                    o[j] += t.x; o[j + 1] += t.y; o[j + 2] += t.z; o[j + 3] += t.w;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (cb + j < p.N_out) o[j] += __ldg(p.bias + cb + j);
                            }
                        }
                    }
                    if (gb != nullptr) {
                        if (cb + 16 <= p.N_out && (reinterpret_cast<uintptr_t>(gb + cb) & 15) == 0) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 t = __ldg(reinterpret_cast<const float4*>(gb + cb + j));
                                o[j] += t.x; o[j + 1] += t.y; o[j + 2] += t.z; o[j + 3] += t.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (cb + j < p.N_out) o[j] += __ldg(gb + cb + j);
                        }
                    }
                    if constexpr (kF32) {
                        if (tma_res) {                                   // fp32 residual chunk prefetched into this slot
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float4 t = *reinterpret_cast<const float4*>(orow + ((c ^ swz) << 4));
                                o[4 * c] += t.x; o[4 * c + 1] += t.y; o[4 * c + 2] += t.z; o[4 * c + 3] += t.w;
                            }
                        }
                    }
                    if constexpr (kCS) {
                        // rows outside the tensor are clipped by the TMA store; zero them so that the column sums read
                        // back from the slot only see real rows (warp-uniform branch: only ragged tiles pay for it)
                        if (warp_ragged) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = row_ok ? o[j] : 0.f;
                        }
                    }
                    if constexpr (kLNP) {
                        // LayerNorm folding, producer side: this row's statistics and the bf16 operand copy of the
                        // final fp32 values (one full 32-byte sector per thread and chunk)
                        if (row_ok) {
                            if (cb + 16 <= p.N_out) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) { ln_s += o[j]; ln_q = fmaf(o[j], o[j], ln_q); }
                                uint4 pk[2];
#pragma unroll
                                for (int c = 0; c < 2; ++c) {
                                    __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * c], o[8 * c + 1]);
                                    __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * c + 2], o[8 * c + 3]);
                                    __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * c + 4], o[8 * c + 5]);
                                    __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * c + 6], o[8 * c + 7]);
                                    pk[c].x = *reinterpret_cast<uint32_t*>(&h0); pk[c].y = *reinterpret_cast<uint32_t*>(&h1);
                                    pk[c].z = *reinterpret_cast<uint32_t*>(&h2); pk[c].w = *reinterpret_cast<uint32_t*>(&h3);
                                }
                                uint4* dst = reinterpret_cast<uint4*>(p.ln_x + grow_ * p.ld_ln_x + cb);
                                dst[0] = pk[0];
                                dst[1] = pk[1];
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (cb + j < p.N_out) {
                                        ln_s += o[j]; ln_q = fmaf(o[j], o[j], ln_q);
                                        p.ln_x[grow_ * p.ld_ln_x + cb + j] = __float2bfloat16_rn(o[j]);
                                    }
                            }
                        }
                    }
                    if constexpr (kF32) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<float4*>(orow + ((c ^ swz) << 4)) =
                                make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint4 pk;
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * c], o[8 * c + 1]);
                            __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * c + 2], o[8 * c + 3]);
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * c + 4], o[8 * c + 5]);
                            __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * c + 6], o[8 * c + 7]);
                            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                            *reinterpret_cast<uint4*>(orow + (((hf * 2 + c) ^ swz) << 4)) = pk;
                        }
                    }
                }
                fence_proxy_async();                                     // generic-proxy writes -> visible to TMA
                // the next chunk writes the other slot: its last store (issued one chunk ago) must have read it;
                // checking before this barrier makes that visible to the whole group without a second barrier
                if (!tma_res && leader) tma_store_wait_read<kSlots - 2>();
                named_bar_sync(bar_id, 128);
                if (leader) {
                    tma_store_4d(&tmOut, slot, col0, o1, o2, o3);
                    tma_store_commit();
                    DCB_TR(2 + grp, 12 + grp, cnt);
                }
                if constexpr (kCS) {
                    // the slot now holds the final values of all 128 rows: fold this warp's 32 rows per column
                    if (p.seg_lw == 32) gn_acc_from_slot<kF32, 32>(slot, q, lane, p.gn_acc, cs_unit, col0, p.N_out, p.fd_cg);
                    else if (p.seg_lw == 16) gn_acc_from_slot<kF32, 16>(slot, q, lane, p.gn_acc, cs_unit, col0, p.N_out, p.fd_cg);
                    else gn_acc_from_slot<kF32, 8>(slot, q, lane, p.gn_acc, cs_unit, col0, p.N_out, p.fd_cg);
                    // the leader's next residual prefetch (TMA write into a slot) must not overtake these reads
                    if (tma_res) named_bar_sync(bar_id, 128);
                }
            }
            if constexpr (kLNP) {
                if (row_ok && grp < nch_t) {
                    atomicAdd(p.ln_acc + 2 * grow_, static_cast<unsigned long long>(__float2ll_rn(ln_s * kLnFixSum)));
                    atomicAdd(p.ln_acc + 2 * grow_ + 1, static_cast<unsigned long long>(__float2ll_rn(ln_q * kLnFixSq)));
                }
            }
            if (!acc_ready) {                                            // group without chunks still tracks the ring
                mbar_wait(&tfull[a], aph);
                tc_fence_after();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (k2) mbar_arrive_cluster(tempty_addr0 + a * 8); else mbar_arrive(&tempty[a]);
            }
        }
        if (leader) tma_store_wait_all<0>();
        if (leader) DCB_TR_DONE(2 + grp);
    }
    if (warp == 0 && lane == 0) DCB_TR_DONE(0);
    if (warp == 1 && lane == 0) DCB_TR_DONE(1);

    tc_fence_before();
    if constexpr (k2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        if constexpr (k2) tmem_dealloc_pair<kTmemCols>(tmem_base); else tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------ host
template <int BN, bool kF32, bool k2, int kMode>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB2, const CUtensorMap& tmOut,
                  const CUtensorMap& tmRes, TapGemmParams& p, cudaStream_t stream) {
    constexpr int kBRows = k2 ? BN / 2 : BN;
    constexpr int kStageBytes = kABytes + kBRows * kBlockK * 2;
    constexpr int smem_bytes = StageCount<BN, k2>::value * kStageBytes + StageCount<BN, k2>::epi_bytes + 1024 + 256;
    static_assert(smem_bytes <= 232448, "shared memory budget");
    static_assert(StageCount<BN, k2>::value >= 4, "pipeline depth");
    static_assert((2 * StageCount<BN, k2>::value + 4 + EpiGroups<k2>::value * EpiSlots<k2>::value) * 8 + 4 <= 256, "barrier area");
    static bool configured[kMaxDevices] = {};
    const int dv = current_device() & (kMaxDevices - 1);
    if (!configured[dv]) {
        cudaError_t e = cudaFuncSetAttribute(tap_gemm_kernel<BN, kF32, k2, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             smem_bytes);
        if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured[dv] = true;
    }
    p.stage_bytes = static_cast<uint32_t>(p.rows_per_tile * kBlockK * 2 + kBRows * kBlockK * 2);
    p.res_bytes = static_cast<uint32_t>(p.rows_per_tile * 64);
    {   // ragged last N tile (see TapGemmParams): UMMA N and the staged B rows shrink to what the tile holds
        const int rem = p.N - (p.num_n_tiles - 1) * BN;                  // 1 .. BN valid accumulator columns
        int n_eff = (rem + 31) / 32 * 32;
        if (n_eff > BN || p.geglu) n_eff = BN;
        p.idesc_last = make_idesc_bf16(k2 ? 2 * kBlockM : kBlockM, n_eff);
        p.b_rows_last = k2 ? n_eff / 2 : n_eff;
        p.stage_bytes_last = static_cast<uint32_t>(p.rows_per_tile * kBlockK * 2 + p.b_rows_last * kBlockK * 2);
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    cfg.blockDim = dim3(Threads<k2>::value);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    const int m_items = k2 ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles;
    {
        const int cap = k2 ? num_sms() / 2 : num_sms();
        const int items = m_items * p.num_n_tiles;
        const int workers = items < cap ? items : cap;
        p.sched_w = workers;
        p.sched_full_groups = m_items / workers;
        const int r = m_items % workers;
        p.fd_gsz = FastDiv::make(static_cast<uint32_t>(workers) * p.num_n_tiles);
        p.fd_w = FastDiv::make(workers);
        p.fd_r = FastDiv::make(r > 0 ? r : 1);
        p.fd_ntiles = FastDiv::make(p.num_n_tiles);
        // DCB_GEMM_ORDER=1 selects the wave-major order.  Measured (r02e): never faster than N-fastest, up to 35 % slower
        // (20480x640x2560): neighbouring workers sharing an A tile at the same time beats one worker re-reading it
        static int wave = -1;
        if (wave < 0) { const char* e = getenv("DCB_GEMM_ORDER"); wave = e ? atoi(e) : 0; }
        p.sched_wave = wave;
    }
    if (k2) {
        const int items = m_items * p.num_n_tiles;
        const int pairs = items < num_sms() / 2 ? items : num_sms() / 2;
        cfg.gridDim = dim3(2 * pairs);
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    } else {
        const int tiles = p.num_m_tiles * p.num_n_tiles;
        cfg.gridDim = dim3(tiles < num_sms() ? tiles : num_sms());
        cfg.attrs = attr;
        cfg.numAttrs = 0;
    }
    if (pdl_enabled(kPdlGemm)) {
        attr[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
        ++cfg.numAttrs;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, tap_gemm_kernel<BN, kF32, k2, kMode>, tmA, tmB, tmB2, tmOut, tmRes, p);
    if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "tap_gemm launch: %s", cudaGetErrorString(e));
    return DCB_OK;
}

template <int BN, bool kF32, int kMode>
static int launch_mode(bool pair, const CUtensorMap& tmA, const CUtensorMap* tmBs, const CUtensorMap& tmOut,
                       const CUtensorMap& tmRes, TapGemmParams& p, cudaStream_t s) {
    return pair ? launch<BN, kF32, true, kMode>(tmA, tmBs[0], tmBs[1], tmOut, tmRes, p, s)
                : launch<BN, kF32, false, kMode>(tmA, tmBs[0], tmBs[1], tmOut, tmRes, p, s);
}

// all tile-shape / dtype / pair variants of one statistics mode
template <int kMode>
static int launch_all(int bn, bool out_bf16, bool pair, const CUtensorMap& tmA, const CUtensorMap* tmB,
                      const CUtensorMap& tmOut, const CUtensorMap& tmRes, TapGemmParams& p, cudaStream_t s) {
    if (out_bf16) {
        if (bn == 256) return launch_mode<256, false, kMode>(pair, tmA, tmB, tmOut, tmRes, p, s);
        if (bn == 192) return launch_mode<192, false, kMode>(pair, tmA, tmB, tmOut, tmRes, p, s);
        return bn == 160 ? launch_mode<160, false, kMode>(pair, tmA, tmB, tmOut, tmRes, p, s)
                         : launch_mode<128, false, kMode>(pair, tmA, tmB, tmOut, tmRes, p, s);
    }
    if (bn == 256) return launch_mode<256, true, kMode>(pair, tmA, tmB, tmOut, tmRes, p, s);
    if (bn == 192) return launch_mode<192, true, kMode>(pair, tmA, tmB, tmOut, tmRes, p, s);
    return bn == 160 ? launch_mode<160, true, kMode>(pair, tmA, tmB, tmOut, tmRes, p, s)
                     : launch_mode<128, true, kMode>(pair, tmA, tmB, tmOut, tmRes, p, s);
}


}  // namespace dcb
