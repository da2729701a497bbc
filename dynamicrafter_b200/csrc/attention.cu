// Attention kernels, head dim 64, bf16 operands, fp32 online softmax.
//
//  * flash_attention_kernel: spatial self-attention (attention.py:101-125, K6) and the dual text+image
//    cross-attention (attention.py:89-94,129-142, K7) as ONE kernel: up to two key/value segments, each with
//    its own softmax, combined as out = seg_scale[0]*softmax(QK0^T)V0 + seg_scale[1]*softmax(QK1^T)V1.
//    The N x N score matrix is never materialised.
//  * temporal_attention_kernel: self-attention over <=16 frames per (batch, location, head)
//    (attention.py:387-390, K8); reads q/k/v rows with the frame stride of the channels-last layout so no
//    '(b t) hw c <-> (b hw) t c' transposes exist anywhere.
//
//  * cross_attention_kernel: text + image cross-attention with all keys resident in shared memory.
//
// These are warp-level mma.sync (HMMA) kernels; the long-sequence spatial self-attention runs on tcgen05 / TMEM
// (flash_tc.cu) and only falls back to flash_attention_kernel for short sequences.
#include <cstdlib>
#include <cuda_bf16.h>

#include "dcb_internal.h"
#include "ptx.cuh"

namespace dcb {

constexpr int kD = 64;   // head dim

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;   // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// tile rows are 64 bf16 = 8 chunks of 16 B; chunk index XOR (row & 7) => conflict-free ldmatrix
__device__ __forceinline__ int swz(int row, int chunk) { return row * 64 + ((chunk ^ (row & 7)) << 3); }

struct FlashParams {
    const __nv_bfloat16* q; long long ldq;
    const __nv_bfloat16* k[2]; const __nv_bfloat16* v[2]; long long ldk[2], ldv[2];
    int nk[2]; int kv_div[2]; float seg_scale[2];
    int nseg, nq, heads;
    __nv_bfloat16* o; long long ldo;
    float scale_log2;
};

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
flash_attention_kernel(const FlashParams p) {
    constexpr int BM = kWarps * 16, BN = 64, kThreads = kWarps * 32;
    extern __shared__ __align__(128) uint8_t fa_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(fa_smem);
    __nv_bfloat16* sK = sQ + BM * kD;            // [2][BN*kD]
    __nv_bfloat16* sV = sK + 2 * BN * kD;        // [2][BN*kD]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
    const int g = lane >> 2, t = lane & 3;
    pdl_launch_dependents();
    pdl_wait();

    // ---- Q tile -> smem (zero filled past nq)
    {
        const __nv_bfloat16* qb = p.q + (static_cast<long long>(b) * p.nq) * p.ldq + h * kD;
        for (int i = tid; i < BM * 8; i += kThreads) {
            const int r = i >> 3, c = i & 7;
            const bool ok = (q0 + r) < p.nq;
            cp_async16(sQ + swz(r, c), qb + static_cast<long long>(ok ? q0 + r : 0) * p.ldq + c * 8, ok);
        }
        cp_async_commit();
    }
    uint32_t qf[4][4];
    float out[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { out[i][0] = out[i][1] = out[i][2] = out[i][3] = 0.f; }
    bool q_loaded = false;

    for (int seg = 0; seg < p.nseg; ++seg) {
        const int nk = p.nk[seg];
        const int kvb = b / p.kv_div[seg];
        const __nv_bfloat16* kb = p.k[seg] + (static_cast<long long>(kvb) * nk) * p.ldk[seg] + h * kD;
        const __nv_bfloat16* vb = p.v[seg] + (static_cast<long long>(kvb) * nk) * p.ldv[seg] + h * kD;
        const int ntiles = (nk + BN - 1) / BN;
        auto load_kv = [&](int tile, int buf) {
            const int k0 = tile * BN;
            for (int i = tid; i < BN * 8; i += kThreads) {
                const int r = i >> 3, c = i & 7;
                const bool ok = (k0 + r) < nk;
                const long long row = ok ? (k0 + r) : 0;
                cp_async16(sK + buf * BN * kD + swz(r, c), kb + row * p.ldk[seg] + c * 8, ok);
                cp_async16(sV + buf * BN * kD + swz(r, c), vb + row * p.ldv[seg] + c * 8, ok);
            }
            cp_async_commit();
        };
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
        float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

        load_kv(0, 0);
        for (int tile = 0; tile < ntiles; ++tile) {
            const int buf = tile & 1;
            if (tile + 1 < ntiles) { load_kv(tile + 1, buf ^ 1); cp_async_wait<1>(); }
            else { cp_async_wait<0>(); }
            __syncthreads();
            if (!q_loaded) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = ks * 2 + (lane >> 4);
                    ldsm_x4(smem_addr(sQ + swz(r, c)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
                }
                q_loaded = true;
            }
            const __nv_bfloat16* tK = sK + buf * BN * kD;
            const __nv_bfloat16* tV = sV + buf * BN * kD;
            // ---- S = Q K^T  (16 x 64 per warp)
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    uint32_t b0, b1, b2, b3;
                    const int r = nt * 8 + (lane & 7), c = kp * 4 + (lane >> 3);
                    ldsm_x4(smem_addr(tK + swz(r, c)), b0, b1, b2, b3);
                    mma_bf16(s[nt], qf[kp * 2], b0, b1);
                    mma_bf16(s[nt], qf[kp * 2 + 1], b2, b3);
                }
            }
            // ---- mask the ragged last tile, online softmax (fp32)
            const int k0 = tile * BN;
            if (k0 + BN > nk) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const int key = k0 + nt * 8 + 2 * t;
                    if (key >= nk) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
                    if (key + 1 >= nk) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
                }
            }
            float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
                mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
            }
            float corr[2], msc[2], rs[2] = {0.f, 0.f};
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                corr[hh] = exp2f((m_run[hh] - mx[hh]) * p.scale_log2);
                msc[hh] = mx[hh] * p.scale_log2;
                m_run[hh] = mx[hh];
            }
            uint32_t pf[4][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float p0 = exp2f(s[nt][0] * p.scale_log2 - msc[0]);
                const float p1 = exp2f(s[nt][1] * p.scale_log2 - msc[0]);
                const float p2 = exp2f(s[nt][2] * p.scale_log2 - msc[1]);
                const float p3 = exp2f(s[nt][3] * p.scale_log2 - msc[1]);
                rs[0] += p0 + p1;
                rs[1] += p2 + p3;
                // C fragment of S (n-tiles 2j, 2j+1) == A fragment of P for key block j
                pf[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
                pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) l_run[hh] = l_run[hh] * corr[hh] + rs[hh];
#pragma unroll
            for (int nd = 0; nd < 8; ++nd) {
                acc[nd][0] *= corr[0]; acc[nd][1] *= corr[0];
                acc[nd][2] *= corr[1]; acc[nd][3] *= corr[1];
            }
            // ---- O += P V
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int ndp = 0; ndp < 4; ++ndp) {
                    uint32_t b0, b1, b2, b3;
                    const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = ndp * 2 + (lane >> 4);
                    ldsm_x4_t(smem_addr(tV + swz(r, c)), b0, b1, b2, b3);
                    mma_bf16(acc[ndp * 2], pf[kk], b0, b1);
                    mma_bf16(acc[ndp * 2 + 1], pf[kk], b2, b3);
                }
            }
            __syncthreads();   // everyone done with buf before it is refilled
        }
        // ---- finish this segment's softmax and fold into the output
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
            l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
        }
        const float i0 = p.seg_scale[seg] / l_run[0], i1 = p.seg_scale[seg] / l_run[1];
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
            out[nd][0] += acc[nd][0] * i0; out[nd][1] += acc[nd][1] * i0;
            out[nd][2] += acc[nd][2] * i1; out[nd][3] += acc[nd][3] * i1;
        }
    }
    // ---- O -> bf16 via this warp's rows of sQ, then 16-byte coalesced stores
    __syncwarp();
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
        const int r0 = warp * 16 + g, cchunk = nd, within = 2 * t;
        *reinterpret_cast<uint32_t*>(sQ + swz(r0, cchunk) + within) = pack_bf16(out[nd][0], out[nd][1]);
        *reinterpret_cast<uint32_t*>(sQ + swz(r0 + 8, cchunk) + within) = pack_bf16(out[nd][2], out[nd][3]);
    }
    __syncwarp();
    __nv_bfloat16* ob = p.o + (static_cast<long long>(b) * p.nq) * p.ldo + h * kD;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + i * 32, r = warp * 16 + (idx >> 3), c = idx & 7;
        if (q0 + r < p.nq)
            *reinterpret_cast<uint4*>(ob + static_cast<long long>(q0 + r) * p.ldo + c * 8) =
                *reinterpret_cast<const uint4*>(sQ + swz(r, c));
    }
}

// -------------------------------------------------------------------------------------------------
// Cross-attention against SHORT key sets (text 77 + image 16 tokens, attention.py:89-94,129-142): the whole K/V of
// both segments of one (batch, head) sits in shared memory (24 KB), so there is no online softmax and no block-level
// synchronisation after the prologue: the job is a stream over Q and O (HBM bound, 4 * 64 * 93 FLOP per 256 bytes).
// Each warp owns 16-row query tiles, prefetches them two tiles ahead with cp.async into a private 3-stage ring,
// computes S for all keys of a segment in registers, normalises P before the P V product (so one accumulator
// serves both segments) and writes O through its Q stage with 16-byte stores.
// NT0 / NT1: 8-key tiles of the two segments (even; keys past nk are zero-filled and masked).
// -------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void cross_segment(const __nv_bfloat16* tK, const __nv_bfloat16* tV, int nk, float scale_log2,
                                              float seg_scale, const uint32_t (&qf)[4][4], float (&out)[8][4], int lane) {
    const int t = lane & 3;
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            uint32_t b0, b1, b2, b3;
            const int r = nt * 8 + (lane & 7), c = kp * 4 + (lane >> 3);
            ldsm_x4(smem_addr(tK + swz(r, c)), b0, b1, b2, b3);
            mma_bf16(s[nt], qf[kp * 2], b0, b1);
            mma_bf16(s[nt], qf[kp * 2 + 1], b2, b3);
        }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int key = nt * 8 + 2 * t;
        if (key >= nk) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (key + 1 >= nk) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
        mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
        mx[hh] *= scale_log2;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = exp2f(s[nt][0] * scale_log2 - mx[0]);
        s[nt][1] = exp2f(s[nt][1] * scale_log2 - mx[0]);
        s[nt][2] = exp2f(s[nt][2] * scale_log2 - mx[1]);
        s[nt][3] = exp2f(s[nt][3] * scale_log2 - mx[1]);
        rs[0] += s[nt][0] + s[nt][1];
        rs[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 1);
        rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 2);
        rs[hh] = seg_scale / rs[hh];
    }
    // O += (P / l * seg_scale) V : C fragments of S (n-tiles 2j, 2j+1) are the A fragment of key block j
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
        uint32_t pf[4];
        pf[0] = pack_bf16(s[2 * kk][0] * rs[0], s[2 * kk][1] * rs[0]);
        pf[1] = pack_bf16(s[2 * kk][2] * rs[1], s[2 * kk][3] * rs[1]);
        pf[2] = pack_bf16(s[2 * kk + 1][0] * rs[0], s[2 * kk + 1][1] * rs[0]);
        pf[3] = pack_bf16(s[2 * kk + 1][2] * rs[1], s[2 * kk + 1][3] * rs[1]);
#pragma unroll
        for (int ndp = 0; ndp < 4; ++ndp) {
            uint32_t b0, b1, b2, b3;
            const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = ndp * 2 + (lane >> 4);
            ldsm_x4_t(smem_addr(tV + swz(r, c)), b0, b1, b2, b3);
            mma_bf16(out[ndp * 2], pf, b0, b1);
            mma_bf16(out[ndp * 2 + 1], pf, b2, b3);
        }
    }
}

constexpr int kCrossWarps = 8, kCrossStages = 3;

template <int NT0, int NT1>
__global__ void __launch_bounds__(kCrossWarps * 32, 2)
cross_attention_kernel(const FlashParams p, int tiles_per_warp) {
    static_assert(NT0 % 2 == 0 && NT1 % 2 == 0, "key tiles come in pairs (16-key P V steps)");
    constexpr int K0 = NT0 * 8, K1 = NT1 * 8;
    extern __shared__ __align__(128) uint8_t fa_smem[];
    __nv_bfloat16* sK0 = reinterpret_cast<__nv_bfloat16*>(fa_smem);
    __nv_bfloat16* sV0 = sK0 + K0 * kD;
    __nv_bfloat16* sK1 = sV0 + K0 * kD;
    __nv_bfloat16* sV1 = sK1 + K1 * kD;
    __nv_bfloat16* sQ = sV1 + K1 * kD;                  // [warp][stage][16 * 64]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int g = lane >> 2, t = lane & 3;
    pdl_launch_dependents();
    pdl_wait();

    // ---- K / V of both segments -> smem (rows past nk zero filled)
    auto load_kv = [&](int seg, __nv_bfloat16* dK, __nv_bfloat16* dV, int rows) {
        const int nk = p.nk[seg];
        const long long kvb = b / p.kv_div[seg];
        const __nv_bfloat16* kb = p.k[seg] + kvb * nk * p.ldk[seg] + h * kD;
        const __nv_bfloat16* vb = p.v[seg] + kvb * nk * p.ldv[seg] + h * kD;
        for (int i = tid; i < rows * 8; i += kCrossWarps * 32) {
            const int r = i >> 3, c = i & 7;
            const bool ok = r < nk;
            const long long row = ok ? r : 0;
            cp_async16(dK + swz(r, c), kb + row * p.ldk[seg] + c * 8, ok);
            cp_async16(dV + swz(r, c), vb + row * p.ldv[seg] + c * 8, ok);
        }
    };
    load_kv(0, sK0, sV0, K0);
    load_kv(1, sK1, sV1, K1);
    cp_async_commit();

    // ---- this warp's query tiles: first + i * kCrossWarps, i < tiles_per_warp
    const int ntiles = (p.nq + 15) / 16;
    const int first = blockIdx.x * kCrossWarps * tiles_per_warp + warp;
    __nv_bfloat16* myQ = sQ + warp * kCrossStages * 16 * kD;
    const __nv_bfloat16* qb = p.q + (static_cast<long long>(b) * p.nq) * p.ldq + h * kD;
    __nv_bfloat16* ob = p.o + (static_cast<long long>(b) * p.nq) * p.ldo + h * kD;
    auto load_q = [&](int i) {              // always commits a group (possibly empty) so the wait counts stay fixed
        const int tile = first + i * kCrossWarps;
        if (i < tiles_per_warp && tile < ntiles) {
            __nv_bfloat16* dst = myQ + (i % kCrossStages) * 16 * kD;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = lane + j * 32, r = idx >> 3, c = idx & 7;
                const int row = tile * 16 + r;
                const bool ok = row < p.nq;
                cp_async16(dst + swz(r, c), qb + static_cast<long long>(ok ? row : 0) * p.ldq + c * 8, ok);
            }
        }
        cp_async_commit();
    };
    load_q(0);
    load_q(1);
    cp_async_wait<2>();                     // this thread's share of K / V has landed
    __syncthreads();

    for (int i = 0; i < tiles_per_warp; ++i) {
        const int tile = first + i * kCrossWarps;
        if (tile >= ntiles) break;
        load_q(i + 2);
        cp_async_wait<2>();                 // tile i (two younger groups may still be in flight)
        __syncwarp();
        __nv_bfloat16* tQ = myQ + (i % kCrossStages) * 16 * kD;
        uint32_t qf[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int r = (lane & 7) + ((lane >> 3) & 1) * 8, c = ks * 2 + (lane >> 4);
            ldsm_x4(smem_addr(tQ + swz(r, c)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        }
        float out[8][4];
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) { out[nd][0] = out[nd][1] = out[nd][2] = out[nd][3] = 0.f; }
        cross_segment<NT0>(sK0, sV0, p.nk[0], p.scale_log2, p.seg_scale[0], qf, out, lane);
        cross_segment<NT1>(sK1, sV1, p.nk[1], p.scale_log2, p.seg_scale[1], qf, out, lane);
        // ---- O -> bf16 through this tile's Q stage, then 16-byte stores
        __syncwarp();
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
            *reinterpret_cast<uint32_t*>(tQ + swz(g, nd) + 2 * t) = pack_bf16(out[nd][0], out[nd][1]);
            *reinterpret_cast<uint32_t*>(tQ + swz(g + 8, nd) + 2 * t) = pack_bf16(out[nd][2], out[nd][3]);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = lane + j * 32, r = idx >> 3, c = idx & 7;
            const int row = tile * 16 + r;
            if (row < p.nq)
                *reinterpret_cast<uint4*>(ob + static_cast<long long>(row) * p.ldo + c * 8) =
                    *reinterpret_cast<const uint4*>(tQ + swz(r, c));
        }
        __syncwarp();                       // stage is refilled by the prefetch of the next iteration
    }
    cp_async_wait<0>();
}

// -------------------------------------------------------------------------------------------------
// Temporal self-attention: one warp per (batch, location, head); L <= 16 frames.
// Row of frame f at location pix of batch b: ((b*L + f)*HW + pix); q/k/v/o are column slices (head h at
// column h*64) of row-major buffers with leading dims ldq/ldk/ldv/ldo.
// -------------------------------------------------------------------------------------------------
struct TemporalParams {
    const __nv_bfloat16 *q, *k, *v;
    __nv_bfloat16* o;
    long long ldq, ldk, ldv, ldo;
    int B, L, HW, heads;
    float scale_log2;
};

__global__ void __launch_bounds__(128)
temporal_attention_kernel(const TemporalParams p) {
    __shared__ __align__(128) __nv_bfloat16 sm[4][3][16 * kD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = blockIdx.x * 4LL + warp;
    const long long items = static_cast<long long>(p.B) * p.HW * p.heads;
    pdl_launch_dependents();
    pdl_wait();
    if (item >= items) return;
    const int h = item % p.heads;
    const long long bp = item / p.heads;
    const int pix = bp % p.HW, b = bp / p.HW;
    const int g = lane >> 2, t = lane & 3;
    __nv_bfloat16* sQ = sm[warp][0];
    __nv_bfloat16* sK = sm[warp][1];
    __nv_bfloat16* sV = sm[warp][2];
    const long long row0 = (static_cast<long long>(b) * p.L) * p.HW + pix;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + i * 32, f = idx >> 3, c = idx & 7;
        const bool ok = f < p.L;
        const long long row = row0 + static_cast<long long>(ok ? f : 0) * p.HW;
        cp_async16(sQ + swz(f, c), p.q + row * p.ldq + h * kD + c * 8, ok);
        cp_async16(sK + swz(f, c), p.k + row * p.ldk + h * kD + c * 8, ok);
        cp_async16(sV + swz(f, c), p.v + row * p.ldv + h * kD + c * 8, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();

    float s[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            uint32_t a0[4], a1[4], b0, b1, b2, b3;
            {
                const int r = (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm_x4(smem_addr(sQ + swz(r, (kp * 2) * 2 + (lane >> 4))), a0[0], a0[1], a0[2], a0[3]);
                ldsm_x4(smem_addr(sQ + swz(r, (kp * 2 + 1) * 2 + (lane >> 4))), a1[0], a1[1], a1[2], a1[3]);
            }
            ldsm_x4(smem_addr(sK + swz(nt * 8 + (lane & 7), kp * 4 + (lane >> 3))), b0, b1, b2, b3);
            mma_bf16(s[nt], a0, b0, b1);
            mma_bf16(s[nt], a1, b2, b3);
        }
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        const int key = nt * 8 + 2 * t;
        if (key >= p.L) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (key + 1 >= p.L) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    }
    float mx[2], sum[2] = {0.f, 0.f};
    mx[0] = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
    mx[1] = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
    }
    uint32_t pf[4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        const float p0 = exp2f((s[nt][0] - mx[0]) * p.scale_log2), p1 = exp2f((s[nt][1] - mx[0]) * p.scale_log2);
        const float p2 = exp2f((s[nt][2] - mx[1]) * p.scale_log2), p3 = exp2f((s[nt][3] - mx[1]) * p.scale_log2);
        sum[0] += p0 + p1;
        sum[1] += p2 + p3;
        pf[nt * 2] = pack_bf16(p0, p1);
        pf[nt * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 1);
        sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 2);
    }
    const float inv0 = 1.f / sum[0], inv1 = 1.f / sum[1];
    float acc[8][4];
#pragma unroll
    for (int ndp = 0; ndp < 4; ++ndp) {
        uint32_t b0, b1, b2, b3;
        const int r = (lane & 7) + ((lane >> 3) & 1) * 8, c = ndp * 2 + (lane >> 4);
        ldsm_x4_t(smem_addr(sV + swz(r, c)), b0, b1, b2, b3);
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[ndp * 2 + j][0] = acc[ndp * 2 + j][1] = acc[ndp * 2 + j][2] = acc[ndp * 2 + j][3] = 0.f;
        mma_bf16(acc[ndp * 2], pf, b0, b1);
        mma_bf16(acc[ndp * 2 + 1], pf, b2, b3);
    }
    __syncwarp();
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
        *reinterpret_cast<uint32_t*>(sQ + swz(g, nd) + 2 * t) = pack_bf16(acc[nd][0] * inv0, acc[nd][1] * inv0);
        *reinterpret_cast<uint32_t*>(sQ + swz(g + 8, nd) + 2 * t) = pack_bf16(acc[nd][2] * inv1, acc[nd][3] * inv1);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + i * 32, f = idx >> 3, c = idx & 7;
        if (f < p.L)
            *reinterpret_cast<uint4*>(p.o + (row0 + static_cast<long long>(f) * p.HW) * p.ldo + h * kD + c * 8) =
                *reinterpret_cast<const uint4*>(sQ + swz(f, c));
    }
}

// fp32 [rows, n] -> softmax(scale * x) bf16 [rows, n]; one warp per row.  VAE mid-block attention
// (ae_modules.py:66-68) materialises its 1-head N x N scores (small: N = h*w of one latent frame).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ x, long long rows, int n, long long ldx, float scale_log2,
                    __nv_bfloat16* __restrict__ y, long long ldy) {
    const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + row * ldx;
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) mx = fmaxf(mx, xr[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < n; i += 32) sum += exp2f((xr[i] - mx) * scale_log2);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    __nv_bfloat16* yr = y + row * ldy;
    for (int i = lane; i < n; i += 32) yr[i] = __float2bfloat16_rn(exp2f((xr[i] - mx) * scale_log2) * inv);
}

}  // namespace dcb

using namespace dcb;

extern "C" int dcb_flash_attention(const dcb_attention_args* a, void* stream) {
    if (a == nullptr || a->q == nullptr || a->o == nullptr || a->nseg < 1 || a->nseg > 2)
        return set_error(DCB_ERR_INVALID, "flash_attention: bad argument");
    if (a->head_dim != kD) return set_error(DCB_ERR_UNSUPPORTED, "flash_attention: head_dim must be 64");
    if (a->batch < 1 || a->heads < 1 || a->nq < 1 || (a->ldq & 7) || (a->ldo & 7))
        return set_error(DCB_ERR_INVALID, "flash_attention: shape/alignment");
    static int tc_min_nq = -1;                  // DCB_FLASH_TC_MIN_NQ: shortest sequence on the tcgen05 kernel
    if (tc_min_nq < 0) { const char* e = getenv("DCB_FLASH_TC_MIN_NQ"); tc_min_nq = e ? atoi(e) : 129; }
    if (a->nq >= tc_min_nq) {
        // spatial self-attention of every UNet level with more than one query tile (2560 / 640 / 160 tokens at 320x512):
        // tcgen05 / TMEM kernel; ragged last query / key tiles are masked (keys) and clipped by TMA (queries)
        static int use_tc = -1;                 // DCB_FLASH_TC=0 forces the mma.sync kernel (A/B testing)
        if (use_tc < 0) { const char* e = getenv("DCB_FLASH_TC"); use_tc = (e == nullptr) ? 1 : atoi(e); }
        static int use_pp = -1;                 // DCB_FLASH_PP=1: ping-pong kernel (flash_pp.cu), two query tiles per CTA
        if (use_pp < 0) { const char* e = getenv("DCB_FLASH_PP"); use_pp = (e == nullptr) ? 0 : atoi(e); }
        if (use_tc && use_pp) {
            const int rc = flash_pp_launch(a, static_cast<cudaStream_t>(stream));
            if (rc != DCB_ERR_UNSUPPORTED) return rc;
        }
        if (use_tc) {
            const int rc = flash_tc_launch(a, static_cast<cudaStream_t>(stream));
            if (rc != DCB_ERR_UNSUPPORTED) return rc;
        }
    }
    FlashParams p;
    p.q = static_cast<const __nv_bfloat16*>(a->q); p.ldq = a->ldq;
    p.o = static_cast<__nv_bfloat16*>(a->o); p.ldo = a->ldo;
    p.nseg = a->nseg; p.nq = a->nq; p.heads = a->heads;
    for (int s = 0; s < 2; ++s) {
        const bool on = s < a->nseg;
        if (on && (a->k[s] == nullptr || a->v[s] == nullptr || a->nk[s] < 1 || a->kv_batch_div[s] < 1 ||
                   (a->ldk[s] & 7) || (a->ldv[s] & 7)))
            return set_error(DCB_ERR_INVALID, "flash_attention: segment %d", s);
        p.k[s] = on ? static_cast<const __nv_bfloat16*>(a->k[s]) : nullptr;
        p.v[s] = on ? static_cast<const __nv_bfloat16*>(a->v[s]) : nullptr;
        p.ldk[s] = a->ldk[s]; p.ldv[s] = a->ldv[s];
        p.nk[s] = on ? a->nk[s] : 0;
        p.kv_div[s] = on ? a->kv_batch_div[s] : 1;
        p.seg_scale[s] = a->seg_scale[s];
    }
    p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (a->heads > 65535 || a->batch > 65535) return set_error(DCB_ERR_INVALID, "flash_attention: grid too large");
    static int use_cross = -1;                  // DCB_CROSS_ATTN=0 forces the tiled kernel (A/B testing)
    if (use_cross < 0) { const char* e = getenv("DCB_CROSS_ATTN"); use_cross = (e == nullptr) ? 1 : atoi(e); }
    if (use_cross && a->nseg == 2 && a->nk[0] <= 80 && a->nk[1] <= 16) {
        // text + image cross-attention: tcgen05 kernel with the 96 key slots resident in shared memory (attn_tc.cu)
        static int use_ctc = -1;                // DCB_CROSS_TC=0 forces the mma.sync resident-key kernel (A/B testing)
        if (use_ctc < 0) { const char* e = getenv("DCB_CROSS_TC"); use_ctc = (e == nullptr) ? 1 : atoi(e); }
        if (use_ctc) {
            const int rc = cross_tc_launch(a, st);
            if (rc != DCB_ERR_UNSUPPORTED) return rc;
        }
        // mma.sync variant: all keys resident in shared memory, warps stream over the queries
        constexpr int NT0 = 10, NT1 = 2;
        constexpr int smem = (2 * (NT0 + NT1) * 8 + kCrossWarps * kCrossStages * 16) * kD * 2;
        static bool cfg[kMaxDevices] = {};
        const int dv = current_device() & (kMaxDevices - 1);
        if (!cfg[dv]) {
            cudaFuncSetAttribute(cross_attention_kernel<NT0, NT1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cfg[dv] = true;
        }
        const int tiles = (a->nq + 15) / 16;
        int tpw = (tiles + kCrossWarps - 1) / kCrossWarps;
        tpw = tpw > 5 ? 5 : tpw;
        dim3 grid((tiles + kCrossWarps * tpw - 1) / (kCrossWarps * tpw), a->heads, a->batch);
        launch_pdl(kPdlAttention, cross_attention_kernel<NT0, NT1>, grid, dim3(kCrossWarps * 32), smem, st, p, tpw);
        return check_launch("cross_attention");
    }
    if (a->nq >= 512) {
        constexpr int W = 8, smem = (W * 16 + 4 * 64) * kD * 2;
        static bool cfg[kMaxDevices] = {};
        const int dv = current_device() & (kMaxDevices - 1);
        if (!cfg[dv]) {
            cudaFuncSetAttribute(flash_attention_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cfg[dv] = true;
        }
        dim3 grid((a->nq + W * 16 - 1) / (W * 16), a->heads, a->batch);
        launch_pdl(kPdlAttention, flash_attention_kernel<W>, grid, dim3(W * 32), smem, st, p);
    } else {
        constexpr int W = 4, smem = (W * 16 + 4 * 64) * kD * 2;
        dim3 grid((a->nq + W * 16 - 1) / (W * 16), a->heads, a->batch);
        launch_pdl(kPdlAttention, flash_attention_kernel<W>, grid, dim3(W * 32), smem, st, p);
    }
    return check_launch("flash_attention");
}

extern "C" int dcb_temporal_attention(const void* q, const void* k, const void* v, void* o, long long ldq,
                                      long long ldk, long long ldv, long long ldo, int B, int L, int HW, int heads,
                                      int head_dim, float softmax_scale, void* stream) {
    if (!q || !k || !v || !o || B < 1 || HW < 1 || heads < 1) return set_error(DCB_ERR_INVALID, "temporal_attention: bad argument");
    if (head_dim != kD || L < 1 || L > 16) return set_error(DCB_ERR_UNSUPPORTED, "temporal_attention: head_dim 64 and <= 16 frames only");
    if ((ldq & 7) || (ldk & 7) || (ldv & 7) || (ldo & 7)) return set_error(DCB_ERR_INVALID, "temporal_attention: alignment");
    {
        static int use_ttc = -1;                // DCB_TEMPORAL_TC=0 forces the mma.sync kernel (A/B testing)
        if (use_ttc < 0) { const char* e = getenv("DCB_TEMPORAL_TC"); use_ttc = (e == nullptr) ? 1 : atoi(e); }
        if (use_ttc) {
            const int rc = temporal_tc_launch(q, k, v, o, ldq, ldk, ldv, ldo, B, L, HW, heads, softmax_scale,
                                              static_cast<cudaStream_t>(stream));
            if (rc != DCB_ERR_UNSUPPORTED) return rc;
        }
    }
    TemporalParams p;
    p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k);
    p.v = static_cast<const __nv_bfloat16*>(v); p.o = static_cast<__nv_bfloat16*>(o);
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
    p.B = B; p.L = L; p.HW = HW; p.heads = heads;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    const long long items = static_cast<long long>(B) * HW * heads;
    launch_pdl(kPdlAttention, temporal_attention_kernel, dim3(static_cast<unsigned>((items + 3) / 4)), dim3(128), 0,
               static_cast<cudaStream_t>(stream), p);
    return check_launch("temporal_attention");
}

extern "C" int dcb_softmax_rows(const float* x, long long rows, int n, long long ldx, float scale, void* y_bf16,
                                long long ldy, void* stream) {
    if (!x || !y_bf16 || rows < 1 || n < 1) return set_error(DCB_ERR_INVALID, "softmax_rows: bad argument");
    softmax_rows_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, rows, n, ldx, scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(y_bf16), ldy);
    return check_launch("softmax_rows");
}
