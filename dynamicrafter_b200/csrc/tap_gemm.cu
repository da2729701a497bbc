// dcb_tap_gemm: persistent, warp-specialised tcgen05 implicit GEMM for sm_100a.
//
//   out[r, n] = epilogue( sum_{tap} sum_{c} A[r shifted by tap, c] * W[n, tap*C + c] )
//
// One kernel serves nn.Linear, 1x1 conv, 3x3 conv (9 taps over y/x) and the (3,1,1) temporal conv
// (3 taps over frames): the A operand is fetched by TMA as a box of a 4-D channels-last tensor and a
// tap is just a coordinate shift of that box; TMA's out-of-bounds zero fill is the conv padding.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (TMEM -> registers -> fused bias / embedding / GEGLU / residual -> HBM).
// Pipelines: smem ring (full/empty mbarriers, 6 stages of {A 128x64, B BNx64} bf16, SWIZZLE_128B)
// and a 2-deep TMEM accumulator ring so the epilogue of tile i overlaps the MMAs of tile i+1.
// Grid = min(#tiles, #SMs); tiles are strided over CTAs with N fastest so an A tile is re-read from L2.
#include "dcb_internal.h"
#include "ptx.cuh"

namespace dcb {

constexpr int kStages = 6;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kThreads = 192;
constexpr int kABytes = kBlockM * kBlockK * 2;

struct TapGemmParams {
    int D1, D2, D3, b1, b2, b3, t1, t2, t3;
    int num_m_tiles, num_n_tiles, rows_per_tile;
    int kchunks, ntaps;
    int taps[DCB_MAX_TAPS][3];
    int N, N_out;
    const float* bias;
    const float* group_bias;
    int rows_per_group, ld_group_bias;
    const float* residual;
    int ld_residual;
    int geglu;
    float* out_f32;
    __nv_bfloat16* out_bf16;
    int ld_out;
    float out_scale;
    uint32_t stage_bytes;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const TapGemmParams p) {
    constexpr int kBBytes = BN * kBlockK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr int kAccStride = (BN <= 128) ? 128 : 256;
    constexpr int kTmemCols = 2 * kAccStride;
    constexpr uint32_t kIdesc = make_idesc_bf16(kBlockM, BN);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tfull = bars + 2 * kStages;
    uint64_t* tempty = bars + 2 * kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int kblocks = p.ntaps * p.kchunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.num_n_tiles;
                const int n_tile = tile - m_tile * p.num_n_tiles;
                const int i1 = (m_tile % p.t1) * p.b1;
                const int i2 = ((m_tile / p.t1) % p.t2) * p.b2;
                const int i3 = (m_tile / (p.t1 * p.t2)) * p.b3;
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const int c1 = i1 + p.taps[tap][0], c2 = i2 + p.taps[tap][1], c3 = i3 + p.taps[tap][2];
                    for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full[s], p.stage_bytes);
                        uint8_t* sa = smem + s * kStageBytes;
                        tma_load_4d(sa, &tmA, &full[s], kc * kBlockK, c1, c2, c3);
                        tma_load_2d(sa + kABytes, &tmB, &full[s], (tap * p.kchunks + kc) * kBlockK, n_tile * BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            uint32_t it = 0, tc = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tc) {
                const uint32_t a = tc & 1, aph = (tc >> 1) & 1;
                mbar_wait(&tempty[a], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * kAccStride;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * kStageBytes);
                    const uint64_t adesc = make_sw128_kmajor_desc(sa);
                    const uint64_t bdesc = make_sw128_kmajor_desc(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) address field
                        umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[s]);       // frees the smem stage when these MMAs retire
                }
                umma_commit(&tfull[a]);           // accumulator complete -> epilogue
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps 2..5
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;              // row of the tile owned by this thread
        constexpr int kCols = 16;
        uint32_t tc = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tc) {
            const int m_tile = tile / p.num_n_tiles;
            const int n_tile = tile - m_tile * p.num_n_tiles;
            const int j1 = r % p.b1, j2 = (r / p.b1) % p.b2, j3 = r / (p.b1 * p.b2);
            const int g1 = (m_tile % p.t1) * p.b1 + j1;
            const int g2 = ((m_tile / p.t1) % p.t2) * p.b2 + j2;
            const int g3 = (m_tile / (p.t1 * p.t2)) * p.b3 + j3;
            const bool row_ok = (r < p.rows_per_tile) && g1 < p.D1 && g2 < p.D2 && g3 < p.D3;
            const long long grow = (static_cast<long long>(g3) * p.D2 + g2) * p.D1 + g1;
            const float* gb = nullptr;
            if (p.group_bias != nullptr && row_ok)
                gb = p.group_bias + (grow / p.rows_per_group) * p.ld_group_bias;

            const uint32_t a = tc & 1, aph = (tc >> 1) & 1;
            mbar_wait(&tfull[a], aph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + a * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
            const int ncols = p.geglu ? BN / 2 : BN;
            const int out_n0 = n_tile * ncols;

            for (int j0 = 0; j0 < ncols; j0 += kCols) {
                uint32_t v[kCols], g[kCols];
                tmem_ld16(t_row + j0, v);
                if (p.geglu) tmem_ld16(t_row + BN / 2 + j0, g);
                tmem_ld_wait();
                const int col0 = out_n0 + j0;
                if (row_ok && col0 < p.N_out) {
                    float o[kCols];
#pragma unroll
                    for (int j = 0; j < kCols; ++j) o[j] = __uint_as_float(v[j]) * p.out_scale;
                    const bool full_chunk = (col0 + kCols <= p.N_out);
                    if (p.geglu) {
                        const int wn = n_tile * BN + j0;   // packed weight column of the value half
#pragma unroll
                        for (int j = 0; j < kCols; ++j) {
                            float val = o[j], gate = __uint_as_float(g[j]) * p.out_scale;
                            if (p.bias != nullptr) {
                                val += p.bias[wn + j];
                                gate += p.bias[wn + BN / 2 + j];
                            }
                            o[j] = val * gelu_erf(gate);
                        }
                    } else if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < kCols; ++j)
                            if (full_chunk || col0 + j < p.N_out) o[j] += p.bias[col0 + j];
                    }
                    if (gb != nullptr) {
#pragma unroll
                        for (int j = 0; j < kCols; ++j)
                            if (full_chunk || col0 + j < p.N_out) o[j] += gb[col0 + j];
                    }
                    if (p.residual != nullptr) {
                        const float* rp = p.residual + grow * p.ld_residual + col0;
                        if (full_chunk && (p.ld_residual & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < kCols; j += 4) {
                                const float4 t = *reinterpret_cast<const float4*>(rp + j);
                                o[j] += t.x; o[j + 1] += t.y; o[j + 2] += t.z; o[j + 3] += t.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < kCols; ++j)
                                if (col0 + j < p.N_out) o[j] += rp[j];
                        }
                    }
                    if (p.out_f32 != nullptr) {
                        float* op = p.out_f32 + grow * p.ld_out + col0;
                        if (full_chunk && (p.ld_out & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < kCols; j += 4)
                                *reinterpret_cast<float4*>(op + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < kCols; ++j)
                                if (col0 + j < p.N_out) op[j] = o[j];
                        }
                    }
                    if (p.out_bf16 != nullptr) {
                        __nv_bfloat16* op = p.out_bf16 + grow * p.ld_out + col0;
                        if (full_chunk && (p.ld_out & 7) == 0) {
#pragma unroll
                            for (int j = 0; j < kCols; j += 8) {
                                uint4 pk;
                                __nv_bfloat162 h0 = __floats2bfloat162_rn(o[j], o[j + 1]);
                                __nv_bfloat162 h1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]);
                                __nv_bfloat162 h3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
                                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                                pk.z = *reinterpret_cast<uint32_t*>(&h2);
                                pk.w = *reinterpret_cast<uint32_t*>(&h3);
                                *reinterpret_cast<uint4*>(op + j) = pk;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < kCols; ++j)
                                if (col0 + j < p.N_out) op[j] = __float2bfloat16_rn(o[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[a]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_elems, const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return set_error(DCB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_elems[i] * 2;
    }
    CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return set_error(DCB_ERR_CUDA, "cuTensorMapEncodeTiled failed (rc=%d)", (int)rc);
    return DCB_OK;
}

template <int BN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, TapGemmParams& p, cudaStream_t stream) {
    constexpr int kStageBytes = kABytes + BN * kBlockK * 2;
    constexpr int smem_bytes = kStages * kStageBytes + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tap_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             smem_bytes);
        if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    p.stage_bytes = static_cast<uint32_t>(p.rows_per_tile * kBlockK * 2 + BN * kBlockK * 2);
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    tap_gemm_kernel<BN><<<grid, kThreads, smem_bytes, stream>>>(tmA, tmB, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "tap_gemm launch: %s", cudaGetErrorString(e));
    return DCB_OK;
}

}  // namespace dcb

using namespace dcb;

extern "C" int dcb_tap_gemm(const dcb_tap_gemm_args* a, void* stream) {
    if (a == nullptr || a->a == nullptr || a->w == nullptr) return set_error(DCB_ERR_INVALID, "tap_gemm: null argument");
    if (a->out_f32 == nullptr && a->out_bf16 == nullptr) return set_error(DCB_ERR_INVALID, "tap_gemm: no output");
    if (a->C <= 0 || a->C % 8 != 0 || (a->ntaps > 1 && a->C % 64 != 0))
        return set_error(DCB_ERR_INVALID, "tap_gemm: C=%d must be a multiple of 8 (of 64 when ntaps > 1)", a->C);
    const int64_t w_ld = a->w_ld > 0 ? a->w_ld : static_cast<int64_t>(a->ntaps) * a->C;
    if (w_ld % 8 != 0 || w_ld < static_cast<int64_t>(a->ntaps) * a->C)
        return set_error(DCB_ERR_INVALID, "tap_gemm: w_ld");
    if (a->ntaps < 1 || a->ntaps > DCB_MAX_TAPS) return set_error(DCB_ERR_INVALID, "tap_gemm: ntaps=%d", a->ntaps);
    if (a->b1 < 1 || a->b2 < 1 || a->b3 < 1 || a->b1 > 256 || a->b2 > 256 || a->b3 > 256 ||
        a->b1 * a->b2 * a->b3 > kBlockM)
        return set_error(DCB_ERR_INVALID, "tap_gemm: box %dx%dx%d must hold <= 128 rows", a->b1, a->b2, a->b3);
    if (a->D1 < 1 || a->D2 < 1 || a->D3 < 1 || a->A1 < a->D1 || a->A2 < a->D2 || a->A3 < a->D3)
        return set_error(DCB_ERR_INVALID, "tap_gemm: bad domain");
    if ((a->a_s1 % 8) || (a->a_s2 % 8) || (a->a_s3 % 8) || (reinterpret_cast<uintptr_t>(a->a) & 15) ||
        (reinterpret_cast<uintptr_t>(a->w) & 15))
        return set_error(DCB_ERR_INVALID, "tap_gemm: A/W need 16-byte aligned base and strides");
    if (a->N < 1) return set_error(DCB_ERR_INVALID, "tap_gemm: N");
    int bn = a->block_n;
    if (bn == 0) bn = (a->N % 160 == 0) ? 160 : 128;
    if (bn != 128 && bn != 160) return set_error(DCB_ERR_UNSUPPORTED, "tap_gemm: block_n must be 128 or 160");
    if (a->geglu && (a->N % bn != 0)) return set_error(DCB_ERR_INVALID, "tap_gemm: geglu needs N %% block_n == 0");
    if (a->group_bias != nullptr && a->rows_per_group < 1) return set_error(DCB_ERR_INVALID, "tap_gemm: rows_per_group");

    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->A1, (uint64_t)a->A2, (uint64_t)a->A3};
        const uint64_t str[4] = {1, (uint64_t)a->a_s1, (uint64_t)a->a_s2, (uint64_t)a->a_s3};
        const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)a->b1, (uint32_t)a->b2, (uint32_t)a->b3};
        int rc = encode_bf16_map(&tmA, a->a, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        const uint64_t K = (uint64_t)a->ntaps * a->C;
        const uint64_t dims[2] = {K, (uint64_t)a->N};
        const uint64_t str[2] = {1, (uint64_t)w_ld};
        const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)bn};
        int rc = encode_bf16_map(&tmB, a->w, 2, dims, str, box);
        if (rc) return rc;
    }
    TapGemmParams p;
    p.D1 = a->D1; p.D2 = a->D2; p.D3 = a->D3;
    p.b1 = a->b1; p.b2 = a->b2; p.b3 = a->b3;
    p.t1 = (a->D1 + a->b1 - 1) / a->b1;
    p.t2 = (a->D2 + a->b2 - 1) / a->b2;
    p.t3 = (a->D3 + a->b3 - 1) / a->b3;
    p.num_m_tiles = p.t1 * p.t2 * p.t3;
    p.num_n_tiles = (a->N + bn - 1) / bn;
    p.rows_per_tile = a->b1 * a->b2 * a->b3;
    p.kchunks = (a->C + 63) / 64;
    p.ntaps = a->ntaps;
    for (int t = 0; t < DCB_MAX_TAPS; ++t)
        for (int d = 0; d < 3; ++d) p.taps[t][d] = (t < a->ntaps) ? a->taps[t][d] : 0;
    p.N = a->N;
    p.N_out = a->geglu ? a->N / 2 : a->N;
    p.bias = a->bias;
    p.group_bias = a->group_bias;
    p.rows_per_group = a->rows_per_group;
    p.ld_group_bias = a->ld_group_bias;
    p.residual = a->residual;
    p.ld_residual = a->ld_residual;
    p.geglu = a->geglu;
    p.out_f32 = a->out_f32;
    p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16);
    p.ld_out = a->ld_out;
    p.out_scale = a->out_scale;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return bn == 160 ? launch<160>(tmA, tmB, p, s) : launch<128>(tmA, tmB, p, s);
}
