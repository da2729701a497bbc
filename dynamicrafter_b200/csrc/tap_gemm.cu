// dcb_tap_gemm: persistent, warp-specialised tcgen05 implicit GEMM for sm_100a.
//
//   out[r, n] = epilogue( sum_{tap} sum_{c} A[r shifted by tap, c] * W[n, tap*C + c] )
//
// One kernel serves nn.Linear, 1x1 conv, 3x3 conv (9 taps over y/x) and the (3,1,1) temporal conv
// (3 taps over frames): the A operand is fetched by TMA as a box of a 4-D channels-last tensor and a
// tap is just a coordinate shift of that box; TMA's out-of-bounds zero fill is the conv padding.
//
// Host entry point: argument validation, tensor maps, tile-shape / launch-mode heuristics.  The kernel and its warp
// roles are described in tap_gemm_kernel.cuh.
#include "tap_gemm_kernel.cuh"

namespace dcb {

// statistics-emitting instantiations live in their own translation unit (tap_gemm_cs.cu): the plain kernels carry
// no code or register cost for the fused GroupNorm statistics
int tap_gemm_launch_cs(int bn, bool out_bf16, bool pair, const CUtensorMap& tmA, const CUtensorMap* tmB,
                       const CUtensorMap& tmOut, const CUtensorMap& tmRes, TapGemmParams& p, cudaStream_t s);

int tap_gemm_launch_ln(int bn, bool out_bf16, bool pair, const CUtensorMap& tmA, const CUtensorMap* tmB,
                       const CUtensorMap& tmOut, const CUtensorMap& tmRes, TapGemmParams& p, cudaStream_t s);

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

static int encode_map(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, CUtensorMapSwizzle swz, const void* base,
                      int rank, const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* box,
                      CUtensorMapL2promotion promo) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return set_error(DCB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_elems[i] * elem_bytes;
    }
    CUresult rc = fn(m, dt, rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return set_error(DCB_ERR_CUDA, "cuTensorMapEncodeTiled failed (rc=%d)", (int)rc);
    return DCB_OK;
}

static int pair_mode_env() {       // DCB_GEMM_PAIR=0 forces single-CTA tiles, =1 forces CTA pairs, unset: heuristic
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("DCB_GEMM_PAIR");
        v = (e == nullptr) ? -1 : atoi(e);
    }
    return v;
}

}  // namespace dcb

using namespace dcb;

extern "C" int dcb_tap_gemm(const dcb_tap_gemm_args* a, void* stream) {
    if (a == nullptr || a->a == nullptr || a->w == nullptr) return set_error(DCB_ERR_INVALID, "tap_gemm: null argument");
    if ((a->out_f32 == nullptr) == (a->out_bf16 == nullptr))
        return set_error(DCB_ERR_INVALID, "tap_gemm: exactly one of out_f32 / out_bf16 must be given");
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
    if (bn == 0 && !a->geglu) {
        static int env_bn = -1;                      // DCB_GEMM_BN: tuning override (tools/bench_kernels.py)
        if (env_bn < 0) { const char* e = getenv("DCB_GEMM_BN"); env_bn = e ? atoi(e) : 0; }
        if (env_bn > 0 && (a->N % env_bn == 0 || env_bn == 192 || env_bn == 256) && a->N >= env_bn) bn = env_bn;
    }
    if (bn == 0) {
        const long long m_tiles = static_cast<long long>((a->D1 + a->b1 - 1) / a->b1) * ((a->D2 + a->b2 - 1) / a->b2) *
                                  ((a->D3 + a->b3 - 1) / a->b3);
        if (a->geglu) {
            bn = (a->N % 256 == 0) ? 256 : 128;
        } else {
            // Tile-width model (CTA pairs, measured with tools/bench_shapes.py, profiles/r02_summary.md): a 64-deep stage
            // costs a CTA ~(256 + BN) clk -- 16 KB of A plus 64 B per column of B against ~64 B/clk of L2 -> SM fill
            // bandwidth; only BN = 256 keeps up with the tensor pipe -- and a launch runs ceil(items / pairs) rounds of
            // such items.  Narrower tiles pay more fill per FLOP but can quantise better (M = 5120, N = 1280: 100 items
            // of 256 columns are 2 rounds on 74 pairs, 140 items of 192 columns are 2 cheaper rounds: -12..18 %).  A
            // ragged last tile (narrow UMMA, fewer B rows) makes every width legal for every N; it gets a 3 % handicap
            // because the narrow tile still stages a full A tile.  DCB_GEMM_RAGGED256=0 restores the round-1 rule.
            static int model = -1;
            if (model < 0) { const char* e = getenv("DCB_GEMM_RAGGED256"); model = e ? atoi(e) : 1; }
            const long long pairs = num_sms() / 2, m_items = (m_tiles + 1) / 2;
            if (!model || m_items * ((a->N + 255) / 256) < 24) {
                if (a->N % 256 == 0 && m_tiles * (a->N / 256) >= 1LL * num_sms()) bn = 256;
                else bn = (a->N % 160 == 0) ? 160 : 128;
            } else {
                const int cand[4] = {256, 192, 160, 128};
                double best = 0.0;
                for (int i = 0; i < 4; ++i) {
                    const int c = cand[i];
                    if (c == 160 && a->N % 160 != 0) continue;                  // 160 only where it divides (320 / 640 / 960 ...)
                    const long long items = m_items * ((a->N + c - 1) / c);
                    double t = static_cast<double>((items + pairs - 1) / pairs) * (256 + c);
                    if (a->N % c != 0) t *= 1.03;
                    if (bn == 0 || t < best) { bn = c; best = t; }
                }
            }
        }
    }
    if (bn != 128 && bn != 160 && bn != 192 && bn != 256)
        return set_error(DCB_ERR_UNSUPPORTED, "tap_gemm: block_n must be 128, 160, 192 or 256");
    // CTA pairs (cta_group::2): each SM stages half of B, halving shared-memory traffic per MMA.  Measured on the
    // UNet's shapes: pairs win down to ~24 work items of 256 rows (M = 1280 layers: 40 items, +10..20 %; whole step -0.5 ms); below that
    // single 128-row tiles spread over more SMs.
    bool pair;
    {
        const long long m_tiles = static_cast<long long>((a->D1 + a->b1 - 1) / a->b1) * ((a->D2 + a->b2 - 1) / a->b2) *
                                  ((a->D3 + a->b3 - 1) / a->b3);
        const long long items = ((m_tiles + 1) / 2) * ((a->N + bn - 1) / bn);
        static int min_items = -1;                   // DCB_GEMM_PAIR_MIN: tuning override
        if (min_items < 0) { const char* e = getenv("DCB_GEMM_PAIR_MIN"); min_items = e ? atoi(e) : 24; }
        pair = items >= min_items;
        if (pair_mode_env() == 0) pair = false;
        if (pair_mode_env() == 1) pair = true;
    }
    if (a->geglu && ((bn != 128 && bn != 256) || a->N % bn != 0))
        return set_error(DCB_ERR_INVALID, "tap_gemm: geglu needs block_n in {128, 256} dividing N");
    if (a->group_bias != nullptr && a->rows_per_group < 1) return set_error(DCB_ERR_INVALID, "tap_gemm: rows_per_group");
    const bool obf = a->out_bf16 != nullptr;
    const void* outp = obf ? a->out_bf16 : static_cast<const void*>(a->out_f32);
    const int oes = obf ? 2 : 4;
    if ((reinterpret_cast<uintptr_t>(outp) & 15) || (static_cast<int64_t>(a->ld_out) * oes) % 16 != 0)
        return set_error(DCB_ERR_INVALID, "tap_gemm: out needs a 16-byte aligned base and row pitch");
    if (a->residual != nullptr &&
        ((reinterpret_cast<uintptr_t>(a->residual) & 15) || (static_cast<int64_t>(a->ld_residual) * 4) % 16 != 0))
        return set_error(DCB_ERR_INVALID, "tap_gemm: residual needs a 16-byte aligned base and row pitch");
    const int n_out = a->geglu ? a->N / 2 : a->N;

    CUtensorMap tmA, tmBs[2], tmOut, tmRes;
    CUtensorMap& tmB = tmBs[0];
    {
        const uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->A1, (uint64_t)a->A2, (uint64_t)a->A3};
        const uint64_t str[4] = {1, (uint64_t)a->a_s1, (uint64_t)a->a_s2, (uint64_t)a->a_s3};
        const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)a->b1, (uint32_t)a->b2, (uint32_t)a->b3};
        int rc = encode_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_128B, a->a, 4, dims, str, box,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    {
        const uint64_t K = (uint64_t)a->ntaps * a->C;
        const uint64_t dims[2] = {K, (uint64_t)a->N};
        const uint64_t str[2] = {1, (uint64_t)w_ld};
        const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)(pair ? bn / 2 : bn)};
        int rc = encode_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_128B, a->w, 2, dims, str, box,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
        // the (possibly narrow) last N tile: same layout, fewer rows per box (mirrors launch<>: n_eff = valid columns
        // rounded up to 32, full width for GEGLU)
        const int rem = a->N - ((a->N + bn - 1) / bn - 1) * bn;
        int n_eff = (rem + 31) / 32 * 32;
        if (n_eff > bn || a->geglu) n_eff = bn;
        const uint32_t box2[2] = {(uint32_t)kBlockK, (uint32_t)(pair ? n_eff / 2 : n_eff)};
        rc = encode_map(&tmBs[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_128B, a->w, 2, dims, str, box2,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    {
        const uint64_t ld = (uint64_t)a->ld_out;
        const bool strided = a->out_s1 != 0 || a->out_s2 != 0 || a->out_s3 != 0;
        if (strided) {
            if (a->out_s1 <= 0 || a->out_s2 <= 0 || a->out_s3 <= 0 || (a->out_s1 * oes) % 16 || (a->out_s2 * oes) % 16 ||
                (a->out_s3 * oes) % 16 || a->residual != nullptr || a->ln_acc != nullptr)
                return set_error(DCB_ERR_INVALID, "tap_gemm: output strides must be positive multiples of 16 bytes (no residual / LayerNorm folding)");
        }
        const uint64_t dims[4] = {(uint64_t)n_out, (uint64_t)a->D1, (uint64_t)a->D2, (uint64_t)a->D3};
        const uint64_t str[4] = {1, strided ? (uint64_t)a->out_s1 : ld, strided ? (uint64_t)a->out_s2 : ld * a->D1,
                                 strided ? (uint64_t)a->out_s3 : ld * a->D1 * a->D2};
        const uint32_t box[4] = {(uint32_t)(obf ? 32 : 16), (uint32_t)a->b1, (uint32_t)a->b2, (uint32_t)a->b3};
        int rc = encode_map(&tmOut, obf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, oes,
                            CU_TENSOR_MAP_SWIZZLE_64B, outp, 4, dims, str, box,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE);
        if (rc) return rc;
    }
    tmRes = tmOut;
    if (a->residual != nullptr && !obf) {
        const uint64_t ld = (uint64_t)a->ld_residual;
        const uint64_t dims[4] = {(uint64_t)n_out, (uint64_t)a->D1, (uint64_t)a->D2, (uint64_t)a->D3};
        const uint64_t str[4] = {1, ld, ld * a->D1, ld * a->D1 * a->D2};
        const uint32_t box[4] = {16u, (uint32_t)a->b1, (uint32_t)a->b2, (uint32_t)a->b3};
        int rc = encode_map(&tmRes, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, CU_TENSOR_MAP_SWIZZLE_64B, a->residual, 4, dims,
                            str, box, CU_TENSOR_MAP_L2_PROMOTION_NONE);
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
    p.N_out = n_out;
    p.bias = a->bias;
    p.group_bias = a->group_bias;
    p.rows_per_group = a->rows_per_group;
    p.ld_group_bias = a->ld_group_bias;
    p.has_residual = a->residual != nullptr ? 1 : 0;
    p.residual = a->residual;
    p.ld_residual = a->ld_residual;
    p.geglu = a->geglu;
    p.out_bf16 = obf ? 1 : 0;
    p.out_scale = a->out_scale;
    p.gn_acc = a->gn_acc;
    p.gn_cg = a->gn_cg;
    p.gn_unit_mode = a->gn_unit_mode;
    p.seg_rows = a->b1 * a->b2;                                   // tile rows per D3 slice
    p.seg_lw = p.seg_rows < 32 ? p.seg_rows : 32;                 // rows folded inside one warp per unit
    if (a->gn_acc != nullptr) {
        const bool seg_ok = p.seg_rows == 8 || p.seg_rows == 16 || p.seg_rows % 32 == 0 || (a->b3 == 1 && !a->gn_unit_mode);
        if (a->geglu || !seg_ok || (a->gn_unit_mode && a->b2 != 1) || a->gn_cg < 1 || n_out % a->gn_cg != 0 ||
            n_out / a->gn_cg != 32 || (reinterpret_cast<uintptr_t>(a->gn_acc) & 7))
            return set_error(DCB_ERR_UNSUPPORTED, "tap_gemm: fused GroupNorm statistics need 32 groups of gn_cg columns, no geglu and "
                             "warp segments inside one unit (b1*b2 = %d, b3 = %d, mode %d)", p.seg_rows, a->b3, a->gn_unit_mode);
        if (p.seg_rows != 8 && p.seg_rows != 16) p.seg_lw = 32;
    }
    p.fd_t1 = FastDiv::make(p.t1);
    p.fd_t2 = FastDiv::make(p.t2);
    p.fd_b1 = FastDiv::make(a->b1);
    p.fd_b2 = FastDiv::make(a->b2);
    p.fd_rpg = FastDiv::make(a->rows_per_group > 0 ? a->rows_per_group : 1);
    p.fd_cg = FastDiv::make(a->gn_cg > 0 ? a->gn_cg : 1);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (a->geglu && !obf) return set_error(DCB_ERR_UNSUPPORTED, "tap_gemm: geglu writes bf16 only");
    if (a->geglu && a->residual != nullptr) return set_error(DCB_ERR_UNSUPPORTED, "tap_gemm: geglu takes no residual");
    p.ln_acc = a->ln_acc;
    p.ln_x = static_cast<__nv_bfloat16*>(a->ln_x_bf16);
    p.ld_ln_x = a->ld_ln_x;
    p.ln_colsum = a->ln_colsum;
    p.ln_inv_c = a->ln_channels > 0 ? 1.0f / static_cast<float>(a->ln_channels) : 0.f;
    p.ln_eps = a->ln_eps;
    if (a->ln_acc != nullptr) {
        if (a->gn_acc != nullptr || (reinterpret_cast<uintptr_t>(a->ln_acc) & 15))
            return set_error(DCB_ERR_INVALID, "tap_gemm: LayerNorm folding excludes GroupNorm statistics; ln_acc must be 16-byte aligned");
        if (!obf) {                                   // producer
            if (a->ln_x_bf16 == nullptr || (reinterpret_cast<uintptr_t>(a->ln_x_bf16) & 15) || a->ld_ln_x % 8 != 0 ||
                a->ld_ln_x < n_out)
                return set_error(DCB_ERR_INVALID, "tap_gemm: LayerNorm producer needs a 16-byte aligned bf16 copy buffer");
        } else {                                      // consumer
            if (a->ln_colsum == nullptr || a->bias == nullptr || a->ln_channels < 1 || a->residual != nullptr ||
                a->group_bias != nullptr || a->out_scale != 1.0f)
                return set_error(DCB_ERR_INVALID, "tap_gemm: LayerNorm consumer needs ln_colsum, the folded bias and ln_channels, "
                                 "and takes no residual / group bias / out_scale");
        }
        return tap_gemm_launch_ln(bn, obf, pair, tmA, tmBs, tmOut, tmRes, p, s);
    }
    if (a->gn_acc != nullptr) return tap_gemm_launch_cs(bn, obf, pair, tmA, tmBs, tmOut, tmRes, p, s);
    return launch_all<0>(bn, obf, pair, tmA, tmBs, tmOut, tmRes, p, s);
}

#ifdef DCB_TRACE
// debug builds only (tools/trace_gemm.py): copy the CTA-0 event log (6 role regions of 4096 entries, entry 0 = count)
extern "C" int dcb_debug_trace(unsigned long long* dst, int cap) {
    cudaDeviceSynchronize();
    if (cap < 6 * 4096) return -1;
    cudaMemcpyFromSymbol(dst, dcb::g_tap_trace, 6 * 4096 * sizeof(unsigned long long));
    return 6 * 4096;
}
#endif
