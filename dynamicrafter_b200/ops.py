"""Python-side launch helpers over the C ABI (device pointers in, nothing allocated by the library).

Tensors are torch CUDA tensors used purely as device-memory handles; every op runs on
``torch.cuda.current_stream()`` so the whole path is CUDA-graph capturable.
"""
import ctypes
import functools
import os

import torch

from . import _lib
from ._lib import TapGemmArgs, check, ptr
from .precision import fp32_mode

BF16 = torch.bfloat16
F32 = torch.float32


try:                                       # raw handle of the current stream: ~0.3 us instead of ~16 us through the
    _raw_stream = torch._C._cuda_getCurrentRawStream          # Stream object (an eager forward asks 2300 times)
    _cur_device = torch._C._cuda_getDevice
except AttributeError:                     # pragma: no cover - older / newer torch without the private accessors
    _raw_stream = _cur_device = None


def _stream():
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(_cur_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.DcbError("dcb200 ops need CUDA tensors (there is no CPU fallback)")


def _divisors_upto(n, cap):
    return [d for d in range(1, min(n, cap) + 1) if n % d == 0]


@functools.lru_cache(maxsize=None)
def choose_box(D1, D2, D3):
    """Pick the M-tile box (b1,b2,b3), b1*b2*b3 <= 128, minimising the number of 128-row tiles."""
    best = None
    c1 = set(_divisors_upto(D1, 128)) | {min(D1, 128)}
    for b1 in sorted(c1, reverse=True):
        cap2 = 128 // b1
        c2 = (set(_divisors_upto(D2, cap2)) | {min(D2, cap2)}) if b1 >= min(D1, 128) or D1 % b1 == 0 else {1}
        for b2 in sorted(c2, reverse=True):
            cap3 = 128 // (b1 * b2)
            c3 = set(_divisors_upto(D3, cap3)) | {min(D3, cap3)}
            for b3 in sorted(c3, reverse=True):
                tiles = -(-D1 // b1) * -(-D2 // b2) * -(-D3 // b3)
                key = (tiles, -b1, -b2)
                if best is None or key < best[0]:
                    best = (key, (b1, b2, b3))
    return best[1]


class ColSums:
    """Fused GroupNorm statistics of a GEMM output, accumulated by the GEMM epilogue: acc = int64 [units, 32, 2] fixed-point
    (sum, sumsq) per unit and group; a unit = `unit_rows` consecutive activation rows (one frame, or one video when the
    producer's tiles span frames).  mode: dcb_tap_gemm_args.gn_unit_mode."""
    __slots__ = ("acc", "unit_rows", "C", "mode")

    def __init__(self, acc, unit_rows, C, mode):
        self.acc, self.unit_rows, self.C, self.mode = acc, unit_rows, C, mode

    def fits(self, S, C):
        """Can a GroupNorm over instances of S rows x C channels be served from these sums?"""
        return C == self.C and S % self.unit_rows == 0


def _unit_mode(dims, b):
    """1: a tile never spans D2 (temporal conv with one frame per tile) -> units = (D3, D2) = frames; else units = D3."""
    return 1 if (dims[1] > 1 and b[1] == 1 and b[2] == 1) else 0


def colsum_supported(dims, N=None):
    b = choose_box(*dims)
    seg = b[0] * b[1]
    if N is not None and N % 32:
        return False
    return seg in (8, 16) or seg % 32 == 0 or b[2] == 1


def stats_enabled():
    """Producer-fused GroupNorm statistics are used in the bf16 mode (in frame-sharded temporal blocks the per-rank sums
    are all-reduced as integers before the apply)."""
    from .lvdm import _packing
    return _packing.FUSED_GN_STATS and not fp32_mode()


class _StatsArena:
    """Zeroed int64 pool the per-launch accumulators are carved from: one memset per forward instead of one per GEMM.
    Buffers that are outgrown are retired, not freed (captured CUDA graphs keep their addresses)."""

    def __init__(self):
        self.buf, self.off, self.active, self.retired = None, 0, False, []


_ARENAS = {}


class stats_scope:
    """with stats_scope(device): every fused-statistics accumulator allocated inside comes from the device's arena.  The
    arena is clean on entry (fresh buffers are zero-filled) and the prefix a forward used is re-zeroed when its scope
    closes -- stream ordered behind the last consumer -- so the memset a CUDA graph captures always covers exactly what
    that forward allocated, also when the arena grew during the warm-up runs.  Not re-entrant: an inner scope is a no-op."""

    def __init__(self, dev):
        self.ar = _ARENAS.setdefault(dev.index, _StatsArena())
        self.owner = False

    def __enter__(self):
        ar = self.ar
        if not ar.active:
            self.owner, ar.active, ar.off = True, True, 0
        return self

    def __exit__(self, *exc):
        if self.owner:
            ar = self.ar
            if ar.buf is not None and ar.off > 0:
                ar.buf[:ar.off].zero_()
            ar.off, ar.active = 0, False


def zero_stats_arena(dev):
    """Re-zero the whole accumulator arena of a device.  Only needed by tools that re-issue statistics-emitting GEMMs
    OUTSIDE a stats_scope (bench.py's class replay): those launches add into accumulators nobody clears afterwards."""
    ar = _ARENAS.get(dev.index)
    if ar is not None and ar.buf is not None:
        ar.buf.zero_()


def _alloc_acc(units, dev):
    return _alloc_i64(units * 64, dev).view(units, 32, 2)


class LnSrc:
    """LayerNorm folding: what the GEMM that produced an fp32 [rows, C] tensor leaves for the GEMM behind the LayerNorm:
    xb = bf16 copy of the tensor, acc = int64 [rows, 2] fixed-point (sum, sum of squares) per row."""
    __slots__ = ("xb", "acc", "C")

    def __init__(self, xb, acc, C):
        self.xb, self.acc, self.C = xb, acc, C


class LnFold:
    """Consumer side of a folded LayerNorm: w = bf16 [N, C] gamma-scaled weights, b = fp32 [N] folded bias,
    s = fp32 [N] row sums of w (as rounded to bf16), eps."""
    __slots__ = ("w", "b", "s", "eps")

    def __init__(self, w, b, s, eps):
        self.w, self.b, self.s, self.eps = w, b, s, eps


def ln_fold_enabled():
    """LayerNorm folded into its neighbour GEMMs (bf16 mode only).  OFF by default (DCB_LN_FOLD=1 enables it): measured on
    B200 (tools/bench_ln_fold.py, profiles/r02_summary.md) the fold removes the layer_norm launches (2.4 ms per step)
    but costs the HBM-bound producer GEMMs +8..20 us each (scattered bf16 copy + per-row atomics) and the issue-bound
    K = 320 consumers +5..17 us: 34.84 vs 34.79 ms per step, i.e. nothing gained."""
    return _LN_FOLD and not fp32_mode()


_LN_FOLD = os.environ.get("DCB_LN_FOLD", "0") == "1"


def _alloc_i64(n, dev):
    """n zeroed int64 (n even: 16-byte aligned) from the per-forward arena, or a fresh tensor outside a stats_scope."""
    ar = _ARENAS.get(dev.index)
    if ar is not None and ar.active:
        if ar.buf is None or ar.off + n > ar.buf.numel():
            if ar.buf is not None:
                ar.retired.append(ar.buf)
            cap = max(2 * (ar.buf.numel() if ar.buf is not None else 0), 1 << 20, n)
            ar.buf, ar.off = torch.zeros(cap, dtype=torch.int64, device=dev), 0
        out = ar.buf[ar.off:ar.off + n]
        ar.off += n
        return out
    return torch.zeros(n, dtype=torch.int64, device=dev)


def _make_colsums(dims, N, dev):
    b = choose_box(*dims)
    mode = _unit_mode(dims, b)
    if mode:
        return ColSums(_alloc_acc(dims[1] * dims[2], dev), dims[0], N, 1)
    return ColSums(_alloc_acc(dims[2], dev), dims[0] * dims[1], N, 0)


def tap_gemm(a, dims, strides, adims, taps, w, N, *, bias=None, group_bias=None, rows_per_group=0,
             residual=None, geglu=False, out_f32=None, out_bf16=None, ld_out=None, out_scale=1.0,
             block_n=0, box=None, colsum=None, ln_acc=None, ln_x=None, ln_colsum=None, ln_channels=0, ln_eps=0.0,
             out_strides=None):
    """Raw launch.  a: bf16 tensor; dims=(D1,D2,D3) output domain; strides=(s1,s2,s3) element strides
    of A; adims=(A1,A2,A3) extents of A; taps=[(o1,o2,o3),...]; w: bf16 [N, ntaps*C]."""
    require_cuda(a, w, bias, group_bias, residual, out_f32, out_bf16)
    if a.dtype == F32:
        # fp32 parity mode: A -> [hi | lo | hi] (3C bf16 columns per row); w was packed as [hi | hi | lo] per tap
        assert fp32_mode() and a.stride(1) == 1 and a.dim() == 2 and out_bf16 is None
        C0, pitch = a.shape[1], a.stride(0)
        assert w.shape[1] == 3 * C0 * len(taps), "weights are not packed for the fp32 mode"
        assert all(s_ % pitch == 0 for s_ in strides)
        a3 = torch.empty((a.shape[0], 3 * C0), dtype=BF16, device=a.device)
        check(_lib.lib().dcb_split_bf16x3(ptr(a), _c_ll(a.shape[0]), _c_i(C0), _c_ll(pitch), ptr(a3), _stream()),
              "dcb_split_bf16x3")
        strides = tuple(s_ // pitch * 3 * C0 for s_ in strides)
        a = a3
    assert a.dtype == BF16 and w.dtype == BF16 and w.stride(1) == 1
    C = w.shape[1] // len(taps)
    args = TapGemmArgs()
    args.a = a.data_ptr()
    args.a_s1, args.a_s2, args.a_s3 = strides
    args.C = C
    args.D1, args.D2, args.D3 = dims
    args.A1, args.A2, args.A3 = adims
    b = box or choose_box(*dims)
    args.b1, args.b2, args.b3 = b
    args.ntaps = len(taps)
    for i, t in enumerate(taps):
        for d in range(3):
            args.taps[i][d] = t[d]
    args.w = w.data_ptr()
    args.w_ld = w.stride(0)
    args.N = N
    args.block_n = block_n
    args.bias = None if bias is None else bias.data_ptr()
    args.group_bias = None if group_bias is None else group_bias.data_ptr()
    args.rows_per_group = rows_per_group
    args.ld_group_bias = 0 if group_bias is None else group_bias.stride(0)
    args.residual = None if residual is None else residual.data_ptr()
    args.ld_residual = 0 if residual is None else residual.stride(0)
    args.geglu = 1 if geglu else 0
    args.out_f32 = None if out_f32 is None else out_f32.data_ptr()
    args.out_bf16 = None if out_bf16 is None else out_bf16.data_ptr()
    args.ld_out = ld_out
    args.out_scale = out_scale
    args.gn_acc = None if colsum is None else colsum.acc.data_ptr()
    args.gn_cg = 0 if colsum is None else colsum.C // 32
    args.gn_unit_mode = 0 if colsum is None else colsum.mode
    args.ln_acc = None if ln_acc is None else ln_acc.data_ptr()
    args.ln_x_bf16 = None if ln_x is None else ln_x.data_ptr()
    args.ld_ln_x = 0 if ln_x is None else ln_x.stride(0)
    args.ln_channels = ln_channels
    args.ln_colsum = None if ln_colsum is None else ln_colsum.data_ptr()
    args.ln_eps = ln_eps
    args.out_s1, args.out_s2, args.out_s3 = out_strides or (0, 0, 0)
    check(_lib.lib().dcb_tap_gemm(ctypes.byref(args), _stream()), "dcb_tap_gemm")


def _outputs(rows, n_out, out, want_f32, want_bf16, dev):
    of = ob = None
    if fp32_mode() and out is None:
        want_f32, want_bf16 = True, False
    if out is not None:
        if out.dtype == F32:
            of = out
        else:
            ob = out
        return of, ob, out.stride(0)
    assert want_f32 != want_bf16, "exactly one output dtype per launch"
    # the TMA store needs a 16-byte row pitch: pad the leading dimension, hand back a [:, :n_out] view
    if want_f32:
        ld = (n_out + 3) // 4 * 4
        of = torch.empty((rows, ld), dtype=F32, device=dev)[:, :n_out]
    else:
        ld = (n_out + 7) // 8 * 8
        ob = torch.empty((rows, ld), dtype=BF16, device=dev)[:, :n_out]
    return of, ob, ld


def linear(a, w, *, bias=None, residual=None, geglu=False, f32=False, bf16=True, out=None, N=None,
           group_bias=None, rows_per_group=0, out_scale=1.0, block_n=0, stats_rows=0, ln_stats=False, ln=None):
    """a: bf16 [M, K] (row stride may exceed K); w: bf16 [N, K] -> [M, N] (or N/2 for geglu).
    stats_rows = rows per frame: also emit the per-frame column sums a consumer GroupNorm needs -> (out, ColSums).
    ln_stats: fp32 output that feeds a LayerNorm -> (out, LnSrc) (bf16 copy + per-row statistics from the epilogue).
    ln: LnFold, with a = LnSrc: the GEMM behind a folded LayerNorm (w / bias are taken from the LnFold)."""
    if ln is not None:
        src = a
        assert isinstance(src, LnSrc) and not f32 and residual is None and out is None and not stats_rows
        M, K = src.xb.shape
        N = ln.w.shape[0]
        n_out = N // 2 if geglu else N
        of, ob, ld = _outputs(M, n_out, None, False, True, src.xb.device)
        s1 = src.xb.stride(0)
        tap_gemm(src.xb, (M, 1, 1), (s1, s1 * M, s1 * M), (M, 1, 1), [(0, 0, 0)], ln.w, N, bias=ln.b, geglu=geglu,
                 out_bf16=ob, ld_out=ld, block_n=block_n, ln_acc=src.acc, ln_colsum=ln.s, ln_channels=src.C, ln_eps=ln.eps)
        return ob
    M, K = a.shape
    if ln_stats and not fp32_mode():
        assert a.stride(1) == 1 and w.shape[1] == K and not geglu and out is None and not stats_rows
        N = N or w.shape[0]
        of, ob, ld = _outputs(M, N, None, True, False, a.device)
        xb = torch.empty((M, N), dtype=BF16, device=a.device)
        acc = _alloc_i64(2 * M, a.device)
        s1 = a.stride(0)
        tap_gemm(a, (M, 1, 1), (s1, s1 * M, s1 * M), (M, 1, 1), [(0, 0, 0)], w, N, bias=bias, residual=residual,
                 out_f32=of, ld_out=ld, group_bias=group_bias, rows_per_group=rows_per_group, out_scale=out_scale,
                 block_n=block_n, ln_acc=acc, ln_x=xb)
        return of, LnSrc(xb, acc, N)
    if a.dtype == F32:                      # fp32 parity mode
        assert a.stride(1) == 1 and w.shape[1] == 3 * K
        N = N or w.shape[0]
        if geglu:                           # plain GEMM (weights not interleaved) + exact-erf gate on the CUDA cores
            assert residual is None and out is None
            both = linear(a, w, bias=bias, N=N, out_scale=out_scale)
            o = torch.empty((M, N // 2), dtype=F32, device=a.device)
            check(_lib.lib().dcb_geglu_f32(ptr(both), _c_ll(M), _c_i(N // 2), _c_ll(both.stride(0)), ptr(o), _stream()),
                  "dcb_geglu_f32")
            return o
        of, ob, ld = _outputs(M, N, out, True, False, a.device)
        s1 = a.stride(0)
        tap_gemm(a, (M, 1, 1), (s1, s1 * M, s1 * M), (M, 1, 1), [(0, 0, 0)], w, N, bias=bias, residual=residual,
                 out_f32=of, ld_out=ld, group_bias=group_bias, rows_per_group=rows_per_group, out_scale=out_scale)
        return (of, None) if stats_rows else of
    assert a.stride(1) == 1 and w.shape[1] == K
    N = N or w.shape[0]
    n_out = N // 2 if geglu else N
    of, ob, ld = _outputs(M, n_out, out, f32, bf16, a.device)
    s1 = a.stride(0)
    res = of if (of is not None and ob is None) else (ob if of is None else (of, ob))
    if stats_rows:
        # same GEMM over a (rows per frame, 1, frames) domain so that the statistics partials are frame aligned
        assert M % stats_rows == 0 and not geglu
        dims = (stats_rows, 1, M // stats_rows)
        cs = _make_colsums(dims, n_out, a.device) if (stats_enabled() and colsum_supported(dims, n_out)) else None
        tap_gemm(a, dims, (s1, s1 * stats_rows, s1 * stats_rows), dims, [(0, 0, 0)], w, N, bias=bias, residual=residual,
                 out_f32=of, out_bf16=ob, ld_out=ld, group_bias=group_bias, rows_per_group=rows_per_group,
                 out_scale=out_scale, block_n=block_n, colsum=cs)
        return res, cs
    tap_gemm(a, (M, 1, 1), (s1, s1 * M, s1 * M), (M, 1, 1), [(0, 0, 0)], w, N, bias=bias, residual=residual,
             geglu=geglu, out_f32=of, out_bf16=ob, ld_out=ld, group_bias=group_bias,
             rows_per_group=rows_per_group, out_scale=out_scale, block_n=block_n)
    return res


TAPS_3X3 = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
TAPS_T3 = [(0, dt, 0) for dt in (-1, 0, 1)]


def conv3x3(a, nf, H, W, w, *, bias=None, residual=None, group_bias=None, rows_per_group=0,
            f32=True, bf16=False, out=None, N=None, stats=False):
    """a: bf16 [nf*H*W, C] channels-last frames; w: bf16 [N, 9*C] ordered (ky, kx, c). Stride 1, pad 1."""
    C = a.shape[1]
    assert a.stride(1) == 1 and a.stride(0) == C and w.shape[1] == (27 if a.dtype == F32 else 9) * C
    N = N or w.shape[0]
    of, ob, ld = _outputs(nf * H * W, N, out, f32, bf16, a.device)
    cs = _make_colsums((W, H, nf), N, a.device) if (stats and stats_enabled() and colsum_supported((W, H, nf), N)) else None
    tap_gemm(a, (W, H, nf), (C, C * W, C * W * H), (W, H, nf), TAPS_3X3, w, N, bias=bias, residual=residual,
             group_bias=group_bias, rows_per_group=rows_per_group, out_f32=of, out_bf16=ob, ld_out=ld, colsum=cs)
    res = of if of is not None else ob
    return (res, cs) if stats else res


# nearest-x2 upsample followed by a 3x3 conv (openaimodel3d.py:98-106, ae_modules.py:123-127) without the 4x tensor: the
# output pixel (2j + py, 2i + px) only sees a 2x2 neighbourhood of the low-resolution input, so each of the four output
# phases is a 2x2-tap conv of the ORIGINAL tensor with pre-summed weights (taps per axis: phase 0 reads offsets -1, 0 with
# k0 and k1 + k2; phase 1 reads 0, +1 with k0 + k1 and k2): 16 instead of 36 MACs per output element and input channel.
_UP2_OFFS = ((-1, 0), (0, 1))


def conv3x3_up2(a, nf, H, W, w4, *, bias=None, f32=True, bf16=False, stats=False):
    """a: bf16 [nf*H*W, C] low-resolution frames; w4: 4 packed phase weights (pack_conv3x3_up2), each bf16 [N, 4*C] ordered
    (a_y, b_x, c).  -> [nf*2H*2W, N] (rows of the upsampled frames), optionally with fused GroupNorm statistics."""
    C = a.shape[1]
    assert a.stride(1) == 1 and a.stride(0) == C and len(w4) == 4
    N = w4[0].shape[0]
    of, ob, ld = _outputs(nf * 4 * H * W, N, None, f32, bf16, a.device)
    full = of if of is not None else ob
    dims = (W, H, nf)
    cs = _make_colsums(dims, N, a.device) if (stats and stats_enabled() and colsum_supported(dims, N)) else None
    if cs is not None:
        cs.unit_rows *= 4                        # a unit (frame, or image row) of the OUTPUT holds the rows of all four phases
    ostr = (2 * ld, 2 * 2 * W * ld, 4 * H * W * ld)
    for py in range(2):
        for px in range(2):
            taps = [(_UP2_OFFS[px][b], _UP2_OFFS[py][a_], 0) for a_ in range(2) for b in range(2)]
            o = full[(py * 2 * W + px):]
            tap_gemm(a, dims, (C, C * W, C * W * H), dims, taps, w4[py * 2 + px], N, bias=bias,
                     out_f32=None if of is None else o, out_bf16=None if ob is None else o, ld_out=ld, colsum=cs,
                     out_strides=ostr)
    return (full, cs) if stats else full


def tconv3(a, B, T, HW, w, *, bias=None, residual=None, f32=True, bf16=False, out=None, stats=False):
    """Temporal (3,1,1) conv.  a: bf16 [B*T*HW, C]; w: bf16 [N, 3*C] ordered (kt, c). Pad 1 frame."""
    C = a.shape[1]
    assert a.stride(1) == 1 and a.stride(0) == C and w.shape[1] == (9 if a.dtype == F32 else 3) * C
    N = w.shape[0]
    of, ob, ld = _outputs(B * T * HW, N, out, f32, bf16, a.device)
    cs = _make_colsums((HW, T, B), N, a.device) if (stats and stats_enabled() and colsum_supported((HW, T, B), N)) else None
    tap_gemm(a, (HW, T, B), (C, C * HW, C * HW * T), (HW, T, B), TAPS_T3, w, N, bias=bias, residual=residual,
             out_f32=of, out_bf16=ob, ld_out=ld, colsum=cs)
    res = of if of is not None else ob
    return (res, cs) if stats else res


# ------------------------------------------------------------------------------------------ norms
_GN_SCRATCH = {}


# GroupNorm inputs up to this many bytes take the single cooperative launch (statistics + apply, second read from
# L2); larger ones keep the two streaming kernels.  DCB_GN_FUSED_MB overrides (0 = never).
_GN_FUSED_MAX_BYTES = int(float(os.environ.get("DCB_GN_FUSED_MB", "4096")) * (1 << 20))


_GN_RETIRED = []         # outgrown scratch buffers: kept alive because captured CUDA graphs bake in their addresses


def _gn_scratch(dev, G):
    """(workspace, counters) reused by every GroupNorm launch on this device.  The engine is single-stream (launches
    and graph replays are stream ordered), so one pair per device is enough.  A buffer that is outgrown is retired,
    not freed -- a CUDA graph captured earlier keeps replaying against the old addresses."""
    key = (dev.index, )
    ent = _GN_SCRATCH.get(key)
    if ent is None or ent[2] < G:
        if ent is not None:
            _GN_RETIRED.append(ent)
        cap = max(2 * G, 256)
        nbytes = _lib.lib().dcb_group_norm_workspace_bytes(cap)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        cnt = torch.zeros(2 * cap, dtype=torch.int32, device=dev)     # fused kernel: arrive + leave counters
        ent = (ws, cnt, cap)
        _GN_SCRATCH[key] = ent
    return ent[0], ent[1]


def _c_i(v):
    return ctypes.c_int(int(v))


def _c_ll(v):
    return ctypes.c_longlong(int(v))


def _c_f(v):
    return ctypes.c_float(float(v))


def _stats_from_colsums(cs, G, S, C, eps, dev):
    """(mean, rstd) [G, 32, 2] from a fused-statistics accumulator (tests / inspection)."""
    assert cs.fits(S, C)
    upi = S // cs.unit_rows
    assert cs.acc.shape[0] == G * upi, (cs.acc.shape, G, upi)
    stats = torch.empty((G, 32, 2), dtype=F32, device=dev)
    check(_lib.lib().dcb_group_norm_stats_from_acc(ptr(cs.acc), _c_i(upi), _c_i(G), ctypes.c_double(float(S) * (C // 32)),
                                                   _c_f(eps), ptr(stats), _stream()), "dcb_group_norm_stats_from_acc")
    return stats


def _group_norm_from_colsums(x, G, S, gamma, beta, eps, silu, cs, shard=None):
    """Producer-fused path: ONE streaming apply pass; no statistics pass over x, no finalize launch, no grid barrier.
    shard: the instances are spread over the ranks of a FrameShard -> the per-instance int64 sums are all-reduced first
    (integer sums: every rank gets bit-identical statistics)."""
    C, dev = x.shape[1], x.device
    upi = S // cs.unit_rows
    assert cs.acc.shape[0] == G * upi, (cs.acc.shape, G, upi)
    acc, stat_rows = cs.acc, 0
    if shard is not None:
        from .parallel import all_reduce_sums
        acc = cs.acc.view(G, upi, 64).sum(1) if upi > 1 else cs.acc.reshape(G, 64).clone()
        all_reduce_sums(acc, shard)
        upi, stat_rows = 1, S * shard.world
    y = torch.empty((G * S, C), dtype=BF16, device=dev)
    check(_lib.lib().dcb_group_norm_stream_apply(ptr(x), _c_i(1 if x.dtype == BF16 else 0), _c_i(C), _c_i(G), _c_i(S),
                                                 _c_ll(stat_rows), ptr(acc), _c_i(upi), _c_f(eps), ptr(gamma), ptr(beta),
                                                 _c_i(1 if silu else 0), ptr(y), _stream()), "dcb_group_norm_stream_apply")
    return y


_SHARDED_GN = [None]      # FrameShard while the instances of group_norm calls are spread over ranks


class sharded_group_norm:
    """Context: every group_norm inside normalises instances whose rows are split over the ranks of `shard`
    (the 5-D GroupNorms of the temporal blocks in frame-sharded mode): local raw sums -> all-reduce -> apply."""

    def __init__(self, shard):
        self.shard = shard

    def __enter__(self):
        self.prev, _SHARDED_GN[0] = _SHARDED_GN[0], self.shard

    def __exit__(self, *exc):
        _SHARDED_GN[0] = self.prev


def _group_norm_sharded(x1, G, S, gamma, beta, eps, silu, shard):
    from .parallel import all_reduce_sums
    L, dev, C = _lib.lib(), x1.device, x1.shape[1]
    assert x1.is_contiguous() and x1.shape[0] == G * S
    ws, cnt = _gn_scratch(dev, G)
    sums = torch.empty((G, 32, 2), dtype=torch.float64, device=dev)
    if x1.dtype == BF16:
        check(L.dcb_group_norm_sums_bf16(ptr(x1), _c_i(C), _c_i(G), _c_i(S), ptr(sums), ptr(ws), ptr(cnt), _stream()),
              "dcb_group_norm_sums_bf16")
    else:
        check(L.dcb_group_norm_sums(ptr(x1), None, _c_i(C), _c_i(0), _c_i(G), _c_i(S), ptr(sums), ptr(ws), ptr(cnt),
                                    _stream()), "dcb_group_norm_sums")
    all_reduce_sums(sums, shard)
    stats = torch.empty((G, 32, 2), dtype=F32, device=dev)
    count = float(S) * shard.world * (C // 32)
    check(L.dcb_group_norm_stats_from_sums(ptr(sums), _c_i(G), ctypes.c_double(count), _c_f(eps), ptr(stats), _stream()),
          "dcb_group_norm_stats_from_sums")
    if fp32_mode():
        assert x1.dtype == F32
        y = torch.empty((G * S, C), dtype=F32, device=dev)
        check(L.dcb_group_norm_apply_f32out(ptr(x1), _c_i(C), _c_i(G), _c_i(S), ptr(stats), ptr(gamma), ptr(beta),
                                            _c_i(1 if silu else 0), ptr(y), _stream()), "dcb_group_norm_apply_f32out")
        return y
    y = torch.empty((G * S, C), dtype=BF16, device=dev)
    if x1.dtype == BF16:
        check(L.dcb_group_norm_apply_bf16(ptr(x1), _c_i(C), _c_i(G), _c_i(S), ptr(stats), ptr(gamma), ptr(beta),
                                          _c_i(1 if silu else 0), ptr(y), _stream()), "dcb_group_norm_apply_bf16")
    else:
        check(L.dcb_group_norm_apply(ptr(x1), None, _c_i(C), _c_i(0), _c_i(G), _c_i(S), ptr(stats), ptr(gamma),
                                     ptr(beta), _c_i(1 if silu else 0), ptr(y), None, _stream()), "dcb_group_norm_apply")
    return y


def _concat_groups_align(C1, C2):
    """Can the 32 groups of cat([x1, x2]) be assembled from whole producer groups (C_i / 32 channels) of one source each?"""
    if C1 % 32 or C2 % 32:
        return False
    cg = (C1 + C2) // 32
    return C1 % cg == 0 and cg % (C1 // 32) == 0 and cg % (C2 // 32) == 0


def group_norm(x1, G, S, gamma, beta, eps, *, x2=None, silu=True, want_raw=False, want_norm=True, colsums=None,
               colsums2=None):
    """x1: fp32 [G*S, C1] (+ x2 [G*S, C2] concatenated on channels) -> bf16 normalised operand
    (and/or bf16 raw copy of the concat).  colsums / colsums2: producer-fused statistics of x1 / x2."""
    require_cuda(x1, x2, gamma, beta)
    if (x2 is not None and colsums is not None and colsums2 is not None and _SHARDED_GN[0] is None and not fp32_mode()
            and want_norm and x1.dtype == F32 and x2.dtype == F32 and x1.is_contiguous() and x2.is_contiguous()
            and x1.shape[1] % 8 == 0 and x2.shape[1] % 8 == 0 and colsums.fits(S, x1.shape[1]) and colsums2.fits(S, x2.shape[1])
            and _concat_groups_align(x1.shape[1], x2.shape[1]) and x1.shape[1] + x2.shape[1] <= 4096):
        C1, C2, dev = x1.shape[1], x2.shape[1], x1.device
        y = torch.empty((G * S, C1 + C2), dtype=BF16, device=dev)
        raw = torch.empty((G * S, C1 + C2), dtype=BF16, device=dev) if want_raw else None
        check(_lib.lib().dcb_group_norm_stream_apply2(ptr(x1), ptr(x2), _c_i(C1), _c_i(C2), _c_i(G), _c_i(S), ptr(colsums.acc),
                                                      _c_i(S // colsums.unit_rows), ptr(colsums2.acc),
                                                      _c_i(S // colsums2.unit_rows), _c_f(eps), ptr(gamma), ptr(beta),
                                                      _c_i(1 if silu else 0), ptr(y), ptr(raw), _stream()),
              "dcb_group_norm_stream_apply2")
        return (y, raw) if want_raw else y
    if _SHARDED_GN[0] is not None:
        assert x2 is None and not want_raw and want_norm
        if colsums is not None and not fp32_mode() and x1.shape[1] % 8 == 0 and x1.is_contiguous() and colsums.fits(S, x1.shape[1]):
            return _group_norm_from_colsums(x1, G, S, gamma, beta, eps, silu, colsums, shard=_SHARDED_GN[0])
        return _group_norm_sharded(x1, G, S, gamma, beta, eps, silu, _SHARDED_GN[0])
    if colsums is not None and (fp32_mode() or x2 is not None or want_raw or not want_norm or x1.shape[1] % 8
                                or not x1.is_contiguous() or not colsums.fits(S, x1.shape[1])):
        colsums = None                      # not usable here: take the self-contained path
    if colsums is not None:
        assert x1.shape[0] == G * S
        return _group_norm_from_colsums(x1, G, S, gamma, beta, eps, silu, colsums)
    if fp32_mode():
        assert x1.dtype == F32 and x1.is_contiguous() and x1.shape[0] == G * S and want_norm
        C1, C2, dev, L = x1.shape[1], (0 if x2 is None else x2.shape[1]), x1.device, _lib.lib()
        ws, cnt = _gn_scratch(dev, G)
        y = torch.empty((G * S, C1 + C2), dtype=F32, device=dev)
        raw = torch.empty((G * S, C1 + C2), dtype=F32, device=dev) if want_raw else None
        check(L.dcb_group_norm_fused_f32out(ptr(x1), ptr(x2), _c_i(C1), _c_i(C2), _c_i(G), _c_i(S), _c_f(eps), ptr(gamma),
                                            ptr(beta), _c_i(1 if silu else 0), ptr(y), ptr(raw), ptr(ws), ptr(cnt),
                                            _stream()), "dcb_group_norm_fused_f32out")
        return (y, raw) if want_raw else y
    if x1.dtype == BF16:
        # block-internal activation kept in bf16 by its producer: single source, normalised output only
        assert x2 is None and not want_raw and x1.is_contiguous() and x1.shape[0] == G * S
        C, dev, L = x1.shape[1], x1.device, _lib.lib()
        if colsums is None and x1.numel() * 2 <= _GN_FUSED_MAX_BYTES:
            ws, cnt = _gn_scratch(dev, G)
            y = torch.empty((G * S, C), dtype=BF16, device=dev)
            check(L.dcb_group_norm_fused_bf16(ptr(x1), _c_i(C), _c_i(G), _c_i(S), _c_f(eps), ptr(gamma), ptr(beta),
                                              _c_i(1 if silu else 0), ptr(y), ptr(ws), ptr(cnt), _stream()),
                  "dcb_group_norm_fused_bf16")
            return y
        if colsums is not None:
            stats = _stats_from_colsums(colsums, G, S, C, eps, dev)
        else:
            ws, cnt = _gn_scratch(dev, G)
            stats = torch.empty((G, 32, 2), dtype=F32, device=dev)
            check(L.dcb_group_norm_stats_bf16(ptr(x1), _c_i(C), _c_i(G), _c_i(S), _c_f(eps), ptr(stats), ptr(ws), ptr(cnt),
                                              _stream()), "dcb_group_norm_stats_bf16")
        y = torch.empty((G * S, C), dtype=BF16, device=dev)
        check(L.dcb_group_norm_apply_bf16(ptr(x1), _c_i(C), _c_i(G), _c_i(S), ptr(stats), ptr(gamma), ptr(beta),
                                          _c_i(1 if silu else 0), ptr(y), _stream()), "dcb_group_norm_apply_bf16")
        return y
    assert x1.dtype == F32 and x1.is_contiguous() and x1.shape[0] == G * S
    C1 = x1.shape[1]
    C2 = 0 if x2 is None else x2.shape[1]
    if x2 is not None:
        assert x2.dtype == F32 and x2.is_contiguous() and x2.shape[0] == G * S
    dev = x1.device
    L = _lib.lib()
    y = raw = stats = None
    if want_norm and colsums is None and (x1.numel() + (0 if x2 is None else x2.numel())) * 4 <= _GN_FUSED_MAX_BYTES \
            and C1 % 8 == 0 and C2 % 8 == 0:
        ws, cnt = _gn_scratch(dev, G)
        y = torch.empty((G * S, C1 + C2), dtype=BF16, device=dev)
        if want_raw:
            raw = torch.empty((G * S, C1 + C2), dtype=BF16, device=dev)
        check(L.dcb_group_norm_fused(ptr(x1), ptr(x2), _c_i(C1), _c_i(C2), _c_i(G), _c_i(S), _c_f(eps), ptr(gamma),
                                     ptr(beta), _c_i(1 if silu else 0), ptr(y), ptr(raw), ptr(ws), ptr(cnt), _stream()),
              "dcb_group_norm_fused")
        return (y, raw) if want_raw else y
    if want_norm and colsums is not None and x2 is None:
        stats = _stats_from_colsums(colsums, G, S, C1, eps, dev)
        y = torch.empty((G * S, C1), dtype=BF16, device=dev)
    elif want_norm:
        ws, cnt = _gn_scratch(dev, G)
        stats = torch.empty((G, 32, 2), dtype=F32, device=dev)
        check(L.dcb_group_norm_stats(ptr(x1), ptr(x2), _c_i(C1), _c_i(C2), _c_i(G), _c_i(S), _c_f(eps), ptr(stats),
                                     ptr(ws), ptr(cnt), _stream()), "dcb_group_norm_stats")
        y = torch.empty((G * S, C1 + C2), dtype=BF16, device=dev)
    if want_raw:
        raw = torch.empty((G * S, C1 + C2), dtype=BF16, device=dev)
    check(L.dcb_group_norm_apply(ptr(x1), ptr(x2), _c_i(C1), _c_i(C2), _c_i(G), _c_i(S), ptr(stats), ptr(gamma),
                                 ptr(beta), _c_i(1 if silu else 0), ptr(y), ptr(raw), _stream()),
          "dcb_group_norm_apply")
    if want_norm and want_raw:
        return y, raw
    return y if want_norm else raw


def layer_norm(x, gamma, beta, eps=1e-5, out_f32=False):
    """fp32 [rows, C] -> bf16 operand (or fp32 when the result leaves the engine, out_f32=True)."""
    require_cuda(x, gamma, beta)
    assert x.dtype == F32 and x.is_contiguous()
    rows, C = x.shape
    if fp32_mode() or out_f32:
        y = torch.empty((rows, C), dtype=F32, device=x.device)
        check(_lib.lib().dcb_layer_norm_f32out(ptr(x), _c_i(rows), _c_i(C), ptr(gamma), ptr(beta), _c_f(eps), ptr(y),
                                               _stream()), "dcb_layer_norm_f32out")
        return y
    y = torch.empty((rows, C), dtype=BF16, device=x.device)
    check(_lib.lib().dcb_layer_norm(ptr(x), _c_i(rows), _c_i(C), ptr(gamma), ptr(beta), _c_f(eps), ptr(y), _stream()),
          "dcb_layer_norm")
    return y


# ------------------------------------------------------------------------------------------ attention
def flash_attention(q, segments, batch, heads, nq, scale, out=None):
    """q: bf16 2-D view [batch*nq, >= heads*64]; segments: list of (k, v, nk, kv_batch_div, seg_scale) with
    k/v 2-D views [kv_batches*nk, >= heads*64].  Returns bf16 [batch*nq, heads*64]."""
    require_cuda(q)
    f32 = q.dtype == F32                    # fp32 parity mode: CUDA-core attention on fp32 q / k / v
    a = _lib.AttentionArgs()
    a.q, a.ldq = q.data_ptr(), q.stride(0)
    a.nseg = len(segments)
    for i, (k, v, nk, div, sc) in enumerate(segments):
        require_cuda(k, v)
        assert k.dtype == q.dtype and v.dtype == q.dtype and k.stride(1) == 1 and v.stride(1) == 1
        a.k[i], a.v[i] = k.data_ptr(), v.data_ptr()
        a.ldk[i], a.ldv[i] = k.stride(0), v.stride(0)
        a.nk[i], a.kv_batch_div[i], a.seg_scale[i] = nk, div, sc
    a.batch, a.heads, a.nq, a.head_dim = batch, heads, nq, 64
    if out is None:
        out = torch.empty((batch * nq, heads * 64), dtype=q.dtype, device=q.device)
    a.o, a.ldo = out.data_ptr(), out.stride(0)
    a.softmax_scale = scale
    if f32:
        assert fp32_mode() and out.dtype == F32
        check(_lib.lib().dcb_attention_f32(ctypes.byref(a), _c_i(1), _stream()), "dcb_attention_f32")
        return out
    check(_lib.lib().dcb_flash_attention(ctypes.byref(a), _stream()), "dcb_flash_attention")
    return out


def temporal_attention(q, k, v, B, L, HW, heads, scale, out=None):
    """q/k/v: bf16 2-D views [B*L*HW, >= heads*64] (rows ordered b, frame, pixel)."""
    require_cuda(q, k, v)
    if q.dtype == F32:                      # fp32 parity mode: B*HW interleaved sequences of L tokens
        assert fp32_mode() and out is None
        out = torch.empty((B * L * HW, heads * 64), dtype=F32, device=q.device)
        a = _lib.AttentionArgs()
        a.q, a.ldq = q.data_ptr(), q.stride(0)
        a.nseg = 1
        a.k[0], a.v[0], a.ldk[0], a.ldv[0] = k.data_ptr(), v.data_ptr(), k.stride(0), v.stride(0)
        a.nk[0], a.kv_batch_div[0], a.seg_scale[0] = L, 1, 1.0
        a.batch, a.heads, a.nq, a.head_dim = B * HW, heads, L, 64
        a.o, a.ldo = out.data_ptr(), out.stride(0)
        a.softmax_scale = scale
        check(_lib.lib().dcb_attention_f32(ctypes.byref(a), _c_i(HW), _stream()), "dcb_attention_f32")
        return out
    if out is None:
        out = torch.empty((B * L * HW, heads * 64), dtype=BF16, device=q.device)
    check(_lib.lib().dcb_temporal_attention(ptr(q), ptr(k), ptr(v), ptr(out), _c_ll(q.stride(0)), _c_ll(k.stride(0)),
                                            _c_ll(v.stride(0)), _c_ll(out.stride(0)), _c_i(B), _c_i(L), _c_i(HW),
                                            _c_i(heads), _c_i(64), _c_f(scale), _stream()), "dcb_temporal_attention")
    return out


def vae_attention(q, k, v, frames, tokens, C, scale, out=None):
    """q/k/v: bf16 2-D views [frames*tokens, >= C]; single head of width C (<= 512) per frame -> bf16 [frames*tokens, C]."""
    require_cuda(q, k, v)
    assert q.dtype == BF16 and k.dtype == BF16 and v.dtype == BF16
    if out is None:
        out = torch.empty((frames * tokens, C), dtype=BF16, device=q.device)
    check(_lib.lib().dcb_vae_attention(ptr(q), ptr(k), ptr(v), ptr(out), _c_ll(q.stride(0)), _c_ll(k.stride(0)),
                                       _c_ll(v.stride(0)), _c_ll(out.stride(0)), _c_i(frames), _c_i(tokens), _c_i(C),
                                       _c_f(scale), _stream()), "dcb_vae_attention")
    return out


def softmax_rows(x, scale):
    require_cuda(x)
    rows, n = x.shape
    y = torch.empty((rows, n), dtype=BF16, device=x.device)
    check(_lib.lib().dcb_softmax_rows(ptr(x), _c_ll(rows), _c_i(n), _c_ll(x.stride(0)), _c_f(scale), ptr(y),
                                      _c_ll(y.stride(0)), _stream()), "dcb_softmax_rows")
    return y


# ------------------------------------------------------------------------------------------ layout / misc
def nchw_to_rows(x, Cpad):
    """fp32 [B, C, T, H, W] (or [B, C, H, W]) -> bf16 [B*T*H*W, Cpad] channels-last, zero padded."""
    require_cuda(x)
    x = x.contiguous().float()
    if x.dim() == 4:
        B, C, H, W = x.shape
        T = 1
    else:
        B, C, T, H, W = x.shape
    if fp32_mode():
        out = torch.empty((B * T * H * W, Cpad), dtype=F32, device=x.device)
        check(_lib.lib().dcb_nchw_to_rows_f32out(ptr(x), _c_i(B), _c_i(C), _c_i(T), _c_i(H * W), _c_i(Cpad), ptr(out),
                                                 _stream()), "dcb_nchw_to_rows_f32out")
        return out
    out = torch.empty((B * T * H * W, Cpad), dtype=BF16, device=x.device)
    check(_lib.lib().dcb_nchw_to_rows(ptr(x), _c_i(B), _c_i(C), _c_i(T), _c_i(H * W), _c_i(Cpad), ptr(out), _stream()),
          "dcb_nchw_to_rows")
    return out


def rows_to_nchw(rows, B, C, T, H, W, five_d=True):
    require_cuda(rows)
    assert rows.dtype == F32 and rows.stride(1) == 1
    out = torch.empty((B, C, T, H, W) if five_d else (B, C, H, W), dtype=F32, device=rows.device)
    check(_lib.lib().dcb_rows_to_nchw(ptr(rows), _c_i(rows.stride(0)), _c_i(B), _c_i(C), _c_i(T), _c_i(H * W), ptr(out),
                                      _stream()), "dcb_rows_to_nchw")
    return out


def upsample2x(x, nf, H, W):
    require_cuda(x)
    C = x.shape[1]
    if fp32_mode():
        out = torch.empty((nf * 4 * H * W, C), dtype=F32, device=x.device)
        check(_lib.lib().dcb_upsample2x_f32out(ptr(x), _c_i(nf), _c_i(H), _c_i(W), _c_i(C), ptr(out), _stream()),
              "dcb_upsample2x_f32out")
        return out
    out = torch.empty((nf * 4 * H * W, C), dtype=BF16, device=x.device)
    check(_lib.lib().dcb_upsample2x(ptr(x), _c_i(nf), _c_i(H), _c_i(W), _c_i(C), ptr(out), _stream()), "dcb_upsample2x")
    return out


def space_to_depth(x, nf, H, W):
    require_cuda(x)
    C = x.shape[1]
    Hp, Wp = (H + 1) // 2, (W + 1) // 2
    if fp32_mode():
        out = torch.empty((4 * nf * Hp * Wp, C), dtype=F32, device=x.device)
        check(_lib.lib().dcb_space_to_depth_f32out(ptr(x), _c_i(nf), _c_i(H), _c_i(W), _c_i(C), ptr(out), _stream()),
              "dcb_space_to_depth_f32out")
        return out
    out = torch.empty((4 * nf * Hp * Wp, C), dtype=BF16, device=x.device)
    check(_lib.lib().dcb_space_to_depth(ptr(x), _c_i(nf), _c_i(H), _c_i(W), _c_i(C), ptr(out), _stream()),
          "dcb_space_to_depth")
    return out


# stride-2 pad-1 3x3 conv on the 2x2 phase planes: tap (ky,kx) reads plane (py,px) at offset (-1 or 0)
_S2 = {0: (1, -1), 1: (0, 0), 2: (1, 0)}   # k -> (phase, offset)


# same conv with the VAE encoder's asymmetric padding (0,1,0,1) (ae_modules.py:103-106): tap k reads pixel 2i + k
_S2_ASYM = {0: (0, 0), 1: (1, 0), 2: (0, 1)}


def conv3x3_s2(planes, nf, H, W, w, *, bias=None, f32=True, bf16=False, asym=False, stats=False):
    """planes: output of space_to_depth ([4*nf*Hp*Wp, C]); w: bf16 [N, 9*C] (ky,kx,c). -> [nf*Ho*Wo, N].
    asym=False: stride 2, pad 1 (UNet Downsample); asym=True: pad right / bottom only, then stride 2 (VAE encoder)."""
    C = planes.shape[1]
    Hp, Wp = (H + 1) // 2, (W + 1) // 2
    if asym:
        Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
    else:
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    tab = _S2_ASYM if asym else _S2
    taps = []
    for ky in range(3):
        py, oy = tab[ky]
        for kx in range(3):
            px, ox = tab[kx]
            taps.append((ox, oy, (py * 2 + px) * nf))
    N = w.shape[0]
    of, ob, ld = _outputs(nf * Ho * Wo, N, None, f32, bf16, planes.device)
    cs = _make_colsums((Wo, Ho, nf), N, planes.device) if (stats and stats_enabled() and colsum_supported((Wo, Ho, nf), N)) else None
    tap_gemm(planes, (Wo, Ho, nf), (C, C * Wp, C * Wp * Hp), (Wp, Hp, 4 * nf), taps, w, N, bias=bias,
             out_f32=of, out_bf16=ob, ld_out=ld, colsum=cs)
    res = of if (of is not None and ob is None) else (ob if of is None else (of, ob))
    return (res, cs) if stats else res


def cast_bf16(x):
    require_cuda(x)
    if fp32_mode():
        return x                            # operands stay fp32; the GEMM splits them
    assert x.dtype == F32 and x.is_contiguous() and x.numel() % 8 == 0
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(_lib.lib().dcb_cast_bf16(ptr(x), _c_ll(x.numel()), ptr(out), _stream()), "dcb_cast_bf16")
    return out


def block_transpose(src, Z, A, B, residual=None):
    """src viewed as [Z, A, B, inner] -> [Z, B, A, inner] (same dtype), flat result of src's shape.  With residual (fp32, in
    the RESULT's order) and bf16 src: fp32 residual + src."""
    require_cuda(src, residual)
    assert src.is_contiguous() and src.numel() % (Z * A * B) == 0
    inner_bytes = src.numel() // (Z * A * B) * src.element_size()
    if residual is not None:
        assert src.dtype == BF16 and residual.dtype == F32 and residual.is_contiguous() and residual.numel() == src.numel()
        out = torch.empty_like(residual)
        check(_lib.lib().dcb_block_transpose(ptr(src), None, _c_i(Z), _c_i(A), _c_i(B), _c_ll(inner_bytes), ptr(residual),
                                             ptr(out), _stream()), "dcb_block_transpose")
        return out
    out = torch.empty_like(src)
    check(_lib.lib().dcb_block_transpose(ptr(src), ptr(out), _c_i(Z), _c_i(A), _c_i(B), _c_ll(inner_bytes), None, None,
                                         _stream()), "dcb_block_transpose")
    return out


def timestep_embedding(t, dim):
    require_cuda(t)
    t = t.to(torch.int64).contiguous()
    out = torch.empty((t.shape[0], dim), dtype=F32, device=t.device)
    check(_lib.lib().dcb_timestep_embedding(ptr(t), _c_i(t.shape[0]), _c_i(dim), ptr(out), _stream()),
          "dcb_timestep_embedding")
    return out


def small_linear(a, w, bias=None, *, silu_in=False, silu_out=False, out=None, accumulate=False):
    """fp32 [M, K] x fp32 [N, K]^T for tiny M (embedding MLPs)."""
    require_cuda(a, w, bias, out)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=F32, device=a.device)
    check(_lib.lib().dcb_small_linear(ptr(a), _c_i(a.stride(0)), ptr(w), ptr(bias), _c_i(M), _c_i(N), _c_i(K),
                                      _c_i(silu_in), _c_i(silu_out), _c_i(accumulate), ptr(out), _c_i(out.stride(0)),
                                      _stream()), "dcb_small_linear")
    return out


def ddim_update(x, e_cond, e_uncond, noise, coeffs, e_uncond_img=None, cfg_img=None):
    """coeffs: _lib.DdimCoeffs.  e_uncond_img / cfg_img: third branch of the multi-condition sampler.
    Returns (x_prev, pred_x0)."""
    require_cuda(x, e_cond, e_uncond, noise, e_uncond_img)
    B = x.shape[0]
    per = x.numel() // B
    x_prev = torch.empty_like(x)
    pred = torch.empty_like(x)
    stds = torch.empty((B, 2), dtype=F32, device=x.device)
    if e_uncond_img is not None:
        check(_lib.lib().dcb_ddim_update_multicond(ptr(x), ptr(e_cond), ptr(e_uncond), ptr(e_uncond_img), _c_f(cfg_img),
                                                   ptr(noise), _c_i(B), _c_ll(per), ctypes.byref(coeffs), ptr(stds),
                                                   ptr(x_prev), ptr(pred), _stream()), "dcb_ddim_update_multicond")
        return x_prev, pred
    check(_lib.lib().dcb_ddim_update(ptr(x), ptr(e_cond), ptr(e_uncond), ptr(noise), _c_i(B), _c_ll(per),
                                     ctypes.byref(coeffs), ptr(stds), ptr(x_prev), ptr(pred), _stream()),
          "dcb_ddim_update")
    return x_prev, pred


def sds_grad(latents, noisy, e_cond, e_uncond, sqrt_a, sqrt_1ma, w_t, cfg_scale, guidance_rescale, weight_type):
    """SDS gradient + loss value (dcb_sds_grad).  fp32 [B, ...] tensors; sqrt_a / sqrt_1ma / w_t: fp32 [B].
    Returns (grad like latents, loss as a 0-dim fp64 tensor = 0.5 * mean((latents - (latents - grad))^2) / B)."""
    require_cuda(latents, noisy, e_cond, e_uncond)
    B = latents.shape[0]
    per = latents.numel() // B
    dev = latents.device
    f = lambda v: v.to(device=dev, dtype=F32).reshape(B).contiguous()
    sa, s1, wt = f(sqrt_a), f(sqrt_1ma), f(w_t)
    grad = torch.empty_like(latents)
    stds = torch.empty(2 * B, dtype=F32, device=dev)
    mabs = torch.empty(B, dtype=F32, device=dev)
    diff = torch.empty_like(latents)
    part = torch.empty(B * per // 256, dtype=torch.float64, device=dev)
    check(_lib.lib().dcb_sds_grad(ptr(latents), ptr(noisy), ptr(e_cond), ptr(e_uncond), _c_i(B), _c_ll(per), ptr(sa), ptr(s1),
                                  ptr(wt), _c_f(cfg_scale), _c_f(guidance_rescale), _c_i(weight_type), ptr(stds), ptr(mabs),
                                  ptr(diff), ptr(grad), ptr(part), _stream()), "dcb_sds_grad")
    loss = 0.5 * part.sum() / (B * per) / B
    return grad, loss


def q_sample(x0, noise, a, c):
    """out[b] = a[b] * x0[b] + c[b] * noise[b]; a, c: fp32 [B] (forward diffusion of clean latents)."""
    require_cuda(x0, noise)                 # a, c: tiny per-sample coefficient vectors, host or device
    x0 = x0.float().contiguous()
    noise = noise.float().contiguous()
    B = x0.shape[0]
    a = a.to(device=x0.device, dtype=F32).reshape(B).contiguous()
    c = c.to(device=x0.device, dtype=F32).reshape(B).contiguous()
    out = torch.empty_like(x0)
    check(_lib.lib().dcb_q_sample(ptr(x0), ptr(noise), ptr(a), ptr(c), _c_i(B), _c_ll(x0.numel() // B), ptr(out),
                                  _stream()), "dcb_q_sample")
    return out


def video_to_uint8(video):
    """Decoded frames fp32 [B, C, T, H, W] in [-1, 1] -> uint8 [B, T, H, W, C] (what torchvision.io.write_video takes)."""
    require_cuda(video)
    v = video.float().contiguous()
    B, C, T, H, W = v.shape
    out = torch.empty((B, T, H, W, C), dtype=torch.uint8, device=v.device)
    check(_lib.lib().dcb_video_to_uint8(ptr(v), _c_i(B), _c_i(C), _c_i(T), _c_ll(H * W), ptr(out), _stream()),
          "dcb_video_to_uint8")
    return out


def gelu_bf16(x):
    """fp32 contiguous -> bf16 gelu(x) (exact erf)."""
    require_cuda(x)
    assert x.dtype == F32 and x.is_contiguous() and x.numel() % 8 == 0
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(_lib.lib().dcb_gelu_bf16(ptr(x), _c_ll(x.numel()), ptr(out), _stream()), "dcb_gelu_bf16")
    return out
