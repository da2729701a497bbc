"""Python-side launch helpers over the C ABI (device pointers in, nothing allocated by the library).

Tensors are torch CUDA tensors used purely as device-memory handles; every op runs on
``torch.cuda.current_stream()`` so the whole path is CUDA-graph capturable.
"""
import ctypes
import functools

import torch

from . import _lib
from ._lib import TapGemmArgs, check, ptr

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
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


def tap_gemm(a, dims, strides, adims, taps, w, N, *, bias=None, group_bias=None, rows_per_group=0,
             residual=None, geglu=False, out_f32=None, out_bf16=None, ld_out=None, out_scale=1.0,
             block_n=0, box=None):
    """Raw launch.  a: bf16 tensor; dims=(D1,D2,D3) output domain; strides=(s1,s2,s3) element strides
    of A; adims=(A1,A2,A3) extents of A; taps=[(o1,o2,o3),...]; w: bf16 [N, ntaps*C]."""
    require_cuda(a, w, bias, group_bias, residual, out_f32, out_bf16)
    assert a.dtype == BF16 and w.dtype == BF16
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
    check(_lib.lib().dcb_tap_gemm(ctypes.byref(args), _stream()), "dcb_tap_gemm")


def _outputs(rows, n_out, out, want_f32, want_bf16, dev):
    of = ob = None
    if out is not None:
        if out.dtype == F32:
            of = out
        else:
            ob = out
        return of, ob, out.stride(0)
    if want_f32:
        of = torch.empty((rows, n_out), dtype=F32, device=dev)
    if want_bf16:
        ob = torch.empty((rows, n_out), dtype=BF16, device=dev)
    return of, ob, n_out


def linear(a, w, *, bias=None, residual=None, geglu=False, f32=False, bf16=True, out=None, N=None,
           group_bias=None, rows_per_group=0, out_scale=1.0, block_n=0):
    """a: bf16 [M, K] (row stride may exceed K); w: bf16 [N, K] -> [M, N] (or N/2 for geglu)."""
    M, K = a.shape
    assert a.stride(1) == 1 and w.shape[1] == K
    N = N or w.shape[0]
    n_out = N // 2 if geglu else N
    of, ob, ld = _outputs(M, n_out, out, f32, bf16, a.device)
    s1 = a.stride(0)
    tap_gemm(a, (M, 1, 1), (s1, s1 * M, s1 * M), (M, 1, 1), [(0, 0, 0)], w, N, bias=bias, residual=residual,
             geglu=geglu, out_f32=of, out_bf16=ob, ld_out=ld, group_bias=group_bias,
             rows_per_group=rows_per_group, out_scale=out_scale, block_n=block_n)
    return of if (of is not None and ob is None) else (ob if of is None else (of, ob))


TAPS_3X3 = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
TAPS_T3 = [(0, dt, 0) for dt in (-1, 0, 1)]


def conv3x3(a, nf, H, W, w, *, bias=None, residual=None, group_bias=None, rows_per_group=0,
            f32=True, bf16=False, out=None, N=None):
    """a: bf16 [nf*H*W, C] channels-last frames; w: bf16 [N, 9*C] ordered (ky, kx, c). Stride 1, pad 1."""
    C = a.shape[1]
    assert a.stride(1) == 1 and a.stride(0) == C and w.shape[1] == 9 * C
    N = N or w.shape[0]
    of, ob, ld = _outputs(nf * H * W, N, out, f32, bf16, a.device)
    tap_gemm(a, (W, H, nf), (C, C * W, C * W * H), (W, H, nf), TAPS_3X3, w, N, bias=bias, residual=residual,
             group_bias=group_bias, rows_per_group=rows_per_group, out_f32=of, out_bf16=ob, ld_out=ld)
    return of if (of is not None and ob is None) else (ob if of is None else (of, ob))


def tconv3(a, B, T, HW, w, *, bias=None, residual=None, f32=True, bf16=False, out=None):
    """Temporal (3,1,1) conv.  a: bf16 [B*T*HW, C]; w: bf16 [N, 3*C] ordered (kt, c). Pad 1 frame."""
    C = a.shape[1]
    assert a.stride(1) == 1 and a.stride(0) == C and w.shape[1] == 3 * C
    N = w.shape[0]
    of, ob, ld = _outputs(B * T * HW, N, out, f32, bf16, a.device)
    tap_gemm(a, (HW, T, B), (C, C * HW, C * HW * T), (HW, T, B), TAPS_T3, w, N, bias=bias, residual=residual,
             out_f32=of, out_bf16=ob, ld_out=ld)
    return of if (of is not None and ob is None) else (ob if of is None else (of, ob))
