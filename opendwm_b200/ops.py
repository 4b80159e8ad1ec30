"""Thin torch-tensor front end of the C ABI: validates shapes/dtypes, passes raw
device pointers + the current CUDA stream.  PyTorch only owns the memory."""
import ctypes

import torch

from . import lib as _l


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return _l.DWM_BF16
    if t.dtype == torch.float16:
        return _l.DWM_F16
    raise TypeError("expected a bf16/fp16 tensor, got {}".format(t.dtype))


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t, name):
    if t is not None:
        if t.dtype != torch.float32 or not t.is_cuda:
            raise TypeError("{} must be a cuda fp32 tensor".format(name))
        if t.stride(-1) != 1:
            raise ValueError("{} must be contiguous in its last dim".format(name))
    return t


def _rows2d(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("{} must be 2-D with contiguous rows".format(name))
    return t


def linear(a, w, bias=None, *, epilogue=_l.EPI_STORE, act=_l.ACT_NONE,
           out=None, rows_per_item=0, out_item_stride=0, out_row_offset=0,
           q_norm_weight=None, k_norm_weight=None, qk_region=0, eps=1e-6,
           resid=None, resid_row_mod=0, gate=None, blend_x=None, alpha=None,
           rows_per_batch=0):
    """out = epilogue(a @ w.T).  a [M,K], w [N,K] 16-bit; see include/dwm_b200.h."""
    _rows2d(a, "a")
    _rows2d(w, "w")
    if not (a.is_cuda and w.is_cuda):
        raise RuntimeError("dwm_b200 kernels need CUDA tensors (no CPU fallback)")
    if a.dtype != w.dtype:
        raise TypeError("a and w dtypes differ: {} vs {}".format(a.dtype, w.dtype))
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise ValueError("K mismatch: a {} vs w {}".format(K, K2))
    out_cols = N // 2 if epilogue == _l.EPI_GEGLU else N
    if out is None:
        if epilogue in (_l.EPI_RESID, _l.EPI_F32):
            out = torch.empty((M, out_cols), device=a.device, dtype=torch.float32)
        else:
            out = torch.empty((M, out_cols), device=a.device, dtype=a.dtype)
    _rows2d(out, "out")
    want = torch.float32 if epilogue in (_l.EPI_RESID, _l.EPI_F32) else a.dtype
    if out.dtype != want:
        raise TypeError("out dtype {} != {}".format(out.dtype, want))
    if out.shape[1] < out_cols:
        raise ValueError("out has {} columns, need {}".format(out.shape[1], out_cols))
    args = _l.LinearArgs()
    args.M, args.N, args.K = M, N, K
    args.A, args.lda = a.data_ptr(), a.stride(0)
    args.W, args.ldw = w.data_ptr(), w.stride(0)
    args.bias = _ptr(_f32(bias, "bias"))
    args.dtype = _dt(a)
    args.epilogue, args.act = epilogue, act
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    args.rows_per_item = rows_per_item
    args.out_item_stride = out_item_stride
    args.out_row_offset = out_row_offset
    args.q_norm_weight = _ptr(_f32(q_norm_weight, "q_norm_weight"))
    args.k_norm_weight = _ptr(_f32(k_norm_weight, "k_norm_weight"))
    args.qk_region = qk_region
    args.eps = eps
    if resid is not None:
        _rows2d(_f32(resid, "resid"), "resid")
        args.resid, args.ldr = resid.data_ptr(), resid.stride(0)
    args.resid_row_mod = resid_row_mod
    if gate is not None:
        _rows2d(_f32(gate, "gate"), "gate")
        args.gate, args.gate_ld = gate.data_ptr(), gate.stride(0)
    if blend_x is not None:
        _rows2d(_f32(blend_x, "blend_x"), "blend_x")
        args.blend_x, args.ldx = blend_x.data_ptr(), blend_x.stride(0)
    args.alpha = _ptr(_f32(alpha, "alpha"))
    args.rows_per_batch = rows_per_batch
    rc = _l.load().dwm_b200_linear(ctypes.byref(args), _stream())
    _l.check(rc, "dwm_b200_linear")
    return out


def pack_geglu(weight: torch.Tensor, bias=None, block=128):
    """Re-orders a diffusers GEGLU projection (rows [0,F) value, [F,2F) gate,
    FeedForward net.0.proj) into blocks [128 value | 128 gate] so one 256-column
    GEMM tile holds matching value/gate columns."""
    two_f = weight.shape[0]
    f = two_f // 2
    if f % block:
        raise ValueError("GEGLU inner dim {} not a multiple of {}".format(f, block))
    idx = torch.arange(two_f, device=weight.device).view(2, f // block, block)
    idx = idx.permute(1, 0, 2).reshape(-1)
    w = weight.index_select(0, idx).contiguous()
    b = None if bias is None else bias.index_select(0, idx).contiguous()
    return w, b
