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
    if _prof is not None:
        _prof["launches"] += 1
    return torch.cuda.current_stream().cuda_stream


_prof = None


def profile_begin():
    """Starts counting kernel launches and timing every dwm_b200_linear launch with a
    CUDA event pair on the launching stream (used by bench.py's roofline figure)."""
    global _prof
    _prof = {"launches": 0, "events": []}


def profile_end():
    """Stops profiling; call after a stream synchronize.  Returns
    {"launches": n, "linear": [{"ms", "flops", "shape", "epilogue"}, ...]}."""
    global _prof
    p, _prof = _prof, None
    out = {"launches": p["launches"], "linear": []}
    for e0, e1, flops, shape, epi in p["events"]:
        out["linear"].append({"ms": e0.elapsed_time(e1), "flops": flops,
                              "shape": shape, "epilogue": epi})
    return out


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
           qk_norm_regions=0, peer_out=None, resid=None, resid_row_mod=0, gate=None, blend_x=None, alpha=None,
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
    args.qk_norm_regions = qk_norm_regions
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
    if peer_out:
        # raw device pointers of the peers' buffers (same layout / pitch as `out`)
        if len(peer_out) > 8:
            raise ValueError("at most 8 peer buffers")
        for i, ptr in enumerate(peer_out):
            args.peer_out[i] = int(ptr)
        args.n_peer_out = len(peer_out)
    if _prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _l.load().dwm_b200_linear(ctypes.byref(args), _stream())
        e1.record()
        _prof["events"].append((e0, e1, 2.0 * M * N * K, (M, N, K), epilogue))
    else:
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


def _code(dtype):
    if dtype == torch.bfloat16:
        return _l.DWM_BF16
    if dtype == torch.float16:
        return _l.DWM_F16
    if dtype == torch.float32:
        return _l.DWM_F32
    raise TypeError("unsupported dtype {}".format(dtype))


def attention(qkv, out, *, D, heads, group_dims, group_strides, seq, inner=None,
              stride_outer=0, stride_inner=1, out_group_strides=None,
              out_stride_outer=None, out_stride_inner=None, split=0, out2=None,
              mask=None, mask_div=1, scale=None, kv=None, k_col=0, v_col=0,
              kv_group_strides=None, seq_kv=None, inner_kv=None,
              kv_stride_outer=0, kv_stride_inner=1):
    """Gathered multi-head attention over the fused q|k|v buffer (or a separate
    key/value buffer `kv`); see dwm_attention_args in include/dwm_b200.h."""
    _rows2d(qkv, "qkv")
    _rows2d(out, "out")
    if not qkv.is_cuda:
        raise RuntimeError("dwm_b200 kernels need CUDA tensors (no CPU fallback)")
    a = _l.AttentionArgs()
    a.qkv, a.ld, a.D = qkv.data_ptr(), qkv.stride(0), D
    a.heads, a.head_dim, a.dtype = heads, D // heads, _dt(qkv)
    gd = list(group_dims) + [1] * (3 - len(group_dims))
    gs = list(group_strides) + [0] * (3 - len(group_strides))
    ogs = gs if out_group_strides is None else \
        list(out_group_strides) + [0] * (3 - len(out_group_strides))
    for i in range(3):
        a.group_dims[i], a.group_strides[i], a.out_group_strides[i] = \
            gd[i], gs[i], ogs[i]
    a.seq, a.inner = seq, seq if inner is None else inner
    a.stride_outer, a.stride_inner = stride_outer, stride_inner
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    a.out_stride_outer = stride_outer if out_stride_outer is None \
        else out_stride_outer
    a.out_stride_inner = stride_inner if out_stride_inner is None \
        else out_stride_inner
    a.split = split
    if out2 is not None:
        _rows2d(out2, "out2")
        a.out2, a.ldo2 = out2.data_ptr(), out2.stride(0)
    if mask is not None:
        if mask.dtype != torch.uint8 or mask.dim() != 3 or not mask.is_contiguous():
            raise TypeError("mask must be a contiguous uint8 [B, n, n] tensor")
        a.mask, a.mask_div, a.n_outer = mask.data_ptr(), mask_div, mask.shape[-1]
    a.scale = (D // heads) ** -0.5 if scale is None else scale
    if kv is not None:
        _rows2d(kv, "kv")
        if kv.dtype != qkv.dtype:
            raise TypeError("kv dtype differs from q dtype")
        a.kv, a.ld_kv, a.k_col, a.v_col = kv.data_ptr(), kv.stride(0), k_col, v_col
        kgs = list(kv_group_strides) + [0] * (3 - len(kv_group_strides))
        for i in range(3):
            a.kv_group_strides[i] = kgs[i]
        a.seq_kv = seq if seq_kv is None else seq_kv
        a.inner_kv = a.seq_kv if inner_kv is None else inner_kv
        a.kv_stride_outer, a.kv_stride_inner = kv_stride_outer, kv_stride_inner
    _l.check(_l.load().dwm_b200_attention(ctypes.byref(a), _stream()),
             "dwm_b200_attention")
    return out


def layernorm(x, out, *, weight=None, bias=None, eps=1e-5, add_item=None,
              add_full=None, rows_per_item=0, sum_out=None, shift=None,
              scale=None, shift2=None, scale2=None, out2=None):
    """LayerNorm (+adds, +AdaLN modulation) of an fp32 stream into 16-bit out."""
    _rows2d(_f32(x, "x"), "x")
    _rows2d(out, "out")
    a = _l.LayerNormArgs()
    a.M, a.D = x.shape
    a.x, a.ldx = x.data_ptr(), x.stride(0)
    if add_item is not None:
        _rows2d(_f32(add_item, "add_item"), "add_item")
        a.add_item, a.add_item_ld = add_item.data_ptr(), add_item.stride(0)
    if add_full is not None:
        _rows2d(_f32(add_full, "add_full"), "add_full")
        a.add_full, a.add_full_ld = add_full.data_ptr(), add_full.stride(0)
    a.rows_per_item = rows_per_item
    if sum_out is not None:
        _rows2d(_f32(sum_out, "sum_out"), "sum_out")
        a.sum_out, a.ld_sum = sum_out.data_ptr(), sum_out.stride(0)
    a.weight, a.bias, a.eps = _ptr(_f32(weight, "weight")), _ptr(_f32(bias, "bias")), eps
    mod_ld = None
    for name, t in (("shift", shift), ("scale", scale), ("shift2", shift2),
                    ("scale2", scale2)):
        if t is not None:
            _rows2d(_f32(t, name), name)
            if mod_ld is not None and t.stride(0) != mod_ld:
                raise ValueError("modulation tensors must share a row pitch")
            mod_ld = t.stride(0)
            setattr(a, name, t.data_ptr())
    a.mod_ld = mod_ld or 0
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    if out2 is not None:
        _rows2d(out2, "out2")
        a.out2, a.ldo2 = out2.data_ptr(), out2.stride(0)
    a.dtype = _dt(out)
    _l.check(_l.load().dwm_b200_layernorm(ctypes.byref(a), _stream()),
             "dwm_b200_layernorm")
    return out


def act_cast(x, out, act=_l.ACT_NONE):
    _f32(x, "x")
    if not x.is_contiguous() or not out.is_contiguous():
        raise ValueError("act_cast needs contiguous tensors")
    _l.check(_l.load().dwm_b200_act_cast(
        x.data_ptr(), out.data_ptr(), x.numel(), act, _dt(out), _stream()),
        "dwm_b200_act_cast")
    return out


def sinusoid(t, channels, out, flip_sin_to_cos=True, downscale_freq_shift=0.0):
    _f32(t, "t")
    _rows2d(out, "out")
    _l.check(_l.load().dwm_b200_sinusoid(
        t.data_ptr(), t.numel(), channels, int(flip_sin_to_cos),
        float(downscale_freq_shift), out.data_ptr(), out.stride(0), _dt(out),
        _stream()), "dwm_b200_sinusoid")
    return out


def patchify(x, patch, out):
    _f32(x, "x")
    if x.dim() != 4 or not x.is_contiguous():
        raise ValueError("x must be contiguous [items, C, H, W]")
    _rows2d(out, "out")
    n, c, h, w = x.shape
    _l.check(_l.load().dwm_b200_patchify(
        x.data_ptr(), n, c, h, w, patch, out.data_ptr(), out.stride(0),
        _dt(out), _stream()), "dwm_b200_patchify")
    return out


def cfg_euler_step(tokens, latents, idx, sigmas, *, cfg, guidance_scale, patch,
                   in_range=None, noise_pred=None, round_dtype=torch.float32):
    _rows2d(_f32(tokens, "tokens"), "tokens")
    _f32(latents, "latents")
    _f32(sigmas, "sigmas")
    if latents.dim() != 6 or not latents.is_contiguous():
        raise ValueError("latents must be contiguous [B,T,V,C,H,W]")
    if idx.dtype != torch.int32 or not idx.is_contiguous():
        raise TypeError("idx must be contiguous int32 [B,T,V]")
    if in_range is not None and (in_range.dtype != torch.uint8 or
                                 in_range.numel() != latents.shape[1]):
        raise TypeError("in_range must be uint8 [T]")
    B, T, V, C, H, W = latents.shape
    _l.check(_l.load().dwm_b200_cfg_euler_step(
        tokens.data_ptr(), tokens.stride(0), cfg, float(guidance_scale),
        B, T, V, C, H, W, patch, idx.data_ptr(), sigmas.data_ptr(),
        sigmas.numel(), _ptr(in_range), latents.data_ptr(), _ptr(noise_pred),
        _code(round_dtype), _stream()), "dwm_b200_cfg_euler_step")
    return latents


def euler_step_by_indices(model_output, sample, idx, sigmas, round_dtype=torch.float32):
    """sample (fp32, in place) += dsigma[idx] * model_output; idx int32 over the
    leading dims of sample."""
    _f32(model_output, "model_output")
    _f32(sample, "sample")
    _f32(sigmas, "sigmas")
    if not (model_output.is_contiguous() and sample.is_contiguous() and idx.is_contiguous()):
        raise ValueError("euler_step_by_indices needs contiguous tensors")
    if idx.dtype != torch.int32:
        raise TypeError("idx must be int32")
    n = sample.numel()
    inner = n // idx.numel()
    _l.check(_l.load().dwm_b200_euler_step_by_indices(
        model_output.data_ptr(), sample.data_ptr(), n, inner, idx.data_ptr(),
        sigmas.data_ptr(), sigmas.numel(), _code(round_dtype), _stream()),
        "dwm_b200_euler_step_by_indices")
    return sample


def pack_conv_weight(weight: torch.Tensor, dtype, pad_out_to=None, pad_in_to=None):
    """torch Conv3d/Conv2d weight [O, I, (kt,) kh, kw] -> tap-major [kt*kh*kw, O', I']
    (optionally zero-padding O / I)."""
    if weight.dim() == 4:
        weight = weight.unsqueeze(2)
    o, i, kt, kh, kw = weight.shape
    w = weight.detach().permute(2, 3, 4, 0, 1).reshape(kt * kh * kw, o, i)
    op = pad_out_to or o
    ip = pad_in_to or i
    if op != o or ip != i:
        wp = torch.zeros(kt * kh * kw, op, ip, device=w.device, dtype=w.dtype)
        wp[:, :o, :i] = w
        w = wp
    return w.to(dtype).contiguous()


def conv(x, weight, bias=None, *, kernel, epilogue=_l.EPI_F32, act=_l.ACT_NONE,
         out=None, resid=None, resid_rows_per_item=0, blend_x=None, alpha=None,
         rows_per_batch=0):
    """x: 16-bit channels-last [nb, tp, h, w, c_in]; weight: tap-major
    [kt*kh*kw, c_out, c_in]; returns [nb*(tp-kt+1)*h*w, c_out]."""
    if x.dim() != 5 or not x.is_contiguous() or not weight.is_contiguous():
        raise ValueError("conv: x must be contiguous [nb, tp, h, w, c]")
    if not x.is_cuda:
        raise RuntimeError("dwm_b200 kernels need CUDA tensors (no CPU fallback)")
    kt, kh, kw = kernel
    nb, tp, h, w, c_in = x.shape
    taps, c_out, c_in2 = weight.shape
    if taps != kt * kh * kw or c_in2 != c_in or weight.dtype != x.dtype:
        raise ValueError("conv: weight shape/dtype mismatch")
    rows = nb * (tp - kt + 1) * h * w
    want = x.dtype if epilogue == _l.EPI_STORE else torch.float32
    if out is None:
        out = torch.empty(rows, c_out, device=x.device, dtype=want)
    _rows2d(out, "out")
    if out.dtype != want or out.shape[0] < rows:
        raise TypeError("conv: bad out tensor")
    a = _l.ConvArgs()
    a.x, a.nb, a.tp, a.h, a.w, a.c_in = x.data_ptr(), nb, tp, h, w, c_in
    a.weight, a.kt, a.kh, a.kw, a.c_out = weight.data_ptr(), kt, kh, kw, c_out
    a.bias = _ptr(_f32(bias, "bias"))
    a.dtype, a.epilogue, a.act = _dt(x), epilogue, act
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    if resid is not None:
        _rows2d(_f32(resid, "resid"), "resid")
        a.resid, a.ldr = resid.data_ptr(), resid.stride(0)
        if resid_rows_per_item:
            a.resid_per_item, a.rows_per_item = 1, resid_rows_per_item
    if blend_x is not None:
        _rows2d(_f32(blend_x, "blend_x"), "blend_x")
        a.blend_x, a.ldx = blend_x.data_ptr(), blend_x.stride(0)
        a.alpha, a.rows_per_batch = _f32(alpha, "alpha").data_ptr(), rows_per_batch
    _l.check(_l.load().dwm_b200_conv(ctypes.byref(a), _stream()), "dwm_b200_conv")
    return out


def groupnorm_stats(x, groups, sums=None):
    """x fp32 channels-last [nb, T, H, W, C] -> double [nb, groups, 2] (sum, sum sq)."""
    _f32(x, "x")
    if x.dim() != 5 or not x.is_contiguous():
        raise ValueError("x must be contiguous [nb, T, H, W, C]")
    nb, T, H, W, C = x.shape
    if sums is None:
        sums = torch.empty(nb, groups, 2, device=x.device, dtype=torch.float64)
    _l.check(_l.load().dwm_b200_groupnorm_stats(
        x.data_ptr(), nb, T * H * W, C, groups, sums.data_ptr(), _stream()),
        "dwm_b200_groupnorm_stats")
    return sums


def spatialnorm_silu(x, sums, gamma, beta, out, *, groups, eps=1e-6, zy=None,
                     zb=None, out_t0=0, silu=True):
    """x fp32 [nb,T,H,W,C]; out 16-bit [nb,out_T,H,W,C]; zy/zb fp32 [nb,Tz,hz,wz,C]."""
    _f32(x, "x")
    nb, T, H, W, C = x.shape
    if out.dim() != 5 or not out.is_contiguous() or out.shape[2:] != x.shape[2:]:
        raise ValueError("out must be contiguous [nb, out_T, H, W, C]")
    Tz = hz = wz = 0
    if zy is not None:
        _f32(zy, "zy")
        _f32(zb, "zb")
        Tz, hz, wz = zy.shape[1:4]
    _l.check(_l.load().dwm_b200_spatialnorm_silu(
        x.data_ptr(), nb, T, H, W, C, groups, sums.data_ptr(), eps,
        _f32(gamma, "gamma").data_ptr(), _f32(beta, "beta").data_ptr(),
        _ptr(zy), _ptr(zb), Tz, hz, wz, int(silu), out.data_ptr(), out.shape[1],
        out_t0, _dt(out), _stream()), "dwm_b200_spatialnorm_silu")
    return out


def upsample_nearest(x, compress_time, dtype):
    _f32(x, "x")
    nb, T, H, W, C = x.shape
    To = T
    if compress_time and T > 1:
        To = 1 + 2 * (T - 1) if T % 2 == 1 else 2 * T
    out = torch.empty(nb, To, 2 * H, 2 * W, C, device=x.device, dtype=dtype)
    _l.check(_l.load().dwm_b200_upsample_nearest(
        x.data_ptr(), nb, T, H, W, C, int(compress_time), out.data_ptr(), _dt(out),
        _stream()), "dwm_b200_upsample_nearest")
    return out


def axpy(x, y, a=1.0):
    """y += a * x (fp32, in place)."""
    _f32(x, "x")
    _f32(y, "y")
    if x.numel() != y.numel() or not (x.is_contiguous() and y.is_contiguous()):
        raise ValueError("axpy needs contiguous tensors of equal size")
    _l.check(_l.load().dwm_b200_axpy(x.data_ptr(), y.data_ptr(), x.numel(), float(a),
                                     _stream()), "dwm_b200_axpy")
    return y


def softmax_rows(x, scale, out):
    """out[r] = softmax(scale * x[r]) — fp32 [rows, cols] (row stride allowed) -> 16-bit."""
    _f32(x, "x")
    if x.dim() != 2 or out.dim() != 2 or x.shape != out.shape or x.stride(1) != 1 or \
            out.stride(1) != 1:
        raise ValueError("softmax_rows needs 2-D row-major tensors of equal shape")
    _l.check(_l.load().dwm_b200_softmax_rows(
        x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), float(scale), out.data_ptr(),
        out.stride(0), _dt(out), _stream()), "dwm_b200_softmax_rows")
    return out


def cfg_ddim_step(pred, latents, timesteps, alphas_cumprod, *, cfg, guidance_scale,
                  step_ratio, final_alpha_cumprod, prediction_type,
                  round_dtype=torch.float32):
    """Fused CFG + DDIM (eta 0) update of fp32 latents [B,T,V,...] in place."""
    _f32(pred, "pred")
    _f32(latents, "latents")
    _f32(alphas_cumprod, "alphas_cumprod")
    if not (pred.is_contiguous() and latents.is_contiguous() and timesteps.is_contiguous()):
        raise ValueError("cfg_ddim_step needs contiguous tensors")
    if timesteps.dtype != torch.int32:
        raise TypeError("timesteps must be int32")
    n_items = timesteps.numel()
    inner = latents.numel() // n_items
    code = {"epsilon": 0, "sample": 1, "v_prediction": 2}[prediction_type]
    _l.check(_l.load().dwm_b200_cfg_ddim_step(
        pred.data_ptr(), cfg, float(guidance_scale), n_items, inner,
        timesteps.data_ptr(), int(step_ratio), alphas_cumprod.data_ptr(),
        alphas_cumprod.numel(), float(final_alpha_cumprod), code, latents.data_ptr(),
        _code(round_dtype), _stream()), "dwm_b200_cfg_ddim_step")
    return latents


def lincomb2(x, y, s0, s1, out):
    """out = s0[item] * x + s1[item] * y with one coefficient pair per item."""
    for t, nme in ((x, "x"), (y, "y"), (s0, "s0"), (s1, "s1"), (out, "out")):
        _f32(t, nme)
        if not t.is_contiguous():
            raise ValueError("lincomb2 needs contiguous tensors")
    n = x.numel()
    _l.check(_l.load().dwm_b200_lincomb2(
        x.data_ptr(), y.data_ptr(), s0.data_ptr(), s1.data_ptr(), n, n // s0.numel(),
        out.data_ptr(), _stream()), "dwm_b200_lincomb2")
    return out
