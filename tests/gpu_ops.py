"""Thin torch <-> C-ABI glue for the kernel-level GPU tests (calls dd3d_op_* with raw device pointers)."""
import ctypes as C

import torch

from dd3d_b200 import lib


ACT = torch.bfloat16  # 16-bit storage type of the operator entry points (set_act_dtype)


def set_act_dtype(name):
    """"bf16" | "fp16": element type of the dd3d_op_* entry points (process-wide dd3d_set_conv_policy("op_fp16"))."""
    global ACT
    ACT = torch.float16 if name == "fp16" else torch.bfloat16
    assert lib.load().dd3d_set_conv_policy(b"op_fp16", int(name == "fp16")) == 0


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_conv_weight(w):
    """[cout][cin][k][k] fp32 -> bf16 [cout_pad16][k*k][cin_pad64] (engine layout, csrc/engine.cu conv_layer)."""
    cout, cin, k, _ = w.shape
    cout_pad = (cout + 15) // 16 * 16
    cin_pad = (cin + 63) // 64 * 64
    out = torch.zeros(cout_pad, k * k, cin_pad, dtype=torch.float32)
    out[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
    return out.to(ACT).contiguous()


def pad16(v, fill):
    n = (v.numel() + 15) // 16 * 16
    out = torch.full((n, ), fill, dtype=torch.float32)
    out[:v.numel()] = v
    return out


def conv2d(x_nhwc, w, scale, bias, stride=1, relu=False, residual=None, res_up2=False, out_f32=False, in_slice=None,
           out_pitch=None, out_offset=0):
    """x_nhwc: bf16 [B,H,W,pitch] cuda (in_slice = (c0, cin) selects a channel slice); w: fp32 [cout,cin,k,k] cpu.
    Returns the output tensor (bf16 [B,Ho,Wo,out_pitch] or fp32 [B,Ho,Wo,cout_pad])."""
    L = lib.load()
    B, H, W, pitch = x_nhwc.shape
    c0, cin = in_slice if in_slice is not None else (0, pitch)
    cout, _, k, _ = w.shape
    wp = pack_conv_weight(w).cuda()
    sc = pad16(scale, 1.0).cuda()
    bi = pad16(bias, 0.0).cuda()
    Ho, Wo = H // stride, W // stride
    cout_pad = (cout + 15) // 16 * 16
    if out_f32:
        out = torch.full((B, Ho, Wo, cout_pad), float("nan"), dtype=torch.float32, device="cuda")
        op = cout_pad
        out_ptr = out.data_ptr()
    else:
        op = out_pitch or cout
        out = torch.full((B, Ho, Wo, op), 7.0, dtype=ACT, device="cuda")
        out_ptr = out.data_ptr() + 2 * out_offset
    in_ptr = x_nhwc.data_ptr() + 2 * c0
    res_ptr, res_pitch = None, 0
    if residual is not None:
        res_ptr, res_pitch = C.c_void_p(residual.data_ptr()), residual.shape[-1]
    st = L.dd3d_op_conv2d(C.c_void_p(in_ptr), B, H, W, cin, pitch, _p(wp), cout, k, stride, _p(sc), _p(bi), int(relu),
                          res_ptr, res_pitch, int(res_up2), C.c_void_p(out_ptr), op, int(out_f32), _stream())
    assert st == 0, f"dd3d_op_conv2d failed: {st}"
    torch.cuda.synchronize()
    return out


def conv2d_ref(x_nhwc, w, scale, bias, stride=1, relu=False, residual=None, res_up2=False, in_slice=None):
    """fp32 torch reference on the SAME bf16-rounded operands; returns fp32 NHWC (not rounded)."""
    import torch.nn.functional as F
    c0, cin = in_slice if in_slice is not None else (0, x_nhwc.shape[-1])
    x = x_nhwc[..., c0:c0 + cin].float().permute(0, 3, 1, 2)
    wq = w.to(ACT).float().to(x.device)
    k = w.shape[-1]
    y = F.conv2d(x, wq, None, stride, (k - 1) // 2)
    y = y * scale.to(x.device).view(1, -1, 1, 1) + bias.to(x.device).view(1, -1, 1, 1)
    if residual is not None:
        r = residual.float().permute(0, 3, 1, 2)
        if res_up2:
            r = F.interpolate(r, scale_factor=2.0, mode="nearest")
        y = y + r
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()
