"""Helpers for the GPU parity tests: ctypes calls into the kernel-level debug entry points
and torch fp32 references of the same ops (on bf16-rounded operands)."""
import ctypes as C

import numpy as np

from satellite_computervision_b200 import _lib


def bf16_round(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def conv3x3_device(x, k, b, relu=True, pooled=False, device=0):
    lib = _lib.load_library()
    N, H, W, Cin = x.shape
    Cout = k.shape[3]
    x = np.ascontiguousarray(x, np.float32)
    k = np.ascontiguousarray(k, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    y = np.empty((N, H, W, Cout), np.float32)
    p = np.empty((N, H // 2, W // 2, Cout), np.float32) if pooled else None
    _lib.check(lib.scv_debug_conv3x3(device, _lib.ptr(x), N, H, W, Cin, _lib.ptr(k), _lib.ptr(b), Cout, int(relu),
                                     _lib.ptr(y), _lib.ptr(p)))
    return (y, p) if pooled else y


def convT_device(x, k, b, relu=True, device=0):
    lib = _lib.load_library()
    N, H, W, Cin = x.shape
    Cout = k.shape[2]
    x = np.ascontiguousarray(x, np.float32)
    k = np.ascontiguousarray(k, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    y = np.empty((N, 2 * H, 2 * W, Cout), np.float32)
    _lib.check(lib.scv_debug_convT2x2(device, _lib.ptr(x), N, H, W, Cin, _lib.ptr(k), _lib.ptr(b), Cout, int(relu),
                                      _lib.ptr(y)))
    return y


def conv3x3_ref(x, k, b, relu=True):
    """torch fp32 reference (cuda if available, TF32 off) on bf16-rounded x and k."""
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = 'cuda' if torch.cuda.is_available() else 'cpu'
    xt = torch.from_numpy(bf16_round(x)).to(dev).permute(0, 3, 1, 2)
    kt = torch.from_numpy(bf16_round(k)).to(dev).permute(3, 2, 0, 1)
    y = F.conv2d(xt, kt, torch.from_numpy(np.asarray(b, np.float32)).to(dev), padding=1)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().cpu().numpy()


def convT_ref(x, k, b, relu=True):
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    dev = 'cuda' if torch.cuda.is_available() else 'cpu'
    xt = torch.from_numpy(bf16_round(x)).to(dev).permute(0, 3, 1, 2)
    kt = torch.from_numpy(bf16_round(k)).to(dev).permute(3, 2, 0, 1)  # (kh,kw,out,in) -> (in,out,kh,kw)
    y = F.conv_transpose2d(xt, kt, torch.from_numpy(np.asarray(b, np.float32)).to(dev), stride=2)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().cpu().numpy()


def maxpool_ref(y):
    N, H, W, Cc = y.shape
    return y.reshape(N, H // 2, 2, W // 2, 2, Cc).max(axis=(2, 4))


def err_stats(got, ref):
    """max |got - ref| relative to bf16 resolution of the reference magnitude."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    nan = int(np.isnan(got).sum())
    d = np.abs(np.nan_to_num(got, nan=1e30) - ref)
    tol = 2.0 ** -7 * np.abs(ref) + 2e-3  # one bf16 ulp (2^-8 rel) + accumulation-order slack
    return dict(max_abs=float(d.max()), n_bad=int((d > tol).sum()), n=int(d.size), nan=nan,
                ref_absmax=float(np.abs(ref).max()))
