"""Shared helpers for the -m gpu parity tests (test infrastructure)."""
import ctypes

import torch

import handmvnet_oracle as O
from handmvnet_b200 import HandMvNet, _lib
from handmvnet_b200.config import release_config


def rel_l2(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def build_pair(views=5, crop=True, precision="bf16", micro_batch=2, seed=0, randomize_norm=True, backbone="resnet"):
    """(product model on cuda:0, oracle cfg, state_dict) sharing identical weights."""
    cfg = release_config(views, crop, backbone=backbone)
    ocfg = O.release_config(views, crop, backbone)
    sd = O.make_state_dict(ocfg, seed=seed, randomize_norm=randomize_norm)
    m = HandMvNet(cfg["train"], cfg["model"], cfg["data"], precision=precision, micro_batch=micro_batch)
    m.load_state_dict(sd, strict=True)
    m.to("cuda:0").eval()
    m.freeze()
    m.prepare("cuda:0")
    return m, ocfg, sd


def conv_bn_act(precision, x, w, scale=None, shift=None, residual=None, stride=1, relu=True, iters=0):
    """hmv_conv_bn_act on CUDA tensors (NCHW fp32)."""
    lib = _lib.load()
    n, cin, h, wd = x.shape
    cout, _, k, _ = w.shape
    out = torch.empty(n, cout, h // stride, wd // stride, device=x.device, dtype=torch.float32)
    ms = ctypes.c_float(0)
    rc = lib.hmv_conv_bn_act(_lib.PRECISION[precision], _lib.ptr(x.contiguous()), _lib.ptr(w.contiguous()),
                             _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(residual), _lib.ptr(out), n, cin, h, wd, cout, k,
                             stride, int(relu), ctypes.byref(ms), iters,
                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "hmv_conv_bn_act")
    return out, ms.value
