"""CPU oracle for the HandMvNet inference forward path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 restatement of the reference algorithm
(pyxploiter/HandMvNet, `src/models`).  It is the checker used by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py`.  Nothing under `handmvnet_b200/` imports it, and the product path
never falls back to it.

Pinning: the reference ships no tests and no golden vectors, so
the oracle is pinned against outputs of the reference itself, executed in the
build container by `oracle/gen_golden.py`, and committed as
`tests/golden/*.npz`.  `tests/test_oracle_golden.py` re-derives every fixture
from `make_state_dict` + `make_inputs` on CPU.

All arithmetic is fp32 and uses the same torch ops the reference calls
(`F.conv2d`, `F.batch_norm`, `F.softmax`, `F.grid_sample`, `F.layer_norm`,
`F.gelu`), cited per function as reference `file:line`.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# reference src/constants.py:37-41 (20-edge MANO / mediapipe hand skeleton)
HAND_EDGES = [(0, 1), (1, 2), (2, 3), (3, 4),
              (0, 5), (5, 6), (6, 7), (7, 8),
              (0, 9), (9, 10), (10, 11), (11, 12),
              (0, 13), (13, 14), (14, 15), (15, 16),
              (0, 17), (17, 18), (18, 19), (19, 20)]

NUM_JOINTS = 21
# reference backbones/hrnet.py:427-495: branch widths; stage2..4 = (modules, branches), 4 BasicBlocks per branch
HR_CHANNELS = {"w40": (40, 80, 160, 320), "w64": (64, 128, 256, 512)}
HR_STAGES = ((1, 2), (4, 3), (3, 4))
BN_EPS = 1e-5
LN_EPS = 1e-5
LAYER_BLOCKS = (3, 4, 6)           # resnet.py:352 layers=[3,4,6,3], variant "paper" drops layer4
LAYER_PLANES = (64, 128, 256)
LAYER_STRIDES = (1, 2, 1)          # resnet.py:164-177 (paper variant: layer3 stride 1)


# ----------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------
def release_config(num_views: int = 5, crop: bool = True, backbone: str = "resnet") -> dict:
    """The fields of configs/release/{HO3D,DexYCB,MVHand}_HandMvNet[_HR][_wo_cam].yaml that the
    forward path reads (reference src/config.py:35-51 derives num_views).  backbone="hrnet" gives the
    `*_HR*.yaml` family (HRNet-w40, four feature levels)."""
    pos_enc = ["pos2d", "crop", "sin"] if crop else ["pos2d", "sin"]
    if backbone == "hrnet":
        return {
            "data": {"name": "ho3d", "batch_size": 16, "heatmap_size": 32, "image_size": 256},
            "model": {"selected_views": list(range(num_views)), "num_views": num_views,
                      "fusion": "cross_attn", "fusion_layers": 5, "pos_enc": pos_enc,
                      "use_gcn": True, "backbone": "hrnet", "backbone_type": "w40",
                      "backbone_pretrained_path": "", "backbone_channels": list(HR_CHANNELS["w40"]),
                      "backbone_pretrained": False},
            "train": {"debug": False, "root_relative": True, "device": "cpu"},
        }
    return {
        "data": {"name": "ho3d", "batch_size": 16, "heatmap_size": 32, "image_size": 256},
        "model": {"selected_views": list(range(num_views)), "num_views": num_views,
                  "fusion": "cross_attn", "fusion_layers": 5, "pos_enc": pos_enc,
                  "use_gcn": True, "backbone": "resnet", "backbone_type": "50_paper",
                  "backbone_early_return": 3, "backbone_channels": [1024],
                  "backbone_pretrained": False},
        "train": {"debug": False, "root_relative": True, "device": "cpu"},
    }


def feat_dim_of(cfg: dict) -> int:
    """reference handmvnet.py:88-95"""
    d = int(sum(cfg["model"]["backbone_channels"]) / 2)
    pe = cfg["model"].get("pos_enc", ["pos2d", "sin"])
    if "pos2d" in pe:
        d += 2
    if "crop" in pe:
        d += 10
    return d


# ----------------------------------------------------------------------------
# deterministic weights / inputs (portable: depends only on torch's CPU RNG)
# ----------------------------------------------------------------------------
def _state_dict_spec(cfg: dict):
    """(key, shape, kind) for the 355 keys of the reference state_dict
    (verified against the live reference by gen_golden.py)."""
    spec = []

    def conv(name, cout, cin, k, bias, kind):
        spec.append((name + ".weight", (cout, cin, k, k), kind))
        if bias:
            spec.append((name + ".bias", (cout,), "bias_fanin:%d" % (cin * k * k)))

    def bn(name, c):
        spec.append((name + ".weight", (c,), "bn_gamma"))
        spec.append((name + ".bias", (c,), "bn_beta"))
        spec.append((name + ".running_mean", (c,), "bn_mean"))
        spec.append((name + ".running_var", (c,), "bn_var"))
        spec.append((name + ".num_batches_tracked", (), "count"))

    if cfg["model"].get("backbone", "resnet") == "hrnet":
        return _hrnet_state_dict_spec(cfg, spec, conv, bn)
    conv("backbone.conv1", 64, 3, 7, False, "kaiming_out")
    bn("backbone.bn1", 64)
    inplanes = 64
    for li, (nblk, planes, stride) in enumerate(zip(LAYER_BLOCKS, LAYER_PLANES, LAYER_STRIDES), 1):
        for b in range(nblk):
            p = f"backbone.layer{li}.{b}"
            conv(p + ".conv1", planes, inplanes, 1, False, "kaiming_out"); bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3, False, "kaiming_out"); bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1, False, "kaiming_out"); bn(p + ".bn3", planes * 4)
            if b == 0 and (stride != 1 or inplanes != planes * 4):
                conv(p + ".downsample.0", planes * 4, inplanes, 1, False, "kaiming_out")
                bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    c = cfg["model"]["backbone_channels"][0]
    conv("pose_net.0", 512, c, 1, True, "default_conv"); bn("pose_net.1", 512)
    conv("pose_net.3", NUM_JOINTS, 512, 1, True, "default_conv")
    conv("sample_nets.0.conv.0", c // 2, c, 1, True, "default_conv"); bn("sample_nets.0.conv.1", c // 2)
    _head_spec(cfg, spec)
    return spec


def _hrnet_state_dict_spec(cfg, spec, conv, bn):
    """Keys of the `*_HR*` models in the reference's registration order (backbones/hrnet.py:241-276 HighResolutionNet,
    :97-120 HighResolutionModule, :26-36 BasicBlock, :61-72 Bottleneck; handmvnet.py:41-56 pose_net, :97 sample_nets);
    1941 keys for w40, verified key-for-key by gen_golden.py.  Every HRNet conv is bias-free and kaiming(fan_out)
    initialised (hrnet.py:412-418)."""
    ch = tuple(cfg["model"]["backbone_channels"])
    conv("backbone.conv1", 64, 3, 3, False, "kaiming_out"); bn("backbone.bn1", 64)
    conv("backbone.conv2", 64, 64, 3, False, "kaiming_out"); bn("backbone.bn2", 64)
    inpl = 64
    for b in range(4):                                   # stage1: 4 Bottlenecks, 64 planes
        p = f"backbone.layer1.{b}"
        conv(p + ".conv1", 64, inpl, 1, False, "kaiming_out"); bn(p + ".bn1", 64)
        conv(p + ".conv2", 64, 64, 3, False, "kaiming_out"); bn(p + ".bn2", 64)
        conv(p + ".conv3", 256, 64, 1, False, "kaiming_out"); bn(p + ".bn3", 256)
        if b == 0:
            conv(p + ".downsample.0", 256, inpl, 1, False, "kaiming_out"); bn(p + ".downsample.1", 256)
        inpl = 256
    pre = [256]
    for si, (nmod, nbr) in enumerate(HR_STAGES, start=2):
        cur = list(ch[:nbr])
        t = f"backbone.transition{si - 1}"
        for i in range(nbr):                             # hrnet.py:318-345
            if i < len(pre):
                if cur[i] != pre[i]:
                    conv(f"{t}.{i}.0", cur[i], pre[i], 3, False, "kaiming_out"); bn(f"{t}.{i}.1", cur[i])
            else:
                for j in range(i + 1 - len(pre)):
                    outc = cur[i] if j == i - len(pre) else pre[-1]
                    conv(f"{t}.{i}.{j}.0", outc, pre[-1], 3, False, "kaiming_out"); bn(f"{t}.{i}.{j}.1", outc)
        for m in range(nmod):
            sname = f"backbone.stage{si}.{m}"
            for i in range(nbr):
                for blk in range(4):
                    q = f"{sname}.branches.{i}.{blk}"
                    conv(q + ".conv1", cur[i], cur[i], 3, False, "kaiming_out"); bn(q + ".bn1", cur[i])
                    conv(q + ".conv2", cur[i], cur[i], 3, False, "kaiming_out"); bn(q + ".bn2", cur[i])
            for i in range(nbr):                         # hrnet.py:177-211
                for j in range(nbr):
                    f = f"{sname}.fuse_layers.{i}.{j}"
                    if j > i:
                        conv(f + ".0", cur[i], cur[j], 1, False, "kaiming_out"); bn(f + ".1", cur[i])
                    elif j < i:
                        for k in range(i - j):
                            outc = cur[i] if k == i - j - 1 else cur[j]
                            conv(f"{f}.{k}.0", outc, cur[j], 3, False, "kaiming_out"); bn(f"{f}.{k}.1", outc)
        pre = cur
    spec.append(("pose_net.weight", (NUM_JOINTS, ch[0], 3, 3), "default_conv"))
    spec.append(("pose_net.bias", (NUM_JOINTS,), "bias_fanin:%d" % (ch[0] * 9)))
    for l, c in enumerate(ch):
        conv(f"sample_nets.{l}.conv.0", c // 2, c, 1, True, "default_conv"); bn(f"sample_nets.{l}.conv.1", c // 2)
    _head_spec(cfg, spec)
    return spec


def _head_spec(cfg, spec):
    """fusion transformer + graph head keys (shared by both backbones)."""
    d = feat_dim_of(cfg)
    for i in range(cfg["model"].get("fusion_layers", 5)):
        p = f"joints_late_fusion.attn_fusion.{i}"
        for nm in ("to_q", "to_k", "to_v"):
            spec.append((f"{p}.{nm}.weight", (1024, d), "default_linear"))
        spec.append((f"{p}.to_out.weight", (d, 1024), "default_linear"))
        spec.append((f"{p}.to_out.bias", (d,), "bias_fanin:1024"))
        for nm in ("norm1", "norm2", "ff.net.0"):
            spec.append((f"{p}.{nm}.weight", (d,), "ln_gamma"))
            spec.append((f"{p}.{nm}.bias", (d,), "ln_beta"))
        spec.append((f"{p}.ff.net.1.weight", (128, d), "default_linear"))
        spec.append((f"{p}.ff.net.1.bias", (128,), "bias_fanin:%d" % d))
        spec.append((f"{p}.ff.net.4.weight", (d, 128), "default_linear"))
        spec.append((f"{p}.ff.net.4.bias", (d,), "bias_fanin:128"))
    for i, (cin, cout) in enumerate(((d, 256), (256, 64), (64, 3)), 1):
        spec.append((f"joints_decoder.joints_gcn{i}.weight", (3, 1, cin, cout), "xavier_cheb"))
        spec.append((f"joints_decoder.joints_gcn{i}.bias", (1, 1, cout), "cheb_bias"))


def make_state_dict(cfg: dict, seed: int = 0, randomize_norm: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic random-init weights with the reference's key names, shapes and
    init distributions (resnet.py:181-187 kaiming fan_out; torch defaults for
    Conv2d/Linear; layers.py:376-381 xavier_normal / zero bias for ChebConv).

    randomize_norm=True additionally randomises BN statistics/affine, LN affine and
    the ChebConv biases so that a wrong BN fold or a dropped bias is visible
    (random-init BN is the identity otherwise).
    """
    sd = OrderedDict()
    for idx, (key, shape, kind) in enumerate(_state_dict_spec(cfg)):
        g = torch.Generator().manual_seed(1_000_003 * (seed + 1) + idx)
        if kind == "kaiming_out":
            cout, cin, k, _ = shape
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (cout * k * k))
        elif kind in ("default_conv", "default_linear"):
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind.startswith("bias_fanin:"):
            bound = 1.0 / math.sqrt(int(kind.split(":")[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "xavier_cheb":
            k, _, cin, cout = shape
            fan_in, fan_out = 1 * cin * cout, k * cin * cout
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (fan_in + fan_out))
        elif kind == "cheb_bias":
            t = torch.randn(shape, generator=g) * 1e-3 if randomize_norm else torch.zeros(shape)
        elif kind in ("bn_gamma", "ln_gamma"):        # U(0.75, 1.25): visible in a wrong fold, yet well conditioned
            t = torch.rand(shape, generator=g) * 0.5 + 0.75 if randomize_norm else torch.ones(shape)
            if randomize_norm and ".branches." in key and key.endswith(".bn2.weight"):
                # HRNet: 32 residual BasicBlocks in a row double the activation variance each with unit gammas (1e8-1e9 at
                # the outputs: every softmax downstream degenerates into a hard max); a small last-BN gain per block - what
                # trained residual nets converge to - keeps the synthetic checkpoint O(10) and the comparison meaningful
                t = t * 0.15
            if randomize_norm and ".fuse_layers." in key and key.endswith(".1.weight"):
                t = t * 0.5                          # same for the 2-4 terms every fusion sum adds up, 8 modules deep
        elif kind in ("bn_beta", "ln_beta", "bn_mean"):
            t = torch.randn(shape, generator=g) * 0.1 if randomize_norm else torch.zeros(shape)
        elif kind == "bn_var":
            t = torch.rand(shape, generator=g) * 0.5 + 0.75 if randomize_norm else torch.ones(shape)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.int64)
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[key] = t
    return sd


def make_inputs(batch: int, num_views: int, seed: int = 1234, image_size: int = 256):
    """Synthetic inputs of SURVEY.md §8d config 2: x ~ N(0,1); square xyxy boxes inside
    a 640x480 frame; fx=fy~U(500,700), cx=320±20, cy=240±20."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, num_views, 3, image_size, image_size, generator=g)
    cx = 160 + 320 * torch.rand(batch, num_views, generator=g)
    cy = 120 + 240 * torch.rand(batch, num_views, generator=g)
    side = 100 + 150 * torch.rand(batch, num_views, generator=g)
    bbox = torch.stack([cx - side / 2, cy - side / 2, cx + side / 2, cy + side / 2], dim=-1)
    f = 500 + 200 * torch.rand(batch, num_views, generator=g)
    px = 320 + 40 * (torch.rand(batch, num_views, generator=g) - 0.5)
    py = 240 + 40 * (torch.rand(batch, num_views, generator=g) - 0.5)
    intr = torch.stack([f, f, px, py], dim=-1)
    return x, bbox, intr


# ----------------------------------------------------------------------------
# backbone (reference backbones/resnet.py)
# ----------------------------------------------------------------------------
def _bn(sd, name, x):
    """eval-mode BatchNorm2d (resnet.py:115 / layers.py:330)."""
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, BN_EPS)


def bottleneck(sd, p, x, stride, taps=None):
    """reference resnet.py:124-144 (Bottleneck.forward)."""
    key = p[len("backbone."):]
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    if taps is not None:
        taps[key + ".conv1"] = out
    out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    if taps is not None:
        taps[key + ".conv2"] = out
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
        if taps is not None:
            taps[key + ".downsample"] = x
    out = F.relu(out + x)
    if taps is not None:
        taps[key + ".conv3"] = out
    return out


def backbone(sd, x, taps=None, per_layer=False):
    """reference resnet.py:216-239, variant "paper": conv7x7/2, BN, ReLU, maxpool3x3/2,
    layer1..layer3 (layer3 stride 1) -> [N,1024,H/8,W/8].  per_layer=True also records every conv output
    (after its BN / ReLU / residual) under the product's step names."""
    x = F.relu(_bn(sd, "backbone.bn1", F.conv2d(x, sd["backbone.conv1.weight"], stride=2, padding=3)))
    if taps is not None:
        taps["conv1"] = x
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps["maxpool"] = x
        taps["stem"] = x
    for li, (nblk, stride) in enumerate(zip(LAYER_BLOCKS, LAYER_STRIDES), 1):
        for b in range(nblk):
            x = bottleneck(sd, f"backbone.layer{li}.{b}", x, stride if b == 0 else 1, taps if per_layer else None)
        if taps is not None:
            taps[f"layer{li}"] = x
    return x


def _conv_bn(sd, conv, bn, x, stride=1, relu=True):
    """bias-free conv (kernel from the weight shape, "same" padding) + eval BatchNorm (+ ReLU)."""
    w = sd[conv + ".weight"]
    y = _bn(sd, bn, F.conv2d(x, w, stride=stride, padding=w.shape[-1] // 2))
    return F.relu(y) if relu else y


def basic_block(sd, p, x, taps=None, key=None):
    """reference hrnet.py:38-55 (BasicBlock.forward, no downsample in HRNet branches)."""
    out = _conv_bn(sd, p + ".conv1", p + ".bn1", x)
    if taps is not None:
        taps[key + ".conv1"] = out
    out = _conv_bn(sd, p + ".conv2", p + ".bn2", out, relu=False)
    out = F.relu(out + x)
    if taps is not None:
        taps[key + ".conv2"] = out
    return out


def hr_module(sd, p, xs, taps=None, key=None):
    """reference hrnet.py:216-233 (HighResolutionModule.forward): 4 BasicBlocks per branch, then every output branch i is
    relu(sum_j f_ij(x_j)) with f_ij = identity (j == i), 1x1 conv + BN + nearest upsample (j > i), or a chain of
    3x3 stride-2 conv + BN (+ ReLU except the last) (j < i) - summed in the reference's order."""
    nbr = len(xs)
    xs = list(xs)
    for i in range(nbr):
        for blk in range(4):
            xs[i] = basic_block(sd, f"{p}.branches.{i}.{blk}", xs[i], taps, f"{key}.b{i}.{blk}")
        if taps is not None:
            taps[f"{key}.branch{i}"] = xs[i]
    outs = []
    for i in range(nbr):
        y = None
        for j in range(nbr):
            f = f"{p}.fuse_layers.{i}.{j}"
            if j == i:
                t = xs[j]
            elif j > i:
                t = _conv_bn(sd, f + ".0", f + ".1", xs[j], relu=False)
                if taps is not None:
                    taps[f"{key}.up{i}{j}"] = t
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:
                t = xs[j]
                for k in range(i - j):
                    t = _conv_bn(sd, f"{f}.{k}.0", f"{f}.{k}.1", t, stride=2, relu=k < i - j - 1)
                    if taps is not None and k < i - j - 1:
                        taps[f"{key}.down{i}{j}.{k}"] = t
            y = t if y is None else y + t
        outs.append(F.relu(y))
        if taps is not None:
            taps[f"{key}.out{i}"] = outs[-1]
    return outs


def hrnet_backbone(sd, x, taps=None):
    """reference hrnet.py:377-409 (HighResolutionNet.forward) for w40 / w64: two stride-2 3x3 stem convs, 4 Bottlenecks,
    then stages 2-4 with their transitions -> [N,C0,64,64], [N,C1,32,32], [N,C2,16,16], [N,C3,8,8]."""
    x = _conv_bn(sd, "backbone.conv1", "backbone.bn1", x, stride=2)
    if taps is not None:
        taps["hr.conv1"] = x
    x = _conv_bn(sd, "backbone.conv2", "backbone.bn2", x, stride=2)
    if taps is not None:
        taps["hr.conv2"] = x
    for b in range(4):
        blk_taps = {} if taps is not None else None
        x = bottleneck(sd, f"backbone.layer1.{b}", x, 1, blk_taps)
        if taps is not None:
            taps.update({"hr." + k: v for k, v in blk_taps.items()})
    if taps is not None:
        taps["hr.layer1"] = x
    ys = [x]
    for si, (nmod, nbr) in enumerate(HR_STAGES, start=2):
        t = f"backbone.transition{si - 1}"
        xs = []
        for i in range(nbr):
            if i < len(ys):
                if f"{t}.{i}.0.weight" in sd:            # hrnet.py:389-392 / :398-401: stage 2 feeds x, later stages y_list[-1]
                    xs.append(_conv_bn(sd, f"{t}.{i}.0", f"{t}.{i}.1", ys[0] if si == 2 else ys[-1]))
                else:
                    xs.append(ys[i])
            else:
                v = ys[0] if si == 2 else ys[-1]
                j = 0
                while f"{t}.{i}.{j}.0.weight" in sd:
                    v = _conv_bn(sd, f"{t}.{i}.{j}.0", f"{t}.{i}.{j}.1", v, stride=2)
                    j += 1
                xs.append(v)
        if taps is not None:
            for i, v in enumerate(xs):
                taps[f"hr.stage{si}.in{i}"] = v
        for m in range(nmod):
            xs = hr_module(sd, f"backbone.stage{si}.{m}", xs, taps, f"hr.stage{si}.{m}")
        ys = xs
    return ys


# ----------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------
def pose_net(sd, feat, taps=None):
    """reference layers.py:318-334 built at handmvnet.py:71: 1x1 conv+BN+ReLU, 1x1 conv."""
    h = F.relu(_bn(sd, "pose_net.1", F.conv2d(feat, sd["pose_net.0.weight"], sd["pose_net.0.bias"])))
    if taps is not None:
        taps["pose_hidden"] = h
    return F.conv2d(h, sd["pose_net.3.weight"], sd["pose_net.3.bias"])


def soft_argmax_2d(heatmap, temperature: float = 1000.0):
    """reference models/utils.py:35-62: softmax(T*hm) over H*W, then E[x], E[y]."""
    n, j, h, w = heatmap.shape
    p = F.softmax(heatmap.reshape(n, j, -1) * temperature, dim=2).reshape(n, j, h, w)
    px = p.sum(dim=2)          # marginal over rows  -> [n,j,w]
    py = p.sum(dim=3)          # marginal over cols  -> [n,j,h]
    ex = (px * torch.arange(w, dtype=torch.float32, device=heatmap.device)[None, None]).sum(dim=2, keepdim=True)
    ey = (py * torch.arange(h, dtype=torch.float32, device=heatmap.device)[None, None]).sum(dim=2, keepdim=True)
    return torch.cat((ex, ey), dim=2)


def sample_net(sd, feat, xy, taps=None, level=0):
    """reference nets.py:55-63 (+ :46-53): dense 1x1 conv+BN+ReLU, then bilinear
    grid_sample(align_corners=True, zero padding) at the joints.  The coordinates are heat-map pixels (0..31) and are
    normalised by THIS level's own width / height (nets.py:47-49): on the HRNet levels that are not 32 x 32 they
    therefore address the level's pixels with the same un-rescaled numbers (64 x 64: the top-left quadrant; 16 x 16 and
    8 x 8: mostly outside the map, i.e. zeros) - the reference's behaviour, reproduced as is."""
    q = f"sample_nets.{level}"
    f = F.relu(_bn(sd, q + ".conv.1", F.conv2d(feat, sd[q + ".conv.0.weight"], sd[q + ".conv.0.bias"])))
    h, w = f.shape[2:]
    gx = xy[:, :, 0] / (w - 1) * 2 - 1
    gy = xy[:, :, 1] / (h - 1) * 2 - 1
    grid = torch.stack((gx, gy), 2)[:, :, None, :]
    s = F.grid_sample(f, grid, align_corners=True)[:, :, :, 0]
    return s.permute(0, 2, 1).contiguous()


def sample_net_gather(sd, feat, xy):
    """Algebraic restatement used by the CUDA path: the conv is
    pointwise, so evaluate conv+BN+ReLU only at the <=4 bilinear neighbours."""
    n, c, h, w = feat.shape
    wgt = sd["sample_nets.0.conv.0.weight"][:, :, 0, 0]
    s = sd["sample_nets.0.conv.1.weight"] / torch.sqrt(sd["sample_nets.0.conv.1.running_var"] + BN_EPS)
    wf = wgt * s[:, None]
    bf = sd["sample_nets.0.conv.1.bias"] + (sd["sample_nets.0.conv.0.bias"] - sd["sample_nets.0.conv.1.running_mean"]) * s
    ix = ((xy[:, :, 0] / (w - 1) * 2 - 1) + 1) / 2 * (w - 1)
    iy = ((xy[:, :, 1] / (h - 1) * 2 - 1) + 1) / 2 * (h - 1)
    x0 = torch.floor(ix); y0 = torch.floor(iy)
    out = torch.zeros(n, xy.shape[1], wf.shape[0])
    fl = feat.permute(0, 2, 3, 1)
    for dy in (0, 1):
        for dx in (0, 1):
            xx = x0 + dx; yy = y0 + dy
            wt = (1 - (ix - xx).abs()) * (1 - (iy - yy).abs())
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            xi = xx.clamp(0, w - 1).long(); yi = yy.clamp(0, h - 1).long()
            px = fl[torch.arange(n)[:, None], yi, xi]            # [n, j, c]
            v = F.relu(px @ wf.t() + bf)
            out += v * (wt * ok)[:, :, None]
    return out


def crop_fov(bbox, intr):
    """reference handmvnet.py:205-222 + utils.py:134-171: 5 bbox points (4 corners, centre)
    minus the principal point, atan(./f), flattened (tx0,ty0,...,tx4,ty4)."""
    b = bbox.reshape(-1, 4).to(torch.float32)
    k = intr.reshape(-1, 4).to(torch.float32)
    pts = torch.stack([b[:, 0], b[:, 1], b[:, 0], b[:, 3], b[:, 2], b[:, 1], b[:, 2], b[:, 3],
                       (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2], dim=1).reshape(-1, 5, 2)
    tx = torch.atan((pts[:, :, 0] - k[:, 2:3]) / k[:, 0:1])
    ty = torch.atan((pts[:, :, 1] - k[:, 3:4]) / k[:, 1:2])
    return torch.stack((tx, ty), dim=2).flatten(start_dim=-2)   # [n, 10]


_CONST_CACHE = {}


def _const(key, device, build):
    """Constant tables live on the tensor's device and are built once (the reference re-copies `pe` and re-derives the
    Chebyshev basis on the CPU every forward, layers.py:157,393-394; caching only favours the baseline when the oracle is
    timed as eager PyTorch on a GPU)."""
    k = (key, str(device))
    if k not in _CONST_CACHE:
        _CONST_CACHE[k] = build().to(device)
    return _CONST_CACHE[k]


def positional_table(d_model: int, max_len: int):
    """reference layers.py:136-150 (even and odd d_model)."""
    pos = torch.arange(max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div) if d_model % 2 == 0 else torch.cos(pos * div[:-1])
    return pe


# ----------------------------------------------------------------------------
# fusion transformer (reference fusion.py:7-30, layers.py:177-237)
# ----------------------------------------------------------------------------
def attention_layer(sd, p, x, custom_query_length=0, heads=8):
    """reference layers.py:202-237 (eval mode: dropout is the identity)."""
    if custom_query_length > 0:
        q_in, kv_in = x[:, :custom_query_length], x[:, custom_query_length:]
    else:
        q_in, kv_in = x, x
    b, nq, _ = q_in.shape
    nk = kv_in.shape[1]
    q = (q_in @ sd[p + ".to_q.weight"].t()).reshape(b, nq, heads, -1).transpose(1, 2)
    k = (kv_in @ sd[p + ".to_k.weight"].t()).reshape(b, nk, heads, -1).transpose(1, 2)
    v = (kv_in @ sd[p + ".to_v.weight"].t()).reshape(b, nk, heads, -1).transpose(1, 2)
    scale = q.shape[-1] ** -0.5
    attn = F.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(b, nq, -1)
    o = o @ sd[p + ".to_out.weight"].t() + sd[p + ".to_out.bias"]
    d = x.shape[-1]
    h = F.layer_norm(o + q_in, (d,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], LN_EPS)
    f = F.layer_norm(h, (d,), sd[p + ".ff.net.0.weight"], sd[p + ".ff.net.0.bias"], LN_EPS)
    f = F.gelu(f @ sd[p + ".ff.net.1.weight"].t() + sd[p + ".ff.net.1.bias"])
    f = f @ sd[p + ".ff.net.4.weight"].t() + sd[p + ".ff.net.4.bias"]
    return F.layer_norm(f + h, (d,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], LN_EPS)


def fusion(sd, tokens, num_layers=5, add_pos=True, query_len=NUM_JOINTS, taps=None):
    """reference fusion.py:26-30: PE add, (L-1)/2 self, 1 cross (first query_len tokens
    query the rest), (L-1)/2 self."""
    x = tokens
    if add_pos:
        d_model, n_tok = x.shape[-1], x.shape[1]
        x = x + _const(("pe", d_model, n_tok), x.device, lambda: positional_table(d_model, n_tok))[None]
    if taps is not None:
        taps["tokens_pe"] = x
    half = (num_layers - 1) // 2
    for i in range(num_layers):
        x = attention_layer(sd, f"joints_late_fusion.attn_fusion.{i}", x,
                            custom_query_length=query_len if i == half else 0)
        if taps is not None:
            taps[f"fusion{i}"] = x
    return x


# ----------------------------------------------------------------------------
# graph head (reference nets.py:119-139, layers.py:363-445, utils.py:89-120)
# ----------------------------------------------------------------------------
def hand_adjacency():
    """utils.py:108-120 with sparse=False: symmetrise, add self loops, row-normalise."""
    a = np.zeros((NUM_JOINTS, NUM_JOINTS), dtype=np.float32)
    for i, j in HAND_EDGES:
        a[i, j] = 1.0
    a = np.maximum(a, a.T) + np.eye(NUM_JOINTS, dtype=np.float32)
    a = a / a.sum(axis=1, keepdims=True)
    return torch.tensor(a, dtype=torch.float32)


def cheb_basis(order: int = 3):
    """layers.py:405-445: L = I - D^-1/2 A D^-1/2 with D = diag(rowsum(A)) (= I here
    because A is row-normalised), T0 = I, T1 = L, Tk = 2 L T(k-1) - T(k-2)."""
    a = hand_adjacency()
    d = torch.diag(torch.sum(a, dim=-1) ** (-1 / 2))
    lap = torch.eye(NUM_JOINTS) - torch.mm(torch.mm(d, a), d)
    t = [torch.eye(NUM_JOINTS), lap]
    for _ in range(2, order):
        t.append(2 * torch.mm(lap, t[-1]) - t[-2])
    return torch.stack(t[:order])


def gcn_decoder(sd, x, taps=None):
    """reference nets.py:133-139: ChebConv(K=2) x3 with LeakyReLU(0.01) after 1 and 2."""
    t = _const("cheb3", x.device, lambda: cheb_basis(3)).unsqueeze(1)   # [3,1,21,21]
    for i in (1, 2, 3):
        w = sd[f"joints_decoder.joints_gcn{i}.weight"]   # [3,1,cin,cout]
        r = torch.matmul(torch.matmul(t, x), w)          # [3,B,21,cout]
        x = torch.sum(r, dim=0) + sd[f"joints_decoder.joints_gcn{i}.bias"]
        if i < 3:
            x = F.leaky_relu(x, 0.01)
        if taps is not None:
            taps[f"gcn{i}"] = x
    return x


# ----------------------------------------------------------------------------
# whole forward (reference handmvnet.py:158-266)
# ----------------------------------------------------------------------------
@torch.no_grad()
def forward(sd, cfg, x, bbox=None, intr=None, return_taps=False, teacher=None):
    """Returns the reference's output dict; with return_taps also every stage tensor.
    `teacher` may override a stage input by name ("backbone_out", "coords"): teacher-forced parity - everything
    downstream of the override is computed from it (soft-argmax at T=1000 is discontinuous, so the
    end-to-end bf16 check conditions both sides on the same joint coordinates)."""
    taps = OrderedDict() if return_taps else None
    b, v, c, h, w = x.shape
    nv = cfg["model"]["num_views"]
    if v != nv:
        raise ValueError(f"input has {v} views, model was built for {nv}")
    pe_list = cfg["model"].get("pos_enc", ["pos2d", "sin"])
    teacher = teacher or {}
    if cfg["model"].get("backbone", "resnet") == "hrnet":    # handmvnet.py:41-56,162,180-187 with the 4-level backbone
        feats = hrnet_backbone(sd, x.reshape(-1, c, h, w), taps)
        if "levels" in teacher:
            feats = list(teacher["levels"])
        feat = feats[0]
        hm = F.conv2d(feat, sd["pose_net.weight"], sd["pose_net.bias"], stride=2, padding=1)
        xy = soft_argmax_2d(hm)
        if "coords" in teacher:
            xy = teacher["coords"].reshape(xy.shape).to(xy.dtype)
        sampled = torch.cat([sample_net(sd, f, xy, None, level=l) for l, f in enumerate(feats)], dim=-1)
        if taps is not None:
            for l, f in enumerate(feats):
                taps[f"level{l}"] = f
    else:
        feat = backbone(sd, x.reshape(-1, c, h, w), taps)
        if "backbone_out" in teacher:                    # [b*v, 1024, 32, 32]
            feat = teacher["backbone_out"]
        hm = pose_net(sd, feat, taps)
        xy = soft_argmax_2d(hm)
        if "coords" in teacher:                          # [b*v, 21, 2] heat-map pixels: conditions everything downstream
            xy = teacher["coords"].reshape(xy.shape).to(xy.dtype)
        sampled = sample_net(sd, feat, xy, taps)
    tok = sampled
    if "pos2d" in pe_list:
        tok = torch.cat([tok, xy], dim=2)
    if "crop" in pe_list:
        fov = crop_fov(bbox, intr)
        tok = torch.cat([tok, fov.unsqueeze(1).expand(-1, NUM_JOINTS, -1)], dim=2)
    tok = tok.reshape(-1, nv * NUM_JOINTS, tok.shape[2])
    fused = fusion(sd, tok, cfg["model"].get("fusion_layers", 5), "sin" in pe_list, NUM_JOINTS, taps)
    joints = gcn_decoder(sd, fused, taps)
    scale = cfg["data"]["image_size"] / cfg["data"]["heatmap_size"]
    out = {
        "joints_crop_img": xy.reshape(-1, nv, NUM_JOINTS, 2) * scale,
        "joints_cam": joints,
        "heatmap": hm.reshape(-1, nv, NUM_JOINTS, hm.shape[2], hm.shape[3]),
    }
    if return_taps:
        taps.update(backbone_out=feat, heatmap=hm, coords=xy, sampled=sampled, tokens=tok,
                    fused=fused, joints_cam=joints)
        return out, taps
    return out


# ----------------------------------------------------------------------------
# image transform in front of the model (reference src/datasets/ho3d.py:35-40, 139-147;
# src/datasets/utils.py:40-77) - the caller side of the hot path
# ----------------------------------------------------------------------------
IMAGENET_MEAN = (0.485, 0.456, 0.406)       # ho3d.py:37-38
IMAGENET_STD = (0.229, 0.224, 0.225)


def crop_and_pad_image(image: np.ndarray, bbox) -> np.ndarray:
    """datasets/utils.py:40-77: crop [y1:y2, x1:x2] of an (h, w, 3) uint8 frame; parts of the box outside the frame
    are zero."""
    height, width = image.shape[:2]
    x1, y1, x2, y2 = (int(v) for v in bbox)
    sx, sy, ex, ey = max(0, x1), max(0, y1), min(width, x2), min(height, y2)
    out = np.zeros((y2 - y1, x2 - x1, 3), dtype=np.uint8)
    if ex > sx and ey > sy:
        px, py = max(0, -x1), max(0, -y1)
        out[py:py + (ey - sy), px:px + (ex - sx)] = image[sy:ey, sx:ex]
    return out


def image_transform(crop: np.ndarray, size: int = 256) -> torch.Tensor:
    """ho3d.py:35-40 `img_transform`: ToTensor (HWC uint8 -> CHW float / 255), Resize((size, size), antialias=True)
    (torchvision resizes tensors with F.interpolate(mode="bilinear", align_corners=False, antialias=True)), Normalize."""
    t = torch.from_numpy(np.ascontiguousarray(crop)).permute(2, 0, 1).to(torch.float32).div(255)
    t = F.interpolate(t[None], size=(size, size), mode="bilinear", align_corners=False, antialias=True)[0]
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    return (t - mean) / std


def preprocess(frames: np.ndarray, bboxes, size: int = 256) -> torch.Tensor:
    """frames (n, h, w, 3) uint8, bboxes (n, 4) integer xyxy -> (n, 3, size, size) fp32: ho3d.py:139-147 for the
    evaluation path (no augmentation; the all-joints-invisible black-image case is dataset logic, not arithmetic)."""
    return torch.stack([image_transform(crop_and_pad_image(f, b), size) for f, b in zip(frames, bboxes)])


def make_frames(n: int, seed: int = 0, height: int = 480, width: int = 640):
    """Seeded synthetic camera frames + integer boxes covering: inside the frame, sticking out on every side,
    smaller than the output (up-sampling), larger than it (anti-aliased down-sampling), non-square."""
    g = torch.Generator().manual_seed(seed)
    # smooth-ish content so that resampling errors are visible but values are not pure noise
    base = torch.rand(n, 3, height // 8, width // 8, generator=g)
    frames = F.interpolate(base, size=(height, width), mode="bilinear", align_corners=False)
    frames = (frames + 0.15 * torch.rand(n, 3, height, width, generator=g)).clamp(0, 1)
    frames = (frames * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().numpy()
    boxes = [(200, 120, 400, 320), (-40, -30, 180, 190), (500, 300, 700, 520), (60, 40, 580, 440),
             (300, 200, 390, 290), (100, 50, 420, 300), (0, 0, 640, 480), (610, 450, 660, 500)]
    bboxes = np.array([boxes[i % len(boxes)] for i in range(n)], dtype=np.int32)
    return frames, bboxes
