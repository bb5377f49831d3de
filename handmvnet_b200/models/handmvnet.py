"""Drop-in replacement for the reference `models.handmvnet.HandMvNet` inference interface
(reference src/models/handmvnet.py:27-266): same constructor arguments, attribute / state_dict
names, `forward(x, bbox=None, cam_params=None)` signature and returned dict - but the forward is
executed by the sm_100a CUDA library behind include/handmvnet_b200.h.

The torch modules below are PARAMETER HOLDERS ONLY (they give the 355-key reference state_dict its
names and shapes so `load_state_dict(strict=True)` of a reference checkpoint works, src/eval.py:27-52);
they are never executed: there is no eager-PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.nn as nn

from .. import _lib

NUM_JOINTS = 21


class _NoEagerPath(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("handmvnet_b200 parameter holder: the forward path runs in the CUDA library only")


class _Bottleneck(_NoEagerPath):
    """Parameters of reference backbones/resnet.py:109-122."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride


class _ResNet50Paper(_NoEagerPath):
    """Parameters of reference ResNet50_Paper (backbones/resnet.py:147-203, 348-357): stem + layers
    [3, 4, 6] with layer3 at stride 1; kaiming_normal(fan_out) convs, unit BatchNorm."""

    def __init__(self):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = self._make_layer(64, 3, 1)
        self.layer2 = self._make_layer(128, 4, 2)
        self.layer3 = self._make_layer(256, 6, 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * 4:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, kernel_size=1, stride=stride, bias=False),
                                       nn.BatchNorm2d(planes * 4))
        layers = [_Bottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * 4
        layers += [_Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)


class _BasicBlock(_NoEagerPath):
    """Parameters of reference backbones/hrnet.py:26-36."""

    def __init__(self, planes):
        super().__init__()
        self.conv1 = nn.Conv2d(planes, planes, kernel_size=3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)


def _conv_bn(cin, cout, k, stride=1, relu=False):
    layers = [nn.Conv2d(cin, cout, k, stride, k // 2, bias=False), nn.BatchNorm2d(cout)]
    if relu:
        layers.append(nn.ReLU(False))
    return nn.Sequential(*layers)


class _HRModule(_NoEagerPath):
    """Parameters of reference hrnet.py:97-211 (HighResolutionModule: 4 BasicBlocks per branch + fuse layers)."""

    def __init__(self, channels):
        super().__init__()
        n = len(channels)
        self.branches = nn.ModuleList([nn.Sequential(*[_BasicBlock(c) for _ in range(4)]) for c in channels])
        fuse = []
        for i in range(n):
            row = []
            for j in range(n):
                if j > i:
                    row.append(nn.Sequential(nn.Conv2d(channels[j], channels[i], 1, 1, 0, bias=False), nn.BatchNorm2d(channels[i]),
                                             nn.Upsample(scale_factor=2 ** (j - i), mode="nearest")))
                elif j == i:
                    row.append(None)
                else:
                    row.append(nn.Sequential(*[_conv_bn(channels[j], channels[i] if k == i - j - 1 else channels[j], 3, 2, relu=k < i - j - 1)
                                               for k in range(i - j)]))
            fuse.append(nn.ModuleList(row))
        self.fuse_layers = nn.ModuleList(fuse)


HR_CHANNELS = {"w40": (40, 80, 160, 320), "w64": (64, 128, 256, 512)}


class _HRNet(_NoEagerPath):
    """Parameters of reference HRNet (backbones/hrnet.py:241-276, 427-495): 3x3/2 stem x2, 4 Bottlenecks, stages 2-4 with
    1 / 4 / 3 modules of 2 / 3 / 4 branches; kaiming_normal(fan_out) convs, unit BatchNorm (hrnet.py:412-418)."""

    def __init__(self, hrnet_type="w40"):
        super().__init__()
        if hrnet_type not in HR_CHANNELS:
            raise Exception("HRNet only supports ['w64', 'w40'] as model_type, found: " + hrnet_type)
        ch = HR_CHANNELS[hrnet_type]
        self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.conv2 = nn.Conv2d(64, 64, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(64)
        ds = nn.Sequential(nn.Conv2d(64, 256, kernel_size=1, bias=False), nn.BatchNorm2d(256))
        self.layer1 = nn.Sequential(_Bottleneck(64, 64, 1, ds), *[_Bottleneck(256, 64) for _ in range(3)])
        pre = [256]
        for si, (nmod, nbr) in enumerate(((1, 2), (4, 3), (3, 4)), start=2):
            cur = list(ch[:nbr])
            trans = []
            for i in range(nbr):
                if i < len(pre):
                    trans.append(_conv_bn(pre[i], cur[i], 3, 1, relu=True) if cur[i] != pre[i] else None)
                else:
                    trans.append(nn.Sequential(*[_conv_bn(pre[-1], cur[i] if j == i - len(pre) else pre[-1], 3, 2, relu=True)
                                                 for j in range(i + 1 - len(pre))]))
            setattr(self, f"transition{si - 1}", nn.ModuleList(trans))
            setattr(self, f"stage{si}", nn.Sequential(*[_HRModule(cur) for _ in range(nmod)]))
            pre = cur
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


class _SampleNet(_NoEagerPath):
    """Parameters of reference nets.py:24-31 (SampleNet([c, c//2]))."""

    def __init__(self, c_in, c_out):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(c_in, c_out, kernel_size=1), nn.BatchNorm2d(c_out), nn.ReLU(inplace=True))


class _FeedForward(_NoEagerPath):
    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(0.0),
                                 nn.Linear(hidden, dim), nn.Dropout(0.0))


class _MultiHeadAttention(_NoEagerPath):
    """Parameters of reference layers.py:178-200 (8 heads x 128, FF hidden 128)."""

    def __init__(self, d_model, n_heads=8, dim_head=128, custom_query_length=0):
        super().__init__()
        inner = n_heads * dim_head
        self.custom_query_length = custom_query_length
        self.heads = n_heads
        self.to_q = nn.Linear(d_model, inner, bias=False)
        self.to_k = nn.Linear(d_model, inner, bias=False)
        self.to_v = nn.Linear(d_model, inner, bias=False)
        self.to_out = nn.Linear(inner, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.ff = _FeedForward(d_model, dim_head)


class _PositionalEncoding(_NoEagerPath):
    """Constant table of reference layers.py:136-150 (a plain attribute there too: not in the state_dict)."""

    def __init__(self, d_model, max_len):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        self.pe = torch.zeros(1, max_len, d_model)
        self.pe[0, :, 0::2] = torch.sin(position * div_term)
        self.pe[0, :, 1::2] = torch.cos(position * div_term) if d_model % 2 == 0 else torch.cos(position * div_term[:-1])


class _CrossAttentionFusion(_NoEagerPath):
    """Parameters of reference fusion.py:7-24."""

    def __init__(self, feat_dim, max_tokens, custom_query_length, num_layers):
        super().__init__()
        assert num_layers % 2 == 1, "num_layers must be an odd number"
        half = (num_layers - 1) // 2
        self.feat_dim = feat_dim
        self.pos_encoding = _PositionalEncoding(feat_dim, max_tokens)
        layers = [_MultiHeadAttention(feat_dim) for _ in range(half)]
        layers.append(_MultiHeadAttention(feat_dim, custom_query_length=custom_query_length))
        layers += [_MultiHeadAttention(feat_dim) for _ in range(half)]
        self.attn_fusion = nn.Sequential(*layers)


class _ChebConv(_NoEagerPath):
    """Parameters of reference layers.py:370-383 (K=2 -> 3 Chebyshev terms)."""

    def __init__(self, in_c, out_c, K=2):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(K + 1, 1, in_c, out_c))
        nn.init.xavier_normal_(self.weight)
        self.bias = nn.Parameter(torch.zeros(1, 1, out_c))


class _JointsDecoderGCN(_NoEagerPath):
    def __init__(self, in_features, out_dim=3):
        super().__init__()
        self.joints_gcn1 = _ChebConv(in_features, 256)
        self.joints_gcn2 = _ChebConv(256, 64)
        self.joints_gcn3 = _ChebConv(64, out_dim)


class HostTicket:
    """One in-flight forward_host_async call.  result() blocks until it has completed and returns the output dict;
    with recycle=True the pinned output buffers go back to the model's pool (they are overwritten by a later call), so
    clone what must outlive the next calls."""

    def __init__(self, model, ticket, inputs, outputs, pool):
        self._model, self._ticket, self._inputs, self._outputs, self._pool = model, ticket, inputs, outputs, pool
        self._done = False

    def result(self, recycle=False):
        if not self._done:
            _lib.check(_lib.load().hmv_host_wait(self._model._handle, self._ticket), "hmv_host_wait")
            self._done = True
            self._inputs = None
        hm, j2d, j3d = self._outputs
        out = {"joints_crop_img": j2d, "joints_cam": j3d}
        if hm is not None:
            out["heatmap"] = hm
        if recycle and self._pool is not None:
            self._pool.append(self._outputs)
            self._pool = None
        return out

    def __del__(self):
        # a ticket dropped without result(): its pinned buffers must outlive the copies still in flight on the
        # library's private streams (torch's pinned allocator does not know about those streams)
        try:
            if not self._done and self._model._handle is not None:
                _lib.load().hmv_host_wait(self._model._handle, self._ticket)
        except Exception:  # interpreter shutdown
            pass


class HandMvNet(nn.Module):
    """`HandMvNet(train_params, model_params, data_params)` - reference handmvnet.py:28.

    Extra keyword arguments (not in the reference): `precision` ("bf16" tensor-core path | "fp32"
    check mode) and `micro_batch` (samples per internal pass; the device workspace is sized for it).
    """

    def __init__(self, train_params, model_params, data_params, precision: str = "bf16", micro_batch: int = 64):
        super().__init__()
        self.train_params = train_params
        self.model_params = model_params
        self.data_params = data_params
        self.debug = train_params.get("debug", False)
        self.num_views = model_params["num_views"]
        self.batch_size = data_params.get("batch_size", 1)

        self.backbone_name = model_params.get("backbone", "hrnet")
        assert self.backbone_name in ["hrnet", "resnet"], "Backbone should be one of ['hrnet', 'resnet']"
        if data_params.get("image_size", 256) != 256 or data_params.get("heatmap_size", 32) != 32:
            raise NotImplementedError("only image_size 256 / heatmap_size 32 (all release configs) are built")
        if self.backbone_name == "hrnet":
            # the `*_HR*` release configs (reference handmvnet.py:41-56): HRNet-w40 / w64, four feature levels
            self.backbone_type = model_params.get("backbone_type", "w40")
            self.backbone_channels = model_params["backbone_channels"]
            self.backbone = _HRNet(self.backbone_type)
            if tuple(self.backbone_channels) != HR_CHANNELS[self.backbone_type]:
                raise ValueError(f"backbone_channels {list(self.backbone_channels)} do not match HRNet-{self.backbone_type}")
            self.pose_net = nn.Conv2d(self.backbone_channels[0], NUM_JOINTS, kernel_size=3, stride=2, padding=1)
        else:
            self.backbone_type = model_params.get("backbone_type", "34")
            assert self.backbone_type in ["18", "34", "50_paper"], "Supports only 18, 34, 50_paper"
            if self.backbone_type != "50_paper":
                raise NotImplementedError("only backbone_type '50_paper' (the release ResNet configs) is built")
            self.backbone_channels = model_params["backbone_channels"]
            if list(self.backbone_channels) != [1024]:
                raise NotImplementedError("backbone_channels must be [1024] for the 50_paper backbone")
            # weights pretrained on ImageNet cannot be fetched offline; a checkpoint is loaded with load_state_dict
            self.backbone = _ResNet50Paper()
            self.pose_net = nn.Sequential(nn.Conv2d(1024, 512, kernel_size=1), nn.BatchNorm2d(512), nn.ReLU(inplace=True),
                                          nn.Conv2d(512, NUM_JOINTS, kernel_size=1))

        self.feat_dim = int(sum(self.backbone_channels) / 2)
        self.pos_enc_list = model_params.get("pos_enc", ["pos2d", "sin"])
        self.sinusoidal_pos = "sin" in self.pos_enc_list
        if "pos2d" in self.pos_enc_list:
            self.feat_dim += 2
        if "crop" in self.pos_enc_list:
            self.feat_dim += 10
        self.sample_nets = nn.ModuleList([_SampleNet(c, c // 2) for c in self.backbone_channels])

        self.fusion_layers = model_params.get("fusion_layers", 5)
        if model_params["fusion"] == "cross_attn":
            self.joints_late_fusion = _CrossAttentionFusion(self.feat_dim, NUM_JOINTS * self.num_views, NUM_JOINTS,
                                                            self.fusion_layers)
        elif model_params["fusion"] == "cross_attn_learnable_query":
            raise NotImplementedError("fusion 'cross_attn_learnable_query' is not selected by any release config")
        else:
            raise NotImplementedError(f"Invalid fusion type: {model_params['fusion']}")
        if not model_params["use_gcn"]:
            raise NotImplementedError("use_gcn: false (JointsDecoderNN) is outside the B200 hot path")
        self.joints_decoder = _JointsDecoderGCN(self.feat_dim)
        if not train_params.get("root_relative", True):
            raise NotImplementedError("root_relative: false - the reference forward itself raises at "
                                      "handmvnet.py:238 (SURVEY.md §8f rank 4); there is no behaviour to match")
        if model_params.get("get_vertices", False):
            raise NotImplementedError("MANO mesh post-processing (get_vertices) is out of scope")
        ds_name = data_params.get("name", "dexycb")
        if ds_name == "ho3d":
            self.auc_thresh = [0.0, 0.05]
        elif ds_name in ("dexycb", "mvhand"):
            self.auc_thresh = [0.0, 0.02]
        else:
            raise NotImplementedError(f"Dataset not found: {ds_name}")

        if precision not in _lib.PRECISION:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISION)}")
        self.precision = precision
        self.micro_batch = int(micro_batch)
        self._handle = None
        self._handle_device = None
        self._input_norm = None              # (mean, std) for uint8 inputs; None = the library's ImageNet defaults

    # ---- Lightning API used by the reference drivers (eval_fps.py:64-65, eval.py:86-87) ----------
    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    # ---- handle management -------------------------------------------------------------------------
    def _release(self):
        if self._handle is not None:
            handle, self._handle, self._handle_device = self._handle, None, None
            _lib.check(_lib.load().hmv_destroy(handle), "hmv_destroy")

    def __del__(self):
        try:
            self._release()
        except RuntimeError as e:            # a failing hmv_destroy must not be silent
            import warnings
            warnings.warn(f"handmvnet_b200: {e}")
        except Exception:                    # interpreter shutdown: modules may already be gone
            pass

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._release()                      # repack on next use
        return out

    def _apply(self, fn, *a, **k):
        self._release()                      # .to(device) / .cuda(): weights move, rebuild the device plan lazily
        return super()._apply(fn, *a, **k)

    def prepare(self, device=None):
        """Create the device handle: fold BN, repack weights, build the TMA descriptors (hmv_prepare)."""
        lib = _lib.load()
        if device is None:
            device = next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("handmvnet_b200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self.training:
            raise RuntimeError("training-mode forward (BatchNorm batch statistics, dropout) is not part of the "
                               "inference path: call .eval() / .freeze() first")
        self._release()
        cfg = _lib.HmvConfig(num_views=self.num_views, image_size=self.data_params.get("image_size", 256),
                             heatmap_size=self.data_params.get("heatmap_size", 32),
                             use_pos2d=int("pos2d" in self.pos_enc_list), use_crop=int("crop" in self.pos_enc_list),
                             use_sin=int(self.sinusoidal_pos), fusion_layers=self.fusion_layers,
                             precision=_lib.PRECISION[self.precision], micro_batch=self.micro_batch, device=index,
                             backbone=_lib.BACKBONE[self.backbone_name])
        if self.backbone_name == "hrnet":
            cfg.hr_channels = (ctypes.c_int32 * 4)(*self.backbone_channels)
        handle = ctypes.c_void_p()
        with torch.cuda.device(index):
            _lib.check(lib.hmv_create(ctypes.byref(cfg), ctypes.byref(handle)), "hmv_create")
            try:
                tensors = dict(self.state_dict())
                tensors["pe"] = self.joints_late_fusion.pos_encoding.pe[0]
                for name, t in tensors.items():
                    if name.endswith("num_batches_tracked"):
                        continue
                    host = t.detach().to("cpu", torch.float32).contiguous()
                    dims = (ctypes.c_int64 * max(host.dim(), 1))(*host.shape)
                    _lib.check(lib.hmv_set_weight(handle, name.encode(), _lib.ptr(host), dims, host.dim()),
                               f"hmv_set_weight({name})")
                _lib.check(lib.hmv_prepare(handle), "hmv_prepare")
            except Exception:
                lib.hmv_destroy(handle)
                raise
        self._handle = handle
        self._handle_device = torch.device("cuda", index)
        if self._input_norm is not None:     # survives .to() / load_state_dict(), which rebuild the handle
            self._apply_input_norm()
        return self

    def _ensure(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("handmvnet_b200 forward needs CUDA tensors (no CPU fallback); use forward_host() "
                               "for host buffers")
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self.training:                    # .train() after prepare(): BatchNorm batch statistics / dropout are not built
            raise RuntimeError("handmvnet_b200 implements the inference path only: call .eval() / .freeze() first")
        if self._handle is None or self._handle_device.index != index:
            self.prepare(torch.device("cuda", index))
        return self._handle

    @staticmethod
    def _f32(t, device):
        return None if t is None else t.to(device=device, dtype=torch.float32).contiguous()

    def _check_inputs(self, x, bbox, cam_params):
        if x.dim() != 5:
            raise ValueError(f"x must be [b, v, 3, H, W], got {tuple(x.shape)}")
        b, v, c, hh, ww = x.shape
        if v != self.num_views:   # the reference fails late or silently regroups views (handmvnet.py:194-196,225)
            raise ValueError(f"input has {v} views but the model was built for num_views={self.num_views}")
        size = self.data_params.get("image_size", 256)
        if c != 3 or hh != size or ww != size:
            raise ValueError(f"x must be [b, v, 3, {size}, {size}], got {tuple(x.shape)}")
        intr = None
        if "crop" in self.pos_enc_list:
            if bbox is None or cam_params is None or "intrinsic" not in cam_params:
                raise ValueError("pos_enc contains 'crop': forward needs bbox and cam_params['intrinsic'] "
                                 "(reference handmvnet.py:205-216)")
            intr = cam_params["intrinsic"]
            if bbox.numel() != b * v * 4 or intr.numel() != b * v * 4:
                raise ValueError("bbox and cam_params['intrinsic'] must be [b, v, 4]")
        return b, v, intr

    # ---- the hot path ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, bbox=None, cam_params=None):
        """x [b, v, 3, 256, 256] (CUDA fp32) -> {"joints_crop_img" [b,v,21,2], "joints_cam" [b,21,3],
        "heatmap" [b,v,21,32,32]} (reference handmvnet.py:158-266).  Asynchronous on the current stream."""
        b, v, intr = self._check_inputs(x, bbox, cam_params)
        handle = self._ensure(x.device)
        dev = x.device
        u8 = x.dtype == torch.uint8                   # raw image bytes: ToTensor + Normalize run inside the stem kernel
        x = x.contiguous() if u8 else self._f32(x, dev)
        crop = "crop" in self.pos_enc_list
        bbox_f = self._f32(bbox, dev) if crop else None
        intr_f = self._f32(intr, dev) if crop else None
        hm = torch.empty((b, v, NUM_JOINTS, 32, 32), device=dev, dtype=torch.float32)
        j2d = torch.empty((b, v, NUM_JOINTS, 2), device=dev, dtype=torch.float32)
        j3d = torch.empty((b, NUM_JOINTS, 3), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            fn = _lib.load().hmv_forward_u8 if u8 else _lib.load().hmv_forward
            _lib.check(fn(handle, _lib.ptr(x), _lib.ptr(bbox_f), _lib.ptr(intr_f), b, _lib.ptr(hm),
                          _lib.ptr(j2d), _lib.ptr(j3d), stream), "hmv_forward")
        return {"joints_crop_img": j2d, "joints_cam": j3d, "heatmap": hm}

    @torch.no_grad()
    def forward_host(self, x, bbox=None, cam_params=None, device=None, want_heatmap=True):
        """Same call on HOST tensors (pinned recommended): host->device copies are pipelined against
        compute inside the library (hmv_forward_host); returns CPU tensors after completion."""
        if x.device.type != "cpu":
            raise ValueError("forward_host expects CPU tensors")
        if x.dtype == torch.uint8:
            return self.forward_host_async(x, bbox, cam_params, device, want_heatmap).result()
        b, v, intr = self._check_inputs(x, bbox, cam_params)
        if self._handle is None:
            self.prepare(device if device is not None else next(self.parameters()).device)
        crop = "crop" in self.pos_enc_list
        x = x.to(torch.float32).contiguous()
        bbox_f = bbox.to(torch.float32).contiguous() if crop else None
        intr_f = intr.to(torch.float32).contiguous() if crop else None
        pin = torch.cuda.is_available()
        hm = torch.empty((b, v, NUM_JOINTS, 32, 32), dtype=torch.float32, pin_memory=pin) if want_heatmap else None
        j2d = torch.empty((b, v, NUM_JOINTS, 2), dtype=torch.float32, pin_memory=pin)
        j3d = torch.empty((b, NUM_JOINTS, 3), dtype=torch.float32, pin_memory=pin)
        _lib.check(_lib.load().hmv_forward_host(self._handle, _lib.ptr(x), _lib.ptr(bbox_f), _lib.ptr(intr_f), b,
                                                _lib.ptr(hm), _lib.ptr(j2d), _lib.ptr(j3d)), "hmv_forward_host")
        out = {"joints_crop_img": j2d, "joints_cam": j3d}
        if want_heatmap:
            out["heatmap"] = hm
        return out

    @torch.no_grad()
    def forward_host_async(self, x, bbox=None, cam_params=None, device=None, want_heatmap=True):
        """Streaming form of forward_host: enqueues one batch and returns a HostTicket at once; the copies of the next
        call overlap this call's compute (hmv_forward_host_async).  `ticket.result()` waits and returns the CPU
        tensors.  The input tensors must be pinned or at least stay alive and unmodified until then."""
        if x.device.type != "cpu":
            raise ValueError("forward_host_async expects CPU tensors")
        b, v, intr = self._check_inputs(x, bbox, cam_params)
        if self._handle is None:
            self.prepare(device if device is not None else next(self.parameters()).device)
        crop = "crop" in self.pos_enc_list
        u8 = x.dtype == torch.uint8
        x = x.contiguous() if u8 else x.to(torch.float32).contiguous()
        bbox_f = bbox.to(torch.float32).contiguous() if crop else None
        intr_f = intr.to(torch.float32).contiguous() if crop else None
        pool = self.__dict__.setdefault("_host_out_pool", {})
        key = (b, v, bool(want_heatmap))
        free = pool.setdefault(key, [])
        if free:
            hm, j2d, j3d = free.pop()
        else:
            pin = torch.cuda.is_available()
            hm = torch.empty((b, v, NUM_JOINTS, 32, 32), dtype=torch.float32, pin_memory=pin) if want_heatmap else None
            j2d = torch.empty((b, v, NUM_JOINTS, 2), dtype=torch.float32, pin_memory=pin)
            j3d = torch.empty((b, NUM_JOINTS, 3), dtype=torch.float32, pin_memory=pin)
        t = ctypes.c_int64(-1)
        fn = _lib.load().hmv_forward_host_u8_async if u8 else _lib.load().hmv_forward_host_async
        _lib.check(fn(self._handle, _lib.ptr(x), _lib.ptr(bbox_f), _lib.ptr(intr_f), b,
                      _lib.ptr(hm), _lib.ptr(j2d), _lib.ptr(j3d), ctypes.byref(t)), "hmv_forward_host_async")
        return HostTicket(self, t.value, (x, bbox_f, intr_f), (hm, j2d, j3d), free)

    @torch.no_grad()
    def preprocess(self, frames, bbox):
        """Camera frames [..., H, W, 3] uint8 (CUDA) + integer boxes [..., 4] xyxy -> normalised crops [..., 3, S, S] fp32:
        the reference dataset's crop_and_pad_image + ToTensor + Resize(antialias) + Normalize (datasets/utils.py:40-77,
        datasets/ho3d.py:35-40) as one kernel.  Feed the result to forward()."""
        if frames.dtype != torch.uint8 or frames.dim() < 4 or frames.shape[-1] != 3:
            raise ValueError("frames must be uint8 [..., H, W, 3]")
        if frames.device.type != "cuda":
            raise RuntimeError("handmvnet_b200 has no CPU fallback: preprocess expects CUDA tensors")
        lead = tuple(frames.shape[:-3])
        hf, wf = int(frames.shape[-3]), int(frames.shape[-2])
        n = 1
        for d in lead:
            n *= int(d)
        if bbox.numel() != n * 4:
            raise ValueError("bbox must hold one xyxy box per frame")
        self._ensure(frames.device)
        dev = frames.device
        size = self.data_params.get("image_size", 256)
        fr = frames.contiguous()
        bb = bbox.to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty(lead + (3, size, size), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.load().hmv_preprocess(self._handle, _lib.ptr(fr), _lib.ptr(bb), n, hf, wf, _lib.ptr(out), stream),
                       "hmv_preprocess")
        return out

    def set_input_norm(self, mean, std):
        """Per-channel mean / std applied to uint8 inputs (defaults: the reference's ImageNet constants, datasets/ho3d.py:35-40)."""
        mean, std = [float(v) for v in mean], [float(v) for v in std]
        if len(mean) != 3 or len(std) != 3 or min(std) <= 0:
            raise ValueError("mean / std must have 3 entries, std > 0")
        self._input_norm = (mean, std)       # kept on the module: re-applied whenever the handle is rebuilt
        if self._handle is not None:
            self._apply_input_norm()

    def _apply_input_norm(self):
        mean, std = self._input_norm
        m = (ctypes.c_float * 3)(*mean)
        sd = (ctypes.c_float * 3)(*std)
        _lib.check(_lib.load().hmv_set_input_norm(self._handle, m, sd), "hmv_set_input_norm")

    def synchronize(self):
        if self._handle is not None:
            _lib.check(_lib.load().hmv_synchronize(self._handle), "hmv_synchronize")

    # ---- per-stage access (teacher-forced parity tests, micro-benchmarks) -----------------------------------
    def stage_run(self, stage, batch, x=None, bbox=None, intr=None):
        dev = self._handle_device
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.load().hmv_stage_run(self._handle, _lib.STAGE[stage], _lib.ptr(self._f32(x, dev)),
                                                 _lib.ptr(self._f32(bbox, dev)), _lib.ptr(self._f32(intr, dev)),
                                                 batch, stream), f"hmv_stage_run({stage})")

    def _tensor_shape(self, name, batch):
        v, d = self.num_views, self.feat_dim
        if self.backbone_name == "hrnet":
            c = self.backbone_channels
            lv = {"feat": (batch * v, c[0], 64, 64), "feat1": (batch * v, c[1], 32, 32), "feat2": (batch * v, c[2], 16, 16),
                  "feat3": (batch * v, c[3], 8, 8)}
            if name in lv:
                return lv[name]
        return {"feat": (batch * v, 1024, 32, 32), "heatmap": (batch * v, NUM_JOINTS, 32, 32),
                "xy": (batch * v, NUM_JOINTS, 2), "tokens": (batch, NUM_JOINTS * v, d),
                "fused": (batch, NUM_JOINTS, d), "joints": (batch, NUM_JOINTS, 3)}[name]

    def tensor_get(self, name, batch):
        dev = self._handle_device
        out = torch.empty(self._tensor_shape(name, batch), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.load().hmv_tensor_get(self._handle, _lib.TENSOR[name], _lib.ptr(out), batch, stream),
                       f"hmv_tensor_get({name})")
        return out

    def tensor_set(self, name, value, batch):
        dev = self._handle_device
        value = self._f32(value, dev)
        if tuple(value.shape) != self._tensor_shape(name, batch):
            raise ValueError(f"{name}: expected shape {self._tensor_shape(name, batch)}, got {tuple(value.shape)}")
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.load().hmv_tensor_set(self._handle, _lib.TENSOR[name], _lib.ptr(value), batch, stream),
                       f"hmv_tensor_set({name})")

    def debug_backbone_steps(self):
        lib = _lib.load()
        return [lib.hmv_debug_step_name(self._handle, i).decode() for i in range(lib.hmv_debug_num_steps(self._handle))]

    def debug_backbone(self, x_img, num_steps):
        """Run the first num_steps backbone steps on x_img [n_img,3,256,256]; returns the last output as NCHW fp32."""
        dev = self._handle_device
        x_img = self._f32(x_img, dev)
        n_img = x_img.shape[0]
        buf = torch.empty(n_img * 1024 * 1024 + 16, device=dev, dtype=torch.float32)   # >= any step output
        chw = (ctypes.c_int32 * 3)()
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.load().hmv_debug_backbone(self._handle, _lib.ptr(x_img), n_img, num_steps, _lib.ptr(buf), chw,
                                                      stream), "hmv_debug_backbone")
        c, hh, ww = chw[0], chw[1], chw[2]
        return buf[: n_img * c * hh * ww].view(n_img, c, hh, ww)

    def debug_step_io(self, step):
        """([(tap, C, H, W)] inputs, [(tap, C, H, W)] outputs) of backbone plan step `step`; the taps are the keys of
        oracle.backbone(per_layer=True).  An empty input list means the step reads the network input."""
        buf = ctypes.create_string_buffer(1024)
        _lib.check(_lib.load().hmv_debug_step_io(self._handle, step, buf, 1024), "hmv_debug_step_io")
        ins, outs = [], []
        for item in buf.value.decode().split(";"):
            if not item:
                continue
            kind, tap, c, hh, ww = item.split(" ")
            (ins if kind == "in" else outs).append((tap, int(c), int(hh), int(ww)))
        return ins, outs

    def debug_step_run(self, step, inputs):
        """Run backbone plan step `step` alone on teacher-forced inputs (NCHW fp32 tensors in debug_step_io order);
        returns its outputs as NCHW fp32 tensors."""
        dev = self._handle_device
        ins, outs = self.debug_step_io(step)
        if len(inputs) != len(ins):
            raise ValueError(f"step {step} takes {len(ins)} inputs")
        inputs = [self._f32(t, dev) for t in inputs]
        n_img = inputs[0].shape[0]
        for t, (tap, c, hh, ww) in zip(inputs, ins):
            if tuple(t.shape) != (n_img, c, hh, ww):
                raise ValueError(f"{tap}: expected {(n_img, c, hh, ww)}, got {tuple(t.shape)}")
        results = [torch.empty((n_img, c, hh, ww), device=dev, dtype=torch.float32) for _, c, hh, ww in outs]
        pin = (ctypes.c_void_p * len(inputs))(*[t.data_ptr() for t in inputs])
        pout = (ctypes.c_void_p * len(results))(*[t.data_ptr() for t in results])
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.load().hmv_debug_step_run(self._handle, step, n_img, pin, pout, stream), "hmv_debug_step_run")
        return results

    def profile(self, enable: bool):
        _lib.check(_lib.load().hmv_profile_enable(self._handle, int(enable)), "hmv_profile_enable")

    def profile_read(self, csv_path=None):
        """(device ms, algorithmic FLOPs, launches) of the tensor-core GEMM kernel since the last read."""
        ms, fl, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int64(0)
        _lib.check(_lib.load().hmv_profile_read(self._handle, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n),
                                                csv_path.encode() if csv_path else None), "hmv_profile_read")
        return ms.value, fl.value, n.value

    def profile_phases(self):
        """Stream ms (gaps included) of the profiled passes: backbone, heads, fusion, graph head."""
        out = (ctypes.c_double * 4)()
        _lib.check(_lib.load().hmv_profile_phases(self._handle, out), "hmv_profile_phases")
        return {"backbone": out[0], "heads": out[1], "fusion": out[2], "gcn": out[3]}

    def launch_count(self):
        return int(_lib.load().hmv_launch_count(self._handle)) if self._handle is not None else 0
