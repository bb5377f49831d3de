import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import handmvnet_oracle as O
from gpu_util import build_pair, rel_l2
b, views = 2, 5
m, ocfg, sd = build_pair(views, True, "fp32", micro_batch=b, seed=3, backbone="hrnet")
x, bbox, intr = O.make_inputs(b, views, seed=41)
ref, taps = O.forward(sd, ocfg, x, bbox, intr, return_taps=True)
m.stage_run("backbone", b, x=x.reshape(-1, 3, 256, 256).cuda())
m.stage_run("pose", b)
hm = m.tensor_get("heatmap", b).cpu()
print("hm finite", torch.isfinite(hm).all().item(), "err", rel_l2(hm, taps["heatmap"]), "max", hm.abs().max().item())
xy = m.tensor_get("xy", b).cpu()
print("xy finite", torch.isfinite(xy).all().item(), "n nan", torch.isnan(xy).sum().item(), "of", xy.numel())
bad = torch.isnan(xy).any(-1)
print("first bad maps", bad.nonzero()[:5].tolist())
if bad.any():
    n, j = bad.nonzero()[0].tolist()
    mp = hm[n, j]
    print("map stats", mp.max().item(), mp.min().item(), (mp * 1000).max().item(), torch.isinf(mp * 1000).any().item())
m.tensor_set("xy", taps["coords"].cuda(), b)
m.stage_run("sample", b, bbox=bbox.reshape(-1, 4).cuda(), intr=intr.reshape(-1, 4).cuda())
tok = m.tensor_get("tokens", b).cpu()
print("tokens finite", torch.isfinite(tok).all().item(), "err", rel_l2(tok, taps["tokens_pe"]))
d = (tok - taps["tokens_pe"]).abs().amax(dim=(0, 1))
print("worst cols", d.topk(5))
m.tensor_set("tokens", taps["tokens_pe"].cuda(), b)
m.stage_run("fusion", b)
f = m.tensor_get("fused", b).cpu()
print("fused finite", torch.isfinite(f).all().item(), "err", rel_l2(f, taps["fused"]))
m.tensor_set("fused", taps["fused"].cuda(), b)
m.stage_run("gcn", b)
j = m.tensor_get("joints", b).cpu()
print("joints finite", torch.isfinite(j).all().item(), "mm", float((j - taps["joints_cam"]).abs().max()) * 1e3)
m.synchronize()
