"""Build experimental variants of the library (-D knobs of the seam kernel) next to the product library; selected at run
time with HMV_LIB_PATH (tools/gpu_r02_variants.sh).  usage: python tools/build_variants.py name:DEF=V,DEF=V ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from handmvnet_b200 import build as B

out_dir = os.path.join(os.path.dirname(B.LIB_PATH), "variants")
os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    name, defs = spec.split(":")
    path = os.path.join(out_dir, f"libhmv_{name}.so")
    B.build(force=True, defines=[d for d in defs.split(",") if d], out_path=path)
    print("built", path)
