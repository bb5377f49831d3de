"""In-tree nvcc build of libhandmvnet_b200.so for sm_100a (no JIT cache, the .so ships with the tree)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhandmvnet_b200.so")
SOURCES = ["conv_gemm_tc.cu", "bottleneck_tc.cu", "bottleneck_next_tc.cu", "stem_pool.cu", "conv_f32.cu", "head_kernels.cu", "fusion_block.cu", "model.cu"]
HEADERS = ["common.cuh", "conv_gemm_tc.cuh", "tc_ptx.cuh", "kernels.cuh", os.path.join("..", "..", "include", "handmvnet_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out_path: str = None) -> str:
    """defines / out_path: experimental variants (tools/gpu_r02_variants.sh) built next to the product library."""
    if out_path is None and not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out_path or LIB_PATH] + \
          [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout)
    if verbose:
        print(r.stdout)
    return out_path or LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
