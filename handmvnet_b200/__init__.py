"""handmvnet_b200 - B200-native (sm_100a) implementation of HandMvNet's inference forward path.

Host side: Python/PyTorch for device memory, streams and torch.distributed plumbing.
Compute: hand-written CUDA (tcgen05/TMEM/TMA) behind the C ABI in include/handmvnet_b200.h,
loaded from handmvnet_b200/lib/libhandmvnet_b200.so.  There is no CPU or eager-PyTorch fallback.
"""
from .models.handmvnet import HandMvNet  # noqa: F401
from .config import load_config  # noqa: F401

__all__ = ["HandMvNet", "load_config"]
