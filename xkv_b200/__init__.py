"""xkv_b200 — B200-native implementation of the xKV cross-layer SVD KV-cache hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); the arithmetic runs in the
hand-written sm_100a kernels of ``libxkv_b200.so`` (C ABI: ``include/xkv_b200.h``).  There is no CPU
fallback: ops raise :class:`xkv_b200._lib.XkvError` if the library is missing.
"""
from ._lib import XkvError  # noqa: F401
from .configurations import (  # noqa: F401
    LayerGroup,
    generate_consecutive_layer_groups,
    generate_consecutive_xKV_config,
    xKVConfig,
)

__all__ = ["XkvError", "LayerGroup", "xKVConfig", "generate_consecutive_layer_groups",
           "generate_consecutive_xKV_config"]
