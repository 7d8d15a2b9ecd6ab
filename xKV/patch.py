from xkv_b200.patch import KVCompress, prepare_cache  # noqa: F401
