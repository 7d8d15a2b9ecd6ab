from xkv_b200.customized_cache import FakeLayerMergingCache, method_to_cache_obj  # noqa: F401
