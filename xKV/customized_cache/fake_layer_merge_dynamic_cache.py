from xkv_b200.customized_cache.fake_layer_merge_dynamic_cache import FakeLayerMergingCache  # noqa: F401
