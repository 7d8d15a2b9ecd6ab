"""Cache classes of the xKV path, keyed by method name (the registry ``prepare_cache`` looks classes up in; the
reference keeps the same mapping in ``xKV/customized_cache/__init__.py:4-6``)."""
from . import fake_layer_merge_dynamic_cache as _fake

FakeLayerMergingCache = _fake.FakeLayerMergingCache

method_to_cache_obj = dict(xKV=FakeLayerMergingCache)

__all__ = ["FakeLayerMergingCache", "method_to_cache_obj"]
