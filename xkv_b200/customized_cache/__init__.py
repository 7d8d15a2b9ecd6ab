"""Cache classes of the xKV path (mirror of the reference's ``xKV/customized_cache/__init__.py:4-6``)."""
from .fake_layer_merge_dynamic_cache import FakeLayerMergingCache  # noqa: F401

method_to_cache_obj = {
    "xKV": FakeLayerMergingCache,
}
