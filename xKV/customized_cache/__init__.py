"""Import shim: ``xKV.customized_cache`` resolves to the B200 implementation."""
import xkv_b200.customized_cache as _impl

FakeLayerMergingCache = _impl.FakeLayerMergingCache
method_to_cache_obj = _impl.method_to_cache_obj
