"""Drop-in import shim: ``import xKV...`` resolves to the B200 implementation in ``xkv_b200`` under the
reference's module paths (xKV.patch, xKV.configurations, xKV.customized_cache, xKV.attn_patch)."""
