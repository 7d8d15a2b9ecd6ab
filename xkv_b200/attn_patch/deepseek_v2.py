"""DeepSeek-V2 (MLA) entry point (reference: xKV/attn_patch/deepseek_v2.py:160-302).

The reference patches the Hub's remote-code ``DeepseekV2FlashAttention2`` so that the cache holds the
latent ``compressed_kv`` (b, 1, l, kv_lora_rank) in the key slot and the RoPE'd ``k_pe`` (b, 1, l, 64) in the
value slot, with ``re_apply_rope=False`` and ``merge_value`` forbidden (:217-232).  That remote code is not
vendored and there is no network here; transformers' native ``deepseek_v2`` caches expanded K/V instead of
latents.  The cache side of this path is implemented and tested (``FakeLayerMergingCache.update`` with
``re_apply_rope=False``, one head of 512, value slot left dense: tests/test_cache_gpu.py); the latent-caching
attention forward itself is the next row and raises until it lands."""
from __future__ import annotations


def enable_deepseek_v2_xKV_eval(model):  # noqa: N802
    raise NotImplementedError(
        "DeepSeek-V2 MLA attention patch is not built yet on the B200 path (SURVEY.md §8 f2): the cache already "
        "supports the latent slot (re_apply_rope=False, merge_value=False), the latent-caching forward is next.")
