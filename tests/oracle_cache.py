"""Test infrastructure: the reference's FakeLayerMergingCache semantics (dense fake_svd cache) restated on
the installed transformers API with the oracle's functions, so model-level tests can compare the B200
cache against it.  Mirrors fake_layer_merge_dynamic_cache.py:127-213."""
from __future__ import annotations

import torch
from transformers.cache_utils import DynamicLayer

from oracle import xkv_oracle as O
from xkv_b200.customized_cache.fake_layer_merge_dynamic_cache import FakeLayerMergingCache


class OracleCache(FakeLayerMergingCache):
    """Dense cache: SVD -> truncate -> multiply back (oracle.merge_group), RoPE after reconstruction."""

    def __init__(self, merge_setup):
        super().__init__(merge_setup)
        self.layer_class_to_replicate = DynamicLayer

    def _layer(self, layer_idx):
        while len(self.layers) <= layer_idx:
            self.layers.append(DynamicLayer())
        return self.layers[layer_idx]

    def attend(self, *args, **kwargs):
        return None

    def latent_slot(self, *args, **kwargs):
        return None

    def update(self, key, value, layer_idx, mode="prefill", cos=None, sin=None, re_apply_rope=True,
               return_dense=True):
        layer = self._layer(layer_idx)
        layer.update(key, value)                                          # cache:129
        if mode == "prefill":
            info = self.merge_setup.get_group_for_layer(layer_idx)
            if info is not None:
                if layer_idx == info.layers[-1]:                          # cache:137-148
                    ids = list(range(info.layers[0], info.layers[-1] + 1))
                    keys = [self.layers[i].keys for i in ids]
                    vals = [self.layers[i].values for i in ids]
                    k_hat, v_hat = O.merge_group(keys, vals, info.rank_k, info.rank_v, self.merge_setup.merge_key,
                                                 self.merge_setup.merge_value)
                    for i, k, v in zip(ids, k_hat, v_hat):
                        self.layers[i].keys = O.apply_rope(k, cos, sin) if re_apply_rope else k
                        self.layers[i].values = v
            elif re_apply_rope:                                            # cache:149-152
                layer.keys = O.apply_rope(layer.keys, cos, sin)
        return layer.keys, layer.values
