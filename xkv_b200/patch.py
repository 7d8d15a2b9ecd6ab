"""Entry point of the xKV patch.

Public surface as in the reference (``xKV/patch.py:12-73``): ``prepare_cache(method, config)`` and
``KVCompress(xKV_config=None, yaml_path=None)`` whose call / ``enable_xKV_patch(model)`` mutates and returns the model:

* every decoder layer's attention forward is rebound to the family's xKV forward,
* ``model.kv_compress_config`` holds the configuration,
* ``generate()`` gets a FRESH cache object per call through ``model._prepare_cache_for_generation``.

Families are recognised by substring of ``model.config.model_type`` in the reference's order (a Mistral checkpoint whose
type string also contains "llama" must take the Mistral branch, patch.py:56-63)."""
from __future__ import annotations

import logging
import types
from typing import Callable, Optional, Sequence, Tuple

from transformers.cache_utils import DynamicCache

from .attn_patch import deepseek_v2 as _deepseek_v2
from .attn_patch import llama as _llama
from .attn_patch import mistral as _mistral
from .attn_patch import qwen as _qwen
from .configurations import xKVConfig
from .customized_cache import method_to_cache_obj

log = logging.getLogger("xkv_b200")

# (substring of config.model_type, function that rebinds the attention forwards), tried top to bottom
_FAMILIES: Sequence[Tuple[str, Callable]] = (
    ("mistral", _mistral.enable_mistral_xKV_eval),
    ("llama", _llama.enable_llama_xKV_eval),
    ("qwen", _qwen.enable_qwen_xKV_eval),
    ("deepseek_v2", _deepseek_v2.enable_deepseek_v2_xKV_eval),
)


def prepare_cache(method: str, config):
    """Build the function that replaces ``GenerationMixin._prepare_cache_for_generation``: it drops a new cache of
    the class registered under ``method`` (``"xKV"`` -> FakeLayerMergingCache) into ``model_kwargs["past_key_values"]``;
    an unregistered method falls back to transformers' DynamicCache."""
    factory = method_to_cache_obj.get(method)

    def _prepare_cache_for_generation(self, generation_config, model_kwargs, *args, **kwargs):
        model_kwargs["past_key_values"] = factory(config) if factory is not None else DynamicCache()

    return _prepare_cache_for_generation


def _enabler_for(model_type: str) -> Callable:
    for needle, enable in _FAMILIES:
        if needle in model_type:
            return enable
    raise ValueError("Model type not supported for xKV patch: {}".format(model_type))


class KVCompress:
    """Callable that installs the xKV cache and attention forwards on a HuggingFace causal LM."""

    def __init__(self, xKV_config: Optional[xKVConfig] = None, yaml_path: Optional[str] = None):  # noqa: N803
        if xKV_config is None and yaml_path is None:
            raise ValueError("Must provide either xKV_config or yaml_path.")
        # an explicit configuration object wins over a YAML path, as in the reference
        self.config = xKV_config if xKV_config is not None else xKVConfig.from_yaml(yaml_path)

    def enable_xKV_patch(self, model):  # noqa: N802
        enable = _enabler_for(model.config.model_type)
        log.info("xKV patch on %s (%s)", getattr(model.config, "architectures", None), model.config.model_type)
        enable(model)
        model.kv_compress_config = self.config
        model._prepare_cache_for_generation = types.MethodType(prepare_cache("xKV", self.config), model)
        return model

    __call__ = enable_xKV_patch
