"""Entry point of the xKV patch (mirror of the reference's ``xKV/patch.py``).

``KVCompress(xKV_config=..., yaml_path=...)(model)`` rebinds each layer's attention forward and makes
``generate()`` build a fresh ``FakeLayerMergingCache`` per call (reference patch.py:12-28, 32-73)."""
from __future__ import annotations

import logging
from typing import Dict, Optional

from transformers.cache_utils import Cache, DynamicCache

from .attn_patch.deepseek_v2 import enable_deepseek_v2_xKV_eval
from .attn_patch.llama import enable_llama_xKV_eval
from .attn_patch.mistral import enable_mistral_xKV_eval
from .attn_patch.qwen import enable_qwen_xKV_eval
from .configurations import xKVConfig
from .customized_cache import method_to_cache_obj

logger = logging.getLogger("xkv_b200")


def prepare_cache(method: str, config):
    """A replacement for ``GenerationMixin._prepare_cache_for_generation`` that installs the method's cache
    class (a plain DynamicCache when the method is unknown), reference patch.py:12-28."""
    cache_obj: Optional[type] = method_to_cache_obj.get(method, None)

    def _prepare_cache_for_generation(self, generation_config, model_kwargs: Dict, *args, **kwargs) -> bool:
        model_kwargs["past_key_values"] = DynamicCache() if cache_obj is None else cache_obj(config)

    return _prepare_cache_for_generation


class KVCompress:
    """Patch a HuggingFace causal LM with the xKV cache, from an ``xKVConfig`` or a YAML file."""

    def __init__(self, xKV_config: Optional[xKVConfig] = None, yaml_path: Optional[str] = None):  # noqa: N803
        if xKV_config is not None:
            self.config = xKV_config
        elif yaml_path is not None:
            self.config = xKVConfig.from_yaml(yaml_path)
        else:
            raise ValueError("Must provide either xKV_config or yaml_path.")

    def __call__(self, model):
        return self.enable_xKV_patch(model)

    def enable_xKV_patch(self, model):  # noqa: N802
        logger.info("Enabling xKV patch for model: %s", getattr(model.config, "architectures", None))
        model_type = model.config.model_type
        if "mistral" in model_type:
            enable_mistral_xKV_eval(model)
        elif "llama" in model_type:
            enable_llama_xKV_eval(model)
        elif "qwen" in model_type:
            enable_qwen_xKV_eval(model)
        elif "deepseek_v2" in model_type:
            enable_deepseek_v2_xKV_eval(model)
        else:
            raise ValueError("Model type not supported for xKV patch: {}".format(model_type))
        model.kv_compress_config = self.config
        prepare_cache_fn = prepare_cache("xKV", self.config)
        model._prepare_cache_for_generation = prepare_cache_fn.__get__(model, model.__class__)
        return model
