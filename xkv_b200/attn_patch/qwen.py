"""Qwen2 attention forward for the xKV cache.

The reference's Qwen patch (xKV/attn_patch/qwen.py:19-75) never compresses: it passes its ``cache_kwargs``
dict positionally into ``mode`` (qwen.py:41), so ``mode == 'prefill'`` is never true.  Only
the entry-point name is part of the surface; here it binds the same forward as Llama (Qwen2 attention has
the same q/k/v/o projection layout), so compression does take effect."""
from __future__ import annotations

from transformers.models.qwen2.modeling_qwen2 import Qwen2Attention

from .llama import _bind, xKV_llama_forward


def xKV_qwen2_forward(self, *args, **kwargs):  # noqa: N802
    # Qwen2 layers of type "sliding_attention" carry their window on the module (None for full attention)
    kwargs.setdefault("sliding_window", getattr(self, "sliding_window", None))
    return xKV_llama_forward(self, *args, **kwargs)


def enable_qwen_xKV_eval(model):  # noqa: N802
    _bind(model, Qwen2Attention, xKV_qwen2_forward, "Qwen2Attention")
