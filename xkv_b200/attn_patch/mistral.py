"""Mistral attention forward for the xKV cache (reference: xKV/attn_patch/mistral.py; identical to the
Llama forward except that ``sliding_window`` is forwarded to the attention interface, mistral.py:69)."""
from __future__ import annotations

from transformers.models.mistral.modeling_mistral import MistralAttention

from .llama import _bind, xKV_llama_forward


def xKV_mistral_forward(self, *args, **kwargs):  # noqa: N802
    kwargs.setdefault("sliding_window", getattr(self.config, "sliding_window", None))
    return xKV_llama_forward(self, *args, **kwargs)


def enable_mistral_xKV_eval(model):  # noqa: N802
    _bind(model, MistralAttention, xKV_mistral_forward, "MistralAttention")
