"""Llama attention forward for the xKV cache (reference: xKV/attn_patch/llama.py:19-88).

Semantics kept from the reference:
* RoPE is applied to Q only before the cache is touched (llama.py:39-40);
* prefill (q_len > 1) hands the PRE-RoPE keys plus cos/sin to ``cache.update(..., mode='prefill')`` and then
  attends on the ORIGINAL keys/values (llama.py:45-50) — compression only affects later decode steps;
* decode appends the post-RoPE key (llama.py:51-53) and attends over the cache;
* only the sdpa attention implementation is accepted (llama.py:55-56).
What differs: the decode step calls ``cache.attend`` — the fused reconstruct + RoPE + GQA-softmax kernel over
the stored factors — instead of SDPA over dense reconstructed tensors.  Written against the installed
transformers (``past_key_values`` kwarg; the reference's ``past_key_value`` spelling is accepted too).
"""
from __future__ import annotations

import types
from typing import Optional, Tuple

import torch
from transformers.cache_utils import Cache
from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
from transformers.models.llama.modeling_llama import LlamaAttention, apply_rotary_pos_emb

from ..customized_cache.fake_layer_merge_dynamic_cache import FakeLayerMergingCache


def xKV_llama_forward(  # noqa: N802
    self,
    hidden_states: torch.Tensor,
    position_embeddings: Tuple[torch.Tensor, torch.Tensor] = None,
    attention_mask: Optional[torch.Tensor] = None,
    past_key_values: Optional[Cache] = None,
    past_key_value: Optional[Cache] = None,
    cache_position: Optional[torch.LongTensor] = None,
    **kwargs,
):
    cache = past_key_values if past_key_values is not None else past_key_value
    input_shape = hidden_states.shape[:-1]
    q_len = hidden_states.shape[1]
    hidden_shape = (*input_shape, -1, self.head_dim)
    query_states = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
    key_states = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
    value_states = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)

    cos, sin = position_embeddings
    is_prefill = q_len > 1  # auto-regressive use, as in the reference
    query_states, _ = apply_rotary_pos_emb(query_states, query_states, cos, sin)

    if cache is not None:
        if is_prefill:
            assert isinstance(cache, FakeLayerMergingCache)
            cache.update(key_states, value_states, self.layer_idx, mode="prefill", cos=cos, sin=sin,
                         return_dense=False)
            key_states, _ = apply_rotary_pos_emb(key_states, key_states, cos, sin)
        else:
            key_pre_rope = key_states
            key_states, _ = apply_rotary_pos_emb(key_states, key_states, cos, sin)
            fused = None
            if isinstance(cache, FakeLayerMergingCache) and getattr(self, "xkv_fused_decode", True):
                fused = cache.attend(query_states, key_states, value_states, self.layer_idx, self.scaling,
                                     key_pre_rope=key_pre_rope, cos=cos, sin=sin)
            if fused is not None:
                attn_output = fused.transpose(1, 2).reshape(*input_shape, -1).contiguous()
                return self.o_proj(attn_output), None
            key_states, value_states = cache.update(key_states, value_states, self.layer_idx, mode="decode")

    if self.config._attn_implementation != "sdpa":
        raise ValueError("Only sdpa is supported for now")
    attention_interface = ALL_ATTENTION_FUNCTIONS["sdpa"]
    attn_output, attn_weights = attention_interface(
        self,
        query_states,
        key_states,
        value_states,
        attention_mask,
        dropout=0.0 if not self.training else self.attention_dropout,
        scaling=self.scaling,
        **kwargs,
    )
    attn_output = attn_output.reshape(*input_shape, -1).contiguous()
    attn_output = self.o_proj(attn_output)
    return attn_output, attn_weights


def _bind(model, expected_cls, forward, what: str):
    for layer in model.model.layers:
        module = layer.self_attn
        if not isinstance(module, expected_cls):
            raise ValueError(f"Only {what} is supported for now")
        module.forward = types.MethodType(forward, module)


def enable_llama_xKV_eval(model):  # noqa: N802
    """Rebind every layer's ``self_attn.forward`` (reference llama.py:77-88)."""
    _bind(model, LlamaAttention, xKV_llama_forward, "LlamaAttention")
