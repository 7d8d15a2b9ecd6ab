"""DeepSeek-V2 (MLA) attention forward for the xKV cache (reference: xKV/attn_patch/deepseek_v2.py:160-302).

Semantics kept from the reference:
* the cache holds the LATENT ``compressed_kv`` (b, 1, l, kv_lora_rank) in the key slot and the RoPE'd ``k_pe``
  (b, 1, l, qk_rope_head_dim) in the value slot (:217-232);
* latents carry no RoPE, so ``re_apply_rope=False`` and ``cos = sin = None`` (:226, :231);
* only the latents are compressed: ``merge_value`` must be off (:222-223);
* the cache's return IS used, in prefill too (:224): the last layer of a group already attends on the
  compressed latents, earlier layers of the group on their exact ones;
* every step the (reconstructed) latents go through ``kv_a_layernorm`` + ``kv_b_proj`` (:235).
What differs: written against transformers' native ``DeepseekV2Attention`` (the reference patches the Hub's
remote-code ``DeepseekV2FlashAttention2``, which is not vendored), attention through the installed sdpa
interface instead of flash-attn 2, and the latents come back from the factors through the tcgen05 GEMM
(``FakeLayerMergingCache.materialize``) instead of being stored dense.
"""
from __future__ import annotations

import types
from typing import Optional

import torch
from transformers.cache_utils import Cache
from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
from transformers.models.deepseek_v2.modeling_deepseek_v2 import DeepseekV2Attention, apply_rotary_emb

from ..customized_cache.fake_layer_merge_dynamic_cache import FakeLayerMergingCache


def xKV_mla_forward(  # noqa: N802
    self,
    hidden_states: torch.Tensor,
    attention_mask: Optional[torch.Tensor] = None,
    past_key_values: Optional[Cache] = None,
    position_embeddings: Optional[torch.Tensor] = None,
    past_key_value: Optional[Cache] = None,
    **kwargs,
):
    cache = past_key_values if past_key_values is not None else past_key_value
    bsz, q_len = hidden_states.shape[:-1]
    is_prefill = q_len > 1  # auto-regressive use, as in the reference

    if self.q_lora_rank is None:
        q = self.q_proj(hidden_states)
    else:
        q = self.q_b_proj(self.q_a_layernorm(self.q_a_proj(hidden_states)))
    q = q.view(bsz, q_len, -1, self.qk_head_dim).transpose(1, 2)
    q_nope, q_pe = torch.split(q, [self.qk_nope_head_dim, self.qk_rope_head_dim], dim=-1)

    compressed_kv = self.kv_a_proj_with_mqa(hidden_states)
    compressed_kv, k_pe = torch.split(compressed_kv, [self.kv_lora_rank, self.qk_rope_head_dim], dim=-1)
    k_pe = k_pe.view(bsz, 1, q_len, self.qk_rope_head_dim)
    q_pe, k_pe = apply_rotary_emb(q_pe, k_pe, position_embeddings.to(q_pe.device))

    latent = compressed_kv.view(bsz, q_len, 1, self.kv_lora_rank).transpose(1, 2)   # (b, 1, l, kv_lora_rank)
    if cache is not None:
        if is_prefill:
            assert isinstance(cache, FakeLayerMergingCache)
            if cache.is_value_merged():
                raise ValueError("DeepseekV2Attention does not support --merge_v")
            latent, k_pe = cache.update(latent, k_pe.contiguous(), self.layer_idx, mode="prefill", cos=None, sin=None,
                                        re_apply_rope=False)
        else:
            latent, k_pe = cache.update(latent, k_pe.contiguous(), self.layer_idx, mode="decode", cos=None, sin=None,
                                        re_apply_rope=False)
    kv_len = latent.shape[-2]
    kv = self.kv_b_proj(self.kv_a_layernorm(latent.squeeze(1)))
    kv = kv.view(bsz, kv_len, -1, self.qk_nope_head_dim + self.v_head_dim).transpose(1, 2)
    k_nope, value_states = torch.split(kv, [self.qk_nope_head_dim, self.v_head_dim], dim=-1)

    query_states = torch.cat((q_nope, q_pe), dim=-1)
    key_states = torch.cat((k_nope, k_pe.expand(*k_nope.shape[:-1], -1)), dim=-1)

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
    attn_output = attn_output.reshape(bsz, q_len, -1).contiguous()
    return self.o_proj(attn_output), attn_weights


def enable_deepseek_v2_xKV_eval(model):  # noqa: N802
    """Rebind every MLA layer's forward (reference deepseek_v2.py:290-302)."""
    for layer in model.model.layers:
        module = layer.self_attn
        if not isinstance(module, DeepseekV2Attention):
            raise ValueError("Only DeepseekV2Attention is supported for now")
        module.forward = types.MethodType(xKV_mla_forward, module)
