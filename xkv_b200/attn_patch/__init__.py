"""Attention forwards that route HuggingFace models through the xKV cache (mirror of xKV/attn_patch)."""
from .llama import enable_llama_xKV_eval, xKV_llama_forward  # noqa: F401
from .mistral import enable_mistral_xKV_eval  # noqa: F401
from .qwen import enable_qwen_xKV_eval  # noqa: F401
from .deepseek_v2 import enable_deepseek_v2_xKV_eval  # noqa: F401
