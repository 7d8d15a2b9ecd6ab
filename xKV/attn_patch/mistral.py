from xkv_b200.attn_patch.mistral import *  # noqa: F401,F403
