from xkv_b200.attn_patch.llama import *  # noqa: F401,F403
