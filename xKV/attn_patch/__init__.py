from xkv_b200.attn_patch import *  # noqa: F401,F403
