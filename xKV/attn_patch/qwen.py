from xkv_b200.attn_patch.qwen import *  # noqa: F401,F403
