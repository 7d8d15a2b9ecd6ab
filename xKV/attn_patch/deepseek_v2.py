from xkv_b200.attn_patch.deepseek_v2 import *  # noqa: F401,F403
