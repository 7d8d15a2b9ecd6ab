from xkv_b200.configurations import *  # noqa: F401,F403
from xkv_b200.configurations import LayerGroup, generate_consecutive_layer_groups, generate_consecutive_xKV_config, xKVConfig  # noqa: F401
