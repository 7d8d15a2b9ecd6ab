#!/usr/bin/env python
"""Same command line as the reference's group_layers.py; the implementation lives in xkv_b200/group_layers.py."""
from xkv_b200.group_layers import main

if __name__ == "__main__":
    main()
