#!/usr/bin/env python3
"""Regenerate envs/T1.yaml from the reference's task config (values only; run in the build container).

    python tools/make_config.py [/root/reference]
"""
import os
import sys

import yaml

HDR = open(os.path.join(os.path.dirname(__file__), "..", "envs", "T1.yaml")).read().split("basic:")[0] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "envs", "T1.yaml")) else ""

if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    cfg = yaml.load(open(os.path.join(ref, "envs", "T1.yaml")), Loader=yaml.FullLoader)
    out = os.path.join(os.path.dirname(__file__), "..", "envs", "T1.yaml")
    with open(out, "w") as f:
        f.write(HDR + yaml.dump(cfg, sort_keys=False, default_flow_style=None, width=120))
    print("wrote", out)
