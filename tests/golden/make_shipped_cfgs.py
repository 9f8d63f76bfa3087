#!/usr/bin/env python
"""Snapshot the `model` and `data` sections of the reference's ten shipped YAML configs (configs/**/*.yml) into
tests/golden/shipped_cfgs.json, so the tests that run on the GPU box (no /root/reference there) can build every
shipped configuration exactly as `get_model(cfg, n_classes)` receives it.

    python tests/golden/make_shipped_cfgs.py
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from multiagentperception_b200 import configs  # noqa: E402

REF = os.environ.get("W2C_REFERENCE_ROOT", "/root/reference")


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, "configs", "*", "*.yml"))):
        cfg = configs.load_yaml(path)
        out[os.path.basename(path)] = {"model": cfg["model"], "data": cfg["data"],
                                       "training": {"batch_size": cfg.get("training", {}).get("batch_size", 1)}}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shipped_cfgs.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", dst, sorted(out))


if __name__ == "__main__":
    main()
