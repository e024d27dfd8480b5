#!/usr/bin/env python3
"""Pins hunyuanvideo_efficiency_b200/sweep.py's enumerators to the UNMODIFIED reference scripts.

Runs /root/reference/dynamic_enumeration.py, dynamic_enumeration_stride.py and dynamic_enumeration_stride_2.py (as
subprocesses, on the reference's own t_ops_config.json) in this container, and commits one sha256 per experiment
list — over the canonical JSON of [exp_1, exp_2, ...] — plus the counts to tests/golden/enumerators.json.
tests/test_metrics_host.py recomputes the same digests from the restated enumerators.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_enumerator_golden.py
"""
import glob
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = {"pool": "dynamic_enumeration.py", "stride": "dynamic_enumeration_stride.py", "stride2": "dynamic_enumeration_stride_2.py"}


def digest(configs):
    return hashlib.sha256(json.dumps(configs, sort_keys=True, separators=(",", ":")).encode()).hexdigest()


def main():
    base = os.path.join(REF, "t_ops_config.json")
    # the base config (the fork's t_ops_config.json, an input data file) travels with the fixture: /root/reference does not
    out = {"base_sha256": hashlib.sha256(open(base, "rb").read()).hexdigest(), "base": json.load(open(base))}
    for mode, script in SCRIPTS.items():
        with tempfile.TemporaryDirectory() as d:
            src = open(os.path.join(REF, script)).read()
            # dynamic_enumeration.py:93 and dynamic_enumeration_stride.py:105 hard-code their author's output directory
            # (output_dir = "/mnt/public/..."): redirect that one literal; _stride_2.py takes the directory as argv[2]
            src, n = re.subn(r'output_dir = "/mnt/public/[^"]*"', "output_dir = " + repr(d), src)
            assert n == (0 if mode == "stride2" else 1), (mode, n)
            subprocess.run([sys.executable, "-c", src, base, d], check=True, stdout=subprocess.DEVNULL)
            files = sorted(glob.glob(os.path.join(d, "exp_*.json")), key=lambda f: int(os.path.basename(f)[4:-5]))
            configs = [json.load(open(f)) for f in files]
        out[mode] = {"count": len(configs), "sha256": digest(configs), "first": configs[0], "last": configs[-1]}
        print(mode, len(configs), out[mode]["sha256"])
    with open(os.path.join(ROOT, "tests", "golden", "enumerators.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
