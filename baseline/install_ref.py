#!/usr/bin/env python
"""Install the UNMODIFIED reference hot-path files into the git-ignored ``baseline/_ref/``.

    python baseline/install_ref.py [--src /root/reference]

The reference is pure Python with no build step, so "installing" it is a byte-for-byte copy of the files on the
SURVEY.md section 8(a) path.  ``baseline/_ref/`` is listed in .gitignore (never committed -- the repo holds no
reference source) but not in .gpurunignore, so it travels to the GPU box, where ``bench.py --impl reference`` and the
``reference_cuda_eager`` leg of ``bench.py`` run it through ``baseline/ref_loader.py``.  A MANIFEST with the sha256
of every copied file is written next to the copy so a run can state exactly what it timed.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

FILES = [
    "model/__init__.py", "model/base.py", "model/embedder.py", "model/encoder.py", "model/head.py",
    "model/baseline.py", "model/mlp.py",
    "loss/__init__.py", "loss/eig.py", "loss/mle.py", "loss/distance.py",
    "tasks/__init__.py", "tasks/base_task.py", "tasks/location_finding.py", "tasks/ces.py", "tasks/psychometric.py",
    "tasks/gaussian_process.py", "tasks/al_benchmarks.py", "tasks/hpo.py",
    "distributions/__init__.py", "distributions/censored_sigmoid_normal.py", "distributions/gmm.py",
    "distributions/truncated_normal.py",
    "utils/eval.py", "utils/target_mask.py", "utils/misc.py",
    "train_aline.py", "requirements.txt",
]


def install(src="/root/reference", quiet=False):
    if not os.path.isdir(src):
        raise FileNotFoundError(f"reference not found at {src}")
    manifest = {}
    for rel in FILES:
        s = os.path.join(src, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"installed {len(manifest)} reference files into {DST}")
    return DST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("ALINE_REFERENCE", "/root/reference"))
    a = ap.parse_args()
    install(a.src)
    sys.exit(0)
