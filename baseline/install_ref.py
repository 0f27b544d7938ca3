"""Recipe for the reference arm of bench.py: places an UNMODIFIED copy of the reference's Python tree under baseline/_ref/.

The reference (Yuki-11/CSBSR) is a script tree without setup.py / pyproject.toml, so `pip install --target baseline/_ref
/root/reference` has nothing to install; the equivalent is a verbatim copy of its importable packages.  baseline/_ref/ is
git-ignored (no reference source enters the history) but NOT gpurun-ignored, so the copy travels to the GPU box, where
`bench.py --impl reference` imports it through oracle/ref_harness.py (shims for the absent yacs / skimage / timm / matplotlib)
and times the reference's own code on the host cores.  Run by __graft_entry__.build() whenever /root/reference is present.

  python baseline/install_ref.py [--src /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
WHAT = ["model", "config", "test.py", "train.py", "LICENSE"]


def _digest(root):
    h = hashlib.sha256()
    n = 0
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".py") or f.endswith(".yaml") or f.endswith(".json"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
                n += 1
    return h.hexdigest(), n


def install(src="/root/reference", quiet=False):
    if not os.path.isdir(os.path.join(src, "model")):
        return False
    os.makedirs(DEST, exist_ok=True)
    for w in WHAT:
        s, d = os.path.join(src, w), os.path.join(DEST, w)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "*.png", "*.jpg"))
        else:
            shutil.copy2(s, d)
    dig, n = _digest(DEST)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": n, "sha256": dig, "modified": False}, f)
    if not quiet:
        print("baseline/_ref: %d reference files copied verbatim from %s (sha256 %s)" % (n, src, dig[:16]))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    if not install(a.src):
        raise SystemExit("no reference tree at %s" % a.src)
