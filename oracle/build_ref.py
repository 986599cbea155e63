#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference package, taken from where it lies under /root/reference.

    python oracle/build_ref.py        ->  oracle/_ref/evenvizion/**.py  +  oracle/_ref/imutils.py

The reference is pure Python (no build step), so "building" it is a byte-for-byte copy of its `evenvizion/**/*.py`
files (sha256 checked after the copy) into oracle/_ref/, which is git-ignored -- the sources are NOT committed -- but
travels to the GPU box with the repository snapshot like a built .so.  One file is written beside it, not copied:
a stand-in for the `imutils` package (not installed in this image; the reference uses `imutils.resize` and
`imutils.is_cv3` only).  `cv2.xfeatures2d.SIFT_create` (opencv-contrib 3.4 API, requirements.txt:3) is aliased to
`cv2.SIFT_create` at import time by `oracle.ref_runner`, not by editing the reference.

Test infrastructure only (see oracle/__init__.py): used by tests/ and by bench.py's CPU arms (`cpu_baseline`,
`--impl reference`, kind "reference").  /root/reference does not exist on the GPU box; this script is a no-op there.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")

IMUTILS = '''"""Stand-in for the two `imutils` functions the reference calls (imutils 0.5.3 semantics)."""
import cv2


def resize(image, width=None, height=None, inter=cv2.INTER_AREA):
    h, w = image.shape[:2]
    if width is None and height is None:
        return image
    if width is None:
        r = height / float(h)
        dim = (int(w * r), height)
    else:
        r = width / float(w)
        dim = (width, int(h * r))
    return cv2.resize(image, dim, interpolation=inter)


def is_cv3(or_better=False):
    return True
'''


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose=True):
    src_root = os.path.join(REF, "evenvizion")
    if not os.path.isdir(src_root):
        if verbose:
            print("oracle/build_ref.py: %s not present (GPU box): using the prebuilt oracle/_ref if any" % REF)
        return os.path.isdir(os.path.join(DST, "evenvizion"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for d, _, files in os.walk(src_root):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            s = os.path.join(d, fn)
            t = os.path.join(DST, os.path.relpath(s, REF))
            os.makedirs(os.path.dirname(t), exist_ok=True)
            shutil.copyfile(s, t)
            assert sha(s) == sha(t), s
            n += 1
    with open(os.path.join(DST, "imutils.py"), "w") as f:
        f.write(IMUTILS)
    if verbose:
        print("oracle/_ref: %d reference files copied unmodified from %s" % (n, src_root))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
