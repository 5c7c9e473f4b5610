#!/usr/bin/env python
"""Installs the UNMODIFIED reference (reczoo/FuxiCTR, read-only at /root/reference) under
baseline/_ref/ so that it travels to the GPU box with `gpurun` (baseline/_ref is git-ignored, not
gpurun-ignored).  Build container only:

    python baseline/install_ref.py

1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of the
   checkout>` — the `fuxictr` package exactly as setup.py ships it.  (--no-deps: keras_preprocessing,
   h5py and polars are not in the offline wheelhouse; they are import-time-only on this path and
   stubbed by baseline/refenv.py, SURVEY.md 8c.  The copy under /tmp is needed because setup.py
   writes build/ and *.egg-info into its source tree and /root/reference is read-only.)
2. What setup.py does NOT package but the reference's own entry points read at run time, copied
   verbatim into baseline/_ref/extras/: the five in-scope `model_zoo` directories (run_expid.py +
   config/*.yaml + src/), `demo/` (example3 = BASELINE config C1) and the `data/tiny_*` fixtures
   those YAMLs point at.

Nothing under baseline/_ref is product source or enters git history; tests and bench.py use it as
the live reference (the checker / the thing our numbers are compared with), never as the product.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FUXICTR_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
ZOO = ["DeepFM", "DCNv2", "DLRM", "DIN", "xDeepFM"]
DATA = ["tiny_npz", "tiny_parquet", "tiny_seq", "tiny_csv"]


def main():
    if not os.path.isdir(os.path.join(REF, "fuxictr")):
        raise SystemExit("no reference checkout at %s" % REF)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "docs"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DST, src]
        subprocess.check_call(cmd)
    extras = os.path.join(DST, "extras")
    for name in ZOO:
        shutil.copytree(os.path.join(REF, "model_zoo", name), os.path.join(extras, "model_zoo", name))
    shutil.copytree(os.path.join(REF, "demo"), os.path.join(extras, "demo"))
    for name in DATA:
        shutil.copytree(os.path.join(REF, "data", name), os.path.join(extras, "data", name))
    for root, dirs, _ in os.walk(DST):
        for d in list(dirs):
            if d == "__pycache__":
                shutil.rmtree(os.path.join(root, d))
                dirs.remove(d)
    print("installed the reference under", DST)


if __name__ == "__main__":
    main()
