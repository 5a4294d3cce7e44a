"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that places the UNMODIFIED reference hot-path modules under oracle/_ref/.

The reference is pure Python, so "building" it is copying the two files its hot path consists of (they import only torch + einops):

    MaXTron_Video-kMaX/maxtron_deeplab/modeling/within_clip_tracking_module/temporal_attention.py   (TrajectoryAttention, layers, TemporalEncoder)
    MaXTron_Video-kMaX/maxtron_deeplab/modeling/within_clip_tracking_module/pos_embeddings.py       (PositionEmbeddingSine3D)

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so the files travel to
the GPU box like the built libaxvs.so does, and `bench.py --impl reference` / `cpu_baseline` / `gpu_eager_baseline` time the
reference ITSELF there (`kind: "reference"`); without them they fall back to the oracle port (`kind: "port"`).
`__graft_entry__.build()` runs this when /root/reference is present.  Usage: python oracle/build_ref.py
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("AXVS_REFERENCE_ROOT", "/root/reference")
WC = os.path.join(REF_ROOT, "MaXTron_Video-kMaX/maxtron_deeplab/modeling/within_clip_tracking_module")
FILES = ["temporal_attention.py", "pos_embeddings.py"]
OUT = os.path.join(HERE, "_ref")


def build() -> bool:
    """Copy the reference files (verbatim) into oracle/_ref/; returns False when the reference tree is not mounted."""
    if not all(os.path.isfile(os.path.join(WC, f)) for f in FILES):
        return False
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for f in FILES:
        src = os.path.join(WC, f)
        shutil.copyfile(src, os.path.join(OUT, f))
        with open(src, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": WC, "sha256": manifest}, fh, indent=1)
    return True


def load():
    """(temporal_attention module, pos_embeddings module) imported from oracle/_ref, or None when the directory is absent."""
    import importlib.util
    import sys
    mods = []
    for f in FILES:
        path = os.path.join(OUT, f)
        if not os.path.isfile(path):
            return None
        name = "axvs_ref_" + f[:-3]
        if name in sys.modules:
            mods.append(sys.modules[name])
            continue
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


if __name__ == "__main__":
    print("oracle/_ref:", "built" if build() else f"reference tree not found under {REF_ROOT}")
