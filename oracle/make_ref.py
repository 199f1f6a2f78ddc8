"""Recipe for oracle/_ref/ -- TEST INFRASTRUCTURE (checker / CPU baseline only, never imported by madtp_b200/).

The reference is pure Python, so "building" it for the GPU box means staging the files of its hot path, unmodified,
where they can be imported there: /root/reference does not exist on the GPU box, oracle/_ref/ travels with the
snapshot (it is git-ignored, NOT gpurun-ignored, exactly like the built .so files). Nothing is edited; every file is
copied byte for byte and its sha256 is recorded in oracle/_ref/MANIFEST.json. Reference sources never enter the git
history.

    python -m oracle.make_ref            # in the build container, where /root/reference exists

What is staged: models/{__init__,vit,utils,med,nlvr_encoder,blip,blip_nlvr,blip_retrieval,
blip_vqa}.py, models/linklink/*, clip/{__init__,model,mock,clip,simple_tokenizer}.py and the two JSON model configs.
oracle/ref_shims.py resolves the reference root as $MADTP_REFERENCE_ROOT, then /root/reference, then oracle/_ref.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference")
DST = ROOT / "oracle" / "_ref"

FILES = [
    "models/__init__.py", "models/vit.py", "models/utils.py", "models/med.py", "models/nlvr_encoder.py",
    "models/blip.py", "models/blip_nlvr.py", "models/blip_retrieval.py", "models/blip_vqa.py",
    "clip/__init__.py", "clip/model.py", "clip/mock.py", "clip/clip.py", "clip/simple_tokenizer.py",
    "clip/bpe_simple_vocab_16e6.txt.gz",     # read by clip/clip.py:30 at import time
    "configs/med_config.json", "configs/bert_config.json", "LICENSE",
]
DIRS = ["models/linklink"]


def stage(src: Path = SRC, dst: Path = DST) -> Path:
    if not (src / "models").is_dir():
        raise RuntimeError(f"{src} is not the reference tree")
    manifest = {}
    files = list(FILES)
    for d in DIRS:
        files += [str(p.relative_to(src)) for p in sorted((src / d).rglob("*.py"))]
    for rel in files:
        s, t = src / rel, dst / rel
        if not s.exists():
            continue
        t.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(s, t)
        manifest[rel] = hashlib.sha256(t.read_bytes()).hexdigest()
    (dst / "MANIFEST.json").write_text(json.dumps({"source": str(src), "files": manifest}, indent=1))
    return dst


if __name__ == "__main__":
    print(stage())
    sys.exit(0)
