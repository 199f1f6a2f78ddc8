"""Drop-in proof on the GPU: the reference's OWN task-model wiring (models/blip_nlvr.py, models/blip_retrieval.py,
models/blip_vqa.py, clip/model.py, staged unmodified in oracle/_ref) runs on top of the madtp_b200 mirrors after the
re-exports of INTEGRATION.md section 2, and reproduces the fixtures the unmodified reference generated. Each case runs
in its own process because it replaces `models.vit` / `models.utils` / `models.nlvr_encoder` / `models.med` in
sys.modules, which is what the integration does."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("case", ["nlvr", "checkpoint", "retrieval", "vqa", "clip"])
def test_reference_wiring_over_the_mirrors(lib, case):
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.fail(f"reference staging missing at {ref_shims.REFERENCE_ROOT}: run `python -m oracle.make_ref` "
                    "(or __graft_entry__.build()) in the build container")
    r = subprocess.run([sys.executable, "-m", "oracle.dropin", case], cwd=ROOT, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, f"oracle.dropin {case} failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["case"] == case
    print(out)
