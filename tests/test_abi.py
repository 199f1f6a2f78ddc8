"""The C-ABI library loads on a CPU-only host and exports exactly what include/madtp_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "madtp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(madtp_\w+)\s*\(", text)))


def test_header_declares_functions():
    names = declared_functions()
    assert "madtp_gemm" in names and "madtp_dtp_select" in names and len(names) >= 18


def test_library_exports_every_declared_symbol(lib):
    cdll = lib.load()
    for name in declared_functions():
        assert hasattr(cdll, name), f"{name} is declared in include/madtp_b200.h but not exported"
    assert cdll.madtp_abi_version() == lib.ABI_VERSION == 4


def test_ctypes_signatures_cover_the_header(lib):
    assert sorted(lib.SIGNATURES) == declared_functions()


def test_binding_argument_counts_match_the_header(lib):
    """Every entry of _lib.SIGNATURES lists exactly as many arguments as the prototype in include/madtp_b200.h (the
    trampoline binding forwards whatever it is given, so a drifted signature would not fail at call time)."""
    text = (ROOT / "include" / "madtp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(madtp_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert sorted(protos) == sorted(lib.SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(lib.SIGNATURES[name]), f"{name}: header has {n} parameters, binding {len(lib.SIGNATURES[name])}"


def test_argument_errors_are_reported_without_a_gpu(lib):
    cdll = lib.load()
    # invalid shapes are rejected before any CUDA call, so this is safe on a CPU-only host
    st = cdll.madtp_dtp_select(1, 0, None, None, None, None, None, None, 0, None, None, 0, None, None, None, None)
    assert st == 1
    assert b"null pointer" in cdll.madtp_last_error_string() or b"out of range" in cdll.madtp_last_error_string()
    st = cdll.madtp_gemm(7, None, None, 0, None, None, 0, None, 0, 0, None, None, 0, 0, ctypes.c_float(1.0), 1, 1, 1,
                         None, 1, None, 1, None)
    assert st == 1


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU instead of silently computing on the CPU."""
    from functools import partial

    from madtp_b200.utils import Query_model, vector_gather
    from madtp_b200.vit import Block
    blk = Block(768, 12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6)).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(torch.randn(1, 5, 768))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Query_model(768, 768)(torch.randn(1, 4, 768), torch.randn(100, 768), return_token_att=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vector_gather(torch.randn(1, 4, 8), torch.zeros(1, 2, dtype=torch.long))


def test_product_does_not_import_the_oracle():
    for p in (ROOT / "madtp_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f"{p} imports the oracle"


def test_state_dict_keys_match_the_reference_layout():
    from madtp_b200.blip_nlvr import BLIP_NLVR
    from madtp_b200 import synthetic
    m = BLIP_NLVR(image_size=224, evaluate=True)
    sd = synthetic.blip_nlvr_state_dict(1234, img_size=224)
    msg = m.load_state_dict(sd, strict=False)
    assert not msg.missing_keys and not msg.unexpected_keys
    keys = set(m.state_dict())
    for k in ("space_dict", "visual_encoder.blocks.0.attn.qkv.weight", "visual_encoder.patch_embed.proj.weight",
              "text_encoder.encoder.layer.0.crossattention.self0.query.weight",
              "text_encoder.encoder.layer.6.crossattention.output.merge_layer.weight",
              "text_encoder.encoder.layer.0.crossattention.output.dense1.bias", "cls_head.2.weight"):
        assert k in keys
    assert "text_encoder.encoder.layer.5.crossattention.output.merge_layer.weight" not in keys
