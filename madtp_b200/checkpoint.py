"""Checkpoint loading for the task-model mirrors: the reference's `load_checkpoint` helpers
restated over state dicts, so that released checkpoints ({'model', 'epoch', 'temperature'},
compress_nlvr_dtp.py:229-236) load into the B200 path with the same key handling.

  load_blip_checkpoint   models/blip.py:254-278   (BLIP retrieval / VQA / caption encoders)
  load_nlvr_checkpoint   models/blip_nlvr.py:131-160 (cross-attention weights fan out to the twin branches)
  clip_model.build_model clip/model.py:678-716     (architecture derived from the shapes in the checkpoint)

Every loader finishes with `functional.clear_caches(model)`: the GEMM-ready weight copies are rebuilt on the next
forward.
"""
from __future__ import annotations

import os
from typing import Mapping, Union

import torch

from . import functional as Fn
from .vit import interpolate_pos_embed


def _state_dict(src: Union[str, os.PathLike, Mapping]):
    """A checkpoint path, a {'model': state_dict, ...} checkpoint or a bare state dict -> (state_dict copy, temperature)."""
    if isinstance(src, (str, os.PathLike)):
        if not os.path.isfile(src):
            raise RuntimeError('checkpoint url or path is invalid')          # models/blip.py:264
        src = torch.load(src, map_location='cpu')
    temperature = src.get('temperature') if isinstance(src, Mapping) else None
    sd = src['model'] if (isinstance(src, Mapping) and 'model' in src and isinstance(src['model'], Mapping)) else src
    return dict(sd), temperature


def load_blip_checkpoint(model, src):
    """models/blip.py:254-278: bicubic resize of the position grid to the model's image size (also for a momentum
    twin when the model has one), keys whose shape does not match the model are dropped, strict=False.
    Returns (model, msg)."""
    sd, _ = _state_dict(src)
    sd['visual_encoder.pos_embed'] = interpolate_pos_embed(sd['visual_encoder.pos_embed'], model.visual_encoder)
    own = model.state_dict()
    if 'visual_encoder_m.pos_embed' in own and 'visual_encoder_m.pos_embed' in sd:
        sd['visual_encoder_m.pos_embed'] = interpolate_pos_embed(sd['visual_encoder_m.pos_embed'], model.visual_encoder_m)
    for key in own.keys():
        if key in sd and sd[key].shape != own[key].shape:
            del sd[key]
    msg = model.load_state_dict(sd, strict=False)
    Fn.clear_caches(model)
    return model, msg


def load_nlvr_checkpoint(model, src):
    """models/blip_nlvr.py:131-160: position-grid resize, then every `crossattention.self.*` tensor is duplicated into
    `self0` / `self1` and every `crossattention.output.dense.*` into `dense0` / `dense1` (the pretrained BLIP has ONE
    cross-attention; BLIP-NLVR runs one per image). Returns (model, msg)."""
    sd, _ = _state_dict(src)
    sd['visual_encoder.pos_embed'] = interpolate_pos_embed(sd['visual_encoder.pos_embed'], model.visual_encoder)
    for key in list(sd.keys()):
        if 'crossattention.self.' in key:
            sd[key.replace('self', 'self0')] = sd[key]
            sd[key.replace('self', 'self1')] = sd[key]
        elif 'crossattention.output.dense.' in key:
            sd[key.replace('dense', 'dense0')] = sd[key]
            sd[key.replace('dense', 'dense1')] = sd[key]
    msg = model.load_state_dict(sd, strict=False)
    Fn.clear_caches(model)
    return model, msg


def load_compressed_checkpoint(model, src):
    """Evaluation of a compressed model (compress_nlvr_dtp.py:153-158): load_state_dict(strict=False) of ckpt['model']
    and the calibrated `temperature` stored next to it. Returns (msg, temperature)."""
    sd, temperature = _state_dict(src)
    msg = model.load_state_dict(sd, strict=False)
    Fn.clear_caches(model)
    return msg, temperature
