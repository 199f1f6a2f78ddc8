"""Mirror of the reference's models/blip_nlvr.py evaluation forward (BLIP_NLVR.forward(train=False), :63-81,99-100):
pruned ViT on cat(image0, image1) -> split -> twin cross-attention text encoder -> cls_head. Same constructor
arguments, attribute names and state-dict keys (`space_dict`, `visual_encoder.*`, `text_encoder.*`, `cls_head.*`).

The HuggingFace tokenizer is host-side I/O outside the hot path: pass `tokenizer=` (any callable with the
BertTokenizer call signature and an `enc_token_id`), or hand `forward` pre-tokenised text (an object with
`input_ids` / `attention_mask`, as the reference's own forward_throughput does, :103-120).
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import functional as Fn
from .configuration import BertConfig
from .nlvr_encoder import BertModel
from .vit import VisionTransformer, interpolate_pos_embed

ENC_TOKEN_ID = 30523    # id of '[ENC]' after models/blip.py:222-224 extends bert-base-uncased (30522 + [DEC] + [ENC])


def create_vit(vit, image_size, use_grad_checkpointing=False, ckpt_layer=0, drop_path_rate=0, evaluate=False,
               sd_dim=768, map_func=False):
    """models/blip.py:228-247."""
    if vit == 'base':
        vision_width = 768
        enc = VisionTransformer(img_size=image_size, patch_size=16, embed_dim=vision_width, depth=12, num_heads=12,
                                evaluate=evaluate, sd_dim=sd_dim, map_func=map_func)
    elif vit == 'large':
        vision_width = 1024
        enc = VisionTransformer(img_size=image_size, patch_size=16, embed_dim=vision_width, depth=24, num_heads=16,
                                evaluate=evaluate, sd_dim=sd_dim, map_func=map_func)
    else:
        raise ValueError(vit)
    return enc, vision_width


class TokenizedText:
    def __init__(self, input_ids, attention_mask):
        self.input_ids, self.attention_mask = input_ids, attention_mask

    def to(self, device):
        return TokenizedText(self.input_ids.to(device), self.attention_mask.to(device))


class BLIP_NLVR(nn.Module):
    def __init__(self, med_config='configs/med_config.json', image_size=480, vit='base', vit_grad_ckpt=False,
                 vit_ckpt_layer=0, evaluate=False, config=None, tokenizer=None):
        super().__init__()
        self.layers = 12 if vit == 'base' else 24
        if config is None:
            self.sd_num, self.sd_dim, self.batch_size = 100, 768, 16
        else:
            self.sd_num, self.sd_dim, self.batch_size = config['sd_num'], config['sd_dim'], config['batch_size_train']
        self.space_dict = nn.Parameter(torch.randn(self.sd_num, self.sd_dim))
        self.world_size = int(os.environ.get('WORLD_SIZE', 1))
        self.visual_encoder, vision_width = create_vit(vit, image_size, vit_grad_ckpt, vit_ckpt_layer,
                                                       drop_path_rate=0.1, evaluate=evaluate, sd_dim=self.sd_dim)
        self.tokenizer = tokenizer
        cfg = BertConfig.from_json_file(med_config) if (isinstance(med_config, str) and os.path.isfile(med_config)) \
            else (med_config if isinstance(med_config, BertConfig) else BertConfig())
        cfg.encoder_width = vision_width
        cfg.evaluate = evaluate
        self.text_encoder = BertModel(config=cfg, add_pooling_layer=False, sd_dim=self.sd_dim)
        hs = cfg.hidden_size
        self.cls_head = nn.Sequential(nn.Linear(hs, hs), nn.ReLU(), nn.Linear(hs, 2))
        self._cache = Fn.WeightCache()
        self.record_states = False      # keep fp32 copies of image_embeds / last_hidden_state in self.last (tests)
        self._graphs = None             # shape key -> graphs.GraphedCall, when enable_cuda_graphs(True)
        self.last = {}

    def enable_cuda_graphs(self, enable: bool = True):
        """Capture the pruned forward (device-resident lengths) in a CUDA graph per input shape and temperature and
        replay it on later calls. The graph bakes in the addresses of the weights and of their GEMM-ready copies: call
        `reset_cuda_graphs()` after loading or changing weights."""
        self._graphs = {} if enable else None
        return self

    def reset_cuda_graphs(self):
        if self._graphs is not None:
            self._graphs = {}

    def _cls_head(self, hidden_state):
        from . import _lib as L
        l0, l2 = self.cls_head[0], self.cls_head[2]
        w0 = self._cache.get("c0", [l0.weight, l0.bias], lambda: Fn.PreparedLinear(l0.weight, l0.bias, f32=True))
        w2 = self._cache.get("c2", [l2.weight, l2.bias], lambda: Fn.PreparedLinear(l2.weight, l2.bias, f32=True))
        return Fn.linear_f32(Fn.linear_f32(hidden_state, w0, act=L.ACT_RELU), w2)

    def _forward_device(self, image, input_ids, attention_mask, temperature):
        """blip_nlvr.py:63-81 with device-resident lengths from the first ViT layer to the logits: nothing is read
        back, every launch argument is data-independent. Returns (logits [P, 2], [image trajectory, text trajectory])."""
        from . import _lib as L
        enc_id = getattr(self.tokenizer, "enc_token_id", ENC_TOKEN_ID)
        input_ids[:, 0] = enc_id                                                            # blip_nlvr.py:69
        keep = self.record_states
        enc = self.visual_encoder.forward_device(image, self.space_dict, temperature, pack_groups=2, keep_f32=keep)
        h, sd_txt_ft, traj, l_dev = self.text_encoder.forward_device(input_ids, attention_mask, enc, self.space_dict,
                                                                     temperature)
        B, Lcap, d = h.shape
        hidden_state = L.take_token(h.view(B * Lcap, d), B, Lcap, 0, n_dev=l_dev)           # last_hidden_state[:, 0, :]
        self.last = _LazyStates(enc, h, traj, sd_txt_ft) if keep else {"sd_img_ft": enc.sd_ft, "sd_txt_ft": sd_txt_ft}
        return self._cls_head(hidden_state), [enc.traj, traj]

    def _tokenize(self, text, device):
        if hasattr(text, "input_ids"):
            return text.input_ids.to(device).clone(), text.attention_mask.to(device)
        if isinstance(text, (tuple, list)) and len(text) == 2 and torch.is_tensor(text[0]):
            return text[0].to(device).clone(), text[1].to(device)
        if self.tokenizer is None:
            raise RuntimeError("madtp_b200.BLIP_NLVR: pass tokenizer= or pre-tokenised text (input_ids, attention_mask)")
        t = self.tokenizer(text, padding='longest', return_tensors="pt").to(device)
        return t.input_ids.clone(), t.attention_mask

    @torch.no_grad()
    def forward(self, image, text, targets, temperature=0, train=True):
        if train:
            raise NotImplementedError("madtp_b200 implements the evaluation forward (train=False) only")
        from .vit import device_lengths_enabled
        if temperature > 0 and device_lengths_enabled():
            Fn.require_cuda(image, "image")
            input_ids, attention_mask = self._tokenize(text, image.device)
            if input_ids.shape[1] <= 64 and image.shape[0] == 2 * input_ids.shape[0]:
                from . import _lib as L
                if self._graphs is None:
                    with L.arena_for(self, (tuple(image.shape), tuple(input_ids.shape), self.record_states)):
                        return self._forward_device(image.contiguous(), input_ids, attention_mask,
                                                    float(temperature))[0].clone()
                from .graphs import GraphedCall
                # one captured graph (with its own buffers) per input shape, temperature AND stream: forwards issued on
                # different streams may overlap on the device (two batches in flight, bench.py --streams 2)
                key = (tuple(image.shape), tuple(input_ids.shape), float(temperature), self.record_states,
                       torch.cuda.current_stream().cuda_stream) + L.mode_key()
                g = self._graphs.get(key)
                if g is None:
                    t = float(temperature)
                    g = self._graphs[key] = GraphedCall(lambda im, ids, m: self._forward_device(im, ids, m, t),
                                                        [image.contiguous(), input_ids, attention_mask])
                return g(image, input_ids, attention_mask)
        image_embeds, sd_img_ft = self.visual_encoder(image, space_dict=self.space_dict, temperature=temperature)
        P = targets.size(0) if torch.is_tensor(targets) else int(targets)
        image0_embeds, image1_embeds = image_embeds[:P], image_embeds[P:]                   # blip_nlvr.py:67
        input_ids, attention_mask = self._tokenize(text, image.device)
        enc_id = getattr(self.tokenizer, "enc_token_id", ENC_TOKEN_ID)
        input_ids[:, 0] = enc_id                                                            # blip_nlvr.py:69
        # image_atts is all ones (blip_nlvr.py:66) -> the additive encoder mask is identically zero: pass None
        output, sd_txt_ft = self.text_encoder(input_ids, attention_mask=attention_mask,
                                              encoder_hidden_states=[image0_embeds, image1_embeds],
                                              encoder_attention_mask=None, return_dict=True,
                                              space_dict=self.space_dict, temperature=temperature)
        hidden_state = output.last_hidden_state[:, 0, :].contiguous()
        self.last = {"image_embeds": image_embeds, "last_hidden_state": output.last_hidden_state,
                     "sd_img_ft": sd_img_ft, "sd_txt_ft": sd_txt_ft}
        return self._cls_head(hidden_state)


class _LazyStates:
    """`model.last` of a device-resident-length forward: narrowed views materialise on first access (one read-back)."""

    def __init__(self, enc, h, traj, sd_txt_ft):
        self._enc, self._h, self._traj, self._sd_txt = enc, h, traj, sd_txt_ft

    def __getitem__(self, key):
        if key == "image_embeds":
            return self._enc.narrowed()
        if key == "last_hidden_state":
            B, _, d = self._h.shape
            n = self._traj.host()[0][-1]
            return self._h.reshape(-1)[:B * n * d].view(B, n, d)
        if key == "sd_img_ft":
            return self._enc.sd_ft
        if key == "sd_txt_ft":
            return self._sd_txt
        raise KeyError(key)


def load_state_dict_from_checkpoint(model: BLIP_NLVR, state_dict):
    """models/blip_nlvr.py:143-158 over a state dict (see madtp_b200.checkpoint.load_nlvr_checkpoint)."""
    from .checkpoint import load_nlvr_checkpoint
    return load_nlvr_checkpoint(model, state_dict)[1]


def load_checkpoint(model, url_or_filename, client=None):
    """models/blip_nlvr.py:131-160 (local files / state dicts; the reference's URL and S3 branches are host I/O)."""
    from .checkpoint import load_nlvr_checkpoint
    return load_nlvr_checkpoint(model, url_or_filename)


def blip_nlvr(pretrained='', **kwargs):
    model = BLIP_NLVR(**kwargs)
    if pretrained:
        model, msg = load_checkpoint(model, pretrained)
        print("missing keys:")
        print(msg.missing_keys)
    return model
