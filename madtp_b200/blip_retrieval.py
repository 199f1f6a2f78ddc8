"""Mirror of the evaluation-path surface of the reference's models/blip_retrieval.py (BLIP_Retrieval) and
models/blip_vqa.py (BLIP_VQA): the attributes and encoder calls that compress_retrieval_dtp.py:85-207 and
blip_vqa.py:58-125 use -- `space_dict`, `visual_encoder`, `text_encoder` (models/med.py), `vision_proj`,
`text_proj`, `itm_head`, `tokenizer`. The training forward (ITC/ITM losses, momentum encoders, queues,
blip_retrieval.py:99-282) and the answer decoder (blip_vqa.py:156-203) are out of scope;
their parameters are simply absent, so released checkpoints load with strict=False.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import _lib as L
from . import functional as Fn
from .blip_nlvr import ENC_TOKEN_ID, create_vit
from .configuration import BertConfig
from .graphs import GraphedForward as _GraphedForward
from .med import BertLMHeadModel, BertModel


def _med_config(med_config, vision_width, evaluate):
    cfg = BertConfig.from_json_file(med_config) if (isinstance(med_config, str) and os.path.isfile(med_config)) \
        else (med_config if isinstance(med_config, BertConfig) else BertConfig())
    cfg.encoder_width = vision_width
    cfg.evaluate = evaluate
    return cfg


class BLIP_Retrieval(_GraphedForward, nn.Module):
    def __init__(self, med_config='configs/med_config.json', image_size=384, vit='base', vit_grad_ckpt=False,
                 vit_ckpt_layer=0, embed_dim=256, queue_size=57600, momentum=0.995, negative_all_rank=False,
                 evaluate=False, config=None, tokenizer=None):
        super().__init__()
        self.sd_num, self.sd_dim = (100, 768) if config is None else (config['sd_num'], config['sd_dim'])
        self.space_dict = nn.Parameter(torch.randn(self.sd_num, self.sd_dim))
        self.layers = 12
        self.visual_encoder, vision_width = create_vit(vit, image_size, vit_grad_ckpt, vit_ckpt_layer, 0,
                                                       evaluate=evaluate, sd_dim=self.sd_dim)
        self.tokenizer = tokenizer
        cfg = _med_config(med_config, vision_width, evaluate)
        self.text_encoder = BertModel(config=cfg, add_pooling_layer=False, sd_dim=self.sd_dim)
        text_width = cfg.hidden_size
        self.vision_proj = nn.Linear(vision_width, embed_dim)
        self.text_proj = nn.Linear(text_width, embed_dim)
        self.itm_head = nn.Linear(text_width, 2)
        self.temp = nn.Parameter(0.07 * torch.ones([]))
        self._cache = Fn.WeightCache()

    def _lin(self, name):
        m = getattr(self, name)
        return self._cache.get(name, [m.weight, m.bias], lambda: Fn.PreparedLinear(m.weight, m.bias, f32=True))

    @torch.no_grad()
    def encode_text(self, input_ids, attention_mask, temperature=0):
        """compress_retrieval_dtp.py:104-105: text-only encoder pass -> L2-normalised text embedding [B, embed_dim]."""
        out, _ = self.text_encoder(input_ids, attention_mask=attention_mask, mode='text', space_dict=self.space_dict,
                                   temperature=temperature)
        e = Fn.linear_f32(out.last_hidden_state[:, 0, :].contiguous(), self._lin("text_proj"))
        return torch.nn.functional.normalize(e, dim=-1)

    @torch.no_grad()
    def encode_image(self, image, temperature=0):
        """compress_retrieval_dtp.py:120-122 -> (image_feat [B, N', d], L2-normalised image embedding)."""
        feat, _ = self.visual_encoder(image, space_dict=self.space_dict, temperature=temperature)
        e = Fn.linear_f32(feat[:, 0, :].contiguous(), self._lin("vision_proj"))
        return feat, torch.nn.functional.normalize(e, dim=-1)

    @torch.no_grad()
    def itm_score(self, input_ids, attention_mask, image_feats, temperature=0):
        """compress_retrieval_dtp.py:166-176: multimodal pass + itm_head -> logits [B, 2]. input_ids[:,0] is
        replaced by [ENC] as the driver does (:112)."""
        ids = input_ids.clone()
        ids[:, 0] = getattr(self.tokenizer, "enc_token_id", ENC_TOKEN_ID)
        out, _ = self.text_encoder(ids, attention_mask=attention_mask, encoder_hidden_states=image_feats,
                                   encoder_attention_mask=None, return_dict=True, space_dict=self.space_dict,
                                   temperature=temperature)
        return Fn.linear_f32(out.last_hidden_state[:, 0, :].contiguous(), self._lin("itm_head"))

    @torch.no_grad()
    def itm_rerank(self, image_feat, input_ids, attention_mask, temperature=0):
        """ITM rerank of ONE image against k_test candidate captions (compress_retrieval_dtp.py:166-177).
        image_feat [N', d] or [1, N', d]; input_ids / attention_mask [k_test, L]. The image's cross-attention K/V
        projections are computed once and broadcast over the candidates (the reference repeats the image k_test times
        and re-projects it for every caption). Returns the ITM logits [k_test, 2]."""
        if image_feat.dim() == 2:
            image_feat = image_feat.unsqueeze(0)
        feats = image_feat.expand(input_ids.shape[0], -1, -1)        # zero batch stride, nothing is copied
        return self.itm_score(input_ids, attention_mask, feats, temperature)

    @torch.no_grad()
    def itm_rerank_t2i(self, image_feats, input_ids_row, attention_mask_row, temperature=0, pad_to=None):
        """ITM rerank of ONE caption against k_test candidate images (compress_retrieval_dtp.py:186-200). The candidates
        come from different evaluation batches, so their pruned lengths differ: `image_feats` is a list of [N_i, d]
        tensors. The reference pads them with CLS copies to the longest image of the whole set (:142-154, `pad_to`) and
        repeats the caption k_test times; here the images stay packed (nlvr_encoder.RaggedImageFeatures: `cu_seqlens`
        layout, the CLS padding evaluated in closed form) and only the caption is repeated. Returns ITM logits [k_test, 2]."""
        from .nlvr_encoder import RaggedImageFeatures
        k = len(image_feats)
        enc = RaggedImageFeatures(image_feats, pad_to)
        ids = input_ids_row.reshape(1, -1).repeat(k, 1)
        ids[:, 0] = getattr(self.tokenizer, "enc_token_id", ENC_TOKEN_ID)
        mask = attention_mask_row.reshape(1, -1).repeat(k, 1)
        out, _ = self.text_encoder(ids, attention_mask=mask, encoder_hidden_states=enc, encoder_attention_mask=None,
                                   return_dict=True, space_dict=self.space_dict, temperature=temperature)
        return Fn.linear_f32(out.last_hidden_state[:, 0, :].contiguous(), self._lin("itm_head"))

    @torch.no_grad()
    def forward(self, image, caption, alpha=0.0, idx=None, temperature=0, train=True):
        """Evaluation-path forward (BASELINE config 3): image encoder + text-only encoder + one multimodal ITM pass
        over the matched pairs. `caption` is pre-tokenised (input_ids, attention_mask) or text for `tokenizer`."""
        if train:
            raise NotImplementedError("madtp_b200 implements the evaluation path only (train=False)")
        if hasattr(caption, "input_ids"):
            ids, mask = caption.input_ids, caption.attention_mask
        elif isinstance(caption, (tuple, list)) and torch.is_tensor(caption[0]):
            ids, mask = caption
        else:
            t = self.tokenizer(caption, padding='max_length', truncation=True, max_length=35,
                               return_tensors="pt").to(image.device)
            ids, mask = t.input_ids, t.attention_mask
        if self._device_path(temperature, image, ids, mask) and ids.shape[1] <= 64:
            t = float(temperature)
            out = self._run_device("retrieval", lambda im, i, m: self._forward_device(im, i, m, t), [image, ids, mask], t)
            return out[0], out[1]
        image_feat, image_embed = self.encode_image(image, temperature)
        text_embed = self.encode_text(ids, mask, temperature)
        itm = self.itm_score(ids, mask, image_feat, temperature)
        return image_embed @ text_embed.t(), itm

    def _forward_device(self, image, ids, mask, temperature):
        """The same three passes with device-resident lengths from the first ViT layer to the logits: nothing is read
        back and every launch argument is data-independent (compress_retrieval_dtp.py:104-122,166-176). Returns
        ((similarity [B, B], ITM logits [B, 2]), [image, text, multimodal trajectories])."""
        B = image.shape[0]
        enc = self.visual_encoder.forward_device(image, self.space_dict, temperature, pack_groups=1, keep_f32=True)
        img_cls = L.take_token(enc.y, B, enc.cap, 0, n_dev=enc.n_dev)
        image_embed = torch.nn.functional.normalize(Fn.linear_f32(img_cls, self._lin("vision_proj")), dim=-1)
        h, _, traj_t, l_dev = self.text_encoder.forward_device(ids, mask, None, self.space_dict, temperature, mode='text')
        Lcap, d = h.shape[1], h.shape[2]
        txt_cls = L.take_token(h.view(B * Lcap, d), B, Lcap, 0, n_dev=l_dev)
        text_embed = torch.nn.functional.normalize(Fn.linear_f32(txt_cls, self._lin("text_proj")), dim=-1)
        ids2 = ids.clone()
        ids2[:, 0] = getattr(self.tokenizer, "enc_token_id", ENC_TOKEN_ID)
        h, _, traj_m, l_dev = self.text_encoder.forward_device(ids2, mask, enc, self.space_dict, temperature)
        itm = Fn.linear_f32(L.take_token(h.view(B * Lcap, d), B, Lcap, 0, n_dev=l_dev), self._lin("itm_head"))
        return (image_embed @ text_embed.t(), itm), [enc.traj, traj_t, traj_m]


class BLIP_VQA(_GraphedForward, nn.Module):
    def __init__(self, med_config='configs/med_config.json', image_size=480, vit='base', vit_grad_ckpt=False,
                 vit_ckpt_layer=0, evaluate=False, config=None, tokenizer=None):
        super().__init__()
        self.sd_num, self.sd_dim = (100, 768) if config is None else (config['sd_num'], config['sd_dim'])
        self.space_dict = nn.Parameter(torch.randn(self.sd_num, self.sd_dim))
        self.visual_encoder, vision_width = create_vit(vit, image_size, vit_grad_ckpt, vit_ckpt_layer,
                                                       drop_path_rate=0.1, evaluate=evaluate, sd_dim=self.sd_dim)
        self.tokenizer = tokenizer
        cfg = _med_config(med_config, vision_width, evaluate)
        self.text_encoder = BertModel(config=cfg, add_pooling_layer=False, sd_dim=self.sd_dim)
        self.text_decoder = BertLMHeadModel(config=_med_config(med_config, vision_width, evaluate), sd_dim=self.sd_dim)

    @torch.no_grad()
    def encode_question(self, image, input_ids, attention_mask, temperature=0):
        """blip_vqa.py:60,119-125: pruned image encoder, then the question encoder cross-attending to the pruned image
        tokens. Returns (question_states [B, L', d], pruned additive question mask is internal, image_embeds)."""
        image_embeds, _ = self.visual_encoder(image, space_dict=self.space_dict, temperature=temperature)
        ids = input_ids.clone()
        ids[:, 0] = getattr(self.tokenizer, "enc_token_id", ENC_TOKEN_ID)
        out, _ = self.text_encoder(ids, attention_mask=attention_mask, encoder_hidden_states=image_embeds,
                                   encoder_attention_mask=None, return_dict=True, space_dict=self.space_dict,
                                   temperature=temperature)
        return out.last_hidden_state, image_embeds

    @torch.no_grad()
    def encode_question_packed(self, image, input_ids, attention_mask, temperature):
        """The two encoders of blip_vqa.py:60,119-125 with device-resident lengths end to end (no read-back; a CUDA
        graph replay when enable_cuda_graphs(True)). Returns (question states as a capacity-sized buffer [B, L, d] of
        packed sequences, device scalar with their length, the [ENC] rows [B, d])."""
        if not self._device_path(temperature, image, input_ids, attention_mask) or input_ids.shape[1] > 64:
            raise RuntimeError("madtp_b200: encode_question_packed is the pruned device-length path (temperature > 0, "
                               "question length <= 64)")
        t = float(temperature)
        return self._run_device("vqa", lambda im, i, m: self._encode_question_device(im, i, m, t),
                                [image, input_ids, attention_mask], t)

    def _encode_question_device(self, image, ids, mask, temperature):
        B = image.shape[0]
        enc = self.visual_encoder.forward_device(image, self.space_dict, temperature, pack_groups=1)
        ids = ids.clone()
        ids[:, 0] = getattr(self.tokenizer, "enc_token_id", ENC_TOKEN_ID)
        h, _, traj, l_dev = self.text_encoder.forward_device(ids, mask, enc, self.space_dict, temperature)
        Lcap, d = h.shape[1], h.shape[2]
        cls = L.take_token(h.view(B * Lcap, d), B, Lcap, 0, n_dev=l_dev)
        return (h, l_dev, cls), [enc.traj, traj]

    @torch.no_grad()
    def rank_answer(self, question_states, question_atts, answer_ids, answer_atts, k):
        """models/blip_vqa.py:156-203: the first-token probabilities pick k candidate answers per question, the decoder
        scores each of them by its label-smoothed log-likelihood, the best one wins. Returns max_ids [num_ques]
        (indices into the answer list); the intermediate results are kept in `self.last_rank`."""
        num_ques = question_states.size(0)
        pad_id = getattr(self.tokenizer, "pad_token_id", 0)
        start_ids = answer_ids[0, 0].repeat(num_ques, 1)                                      # bos token
        start = self.text_decoder(start_ids, encoder_hidden_states=question_states,
                                  encoder_attention_mask=question_atts, return_dict=True, reduction='none')
        logits = start.logits[:, 0, :]                                                        # first token's logits
        _, lse = L.lm_nll(logits)
        answer_first_token = answer_ids[:, 1]
        prob_first_token = torch.exp(logits.index_select(1, answer_first_token) - lse[:, None])
        topk_probs, topk_ids = prob_first_token.topk(k, dim=1)
        input_ids = torch.cat([answer_ids.index_select(0, t) for t in topk_ids], dim=0)
        input_atts = torch.cat([answer_atts.index_select(0, t) for t in topk_ids], dim=0)
        targets_ids = input_ids.masked_fill(input_ids == pad_id, -100)
        # repeat the encoder's output for the k candidates of every question (tile(), :214-220)
        q_states = question_states.repeat_interleave(k, dim=0)
        q_atts = None if question_atts is None else question_atts.repeat_interleave(k, dim=0)
        out = self.text_decoder(input_ids, attention_mask=input_atts, encoder_hidden_states=q_states,
                                encoder_attention_mask=q_atts, labels=targets_ids, return_dict=True, reduction='none')
        log_probs_sum = (-out.loss).view(num_ques, k)
        max_topk_ids = log_probs_sum.argmax(dim=1)
        max_ids = topk_ids[max_topk_ids >= 0, max_topk_ids]
        self.last_rank = {"prob_first_token": prob_first_token, "topk_ids": topk_ids, "log_probs_sum": log_probs_sum}
        return max_ids

    @torch.no_grad()
    def forward(self, image, question, answer=None, temperature=0, train=True, n=None, weights=None,
                inference='rank', k_test=128):
        """Evaluation path of models/blip_vqa.py:58-154 with inference='rank'. `question` and `answer` are tokenised:
        objects with .input_ids / .attention_mask (the HuggingFace tokenizer stays on the host, outside the hot path)."""
        if train:
            raise NotImplementedError("madtp_b200: evaluation only (train=False)")
        if inference != 'rank':
            raise NotImplementedError("madtp_b200: inference='generate' (beam search) is out of scope")
        q_states, _ = self.encode_question(image, question.input_ids, question.attention_mask, temperature)
        return self.rank_answer(q_states, question.attention_mask, answer.input_ids, answer.attention_mask, k_test)


def load_checkpoint(model, url_or_filename, client=None):
    """models/blip.py:254-278 (see madtp_b200.checkpoint.load_blip_checkpoint)."""
    from .checkpoint import load_blip_checkpoint
    return load_blip_checkpoint(model, url_or_filename)


def blip_retrieval(pretrained='', **kwargs):
    """models/blip_retrieval.py:285-291."""
    model = BLIP_Retrieval(**kwargs)
    if pretrained:
        model, msg = load_checkpoint(model, pretrained)
        print("missing keys:")
        print(msg.missing_keys)
    return model


def blip_vqa(pretrained='', **kwargs):
    """models/blip_vqa.py:206-211."""
    model = BLIP_VQA(**kwargs)
    if pretrained:
        model, msg = load_checkpoint(model, pretrained)
    return model
