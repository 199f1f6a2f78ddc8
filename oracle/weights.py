"""Seeded synthetic weights / inputs shared by the oracle, the tests and bench.py (they live in the package so that
bench.py's product arm does not import anything from oracle/)."""
from madtp_b200.synthetic import *  # noqa: F401,F403
from madtp_b200.synthetic import clip_state_dict, clip_inputs, med_text_state_dict, retrieval_state_dict, retrieval_inputs, block_inputs, block_state_dict, blip_nlvr_state_dict, nlvr_inputs, tensor_digest, vqa_state_dict, vqa_answer_candidates  # noqa: F401
