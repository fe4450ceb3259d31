"""Seeded synthetic inputs for the VIMA policy path (TEST / BENCH INFRASTRUCTURE).

Shapes and value ranges follow SURVEY.md Appendix C / section 8(d): uint8 crops (a pixel > 2 always exists),
int64 bboxes in [0,128), ee in {0,1}, word ids in [0,32100), bool object masks whose first slot is valid.
Everything comes from `oracle.detgen` so the same bits are produced in the build container (golden minting)
and on the GPU box (parity tests, bench).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import detgen

VIEWS = ("front", "top")

MODEL_CFGS = {  # embed_dim, xf_n_layers, heads (head_dim is 32 everywhere)
    "2M": dict(embed_dim=256, xf_n_layers=1, sattn_n_heads=8, xattn_n_heads=8),
    "4M": dict(embed_dim=256, xf_n_layers=2, sattn_n_heads=8, xattn_n_heads=8),
    "9M": dict(embed_dim=320, xf_n_layers=3, sattn_n_heads=10, xattn_n_heads=10),
    "20M": dict(embed_dim=384, xf_n_layers=4, sattn_n_heads=12, xattn_n_heads=12),
    "43M": dict(embed_dim=512, xf_n_layers=5, sattn_n_heads=16, xattn_n_heads=16),
    "92M": dict(embed_dim=640, xf_n_layers=7, sattn_n_heads=20, xattn_n_heads=20),
    "200M": dict(embed_dim=768, xf_n_layers=11, sattn_n_heads=24, xattn_n_heads=24),
}


@dataclass
class Case:
    name: str
    model: str
    B: int
    T: int  # obs steps in the history (incl. the new one)
    n_slots: int  # object slots per view; Q = 2 * n_slots
    n_words: int  # words per prompt (max, when ragged)
    n_imgs: int  # image placeholders per prompt (max, when ragged)
    ragged: bool = False
    seed: int = 0

    @property
    def Q(self):
        return 2 * self.n_slots

    @property
    def Lp(self):
        return self.n_words + self.n_imgs * self.Q

    @property
    def L(self):
        return self.T * self.Q + self.T - 1


CASES = {
    # BASELINE.json configs[0]: 8 text + 2 object tokens, 1 obs step (plus the T=2 variant)
    "cfg1": Case("cfg1", "2M", B=1, T=1, n_slots=3, n_words=8, n_imgs=1, seed=11),
    "cfg1_t2": Case("cfg1_t2", "2M", B=1, T=2, n_slots=3, n_words=8, n_imgs=1, seed=12),
    # ragged prompts + random object masks
    "ragged_4M": Case("ragged_4M", "4M", B=3, T=3, n_slots=2, n_words=9, n_imgs=2, ragged=True, seed=13),
    # BASELINE.json configs[1] shape at a CPU-sized batch
    "cfg2_small": Case("cfg2_small", "20M", B=4, T=4, n_slots=8, n_words=48, n_imgs=1, ragged=True, seed=14),
    # BASELINE.json configs[2] shape at a CPU-sized batch (Lp=256, L=263)
    "cfg3_small": Case("cfg3_small", "200M", B=2, T=8, n_slots=16, n_words=224, n_imgs=1, ragged=True, seed=15),
    # full-size configs (bench / property tests; never run on the CPU oracle in full)
    "cfg2": Case("cfg2", "20M", B=64, T=4, n_slots=8, n_words=48, n_imgs=1, seed=16),
    "cfg3": Case("cfg3", "200M", B=256, T=8, n_slots=16, n_words=224, n_imgs=1, seed=17),
}
# cfg1's prompt: 8 words + 1 placeholder with 1 slot per view -> Lp = 10; its obs uses 3 slots per view (Q=6)
CFG1_PROMPT_SLOTS = 1


def _objects(tag: str, lead: tuple, n_slots: int, seed: int, random_masks: bool) -> Dict[str, Dict[str, torch.Tensor]]:
    out = {"cropped_img": {}, "bbox": {}, "mask": {}}
    for v in VIEWS:
        img = detgen.randint(f"{tag}.img.{v}", lead + (n_slots, 3, 32, 32), 0, 256, seed).to(torch.uint8)
        img.view(-1)[0] = 200  # preprocess.py:28 requires max() > 2
        out["cropped_img"][v] = img
        out["bbox"][v] = detgen.randint(f"{tag}.bbox.{v}", lead + (n_slots, 4), 0, 128, seed)
        if random_masks:
            m = detgen.randint(f"{tag}.mask.{v}", lead + (n_slots,), 0, 4, seed) > 0  # ~75 % valid
            m[..., 0] = True
        else:
            m = torch.ones(lead + (n_slots,), dtype=torch.bool)
        out["mask"][v] = m
    return out


def make_prompt(case: Case):
    """-> (token_types, word_batch, image_batch) as VIMAPolicy.forward_prompt_assembly takes them."""
    prompt_slots = CFG1_PROMPT_SLOTS if case.name.startswith("cfg1") else case.n_slots
    token_types: List[List[int]] = []
    n_words_total = n_imgs_total = 0
    for b in range(case.B):
        nw, ni = case.n_words, case.n_imgs
        if case.ragged and b > 0:  # episode 0 keeps the full length so Lp is the nominal one
            nw = int(detgen.randint(f"{case.name}.nw.{b}", (1,), max(1, case.n_words // 2), case.n_words + 1, case.seed))
            ni = int(detgen.randint(f"{case.name}.ni.{b}", (1,), 0 if case.n_imgs > 1 else case.n_imgs, case.n_imgs + 1, case.seed))
        tt = [0] * nw
        # spread the placeholders through the sentence
        for j in range(ni):
            tt.insert(min(len(tt), 2 + j * max(1, nw // max(ni, 1))), 1)
        token_types.append(tt)
        n_words_total += nw
        n_imgs_total += ni
    word_batch = detgen.randint(f"{case.name}.words", (n_words_total,), 0, 32100, case.seed)
    image_batch = _objects(f"{case.name}.prompt", (max(n_imgs_total, 1),), prompt_slots, case.seed, case.ragged)
    if n_imgs_total == 0:
        raise ValueError("synthetic prompts need at least one image placeholder in the batch")
    image_batch = {k: {v: t[:n_imgs_total] for v, t in d.items()} for k, d in image_batch.items()}
    return token_types, word_batch, image_batch


def make_obs(case: Case, T: Optional[int] = None, tag: str = "obs"):
    """-> {"ee": (T,B) i64, "objects": {cropped_img/bbox/mask: {view: (T,B,n_slots,...)}}}."""
    T = case.T if T is None else T
    objects = _objects(f"{case.name}.{tag}", (T, case.B), case.n_slots, case.seed, case.ragged)
    ee = detgen.randint(f"{case.name}.{tag}.ee", (T, case.B), 0, 2, case.seed)
    return {"ee": ee, "objects": objects}


def slice_obs(obs, t0: int, t1: int):
    return {
        "ee": obs["ee"][t0:t1],
        "objects": {k: {v: x[t0:t1] for v, x in d.items()} for k, d in obs["objects"].items()},
    }


def make_actions(case: Case, T: int):
    """T-1 past discrete actions (int64) per episode, as example.py feeds back via forward_action_token."""
    out = {}
    for k, dims in (("pose0_position", [50, 100]), ("pose0_rotation", [50] * 4), ("pose1_position", [50, 100]), ("pose1_rotation", [50] * 4)):
        cols = [detgen.randint(f"{case.name}.act.{k}.{j}", (max(T - 1, 0), case.B, 1), 0, n, case.seed) for j, n in enumerate(dims)]
        out[k] = torch.cat(cols, dim=-1)
    return out


# ------------------------------------------------------------------------------------------------
# VIMA-Gato baseline inputs: whole 64x128 RGB views instead of object crops
# ------------------------------------------------------------------------------------------------
GATO_CFGS = {"gato_tiny": dict(embed_dim=256, n_layer=2, n_head=8), "gato_200M": dict(embed_dim=768, n_layer=22, n_head=24)}
GATO_CASES = {
    "gato_small": Case("gato_small", "gato_tiny", B=2, T=2, n_slots=8, n_words=6, n_imgs=1, ragged=True, seed=21),
    "gato_cfg5": Case("gato_cfg5", "gato_200M", B=256, T=8, n_slots=8, n_words=240, n_imgs=1, seed=22),
}


# VIMA-GPT baseline (one token per observation): same decoder family / image shapes as the Gato cases
GPT_CASES = {"gpt_small": Case("gpt_small", "gato_tiny", B=2, T=3, n_slots=8, n_words=6, n_imgs=2, ragged=True, seed=31)}


# VIMA-Flamingo baseline: XAttnGPT decoder + Perceiver-resampled image tokens (4 per image)
FLAMINGO_CFGS = {"flamingo_tiny": dict(embed_dim=256, dt_n_layers=2, dt_n_heads=8, xattn_n_heads=8)}
FLAMINGO_CASES = {"flamingo_small": Case("flamingo_small", "flamingo_tiny", B=2, T=2, n_slots=8, n_words=6, n_imgs=2, ragged=True, seed=41)}


def _rgb(tag: str, lead: tuple, seed: int):
    out = {}
    for v in VIEWS:
        img = detgen.randint(f"{tag}.rgb.{v}", lead + (3, 64, 128), 0, 256, seed).to(torch.uint8)
        img.view(-1)[0] = 200
        out[v] = img
    return out


def make_gato_prompt(case: Case):
    token_types, n_words_total, n_imgs_total = [], 0, 0
    for b in range(case.B):
        nw = case.n_words
        if case.ragged and b > 0:
            nw = int(detgen.randint(f"{case.name}.nw.{b}", (1,), max(1, case.n_words // 2), case.n_words + 1, case.seed))
        tt = [0] * nw
        for j in range(case.n_imgs):
            tt.insert(min(len(tt), 2 + j), 1)
        token_types.append(tt)
        n_words_total += nw
        n_imgs_total += case.n_imgs
    word_batch = detgen.randint(f"{case.name}.words", (n_words_total,), 0, 32100, case.seed)
    return token_types, word_batch, {"rgb": _rgb(f"{case.name}.prompt", (n_imgs_total,), case.seed)}


def make_gato_obs(case: Case, T: Optional[int] = None, tag: str = "obs"):
    T = case.T if T is None else T
    return {"ee": detgen.randint(f"{case.name}.{tag}.ee", (T, case.B), 0, 2, case.seed), "rgb": _rgb(f"{case.name}.{tag}", (T, case.B), case.seed)}
