"""GPU versions of the two helpers that sit immediately before the policy path in the reference's inference script
(/root/reference/scripts/example.py:243-375 `prepare_prompt`, :377-473 `prepare_obs`).

Same arguments, same nested outputs (`DataDict`s of uint8 crops, int64 boxes, bool masks, leading dims as the script
builds them) -- but the per-object numpy / cv2 loop (mask -> bounding box -> crop -> zero-pad to square ->
`cv2.resize(..., (32, 32), INTER_AREA)` -> visible objects first) runs as two CUDA kernels per view
(`vima_object_stats`, `vima_crop_resize`), bit-exact with OpenCV's 8-bit INTER_AREA paths, and the results stay on the
device for `forward_obs_token` / `forward_prompt_assembly`.  There is no CPU path here.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import engine as eng
from .utils import DataDict, any_to_datadict, get_batch_size

__all__ = ["crop_objects", "prepare_obs", "prepare_prompt"]


def _dev(x, device, dtype=None) -> torch.Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def crop_objects(rgb, segm, obj_ids, *, device="cuda"):
    """rgb (N,3,H,W) uint8, segm (N,H,W) integer, obj_ids (n,) or (N,n) -> cropped_img (N,n,3,32,32) uint8, bbox (N,n,4) int64,
    mask (N,n) bool, n_valid (N,) int32 -- one (step, view) image per leading index, visible objects first (example.py:407-456)."""
    rgb = _dev(rgb, device)
    segm = _dev(segm, device)
    if rgb.dtype != torch.uint8 or rgb.dim() != 4 or rgb.shape[1] != 3:
        raise ValueError(f"rgb must be uint8 (N,3,H,W), got {rgb.dtype} {tuple(rgb.shape)}")
    N, _, H, W = rgb.shape
    if segm.shape != (N, H, W) or segm.dtype not in (torch.uint8, torch.int8, torch.int32, torch.int64):
        raise ValueError(f"segm must be an integer (N,H,W) map matching rgb, got {segm.dtype} {tuple(segm.shape)}")
    if segm.dtype == torch.int8:
        segm = segm.to(torch.int32)
    ids = _dev(np.asarray(obj_ids) if not isinstance(obj_ids, torch.Tensor) else obj_ids, device, torch.int64)
    per_image = ids.dim() == 2
    n = ids.shape[-1]
    if per_image and ids.shape[0] != N:
        raise ValueError("per-image object ids must have one row per image")
    ctx = eng.ctx_for(rgb)
    stats = torch.empty((N, n, 5), dtype=torch.int32, device=rgb.device)
    crops = torch.empty((N, n, 3, 32, 32), dtype=torch.uint8, device=rgb.device)
    bbox = torch.empty((N, n, 4), dtype=torch.int64, device=rgb.device)
    mask = torch.empty((N, n), dtype=torch.uint8, device=rgb.device)
    n_valid = torch.zeros((N,), dtype=torch.int32, device=rgb.device)  # zeros: an empty object list launches nothing
    if N == 0 or n == 0:
        return crops, bbox, mask.view(torch.bool), n_valid
    ctx.object_stats(segm, N, H, W, ids, n, per_image, stats)
    ctx.crop_resize(rgb, N, H, W, stats, n, crops, bbox, mask, n_valid)
    return crops, bbox, mask.view(torch.bool), n_valid


def prepare_obs(*, obs: dict, rgb_dict: Optional[dict] = None, meta: dict, device="cuda") -> DataDict:
    """example.py:377-473.  obs: {"rgb": {view: (L,3,H,W) u8}, "segm": {view: (L,H,W)}, "ee": (L,)} (numpy or torch; "rgb" /
    "segm" are popped as the reference does), meta: {"n_objects", "obj_id_to_info"} -> DataDict with leading dims (L, 1)."""
    assert not (rgb_dict is not None and "rgb" in obs)
    rgb_dict = rgb_dict or obs.pop("rgb")
    segm_dict = obs.pop("segm")
    views = sorted(rgb_dict.keys())
    assert meta["n_objects"] == len(meta["obj_id_to_info"])
    objects = list(meta["obj_id_to_info"].keys())
    L_obs = get_batch_size(obs)
    out = {"ee": _dev(obs["ee"], device).reshape(L_obs, 1), "objects": {"cropped_img": {}, "bbox": {}, "mask": {}}}
    for view in views:
        crops, bbox, mask, _ = crop_objects(rgb_dict[view], segm_dict[view], objects, device=device)
        assert crops.shape[0] == L_obs
        out["objects"]["cropped_img"][view] = crops.unsqueeze(1)
        out["objects"]["bbox"][view] = bbox.unsqueeze(1)
        out["objects"]["mask"][view] = mask.unsqueeze(1)
    return any_to_datadict(out)


def prepare_prompt(*, prompt: Optional[str] = None, prompt_assets: dict, views: Sequence[str], tokenizer=None,
                   prompt_ids: Optional[List[int]] = None, prompt_tokens: Optional[List[str]] = None, device="cuda"):
    """example.py:243-375 -> (raw_prompt_token_type, word_batch (n_words,) int64, image_batch DataDict).  Tokenisation is outside
    this package: pass the reference's `tokenizers.Tokenizer` as `tokenizer`, or its output as `prompt_ids` / `prompt_tokens`.
    A token of the form "{name}" is a placeholder filled from `prompt_assets[name]` ({"rgb": {view: (3,H,W)}, "segm":
    {view: (H,W), "obj_info": ...}, "placeholder_type": "object" | "scene"})."""
    views = sorted(views)
    if prompt_ids is None:
        if tokenizer is None or prompt is None:
            raise ValueError("prepare_prompt needs either (prompt, tokenizer) or (prompt_ids, prompt_tokens)")
        enc = tokenizer.encode(prompt, add_special_tokens=True)
        prompt_ids, prompt_tokens = enc.ids, enc.tokens
    is_ph = lambda t: t.startswith("{") and t.endswith("}")
    assert set(prompt_assets.keys()) == set(t[1:-1] for t in prompt_tokens if is_ph(t))
    token_type, words, per_token = [], [], []
    for id_, token in zip(prompt_ids, prompt_tokens):
        if not is_ph(token):
            assert "{" not in token and "}" not in token
            token_type.append(0)
            words.append(int(id_))
            continue
        asset = prompt_assets[token[1:-1]]
        obj_info = asset["segm"]["obj_info"]
        kind = asset["placeholder_type"]
        objects = [obj_info["obj_id"]] if kind == "object" else [info["obj_id"] for info in obj_info]
        token_type.append(1)
        entry = {}
        for view in views:
            rgb, segm = asset["rgb"][view], asset["segm"][view]
            crops, bbox, mask, n_valid = crop_objects(_dev(rgb, device)[None], _dev(segm, device)[None], objects, device=device)
            entry[view] = (crops[0], bbox[0], mask[0], n_valid)
        per_token.append(entry)
    # objects that are not visible are dropped from a prompt token (example.py:281-282); every token is then padded to the
    # prompt's per-view maximum (example.py:304-352).  The kernels already put the visible ones first and zero the rest.
    counts = {v: [int(e[v][3].item()) for e in per_token] for v in views}  # one host read per asset view, once per episode
    image_batch = {"cropped_img": {}, "bbox": {}, "mask": {}}
    for v in views:
        mx = max(counts[v]) if per_token else 0
        fit = lambda t: t[:mx] if t.shape[0] >= mx else torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))], 0)
        image_batch["cropped_img"][v] = torch.stack([fit(e[v][0]) for e in per_token], 0) if per_token else torch.zeros((0, 0, 3, 32, 32), dtype=torch.uint8, device=device)
        image_batch["bbox"][v] = torch.stack([fit(e[v][1]) for e in per_token], 0) if per_token else torch.zeros((0, 0, 4), dtype=torch.int64, device=device)
        image_batch["mask"][v] = torch.stack([fit(e[v][2]) for e in per_token], 0) if per_token else torch.zeros((0, 0), dtype=torch.bool, device=device)
    assert len(token_type) == len(words) + len(per_token)
    word_batch = torch.tensor(words, dtype=torch.int64, device=device)
    return [token_type], word_batch, any_to_datadict(image_batch)
