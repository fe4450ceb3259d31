"""GPU: `vima_b200.prepare` (the step before the path, scripts/example.py:243-473) against the oracle restatement
(oracle/prepare_oracle.py, pinned to cv2) and the cv2-minted fixtures -- bit-exact: uint8 crops, int64 boxes, bool masks."""
import os

import numpy as np
import pytest
import torch

from oracle import prepare_oracle as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_resize_every_square_size_bit_exact():
    """One n x n object per image, n = 2..256: exercises the copy, enlarge (fixed point), integer-factor and float-area paths."""
    from vima_b200.prepare import crop_objects

    rng = np.random.default_rng(5)
    H = W = 256
    sizes = list(range(2, 257))
    rgb = rng.integers(0, 256, (len(sizes), 3, H, W), dtype=np.uint8)
    rgb[::7, :, :128] = 255
    segm = np.zeros((len(sizes), H, W), np.uint8)
    for i, n in enumerate(sizes):
        y0, x0 = int(rng.integers(0, H - n + 1)), int(rng.integers(0, W - n + 1))
        segm[i, y0:y0 + n, x0:x0 + n] = 3
    crops, bbox, mask, n_valid = crop_objects(rgb, segm, [3])
    crops, bbox = crops.cpu().numpy(), bbox.cpu().numpy()
    assert mask.all() and (n_valid == 1).all()
    for i, n in enumerate(sizes):
        ref = P.crop_object(rgb[i], segm[i], 3)
        assert bbox[i, 0].tolist() == ref[0], n
        assert np.array_equal(crops[i, 0], ref[1]), n


def test_scene_fixtures_bit_exact():
    from vima_b200.prepare import crop_objects

    sc = np.load(os.path.join(GOLD, "prepare_scenes.npz"))
    for seed in range(6):
        for cast in (np.uint8, np.int32, np.int64):
            crops, bbox, mask, n_valid = crop_objects(sc[f"rgb_{seed}"][None], sc[f"segm_{seed}"].astype(cast)[None], sc[f"ids_{seed}"])
            assert crops.dtype == torch.uint8 and bbox.dtype == torch.int64 and mask.dtype == torch.bool
            assert np.array_equal(mask[0].cpu().numpy(), sc[f"mask_{seed}"])
            assert np.array_equal(bbox[0].cpu().numpy(), sc[f"bbox_{seed}"])
            assert np.array_equal(crops[0].cpu().numpy(), sc[f"crops_{seed}"])
            assert int(n_valid[0]) == int(sc[f"mask_{seed}"].sum())


def _scene_batch(L, seed0):
    sc = [P.synthetic_scene(seed0 + l, 128, 256) for l in range(L)]
    return np.stack([s[0] for s in sc]), np.stack([s[1] for s in sc]), sc[0][2]


def test_prepare_obs_matches_oracle_and_feeds_the_policy():
    import vima_b200
    from vima_b200.prepare import prepare_obs
    from tests.policy_runner import build_policy

    L = 3
    rgb_f, segm_f, ids = _scene_batch(L, 100)
    rgb_t, segm_t, _ = _scene_batch(L, 200)
    ee = np.array([0, 1, 1], dtype=np.int64)
    meta = {"n_objects": len(ids), "obj_id_to_info": {i: {"obj_name": str(i)} for i in ids}}
    obs = {"rgb": {"front": rgb_f, "top": rgb_t}, "segm": {"front": segm_f, "top": segm_t}, "ee": ee}
    got = prepare_obs(obs=obs, meta=meta)
    assert "rgb" not in obs and "segm" not in obs  # popped, as the reference does
    ref = P.prepare_obs({"front": rgb_f, "top": rgb_t}, {"front": segm_f, "top": segm_t}, ee, ids)
    assert np.array_equal(got["ee"].cpu().numpy(), ref["ee"])
    for k in ("cropped_img", "bbox", "mask"):
        for v in ("front", "top"):
            g = got["objects"][k][v]
            assert g.is_cuda and tuple(g.shape) == ref["objects"][k][v].shape
            assert np.array_equal(g.cpu().numpy(), ref["objects"][k][v]), (k, v)
    # per-image object ids (a batch of different episodes) == per-image calls
    from vima_b200.prepare import crop_objects
    ids2 = np.stack([np.roll(np.asarray(ids), l) for l in range(L)])
    c2, b2, m2, _ = crop_objects(rgb_f, segm_f, ids2)
    for l in range(L):
        r = P.prepare_obs_view(rgb_f[l], segm_f[l], ids2[l].tolist())
        assert np.array_equal(c2[l].cpu().numpy(), r[0]) and np.array_equal(b2[l].cpu().numpy(), r[1]) and np.array_equal(m2[l].cpu().numpy(), r[2])
    # the DataDict plugs straight into the policy entry point that follows it
    vima_b200.set_precision("f16x3")
    pol = build_policy("2M")
    with torch.no_grad():
        tok, msk = pol.forward_obs_token(got)
    assert tok.shape == (L, 1, 2 * len(ids), pol.embed_dim) and msk.shape == (L, 1, 2 * len(ids))


class _Enc:
    def __init__(self, ids, tokens):
        self.ids, self.tokens = ids, tokens


class _Tok:
    def encode(self, prompt, add_special_tokens=True):
        toks = prompt.split() + ["</s>"]
        return _Enc([100 + i for i in range(len(toks))], toks)


def test_prepare_prompt_matches_reference_padding():
    from vima_b200.prepare import prepare_prompt

    views = ["top", "front"]
    assets = {}
    for j, (name, kind) in enumerate([("dragged_obj", "object"), ("scene", "scene"), ("base_obj", "object")]):
        rgb, segm = {}, {}
        for v in views:
            r, s, ids = P.synthetic_scene(300 + 10 * j + (v == "top"), 128, 256)
            rgb[v], segm[v] = r, s
        present = [i for i in ids if (segm["front"] == i).sum() >= 2 and (segm["top"] == i).sum() >= 2]
        segm["obj_info"] = {"obj_id": present[0]} if kind == "object" else [{"obj_id": i} for i in ids]
        assets[name] = {"rgb": rgb, "segm": segm, "placeholder_type": kind}
    prompt = "Put the {dragged_obj} into {base_obj} as in {scene} ."
    tt, words, img = prepare_prompt(prompt=prompt, prompt_assets=assets, views=views, tokenizer=_Tok())
    toks = prompt.split() + ["</s>"]
    assert tt == [[1 if t.startswith("{") else 0 for t in toks]]
    assert words.dtype == torch.int64 and words.tolist() == [100 + i for i, t in enumerate(toks) if not t.startswith("{")]
    order = [t[1:-1] for t in toks if t.startswith("{")]
    for v in sorted(views):
        per = []
        for name in order:
            a = assets[name]
            objs = [a["segm"]["obj_info"]["obj_id"]] if a["placeholder_type"] == "object" else [i["obj_id"] for i in a["segm"]["obj_info"]]
            per.append(P.prompt_asset_objects(a["rgb"][v], a["segm"][v], objs))
        mx = max(p[0].shape[0] for p in per)
        assert tuple(img["cropped_img"][v].shape) == (3, mx, 3, 32, 32) and mx >= 2
        for t, (c, b) in enumerate(per):
            n = c.shape[0]
            assert np.array_equal(img["cropped_img"][v][t, :n].cpu().numpy(), c) and not img["cropped_img"][v][t, n:].any()
            assert np.array_equal(img["bbox"][v][t, :n].cpu().numpy(), b) and not img["bbox"][v][t, n:].any()
            assert img["mask"][v][t].tolist() == [True] * n + [False] * (mx - n)


def test_crop_objects_argument_errors():
    from vima_b200.prepare import crop_objects

    rgb = np.zeros((1, 3, 8, 8), np.uint8)
    with pytest.raises(ValueError):
        crop_objects(rgb.astype(np.float32), np.zeros((1, 8, 8), np.uint8), [1])
    with pytest.raises(ValueError):
        crop_objects(rgb, np.zeros((1, 8, 9), np.uint8), [1])
    with pytest.raises(RuntimeError):
        crop_objects(rgb, np.zeros((1, 8, 8), np.uint8), list(range(65)))
    c, b, m, n = crop_objects(rgb, np.zeros((1, 8, 8), np.uint8), [5, 6])  # nothing visible
    assert not m.any() and not c.any() and not b.any() and int(n[0]) == 0
