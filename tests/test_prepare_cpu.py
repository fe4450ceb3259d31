"""CPU: pin oracle/prepare_oracle.py (the restatement of scripts/example.py:243-473's object loop and of OpenCV's INTER_AREA
resize) to the fixtures minted with cv2 (tests/golden/make_prepare_golden.py) and, where cv2 is importable, to cv2 itself."""
import os

import numpy as np
import pytest

from oracle import prepare_oracle as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_resize_known_answers():
    kat = np.load(os.path.join(GOLD, "prepare_resize_kat.npz"))
    names = sorted(k for k in kat.files if k.startswith("in_"))
    assert len(names) >= 60
    for k in names:
        got = P.resize_area_u8(kat[k])
        assert np.array_equal(got, kat["out" + k[2:]]), k


def test_resize_every_size_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for n in range(1, 257):
        img = rng.integers(0, 256, (n, n, 3), dtype=np.uint8)
        if n % 5 == 0:
            img[: n // 2] = 255  # saturated half: rounding at the top of the range
        ref = cv2.resize(img, (32, 32), interpolation=cv2.INTER_AREA)
        assert np.array_equal(P.resize_area_u8(img), ref), n


def test_scene_fixtures():
    sc = np.load(os.path.join(GOLD, "prepare_scenes.npz"))
    seen_valid = seen_pad = 0
    for seed in range(6):
        crops, bbox, mask = P.prepare_obs_view(sc[f"rgb_{seed}"], sc[f"segm_{seed}"], sc[f"ids_{seed}"].tolist())
        assert np.array_equal(mask, sc[f"mask_{seed}"])
        assert np.array_equal(bbox, sc[f"bbox_{seed}"]) and bbox.dtype == np.int64
        assert np.array_equal(crops, sc[f"crops_{seed}"]) and crops.dtype == np.uint8
        seen_valid += int(mask.sum())
        seen_pad += int((~mask).sum())
        assert not mask[int(mask.sum()):].any() and mask[: int(mask.sum())].all()  # visible objects first
    assert seen_valid >= 12 and seen_pad >= 6


def test_bbox_and_padding_semantics():
    """Hand-checkable case of example.py:412-431: bbox = [int((xmin+xmax)/2), int((ymin+ymax)/2), ymax-ymin, xmax-xmin];
    the shorter side is zero-padded int(diff/2) before and the rest after."""
    rgb = np.zeros((3, 40, 50), np.uint8)
    rgb[:, 5:9, 10:42] = 200  # 4 rows x 32 columns of object 7
    segm = np.zeros((40, 50), np.int64)
    segm[5:9, 10:42] = 7
    segm[30, 30] = 9  # single pixel: dropped
    crops, bbox, mask = P.prepare_obs_view(rgb, segm, [9, 7, 3])
    assert mask.tolist() == [True, False, False]
    assert bbox[0].tolist() == [25, 6, 3, 31] and not bbox[1:].any()
    # 4x32 -> pad rows to 32 (14 before, 14 after), no resize needed
    assert (crops[0][:, 14:18] == 200).all() and not crops[0][:, :14].any() and not crops[0][:, 18:].any()
    out = P.prepare_obs({"front": rgb[None], "top": rgb[None]}, {"front": segm[None], "top": segm[None]}, np.array([1]), [9, 7, 3])
    assert out["objects"]["cropped_img"]["top"].shape == (1, 1, 3, 3, 32, 32) and out["ee"].shape == (1, 1)
    c, b = P.prompt_asset_objects(rgb, segm, [9, 7, 3])
    assert c.shape == (1, 3, 32, 32) and b.tolist() == [[25, 6, 3, 31]]
