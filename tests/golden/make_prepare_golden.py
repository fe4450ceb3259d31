"""Mint tests/golden/prepare_*.npz with the real third-party dependency (cv2.resize INTER_AREA, OpenCV 4.13.0 in the build
container): synthetic scenes (rgb, segmentation) -> what scripts/example.py:401-456 computes per view, with the resize done
by cv2 itself.  Run from the repo root:  python tests/golden/make_prepare_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import prepare_oracle as P  # noqa: E402


def cv2_resize(img):
    import cv2

    return cv2.resize(img, (32, 32), interpolation=cv2.INTER_AREA)


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    # (1) resize known-answers: every regime, incl. flat / gradient images that sit on rounding ties
    rng = np.random.default_rng(7)
    sizes = [1, 2, 3, 5, 11, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 96, 100, 127, 128, 129, 160, 200, 255, 256]
    kat = {}
    for n in sizes:
        imgs = [rng.integers(0, 256, (n, n, 3), dtype=np.uint8), np.full((n, n, 3), 255, np.uint8),
                (np.add.outer(np.arange(n), np.arange(n))[..., None] * np.array([1, 2, 3]) % 256).astype(np.uint8)]
        for j, im in enumerate(imgs):
            kat[f"in_{n}_{j}"] = im
            kat[f"out_{n}_{j}"] = cv2_resize(im)
    np.savez_compressed(os.path.join(out, "prepare_resize_kat.npz"), **kat)
    # (2) whole scenes
    sc = {}
    for seed in range(6):
        H, W = ((128, 256), (128, 256), (64, 64), (100, 37), (256, 256), (128, 256))[seed]
        rgb, segm, ids = P.synthetic_scene(seed, H, W)
        crops, bbox, mask = P.prepare_obs_view(rgb, segm, ids, resize=cv2_resize)
        sc[f"rgb_{seed}"], sc[f"segm_{seed}"], sc[f"ids_{seed}"] = rgb, segm, np.asarray(ids, np.int64)
        sc[f"crops_{seed}"], sc[f"bbox_{seed}"], sc[f"mask_{seed}"] = crops, bbox, mask
    np.savez_compressed(os.path.join(out, "prepare_scenes.npz"), **sc)
    print("wrote", {k: os.path.getsize(os.path.join(out, k)) for k in ("prepare_resize_kat.npz", "prepare_scenes.npz")})


if __name__ == "__main__":
    main()
