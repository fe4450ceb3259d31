"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as oracle/vima_oracle.py: only tests/ may import it; the
product path `vima_b200.prepare` never does).  CPU restatement of the step immediately BEFORE the policy path --
`prepare_obs` / the object loop of `prepare_prompt` in /root/reference/scripts/example.py:243-473:
per view and object id, bounding box of the segmentation mask, crop, zero-pad to a square, `cv2.resize(..., (32, 32),
interpolation=cv2.INTER_AREA)`, valid objects first / padded slots after.

The resize lives in a third-party dependency that is not under /root/reference: OpenCV (`opencv-python`, unpinned in the
reference's requirements; 4.13.0 in this image).  `resize_area_u8` restates its published algorithm for 8-bit images
(modules/imgproc/src/resize.cpp):

* shrink by an integer factor k      -> `ResizeAreaFast_`: integer box sums; k == 2 rounds `(s + 2) >> 2`, other k
                                        `cvRound(float(s) * (1.f / k^2))`;
* shrink by a non-integer factor     -> `ResizeArea_<uchar, float>` with `computeResizeAreaTab`: float32 partial-cell weights,
                                        row buffer `buf += S * alpha` then `sum (+)= beta * buf`, `cvRound` to uint8;
* enlarge (crop smaller than 32)     -> INTER_AREA falls back to the bilinear kernel with "area" coefficients
                                        (`fx = (dx+1) - (sx+1)*inv_scale`), 11-bit fixed point:
                                        `((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2`;
* same size                          -> copy.

Pinned: tests/test_prepare_cpu.py checks it bit-for-bit against cv2 itself for every square size 1..256 and against the
committed fixtures tests/golden/prepare_*.npz (minted by tests/golden/make_prepare_golden.py with cv2).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32
OUT = 32


# --------------------------------------------------------------------------------------------------------------------
# cv2.resize(img, (32, 32), interpolation=cv2.INTER_AREA) for square uint8 HxWxC images
# --------------------------------------------------------------------------------------------------------------------
def _area_tab(ssize: int, dsize: int) -> List[Tuple[int, int, np.float32]]:
    """computeResizeAreaTab: (dst index, src index, weight) triples in source order."""
    scale = 1.0 / (float(dsize) / ssize)
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, f32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, f32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, f32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _round_u8(x: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)  # cvRound (half to even) + saturate


def _resize_area_general(img: np.ndarray, d: int) -> np.ndarray:
    """ResizeArea_<uchar, float>: per source row `buf[dx] += S[sx] * alpha` over the row's table entries (float32, in table
    order), then per destination row `sum = beta * buf` for its first source row and `sum += beta * buf` for the rest.
    Vectorised over rows / columns; the k-th entry of every destination index is applied in step k, which keeps each
    accumulator's own order of float additions."""
    n, cn = img.shape[0], img.shape[2]
    tab = _area_tab(n, d)
    per = [[(s, a) for (dd, s, a) in tab if dd == j] for j in range(d)]
    depth = max(len(p) for p in per)
    src = img.astype(f32)
    buf = np.zeros((n, d, cn), f32)  # horizontal pass of every source row
    for k in range(depth):
        js = [j for j in range(d) if len(per[j]) > k]
        sx = [per[j][k][0] for j in js]
        al = np.asarray([per[j][k][1] for j in js], f32)[None, :, None]
        buf[:, js] = buf[:, js] + src[:, sx] * al
    sums = np.zeros((d, d, cn), f32)
    for k in range(depth):
        js = [j for j in range(d) if len(per[j]) > k]
        sy = [per[j][k][0] for j in js]
        be = np.asarray([per[j][k][1] for j in js], f32)[:, None, None]
        sums[js] = be * buf[sy] if k == 0 else sums[js] + be * buf[sy]
    return _round_u8(sums)


def _resize_area_fast(img: np.ndarray, d: int) -> np.ndarray:
    k = img.shape[0] // d
    s = img.reshape(d, k, d, k, -1).astype(np.int64).sum(axis=(1, 3))
    if k == 2:
        return ((s + 2) >> 2).astype(np.uint8)
    return _round_u8(s.astype(f32) * f32(1.0 / (k * k)))


def _sat_short(v) -> int:
    return max(-32768, min(32767, int(np.rint(v))))


def _linear_area_tab(ssize: int, dsize: int):
    inv = float(dsize) / ssize
    scale = 1.0 / inv
    ofs, coef, xmax = [], [], dsize
    for dx in range(dsize):
        sx = math.floor(dx * scale)
        fx = f32((dx + 1) - (sx + 1) * inv)
        fx = f32(0) if fx <= 0 else f32(fx - f32(math.floor(fx)))
        if sx < 0:
            fx, sx = f32(0), 0
        if sx + 1 >= ssize:
            xmax = min(xmax, dx)
            if sx >= ssize - 1:
                fx, sx = f32(0), ssize - 1
        ofs.append(sx)
        coef.append((_sat_short(f32((f32(1.0) - fx) * f32(2048))), _sat_short(f32(fx * f32(2048)))))
    return ofs, coef, xmax


def _resize_enlarge(img: np.ndarray, d: int) -> np.ndarray:
    n, cn = img.shape[0], img.shape[2]
    ofs, coef, xmax = _linear_area_tab(n, d)
    src = img.astype(np.int64)
    hor = np.zeros((n, d, cn), np.int64)
    for dx in range(d):
        sx = ofs[dx]
        hor[:, dx] = src[:, sx] * coef[dx][0] + src[:, sx + 1] * coef[dx][1] if dx < xmax else src[:, sx] * 2048
    out = np.zeros((d, d, cn), np.uint8)
    for dy in range(d):
        s0 = min(max(ofs[dy], 0), n - 1)
        s1 = min(max(ofs[dy] + 1, 0), n - 1)
        b0, b1 = coef[dy]
        v = (((b0 * (hor[s0] >> 4)) >> 16) + ((b1 * (hor[s1] >> 4)) >> 16) + 2) >> 2
        out[dy] = np.clip(v, 0, 255).astype(np.uint8)
    return out


def resize_area_u8(img: np.ndarray, d: int = OUT) -> np.ndarray:
    """cv2.resize(img, (d, d), interpolation=cv2.INTER_AREA) for a square uint8 (n, n, C) image (example.py:293-297,434-438)."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[0] == img.shape[1]
    n = img.shape[0]
    if n == d:
        return img.copy()
    if n < d:
        return _resize_enlarge(img, d)
    if n % d == 0:
        return _resize_area_fast(img, d)
    return _resize_area_general(img, d)


# --------------------------------------------------------------------------------------------------------------------
# object loop (example.py:273-299 / 407-439)
# --------------------------------------------------------------------------------------------------------------------
def crop_object(rgb: np.ndarray, segm: np.ndarray, obj_id: int, resize=resize_area_u8) -> Optional[Tuple[List[int], np.ndarray]]:
    """rgb (3,H,W) uint8, segm (H,W) int -> ([x_center, y_center, h, w], crop (3,32,32) uint8), or None when fewer than two
    pixels carry `obj_id` (example.py:409-411)."""
    ys, xs = np.nonzero(segm == obj_id)
    if len(xs) < 2 or len(ys) < 2:
        return None
    xmin, xmax, ymin, ymax = int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())
    bbox = [int((xmin + xmax) / 2), int((ymin + ymax) / 2), ymax - ymin, xmax - xmin]
    crop = rgb[:, ymin:ymax + 1, xmin:xmax + 1]
    ch, cw = crop.shape[1], crop.shape[2]
    if ch != cw:
        diff = abs(ch - cw)
        before, after = diff // 2, diff - diff // 2
        pad = ((0, 0), (0, 0), (before, after)) if ch > cw else ((0, 0), (before, after), (0, 0))
        crop = np.pad(crop, pad, mode="constant", constant_values=0)
    out = resize(np.ascontiguousarray(crop.transpose(1, 2, 0)))
    return bbox, np.ascontiguousarray(out.transpose(2, 0, 1))


def prepare_obs_view(rgb: np.ndarray, segm: np.ndarray, objects: Sequence[int], resize=resize_area_u8):
    """One (step, view) of prepare_obs (example.py:401-456): rgb (3,H,W), segm (H,W) -> cropped_img (n,3,32,32) uint8,
    bbox (n,4) int64, mask (n,) bool with the visible objects first and zero-filled slots for the others."""
    n = len(objects)
    crops = np.zeros((n, 3, OUT, OUT), np.uint8)
    bbox = np.zeros((n, 4), np.int64)
    mask = np.zeros((n,), bool)
    k = 0
    for obj_id in objects:
        r = crop_object(rgb, segm, obj_id, resize)
        if r is None:
            continue
        bbox[k], crops[k], mask[k] = r[0], r[1], True
        k += 1
    return crops, bbox, mask


def prepare_obs(rgb_dict: Dict[str, np.ndarray], segm_dict: Dict[str, np.ndarray], ee: np.ndarray, objects: Sequence[int],
                resize=resize_area_u8):
    """prepare_obs (example.py:377-473) on arrays: rgb_dict[view] (L,3,H,W) uint8, segm_dict[view] (L,H,W), ee (L,) ->
    the nested dict the policy's forward_obs_token takes, leading dims (L, 1)."""
    views = sorted(rgb_dict.keys())
    L = rgb_dict[views[0]].shape[0]
    out = {"ee": np.asarray(ee)[:, None], "objects": {"cropped_img": {}, "bbox": {}, "mask": {}}}
    for v in views:
        per = [prepare_obs_view(rgb_dict[v][l], segm_dict[v][l], objects, resize) for l in range(L)]
        out["objects"]["cropped_img"][v] = np.stack([p[0] for p in per])[:, None]
        out["objects"]["bbox"][v] = np.stack([p[1] for p in per])[:, None]
        out["objects"]["mask"][v] = np.stack([p[2] for p in per])[:, None]
    return out


def prompt_asset_objects(rgb: np.ndarray, segm: np.ndarray, objects: Sequence[int], resize=resize_area_u8):
    """The object loop of prepare_prompt for one asset view (example.py:273-302): invisible objects are DROPPED (no padding
    here; padding to the per-prompt maximum happens later, example.py:311-352)."""
    got = [r for r in (crop_object(rgb, segm, o, resize) for o in objects) if r is not None]
    bbox = np.asarray([g[0] for g in got], dtype=np.int64).reshape(-1, 4)
    crops = np.asarray([g[1] for g in got], dtype=np.uint8).reshape(-1, 3, OUT, OUT)
    return crops, bbox


# --------------------------------------------------------------------------------------------------------------------
# synthetic inputs for the fixtures and the GPU parity tests
# --------------------------------------------------------------------------------------------------------------------
def synthetic_scene(seed: int, H: int = 128, W: int = 256, n_obj: int = 8):
    """Random rectangles / blobs of object ids 1..n_obj over background 0; some ids absent, one a single pixel."""
    rng = np.random.default_rng(seed)
    rgb = rng.integers(0, 256, (3, H, W), dtype=np.uint8)
    segm = np.zeros((H, W), np.uint8)
    ids = list(range(1, n_obj + 1))
    for oid in ids:
        kind = rng.integers(0, 6)
        if kind == 0:
            continue  # invisible object
        if kind == 1:
            segm[rng.integers(0, H), rng.integers(0, W)] = oid  # a single pixel: dropped (fewer than 2)
            continue
        h, w = int(rng.integers(1, H // 2 + 60)), int(rng.integers(1, W // 2 + 60))
        if kind == 2:
            h = w = int(rng.choice([2, 7, 16, 31, 32, 33, 64, 96, 100]))
        h, w = min(h, H), min(w, W)
        y0, x0 = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
        blob = rng.random((h, w)) < (1.0 if kind < 4 else 0.6)
        blob[0, 0] = blob[-1, -1] = True
        segm[y0:y0 + h, x0:x0 + w][blob] = oid
    return rgb, segm, ids
