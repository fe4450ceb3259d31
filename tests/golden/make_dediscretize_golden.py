#!/usr/bin/env python
"""Mint the de-discretisation fixture from the UNMODIFIED reference method `VIMAPolicy._de_discretize_actions`
(/root/reference/vima/policy/vima_policy.py:301-322; run in the build container only).

The method only reads the three bin counts from `self`, so it is called unbound on a stub carrying the reference's default bin
counts (vima_policy.py:77-79: 50 x-bins, 100 y-bins, 50 rotation bins): no model is built.  Every bin index of every head is
covered.  Output: tests/golden/dediscretize.npz (int64 indices in, float32 values out)."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    import torch

    from oracle import ref_shim

    ref = ref_shim.load_reference()
    stub = types.SimpleNamespace(_n_discrete_x_bins=50, _n_discrete_y_bins=100, _n_discrete_rot_bins=50)
    n = 100
    idx = torch.arange(n, dtype=torch.int64)
    actions = {
        "pose0_position": torch.stack([idx % 50, idx], dim=-1)[None],                     # [1, 100, 2]: x in 0..49, y in 0..99
        "pose0_rotation": torch.stack([idx % 50, (idx * 7) % 50, (idx * 3 + 1) % 50, 49 - idx % 50], dim=-1)[None],
        "pose1_position": torch.stack([49 - idx % 50, 99 - idx], dim=-1)[None],
        "pose1_rotation": torch.stack([(idx * 11) % 50, idx % 50, (idx + 25) % 50, (idx * 13 + 5) % 50], dim=-1)[None],
    }
    out = ref.VIMAPolicy._de_discretize_actions(stub, {k: v.clone() for k, v in actions.items()})
    pack = {}
    for k in actions:
        pack[f"in.{k}"] = actions[k].numpy()
        pack[f"out.{k}"] = out[k].numpy()
        assert out[k].dtype == torch.float32
    np.savez_compressed(os.path.join(HERE, "dediscretize.npz"), **pack)
    print("wrote dediscretize.npz", {k: v.shape for k, v in pack.items()})


if __name__ == "__main__":
    main()
