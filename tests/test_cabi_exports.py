"""CPU: the C-ABI library loads and exports every symbol include/vima_b200.h declares (no compute without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vima_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vima_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import __graft_entry__

    __graft_entry__.build()
    from vima_b200 import _C

    lib = _C.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vima_b200.h but not exported"
    assert sorted(_C.EXPORTS) == names
    assert lib.vima_abi_version() == 3


def test_no_cpu_fallback():
    """The product path refuses to run without an sm_100 device instead of silently falling back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU box")
    from vima_b200 import _C

    with pytest.raises(RuntimeError):
        _C.Context.get(torch.device("cpu"))
    with pytest.raises(RuntimeError):
        _C.Context(0)


def test_new_entry_points_refuse_cpu_inputs():
    """The rows added around the path (prepare, incremental decode, action post-processing, the baselines' encoders) have no
    CPU route either: CPU tensors raise instead of being computed by torch."""
    import numpy as np
    import torch

    import vima_b200
    from vima_b200 import nn as vnn
    from vima_b200.prepare import crop_objects

    if torch.cuda.is_available():
        pytest.skip("GPU box")
    with pytest.raises((RuntimeError, AssertionError)):
        crop_objects(np.zeros((1, 3, 8, 8), np.uint8), np.zeros((1, 8, 8), np.uint8), [1], device="cpu")
    pol = vima_b200.VIMAPolicy(embed_dim=256, xf_n_layers=1, sattn_n_heads=8, xattn_n_heads=8)
    with pytest.raises(RuntimeError, match="no CPU|CUDA"):
        pol.start_decode(torch.zeros(4, 1, 256), torch.ones(1, 4, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU|CUDA"):
        pol.postprocess_actions({"pose0_position": torch.zeros(1, 1, 2, dtype=torch.int64)}, torch.zeros(1, 2), torch.ones(1, 2))
    enc = vnn.ObjectsPerceiverEncoder(64, num_latents=4, num_blocks=1, num_self_attends_per_block=1, num_self_attention_heads=8,
                                      num_cross_attention_heads=8, attention_probs_dropout_prob=0.1)
    with pytest.raises(RuntimeError, match="no CPU|CUDA"):
        enc(torch.zeros(2, 16, 64))
