"""CPU: host-side logic of the product that needs no GPU -- utils/DataDict surface, T5 bucket table, state-dict contract."""
import numpy as np
import torch

from oracle import synth
from oracle.state_dict_spec import state_dict_spec


def test_policy_state_dict_contract():
    """Keys, order and shapes equal the reference's (strict=True loading of released checkpoints)."""
    import vima_b200

    cfg = synth.MODEL_CFGS["2M"]
    pol = vima_b200.VIMAPolicy(**cfg)
    sd = pol.state_dict()
    spec = state_dict_spec(**cfg)
    assert list(sd.keys()) == list(spec.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(spec[k]), k
    # round trip through the checkpoint format of vima/__init__.py:7-16
    ck = {"cfg": cfg, "state_dict": {"policy." + k: v for k, v in sd.items()}}
    import tempfile, os

    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ck.pt")
        torch.save(ck, path)
        pol2 = vima_b200.create_policy_from_ckpt(path, "cpu")
    assert not pol2.training
    assert torch.equal(pol2.state_dict()["xattn_gpt.h.0.attn.c_attn.weight"], sd["xattn_gpt.h.0.attn.c_attn.weight"])


def test_vima_alias_package_surface():
    import vima
    import vima.nn as vnn
    from vima.utils import any_to_datadict  # noqa: F401

    for name in ["ActionDecoder", "ActionEmbedding", "ContinuousActionEmbedding", "ObjEncoder", "T5PromptEncoder", "WordEmbedding", "XAttnGPT",
                 "Embedding", "build_mlp"]:
        assert hasattr(vnn, name), name
    assert hasattr(vima, "create_policy_from_ckpt") and hasattr(vima, "VIMAPolicy")


def test_t5_buckets_match_hf():
    from transformers.models.t5.modeling_t5 import T5Attention

    from vima_b200.nn.t5_encoder import relative_position_buckets

    for L in (1, 10, 256, 300):
        ours = relative_position_buckets(L)
        rel = torch.arange(-(L - 1), L)[None, :]
        hf = T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128)[0]
        assert ours.dtype == torch.int64 and torch.equal(ours, hf)


def test_utils_surface():
    from vima_b200.utils import (DataDict, add_batch_dim, any_concat, any_slice, any_stack, any_to_datadict, any_to_torch_tensor,
                                 any_transpose_first_two_axes, get_batch_size, stack_sequence_fields)

    d = any_to_datadict({"objects": {"mask": {"top": np.ones((2, 3), bool), "front": np.zeros((2, 3), bool)}}, "ee": np.array([0, 1])})
    assert isinstance(d["objects"], DataDict) and d["objects.mask.top"].shape == (2, 3) and d.objects.mask.front.sum() == 0
    t = d.to_torch_tensor(device="cpu")
    assert torch.is_tensor(t["ee"]) and t["objects"]["mask"]["top"].dtype == torch.bool
    r = t["objects"].map_structure(func=lambda x: x.reshape(-1, *x.shape[2:]))
    assert r["mask"]["front"].shape == (6,)
    assert list(dict(**t["objects"]).keys()) == ["mask"]
    s = any_stack([{"a": np.zeros(3)}, {"a": np.ones(3)}], dim=0)
    assert s["a"].shape == (2, 3)
    c = any_concat([torch.zeros(2, 1), torch.ones(2, 2)], dim=-1)
    assert c.shape == (2, 3)
    assert any_slice({"a": np.arange(6).reshape(2, 3)}, np.s_[0, 1])["a"] == 1
    assert add_batch_dim({"a": np.zeros(3)})["a"].shape == (1, 3)
    assert any_transpose_first_two_axes({"a": torch.zeros(2, 5, 1)})["a"].shape == (5, 2, 1)
    assert get_batch_size({"b": np.zeros((4, 2)), "a": np.zeros((4, 1))}) == 4
    st = stack_sequence_fields([{"x": np.zeros(2), "y": {"z": np.ones(1)}}, {"x": np.ones(2), "y": {"z": np.ones(1)}}])
    assert st["x"].shape == (2, 2) and st["y"]["z"].shape == (2, 1)
    assert any_to_torch_tensor([True, False], dtype=torch.bool).dtype == torch.bool
