"""CPU: pins oracle/vima_oracle.py (the restatement) to golden vectors minted from the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import detgen, synth, vima_oracle as O
from tests.util import assert_close, golden_pick, load_golden

TOL = 2e-5  # fp32 CPU vs fp32 CPU: only summation-order noise

_sd_cache = {}


def oracle_state_dict(model: str):
    """State dict with the reference's keys/shapes, filled by detgen (no reference import needed)."""
    if model not in _sd_cache:
        from oracle.state_dict_spec import state_dict_spec

        sd = {}
        for k, shape in state_dict_spec(**synth.MODEL_CFGS[model]).items():
            w = detgen.weight_for(k, shape)
            if w is not None:
                sd[k] = w
        _sd_cache[model] = sd
    return _sd_cache[model]


def run_oracle_case(name):
    case = synth.CASES[name]
    cfg = synth.MODEL_CFGS[case.model]
    sd = oracle_state_dict(case.model)
    with torch.no_grad():
        prompt_tokens, prompt_masks, _ = O.forward_prompt_assembly(sd, synth.make_prompt(case))
        obs_tokens, obs_masks = O.forward_obs_token(sd, synth.make_obs(case))
        action_tokens = None
        if case.T > 1:
            action_tokens = O.forward_action_token(sd, synth.make_actions(case, case.T))
        predicted = O.policy_forward(sd, obs_tokens, obs_masks, action_tokens, prompt_tokens, prompt_masks,
                                     n_head=cfg["sattn_n_heads"], xattn_n_head=cfg["xattn_n_heads"])
        logits = O.action_decoder_logits(sd, predicted[-1:])
        modes = O.action_modes(logits)
        nxt = O.forward_action_token(sd, modes)
    return dict(prompt_tokens=prompt_tokens, prompt_masks=prompt_masks, obs_tokens=obs_tokens, obs_masks=obs_masks,
                action_tokens=action_tokens, predicted=predicted, logits_raw=logits, modes=modes, next_action_token=nxt)


@pytest.mark.parametrize("name", ["cfg1", "cfg1_t2", "ragged_4M", "cfg2_small"])
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    r = run_oracle_case(name)
    for key in ["prompt_masks", "obs_masks"]:
        e, a = golden_pick(g, key, r[key])
        assert np.array_equal(e, a), key  # masks are bit-exact
    for key in ["prompt_tokens", "obs_tokens", "predicted", "logits_raw", "next_action_token"]:
        e, a = golden_pick(g, key, r[key])
        assert_close(f"{name}.{key}", e, a, TOL)
    if r["action_tokens"] is not None:
        e, a = golden_pick(g, "action_tokens", r["action_tokens"])
        assert_close(f"{name}.action_tokens", e, a, TOL)
    norm = torch.cat([torch.log_softmax(x, -1) for x in torch.split(r["logits_raw"], [n for d in O.ACTION_DIMS.values() for n in d], dim=-1)], -1)
    e, a = golden_pick(g, "logits_normalised", norm)
    assert_close(f"{name}.logits_normalised", e, a, TOL)
    for k, v in r["modes"].items():
        e, a = golden_pick(g, f"mode.{k}", v)
        assert v.dtype == torch.int64
        assert np.array_equal(e, a), f"action indices differ for {k}"


def test_oracle_matches_reference_golden_cfg3_small():
    """BASELINE configs[2] shapes (200M, Lp=256, L=263) at B=2 -- the slowest CPU test (~20 s)."""
    test_oracle_matches_reference_golden("cfg3_small")


def test_position_id_kats():
    """mask 111011011 -> ids 012234456 ; prompt mask 111100 -> 012333."""
    m = torch.tensor([1, 1, 1, 0, 1, 1, 0, 1, 1], dtype=torch.bool)
    assert (torch.cumsum(m, 0) - 1).tolist() == [0, 1, 2, 2, 3, 4, 4, 5, 6]
    obs_mask = torch.tensor([[[1, 1, 1, 0]], [[1, 1, 0, 1]]], dtype=torch.bool)  # (T=2,B=1,Q=4)
    toks, masks, ids = O.assemble_history(torch.zeros(2, 1, 4, 8), obs_mask, torch.zeros(1, 1, 8))
    assert masks[:, 0].tolist() == [True, True, True, False, True, True, True, False, True]
    assert ids[:, 0].tolist() == [0, 1, 2, 2, 3, 4, 5, 5, 6]
    pm = torch.tensor([[1, 1, 1, 1, 0, 0]], dtype=torch.bool)
    assert (torch.cumsum(pm, 1) - 1).tolist() == [[0, 1, 2, 3, 3, 3]]


def test_padded_key_perturbation_is_exactly_invisible():
    """perturbing a padded prompt token / padded obs key changes nothing, bit for bit."""
    sd = oracle_state_dict("2M")
    E = 256
    torch.manual_seed(0)
    L, Lp, B = 7, 6, 2
    x = torch.randn(L, B, E)
    pr = torch.randn(Lp, B, E)
    pm = torch.tensor([[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1]], dtype=torch.bool)
    om = torch.tensor([[1, 1, 0, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1, 1]], dtype=torch.bool)
    kw = dict(n_layer=1, n_head=8, xattn_n_head=8, prompt_mask=pm, obs_action_masks=om,
              obs_action_position_ids=(torch.cumsum(om, 1) - 1), prompt_position_ids=(torch.cumsum(pm, 1) - 1))
    with torch.no_grad():
        y0 = O.xattn_gpt_forward(sd, "xattn_gpt.", obs_action_tokens=x, prompt_tokens=pr, **kw)
        pr2 = pr.clone(); pr2[4:, 0] += 3.0
        y1 = O.xattn_gpt_forward(sd, "xattn_gpt.", obs_action_tokens=x, prompt_tokens=pr2, **kw)
        assert torch.equal(y0, y1)
        x2 = x.clone(); x2[2, 0] += 3.0  # padded obs token: only its own row may change
        y2 = O.xattn_gpt_forward(sd, "xattn_gpt.", obs_action_tokens=x2, prompt_tokens=pr, **kw)
        keep = [0, 1, 3, 4, 5, 6]
        assert torch.equal(y0[keep, 0], y2[keep, 0]) and torch.equal(y0[:, 1], y2[:, 1])
        assert not torch.equal(y0[2, 0], y2[2, 0])


def test_t5_bucket_table_matches_hf():
    """Bucket ids are int64 and equal HF's own `_relative_position_bucket` for every distance in range."""
    from transformers.models.t5.modeling_t5 import T5Attention

    rel = torch.arange(-300, 301)[None, :]
    ours = O.t5_relative_position_bucket(rel)
    hf = T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128)
    assert ours.dtype == torch.int64 and torch.equal(ours, hf)


def test_de_discretize_matches_reference_fixture():
    """oracle.de_discretize_actions vs the fixture minted from the unmodified VIMAPolicy._de_discretize_actions
    (tests/golden/make_dediscretize_golden.py; vima_policy.py:301-322): bit-exact over every bin index of every head."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dediscretize.npz"))
    keys = ["pose0_position", "pose0_rotation", "pose1_position", "pose1_rotation"]
    got = O.de_discretize_actions({k: torch.from_numpy(g[f"in.{k}"]) for k in keys})
    for k in keys:
        assert got[k].dtype == torch.float32
        assert np.array_equal(got[k].numpy(), g[f"out.{k}"]), k


def _nongeglu_modules():
    from tests.golden.make_nongeglu_golden import CFG as c
    import vima_b200.nn as vnn

    xg = vnn.XAttnGPT(c["E"], n_layer=c["n_layer"], n_head=c["n_head"], dropout=0.1, xattn_n_head=c["xattn_n_head"], xattn_ff_expanding=4,
                      xattn_n_positions=c["xattn_n_positions"], n_positions=c["n_positions"], use_geglu=False).eval()
    hf = vnn.HFGPT(n_positions=c["n_positions"], n_embd=c["E"], n_layer=c["n_layer"], n_head=c["n_head"], dropout=0.1, use_geglu=False).eval()
    detgen.fill_module_(xg)
    detgen.fill_module_(hf)
    return c, xg, hf


def test_nongeglu_oracle_and_state_dict_match_reference_fixture():
    """`use_geglu=False` (reference components.py:92-98,139-142,218-225; gpt.py:255-259): act(c_fc(x)) with HF gelu_new in the GPT
    blocks, gelu(linear1(ln(a))) without a gate in XAttention.  The oracle and our modules' state-dict keys vs the fixture minted
    from the unmodified reference modules (tests/golden/make_nongeglu_golden.py)."""
    import os

    from tests.golden.make_nongeglu_golden import inputs

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nongeglu.npz"))
    c, xg, hf = _nongeglu_modules()
    assert sorted(xg.state_dict().keys()) == list(g["xattn_gpt_keys"])
    assert sorted(k for k in hf.state_dict().keys()) == list(g["hfgpt_keys"])
    x, pr, pmask, omask, pos, ppos = inputs()
    sd = {k: v.detach() for k, v in xg.state_dict().items()}
    y = O.xattn_gpt_forward(sd, "", obs_action_tokens=x, obs_action_position_ids=pos, prompt_tokens=pr, prompt_mask=pmask,
                            prompt_position_ids=ppos, obs_action_masks=omask, n_layer=c["n_layer"], n_head=c["n_head"],
                            xattn_n_head=c["xattn_n_head"])
    assert_close("xattn_gpt(use_geglu=False)", g["xattn_gpt"], y.numpy(), 2e-5)
    sdh = {k: v.detach() for k, v in hf.state_dict().items()}
    yh = O.hfgpt_forward(sdh, "", x, omask, pos, c["n_head"])
    assert_close("hfgpt(use_geglu=False)", g["hfgpt"], yh.numpy(), 2e-5)
