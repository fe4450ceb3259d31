"""GPU parity: vima_b200.VIMAPolicy (CUDA path through the C ABI) vs golden vectors minted from the unmodified
reference, on shared deterministic weights.  Tolerance from BASELINE.json north_star: 1e-3 rel (fp32 reference),
bit-exact masks / indices."""
import numpy as np
import pytest
import torch

from oracle import synth, vima_oracle as O
from tests.policy_runner import build_policy, run_policy_case
from tests.util import allclose_ratio, argmax_safe_mask, golden_pick, load_golden, max_rel, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-3  # north_star tolerance; the default "f16x3" mode lands around 1e-5
DIMS = [n for d in O.ACTION_DIMS.values() for n in d]


def check_against_golden(name, r, tol, exact_modes=True):
    g = load_golden(name)
    for key in ["prompt_masks", "obs_masks"]:
        e, a = golden_pick(g, key, r[key])
        assert np.array_equal(e, a), f"{name}.{key}: masks must be bit-exact"
    errs = {}
    for key in ["prompt_tokens", "obs_tokens", "action_tokens", "predicted", "logits_raw", "logits_normalised", "next_action_token"]:
        if r.get(key) is None:
            continue
        e, a = golden_pick(g, key, r[key])
        assert np.isfinite(a).all(), f"{name}.{key} has non-finite values"
        errs[key] = rel_l2(e, a)
    worst = max(errs.values())
    assert worst <= tol, f"{name}: rel-L2 errors {errs}"
    # action indices: bit-exact wherever the reference's top-2 logit gap exceeds the fp tolerance (ties are not defined)
    raw = g["logits_raw"] if "logits_raw" in g else None
    got = torch.cat([r["modes"][k] for k in O.ACTION_DIMS], dim=-1).cpu().numpy()
    exp = np.concatenate([g[f"mode.{k}"] for k in O.ACTION_DIMS], axis=-1)
    assert got.dtype == np.int64
    safe = argmax_safe_mask(raw, DIMS, margin=4 * tol * np.abs(raw).max()) if exact_modes else np.zeros_like(exp, dtype=bool)
    assert np.array_equal(got[safe], exp[safe]), f"{name}: action indices differ outside numerical ties"
    if exact_modes:
        assert safe.mean() > 0.9
    return errs


@pytest.mark.parametrize("name", ["cfg1", "cfg1_t2", "ragged_4M", "cfg2_small", "cfg3_small"])
def test_policy_matches_reference_golden(name):
    import vima_b200

    vima_b200.set_precision("f16x3")
    case = synth.CASES[name]
    pol = build_policy(case.model)
    errs = check_against_golden(name, run_policy_case(pol, case), TOL)
    print(name, {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.parametrize("name", ["cfg2_small", "cfg3_small"])
def test_f16f8_mode_within_north_star_tolerance(name):
    """fp16 hi*hi + e4m3 cross terms for the decoder GEMMs (2 tensor pass-equivalents instead of 3) stays inside 1e-3."""
    import vima_b200

    case = synth.CASES[name]
    pol = build_policy(case.model)
    vima_b200.set_precision("f16f8")
    try:
        errs = check_against_golden(name, run_policy_case(pol, case), TOL)
    finally:
        vima_b200.set_precision("f16x3")
    print("f16f8", name, {k: f"{v:.1e}" for k, v in errs.items()})


ELEM_RTOL, ELEM_ATOL_FRAC = 1e-3, 1e-4  # element-wise: |a - e| <= 1e-3 * |e| + 1e-4 * max|e|  (numpy.allclose form)


@pytest.mark.parametrize("mode", ["f16x3", "f16f8"])
@pytest.mark.parametrize("name", ["cfg2_small", "cfg3_small"])
def test_elementwise_tolerance(name, mode):
    """north_star says "1e-3 rel": beside the aggregate rel-L2, every element of the predicted tokens and of the raw logits is
    held to numpy.allclose(rtol = 1e-3, atol = 1e-4 of the tensor's largest magnitude).  The absolute term is needed because an
    element-wise relative error is undefined near zero; its size is set by fp32 itself, not by the operand format: the 3-pass
    f16x3 mode (products exact to ~2^-22) already differs from the CPU reference by up to 6.6e-4 of an element's own magnitude
    once elements down to 5 % of the maximum are held to a pure relative bound (measured table in DESIGN.md section 3; the
    per-floor maxima are printed here for the record)."""
    import vima_b200

    case = synth.CASES[name]
    pol = build_policy(case.model)
    vima_b200.set_precision(mode)
    try:
        r = run_policy_case(pol, case)
    finally:
        vima_b200.set_precision("f16x3")
    g = load_golden(name)
    worst, ratio = {}, {}
    for key in ["predicted", "logits_raw", "next_action_token"]:
        e, a = golden_pick(g, key, r[key])
        worst[key] = {f: max_rel(e, a, floor=f) for f in (0.1, 0.05, 0.01)}
        ratio[key] = allclose_ratio(e, a, ELEM_RTOL, ELEM_ATOL_FRAC)
    print(mode, name, {k: {f: f"{v:.1e}" for f, v in d.items()} for k, d in worst.items()}, {k: f"{v:.2f}" for k, v in ratio.items()})
    for key, v in ratio.items():
        assert v <= 1.0, (key, v, worst[key])


@pytest.mark.parametrize("mode,tol", [("bf16x3", 2e-3), ("f16", 6e-2), ("bf16", 0.5)])
def test_other_precision_modes(mode, tol):
    """bf16x3 stays near the fp32 bar; the single-pass modes are TF32/bf16-class and only reported."""
    import vima_b200

    case = synth.CASES["cfg2_small"]
    pol = build_policy(case.model)
    vima_b200.set_precision(mode)
    try:
        errs = check_against_golden("cfg2_small", run_policy_case(pol, case), tol, exact_modes=False)
    finally:
        vima_b200.set_precision("f16x3")
    print(mode, {k: f"{v:.1e}" for k, v in errs.items()})


def test_policy_matches_oracle_directly():
    """Same comparison against the CPU oracle run in this process (no fixtures involved)."""
    import vima_b200
    from tests.test_oracle_golden import run_oracle_case

    vima_b200.set_precision("f16x3")
    name = "ragged_4M"
    case = synth.CASES[name]
    r = run_policy_case(build_policy(case.model), case)
    o = run_oracle_case(name)
    for key in ["prompt_tokens", "obs_tokens", "predicted", "logits_raw", "next_action_token"]:
        assert rel_l2(o[key].numpy(), r[key].cpu().numpy()) < TOL, key
    assert torch.equal(o["prompt_masks"], r["prompt_masks"].cpu()) and torch.equal(o["obs_masks"], r["obs_masks"].cpu())


def test_cpu_tensors_are_refused():
    import vima_b200

    pol = vima_b200.nn.XAttnGPT(64, n_layer=1, n_head=2, xattn_n_head=2, xattn_n_positions=16, use_geglu=True)
    with pytest.raises(RuntimeError, match="no CPU"):
        pol(obs_action_tokens=torch.zeros(3, 1, 64), prompt_tokens=torch.zeros(2, 1, 64))


def test_module_level_xattn_gpt_defaults():
    """vnn.XAttnGPT without masks / position ids (defaults of xattn_gpt.py:99-112), batch_first both ways."""
    import vima_b200
    from oracle import detgen
    from oracle.state_dict_spec import xattn_gpt_spec

    vima_b200.set_precision("f16x3")
    E, nl, H = 128, 2, 4
    m = vima_b200.nn.XAttnGPT(E, n_layer=nl, n_head=H, xattn_n_head=H, xattn_n_positions=256, use_geglu=True)
    detgen.fill_module_(m)
    m = m.cuda().eval()
    sd = {"x." + k: v.cpu() for k, v in m.state_dict().items()}
    L, Lp, B = 37, 19, 3
    x = detgen.uniform("mx", (L, B, E)); pr = detgen.uniform("mp", (Lp, B, E))
    with torch.no_grad():
        y = m(obs_action_tokens=x.cuda(), prompt_tokens=pr.cuda())
        yb = m(obs_action_tokens=x.transpose(0, 1).contiguous().cuda(), prompt_tokens=pr.transpose(0, 1).contiguous().cuda(), batch_first=True)
        ref = O.xattn_gpt_forward(sd, "x.", obs_action_tokens=x, prompt_tokens=pr, obs_action_position_ids=torch.arange(L)[None].expand(B, L),
                                  prompt_position_ids=torch.arange(Lp)[None].expand(B, Lp), prompt_mask=torch.ones(B, Lp, dtype=torch.bool),
                                  obs_action_masks=torch.ones(B, L, dtype=torch.bool), n_layer=nl, n_head=H, xattn_n_head=H)
    assert rel_l2(ref.numpy(), y.cpu().numpy()) < TOL
    assert torch.equal(y, yb.transpose(0, 1))


def test_postprocess_actions_bit_exact():
    """scripts/example.py:199-232 (de-discretise, bounds affine, clamp) as one kernel per key vs the oracle: bit-exact."""
    pol = build_policy("2M")
    g = torch.Generator().manual_seed(4)
    B = 37
    idx = {"pose0_position": torch.stack([torch.randint(0, 50, (1, B), generator=g), torch.randint(0, 100, (1, B), generator=g)], -1),
           "pose0_rotation": torch.randint(0, 50, (1, B, 4), generator=g),
           "pose1_position": torch.stack([torch.randint(0, 50, (1, B), generator=g), torch.randint(0, 100, (1, B), generator=g)], -1),
           "pose1_rotation": torch.randint(0, 50, (1, B, 4), generator=g)}
    idx["pose0_position"][0, 0] = torch.tensor([49, 99])
    idx["pose1_position"][0, 0] = torch.tensor([0, 0])
    low = torch.tensor([[0.25, -0.5]])
    high = torch.tensor([[0.75, 0.5]])
    ref = O.postprocess_actions(idx, low, high)
    got = pol.postprocess_actions({k: v.cuda() for k, v in idx.items()}, low.cuda(), high.cuda())
    for k in ref:
        assert got[k].dtype == torch.float32 and got[k].shape == ref[k].shape
        assert torch.equal(got[k].cpu(), ref[k]), k
    # one bounds row per episode
    low_b = low + torch.rand(B, 2, generator=g) * 0.1
    high_b = high + torch.rand(B, 2, generator=g) * 0.1
    ref = O.postprocess_actions(idx, low_b[None], high_b[None])
    got = pol.postprocess_actions({k: v.cuda() for k, v in idx.items()}, low_b.cuda(), high_b.cuda())
    for k in ref:
        assert torch.equal(got[k].cpu(), ref[k]), k


def test_xattn_gpt_512_prompt_tokens():
    """north_star allows prompts of up to 512 tokens; the reference's VIMAPolicy caps XAttnGPT at xattn_n_positions=256
    (vima_policy.py:26-38), so the 512 case exists only at the module level (SURVEY.md 8(d) row #3x): vnn.XAttnGPT built
    with xattn_n_positions=512 against the oracle's restatement of xattn_gpt.py:73-139."""
    import vima_b200
    from oracle import detgen
    from vima_b200 import nn as vnn

    vima_b200.set_precision("f16x3")
    E, nl, H, B, L, Lp = 256, 2, 8, 2, 67, 512
    mod = vnn.XAttnGPT(E, n_layer=nl, n_head=H, dropout=0.1, xattn_n_head=H, xattn_ff_expanding=4, xattn_n_positions=512, use_geglu=True)
    detgen.fill_module_(mod)
    sd = {"xattn_gpt." + k: v.detach().clone() for k, v in mod.state_dict().items()}
    mod = mod.cuda().eval()
    g = torch.Generator().manual_seed(9)
    tok = torch.randn(L, B, E, generator=g)
    ptk = torch.randn(Lp, B, E, generator=g)
    pmask = torch.ones(B, Lp, dtype=torch.bool)
    pmask[1, 300:] = False
    omask = torch.rand(B, L, generator=g) > 0.15
    omask[:, 0] = True
    oa_pos = torch.cumsum(omask, dim=1) - 1
    p_pos = torch.cumsum(pmask, dim=1) - 1
    with torch.no_grad():
        ref = O.xattn_gpt_forward(sd, "xattn_gpt.", obs_action_tokens=tok, obs_action_position_ids=oa_pos, prompt_tokens=ptk, prompt_mask=pmask,
                                  prompt_position_ids=p_pos, obs_action_masks=omask, n_layer=nl, n_head=H, xattn_n_head=H)
        got = mod(obs_action_tokens=tok.cuda(), obs_action_position_ids=oa_pos.cuda(), prompt_tokens=ptk.cuda(), prompt_mask=pmask.cuda(),
                  prompt_position_ids=p_pos.cuda(), obs_action_masks=omask.cuda())
    assert rel_l2(ref, got.cpu()) < 1e-3, rel_l2(ref, got.cpu())


def test_de_discretize_bit_exact_vs_reference_fixture():
    """policy._de_discretize_actions (one `vima_action_scale` launch per key) vs the fixture minted from the unmodified reference
    method (vima_policy.py:301-322, tests/golden/make_dediscretize_golden.py): bit-exact over every bin index of every head."""
    import os

    pol = build_policy("2M")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dediscretize.npz"))
    keys = ["pose0_position", "pose0_rotation", "pose1_position", "pose1_rotation"]
    got = pol._de_discretize_actions({k: torch.from_numpy(g[f"in.{k}"]).cuda() for k in keys})
    for k in keys:
        assert got[k].dtype == torch.float32
        assert np.array_equal(got[k].cpu().numpy(), g[f"out.{k}"]), k


def test_nongeglu_sequence_models_match_reference_fixture():
    """XAttnGPT / HFGPT with use_geglu=False (GEMM epilogue VIMA_ACT_GELU_TANH = HF gelu_new in the blocks, erf GELU without a gate in
    XAttention, LayerNorms folded as in the GEGLU path) vs the fixture minted from the unmodified reference modules."""
    import os

    import vima_b200
    from tests.golden.make_nongeglu_golden import inputs
    from tests.test_oracle_golden import _nongeglu_modules

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nongeglu.npz"))
    c, xg, hf = _nongeglu_modules()
    xg, hf = xg.cuda(), hf.cuda()
    x, pr, pmask, omask, pos, ppos = (t.cuda() for t in inputs())
    for mode in ("f16x3", "f16f8"):
        vima_b200.set_precision(mode)
        try:
            with torch.no_grad():
                y = xg(obs_action_tokens=x, obs_action_position_ids=pos, prompt_tokens=pr, prompt_mask=pmask, prompt_position_ids=ppos,
                       batch_first=False, obs_action_masks=omask)
                yh = hf(x, custom_mask=omask, position_ids=pos, batch_first=False)
        finally:
            vima_b200.set_precision("f16x3")
        ex, eh = rel_l2(g["xattn_gpt"], y.cpu().numpy()), rel_l2(g["hfgpt"], yh.cpu().numpy())
        print(mode, f"xattn_gpt {ex:.1e} hfgpt {eh:.1e}")
        assert ex <= TOL and eh <= TOL, (mode, ex, eh)
