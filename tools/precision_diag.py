"""GPU diagnostic: per-stage error of the CUDA path vs the reference goldens in each precision mode, with the prompt encoder
optionally held at f16x3 (isolates the T5/ViT f16f8 contribution from the decoder's)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vima_b200
from oracle import synth
from tests.policy_runner import build_policy, to_dev
from tests.util import golden_pick, load_golden, max_rel, rel_l2
from vima_b200.utils import DataDict


@torch.no_grad()
def run(policy, case, mode_prompt, mode_obs, mode_dec, dev="cuda"):
    token_types, word_batch, image_batch = synth.make_prompt(case)
    vima_b200.set_precision(mode_prompt)
    pt, pm = policy.forward_prompt_assembly((token_types, word_batch.to(dev), DataDict(to_dev(image_batch, dev))))
    vima_b200.set_precision(mode_obs)
    ot, om = policy.forward_obs_token(DataDict(to_dev(synth.make_obs(case), dev)))
    vima_b200.set_precision(mode_dec)
    at = policy.forward_action_token(to_dev(synth.make_actions(case, case.T), dev)) if case.T > 1 else None
    pred = policy.forward(obs_token=ot, obs_mask=om, action_token=at, prompt_token=pt, prompt_token_mask=pm)
    dists = policy.forward_action_decoder(pred[-1:])
    raw = torch.cat([dists[k].raw_logits for k in dists], dim=-1)
    vima_b200.set_precision("f16x3")
    return dict(prompt_tokens=pt, obs_tokens=ot, predicted=pred, logits_raw=raw)


for name in ["cfg2_small", "cfg3_small"]:
    case = synth.CASES[name]
    pol = build_policy(case.model)
    g = load_golden(name)
    for combo in [("f16x3",) * 3, ("f16f8",) * 3, ("f16x3", "f16f8", "f16f8"), ("f16x3", "f16x3", "f16f8"), ("f16f8", "f16f8", "f16x3")]:
        r = run(pol, case, *combo)
        out = {}
        for key in ["prompt_tokens", "obs_tokens", "predicted", "logits_raw"]:
            e, a = golden_pick(g, key, r[key])
            out[key] = f"l2 {rel_l2(e, a):.1e} max@.1 {max_rel(e, a, 0.1):.1e} @.05 {max_rel(e, a, 0.05):.1e} @.01 {max_rel(e, a, 0.01):.1e}"
        print(name, "prompt/obs/decoder =", combo)
        for k, v in out.items():
            print("   ", k, v)
