"""Shared driver: one full pass over the policy's public methods (as scripts/example.py chains them) for a synthetic case."""
from __future__ import annotations

import torch

from oracle import detgen, synth


def to_dev(x, dev):
    if isinstance(x, dict):
        return {k: to_dev(v, dev) for k, v in x.items()}
    return x.to(dev)


_policy_cache = {}


def build_policy(model: str, device="cuda"):
    """vima_b200.VIMAPolicy with the deterministic shared weights, on `device`, eval mode."""
    import vima_b200

    if model not in _policy_cache:
        _policy_cache.clear()  # one resident model at a time
        pol = vima_b200.VIMAPolicy(**synth.MODEL_CFGS[model])
        detgen.fill_module_(pol)
        _policy_cache[model] = pol.to(device).eval()
    return _policy_cache[model]


@torch.no_grad()
def run_policy_case(policy, case: synth.Case, dev="cuda"):
    from vima_b200.utils import DataDict

    token_types, word_batch, image_batch = synth.make_prompt(case)
    prompt_tokens, prompt_masks = policy.forward_prompt_assembly((token_types, word_batch.to(dev), DataDict(to_dev(image_batch, dev))))
    obs = synth.make_obs(case)
    obs_dd = DataDict(to_dev(obs, dev))
    obs_tokens, obs_masks = policy.forward_obs_token(obs_dd)
    action_tokens = None
    if case.T > 1:
        action_tokens = policy.forward_action_token(to_dev(synth.make_actions(case, case.T), dev))
    predicted = policy.forward(obs_token=obs_tokens, obs_mask=obs_masks, action_token=action_tokens, prompt_token=prompt_tokens,
                               prompt_token_mask=prompt_masks)
    dists = policy.forward_action_decoder(predicted[-1:])
    logits_norm = torch.cat([d.logits for k in dists for d in dists[k]._dists], dim=-1)
    logits_raw = torch.cat([dists[k].raw_logits for k in dists], dim=-1)
    modes = {k: v.mode() for k, v in dists.items()}
    nxt = policy.forward_action_token(modes)
    return dict(prompt_tokens=prompt_tokens, prompt_masks=prompt_masks, obs_tokens=obs_tokens, obs_masks=obs_masks, action_tokens=action_tokens,
                predicted=predicted, logits_normalised=logits_norm, logits_raw=logits_raw, modes=modes, next_action_token=nxt)
