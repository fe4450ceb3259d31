#!/usr/bin/env python
"""Mint golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # all CPU-sized cases
    python tests/golden/make_golden.py cfg1 ...   # a subset

For every case in `oracle.synth.CASES` that fits the CPU it builds the reference `VIMAPolicy`
(`/root/reference`, shimmed by `oracle/ref_shim.py`), fills it with the deterministic weights of
`oracle.detgen`, feeds the deterministic inputs of `oracle.synth` through the reference's own public
methods (forward_prompt_assembly, forward_obs_token, forward_action_token, forward, forward_action_decoder,
.mode()) and stores the stage outputs as `tests/golden/<case>.npz`.  Large tensors are stored strided
(`<name>__stride<k>`): flat[::k].  The fixtures are what pins `oracle/vima_oracle.py` (CPU tests) and the
CUDA path (GPU tests) to the reference.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

MAX_ELEMS = 1 << 16
CPU_CASES = ["cfg1", "cfg1_t2", "ragged_4M", "cfg2_small", "cfg3_small"]


def pack(out: dict, name: str, t):
    import torch

    a = t.detach().cpu()
    if a.dtype == torch.bool:
        a = a.to(torch.uint8)
    a = a.numpy()
    if a.size > MAX_ELEMS:
        k = -(-a.size // MAX_ELEMS)
        out[f"{name}__stride{k}"] = a.reshape(-1)[::k].copy()
        out[f"{name}__shape"] = np.array(a.shape, dtype=np.int64)
    else:
        out[name] = a


def run_case(ref_vima, case_name: str):
    import torch

    from oracle import detgen, synth

    case = synth.CASES[case_name]
    cfg = synth.MODEL_CFGS[case.model]
    torch.manual_seed(0)
    policy = ref_vima.VIMAPolicy(**cfg).eval()
    detgen.fill_module_(policy)
    DataDict = sys.modules["vima.utils"].DataDict

    out = {}
    with torch.no_grad():
        token_types, word_batch, image_batch = synth.make_prompt(case)
        prompt_tokens, prompt_masks = policy.forward_prompt_assembly((token_types, word_batch, DataDict(image_batch)))
        pack(out, "prompt_tokens", prompt_tokens)
        pack(out, "prompt_masks", prompt_masks)

        obs = synth.make_obs(case)
        obs_dd = DataDict({"ee": obs["ee"], "objects": DataDict(obs["objects"])})
        obs_tokens, obs_masks = policy.forward_obs_token(obs_dd)
        pack(out, "obs_tokens", obs_tokens)
        pack(out, "obs_masks", obs_masks)

        action_tokens = None
        if case.T > 1:
            acts = synth.make_actions(case, case.T)
            action_tokens = policy.forward_action_token(acts)
            pack(out, "action_tokens", action_tokens)

        predicted = policy.forward(
            obs_token=obs_tokens,
            obs_mask=obs_masks,
            action_token=action_tokens,
            prompt_token=prompt_tokens,
            prompt_token_mask=prompt_masks,
        )
        pack(out, "predicted", predicted)

        dists = policy.forward_action_decoder(predicted[-1:])
        norm_logits = torch.cat([d.logits for k in dists for d in dists[k]._dists], dim=-1)
        raw_logits = torch.cat(
            [mlp(predicted[-1:]) for k in policy.action_decoder._decoders for mlp in policy.action_decoder._decoders[k].mlps],
            dim=-1,
        )
        pack(out, "logits_normalised", norm_logits)
        pack(out, "logits_raw", raw_logits)
        modes = {k: v.mode() for k, v in dists.items()}
        for k, v in modes.items():
            assert v.dtype == torch.int64
            pack(out, f"mode.{k}", v)
        next_action_token = policy.forward_action_token(modes)
        pack(out, "next_action_token", next_action_token)
    return out


def run_gato_case(ref_vima, case_name: str):
    """VIMAGatoPolicy (gato_*) or VIMAGPTPolicy (gpt_*): same method chain, one token per observation for the latter."""
    import torch

    from oracle import detgen, synth

    if case_name.startswith("flamingo"):
        case = synth.FLAMINGO_CASES[case_name]
        policy = sys.modules["vima.policy"].VIMAFlamingoPolicy(**synth.FLAMINGO_CFGS[case.model]).eval()
    elif case_name.startswith("gpt"):
        case = synth.GPT_CASES[case_name]
        policy = sys.modules["vima.policy"].VIMAGPTPolicy(**synth.GATO_CFGS[case.model]).eval()
    else:
        case = synth.GATO_CASES[case_name]
        policy = ref_vima.VIMAGatoPolicy(**synth.GATO_CFGS[case.model]).eval()
    detgen.fill_module_(policy)
    DataDict = sys.modules["vima.utils"].DataDict
    out = {}
    with torch.no_grad():
        token_types, word_batch, image_batch = synth.make_gato_prompt(case)
        prompt_tokens, prompt_masks = policy.forward_prompt_assembly((token_types, word_batch, DataDict(image_batch)))
        pack(out, "prompt_tokens", prompt_tokens)
        pack(out, "prompt_masks", prompt_masks)
        obs = synth.make_gato_obs(case)
        obs_tokens = policy.forward_obs_token(DataDict({"ee": obs["ee"], "rgb": DataDict(obs["rgb"])}))
        pack(out, "obs_tokens", obs_tokens)
        action_tokens = policy.forward_action_token(synth.make_actions(case, case.T)) if case.T > 1 else None
        if action_tokens is not None:
            pack(out, "action_tokens", action_tokens)
        predicted = policy.forward(obs_token=obs_tokens, action_token=action_tokens, prompt_token=prompt_tokens, prompt_token_mask=prompt_masks)
        pack(out, "predicted", predicted)
        dists = policy.forward_action_decoder(predicted[-1:])
        raw_logits = torch.cat([mlp(predicted[-1:]) for k in policy.action_decoder._decoders for mlp in policy.action_decoder._decoders[k].mlps], dim=-1)
        pack(out, "logits_raw", raw_logits)
        for k, v in dists.items():
            pack(out, f"mode.{k}", v.mode())
    return out


def main():
    names = sys.argv[1:] or (CPU_CASES + ["gato_small", "gpt_small", "flamingo_small"])
    from oracle.ref_shim import load_reference

    ref_vima = load_reference()
    import torch

    torch.set_num_threads(os.cpu_count())
    for n in names:
        t0 = time.time()
        out = run_gato_case(ref_vima, n) if n.startswith(("gato", "gpt", "flamingo")) else run_case(ref_vima, n)
        path = os.path.join(HERE, f"{n}.npz")
        np.savez_compressed(path, **out)
        print(f"{n}: {len(out)} arrays -> {path} ({os.path.getsize(path)/1e3:.0f} kB) in {time.time()-t0:.1f}s")


if __name__ == "__main__":
    main()
