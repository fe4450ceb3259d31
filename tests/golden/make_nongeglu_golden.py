#!/usr/bin/env python
"""Mint the fixture for the NON-GEGLU sequence models from the unmodified reference modules (run in the build container only):
`vima.nn.XAttnGPT(use_geglu=False)` (xattn_gpt.py:13-139, components.py:82-103,106-228) and `vima.nn.HFGPT(use_geglu=False)`
(gpt.py:15-301): MLPs are act(c_fc(x)) with ACT_FNS["gelu"] (HF NewGELUActivation), XAttention's feed-forward is
linear2(gelu(linear1(ln(a)))) without a gate.  No released VIMA checkpoint uses this configuration; the fixture pins the oracle
(tests/test_oracle_golden.py) and the CUDA path (tests/test_policy_gpu.py) to the reference for it anyway.
Output: tests/golden/nongeglu.npz (deterministic weights from oracle.detgen keyed by state-dict key, deterministic inputs)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CFG = dict(E=128, n_layer=2, n_head=4, xattn_n_head=4, n_positions=96, xattn_n_positions=40, B=3, L=70, Lp=33)


def inputs():
    import torch

    from oracle import detgen

    c = CFG
    x = detgen.uniform("nongeglu.x", (c["L"], c["B"], c["E"]), 3)
    pr = detgen.uniform("nongeglu.prompt", (c["Lp"], c["B"], c["E"]), 3)
    n_valid = torch.tensor([c["Lp"], c["Lp"] - 7, 5])
    pmask = torch.arange(c["Lp"])[None, :] < n_valid[:, None]
    omask = detgen.randint("nongeglu.omask", (c["B"], c["L"]), 0, 10, 3) > 1
    omask[:, 0] = True
    pos = (torch.cumsum(omask.long(), dim=1) - 1).clamp(min=0)
    ppos = torch.arange(c["Lp"])[None, :].expand(c["B"], c["Lp"]).contiguous()
    return x, pr, pmask, omask, pos, ppos


def main():
    import torch

    from oracle import detgen, ref_shim

    ref_shim.load_reference()
    vnn = sys.modules["vima.nn"]
    c = CFG
    torch.manual_seed(0)
    xg = vnn.XAttnGPT(c["E"], n_layer=c["n_layer"], n_head=c["n_head"], dropout=0.1, xattn_n_head=c["xattn_n_head"], xattn_ff_expanding=4,
                      xattn_n_positions=c["xattn_n_positions"], n_positions=c["n_positions"], use_geglu=False).eval()
    hf = vnn.HFGPT(n_positions=c["n_positions"], n_embd=c["E"], n_layer=c["n_layer"], n_head=c["n_head"], dropout=0.1, use_geglu=False).eval()
    detgen.fill_module_(xg)
    detgen.fill_module_(hf)
    assert not any("gated_layer" in k for k in xg.state_dict()) and not any("gated_layer" in k for k in hf.state_dict())
    x, pr, pmask, omask, pos, ppos = inputs()
    with torch.no_grad():
        y_x = xg(obs_action_tokens=x, obs_action_position_ids=pos, prompt_tokens=pr, prompt_mask=pmask, prompt_position_ids=ppos,
                 batch_first=False, obs_action_masks=omask)
        y_h = hf(x, custom_mask=omask, position_ids=pos, batch_first=False)
    out = {"xattn_gpt": y_x.numpy(), "hfgpt": y_h.numpy(), "xattn_gpt_keys": np.array(sorted(xg.state_dict().keys())),
           "hfgpt_keys": np.array(sorted(hf.state_dict().keys()))}
    np.savez_compressed(os.path.join(HERE, "nongeglu.npz"), **out)
    print("wrote nongeglu.npz", {k: getattr(v, "shape", None) for k, v in out.items()}, float(np.abs(out["xattn_gpt"]).max()), float(np.abs(out["hfgpt"]).max()))


if __name__ == "__main__":
    main()
