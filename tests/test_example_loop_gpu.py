"""GPU: the inference loop of the reference's scripts/example.py:112-198 (B=1, growing history, per-step padding to the running
max object count, tokens cached across steps) driven through the `vima` drop-in package, against the CPU oracle running the same loop."""
import pytest
import torch

from oracle import detgen, synth, vima_oracle as O
from tests.test_oracle_golden import oracle_state_dict
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _obs(step, n_slots):
    lead = (1, 1)
    objs = {"cropped_img": {}, "bbox": {}, "mask": {}}
    for v in ("front", "top"):
        img = detgen.randint(f"loop.{step}.img.{v}", lead + (n_slots, 3, 32, 32), 0, 256).to(torch.uint8)
        img.view(-1)[0] = 250
        objs["cropped_img"][v] = img
        objs["bbox"][v] = detgen.randint(f"loop.{step}.bbox.{v}", lead + (n_slots, 4), 0, 128)
        m = torch.ones(lead + (n_slots,), dtype=torch.bool)
        if n_slots > 1:
            m[..., -1] = step % 2 == 0  # an invisible (padded) object appended after the visible ones
        objs["mask"][v] = m
    return {"ee": detgen.randint(f"loop.{step}.ee", lead, 0, 2), "objects": objs}


def _pad_cat(tokens, masks, cat, stack, zeros):
    max_objs = max(x.shape[0] for x in tokens)
    tt, mm = [], []
    for t, m in zip(tokens, masks):
        pad = max_objs - t.shape[0]
        tt.append(cat([t, zeros((pad, t.shape[1]), t.dtype, t.device)], 0))
        mm.append(cat([m, zeros((pad,), m.dtype, m.device)], 0))
    return stack(tt, 0)[:, None], stack(mm, 0)[:, None]  # (T, B=1, Q, E), (T, 1, Q)


def test_example_style_loop_matches_oracle():
    import vima  # the drop-in alias package
    from vima.utils import DataDict, any_concat, any_stack

    vima.set_precision("f16x3")
    cfg = synth.MODEL_CFGS["4M"]
    pol = vima.VIMAPolicy(**cfg)
    detgen.fill_module_(pol)
    pol = pol.cuda().eval()
    sd = oracle_state_dict("4M")
    case = synth.CASES["cfg1"]
    tt, wb, ib = synth.make_prompt(case)
    dev = lambda x: {k: dev(v) for k, v in x.items()} if isinstance(x, dict) else x.cuda()
    z = lambda shape, dt, d: torch.zeros(shape, dtype=dt, device=d)
    with torch.no_grad():
        p_tok, p_msk = pol.forward_prompt_assembly((tt, wb.cuda(), DataDict(dev(ib))))
        rp_tok, rp_msk, _ = O.forward_prompt_assembly(sd, (tt, wb, ib))
        cache = {"t": [], "m": [], "a": []}
        ref = {"t": [], "m": [], "a": []}
        for step, n_slots in enumerate([2, 3, 1, 3]):
            obs = _obs(step, n_slots)
            t, m = pol.forward_obs_token(DataDict(dev(obs)))
            cache["t"].append(t.squeeze(0)[0]); cache["m"].append(m.squeeze(0)[0])
            ot, om = _pad_cat(cache["t"], cache["m"], lambda xs, d: any_concat(xs, dim=d), lambda xs, d: any_stack(xs, dim=d), z)
            at = None if step == 0 else any_stack(cache["a"], dim=0)[:, None]
            pred = pol.forward(obs_token=ot, action_token=at, prompt_token=p_tok, prompt_token_mask=p_msk, obs_mask=om)
            dist = pol.forward_action_decoder(pred[-1].unsqueeze(0))
            actions = {k: v.mode() for k, v in dist.items()}
            cache["a"].append(pol.forward_action_token(actions).squeeze(0)[0])
            # ---- oracle, same loop ----
            rt, rm = O.forward_obs_token(sd, obs)
            ref["t"].append(rt.squeeze(0)[0]); ref["m"].append(rm.squeeze(0)[0])
            rot, rom = _pad_cat(ref["t"], ref["m"], lambda xs, d: torch.cat(xs, d), lambda xs, d: torch.stack(xs, d), z)
            rat = None if step == 0 else torch.stack(ref["a"], 0)[:, None]
            rpred = O.policy_forward(sd, rot, rom, rat, rp_tok, rp_msk, n_head=cfg["sattn_n_heads"], xattn_n_head=cfg["xattn_n_heads"])
            rlogits = O.action_decoder_logits(sd, rpred[-1].unsqueeze(0))
            ractions = O.action_modes(rlogits)
            ref["a"].append(O.forward_action_token(sd, ractions).squeeze(0)[0])
            assert pred.shape == rpred.shape and torch.equal(om.cpu(), rom)
            assert rel_l2(rpred.numpy(), pred.cpu().numpy()) < 1e-3, step
            for k in actions:
                assert torch.equal(actions[k].cpu(), ractions[k]), (step, k)
            # what example.py does next with the indices
            cont = pol._de_discretize_actions(actions)
            assert all(v.dtype == torch.float32 for v in cont.values())
