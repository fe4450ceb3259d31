"""GPU: a policy step captured into a CUDA graph (vima_b200.graphs.GraphedStep, what bench.py replays by default) returns bit for
bit what the same step returns when every kernel is launched from Python -- for the inputs it was captured with, for new inputs
copied from device tensors, and for new inputs loaded straight from pinned host memory."""
import pytest
import torch

from oracle import synth
from tests.policy_runner import build_policy, to_dev

pytestmark = pytest.mark.gpu


def _pin(x):
    if isinstance(x, dict):
        return {k: _pin(v) for k, v in x.items()}
    return x.pin_memory()


@pytest.mark.parametrize("mode", ["f16x3", "f16f8"])
def test_graph_replay_equals_eager_launches(mode):
    import vima_b200
    from vima_b200.graphs import GraphedStep
    from vima_b200.utils import DataDict

    case = synth.CASES["cfg2_small"]
    vima_b200.set_precision(mode)
    try:
        pol = build_policy(case.model)
        with torch.no_grad():
            tt, wb, ib = synth.make_prompt(case)
            p_tok, p_msk = pol.forward_prompt_assembly((tt, wb.cuda(), DataDict(to_dev(ib, "cuda"))))
            obs_all = synth.make_obs(case)                      # nest of (T, B, ...) tensors
            hist = DataDict(to_dev(synth.slice_obs(obs_all, 0, case.T - 1), "cuda"))
            # two different "newest observation" inputs of the same shape
            new_a, new_b = synth.slice_obs(obs_all, case.T - 1, case.T), synth.slice_obs(obs_all, 0, 1)
            h_tok, h_msk = pol.forward_obs_token(hist)
            a_tok = pol.forward_action_token(to_dev(synth.make_actions(case, case.T), "cuda"))

            def step(obs_dev):
                n_tok, n_msk = pol.forward_obs_token(DataDict(obs_dev))
                pred = pol.forward(obs_token=torch.cat([h_tok, n_tok], 0), obs_mask=torch.cat([h_msk, n_msk], 0), action_token=a_tok,
                                   prompt_token=p_tok, prompt_token_mask=p_msk)
                dists = pol.forward_action_decoder(pred[-1:])
                raw = torch.cat([dists[k].raw_logits for k in dists], dim=-1)
                modes = torch.cat([v.mode() for v in dists.values()], dim=-1)
                return pred[-1:], raw, modes

            dev_a, dev_b = to_dev(new_a, "cuda"), to_dev(new_b, "cuda")
            eager_a = [t.clone() for t in step(dev_a)]
            eager_b = [t.clone() for t in step(dev_b)]
            assert not torch.equal(eager_a[0], eager_b[0])      # the two inputs really differ

            g = GraphedStep(step, dev_a, warmup=2)
            assert g.kernels_per_replay > 50
            for name, inputs, want in (("captured inputs", dev_a, eager_a), ("device inputs", dev_b, eager_b), ("captured inputs again", dev_a, eager_a)):
                got = g(inputs)
                torch.cuda.synchronize()
                for w, x in zip(want, got):
                    assert torch.equal(w, x), name
            got = g(g.load_inputs(_pin(new_b)))                 # pinned host -> static buffers, then replay
            torch.cuda.synchronize()
            for w, x in zip(eager_b, got):
                assert torch.equal(w, x), "pinned host inputs"
    finally:
        vima_b200.set_precision("f16x3")
