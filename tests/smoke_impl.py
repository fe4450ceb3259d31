"""__graft_entry__.smoke(): one tiny policy step on cuda:0 through the C ABI, checked against the CPU oracle."""
import torch


def run_smoke():
    import vima_b200
    from oracle import synth
    from tests.policy_runner import build_policy, run_policy_case
    from tests.test_oracle_golden import run_oracle_case
    from tests.util import rel_l2

    vima_b200.set_precision("f16x3")
    name = "cfg1_t2"
    case = synth.CASES[name]
    pol = build_policy(case.model, "cuda:0")
    r = run_policy_case(pol, case, "cuda:0")
    torch.cuda.synchronize()
    o = run_oracle_case(name)
    for key in ["prompt_tokens", "obs_tokens", "predicted", "logits_raw", "next_action_token"]:
        e = rel_l2(o[key].numpy(), r[key].cpu().numpy())
        assert e < 1e-3, (key, e)
    for k, v in o["modes"].items():
        assert torch.equal(v, r["modes"][k].cpu()), k
    from vima_b200 import _C

    print(f"smoke ok: {name} on {torch.cuda.get_device_name(0)}, kernels launched: {_C.Context.get(0).launches}")
