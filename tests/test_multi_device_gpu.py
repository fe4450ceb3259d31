"""GPU (needs 2 devices; skipped otherwise): a policy living on a device that is NOT the thread's current CUDA device
(the C ABI used torch's current stream of the wrong device and left the thread switched to the context's device)."""
import pytest
import torch

from oracle import synth
from tests.policy_runner import run_policy_case

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_policy_on_non_current_device_matches_device0():
    import vima_b200
    from oracle import detgen

    vima_b200.set_precision("f16x3")
    case = synth.CASES["ragged_4M"]
    outs = {}
    for dev in ("cuda:0", "cuda:1"):
        torch.cuda.set_device(0)  # the thread's current device stays 0 throughout
        pol = vima_b200.VIMAPolicy(**synth.MODEL_CFGS[case.model])
        detgen.fill_module_(pol)
        pol = pol.to(dev).eval()
        r = run_policy_case(pol, case, dev)
        torch.cuda.synchronize(dev)
        assert torch.cuda.current_device() == 0, "a C-ABI call left the thread on another device"
        assert r["predicted"].device == torch.device(dev)
        outs[dev] = {k: (v.cpu() if torch.is_tensor(v) else {kk: vv.cpu() for kk, vv in v.items()}) for k, v in r.items() if v is not None}
    for k in ("prompt_tokens", "obs_tokens", "predicted", "logits_raw"):
        assert torch.equal(outs["cuda:0"][k], outs["cuda:1"][k]), k
    for k, v in outs["cuda:0"]["modes"].items():
        assert torch.equal(v, outs["cuda:1"]["modes"][k])
    # side stream on the non-current device: kernels must be enqueued on THAT device's current stream
    with torch.cuda.device(0):
        s = torch.cuda.Stream(device="cuda:1")
        with torch.cuda.stream(s):
            pol = vima_b200.VIMAPolicy(**synth.MODEL_CFGS[case.model])
            detgen.fill_module_(pol)
            pol = pol.to("cuda:1").eval()
            r2 = run_policy_case(pol, case, "cuda:1")
        s.synchronize()
        assert torch.equal(r2["predicted"].cpu(), outs["cuda:1"]["predicted"])
