"""State-dict contract: the hand-written spec equals the real reference (build container only)."""
import pytest


@pytest.mark.reference
@pytest.mark.parametrize("model", ["2M", "20M"])
def test_spec_matches_reference(model):
    from oracle import synth
    from oracle.ref_shim import load_reference
    from oracle.state_dict_spec import state_dict_spec

    ref = load_reference()
    cfg = synth.MODEL_CFGS[model]
    sd = ref.VIMAPolicy(**cfg).state_dict()
    spec = state_dict_spec(**cfg)
    assert list(sd.keys()) == list(spec.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(spec[k]), k
