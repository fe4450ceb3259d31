#!/usr/bin/env python
"""Which tensor-core operand format keeps the decoder within 1e-3 rel of the fp32 reference?

Emulates reduced-precision GEMM multiplicands inside the CPU oracle (fp32 accumulate, fp32 softmax/LN/
residual) on the 200M decoder at L=263, Lp=256 and reports rel-L2 of predicted tokens and logits.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import detgen, synth, vima_oracle as O
from oracle.state_dict_spec import xattn_gpt_spec, mlp_spec, ACTION_DIMS

def rnd_tf32(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF   # round-to-nearest (ties away) to 10 explicit mantissa bits
    return i.view(torch.float32)
def trunc_tf32(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
MODES = {
    "fp16": lambda x: x.to(torch.float16).float(),
    "bf16": lambda x: x.to(torch.bfloat16).float(),
    "tf32_rn": rnd_tf32,
    "tf32_trunc": trunc_tf32,
}
def make_split_mm(dt, terms):
    def split(x):
        hi = x.to(dt).float(); lo = (x - hi).to(dt).float(); return hi, lo
    def mm(x, w_t):
        xh, xl = split(x); wh, wl = split(w_t)
        y = xh @ wh
        if terms >= 2: y = y + xl @ wh
        if terms >= 3: y = y + xh @ wl
        return y
    return mm
SPLIT = {"fp16x3": (torch.float16, 3), "bf16x3": (torch.bfloat16, 3), "fp16x2(act only)": (torch.float16, 2)}
def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "200M"
    cfg = synth.MODEL_CFGS[model]; E = cfg["embed_dim"]; nl = cfg["xf_n_layers"]; H = cfg["sattn_n_heads"]
    sd = {}
    spec = dict(xattn_gpt_spec("xattn_gpt.", E, nl))
    for k, dims in ACTION_DIMS.items():
        for j, n in enumerate(dims):
            spec.update(mlp_spec(f"action_decoder._decoders.{k}.mlps.{j}.", [E, 512, 512, n]))
    for k, s in spec.items():
        w = detgen.weight_for(k, s)
        if w is not None: sd[k] = w
    B, T, Q, Lp = 2, 8, 32, 256
    obs = detgen.uniform("ps.obs", (T, B, Q, E)) * 1.7
    act = detgen.uniform("ps.act", (T - 1, B, E)) * 1.7
    pr = detgen.uniform("ps.prompt", (Lp, B, E)) * 1.7
    om = detgen.randint("ps.om", (T, B, Q), 0, 4) > 0; om[:, :, 0] = True
    pm = torch.ones(B, Lp, dtype=torch.bool); pm[1, 200:] = False
    torch.set_num_threads(os.cpu_count())
    def run():
        with torch.no_grad():
            p = O.policy_forward(sd, obs, om, act, pr, pm, n_head=H, xattn_n_head=H)
            return p, O.action_decoder_logits(sd, p[-1:])
    O.set_operand_rounding(None)
    p0, l0 = run()
    for name, fn in MODES.items():
        O.set_operand_rounding(fn)
        p, l = run()
        rp = ((p - p0).norm() / p0.norm()).item(); rl = ((l - l0).norm() / l0.norm()).item()
        mx = ((p - p0).abs().max() / p0.abs().max()).item()
        am = (O.action_modes(l)["pose0_rotation"] == O.action_modes(l0)["pose0_rotation"]).float().mean().item()
        print(f"{model} {name:11s} predicted rel-L2 {rp:.2e} (max-abs/max {mx:.2e})  logits rel-L2 {rl:.2e}  rot-idx match {am:.2f}")
    O.set_operand_rounding(None)
    orig = O._mm
    for name, (dt, terms) in SPLIT.items():
        O._mm = make_split_mm(dt, terms)
        p, l = run()
        rp = ((p - p0).norm() / p0.norm()).item(); rl = ((l - l0).norm() / l0.norm()).item()
        print(f"{model} {name:11s} predicted rel-L2 {rp:.2e}  logits rel-L2 {rl:.2e}")
    O._mm = orig
if not (len(sys.argv) > 2): main()

def attention_only():
    """GEMMs exact; only QK^T / PV operands rounded (what an fp16 / bf16 tensor-core attention would do)."""
    import torch as T
    model = sys.argv[1] if len(sys.argv) > 1 else "200M"
    real = T.matmul
    for name, dt in (("attn-fp16", T.float16), ("attn-bf16", T.bfloat16)):
        T.matmul = lambda a, b: real(a.to(dt).float(), b.to(dt).float())
        yield name
    T.matmul = real


def ftz_study():
    """fp16 split products when the tensor core flushes fp16 subnormal inputs, with a power-of-two activation pre-scale."""
    import torch
    model = "200M"
    cfg = synth.MODEL_CFGS[model]; E = cfg["embed_dim"]; nl = cfg["xf_n_layers"]; H = cfg["sattn_n_heads"]
    sd = {k: detgen.weight_for(k, s) for k, s in xattn_gpt_spec("xattn_gpt.", E, nl).items() if detgen.weight_for(k, s) is not None}
    B, T, Q, Lp = 2, 8, 32, 256
    obs = detgen.uniform("ps.obs", (T, B, Q, E)) * 1.7
    act = detgen.uniform("ps.act", (T - 1, B, E)) * 1.7
    pr = detgen.uniform("ps.prompt", (Lp, B, E)) * 1.7
    om = detgen.randint("ps.om", (T, B, Q), 0, 4) > 0; om[:, :, 0] = True
    pm = torch.ones(B, Lp, dtype=torch.bool); pm[1, 200:] = False
    torch.set_num_threads(os.cpu_count())
    def run():
        with torch.no_grad():
            return O.policy_forward(sd, obs, om, act, pr, pm, n_head=H, xattn_n_head=H)
    p0 = run()
    def ftz(h):  # flush fp16 subnormals
        return torch.where(h.abs() < 6.103515625e-05, torch.zeros_like(h), h)
    def mk(act_scale, flush):
        def split(x, s):
            xs = x * s
            hi = xs.to(torch.float16).float(); lo = (xs - hi).to(torch.float16).float()
            if flush: hi, lo = ftz(hi), ftz(lo)
            return hi, lo
        def mm(x, w_t):
            ws = 2.0 ** torch.floor(torch.log2(1024.0 / w_t.abs().max()))
            xh, xl = split(x, act_scale); wh, wl = split(w_t, ws)
            return (xh @ wh + xl @ wh + xh @ wl) / (act_scale * ws)
        return mm
    orig = O._mm
    for s, fl in [(1, False), (1, True), (16, True), (64, True), (256, True)]:
        O._mm = mk(float(s), fl)
        p = run()
        print(f"fp16x3 act_scale={s:4d} ftz={fl}: predicted rel-L2 {((p - p0).norm() / p0.norm()).item():.2e}")
    O._mm = orig

if len(sys.argv) > 2 and sys.argv[2] == "ftz":
    ftz_study()
