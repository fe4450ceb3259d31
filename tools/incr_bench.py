"""Step-by-step decode through the K/V cache (VIMAPolicy.start_decode / forward_step) vs the reference-shaped full
re-forward of the history, on the cfg3 workload shape: per environment step t the full path pushes t*(Q+1)-1 tokens per
episode through the decoder, the cached path Q+1.  Prints ms per environment step for both (decoder only; CUDA events)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vima_b200
from oracle import synth  # shapes only


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="200M")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--n-obj", type=int, default=32)
    ap.add_argument("--prompt-len", type=int, default=256)
    ap.add_argument("--precision", default="f16f8")
    a = ap.parse_args()
    vima_b200.set_precision(a.precision)
    pol = vima_b200.VIMAPolicy(**synth.MODEL_CFGS[a.model]).cuda().eval()
    E, B, T, Q, Lp = pol.embed_dim, a.batch, a.steps, a.n_obj, a.prompt_len
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.randn(T, B, Q, E, device="cuda", generator=g)
    msk = torch.rand(T, B, Q, device="cuda", generator=g) > 0.1
    msk[:, :, 0] = True
    act = torch.randn(T - 1, B, E, device="cuda", generator=g)
    ptk = torch.randn(Lp, B, E, device="cuda", generator=g)
    pmk = torch.ones(B, Lp, dtype=torch.bool, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for rep in range(2):  # first repetition warms up
            cache = pol.start_decode(ptk, pmk, max_tokens=T * (Q + 1) - 1)
            t_inc, t_full = [], []
            for t in range(T):
                e0, e1, e2 = ev(), ev(), ev()
                e0.record()
                s = pol.forward_step(cache, obs[t:t + 1], msk[t:t + 1], None if t == 0 else act[t - 1:t])
                e1.record()
                f = pol.forward(obs_token=obs[:t + 1], obs_mask=msk[:t + 1], action_token=None if t == 0 else act[:t], prompt_token=ptk,
                                prompt_token_mask=pmk)[-1:]
                e2.record()
                torch.cuda.synchronize()
                t_inc.append(e0.elapsed_time(e1)); t_full.append(e1.elapsed_time(e2))
                d = ((s - f).norm() / f.norm()).item()
                assert d < 1e-5, d
        for t in range(T):
            print(f"env step {t}: cached {t_inc[t]:8.2f} ms   full re-forward {t_full[t]:8.2f} ms   x{t_full[t] / t_inc[t]:.1f}")
        print(f"episode of {T} steps, {B} episodes: cached {sum(t_inc):.1f} ms, full {sum(t_full):.1f} ms  "
              f"({B * T / sum(t_inc) * 1e3:.0f} vs {B * T / sum(t_full) * 1e3:.0f} env-steps/s, decoder only)")


if __name__ == "__main__":
    main()
