"""GPU microbenchmark of attention_kernel at the cfg3 decoder shapes (B=256, H=24, d=32, L=263, Lp=256)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vima_b200 import _C
ctx = _C.Context.get(torch.device("cuda", 0))
B, H, D, L, Lp = int(os.environ.get("AB_B", 256)), 24, 32, int(os.environ.get("AB_L", 263)), 256
E = H * D
for split in [int(x) for x in os.environ.get("AB_SPLIT", "0,1").split(",")]:
    for name, Lq, Lk, causal in (("self causal", L, L, True), ("cross", L, Lp, False)):
        mk = lambda r, c: torch.randint(-3000, 3000, (r, c), dtype=torch.int16, device="cuda")
        if causal:
            qkv = mk(B * Lq, 3 * E); qkl = mk(B * Lq, 3 * E) if split else None
            q = (qkv, qkl, 3 * E, 0); k = (qkv, qkl, 3 * E, E); v = (qkv, qkl, 3 * E, 2 * E)
        else:
            qq = mk(B * Lq, E); ql = mk(B * Lq, E) if split else None
            kv = mk(B * Lk, 2 * E); kl = mk(B * Lk, 2 * E) if split else None
            q = (qq, ql, E, 0); k = (kv, kl, 2 * E, 0); v = (kv, kl, 2 * E, E)
        o_hi = torch.empty(B * Lq, E, dtype=torch.int16, device="cuda"); o_lo = torch.empty_like(o_hi) if split else None
        mask = torch.ones(B, Lk, dtype=torch.uint8, device="cuda")
        if causal and os.environ.get("AB_MASKED", "1") == "1":  # the bench workload pads ~10 % of the history's object slots
            mask = (torch.rand(B, Lk, device="cuda") > 0.1).to(torch.uint8)
            mask[:, 0] = 1
        kw = dict(q=q, k=k, v=v, o=(o_hi, o_lo, E, 0), B=B, H=H, Lq=Lq, Lk=Lk, D=D, scale=1 / math.sqrt(D), causal=causal, key_mask=mask, dtype=0)
        ctx.attention(**kw); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ctx.attention(**kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 4.0 * B * H * Lq * Lk * D * (0.5 if causal else 1.0)
        print(f"split={split} {name:12s} {ms:7.3f} ms   {fl/ms/1e9:7.1f} TF/s algorithmic (causal counted as half)", flush=True)
