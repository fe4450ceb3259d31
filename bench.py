#!/usr/bin/env python
"""bench.py -- policy steps/sec of the VIMA policy forward pass on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg3] [--precision f16x3]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one policy step for every episode of the batch, exactly as scripts/example.py chains the policy's
public methods with a full-history re-forward:
    forward_obs_token(new obs) -> forward(T obs steps, T-1 actions, prompt) -> forward_action_decoder(last row)
    -> .mode() -> forward_action_token(action)                       [prompt encode is once per episode: untimed, reported]
Workload = BASELINE.json configs[2]: VIMA-200M, 256 episodes per GPU, Q=32 object tokens, Lp=256, T=8 (L=263).
`value`   : device-timed (CUDA events), inputs resident in HBM.
`e2e`     : same step through the same public methods but the new observation comes from pinned HOST memory
            (H2D inside the timed region) and the action indices are read back to the host every step.
`--impl reference`: the CPU oracle port of the reference (oracle/vima_oracle.py, torch fp32, all host threads)
            on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "policy steps/sec (batched action decode)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--precision", default="f16f8")
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU (default: the workload's)")
    ap.add_argument("--cpu-episodes", type=int, default=4, help="episodes in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast-modes", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------------
def workload_case(name: str, batch: int):
    from dataclasses import replace

    from oracle import synth

    case = synth.CASES[name]
    if batch:
        case = replace(case, B=batch)
    return case


def algorithmic_flops_per_episode_step(case):
    """decoder + object encoder (new obs) + heads, per episode-step."""
    from oracle import synth

    cfg = synth.MODEL_CFGS[case.model]
    E, nl = cfg["embed_dim"], cfg["xf_n_layers"]
    L, Lp, Q = case.L, case.Lp, case.Q
    dec = nl * (60 * L * E * E + 4 * Lp * E * E + 4 * L * Lp * E + 4 * L * L * E)
    obj = Q * (0.286e9 + 2.4e6 + 2 * 1536 * E)
    heads = 12 * 2 * (512 * E + 512 * 512) + 2 * 512 * 700
    return dec + obj + heads


def sample_clocks(stop_evt, out):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    dev = os.environ.get("LOCAL_RANK", "0")
    try:
        pr = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", dev],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    out["proc"] = pr

    def rd():
        for line in pr.stdout:
            out.setdefault("lines", []).append(line.strip())

    t = threading.Thread(target=rd, daemon=True)
    t.start()
    stop_evt.wait()
    pr.terminate()
    t.join(timeout=2)


def summarise_clocks(lines):
    sm, mx, reasons = [], [], set()
    for ln in lines or []:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0])); mx.append(float(f[1]))
        except ValueError:
            continue
        for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_oracle_steps(case_name: str, n_episodes: int, reps: int):
    """Times `reps` policy steps of `n_episodes` episodes on the CPU oracle. Returns (steps/sec, cores, seconds list)."""
    from dataclasses import replace

    from oracle import detgen, synth, vima_oracle as O
    from oracle.state_dict_spec import state_dict_spec

    case = replace(synth.CASES[case_name], B=n_episodes)
    cfg = synth.MODEL_CFGS[case.model]
    skip = ("t5_prompt_encoder", "prompt_embedding", "prompt_obj_post_layer")  # prompt encode is outside the step
    sd = {}
    for k, shape in state_dict_spec(**cfg).items():
        if k.startswith(skip):
            continue
        w = detgen.weight_for(k, shape)
        if w is not None:
            sd[k] = w
    E = cfg["embed_dim"]
    with torch.no_grad():
        prompt_tokens = detgen.uniform("bench.prompt", (case.Lp, case.B, E))
        prompt_masks = torch.ones(case.B, case.Lp, dtype=torch.bool)
        hist = synth.make_obs(case, T=case.T - 1, tag="hist")
        h_tok, h_msk = O.forward_obs_token(sd, hist)
        a_tok = O.forward_action_token(sd, synth.make_actions(case, case.T))
        new_obs = synth.make_obs(case, T=1, tag="new")

        def one_step():
            t0 = time.perf_counter()
            O.policy_step(sd, obs=new_obs, history_obs_tokens=h_tok, history_obs_masks=h_msk, history_action_tokens=a_tok,
                          prompt_tokens=prompt_tokens, prompt_masks=prompt_masks, n_head=cfg["sattn_n_heads"], xattn_n_head=cfg["xattn_n_heads"])
            return time.perf_counter() - t0

        # "all the host threads it can use": torch's CPU kernels stop scaling (and collapse when oversubscribed) well
        # before 100+ threads, so pick the fastest thread count among powers of two up to the core count.
        ncpu = os.cpu_count() or 1
        cands = sorted({c for c in (8, 16, 32, 64, 128, 256) if c <= ncpu} | {min(ncpu, 8)})
        best_t, best_n = None, cands[0]
        for c in cands:
            torch.set_num_threads(c)
            one_step()
            dt = one_step()
            if best_t is None or dt < best_t:
                best_t, best_n = dt, c
            elif dt > 3 * best_t:
                break
        torch.set_num_threads(best_n)
        times = [one_step() for _ in range(reps)]
    med = statistics.median(times)
    return n_episodes / med, torch.get_num_threads(), times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    case = workload_case(args.workload, args.batch)
    n_ep = args.cpu_episodes
    t0 = time.perf_counter()
    val, cores, times = cpu_oracle_steps(args.workload, n_ep, max(args.steps, 1) + max(args.warmup - 1, 0))
    times = times[-max(args.steps, 1):]
    ms = statistics.median(times) * 1e3
    val = n_ep / (ms / 1e3)
    sample = f"{n_ep} episodes x {len(times)} timed steps of {args.workload} (L={case.L}, Lp={case.Lp}) on the CPU oracle port"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: VIMA-{case.model} policy step, {n_ep} episodes (bounded CPU sample), Q={case.Q}, Lp={case.Lp}, T={case.T}",
                   "parallelism": "cpu"},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def to_dev(x, dev, non_blocking=False):
    if isinstance(x, dict):
        return {k: to_dev(v, dev, non_blocking) for k, v in x.items()}
    return x.to(dev, non_blocking=non_blocking)


def pin(x):
    if isinstance(x, dict):
        return {k: pin(v) for k, v in x.items()}
    return x.pin_memory()


def nbytes(x):
    if isinstance(x, dict):
        return sum(nbytes(v) for v in x.values())
    return x.numel() * x.element_size()


class GemmTimer:
    """Wraps Context.gemm with CUDA events on the launching stream: per-launch durations + algorithmic FLOPs."""

    def __init__(self, ctx):
        self.ctx, self.orig, self.rec, self.on = ctx, ctx.gemm, [], False

        def timed(**kw):
            if not self.on:
                return self.orig(**kw)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            self.orig(**kw)
            e1.record()
            self.rec.append((e0, e1, 2.0 * kw["M"] * kw["N"] * kw["K"], (kw["M"], kw["N"], kw["K"])))

        ctx.gemm = timed

    def summary(self):
        tot_ms = sum(a.elapsed_time(b) for a, b, _, _ in self.rec)
        tot_fl = sum(f for _, _, f, _ in self.rec)
        return tot_ms, tot_fl, len(self.rec)


def run_ours(args):
    import torch.distributed as dist

    import vima_b200
    from oracle import detgen, synth
    from vima_b200 import _C
    from vima_b200.utils import DataDict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vima_b200.set_precision(args.precision)
    case = workload_case(args.workload, args.batch)
    case = type(case)(**{**case.__dict__, "seed": case.seed + rank})  # independent episodes per rank
    cfg = synth.MODEL_CFGS[case.model]
    t_setup = time.perf_counter()
    policy = vima_b200.VIMAPolicy(**cfg)
    detgen.fill_module_(policy)
    policy = policy.to(dev).eval()
    ctx = _C.Context.get(dev)
    B, T, Q, E = case.B, case.T, case.Q, cfg["embed_dim"]

    with torch.no_grad():
        # ---- once per episode: prompt encode (untimed here, reported) ----
        prompt = synth.make_prompt(case)
        pr_in = (prompt[0], prompt[1].to(dev), DataDict(to_dev(prompt[2], dev)))
        policy.forward_prompt_assembly(pr_in)  # warm (weight packing)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        prompt_tokens, prompt_masks = policy.forward_prompt_assembly(pr_in)
        e1.record(); torch.cuda.synchronize()
        prompt_ms = e0.elapsed_time(e1)
        # ---- history cache, as example.py keeps it: tokens of the T-1 earlier obs steps and actions ----
        hist = synth.make_obs(case, T=T - 1, tag="hist")
        h_tok, h_msk = policy.forward_obs_token(DataDict(to_dev(hist, dev)))
        a_tok = policy.forward_action_token(to_dev(synth.make_actions(case, T), dev))
        new_obs_host = pin(synth.make_obs(case, T=1, tag="new"))
        new_obs_dev = to_dev(new_obs_host, dev)
        gathered = torch.empty((world * B, 700), dtype=torch.float32, device=dev) if world > 1 else None

        def step(obs_dev):
            n_tok, n_msk = policy.forward_obs_token(DataDict(obs_dev))
            obs_tok = torch.cat([h_tok, n_tok], dim=0)
            obs_msk = torch.cat([h_msk, n_msk], dim=0)
            pred = policy.forward(obs_token=obs_tok, obs_mask=obs_msk, action_token=a_tok, prompt_token=prompt_tokens,
                                  prompt_token_mask=prompt_masks)
            dists = policy.forward_action_decoder(pred[-1:])
            if world > 1:  # the path's one exchange: all-gather of the action logits over NVLink
                logits = torch.cat([dists[k].raw_logits for k in dists], dim=-1).reshape(B, 700).contiguous()
                dist.all_gather_into_tensor(gathered, logits)
            modes = {k: v.mode() for k, v in dists.items()}
            nxt = policy.forward_action_token(modes)
            return modes, nxt

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        gt = GemmTimer(ctx)
        for _ in range(max(args.warmup, 3)):
            step(new_obs_dev)
        barrier()
        setup_s = time.perf_counter() - t_setup

        # ---- timed: device-resident inputs ----
        stop = threading.Event(); clk = {}
        th = threading.Thread(target=sample_clocks, args=(stop, clk), daemon=True); th.start()
        time.sleep(0.3)
        launches0 = ctx.launches
        gt.on = True
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.nvtx.range_push("timed")
        e0.record()
        for _ in range(args.steps):
            step(new_obs_dev)
        e1.record()
        barrier()
        torch.cuda.nvtx.range_pop()
        gt.on = False
        ms_total = e0.elapsed_time(e1)
        launches = ctx.launches - launches0
        gemm_ms, gemm_fl, n_gemm = gt.summary()

        # ---- timed: end to end (pinned host obs -> device, action indices -> host, every step) ----
        h2d = nbytes(new_obs_host)
        barrier()
        t0 = time.perf_counter()
        e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
        e2.record()
        d2h = 0
        for _ in range(args.steps):
            obs_d = to_dev(new_obs_host, dev, non_blocking=True)
            modes, _ = step(obs_d)
            host_modes = {k: v.cpu() for k, v in modes.items()}  # blocking read of the step's result
            d2h = sum(v.numel() * v.element_size() for v in host_modes.values())
        e3.record()
        barrier()
        e2e_ms_total = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3 * 0.0)
        stop.set(); th.join(timeout=3)

        # ---- separately reported: the same T-step episode decoded step by step through the K/V cache ----
        all_tok = torch.cat([h_tok, policy.forward_obs_token(DataDict(new_obs_dev))[0]], dim=0)
        all_msk = torch.cat([h_msk, policy.forward_obs_token(DataDict(new_obs_dev))[1]], dim=0)
        incr_ms = None
        for rep in range(2):  # first repetition warms the allocator
            barrier()
            e4 = torch.cuda.Event(enable_timing=True); e5 = torch.cuda.Event(enable_timing=True)
            e4.record()
            cache = policy.start_decode(prompt_tokens, prompt_masks, max_tokens=T * (Q + 1) - 1)
            for t in range(T):
                pred_t = policy.forward_step(cache, all_tok[t:t + 1], all_msk[t:t + 1], None if t == 0 else a_tok[t - 1:t])
                policy.forward_action_token({k: v.mode() for k, v in policy.forward_action_decoder(pred_t).items()})
            e5.record(); barrier()
            incr_ms = e4.elapsed_time(e5)
            del cache

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_total = maxr(ms_total)
    e2e_ms_total = maxr(e2e_ms_total)
    incr_ms = maxr(incr_ms)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms_total / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0  # the GEMMs run inside a long step -> sustained figure
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "B200_PROFILING.md fallback 1.4 PF sustained (of fallback)"
        achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        step_flops = algorithmic_flops_per_episode_step(case) * B
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"f16x3": "f16 hi/lo operand pairs (3-term products, fp32-equivalent), fp32 accumulate/softmax/LN",
                      "bf16x3": "bf16 hi/lo operand pairs (3-term), fp32 accumulate", "f16": "f16 operands, fp32 accumulate",
                      "bf16": "bf16 operands, fp32 accumulate",
                      "f16f8": "f16 hi*hi + e4m3 cross terms (decoder GEMMs; 2 tensor pass-equivalents), f16 3-term elsewhere, fp32 accumulate/softmax/LN"}[args.precision],
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: VIMA-{case.model} policy step (full-history re-forward), {B} episodes/GPU, Q={Q} object tokens, "
                                   f"Lp={case.Lp} prompt tokens, T={T}-step history (L={case.L})",
                       "global_batch": world * B, "parallelism": f"dp{world}", "precision_mode": args.precision,
                       "l2": "inputs larger than L2: >4 GB of activations + 1.6 GB of packed weights stream per step (L2 = 126 MB)",
                       "prompt_len_note": "Lp=256 is BASELINE.md section 4 row #3 / SURVEY 8(d) #3: the reference VIMAPolicy caps prompts at "
                                          "xattn_n_positions=256 (vima_policy.py:26-38), longer prompts raise in the reference itself",
                       "prompt_encode_ms_per_batch": prompt_ms, "setup_s": setup_s,
                       "steps_per_s_with_prompt_amortised": world * B * T / ((prompt_ms + T * ms_step) / 1e3)},
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_total / args.steps},
            "gpu_launches": int(launches),
            "clocks": summarise_clocks(clk.get("lines")),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": 1.026e9 if args.precision == "f16f8" else (1.010e9 if args.precision == "f16x3" else None),
                         "traffic_note": "ncu dram read+write bytes of one c_fc||gated_layer launch (M=67840,N=6144,K=768), profiles/r1_summary.md sections 2/2b; algorithmic 1.05-1.06e9",
                         "kernel": "gemm_tc_kernel (tcgen05)", "launches_per_step": n_gemm / args.steps, "share_of_step": gemm_ms / ms_total,
                         "peak_source": peak_src,
                         "note": ("algorithmic FLOPs 2MNK per launch; in *x3 modes every product is issued as 3 tensor-core passes "
                                  "(f16f8: 1 fp16 pass + 2 fp8 passes), so tensor-pipe work is 3x (2x) the algorithmic figure"),
                         "step_algorithmic_tflop": step_flops / 1e12, "step_tflops": step_flops / (ms_step / 1e3) / 1e12},
        }
        line["incremental"] = {"value": world * B * T / (incr_ms / 1e3), "unit": "env steps/s", "episode_ms": incr_ms,
                               "note": f"not the graded metric: {T}-step episode decoded through the K/V cache (start_decode/forward_step, "
                                       "decoder + heads + action embed per step; obs tokens precomputed); same predictions as the full re-forward"}
        if not args.no_cpu_baseline and world == 1:
            v, cores, times = cpu_oracle_steps(args.workload, args.cpu_episodes, 3)
            line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_episodes} episodes x 3 timed steps of {args.workload} on the CPU oracle port (median)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
