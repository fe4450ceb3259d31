"""XAttnGPT: cross-attention to the prompt alternating with causal self-attention over the obs/action history.

Module surface and state-dict keys of /root/reference/vima/nn/seq_modeling/xattn_gpt/{xattn_gpt.py:13-177,
components.py:14-263}; the arithmetic runs on the sm_100a kernels (tcgen05 GEMMs with fused bias / GELU / GEGLU /
residual epilogues, fused masked attention, warp-shuffle LayerNorm).  Per layer (reference order, xattn_gpt.py:123-132):

    XAttention (pre-LN, bias-free, components.py:158-228)        Block (GPT-1 post-LN, components.py:23-37)
      q  = Wq LN(x)            k,v = Wkv (prompt + pos)             qkv = c_attn(x)
      a  = Wo attn(q,k,v) + x          [+ row sums of a]            s   = c_proj(causal_attn(qkv)) + x   [+ row sums of s]
      h  = gelu(W1 LN2(a)) * (Wg a)    [ONE GEGLU GEMM over a:      h   = gelu(c_fc LN1(s)) * (Wg LN1(s))  [one GEGLU GEMM over s,
           LN2 folded into W1, the gate reads UN-normalised a]            LN1 folded into both halves]
      x' = W2 h + a                                                 x'' = LN2(c_proj h + LN1(s))   [LN1(s) rebuilt in the epilogue]

LayerNorm folding: W (gamma (x - mean) rstd + beta) = rstd ((W gamma) x - mean rowsum(W gamma)) + W beta, so the GEMM
runs on the un-normalised rows and its epilogue applies (mean, rstd); the producing GEMM's epilogue emits the row sums.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from .. import _C
from .. import engine as eng
from .basic import Conv1D


class _SelfAttention(nn.Module):
    """Parameter holder for HF openai `Attention` + the reference's persistent causal `bias` buffer (components.py:40-49)."""

    def __init__(self, nx: int, n_positions: int):
        super().__init__()
        self.register_buffer("bias", torch.tril(torch.ones(n_positions, n_positions)).view(1, 1, n_positions, n_positions), persistent=True)
        self.c_attn = Conv1D(3 * nx, nx)
        self.c_proj = Conv1D(nx, nx)


class _MLP(nn.Module):
    def __init__(self, nx: int, geglu: bool):
        super().__init__()
        self.c_fc = Conv1D(4 * nx, nx)
        self.c_proj = Conv1D(nx, 4 * nx)
        self.gated_layer = nn.Linear(nx, 4 * nx, bias=False) if geglu else None


class Block(nn.Module):
    def __init__(self, nx: int, n_positions: int, n_head: int, geglu: bool, eps: float = 1e-5):
        super().__init__()
        self.n_head = n_head
        self.attn = _SelfAttention(nx, n_positions)
        self.ln_1 = nn.LayerNorm(nx, eps=eps)
        self.mlp = _MLP(nx, geglu)
        self.ln_2 = nn.LayerNorm(nx, eps=eps)


def pack_block(ctx, blk: "Block", p) -> dict:
    """Packed weights of one GPT block.  ln_1 is folded into c_fc || gated_layer (both read ln_1(s), components.py:31-36); without
    GEGLU (`afn != "geglu"`, components.py:92-94) into c_fc alone."""
    ln1 = (blk.ln_1.weight.detach(), blk.ln_1.bias.detach())
    d = {
        "c_attn": eng.pack_linear(ctx, blk.attn.c_attn.weight, blk.attn.c_attn.bias, transposed=True, p=p, f8=True),
        "c_proj": eng.pack_linear(ctx, blk.attn.c_proj.weight, blk.attn.c_proj.bias, transposed=True, p=p, f8=True),
        "mlp_proj": eng.pack_linear(ctx, blk.mlp.c_proj.weight, blk.mlp.c_proj.bias, transposed=True, p=p, f8=True),
    }
    if blk.mlp.gated_layer is not None:
        d["fc_glu"] = eng.pack_glu(ctx, blk.mlp.c_fc.weight, blk.mlp.c_fc.bias, blk.mlp.gated_layer.weight, val_transposed=True,
                                   gate_transposed=False, p=p, f8=True, ln=ln1, ln_gate=True)
    else:
        d["fc"] = eng.pack_linear(ctx, blk.mlp.c_fc.weight, blk.mlp.c_fc.bias, transposed=True, p=p, f8=True, ln=ln1)
    return d


class DecodeCache:
    """Per-layer K/V cache of one batch of episodes for step-by-step decode.

    The reference re-runs the whole history every environment step (scripts/example.py:139-171); with the cache a step
    only pushes its Q+1 new tokens through the stack: the prompt's key/value projections are computed once, and every
    causal block appends the new tokens' keys/values to `kv[i]` ([B*Lmax, 2E] (hi, lo) pairs) and attends over the
    cached prefix (`vima_attention` with kv_batch_rows / mask_ld / q_pos0)."""

    def __init__(self, *, B: int, Lmax: int, E: int, n_layer: int, device, split: bool, precision: str = ""):
        self.B, self.Lmax, self.E, self.L = B, Lmax, E, 0
        self.precision = precision  # the operand format the K/V rows are stored in; forward_step refuses any other mode
        mk = lambda: torch.zeros((B * Lmax, 2 * E), dtype=torch.int16, device=device)
        self.kv_hi = [mk() for _ in range(n_layer)]
        self.kv_lo = [mk() if split else None for _ in range(n_layer)]
        self.mask = torch.zeros((B, Lmax), dtype=torch.uint8, device=device)
        self.n_valid = torch.zeros((B,), dtype=torch.int64, device=device)  # valid tokens so far -> next position id
        self.prompt_kv = None  # per-layer projected prompt keys/values (filled by the first step)

    def append_kv(self, i: int, qkv16, L0: int, Ln: int):
        E, B = self.E, self.B
        self.kv_hi[i].view(B, self.Lmax, 2 * E)[:, L0:L0 + Ln].copy_(qkv16.hi.view(B, Ln, -1)[:, :, E:3 * E])
        if self.kv_lo[i] is not None:
            self.kv_lo[i].view(B, self.Lmax, 2 * E)[:, L0:L0 + Ln].copy_(qkv16.lo.view(B, Ln, -1)[:, :, E:3 * E])


def check_cache_append(cache: "DecodeCache", B: int, L: int, E: int, p) -> None:
    """Everything that can refuse an append, BEFORE any cache state is touched (capacity, shapes, precision mode)."""
    if cache.B != B or cache.E != E:
        raise ValueError(f"DecodeCache(B={cache.B}, E={cache.E}) does not match the step's batch {B} / width {E}")
    if cache.L + L > cache.Lmax:
        raise ValueError(f"DecodeCache(B={cache.B}, Lmax={cache.Lmax}) cannot take {L} more tokens at length {cache.L}")
    if cache.precision and cache.precision != p.name:
        raise ValueError(f"DecodeCache was opened in precision mode {cache.precision!r}; the current mode is {p.name!r}")


def run_block(ctx, p, W, blk: "Block", x32, x16, c16, *, B, L, E, H, omask, chain_ln=None, want16=False, out_f32=None, cache=None,
              layer=0):
    """GPT-1 post-LN block (components.py:23-37 / gpt.py:223-249): returns (LN2 output fp32, operands of the NEXT consumer):
    with `chain_ln` the operands are chain_ln(LN2(...)) (next layer's query LayerNorm), with `want16` they are LN2(...) itself.
    With `cache` the L rows are the NEW tokens of each episode and attention runs over the cached prefix + themselves.

    ln_1 never runs as a kernel: c_proj's epilogue emits s = attn + x as fp32 + operands together with per-row partial sums, the
    GEGLU GEMM takes the un-normalised s with ln_1 folded into its weights (rstd / mean applied in its epilogue), and the MLP's
    c_proj normalises its residual ln_1(s) on the fly from the same (mean, rstd)."""
    M = B * L
    d = E // H
    dev = x32.device
    # operand formats: attention inputs keep the 16-bit (hi, lo) pair; everything that only feeds a GEMM carries e4m3
    # cross-term views in "f16f8" mode (out_f8=True is a no-op in the other modes)
    _, qkv16 = eng.gemm(ctx, x16, W["c_attn"], p, want16=True)
    o8 = None if c16.lo8 is None else (c16.lo8, c16.hi8)
    if cache is None:
        ctx.attention(q=(qkv16.hi, qkv16.lo, qkv16.ld, 0), k=(qkv16.hi, qkv16.lo, qkv16.ld, E), v=(qkv16.hi, qkv16.lo, qkv16.ld, 2 * E),
                      o=(c16.hi, c16.lo, c16.ld, 0), B=B, H=H, Lq=L, Lk=L, D=d, scale=1.0 / math.sqrt(d), causal=True, key_mask=omask,
                      dtype=p.dtype, o8=o8)
    else:
        L0 = cache.L
        cache.append_kv(layer, qkv16, L0, L)
        khi, klo = cache.kv_hi[layer], cache.kv_lo[layer]
        ctx.attention(q=(qkv16.hi, qkv16.lo, qkv16.ld, 0), k=(khi, klo, 2 * E, 0), v=(khi, klo, 2 * E, E), o=(c16.hi, c16.lo, c16.ld, 0),
                      B=B, H=H, Lq=L, Lk=L0 + L, D=d, scale=1.0 / math.sqrt(d), causal=True, key_mask=cache.mask, dtype=p.dtype, o8=o8,
                      kv_batch_rows=cache.Lmax, mask_ld=cache.Lmax, q_pos0=L0)
    part = eng.stats_buffer(ctx, M, W["c_proj"], dev)
    s32, s16 = eng.gemm(ctx, c16, W["c_proj"], p, residual=x32, want_f32=True, want16=True, out_f8=True, stats_out=part)
    st = eng.row_stats_of(ctx, part, M, E, blk.ln_1.eps)
    if "fc_glu" in W:
        _, h16 = eng.gemm(ctx, s16, W["fc_glu"], p, act=_C.ACT_GELU, want16=True, out_f8=True, row_stats=st)
    else:  # afn = "gelu" (HF NewGELUActivation, the OpenAIGPTConfig default): act(c_fc(ln_1(s))), components.py:92-98
        _, h16 = eng.gemm(ctx, s16, W["fc"], p, act=_C.ACT_GELU_TANH, want16=True, out_f8=True, row_stats=st)
    t32, _ = eng.gemm(ctx, h16, W["mlp_proj"], p, residual=s32, res_ln=(st, blk.ln_1.weight.detach(), blk.ln_1.bias.detach()), want_f32=True)
    del h16, s32, s16
    w, b = blk.ln_2.weight.detach(), blk.ln_2.bias.detach()
    if chain_ln is not None:
        y32, _, nxt16 = eng.norm(ctx, t32, p, rows=M, cols=E, w=w, b=b, eps=blk.ln_2.eps, w2=chain_ln.weight.detach(), b2=chain_ln.bias.detach(),
                                 eps2=chain_ln.eps, want16=True, out_f32=out_f32, want_f32=out_f32 is None, out_f8=True)
    else:
        y32, _, nxt16 = eng.norm(ctx, t32, p, rows=M, cols=E, w=w, b=b, eps=blk.ln_2.eps, want16=want16, out_f32=out_f32, want_f32=out_f32 is None,
                                 out_f8=True)
    return y32, nxt16


class PosIdGuard:
    """Deferred report of out-of-range position ids (the reference's nn.Embedding raises IndexError on every call, xattn_gpt.py:
    103-114; ids of -1 arise when an episode's first history slot is masked).  The kernels set a persistent device flag; the
    first call of a module checks it synchronously, later calls copy it to pinned host memory asynchronously and the NEXT call
    (or `check()`) raises -- no host synchronisation on the steady-state path."""

    def __init__(self):
        self.flag = None      # int32[1] on the device
        self.host = None      # pinned int32[1]
        self.event = None
        self.checked_sync = False

    def device_flag(self, dev) -> torch.Tensor:
        if self.flag is None or self.flag.device != dev:
            self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
            self.host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self.event, self.checked_sync = None, False
        return self.flag

    def poll(self):
        """Raise if a PREVIOUS call saw a bad id (its flag copy has landed)."""
        if self.event is not None and not torch.cuda.is_current_stream_capturing() and self.event.query():
            self.event = None
            if int(self.host[0]) != 0:
                self.reset()
                raise IndexError("index out of range in self (a previous call passed a position id outside the embedding table)")

    def after_launch(self):
        if not self.checked_sync:
            self.checked_sync = True
            if int(self.flag.item()) != 0:
                self.reset()
                raise IndexError("index out of range in self (position id outside the embedding table)")
            return
        if torch.cuda.is_current_stream_capturing():
            return
        self.host.copy_(self.flag, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def check(self):
        """Synchronous check (host sync)."""
        if self.flag is not None and int(self.flag.item()) != 0:
            self.reset()
            raise IndexError("index out of range in self (position id outside the embedding table)")

    def reset(self):
        if self.flag is not None:
            self.flag.zero_()
        self.event = None


class XAttention(nn.Module):
    def __init__(self, dim: int, *, num_heads: int, ff_expanding: int, kv_n_positions: int, use_geglu: bool):
        super().__init__()
        if dim % num_heads != 0:
            raise ValueError(f"dim ({dim}) must be divisible by num_heads ({num_heads}).")
        self.num_heads = num_heads
        self.dim = dim
        inner = int(dim * ff_expanding)
        self.layernorm = nn.LayerNorm(dim)
        self.query = nn.Linear(dim, dim, bias=False)
        self.key_value = nn.Linear(dim, 2 * dim, bias=False)
        self.attention_out = nn.Linear(dim, dim, bias=False)
        self.ln = nn.LayerNorm(dim)
        self.linear1 = nn.Linear(dim, inner, bias=False)
        self.linear2 = nn.Linear(inner, dim, bias=False)
        self.gated_layer = nn.Linear(dim, inner, bias=False) if use_geglu else None
        self.register_buffer("kv_position_ids", torch.arange(kv_n_positions))


class XAttnGPT(nn.Module):
    def __init__(
        self,
        embd_dim: int = 768,
        *,
        n_positions: int = 512,
        n_layer: int = 12,
        n_head: int = 12,
        dropout: float = 0.1,
        xattn_n_head: int = 8,
        xattn_ff_expanding: int = 4,
        xattn_detach_qk: bool = False,
        xattn_n_positions: int,
        use_geglu: bool = False,
    ):
        super().__init__()
        self.embd_dim, self.n_layer, self.n_head, self.xattn_n_head = embd_dim, n_layer, n_head, xattn_n_head
        self.n_positions, self.xattn_n_positions = n_positions, xattn_n_positions
        self.positions_embed = nn.Embedding(n_positions, embd_dim)
        self.xattn_positions_embed = nn.Embedding(xattn_n_positions, embd_dim)
        self.h = nn.ModuleList([Block(embd_dim, n_positions, n_head, use_geglu) for _ in range(n_layer)])
        self.xattns = nn.ModuleList(
            [XAttention(embd_dim, num_heads=xattn_n_head, ff_expanding=xattn_ff_expanding, kv_n_positions=xattn_n_positions, use_geglu=use_geglu)
             for _ in range(n_layer)]
        )
        self.register_buffer("position_ids", torch.arange(n_positions))
        self.register_buffer("xattn_position_ids", torch.arange(xattn_n_positions))
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                nn.init.normal_(m.weight, std=0.02)
        self._input_checked = False
        self._wc = eng.WeightCache()
        self._pos_guard = PosIdGuard()

    def check_errors(self):
        """Host-synchronising check of the deferred position-id error flag (see PosIdGuard)."""
        self._pos_guard.check()

    # ---------------------------------------------------------------------------------------------
    def _packed(self, ctx, p):
        def build():
            L = []
            for blk, xa in zip(self.h, self.xattns):
                d = {}
                d["wq"] = eng.pack_linear(ctx, xa.query.weight, None, transposed=False, p=p, f8=True)
                d["wkv"] = eng.pack_linear(ctx, xa.key_value.weight, None, transposed=False, p=p, f8=True)
                d["wo"] = eng.pack_linear(ctx, xa.attention_out.weight, None, transposed=False, p=p, f8=True)
                # linear1 reads ln(a), the gate reads a itself (components.py:218-221): one GEGLU GEMM over the un-normalised a with
                # `ln` folded into the value half only
                xln = (xa.ln.weight.detach(), xa.ln.bias.detach())
                if xa.gated_layer is not None:
                    d["w1g"] = eng.pack_glu(ctx, xa.linear1.weight, None, xa.gated_layer.weight, val_transposed=False, gate_transposed=False,
                                            p=p, f8=True, ln=xln, ln_gate=False)
                else:  # use_geglu=False (components.py:139-142,218-223): gelu(linear1(ln(a))), no gate
                    d["w1"] = eng.pack_linear(ctx, xa.linear1.weight, None, transposed=False, p=p, f8=True, ln=xln)
                d["w2"] = eng.pack_linear(ctx, xa.linear2.weight, None, transposed=False, p=p, f8=True)
                d.update(pack_block(ctx, blk, p))
                L.append(d)
            return L

        params = tuple(t for t in self.parameters())
        return self._wc.get("layers", params, build)

    def _check_input(self, obs_action_tokens, prompt_tokens, prompt_mask, batch_first, obs_action_masks):
        """xattn_gpt.py:141-177 (first call only; host syncs)."""
        assert obs_action_tokens.dim() == 3 and obs_action_tokens.dtype == torch.float32
        assert prompt_tokens.dim() == 3 and prompt_tokens.dtype == torch.float32
        if batch_first:
            B_oa, L_oa, E_oa = obs_action_tokens.shape
            B_p, L_p, E_p = prompt_tokens.shape
        else:
            L_oa, B_oa, E_oa = obs_action_tokens.shape
            L_p, B_p, E_p = prompt_tokens.shape
        assert B_oa == B_p and E_oa == E_p
        if prompt_mask is not None:
            assert prompt_mask.shape == (B_oa, L_p) or prompt_mask.shape == (B_oa, 1, L_p), \
                f"Expect `prompt_mask` to have shape of either ({B_oa, 1, L_p}) or ({B_oa, L_p}), but got {prompt_mask.shape}"
            assert torch.all(prompt_mask.sum(dim=-1) > 0), "each source token should attend to at least one target token"
            assert prompt_mask.dtype == torch.bool
        if obs_action_masks is not None:
            assert obs_action_masks.shape == (B_oa, L_oa)
            assert torch.all(obs_action_masks.sum(dim=-1) > 0)
            assert obs_action_masks.dtype == torch.bool

    # ---------------------------------------------------------------------------------------------
    def forward(
        self,
        *,
        obs_action_tokens: torch.Tensor,
        obs_action_position_ids: Optional[torch.Tensor] = None,
        prompt_tokens: torch.Tensor,
        prompt_mask: Optional[torch.Tensor] = None,
        prompt_position_ids: Optional[torch.Tensor] = None,
        batch_first: bool = False,
        obs_action_masks: Optional[torch.Tensor] = None,
        cache: Optional[DecodeCache] = None,
    ):
        """Reference signature (xattn_gpt.py:89-99) plus `cache`: with a DecodeCache the obs/action arguments describe only the
        tokens appended this step (their position ids are absolute) and the return value holds only their rows."""
        ctx = eng.ctx_for(obs_action_tokens)
        p = eng.prec()
        if not self._input_checked and cache is None:
            self._check_input(obs_action_tokens, prompt_tokens, prompt_mask, batch_first, obs_action_masks)
        dev = obs_action_tokens.device
        if batch_first:
            B, L, E = obs_action_tokens.shape
            Lp = prompt_tokens.shape[1]
        else:
            L, B, E = obs_action_tokens.shape
            Lp = prompt_tokens.shape[0]
        assert E == self.embd_dim
        assert Lp <= self.xattn_n_positions and L <= self.n_positions
        if cache is not None:
            if obs_action_position_ids is None or obs_action_masks is None:
                raise ValueError("cached decode needs absolute position ids and masks for the appended tokens")
            check_cache_append(cache, B, L, E, p)
        if obs_action_tokens.dtype != torch.float32 or prompt_tokens.dtype != torch.float32:
            raise TypeError("XAttnGPT expects float32 tokens (xattn_gpt.py:150,152)")
        tok = obs_action_tokens if obs_action_tokens.stride(-1) == 1 else obs_action_tokens.contiguous()
        ptk = prompt_tokens if prompt_tokens.stride(-1) == 1 else prompt_tokens.contiguous()
        sb, sl = (tok.stride(0), tok.stride(1)) if batch_first else (tok.stride(1), tok.stride(0))
        psb, psl = (ptk.stride(0), ptk.stride(1)) if batch_first else (ptk.stride(1), ptk.stride(0))
        if obs_action_position_ids is None:
            obs_action_position_ids = self.position_ids[None, :L].expand(B, L)
        if prompt_position_ids is None:
            prompt_position_ids = self.xattn_position_ids[None, :Lp].expand(B, Lp)
        oa_ids = obs_action_position_ids.to(torch.int64).contiguous()
        pr_ids = prompt_position_ids.to(torch.int64).contiguous()
        if prompt_mask is not None and prompt_mask.dim() == 3:
            prompt_mask = prompt_mask.squeeze(1)
        pmask = None if prompt_mask is None else eng.as_u8(prompt_mask)
        omask = None if obs_action_masks is None else eng.as_u8(obs_action_masks)
        if cache is not None:
            cache.mask[:, cache.L:cache.L + L].copy_(omask)

        M, Mp, H, Hx = B * L, B * Lp, self.n_head, self.xattn_n_head
        d_s, d_x = E // H, E // Hx
        self._pos_guard.poll()
        err = self._pos_guard.device_flag(dev)
        # x = tokens + positions_embed[ids] (fp32 residual stream); kv = prompt + xattn_positions_embed[ids] (operands only)
        x32 = torch.empty((M, E), dtype=torch.float32, device=dev)
        ctx.add_pos_embed(tok, sb, sl, oa_ids, self.positions_embed.weight.detach(), B, L, E, out_f32=x32, err_flag=err)
        need_prompt = cache is None or cache.prompt_kv is None
        kv16 = eng.Opnd(Mp, E, dev, p.split, f8=p.f8) if need_prompt else None
        if not need_prompt:
            pass
        elif p.f8:  # prompt + position embedding feeds only the key_value GEMM: fp32 once, then hi16 + e4m3 views
            kv32 = torch.empty((Mp, E), dtype=torch.float32, device=dev)
            ctx.add_pos_embed(ptk, psb, psl, pr_ids, self.xattn_positions_embed.weight.detach(), B, Lp, E, out_f32=kv32, hi=kv16.hi, lo=None,
                              dtype=p.dtype, err_flag=err)
            ctx.split_f8(kv32, kv16.lo8, kv16.hi8)
            del kv32
        else:
            ctx.add_pos_embed(ptk, psb, psl, pr_ids, self.xattn_positions_embed.weight.detach(), B, Lp, E, hi=kv16.hi, lo=kv16.lo,
                              dtype=p.dtype, err_flag=err)
        self._pos_guard.after_launch()  # first call: synchronous check; later: asynchronous copy, reported by the next call
        if cache is None:
            self._input_checked = True

        layers = self._packed(ctx, p)
        lnw = lambda ln: (ln.weight.detach(), ln.bias.detach())
        # first layer's query LayerNorm; later ones are chained onto the previous block's LN2
        w, b = lnw(self.xattns[0].layernorm)
        _, _, qin16 = eng.norm(ctx, x32, p, rows=M, cols=E, w=w, b=b, eps=self.xattns[0].layernorm.eps, want16=True, out_f8=True)
        if cache is not None and need_prompt:
            cache.prompt_kv = [eng.gemm(ctx, kv16, W["wkv"], p, want16=True)[1] for W in layers]
        for i, (blk, xa, W) in enumerate(zip(self.h, self.xattns, layers)):
            # ---------------- XAttention ----------------
            _, q16 = eng.gemm(ctx, qin16, W["wq"], p, want16=True)
            kvp16 = cache.prompt_kv[i] if cache is not None else eng.gemm(ctx, kv16, W["wkv"], p, want16=True)[1]
            c16 = eng.Opnd(M, E, dev, p.split, f8=p.f8)
            ctx.attention(q=(q16.hi, q16.lo, q16.ld, 0), k=(kvp16.hi, kvp16.lo, kvp16.ld, 0), v=(kvp16.hi, kvp16.lo, kvp16.ld, E),
                          o=(c16.hi, c16.lo, c16.ld, 0), B=B, H=Hx, Lq=L, Lk=Lp, D=d_x, scale=1.0 / math.sqrt(d_x), causal=False,
                          key_mask=pmask, dtype=p.dtype, o8=None if c16.lo8 is None else (c16.lo8, c16.hi8))
            part = eng.stats_buffer(ctx, M, W["wo"], dev)
            a32, a16 = eng.gemm(ctx, c16, W["wo"], p, residual=x32, want_f32=True, want16=True, out_f8=True, stats_out=part)
            st = eng.row_stats_of(ctx, part, M, E, xa.ln.eps)
            _, h16 = eng.gemm(ctx, a16, W["w1g"] if "w1g" in W else W["w1"], p, act=_C.ACT_GELU, want16=True, out_f8=True, row_stats=st)
            xb32, xb16 = eng.gemm(ctx, h16, W["w2"], p, residual=a32, want_f32=True, want16=True, out_f8=True)
            del h16, a32, a16
            # ---------------- causal Block ----------------
            nxt = self.xattns[i + 1].layernorm if i + 1 < self.n_layer else None
            x32, qin16 = run_block(ctx, p, W, blk, xb32, xb16, c16, B=B, L=L, E=E, H=H, omask=omask, chain_ln=nxt, out_f32=x32, cache=cache,
                                   layer=i)
        if cache is not None:
            cache.L += L
        out = x32.view(B, L, E)
        return out if batch_first else out.transpose(0, 1)


class _OpenAIGPTModel(nn.Module):
    """Parameter holder with the key layout of the reference's OpenAIGPTModel (gpt.py:83-101)."""

    def __init__(self, vocab_size, n_positions, n_embd, n_layer, n_head, geglu):
        super().__init__()
        self.tokens_embed = nn.Embedding(vocab_size, n_embd)
        self.positions_embed = nn.Embedding(n_positions, n_embd)
        self.h = nn.ModuleList([Block(n_embd, n_positions, n_head, geglu) for _ in range(n_layer)])
        for blk in self.h:  # HF >= 4.3x keeps the causal buffer out of the state dict (checked against the reference here)
            buf = blk.attn._buffers.pop("bias")
            blk.attn.register_buffer("bias", buf, persistent=False)
        self.register_buffer("position_ids", torch.arange(n_positions))


class HFGPT(nn.Module):
    """Decoder-only GPT-1 stack, GEGLU or plain gelu_new MLPs (reference: vima/nn/seq_modeling/gpt/gpt.py:15-301) -- the VIMA-Gato baseline's
    sequence model.  Same Block kernels as XAttnGPT (causal-only path, BASELINE.json configs[4])."""

    def __init__(self, *, vocab_size=40478, n_positions=512, n_embd=768, n_layer=12, n_head=12, dropout: float = 0.1, use_geglu: bool = False):
        super().__init__()
        self.n_embd, self.n_layer, self.n_head, self.n_positions = n_embd, n_layer, n_head, n_positions
        self.lm = _OpenAIGPTModel(vocab_size, n_positions, n_embd, n_layer, n_head, use_geglu)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                nn.init.normal_(m.weight, std=0.02)
        self._wc = eng.WeightCache()
        # checkpoints written with transformers 4.x carry `lm.h.N.attn.bias`; accept and ignore it
        self._register_load_state_dict_pre_hook(self._drop_causal_buffers)

    @staticmethod
    def _drop_causal_buffers(state_dict, prefix, *args):
        for k in [k for k in state_dict if k.startswith(prefix + "lm.h.") and k.endswith(".attn.bias")]:
            del state_dict[k]

    def forward(self, x: torch.Tensor, *, custom_mask: Optional[torch.Tensor] = None, position_ids: Optional[torch.Tensor] = None,
                batch_first: bool = False):
        """x: (L,B,E) if not batch_first else (B,L,E); custom_mask (B,L) or (B,1,L) combined with the causal mask (gpt.py:46-79)."""
        ctx = eng.ctx_for(x)
        p = eng.prec()
        if batch_first:
            B, L, E = x.shape
        else:
            L, B, E = x.shape
        assert E == self.n_embd and L <= self.n_positions
        dev = x.device
        xf = x.float()
        if xf.stride(-1) != 1:
            xf = xf.contiguous()
        sb, sl = (xf.stride(0), xf.stride(1)) if batch_first else (xf.stride(1), xf.stride(0))
        if position_ids is None:
            position_ids = self.lm.position_ids[None, :L].expand(B, L)
        ids = position_ids.to(torch.int64).contiguous()
        omask = None
        if custom_mask is not None:
            if custom_mask.dim() == 3:
                custom_mask = custom_mask.squeeze(dim=1)
            omask = eng.as_u8(custom_mask != 0)
        M, H = B * L, self.n_head
        x32 = torch.empty((M, E), dtype=torch.float32, device=dev)
        x16 = eng.Opnd(M, E, dev, p.split, f8=p.f8)
        ctx.add_pos_embed(xf, sb, sl, ids, self.lm.positions_embed.weight.detach(), B, L, E, out_f32=x32, hi=x16.hi, lo=x16.lo, dtype=p.dtype)
        if p.f8:
            ctx.split_f8(x32, x16.lo8, x16.hi8)
        layers = self._wc.get("blocks", tuple(self.lm.h.parameters()), lambda: [pack_block(ctx, blk, p) for blk in self.lm.h])
        c16 = eng.Opnd(M, E, dev, p.split, f8=p.f8)
        for i, (blk, W) in enumerate(zip(self.lm.h, layers)):
            x32, x16 = run_block(ctx, p, W, blk, x32, x16, c16, B=B, L=L, E=E, H=H, omask=omask, want16=i + 1 < self.n_layer)
        out = x32.view(B, L, E)
        return out if batch_first else out.transpose(0, 1)
