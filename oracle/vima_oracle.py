"""CPU ORACLE for the VIMA policy forward pass -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain torch-CPU fp32 restatement of the reference algorithm (no HF / reference imports), written as
functions over a state dict with the reference's key names.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this file; the product path
(`vima_b200`) never does and fails loudly if its CUDA library is missing.

Parity pin: `tests/test_oracle_golden.py` checks every function here against golden vectors minted from
the UNMODIFIED reference source run in the build container (`tests/golden/make_golden.py`, via
`oracle/ref_shim.py`).  The reference itself ships no tests or golden vectors.

Each function cites the reference lines it restates (paths relative to /root/reference; `HF:` = the
`transformers` package the reference subclasses).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

FP32_MIN = torch.finfo(torch.float32).min
VIMA_IMG_MEAN = (0.3471, 0.3429, 0.3383)  # vima/nn/obj_encoder/vit/vit.py:9
VIMA_IMG_STD = (0.3011, 0.2961, 0.2956)  # vima/nn/obj_encoder/vit/vit.py:10
VIEWS = ("front", "top")  # sorted(["front","top"]), vima/nn/obj_encoder/obj_encoder.py:30
ACTION_DIMS = {  # vima/policy/vima_policy.py:82-87 (dict order == logits order)
    "pose0_position": [50, 100],
    "pose0_rotation": [50, 50, 50, 50],
    "pose1_position": [50, 100],
    "pose1_rotation": [50, 50, 50, 50],
}

# ------------------------------------------------------------------------------------------------
# matmul hook: identity in the oracle proper; the precision study (tools/precision_study.py) swaps
# it to emulate reduced-precision multiplicands.
# ------------------------------------------------------------------------------------------------
_operand_round: Optional[Callable[[torch.Tensor], torch.Tensor]] = None


def set_operand_rounding(fn: Optional[Callable[[torch.Tensor], torch.Tensor]]) -> None:
    global _operand_round
    _operand_round = fn


def _mm(x: torch.Tensor, w_t: torch.Tensor) -> torch.Tensor:
    """x[..., K] @ w_t[K, N]"""
    if _operand_round is not None:
        x = _operand_round(x)
        w_t = _operand_round(w_t)
    return x @ w_t


def linear(x, w, b=None):
    """nn.Linear: weight is [out, in]."""
    y = _mm(x, w.t())
    return y if b is None else y + b


def conv1d_hf(x, w, b):
    """HF Conv1D (HF:pytorch_utils.py:119-123): weight is [in, out]; y = addmm(b, x, w)."""
    return _mm(x, w) + b


def layer_norm(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_new(x):
    """HF NewGELUActivation (transformers/activations.py; ACT_FNS["gelu"] of HF:modeling_openai.py, the OpenAIGPTConfig.afn default)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def mlp_seq(sd: SD, prefix: str, x: torch.Tensor, idx: Sequence[int]) -> torch.Tensor:
    """build_mlp Sequential (vima/nn/utils.py:84-91): Linear, Identity, ReLU, ..., Linear; Linear sits at idx."""
    for j, i in enumerate(idx):
        x = linear(x, sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"])
        if j + 1 < len(idx):
            x = torch.relu(x)
    return x


# ------------------------------------------------------------------------------------------------
# XAttnGPT decoder
# ------------------------------------------------------------------------------------------------
def xattention(sd: SD, p: str, q: torch.Tensor, kv: torch.Tensor, kv_mask: torch.Tensor, n_head: int):
    """XAttention.forward, vima/nn/seq_modeling/xattn_gpt/components.py:158-228.

    q (B,L,E), kv (B,Lp,E) un-normalised, kv_mask (B,Lp) bool.  All six Linears are bias-free (:130-142).
    """
    B, L, E = q.shape
    Lp = kv.shape[1]
    d = E // n_head
    queries = linear(layer_norm(q, sd[p + "layernorm.weight"], sd[p + "layernorm.bias"]), sd[p + "query.weight"])  # :166-167
    k, v = linear(kv, sd[p + "key_value.weight"]).chunk(2, dim=-1)  # :175
    qh = queries.view(B, L, n_head, d).permute(0, 2, 1, 3)
    kh = k.reshape(B, Lp, n_head, d).permute(0, 2, 1, 3)
    vh = v.reshape(B, Lp, n_head, d).permute(0, 2, 1, 3)
    scores = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(d)  # :187-195
    scores = scores + ((1.0 - kv_mask[:, None, None, :].to(torch.float32)) * FP32_MIN)  # :197-202, :252-254
    probs = torch.softmax(scores, dim=-1)  # :207
    ctx = torch.matmul(probs, vh).permute(0, 2, 1, 3).reshape(B, L, E)  # :210-214
    a = linear(ctx, sd[p + "attention_out.weight"]) + q  # :217-218 (residual with UN-normalised q)
    ff = linear(layer_norm(a, sd[p + "ln.weight"], sd[p + "ln.bias"]), sd[p + "linear1.weight"])  # :220-221
    ff = gelu_erf(ff)  # :222 nn.GELU() exact erf
    if (p + "gated_layer.weight") in sd:  # use_geglu (:139-142)
        ff = ff * linear(a, sd[p + "gated_layer.weight"])  # :223-224 gate reads UN-normalised a
    ff = linear(ff, sd[p + "linear2.weight"])  # :225
    return ff + a  # :227


def causal_self_attention(sd: SD, p: str, x: torch.Tensor, add_mask: Optional[torch.Tensor], n_head: int):
    """HF openai Attention.forward (HF:modeling_openai.py:100-115) with `_attn` overridden at
    vima/nn/seq_modeling/xattn_gpt/components.py:51-80.  add_mask is additive (B,1,1,L) or None."""
    B, L, E = x.shape
    d = E // n_head
    qkv = conv1d_hf(x, sd[p + "c_attn.weight"], sd[p + "c_attn.bias"])
    q, k, v = qkv.split(E, dim=2)
    q = q.reshape(B, L, n_head, d).permute(0, 2, 1, 3)
    k = k.reshape(B, L, n_head, d).permute(0, 2, 3, 1)
    v = v.reshape(B, L, n_head, d).permute(0, 2, 1, 3)
    w = torch.matmul(q, k) / math.sqrt(d)  # :56-58, scale=True (xattn_gpt.py:49)
    tril = torch.tril(torch.ones(L, L, dtype=w.dtype, device=w.device))
    w = w * tril + -1e4 * (1 - tril)  # :61-63
    if add_mask is not None:
        w = w + add_mask  # :65-67
    w = torch.softmax(w, dim=-1)
    a = torch.matmul(w, v).permute(0, 2, 1, 3).reshape(B, L, E)
    return conv1d_hf(a, sd[p + "c_proj.weight"], sd[p + "c_proj.bias"])


def gpt_block(sd: SD, p: str, x: torch.Tensor, add_mask: Optional[torch.Tensor], n_head: int):
    """Block.forward (post-LN), components.py:23-37; MLP.forward GEGLU, components.py:97-102."""
    a = causal_self_attention(sd, p + "attn.", x, add_mask, n_head)
    n = layer_norm(x + a, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
    h = conv1d_hf(n, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
    if (p + "mlp.gated_layer.weight") in sd:  # afn == "geglu": nn.GELU() (:89-91), gate reads the NORMALISED n (:100)
        h = gelu_erf(h) * linear(n, sd[p + "mlp.gated_layer.weight"])
    else:  # afn = "gelu" (the OpenAIGPTConfig default) -> ACT_FNS["gelu"] = gelu_new (:92-94)
        h = gelu_new(h)
    m = conv1d_hf(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    return layer_norm(n + m, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])


def xattn_gpt_forward(
    sd: SD,
    p: str,
    *,
    obs_action_tokens: torch.Tensor,  # (L,B,E) seq-first
    obs_action_position_ids: torch.Tensor,  # (B,L) int64
    prompt_tokens: torch.Tensor,  # (Lp,B,E)
    prompt_mask: torch.Tensor,  # (B,Lp) bool
    prompt_position_ids: torch.Tensor,  # (B,Lp) int64
    obs_action_masks: torch.Tensor,  # (B,L) bool
    n_layer: int,
    n_head: int,
    xattn_n_head: int,
):
    """XAttnGPT.forward, vima/nn/seq_modeling/xattn_gpt/xattn_gpt.py:73-139 (batch_first=False)."""
    x = obs_action_tokens.transpose(0, 1)
    kv = prompt_tokens.transpose(0, 1)
    x = x + sd[p + "positions_embed.weight"][obs_action_position_ids]  # :103-105
    kv = kv + sd[p + "xattn_positions_embed.weight"][prompt_position_ids]  # :110-114, once, outside the layers
    add_mask = (1.0 - obs_action_masks[:, None, None, :].to(torch.float32)) * FP32_MIN  # :116-121
    for i in range(n_layer):  # :123-132 cross-attention FIRST, then the causal block
        x = xattention(sd, f"{p}xattns.{i}.", x, kv, prompt_mask, xattn_n_head)
        x = gpt_block(sd, f"{p}h.{i}.", x, add_mask, n_head)
    return x.transpose(0, 1)


# ------------------------------------------------------------------------------------------------
# Object encoder (ViT + bbox MLP)
# ------------------------------------------------------------------------------------------------
def image_preprocess(img_u8: torch.Tensor) -> torch.Tensor:
    """basic_image_tensor_preprocess, vima/nn/obj_encoder/vit/preprocess.py:9-43 with VIMA mean/std."""
    x = img_u8.float() / 255.0
    mean = torch.tensor(VIMA_IMG_MEAN, dtype=torch.float32).view(3, 1, 1)
    std = torch.tensor(VIMA_IMG_STD, dtype=torch.float32).view(3, 1, 1)
    return (x - mean) / std


def vit_forward(sd: SD, p: str, img: torch.Tensor, heads: int = 24, patch: int = 16) -> torch.Tensor:
    """VisionTransformer.forward vit.py:171-191 + ResidualAttentionBlock vit.py:199-236. img (N,3,H,W) float."""
    N = img.shape[0]
    w = sd[p + "conv1.weight"]
    width = w.shape[0]
    # conv k=s=patch, no bias == per-patch GEMM (vit.py:151-157,172)
    patches = F.unfold(img, kernel_size=patch, stride=patch).transpose(1, 2)  # (N, n_patch, 3*p*p)
    x = _mm(patches, w.reshape(width, -1).t())
    x = torch.cat([sd[p + "cls_token"].expand(N, 1, width), x], dim=1) + sd[p + "pos_embed"]  # :176-179
    x = layer_norm(x, sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"])
    S = x.shape[1]
    d = width // heads
    i = 0
    while f"{p}blocks.{i}.ln_1.weight" in sd:
        b = f"{p}blocks.{i}."
        y = layer_norm(x, sd[b + "ln_1.weight"], sd[b + "ln_1.bias"])
        qkv = linear(y, sd[b + "attn.in_proj_weight"], sd[b + "attn.in_proj_bias"])  # nn.MultiheadAttention
        q, k, v = qkv.split(width, dim=-1)
        q = q.view(N, S, heads, d).transpose(1, 2)
        k = k.view(N, S, heads, d).transpose(1, 2)
        v = v.view(N, S, heads, d).transpose(1, 2)
        att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
        o = torch.matmul(att, v).transpose(1, 2).reshape(N, S, width)
        x = x + linear(o, sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"])
        y = layer_norm(x, sd[b + "ln_2.weight"], sd[b + "ln_2.bias"])
        h = linear(y, sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)  # QuickGELU vit.py:194-196
        x = x + linear(h, sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"])
        i += 1
    x = layer_norm(x[:, 0, :], sd[p + "ln_post.weight"], sd[p + "ln_post.bias"])  # :187
    return _mm(x, sd[p + "projection"])  # :189-190


def obj_encoder_forward(sd: SD, p: str, cropped_img: dict, bbox: dict, mask: dict = None) -> torch.Tensor:
    """ObjEncoder.forward, vima/nn/obj_encoder/obj_encoder.py:66-95.  Leading dims arbitrary."""
    outs = []
    for view in VIEWS:
        img = cropped_img[view]
        lead = img.shape[:-3]
        feats = vit_forward(sd, p + "cropped_img_encoder.vit.", image_preprocess(img).flatten(0, img.dim() - 4))
        feats = feats.view(*lead, -1)
        bb = bbox[view].float() / torch.tensor([256.0, 128.0, 128.0, 256.0])  # :79-85  [w,h,h,w] maxima
        bb = mlp_seq(sd, f"{p}bbox_mlp.{view}.", bb, (0, 3, 6))  # :86
        outs.append(
            linear(
                torch.cat([feats, bb], dim=-1),
                sd[f"{p}pre_transformer_layer.{view}.weight"],
                sd[f"{p}pre_transformer_layer.{view}.bias"],
            )
        )  # :88-93
    return torch.cat(outs, dim=-2)  # :94


# ------------------------------------------------------------------------------------------------
# T5 prompt encoder
# ------------------------------------------------------------------------------------------------
def t5_relative_position_bucket(relative_position: torch.Tensor, num_buckets: int = 32, max_distance: int = 128):
    """HF:modeling_t5.py `_relative_position_bucket`, bidirectional=True (encoder)."""
    num_buckets //= 2
    rb = (relative_position > 0).to(torch.long) * num_buckets
    n = torch.abs(relative_position)
    max_exact = num_buckets // 2
    is_small = n < max_exact
    large = max_exact + (
        torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return rb + torch.where(is_small, n, large)


def t5_rms_norm(x, w, eps=1e-6):
    """HF T5LayerNorm (HF:modeling_t5.py:46-68): no mean subtraction, no bias, fp32 variance."""
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def t5_encoder_forward(sd: SD, p: str, x: torch.Tensor, mask: torch.Tensor, n_heads: int = 12) -> torch.Tensor:
    """T5PromptEncoder.forward -> T5Stack (encoder) -> T5Block, vima/nn/prompt_encoder/prompt_encoder.py:30-58,
    212-473, 491-604, 654-678, 682-825.  x (B,Lp,768) batch-first, mask (B,Lp) bool.  `p` ends with 't5.encoder.'"""
    B, Lp, D = x.shape
    d_kv = sd[p + "block.0.layer.0.SelfAttention.q.weight"].shape[0] // n_heads
    ctx_pos = torch.arange(Lp)[:, None]
    mem_pos = torch.arange(Lp)[None, :]
    buckets = t5_relative_position_bucket(mem_pos - ctx_pos)  # HF compute_bias: memory - context
    bias = sd[p + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"][buckets]  # (Lp,Lp,H)
    bias = bias.permute(2, 0, 1).unsqueeze(0)
    ext = (1.0 - mask[:, None, None, :].to(torch.float32)) * FP32_MIN  # get_extended_attention_mask
    position_bias = bias + ext  # prompt_encoder.py:794-797, shared by all layers (:426)
    h = x
    i = 0
    while f"{p}block.{i}.layer.0.layer_norm.weight" in sd:
        b = f"{p}block.{i}."
        n = t5_rms_norm(h, sd[b + "layer.0.layer_norm.weight"])
        q = linear(n, sd[b + "layer.0.SelfAttention.q.weight"]).view(B, Lp, n_heads, d_kv).transpose(1, 2)
        k = linear(n, sd[b + "layer.0.SelfAttention.k.weight"]).view(B, Lp, n_heads, d_kv).transpose(1, 2)
        v = linear(n, sd[b + "layer.0.SelfAttention.v.weight"]).view(B, Lp, n_heads, d_kv).transpose(1, 2)
        scores = torch.matmul(q, k.transpose(3, 2)) + position_bias  # no 1/sqrt(d) (:771-773,799-800)
        w = torch.softmax(scores, dim=-1)
        o = torch.matmul(w, v).transpose(1, 2).reshape(B, Lp, n_heads * d_kv)
        h = h + linear(o, sd[b + "layer.0.SelfAttention.o.weight"])
        n = t5_rms_norm(h, sd[b + "layer.1.layer_norm.weight"])
        ff = torch.relu(linear(n, sd[b + "layer.1.DenseReluDense.wi.weight"]))
        h = h + linear(ff, sd[b + "layer.1.DenseReluDense.wo.weight"])
        i += 1
    return t5_rms_norm(h, sd[p + "final_layer_norm.weight"])  # :448


# ------------------------------------------------------------------------------------------------
# Policy-level entry points (vima/policy/vima_policy.py)
# ------------------------------------------------------------------------------------------------
def policy_dims(sd: SD):
    E = sd["xattn_gpt.positions_embed.weight"].shape[1]
    n_layer = 0
    while f"xattn_gpt.h.{n_layer}.ln_1.weight" in sd:
        n_layer += 1
    return E, n_layer


def forward_prompt_assembly(sd: SD, prompts):
    """VIMAPolicy.forward_prompt_assembly, vima_policy.py:161-240.  Returns (Lp,B,E), (B,Lp) bool and the
    pre-T5 assembled tokens (B,Lp,768) for stage-level checks."""
    token_types, word_batch, image_batch = prompts
    word_emb = sd["prompt_embedding._embed_layer.weight"][word_batch]  # :163
    img_emb = obj_encoder_forward(sd, "obj_encoder.", image_batch["cropped_img"], image_batch["bbox"])  # :164
    img_emb = mlp_seq(sd, "prompt_obj_post_layer.", img_emb, (0, 3, 6))  # :165
    n_max_objs = img_emb.shape[-2]
    lens = [sum(1 if t == 0 else n_max_objs for t in tt) for tt in token_types]
    L_max = max(lens)
    B = len(token_types)
    toks = torch.zeros(B, L_max, img_emb.shape[-1])
    masks = torch.zeros(B, L_max, dtype=torch.bool)
    wp = ip = 0
    for b, tt in enumerate(token_types):
        pos = 0
        for t in tt:
            if t == 0:
                toks[b, pos] = word_emb[wp]
                masks[b, pos] = True
                wp += 1
                pos += 1
            elif t == 1:
                om = torch.cat([image_batch["mask"][v][ip] for v in VIEWS], dim=-1)  # :191-197
                toks[b, pos : pos + n_max_objs] = img_emb[ip]
                masks[b, pos : pos + n_max_objs] = om
                ip += 1
                pos += n_max_objs
            else:
                raise ValueError(f"Invalid prompt token type {t}")
    enc = t5_encoder_forward(sd, "t5_prompt_encoder.t5.encoder.", toks, masks)  # :236-238
    if "t5_prompt_encoder_post_layer.weight" in sd:
        enc = linear(enc, sd["t5_prompt_encoder_post_layer.weight"])  # :239
    return enc.transpose(0, 1), masks, toks


def forward_obs_token(sd: SD, obs):
    """VIMAPolicy.forward_obs_token, vima_policy.py:242-259. obs = {"ee": (T,B) i64, "objects": {...}}."""
    objects, ee = obs["objects"], obs["ee"]
    img_feats = obj_encoder_forward(sd, "obj_encoder.", objects["cropped_img"], objects["bbox"])  # (T,B,Q,E)
    ee_feats = sd["end_effector_encoder.weight"][ee]  # (T,B,2)
    ee_feats = ee_feats.unsqueeze(2).repeat(1, 1, img_feats.shape[-2], 1)
    feats = linear(torch.cat([img_feats, ee_feats], dim=-1), sd["obs_fusion_layer.weight"], sd["obs_fusion_layer.bias"])
    mask = torch.cat([objects["mask"][v] for v in VIEWS], dim=-1)
    return feats, mask


def assemble_history(obs_token, obs_mask, action_token):
    """Token interleave + masks + cumsum position ids, vima_policy.py:124-147."""
    T, B, Q, E = obs_token.shape
    La = 0 if action_token is None else action_token.shape[0]
    L = T * Q + La
    tokens = torch.zeros(L, B, E)
    masks = torch.ones(L, B, dtype=torch.bool)
    for t in range(T):
        tokens[t * (Q + 1) : t * (Q + 1) + Q] = obs_token[t].transpose(0, 1)
        masks[t * (Q + 1) : t * (Q + 1) + Q] = obs_mask[t].transpose(0, 1)
    for t in range(La):
        tokens[t * (Q + 1) + Q] = action_token[t]
    position_ids = (torch.cumsum(masks, dim=0) - 1).long()
    return tokens, masks, position_ids


def policy_forward(sd: SD, obs_token, obs_mask, action_token, prompt_token, prompt_token_mask, *, n_head: int, xattn_n_head: int):
    """VIMAPolicy.forward, vima_policy.py:116-159."""
    E, n_layer = policy_dims(sd)
    Q = obs_token.shape[-2]
    tokens, masks, position_ids = assemble_history(obs_token, obs_mask, action_token)
    prompt_position_ids = torch.cumsum(prompt_token_mask, dim=1) - 1
    out = xattn_gpt_forward(
        sd,
        "xattn_gpt.",
        obs_action_tokens=tokens,
        obs_action_position_ids=position_ids.transpose(0, 1),
        prompt_tokens=prompt_token,
        prompt_mask=prompt_token_mask,
        prompt_position_ids=prompt_position_ids,
        obs_action_masks=masks.transpose(0, 1),
        n_layer=n_layer,
        n_head=n_head,
        xattn_n_head=xattn_n_head,
    )
    return out[Q - 1 :: Q + 1]  # :158


def action_decoder_logits(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """ActionDecoder.forward (vima/nn/action_decoder/action_decoder.py:51-52,165-166): raw logits (...,700)
    in dict order, i.e. before Categorical's logsumexp normalisation (dists.py:20-23)."""
    outs = []
    for key, dims in ACTION_DIMS.items():
        for j in range(len(dims)):
            outs.append(mlp_seq(sd, f"action_decoder._decoders.{key}.mlps.{j}.", x, (0, 3, 6)))
    return torch.cat(outs, dim=-1)


def action_modes(logits: torch.Tensor) -> Dict[str, torch.Tensor]:
    """MultiCategorical.mode(), vima/nn/action_decoder/dists.py:25-28: argmax of softmax probs per head."""
    out, off = {}, 0
    for key, dims in ACTION_DIMS.items():
        idx = []
        for n in dims:
            probs = torch.softmax(logits[..., off : off + n], dim=-1)
            idx.append(torch.argmax(probs, dim=-1))
            off += n
        out[key] = torch.stack(idx, dim=-1)
    return out


def de_discretize_actions(actions: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """VIMAPolicy._de_discretize_actions, vima_policy.py:301-322: bin index / bin count as float32 (50 x-bins, 100 y-bins,
    50 rotation bins: vima_policy.py:77-79).  Pinned bit-exact by tests/golden/dediscretize.npz (minted from the reference method)."""
    a = {k: v.float().clone() for k, v in actions.items()}
    for k in ("pose0_position", "pose1_position"):
        a[k][..., 0] = a[k][..., 0] / 50
        a[k][..., 1] = a[k][..., 1] / 100
    for k in ("pose0_rotation", "pose1_rotation"):
        a[k] = a[k] / 50
    return a


def forward_action_token(sd: SD, actions: Dict[str, torch.Tensor]) -> torch.Tensor:
    """VIMAPolicy.forward_action_token = ActionEmbedding(_de_discretize_actions(a)), vima_policy.py:261-262,
    301-322; vima/nn/action_embd/action_embd.py:29-37,55-56."""
    a = de_discretize_actions(actions)
    feats = [mlp_seq(sd, f"action_encoder._embed_dict.{k}._layer.", a[k], (0, 3)) for k in sorted(a.keys())]
    return linear(torch.cat(feats, dim=-1), sd["action_encoder._post_layer.weight"], sd["action_encoder._post_layer.bias"])


def postprocess_actions(actions: Dict[str, torch.Tensor], bounds_low: torch.Tensor, bounds_high: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The environment-facing arithmetic after the heads, scripts/example.py:199-232: `_de_discretize_actions`
    (vima_policy.py:301-322), positions scaled into the action bounds and clamped, rotations mapped to [-1, 1]."""
    a = {k: v.float().clone() for k, v in actions.items()}
    for k in ("pose0_position", "pose1_position"):
        a[k][..., 0] = a[k][..., 0] / 50
        a[k][..., 1] = a[k][..., 1] / 100
        a[k] = a[k] * (bounds_high - bounds_low) + bounds_low
        a[k] = torch.clamp(a[k], min=bounds_low, max=bounds_high)
    for k in ("pose0_rotation", "pose1_rotation"):
        a[k] = a[k] / 50
        a[k] = a[k] * 2 - 1
        a[k] = torch.clamp(a[k], min=-1, max=1)
    return a


def policy_step(sd: SD, *, obs, history_obs_tokens, history_obs_masks, history_action_tokens, prompt_tokens,
                prompt_masks, n_head: int, xattn_n_head: int):
    """One policy step as scripts/example.py:125-198 runs it (full-history re-forward)."""
    new_tok, new_mask = forward_obs_token(sd, obs)
    obs_tok = new_tok if history_obs_tokens is None else torch.cat([history_obs_tokens, new_tok], dim=0)
    obs_msk = new_mask if history_obs_masks is None else torch.cat([history_obs_masks, new_mask], dim=0)
    pred = policy_forward(sd, obs_tok, obs_msk, history_action_tokens, prompt_tokens, prompt_masks,
                          n_head=n_head, xattn_n_head=xattn_n_head)
    logits = action_decoder_logits(sd, pred[-1:])
    modes = action_modes(logits)
    act_tok = forward_action_token(sd, modes)
    return dict(obs_tokens=obs_tok, obs_masks=obs_msk, predicted=pred, logits=logits, actions=modes, action_token=act_tok)


# ------------------------------------------------------------------------------------------------
# VIMA-Gato baseline (decoder-only; BASELINE.json configs[4])
# ------------------------------------------------------------------------------------------------
def gato_vit_forward(sd: SD, p: str, img: torch.Tensor, heads: int = 24, patch: int = 32) -> torch.Tensor:
    """GatoVisionTransformerRectangular.forward, vima/nn/obj_encoder/vit/vit.py:120-134 (no CLS; every patch token kept)."""
    N = img.shape[0]
    w = sd[p + "conv1.weight"]
    width = w.shape[0]
    patches = F.unfold(img, kernel_size=patch, stride=patch).transpose(1, 2)
    x = _mm(patches, w.reshape(width, -1).t()) + sd[p + "pos_embed"]
    x = layer_norm(x, sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"])
    S = x.shape[1]
    d = width // heads
    i = 0
    while f"{p}blocks.{i}.ln_1.weight" in sd:
        b = f"{p}blocks.{i}."
        y = layer_norm(x, sd[b + "ln_1.weight"], sd[b + "ln_1.bias"])
        q, k, v = linear(y, sd[b + "attn.in_proj_weight"], sd[b + "attn.in_proj_bias"]).split(width, dim=-1)
        q = q.view(N, S, heads, d).transpose(1, 2)
        k = k.view(N, S, heads, d).transpose(1, 2)
        v = v.view(N, S, heads, d).transpose(1, 2)
        att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
        o = torch.matmul(att, v).transpose(1, 2).reshape(N, S, width)
        x = x + linear(o, sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"])
        y = layer_norm(x, sd[b + "ln_2.weight"], sd[b + "ln_2.bias"])
        h = linear(y, sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)
        x = x + linear(h, sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"])
        i += 1
    x = layer_norm(x, sd[p + "ln_post.weight"], sd[p + "ln_post.bias"])
    return _mm(x, sd[p + "projection"])


def gato_obj_encoder(sd: SD, p: str, rgb: dict) -> torch.Tensor:
    """GatoMultiViewRGBEncoder.forward, vima/nn/obj_encoder/obj_encoder.py:127-142."""
    outs = []
    for view in VIEWS:
        img = rgb[view]
        lead = img.shape[:-3]
        f = gato_vit_forward(sd, p + "cropped_img_encoder.vit.", image_preprocess(img).flatten(0, img.dim() - 4))
        outs.append(f.view(*lead, *f.shape[-2:]))
    return torch.cat(outs, dim=-2)


def hfgpt_forward(sd: SD, p: str, x: torch.Tensor, mask: torch.Tensor, position_ids: torch.Tensor, n_head: int) -> torch.Tensor:
    """HFGPT.forward -> OpenAIGPTModel.forward, vima/nn/seq_modeling/gpt/gpt.py:46-221. x (L,B,E), mask (B,L), ids (B,L)."""
    h = x.transpose(0, 1) + sd[p + "lm.positions_embed.weight"][position_ids]
    add_mask = (1.0 - mask[:, None, None, :].to(torch.float32)) * FP32_MIN
    i = 0
    while f"{p}lm.h.{i}.ln_1.weight" in sd:
        h = gpt_block(sd, f"{p}lm.h.{i}.", h, add_mask, n_head)
        i += 1
    return h.transpose(0, 1)


def gato_forward_prompt_assembly(sd: SD, prompts):
    """VIMAGatoPolicy.forward_prompt_assembly, vima/policy/vima_gato_policy.py:193-251."""
    token_types, word_batch, image_batch = prompts
    word_emb = sd["prompt_embedding._embed_layer.weight"][word_batch]
    img_emb = mlp_seq(sd, "prompt_obj_post_layer.", gato_obj_encoder(sd, "obj_encoder.", image_batch["rgb"]), (0, 3, 6))
    nq = img_emb.shape[-2]
    lens = [sum(1 if t == 0 else nq for t in tt) for tt in token_types]
    B, L_max = len(token_types), max(lens)
    toks = torch.zeros(B, L_max, img_emb.shape[-1])
    masks = torch.zeros(B, L_max, dtype=torch.bool)
    wp = ip = 0
    for b, tt in enumerate(token_types):
        pos = 0
        for t in tt:
            if t == 0:
                toks[b, pos] = word_emb[wp]; wp += 1; pos += 1
            else:
                toks[b, pos:pos + nq] = img_emb[ip]; ip += 1; pos += nq
        masks[b, :pos] = True
    enc = t5_encoder_forward(sd, "t5_prompt_encoder.t5.encoder.", toks, masks)
    if "t5_prompt_encoder_post_layer.weight" in sd:
        enc = linear(enc, sd["t5_prompt_encoder_post_layer.weight"])
    return enc.transpose(0, 1), masks


def gato_forward_obs_token(sd: SD, obs):
    """VIMAGatoPolicy.forward_obs_token, vima_gato_policy.py:253-262."""
    img_feats = gato_obj_encoder(sd, "obj_encoder.", obs["rgb"])
    ee_feats = sd["end_effector_encoder.weight"][obs["ee"]].unsqueeze(2).repeat(1, 1, img_feats.shape[-2], 1)
    return linear(torch.cat([img_feats, ee_feats], dim=-1), sd["obs_fusion_layer.weight"], sd["obs_fusion_layer.bias"])


def gato_policy_forward(sd: SD, obs_token, action_token, prompt_token, prompt_token_mask, *, n_head: int):
    """VIMAGatoPolicy.forward, vima_gato_policy.py:120-191."""
    T, B, Q, E = obs_token.shape
    Lp = prompt_token.shape[0]
    hist, _, _ = assemble_history(obs_token, torch.ones(T, B, Q, dtype=torch.bool), action_token)
    tokens = torch.cat([prompt_token, sd["prompt_sep_token"].expand(1, B, E), hist], dim=0)
    L = tokens.shape[0]
    mask = torch.cat([prompt_token_mask, torch.ones(B, L - Lp, dtype=torch.bool)], dim=1)
    n_valid = prompt_token_mask.sum(dim=1)
    ids = []
    for n in n_valid.tolist():
        ids.append(torch.cat([torch.arange(n), torch.full((Lp - n,), n - 1), torch.arange(n, n + L - Lp)]))
    out = hfgpt_forward(sd, "transformer.", tokens, mask, torch.stack(ids).long(), n_head)
    return out[Lp + 1 + Q - 1 :: Q + 1]


# ------------------------------------------------------------------------------------------------
# VIMA-GPT baseline (decoder-only, ONE token per observation / prompt image; vima/policy/vima_gpt_policy.py)
# ------------------------------------------------------------------------------------------------
def rect_vit_forward(sd: SD, p: str, img: torch.Tensor, heads: int = 24, patch: int = 32) -> torch.Tensor:
    """VisionTransformerRectangular.forward, vima/nn/obj_encoder/vit/vit.py:310-330: CLS token + patch tokens, the CLS row
    after ln_post @ projection is the image feature."""
    N = img.shape[0]
    w = sd[p + "conv1.weight"]
    width = w.shape[0]
    patches = F.unfold(img, kernel_size=patch, stride=patch).transpose(1, 2)
    x = _mm(patches, w.reshape(width, -1).t())
    x = torch.cat([sd[p + "cls_token"].expand(N, 1, width), x], dim=1) + sd[p + "pos_embed"]
    x = layer_norm(x, sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"])
    S = x.shape[1]
    d = width // heads
    i = 0
    while f"{p}blocks.{i}.ln_1.weight" in sd:
        b = f"{p}blocks.{i}."
        y = layer_norm(x, sd[b + "ln_1.weight"], sd[b + "ln_1.bias"])
        q, k, v = linear(y, sd[b + "attn.in_proj_weight"], sd[b + "attn.in_proj_bias"]).split(width, dim=-1)
        q = q.view(N, S, heads, d).transpose(1, 2)
        k = k.view(N, S, heads, d).transpose(1, 2)
        v = v.view(N, S, heads, d).transpose(1, 2)
        att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
        o = torch.matmul(att, v).transpose(1, 2).reshape(N, S, width)
        x = x + linear(o, sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"])
        y = layer_norm(x, sd[b + "ln_2.weight"], sd[b + "ln_2.bias"])
        h = linear(y, sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)
        x = x + linear(h, sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"])
        i += 1
    x = layer_norm(x[:, 0, :], sd[p + "ln_post.weight"], sd[p + "ln_post.bias"])
    return _mm(x, sd[p + "projection"])


def multiview_rgb_encoder(sd: SD, p: str, rgb: dict) -> torch.Tensor:
    """MultiViewRGBEncoder.forward, vima/nn/obj_encoder/obj_encoder.py:236-242: per-view CLS features concatenated on the
    FEATURE axis -> (..., 2 * emb_dim)."""
    outs = []
    for view in VIEWS:
        img = rgb[view]
        lead = img.shape[:-3]
        f = rect_vit_forward(sd, p + "cropped_img_encoder.vit.", image_preprocess(img).flatten(0, img.dim() - 4))
        outs.append(f.view(*lead, f.shape[-1]))
    return torch.cat(outs, dim=-1)


def gpt_forward_prompt_assembly(sd: SD, prompts):
    """VIMAGPTPolicy.forward_prompt_assembly, vima/policy/vima_gpt_policy.py:178-238 (one token per prompt image)."""
    token_types, word_batch, image_batch = prompts
    word_emb = sd["prompt_embedding._embed_layer.weight"][word_batch]
    img_emb = mlp_seq(sd, "prompt_obj_post_layer.", multiview_rgb_encoder(sd, "obj_encoder.", image_batch["rgb"]), (0, 3, 6))
    B, L_max = len(token_types), max(len(tt) for tt in token_types)
    toks = torch.zeros(B, L_max, img_emb.shape[-1])
    masks = torch.zeros(B, L_max, dtype=torch.bool)
    wp = ip = 0
    for b, tt in enumerate(token_types):
        for pos, t in enumerate(tt):
            if t == 0:
                toks[b, pos] = word_emb[wp]; wp += 1
            else:
                toks[b, pos] = img_emb[ip]; ip += 1
        masks[b, :len(tt)] = True
    enc = t5_encoder_forward(sd, "t5_prompt_encoder.t5.encoder.", toks, masks)
    if "t5_prompt_encoder_post_layer.weight" in sd:
        enc = linear(enc, sd["t5_prompt_encoder_post_layer.weight"])
    return enc.transpose(0, 1), masks


def gpt_forward_obs_token(sd: SD, obs):
    """VIMAGPTPolicy.forward_obs_token, vima_gpt_policy.py:240-251 -> (T, B, E)."""
    img_feats = multiview_rgb_encoder(sd, "obj_encoder.", obs["rgb"])
    ee_feats = sd["end_effector_encoder.weight"][obs["ee"]]
    return linear(torch.cat([img_feats, ee_feats], dim=-1), sd["obs_fusion_layer.weight"], sd["obs_fusion_layer.bias"])


def gpt_policy_forward(sd: SD, obs_token, action_token, prompt_token, prompt_token_mask, *, n_head: int):
    """VIMAGPTPolicy.forward, vima_gpt_policy.py:119-176: [prompt | sep | o0 a0 o1 a1 ...], predictions at the obs rows."""
    return gato_policy_forward(sd, obs_token.unsqueeze(2), action_token, prompt_token, prompt_token_mask, n_head=n_head)


# ------------------------------------------------------------------------------------------------
# VIMA-Flamingo baseline: XAttnGPT decoder over Perceiver-resampled image tokens (vima/policy/vima_flamingo_policy.py)
# ------------------------------------------------------------------------------------------------
def _perceiver_layer(sd: SD, p: str, x: torch.Tensor, inputs: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """HF:modeling_perceiver.py PerceiverLayer.forward (PerceiverSelfAttention + PerceiverSelfOutput + query residual, then
    LayerNorm -> dense1 -> GELU -> dense2 + residual).  x (N,Lq,E) latents; inputs (N,Lk,E) for the cross-attention layer."""
    a = p + "attention.self."
    h = layer_norm(x, sd[a + "layernorm1.weight"], sd[a + "layernorm1.bias"])
    kv = h
    if inputs is not None:
        kv = layer_norm(inputs, sd[a + "layernorm2.weight"], sd[a + "layernorm2.bias"])
    q = linear(h, sd[a + "query.weight"], sd[a + "query.bias"])
    k = linear(kv, sd[a + "key.weight"], sd[a + "key.bias"])
    v = linear(kv, sd[a + "value.weight"], sd[a + "value.bias"])
    N, Lq, E = q.shape
    Lk, d = k.shape[1], E // heads
    q = q.view(N, Lq, heads, d).transpose(1, 2)
    k = k.view(N, Lk, heads, d).transpose(1, 2)
    v = v.view(N, Lk, heads, d).transpose(1, 2)
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d), dim=-1)  # inputs mask is all ones (obj_encoder.py:199-203)
    ctx = torch.matmul(att, v).transpose(1, 2).reshape(N, Lq, E)
    x = linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]) + x  # use_query_residual
    h = layer_norm(x, sd[p + "layernorm.weight"], sd[p + "layernorm.bias"])
    h = F.gelu(linear(h, sd[p + "mlp.dense1.weight"], sd[p + "mlp.dense1.bias"]))
    return linear(h, sd[p + "mlp.dense2.weight"], sd[p + "mlp.dense2.bias"]) + x


def perceiver_forward(sd: SD, p: str, inputs: torch.Tensor, *, num_blocks: int = 4, heads: int = 8) -> torch.Tensor:
    """ObjectsPerceiverEncoder.forward = HF PerceiverModel(inputs, all-ones mask).last_hidden_state
    (vima/nn/obj_encoder/perceiver/perceiver.py:11-41; HF PerceiverEncoder.forward: one cross-attention layer, then the SAME
    stack of self-attention layers applied `num_blocks` times).  inputs (N,L,E) -> (N,num_latents,E)."""
    x = sd[p + "model.embeddings.latents"].expand(inputs.shape[0], -1, -1)
    x = _perceiver_layer(sd, p + "model.encoder.cross_attention.", x, inputs, heads)
    n_self = 0
    while f"{p}model.encoder.self_attends.{n_self}.layernorm.weight" in sd:
        n_self += 1
    for _ in range(num_blocks):
        for i in range(n_self):
            x = _perceiver_layer(sd, f"{p}model.encoder.self_attends.{i}.", x, None, heads)
    return x


def perceiver_obj_encoder(sd: SD, p: str, rgb: dict) -> torch.Tensor:
    """MultiViewRGBPerceiverEncoder.forward, vima/nn/obj_encoder/obj_encoder.py:192-205: Gato ViT patch tokens of both views
    (16 per image) resampled to 4 latent tokens.  (The reference spells the sub-module `peceiver`.)"""
    feats = gato_obj_encoder(sd, p, rgb)  # (..., 16, E)
    lead = feats.shape[:-2]
    out = perceiver_forward(sd, p + "peceiver.", feats.reshape(-1, *feats.shape[-2:]))
    return out.view(*lead, *out.shape[-2:])


def flamingo_forward_prompt_assembly(sd: SD, prompts):
    """VIMAFlamingoPolicy.forward_prompt_assembly, vima_flamingo_policy.py:165-228 (4 tokens per prompt image)."""
    token_types, word_batch, image_batch = prompts
    word_emb = sd["prompt_embedding._embed_layer.weight"][word_batch]
    img_emb = mlp_seq(sd, "prompt_obj_post_layer.", perceiver_obj_encoder(sd, "obj_encoder.", image_batch["rgb"]), (0, 3, 6))
    nq = img_emb.shape[-2]
    lens = [sum(1 if t == 0 else nq for t in tt) for tt in token_types]
    B, L_max = len(token_types), max(lens)
    toks = torch.zeros(B, L_max, img_emb.shape[-1])
    masks = torch.zeros(B, L_max, dtype=torch.bool)
    wp = ip = 0
    for b, tt in enumerate(token_types):
        pos = 0
        for t in tt:
            if t == 0:
                toks[b, pos] = word_emb[wp]; wp += 1; pos += 1
            else:
                toks[b, pos:pos + nq] = img_emb[ip]; ip += 1; pos += nq
        masks[b, :pos] = True
    enc = t5_encoder_forward(sd, "t5_prompt_encoder.t5.encoder.", toks, masks)
    if "t5_prompt_encoder_post_layer.weight" in sd:
        enc = linear(enc, sd["t5_prompt_encoder_post_layer.weight"])
    return enc.transpose(0, 1), masks


def flamingo_forward_obs_token(sd: SD, obs):
    """VIMAFlamingoPolicy.forward_obs_token, vima_flamingo_policy.py:230-240 -> (T, B, 4, E)."""
    img_feats = perceiver_obj_encoder(sd, "obj_encoder.", obs["rgb"])
    ee_feats = sd["end_effector_encoder.weight"][obs["ee"]].unsqueeze(2).repeat(1, 1, img_feats.shape[-2], 1)
    return linear(torch.cat([img_feats, ee_feats], dim=-1), sd["obs_fusion_layer.weight"], sd["obs_fusion_layer.bias"])


def flamingo_policy_forward(sd: SD, obs_token, action_token, prompt_token, prompt_token_mask, *, n_head: int, xattn_n_head: int):
    """VIMAFlamingoPolicy.forward, vima_flamingo_policy.py:129-163: the VIMA history layout with every token valid and DEFAULT
    position ids (arange over the history and over the padded prompt; no cumsum of masks here)."""
    T, B, Q, E = obs_token.shape
    E2, n_layer = policy_dims(sd)
    tokens, masks, _ = assemble_history(obs_token, torch.ones(T, B, Q, dtype=torch.bool), action_token)
    L, Lp = tokens.shape[0], prompt_token.shape[0]
    out = xattn_gpt_forward(
        sd, "xattn_gpt.", obs_action_tokens=tokens, obs_action_position_ids=torch.arange(L).expand(B, L), prompt_tokens=prompt_token,
        prompt_mask=prompt_token_mask, prompt_position_ids=torch.arange(Lp).expand(B, Lp), obs_action_masks=masks.transpose(0, 1),
        n_layer=n_layer, n_head=n_head, xattn_n_head=xattn_n_head)
    return out[Q - 1 :: Q + 1]
