"""VIMAPolicy: the drop-in policy class (reference: /root/reference/vima/policy/vima_policy.py:11-322).

Same constructor, sub-module names (state-dict prefixes) and five entry methods as the reference; every tensor op
between the inputs and the returned tensors runs in the sm_100a kernels of libvima_b200.so.  Host Python only builds
the prompt index map from `token_types` (lists of ints) and launches kernels; there are no per-token Python loops on
tensors and no host synchronisation after the first call's input checks.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import engine as eng
from .. import nn as vnn
from ..utils import *  # noqa: F401,F403  (the reference re-exports vima.utils here)


class VIMAPolicy(nn.Module):
    def __init__(self, *, embed_dim: int, xf_n_layers: int, sattn_n_heads: int, xattn_n_heads: int):
        super().__init__()
        self.embed_dim = embed_dim
        self.xattn_gpt = vnn.XAttnGPT(embed_dim, n_layer=xf_n_layers, n_head=sattn_n_heads, dropout=0.1, xattn_n_head=xattn_n_heads,
                                      xattn_ff_expanding=4, xattn_n_positions=256, use_geglu=True)
        self.obj_encoder = vnn.ObjEncoder(transformer_emb_dim=embed_dim, views=["front", "top"], vit_output_dim=768, vit_resolution=32,
                                          vit_patch_size=16, vit_width=768, vit_layers=4, vit_heads=24, bbox_mlp_hidden_dim=768,
                                          bbox_mlp_hidden_depth=2)
        self.end_effector_encoder = vnn.Embedding(num_embeddings=2, embedding_dim=2)
        self.obs_fusion_layer = vnn.Linear(self.obj_encoder.output_dim + 2, embed_dim)
        self.action_encoder = vnn.ActionEmbedding(
            output_dim=embed_dim,
            embed_dict={
                "pose0_position": vnn.ContinuousActionEmbedding(output_dim=256, input_dim=2, hidden_dim=256, hidden_depth=1),
                "pose0_rotation": vnn.ContinuousActionEmbedding(output_dim=256, input_dim=4, hidden_dim=256, hidden_depth=1),
                "pose1_position": vnn.ContinuousActionEmbedding(output_dim=256, input_dim=2, hidden_dim=256, hidden_depth=1),
                "pose1_rotation": vnn.ContinuousActionEmbedding(output_dim=256, input_dim=4, hidden_dim=256, hidden_depth=1),
            },
        )
        self.action_decoder = vnn.ActionDecoder(
            input_dim=embed_dim,
            action_dims={"pose0_position": [50, 100], "pose0_rotation": [50] * 4, "pose1_position": [50, 100], "pose1_rotation": [50] * 4},
            hidden_dim=512, hidden_depth=2, activation="relu", norm_type=None, last_layer_gain=0.01,
        )
        self.prompt_embedding = vnn.WordEmbedding()
        self.t5_prompt_encoder = vnn.T5PromptEncoder()
        self.t5_prompt_encoder_post_layer = (
            nn.Identity() if embed_dim == self.t5_prompt_encoder.output_dim else vnn.Linear(self.t5_prompt_encoder.output_dim, embed_dim, bias=False)
        )
        self.prompt_obj_post_layer = vnn.build_mlp(self.obj_encoder.output_dim, hidden_dim=768, output_dim=768, hidden_depth=2)
        self._views = ["front", "top"]
        self._n_discrete_x_bins = 50
        self._n_discrete_y_bins = 100
        self._n_discrete_z_bins = 50
        self._n_discrete_rot_bins = 50
        self._wc = eng.WeightCache()
        self._bins = {}

    # --------------------------------------------------------------------------------------------------
    def forward(self, obs_token: torch.Tensor, obs_mask: torch.Tensor, action_token: Optional[torch.Tensor], prompt_token: torch.Tensor,
                prompt_token_mask: torch.Tensor):
        """obs_token (T,B,Q,E), obs_mask (T,B,Q) bool, action_token (T-1,B,E)|None, prompt_token (Lp,B,E),
        prompt_token_mask (B,Lp) bool -> predicted action tokens (T,B,E)   (vima_policy.py:116-159)."""
        ctx = eng.ctx_for(obs_token)
        T, B, Q, E = obs_token.shape
        La = 0 if action_token is None else action_token.shape[0]
        L = T * Q + La
        dev = obs_token.device
        obs = obs_token.float().contiguous()
        act = None if action_token is None else action_token.float().contiguous()
        tokens = torch.empty((L, B, E), dtype=torch.float32, device=dev)
        masks_bl = torch.empty((B, L), dtype=torch.uint8, device=dev)
        pos_bl = torch.empty((B, L), dtype=torch.int64, device=dev)
        ctx.assemble_history(obs, eng.as_u8(obs_mask), act, tokens, masks_bl, pos_bl)
        pmask_u8 = eng.as_u8(prompt_token_mask)
        prompt_pos = torch.empty(pmask_u8.shape, dtype=torch.int64, device=dev)
        ctx.mask_cumsum(pmask_u8, prompt_pos)
        tokens_out = self.xattn_gpt(
            obs_action_tokens=tokens,
            prompt_tokens=prompt_token,
            prompt_mask=pmask_u8.view(torch.bool),
            obs_action_masks=masks_bl.view(torch.bool),
            obs_action_position_ids=pos_bl,
            prompt_position_ids=prompt_pos,
        )
        return tokens_out[Q - 1 :: Q + 1]

    # --------------------------------------------------------------------------------------------------
    # Incremental decode (not in the reference, which re-runs the whole history each step).
    def start_decode(self, prompt_token: torch.Tensor, prompt_token_mask: torch.Tensor, *, max_tokens: Optional[int] = None):
        """Open a K/V cache for a batch of episodes: prompt_token (Lp,B,E), prompt_token_mask (B,Lp); room for `max_tokens`
        history tokens per episode (default: the decoder's n_positions).  Feed it to `forward_step` once per environment step."""
        Lp, B, E = prompt_token.shape
        ctx = eng.ctx_for(prompt_token)
        Lmax = self.xattn_gpt.n_positions if max_tokens is None else int(max_tokens)
        if not 0 < Lmax <= self.xattn_gpt.n_positions:
            raise ValueError(f"max_tokens={Lmax} outside (0, n_positions={self.xattn_gpt.n_positions}]")
        cache = vnn.DecodeCache(B=B, Lmax=Lmax, E=E, n_layer=self.xattn_gpt.n_layer, device=prompt_token.device, split=eng.prec().split,
                                precision=eng.prec().name)
        pmask_u8 = eng.as_u8(prompt_token_mask)
        cache.prompt = (prompt_token, pmask_u8, torch.empty(pmask_u8.shape, dtype=torch.int64, device=prompt_token.device))
        ctx.mask_cumsum(pmask_u8, cache.prompt[2])
        return cache

    def forward_step(self, cache, obs_token: torch.Tensor, obs_mask: torch.Tensor, prev_action_token: Optional[torch.Tensor]):
        """One environment step through the cache: obs_token (1,B,Q,E), obs_mask (1,B,Q), prev_action_token (1,B,E) (None at
        the first step) -> predicted action token (1,B,E); equals `forward(...)[-1:]` over the whole history.  Q may differ
        from step to step: scripts/example.py:139-171 re-pads every earlier step to the running maximum, but padded slots are
        masked keys with weight exactly 0 that do not advance position ids, so earlier steps can stay at the width they were
        appended with.  The prediction is read at the LAST slot passed (a padded one if the caller padded), as the reference does."""
        ctx = eng.ctx_for(obs_token)
        _, B, Q, E = obs_token.shape
        if (prev_action_token is None) != (cache.L == 0):
            raise ValueError("forward_step: exactly one action token per previous step is required")
        dev = obs_token.device
        # every refusal happens before the cache is touched (a failed step must leave n_valid / L consistent)
        from ..nn.xattn_gpt import check_cache_append

        check_cache_append(cache, B, Q + (0 if prev_action_token is None else 1), E, eng.prec())
        new = obs_token[0].float().transpose(0, 1)  # (Q,B,E)
        m_new = eng.as_u8(obs_mask[0])
        if prev_action_token is not None:
            new = torch.cat([prev_action_token.float(), new], dim=0)
            m_new = torch.cat([torch.ones((B, 1), dtype=torch.uint8, device=dev), m_new], dim=1)
        new, m_new = new.contiguous(), m_new.contiguous()
        pos = torch.empty(m_new.shape, dtype=torch.int64, device=dev)
        ctx.mask_cumsum(m_new, pos)
        pos += cache.n_valid[:, None]
        prompt_token, pmask_u8, prompt_pos = cache.prompt
        out = self.xattn_gpt(obs_action_tokens=new, prompt_tokens=prompt_token, prompt_mask=pmask_u8.view(torch.bool),
                             obs_action_masks=m_new.view(torch.bool), obs_action_position_ids=pos, prompt_position_ids=prompt_pos,
                             cache=cache)
        cache.n_valid += m_new.sum(dim=1)  # only after the decoder accepted the step
        return out[-1:]

    # --------------------------------------------------------------------------------------------------
    def forward_prompt_assembly(self, prompts):
        """(token_types, word_batch, image_batch) -> prompt tokens (Lp,B,E), masks (B,Lp) bool  (vima_policy.py:161-240)."""
        raw_prompts_token_type, word_batch, image_batch = prompts
        ref = image_batch["cropped_img"][sorted(self._views)[0]]
        ctx = eng.ctx_for(ref)
        p = eng.prec()
        dev = ref.device
        word_ids = word_batch.to(device=dev, dtype=torch.int64).contiguous()
        img_feats = self.obj_encoder(**image_batch)                       # (n_img, Q, E) fp32
        n_img, n_max_objs = img_feats.shape[0], img_feats.shape[-2]
        img_emb = self.prompt_obj_post_layer(img_feats)                   # (n_img, Q, 768) fp32
        D = img_emb.shape[-1]
        obj_mask = torch.cat([image_batch["mask"][v].reshape(n_img, -1) for v in sorted(self._views)], dim=-1)

        # host: index map from the token-type lists (ints only)
        lens = []
        for raw in raw_prompts_token_type:
            n = 0
            for item in raw:
                if item == 0:
                    n += 1
                elif item == 1:
                    n += n_max_objs
                else:
                    raise ValueError(f"Invalid prompt token type {item}")
            lens.append(n)
        B, Lp = len(raw_prompts_token_type), max(lens)
        kind = np.zeros((B, Lp), dtype=np.int32)
        index = np.zeros((B, Lp), dtype=np.int32)
        wp = ip = 0
        for b, raw in enumerate(raw_prompts_token_type):
            pos = 0
            for item in raw:
                if item == 0:
                    kind[b, pos], index[b, pos] = 1, wp
                    wp += 1
                    pos += 1
                else:
                    kind[b, pos : pos + n_max_objs] = 2
                    index[b, pos : pos + n_max_objs] = ip * n_max_objs + np.arange(n_max_objs)
                    ip += 1
                    pos += n_max_objs
        kind_d = torch.from_numpy(kind).to(dev)
        index_d = torch.from_numpy(index).to(dev)
        tokens = torch.empty((B, Lp, D), dtype=torch.float32, device=dev)
        masks_u8 = torch.empty((B, Lp), dtype=torch.uint8, device=dev)
        ctx.gather_prompt(kind_d, index_d, word_ids, self.prompt_embedding._embed_layer.weight.detach(), img_emb.reshape(-1, D).contiguous(),
                          eng.as_u8(obj_mask.reshape(-1)), B, Lp, D, tokens, masks_u8)
        prompt_masks = masks_u8.view(torch.bool)
        if self.t5_prompt_encoder is None:
            return tokens.transpose(0, 1), prompt_masks
        need_post = not isinstance(self.t5_prompt_encoder_post_layer, nn.Identity)
        out32, out16 = self.t5_prompt_encoder.encode(tokens, prompt_masks, want16=need_post)
        if need_post:
            pl = self.t5_prompt_encoder_post_layer
            pw = self._wc.get("t5post", (pl.weight,), lambda: eng.pack_linear(ctx, pl.weight, None, transposed=False, p=p))
            out32, _ = eng.gemm(ctx, out16, pw, p, want_f32=True)
        prompt_tokens = out32.view(B, Lp, -1).transpose(0, 1)
        return prompt_tokens, prompt_masks

    # --------------------------------------------------------------------------------------------------
    def forward_obs_token(self, obs):
        """obs {"ee": (T,B) i64, "objects": {cropped_img,bbox,mask}x{front,top}} -> (T,B,Q,E), (T,B,Q) bool  (:242-259)."""
        objects, ee = obs["objects"], obs["ee"]
        lead = tuple(ee.shape[:2])
        ctx = eng.ctx_for(ee)
        p = eng.prec()
        img_feats = self.obj_encoder(cropped_img=objects["cropped_img"], bbox=objects["bbox"], mask=objects["mask"])  # (T,B,Q,E)
        Q, E = img_feats.shape[-2], img_feats.shape[-1]
        rows = lead[0] * lead[1] * Q
        a = eng.to_operand(ctx, img_feats.reshape(rows, E), p, pad_cols=E + 2)   # [rows, E+8]: zero padded
        ctx.fill_ee(ee.to(torch.int64).contiguous(), self.end_effector_encoder.weight.detach().float().contiguous(), lead[0] * lead[1], Q,
                    a.hi, a.lo, E, 0, dtype=p.dtype)
        fl = self.obs_fusion_layer
        pw = self._wc.get("fusion", (fl.weight, fl.bias), lambda: eng.pack_linear(ctx, fl.weight, fl.bias, transposed=False, p=p))
        out32, _ = eng.gemm(ctx, a, pw, p, want_f32=True)
        obs_feats = out32.view(*lead, Q, self.embed_dim)
        obj_mask = torch.cat([objects["mask"][v].reshape(*lead, -1) for v in sorted(self._views)], dim=-1)
        return obs_feats, obj_mask

    # --------------------------------------------------------------------------------------------------
    def forward_action_token(self, action):
        return self.action_encoder(self._de_discretize_actions(action))

    def forward_action_decoder(self, predicted_action_tokens: torch.Tensor):
        return self.action_decoder(predicted_action_tokens)

    def discretize_action(self, action):
        """Training-side helper (vima_policy.py:267-299); not on the inference path."""
        device = action["pose0_position"].device
        bx = torch.linspace(0, 1, self._n_discrete_x_bins, device=device)
        by = torch.linspace(0, 1, self._n_discrete_y_bins, device=device)
        br = torch.linspace(0, 1, self._n_discrete_rot_bins, device=device)
        for k in ("pose0_position", "pose1_position"):
            action[k][..., 0] = torch.bucketize(action[k][..., 0].contiguous(), bx)
            action[k][..., 1] = torch.bucketize(action[k][..., 1].contiguous(), by)
        for k in ("pose0_rotation", "pose1_rotation"):
            action[k] = torch.bucketize(action[k].contiguous(), br)
        return {k: v.long() for k, v in action.items()}

    def postprocess_actions(self, actions, action_bounds_low: torch.Tensor, action_bounds_high: torch.Tensor):
        """The environment-facing step after the heads, scripts/example.py:199-232, in one kernel per action key: int64 bin
        indices -> de-discretised, scaled to the action bounds and clamped positions; rotations mapped to [-1, 1].
        Bounds: float32 (..., 2) broadcast over the leading dims of the position indices, or one row per episode."""
        out = {}
        for k, v in actions.items():
            ctx = eng.ctx_for(v)
            width = v.shape[-1]
            idx = v.to(torch.int64).contiguous()
            n = idx.numel() // width
            if k.endswith("position"):
                bins = self._bin_tensor(True, width, v.device)
                lo = action_bounds_low.to(device=v.device, dtype=torch.float32).reshape(-1, width).contiguous()
                hi = action_bounds_high.to(device=v.device, dtype=torch.float32).reshape(-1, width).contiguous()
                if lo.shape != hi.shape or lo.shape[0] not in (1, n):
                    raise ValueError(f"action bounds must hold 1 or {n} rows of {width} values, got {tuple(lo.shape)} / {tuple(hi.shape)}")
                stride = 0 if lo.shape[0] == 1 else width
            else:
                bins = self._bin_tensor(False, width, v.device)
                key = ("rot_bounds", width, str(v.device))
                if key not in self._bins:
                    self._bins[key] = (torch.full((1, width), -1.0, device=v.device), torch.full((1, width), 1.0, device=v.device))
                lo, hi = self._bins[key]
                stride = 0
            o = torch.empty(idx.shape, dtype=torch.float32, device=v.device)
            ctx.action_postprocess(idx.view(-1, width), n, width, bins, lo, hi, stride, o)
            out[k] = o
        return out

    def _bin_tensor(self, is_position: bool, width: int, device):
        key = (is_position, width, str(device))
        if key not in self._bins:
            b = [float(self._n_discrete_x_bins), float(self._n_discrete_y_bins)] if is_position else [float(self._n_discrete_rot_bins)] * width
            self._bins[key] = torch.tensor(b, dtype=torch.float32).to(device)
        return self._bins[key]

    def _de_discretize_actions(self, actions):
        """int64 indices -> float / bins  (vima_policy.py:301-322)."""
        out = {}
        for k, v in actions.items():
            ctx = eng.ctx_for(v)
            width = v.shape[-1]
            bins = self._bin_tensor(k.endswith("position"), width, v.device)
            idx = v.to(torch.int64).contiguous()
            o = torch.empty(idx.shape, dtype=torch.float32, device=v.device)
            ctx.action_scale(idx.view(-1, width), idx.numel() // width, width, bins, o)
            out[k] = o
        return out
