"""VIMAGatoPolicy: the decoder-only baseline (reference: /root/reference/vima/policy/vima_gato_policy.py:11-326).

One causal sequence [encoded prompt | separator | interleaved obs / action history] through `HFGPT` -- BASELINE.json
configs[4], the causal-only kernel path.  Same surface as the reference (constructor, sub-module names, methods);
the reference's `self.device` bug (:133) does not exist here -- the device comes from the inputs.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .. import engine as eng
from .. import nn as vnn
from ..utils import *  # noqa: F401,F403
from .vima_policy import VIMAPolicy


class VIMAGatoPolicy(nn.Module):
    def __init__(self, *, embed_dim: int, vocab_size=40478, n_positions=512, n_layer=12, n_head=12, dropout: float = 0.1):
        super().__init__()
        self.embed_dim = embed_dim
        self.transformer = vnn.HFGPT(n_embd=embed_dim, use_geglu=True, vocab_size=vocab_size, n_positions=n_positions, n_layer=n_layer,
                                     n_head=n_head, dropout=dropout)
        self.prompt_sep_token = nn.Parameter(torch.zeros(embed_dim))
        self.obj_encoder = vnn.GatoMultiViewRGBEncoder(emb_dim=embed_dim, views=["front", "top"], img_size=(64, 128), vit_patch_size=32,
                                                       vit_width=768, vit_layers=4, vit_heads=24)
        self._obj_xf_num_queries = self.obj_encoder.img_patch_len
        self.end_effector_encoder = vnn.Embedding(num_embeddings=2, embedding_dim=2)
        obs_feat_dim = self.obj_encoder.output_dim + 2
        self.obs_fusion_layer = nn.Identity() if obs_feat_dim == embed_dim else vnn.Linear(obs_feat_dim, embed_dim)
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

    @property
    def device(self):
        return next(self.parameters()).device

    # --------------------------------------------------------------------------------------------------
    def forward(self, obs_token: torch.Tensor, action_token: Optional[torch.Tensor], prompt_token: torch.Tensor, prompt_token_mask: torch.Tensor):
        """obs_token (T,B,Q,E), action_token (T-1,B,E)|None, prompt_token (Lp,B,E), prompt_token_mask (B,Lp) -> (T,B,E)
        (vima_gato_policy.py:120-191)."""
        ctx = eng.ctx_for(obs_token)
        T, B, Q, E = obs_token.shape
        assert Q == self._obj_xf_num_queries
        Lp = prompt_token.shape[0]
        La = 0 if action_token is None else action_token.shape[0]
        Ls = T * Q + La
        L = Lp + 1 + Ls
        dev = obs_token.device
        tokens = torch.empty((L, B, E), dtype=torch.float32, device=dev)
        tokens[:Lp].copy_(prompt_token)
        tokens[Lp].copy_(self.prompt_sep_token.detach().unsqueeze(0).expand(B, E))
        ones = torch.ones((T, B, Q), dtype=torch.uint8, device=dev)
        scratch_m = torch.empty((B, Ls), dtype=torch.uint8, device=dev)
        scratch_p = torch.empty((B, Ls), dtype=torch.int64, device=dev)
        ctx.assemble_history(obs_token.float().contiguous(), ones, None if action_token is None else action_token.float().contiguous(),
                             tokens[Lp + 1:], scratch_m, scratch_p)
        mask = torch.empty((B, L), dtype=torch.uint8, device=dev)
        position_ids = torch.empty((B, L), dtype=torch.int64, device=dev)
        ctx.gato_positions(eng.as_u8(prompt_token_mask), L, mask, position_ids)
        tokens_out = self.transformer(tokens, custom_mask=mask.view(torch.bool), batch_first=False, position_ids=position_ids)
        return tokens_out[Lp + 1 + Q - 1 :: Q + 1]

    def forward_prompt_assembly(self, prompts):
        """(token_types, word_batch, image_batch{"rgb": {view: (n_img,3,64,128)}}) -> (Lp,B,E), (B,Lp) bool (:193-251)."""
        raw_prompts_token_type, word_batch, image_batch = prompts
        ref = image_batch["rgb"][sorted(self._views)[0]]
        ctx = eng.ctx_for(ref)
        p = eng.prec()
        dev = ref.device
        nq = self._obj_xf_num_queries
        word_ids = word_batch.to(device=dev, dtype=torch.int64).contiguous()
        img_emb = self.prompt_obj_post_layer(self.obj_encoder(**image_batch))  # (n_img, nq, 768)
        D = img_emb.shape[-1]
        lens = []
        for raw in raw_prompts_token_type:
            n = 0
            for item in raw:
                if item == 0:
                    n += 1
                elif item == 1:
                    n += nq
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
                    kind[b, pos:pos + nq] = 2
                    index[b, pos:pos + nq] = ip * nq + np.arange(nq)
                    ip += 1
                    pos += nq
        tokens = torch.empty((B, Lp, D), dtype=torch.float32, device=dev)
        masks_u8 = torch.empty((B, Lp), dtype=torch.uint8, device=dev)
        all_valid = torch.ones(max(img_emb.shape[0] * nq, 1), dtype=torch.uint8, device=dev)
        ctx.gather_prompt(torch.from_numpy(kind).to(dev), torch.from_numpy(index).to(dev), word_ids,
                          self.prompt_embedding._embed_layer.weight.detach(), img_emb.reshape(-1, D).contiguous(), all_valid, B, Lp, D, tokens, masks_u8)
        prompt_masks = masks_u8.view(torch.bool)
        need_post = not isinstance(self.t5_prompt_encoder_post_layer, nn.Identity)
        out32, out16 = self.t5_prompt_encoder.encode(tokens, prompt_masks, want16=need_post)
        if need_post:
            pl = self.t5_prompt_encoder_post_layer
            pw = self._wc.get("t5post", (pl.weight,), lambda: eng.pack_linear(ctx, pl.weight, None, transposed=False, p=p))
            out32, _ = eng.gemm(ctx, out16, pw, p, want_f32=True)
        return out32.view(B, Lp, -1).transpose(0, 1), prompt_masks

    def forward_obs_token(self, obs):
        """obs {"rgb": {view: (T,B,3,64,128) u8}, "ee": (T,B)} -> (T,B,Q,E)  (:253-262)."""
        rgbs, ee = obs["rgb"], obs["ee"]
        lead = tuple(ee.shape[:2])
        ctx = eng.ctx_for(ee)
        p = eng.prec()
        img_feats = self.obj_encoder(rgb=rgbs)  # (T,B,Q,E)
        Q, E = img_feats.shape[-2], img_feats.shape[-1]
        if isinstance(self.obs_fusion_layer, nn.Identity):
            raise NotImplementedError("embed_dim == obs feature dim never happens for VIMA-Gato (E + 2 != E)")
        rows = lead[0] * lead[1] * Q
        a = eng.to_operand(ctx, img_feats.reshape(rows, E), p, pad_cols=E + 2)
        ctx.fill_ee(ee.to(torch.int64).contiguous(), self.end_effector_encoder.weight.detach().float().contiguous(), lead[0] * lead[1], Q,
                    a.hi, a.lo, E, 0, dtype=p.dtype)
        fl = self.obs_fusion_layer
        pw = self._wc.get("fusion", (fl.weight, fl.bias), lambda: eng.pack_linear(ctx, fl.weight, fl.bias, transposed=False, p=p))
        out32, _ = eng.gemm(ctx, a, pw, p, want_f32=True)
        return out32.view(*lead, Q, self.embed_dim)

    def forward_action_token(self, action):
        return self.action_encoder(self._de_discretize_actions(action))

    def forward_action_decoder(self, predicted_action_tokens: torch.Tensor):
        return self.action_decoder(predicted_action_tokens)

    discretize_action = VIMAPolicy.discretize_action
    _de_discretize_actions = VIMAPolicy._de_discretize_actions
    _bin_tensor = VIMAPolicy._bin_tensor
    postprocess_actions = VIMAPolicy.postprocess_actions
