"""The reference's `VIMAPolicy.state_dict()` contract: key -> shape (TEST INFRASTRUCTURE).

Written out by hand from the reference constructors (vima/policy/vima_policy.py:12-114 and the vima.nn
modules it wires up) so that tests on the GPU box -- where /root/reference does not exist -- can build
shared weights and check the product's `load_state_dict(strict=True)` compatibility.
`tests/test_state_dict_contract.py::test_spec_matches_reference` compares it with the real thing in the
build container.
"""
from __future__ import annotations

from collections import OrderedDict

ACTION_DIMS = OrderedDict(
    pose0_position=[50, 100], pose0_rotation=[50] * 4, pose1_position=[50, 100], pose1_rotation=[50] * 4
)
ACTION_IN = OrderedDict(pose0_position=2, pose0_rotation=4, pose1_position=2, pose1_rotation=4)


def xattn_gpt_spec(prefix: str, E: int, n_layer: int, n_positions: int = 512, xattn_n_positions: int = 256, geglu=True):
    sd = OrderedDict()
    sd[prefix + "position_ids"] = (n_positions,)
    sd[prefix + "xattn_position_ids"] = (xattn_n_positions,)
    sd[prefix + "positions_embed.weight"] = (n_positions, E)
    sd[prefix + "xattn_positions_embed.weight"] = (xattn_n_positions, E)
    for i in range(n_layer):
        h = f"{prefix}h.{i}."
        sd[h + "attn.bias"] = (1, 1, n_positions, n_positions)
        sd[h + "attn.c_attn.weight"] = (E, 3 * E)
        sd[h + "attn.c_attn.bias"] = (3 * E,)
        sd[h + "attn.c_proj.weight"] = (E, E)
        sd[h + "attn.c_proj.bias"] = (E,)
        sd[h + "ln_1.weight"] = (E,)
        sd[h + "ln_1.bias"] = (E,)
        sd[h + "mlp.c_fc.weight"] = (E, 4 * E)
        sd[h + "mlp.c_fc.bias"] = (4 * E,)
        sd[h + "mlp.c_proj.weight"] = (4 * E, E)
        sd[h + "mlp.c_proj.bias"] = (E,)
        if geglu:
            sd[h + "mlp.gated_layer.weight"] = (4 * E, E)
        sd[h + "ln_2.weight"] = (E,)
        sd[h + "ln_2.bias"] = (E,)
    for i in range(n_layer):
        x = f"{prefix}xattns.{i}."
        sd[x + "kv_position_ids"] = (xattn_n_positions,)
        sd[x + "layernorm.weight"] = (E,)
        sd[x + "layernorm.bias"] = (E,)
        sd[x + "query.weight"] = (E, E)
        sd[x + "key_value.weight"] = (2 * E, E)
        sd[x + "attention_out.weight"] = (E, E)
        sd[x + "ln.weight"] = (E,)
        sd[x + "ln.bias"] = (E,)
        sd[x + "linear1.weight"] = (4 * E, E)
        sd[x + "linear2.weight"] = (E, 4 * E)
        if geglu:
            sd[x + "gated_layer.weight"] = (4 * E, E)
    return sd


def vit_spec(prefix: str, width=768, layers=4, res=32, patch=16, out=768):
    sd = OrderedDict()
    sd[prefix + "cls_token"] = (width,)
    sd[prefix + "pos_embed"] = ((res // patch) ** 2 + 1, width)
    sd[prefix + "projection"] = (width, out)
    sd[prefix + "conv1.weight"] = (width, 3, patch, patch)
    sd[prefix + "ln_pre.weight"] = (width,)
    sd[prefix + "ln_pre.bias"] = (width,)
    for i in range(layers):
        b = f"{prefix}blocks.{i}."
        sd[b + "attn.in_proj_weight"] = (3 * width, width)
        sd[b + "attn.in_proj_bias"] = (3 * width,)
        sd[b + "attn.out_proj.weight"] = (width, width)
        sd[b + "attn.out_proj.bias"] = (width,)
        sd[b + "ln_1.weight"] = (width,)
        sd[b + "ln_1.bias"] = (width,)
        sd[b + "mlp.c_fc.weight"] = (4 * width, width)
        sd[b + "mlp.c_fc.bias"] = (4 * width,)
        sd[b + "mlp.c_proj.weight"] = (width, 4 * width)
        sd[b + "mlp.c_proj.bias"] = (width,)
        sd[b + "ln_2.weight"] = (width,)
        sd[b + "ln_2.bias"] = (width,)
    sd[prefix + "ln_post.weight"] = (width,)
    sd[prefix + "ln_post.bias"] = (width,)
    return sd


def mlp_spec(prefix: str, dims):
    """build_mlp Sequential: Linear at 0,3,6,... (Identity norm + ReLU in between)."""
    sd = OrderedDict()
    for j in range(len(dims) - 1):
        sd[f"{prefix}{3 * j}.weight"] = (dims[j + 1], dims[j])
        sd[f"{prefix}{3 * j}.bias"] = (dims[j + 1],)
    return sd


def t5_spec(prefix: str, d_model=768, d_ff=3072, n_layers=12, n_heads=12, d_kv=64, vocab=32128, buckets=32):
    sd = OrderedDict()
    sd[prefix + "shared.weight"] = (vocab, d_model)
    sd[prefix + "encoder.embed_tokens.weight"] = (vocab, d_model)
    for i in range(n_layers):
        b = f"{prefix}encoder.block.{i}."
        for n in "qkv":
            sd[b + f"layer.0.SelfAttention.{n}.weight"] = (n_heads * d_kv, d_model)
        sd[b + "layer.0.SelfAttention.o.weight"] = (d_model, n_heads * d_kv)
        if i == 0:
            sd[b + "layer.0.SelfAttention.relative_attention_bias.weight"] = (buckets, n_heads)
        sd[b + "layer.0.layer_norm.weight"] = (d_model,)
        sd[b + "layer.1.DenseReluDense.wi.weight"] = (d_ff, d_model)
        sd[b + "layer.1.DenseReluDense.wo.weight"] = (d_model, d_ff)
        sd[b + "layer.1.layer_norm.weight"] = (d_model,)
    sd[prefix + "encoder.final_layer_norm.weight"] = (d_model,)
    return sd


def state_dict_spec(*, embed_dim: int, xf_n_layers: int, sattn_n_heads: int = 0, xattn_n_heads: int = 0):
    E = embed_dim
    sd = OrderedDict()
    sd.update(xattn_gpt_spec("xattn_gpt.", E, xf_n_layers))
    sd.update(vit_spec("obj_encoder.cropped_img_encoder.vit."))
    for v in ("front", "top"):
        sd.update(mlp_spec(f"obj_encoder.bbox_mlp.{v}.", [4, 768, 768, 768]))
    for v in ("front", "top"):
        sd[f"obj_encoder.pre_transformer_layer.{v}.weight"] = (E, 1536)
        sd[f"obj_encoder.pre_transformer_layer.{v}.bias"] = (E,)
    sd["end_effector_encoder.weight"] = (2, 2)
    sd["obs_fusion_layer.weight"] = (E, E + 2)
    sd["obs_fusion_layer.bias"] = (E,)
    for k, n_in in ACTION_IN.items():
        sd.update(mlp_spec(f"action_encoder._embed_dict.{k}._layer.", [n_in, 256, 256]))
    sd["action_encoder._post_layer.weight"] = (E, 1024)
    sd["action_encoder._post_layer.bias"] = (E,)
    for k, dims in ACTION_DIMS.items():
        for j, n in enumerate(dims):
            sd.update(mlp_spec(f"action_decoder._decoders.{k}.mlps.{j}.", [E, 512, 512, n]))
    sd["prompt_embedding._embed_layer.weight"] = (32128, 768)
    sd.update(t5_spec("t5_prompt_encoder.t5."))
    if E != 768:
        sd["t5_prompt_encoder_post_layer.weight"] = (E, 768)
    sd.update(mlp_spec("prompt_obj_post_layer.", [E, 768, 768, 768]))
    return sd


def gato_state_dict_spec(*, embed_dim: int, n_layer: int, n_head: int = 0, vocab_size: int = 40478, n_positions: int = 512):
    """`VIMAGatoPolicy.state_dict()` of the reference under the transformers version in this image (no `attn.bias`)."""
    E = embed_dim
    sd = OrderedDict()
    sd["prompt_sep_token"] = (E,)
    sd["transformer.lm.position_ids"] = (n_positions,)
    sd["transformer.lm.tokens_embed.weight"] = (vocab_size, E)
    sd["transformer.lm.positions_embed.weight"] = (n_positions, E)
    for i in range(n_layer):
        h = f"transformer.lm.h.{i}."
        sd[h + "attn.c_attn.weight"] = (E, 3 * E)
        sd[h + "attn.c_attn.bias"] = (3 * E,)
        sd[h + "attn.c_proj.weight"] = (E, E)
        sd[h + "attn.c_proj.bias"] = (E,)
        sd[h + "ln_1.weight"] = (E,)
        sd[h + "ln_1.bias"] = (E,)
        sd[h + "mlp.c_fc.weight"] = (E, 4 * E)
        sd[h + "mlp.c_fc.bias"] = (4 * E,)
        sd[h + "mlp.c_proj.weight"] = (4 * E, E)
        sd[h + "mlp.c_proj.bias"] = (E,)
        sd[h + "mlp.gated_layer.weight"] = (4 * E, E)
        sd[h + "ln_2.weight"] = (E,)
        sd[h + "ln_2.bias"] = (E,)
    v = "obj_encoder.cropped_img_encoder.vit."
    full = vit_spec(v, width=768, layers=4, res=32, patch=16, out=E)
    sd[v + "pos_embed"] = (8, 768)
    sd[v + "projection"] = (768, E)
    sd[v + "conv1.weight"] = (768, 3, 32, 32)
    for k, shp in full.items():
        if k.split(".")[-2] in ("ln_pre",) or ".blocks." in k or ".ln_post." in k:
            sd[k] = shp
    sd["end_effector_encoder.weight"] = (2, 2)
    sd["obs_fusion_layer.weight"] = (E, E + 2)
    sd["obs_fusion_layer.bias"] = (E,)
    for k, n_in in ACTION_IN.items():
        sd.update(mlp_spec(f"action_encoder._embed_dict.{k}._layer.", [n_in, 256, 256]))
    sd["action_encoder._post_layer.weight"] = (E, 1024)
    sd["action_encoder._post_layer.bias"] = (E,)
    for k, dims in ACTION_DIMS.items():
        for j, n in enumerate(dims):
            sd.update(mlp_spec(f"action_decoder._decoders.{k}.mlps.{j}.", [E, 512, 512, n]))
    sd["prompt_embedding._embed_layer.weight"] = (32128, 768)
    sd.update(t5_spec("t5_prompt_encoder.t5."))
    if E != 768:
        sd["t5_prompt_encoder_post_layer.weight"] = (E, 768)
    sd.update(mlp_spec("prompt_obj_post_layer.", [E, 768, 768, 768]))
    return sd


def gpt_state_dict_spec(*, embed_dim: int, n_layer: int, n_head: int = 0, vocab_size: int = 40478, n_positions: int = 512):
    """`VIMAGPTPolicy.state_dict()` of the reference (vima/policy/vima_gpt_policy.py:10-117): the Gato layout with a CLS-token
    rectangular ViT (9 position rows), a 2E-wide image feature (views concatenated on the feature axis) and therefore a
    (2E + 2)-wide fusion layer and a 2E-wide prompt object MLP."""
    E = embed_dim
    sd = gato_state_dict_spec(embed_dim=embed_dim, n_layer=n_layer, n_head=n_head, vocab_size=vocab_size, n_positions=n_positions)
    v = "obj_encoder.cropped_img_encoder.vit."
    out = OrderedDict()
    for k, shp in sd.items():
        if k == v + "pos_embed":
            out[v + "cls_token"] = (768,)
            out[k] = (9, 768)
        elif k == "obs_fusion_layer.weight":
            out[k] = (E, 2 * E + 2)
        elif k.startswith("prompt_obj_post_layer."):
            continue
        else:
            out[k] = shp
    out.update(mlp_spec("prompt_obj_post_layer.", [2 * E, 768, 768, 768]))
    return out


def perceiver_spec(p: str, E: int, *, num_latents: int = 4, n_self: int = 4):
    """HF PerceiverModel keys as ObjectsPerceiverEncoder holds them (perceiver.py:25-38; widening factors 1, qk/v channels = E)."""
    sd = OrderedDict()
    sd[p + "model.embeddings.latents"] = (num_latents, E)

    def layer(q: str, cross: bool):
        a = q + "attention.self."
        for ln in ("layernorm1",) + (("layernorm2",) if cross else ()):
            sd[a + ln + ".weight"] = (E,)
            sd[a + ln + ".bias"] = (E,)
        for lin in ("query", "key", "value"):
            sd[a + lin + ".weight"] = (E, E)
            sd[a + lin + ".bias"] = (E,)
        for lin in ("attention.output.dense", "mlp.dense1", "mlp.dense2"):
            sd[q + lin + ".weight"] = (E, E)
            sd[q + lin + ".bias"] = (E,)
        sd[q + "layernorm.weight"] = (E,)
        sd[q + "layernorm.bias"] = (E,)

    layer(p + "model.encoder.cross_attention.", True)
    for i in range(n_self):
        layer(f"{p}model.encoder.self_attends.{i}.", False)
    return sd


def flamingo_state_dict_spec(*, embed_dim: int, dt_n_layers: int, dt_n_heads: int = 0, xattn_n_heads: int = 0):
    """`VIMAFlamingoPolicy.state_dict()` of the reference (vima/policy/vima_flamingo_policy.py:10-127): the XAttnGPT decoder of
    VIMAPolicy, the Gato ViT + HF Perceiver resampler as object encoder (sub-module spelled `peceiver`), Gato-style heads."""
    E = embed_dim
    sd = OrderedDict()
    sd.update(xattn_gpt_spec("xattn_gpt.", E, dt_n_layers))
    g = gato_state_dict_spec(embed_dim=E, n_layer=0)
    for k, shp in g.items():
        if k.startswith("obj_encoder."):
            sd[k] = shp
    sd.update(perceiver_spec("obj_encoder.peceiver.", E))
    for k, shp in g.items():
        if k.startswith(("end_effector_encoder", "obs_fusion_layer", "action_encoder", "action_decoder", "prompt_embedding", "t5_prompt_encoder",
                         "prompt_obj_post_layer")):
            sd[k] = shp
    return sd
