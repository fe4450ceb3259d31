"""Portable deterministic number generator for weights and synthetic inputs.

TEST / BENCH INFRASTRUCTURE.  The real VIMA checkpoints are not reachable offline, and HF `post_init`
differs between transformers versions, so parity is defined on *shared* weights: the same generator fills
the reference model (golden minting, in this container) and the product model (everywhere), keyed by the
state-dict key.  It is a counter-based SplitMix64 in numpy integer arithmetic, so it gives the same bits
on any machine / numpy / torch version.
"""
from __future__ import annotations

import re
import zlib

import numpy as np
import torch

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return z


def raw_u64(key: str, n: int, seed: int = 0) -> np.ndarray:
    h = np.uint64(zlib.crc32(key.encode()) & 0xFFFFFFFF)
    base = _splitmix64(np.array([(int(h) << 20) ^ (seed * 0x632BE5AB)], dtype=np.uint64))[0]
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return _splitmix64((idx * np.uint64(0xD1342543DE82EF95) + base) & _M64)


def uniform(key: str, shape, seed: int = 0, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    """float32 uniform in [lo, hi) with 24 random bits per value."""
    n = int(np.prod(shape)) if len(shape) else 1
    u = (raw_u64(key, n, seed) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    v = (lo + (hi - lo) * u).astype(np.float32)
    return torch.from_numpy(v.reshape(tuple(shape)))


def randint(key: str, shape, low: int, high: int, seed: int = 0, dtype=torch.int64) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    r = (raw_u64(key, n, seed) >> np.uint64(33)).astype(np.int64)
    v = low + (r % (high - low))
    return torch.from_numpy(v.reshape(tuple(shape))).to(dtype)


# ------------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------------
_CONV1D = re.compile(r"(^|\.)h\.\d+\.(attn|mlp)\.c_(attn|fc|proj)\.weight$")  # HF Conv1D: weight is [in, out]
_SKIP = re.compile(r"((^|\.)(position_ids|kv_position_ids|xattn_position_ids)|(^|\.)h\.\d+\.attn\.bias)$")  # structural buffers


def weight_for(key: str, shape, seed: int = 0) -> torch.Tensor | None:
    """Deterministic value for state-dict entry `key`; None for structural buffers that keep their own value."""
    if _SKIP.search(key):
        return None
    shape = tuple(shape)
    nd = len(shape)
    if key.endswith("t5.shared.weight") or key.endswith("embed_tokens.weight"):
        key = "T5.SHARED"  # HF aliases these two entries to one tensor
    if nd == 1:
        if key.endswith(".weight"):  # norm gains
            return 1.0 + 0.1 * uniform(key, shape, seed)
        if key.endswith("cls_token") or key.endswith("prompt_sep_token"):
            return 0.5 * uniform(key, shape, seed)
        return 0.1 * uniform(key, shape, seed)  # biases
    if "relative_attention_bias" in key:
        return 0.5 * uniform(key, shape, seed)
    if key.endswith("tokens_embed.weight"):
        return 0.02 * uniform(key, shape, seed)
    if "positions_embed" in key or key.endswith("pos_embed"):
        return 0.1 * uniform(key, shape, seed)
    if "end_effector_encoder" in key:
        return uniform(key, shape, seed)
    if "_embed_layer" in key or key == "T5.SHARED":
        return uniform(key, shape, seed)
    if nd == 4:  # conv1 [out, in, kh, kw]
        fan_in = shape[1] * shape[2] * shape[3]
    elif _CONV1D.search(key):
        fan_in = shape[0]
    elif key.endswith("vit.projection"):
        fan_in = shape[0]
    else:
        fan_in = shape[1]  # nn.Linear [out, in]
    gain = 1.0
    if key.endswith("SelfAttention.q.weight"):
        gain = 0.125  # T5 does not scale scores by 1/sqrt(d_kv); keep logits O(1)
    a = gain * (3.0 / fan_in) ** 0.5
    return a * uniform(key, shape, seed)


@torch.no_grad()
def fill_module_(module: torch.nn.Module, seed: int = 0) -> None:
    """Overwrite every parameter/buffer of `module` in place with the deterministic values."""
    sd = module.state_dict()
    for k, v in sd.items():
        w = weight_for(k, v.shape, seed)
        if w is not None:
            v.copy_(w.to(v.dtype))
