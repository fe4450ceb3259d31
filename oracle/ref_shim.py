"""Import the UNMODIFIED reference (`/root/reference/vima`, or its staged copy `oracle/_ref/vima`).

TEST / BENCH INFRASTRUCTURE ONLY -- used by `tests/golden/make_golden.py` to mint golden vectors, by the `reference`-marked
tests, and by `bench.py`'s reference legs (`--impl reference`, `gpu_eager`).  `/root/reference` does not exist on the GPU
box; there the bytes-identical copy staged by the committed recipe `oracle/make_ref.py` (git-ignored build artefact,
verified against its sha256 manifest before import) is used instead.  Nothing under `vima_b200/` imports this file.

The reference does not import as-is under transformers 5.x / without kornia, dm-tree and network access.
The shims below restore the transformers-4.x symbols it imports and replace the two `from_pretrained`
calls with a random-initialised t5-base-shaped model.  No reference source file is modified or copied.

Shim list (reference file:line that needs it):
  * kornia                      vima/nn/obj_encoder/vit/preprocess.py:6      -> oracle/shims/kornia.py
  * tree (dm-tree)              vima/utils.py:4                              -> oracle/shims/tree.py
  * modeling_t5.{get_device_map,assert_device_map,checkpoint}
                                vima/nn/prompt_encoder/prompt_encoder.py:8-19 (dead model-parallel code)
  * openai Attention.forward(head_mask=...)   vima/nn/seq_modeling/xattn_gpt/components.py:24-29
  * PreTrainedModel.get_head_mask             vima/nn/prompt_encoder/prompt_encoder.py:329-332
  * from_pretrained("t5-base")                vima/nn/prompt_encoder/word_embd.py:11, prompt_encoder.py:26
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = "/root/reference"
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

T5_BASE = dict(
    vocab_size=32128,
    d_model=768,
    d_kv=64,
    d_ff=3072,
    num_layers=12,
    num_heads=12,
    relative_attention_num_buckets=32,
    relative_attention_max_distance=128,
    dropout_rate=0.1,
    layer_norm_epsilon=1e-6,
    feed_forward_proj="relu",
)

_loaded = None


def reference_root():
    """/root/reference when present (build container), else the verified staged copy, else None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "vima")):
        return REFERENCE_ROOT
    from . import make_ref

    if make_ref.verify():
        return STAGED_ROOT
    return None


def reference_available() -> bool:
    return reference_root() is not None


def load_reference():
    """Returns the reference's `vima` package (as module object), shimmed. Idempotent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present and no verified staged copy under oracle/_ref (run oracle/make_ref.py in the build container)")

    shim_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
    # our own drop-in alias package is also called `vima`; make sure the reference wins here
    for name in [m for m in sys.modules if m == "vima" or m.startswith("vima.")]:
        del sys.modules[name]
    sys.path.insert(0, root)
    sys.path.insert(0, shim_dir)

    import torch  # noqa: F401
    import torch.nn as nn
    import transformers.models.t5.modeling_t5 as mt5
    from torch.utils.checkpoint import checkpoint

    if not hasattr(mt5, "checkpoint"):
        mt5.checkpoint = checkpoint
    if not hasattr(mt5, "get_device_map"):
        mt5.get_device_map = lambda *a, **k: None
    if not hasattr(mt5, "assert_device_map"):
        mt5.assert_device_map = lambda *a, **k: None

    import transformers.models.openai.modeling_openai as moa

    def _attention_forward_4x(self, x, attention_mask=None, head_mask=None, output_attentions=False):
        # transformers 4.x signature of modeling_openai.Attention.forward
        x = self.c_attn(x)
        query, key, value = x.split(self.split_size, dim=2)
        query = self.split_heads(query)
        key = self.split_heads(key, k=True)
        value = self.split_heads(value)
        attn_outputs = self._attn(query, key, value, attention_mask, head_mask, output_attentions)
        a = attn_outputs[0]
        a = self.merge_heads(a)
        a = self.c_proj(a)
        a = self.resid_dropout(a)
        return [a] + attn_outputs[1:]

    moa.Attention.forward = _attention_forward_4x

    from transformers.modeling_utils import PreTrainedModel

    if not hasattr(PreTrainedModel, "get_head_mask"):
        PreTrainedModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n

    import vima as ref_vima  # noqa: E402  (the reference)
    from transformers import T5Config

    pe = sys.modules["vima.nn.prompt_encoder.prompt_encoder"]
    we = sys.modules["vima.nn.prompt_encoder.word_embd"]

    pe.T5EncoderModel.from_pretrained = classmethod(lambda cls, name, *a, **k: cls(T5Config(**T5_BASE)))

    class _FakeAutoModel:
        @staticmethod
        def from_pretrained(name, *a, **k):
            class _M:
                def get_input_embeddings(self):
                    return nn.Embedding(T5_BASE["vocab_size"], T5_BASE["d_model"])

            return _M()

    we.AutoModel = _FakeAutoModel

    from vima.policy.vima_gato_policy import VIMAGatoPolicy

    if not hasattr(VIMAGatoPolicy, "device"):
        VIMAGatoPolicy.device = property(lambda s: next(s.parameters()).device)
    from vima.policy.vima_gpt_policy import VIMAGPTPolicy

    if not hasattr(VIMAGPTPolicy, "device"):  # same missing attribute (vima_gpt_policy.py:131)
        VIMAGPTPolicy.device = property(lambda s: next(s.parameters()).device)
    from vima.policy.vima_flamingo_policy import VIMAFlamingoPolicy

    if not hasattr(VIMAFlamingoPolicy, "device"):  # vima_flamingo_policy.py:141
        VIMAFlamingoPolicy.device = property(lambda s: next(s.parameters()).device)

    assert ref_vima.__file__.startswith(root), ref_vima.__file__
    _loaded = ref_vima
    # leave sys.path as is: sub-imports inside the reference are lazy in places
    return ref_vima
