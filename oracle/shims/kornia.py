"""Empty stand-in for `kornia` (TEST INFRASTRUCTURE ONLY).

The reference imports kornia at /root/reference/vima/nn/obj_encoder/vit/preprocess.py:6 but only
touches it in the resize branch (preprocess.py:30-36), which ViTEncoder never takes (no `shape`
argument, vit.py:42).  This file only lets the unmodified reference import in this container.
"""
