"""Drop-in alias: `import vima` resolves to vima_b200 so scripts/example.py of the reference runs unchanged."""
import sys

import vima_b200 as _impl
import vima_b200.nn as _nn
import vima_b200.policy as _policy
import vima_b200.utils as _utils
from vima_b200 import *  # noqa: F401,F403
from vima_b200 import create_policy_from_ckpt  # noqa: F401

sys.modules[__name__ + ".nn"] = _nn
sys.modules[__name__ + ".policy"] = _policy
sys.modules[__name__ + ".utils"] = _utils
nn = _nn
policy = _policy
utils = _utils
