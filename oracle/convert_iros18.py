"""TEST INFRASTRUCTURE — converts the reference's trained GA3C-CADRL checkpoint
(GCA/envs/policies/GA3C_CADRL/checkpoints/IROS18/network_01900000.{index,data-00000-of-00001}) into
tests/golden/iros18_weights.npz with the product's own TensorFlow-checkpoint reader (ga3c/tf_checkpoint.py).
Build container only.  The weights are a fixture for the end-to-end policy test (a trained policy must reach its goals
in this environment), like the golden vectors they are derived from the reference, not part of the product."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from rl_collision_avoidance_b200.ga3c import tf_checkpoint  # noqa: E402

PREFIX = os.environ.get("CA_REFERENCE_ROOT", "/root/reference") + \
    "/gym-collision-avoidance/gym_collision_avoidance/envs/policies/GA3C_CADRL/checkpoints/IROS18/network_01900000"

if __name__ == "__main__":
    v = tf_checkpoint.network_variables(PREFIX)
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "iros18_weights.npz")
    np.savez_compressed(out, **{k.replace("/", "__"): a for k, a in v.items()})
    print("wrote", out, os.path.getsize(out) // 1024, "KB", {k: a.shape for k, a in v.items()})
