"""B200-native (sm_100a) vectorised multi-agent collision-avoidance environment.

Drop-in for the env.step() hot path of mit-acl/rl_collision_avoidance.
The compute lives in csrc/ (hand-written CUDA behind the C-ABI of include/ca_step.h);
this package is the thin Python host side.  There is no CPU fallback: the CUDA library
must be built (python -c "import __graft_entry__ as g; g.build()") and a GPU present.
"""
__version__ = "0.1.0"
