from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn  # noqa: F401
