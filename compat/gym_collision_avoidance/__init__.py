"""Namespace shim: when the `gym` package is installed the environment is made available under the id the reference
uses (gym.make("CollisionAvoidance-v0")); without gym, import CollisionAvoidanceEnv from .envs.collision_avoidance_env."""
ENV_ID = "CollisionAvoidance-v0"
ENTRY_POINT = "%s.envs.collision_avoidance_env:CollisionAvoidanceEnv" % __name__


def _register_with_gym():
    try:
        import gym.envs.registration as registration
    except Exception:           # gym is optional here
        return False
    if ENV_ID not in getattr(registration, "registry", {}):
        registration.register(id=ENV_ID, entry_point=ENTRY_POINT)
    return True


_register_with_gym()
