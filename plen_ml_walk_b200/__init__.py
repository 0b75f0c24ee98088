"""B200-native batched PLEN walking environment (hot path of moribots/plen_ml_walk behind its Gym surface)."""


def register_gym():
    """Register the 1-env adapter under the reference's id -- `PlenWalkEnv-v1`, max_episode_steps=500
    (plen_bullet/src/plen_bullet/plen_env.py:15-19) -- when a gym / gymnasium package is importable, so
    `gym.make("PlenWalkEnv-v1", joint_act=...)` (plen_td3.py:43, walk_eval.py:35, trajectory_eval.py:37) resolves to this
    package.  gym is NOT a dependency: returns the module used, or None when neither is installed."""
    for name in ("gym", "gymnasium"):
        try:
            mod = __import__(name)
        except ImportError:
            continue
        try:
            mod.register(id="PlenWalkEnv-v1", entry_point="plen_ml_walk_b200.vec_env:PlenWalkEnv", max_episode_steps=500)
        except Exception as e:          # gym raises when the id is registered already (second import): keep the first
            if "regist" not in str(e).lower():
                raise
        return mod
    return None


try:
    register_gym()
except Exception:                       # a broken gym install must not take the package down with it
    pass
