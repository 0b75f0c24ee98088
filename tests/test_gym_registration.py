"""The package registers its 1-env adapter under the reference's gym id when gym is importable (plen_env.py:15-19)."""
import sys
import types


def test_registers_under_the_reference_id_when_gym_is_present(monkeypatch):
    calls = []
    fake = types.ModuleType("gym")
    fake.register = lambda **kw: calls.append(kw)
    monkeypatch.setitem(sys.modules, "gym", fake)
    import plen_ml_walk_b200
    assert plen_ml_walk_b200.register_gym() is fake
    assert calls and all(c == {"id": "PlenWalkEnv-v1", "entry_point": "plen_ml_walk_b200.vec_env:PlenWalkEnv", "max_episode_steps": 500} for c in calls)


def test_no_gym_is_not_an_error(monkeypatch):
    monkeypatch.setitem(sys.modules, "gym", None)
    monkeypatch.setitem(sys.modules, "gymnasium", None)
    import plen_ml_walk_b200
    assert plen_ml_walk_b200.register_gym() is None
