"""Size-independent properties of the batched env step at BASELINE.json's full sizes (configs 2, 3 and 5).

The oracle finishes a few hundred robots in seconds; at 4,096 / 65,536 / 1,048,576 robots parity is carried by
properties the domain offers: robots are independent, so a robot's result may depend neither on WHERE it sits in the
batch (k_rank sorts robots by contact load, k_solve packs eight of them per warp), nor on how many other robots the
context holds, nor on which context / shard owns it -- and all of it must hold BIT FOR BIT, since every robot runs the
same float32 instruction sequence.  A 64-robot sample of the large batch is then checked against the oracle
(teacher-forced, the tolerance of test_gpu_parity.py), which ties the whole batch to the checker.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _mk(n, **kw):
    from plen_ml_walk_b200.vec_env import PlenVecEnv
    return PlenVecEnv(n, device="cuda:0", **kw)


def _rollout(env, steps, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    env.reset()
    for _ in range(steps):
        env.step(torch.empty((env.num_envs, 18), device="cuda").uniform_(-1, 1, generator=g))


def _snapshot(env):
    return [t.clone() for t in env.get_state()]


def _step_out(env, act):
    obs, rew, done, info = env.step(act)
    term = torch.where(done[:, None], info["terminal_obs"], torch.zeros_like(obs))      # only written where done
    return obs.clone(), rew.clone(), done.clone(), info["timeout"].clone(), term


def _same(a, b):
    # bit-exact, NaN-aware (the cosine-similarity reward term may be 0/0 = NaN in the reference too, plen_env.py:932)
    return bool(torch.equal(a, b) or torch.equal(torch.nan_to_num(a.float(), nan=1234.5), torch.nan_to_num(b.float(), nan=1234.5)))


def test_result_does_not_depend_on_batch_position_config2():
    """4096 robots in mixed contact phases; a random permutation of the batch permutes the results bit for bit."""
    n = 4096
    env = _mk(n)
    _rollout(env, 37, seed=1)
    st = _snapshot(env)
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    act = torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g)
    ref = _step_out(env, act)
    ref_state = _snapshot(env)
    perm = torch.randperm(n, device="cuda", generator=g)
    env2 = _mk(n)
    env2.reset()
    env2.set_state(*[t[perm] for t in st])
    got = _step_out(env2, act[perm])
    for a, b in zip(ref, got):
        assert _same(a[perm], b)
    for a, b in zip(ref_state, _snapshot(env2)):
        assert _same(a[perm], b)
    assert 0.02 < ref[0][:, 24:26].mean().item() < 0.98      # the sample really mixes flight and contact


def test_shards_reproduce_the_single_context_result():
    """Config 5's sharding: two contexts of N/2 robots give exactly the rows one context of N gives (no cross-robot
    coupling, no dependence on the batch size or the tile a robot falls in)."""
    n = 6144 + 8                                              # not a multiple of the 1024-robot sort tile
    env = _mk(n)
    _rollout(env, 21, seed=3)
    st = _snapshot(env)
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    acts = [torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g) for _ in range(3)]
    outs = [_step_out(env, a) for a in acts]
    lo = 0
    for size in (n // 2 - 5, n - (n // 2 - 5)):               # ragged split
        sh = _mk(size)
        sh.reset()
        sh.set_state(*[t[lo:lo + size] for t in st])
        for a, ref in zip(acts, outs):
            got = _step_out(sh, a[lo:lo + size])
            for x, y in zip(ref, got):
                assert _same(x[lo:lo + size], y)
        lo += size


def test_step_is_deterministic_run_to_run():
    n = 8192
    res = []
    for _ in range(2):
        env = _mk(n)
        _rollout(env, 25, seed=5)
        res.append(_snapshot(env) + [env._obs.clone(), env._reward.clone()])
    for a, b in zip(*res):
        assert _same(a, b)


def test_box_contact_solver_instances_agree_bit_for_bit(monkeypatch):
    """Robots whose link boxes touch the ground are solved by the EXT instance of the PGS: inside k_solve for small ranges
    (rows read from global memory), in k_solve_x for large ones (rows staged in shared memory).  The choice (PLEN_MERGE_MAX,
    read at plen_create) must not change one bit; the rollout is long enough for most robots to have fallen."""
    n = 8192
    g = torch.Generator(device="cuda"); g.manual_seed(23)
    acts = [torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g) for _ in range(70)]
    res = []
    for merge_max in ("0", "1000000000"):
        monkeypatch.setenv("PLEN_MERGE_MAX", merge_max)
        env = _mk(n, auto_reset=False)
        env.reset()
        ever_done = torch.zeros(n, dtype=torch.bool, device="cuda")
        for a in acts:
            _, _, d, _ = env.step(a)
            ever_done |= d
        res.append(_snapshot(env) + [env._obs.clone(), env._reward.clone()])
        assert ever_done.float().mean().item() > 0.3          # plenty of robots went down (and kept lying there)
    for a, b in zip(*res):
        assert _same(a, b)


def test_one_million_robots_config5_full_size():
    """1,048,576 robots on ONE GPU (6.7 GB of solve records): the first 4096 robots reproduce a 4096-robot context bit
    for bit over 6 steps with auto-reset, and robots with identical inputs stay identical across the whole batch."""
    n, m = 1048576, 4096
    big, small = _mk(n), _mk(m)
    g = torch.Generator(device="cuda"); g.manual_seed(13)
    big.reset(); small.reset()
    for s in range(6):
        a_small = torch.empty((m, 18), device="cuda").uniform_(-1, 1, generator=g)
        a_big = a_small.repeat(n // m, 1)                     # robot i of every 4096-block gets the same action
        ob, rb, db, _ = big.step(a_big)
        os_, rs, ds, _ = small.step(a_small)
        assert _same(ob[:m], os_) and _same(rb[:m], rs) and _same(db[:m], ds)
        assert _same(ob.view(n // m, m, 26)[-1], os_)         # last block too (different tiles, different SMs)
        assert _same(rb.view(n // m, m).amax(0), rb.view(n // m, m).amin(0)) or torch.isnan(rb).any()
    big.close(); small.close()


def test_sample_of_large_batch_against_oracle_config3_size():
    """65,536 robots (config 3's size), random actions with auto-reset: a 64-robot sample from the middle of the batch
    after 30 steps is stepped once more by the oracle from the same state -- same bound as the small teacher-forced test."""
    from oracle.oracle import PlenOracle
    from parity_util import oracle_from_abi
    n, k = 65536, 64
    env = _mk(n)
    _rollout(env, 30, seed=17)
    idx = torch.arange(n // 2 - k // 2, n // 2 + k // 2, device="cuda")
    st = oracle_from_abi(*[t[idx].cpu().numpy() for t in env.get_state()])
    g = torch.Generator(device="cuda"); g.manual_seed(19)
    act = torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g)
    obs, rew, done, info = env.step(act)
    gobs = torch.where(done[:, None], info["terminal_obs"], obs)[idx].cpu().numpy()      # pre-reset observation where done
    o = PlenOracle(k, n_threads=4)
    o.reset()
    o.set_state(st)
    oo, orw, od, _ = o.step(act[idx].cpu().numpy().astype(np.float64), auto_reset=False)
    err = np.abs(gobs[:, :24] - oo[:, :24]).max(axis=1)
    assert np.median(err) < 1e-4
    assert (err < 1e-3).mean() >= 0.6
    assert (gobs[:, 24:26] != oo[:, 24:26]).mean() <= 0.05
    assert (done[idx].cpu().numpy() != od).mean() <= 0.05


def test_host_buffer_step_equals_device_step():
    """plen_step_host (pinned host buffers, the batch pipelined in ranges over several streams so the copies hide under
    the kernels) returns exactly what plen_step returns for the same robots: ragged N (ranges of 2048 + 2048 + 904)."""
    n = 5000
    a, b = _mk(n), _mk(n)
    _rollout(a, 12, seed=23)
    b.reset(); b.set_state(*_snapshot(a))
    g = torch.Generator(device="cuda"); g.manual_seed(29)
    for _ in range(3):
        act = torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g)
        obs, rew, done, info = a.step(act)
        torch.cuda.synchronize()
        h_act = act.cpu().pin_memory()
        h_obs = torch.empty((n, 26)).pin_memory(); h_rew = torch.empty(n).pin_memory()
        h_done = torch.empty(n, dtype=torch.uint8).pin_memory(); h_tmo = torch.empty(n, dtype=torch.uint8).pin_memory()
        b.step_host(h_act, h_obs, h_rew, h_done, h_tmo)
        bad = (torch.nan_to_num(obs.cpu(), nan=1234.5) != torch.nan_to_num(h_obs, nan=1234.5)).any(1)
        assert not bool(bad.any()), "step %d: %d robots differ, first %s" % (_, int(bad.sum()), bad.nonzero()[:8, 0].tolist())
        assert _same(rew.cpu(), h_rew)
        assert torch.equal(done.cpu(), h_done.bool()) and torch.equal(info["timeout"].cpu(), h_tmo.bool())
    for x, y in zip(_snapshot(a), _snapshot(b)):
        assert _same(x, y)


def test_host_step_orders_itself_after_unsynchronised_device_calls():
    """Ordering contract of plen_step_host (include/plen_b200.h): it runs on the context's private streams and must wait
    for whatever reset / set_state / step queued on the caller's stream -- no torch.cuda.synchronize() in between."""
    n = 9000                                            # ranges of 3072 + 3072 + 2856: two of them beyond plen_step's two ranges
    a, b = _mk(n), _mk(n)
    _rollout(a, 6, seed=31)
    snap = _snapshot(a)
    g = torch.Generator(device="cuda"); g.manual_seed(37)
    acts = [torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g) for _ in range(3)]
    h_acts = [x.cpu().pin_memory() for x in acts]
    torch.cuda.synchronize()
    h_obs = torch.empty((n, 26)).pin_memory(); h_rew = torch.empty(n).pin_memory()
    h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
    for rep in range(5):
        # reference sequence on ONE stream: reset -> set_state -> step -> step
        a.reset(); a.set_state(*snap)
        a.step(acts[0])
        obs, rew, done, _ = a.step(acts[1])
        obs, rew, done = obs.clone(), rew.clone(), done.clone()
        # same thing with the last step through the host path, queued right behind the device calls
        b.reset(); b.set_state(*snap)
        b.step(acts[0])
        b.step_host(h_acts[1], h_obs, h_rew, h_done)
        torch.cuda.synchronize()
        assert _same(obs.cpu(), h_obs) and _same(rew.cpu(), h_rew) and torch.equal(done.cpu(), h_done.bool()), rep


def test_numeric_guard_retires_a_robot_with_a_non_finite_state():
    """a robot whose state went NaN / Inf is reported done with reward -100, reset from the snapshot
    (also with auto_reset off) and counted; its neighbours in the batch / solver warp are untouched."""
    n = 64
    a, b = _mk(n, auto_reset=False), _mk(n, auto_reset=False)
    _rollout(a, 5, seed=41)
    qpos, qvel, aux = _snapshot(a)
    b.reset(); b.set_state(qpos, qvel, aux)
    # (a non-finite VELOCITY is not a fault by itself: the +-100 coordinate-velocity clamp -- Bullet's maxCoordinateVelocity,
    # fminf / fmaxf drop a NaN operand -- turns it into a finite one; the guard watches the pose)
    bad_qpos = qpos.clone(); bad_qpos[9, 7 + 1] = float("nan"); bad_qpos[40, 0] = float("inf")
    a.set_state(bad_qpos, qvel, aux)
    reset_obs = _mk(1).reset().clone()                   # the post-reset observation (a constant)
    act = torch.zeros((n, 18), device="cuda")
    oa, ra, da, _ = a.step(act)
    ob, rb, db, _ = b.step(act)
    ok = torch.ones(n, dtype=torch.bool, device="cuda"); ok[9] = ok[40] = False
    assert _same(oa[ok], ob[ok]) and _same(ra[ok], rb[ok]) and torch.equal(da[ok], db[ok])
    assert bool(da[9]) and bool(da[40]) and float(ra[9]) == -100.0 and float(ra[40]) == -100.0
    assert torch.isfinite(oa).all() and _same(oa[9], reset_obs[0]) and _same(oa[40], reset_obs[0])
    assert a.fault_count() == 2 and b.fault_count() == 0
    assert torch.isfinite(torch.cat([t.flatten() for t in a.get_state()[:2]])).all()


def test_box_foot_model_variant_stands_and_steps():
    """plen_new.urdf (box feet, no collision margin, +-1.0 rad limits) through the same kernels: the robot settles on its
    soles at reset (both feet in contact, torso height within 5 mm of the stock model's) and zero joint targets
    (joint_act: raw radians, the standing pose) hold it up."""
    from plen_ml_walk_b200.urdf_loader import packaged_model
    from plen_ml_walk_b200.vec_env import PlenVecEnv
    env = PlenVecEnv(64, device="cuda:0", auto_reset=False, joint_act=True, model=packaged_model("plen_new"))
    ref = PlenVecEnv(64, device="cuda:0", auto_reset=False, joint_act=True)
    o, r = env.reset().clone(), ref.reset().clone()
    assert torch.isfinite(o).all() and abs(float(o[0, 18]) - float(r[0, 18])) < 5e-3
    zero = torch.zeros((64, 18), device="cuda")
    for _ in range(60):
        o, rew, done, _ = env.step(zero)
    assert torch.isfinite(o).all() and not bool(done.any())
    assert float(o[0, 24]) == 1.0 and float(o[0, 25]) == 1.0 and float(o[0, 18]) > 0.14


def test_snapshot_to_disk_and_back_continues_bit_exact(tmp_path):
    """save_state / load_state: a run restored from disk into a fresh context continues exactly."""
    n = 3000
    a = _mk(n)
    # per-env domain-randomisation scales travel with the snapshot (they change the physics of every step after it)
    gs = torch.Generator(device="cuda"); gs.manual_seed(30)
    a.set_env_scales(friction=torch.empty(n, device="cuda").uniform_(0.6, 1.2, generator=gs),
                     motor_gain=torch.empty(n, device="cuda").uniform_(0.8, 1.1, generator=gs))
    _rollout(a, 15, seed=31)
    path = str(tmp_path / "plen_state.pt")
    a.save_state(path)
    b = _mk(n)
    b.reset()
    b.load_state(path)
    g = torch.Generator(device="cuda"); g.manual_seed(37)
    for _ in range(4):
        act = torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g)
        ra, rb = _step_out(a, act), _step_out(b, act)
        for x, y in zip(ra, rb):
            assert _same(x, y)
    with pytest.raises(ValueError):
        _mk(n + 1).load_state(path)


def test_long_permuted_rollout_stays_bit_exact_config5_shard():
    """131,072 robots (one config-5 shard) for 100 steps against a randomly PERMUTED copy of the batch, so every robot shares
    its solver warp with different neighbours: 13 M robot-steps, every record word of every robot bit-equal after every
    step.  This is the test that catches a converged (frozen) robot that is not an exact no-op while the rest of its warp
    keeps iterating (an impulse one ulp beyond its friction bound used to be re-clamped: ~1 robot per 5 M robot-steps)."""
    n, steps = 131072, 100
    a, b = _mk(n), _mk(n)
    g = torch.Generator(device="cuda"); g.manual_seed(41)
    perm = torch.randperm(n, device="cuda", generator=g)
    a.reset(); b.reset()
    for s in range(steps):
        act = torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g)
        a.step(act); b.step(act[perm])
        ra, rb = a.debug_records()[perm, :80], b.debug_records()[:, :80]
        same = (ra.view(torch.int32) == rb.view(torch.int32)).all(1)
        assert bool(same.all()), "step %d: %d robots differ from their permuted twins" % (s, int((~same).sum()))


@pytest.mark.parametrize("n", [1024, 4096])
def test_step_captured_in_a_cuda_graph_replays_bit_exact(n):
    """plen_step launches only kernels on the caller's stream (and, for 2,048-4,096 robots, on the context's streams behind
    an event fork / join), so a caller may capture it in a CUDA graph: 20 replays equal 20 direct calls bit for bit."""
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    acts = [torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g) for _ in range(20)]
    a, b = _mk(n), _mk(n)
    _rollout(a, 30, 5)
    _rollout(b, 30, 5)
    direct = [_step_out(a, x) for x in acts]
    static_act = acts[0].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    state0 = _snapshot(b)
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            out = b.step(static_act)
    torch.cuda.current_stream().wait_stream(side)
    b.set_state(*state0)                                        # capture runs nothing, but keep the start state explicit
    for k, x in enumerate(acts):
        static_act.copy_(x)
        graph.replay()
        obs, rew, done, info = out
        term = torch.where(done[:, None], info["terminal_obs"], torch.zeros_like(obs))
        for u, v in zip(direct[k], (obs, rew, done, b._timeout.bool(), term)):
            assert _same(u, v), "graph replay %d differs from the direct call" % k


def test_persistent_sole_manifold_properties(tmp_path):
    """sole_manifold = 1: the manifolds are per-robot state like everything else -- a permutation of the batch permutes the
    results bit for bit (auto-reset restores the snapshot's manifolds), a snapshot to disk and back continues exactly, and
    the option changes the physics (the default context differs after a few steps)."""
    n = 3000
    kw = dict(config_overrides={"sole_manifold": 1})
    a = _mk(n, **kw)
    _rollout(a, 40, seed=51)
    st, man = _snapshot(a), a.get_manifold().clone()
    assert 0 < float(man[:, 48:50].mean()) < 4 and int(man[:, 48:50].max()) == 4
    path = str(tmp_path / "plen_state_manifold.pt")
    a.save_state(path)
    g = torch.Generator(device="cuda"); g.manual_seed(53)
    acts = [torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g) for _ in range(6)]
    ref = [_step_out(a, x) for x in acts]
    ref_man = a.get_manifold().clone()
    perm = torch.randperm(n, device="cuda", generator=g)
    b = _mk(n, **kw)
    b.reset()
    b.set_state(*[t[perm] for t in st]); b.set_manifold(man[perm])
    for x, r in zip(acts, ref):
        got = _step_out(b, x[perm])
        for p, q in zip(r, got):
            assert _same(p[perm], q)
    assert _same(ref_man[perm], b.get_manifold())
    assert any(bool(r[2].any()) for r in ref)                       # some episodes ended: the auto-reset path ran
    c = _mk(n, **kw)
    c.reset()
    c.load_state(path)
    for x, r in zip(acts, ref):
        got = _step_out(c, x)
        for p, q in zip(r, got):
            assert _same(p, q)
    d = _mk(n)
    _rollout(d, 40, seed=51)
    assert not _same(_snapshot(d)[0], st[0])
    with pytest.raises(RuntimeError):
        d.get_manifold()
