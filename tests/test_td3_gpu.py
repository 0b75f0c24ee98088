"""GPU: device replay ring, fused actor-forward kernel and the TD3 update against the reference semantics."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "td3_golden.npz")


def _tuples(lo, hi, dev):
    k = torch.arange(lo, hi, device=dev, dtype=torch.float32)
    return (k[:, None].expand(-1, 26).contiguous(), (k[:, None] + 0.25).expand(-1, 18).contiguous(),
            (k[:, None] + 0.5).expand(-1, 26).contiguous(), k.clone(), (k.long() % 3 == 0))


def test_replay_ring_matches_reference_order():
    """Batched adds of ragged sizes reproduce the reference's one-by-one ring (append, then overwrite from index 0)."""
    from oracle.replay_oracle import ReplayRing
    from plen_ml_walk_b200.td3 import ReplayBuffer
    g = np.load(GOLD)
    dev = torch.device("cuda:0")
    rb = ReplayBuffer(max_size=3, device=dev)
    for k in range(5):
        rb.add(*_tuples(k, k + 1, dev))
    assert rb.storage()[:, 70].cpu().tolist() == list(g["ring_rewards"]) and rb.ptr == int(g["ring_ptr"])
    cap = 1000
    rb, ring = ReplayBuffer(max_size=cap, device=dev), ReplayRing(cap)
    lo = 0
    for n in (1, 7, 640, 352, 0, 999, 1000, 2500, 3):        # fills, wraps, and adds larger than the ring
        rb.add(*_tuples(lo, lo + n, dev))
        for k in range(lo, lo + n):
            ring.add((k, k + 0.25, k + 0.5, float(k), float(k % 3 == 0)))
        lo += n
        st = rb.storage().cpu().numpy()
        assert len(rb) == len(ring.storage) and rb.ptr == ring.ptr
        assert (st[:, 70] == np.array([t[3] for t in ring.storage])).all()
        assert (st[:, 0] == st[:, 70]).all() and (st[:, 26] == st[:, 70] + 0.25).all() and (st[:, 44] == st[:, 70] + 0.5).all()
        assert (st[:, 71] == np.array([t[4] for t in ring.storage])).all()


def test_replay_sample_is_uniform_with_replacement_and_consistent():
    from plen_ml_walk_b200.td3 import ReplayBuffer
    dev = torch.device("cuda:0")
    rb = ReplayBuffer(max_size=4096, device=dev, seed=1)
    rb.add(*_tuples(0, 3000, dev))                                  # partially filled: only rows < len may be drawn
    s, a, s2, r, nd, idx = rb.sample(200000, return_index=True)
    idx = idx.long()
    assert int(idx.min()) >= 0 and int(idx.max()) < 3000
    assert torch.equal(r[:, 0], idx.float()) and torch.equal(s[:, 5], idx.float()) and torch.equal(a[:, 17], idx.float() + 0.25)
    assert torch.equal(s2[:, 25], idx.float() + 0.5) and torch.equal(nd[:, 0], 1.0 - (idx % 3 == 0).float())
    counts = torch.bincount(idx, minlength=3000).float()
    assert counts.min() > 20 and abs(float(counts.mean()) - 200000 / 3000) < 1e-3      # every row reachable
    assert float(counts.std()) < 1.3 * (200000 / 3000) ** 0.5                           # Poisson-like spread
    i1 = rb.sample(100, return_index=True)[5]
    i2 = rb.sample(100, return_index=True)[5]
    assert not torch.equal(i1, i2)                                                      # fresh draw per call
    with pytest.raises(RuntimeError):
        ReplayBuffer(max_size=8, device=dev).sample(4)                                  # empty buffer (np.random.randint(0, 0) raises)


def test_actor_forward_kernel_vs_reference_checkpoint_and_torch():
    from plen_ml_walk_b200.td3 import Actor, actor_forward
    g = np.load(GOLD)
    dev = torch.device("cuda:0")
    a = Actor().to(dev)
    a.load_state_dict({k: torch.from_numpy(g["actor_" + k.replace(".", "_")]) for k in a.state_dict().keys()})
    out = actor_forward(a, torch.from_numpy(g["obs"]).to(dev)).cpu().numpy()
    assert np.abs(out - g["actor_out"]).max() < 1e-5                 # vs the reference Actor on the shipped checkpoint
    torch.manual_seed(0)
    b = Actor(max_action=1.0).to(dev)
    for n in (1, 31, 32, 33, 4096, 70001):                           # ragged tails of the 32-row CTA tile
        obs = torch.randn(n, 26, device=dev)
        with torch.no_grad():
            ref = b(obs)
        got = actor_forward(b, obs)
        assert float((got - ref).abs().max()) < 1e-5, n
    # exploration noise of plen_td3.py:101-104: N(0, 0.1) added, then clipped to +-max_action
    obs = torch.zeros(200000, 26, device=dev)
    clean = actor_forward(b, obs)
    noisy = actor_forward(b, obs, noise_std=0.1, seed=7)
    d = (noisy - clean)
    inside = noisy.abs() < 0.999
    assert float(noisy.abs().max()) <= 1.0
    assert abs(float(d[inside].mean())) < 2e-3 and abs(float(d[inside].std()) - 0.1) < 5e-3
    assert not torch.equal(noisy, actor_forward(b, obs, noise_std=0.1, seed=8))


def _abi_lib():
    from plen_ml_walk_b200 import _abi
    return _abi.load_library()


def _twin_agents(dev, seed=0, precision="fp32", twin_precision="fp32"):
    """Two agents with identical parameters: one stepped by the CUDA learner, one by the PyTorch fp32 reference."""
    from plen_ml_walk_b200.td3 import TD3Agent
    torch.manual_seed(seed)
    a, b = TD3Agent(device=dev, precision=precision), TD3Agent(device=dev, precision=twin_precision)
    for k in ("actor", "actor_target", "critic", "critic_target"):
        b._flat[k].copy_(a._flat[k])
    return a, b


def test_td3_cuda_update_matches_pytorch_reference():
    """plen_td3_* (hand-written CUDA forward / backward / Adam / Polyak) against the same update in PyTorch autograd +
    torch.optim.Adam (the reference's td3.py:259-356 rule), same minibatch and the same policy-smoothing noise, over six
    consecutive updates (three of them policy updates).  fp32 on both sides: every parameter agrees to 1e-4 = a third
    of ONE Adam step (lr 3e-4: where a gradient entry is within a few eps of zero, m / (sqrt(v) + eps) amplifies the
    last-bit differences of the two summation orders), and the mean difference stays below 2e-6."""
    dev = torch.device("cuda:0")
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for B in (100, 37, 256, 1024):                               # the reference's batch, a ragged one, a tile multiple, split-K
            a, b = _twin_agents(dev, seed=B)
            g = torch.Generator(device=dev); g.manual_seed(B)
            for it in range(6):
                s = torch.randn(B, 26, device=dev, generator=g)
                ac = torch.rand(B, 18, device=dev, generator=g) * 2 - 1
                s2 = s + 0.05 * torch.randn(B, 26, device=dev, generator=g)
                r = torch.randn(B, 1, device=dev, generator=g)
                nd = (torch.rand(B, 1, device=dev, generator=g) > 0.1).float()
                nz = torch.randn(B, 18, device=dev, generator=g)
                al_c, cl_c = a.train(None, batch=(s, ac, s2, r, nd), noise=nz, return_losses=True)
                al_t, cl_t = b.train_torch((s, ac, s2, r, nd), noise=nz)
                assert abs(float(cl_c) - float(cl_t.detach())) <= 1e-4 * max(1.0, abs(float(cl_t.detach())))
                assert (al_c is None) == (al_t is None)
                if al_t is not None:
                    assert abs(float(al_c) - float(al_t.detach())) <= 1e-4 * max(1.0, abs(float(al_t.detach())))
                for k in ("critic", "actor", "critic_target", "actor_target"):
                    d = (a._flat[k] - b._flat[k]).abs()
                    assert float(d.max()) < 1e-4 and float(d.mean()) < 2e-6, (B, it, k, float(d.max()), float(d.mean()))
                if it == 0:
                    m_t = torch.cat([b.critic_optimizer.state[p]["exp_avg"].reshape(-1) for p in b.critic.parameters()])
                    v_t = torch.cat([b.critic_optimizer.state[p]["exp_avg_sq"].reshape(-1) for p in b.critic.parameters()])
                    assert float((a._adam["critic_m"] - m_t).abs().max()) < 1e-4 * float(m_t.abs().max()) + 1e-8
                    assert float((a._adam["critic_v"] - v_t).abs().max()) < 1e-4 * float(v_t.abs().max()) + 1e-10
            # Adam moments against torch's optimizer state: tight after the first update (gradients taken at identical
            # parameters, checked inside the loop), loose after six (parameters already differ by up to 1e-4)
            m_t = torch.cat([b.critic_optimizer.state[p]["exp_avg"].reshape(-1) for p in b.critic.parameters()])
            v_t = torch.cat([b.critic_optimizer.state[p]["exp_avg_sq"].reshape(-1) for p in b.critic.parameters()])
            assert float((a._adam["critic_m"] - m_t).abs().max()) < 2e-2 * float(m_t.abs().max())
            assert float((a._adam["critic_v"] - v_t).abs().max()) < 2e-2 * float(v_t.abs().max())
            extra = 1 if B >= 512 else 0                             # gradient memset ahead of the split-K dW products
            assert a.kernel_launches() == 6 * (1 + 15 + extra) + 3 * (15 + extra)      # set_batch + critic graph; actor graph on policy steps
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_td3_gradients_match_autograd():
    """critic_grad / actor_grad (the vectors a data-parallel learner all-reduces) against autograd, entry by entry."""
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    a, b = _twin_agents(dev, seed=3)
    B = 100
    g = torch.Generator(device=dev); g.manual_seed(5)
    s = torch.randn(B, 26, device=dev, generator=g); ac = torch.rand(B, 18, device=dev, generator=g) * 2 - 1
    s2 = torch.randn(B, 26, device=dev, generator=g); r = torch.randn(B, 1, device=dev, generator=g)
    nd = torch.ones(B, 1, device=dev); nz = torch.randn(B, 18, device=dev, generator=g)
    seen = {}
    a.total_it = 1                                                  # next call is a policy update
    a.train(None, batch=(s, ac, s2, r, nd), noise=nz, grad_hook=lambda gflat: seen.setdefault(gflat.numel(), gflat.clone()))
    with torch.no_grad():
        n2 = (nz * b.policy_noise).clamp(-b.noise_clip, b.noise_clip)
        a2 = (b.actor_target(s2) + n2).clamp(-1, 1)
        q1t, q2t = b.critic_target(s2, a2)
        y = r + nd * b.discount * torch.min(q1t, q2t)
    q1, q2 = b.critic(s, ac)
    loss = F.mse_loss(q1, y) + F.mse_loss(q2, y)
    gc = torch.cat([x.reshape(-1) for x in torch.autograd.grad(loss, list(b.critic.parameters()), retain_graph=True)])
    assert float((seen[155138] - gc).abs().max()) < 1e-5 * max(1.0, float(gc.abs().max()))
    # the actor gradient is taken after the critic's Adam step, as in the reference: apply the same step to b first
    b.critic_optimizer.zero_grad(); loss.backward(); b.critic_optimizer.step()
    la = -b.critic.Q1(s, b.actor(s)).mean()
    ga = torch.cat([x.reshape(-1) for x in torch.autograd.grad(la, list(b.actor.parameters()))])
    assert float((seen[77330] - ga).abs().max()) < 1e-5 * max(1.0, float(ga.abs().max()))


def _layer_slices(module):
    out, off = [], 0
    for name, p in module.named_parameters():
        out.append((name, off, off + p.numel()))
        off += p.numel()
    return out


@pytest.mark.parametrize("B", [1024, 4096, 1000])
def test_td3_tensor_core_gradients_vs_autograd(B):
    """precision="tf32": the learner's products on tcgen05 (kind::tf32).  Every gradient tensor (per layer, weight and bias)
    stays within 1e-3 of fp32 autograd relative to the tensor's largest entry, the losses within 1e-3 relative; B = 1000 has
    a ragged last 128-row tile.  The same call with precision="fp32" must hold 1e-5 (it does at B = 100 above)."""
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    a, b = _twin_agents(dev, seed=11, precision="tf32")
    g = torch.Generator(device=dev); g.manual_seed(17)
    s = torch.randn(B, 26, device=dev, generator=g); ac = torch.rand(B, 18, device=dev, generator=g) * 2 - 1
    s2 = torch.randn(B, 26, device=dev, generator=g); r = torch.randn(B, 1, device=dev, generator=g)
    nd = (torch.rand(B, 1, device=dev, generator=g) > 0.1).float(); nz = torch.randn(B, 18, device=dev, generator=g)
    seen = {}
    a.total_it = 1                                                  # next call is a policy update
    la_g, lc_g = a.train(None, batch=(s, ac, s2, r, nd), noise=nz, return_losses=True,
                         grad_hook=lambda gflat: seen.setdefault(gflat.numel(), gflat.clone()))
    with torch.no_grad():
        n2 = (nz * b.policy_noise).clamp(-b.noise_clip, b.noise_clip)
        a2 = (b.actor_target(s2) + n2).clamp(-1, 1)
        q1t, q2t = b.critic_target(s2, a2)
        y = r + nd * b.discount * torch.min(q1t, q2t)
    q1, q2 = b.critic(s, ac)
    loss = F.mse_loss(q1, y) + F.mse_loss(q2, y)
    gc = torch.cat([x.reshape(-1) for x in torch.autograd.grad(loss, list(b.critic.parameters()), retain_graph=True)])
    assert abs(float(lc_g) - float(loss)) < 1e-3 * max(1.0, abs(float(loss)))
    worst = 0.0
    for name, lo, hi in _layer_slices(b.critic):
        err = float((seen[155138][lo:hi] - gc[lo:hi]).abs().max()) / max(float(gc[lo:hi].abs().max()), 1e-12)
        worst = max(worst, err)
        assert err < 1e-3, (name, err)
    b.critic_optimizer.zero_grad(); loss.backward(); b.critic_optimizer.step()
    la = -b.critic.Q1(s, b.actor(s)).mean()
    ga = torch.cat([x.reshape(-1) for x in torch.autograd.grad(la, list(b.actor.parameters()))])
    assert abs(float(la_g) - float(la)) < 1e-3 * max(1.0, abs(float(la)))
    for name, lo, hi in _layer_slices(b.actor):
        err = float((seen[77330][lo:hi] - ga[lo:hi]).abs().max()) / max(float(ga[lo:hi].abs().max()), 1e-12)
        worst = max(worst, err)
        assert err < 1e-3, (name, err)
    assert _abi_lib().plen_td3_tc_timed_out() == 0
    print("tf32 learner, B = %d: worst per-tensor gradient error %.2e" % (B, worst))


def test_td3_tensor_core_update_trains_like_fp32():
    """20 whole updates (graph path, minibatches drawn from the replay ring with the same seeds) in tf32 and fp32: the
    parameters stay close (Adam normalises the step, so a 1e-3 gradient error moves a parameter by << lr per step)."""
    dev = torch.device("cuda:0")
    from plen_ml_walk_b200.td3 import ReplayBuffer
    a, b = _twin_agents(dev, seed=21, precision="tf32", twin_precision="fp32")
    g = torch.Generator(device=dev); g.manual_seed(23)
    n = 20000
    rbs = [ReplayBuffer(n, device=dev, seed=5) for _ in range(2)]
    st = torch.randn(n, 26, device=dev, generator=g); ac = torch.rand(n, 18, device=dev, generator=g) * 2 - 1
    s2 = torch.randn(n, 26, device=dev, generator=g); r = torch.randn(n, device=dev, generator=g)
    dn = torch.rand(n, device=dev, generator=g) < 0.05
    for rb in rbs:
        rb.add(st, ac, s2, r, dn)
    for _ in range(20):
        a.train(rbs[0], 2048)
        b.train(rbs[1], 2048)
    for k in ("actor", "critic", "actor_target", "critic_target"):
        d = float((a._flat[k] - b._flat[k]).abs().max())
        assert d < 20 * 3e-4 * 0.5, (k, d)      # far below 20 full Adam steps of lr = 3e-4
        assert torch.isfinite(a._flat[k]).all()
    assert a.kernel_launches() > 0 and _abi_lib().plen_td3_tc_timed_out() == 0


def test_td3_update_branches_equal_the_serial_order(monkeypatch):
    """plen_td3_train enqueues an update as three branches (current-Q forward and every dW next to the target / dX chain, the
    actor forward and the critic's Polyak update next to the critic step; captured as parallel graph branches).  With
    PLEN_TD3_SERIAL=1 the same launches go down ONE stream: below 512 rows (no split-K atomics) 40 updates must agree bit for
    bit, i.e. no branch runs ahead of a producer it depends on."""
    from plen_ml_walk_b200.td3 import ReplayBuffer, TD3Agent
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(41)
    n = 20000
    st = torch.randn(n, 26, device=dev, generator=g); ac = torch.rand(n, 18, device=dev, generator=g) * 2 - 1
    s2 = torch.randn(n, 26, device=dev, generator=g); r = torch.randn(n, device=dev, generator=g)
    dn = torch.rand(n, device=dev, generator=g) < 0.05
    for B in (100, 384):
        res = []
        for serial in ("1", "0"):
            monkeypatch.setenv("PLEN_TD3_SERIAL", serial)
            torch.manual_seed(43)
            agent = TD3Agent(device=dev, seed=7)
            rb = ReplayBuffer(n, device=dev, seed=5)
            rb.add(st, ac, s2, r, dn)
            losses = [agent.train(rb, B, return_losses=True) for _ in range(40)]
            res.append((agent, [x[1] for x in losses]))
        (a, la), (b, lb) = res
        for k in ("actor", "critic", "actor_target", "critic_target"):
            assert torch.equal(a._flat[k], b._flat[k]), (B, k)
        assert all(torch.equal(x, y) for x, y in zip(la, lb))
        for k in ("actor_m", "actor_v", "critic_m", "critic_v"):
            assert torch.equal(a._adam[k], b._adam[k]), (B, k)


def test_td3_update_follows_reference_rule():
    """TD3Agent.train (td3.py:259-356) sampling from the device replay ring inside the library: critic step every call,
    actor + Polyak every policy_freq calls; the critic loss on a fixed synthetic buffer goes down; checkpoints round-trip
    through the reference's four-file format including the Adam state."""
    from plen_ml_walk_b200.td3 import ReplayBuffer, TD3Agent
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    agent = TD3Agent(device=dev)
    rb = ReplayBuffer(max_size=5000, device=dev)
    s = torch.randn(5000, 26, device=dev)
    a = torch.rand(5000, 18, device=dev) * 2 - 1
    rb.add(s, a, s + 0.01, -(a ** 2).sum(1), torch.zeros(5000, dtype=torch.bool, device=dev))
    actor0 = [p.clone() for p in agent.actor.parameters()]
    tgt0 = [p.clone() for p in agent.critic_target.parameters()]
    al, cl0 = agent.train(rb, 100, return_losses=True)
    assert al is None and all(torch.equal(p, q) for p, q in zip(agent.actor.parameters(), actor0))      # delayed policy update
    assert all(torch.equal(p, q) for p, q in zip(agent.critic_target.parameters(), tgt0))
    al, _ = agent.train(rb, 100, return_losses=True)
    assert al is not None and not all(torch.equal(p, q) for p, q in zip(agent.actor.parameters(), actor0))
    d = [float((p - q).abs().max()) for p, q in zip(agent.critic_target.parameters(), tgt0)]
    assert 0 < max(d) < 0.01                                                                              # tau = 0.005
    losses = [float(agent.train(rb, 100, return_losses=True)[1]) for _ in range(300)]
    assert np.mean(losses[-50:]) < 0.5 * float(cl0)
    act = agent.select_action(s[:1000])
    with torch.no_grad():
        assert float((act - agent.actor(s[:1000])).abs().max()) < 1e-5
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        agent.save(td + "/ck")
        other = TD3Agent(device=dev)
        other.load(td + "/ck")
        assert torch.equal(other._flat["actor"], agent._flat["actor"]) and torch.equal(other._flat["critic"], agent._flat["critic"])
        assert other.critic_steps == agent.critic_steps == 302 and other.actor_steps == agent.actor_steps == 151
        assert torch.equal(other._adam["critic_v"], agent._adam["critic_v"])


def test_actor_forward_tensor_core_path_vs_fp32_kernel():
    """plen_actor_forward_tc (tcgen05.mma, FP16 operands, FP32 accumulation in TMEM) against the fp32 CUDA-core kernel on
    the reference's shipped checkpoint: FP16 operand rounding (2^-12 relative per operand, three layers) moves an action
    by 1e-3 on average at most; single actions that sit on a steep part of a tanh of this trained, large-weight network
    may move more (bounded at 0.15 here).  Ragged N (not a multiple of the 128-row tile), more tiles than SMs, and the
    exploration-noise stream is the same as the fp32 kernel's."""
    from plen_ml_walk_b200 import _abi
    from plen_ml_walk_b200.td3 import Actor, actor_forward
    dev = torch.device("cuda:0")
    g = np.load(GOLD)
    a = Actor().to(dev)
    a.load_state_dict({k: torch.from_numpy(g["actor_" + k.replace(".", "_")]).to(dev) for k in a.state_dict().keys()})
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    for n in (1, 127, 1000, 148 * 128 * 2 + 77):
        obs = torch.randn(n, 26, device=dev, generator=gen) * 0.5
        ref = actor_forward(a, obs)
        got = actor_forward(a, obs, precision="fp16")
        torch.cuda.synchronize()
        assert _abi.load_library().plen_actor_tc_timed_out() == 0
        d = (got - ref).abs()
        print("n %d: max %.3e mean %.3e" % (n, float(d.max()), float(d.mean())))
        assert float(d.max()) < 0.15 and float(d.mean()) < 1e-3, (n, float(d.max()), float(d.mean()))
    obs = torch.from_numpy(g["obs"]).to(dev)
    got = actor_forward(a, obs, precision="fp16")
    d = (got - torch.from_numpy(g["actor_out"]).to(dev)).abs()                                  # reference td3.py outputs
    assert float(d.max()) < 0.15 and float(d.mean()) < 1e-3, (float(d.max()), float(d.mean()))
    torch.manual_seed(0)
    fresh = Actor().to(dev)                                                                     # default-initialised network
    obs = torch.randn(4096, 26, device=dev, generator=gen)
    assert float((actor_forward(fresh, obs, precision="fp16") - actor_forward(fresh, obs)).abs().max()) < 1e-3
    obs = torch.randn(4096, 26, device=dev, generator=gen) * 0.1
    c32, n32 = actor_forward(a, obs), actor_forward(a, obs, noise_std=0.1, seed=11)
    c16, n16 = actor_forward(a, obs, precision="fp16"), actor_forward(a, obs, noise_std=0.1, seed=11, precision="fp16")
    inside = (n32.abs() < 0.999) & (n16.abs() < 0.999)
    assert float(((n16 - c16) - (n32 - c32))[inside].abs().max()) < 1e-5
