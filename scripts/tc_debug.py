"""Dev tool: per-tensor gradient error of the tf32 learner against fp32 autograd.  python scripts/tc_debug.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from plen_ml_walk_b200.td3 import TD3Agent

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
torch.manual_seed(11)
a, b = TD3Agent(device=dev, precision=os.environ.get("PREC", "tf32")), TD3Agent(device=dev)
for k in ("actor", "actor_target", "critic", "critic_target"):
    b._flat[k].copy_(a._flat[k])
g = torch.Generator(device=dev); g.manual_seed(17)
s = torch.randn(B, 26, device=dev, generator=g); ac = torch.rand(B, 18, device=dev, generator=g) * 2 - 1
s2 = torch.randn(B, 26, device=dev, generator=g); r = torch.randn(B, 1, device=dev, generator=g)
nd = (torch.rand(B, 1, device=dev, generator=g) > 0.1).float(); nz = torch.randn(B, 18, device=dev, generator=g)
seen = {}
a.total_it = 1
la_g, lc_g = a.train(None, batch=(s, ac, s2, r, nd), noise=nz, return_losses=True,
                     grad_hook=lambda gf: seen.setdefault(gf.numel(), gf.clone()))
with torch.no_grad():
    n2 = (nz * b.policy_noise).clamp(-b.noise_clip, b.noise_clip)
    a2 = (b.actor_target(s2) + n2).clamp(-1, 1)
    q1t, q2t = b.critic_target(s2, a2)
    y = r + nd * b.discount * torch.min(q1t, q2t)
q1, q2 = b.critic(s, ac)
loss = F.mse_loss(q1, y) + F.mse_loss(q2, y)
gc = torch.cat([x.reshape(-1) for x in torch.autograd.grad(loss, list(b.critic.parameters()), retain_graph=True)])
print("critic loss", float(lc_g), float(loss))
off = 0
for name, p in b.critic.named_parameters():
    n = p.numel()
    d = (seen[155138][off:off + n] - gc[off:off + n]).abs()
    print("critic %-12s max|ref| %.3e  max err %.3e  rel %.2e  argmax %d" % (name, float(gc[off:off + n].abs().max()), float(d.max()),
          float(d.max()) / max(float(gc[off:off + n].abs().max()), 1e-12), int(d.argmax())))
    off += n
b.critic_optimizer.zero_grad(); loss.backward(); b.critic_optimizer.step()
la = -b.critic.Q1(s, b.actor(s)).mean()
ga = torch.cat([x.reshape(-1) for x in torch.autograd.grad(la, list(b.actor.parameters()))])
print("actor loss", float(la_g), float(la))
off = 0
for name, p in b.actor.named_parameters():
    n = p.numel()
    d = (seen[77330][off:off + n] - ga[off:off + n]).abs()
    print("actor  %-12s max|ref| %.3e  max err %.3e  rel %.2e  argmax %d" % (name, float(ga[off:off + n].abs().max()), float(d.max()),
          float(d.max()) / max(float(ga[off:off + n].abs().max()), 1e-12), int(d.argmax())))
    off += n
