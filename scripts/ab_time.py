"""Dev tool: time plen_step of an arbitrary build of libplen_b200.so (A/B comparisons on one GPU box).
    python scripts/ab_time.py <lib.so> [envs] [steps]
Uses only the entry points every build has (plen_default_config / plen_create / plen_reset / plen_step)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plen_ml_walk_b200 import _abi
from plen_ml_walk_b200.urdf_loader import packaged_model

lib = C.CDLL(os.path.abspath(sys.argv[1]))
E = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
K = int(sys.argv[3]) if len(sys.argv) > 3 else 30
vp = C.c_void_p
lib.plen_create.restype = vp
lib.plen_create.argtypes = [vp, vp, C.c_int, C.c_int]
lib.plen_step.argtypes = [vp] * 8
lib.plen_reset.argtypes = [vp] * 4
cfg = _abi.PlenConfigC()
lib.plen_default_config(C.byref(cfg), 0)
if os.environ.get("PLEN_AB_NOLINKS"):
    cfg.link_contacts = 0      # round-1 behaviour: only the soles touch the ground
if os.environ.get("PLEN_AB_MANIFOLD"):
    cfg.sole_manifold = 1      # the persistent sole manifold option (needs a build that has the field)
model = _abi.model_to_c(packaged_model())
ctx = lib.plen_create(C.byref(cfg), C.byref(model), E, 0)
assert ctx
dev = torch.device("cuda:0")
obs = torch.empty((E, 26), device=dev); rew = torch.empty(E, device=dev)
done = torch.empty(E, dtype=torch.uint8, device=dev); tmo = torch.empty(E, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev); g.manual_seed(0)
amp = float(os.environ.get("PLEN_AB_AMP", "1"))      # < 1: small random actions, the robots stay on their feet (8 sole points)
acts = [torch.empty((E, 18), device=dev).uniform_(-amp, amp, generator=g) for _ in range(8)]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
lib.plen_reset(ctx, None, P(obs), st)
def step(a):
    rc = lib.plen_step(ctx, P(a), P(obs), P(rew), P(done), P(tmo), None, st)
    assert rc == 0
for w in range(int(os.environ.get("PLEN_AB_WARM", "5"))):
    step(acts[w % 8])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(K):
    step(acts[k % 8])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("%s: %.3f ms/step  %.3f M env-steps/s  (reward mean %.4f, done frac %.4f)" % (sys.argv[1], ms, E / ms / 1e3, rew.mean().item(), done.float().mean().item()))
