"""Dev tool for ncu: E robots, W warm-up env steps with random actions (robots fall, lie, reset: the steady state of the bench
workload), then S more.  PLEN_AB_NOLINKS=1 switches the link-box contacts off.   python scripts/profile_steady.py E W S
The last S steps sit between cudaProfilerStart / cudaProfilerStop (ncu --profile-from-start off captures exactly them)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plen_ml_walk_b200 import _abi
from plen_ml_walk_b200.urdf_loader import packaged_model

E, W, S = (int(sys.argv[k]) if len(sys.argv) > k else d for k, d in ((1, 32768), (2, 60), (3, 2)))
lib = _abi.load_library()
cfg = _abi.PlenConfigC()
lib.plen_default_config(C.byref(cfg), 0)
if os.environ.get("PLEN_AB_NOLINKS"):
    cfg.link_contacts = 0
if os.environ.get("PLEN_AB_MANIFOLD"):
    cfg.sole_manifold = 1      # persistent sole manifolds (plen_config.sole_manifold)
model = _abi.model_to_c(packaged_model())
ctx = lib.plen_create(C.byref(cfg), C.byref(model), E, 0)
assert ctx
dev = torch.device("cuda:0")
obs = torch.empty((E, 26), device=dev); rew = torch.empty(E, device=dev)
done = torch.empty(E, dtype=torch.uint8, device=dev); tmo = torch.empty(E, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev); g.manual_seed(0)
acts = [torch.empty((E, 18), device=dev).uniform_(-1, 1, generator=g) for _ in range(8)]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
lib.plen_reset(ctx, None, P(obs), st)
for k in range(W + S):
    if k == W:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    assert lib.plen_step(ctx, P(acts[k % 8]), P(obs), P(rew), P(done), P(tmo), None, st) == 0
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(done.float().mean()))
