"""Dev tool: Actor.forward for N observations -- fp32 CUDA-core kernel vs tcgen05 FP16 kernel vs torch (cuBLAS) eager.
    python scripts/actor_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plen_ml_walk_b200 import _abi
from plen_ml_walk_b200.td3 import Actor, actor_forward

dev = torch.device("cuda:0")
torch.manual_seed(0)
a = Actor().to(dev)
out = {}
for n in (16384, 131072, 1048576):
    obs = torch.randn(n, 26, device=dev)
    res = {}
    for name, fn in (("fp32_kernel", lambda: actor_forward(a, obs)), ("tcgen05_fp16", lambda: actor_forward(a, obs, precision="fp16")),
                     ("torch_eager_fp32", lambda: a(obs))):
        with torch.no_grad():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 20
        flop = 2.0 * n * (26 * 256 + 256 * 256 + 256 * 18)
        res[name] = {"us": us, "tflops": flop / us / 1e6}
    with torch.no_grad():
        res["max_abs_diff_fp16_vs_fp32"] = float((actor_forward(a, obs, precision="fp16") - actor_forward(a, obs)).abs().max())
    res["timed_out"] = _abi.load_library().plen_actor_tc_timed_out()
    out["n_%d" % n] = res
print(json.dumps(out))
