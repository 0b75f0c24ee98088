O=gpurun_out/${1:-r2_tc6}; mkdir -p $O
for m in 0 1 2; do
  echo "== PLEN_TC_ONEPASS=$m" | tee -a $O/onepass.txt
  PLEN_TC_ONEPASS=$m python scripts/tc_debug.py 4096 2>&1 | grep "critic fc[1245]\|actor  fc[12]" | awk '{print $1,$2,$9,$10}' | tr '\n' ';' | tee -a $O/onepass.txt; echo | tee -a $O/onepass.txt
  PLEN_TC_ONEPASS=$m python - <<'PY' 2>&1 | tail -1 | tee -a $O/onepass.txt
import torch, sys, os
sys.path.insert(0, os.getcwd())
from plen_ml_walk_b200.td3 import ReplayBuffer, TD3Agent
dev = torch.device("cuda:0"); torch.manual_seed(0)
rb = ReplayBuffer(max_size=200000, device=dev)
s = torch.randn(200000, 26, device=dev); a = torch.rand(200000, 18, device=dev) * 2 - 1
rb.add(s, a, s + 0.01, -(a ** 2).sum(1), torch.zeros(200000, dtype=torch.bool, device=dev))
out = []
for B in (1024, 4096, 16384):
    tc = TD3Agent(device=dev, max_batch=16384, precision="tf32")
    for _ in range(20): tc.train(rb, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300): tc.train(rb, B)
    e1.record(); torch.cuda.synchronize()
    out.append("B=%d %.1f us" % (B, 1e3 * e0.elapsed_time(e1) / 300))
print("; ".join(out))
PY
done
