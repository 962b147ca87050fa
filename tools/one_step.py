"""One warm + N profiled train steps at the bench configuration (for ncu)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from transfer_em_b200 import EM2EM
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = EM2EM(74, "ncu", wf=8, max_batch=B, seed=1, meanstd_x=(0, .58), meanstd_y=(0, .35), checkpoint_dir="/tmp/none_ncu")
x = torch.randint(0, 256, (B, 74, 74, 74, 1), dtype=torch.uint8, device="cuda")
y = torch.randint(0, 256, (B, 74, 74, 74, 1), dtype=torch.uint8, device="cuda")
for i in range(n):
    print(m.train_step(x, y))
