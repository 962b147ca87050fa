#!/usr/bin/env python
"""Per-layer timing of the tiled-inference path (predict_volume) at several tile-batch sizes.
   python tools/infer_profile.py [S=288] [batches...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transfer_em_b200 import Engine
S = int(sys.argv[1]) if len(sys.argv) > 1 else 288
batches = [int(v) for v in sys.argv[2:]] or [64, 128, 256]
ms = (0.0, 0.5774)
dev = torch.device("cuda", 0)
vol = torch.randint(0, 256, (S + 38,) * 3, dtype=torch.uint8, device=dev)
for B in batches:
    eng = Engine(dimsize=74, max_batch=B, train=False, seed=1)
    out = torch.zeros((S, S, S), dtype=torch.uint8, device=dev)
    eng.predict_volume(vol, (19, 19, 19), (S, S, S), ms, ms, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.predict_volume(vol, (19, 19, 19), (S, S, S), ms, ms, out=out); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print(f"== tile batch {B}: {S}^3 in {t:.2f} ms = {S ** 3 / t / 1e3:.1f} Mvox/s")
    eng.profile(True)
    eng.predict_volume(vol, (19, 19, 19), (S, S, S), ms, ms, out=out)
    rep = eng.profile_report(); eng.profile(False)
    tot = sum(v["ms"] for v in rep.values())
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"   {k:10s} {v['ms']:8.3f} ms  {100 * v['ms'] / tot:5.1f} %  {v['ms'] / v['count'] * 1e3:8.1f} us/launch  {v['bytes'] * v['count'] / v['ms'] / 1e6:8.1f} GB/s  {v['kernel']}")
    eng.close()
