#!/usr/bin/env python
"""What would a CUDA graph of the train step buy?  Captures one step and replays it (tem_debug_graph_replay), next to the
normal stream-launched step.   python tools/graph_probe.py [batch=8]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from transfer_em_b200 import EM2EM, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ms = (0.0, 0.5774)
m = EM2EM(74, "probe", max_batch=B, seed=1, meanstd_x=ms, meanstd_y=ms, checkpoint_dir="/tmp/tem_probe_none")
r = np.random.default_rng(0)
x = torch.from_numpy(r.integers(0, 256, (B, 74, 74, 74, 1), dtype=np.uint8)).cuda()
y = torch.from_numpy(r.integers(0, 256, (B, 74, 74, 74, 1), dtype=np.uint8)).cuda()
eng = m.engine
for _ in range(5): eng.train_step_async(x, y, ms, ms)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): eng.train_step_async(x, y, ms, ms)
e1.record(); torch.cuda.synchronize()
print(f"stream-launched step: {e0.elapsed_time(e1) / 20:.3f} ms")
out = C.c_float()
lib = _lib.load()
_lib.check(lib.tem_debug_graph_replay(eng._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), _lib.TEM_U8, _lib.fptr2(ms), _lib.fptr2(ms), B, 20, C.byref(out)))
print(f"graph replay        : {out.value:.3f} ms   ({lib.tem_last_error().decode()})")
