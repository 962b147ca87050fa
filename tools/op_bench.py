"""Per-op micro-benchmark at the bench shapes (batch 8, wf=8), through the per-op C ABI.
   python tools/op_bench.py [names...]     names: g1.fwd g1.dgrad g1.wgrad g0.fwd g0.wgrad g2.fwd ..."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.gpu_helpers import make_desc, conv_forward, conv_dgrad, conv_wgrad
B = 8
# name: (k, s, cin, cout, transposed, in_dim)
L = {'g0': (3,1,1,8,False,74), 'g1': (3,1,8,8,False,72), 'g2': (4,2,8,8,False,70), 'g3': (3,1,8,16,False,34), 'g4': (4,2,16,16,False,32),
     'g5': (3,1,16,32,False,15), 'g6': (4,2,32,16,True,13), 'g7': (3,1,32,32,False,26), 'g8': (3,1,32,16,False,24), 'g9': (4,2,16,8,True,22),
     'g10': (3,1,16,16,False,44), 'g11': (3,1,16,1,False,42), 'w1': (3,1,64,64,False,72), 'w2': (4,2,64,64,False,70), 'w4': (4,2,128,128,False,32), 'w6': (4,2,256,128,True,13), 'w9': (4,2,128,64,True,22), 'w3': (3,1,64,128,False,34), 'w5': (3,1,128,256,False,15), 'w7': (3,1,256,256,False,26), 'w8': (3,1,256,128,False,24), 'w10': (3,1,128,128,False,44), 'd2': (3,1,8,16,False,18), 'd3': (3,1,16,32,False,16), 'd5': (3,1,32,32,False,6), 'd0': (3,1,1,8,False,40), 'd1': (4,2,8,8,False,38), 'd4': (4,2,32,32,False,14), 'd6': (4,2,32,32,False,4)}
if len(sys.argv) > 1 and sys.argv[1].startswith('B='): B = int(sys.argv.pop(1)[2:])
names = sys.argv[1:] or ['g1.fwd', 'g1.dgrad', 'g1.wgrad', 'g0.fwd', 'g0.wgrad', 'g10.wgrad', 'g7.wgrad', 'g2.fwd', 'g2.dgrad', 'g2.wgrad', 'g11.fwd', 'g11.wgrad', 'd6.fwd', 'd4.fwd']
dev = 'cuda'
import ctypes as C
from transfer_em_b200 import _lib
from tests.gpu_helpers import ptr, stream, DT
lib = _lib.load()
def timeit(fn, n=20):
    """back-to-back asynchronous launches between two events: host overhead is hidden behind the GPU queue"""
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for nm in names:
    lay, op = nm.split('.')
    k, s, ci, co, tr, n = L[lay]
    dims = (n, n, n)
    in_dt = torch.uint8 if ci == 1 else torch.bfloat16
    out_dt = torch.float32 if co == 1 else torch.bfloat16
    d = make_desc(B, dims, ci, co, k, s, tr, 0.3, 0, in_dt, out_dt, (0.0, 0.58), tc=1)
    x = torch.randint(0, 255, (B,) + dims + (ci,), device=dev).to(in_dt) if ci == 1 else torch.randn((B,) + dims + (ci,), device=dev).to(torch.bfloat16)
    wshape = (k, k, k) + ((co, ci) if tr else (ci, co))
    w = (torch.randn(wshape, device=dev) * 0.1).float()
    y = conv_forward(x, w, d)
    od = y.shape[1]
    dy = torch.randn(y.shape, device=dev).to(y.dtype)
    act = torch.randn(x.shape, device=dev).to(torch.bfloat16)
    odims = (C.c_int32 * 3)(*y.shape[1:4])
    st = stream()
    if op == 'fwd':
        out = torch.empty_like(y)
        us = timeit(lambda: _lib.check(lib.tem_conv_forward(C.byref(d), ptr(x), ptr(w), None, ptr(out), odims, st)))
    elif op == 'dgrad':
        dx_dt = torch.bfloat16 if ci > 1 else torch.float32
        dx = torch.empty(x.shape, dtype=dx_dt, device=dev)
        a_ = act if ci > 1 else None
        us = timeit(lambda: _lib.check(lib.tem_conv_dgrad(C.byref(d), ptr(dy), DT[dy.dtype], ptr(w), ptr(a_), 0.3, ptr(dx), DT[dx_dt], st)))
    else:
        dw = torch.zeros(wshape, dtype=torch.float32, device=dev)
        us = timeit(lambda: _lib.check(lib.tem_conv_wgrad(C.byref(d), ptr(x), ptr(dy), DT[dy.dtype], ptr(dw), st)))
    vox_in, vox_out = B * n ** 3, B * od ** 3
    byt = vox_in * ci * (1 if ci == 1 else 2) + vox_out * co * (4 if co == 1 else 2)
    macs = (vox_in if tr else vox_out) * ci * co * k ** 3
    print(f"{nm:10s} {us:9.1f} us   {byt / us / 1e3:8.1f} GB/s   {2 * macs / us / 1e6:8.2f} TFLOP/s")
