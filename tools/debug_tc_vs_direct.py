"""Runs the same 3D train_grads with the tcgen05 conv path on/off in two subprocesses per pass limit and
compares the backward scratch (dP per layer) of the k-th generator backward pass."""
import sys, os, subprocess, numpy as np
sys.path.insert(0, '.')
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    from oracle import tem_oracle as O
    from transfer_em_b200 import EM2EM
    from transfer_em_b200._lib import NET_G, NET_F, NET_DX, NET_DY
    NETS = {'g': NET_G, 'f': NET_F, 'dx': NET_DX, 'dy': NET_DY}
    r = np.random.default_rng(21)
    P = {k: [p*2.0 for p in O.init_params(O.generator_layers(8) if k in 'gf' else O.discriminator_layers(8, True), True, r)] for k in NETS}
    model = EM2EM(74, "dbg", is3d=True, wf=8, max_batch=1, dropout=False, checkpoint_dir="/tmp/none_dbg")
    for k, n in NETS.items(): model.engine.set_weights(n, P[k])
    rx = r.standard_normal((1,74,74,74,1)).astype(np.float32); ry = (r.standard_normal((1,74,74,74,1))*0.8+0.1).astype(np.float32)
    model.engine.train_grads(rx, ry)
    out = {f'dP{i}': model.engine.debug_backward_scratch(True, i) for i in range(11)}
    for k, n in NETS.items(): out['grad_'+k] = model.engine.get_vector(n, 1)
    np.savez(sys.argv[2], **out)
    sys.exit(0)
def rel(a, b): return float(np.linalg.norm(a-b)/max(np.linalg.norm(b), 1e-30))
for lim in (5, 6):
    res = {}
    for tag, env in (('tc', {}), ('direct', {'TEM_NO_CONV_TC': '1'})):
        e = dict(os.environ); e.update(env); e['TEM_DEBUG_GEN_BWD'] = str(lim)
        fn = f'/tmp/dbg_{tag}_{lim}.npz'
        subprocess.check_call([sys.executable, __file__, 'child', fn], env=e, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        res[tag] = np.load(fn)
    print('pass limit', lim, ' '.join('dP%d:%.1e' % (i, rel(res['tc'][f'dP{i}'], res['direct'][f'dP{i}'])) for i in range(10, -1, -1)))
    print('   grads tc-vs-direct', {k: '%.2e' % rel(res['tc']['grad_'+k], res['direct']['grad_'+k]) for k in 'gf'})
