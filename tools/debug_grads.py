import sys, os, numpy as np, torch
sys.path.insert(0, '.')
from oracle import tem_oracle as O
from transfer_em_b200 import EM2EM
from transfer_em_b200._lib import NET_G, NET_F, NET_DX, NET_DY
from tests.gpu_helpers import rel_l2
NETS = {'g': NET_G, 'f': NET_F, 'dx': NET_DX, 'dy': NET_DY}
is3d = '--3d' in sys.argv
B = 1 if is3d else 2
scale = float(os.environ.get("SCALE", "2.0"))
r = np.random.default_rng(21)
P = {}
for k in NETS:
    layers = O.generator_layers(8) if k in ('g','f') else O.discriminator_layers(8, is3d)
    P[k] = [p*scale for p in O.init_params(layers, is3d, r)]
model = EM2EM(74, "dbg", is3d=is3d, wf=8, max_batch=B, dropout=False, checkpoint_dir="/tmp/none_dbg")
for k, n in NETS.items(): model.engine.set_weights(n, P[k])
shape = (B,) + (74,)*(3 if is3d else 2) + (1,)
rx = r.standard_normal(shape).astype(np.float32); ry = (r.standard_normal(shape)*0.8+0.1).astype(np.float32)
losses = model.engine.train_grads(rx, ry)
ov = {'fake_y': model.engine.train_output('fake_y'), 'fake_x': model.engine.train_output('fake_x')} if '--ov' in sys.argv else None
ref = O.train_step_grads(P, rx, ry, 8, is3d, dtype=torch.float32, keep_outputs=True, quant=O.bf16_round if '--q' in sys.argv else None, qweights='--q' in sys.argv, override_fakes=ov)
print('losses gpu', [float(v) for v in losses]); print('losses ref', ref.losses)
for name in ("fake_y","cycled_x","fake_x","cycled_y","same_x","same_y"):
    print(name, rel_l2(model.engine.train_output(name), ref.outputs[name]))
for k, net in NETS.items():
    got = model.engine.get_weights(net, which=1)
    for (vn,_,_), a, b in zip(model.engine.variables(net), got, ref.grads[k]):
        print(k, vn, 'rel', '%.3e' % rel_l2(a,b), 'norm ref %.3e got %.3e' % (np.linalg.norm(b), np.linalg.norm(a)))
