import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import tem_oracle as O
from transfer_em_b200 import EM2EM
from transfer_em_b200._lib import NET_G, NET_F, NET_DX, NET_DY
from tests.gpu_helpers import rel_l2
NETS = {'g': NET_G, 'f': NET_F, 'dx': NET_DX, 'dy': NET_DY}
is3d = '--3d' in sys.argv
B = 1
r = np.random.default_rng(21)
P = {}
for k in NETS:
    layers = O.generator_layers(8) if k in ('g','f') else O.discriminator_layers(8, is3d)
    P[k] = [p*4.0 for p in O.init_params(layers, is3d, r)]
model = EM2EM(74, "dbg", is3d=is3d, wf=8, max_batch=B, dropout=False, checkpoint_dir="/tmp/none_dbg")
for k, n in NETS.items(): model.engine.set_weights(n, P[k])
shape = (B,) + (74,)*(3 if is3d else 2) + (1,)
rx = r.standard_normal(shape).astype(np.float32); ry = (r.standard_normal(shape)*0.8+0.1).astype(np.float32)
model.engine.train_grads(rx, ry)
# oracle: identity loss of pass same_y = G(real_y) only
T = [torch.tensor(p, requires_grad=True) for p in P['g']]
ga = {}
same_y = O.generator_forward(T, torch.tensor(ry), 8, is3d, quant=O.bf16_round, qweights=True, graph_acts=ga)
loss = O.identity_loss(O.crop_cl(torch.tensor(ry), 17), same_y)
loss.backward()
layers = O.generator_layers(8)
for i in range(11):
    y = ga[f'g{i}']
    dA = O._to_cl(y.grad).numpy(); a = O._to_cl(y.detach()).numpy()
    ref = dA * np.where(a > 0, 1.0, layers[i].slope)
    got = model.engine.debug_backward_scratch(True, i).reshape(ref.shape)
    err = np.abs(got-ref)
    # where is the error concentrated?
    e2 = (err**2).sum(axis=(0, -1))
    tot = e2.sum()
    sl = (slice(2, -2),)*(3 if is3d else 2)
    print(f'g{i}', 'rel %.3e' % rel_l2(got, ref), 'norms got %.3e ref %.3e' % (np.linalg.norm(got), np.linalg.norm(ref)), 'interior share of err^2: %.3f' % (e2[sl].sum()/max(tot,1e-300)), 'shape', ref.shape)
