import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import tem_oracle as O
from transfer_em_b200 import EM2EM
from transfer_em_b200._lib import NET_G, NET_F, NET_DX, NET_DY
NETS = {'g': NET_G, 'f': NET_F, 'dx': NET_DX, 'dy': NET_DY}
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
is3d = False; B = 2
r = np.random.default_rng(41)
P = {}
for k in NETS:
    layers = O.generator_layers(8) if k in ('g','f') else O.discriminator_layers(8, is3d)
    P[k] = [p*scale for p in O.init_params(layers, is3d, r)]
model = EM2EM(74, "dbg", is3d=is3d, wf=8, max_batch=B, dropout=False, checkpoint_dir="/tmp/none_dbg")
for k, n in NETS.items(): model.engine.set_weights(n, P[k])
def mk(perturb=0.0):
    o = O.OracleEM2EM(74, is3d=False, wf=8)
    rr = np.random.default_rng(7)
    o.P = {k: [(p*(1+perturb*rr.standard_normal(p.shape))).astype(np.float32) for p in v] for k, v in P.items()}
    o.M = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
    o.V = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
    return o
o1, o2 = mk(0.0), mk(1e-3)
shape = (B,74,74,1)
r = np.random.default_rng(42)
# "natural-image-like" smooth data rather than white noise
def smooth(a):
    t = torch.tensor(a).permute(0,3,1,2)
    t = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(t,(2,2,2,2),mode='reflect'),5,1)
    t = (t - t.mean())/t.std()
    return t.permute(0,2,3,1).numpy().astype(np.float32)
data = [(smooth(r.standard_normal(shape).astype(np.float32)), r.standard_normal(shape).astype(np.float32)*0.8) for _ in range(8)]
G=[];R1=[];R2=[]
for s in range(nsteps):
    bx, by = data[s % 8]
    G.append(model.train_step(bx, by)); R1.append(o1.train_step(bx, by)); R2.append(o2.train_step(bx, by))
G=np.array(G,np.float64);R1=np.array(R1);R2=np.array(R2)
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for s in list(range(0,nsteps,max(nsteps//10,1)))+[nsteps-1]:
    print(s, 'gpu', G[s]); print(s, 'ref', R1[s]); print(s, 'prt', R2[s])
relG = np.abs(G-R1)/np.maximum(np.abs(R1),1e-3); relP = np.abs(R2-R1)/np.maximum(np.abs(R1),1e-3)
print('max rel dev gpu-vs-ref per loss', relG.max(axis=0)); print('max rel dev perturbed-ref-vs-ref', relP.max(axis=0))
