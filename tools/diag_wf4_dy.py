#!/usr/bin/env python
"""Isolates the wf=4 D_y gradient discrepancy of tests/test_gpu_model.py::test_train_step_gradients_wide_model:
per-layer activations of the D_y passes (GPU stored values vs the oracle at stored values), LeakyReLU' sign mismatches,
per-variable gradient errors.  Run on a GPU box: python tools/diag_wf4_dy.py > gpurun_out/diag_wf4_dy.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tem_oracle as O                                 # noqa: E402
from tests.test_gpu_model import _train_case, NETS, _tt           # noqa: E402
from tests.gpu_helpers import rel_l2                               # noqa: E402
from transfer_em_b200._lib import NET_DX, NET_DY                   # noqa: E402


def main(seed=27, wf=4, scale=1.5):
    model, P, rx, ry = _train_case(True, 1, False, seed, scale=scale, wf=wf)
    eng = model.engine
    losses = eng.train_grads(rx, ry)
    ov = {'fake_y': eng.train_output('fake_y'), 'fake_x': eng.train_output('fake_x')}
    refq = O.train_step_grads(P, rx, ry, wf, True, dtype=torch.float32, quant=O.bf16_round, qweights=True, override_fakes=ov, keep_outputs=True)
    print("losses gpu", losses); print("losses ref", refq.losses)
    for k, net in NETS.items():
        got = eng.get_weights(net, which=1)
        print(f"net {k}: whole-gradient rel-L2 {rel_l2(np.concatenate([g.ravel() for g in got]), np.concatenate([g.ravel() for g in refq.grads[k]])):.3e}")
        for (vname, _, _), a, b in zip(eng.variables(net), got, refq.grads[k]):
            print(f"   {vname:12s} rel-L2 {rel_l2(a, b):.3e}  |ref| {np.linalg.norm(b):.3e}")
    b = 17
    inputs = {'real_y': ry[:, b:-b, b:-b, b:-b, :], 'fake_y': ov['fake_y'], 'real_x': rx[:, b:-b, b:-b, b:-b, :], 'fake_x': ov['fake_x']}
    for name, x in inputs.items():
        net, key = (NET_DY, 'dy') if name.endswith('_y') else (NET_DX, 'dx')
        lg = eng.disc_forward(net, x.astype(np.float32))
        acts = {}
        with torch.no_grad():
            ref = O.discriminator_forward(_tt(P, key), torch.tensor(x.astype(np.float32)), wf, True, quant=O.bf16_round, qweights=True, acts=acts)
        print(f"{key}({name}): logit gpu {lg.ravel()} ref {ref.numpy().ravel()}")
        for li in range(8):
            a = eng.last_activation(net, li).reshape(acts[f'd{li}'].shape)
            r = acts[f'd{li}'].numpy()
            flips = int(((a > 0) != (r > 0)).sum())
            small = np.abs(r[(a > 0) != (r > 0)])
            print(f"   d{li}: shape {r.shape[1:]} rel-L2 {rel_l2(a, r):.3e} sign flips {flips}/{r.size}" +
                  (f" (|ref| of flipped: max {small.max():.2e}, typical |act| {np.abs(r).mean():.2e})" if flips else ""))


if __name__ == "__main__":
    main(*[type(d)(v) for d, v in zip((27, 4, 1.5), sys.argv[1:])])
