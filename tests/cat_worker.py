"""Worker of tests/test_gpu_model.py::test_concat_layer_merged_launches_match_split_launches: one 3-D train-step gradient
evaluation (wf = 8, batch 1, fixed dropout keys) whose generator / discriminator gradient vectors are written to an .npz.
The debug knobs that select the kernel paths (TEM_NO_DGRAD_CAT, TEM_NO_WGRAD_CAT, TEM_CONV_DOWN_V1, ...) are read once per
process, hence the subprocess."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tem_oracle as O                        # noqa: E402  (test infrastructure: parameter recipe only)
from transfer_em_b200 import Engine                       # noqa: E402
from transfer_em_b200._lib import NET_G, NET_F, NET_DX, NET_DY   # noqa: E402

NETS = {'g': NET_G, 'f': NET_F, 'dx': NET_DX, 'dy': NET_DY}


def main():
    out = sys.argv[1]
    r = np.random.default_rng(77)
    P = {}
    for k in NETS:
        layers = O.generator_layers(8) if k in ('g', 'f') else O.discriminator_layers(8, True)
        P[k] = [p * 2.0 for p in O.init_params(layers, True, r)]
    # narrow uint8 range: standardised inputs within +-0.5, away from the 1/(t + 1e-7) singularity of the focal identity / cycle
    # loss at |a - b| -> 2 (tests/test_gpu_model.py:_train_case)
    rx = r.integers(90, 166, (1, 74, 74, 74, 1), dtype=np.uint8)
    ry = r.integers(100, 176, (1, 74, 74, 74, 1), dtype=np.uint8)
    eng = Engine(dimsize=74, is3d=True, wf=8, max_batch=1, train=True)
    for k, net in NETS.items():
        eng.set_weights(net, P[k])
    eng.set_dropout_keys([2 * i + 101 for i in range(12)])
    losses = eng.train_grads(rx, ry, meanstd_x=(0.0, 0.6), meanstd_y=(0.08, 0.6))
    np.savez(out, losses=np.asarray(losses, np.float64), **{k: eng.get_vector(net, which=1) for k, net in NETS.items()})
    eng.close()


if __name__ == "__main__":
    main()
