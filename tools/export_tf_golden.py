#!/usr/bin/env python
"""Dump golden vectors from the REAL reference (needs tensorflow 2.x + tensorflow_addons and the reference checkout).

Cannot run in the build image (no TF).  A third party runs it once:
    python tools/export_tf_golden.py /path/to/transfer_em tests/golden/tf_golden_3d.npz
and tests/test_tf_golden.py (skipped when the file is absent) then compares the oracle AND the CUDA path with it,
which is what would pin parity against TensorFlow.
"""
import sys

import numpy as np


def main(ref_root, out_path, is3d=True, seed=0):
    sys.path.insert(0, ref_root)
    import tensorflow as tf
    from transfer_em.cgan import EM2EM
    tf.random.set_seed(seed)
    model = EM2EM(74, "golden", is3d=is3d, wf=8)
    rng = np.random.default_rng(seed)
    shape = (1,) + (74,) * (3 if is3d else 2) + (1,)
    rx = np.clip(rng.standard_normal(shape) * 0.4, -0.9, 0.9).astype(np.float32)
    ry = np.clip(rng.standard_normal(shape) * 0.35 + 0.1, -0.9, 0.9).astype(np.float32)
    out = {"real_x": rx, "real_y": ry}
    for name, net in (("g", model.generator_g), ("f", model.generator_f), ("dx", model.discriminator_x), ("dy", model.discriminator_y)):
        for i, w in enumerate(net.get_weights()):
            out[f"w_{name}_{i}"] = w
    out["fake_y_inference"] = model.generator_g(rx, training=False).numpy()
    out["logit_dy"] = model.discriminator_y(out["fake_y_inference"], training=False).numpy()
    out["gen_loss"] = np.float32(model.generator_loss(out["logit_dy"]))
    out["identity_loss"] = np.float32(model.identity_loss(ry[:, 17:-17, 17:-17, 17:-17] if is3d else ry[:, 17:-17, 17:-17], out["fake_y_inference"]))
    np.savez_compressed(out_path, **out)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
