#!/usr/bin/env python
"""Convert a checkpoint of the REAL reference (tf.train.Checkpoint written by transfer_em/cgan.py:84-107) into the flat
`ckpt-N.npz` that transfer_em_b200.EM2EM.restore() reads.

Needs tensorflow 2.x + tensorflow_addons and the reference checkout, so it cannot run in the build image; the packing half
(`pack_checkpoint`, numpy only) is what tests/test_abi.py exercises.  Usage on a machine that has TensorFlow:

    python tools/convert_tf_checkpoint.py /path/to/transfer_em <exp_name> <tf_ckpt_prefix> out_dir [--2d] [--wf 8]

File format (transfer_em_b200/cgan.py:make_checkpoint): per network one flat float32 vector of its variables in layer order
with the Keras layouts kept (conv kernels [k,k,k,Cin,Cout], transposed conv [k,k,k,Cout,Cin], the last discriminator layer
kernel then bias), the Adam first / second moments in the same order, and the scalars step / wf / is3d / dimsize.
"""
import os
import sys

import numpy as np

NETS = ("generator_g", "generator_f", "discriminator_x", "discriminator_y")


def pack_checkpoint(weights, adam_m=None, adam_v=None, step=0, wf=8, is3d=True, dimsize=74):
    """weights / adam_m / adam_v: {net name: list of arrays in `model.trainable_variables` order}.  Missing optimizer
    state (a checkpoint saved before the first step) becomes zeros, as Keras creates the slots."""
    data = {"step": np.int64(step), "wf": np.int64(wf), "is3d": np.int64(bool(is3d)), "dimsize": np.int64(dimsize)}
    for name in NETS:
        if name not in weights:
            raise KeyError(f"weights of {name} missing")
        flat = np.concatenate([np.asarray(w, np.float32).reshape(-1) for w in weights[name]])
        data[name] = flat
        for slot, src in (("_optimizer_m", adam_m), ("_optimizer_v", adam_v)):
            if src is not None and name in src:
                s = np.concatenate([np.asarray(w, np.float32).reshape(-1) for w in src[name]])
                if s.shape != flat.shape:
                    raise ValueError(f"{name}{slot}: {s.shape} values for {flat.shape} parameters")
            else:
                s = np.zeros_like(flat)
            data[name + slot] = s
    return data


def main(argv):
    ref_root, exp_name, prefix, out_dir = argv[:4]
    is3d = "--2d" not in argv
    wf = int(argv[argv.index("--wf") + 1]) if "--wf" in argv else 8
    sys.path.insert(0, ref_root)
    import tensorflow as tf  # noqa: F401
    from transfer_em.cgan import EM2EM
    model = EM2EM(74, exp_name, is3d=is3d, wf=wf)
    for net, opt in ((model.generator_g, model.generator_g_optimizer), (model.generator_f, model.generator_f_optimizer),
                     (model.discriminator_x, model.discriminator_x_optimizer), (model.discriminator_y, model.discriminator_y_optimizer)):
        try:                                                # Keras creates the Adam slots lazily; without them the restore of m / v is deferred
            opt._create_all_weights(net.trainable_variables)   # private TF 2.2-2.10 API [unverified: no TF install in the build image]
        except Exception:
            pass
    model.ckpt.restore(prefix).expect_partial()          # cgan.py:84-97: generators, discriminators and the four optimizers
    nets = {"generator_g": (model.generator_g, model.generator_g_optimizer), "generator_f": (model.generator_f, model.generator_f_optimizer),
            "discriminator_x": (model.discriminator_x, model.discriminator_x_optimizer),
            "discriminator_y": (model.discriminator_y, model.discriminator_y_optimizer)}
    weights, m, v = {}, {}, {}
    step = 0
    for name, (net, opt) in nets.items():
        tv = net.trainable_variables
        weights[name] = [x.numpy() for x in tv]
        try:
            m[name] = [opt.get_slot(x, "m").numpy() for x in tv]
            v[name] = [opt.get_slot(x, "v").numpy() for x in tv]
            step = max(step, int(opt.iterations.numpy()))
        except Exception:                                   # no step taken yet: slots do not exist
            pass
    data = pack_checkpoint(weights, m or None, v or None, step, wf, is3d, 74)
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "ckpt-1.npz")
    np.savez(out, **data)
    print("wrote", out, {k: (tuple(a.shape) if hasattr(a, "shape") else a) for k, a in data.items()})


if __name__ == "__main__":
    if len(sys.argv) < 5:
        sys.exit(__doc__)
    main(sys.argv[1:])
