#!/usr/bin/env python
"""Generate tests/golden/{ref,tf}_*.npz by running the REFERENCE'S OWN Python (/root/reference/transfer_em).

Two modes, one schema:
  python tools/make_reference_golden.py              build container: the reference runs on oracle/tf_shim -> tests/golden/ref_*.npz
  python tools/make_reference_golden.py --real-tf    anywhere tensorflow 2.x + tensorflow_addons are installed (TEM_REFERENCE points to
                                                     the reference checkout): the reference runs on REAL TensorFlow, eagerly
                                                     -> tests/golden/tf_*.npz, which pin parity against TensorFlow's own kernels
tests/test_reference_golden.py (oracle, CPU) and tests/test_gpu_golden.py (CUDA path) check every prefix that is present; the
tf_ files cannot be produced in the build image (no TensorFlow wheel, no network) and are skipped while absent.

TensorFlow / tensorflow_addons are not installable here, so the reference is imported on top of oracle/tf_shim (a
torch-CPU stand-in for the few dozen TF symbols its hot path touches; see oracle/tf_shim/tensorflow/__init__.py for what
that pins and what it leaves unpinned).  Everything below calls reference code only: EM2EM.__init__ / train_step /
predict (transfer_em/cgan.py), the model builders, predict_ng_cube (transfer_em/utils.py), scale_tensor /
standardize_population / unstandardize_population / get_meanstd (datasets/datasets.py), warp_tensor / accuracy (debug.py).
The two network-bound helpers of predict_ng_cube (volume3d_ng: HTTP fetch; create_dataset_from_generator: tf.data) are
replaced by an in-memory source that feeds the reference's own scale_tensor + standardize_population.

Run:  python tools/make_reference_golden.py            (needs /root/reference; the GPU box never runs this)
The committed fixtures are read by tests/test_reference_golden.py (oracle, CPU) and tests/test_gpu_golden.py (CUDA path).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TEM_REFERENCE", "/root/reference")
REAL_TF = "--real-tf" in sys.argv
if not REAL_TF:
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import torch                                        # noqa: E402
import tensorflow as tf                             # noqa: E402  (the shim)
from transfer_em.cgan import EM2EM                  # noqa: E402  (reference code)
from transfer_em import utils as ref_utils          # noqa: E402
from transfer_em.datasets import datasets as ref_ds  # noqa: E402
from transfer_em import debug as ref_debug          # noqa: E402
from oracle import tem_oracle as O                  # noqa: E402  (only for init_params / the dropout hash)

PREFIX = "tf" if REAL_TF else "ref"
if REAL_TF:
    tf.config.run_functions_eagerly(True)           # train_step is a @tf.function: the gradient recorder needs eager tensors
    tf.random.set_seed(0)
    _t = tf.constant
else:
    _t = tf._t


def to_np(t):
    if hasattr(t, "detach"):
        return t.detach().numpy()
    return t.numpy() if hasattr(t, "numpy") else np.asarray(t)

GOLD = os.path.join(ROOT, "tests", "golden")
NETS = (("g", "generator_g"), ("f", "generator_f"), ("dx", "discriminator_x"), ("dy", "discriminator_y"))
PROBE_SEED = 977


def params_for(wf, is3d, seed, scale):
    """The parity tests' parameter recipe (tests/test_gpu_model.py:_params)."""
    r = np.random.default_rng(seed)
    P = {}
    for k, _ in NETS:
        layers = O.generator_layers(wf) if k in ("g", "f") else O.discriminator_layers(wf, is3d)
        P[k] = [p * scale for p in O.init_params(layers, is3d, r)]
    P["dx"][9] = np.array([0.1], np.float32); P["dy"][9] = np.array([-0.2], np.float32)
    return P


def load_into_reference(model, P, is3d):
    for k, attr in NETS:
        ws = P[k]
        if k in ("dx", "dy") and not is3d:
            ws = ws[2:]            # 2-D: block "1" is not part of the Keras graph (discriminator.py:49-51)
        getattr(model, attr).set_weights(ws)


def inputs_for(is3d, B, seed):
    r = np.random.default_rng(seed + 1)
    shape = (B,) + (74,) * (3 if is3d else 2) + (1,)
    rx = np.clip(r.standard_normal(shape) * 0.4, -0.9, 0.9).astype(np.float32)
    ry = np.clip(r.standard_normal(shape) * 0.35 + 0.1, -0.9, 0.9).astype(np.float32)
    return rx, ry


def probes(arrs, tag):
    """Exact scalar invariants of a list of arrays: L2 norm and the dot product with a seeded Gaussian vector (fp64)."""
    r = np.random.default_rng(PROBE_SEED)
    out = []
    for a in arrs:
        a = np.asarray(a, np.float64).reshape(-1)
        out.append([np.linalg.norm(a), float(a @ r.standard_normal(a.size))])
    return {tag: np.array(out, np.float64)}


class GradRecorder:
    """The shim's Adam sees exactly what the reference hands to apply_gradients (cgan.py:218-228)."""

    def __init__(self, opt):
        self.opt, self.last = opt, None
        self._orig = opt.apply_gradients
        opt.apply_gradients = self.apply

    def apply(self, gv):
        gv = list(gv)
        self.last = [None if g is None else to_np(g).copy() for g, _ in gv]
        self._orig(gv)


def train_case(name, is3d, B, seed, scale, wf=8, keys=None, steps=3, store_full_grads=True):
    if REAL_TF and keys is not None:
        return                          # mask injection needs the shim's Dropout layer
    if not REAL_TF:
        tf.DROPOUT.update(enabled=keys is not None, keys=keys, calls=0,
                          mask_fn=lambda key, shape: O.dropout_keep_mask(key, shape))
    model = EM2EM(74, f"golden_{name}", is3d=is3d, wf=wf)
    P = params_for(wf, is3d, seed, scale)
    load_into_reference(model, P, is3d)
    rec = {k: GradRecorder(getattr(model, attr + "_optimizer")) for k, attr in NETS}
    rx, ry = inputs_for(is3d, B, seed)
    out = {"is3d": np.int64(is3d), "B": np.int64(B), "seed": np.int64(seed), "scale": np.float64(scale), "wf": np.int64(wf),
           "buffer": np.int64(model.buffer), "outdimsize": np.int64(model.outdimsize),
           "weights_checksum": np.array([[float(np.sum(np.asarray(w, np.float64))), float(np.sum(np.asarray(w, np.float64) ** 2))]
                                         for k, _ in NETS for w in P[k]])}
    if keys is not None:
        out["dropout_keys"] = np.array(keys, np.uint32)
    losses = []
    for step in range(steps):
        if keys is not None:
            tf.DROPOUT["calls"] = 0              # the same 12 masks every step (the CUDA test re-injects the same keys)
        before = {k: getattr(model, attr).get_weights() for k, attr in NETS}
        l = model.train_step(_t(rx), _t(ry))
        losses.append([float(v) for v in l])
        after = {k: getattr(model, attr).get_weights() for k, attr in NETS}
        for k, _ in NETS:
            out.update(probes([a - b for a, b in zip(after[k], before[k])], f"delta_{k}_step{step + 1}"))
        if step == 0:
            for k, _ in NETS:
                g = [np.zeros_like(w) if gi is None else gi for gi, w in zip(rec[k].last, before[k])]
                out.update(probes(g, f"grad_probe_{k}"))
                if store_full_grads:
                    for i, gi in enumerate(g):
                        out[f"grad_{k}_{i}"] = gi.astype(np.float32)
                else:                                 # fp16 per-variable max-normalised copy (rel. precision 5e-4)
                    for i, gi in enumerate(g):
                        m = float(np.abs(gi).max()) or 1.0
                        out[f"grad16_{k}_{i}"] = (gi / m).astype(np.float16); out[f"gradmax_{k}_{i}"] = np.float64(m)
    out["losses"] = np.array(losses, np.float64)
    if not REAL_TF:
        tf.DROPOUT.update(enabled=False, keys=None)
    y = model.predict(_t(rx))                                          # cgan.py:289-293 after `steps` updates
    out["predict_after"] = np.asarray(to_np(y), np.float32)[:1, ::3, ::3]
    np.savez_compressed(os.path.join(GOLD, f"{PREFIX}_train_{name}.npz"), **out)
    print(name, "losses step 1:", losses[0])


class _MemDataset:
    """In-memory replacement for volume3d_ng + create_dataset_from_generator (utils.py:88-89): yields, per ROI, the
    reference's scale_tensor + standardize_population of the uint8 cube, batch 1."""

    def __init__(self, vol, size, rois, meanstd):
        self.vol, self.size, self.rois, self.meanstd = vol, size, rois, meanstd

    def __iter__(self):
        s = self.size
        for (x, y, z) in self.rois:
            cube = self.vol[z:z + s, y:y + s, x:x + s]
            t = ref_ds.scale_tensor(_t(cube.copy()))
            t = ref_ds.standardize_population(t, self.meanstd)
            yield tf.expand_dims(t, 0)


def run_reference_predict_ng_cube(vol, start, size, model, ms_x, ms_y, fetch_input):
    state = {}

    def fake_volume3d_ng(location, bbox, size=132, seed=None, array=None, cloudrun=None, **kw):
        state["size"], state["rois"] = size, array
        return "source"

    def fake_create(dataset, custom_map, batch_size=1, epoch_size=None, meanstd=None, **kw):
        return _MemDataset(vol, state["size"], state["rois"], meanstd), None
    ref_utils.volume3d_ng, ref_utils.create_dataset_from_generator = fake_volume3d_ng, fake_create
    return ref_utils.predict_ng_cube("mem", start, size, model, ms_x, ms_y, fetch_input=fetch_input)


class _ExactModel:
    """A 'generator' both sides can evaluate bit-exactly: y = 0.5 * centre_crop(x) + 0.1 in fp32."""
    outdimsize, buffer = 40, 17

    def predict(self, x):
        a = np.asarray(to_np(x), np.float32)
        return _t(np.float32(0.5) * a[:, 17:-17, 17:-17, 17:-17, :] + np.float32(0.1))


def tiling_cases():
    r = np.random.default_rng(34)
    vol = r.integers(0, 256, (72 + 38, 72 + 38, 72 + 38), dtype=np.uint8)      # every 74^3 tile of the ragged request is in bounds
    start, size = (19, 19, 19), (72, 50, 36 + 3)
    ms_x, ms_y = (0.0, 0.5774), (0.03, 0.4)
    inb, out = run_reference_predict_ng_cube(vol, start, size, _ExactModel(), ms_x, ms_y, True)
    res = {"vol_seed": np.int64(34), "start": np.array(start), "size": np.array(size), "ms_x": np.array(ms_x), "ms_y": np.array(ms_y),
           "exact_in": inb, "exact_out": out}
    # the same request through the reference with a real generator (weights of the parity recipe, last layer rescaled)
    model = EM2EM(74, "golden_tile", is3d=True, wf=8)
    P = params_for(8, True, 33, 5.0)
    load_into_reference(model, P, True)
    probe = model.generator_g(_t(np.random.default_rng(1).standard_normal((1, 74, 74, 74, 1)).astype(np.float32)))
    g = model.generator_g.get_weights()
    k11 = np.float32(0.8 / float(np.std(to_np(probe))))
    g[11] = (g[11] * k11).astype(np.float32)
    model.generator_g.set_weights(g)
    res["g11_scale"] = np.float64(k11)
    with torch.no_grad():
        res["gen_out"] = run_reference_predict_ng_cube(vol, start, size, model, ms_x, ms_y, False)
    np.savez_compressed(os.path.join(GOLD, f"{PREFIX}_predict_ng_cube.npz"), **res)
    print("predict_ng_cube goldens:", out.shape, res["gen_out"].shape)


def conversion_cases():
    u = np.arange(256, dtype=np.uint8)
    out = {}
    for i, ms in enumerate([(0.0, 1.0), (0.0, 0.5774), (0.02, 0.55), (-0.113, 0.731)]):
        t = ref_ds.standardize_population(ref_ds.scale_tensor(_t(u.copy())), ms)
        out[f"std_{i}"] = to_np(t)[:, 0]; out[f"ms_{i}"] = np.array(ms)
    # the uint8 conversion of utils.py:109,118 on a dense sweep that includes .5 ties and out-of-range values (wraps)
    y = np.concatenate([np.linspace(-3.5, 3.5, 4001), (np.arange(-40, 300) + 0.5) / 127.5 / 0.4 - (1 + 0.03) / 0.4]).astype(np.float32)
    ms_y = (0.03, 0.4)
    v = to_np((ref_ds.unstandardize_population(_t(y.copy()), ms_y) + 1) * 127.5)
    out["y_sweep"] = y; out["ms_y"] = np.array(ms_y)
    out["y_u8"] = np.around(v).astype(np.uint8)
    out["y_trunc"] = v.astype(np.uint8)                      # the fetch_input path truncates (utils.py:123-125)
    # get_meanstd (datasets.py:173-190)
    r = np.random.default_rng(5)
    tensors = [r.standard_normal((9, 11, 13, 1)).astype(np.float32) * (1 + 0.1 * i) + 0.01 * i for i in range(5)]
    m, s = ref_ds.get_meanstd([_t(t) for t in tensors])
    out["meanstd"] = np.array([float(m), float(s)])
    # warp_tensor (debug.py:7-63) with the hole seeds fixed, and accuracy (debug.py:65-71)
    for nd, shape in ((3, (12, 14, 16, 1)), (2, (20, 24, 1))):
        t = r.standard_normal(shape).astype(np.float32)
        uni = r.uniform(0, 1, int(np.prod(shape))).astype(np.float32)
        uni[:: 97] = 1e-5                                   # make sure some holes are seeded at this small size
        orig = tf.random.uniform
        tf.random.uniform = lambda shp, lo=0.0, hi=1.0: _t(uni.copy())
        w = ref_debug.warp_tensor(_t(t.copy()))
        tf.random.uniform = orig
        out[f"warp_in_{nd}"] = t; out[f"warp_uniform_{nd}"] = uni; out[f"warp_out_{nd}"] = to_np(w)
    out["accuracy"] = np.float64(ref_debug.accuracy(_t(out["warp_in_3"]), _t(out["warp_out_3"])))
    np.savez_compressed(os.path.join(GOLD, f"{PREFIX}_conversions.npz"), **out)
    print("conversion goldens written")


def structure_case():
    """Facts the reference's builders decide: variable shapes / order, parameter counts, geometry, error behaviour."""
    out = {}
    for is3d in (True, False):
        m = EM2EM(74, "golden_struct", is3d=is3d, wf=8)
        tag = "3d" if is3d else "2d"
        for k, attr in NETS[::2]:
            shapes = [tuple(w.shape) for w in getattr(m, attr).get_weights()]
            out[f"shapes_{k}_{tag}"] = np.array([list(s) + [0] * (6 - len(s)) for s in shapes])
            out[f"count_{k}_{tag}"] = np.int64(sum(int(np.prod(s)) for s in shapes))
        out[f"buffer_{tag}"] = np.int64(m.buffer); out[f"outdimsize_{tag}"] = np.int64(m.outdimsize)
        n = 74
        x = _t(np.zeros((1,) + (n,) * (3 if is3d else 2) + (1,), np.float32))
        out[f"gen_out_shape_{tag}"] = np.array(m.generator_g(x).shape)
        out[f"disc_out_shape_{tag}"] = np.array(m.discriminator_x(_t(np.zeros((1,) + (40,) * (3 if is3d else 2) + (1,), np.float32))).shape)
    errs = []
    for d in (70, 76, 132):
        try:
            EM2EM(d, "golden_err")
            errs.append(0)
        except RuntimeError:
            errs.append(1)
    out["raises_runtime_error_70_76_132"] = np.array(errs)
    np.savez_compressed(os.path.join(GOLD, f"{PREFIX}_structure.npz"), **out)
    print("structure goldens written", {k: v for k, v in out.items() if k.startswith("count")})


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, (os.cpu_count() or 2)))
    structure_case()
    conversion_cases()
    tiling_cases()
    train_case("2d", False, 2, 21, 2.0)
    train_case("3d", True, 1, 21, 2.0, store_full_grads=False)
    keys = [0x1001 + 7919 * i for i in range(12)]
    train_case("3d_dropout", True, 1, 23, 2.0, keys=keys, steps=1, store_full_grads=False)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
