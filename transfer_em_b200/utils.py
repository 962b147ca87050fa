"""Tiled inference + export mirroring transfer_em/utils.py."""
import json
import os

import numpy as np
import torch

from ._lib import NET_G
from .cgan import EM2EM
from .engine import Engine


def z_slab_range(nz, rank, world):
    """Contiguous split of nz z tile-layers over `world` ranks: sizes differ by at most one."""
    per, rem = divmod(nz, world)
    zb = rank * per + min(rank, rem)
    return zb, zb + per + (1 if rank < rem else 0)


def _engine_of(model):
    if isinstance(model, Engine):
        return model, NET_G
    if hasattr(model, "engine") and hasattr(model, "net"):
        return model.engine, model.net          # a generator NetModel
    if hasattr(model, "engine"):
        return model.engine, NET_G              # an EM2EM: predict() uses generator_g (cgan.py:289-293)
    raise TypeError("model must be an EM2EM, a generator model or an Engine")


def predict_ng_cube(location, start, size, model, meanstd_x, meanstd_y, cloudrun=None, fetch_input=False,
                    outdimsize=None, buffer=None, *, rank=0, world=1, as_numpy=None):
    """transfer_em/utils.py:41-130 with `location` an in-memory uint8 [z,y,x] volume (numpy or cuda tensor)
    instead of a bucket path (`cloudrun` is accepted and ignored).  start / size are (x,y,z) tuples; the
    result is uint8 [size_z,size_y,size_x] (and the truncated input if fetch_input).

    Tiling, un-standardisation, crop, round-half-even and the uint8 wrap follow the reference exactly;
    the per-tile loop runs on the GPU in batches of tiles read straight from the resident volume.
    world > 1 shards the z tile-layers across ranks (no communication); each rank returns the full-size
    buffer with only its slab filled."""
    if isinstance(location, str):
        raise NotImplementedError("neuroglancer/DVID fetch (datasets/generators.py) is out of scope: pass a uint8 array")
    eng, net = _engine_of(model)
    od = outdimsize if outdimsize is not None else getattr(model, "outdimsize", eng.outdimsize)
    buf = buffer if buffer is not None else getattr(model, "buffer", eng.buffer)
    zr = None
    if world > 1:
        od2 = od - (od % 6) if (od // 6) != 0 else od
        nz = (int(size[2]) + od2 - 1) // od2
        zr = z_slab_range(nz, rank, world)
    res = eng.predict_volume(location, start, size, meanstd_x, meanstd_y, net=net, outdimsize=od, buffer=buf,
                             fetch_input=fetch_input, tile_z_range=zr)
    np_out = isinstance(location, np.ndarray) if as_numpy is None else as_numpy
    if np_out:
        torch.cuda.current_stream().synchronize()
        return tuple(r.cpu().numpy() for r in res) if fetch_input else res.cpu().numpy()
    return res


def save_model(name, ckpt_dir, meanstd_x, meanstd_y, size=74, is3d=True, wf=8):
    """transfer_em/utils.py:133-167: export generator_g + meta.json {buffer,outdimsize,meanstd_x,meanstd_y}.
    (The reference's default size=132 no longer constructs there; 74 is the only valid size.)"""
    model = EM2EM(size, name, is3d=is3d, ckpt_restore=ckpt_dir, wf=wf, max_batch=1, train=False)
    os.makedirs(name, exist_ok=True)
    np.savez(os.path.join(name, "generator_g.npz"), weights=model.engine.get_vector(NET_G), wf=np.int64(wf),
             is3d=np.int64(is3d), dimsize=np.int64(size))
    meta = {"buffer": model.buffer, "outdimsize": model.outdimsize,
            "meanstd_x": [float(meanstd_x[0]), float(meanstd_x[1])],
            "meanstd_y": [float(meanstd_y[0]), float(meanstd_y[1])]}
    with open(os.path.join(name, "meta.json"), "w") as f:
        f.write(json.dumps(meta))
    return name


def load_saved_model(model_dir, max_batch=32, device=None):
    z = np.load(os.path.join(model_dir, "generator_g.npz"))
    eng = Engine(dimsize=int(z["dimsize"]), is3d=bool(z["is3d"]), wf=int(z["wf"]), max_batch=max_batch, train=False, device=device)
    eng.set_vector(NET_G, z["weights"])
    return eng


def predict_cube_from_saved_model(location, start, size, cloudrun, model_dir, fetch_input=False):
    """transfer_em/utils.py:12-38."""
    meta = json.load(open(os.path.join(model_dir, 'meta.json')))
    eng = load_saved_model(model_dir)
    return predict_ng_cube(location, start, size, eng, meta["meanstd_x"], meta["meanstd_y"], cloudrun,
                           outdimsize=meta["outdimsize"], buffer=meta["buffer"], fetch_input=fetch_input)


def chunk_volume(volume_zyx, chunk=64, device=None):
    """The blocks of model_cloudrun/transferem.py:171-184, re-tiled on device: returns a list of
    ((x0, y0, z0), bytes) in the reference's iteration order (z outer, y, x inner); every block is the C-order bytes of
    volume[z0:z0+chunk, y0:y0+chunk, x0:x0+chunk] (clipped at the volume edge), i.e. `block.tobytes()` of the reference."""
    import ctypes as C
    import torch
    from . import _lib
    from .engine import _as_device, _stream
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    t, _ = _as_device(volume_zyx, dev, (torch.uint8,))
    t = t.contiguous()
    Z, Y, X = [int(v) for v in t.shape]
    out = torch.empty(Z * Y * X, dtype=torch.uint8, device=t.device)
    dims = (C.c_int64 * 3)(Z, Y, X)
    _lib.check(_lib.load().tem_chunk_volume(C.c_void_p(t.data_ptr()), dims, chunk, C.c_void_p(out.data_ptr()), _stream()))
    host = out.cpu().numpy()
    blocks = []
    for z0 in range(0, Z, chunk):
        cz = min(chunk, Z - z0)
        for y0 in range(0, Y, chunk):
            cy = min(chunk, Y - y0)
            for x0 in range(0, X, chunk):
                cx = min(chunk, X - x0)
                off = z0 * Y * X + cz * (y0 * X + cy * x0)
                blocks.append(((x0, y0, z0), host[off:off + cz * cy * cx].tobytes()))
    return blocks


def write_ng_chunks(volume_zyx, dest_dir, offset_xyz=(0, 0, 0), chunk=64, compress=True, device=None):
    """model_cloudrun/transferem.py:158-184 with a local directory in place of the GCS bucket: one file per 64^3 block named
    "{x0}-{x0+64}_{y0}-{y0+64}_{z0}-{z0+64}" (offsets added, the +64 is NOT clipped, as in the reference), gzip-compressed
    raw bytes.  Returns the list of file names."""
    import gzip
    os.makedirs(dest_dir, exist_ok=True)
    names = []
    ox, oy, oz = offset_xyz
    for (x0, y0, z0), raw in chunk_volume(volume_zyx, chunk, device):
        name = f"{x0 + ox}-{x0 + chunk + ox}_{y0 + oy}-{y0 + chunk + oy}_{z0 + oz}-{z0 + chunk + oz}"
        with open(os.path.join(dest_dir, name), "wb") as f:
            f.write(gzip.compress(raw) if compress else raw)
        names.append(name)
    return names
