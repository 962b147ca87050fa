"""Discriminator builder mirroring transfer_em/models/discriminator.py:14-105 (disc_prior=None path)."""
from ..engine import Engine
from .utils import NetModel, NET_DX


def discriminator(is3d=True, norm_type='instancenorm', wf=8, disc_prior=None, *, engine=None, net=NET_DX,
                  max_batch=1, device=None, seed=0):
    if disc_prior is not None:
        raise NotImplementedError("disc_prior (discriminator.py:62-66) needs a Keras h5 prior model: out of scope")
    if engine is None:
        engine = Engine(dimsize=74, is3d=is3d, wf=wf, max_batch=max_batch, train=False, device=device, seed=seed)
    return NetModel(engine, net, 'discriminator')
