"""Generator builder mirroring transfer_em/models/generator.py:22-117."""
from ..engine import Engine
from .utils import NetModel, NET_G

# generator.py:18,20.  The built graph is shape-polymorphic: any n = 2 (mod 4) >= 74 gives n-34, which the
# CUDA path accepts as a documented superset; 74 remains the default.
VALID_DIMS = [74]
VALID_OUT = [40]


def unet_generator(dimsize, is3d=True, norm_type='instancenorm', wf=8, *, engine=None, net=NET_G, max_batch=1,
                   device=None, seed=0):
    """Returns (model, out_dim) like the reference.  Raises RuntimeError for sizes that do not allow
    valid convolutions (generator.py:37-38)."""
    if dimsize % 4 != 2 or dimsize < 74:
        raise RuntimeError(f"{dimsize} does not allow for valid convolutions")
    if engine is None:
        engine = Engine(dimsize=dimsize, is3d=is3d, wf=wf, max_batch=max_batch, train=False, device=device, seed=seed)
    return NetModel(engine, net, 'generator'), engine.outdimsize
