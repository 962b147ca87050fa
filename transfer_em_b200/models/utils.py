"""Model objects and block helpers mirroring transfer_em/models/utils.py.

The reference returns Keras models; here a model is a view (engine, net id) onto a tem_handle whose
arithmetic runs in libtem_b200.  ``downsample`` / ``upsample`` (utils.py:41,89) are kept as
descriptors of the two block types so that code written against the reference can introspect the
layer list; the blocks themselves only execute as part of a generator / discriminator.
"""
from dataclasses import dataclass
from typing import List

from .._lib import NET_G, NET_F, NET_DX, NET_DY  # noqa: F401


@dataclass
class BlockLayer:
    kind: str          # 'conv' | 'convT'
    kernel: int
    stride: int
    padding: str       # 'valid' | 'same'
    filters: int
    activation: str    # 'leaky_relu(0.3)'
    dropout: float = 0.0


@dataclass
class Block:
    name: str
    infilters: int
    layers: List[BlockLayer]


def downsample(id, infilters, outfilters, is3d, filter_size=4, norm_type='instancenorm', apply_norm=True):
    """transfer_em/models/utils.py:41-85: conv3 VALID + LReLU (skip output) -> conv(filter_size) stride 2 VALID + LReLU.
    Returns (down_block, skip_block) like the reference's pair of models.  norm_type / apply_norm are accepted
    and ignored, as in the reference (every normalisation call is commented out there)."""
    skip = Block(f"Downsample_{id}_skip", infilters, [BlockLayer('conv', 3, 1, 'valid', outfilters, 'leaky_relu(0.3)')])
    down = Block(f"Downsample_{id}", infilters, skip.layers + [BlockLayer('conv', filter_size, 2, 'valid', outfilters, 'leaky_relu(0.3)')])
    return down, skip


def upsample(id, infilters, outfilters, is3d, filter_size=4, norm_type='instancenorm', apply_dropout=True):
    """transfer_em/models/utils.py:89-137: conv3 VALID (2*outfilters) + LReLU -> convT(filter_size) stride 2 SAME
    -> Dropout(0.5) -> LReLU."""
    return Block(f"Upsample_{id}", infilters, [
        BlockLayer('conv', 3, 1, 'valid', outfilters * 2, 'leaky_relu(0.3)'),
        BlockLayer('convT', filter_size, 2, 'same', outfilters, 'leaky_relu(0.3)', 0.5 if apply_dropout else 0.0)])


class NetModel:
    """Callable stand-in for the Keras model of one network of an Engine."""

    def __init__(self, engine, net, kind):
        self.engine, self.net, self.kind = engine, net, kind

    @property
    def trainable_variables(self):
        return self.engine.get_weights(self.net)

    def get_weights(self):
        return self.engine.get_weights(self.net)

    def set_weights(self, weights):
        self.engine.set_weights(self.net, weights)

    def count_params(self):
        return self.engine.param_count(self.net)

    def variable_info(self):
        return self.engine.variables(self.net)

    def __call__(self, x, training=False, meanstd=None, dropout_key=None):
        if self.kind == 'generator':
            key = 0
            if training and self.engine_dropout:
                key = dropout_key if dropout_key is not None else self._next_key()
            return self.engine.gen_forward(self.net, x, meanstd=meanstd, dropout_key=key)
        return self.engine.disc_forward(self.net, x)

    predict = __call__
    engine_dropout = True
    _ctr = 0

    def _next_key(self):
        NetModel._ctr += 1
        return (0x9E3779B1 * NetModel._ctr) & 0xFFFFFFFF or 1
