from .generator import unet_generator, VALID_DIMS, VALID_OUT      # noqa: F401
from .discriminator import discriminator                           # noqa: F401
from .utils import downsample, upsample                            # noqa: F401
