"""kinetic-gan_b200: B200-native (sm_100a) implementation of Kinetic-GAN's spatial-temporal graph-convolution
hot path behind the reference's own module surface.  See DESIGN.md.

The directory name carries a hyphen (fixed by the project layout); import it with
`importlib.import_module("kinetic-gan_b200")` or through the `kgan_b200` alias module at the repo root."""
from . import functional, geometry, ops  # noqa: F401
from .models.discriminator import Discriminator  # noqa: F401
from .models.generator import Generator, Mapping_Net, NoiseInjection  # noqa: F401
from .models.init_gan.tgcn import ConvTemporalGraphical  # noqa: F401
from .ops import get_precision, set_precision  # noqa: F401

__version__ = "0.1.0"
