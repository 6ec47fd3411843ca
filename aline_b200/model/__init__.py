from .base import Aline  # noqa: F401
from .embedder import Embedder  # noqa: F401
from .encoder import Encoder  # noqa: F401
from .head import AcquisitionHead, GMMTargetHead, OutputHead  # noqa: F401
