from .base_task import Task  # noqa: F401
from .location_finding import HiddenLocation  # noqa: F401
from .ces import CESTask  # noqa: F401
from .psychometric import PsychometricTask  # noqa: F401
from .gaussian_process import GPTask  # noqa: F401
