from .censored_sigmoid_normal import CensoredSigmoidNormal  # noqa: F401
