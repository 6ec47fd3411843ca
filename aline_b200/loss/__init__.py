from .eig import EIGBounds, EIGStepLoss, NMCLoss, PCELoss  # noqa: F401
