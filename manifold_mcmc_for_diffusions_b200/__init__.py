"""B200-native constrained-HMC hot path of ``sde.mici_extensions``."""

from ._lib import MmdError, lib  # noqa: F401
from .batched import BatchedChains  # noqa: F401
