"""synference_b200: B200-native implementation of synference's mock-library hot path."""
from .units import *  # noqa: F401,F403
