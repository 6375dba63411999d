"""B200-native parallel runtime pieces (arenas, symmetric memory, fused collectives)."""
from .arena import ParamArena  # noqa: F401
