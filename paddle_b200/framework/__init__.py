from . import dtype, flags, place, random, unique_name  # noqa: F401
from .flags import get_flags, set_flags  # noqa: F401
from .random import seed, get_rng_state, set_rng_state, get_cuda_rng_state, set_cuda_rng_state  # noqa: F401
