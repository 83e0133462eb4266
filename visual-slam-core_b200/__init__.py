"""B200-native local bundle adjustment behind visual-slam-core's optimizer API."""
from . import synth  # noqa: F401
from .optimizer import B200Optimizer, BaseOptimizer, lm_config_from  # noqa: F401
from .solver import LMConfig, LocalBASolver, SolveResult  # noqa: F401
from . import dist, handlers, motion, packing  # noqa: F401
from .motion import refine_motion_from_2d3d  # noqa: F401
