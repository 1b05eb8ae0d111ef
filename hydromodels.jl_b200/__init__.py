"""hydromodels.jl_b200 — B200-native forward-run engine behind the HydroModels.jl component API.

Scope: `model(input, params, config; initstates)` for bucket models under the
explicit `MutableSolver`, lowered to CUDA C, compiled with NVRTC for sm_100a and launched
through the C-ABI `libhydrob200.so` (`include/hydrob200.h`).  No CPU execution path.
"""
from .symbolic import (Const, Eq, Equation, Expr, Sym, Symbols, clamp, exp, hmax, hmin, log, parameters,
                       sqrt, step_func, tanh, variables)
from .components import (Chain, ConstantInterpolation, Dense, DiscreteSolver, GraphAggregation, GridAggregation,
                         HydroBucket, HydroConfig, HydroFlux, HydroModel, HydroRoute, ImmutableSolver,
                         LinearInterpolation, MutableSolver, NeuralBucket, NeuralFlux, ODESolver, SolverType, StateFlux,
                         UnitHydrograph, build_aggr_func, default_config, normalize_config)
from . import calibration, irf, models, rasters
from .lowering import LoweringError, lower_stepper
from .engine import B200Model, HydroB200Error, NoDeviceError, device_count

__version__ = "0.1.0"
