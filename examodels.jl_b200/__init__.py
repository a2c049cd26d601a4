"""B200-native evaluator for the ExaModels.jl per-pattern NLP callback path.

The package directory is named after the reference (`examodels.jl_b200`); because of the
dot it is imported through the shim module `examodels_jl_b200` at the repo root.
"""
from .graph import *  # noqa: F401,F403
from .graph import (Constant, Null, Var, ParameterNode, DataSource, Node1, Node2, exa_sum,
                    exa_prod, atan2, hypot, max_, min_, abs_, pow_runtime)
from .nlp import ExaCore, product, Iterator
from .backend import ExaModel, Plan, ExbError, build_library  # noqa: E402
from . import models  # noqa: E402
