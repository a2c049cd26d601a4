"""The compiled CPU baseline ("port-compiled"): the oracle unrolls its own expanded tree of every pattern into straight-line
C++ (exa_oracle.cpp `ora_emit_source`), g++ compiles it with the interpreter's flags, and the result must be BIT-IDENTICAL to
the interpreter's hess_coord! -- it is the same restatement of src/hessian.jl:337-517,681-717, only specialised on the tree
the way Julia's compiler specialises the reference.  Nothing of the product's generator is involved."""
import os
import re

import numpy as np
import pytest

from examodels_jl_b200 import models as M
from oracle.oracle_api import Oracle
from util import inputs

MODELS = {
    "lv": lambda: M.luksan_vlcek(500),
    "lv_param": lambda: M.luksan_vlcek_param(40),
    "lv_aug": lambda: M.luksan_vlcek_aug(21, 3),
    "opf": lambda: M.ac_power(M.synthetic_power_data(30, 41, 6, seed=5)),
    "rocket": lambda: M.goddard_rocket(50),
    "family": lambda: M.pattern_family(200, 8),
    "all_ops_2": lambda: M.all_ops(32, 2),
}


@pytest.mark.parametrize("name", sorted(MODELS))
def test_compiled_port_equals_the_interpreter_bitwise(name):
    core = MODELS[name]()
    ora = Oracle.from_core(core)
    x, y = inputs(core)
    comp = ora.compile()
    for sigma in (1.0, 0.5):
        assert np.array_equal(comp.hess_coord(x, y, sigma), ora.hess_coord(x, y, sigma), equal_nan=True)
    assert np.array_equal(comp.hess_coord(x, None, 1.0), ora.hess_coord(x, None, 1.0), equal_nan=True)   # objective-only form
    ora.set_threads(4)                                                                                 # threaded partition of the points
    assert np.array_equal(comp.hess_coord(x, y, 1.0), ora.hess_coord(x, y, 1.0), equal_nan=True)
    ora.set_shard(1, 3)                                                                                # shard mode: the same slices are written
    a, b = np.full(ora.nnzh, 7.0), np.full(ora.nnzh, 7.0)
    comp.hess_coord(x, y, 1.0, a)
    ref = ora.hess_coord(x, y, 1.0)
    lib_written = a != 7.0
    assert np.array_equal(a[lib_written], ref[lib_written]) and not np.any(ref[~lib_written])


def test_emitted_source_is_self_contained_and_independent_of_the_product_generator():
    ora = Oracle.from_core(M.luksan_vlcek(10))
    src = ora.compile().source
    assert '#include "ora_tables.hpp"' in src and "exb_" not in src and "examodels" not in src
    # LV constraint: 2 sin + 1 exp table calls (op codes 16 and 8), unrolled -- no loop over tree nodes, no switch on a runtime tag
    con = src[src.index("hess_p0("):src.index("hess_p1(")]
    assert len(re.findall(r"uni\(16,", con)) == 2 and len(re.findall(r"uni\(8,", con)) == 1 and "switch" not in con
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "..", "oracle", "exa_oracle.cpp")) as f:
        assert "exb_plan" not in f.read().replace("examodels.jl_b200/csrc/exb_plan.hpp) is used", "")
