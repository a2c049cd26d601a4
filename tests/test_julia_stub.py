"""The Julia binding stub (integration/ExaModelsB200.jl) cannot run here (no Julia toolchain): what CAN be checked statically is
that it covers every concrete node type of the reference's src/graph.jl, only calls entry points that include/exa_b200.h
declares, and builds the exb_options tuple with the header's field count (scripts/check_julia_stub.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_stub_covers_every_reference_node_type_and_only_declared_symbols():
    import check_julia_stub as C
    types = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_node_types.json")))["types"]
    assert {"SumNode", "ProdNode", "Node1", "Node2", "Var", "DataIndexed", "ArgLeaf"} <= set(types)
    if os.path.exists(C.REF):   # CPU container: the committed list is what the reference tree says
        assert C.node_types() == types
    missing, undeclared, (nf, nt) = C.check(types)
    assert not missing and not undeclared and nf == nt == 6


def test_stub_handles_the_cases_the_round_1_review_found_missing():
    stub = open(os.path.join(ROOT, "integration", "ExaModelsB200.jl")).read()
    assert "emit_fold!(e, n.inners, :+, 0.0)" in stub and "emit_fold!(e, n.inners, :*, 1.0)" in stub   # left folds (graph.jl:520-567)
    assert "Pair{<:Integer}" in stub                                                                 # Int augmentation index (nlp.jl:1994-1997)
    assert "sort!(pats" not in stub and "take_obj" in stub                                           # add order: merge of the core's own lists
    assert "struct CompressedB200Model" in stub and ":exb_hess_compressed" in stub
    assert ":exb_comm_init" in stub and ":exb_eval" in stub
