#!/usr/bin/env python
"""Static completeness check of the Julia binding stub (integration/ExaModelsB200.jl), which cannot be executed here (no Julia):
every concrete `struct ... <: AbstractNode` of the reference's src/graph.jl must have an `emit!` method in the stub, every C-ABI
symbol the stub `ccall`s must be declared in include/exa_b200.h, and the exb_options tuple must have the header's field count.
Run from the repo root; needs /root/reference (CPU container only).  tests/test_julia_stub.py runs the same checks against a
committed list of the node types (tests/golden/reference_node_types.json) so that the GPU box needs no reference tree."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/graph.jl"


def node_types(path=REF):
    src = open(path).read()
    return sorted(set(re.findall(r"^struct\s+(\w+)(?:\{[^}]*\})?\s*<:\s*AbstractNode", src, flags=re.M)))


def check(types):
    stub = open(os.path.join(ROOT, "integration", "ExaModelsB200.jl")).read()
    hdr = open(os.path.join(ROOT, "include", "exa_b200.h")).read()
    missing = []
    for t in types:
        # emit!(e, n::T), emit!(e, ::T), emit!(e, n::T{...}) or T inside a Union{...}
        pat = re.compile(r"emit!\(e,\s*(?:\w+)?::(?:Union\{[^}]*\b%s\b[^}]*\}|%s\b)" % (t, t))
        if not pat.search(stub):
            missing.append(t)
    called = set(re.findall(r"\(:(exb_\w+),\s*LIB\)", stub)) | set(re.findall(r":(exb_\w+)\)", stub))
    declared = set(re.findall(r"\b(exb_[a-z0-9_]+)\s*\(", hdr))
    undeclared = sorted(c for c in called if c not in declared)
    body = re.sub(r"/\*.*?\*/", "", re.search(r"typedef struct exb_options \{(.*?)\} exb_options;", hdr, flags=re.S).group(1), flags=re.S)
    nfields = len(re.findall(r";", body))
    opt = re.search(r"opt = Ref\(\((.*?)\)\)\s*#", stub, flags=re.S).group(1)
    ntuple = len([x for x in re.split(r",(?![^()]*\))", opt) if x.strip()])
    return missing, undeclared, (nfields, ntuple)


if __name__ == "__main__":
    types = node_types()
    if "--write" in sys.argv:
        with open(os.path.join(ROOT, "tests", "golden", "reference_node_types.json"), "w") as f:
            json.dump({"source": "/root/reference/src/graph.jl: struct ... <: AbstractNode", "types": types}, f, indent=1)
    missing, undeclared, (nf, nt) = check(types)
    print("node types:", types)
    print("without an emit! method:", missing or "none")
    print("ccall'ed but not declared in include/exa_b200.h:", undeclared or "none")
    print(f"exb_options fields: header {nf}, stub tuple {nt}")
    sys.exit(1 if (missing or undeclared or nf != nt) else 0)
