"""Generates tests/golden/ipopt_logs.json from the Ipopt logs the reference's own documentation build printed
(/root/reference/docs/src/parameters.md: the parametric Luksan-Vlcek N=10 model solved for three parameter settings).
These are REFERENCE-PRODUCED numbers: iteration tables (objective, inf_pr, inf_du, ||d||, step lengths), final scaled and
unscaled objective, dual infeasibility and constraint violation.  Run from the repo root in the build container:
    python tests/golden/make_ipopt_log_golden.py"""
import json
import os
import re

SRC = "/root/reference/docs/src/parameters.md"
ROW = re.compile(r"^\s*(\d+)\s+([-\d.e+]+)\s+([-\d.e+]+)\s+([-\d.e+]+)\s+(-?[\d.]+)\s+([-\d.e+]+)\s+(\S+)\s+([-\d.e+]+)\s+([-\d.e+]+)\S*\s+(\d+)\s*$")


def main():
    text = open(SRC).read()
    runs = text.split("This is Ipopt version")[1:]
    thetas = [[100.0, 1.0], [200.0, 1.0], [200.0, 0.5]]      # parameters.md: add_par, then two set_parameter! calls
    assert len(runs) == len(thetas)
    out = {"_source": "docs/src/parameters.md of the reference (Ipopt 3.14.19 logs printed by its documentation build)", "runs": []}
    for theta, run in zip(thetas, runs):
        rows = []
        for line in run.splitlines():
            m = ROW.match(line)
            if m:
                rows.append({"iter": int(m.group(1)), "objective": m.group(2), "inf_pr": m.group(3), "inf_du": m.group(4),
                             "d_norm": m.group(6), "lg_rg": m.group(7), "alpha_du": m.group(8), "alpha_pr": m.group(9), "ls": int(m.group(10))})
        fin = {}
        for key, pat in (("objective", r"Objective\.+:\s+(\S+)\s+(\S+)"), ("dual_infeasibility", r"Dual infeasibility\.+:\s+(\S+)\s+(\S+)"),
                         ("constraint_violation", r"Constraint violation\.+:\s+(\S+)\s+(\S+)")):
            m = re.search(pat, run)
            fin[key] = {"scaled": m.group(1), "unscaled": m.group(2)}
        nnz = [int(v) for v in re.findall(r"Number of nonzeros in (?:equality constraint Jacobian|Lagrangian Hessian)\.+:\s+(\d+)", run)]
        out["runs"].append({"theta": theta, "nnzj": nnz[0], "nnzh": nnz[1], "iterations": rows, "final": fin})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ipopt_logs.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, [len(r["iterations"]) for r in out["runs"]])


if __name__ == "__main__":
    main()
