#!/usr/bin/env python
"""Static evidence per generated kernel and launch-shape variant (no GPU needed: nvcc cross-compiles): registers, spill / stack
bytes from `ptxas -v`, and the SASS mnemonics that prove the TMA bulk store (UBLKCP + UTMACMDFLUSH / fence.proxy.async) and
FP64 FMA / shared-memory staging.  Writes a markdown table (profiles/r02_sass_registers.md).

    python scripts/sass_report.py lv rocket opf family > profiles/r02_sass_registers.md
"""
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KERNELS = ("exb_hess_g0", "exb_hessc_g0", "exb_eval_g0", "exb_jac_g0", "exb_cons_g0", "exb_obj_g0", "exb_ggrad_g0", "exb_gradt_g0", "exb_sgrad_g0",
           "exb_jprod_g0", "exb_jtprod_g0", "exb_hprod_g0")


def main():
    import examodels_jl_b200 as E
    from examodels_jl_b200 import models as M
    which = sys.argv[1:] or ["lv", "rocket", "opf", "family"]
    models = {"lv": lambda: M.luksan_vlcek(10_000_000), "rocket": lambda: M.goddard_rocket(1_000_000),
              "opf": lambda: M.ac_power(M.synthetic_power_data()), "family": lambda: M.pattern_family(1000, 32)}
    print("# Generated kernels: registers, spills and SASS evidence (round 2)\n")
    print("`nvcc -cubin -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v`, three launch-shape variants per model "
          "(`__launch_bounds__(128, 16 | 12 | 1)`); the first-call tuner picks one per kernel.  Columns: registers / stack bytes / spill stores+loads "
          "(ptxas -v), static SASS instruction count, and whether the kernel contains the TMA bulk store (`UBLKCP` + `UTMACMDFLUSH`).\n")
    for key in which:
        with tempfile.TemporaryDirectory() as d:
            os.environ["EXB_CACHE_DIR"], os.environ["EXB_KEEP_CUBIN"] = d, "1"
            E.build_library()
            p = E.Plan(models[key]())
            p.compile()
            print(f"## {key}: module {os.path.basename(p.module_path())}\n")
            print("| kernel | minb | regs | stack B | spill st/ld B | SASS instr | DFMA+DMUL+DADD | UBLKCP | UTMACMDFLUSH | STS/LDS |")
            print("|---|---|---|---|---|---|---|---|---|---|")
            for cu in sorted(glob.glob(os.path.join(d, "*.cu"))):
                src = open(cu).read()
                minb = re.search(r"#define EXB_MINB (\d+)", src).group(1)
                cubin = cu[:-3] + ".cubin"
                log = open(cubin + ".log").read()
                for k in KERNELS:
                    m = re.search(r"Compiling entry function '%s' for 'sm_100a'.*?Function properties for %s\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers"
                                  % (k, k), log, flags=re.S)
                    if not m:
                        continue
                    sass = subprocess.run(["cuobjdump", "-sass", "-fun", k, cubin], capture_output=True, text=True).stdout
                    ins = [ln.split()[1] if not ln.split()[1].startswith("@") else ln.split()[2]
                           for ln in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln) and len(ln.split()) > 2]
                    cnt = lambda pre: sum(1 for i in ins if i.startswith(pre))  # noqa: E731
                    print(f"| {k} | {minb} | {m.group(4)} | {m.group(1)} | {m.group(2)}/{m.group(3)} | {len(ins)} | {cnt('DFMA') + cnt('DMUL') + cnt('DADD')} | "
                          f"{cnt('UBLKCP')} | {cnt('UTMACMDFLUSH')} | {cnt('STS')}/{cnt('LDS')} |")
            print()


if __name__ == "__main__":
    main()
