#!/usr/bin/env python
"""Write profiles/traffic.json from `ncu --set full` captures: per kernel, dram__bytes_read.sum + dram__bytes_write.sum of ONE
launch, tied to the hash of the generated module the capture was taken on (bench.py ignores -- loudly -- a figure whose module
is not the one it runs).

    python scripts/traffic_from_ncu.py <module.cubin> <kernel>=<file.ncu-rep> ...
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    module = os.path.basename(sys.argv[1])
    out = {}
    for spec in sys.argv[2:]:
        kernel, rep = spec.split("=")
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units, r = rows[0], rows[1], rows[2]
        assert kernel in r[hdr.index("Kernel Name")], (kernel, r[hdr.index("Kernel Name")])
        tot, parts = 0.0, []
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            b = float(r[i].replace(",", "")) * UNIT[units[i]]
            tot += b; parts.append(f"{name} {b / 1e6:.3f} MB")
        out[kernel] = {"bytes": int(round(tot)), "module": module,
                       "capture": f"{os.path.basename(rep)} (ncu --set full --clock-control none, LV N=1e7: {' + '.join(parts)} per launch)"}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
