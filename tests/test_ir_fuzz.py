"""The C ABI must never crash on a malformed IR: random word mutations and truncations of valid IRs either
build a plan or return EXB_ERR_IR (run in a subprocess so that a crash is a test failure, not a dead pytest)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mutated_ir_never_crashes():
    code = textwrap.dedent(f"""
        import sys, struct, ctypes as C
        sys.path.insert(0, {ROOT!r})
        import numpy as np
        import examodels_jl_b200 as E
        from examodels_jl_b200 import models as M, backend as B
        lib = B.lib()
        rng = np.random.default_rng(0)
        ok = bad = 0
        for core in (M.luksan_vlcek(10), M.luksan_vlcek_aug(6, 2), M.ac_power(M.synthetic_power_data(5, 6, 2, seed=1))):
            ir, _ = core.to_ir()
            w = np.frombuffer(ir, dtype=np.int64).copy()
            for trial in range(1500):
                v = w.copy()
                for _ in range(int(rng.integers(1, 4))):
                    k = int(rng.integers(0, v.size))
                    v[k] = rng.choice([0, -1, 1, 2, 7, 9, 50, 2**31, -2**40, int(rng.integers(-5, 300))])
                if trial % 7 == 0:
                    v = v[: int(rng.integers(1, v.size))]
                buf = v.tobytes()
                h = C.c_void_p()
                rc = lib.exb_plan_create(buf, C.c_size_t(len(buf)), None, C.byref(h))
                if rc == 0:
                    ok += 1
                    lib.exb_plan_destroy(h)
                else:
                    assert rc == 3, rc
                    bad += 1
        print("done", ok, bad)
    """)
    env = dict(os.environ, EXB_CACHE_DIR="/tmp/exb_fuzz_cache")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "done" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
