"""world_size-2 (and 3) gloo tests of the sharded path on CPU: the collective assembly of
examodels.jl_b200/parallel.py, driven by a stand-in rank-local evaluator built on the oracle's
shard mode, must reproduce the unsharded callbacks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleShard:
    """Rank-local evaluator with the ExaModel callback surface, on CPU tensors (tests only)."""

    def __init__(self, core, rank, world):
        from oracle.oracle_api import Oracle
        self.o = Oracle.from_core(core)
        self.o.set_shard(rank, world)
        self.rank, self.world = rank, world
        for a in ("nvar", "ncon", "nnzj", "nnzh"):
            setattr(self, a, getattr(self.o, a))

    def obj(self, x):
        return self.o.obj(x.numpy())

    def grad(self, x, g):
        g.copy_(torch.from_numpy(self.o.grad(x.numpy()))); return g

    def cons_nln(self, x, c):
        c.copy_(torch.from_numpy(self.o.cons(x.numpy()))); return c

    def jac_coord(self, x, v):
        v.copy_(torch.from_numpy(self.o.jac_coord(x.numpy()))); return v

    def hess_coord(self, x, y, v, obj_weight=1.0):
        v.copy_(torch.from_numpy(self.o.hess_coord(x.numpy(), y.numpy(), obj_weight))); return v

    def jprod_nln(self, x, v, out):
        out.copy_(torch.from_numpy(self.o.jprod(x.numpy(), v.numpy()))); return out

    def jtprod_nln(self, x, v, out):
        out.copy_(torch.from_numpy(self.o.jtprod(x.numpy(), v.numpy()))); return out

    def hprod(self, x, y, v, out, obj_weight=1.0):
        out.copy_(torch.from_numpy(self.o.hprod(x.numpy(), y.numpy(), v.numpy(), obj_weight))); return out


def _worker(rank, world, port, which, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import examodels_jl_b200 as E
        from examodels_jl_b200 import models as M
        from examodels_jl_b200.parallel import ShardedExaModel
        from oracle.oracle_api import Oracle
        core = {"lv": lambda: M.luksan_vlcek(57), "opf": lambda: M.ac_power(M.synthetic_power_data(23, 31, 5, seed=4)),
                "aug": lambda: M.luksan_vlcek_aug(11, 3)}[which]()
        plan = E.Plan(core)
        pats = [plan.pattern_info(k) for k in range(plan.npatterns())]
        full = Oracle.from_core(core)
        sm = ShardedExaModel(_OracleShard(core, rank, world), pats, gather=True)
        rng = np.random.default_rng(0)
        x = torch.from_numpy(core.meta()["x0"] + 0.01 * rng.uniform(-1, 1, full.nvar))
        y = torch.from_numpy(np.random.default_rng(1).standard_normal(full.ncon))
        tol = dict(rtol=1e-10, atol=1e-12)
        assert abs(sm.obj(x) - full.obj(x.numpy())) <= 1e-10 * max(1.0, abs(full.obj(x.numpy())))
        np.testing.assert_allclose(sm.grad(x, torch.empty(full.nvar, dtype=torch.float64)).numpy(), full.grad(x.numpy()), **tol)
        np.testing.assert_allclose(sm.cons_nln(x, torch.empty(full.ncon, dtype=torch.float64)).numpy(), full.cons(x.numpy()), **tol)
        j = sm.jac_coord(x, torch.full((full.nnzj,), float("nan"), dtype=torch.float64))
        assert np.array_equal(j.numpy(), full.jac_coord(x.numpy()))
        h = sm.hess_coord(x, y, torch.full((full.nnzh,), float("nan"), dtype=torch.float64), obj_weight=0.5)
        assert np.array_equal(h.numpy(), full.hess_coord(x.numpy(), y.numpy(), 0.5))
        # matrix-free products: partial products of the shards, summed
        v = torch.from_numpy(np.random.default_rng(2).standard_normal(full.nvar)); w = torch.from_numpy(np.random.default_rng(3).standard_normal(full.ncon))
        e = lambda n: torch.empty(n, dtype=torch.float64)   # noqa: E731
        np.testing.assert_allclose(sm.jprod_nln(x, v, e(full.ncon)).numpy(), full.jprod(x.numpy(), v.numpy()), **tol)
        np.testing.assert_allclose(sm.jtprod_nln(x, w, e(full.nvar)).numpy(), full.jtprod(x.numpy(), w.numpy()), **tol)
        np.testing.assert_allclose(sm.hprod(x, y, v, e(full.nvar), obj_weight=0.5).numpy(), full.hprod(x.numpy(), y.numpy(), v.numpy(), 0.5), **tol)
        # slices of all ranks tile each buffer exactly once
        for w_, n in ((1, full.nnzj), (2, full.nnzh)):
            cover = np.zeros(n, dtype=np.int64)
            for r in range(world):
                for lo, hi in sm.slices(w_, r):
                    cover[lo:hi] += 1
            assert (cover == 1).all()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world,which", [(2, "lv"), (2, "opf"), (3, "aug")])
def test_sharded_callbacks_gloo(world, which):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, which, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
