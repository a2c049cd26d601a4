"""Sharded evaluation on real GPU handles: two ranks (NCCL over two GPUs when the box has them, else both ranks
on GPU 0 with gloo), each an `ExaModel(core, rank=r, world=2)`, completed by the collectives of
examodels.jl_b200/parallel.py, against the unsharded oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, which, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    ngpu = torch.cuda.device_count()
    dev = rank % ngpu
    torch.cuda.set_device(dev)
    backend = "nccl" if ngpu >= world else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        import examodels_jl_b200 as E
        from examodels_jl_b200 import models as M
        from examodels_jl_b200.parallel import ShardedExaModel
        from oracle.oracle_api import Oracle
        from util import assert_close
        core = {"lv": lambda: M.luksan_vlcek(1003), "opf": lambda: M.ac_power(M.synthetic_power_data(300, 420, 70, seed=2)),
                "aug": lambda: M.luksan_vlcek_aug(21, 3)}[which]()
        plan = E.Plan(core)
        pats = [plan.pattern_info(k) for k in range(plan.npatterns())]
        full = Oracle.from_core(core)
        local = E.ExaModel(core, device=dev, rank=rank, world=world)
        sm = ShardedExaModel(local, pats, gather=True)
        x = core.meta()["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, full.nvar)
        y = np.random.default_rng(1).standard_normal(full.ncon)
        dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        ref = full.obj(x)
        assert abs(sm.obj(dx) - ref) <= 1e-10 * max(1.0, abs(ref))
        assert_close(sm.grad(dx, local.new(local.nvar)).cpu().numpy(), full.grad(x), "grad")
        assert_close(sm.cons_nln(dx, local.new(local.ncon)).cpu().numpy(), full.cons(x), "cons")
        assert_close(sm.jac_coord(dx, local.new(local.nnzj).fill_(float("nan"))).cpu().numpy(), full.jac_coord(x), "jac")
        assert_close(sm.hess_coord(dx, dy, local.new(local.nnzh).fill_(float("nan")), obj_weight=0.5).cpu().numpy(),
                     full.hess_coord(x, y, 0.5), "hess")
        v = torch.from_numpy(np.random.default_rng(2).standard_normal(full.nvar)).cuda()
        w = torch.from_numpy(np.random.default_rng(3).standard_normal(full.ncon)).cuda()
        assert_close(sm.jprod_nln(dx, v, local.new(local.ncon)).cpu().numpy(), full.jprod(x, v.cpu().numpy()), "jprod")
        assert_close(sm.jtprod_nln(dx, w, local.new(local.nvar)).cpu().numpy(), full.jtprod(x, w.cpu().numpy()), "jtprod")
        assert_close(sm.hprod(dx, dy, v, local.new(local.nvar), obj_weight=0.5).cpu().numpy(), full.hprod(x, y, v.cpu().numpy(), 0.5), "hprod")
        # sharded output left in place: only this rank's slices are written
        sm.gather = False
        h = sm.hess_coord(dx, dy, local.new(local.nnzh).fill_(float("nan")), obj_weight=0.5).cpu().numpy()
        mine = np.zeros(local.nnzh, dtype=bool)
        for lo, hi in sm.slices(2, rank):
            mine[lo:hi] = True
        assert not np.isnan(h[mine]).any() and np.isnan(h[~mine]).all()
        assert_close(h[mine], full.hess_coord(x, y, 0.5)[mine], "hess shard")
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("which", ["lv", "opf", "aug"])
def test_sharded_gpu_handles(which):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, which, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
