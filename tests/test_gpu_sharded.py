"""Sharded evaluation on real GPU handles: two ranks, each an `ExaModel(core, rank=r, world=2)`, against the unsharded
oracle.  On a box with >= 2 GPUs: NCCL, and the collectives are exercised three ways -- on the host side
(examodels.jl_b200/parallel.py over torch.distributed), inside the library through its own communicator (exb_comm_*, replicate
mode), and in owner mode (sharded consumer).  On a 1-GPU box both ranks share GPU 0 and only the host-side form runs, over
gloo (NCCL refuses two ranks on one device)."""
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
        x = core.meta()["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, full.nvar)
        y = np.random.default_rng(1).standard_normal(full.ncon)
        dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        v = torch.from_numpy(np.random.default_rng(2).standard_normal(full.nvar)).cuda()
        w = torch.from_numpy(np.random.default_rng(3).standard_normal(full.ncon)).cuda()
        nan = float("nan")

        def checks(sm):
            ref = full.obj(x)
            assert abs(sm.obj(dx) - ref) <= 1e-10 * max(1.0, abs(ref))
            assert_close(sm.grad(dx, local.new(local.nvar).fill_(nan)).cpu().numpy(), full.grad(x), "grad")
            assert_close(sm.cons_nln(dx, local.new(local.ncon).fill_(nan)).cpu().numpy(), full.cons(x), "cons")
            assert_close(sm.jac_coord(dx, local.new(local.nnzj).fill_(nan)).cpu().numpy(), full.jac_coord(x), "jac")
            assert_close(sm.hess_coord(dx, dy, local.new(local.nnzh).fill_(nan), obj_weight=0.5).cpu().numpy(),
                         full.hess_coord(x, y, 0.5), "hess")
            assert_close(sm.jprod_nln(dx, v, local.new(local.ncon).fill_(nan)).cpu().numpy(), full.jprod(x, v.cpu().numpy()), "jprod")
            assert_close(sm.jtprod_nln(dx, w, local.new(local.nvar).fill_(nan)).cpu().numpy(), full.jtprod(x, w.cpu().numpy()), "jtprod")
            assert_close(sm.hprod(dx, dy, v, local.new(local.nvar).fill_(nan), obj_weight=0.5).cpu().numpy(), full.hprod(x, y, v.cpu().numpy(), 0.5), "hprod")
            # sharded output left in place: only this rank's slices are written
            sm.gather = False
            h = sm.hess_coord(dx, dy, local.new(local.nnzh).fill_(nan), obj_weight=0.5).cpu().numpy()
            mine = np.zeros(local.nnzh, dtype=bool)
            for lo, hi in sm.slices(2, rank):
                mine[lo:hi] = True
            assert not np.isnan(h[mine]).any() and np.isnan(h[~mine]).all()
            assert_close(h[mine], full.hess_coord(x, y, 0.5)[mine], "hess shard")
            sm.gather = True

        # (1) collectives on the host side (torch.distributed): the handle returns partial results
        sm = ShardedExaModel(local, pats, gather=True)
        assert not sm.abi
        checks(sm)
        if backend == "nccl":
            # (2) the library's own NCCL communicator (exb_comm_init): every reducing callback completes itself
            local.comm_init(mode="replicate")
            sm = ShardedExaModel(local, pats, gather=True)
            assert sm.abi
            checks(sm)
            st = local.comm_stats()
            assert st["attached"] == 1 and st["collectives"] >= 7
            # the fused sweep on a sharded handle: the objective's all-reduce runs on a side stream next to the finishing steps
            od, g2, c2 = local.new(1).fill_(nan), local.new(local.nvar).fill_(nan), local.new(local.ncon).fill_(nan)
            j2, h2 = local.new(local.nnzj).fill_(nan), local.new(local.nnzh).fill_(nan)
            for _ in range(2):
                local.eval_all(dx, dy, od, g2, c2, j2, h2, obj_weight=0.5)
            torch.cuda.synchronize()
            assert abs(float(od.item()) - full.obj(x)) <= 1e-10 * max(1.0, abs(full.obj(x)))
            assert_close(g2.cpu().numpy(), full.grad(x), "eval grad (replicate)")
            assert_close(c2.cpu().numpy(), full.cons(x), "eval cons (replicate)")
            hh = h2.cpu().numpy(); own = ~np.isnan(hh)
            assert own.any() and not own.all()
            assert_close(hh[own], full.hess_coord(x, y, 0.5)[own], "eval hess (own slices)")
            o1 = local.new(1).fill_(nan)
            local.eval_all(dx, None, o1, g2, c2, j2, None, mask=15)
            torch.cuda.synchronize()
            assert abs(float(o1.item()) - full.obj(x)) <= 1e-10 * max(1.0, abs(full.obj(x)))
            # (3) owner mode, the sharded consumer: g on the owned variables, c on the rows of the own points
            local.comm_set_mode("owner")
            lo, hi = local.owned()
            assert (lo, hi) == (full.nvar * rank // world, full.nvar * (rank + 1) // world)
            g = local.grad(dx, local.new(local.nvar).fill_(nan)).cpu().numpy()
            assert_close(g[lo:hi], full.grad(x)[lo:hi], "grad (owned variables)")
            if which == "lv":   # shift-indexed objective: owner-computed per variable, nothing is exchanged
                assert local.comm_stats()["last_collectives"] == 0
            c = local.cons_nln(dx, local.new(local.ncon).fill_(nan)).cpu().numpy()
            cref = full.cons(x)
            for k, pt in enumerate(pats):
                if pt["kind"] == 1:
                    sh = local.shard(k)
                    assert_close(c[pt["o0"] + sh["lo"]:pt["o0"] + sh["hi"]], cref[pt["o0"] + sh["lo"]:pt["o0"] + sh["hi"]], "cons (own rows)")
            if which == "lv":
                assert local.comm_stats()["last_collectives"] == 0
            assert abs(local.obj(dx) - full.obj(x)) <= 1e-10 * max(1.0, abs(full.obj(x)))
            local.comm_destroy()
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
