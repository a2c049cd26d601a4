"""Strong-scaling sweep of the sharded path for BASELINE configs 4 and 5, launched with torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_scaling.py <tag>
Total work is FIXED (the config's model); every pattern's iterator is split into N contiguous shards.  Reports, as the
max over ranks of CUDA-event times: hess_coord! with the COO buffer left sharded (no collective), hess_coord! with the
slices replicated (broadcast per pattern and owner), and the full five-callback evaluation with its collectives
(all_reduce of obj / grad / cons; jac and hess left sharded)."""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
from examodels_jl_b200.parallel import ShardedExaModel

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


CFG = {"opf10k": lambda: M.ac_power(M.synthetic_power_data()), "family32x1e6": lambda: M.pattern_family(1_000_000, 32)}
out = []
for key, build in CFG.items():
    core = build()
    plan = E.Plan(core)
    pats = [plan.pattern_info(k) for k in range(plan.npatterns())]
    m = E.ExaModel(core, device=local, rank=rank, world=world)
    sm = ShardedExaModel(m, pats, gather=False)
    meta = core.meta()
    x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
    y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
    h, j, g, c = m.new(m.nnzh), m.new(m.nnzj), m.new(m.nvar), m.new(m.ncon)
    res = {"config": key, "n_gpus": world, "nnzh": m.nnzh}
    res["hess_sharded_ms"] = timeit(lambda: sm.hess_coord(x, y, h))
    sm.gather = True
    res["hess_replicated_ms"] = timeit(lambda: sm.hess_coord(x, y, h), 10)
    sm.gather = False

    def full():
        sm.obj(x); sm.grad(x, g); sm.cons_nln(x, c); sm.jac_coord(x, j); sm.hess_coord(x, y, h)
    res["full_eval_ms"] = timeit(full, 20)
    res["hess_nnz_per_s_sharded"] = m.nnzh / (res["hess_sharded_ms"] * 1e-3)
    res["evals_per_s"] = 1e3 / res["full_eval_ms"]
    if rank == 0:
        print(json.dumps(res), flush=True)
        out.append(res)
    del m, sm, h, j
    torch.cuda.empty_cache()
if rank == 0:
    json.dump(out, open(f"gpurun_out/scaling_{tag}_n{world}.json", "w"), indent=1)
dist.destroy_process_group()
