"""Multi-GPU evaluation: every pattern's iterator is split into `world` contiguous shards, one
process per GPU evaluates its shard, and the callbacks are completed with the collectives of
SURVEY.md §8e over torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).

  obj          partial scalar                       -> all_reduce(SUM), 8 bytes
  grad!        dense partial (neighbouring points share variables) -> all_reduce(SUM) over nvar
  cons!        own base rows + own augmentation terms, zeros elsewhere -> all_reduce(SUM) over ncon
  jac_coord! / hess_coord!
               each rank writes a CONTIGUOUS, non-overlapping slice per pattern
               [o + step*lo_r, o + step*hi_r): nothing to reduce.  `gather=False` leaves the COO
               buffer sharded (what a sharded consumer wants and what bench.py times);
               `gather=True` replicates it with one broadcast per (pattern, owner rank).
  jprod_nln! / jtprod_nln! / hprod!
               the fused product kernels of a shard return its partial product -> all_reduce(SUM) over ncon / nvar
  structures   computed replicated (each rank evaluates every point once, at build time).

When the rank-local evaluator carries the library's own NCCL communicator (`ExaModel.comm_init`, the C ABI's
exb_comm_*), every reducing callback completes itself inside the library on the caller's stream and this class only
forwards; the torch.distributed collectives below are the host-side fallback (gloo in the CPU tests, or a host that
prefers to own the communication).

The reference has no multi-device path at all (single device, ext/ExaModelsKernelAbstractions.jl);
this layer is new.  Summation order across ranks differs from the single-GPU order, so sharded
results are compared at 1e-10, not bitwise.
"""
from __future__ import annotations


def bind_near_gpu(device_index):
    """Pin the calling process to the CPU cores NVML reports as local to `device_index` (its NUMA node / PCIe root), so
    that page-locked staging buffers allocated afterwards are first-touched next to the GPU they feed.  With one process
    per GPU on a multi-socket host this is what keeps the host-buffer entry points (exb_host_*) at PCIe speed instead of
    crossing the socket interconnect.  Returns the CPU list, or None when NVML / affinity control is unavailable."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(device_index)
            bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def shard_range(n, rank, world):
    """Contiguous shard of an n-point iterator (same rule as exb_create: lo = n*r/W)."""
    return n * rank // world, n * (rank + 1) // world


class ShardedExaModel:
    """Wraps a rank-local evaluator (an `ExaModel(core, rank=r, world=W)`, or any object with the
    same callback methods) and a process group."""

    def __init__(self, local, patterns, group=None, gather=True):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.local, self.gather = local, gather
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert (local.rank, local.world) == (self.rank, self.world), "local evaluator shard != process rank"
        self.abi = bool(getattr(local, "has_comm", False))   # collectives run inside libexa_b200.so
        # patterns: list of dicts with kind, nitr, o1, o2, o1step, o2step (Plan.pattern_info)
        self.patterns = patterns
        for a in ("nvar", "ncon", "nnzj", "nnzh"):
            setattr(self, a, getattr(local, a))

    # -- reductions -------------------------------------------------------------------
    def obj(self, x):
        import torch
        if self.abi:
            return self.local.obj(x)
        t = torch.tensor([self.local.obj(x)], dtype=torch.float64, device=x.device)
        self.dist.all_reduce(t, group=self.group)
        return float(t.item())

    def grad(self, x, g):
        self.local.grad(x, g)
        if not self.abi:
            self.dist.all_reduce(g, group=self.group)
        return g

    def cons_nln(self, x, c):
        self.local.cons_nln(x, c)
        if not self.abi:
            self.dist.all_reduce(c, group=self.group)
        return c

    # -- matrix-free products: partial products of the shards add up ----------------------
    def jprod_nln(self, x, v, Jv):
        self.local.jprod_nln(x, v, Jv)
        if not self.abi:
            self.dist.all_reduce(Jv, group=self.group)
        return Jv

    def jtprod_nln(self, x, v, Jtv):
        self.local.jtprod_nln(x, v, Jtv)
        if not self.abi:
            self.dist.all_reduce(Jtv, group=self.group)
        return Jtv

    def hprod(self, x, y, v, Hv, obj_weight=1.0):
        self.local.hprod(x, y, v, Hv, obj_weight=obj_weight)
        if not self.abi:
            self.dist.all_reduce(Hv, group=self.group)
        return Hv

    # -- sharded COO outputs -------------------------------------------------------------
    def slices(self, which, rank):
        """Half-open slices of the jac (which=1) / hess (which=2) COO buffer that `rank` writes."""
        out = []
        for p in self.patterns:
            if which == 1 and p["kind"] == 0:
                continue
            o, step = (p["o1"], p["o1step"]) if which == 1 else (p["o2"], p["o2step"])
            lo, hi = shard_range(p["nitr"], rank, self.world)
            if step > 0 and hi > lo:
                out.append((o + step * lo, o + step * hi))
        return out

    def _replicate(self, vals, which):
        if not self.gather or self.world == 1:
            return vals
        if self.abi:
            return self.local.gather_coo(which, vals)
        jobs = []
        for r in range(self.world):
            src = self.dist.get_global_rank(self.group, r) if self.group is not None else r
            jobs += [(vals[lo:hi], src) for lo, hi in self.slices(which, r)]
        # host-side fallback (no communicator inside the library): one broadcast per (pattern, owner).  The sequence of
        # collectives is the same on every rank by construction; no try / except around live collectives.
        works = [self.dist.broadcast(t, src=src, group=self.group, async_op=True) for t, src in jobs]
        for w in works:
            w.wait()
        return vals

    def jac_coord(self, x, vals):
        self.local.jac_coord(x, vals)
        return self._replicate(vals, 1)

    def hess_coord(self, x, y, vals, obj_weight=1.0):
        self.local.hess_coord(x, y, vals, obj_weight=obj_weight)
        return self._replicate(vals, 2)
