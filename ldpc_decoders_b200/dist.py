"""Multi-GPU plumbing: one process per GPU, frames sharded, counters all-reduced.

Frames are independent, so there is no data-path collective.  The only exchanges are
  * all_gather of the per-frame (bit errors, iteration count) of a round, so that every rank can apply the
    reference's sequential stopping rule `while wec < min_wec` (src/main.py:37) to the GLOBAL frame order, and
  * all_reduce(SUM) of Monte-Carlo counters.
Backend: NCCL over NVLink when the process group lives on GPUs, gloo in the CPU tests.
"""
import os

import numpy as np


def quiet_nccl_stdout():
    """NCCL writes its version banner (NCCL_DEBUG >= VERSION) and warnings to stdout; route them to stderr so that
    stdout carries results only (bench.py prints exactly one JSON line).  An explicit NCCL_DEBUG_FILE wins."""
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")


def bind_near_gpu(index):
    """Restrict this process to the CPUs local to CUDA device `index` (sysfs `local_cpulist` of its PCI function), so
    that the pinned host buffers it allocates afterwards, and the thread that feeds the copy engines, sit on the
    GPU's own NUMA node — with one process per GPU the host side of decode_host otherwise crosses the socket link.
    Returns the CPU set applied, or None when the topology is not visible (containers) or LDPC_NUMA_BIND=0."""
    if os.environ.get("LDPC_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as fp:
            cpus = parse_cpulist(fp.read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def bind_memory_near_gpu(index):
    """Prefer the NUMA node of CUDA device `index` for every page this process allocates from now on — in particular the
    pinned host buffers of decode_host, whose pages are placed when cudaHostAlloc pins them — via
    set_mempolicy(MPOL_PREFERRED) (CPU affinity alone does not move memory).  Returns the node, or None when the
    topology is not visible / the box is a single node / the syscall is refused (containers) / LDPC_NUMA_BIND=0."""
    if os.environ.get("LDPC_NUMA_BIND", "1") == "0":
        return None
    try:
        import ctypes
        import platform
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fp:
            node = int(fp.read().strip())
        if node < 0 or not os.path.isdir("/sys/devices/system/node/node%d" % node):
            return None
        nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
        if nr is None:
            return None
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED = 1
        rc = ctypes.CDLL(None, use_errno=True).syscall(nr, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
        return node if rc == 0 else None
    except Exception:
        return None


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0,1,2,3,8,10,11} (the kernel's cpulist format)."""
    out = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


class Comm:
    """Thin wrapper over torch.distributed that also works as a 1-rank no-op."""

    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if not dist.is_initialized():
                if backend == "nccl":
                    quiet_nccl_stdout()
                    torch.cuda.set_device(self.local_rank)
                    bind_near_gpu(self.local_rank)
                    bind_memory_near_gpu(self.local_rank)
                    dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
                else:
                    dist.init_process_group(backend)
            self.dist = dist
            self.device = "cuda" if dist.get_backend() == "nccl" else "cpu"

    def allreduce_sum(self, counters):
        """int64 numpy vector -> element-wise sum over ranks."""
        counters = np.asarray(counters, np.int64)
        if self.dist is None:
            return counters.copy()
        import torch
        t = torch.from_numpy(counters.copy()).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def allgather(self, arr):
        """Equal-length numpy vector per rank -> [world, len] in rank order."""
        arr = np.ascontiguousarray(arr)
        if self.dist is None:
            return arr[None, :].copy()
        import torch
        t = torch.from_numpy(arr.copy()).to(self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    # ---- the same collectives on torch tensors that already live where the backend wants them (NCCL: on the GPU):
    # no host round trip on the way in, the caller decides when (and whether) the result is read back
    def _native(self, t):
        return t if (self.device == "cuda") == t.is_cuda else t.to(self.device)

    def allreduce_tensor(self, t):
        """Element-wise SUM over ranks of an int64 tensor, returned on the tensor's own device (in place when possible)."""
        if self.dist is None:
            return t
        w = self._native(t)
        self.dist.all_reduce(w, op=self.dist.ReduceOp.SUM)
        if w is not t:
            t.copy_(w)
        return t

    def allgather_tensor(self, t):
        """Equal-length 1-D tensor per rank -> [world, len] in rank order, on the tensor's own device."""
        import torch
        if self.dist is None:
            return t[None, :]
        w = self._native(t.contiguous())
        out = torch.empty((self.world,) + tuple(w.shape), dtype=w.dtype, device=w.device)
        self.dist.all_gather_into_tensor(out, w)
        return out if out.device == t.device else out.to(t.device)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
            self.dist = None


def round_slice(round_idx, rank, world, batch):
    """Global frame indices of `rank` in round `round_idx`: rounds of world*batch frames, rank-major inside a
    round: g in [k*R*B + r*B, k*R*B + (r+1)*B)."""
    g0 = (round_idx * world + rank) * batch
    return g0, g0 + batch


def sequential_stop(bit_errs_global, wec_so_far, min_wec):
    """The reference counts frames one by one and stops as soon as `wec >= min_wec` (src/main.py:37-45).
    Given the bit-error counts of the next frames in GLOBAL order, return how many of them are consumed."""
    we = np.asarray(bit_errs_global) > 0
    need = min_wec - wec_so_far
    if need <= 0:
        return 0
    c = np.cumsum(we)
    hit = np.flatnonzero(c >= need)
    return int(hit[0]) + 1 if hit.size else int(we.size)
