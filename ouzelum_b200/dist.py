"""Multi-GPU plumbing: env sharding and the metrics all-reduce (the only collective on the path).

Envs are independent, so N_total envs are split into contiguous blocks of global env ids, one process per
GPU; the counter RNG is keyed by GLOBAL env id, so results do not depend on the number of ranks.  No per-step
communication.  `torch.distributed` (NCCL over NVLink on the GPUs, gloo in the CPU tests) is plumbing only.
"""
import os

import torch
import torch.distributed as dist


def rank_info():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin the calling process to the CPUs (and therefore, by first touch, the memory) of the NUMA node the GPU hangs off.

    A host consumer exchanges ~1.3 MB per 16384-env step with its GPU through pinned buffers (zero-copy PCIe reads / writes);
    torchrun does not bind its workers, so on a two-socket box half of the ranks would otherwise reach their GPU across the
    inter-socket link.  Returns the CPU list it bound to (None if the topology could not be read -- nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                       # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus = sorted(cpus & allowed) or None
        if cpus:
            os.sched_setaffinity(0, cpus)
        _prefer_numa_node(f"/sys/bus/pci/devices/{bus}/numa_node")
        return cpus
    except Exception:  # noqa: BLE001 -- best effort: containers may hide sysfs / NVML
        return None


def _prefer_numa_node(numa_node_file):
    """set_mempolicy(MPOL_PREFERRED, {node}) for the calling thread: pages touched from now on (the pinned I/O buffers) come from the
    GPU's own NUMA node even when the CPUs this process may run on sit on another one (containers often expose the CPUs of a single
    socket).  Best effort: a raw syscall, no libnuma needed; silently does nothing when the node is unknown or not allowed."""
    try:
        import ctypes
        with open(numa_node_file) as f:
            node = int(f.read().strip())
        if node < 0:
            return False
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        libc = ctypes.CDLL(None, use_errno=True)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238                      # x86_64
        return libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, mask, 16 * 64) == 0
    except Exception:  # noqa: BLE001
        return False


def shard(total_envs: int, rank: int, world: int):
    """Contiguous block owned by `rank`: returns (num_local_envs, env_id_base).  The first `total % world` ranks get one
    extra env, so any total is covered exactly once."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    q, r = divmod(int(total_envs), int(world))
    n = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return n, base


def allreduce_metrics(metrics: torch.Tensor, async_op: bool = False):
    """Sum the 16-double metrics vector over all ranks (in place).  No-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(metrics, op=dist.ReduceOp.SUM, async_op=async_op)
    return None


def summarize(metrics: torch.Tensor):
    """Human-readable episode / reward statistics from a (reduced) metrics vector."""
    m = metrics.detach().cpu().tolist()
    steps, eps = max(m[8], 1.0), max(m[9], 1.0)
    return {"env_steps": m[8], "episodes": m[9], "mean_reward": m[0] / steps, "mean_episode_return": m[1] / eps,
            "mean_episode_length": m[10] / eps, "timeouts": m[11], "crash_dist": m[12], "crash_z": m[13],
            "fault_active_frac": m[14] / steps, "landed_episodes": m[2], "resets": m[15]}


class PeerMetrics:
    """Sum of the 16-double metrics vector over all ranks through NVLink peer memory (csrc/peer_metrics.cu): every rank's
    metrics-read kernel stores its values straight into every peer's mailbox; no NCCL kernel, no side stream, CUDA-graph
    capturable.  One process per GPU (`torch.distributed` only carries the 64-byte IPC handles at start-up), or several devices
    driven by one process (`connect_same_process`).

        pm = PeerMetrics(device)                  # after init_process_group(...)
        pm.push(sim, prev_sum=buf16)              # every few steps: one launch; buf16 <- sum of the PREVIOUS exchange
        pm.sum(sim, out=buf16)                    # when the reduced vector of the last push is needed
    """

    def __init__(self, device, rank=None, world=None, connect=True):
        import ctypes as C
        from ._lib import check, lib
        self.device = torch.device(device)
        have_pg = dist.is_available() and dist.is_initialized()
        self.rank = int(rank if rank is not None else (dist.get_rank() if have_pg else 0))
        self.world = int(world if world is not None else (dist.get_world_size() if have_pg else 1))
        h = C.c_void_p()
        check(lib.ozl_metrics_xchg_create(self.rank, self.world, self.device.index or 0, C.byref(h)))
        self._h = h
        if connect and self.world > 1:
            self._connect_ipc()

    def _connect_ipc(self):
        """Collective: every rank of the process group must call it.  Either every rank ends up connected or every rank raises (a
        rank that fails to map a peer's mailbox still takes part in the collectives below, so nobody is left waiting)."""
        import ctypes as C
        from ._lib import lib, last_error
        mine = (C.c_ubyte * 64)()
        ok, why = True, ""
        if lib.ozl_metrics_xchg_ipc_handle(self._h, mine) != 0:
            ok, why = False, last_error()
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(mine) if ok else None)          # host-side, once: plumbing
        if ok and all(g is not None for g in gathered):
            blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(gathered))
            if lib.ozl_metrics_xchg_connect_ipc(self._h, blob) != 0:
                ok, why = False, last_error()
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)            # also the barrier: every mailbox is mapped before anyone pushes
        if int(flag.item()) == 0:
            raise RuntimeError(f"ouzelum_b200: NVLink metrics exchange could not be connected on every rank ({why or 'a peer failed'})")

    @staticmethod
    def connect_same_process(exchanges):
        """Several devices driven by ONE process: enable peer access and hand every exchange the others' mailboxes."""
        import ctypes as C
        from ._lib import check, lib
        boxes = (C.c_void_p * len(exchanges))()
        devs = (C.c_int32 * len(exchanges))()
        for i, e in enumerate(exchanges):
            b = C.c_void_p()
            check(lib.ozl_metrics_xchg_box(e._h, C.byref(b)))
            boxes[i], devs[i] = b.value, e.device.index or 0
        for e in exchanges:
            check(lib.ozl_metrics_xchg_connect_ptrs(e._h, boxes, devs))

    def push(self, sim, local=None, prev_sum=None, clear=False):
        from ._lib import check, lib, ptr
        check(lib.ozl_metrics_push(sim._h, self._h, ptr(local), ptr(prev_sum), 1 if clear else 0,
                                   torch._C._cuda_getCurrentRawStream(self.device.index or 0)))

    def sum(self, sim, out=None):
        from ._lib import check, lib
        if out is None:
            out = torch.zeros(16, dtype=torch.float64, device=self.device)
        check(lib.ozl_metrics_sum(sim._h, self._h, out.data_ptr(), torch._C._cuda_getCurrentRawStream(self.device.index or 0)))
        return out

    def status(self):
        import ctypes as C
        from ._lib import check, lib
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib.ozl_metrics_xchg_status(self._h, C.byref(a), C.byref(b), C.byref(c),
                                          torch._C._cuda_getCurrentRawStream(self.device.index or 0)))
        return {"pushed": a.value, "summed": b.value, "error": c.value}

    def close(self):
        from ._lib import lib
        if getattr(self, "_h", None):
            lib.ozl_metrics_xchg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
