"""Where a host-consumer step (16384 envs, ozl_step_host + stream sync) spends its time: the same step with the action reads and /
or the result writes redirected to device memory.  Prints one JSON line (us per step, host wall clock incl. launch + sync)."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ouzelum_b200 import _lib  # noqa: E402
from ouzelum_b200.sim import QuadSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda:0")
sim = QuadSim(_lib.default_cfg(n, fault_mode=1, seed=0), dev)
rs, pg = torch.ones(n, dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.int64, device=dev)
to, er = torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, device=dev)
pin = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt).pin_memory()
h = dict(act=(torch.rand(n, 4) * 2 - 1).pin_memory(), obs=pin(n, 13), rew=pin(n), done=pin(n, dt=torch.uint8), rst=pin(n, dt=torch.int64))
d = dict(act=h["act"].to(dev), obs=torch.empty(n, 13, device=dev), rew=torch.empty(n, device=dev),
         done=torch.empty(n, dtype=torch.uint8, device=dev), rst=torch.empty(n, dtype=torch.int64, device=dev))
stream = torch.cuda.current_stream().cuda_stream


def run(act, out, label, steps=2000):
    io = _lib.OzlHostIo(act["act"].data_ptr(), out["obs"].data_ptr(), out["rew"].data_ptr(), out["done"].data_ptr(), out["rst"].data_ptr(),
                        rs.data_ptr(), pg.data_ptr(), to.data_ptr(), er.data_ptr())
    ref = C.byref(io)
    for _ in range(200):
        _lib.lib.ozl_step_host_sync(sim._h, ref, stream)
    t0 = time.perf_counter()
    for _ in range(steps):
        _lib.lib.ozl_step_host_sync(sim._h, ref, stream)
    return label, (time.perf_counter() - t0) / steps * 1e6


res = dict([run(h, h, "host_actions__host_results"), run(d, h, "device_actions__host_results"),
            run(h, d, "host_actions__device_results"), run(d, d, "device_actions__device_results")])
res["n_envs"] = n
res["h2d_bytes"], res["d2h_bytes"] = n * 16, n * (13 * 4 + 4 + 1 + 8)
print(json.dumps(res), flush=True)
