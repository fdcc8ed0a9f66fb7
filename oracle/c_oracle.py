"""ctypes front-end of the C restatement `oracle/quad_step_c.c` -- TEST INFRASTRUCTURE.

Two uses: (1) a second, independent CPU oracle (tests check it bit-for-bit against the torch oracle and, on the GPU box,
against the kernel); (2) the multi-threaded CPU baseline of bench.py (`--impl reference`, `cpu_baseline`): the same step
on all host cores, which is the strongest CPU arm this repo can field (Isaac Gym's CPU pipeline cannot run).
The shared object is built on demand into oracle/_build/ (git-ignored) with gcc; nothing here touches the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "quad_step_c.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle_quad.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    hdr = os.path.join(ROOT, "include", "ouzelum_b200.h")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        return OUT
    # -mfma: the explicit fmaf() calls of the integrator become one vfmadd instruction (without it glibc's software fmaf, equally
    # exact, is called); -ffp-contract=off keeps every OTHER a*b+c unfused.  Only when the build host has FMA3.
    try:
        has_fma = " fma " in open("/proc/cpuinfo").read()
    except OSError:
        has_fma = False
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math"] + (["-mfma"] if has_fma else []) + \
          ["-fopenmp", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", OUT, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed: " + r.stderr)
    return OUT


class OracleDrParam(C.Structure):
    _fields_ = [("distribution", C.c_int32), ("operation", C.c_int32), ("range", C.c_float * 2),
                ("schedule", C.c_int32), ("schedule_steps", C.c_int32)]


class OracleCfg(C.Structure):
    """Independent mirror of `struct ozl_cfg` (include/ouzelum_b200.h) so that the CPU arm needs nothing from the product."""
    _fields_ = [
        ("abi_version", C.c_int32), ("reserved0", C.c_int32),
        ("num_envs", C.c_int64), ("env_id_base", C.c_int64), ("seed", C.c_uint64),
        ("max_episode_length", C.c_int32), ("target_period", C.c_int32), ("target_fixed", C.c_int32),
        ("substeps", C.c_int32), ("control_freq_inv", C.c_int32),
        ("dt", C.c_float), ("gravity_z", C.c_float), ("clip_actions", C.c_float), ("clip_obs", C.c_float),
        ("thrust_rate", C.c_float), ("thrust_max", C.c_float),
        ("die_dist", C.c_float), ("die_z", C.c_float), ("up_coef", C.c_float),
        ("spawn_base", C.c_float * 3), ("spawn_lo", C.c_float * 3), ("spawn_range", C.c_float * 3),
        ("target_scale", C.c_float * 3), ("target_off", C.c_float * 3),
        ("mass", C.c_float), ("ixx", C.c_float), ("iyy", C.c_float), ("izz", C.c_float),
        ("arm", C.c_float), ("com_z", C.c_float), ("max_angvel", C.c_float),
        ("lin_drag", C.c_float), ("yaw_km", C.c_float),
        ("fault_mode", C.c_int32), ("fault_eff_lo", C.c_float), ("fault_eff_range", C.c_float),
        ("dr_enable", C.c_int32), ("dr", OracleDrParam * 7),
        ("pomdp_mode", C.c_int32), ("pomdp_prob", C.c_float), ("noise_sigma", C.c_float),
        ("collect_metrics", C.c_int32), ("plate_enable", C.c_int32), ("plate_z", C.c_float), ("plate_radius", C.c_float),
        ("land_cutoff", C.c_float), ("wrench_warmup_steps", C.c_int32), ("reserved1", C.c_int32),
    ]


def make_cfg(cfg_dict):
    """Build the C struct from an oracle config dict (oracle.quad_step.default_cfg)."""
    c = OracleCfg()
    names = dict(OracleCfg._fields_)
    for k, v in cfg_dict.items():
        if k not in names:
            continue
        if k == "dr":
            for j, spec in enumerate(v):
                d = c.dr[j]
                d.distribution, d.operation, d.schedule, d.schedule_steps = int(spec[0]), int(spec[1]), int(spec[4]), int(spec[5])
                d.range[0], d.range[1] = float(spec[2]), float(spec[3])
        elif isinstance(v, (tuple, list)):
            getattr(c, k)[:] = [float(x) for x in v]
        else:
            setattr(c, k, v)
    c.abi_version = 3
    return c


class COracle:
    """State-holding wrapper with the interface of oracle.quad_step.QuadStepOracle (rotor-action mode)."""

    def __init__(self, cfg_struct, threads=None):
        """cfg_struct: a ctypes structure laid out as `struct ozl_cfg` (tests pass ouzelum_b200._lib.OzlCfg; bench builds one
        from the defaults through `make_cfg`)."""
        self.lib = C.CDLL(build())
        assert self.lib.ozl_oracle_cfg_size() == C.sizeof(cfg_struct), "ozl_cfg layout mismatch"
        self.cfg = cfg_struct
        # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host thread it can
        self.threads = self.lib.ozl_oracle_set_threads(int(threads or os.cpu_count() or 1))
        n = self.n = int(cfg_struct.num_envs)
        f = np.float32
        self.root = np.zeros((n, 13), f)
        self.root[:, 0:3] = np.asarray(list(cfg_struct.spawn_base), f)
        self.root[:, 6] = 1
        self.thrust, self.target, self.ep_ret = np.zeros((n, 4), f), np.zeros((n, 3), f), np.zeros(n, f)
        self.target[:, 2] = 1
        self.params = np.zeros((n, 8), f)     # mass, ixx, iyy, izz, arm, thrust scale, fault effectiveness, yaw_km
        self.params[:] = [cfg_struct.mass, cfg_struct.ixx, cfg_struct.iyy, cfg_struct.izz, cfg_struct.arm, 1.0, 1.0, cfg_struct.yaw_km]
        self.fault = np.zeros((n, 2), np.int32)
        self.fault[:, 1] = 0x1FFFFFFF
        self.obs_buf, self.rew_buf = np.zeros((n, 13), f), np.zeros(n, f)
        self.reset_buf, self.progress_buf = np.ones(n, np.int64), np.zeros(n, np.int64)
        self.timeout_buf = np.zeros(n, np.uint8)
        self.step_count = 0
        self.lib.ozl_oracle_step.restype = None
        self.lib.ozl_oracle_step.argtypes = [C.c_void_p, C.c_uint64, C.c_int64] + [C.c_void_p] * 12

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.float32))
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        self.lib.ozl_oracle_step(C.byref(self.cfg), self.step_count, self.n, p(self.root), p(self.thrust), p(self.target),
                                 p(self.ep_ret), p(self.params), p(self.fault), p(a), p(self.obs_buf), p(self.rew_buf),
                                 p(self.reset_buf), p(self.progress_buf), p(self.timeout_buf))
        self.step_count += 1
        return self.obs_buf, self.rew_buf, self.reset_buf, self.timeout_buf
