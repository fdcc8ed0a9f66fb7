"""Landing-family x500 tasks: the target rides on a ground vehicle.

Mirrors isaacgymenvs/tasks/lando.py, landing.py and landed.py.  All three share Ouzelum's pre/post-physics code except
  * target source: `target.xy = husky.xy; target.x += 0.08; target.z = 0.377` every step (landing.py:76,373-374)
  * reward: die when z < 0.3 instead of 0.5 (landing.py:447)
  * Landing: the Husky follows lemniscate / circle / square waypoints (landing.py:108-112,208-244,319-364)
  * Landed: evaluation variant -- observation through the sensor-fault model env-side (landed.py:62,340), landing
    detector that cuts thrust within 0.2 m of the target (landed.py:288-295), landing counter (landed.py:265-271)
One step = ONE launch (`ozl_landing_step`: vehicle + target + tracking step per env in one thread); `env.fusedStep = False`
keeps the two-launch sequence `ozl_husky_step` -> `ozl_step_tracking` for A/B tests (identical bits).
The PhysX Husky and the leg/plate contact are replaced by a kinematic unicycle and an inelastic plate.
"""
import torch

from ..trajectories import HuskyFollower
from .ouzelum import X500Task, x500_cfg_from_task

TARGET_Z = 0.377          # landing.py:76
TOP_PLATE_X_SHIFT = 0.08  # landing.py:374


class _VehicleTargetTask(X500Task):
    die_z = 0.3
    x_offset = TOP_PLATE_X_SHIFT
    vehicle_moves = True
    land_cutoff = 0.0

    def _native_cfg(self):
        return x500_cfg_from_task(self.cfg, self.num_envs, target_fixed=1, die_z=self.die_z, plate_enable=1,
                                  plate_z=TARGET_Z, plate_radius=0.35, land_cutoff=self.land_cutoff)

    def create_sim(self):
        super().create_sim()
        env = self.cfg["env"]
        self.husky = HuskyFollower(self.num_envs, self.device, seed=int(env.get("seed", 0)),
                                   env_id_base=int(env.get("envIdBase", 0)), dt=float(self.cfg["sim"]["dt"]),
                                   x_offset=self.x_offset, target_z=TARGET_Z, env_spacing=float(env.get("envSpacing", 2.5)))
        # target_root_positions starts at (0,0,0.377) (landing.py:75-76)
        t0 = torch.zeros(self.num_envs, 3, device=self.device)
        t0[:, 2] = TARGET_Z
        self.sim.set_state(target=t0)
        self._target = self.husky.target
        # the vehicle kernel follows the env's device step counter: no host-changing launch argument (CUDA-graph capturable)
        import ctypes as C
        from .._lib import check, lib
        ctr = C.c_void_p()
        check(lib.ozl_step_counter_ptr(self.sim._h, C.byref(ctr)))
        self.husky.follow_step_counter(ctr.value)

    def _launch(self, actions):
        if self.vehicle_moves and bool(self.cfg["env"].get("fusedStep", True)):
            import ctypes as C
            from .._lib import check, lib
            check(lib.ozl_landing_step(self.sim._h, actions.data_ptr(), C.byref(self.husky._a), self.obs_buf.data_ptr(),
                                       self.rew_buf.data_ptr(), self.reset_buf.data_ptr(), self.progress_buf.data_ptr(),
                                       self._timeout_u8.data_ptr(), self.episode_return_buf.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream))
            self._target = self.husky.target
            return
        if self.vehicle_moves:
            self._target = self.husky.step(self.reset_buf)
        self.sim.step_tracking(actions, self._target, self.obs_buf, self.rew_buf, self.reset_buf, self.progress_buf,
                               self._timeout_u8, self.episode_return_buf)

    @property
    def husky_positions(self):
        return self.husky.husky_positions

    # ---- checkpoint / resume: the vehicle is part of the env state
    def state_dict(self):
        sd = super().state_dict()
        sd.update(husky_pose=self.husky.pose.cpu(), husky_idx=self.husky.idx.cpu(), husky_target=self.husky.target.cpu(),
                  husky_step_count=int(self.husky.step_count))
        return sd

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        self.husky.pose.copy_(sd["husky_pose"])
        self.husky.idx.copy_(sd["husky_idx"])
        self.husky.target.copy_(sd["husky_target"])
        self.husky.step_count = int(sd["husky_step_count"])
        self._target = self.husky.target


class Lando(_VehicleTargetTask):
    """isaacgymenvs/tasks/lando.py: the Husky is driven with constant opposing wheel targets (lando.py, create_envs:
    set_actor_dof_velocity_targets [10,-20,20,-10]) -- it turns on the spot, so the target stays at the spawn point."""
    vehicle_moves = False


class Landing(_VehicleTargetTask):
    """isaacgymenvs/tasks/landing.py: waypoint-following Husky; `extras["reset_ids"]` as at landing.py:379."""

    def step(self, actions):
        out = super().step(actions)
        return out

    def get_reset_ids(self):
        return self.reset_buf.nonzero(as_tuple=False).squeeze(-1)       # landing.py:283-284 (host sync; only when asked)


class Landed(_VehicleTargetTask):
    """isaacgymenvs/tasks/landed.py: evaluation variant (the reference supports num_envs == 1 only, landed.py:290)."""
    land_cutoff = 0.2                                                  # landed.py:290

    def __init__(self, cfg, *a, **k):
        cfg["env"].setdefault("POMDP", "flicker")                      # landed.py:62: POMDPWrapper(pomdp='flicker', ...)
        if cfg["env"].get("POMDP") in (None, "none"):
            cfg["env"]["POMDP"] = "flicker"
        cfg["env"].setdefault("pomdp_prob", 0.01)
        super().__init__(cfg, *a, **k)

    @property
    def landings(self):
        """Episodes that ended after the vehicle was reached (the reference writes this to metrics/<pomdp>_<p>.txt)."""
        return int(self.sim.metrics()[2].item())

    # ---- on-disk formats of the reference's evaluation runs (opt-in: they cost a host sync per step) ----------------
    def enable_logging(self, log_dir):
        """Reproduce the files `Landed` writes (landed.py:114-117,265-271,346-353):
          <log_dir>/trajectories/<pomdp>_<prob>_ep_<k>.csv   one row per step: drone x,y,z, target x,y,z of env 0
                                                            (ep_0 starts with the header 'Position X,Position Y,Position Z')
          <log_dir>/metrics/<pomdp>_<prob>.txt               running count of episodes that ended after a landing
        so the MATLAB plotting scripts (isaacgymenvs/trajectories/csvreadf.m) keep working."""
        import csv
        import os
        self._log_dir = log_dir
        self._csv = csv
        env = self.cfg["env"]
        self._log_tag = f"{env.get('POMDP', 'flicker')}_{env.get('pomdp_prob', 0.01)}"
        os.makedirs(os.path.join(log_dir, "trajectories"), exist_ok=True)
        os.makedirs(os.path.join(log_dir, "metrics"), exist_ok=True)
        self.epi = 0
        with open(self._traj_path(), "w") as f:
            csv.writer(f).writerow(["Position X", "Position Y", "Position Z"])

    def _traj_path(self):
        import os
        return os.path.join(self._log_dir, "trajectories", f"{self._log_tag}_ep_{self.epi}.csv")

    def _launch(self, actions):
        log = getattr(self, "_log_dir", None)
        if log and bool(self.reset_buf[0].item()):                 # pre_physics_step: env 0 is being reset (landed.py:262-264)
            self.epi += 1
        super()._launch(actions)
        if log:
            import os
            st = self.sim.get_state()
            row = torch.cat([st["root"][0, 0:3], st["target"][0]]).cpu().tolist()          # landed.py:342-353
            with open(self._traj_path(), "a") as f:
                self._csv.writer(f).writerow(row)
            with open(os.path.join(log, "metrics", f"{self._log_tag}.txt"), "w") as f:      # landed.py:269-271
                f.write(str(self.landings))
