"""`Controller`: same constructor / call contract as isaacgymenvs/controllers/controller.py:19-48, but the ~60 batched
torch launches of a Lee controller call are ONE CUDA kernel (`ozl_lee_control`)."""
import ctypes as C

import torch

from .._lib import check, lib

control_class_dict = {"lee_position_control": 0, "lee_velocity_control": 1, "lee_attitude_control": 2}


class Controller:
    def __init__(self, control_config, device):
        self.control_config = control_config
        self.device = device
        self.controller_name = control_config.controller
        if self.controller_name not in control_class_dict:
            raise ValueError("Invalid controller name: {}".format(self.controller_name))      # controller.py:32-33
        if torch.device(device).type != "cuda":
            raise RuntimeError("ouzelum_b200 controllers run on CUDA only (no CPU fallback)")
        self.mode = control_class_dict[self.controller_name]
        self.kP = torch.tensor(control_config.kP, dtype=torch.float32)
        self.kV = torch.tensor(control_config.kV, dtype=torch.float32)
        self.kR = torch.tensor(control_config.kR, dtype=torch.float32)
        self.kOmega = torch.tensor(control_config.kOmega, dtype=torch.float32)
        self.scale_input = torch.tensor(control_config.scale_input, dtype=torch.float32)
        g = list(control_config.kP) + list(control_config.kV) + list(control_config.kR) + list(control_config.kOmega) + \
            list(control_config.scale_input)
        self._gains = (C.c_float * 16)(*[float(x) for x in g])

    def __call__(self, robot_state, command_actions, out=None):
        """robot_state [N,13], command_actions [N,4] -> (thrust [N], torque [N,3])."""
        st = robot_state.to(dtype=torch.float32).contiguous()
        cmd = command_actions.to(dtype=torch.float32).contiguous()
        n = st.shape[0]
        if out is None:
            thrust = torch.empty(n, dtype=torch.float32, device=st.device)
            torque = torch.empty(n, 3, dtype=torch.float32, device=st.device)
        else:
            thrust, torque = out
        check(lib.ozl_lee_control(self.mode, n, st.data_ptr(), cmd.data_ptr(), self._gains, thrust.data_ptr(),
                                  torque.data_ptr(), torch.cuda.current_stream().cuda_stream), ValueError)
        return thrust, torque

    def wrench(self, robot_state, command_actions, thrust_scale, out=None):
        """(thrust_scale * thrust, torque) packed as [N,4] -- the body wrench of tasks/lee_landed.py:313-314."""
        st = robot_state.to(dtype=torch.float32).contiguous()
        cmd = command_actions.to(dtype=torch.float32).contiguous()
        n = st.shape[0]
        if out is None:
            out = torch.empty(n, 4, dtype=torch.float32, device=st.device)
        check(lib.ozl_lee_wrench(self.mode, n, st.data_ptr(), cmd.data_ptr(), self._gains, float(thrust_scale),
                                 out.data_ptr(), torch.cuda.current_stream().cuda_stream), ValueError)
        return out
