"""Lee geometric controllers (mirror of isaacgymenvs/controllers/)."""
from .control_config import control
from .controller import Controller, control_class_dict

__all__ = ["control", "Controller", "control_class_dict"]
