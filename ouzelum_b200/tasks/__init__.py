"""Task registry (mirror of `isaacgym_task_map`, isaacgymenvs/tasks/__init__.py:58-85, quadcopter family only)."""
from .ekf_lee_landed import EKFLeeLanded
from .landing import Landed, Landing, Lando
from .lee_landed import LeeLanded
from .ouzelum import Ouzelum
from .quadcopter import Quadcopter

task_map = {
    "Ouzelum": Ouzelum,
    "Lando": Lando,
    "Landing": Landing,
    "Landed": Landed,
    "LeeLanded": LeeLanded,
    "EKFLeeLanded": EKFLeeLanded,
    "Quadcopter": Quadcopter,
}
