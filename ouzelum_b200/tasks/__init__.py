"""Task registry (mirror of `isaacgym_task_map`, isaacgymenvs/tasks/__init__.py:58-85, quadcopter family only)."""
from .landing import Landed, Landing, Lando
from .lee_landed import LeeLanded
from .ouzelum import Ouzelum

task_map = {
    "Ouzelum": Ouzelum,
    "Lando": Lando,
    "Landing": Landing,
    "Landed": Landed,
    "LeeLanded": LeeLanded,
}
