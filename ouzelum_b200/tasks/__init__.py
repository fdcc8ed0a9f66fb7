"""Task registry (mirror of `isaacgym_task_map`, isaacgymenvs/tasks/__init__.py:58-85, x500 family only)."""
from .ouzelum import Ouzelum

task_map = {
    "Ouzelum": Ouzelum,
}
