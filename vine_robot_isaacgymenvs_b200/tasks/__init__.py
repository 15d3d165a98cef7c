"""Task registry (reference: isaacgymenvs/tasks/__init__.py:53-74); only the vine path exists here."""
from .vine5link_moving_base import Vine5LinkMovingBase

isaacgym_task_map = {
    "Vine5LinkMovingBase": Vine5LinkMovingBase,
}
