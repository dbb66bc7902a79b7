"""Stand-in `isaacgym` package.  TEST INFRASTRUCTURE ONLY.

Isaac Gym Preview 4 is a closed, un-vendored third-party binary.  This stub exists so that the reference's
OWN Python (`envs/t1.py`, `utils/terrain.py`, `envs/base_task.py`) can be imported, read-only from /root/reference, and
its observation / reward / termination / reset methods executed on hand-filled tensors as the parity oracle
(tools/make_golden.py).  Only `torch_utils` has real bodies (restated from the public Isaac Gym semantics);
`gymapi`, `gymtorch`, `gymutil` and `terrain_utils` carry the names the imports need.
"""
from . import gymapi, gymtorch, gymutil, terrain_utils, torch_utils  # noqa: F401
