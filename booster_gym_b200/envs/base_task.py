"""BaseTask: device / sim-parameter plumbing of a task (mirror of envs/base_task.py:7-140).

The reference acquires Isaac Gym, creates a PhysX sim and an optional viewer here.  In this build the "sim" is the
CUDA library behind include/b200_t1.h, so what remains is: pick the device, validate the `sim` section with the
reference's error behaviour, build the Terrain, and keep the attribute surface (`device, headless, up_axis_idx, dt`
source values, `enable_viewer_sync, viewer, camera_frames`) that Runner / play code touches.  There is no renderer on
a B200 box: `render()` is a no-op and `camera_frames` stays empty.
"""
import torch

from .. import _lib
from ..utils.terrain import Terrain


def parse_device_str(device):
    """gymutil.parse_device_str semantics: 'cuda:1' -> ('cuda', 1), 'cpu' -> ('cpu', 0)"""
    s = str(device).lower()
    if s == "cpu":
        return "cpu", 0
    if s.startswith("cuda"):
        parts = s.split(":")
        return "cuda", int(parts[1]) if len(parts) > 1 else 0
    raise ValueError(f"Invalid device string: {device}")


class BaseTask:
    def __init__(self, cfg):
        self.cfg = cfg
        self.gym = None
        self.create_sim()
        self.terrain = Terrain(self.gym, self.sim, self.device, self.cfg["terrain"])
        self.set_viewer()

    def create_sim(self):
        sim_cfg = self.cfg["sim"]
        sim_device = self.cfg["basic"]["sim_device"]
        sim_device_type, self.sim_device_id = parse_device_str(sim_device)
        if sim_device_type != "cuda":
            # the reference falls back to CPU PhysX here (envs/base_task.py:27-30); this build has no CPU path
            raise _lib.B200Error("booster_gym_b200 runs the simulation on a CUDA device only (sim_device must be 'cuda:N')")
        self.device = sim_device
        self.headless = self.cfg["basic"]["headless"]
        self.graphics_device_id = -1
        if sim_cfg["up_axis"] == "z":
            self.up_axis_idx = 2
        elif sim_cfg["up_axis"] == "y":
            raise NotImplementedError("up_axis 'y' is not supported: the T1 rigid-body model is z-up")
        else:
            raise ValueError(f"Invalid physics up-axis: {sim_cfg['up_axis']}")
        if sim_cfg["physics_engine"] not in ("physx", "flex"):
            raise ValueError(f"Invalid physics engine backend: {sim_cfg['physics_engine']}")
        self.physics_engine = "b200"  # the PhysX/Flex solver options of the YAML have no meaning for this engine
        self.sim = None
        _lib.load()
        torch.cuda.set_device(self.sim_device_id)

    def set_viewer(self):
        self.enable_viewer_sync = True
        self.viewer = None
        self.camera_frames = []

    def render(self):
        return
