"""Terrain: plane or int16 heightfield, and the bilinear height lookup (mirror of utils/terrain.py:6-121).

Same constructor signature and attributes as the reference class (`type, env_width, env_length, border_size,
border_pixels, horizontal_scale, vertical_scale, height_field_raw`), but nothing is uploaded to a physics engine:
the int16 array itself is what the CUDA contact and lookup kernels read (b200_t1_create uploads it once).
`terrain_heights()` runs on the device through the C-ABI (b200_terrain_heights) - no `.cpu().numpy()` round trip.

The sub-terrain generators restate the published semantics of Isaac Gym's `terrain_utils` (third-party, not in the
reference tree).  Bit-identical generation is not a contract (the reference relies on scipy's removed
interp2d); the LOOKUP on a given array is.
"""
import numpy as np
import torch

from .. import _lib


from .terrain_gen import (SubTerrain, discrete_obstacles_terrain, pyramid_sloped_terrain,  # noqa: E402,F401
                          random_uniform_terrain)


class Terrain:

    def __init__(self, gym, sim, device, terrain_cfg):
        self.terrain_cfg = terrain_cfg
        self.gym = gym      # unused: kept for signature compatibility (utils/terrain.py:8)
        self.sim = sim
        self.device = device
        self.type = self.terrain_cfg["type"]
        self.height_field_raw = None
        self._handle = None  # B200T1Handle* owning the device copy of the heightfield (set by the task)

        if self.type == "plane":
            self._create_ground_plane()
        elif self.type == "trimesh":
            self._create_trimesh()
        else:
            raise ValueError(f"Invalid terrain type: {self.type}")

    def _create_ground_plane(self):
        # friction / restitution of the plane are read from terrain_cfg by the contact model (config.t1_config)
        self.static_friction = self.terrain_cfg["static_friction"]

    def _create_trimesh(self):
        tc = self.terrain_cfg
        self.env_width = tc["num_terrains"] * tc["terrain_width"]
        self.env_length = tc["terrain_length"]
        self.border_size = tc["border_size"]
        self.horizontal_scale = tc["horizontal_scale"]
        self.vertical_scale = tc["vertical_scale"]
        self.border_pixels = int(self.border_size / self.horizontal_scale)
        wpx = int(tc["terrain_width"] / self.horizontal_scale)
        lpx = int(tc["terrain_length"] / self.horizontal_scale)
        self.height_field_raw = np.zeros((tc["num_terrains"] * wpx + 2 * self.border_pixels, lpx + 2 * self.border_pixels),
                                         dtype=np.int16)
        cum = np.cumsum(tc["terrain_proportions"]) / np.sum(tc["terrain_proportions"]) * tc["num_terrains"]
        for i in range(tc["num_terrains"]):
            sub = SubTerrain("terrain", width=wpx, length=lpx, vertical_scale=self.vertical_scale,
                             horizontal_scale=self.horizontal_scale)
            if i < cum[0]:
                pass  # flat
            elif i < cum[1]:
                pyramid_sloped_terrain(sub, slope=tc["slope"], platform_size=3.0)
            elif i < cum[2]:
                random_uniform_terrain(sub, min_height=-0.5 * tc["random_height"], max_height=0.5 * tc["random_height"],
                                       step=0.005, downsampled_scale=0.2)
            else:
                discrete_obstacles_terrain(sub, max_height=tc["discrete_height"], min_size=1.0, max_size=2.0,
                                           num_rects=20, platform_size=3.0)
            x0 = self.border_pixels + i * wpx
            self.height_field_raw[x0:x0 + wpx, self.border_pixels:self.border_pixels + lpx] = sub.height_field_raw

    def terrain_heights(self, base_pos):
        """heights under the first two columns of `base_pos` ([K, >=2] float32 CUDA tensor) -> float32 [K]"""
        if self.type == "plane":
            return torch.zeros(len(base_pos), dtype=torch.float, device=self.device)
        if self._handle is None:
            raise _lib.B200Error("Terrain.terrain_heights: the device heightfield is not bound to a task yet")
        if not base_pos.is_cuda:
            raise _lib.B200Error("Terrain.terrain_heights needs a CUDA tensor (there is no CPU path)")
        pos = base_pos if base_pos.dtype == torch.float32 else base_pos.float()
        if pos.stride(-1) != 1:
            pos = pos.contiguous()
        out = torch.empty(pos.shape[0], dtype=torch.float32, device=pos.device)
        _lib.check(_lib.load().b200_terrain_heights(self._handle, pos.data_ptr(), pos.stride(0), pos.shape[0], out.data_ptr(),
                                                    torch.cuda.current_stream(pos.device).cuda_stream), "terrain_heights")
        return out
