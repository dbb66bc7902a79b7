"""Sub-terrain generators (restated public semantics of Isaac Gym's `terrain_utils`): plain NumPy, no
package-relative imports, so oracle/shims/isaacgym/terrain_utils.py can load this file by path."""
import numpy as np


class SubTerrain:
    def __init__(self, terrain_name="terrain", width=256, length=256, vertical_scale=1.0, horizontal_scale=1.0):
        self.terrain_name = terrain_name
        self.vertical_scale = vertical_scale
        self.horizontal_scale = horizontal_scale
        self.width = width
        self.length = length
        self.height_field_raw = np.zeros((self.width, self.length), dtype=np.int16)


def _bilinear_resample(coarse, nx, ny):
    """linear interpolation of a coarse grid (spanning the same extent) onto nx x ny points"""
    cx, cy = coarse.shape
    fx = np.linspace(0.0, cx - 1.0, nx)
    fy = np.linspace(0.0, cy - 1.0, ny)
    x0 = np.clip(np.floor(fx).astype(int), 0, cx - 2) if cx > 1 else np.zeros(nx, int)
    y0 = np.clip(np.floor(fy).astype(int), 0, cy - 2) if cy > 1 else np.zeros(ny, int)
    tx = (fx - x0)[:, None]
    ty = (fy - y0)[None, :]
    x1 = np.minimum(x0 + 1, cx - 1)
    y1 = np.minimum(y0 + 1, cy - 1)
    c = coarse.astype(np.float64)
    return ((1 - tx) * (1 - ty) * c[np.ix_(x0, y0)] + tx * (1 - ty) * c[np.ix_(x1, y0)]
            + (1 - tx) * ty * c[np.ix_(x0, y1)] + tx * ty * c[np.ix_(x1, y1)])


def random_uniform_terrain(terrain, min_height, max_height, step=1, downsampled_scale=None):
    """discrete random heights on a coarse grid, bilinearly up-sampled, rounded, added as int16"""
    if downsampled_scale is None:
        downsampled_scale = terrain.horizontal_scale
    min_h = int(min_height / terrain.vertical_scale)
    max_h = int(max_height / terrain.vertical_scale)
    stp = int(step / terrain.vertical_scale)
    heights_range = np.arange(min_h, max_h + stp, stp)
    shape = (int(terrain.width * terrain.horizontal_scale / downsampled_scale),
             int(terrain.length * terrain.horizontal_scale / downsampled_scale))
    coarse = np.random.choice(heights_range, shape)
    up = np.rint(_bilinear_resample(coarse, terrain.width, terrain.length))
    terrain.height_field_raw += up.astype(np.int16)
    return terrain


def pyramid_sloped_terrain(terrain, slope=1, platform_size=1.0):
    x = np.arange(0, terrain.width)
    y = np.arange(0, terrain.length)
    center_x = int(terrain.width / 2)
    center_y = int(terrain.length / 2)
    xx = ((center_x - np.abs(center_x - x)) / center_x).reshape(terrain.width, 1)
    yy = ((center_y - np.abs(center_y - y)) / center_y).reshape(1, terrain.length)
    max_height = int(slope * (terrain.horizontal_scale / terrain.vertical_scale) * (terrain.width / 2))
    terrain.height_field_raw += (max_height * xx * yy).astype(terrain.height_field_raw.dtype)
    platform_size = int(platform_size / terrain.horizontal_scale / 2)
    x1 = terrain.width // 2 - platform_size
    x2 = terrain.width // 2 + platform_size
    y1 = terrain.length // 2 - platform_size
    y2 = terrain.length // 2 + platform_size
    min_h = min(terrain.height_field_raw[x1, y1], 0)
    max_h = max(terrain.height_field_raw[x1, y1], 0)
    terrain.height_field_raw = np.clip(terrain.height_field_raw, min_h, max_h)
    return terrain


def discrete_obstacles_terrain(terrain, max_height, min_size, max_size, num_rects, platform_size=1.0):
    max_height = int(max_height / terrain.vertical_scale)
    min_size = int(min_size / terrain.horizontal_scale)
    max_size = int(max_size / terrain.horizontal_scale)
    platform_size = int(platform_size / terrain.horizontal_scale)
    (i, j) = terrain.height_field_raw.shape
    height_range = [-max_height, -max_height // 2, max_height // 2, max_height]
    width_range = range(min_size, max_size, 4)
    length_range = range(min_size, max_size, 4)
    for _ in range(num_rects):
        width = np.random.choice(width_range)
        length = np.random.choice(length_range)
        start_i = np.random.choice(range(0, i - width, 4))
        start_j = np.random.choice(range(0, j - length, 4))
        terrain.height_field_raw[start_i:start_i + width, start_j:start_j + length] = np.random.choice(height_range)
    x1 = (terrain.width - platform_size) // 2
    x2 = (terrain.width + platform_size) // 2
    y1 = (terrain.length - platform_size) // 2
    y2 = (terrain.length + platform_size) // 2
    terrain.height_field_raw[x1:x2, y1:y2] = 0
    return terrain
