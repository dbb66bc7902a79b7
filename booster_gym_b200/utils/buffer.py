"""ExperienceBuffer: rollout storage behind the interface of utils/buffer.py:4-25.

Time-major `[horizon, num_envs, *shape]` device tensors keyed by name, zero-initialised.  Beyond the reference's five
methods it exposes what the CUDA learner needs without copies: `raw(name)` returns the storage reinterpreted for the
C-ABI (bool buffers as uint8), and `row(name, t)` the contiguous `[num_envs, *shape]` slice a kernel can write into.
"""
import torch


class ExperienceBuffer:
    __slots__ = ("tensor_dict", "horizon_length", "num_envs", "device")

    def __init__(self, horizon_length, num_envs, device):
        self.horizon_length, self.num_envs, self.device = int(horizon_length), int(num_envs), device
        self.tensor_dict = {}

    # ---- reference interface -----------------------------------------------------------------------------------------
    def add_buffer(self, name, shape, dtype=None):
        full_shape = (self.horizon_length, self.num_envs) + tuple(shape)
        self.tensor_dict[name] = torch.zeros(full_shape, dtype=dtype, device=self.device)

    def update_data(self, name, idx, data):
        self.tensor_dict[name][idx].copy_(data)

    def keys(self):
        return self.tensor_dict.keys()

    def __getitem__(self, buf_name):
        return self.tensor_dict[buf_name]

    def __len__(self):
        return len(self.tensor_dict)

    # ---- zero-copy accessors for the kernels ---------------------------------------------------------------------------
    def raw(self, name):
        t = self.tensor_dict[name]
        return t.view(torch.uint8) if t.dtype == torch.bool else t

    def row(self, name, t):
        return self.raw(name)[t]
