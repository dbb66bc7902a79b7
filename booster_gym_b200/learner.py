"""Learner: host-side owner of the PPO learner state (flat parameter / gradient / Adam buffers, workspace) and thin
wrappers over the learner half of the C-ABI (include/b200_t1.h).  torch allocates every buffer; the library only
launches kernels on torch's current stream.  Replaces what utils/runner.py does with ActorCritic + torch.optim.Adam
(utils/runner.py:32-33,109-111,123-185).
"""
import ctypes as C

import torch

from . import _abi, _lib, config


class Learner:
    def __init__(self, cfg, num_envs, device, world_size=1, env_base=0, learning_rate=None, seed=0):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.B200Error("the learner kernels run on a CUDA device only")
        self.cfg = cfg
        self.horizon = int(cfg["runner"]["horizon_length"])
        self.num_envs = int(num_envs)
        self.world_size = int(world_size)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        n = self._lib.b200_ppo_num_params()
        dev = self.device
        self.params = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.scalars = torch.zeros(_abi.SC["COUNT"], dtype=torch.float32, device=dev)
        self.dstats = torch.zeros(_abi.DS["COUNT"], dtype=torch.float64, device=dev)
        ws_bytes = self._lib.b200_ppo_workspace_bytes(self.horizon, self.num_envs)
        # the library needs a 1 KiB-aligned workspace (TMA / 128-byte swizzle); the caching allocator only promises 512 B
        self._workspace_raw = torch.empty(ws_bytes // 4 + 256, dtype=torch.float32, device=dev)
        off = (-self._workspace_raw.data_ptr() % 1024) // 4
        self.workspace = self._workspace_raw[off:off + ws_bytes // 4]
        M = self.horizon * self.num_envs
        self.old_mu = torch.empty(M, 12, dtype=torch.float32, device=dev)
        self.old_logp = torch.empty(M, dtype=torch.float32, device=dev)
        self.table = _lib.param_table()
        self._c_cfg = config.ppo_config(cfg, self.num_envs, world_size, env_base)
        h = C.c_void_p()
        _lib.check(self._lib.b200_ppo_create(C.byref(self._c_cfg), self.params.data_ptr(), self.grads.data_ptr(),
                                             self.adam_m.data_ptr(), self.adam_v.data_ptr(), self.scalars.data_ptr(),
                                             self.dstats.data_ptr(), self.workspace.data_ptr(), self.device.index or 0,
                                             C.byref(h)), "b200_ppo_create")
        self._h = h
        self.set_lr(cfg["algorithm"]["learning_rate"] if learning_rate is None else learning_rate)
        self.act_step = 0
        self.peers_bound = False
        self._peer_buf = self._peer_hdl = None

    def bind_peers(self, group=None):
        """Multi-GPU: exchange advantage moments / gradients / loss sums over NVLink peer memory inside the library's own kernels
        (b200_ppo_bind_peers) instead of three NCCL all-reduces per epoch.  Needs torch.distributed (NCCL) initialised and a
        symmetric-memory allocation peer-mapped across the node's ranks.  Raises if the buffers cannot be bound (no silent NCCL
        fall-back: B200_PEER_EXCHANGE=0 selects that protocol explicitly); returns False only for single-process runs."""
        import torch.distributed as dist

        if self.world_size <= 1 or not dist.is_initialized():
            return False
        try:
            import torch.distributed._symmetric_memory as symm_mem

            group = group or dist.group.WORLD
            nfloat = self._lib.b200_ppo_peer_buffer_bytes() // 4
            buf = symm_mem.empty(nfloat, dtype=torch.float32, device=self.device)
            buf.zero_()
            torch.cuda.synchronize(self.device)
            hdl = symm_mem.rendezvous(buf, group)
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            if len(ptrs) != self.world_size:
                raise RuntimeError(f"symmetric memory spans {len(ptrs)} ranks, the learner {self.world_size}")
            arr = (C.c_ulonglong * len(ptrs))(*ptrs)
            dist.barrier(group)            # every rank's flags are zero before anyone can post
            _lib.check(self._lib.b200_ppo_bind_peers(self._h, arr, int(hdl.rank), len(ptrs)), "b200_ppo_bind_peers")
            self._peer_buf, self._peer_hdl = buf, hdl
            self.peers_bound = True
        except Exception as ex:  # noqa: BLE001
            # A silent fall-back would make a scaling run unable to say which data plane it measured: failing to bind is an
            # error.  The NCCL protocol of Runner.update is selected explicitly with B200_PEER_EXCHANGE=0.
            raise _lib.B200Error(f"peer-memory gradient exchange could not be bound ({type(ex).__name__}: {ex}); "
                                 f"set B200_PEER_EXCHANGE=0 to run the NCCL all-reduce protocol instead") from ex
        # all ranks must agree, otherwise the lockstep protocol deadlocks
        flag = torch.tensor([1 if self.peers_bound else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() == 0 and self.peers_bound:
            raise _lib.B200Error("peer-memory exchange bound on this rank but not on all ranks")
        return self.peers_bound

    # ---- parameter plumbing ---------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def views(self, flat=None):
        """{state_dict name: view into the flat buffer with the reference's shape} (utils/model.py parameter names)"""
        flat = self.params if flat is None else flat
        out = {}
        for name, off, rows, cols in self.table:
            v = flat[off:off + rows * cols]
            out[name] = v.view(rows, cols) if (name.endswith("weight") or name == "logstd") else v.view(rows)
        return out

    def load_state_dict(self, sd):
        views = self.views()
        for name, v in views.items():
            v.copy_(sd[name].to(device=self.device, dtype=torch.float32).reshape(v.shape))

    def state_dict(self):
        return {k: v.clone() for k, v in self.views().items()}

    def set_lr(self, lr):
        self.scalars[_abi.SC["LR"]] = float(lr)

    # ---- kernels -------------------------------------------------------------------------------------------------------
    def act(self, obs, actions_out, mu_out=None, eps=None, deterministic=False, step=None, reuse_weights=False):
        """reuse_weights: the caller vouches that the parameters have not changed since its previous act() call (B200_ACT_REUSE_WEIGHTS:
        the tensor-core path of >= 2048 rows then skips rebuilding its weight operands); the default rebuilds them on every call."""
        if step is None:
            step = 0xFFFFFFFFFFFFFFFF  # B200_STEP_AUTO: device-side counter, so captured graphs draw fresh noise on replay
        flags = (1 if deterministic else 0) | (2 if reuse_weights else 0)
        _lib.check(self._lib.b200_policy_act(self._h, obs.data_ptr(), obs.shape[0], actions_out.data_ptr(),
                                             mu_out.data_ptr() if mu_out is not None else None,
                                             eps.data_ptr() if eps is not None else None, self.seed, int(step),
                                             flags, self._stream()), "b200_policy_act")
        return actions_out

    def value(self, obs, priv, out=None):
        out = torch.empty(obs.shape[0], dtype=torch.float32, device=self.device) if out is None else out
        _lib.check(self._lib.b200_critic_value(self._h, obs.data_ptr(), priv.data_ptr(), obs.shape[0], out.data_ptr(),
                                               self._stream()), "b200_critic_value")
        return out

    def old_dist(self, obses, privs, actions):
        _lib.check(self._lib.b200_ppo_old_dist(self._h, obses.data_ptr(), privs.data_ptr(), actions.data_ptr(),
                                               self.old_mu.data_ptr(), self.old_logp.data_ptr(), self._stream()), "b200_ppo_old_dist")

    def epoch_a(self, rewards, dones_u8, time_outs_u8, last_obs, last_priv):
        _lib.check(self._lib.b200_ppo_epoch_a(self._h, rewards.data_ptr(), dones_u8.data_ptr(), time_outs_u8.data_ptr(),
                                              last_obs.data_ptr(), last_priv.data_ptr(), self._stream()), "b200_ppo_epoch_a")

    def epoch_b(self, actions):
        _lib.check(self._lib.b200_ppo_epoch_b(self._h, actions.data_ptr(), self.old_mu.data_ptr(), self.old_logp.data_ptr(),
                                              self._stream()), "b200_ppo_epoch_b")

    def apply(self):
        _lib.check(self._lib.b200_ppo_apply(self._h, self._stream()), "b200_ppo_apply")

    def buffer(self, which, shape):
        """test helper: copy of a workspace view (0 values, 1 adv, 2 returns, 3 mu, 4 last values, 5 dV, 6 dmu)"""
        ptr = self._lib.b200_ppo_buffer(self._h, which)
        base = self.workspace.data_ptr()
        off = (ptr - base) // 4
        n = 1
        for s in shape:
            n *= s
        return self.workspace[off:off + n].view(*shape).clone()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200_ppo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gae(rewards, dones_u8, time_outs_u8, values, last_values, gamma, lam, stats=None):
    """standalone GAE kernel (utils/utils.py:33-44 + the in-place time-out bootstrap of utils/runner.py:135)"""
    lib = _lib.load()
    T, N = rewards.shape
    adv = torch.empty_like(rewards)
    ret = torch.empty_like(rewards)
    _lib.check(lib.b200_gae(rewards.data_ptr(), dones_u8.data_ptr(), time_outs_u8.data_ptr(), values.data_ptr(),
                            last_values.data_ptr(), float(gamma), float(lam), adv.data_ptr(), ret.data_ptr(),
                            stats.data_ptr() if stats is not None else None, T, N,
                            torch.cuda.current_stream(rewards.device).cuda_stream), "b200_gae")
    return adv, ret
