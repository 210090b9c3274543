"""TD3 gradient steps on the device behind the reference's ``TD3.train`` surface.

``FusedTD3Update`` owns five flat float32 device blocks (``params, targets, grads, adam_m, adam_v``, layout from
``cstr_td3_layout``) and runs one iteration of the loop body of ``TD3.train`` (``core/td3/td3.py:162-206``) per
``update()`` call through ``cstr_td3_update`` — target smoothing, twin-min target, critic forward/MSE/backward, Adam,
delayed actor update, polyak — as hand-written float32 CUDA kernels.  No torch autograd, no cuBLAS.

``gemm="fp32"`` (default) keeps the reference's float32 FMA arithmetic; ``gemm="tensor"`` runs the three hidden-layer GEMM roles on
tcgen05 with every fp32 operand split into three bf16 planes (fp32-grade accuracy, see csrc/cstr_td3_tc.cuh); ``gemm="bf16"`` uses plain bf16
operands with fp32 accumulation (reduced precision, throughput mode; no per-step parity bar).

``adopt_policy`` re-points the parameters of the reference's ``TD3Policy`` modules at views of the flat blocks, so
``policy.predict``, ``model.save`` and the fused rollout keep seeing the live weights without copies.
``bind_td3_class(TD3)`` returns a subclass of the reference algorithm whose ``train()`` runs here.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int64
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _lib

NET_NAMES = ("actor", "critic0", "critic1")
_GEMM_MODES = {"fp32": 0, "tensor": 1, "bf16": 2}
_TENSORS = ("W1", "b1", "W2", "b2", "W3", "b3")


class FusedTD3Update:
    def __init__(self, net_arch: Sequence[int] = (400, 300), batch_size: int = 256, device: Any = "cuda", gamma: float = 0.99, tau: float = 0.005,
                 learning_rate: float = 1e-3, policy_delay: int = 2, target_policy_noise: float = 0.2, target_noise_clip: float = 0.5,
                 betas=(0.9, 0.999), eps: float = 1e-8, seed: int = 0, gemm: str = "fp32", n_critics: int = 2, dp_rank: int = 0):
        torch = _lib.require_cuda()
        self._torch = torch
        self._libc = _lib.load()
        if len(net_arch) != 2:
            raise ValueError("FusedTD3Update supports the two-hidden-layer MLPs of TD3's MlpPolicy (net_arch=[h1, h2])")
        self.h1, self.h2 = int(net_arch[0]), int(net_arch[1])
        if self.h1 % 4 or self.h2 % 4:
            raise ValueError("hidden sizes must be multiples of 4")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CstrLibraryError("FusedTD3Update needs a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.gamma, self.tau, self.learning_rate = float(gamma), float(tau), float(learning_rate)
        self.policy_delay, self.target_policy_noise, self.target_noise_clip = int(policy_delay), float(target_policy_noise), float(target_noise_clip)
        self.betas, self.eps, self.seed = (float(betas[0]), float(betas[1])), float(eps), int(seed)
        # data-parallel training: every rank holds the same `seed`, and the in-kernel noise is keyed by (seed, batch row, update), so
        # without the rank in the key row b of every shard would draw the same smoothing noise / eps (perfectly correlated global batch)
        self.dp_rank = int(dp_rank)
        if gemm not in _GEMM_MODES:
            raise ValueError("gemm must be 'fp32' (FFMA tiles, the reference's arithmetic), 'tensor' (tcgen05 bf16x3 split, fp32-grade) or "
                             "'bf16' (tcgen05, plain bf16 operands: reduced precision)")
        self.gemm = gemm
        if n_critics not in (1, 2):
            raise ValueError("n_critics must be 2 (TD3) or 1 (DDPG)")
        self.n_critics = int(n_critics)
        offs = (c_int64 * 19)()
        _lib.check(self._libc.cstr_td3_layout(self.h1, self.h2, offs), "cstr_td3_layout")
        self.param_count = int(offs[18])
        self._offsets = {(NET_NAMES[n], _TENSORS[k]): int(offs[n * 6 + k]) for n in range(3) for k in range(6)}
        self.actor_range = (0, self._offsets[("critic0", "W1")])
        self.critic_range = (self._offsets[("critic0", "W1")], self.param_count)
        with torch.cuda.device(self.device):
            z = lambda: torch.zeros(self.param_count, dtype=torch.float32, device=self.device)  # noqa: E731
            self.params, self.targets, self.grads, self.adam_m, self.adam_v = z(), z(), z(), z(), z()
            self.loss_sums = torch.zeros(4, dtype=torch.float32, device=self.device)
            self._counters = torch.zeros(4, dtype=torch.int64, device=self.device)  # graph mode: n_updates, critic_step, actor_step, sample_draw
        self._graph = None
        self._graph_key = None
        self._graph_out = None
        self._peer = None  # cstr_peer_comm of enable_peer_allreduce()
        self._workspace = None
        self._batch = 0
        self._set_batch(int(batch_size))
        self.n_updates = 0
        self.critic_step = 0
        self.actor_step = 0
        self.launches = 0

    # ---- layout ---------------------------------------------------------------------------------------------------
    def _shape(self, net: str, tensor: str):
        i, o = (4, 2) if net == "actor" else (6, 1)
        return {"W1": (self.h1, i), "b1": (self.h1,), "W2": (self.h2, self.h1), "b2": (self.h2,), "W3": (o, self.h2), "b3": (o,)}[tensor]

    def views(self, block: str = "params") -> Dict[str, List[Any]]:
        """``{net: [W1, b1, W2, b2, W3, b3]}`` as views of one flat block (``params|targets|grads|adam_m|adam_v``)."""
        flat = getattr(self, block)
        out = {}
        for net in NET_NAMES:
            out[net] = []
            for t in _TENSORS:
                shape = self._shape(net, t)
                off = self._offsets[(net, t)]
                out[net].append(flat[off:off + int(np.prod(shape))].view(shape))
        return out

    def _set_batch(self, batch: int) -> None:
        if batch == self._batch:
            return
        need = self._workspace_bytes(batch)
        if need < 0:
            msg = self._libc.cstr_last_error()
            raise ValueError(msg.decode() if msg else "bad configuration")
        with self._torch.cuda.device(self.device):
            self._workspace = self._torch.empty(need // 4, dtype=self._torch.float32, device=self.device)
        self._batch = batch

    def _workspace_bytes(self, batch: int) -> int:
        return int(self._libc.cstr_td3_workspace_bytes(byref(self._config(batch))))

    def _keyed_seed(self) -> int:
        return (self.seed + 0x9E3779B97F4A7C15 * self.dp_rank) & (2**64 - 1)

    def _config(self, batch: int) -> "_lib.Td3Config":
        return _lib.Td3Config(h1=self.h1, h2=self.h2, batch=batch, policy_delay=self.policy_delay, gamma=self.gamma, tau=self.tau, lr=self.learning_rate,
                              beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, target_policy_noise=self.target_policy_noise,
                              target_noise_clip=self.target_noise_clip, seed=self._keyed_seed(), gemm_mode=_GEMM_MODES[self.gemm],
                              n_critics=getattr(self, "n_critics", 2))

    # ---- weights in / out -------------------------------------------------------------------------------------------
    def load_nets(self, nets: Dict[str, Sequence[Any]]) -> None:
        """``nets`` maps ``actor, critic0, critic1`` (and optionally ``*_target``) to six arrays/tensors in nn.Linear layout."""
        torch = self._torch
        for block, suffix in (("params", ""), ("targets", "_target")):
            v = self.views(block)
            for net in NET_NAMES[:1 + self.n_critics]:
                src = nets.get(net + suffix, nets[net] if suffix else None)
                for dst, s in zip(v[net], src):
                    dst.copy_(torch.as_tensor(np.asarray(s) if not isinstance(s, torch.Tensor) else s).to(self.device, torch.float32).reshape(dst.shape))

    def nets(self) -> Dict[str, List[np.ndarray]]:
        out = {}
        for block, suffix in (("params", ""), ("targets", "_target")):
            for net, ts in self.views(block).items():
                out[net + suffix] = [t.cpu().numpy() for t in ts]
        return out

    def adopt_modules(self, actor, critics: Sequence[Any], actor_target, critic_targets: Sequence[Any]) -> None:
        """Copy the weights of ``nn.Sequential(Linear, ReLU, Linear, ReLU, Linear[, Tanh])`` modules in and re-point their
        parameters at the flat blocks (zero-copy sharing from then on)."""
        if len(critics) != self.n_critics or len(critic_targets) != self.n_critics:
            raise ValueError(f"expected {self.n_critics} critic network(s)")
        pairs = [("params", "actor", actor), ("targets", "actor", actor_target)]
        for z in range(self.n_critics):
            pairs += [("params", f"critic{z}", critics[z]), ("targets", f"critic{z}", critic_targets[z])]
        for block, net, module in pairs:
            ps = list(module.parameters())
            if len(ps) != 6:
                raise ValueError("expected a 3-layer MLP (6 parameter tensors)")
            for view, p in zip(self.views(block)[net], ps):
                if tuple(p.shape) != tuple(view.shape):
                    raise ValueError(f"shape mismatch for {net}: module {tuple(p.shape)} vs layout {tuple(view.shape)}")
                view.copy_(p.data.to(self.device, self._torch.float32))
                p.data = view

    def adopt_policy(self, policy) -> None:
        """The reference's ``TD3Policy`` (core/td3/policies.py:172-210): ``actor.mu``, ``critic.q_networks`` and the targets."""
        if len(policy.critic.q_networks) != self.n_critics:
            raise ValueError(f"this engine was built for n_critics={self.n_critics}, the policy has {len(policy.critic.q_networks)}")
        self.adopt_modules(policy.actor.mu, list(policy.critic.q_networks), policy.actor_target.mu, list(policy.critic_target.q_networks))
        self._opt_params = {"actor": list(policy.actor.mu.parameters()), "critic": [p for q in policy.critic.q_networks for p in q.parameters()]}

    def import_optimizer_state(self, actor_optimizer, critic_optimizer) -> None:
        """Take over Adam moments / step counts of ``torch.optim.Adam`` optimisers created over the adopted modules."""
        for tag, opt, nets in (("actor", actor_optimizer, ("actor",)), ("critic", critic_optimizer, ("critic0", "critic1")[:self.n_critics])):
            group = opt.param_groups[0]
            self.betas, self.eps = (float(group["betas"][0]), float(group["betas"][1])), float(group["eps"])
            m = [t for n in nets for t in self.views("adam_m")[n]]
            v = [t for n in nets for t in self.views("adam_v")[n]]
            step = 0
            for p, mv, vv in zip(group["params"], m, v):
                st = opt.state.get(p)
                if st:
                    mv.copy_(st["exp_avg"])
                    vv.copy_(st["exp_avg_sq"])
                    step = int(float(st["step"]))
            if tag == "actor":
                self.actor_step = step
            else:
                self.critic_step = step

    def export_optimizer_state(self, actor_optimizer, critic_optimizer) -> None:
        """Write moments / step counts back so ``model.save`` and a later torch ``optimizer.step()`` continue from here."""
        torch = self._torch
        for opt, nets, step in ((actor_optimizer, ("actor",), self.actor_step), (critic_optimizer, ("critic0", "critic1")[:self.n_critics], self.critic_step)):
            if step == 0:
                continue
            m = [t for n in nets for t in self.views("adam_m")[n]]
            v = [t for n in nets for t in self.views("adam_v")[n]]
            for p, mv, vv in zip(opt.param_groups[0]["params"], m, v):
                opt.state[p] = {"step": torch.tensor(float(step)), "exp_avg": mv, "exp_avg_sq": vv}

    # ---- the gradient step ----------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def _f32(self, t, cols):
        torch = self._torch
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.ascontiguousarray(t), device=self.device)
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        if t.numel() != self._batch * cols:
            raise ValueError(f"batch tensor has {t.numel()} elements, expected {self._batch}x{cols}")
        return t

    def _state(self, counters: bool) -> "_lib.Td3State":
        return _lib.Td3State(params=self.params.data_ptr(), targets=self.targets.data_ptr(), grads=self.grads.data_ptr(), adam_m=self.adam_m.data_ptr(),
                             adam_v=self.adam_v.data_ptr(), workspace=self._workspace.data_ptr(), workspace_bytes=self._workspace.numel() * 4,
                             losses=self.loss_sums.data_ptr(), counters=self._counters.data_ptr() if counters else None,
                             peer=ctypes.addressof(self._peer) if self._peer is not None else None)

    # ---- data-parallel training: the gradient all-reduce inside the Adam kernel (include/cstr_b200.h, cstr_peer_comm) ---------------
    def enable_peer_allreduce(self, group=None) -> bool:
        """Move the flat gradient block into memory every rank of ``group`` maps (CUDA IPC over NVLink/NVSwitch, one process per GPU on one
        node) so that the APPLY phases average the gradient over the ranks inside the Adam kernel — no collective launch at all, and the
        update stays capturable as one CUDA graph.  Every rank must then call ``update``/``train`` in lockstep (same number of updates, same
        policy steps).  Returns False (and changes nothing) for a single-process group."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return False
        torch, lib = self._torch, self._libc
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > _lib.PEER_MAX_WORLD:
            raise ValueError(f"peer all-reduce spans one NVSwitch node: world size {world} > {_lib.PEER_MAX_WORLD}")
        grad_bytes = (self.param_count * 4 + 255) // 256 * 256
        total = grad_bytes + int(lib.cstr_peer_flag_bytes())
        base, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _lib.check(lib.cstr_peer_alloc(total, byref(base), handle), "cstr_peer_alloc")
            handles = [None] * world
            dist.all_gather_object(handles, handle.raw, group=group)
            bases = []
            for r in range(world):
                if r == rank:
                    bases.append(base.value)
                else:
                    p = ctypes.c_void_p()
                    _lib.check(lib.cstr_peer_open(handles[r], byref(p)), f"cstr_peer_open(rank {r})")
                    bases.append(p.value)
            comm = _lib.PeerComm(world=world, rank=rank)
            for r in range(world):
                comm.grads[r], comm.flags[r] = bases[r], bases[r] + grad_bytes
            local = torch.as_tensor(_DeviceMemory(base.value, self.param_count), device=self.device)
            local.copy_(self.grads)
            torch.cuda.synchronize(self.device)
        self.grads, self._peer, self._peer_bases, self._peer_group = local, comm, bases, group
        self._graph = None
        dist.barrier(group)  # nobody starts an update before every rank has mapped every block
        return True

    def peer_error(self) -> int:
        """Non-zero after a rank waited in vain for a peer inside the fused all-reduce (the update of that step was skipped)."""
        if self._peer is None:
            return 0
        err = ctypes.c_uint32(0)
        with self._torch.cuda.device(self.device):
            _lib.check(self._libc.cstr_peer_error(byref(self._peer), byref(err), self._stream()), "cstr_peer_error")
        return int(err.value)

    def close_peer_allreduce(self) -> None:
        """Unmap the peers' blocks and free the local one (collective: every rank calls it)."""
        if self._peer is None:
            return
        import torch.distributed as dist

        torch = self._torch
        torch.cuda.synchronize(self.device)
        dist.barrier(self._peer_group)
        keep = torch.empty(self.param_count, dtype=torch.float32, device=self.device)
        keep.copy_(self.grads)
        rank = self._peer.rank
        self.grads, self._graph = keep, None
        torch.cuda.synchronize(self.device)
        for r, b in enumerate(self._peer_bases):
            if r != rank:
                self._libc.cstr_peer_close(b)
        dist.barrier(self._peer_group)  # a block is freed only after every peer has unmapped it
        self._libc.cstr_peer_free(self._peer_bases[rank])
        self._peer = None

    # ---- CUDA-graph path: one captured cycle of policy_delay x (sample + update), replayed ------------------------------------
    def _capture(self, buffer, batch_size: int, env, allreduce=None) -> None:
        torch = self._torch
        self._set_batch(batch_size)
        f32 = dict(dtype=torch.float32, device=self.device)
        out = (torch.empty((batch_size, 4), **f32), torch.empty((batch_size, 2), **f32), torch.empty((batch_size, 4), **f32),
               torch.empty((batch_size, 1), **f32), torch.empty((batch_size, 1), **f32))
        st = self._state(counters=True)
        draw = self._counters[3:4]

        def cycle():
            for k in range(1, self._cycle_len() + 1):  # by-value counters only choose the launch structure: update k of the cycle
                buffer.sample_into(out, draw, env=env)
                self._graph_launch(batch_size, st, out, k, allreduce)

        snapshot = [t.clone() for t in (self.params, self.targets, self.adam_m, self.adam_v, self.loss_sums, self._counters)]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):  # warm-up outside capture (first-launch attribute calls, lazy module loading)
            cycle()
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        # thread_local: a collective inside the cycle keeps NCCL's watchdog thread polling events while this thread captures
        with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
            cycle()
        for dst, src in zip((self.params, self.targets, self.adam_m, self.adam_v, self.loss_sums, self._counters), snapshot):
            dst.copy_(src)  # the warm-up cycle was a real update: undo it
        self._graph, self._graph_out = graph, out
        self._graph_key = self._make_graph_key(buffer, batch_size, env, allreduce)

    def _make_graph_key(self, buffer, batch_size: int, env, allreduce):
        return (id(buffer), batch_size, buffer.size(), buffer.n_envs, id(env), self.policy_delay, self.learning_rate, self.params.data_ptr(),
                self._workspace.data_ptr() if self._workspace is not None else 0, self.grads.data_ptr(), id(allreduce) if allreduce is not None else 0)

    def _cycle_len(self) -> int:
        return self.policy_delay

    def _graph_launch(self, batch_size: int, st, out, k: int, allreduce=None) -> None:
        cfg = self._config(batch_size)

        def run(phases):
            rc = self._libc.cstr_td3_update(byref(cfg), byref(st), _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(out[3]),
                                            _lib.ptr(out[4]), None, k, 1, 1, phases, self._stream())
            _lib.check(rc, "cstr_td3_update (graph capture)")

        if allreduce is None:
            run(_lib.TD3_ALL)
        else:  # the collective is captured between the phases: GRAD -> all-reduce -> APPLY is ONE graph
            run(_lib.TD3_CRITIC_GRAD)
            allreduce(self.grads[self.critic_range[0]:self.critic_range[1]])
            run(_lib.TD3_CRITIC_APPLY | _lib.TD3_ACTOR_GRAD)
            if k % self.policy_delay == 0:
                allreduce(self.grads[self.actor_range[0]:self.actor_range[1]])
            run(_lib.TD3_ACTOR_APPLY)

    def _train_graph(self, gradient_steps: int, buffer, batch_size: int, env, allreduce=None) -> int:
        """Replays whole cycles while the update count is cycle-aligned; returns the number of gradient steps done."""
        if self.n_updates % self._cycle_len() or gradient_steps < self._cycle_len():
            return 0
        key = self._make_graph_key(buffer, batch_size, env, allreduce)
        if self._graph is None or key != self._graph_key:
            self._capture(buffer, batch_size, env, allreduce)
        cycles = gradient_steps // self._cycle_len()
        self._counters.copy_(self._torch.tensor([self.n_updates, self.critic_step, self.actor_step, buffer._draw], dtype=self._torch.int64),
                             non_blocking=True)
        with self._torch.cuda.device(self.device):
            for _ in range(cycles):
                self._graph.replay()
        done = cycles * self._cycle_len()
        self.n_updates += done
        self.critic_step += done
        self.actor_step += done // self.policy_delay
        buffer._draw += done
        self.launches += cycles * (23 * (self.policy_delay - 1) + 44 + self.policy_delay)
        return done

    def update(self, batch, noise=None, allreduce: Optional[Callable[[Any], None]] = None) -> None:
        """One iteration of the loop body (td3.py:162-206) on ``batch`` (``ReplayBufferSamples`` or a 5-tuple in that order).
        ``noise``: explicit N(0, target_policy_noise) draws (B,2) for parity tests; default = Philox in the kernel.
        ``allreduce``: called on the flat gradient range between backward and Adam (data-parallel training)."""
        obs, act, nobs, dones, rew = batch
        self._set_batch(int(obs.shape[0]))
        obs, act, nobs = self._f32(obs, 4), self._f32(act, 2), self._f32(nobs, 4)
        dones, rew = self._f32(dones, 1), self._f32(rew, 1)
        nz = None if noise is None else self._f32(noise, 2)
        self.n_updates += 1
        self.critic_step += 1
        policy_step = self.n_updates % self.policy_delay == 0
        if policy_step:
            self.actor_step += 1
        cfg = self._config(self._batch)
        st = self._state(counters=False)

        def run(phases):
            rc = self._libc.cstr_td3_update(byref(cfg), byref(st), _lib.ptr(obs), _lib.ptr(act), _lib.ptr(nobs), _lib.ptr(dones), _lib.ptr(rew),
                                            _lib.ptr(nz), self.n_updates, self.critic_step, max(self.actor_step, 0), phases, self._stream())
            _lib.check(rc, "cstr_td3_update")

        with self._torch.cuda.device(self.device):
            if allreduce is None or self._peer is not None:  # a peer communicator averages inside the APPLY kernels
                run(_lib.TD3_ALL)
            else:
                run(_lib.TD3_CRITIC_GRAD)
                allreduce(self.grads[self.critic_range[0]:self.critic_range[1]])
                run(_lib.TD3_CRITIC_APPLY | _lib.TD3_ACTOR_GRAD)
                if policy_step:
                    allreduce(self.grads[self.actor_range[0]:self.actor_range[1]])
                run(_lib.TD3_ACTOR_APPLY)
        self.launches += 23 if not policy_step else 44  # without the split-K finish kernels of small batches

    def train(self, gradient_steps: int, buffer, batch_size: Optional[int] = None, env=None, allreduce=None, graph: bool = False) -> None:
        """``TD3.train(gradient_steps, batch_size)`` (td3.py:154-206): sample + update, ``gradient_steps`` times.
        ``graph=True`` (Philox-index buffer): whole cycles of ``policy_delay`` updates are replayed from ONE captured CUDA graph (24-45
        launches per update become one graph launch per cycle); the remainder runs launch by launch.  Data-parallel training is captured
        too when the averaging is part of the Adam kernels (``enable_peer_allreduce``): sample -> GRAD -> mean over ranks -> APPLY is one
        graph on every rank.  An ``allreduce=`` hook (``dist.all_reduce`` over NCCL) runs launch by launch between the phases unless
        ``CSTR_NCCL_GRAPH=1`` (see ``_hook_capturable``)."""
        bs = int(batch_size or self._batch)
        done = 0
        if self._peer is not None:
            allreduce = None
        # a captured cycle bakes the sampling range in: only worth capturing once the ring is full (its range is constant from then on)
        if graph and _hook_capturable(allreduce) and getattr(buffer, "index_mode", None) == "philox" and buffer.full and _graph_env_ok(env):
            done = self._train_graph(gradient_steps, buffer, bs, env, allreduce)
        for _ in range(gradient_steps - done):
            self.update(buffer.sample(bs, env=env), allreduce=allreduce)

    def pop_losses(self):
        """(mean critic loss, mean actor loss or None) since the last call — what TD3.train logs (td3.py:207-210)."""
        s = self.loss_sums.cpu().numpy().astype(np.float64)
        self.loss_sums.zero_()
        critic = s[0] / s[1] if s[1] else None
        actor = s[2] / s[3] if s[3] else None
        return critic, actor


class _DeviceMemory:
    """``__cuda_array_interface__`` view of float32 device memory the library allocated (``cstr_peer_alloc``), for ``torch.as_tensor``."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _dist_rank() -> int:
    """Rank of this process in the default process group (0 when torch.distributed is not initialised)."""
    try:
        import torch.distributed as dist

        return int(dist.get_rank()) if dist.is_available() and dist.is_initialized() else 0
    except Exception:
        return 0


def _hook_capturable(allreduce) -> bool:
    """Whether ``train(graph=True)`` may capture the update.  With no host-side hook (single GPU, or the peer-memory all-reduce inside the Adam
    kernels) always.  A ``dist.all_reduce`` hook captured between the phases worked on 2 GPUs (tests, bench) but HUNG on 8 (ProcessGroupNCCL's
    watchdog stuck while eight ranks captured; profiles/r02_multi_gpu.log), so it is opt-in (``CSTR_NCCL_GRAPH=1``): by default a hook means
    launch by launch, as in round 1 — use ``enable_peer_allreduce()`` for a captured data-parallel update."""
    import os

    return allreduce is None or os.environ.get("CSTR_NCCL_GRAPH", "0") == "1"


def _graph_env_ok(env) -> bool:
    """A captured sample can only normalise on the device (``GpuVecNormalize.norm_params``); the reference's host-side ``VecNormalize``
    needs the launch-by-launch path, whose ``sample()`` round-trips through ``env.normalize_obs`` (buffer.py ``_finish``)."""
    return env is None or getattr(env, "norm_params", None) is not None


def fused_update_unsupported(model, max_critics: int = 2) -> Optional[str]:
    """Why the fused kernels can NOT stand in for ``model.train()`` — or None when they can.  The kernels are specialised to what the
    reference's TD3/DDPG/SAC build by default for the CSTR task (core/td3/policies.py:117-170, core/sac/policies.py:218-262): a flat
    4-value observation, a 2-value action, two ReLU hidden layers shared by actor and critics (multiples of 4), 1-2 critics and plain
    Adam.  Anything else keeps the reference's own torch ``train()``."""
    import torch.nn as nn
    import torch.optim as optim
    pol = model.policy
    if tuple(model.observation_space.shape or ()) != (4,) or tuple(model.action_space.shape or ()) != (2,):
        return f"spaces {model.observation_space.shape} -> {model.action_space.shape} are not the CSTR's (4,) -> (2,)"
    if getattr(pol, "activation_fn", nn.ReLU) is not nn.ReLU:
        return f"activation_fn {pol.activation_fn.__name__} (kernels are ReLU)"
    arch = pol.net_arch
    if isinstance(arch, dict):
        if list(arch.get("pi", [])) != list(arch.get("qf", [])):
            return "different actor and critic net_arch"
        arch = arch["pi"]
    arch = list(arch)
    if len(arch) != 2 or any(int(h) < 4 or int(h) > 4096 or int(h) % 4 for h in arch):
        return f"net_arch {arch}: two hidden layers, multiples of 4 in [4, 4096]"
    n_critics = len(pol.critic.q_networks)
    if not 1 <= n_critics <= max_critics:
        return f"n_critics={n_critics}"
    if type(pol.actor.features_extractor).__name__ != "FlattenExtractor":
        return f"features extractor {type(pol.actor.features_extractor).__name__}"
    for opt in (pol.actor.optimizer, pol.critic.optimizer):
        g = opt.param_groups[0]
        if type(opt) is not optim.Adam or g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False):
            return f"optimizer {type(opt).__name__} {({k: v for k, v in g.items() if k in ('weight_decay', 'amsgrad', 'maximize')})} (kernels are plain Adam)"
    return None


def _fallback_once(model, why: str) -> None:
    if not getattr(model, "_fused_fallback_warned", False):
        import warnings
        warnings.warn(f"fused update not used, the reference's torch train() runs instead: {why}", RuntimeWarning, stacklevel=3)
        model._fused_fallback_warned = True


def bind_td3_class(td3_base: type) -> type:
    """Return a subclass of the reference's ``TD3`` (or of ``DDPG``, which is TD3 with one critic, core/ddpg/ddpg.py) whose ``train()``
    (core/td3/td3.py:154-211) runs on ``cstr_td3_update``.
    Rollout collection, logging, saving and ``predict`` stay the reference's code; the policy modules share memory with the
    flat parameter blocks."""

    class FusedTD3(td3_base):  # type: ignore[misc, valid-type]
        _fused: Optional[FusedTD3Update] = None

        def _fused_engine(self, batch_size: int) -> FusedTD3Update:
            if self._fused is None:
                arch = self.policy.net_arch if isinstance(self.policy.net_arch, (list, tuple)) else self.policy.net_arch["pi"]
                eng = FusedTD3Update(arch, batch_size, self.device, self.gamma, self.tau, float(self.lr_schedule(self._current_progress_remaining)),
                                     self.policy_delay, self.target_policy_noise, self.target_noise_clip, seed=int(self.seed or 0),
                                     n_critics=len(self.policy.critic.q_networks),  # DDPG (a TD3 subclass) has one
                                     dp_rank=_dist_rank())
                eng.adopt_policy(self.policy)
                eng.import_optimizer_state(self.actor.optimizer, self.critic.optimizer)
                eng.n_updates = int(self._n_updates)
                self._fused = eng
            return self._fused

        def train(self, gradient_steps: int, batch_size: int = 100) -> None:
            why = fused_update_unsupported(self) if self._fused is None else None
            if why:
                _fallback_once(self, why)
                return super().train(gradient_steps, batch_size)
            self.policy.set_training_mode(True)
            self._update_learning_rate([self.actor.optimizer, self.critic.optimizer])
            eng = self._fused_engine(batch_size)
            eng.learning_rate = float(self.lr_schedule(self._current_progress_remaining))
            # a full ring has a constant sampling range: cycles of policy_delay updates replay from one captured CUDA graph
            eng.train(gradient_steps, self.replay_buffer, batch_size, env=self._vec_normalize_env, graph=bool(getattr(self.replay_buffer, "full", False)))
            self._n_updates = eng.n_updates
            critic_loss, actor_loss = eng.pop_losses()
            self.logger.record("train/n_updates", self._n_updates, exclude="tensorboard")
            if actor_loss is not None:
                self.logger.record("train/actor_loss", actor_loss)
            self.logger.record("train/critic_loss", critic_loss)

        def _excluded_save_params(self):
            return super()._excluded_save_params() + ["_fused"]

        def save(self, *args, **kwargs):
            if self._fused is not None:
                self._fused.export_optimizer_state(self.actor.optimizer, self.critic.optimizer)
            return super().save(*args, **kwargs)

    FusedTD3.__name__ = "TD3"
    FusedTD3.__qualname__ = "TD3"
    return FusedTD3


class FusedSACUpdate(FusedTD3Update):
    """SAC gradient steps on the device: one iteration of the loop body of ``SAC.train`` (``core/sac/sac.py:213-288``) per ``update()``
    through ``cstr_sac_update`` — squashed-Gaussian actor sample and log-prob, automatic entropy coefficient, soft twin-min target,
    critics, actor (backward through both critics and the tanh-Gaussian), polyak.  Same flat-block design as :class:`FusedTD3Update`;
    the actor head is ONE (4, h2) matrix [mu; log_std], ``log_ent_coef`` lives in a slot after the three nets."""

    def __init__(self, net_arch: Sequence[int] = (256, 256), batch_size: int = 256, device: Any = "cuda", gamma: float = 0.99, tau: float = 0.005,
                 learning_rate: float = 3e-4, target_entropy: float = -2.0, ent_coef_init: float = 1.0, target_update_interval: int = 1,
                 betas=(0.9, 0.999), eps: float = 1e-8, seed: int = 0, gemm: str = "fp32", dp_rank: int = 0):
        self.target_entropy, self.target_update_interval = float(target_entropy), int(target_update_interval)
        super().__init__(net_arch, batch_size, device, gamma, tau, learning_rate, 1, 0.0, 0.0, betas, eps, seed, gemm, dp_rank=dp_rank)
        torch = self._torch
        offs = (c_int64 * 20)()
        _lib.check(self._libc.cstr_sac_layout(self.h1, self.h2, offs), "cstr_sac_layout")
        self.param_count = int(offs[19])
        self._offsets = {(NET_NAMES[n], _TENSORS[k]): int(offs[n * 6 + k]) for n in range(3) for k in range(6)}
        self._ent_offset = int(offs[18])
        self.actor_range = (0, self._offsets[("critic0", "W1")])
        self.critic_range = (self._offsets[("critic0", "W1")], self._ent_offset)
        with torch.cuda.device(self.device):
            z = lambda: torch.zeros(self.param_count, dtype=torch.float32, device=self.device)  # noqa: E731
            self.params, self.targets, self.grads, self.adam_m, self.adam_v = z(), z(), z(), z(), z()
            self.loss_sums = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.params[self._ent_offset] = float(np.log(ent_coef_init))
        self._batch = 0
        self._set_batch(int(batch_size))

    def _shape(self, net: str, tensor: str):
        i, o = (4, 4) if net == "actor" else (6, 1)
        return {"W1": (self.h1, i), "b1": (self.h1,), "W2": (self.h2, self.h1), "b2": (self.h2,), "W3": (o, self.h2), "b3": (o,)}[tensor]

    @property
    def log_ent_coef(self):
        return self.params[self._ent_offset:self._ent_offset + 1]

    def _sac_config(self, batch: int) -> "_lib.SacConfig":
        return _lib.SacConfig(h1=self.h1, h2=self.h2, batch=batch, target_update_interval=self.target_update_interval, gamma=self.gamma, tau=self.tau,
                              lr=self.learning_rate, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, target_entropy=self.target_entropy,
                              seed=self._keyed_seed(), gemm_mode=_GEMM_MODES[self.gemm])

    def _workspace_bytes(self, batch: int) -> int:
        return int(self._libc.cstr_sac_workspace_bytes(byref(self._sac_config(batch))))

    def load_nets(self, nets: Dict[str, Sequence[Any]]) -> None:
        """``actor`` (head = [mu; log_std] stacked), ``critic0``, ``critic1`` and optionally ``critic*_target``."""
        torch = self._torch
        for block, suffix, names in (("params", "", NET_NAMES), ("targets", "_target", NET_NAMES[1:])):
            v = self.views(block)
            for net in names:
                src = nets.get(net + suffix, nets[net])
                for dst, s in zip(v[net], src):
                    dst.copy_(torch.as_tensor(np.asarray(s) if not isinstance(s, torch.Tensor) else s).to(self.device, torch.float32).reshape(dst.shape))

    def nets(self) -> Dict[str, List[np.ndarray]]:
        out = {net: [t.cpu().numpy() for t in ts] for net, ts in self.views("params").items()}
        out.update({net + "_target": [t.cpu().numpy() for t in ts] for net, ts in self.views("targets").items() if net != "actor"})
        return out

    def adopt_policy(self, policy) -> None:
        """The reference's ``SACPolicy`` (core/sac/policies.py): ``actor.latent_pi`` + ``actor.mu`` + ``actor.log_std``, twin critics."""
        a, v, t = policy.actor, self.views("params"), self.views("targets")
        lat = list(a.latent_pi.parameters())
        pairs = list(zip(v["actor"][:4], lat)) + [(v["actor"][4][0:2], a.mu.weight), (v["actor"][4][2:4], a.log_std.weight),
                                                 (v["actor"][5][0:2], a.mu.bias), (v["actor"][5][2:4], a.log_std.bias)]
        for z, q in enumerate(policy.critic.q_networks):
            pairs += list(zip(v[f"critic{z}"], q.parameters()))
        for z, q in enumerate(policy.critic_target.q_networks):
            pairs += list(zip(t[f"critic{z}"], q.parameters()))
        for view, p in pairs:
            if tuple(p.shape) != tuple(view.shape):
                raise ValueError(f"shape mismatch: module {tuple(p.shape)} vs layout {tuple(view.shape)}")
            view.copy_(p.data.to(self.device, self._torch.float32))
            p.data = view
        # Adam moment views per module parameter (same slicing as the weights), for import/export_optimizer_state
        am, av = self.views("adam_m"), self.views("adam_v")
        head = lambda blk: [blk["actor"][4][0:2], blk["actor"][4][2:4], blk["actor"][5][0:2], blk["actor"][5][2:4]]  # noqa: E731
        actor_params = lat + [a.mu.weight, a.log_std.weight, a.mu.bias, a.log_std.bias]
        critic_params = [p for q in policy.critic.q_networks for p in q.parameters()]
        self._moments = {id(p): mv for p, mv in zip(actor_params, zip(am["actor"][:4] + head(am), av["actor"][:4] + head(av)))}
        self._moments.update({id(p): mv for p, mv in zip(critic_params, zip(am["critic0"] + am["critic1"], av["critic0"] + av["critic1"]))})

    def import_optimizer_state(self, actor_optimizer, critic_optimizer, ent_coef_optimizer=None) -> None:  # type: ignore[override]
        """Take over Adam moments / the step count of the reference's three optimisers (sac.py:187, policies.py:264-284) after ``adopt_policy``.
        One step count serves all three: the reference steps each exactly once per gradient step."""
        step = 0
        for opt in (actor_optimizer, critic_optimizer):
            group = opt.param_groups[0]
            self.betas, self.eps = (float(group["betas"][0]), float(group["betas"][1])), float(group["eps"])
            for p in group["params"]:
                st = opt.state.get(p)
                if st and id(p) in self._moments:
                    m, v = self._moments[id(p)]
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                    step = max(step, int(float(st["step"])))
        if ent_coef_optimizer is not None:
            for p in ent_coef_optimizer.param_groups[0]["params"]:
                st = ent_coef_optimizer.state.get(p)
                if st:
                    self.adam_m[self._ent_offset] = float(st["exp_avg"])
                    self.adam_v[self._ent_offset] = float(st["exp_avg_sq"])
                    step = max(step, int(float(st["step"])))
        self.critic_step = self.actor_step = step

    def export_optimizer_state(self, actor_optimizer, critic_optimizer, ent_coef_optimizer=None) -> None:  # type: ignore[override]
        """Write moments / step count back so ``model.save`` and a later torch ``optimizer.step()`` continue from here."""
        torch = self._torch
        if self.critic_step == 0:
            return
        step = lambda: torch.tensor(float(self.critic_step))  # noqa: E731
        for opt in (actor_optimizer, critic_optimizer):
            for p in opt.param_groups[0]["params"]:
                if id(p) in self._moments:
                    m, v = self._moments[id(p)]
                    opt.state[p] = {"step": step(), "exp_avg": m, "exp_avg_sq": v}
        if ent_coef_optimizer is not None:
            e = self._ent_offset
            for p in ent_coef_optimizer.param_groups[0]["params"]:
                ent_coef_optimizer.state[p] = {"step": step(), "exp_avg": self.adam_m[e:e + 1].reshape(p.shape), "exp_avg_sq": self.adam_v[e:e + 1].reshape(p.shape)}

    def update(self, batch, eps_pi=None, eps_next=None, allreduce: Optional[Callable[[Any], None]] = None,  # type: ignore[override]
               gradient_step: Optional[int] = None) -> None:
        """One iteration of sac.py:213-288.  ``eps_pi`` / ``eps_next``: explicit standard-normal draws (B,2) of the two rsample() calls
        (parity tests); default = Philox inside the kernel.  ``allreduce``: called on grads[critics .. log_ent_coef] and on grads[actor]
        between backward and Adam (data-parallel training: every rank ends with the gradient of the global batch).
        ``gradient_step``: the index of this step inside the current ``train()`` call — what the reference's target sync tests
        (``gradient_step % target_update_interval``, sac.py:284); None = the engine's own running count of updates."""
        obs, act, nobs, dones, rew = batch
        self._set_batch(int(obs.shape[0]))
        obs, act, nobs = self._f32(obs, 4), self._f32(act, 2), self._f32(nobs, 4)
        dones, rew = self._f32(dones, 1), self._f32(rew, 1)
        e1 = None if eps_pi is None else self._f32(eps_pi, 2)
        e2 = None if eps_next is None else self._f32(eps_next, 2)
        self.n_updates += 1
        self.critic_step += 1
        self.actor_step += 1
        cfg, st = self._sac_config(self._batch), self._state(counters=False)
        cfg.local_step = 0 if gradient_step is None else int(gradient_step) + 1

        def run(phases):
            rc = self._libc.cstr_sac_update(byref(cfg), byref(st), _lib.ptr(obs), _lib.ptr(act), _lib.ptr(nobs), _lib.ptr(dones), _lib.ptr(rew),
                                            _lib.ptr(e1), _lib.ptr(e2), self.n_updates, self.critic_step, phases, self._stream())
            _lib.check(rc, "cstr_sac_update")

        with self._torch.cuda.device(self.device):
            if allreduce is None or self._peer is not None:
                run(_lib.TD3_ALL)
            else:
                run(_lib.TD3_CRITIC_GRAD)
                allreduce(self.grads[self.critic_range[0]:self._ent_offset + 4])  # both critics and the log_ent_coef slot, contiguous
                run(_lib.TD3_CRITIC_APPLY | _lib.TD3_ACTOR_GRAD)
                allreduce(self.grads[self.actor_range[0]:self.actor_range[1]])
                run(_lib.TD3_ACTOR_APPLY)
        self.launches += 50

    def _graph_launch(self, batch_size: int, st, out, k: int, allreduce=None) -> None:
        cfg = self._sac_config(batch_size)

        def run(phases):
            rc = self._libc.cstr_sac_update(byref(cfg), byref(st), _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(out[3]),
                                            _lib.ptr(out[4]), None, None, 1, 1, phases, self._stream())
            _lib.check(rc, "cstr_sac_update (graph capture)")

        if allreduce is None:
            run(_lib.TD3_ALL)
        else:
            run(_lib.TD3_CRITIC_GRAD)
            allreduce(self.grads[self.critic_range[0]:self._ent_offset + 4])
            run(_lib.TD3_CRITIC_APPLY | _lib.TD3_ACTOR_GRAD)
            allreduce(self.grads[self.actor_range[0]:self.actor_range[1]])
            run(_lib.TD3_ACTOR_APPLY)

    def train(self, gradient_steps: int, buffer, batch_size: Optional[int] = None, env=None, allreduce=None, graph: bool = False) -> None:  # type: ignore[override]
        """``graph=True`` (single GPU, Philox-index buffer, full ring, target_update_interval 1): every update replays one captured CUDA graph."""
        bs = int(batch_size or self._batch)
        done = 0
        if self._peer is not None:
            allreduce = None
        if (graph and _hook_capturable(allreduce) and getattr(buffer, "index_mode", None) == "philox" and buffer.full and self.target_update_interval == 1
                and _graph_env_ok(env)):
            done = self._train_graph(gradient_steps, buffer, bs, env, allreduce)
        for g in range(done, gradient_steps):  # g = the reference's loop index: the target sync tests `g % target_update_interval` (sac.py:284)
            self.update(buffer.sample(bs, env=env), allreduce=allreduce, gradient_step=g)

    def pop_losses(self):
        """(critic loss, actor loss, ent_coef_loss, ent_coef) means since the last call — the keys SAC.train logs (sac.py:290-296)."""
        s = self.loss_sums.cpu().numpy().astype(np.float64)
        self.loss_sums.zero_()
        return tuple(s[2 * i] / s[2 * i + 1] if s[2 * i + 1] else None for i in range(4))


def bind_sac_class(sac_base: type) -> type:
    """Subclass of the reference's ``SAC`` (``ent_coef="auto"``, no gSDE) whose ``train()`` (core/sac/sac.py:199-296) runs on ``cstr_sac_update``."""

    class FusedSAC(sac_base):  # type: ignore[misc, valid-type]
        _fused: Optional[FusedSACUpdate] = None

        def train(self, gradient_steps: int, batch_size: int = 64) -> None:
            why = None
            if self._fused is None:
                why = ("use_sde" if self.use_sde else "fixed ent_coef" if self.ent_coef_optimizer is None else
                       "n_critics != 2" if len(self.policy.critic.q_networks) != 2 else fused_update_unsupported(self))
            if why:  # the reference's torch path
                _fallback_once(self, why)
                return super().train(gradient_steps, batch_size)
            self.policy.set_training_mode(True)
            self._update_learning_rate([self.actor.optimizer, self.critic.optimizer, self.ent_coef_optimizer])
            if self._fused is None:
                arch = self.policy.net_arch if isinstance(self.policy.net_arch, (list, tuple)) else self.policy.net_arch["pi"]
                eng = FusedSACUpdate(arch, batch_size, self.device, self.gamma, self.tau, float(self.lr_schedule(self._current_progress_remaining)),
                                     float(self.target_entropy), float(self.log_ent_coef.detach().exp().item()), self.target_update_interval,
                                     seed=int(self.seed or 0), dp_rank=_dist_rank())
                eng.adopt_policy(self.policy)
                self.log_ent_coef.data = eng.log_ent_coef  # shared storage
                eng.import_optimizer_state(self.actor.optimizer, self.critic.optimizer, self.ent_coef_optimizer)
                eng.n_updates = int(self._n_updates)
                self._fused = eng
            eng = self._fused
            eng.learning_rate = float(self.lr_schedule(self._current_progress_remaining))
            eng.train(gradient_steps, self.replay_buffer, batch_size, env=self._vec_normalize_env, graph=bool(getattr(self.replay_buffer, "full", False)))
            self._n_updates += gradient_steps
            critic_loss, actor_loss, ent_loss, ent_coef = eng.pop_losses()
            self.logger.record("train/n_updates", self._n_updates, exclude="tensorboard")
            self.logger.record("train/ent_coef", ent_coef)
            self.logger.record("train/actor_loss", actor_loss)
            self.logger.record("train/critic_loss", critic_loss)
            self.logger.record("train/ent_coef_loss", ent_loss)

        def _excluded_save_params(self):
            return super()._excluded_save_params() + ["_fused"]

        def save(self, *args, **kwargs):
            if self._fused is not None:
                self._fused.export_optimizer_state(self.actor.optimizer, self.critic.optimizer, self.ent_coef_optimizer)
            return super().save(*args, **kwargs)

    FusedSAC.__name__ = "SAC"
    FusedSAC.__qualname__ = "SAC"
    return FusedSAC
