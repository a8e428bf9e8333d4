"""Device-resident rollout worker: what an RLlib RolloutWorker does around the reference env
(pioneer/launch/pioneer_knm_train.py:43-73 -- sample `train_batch_size` env steps with the current policy, observation
filter 'ConcurrentMeanStdFilter', episode metrics back to the trainer), with every tensor staying on the GPU:

    obs --(pnr_filter_apply)--> normalised obs --(policy MLP)--> action --(pnr_step)--> obs', reward, done

The policy here is the launcher's model shape (fcnet_hiddens [256, 256], tanh; `pioneer_knm_train.py:59-61`) as a
plain torch module with a diagonal-Gaussian head -- a stand-in for the PPO policy, which is RLlib's (out of scope);
PyTorch / cuBLAS are plumbing on this side of the boundary.  The trainer is not reproduced: ``collect()`` returns
the fragment tensors a trainer would consume.  Once per iteration ``sync()`` all-reduces the episode statistics and
the filter statistics over the ranks (the only collectives of the path)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .batched_env import DOF, OBS_DIM, BatchedPioneerEnv
from .distributed import IterationSync, summarize
from .obs_filter import MeanStdObsFilter


class GaussianMlpPolicy(nn.Module):
    """137 -> 256 -> 256 -> 12 (mean, log_std), tanh: the launcher's `fcnet_hiddens` / `fcnet_activation`.

    The input is zero-padded to 144 columns and the head to 16 outputs: cuBLAS picks its 16-byte-aligned bf16 kernels
    only when K and N are multiples of 8 (with K = 137 the first layer ran a legacy align1 kernel, 112 us per
    131,072-row call against 16 us padded; profiles/r01_rollout_launches.md).  The padding columns of the first weight
    multiply zeros and the padding outputs are dropped, so the function is the 137 -> 12 network."""

    IN_PAD = (OBS_DIM + 7) // 8 * 8           # 144
    OUT_PAD = (2 * DOF + 7) // 8 * 8          # 16

    def __init__(self, hiddens=(256, 256), dtype=torch.float32):
        super().__init__()
        layers, d = [], self.IN_PAD
        for h in hiddens:
            layers += [nn.Linear(d, h), nn.Tanh()]
            d = h
        layers.append(nn.Linear(d, self.OUT_PAD))
        self.net = nn.Sequential(*layers).to(dtype)
        with torch.no_grad():
            self.net[0].weight[:, OBS_DIM:].zero_()
            self.net[-1].weight.mul_(0.01)
            self.net[-1].bias.zero_()
        self._x: Optional[torch.Tensor] = None

    def forward(self, obs: torch.Tensor):
        w = self.net[0].weight
        if self._x is None or self._x.shape[0] != obs.shape[0] or self._x.device != obs.device or self._x.dtype != w.dtype:
            self._x = torch.zeros((obs.shape[0], self.IN_PAD), dtype=w.dtype, device=obs.device)
        self._x[:, :OBS_DIM].copy_(obs)                                # cast + pad in one strided copy
        out = self.net(self._x).float()
        return out[:, :DOF], out[:, DOF:2 * DOF].clamp(-5.0, 2.0)


class RolloutWorker:
    def __init__(self, env: BatchedPioneerEnv, fragment_length: int = 8, policy: Optional[nn.Module] = None,
                 use_filter: bool = True, seed: int = 0, policy_dtype=torch.bfloat16, fused_filter: bool = True,
                 cuda_graph: bool = False, keep_terminal_obs: bool = False):
        """``cuda_graph=True``: the T steps of a fragment (policy, sampling, env step) are captured once and replayed,
        which removes the ~25 launch gaps per env step; the noise then comes from the device's default generator
        (graph-safe) and the env's reset counter advances on the device (pnr_tick_advance).  Needs the fused filter (or
        none): a separate filter pass would be captured too, but its synchronisation is not.
        ``keep_terminal_obs``: also return the terminal observations of finished rows (``terminal_obs`` [T,N,137], written
        for those rows only)."""
        self.env, self.T = env, int(fragment_length)
        self.use_graph, self._graph = bool(cuda_graph), None
        dev, n = env.device, env.n_envs
        self.policy = (policy or GaussianMlpPolicy(dtype=policy_dtype)).to(dev)
        self.fused = bool(use_filter and fused_filter)          # the step kernel normalises and pushes statistics itself
        self.filter = MeanStdObsFilter(env, fused=self.fused) if use_filter else None
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.a_max = torch.as_tensor(env.a_max, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        # fragment storage [T, N, ...]: raw observations are overwritten in place by their normalised version.  The row
        # count of the allocation is padded to a multiple of 4 so that every obs[t] starts on a 16-byte boundary (the step
        # kernel stores observation tiles with 16-byte bulk copies) whatever n is.
        n_pad = (n + 3) // 4 * 4
        self._obs_store = torch.empty((self.T + 1, n_pad, OBS_DIM), **f32)
        self.obs = self._obs_store[:, :n]
        # terminal observations of the rows that finished (written for those rows only), on request: the policy input after
        # a done is the RESET observation (RLlib's sampler, bullet_env.py:187-190), which replaces the row in `obs`
        self.terminal_obs = torch.zeros((self.T, n_pad, OBS_DIM), **f32)[:, :n] if keep_terminal_obs else None
        self.actions = torch.empty((self.T, n, DOF), **f32)
        self.logp = torch.empty((self.T, n), **f32)
        self.reward = torch.empty((self.T, n), **f32)
        self.flags = torch.empty((self.T, n), dtype=torch.uint8, device=dev)
        self._have_first = False
        self._sync: Optional[IterationSync] = None
        self._last_done: Optional[torch.Tensor] = None
        # terminal-observation envs with in-kernel auto-reset need their done rows patched; 'autoreset' envs already return
        # the fresh observation (and have no terminal one to offer)
        self._patch_done = env.batch_config.auto_reset and env.batch_config.obs_mode == "terminal"
        if keep_terminal_obs:
            assert self._patch_done, "terminal observations exist only in obs_mode='terminal' with auto_reset"

    def _filter_rows(self, obs: torch.Tensor, flags: torch.Tensor) -> None:
        """Separate-pass filter: normalise (and push) just the rows pnr_observe_done replaced.  Eager only (dynamic shape)."""
        idx = torch.nonzero(flags & 1).flatten()
        if idx.numel():
            rows = obs[idx].contiguous()
            self.filter(rows)
            obs[idx] = rows

    def _steps(self) -> None:
        env = self.env
        self.obs[0].copy_(self.obs[self.T])
        for t in range(self.T):
            mean, log_std = self.policy(self.obs[t])
            noise = torch.randn(mean.shape, device=mean.device, generator=None if self.use_graph else self.gen)
            raw = mean + log_std.exp() * noise
            self.logp[t] = (-0.5 * noise.pow(2) - log_std).sum(-1)
            torch.mul(torch.tanh(raw), self.a_max, out=self.actions[t])          # squash into the action space
            env.step_tensor(self.actions[t], out=(self.obs[t + 1], self.reward[t], self.flags[t]))
            if self.filter is not None and not self.fused:
                self.filter(self.obs[t + 1])                                      # push + normalise in place
            if self._patch_done:
                # rows that finished: the next policy input is the first observation of the new episode, not the terminal one
                env.observe_done(self.flags[t], self.obs[t + 1],
                                 None if self.terminal_obs is None else self.terminal_obs[t])
                if self.filter is not None and not self.fused:
                    self._filter_rows(self.obs[t + 1], self.flags[t])

    @torch.no_grad()
    def collect(self) -> Dict[str, torch.Tensor]:
        """T env steps of every env; returns views of the fragment buffers (valid until the next call)."""
        env = self.env
        if not self._have_first:
            self.obs[self.T].copy_(env.reset())
            if self.filter is not None:
                # RLlib's filter pushes a sample before it normalises it: push the reset observations, make them the running
                # statistics, only then normalise (without pushing them twice)
                self.filter.push(self.obs[self.T])
                self.filter.sync()
                self.filter(self.obs[self.T], update=False)
            self._have_first = True
            if self.use_graph:
                with torch.cuda.device(env.device):
                    self._steps()                                                 # warm-up: allocations, kernel attributes
                    torch.cuda.synchronize(env.device)
                    self._graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self._graph):
                        env.advance_reset_counter(self.T)
                        self._steps()
        if self._graph is not None:
            self._graph.replay()
        else:
            self._steps()
        done = (self.flags & 1).bool()
        out = dict(obs=self.obs[:self.T], next_obs=self.obs[1:], actions=self.actions, logp=self.logp,
                   reward=self.reward, done=done, truncated=(self.flags & 2).bool())
        # obs[t] is what the policy saw at step t.  After a done that is the first observation of the NEW episode, so
        # next_obs[t] of a finished row belongs to the next episode; its terminal observation is in `terminal_obs` (when
        # kept).  episode_start[t]: row starts an episode at step t (within the fragment; the previous fragment's last
        # `done` marks the starts of step 0).
        start = torch.zeros_like(done)
        start[1:] = done[:-1]
        if self._last_done is not None:
            start[0] = self._last_done
        self._last_done = done[-1].clone()
        out["episode_start"] = start
        if self.terminal_obs is not None:
            out["terminal_obs"] = self.terminal_obs
        return out

    def sync(self, group=None, summary: bool = True):
        """Once per training iteration: episode statistics and filter statistics of all ranks in ONE collective
        (``distributed.IterationSync``: snapshot + filter delta -> one all-gather -> one merge kernel -> filter merge, all on
        the device; replayed from a CUDA graph when the worker runs its fragments from one).  ``summary=False`` returns the
        reduced float64[8] statistics tensor without reading it back, so the loop never waits for the GPU."""
        if self._sync is None or self._sync.group is not group:
            self._sync = IterationSync(self.env, self.filter, group, clear=True, cuda_graph=self.use_graph)
        stats = self._sync()
        return summarize(stats) if summary else stats
