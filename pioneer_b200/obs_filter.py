"""Observation normaliser for the rollout loop: the ``'observation_filter': 'ConcurrentMeanStdFilter'`` the reference launcher
configures (pioneer/launch/pioneer_knm_train.py:66; the filter itself is RLlib's).  One streaming CUDA pass per batch
(pnr_filter_apply) pushes the rows into the running statistics and rewrites them as
``clip((x - mean) / (std + 1e-8), +-clip)``; ``sync()`` merges what was pushed since the last call -- summed over the
ranks with one all-reduce when torch.distributed is initialised -- into the running mean / variance, which is the
once-per-iteration filter synchronisation RLlib performs between its rollout workers and the trainer.

Between two ``sync()`` calls every rank normalises with the same statistics (RLlib workers keep updating their local
copy in between; the difference vanishes once the statistics have converged)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi
from .batched_env import OBS_DIM, BatchedPioneerEnv


class MeanStdObsFilter:
    def __init__(self, env: BatchedPioneerEnv, clip: float = 10.0, demean: bool = True, destd: bool = True,
                 fused: bool = False, update: bool = True):
        """``fused=True``: the env's step kernel itself normalises the observations it writes and pushes their raw values
        into the statistics (pnr_filter_fuse) -- ``env.step_tensor`` then returns normalised observations and this
        object is only needed for ``sync()``; the first observations of a run (``env.reset()``) still go through
        ``__call__``.  Kinematic mode: float32 arithmetic and terminal observations only; dynamic mode: both observation modes."""
        self.env = env
        self._lib, self._h = env._lib, env._h
        env._filters.append(self)               # env.close() clears _h: a filter must not outlive the handle
        self.clip, self.demean, self.destd, self.fused = float(clip), bool(demean), bool(destd), bool(fused)
        with torch.cuda.device(env.device):
            _cabi.check(self._lib.pnr_filter_configure(self._h, self.clip, int(demean), int(destd)), "pnr_filter_configure")
            if fused:
                _cabi.check(self._lib.pnr_filter_fuse(self._h, 1, int(update)), "pnr_filter_fuse")

    def _handle(self):
        if self._h is None or not self._h.value:
            raise _cabi.PioneerB200Error("MeanStdObsFilter: the env it was created for has been closed")
        return self._h

    def set_fused(self, on: bool, update: bool = True) -> None:
        self._handle()
        _cabi.check(self._lib.pnr_filter_fuse(self._h, int(on), int(update)), "pnr_filter_fuse")
        self.fused = bool(on)

    def __call__(self, obs: torch.Tensor, update: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Normalise ``obs`` [n, 137] (float32, on the env's device); ``out`` defaults to in-place."""
        out = obs if out is None else out
        self._handle()
        assert obs.dtype == torch.float32 and obs.is_contiguous() and obs.shape[-1] == OBS_DIM and obs.device == self.env.device
        assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == obs.shape
        with torch.cuda.device(self.env.device):
            _cabi.check(self._lib.pnr_filter_apply(self._h, obs.data_ptr(), out.data_ptr(), obs.numel() // OBS_DIM,
                                                   int(update), 1, self.env._stream()), "pnr_filter_apply")
        return out

    def push(self, obs: torch.Tensor) -> None:
        """Statistics only; ``obs`` is left untouched."""
        self._handle()
        assert obs.dtype == torch.float32 and obs.is_contiguous() and obs.shape[-1] == OBS_DIM
        with torch.cuda.device(self.env.device):
            _cabi.check(self._lib.pnr_filter_apply(self._h, obs.data_ptr(), obs.data_ptr(), obs.numel() // OBS_DIM,
                                                   1, 0, self.env._stream()), "pnr_filter_apply")

    def sync(self, group: Optional[dist.ProcessGroup] = None) -> None:
        """Merge the rows pushed since the last sync (of ALL ranks) into the running statistics.  Stays on the device:
        one small kernel (after one all-reduce when distributed), no host round trip; ``n`` / ``mean`` / ``var`` read the
        result back on demand."""
        self._handle()
        with torch.cuda.device(self.env.device):
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                delta = torch.empty(_cabi.PNR_FILTER_DELTA_LEN, dtype=torch.float64, device=self.env.device)
                _cabi.check(self._lib.pnr_filter_delta_device(self._h, delta.data_ptr(), self.env._stream()),
                            "pnr_filter_delta_device")
                dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=group)
                _cabi.check(self._lib.pnr_filter_sync_device(self._h, delta.data_ptr(), self.env._stream()),
                            "pnr_filter_sync_device")
                self._keep = delta                      # the kernel reads it asynchronously
            else:
                _cabi.check(self._lib.pnr_filter_sync_device(self._h, None, self.env._stream()), "pnr_filter_sync_device")

    def delta(self) -> torch.Tensor:
        """The statistics pushed on THIS rank since the last sync: float64[1 + 2 * 137] on the device (rows, sum(x - mean),
        sum((x - mean)^2)); additive over ranks."""
        self._handle()
        with torch.cuda.device(self.env.device):
            out = torch.empty(_cabi.PNR_FILTER_DELTA_LEN, dtype=torch.float64, device=self.env.device)
            _cabi.check(self._lib.pnr_filter_delta_device(self._h, out.data_ptr(), self.env._stream()),
                        "pnr_filter_delta_device")
        return out

    def apply_merged(self, merged: torch.Tensor) -> None:
        """Finish a synchronisation with a delta that was summed over the ranks elsewhere (e.g. packed into the
        rollout worker's one collective per iteration)."""
        assert merged.dtype == torch.float64 and merged.numel() == _cabi.PNR_FILTER_DELTA_LEN and merged.is_contiguous()
        self._handle()
        with torch.cuda.device(self.env.device):
            _cabi.check(self._lib.pnr_filter_sync_device(self._h, merged.data_ptr(), self.env._stream()),
                        "pnr_filter_sync_device")
        self._keep = merged

    def _get(self):
        self._handle()
        n = C.c_double()
        mean = (C.c_double * OBS_DIM)()
        var = (C.c_double * OBS_DIM)()
        _cabi.check(self._lib.pnr_filter_get(self._h, C.byref(n), mean, var), "pnr_filter_get")
        return n.value, np.array(mean), np.array(var)

    @property
    def n(self) -> float:
        return self._get()[0]

    @property
    def mean(self) -> np.ndarray:
        return self._get()[1]

    @property
    def var(self) -> np.ndarray:
        return self._get()[2]

    @property
    def std(self) -> np.ndarray:
        return np.sqrt(self.var)

    def set_stats(self, count: float, mean, var) -> None:
        m = np.ascontiguousarray(mean, dtype=np.float64)
        v = np.ascontiguousarray(var, dtype=np.float64)
        assert m.shape == (OBS_DIM,) and v.shape == (OBS_DIM,)
        self._handle()
        with torch.cuda.device(self.env.device):
            _cabi.check(self._lib.pnr_filter_set(self._h, float(count), m.ctypes.data_as(C.POINTER(C.c_double)),
                                                 v.ctypes.data_as(C.POINTER(C.c_double)), self.env._stream()), "pnr_filter_set")
