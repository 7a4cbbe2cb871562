"""GPU-resident replay buffer with the call surface the reference uses from torchrl
(``TensorDictReplayBuffer(storage=LazyTensorStorage(capacity, device))``, main.py:167-171):
``extend(td)`` (orchestrator.py:100-113), ``sample(batch_size)`` (orchestrator.py:338), ``len()``
(orchestrator.py:385). Semantics restated from torchrl's defaults: round-robin writer,
uniform-with-replacement RandomSampler, storage shaped lazily by the first ``extend``.

Storage is one padded array-of-structs row per transition (include/b2rl.h "transition rows");
sampling is a single CUDA launch (csrc/replay.cu) that draws the indices (Philox) or takes them
from the caller, and gathers whole rows with 128-bit loads.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional

import torch

from . import _lib as L

KEYS = ("observations", "next_observations", "actions", "rewards", "terminations", "dones")


def row_format(ob_dim: int, ac_dim: int) -> L.RowFmt:
    return L.RowFmt(ob_dim, ac_dim, (2 * ob_dim + ac_dim + 2 + 3) & ~3, 0)


class Batch(dict):
    """A sampled batch: the packed rows (what the kernels consume) plus per-key views with the
    reference's shapes and dtypes. ``dones``/``terminations`` are materialised as bool on access."""

    def __init__(self, rows: torch.Tensor, fmt: L.RowFmt, index: Optional[torch.Tensor] = None):
        O, A = fmt.ob_dim, fmt.ac_dim
        super().__init__(observations=rows[:, :O], actions=rows[:, O:O + A], rewards=rows[:, O + A:O + A + 1],
                         next_observations=rows[:, O + A + 2:2 * O + A + 2])
        if index is not None:
            self["index"] = index
        self.rows, self.fmt = rows, fmt

    def __missing__(self, key):
        if key in ("dones", "terminations"):
            O, A = self.fmt.ob_dim, self.fmt.ac_dim
            return self.rows[:, O + A + 1:O + A + 2] != 0
        raise KeyError(key)

    def __contains__(self, key):
        return key in ("dones", "terminations") or super().__contains__(key)


def pack_rows(td: Mapping[str, torch.Tensor], fmt: L.RowFmt, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[n] transitions given as the reference's six keys -> [n, row_stride] packed rows.
    Only ``dones`` is stored: the reference's writer sets ``dones = terminations`` (orchestrator.py:107-108) and the
    update reads ``dones`` alone (agents/agent.py:226-228), so ``Batch["terminations"]`` returns the same column."""
    obs = td["observations"]
    n = obs.shape[0]
    if out is None:
        out = torch.zeros(n, fmt.row_stride, dtype=torch.float32, device=obs.device)
    O, A = fmt.ob_dim, fmt.ac_dim
    out[:, :O] = obs
    out[:, O:O + A] = td["actions"]
    out[:, O + A] = td["rewards"].reshape(n)
    out[:, O + A + 1] = td["dones"].reshape(n).to(torch.float32)
    out[:, O + A + 2:2 * O + A + 2] = td["next_observations"]
    return out


class ReplayBuffer:
    def __init__(self, capacity: int, device, seed: int = 0, agent_id: int = 0):
        self.capacity, self.device, self.seed = int(capacity), torch.device(device), int(seed)
        self.agent_id = int(agent_id)  # global learner id: part of the Philox key of the index draws
        self.fmt: Optional[L.RowFmt] = None
        self.storage: Optional[torch.Tensor] = None
        self._size = 0
        self._cursor = 0
        self._lib = L.load()
        L.init_device(self.device)
        self.counters = torch.zeros(8, dtype=torch.int64, device=self.device)
        self._rows: dict[int, torch.Tensor] = {}
        self._idx: dict[int, torch.Tensor] = {}

    def __len__(self) -> int:
        return self._size

    def _ensure(self, ob_dim: int, ac_dim: int) -> None:
        if self.storage is None:
            self.fmt = row_format(ob_dim, ac_dim)
            self.storage = torch.zeros(self.capacity, self.fmt.row_stride, dtype=torch.float32, device=self.device)

    def extend(self, td: Mapping[str, torch.Tensor]) -> None:
        obs = td["observations"]
        self._ensure(obs.shape[-1], td["actions"].shape[-1])
        rows = pack_rows(td, self.fmt)
        self.extend_rows(rows)

    def extend_rows(self, rows: torch.Tensor) -> None:
        n = rows.shape[0]
        assert rows.is_contiguous() and rows.shape[1] == self.fmt.row_stride and n <= self.capacity
        st = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self._lib.b2rl_replay_extend(self.storage.data_ptr(), self.capacity, self._cursor, self.fmt,
                                             rows.data_ptr(), n, st), "replay_extend")
        self._cursor = (self._cursor + n) % self.capacity
        self._size = min(self.capacity, self._size + n)

    def fill_(self, td: Mapping[str, torch.Tensor]) -> None:
        """Bulk-load transitions (tests / benchmarks)."""
        self.extend(td)

    def sample(self, batch_size: int, idx: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
               idx_out: Optional[torch.Tensor] = None) -> Batch:
        if self._size == 0:
            raise RuntimeError("cannot sample from an empty replay buffer")
        # (the returned Batch ALIASES these cached buffers: the next sample() of the same size overwrites it — pass
        # `out` / `idx_out` to keep a batch)
        if out is None:
            if batch_size not in self._rows:
                self._rows[batch_size] = torch.empty(batch_size, self.fmt.row_stride, dtype=torch.float32, device=self.device)
            out = self._rows[batch_size]
        if idx_out is None:
            if batch_size not in self._idx:
                self._idx[batch_size] = torch.empty(batch_size, dtype=torch.int64, device=self.device)
            idx_out = self._idx[batch_size]
        if idx is not None:
            idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        st = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self._lib.b2rl_replay_sample_gather(
            self.storage.data_ptr(), 0, self._size, self.fmt, batch_size, 1, L.ptr(idx), idx_out.data_ptr(),
            out.data_ptr(), C.c_uint64(self.seed), self.counters.data_ptr(), L.CTR_SAMPLE, 1, self.agent_id, st),
            "replay_sample_gather")
        return Batch(out, self.fmt, idx_out)
