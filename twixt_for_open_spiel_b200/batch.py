"""TwixTBatch -- Python face of the C++ host class of the same name
(csrc/twixt_batch.cu) over the C ABI in include/twixt_b200.h.

One batch = `num_envs` independent TwixT games of one board size resident in
the HBM of one GPU.  Every method is one batched call of the open_spiel State
surface the reference implements (twixt.h:31-112).  Array arguments may be
numpy arrays (host path: staged, synchronous) or torch CUDA tensors (device
path: the kernel reads/writes them directly on the batch's stream).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib

# torch is plumbing only (device buffers, streams) and is imported LAZILY: a host that hands over numpy
# arrays -- e.g. the per-board-size worker processes of the reference-parity tests -- never pays for it.
torch = None


def _torch():
    global torch
    if torch is None:
        import torch as _t
        torch = _t
    return torch


class SpielFatalError(RuntimeError):
    """Raised where the reference calls SpielFatalError (twixt.h:96, twixt.cc:140-143)."""


class TwixTCudaError(RuntimeError):
    pass


def _is_torch(a) -> bool:
    if not type(a).__module__.startswith("torch"):
        return False
    return isinstance(a, _torch().Tensor)


def _ptr(a, dtype=None, min_elems: int = 0) -> int:
    """Address of a numpy array or torch tensor (None -> 0)."""
    if a is None:
        return 0
    if _is_torch(a):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if dtype is not None and a.dtype != getattr(_torch(), dtype[1]):
            raise TypeError("expected torch dtype %s, got %s" % (dtype[1], a.dtype))
        if a.numel() < min_elems:
            raise ValueError("tensor too small: %d < %d" % (a.numel(), min_elems))
        return a.data_ptr()
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        if dtype is not None and a.dtype != dtype[0]:
            raise TypeError("expected numpy dtype %s, got %s" % (dtype[0], a.dtype))
        if a.size < min_elems:
            raise ValueError("array too small: %d < %d" % (a.size, min_elems))
        return a.ctypes.data
    raise TypeError("expected numpy array or torch tensor, got %r" % type(a))


def _dt(np_dtype, torch_name: str):
    return (np.dtype(np_dtype), torch_name)  # the torch dtype is resolved by _ptr, only for torch tensors


def game_info(board_size: int) -> _lib.GameInfo:
    """Game constants + the reference's board_size range check (twixt.cc:134-145); needs no GPU."""
    lib = _lib.load()
    info = _lib.GameInfo()
    rc = lib.twixt_game_info_for(int(board_size), C.byref(info))
    if rc != 0:
        raise SpielFatalError(lib.twixt_last_error().decode())
    return info


class TwixTBatch:
    def __init__(self, board_size: int = 8, num_envs: int = 1, device: int = 0, seed: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.info = game_info(board_size)  # raises the reference's message for a bad size
        rc = self._lib.twixt_create(int(board_size), int(num_envs), int(device), int(seed) & (2**64 - 1),
                                    C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            self._raise(rc)
        self.board_size = int(board_size)
        self.num_envs = int(num_envs)
        self.device = int(device)
        self.max_legal_actions = self.info.max_legal_actions
        self.obs_shape = tuple(self.info.obs_shape)
        self.record_words = self.info.record_words
        self.max_game_length = self.info.max_game_length

    # -- plumbing ----------------------------------------------------------
    def _raise(self, rc: int):
        msg = self._lib.twixt_last_error().decode()
        if rc == _lib.EILLEGAL or (rc == _lib.EINVAL and msg.startswith("board_size out of range")):
            raise SpielFatalError(msg)
        if rc == _lib.EINVAL:
            raise ValueError(msg)
        if rc == _lib.ENOMEM:
            raise MemoryError(msg)
        raise TwixTCudaError(msg)

    def _check(self, rc: int):
        if rc != 0:
            self._raise(rc)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.twixt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _range(self, first: int, count: Optional[int]) -> Tuple[int, int]:
        if count is None:
            count = self.num_envs - first
        return int(first), int(count)

    def set_stream(self, cuda_stream: int):
        self._check(self._lib.twixt_set_stream(self._h, int(cuda_stream)))

    def use_torch_stream(self):
        """Launch on torch's current stream so torch.cuda.Event timing brackets our kernels."""
        self.set_stream(_torch().cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        self._check(self._lib.twixt_synchronize(self._h))

    def set_seed(self, seed: int):
        self._check(self._lib.twixt_set_seed(self._h, int(seed) & (2**64 - 1)))

    def set_stream_base(self, base: int):
        self._check(self._lib.twixt_set_stream_base(self._h, int(base) & (2**64 - 1)))

    # -- State surface -----------------------------------------------------
    def reset(self, first: int = 0, count: Optional[int] = None):
        first, count = self._range(first, count)
        self._check(self._lib.twixt_reset(self._h, first, count))

    def clone(self, src_first: int, dst_first: int, count: int = 1):
        self._check(self._lib.twixt_clone(self._h, int(src_first), int(dst_first), int(count)))

    def clone_gather(self, src_ids, dst_first: int):
        n = src_ids.numel() if _is_torch(src_ids) else int(np.asarray(src_ids).size)
        if not _is_torch(src_ids):
            src_ids = np.ascontiguousarray(src_ids, dtype=np.int64)
        self._check(self._lib.twixt_clone_gather(self._h, _ptr(src_ids, _dt(np.int64, "int64")), int(dst_first), n))

    def clone_from(self, dst_first: int, src: "TwixTBatch", src_first: int, count: int = 1):
        self._check(self._lib.twixt_clone_from(self._h, int(dst_first), src._h, int(src_first), int(count)))

    def legal_actions(self, first: int = 0, count: Optional[int] = None, out_actions=None, out_counts=None):
        """Returns (actions [count, max_legal_actions], counts [count]); rows are ascending, valid up to counts."""
        first, count = self._range(first, count)
        if out_actions is None:
            out_actions = np.full((count, self.max_legal_actions), -1, dtype=np.int64)
        if out_counts is None:
            out_counts = np.zeros(count, dtype=np.int32)
        if _is_torch(out_actions):
            elem = out_actions.element_size()
            stride = out_actions.shape[1] if out_actions.dim() == 2 else self.max_legal_actions
        else:
            elem = out_actions.dtype.itemsize
            stride = out_actions.shape[1] if out_actions.ndim == 2 else self.max_legal_actions
        self._check(self._lib.twixt_legal_actions(self._h, first, count, _ptr(out_actions), elem, stride,
                                                  _ptr(out_counts, _dt(np.int32, "int32"), count)))
        return out_actions, out_counts

    def legal_mask(self, first: int = 0, count: Optional[int] = None, out=None):
        first, count = self._range(first, count)
        if out is None:
            out = np.zeros((count, self.board_size * self.board_size), dtype=np.uint8)
        self._check(self._lib.twixt_legal_mask(self._h, first, count,
                                               _ptr(out, _dt(np.uint8, "uint8"), count * self.board_size ** 2)))
        return out

    def apply(self, actions, first: int = 0, out_status=None, raise_on_illegal: bool = True):
        """ApplyAction for envs first..first+len(actions); a negative action skips its env."""
        if not _is_torch(actions):
            actions = np.ascontiguousarray(actions, dtype=np.int32)
            count = int(actions.size)
        else:
            count = int(actions.numel())
        if out_status is None and not _is_torch(actions):
            out_status = np.zeros(count, dtype=np.int32)
        rc = self._lib.twixt_apply(self._h, int(first), count, _ptr(actions, _dt(np.int32, "int32")),
                                   _ptr(out_status, _dt(np.int32, "int32"), count))
        if rc == _lib.EILLEGAL and not raise_on_illegal:
            return out_status
        self._check(rc)
        return out_status

    def step(self, env: int, action: int = -1, out_legal=None):
        """One State step for an unbatched caller (twixt_step): `action` >= 0 is applied (illegal ->
        SpielFatalError "Not a legal action: N"), STEP_QUERY (-1) applies nothing, STEP_RESET (-2) resets the
        env first; returns (status, current_player, is_terminal, returns [2], legal actions int64 ascending)
        of the resulting state -- one kernel launch for everything the caller asks between two moves."""
        if out_legal is None:
            out_legal = np.empty(self.max_legal_actions, dtype=np.int64)
        res = _lib.StepResult()
        self._check(self._lib.twixt_step(self._h, int(env), int(action), C.byref(res),
                                         _ptr(out_legal, _dt(np.int64, "int64"), self.max_legal_actions)))
        return (int(res.status), int(res.current_player), bool(res.is_terminal), [float(res.returns[0]), float(res.returns[1])],
                out_legal[:res.num_legal])

    def current_player(self, first: int = 0, count: Optional[int] = None, out=None):
        first, count = self._range(first, count)
        if out is None:
            out = np.zeros(count, dtype=np.int8)
        self._check(self._lib.twixt_current_player(self._h, first, count, _ptr(out, _dt(np.int8, "int8"), count)))
        return out

    def is_terminal(self, first: int = 0, count: Optional[int] = None, out=None):
        first, count = self._range(first, count)
        if out is None:
            out = np.zeros(count, dtype=np.uint8)
        self._check(self._lib.twixt_is_terminal(self._h, first, count, _ptr(out, _dt(np.uint8, "uint8"), count)))
        return out

    def returns(self, first: int = 0, count: Optional[int] = None, out=None):
        first, count = self._range(first, count)
        if out is None:
            out = np.zeros((count, 2), dtype=np.float32)
        self._check(self._lib.twixt_returns(self._h, first, count, _ptr(out, _dt(np.float32, "float32"), 2 * count)))
        return out

    def observation(self, first: int = 0, count: Optional[int] = None, out=None):
        first, count = self._range(first, count)
        if out is None:
            out = np.empty((count,) + self.obs_shape, dtype=np.float32)
        self._check(self._lib.twixt_observation(self._h, first, count,
                                                _ptr(out, _dt(np.float32, "float32"), count * self.info.obs_size)))
        return out

    def observation_and_mask(self, first: int = 0, count: Optional[int] = None, out_obs=None, out_mask=None):
        """ObservationTensor + LegalActionsMask in one pass over the records (twixt_observation_and_mask)."""
        first, count = self._range(first, count)
        if out_obs is None:
            out_obs = np.empty((count,) + self.obs_shape, dtype=np.float32)
        if out_mask is None:
            out_mask = np.zeros((count, self.board_size * self.board_size), dtype=np.uint8)
        self._check(self._lib.twixt_observation_and_mask(
            self._h, first, count, _ptr(out_obs, _dt(np.float32, "float32"), count * self.info.obs_size),
            _ptr(out_mask, _dt(np.uint8, "uint8"), count * self.board_size ** 2)))
        return out_obs, out_mask

    def replay(self, actions, first: int = 0, lengths=None, out_applied=None, raise_on_illegal: bool = True):
        """Applies a whole action history per env in one launch (twixt_replay).

        actions: [count, T] int32 (numpy or torch), rows padded with negative entries or cut by `lengths`.
        Returns out_applied [count] int32 = moves made per env."""
        if _is_torch(actions):
            count, stride = int(actions.shape[0]), int(actions.shape[1])
        else:
            actions = np.ascontiguousarray(actions, dtype=np.int32)
            if actions.ndim != 2:
                raise ValueError("actions must be [count, T]")
            count, stride = actions.shape
        if lengths is not None and not _is_torch(lengths):
            lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        if out_applied is None and not _is_torch(actions):
            out_applied = np.zeros(count, dtype=np.int32)
        rc = self._lib.twixt_replay(self._h, int(first), count, _ptr(actions, _dt(np.int32, "int32")), stride,
                                    _ptr(lengths, _dt(np.int32, "int32"), count) if lengths is not None else 0,
                                    _ptr(out_applied, _dt(np.int32, "int32"), count) if out_applied is not None else 0)
        if rc == _lib.EILLEGAL and not raise_on_illegal:
            return out_applied
        self._check(rc)
        return out_applied

    def playout(self, first: int = 0, count: Optional[int] = None, max_plies: Optional[int] = None, stream_ids=None,
                out_returns=None, out_lengths=None, out_actions=None, want_returns: bool = True,
                want_lengths: bool = True, trace: bool = False):
        """Fused random playout of every env in the range from its current state.

        Returns (returns [count,2] f32, lengths [count] i32, actions [trace_plies,count] u16 or None).
        """
        first, count = self._range(first, count)
        if max_plies is None:
            max_plies = self.max_game_length
        if want_returns and out_returns is None:
            out_returns = np.zeros((count, 2), dtype=np.float32)
        if want_lengths and out_lengths is None:
            out_lengths = np.zeros(count, dtype=np.int32)
        trace_plies = 0
        if trace and out_actions is None:
            out_actions = np.zeros((min(max_plies, self.max_game_length), count), dtype=np.uint16)
        if out_actions is not None:
            trace_plies = int(out_actions.shape[0])
        if stream_ids is not None and not _is_torch(stream_ids):
            stream_ids = np.ascontiguousarray(stream_ids, dtype=np.uint64)
        self._check(self._lib.twixt_playout(
            self._h, first, count, int(max_plies),
            _ptr(stream_ids) if stream_ids is not None else 0,
            _ptr(out_returns, _dt(np.float32, "float32"), 2 * count) if out_returns is not None else 0,
            _ptr(out_lengths, _dt(np.int32, "int32"), count) if out_lengths is not None else 0,
            _ptr(out_actions, _dt(np.uint16, "uint16") if not _is_torch(out_actions) else None,
                 trace_plies * count) if out_actions is not None else 0,
            trace_plies))
        return out_returns, out_lengths, out_actions

    def export_state(self, first: int = 0, count: Optional[int] = None, out=None):
        first, count = self._range(first, count)
        if out is None:
            out = np.zeros((count, self.record_words), dtype=np.uint32)
        self._check(self._lib.twixt_export_state(self._h, first, count, _ptr(out)))
        return out

    def set_validation(self, enabled: bool):
        """import_state validates records by default (twixt_import_state); off = trusted records."""
        self._check(self._lib.twixt_set_validation(self._h, 1 if enabled else 0))

    def import_state(self, records, first: int = 0):
        if not _is_torch(records):
            records = np.ascontiguousarray(records, dtype=np.uint32)
            count = records.size // self.record_words
        else:
            count = records.numel() // self.record_words
        self._check(self._lib.twixt_import_state(self._h, int(first), int(count), _ptr(records)))

    def stats(self) -> dict:
        s = _lib.Stats()
        self._check(self._lib.twixt_get_stats(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in _lib.Stats._fields_}

    def stats_reset(self):
        self._check(self._lib.twixt_stats_reset(self._h))
