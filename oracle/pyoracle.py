"""TEST INFRASTRUCTURE ONLY (oracle/).

ctypes bindings for the two CPU checkers:

* ``Oracle*``  -> oracle/liboracle.so, the plain-C restatement
  (oracle/twixt_oracle.c), always buildable (gcc only);
* ``Ref*``     -> oracle/_ref/libtwixt_ref.so, the UNMODIFIED reference
  (/root/reference/open_spiel/games/twixt/*.cc) compiled against the header
  shim; only buildable where /root/reference exists, but the built file
  travels to the GPU box.

Both expose the same small interface so tests can run one differential loop
over either.  Nothing under twixt_for_open_spiel_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtwixt_ref.so")
REF_BENCH = os.path.join(HERE, "_ref", "ref_bench")
REFERENCE_ROOT = "/root/reference"

TERMINAL_PLAYER = -4
_ERRCAP = 256


def build(verbose: bool = False) -> None:
    """(Re)build liboracle.so always, and oracle/_ref when /root/reference exists."""
    cmd = ["make", "-C", HERE, "all"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed")


def _ensure_oracle() -> None:
    src = os.path.join(HERE, "twixt_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        build()


def have_reference() -> bool:
    return os.path.exists(REF_SO)


_oracle_lib = None
_ref_lib = None


def oracle_lib() -> C.CDLL:
    global _oracle_lib
    if _oracle_lib is None:
        _ensure_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.oracle_game_new.restype = C.c_void_p
        lib.oracle_game_new.argtypes = [C.c_int, C.c_char_p, C.c_int]
        lib.oracle_game_free.argtypes = [C.c_void_p]
        for name in ("oracle_game_board_size", "oracle_num_distinct_actions", "oracle_max_game_length",
                     "oracle_observation_size", "oracle_record_words"):
            getattr(lib, name).restype = C.c_int
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.oracle_blockers.restype = C.c_int
        lib.oracle_blockers.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.oracle_state_new.restype = C.c_void_p
        lib.oracle_state_new.argtypes = [C.c_void_p]
        lib.oracle_state_clone.restype = C.c_void_p
        lib.oracle_state_clone.argtypes = [C.c_void_p]
        lib.oracle_state_copy.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_state_free.argtypes = [C.c_void_p]
        lib.oracle_legal_actions.restype = C.c_int
        lib.oracle_legal_actions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_legal_list_of.restype = C.c_int
        lib.oracle_legal_list_of.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_apply.restype = C.c_int
        lib.oracle_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int]
        lib.oracle_current_player.restype = C.c_int
        lib.oracle_current_player.argtypes = [C.c_void_p]
        lib.oracle_is_terminal.restype = C.c_int
        lib.oracle_is_terminal.argtypes = [C.c_void_p]
        lib.oracle_returns.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_observation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_board_header.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_export_cells.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_export_record.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_replay.restype = C.c_int
        lib.oracle_replay.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_playout_philox.restype = C.c_int
        lib.oracle_playout_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        lib.oracle_bench_playouts.restype = C.c_int64
        lib.oracle_bench_playouts.argtypes = [C.c_void_p, C.c_double, C.c_uint64, C.c_void_p, C.c_void_p]
        lib.oracle_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _oracle_lib = lib
    return _oracle_lib


def ref_lib() -> C.CDLL:
    global _ref_lib
    if _ref_lib is None:
        if not os.path.exists(REF_SO):
            if os.path.isdir(REFERENCE_ROOT):
                build()
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO)
        lib = C.CDLL(REF_SO)
        lib.ref_game_new.restype = C.c_void_p
        lib.ref_game_new.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
        lib.ref_game_new_with_param.restype = C.c_int
        lib.ref_game_new_with_param.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        lib.ref_game_free.argtypes = [C.c_void_p]
        lib.ref_game_info.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_game_utils.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_state_new.restype = C.c_void_p
        lib.ref_state_new.argtypes = [C.c_void_p]
        lib.ref_state_clone.restype = C.c_void_p
        lib.ref_state_clone.argtypes = [C.c_void_p]
        lib.ref_state_free.argtypes = [C.c_void_p]
        lib.ref_legal_actions.restype = C.c_int
        lib.ref_legal_actions.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_legal_list_of.restype = C.c_int
        lib.ref_legal_list_of.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.ref_apply.restype = C.c_int
        lib.ref_apply.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int]
        lib.ref_current_player.restype = C.c_int
        lib.ref_current_player.argtypes = [C.c_void_p]
        lib.ref_is_terminal.restype = C.c_int
        lib.ref_is_terminal.argtypes = [C.c_void_p]
        lib.ref_returns.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_observation.restype = C.c_int
        lib.ref_observation.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        lib.ref_to_string.restype = C.c_int
        lib.ref_to_string.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.ref_action_to_string.restype = C.c_int
        lib.ref_action_to_string.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_char_p, C.c_int]
        lib.ref_board_header.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_export_cells.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_replay.restype = C.c_int
        lib.ref_replay.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_blockers.restype = C.c_int
        lib.ref_blockers.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.ref_playout_philox.restype = C.c_int
        lib.ref_playout_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        _ref_lib = lib
    return _ref_lib


class SpielError(RuntimeError):
    """What upstream raises through SpielFatalError."""


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _StateBase:
    n: int

    def observation_shape(self):
        return (12, self.n, self.n - 2)


class OracleGame:
    def __init__(self, board_size: int = 8):
        self.lib = oracle_lib()
        err = C.create_string_buffer(_ERRCAP)
        self.h = self.lib.oracle_game_new(board_size, err, _ERRCAP)
        if not self.h:
            raise SpielError(err.value.decode())
        self.n = board_size

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_game_free(self.h)
            self.h = None

    def num_distinct_actions(self) -> int:
        return self.lib.oracle_num_distinct_actions(self.h)

    def max_game_length(self) -> int:
        return self.lib.oracle_max_game_length(self.h)

    def observation_tensor_shape(self):
        return [12, self.n, self.n - 2]

    def record_words(self) -> int:
        return self.lib.oracle_record_words(self.h)

    def blockers(self, x: int, y: int, d: int):
        out = np.zeros(27, dtype=np.int32)
        c = self.lib.oracle_blockers(self.h, x, y, d, _ptr(out))
        return [tuple(int(v) for v in out[3 * i:3 * i + 3]) for i in range(c)]

    def new_initial_state(self) -> "OracleState":
        return OracleState(self, self.lib.oracle_state_new(self.h))


class OracleState(_StateBase):
    def __init__(self, game: OracleGame, handle):
        self.game = game
        self.lib = game.lib
        self.h = handle
        self.n = game.n

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_state_free(self.h)
            self.h = None

    def clone(self) -> "OracleState":
        return OracleState(self.game, self.lib.oracle_state_clone(self.h))

    def legal_actions(self) -> List[int]:
        out = np.zeros(self.n * self.n, dtype=np.int64)
        c = self.lib.oracle_legal_actions(self.game.h, self.h, _ptr(out))
        return out[:c].tolist()

    def legal_list_of(self, player: int) -> List[int]:
        out = np.zeros(self.n * self.n, dtype=np.int64)
        c = self.lib.oracle_legal_list_of(self.game.h, self.h, player, _ptr(out))
        return out[:c].tolist()

    def apply_action(self, action: int) -> None:
        err = C.create_string_buffer(_ERRCAP)
        if self.lib.oracle_apply(self.game.h, self.h, int(action), err, _ERRCAP) != 0:
            raise SpielError(err.value.decode())

    def current_player(self) -> int:
        return self.lib.oracle_current_player(self.h)

    def is_terminal(self) -> bool:
        return bool(self.lib.oracle_is_terminal(self.h))

    def returns(self) -> List[float]:
        out = np.zeros(2, dtype=np.float64)
        self.lib.oracle_returns(self.h, _ptr(out))
        return out.tolist()

    def observation_tensor(self, player: int = 0) -> np.ndarray:
        if player < 0 or player >= 2:
            raise SpielError("CHECK failed: player")
        out = np.empty(12 * self.n * (self.n - 2), dtype=np.float32)
        self.lib.oracle_observation(self.game.h, self.h, _ptr(out))
        return out

    def board_header(self) -> List[int]:
        out = np.zeros(5, dtype=np.int32)
        self.lib.oracle_board_header(self.h, _ptr(out))
        return out.tolist()

    def export_cells(self) -> np.ndarray:
        out = np.zeros(4 * self.n * self.n, dtype=np.int32)
        self.lib.oracle_export_cells(self.game.h, self.h, _ptr(out))
        return out.reshape(self.n * self.n, 4)

    def export_record(self) -> np.ndarray:
        out = np.zeros(self.game.record_words(), dtype=np.uint32)
        self.lib.oracle_export_record(self.game.h, self.h, _ptr(out))
        return out

    def replay(self, actions: Sequence[int]) -> int:
        a = np.asarray(actions, dtype=np.int64)
        return self.lib.oracle_replay(self.game.h, self.h, _ptr(a), len(a))

    def playout_philox(self, seed: int, stream: int, max_plies: int = 1 << 30) -> List[int]:
        out = np.zeros(self.n * self.n, dtype=np.int64)
        c = self.lib.oracle_playout_philox(self.game.h, self.h, seed, stream, max_plies, _ptr(out))
        return out[:c].tolist()


class RefGame:
    """The unmodified reference.  Keep only ONE board size alive at a time."""

    def __init__(self, board_size: int = 8, ansi: bool = True):
        self.lib = ref_lib()
        err = C.create_string_buffer(_ERRCAP)
        self.h = self.lib.ref_game_new(board_size, 1 if ansi else 0, err, _ERRCAP)
        if not self.h:
            raise SpielError(err.value.decode())
        self.n = board_size

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_game_free(self.h)
            self.h = None

    def info(self) -> List[int]:
        out = np.zeros(6, dtype=np.int32)
        self.lib.ref_game_info(self.h, _ptr(out))
        return out.tolist()

    def num_distinct_actions(self) -> int:
        return self.info()[0]

    def max_game_length(self) -> int:
        return self.info()[1]

    def utilities(self) -> List[float]:
        out = np.zeros(3, dtype=np.float64)
        self.lib.ref_game_utils(self.h, _ptr(out))
        return out.tolist()

    def new_initial_state(self) -> "RefState":
        return RefState(self, self.lib.ref_state_new(self.h))

    def blockers(self, x: int, y: int, d: int):
        """BlockerMap entries of link (x,y,d); valid only after a state of this size was constructed."""
        out = np.zeros(3 * 32, dtype=np.int32)
        c = self.lib.ref_blockers(x, y, d, _ptr(out), 32)
        return [tuple(int(v) for v in out[3 * i:3 * i + 3]) for i in range(c)]


class RefState(_StateBase):
    def __init__(self, game: RefGame, handle):
        self.game = game
        self.lib = game.lib
        self.h = handle
        self.n = game.n

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_state_free(self.h)
            self.h = None

    def clone(self) -> "RefState":
        return RefState(self.game, self.lib.ref_state_clone(self.h))

    def legal_actions(self) -> List[int]:
        out = np.zeros(self.n * self.n, dtype=np.int64)
        c = self.lib.ref_legal_actions(self.h, _ptr(out), len(out))
        return out[:c].tolist()

    def legal_list_of(self, player: int) -> List[int]:
        out = np.zeros(self.n * self.n, dtype=np.int64)
        c = self.lib.ref_legal_list_of(self.h, player, _ptr(out), len(out))
        return out[:c].tolist()

    def apply_action(self, action: int) -> None:
        err = C.create_string_buffer(_ERRCAP)
        if self.lib.ref_apply(self.h, int(action), err, _ERRCAP) != 0:
            raise SpielError(err.value.decode())

    def current_player(self) -> int:
        return self.lib.ref_current_player(self.h)

    def is_terminal(self) -> bool:
        return bool(self.lib.ref_is_terminal(self.h))

    def returns(self) -> List[float]:
        out = np.zeros(2, dtype=np.float64)
        self.lib.ref_returns(self.h, _ptr(out))
        return out.tolist()

    def observation_tensor(self, player: int = 0) -> np.ndarray:
        out = np.empty(12 * self.n * (self.n - 2), dtype=np.float32)
        err = C.create_string_buffer(_ERRCAP)
        if self.lib.ref_observation(self.h, player, _ptr(out), len(out), err, _ERRCAP) != 0:
            raise SpielError(err.value.decode())
        return out

    def to_string(self) -> str:
        ln = self.lib.ref_to_string(self.h, None, 0)
        buf = C.create_string_buffer(ln + 1)
        self.lib.ref_to_string(self.h, buf, ln + 1)
        return buf.raw[:ln].decode("utf-8")

    def action_to_string(self, player: int, action: int) -> str:
        buf = C.create_string_buffer(32)
        self.lib.ref_action_to_string(self.h, player, int(action), buf, 32)
        return buf.value.decode()

    def board_header(self) -> List[int]:
        out = np.zeros(5, dtype=np.int32)
        self.lib.ref_board_header(self.h, _ptr(out))
        return out.tolist()

    def export_cells(self) -> np.ndarray:
        out = np.zeros(4 * self.n * self.n, dtype=np.int32)
        self.lib.ref_export_cells(self.h, _ptr(out))
        return out.reshape(self.n * self.n, 4)

    def replay(self, actions: Sequence[int]) -> int:
        a = np.asarray(actions, dtype=np.int64)
        return self.lib.ref_replay(self.h, _ptr(a), len(a))

    def export_record(self) -> np.ndarray:
        """The packed state record (include/twixt_b200.h) of the REFERENCE's board, assembled here from its
        cell internals (ref_export_cells) and header -- the same mapping as oracle_export_record in
        twixt_oracle.c -- so a CUDA record can be compared with the compiled reference directly."""
        n = self.n
        cells = self.export_cells().reshape(n, n, 4).astype(np.int64)  # [x, y, field]
        hdr = self.board_header()  # move_counter, swapped, result, move_one x, y
        words = (4 + 9 * n + 3) // 4 * 4
        rec = np.zeros(words, dtype=np.uint32)
        color, links, blocked, flags = (cells[:, :, k] for k in range(4))
        bit = (np.int64(1) << np.arange(n, dtype=np.int64))[None, :]  # bit y of a column word

        def column_words(sel):  # [x, y] bool -> [x] uint32
            return (sel * bit).sum(axis=1).astype(np.uint32)

        peg = color < 2
        own = np.where(peg, color, 0)
        planes = [column_words(color == 0), column_words(color == 1)]
        planes += [column_words(peg & (((links >> d) & 1) == 1)) for d in range(4)]
        planes.append(column_words(peg & (((flags >> (2 * own)) & 1) == 1)))      # owner's start line
        planes.append(column_words(peg & (((flags >> (2 * own + 1)) & 1) == 1)))  # owner's end line
        planes.append(column_words(peg & ((blocked & 15) != 0)))
        rec[4:4 + 9 * n] = np.concatenate(planes)
        empty = color == 2  # corners are 3 (off-board)
        rec[0] = hdr[0]
        rec[1] = hdr[2] | (hdr[1] << 2)
        rec[2] = 0xFFFFFFFF if hdr[0] == 0 else hdr[3] * n + hdr[4]
        rec[3] = int(empty[1:n - 1, :].sum()) | (int(empty[:, 1:n - 1].sum()) << 16)
        return rec

    def playout_philox(self, seed: int, stream: int, max_plies: int = 1 << 30) -> List[int]:
        out = np.zeros(self.n * self.n, dtype=np.int64)
        c = self.lib.ref_playout_philox(self.h, seed, stream, max_plies, _ptr(out))
        return out[:c].tolist()


def philox(ctr: Sequence[int], key: Sequence[int]) -> List[int]:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    oracle_lib().oracle_philox(_ptr(c), _ptr(k), _ptr(out))
    return [int(v) for v in out]


def run_ref_bench(board_size: int, workers: int, seconds: float, mode: str = "clone", seed: int = 1) -> Optional[dict]:
    """Run oracle/_ref/ref_bench (the compiled reference's CPU playout loop)."""
    import json

    if not os.path.exists(REF_BENCH):
        return None
    res = subprocess.run([REF_BENCH, str(board_size), str(workers), str(seconds), mode, str(seed)],
                         capture_output=True, text=True, timeout=max(120.0, seconds * 4 + 60))
    if res.returncode != 0:
        return None
    return json.loads(res.stdout.strip().splitlines()[-1])
