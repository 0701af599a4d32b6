"""open_spiel-shaped adapter: `TwixTGame` / `TwixTState` look-alikes of the
reference's classes (twixt.h:31-146, twixt.cc:35-145) whose every method is a
call into the batched CUDA engine with a batch of ONE env.

This is the drop-in surface for code written against open_spiel's Python API
(`pyspiel.load_game("twixt(board_size=12)")`, `state.legal_actions()`,
`state.apply_action(a)`, ...): same method names, argument meaning and error
text.  It exists so existing single-game drivers (random playout example, MCTS
with rollouts) run unchanged; throughput work should use `TwixTBatch` directly.

States of one game share a device pool (one TwixTBatch per `pool_size` envs);
a state is a slot in that pool, `clone()` is a device-side record copy.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional

import numpy as np

from .batch import SpielFatalError, TwixTBatch, game_info

K_TERMINAL_PLAYER_ID = -4
DEFAULT_BOARD_SIZE = 8           # twixtboard.h:34
DEFAULT_ANSI_COLOR_OUTPUT = True  # twixtboard.h:36
_KNOWN_PARAMS = ("ansi_color_output", "board_size")  # twixt.cc:50-51


class GameType:
    """kGameType, twixt.cc:35-52."""
    short_name = "twixt"
    long_name = "TwixT"
    dynamics = "SEQUENTIAL"
    chance_mode = "DETERMINISTIC"
    information = "PERFECT_INFORMATION"
    utility = "ZERO_SUM"
    reward_model = "TERMINAL"
    max_num_players = 2
    min_num_players = 2
    provides_information_state_string = True
    provides_information_state_tensor = False
    provides_observation_string = True
    provides_observation_tensor = True
    parameter_specification = list(_KNOWN_PARAMS)


def _parse_game_string(s: str) -> Dict[str, object]:
    m = re.fullmatch(r"\s*(\w+)\s*(?:\((.*)\))?\s*", s)
    if not m or m.group(1) != "twixt":
        raise SpielFatalError("Unknown game '%s'. Available games are:\ntwixt" % s)
    params: Dict[str, object] = {}
    body = m.group(2)
    if body:
        for item in body.split(","):
            if not item.strip():
                continue
            k, _, v = item.partition("=")
            k, v = k.strip(), v.strip()
            if v.lower() in ("true", "false"):
                params[k] = v.lower() == "true"
            else:
                params[k] = int(v)
    return params


class TwixTGame:
    def __init__(self, params: Optional[Dict[str, object]] = None, device: int = 0, pool_size: int = 256,
                 seed: int = 0):
        params = dict(params or {})
        for k in params:
            if k not in _KNOWN_PARAMS:  # upstream Game ctor; text pinned by twixt_test.cc:88-89
                raise SpielFatalError("Unknown parameter '%s'. Available parameters are: %s"
                                      % (k, ", ".join(_KNOWN_PARAMS)))
        self._board_size = int(params.get("board_size", DEFAULT_BOARD_SIZE))
        self._ansi = bool(params.get("ansi_color_output", DEFAULT_ANSI_COLOR_OUTPUT))
        self._info = game_info(self._board_size)  # "board_size out of range [5..24]: N" (twixt.cc:139-144)
        self._params = params
        self._device = device
        self._pool_size = pool_size
        self._seed = seed
        self._pools: List[TwixTBatch] = []
        self._free: List[tuple] = []

    # -- Game surface (twixt.h:116-141) -------------------------------------
    def get_type(self):
        return GameType

    def get_parameters(self):
        return {"ansi_color_output": self._ansi, "board_size": self._board_size}

    def new_initial_state(self) -> "TwixTState":
        pool, idx = self._take_slot()
        return TwixTState(self, pool, idx, step=STEP_RESET)

    def num_distinct_actions(self) -> int:
        return self._info.num_distinct_actions

    def num_players(self) -> int:
        return 2

    def min_utility(self) -> float:
        return -1.0

    def max_utility(self) -> float:
        return 1.0

    def utility_sum(self) -> float:
        return 0.0

    def observation_tensor_shape(self) -> List[int]:
        return list(self._info.obs_shape)

    def observation_tensor_size(self) -> int:
        return self._info.obs_size

    def max_game_length(self) -> int:
        return self._info.max_game_length

    def board_size(self) -> int:
        return self._board_size

    def ansi_color_output(self) -> bool:
        return self._ansi

    def __str__(self):
        if not self._params:
            return "twixt()"
        return "twixt(%s)" % ",".join("%s=%s" % (k, self._params[k]) for k in sorted(self._params))

    # -- serialisation (upstream serialises a state as its action history) --------
    def deserialize_state(self, text: str) -> "TwixTState":
        """Inverse of TwixTState.serialize(): the whole action history is replayed on the device by ONE
        launch (twixt_replay); an illegal action in it raises "Not a legal action: N" like ApplyAction."""
        history = [int(tok) for tok in text.replace(",", " ").split()]
        for a in history:
            if a < 0 or a > 0x7FFFFFFF:
                raise SpielFatalError("Not a legal action: %d" % a)
        state = self.new_initial_state()
        if history:
            state._pool.replay(np.asarray([history], dtype=np.int32), state._idx)
            state._history = history
            state._step(STEP_QUERY)
        return state

    def new_state_from_record(self, record) -> "TwixTState":
        """A state from an exported packed record (twixt_import_state); its history is unknown."""
        pool, idx = self._take_slot()
        pool.import_state(np.ascontiguousarray(record, dtype=np.uint32), idx)
        return TwixTState(self, pool, idx, history=None, step=STEP_QUERY)

    # -- slot pool -------------------------------------------------------------
    def _take_slot(self):
        if not self._free:
            pool = TwixTBatch(self._board_size, self._pool_size, self._device, self._seed)
            self._pools.append(pool)
            self._free.extend((pool, i) for i in reversed(range(self._pool_size)))
        return self._free.pop()

    def _give_slot(self, pool, idx):
        self._free.append((pool, idx))


STEP_QUERY, STEP_RESET = -1, -2  # include/twixt_b200.h


class TwixTState:
    """Like the C++ adapter (adapter/twixt_b200_game.h) a state keeps the answer of its last twixt_step --
    player, terminal flag, returns, legal actions -- so apply_action is one kernel launch and the four query
    methods touch no GPU."""

    def __init__(self, game: TwixTGame, pool: TwixTBatch, idx: int, history: Optional[List[int]] = None,
                 step: Optional[int] = None, now=None):
        self._game = game
        self._pool = pool
        self._idx = idx
        self._history: List[int] = list(history or [])
        self._now = now
        if step is not None:
            self._step(step)

    def _step(self, action: int) -> None:
        _, player, terminal, returns, legal = self._pool.step(self._idx, action)
        self._now = (player, terminal, returns, [int(a) for a in legal])

    def __del__(self):
        try:
            self._game._give_slot(self._pool, self._idx)
        except Exception:
            pass

    # -- State surface (twixt.h:31-112) -----------------------------------------
    def current_player(self) -> int:
        return self._now[0]

    def is_terminal(self) -> bool:
        return self._now[1]

    def returns(self) -> List[float]:
        return list(self._now[2])

    def rewards(self) -> List[float]:
        return self.returns()

    def player_return(self, player: int) -> float:
        return self.returns()[player]

    def legal_actions(self, player: Optional[int] = None) -> List[int]:
        return list(self._now[3])

    def legal_actions_mask(self, player: Optional[int] = None) -> List[int]:
        return [int(v) for v in self._pool.legal_mask(self._idx, 1)[0]]

    def apply_action(self, action: int) -> None:
        a = int(action)
        if a < 0 or a > 0x7FFFFFFF:
            raise SpielFatalError("Not a legal action: %d" % a)
        self._step(a)  # raises "Not a legal action: N" and leaves the state as it was
        self._history.append(a)

    def clone(self) -> "TwixTState":
        pool, idx = self._game._take_slot()
        if pool is self._pool:
            pool.clone(self._idx, idx, 1)
        else:
            pool.clone_from(idx, self._pool, self._idx, 1)
        return TwixTState(self._game, pool, idx, self._history, now=self._now)

    def observation_tensor(self, player: int = 0) -> List[float]:
        if player < 0 or player >= 2:  # SPIEL_CHECK_GE / _LT, twixt.cc:103-104
            raise SpielFatalError("ObservationTensor: player %d out of range" % player)
        return self._pool.observation(self._idx, 1).reshape(-1).tolist()

    def observation_array(self, player: int = 0) -> np.ndarray:
        if player < 0 or player >= 2:
            raise SpielFatalError("ObservationTensor: player %d out of range" % player)
        return self._pool.observation(self._idx, 1)[0]

    def action_to_string(self, player: int, action: int) -> str:
        """twixt.cc:67-74: 'x'|'o' + column letter + (n - y)."""
        n = self._game.board_size()
        x, y = divmod(int(action), n)
        return ("x" if player == 0 else "o") + chr(ord("a") + x) + str(n - y)

    def history(self) -> List[int]:
        return list(self._history)

    def history_str(self) -> str:
        return ", ".join(str(a) for a in self._history)

    def serialize(self) -> str:
        """One action per line, like upstream State::Serialize() for a game without chance nodes."""
        return "\n".join(str(a) for a in self._history)

    def is_chance_node(self) -> bool:
        return False

    def is_simultaneous_node(self) -> bool:
        return False

    def undo_action(self, player: int, action: int) -> None:
        """An empty stub in the reference as well (twixt.h:84)."""

    def get_game(self) -> TwixTGame:
        return self._game

    def export_record(self) -> np.ndarray:
        return self._pool.export_state(self._idx, 1)[0]

    def to_string(self) -> str:
        from .render import board_to_string
        return board_to_string(self.export_record(), self._game.board_size(), self._game.ansi_color_output())

    def __str__(self) -> str:
        return self.to_string()

    def observation_string(self, player: int = 0) -> str:
        if player < 0 or player >= 2:
            raise SpielFatalError("ObservationString: player %d out of range" % player)
        return self.to_string()

    def information_state_string(self, player: int = 0) -> str:
        if player < 0 or player >= 2:
            raise SpielFatalError("InformationStateString: player %d out of range" % player)
        return self.to_string()


def load_game(game_string: str = "twixt", params: Optional[Dict[str, object]] = None, **kwargs) -> TwixTGame:
    """pyspiel.load_game look-alike: load_game("twixt(board_size=12,ansi_color_output=False)")."""
    p = _parse_game_string(game_string)
    if params:
        p.update(params)
    return TwixTGame(p, **kwargs)
