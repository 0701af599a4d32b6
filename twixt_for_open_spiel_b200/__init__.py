"""twixt_for_open_spiel_b200 -- B200-native batched TwixT engine.

Host-side mirror of the reference's open_spiel plug-in surface over
libtwixt_b200.so (hand-written sm_100a kernels, C ABI in include/twixt_b200.h).
"""
from .batch import SpielFatalError, TwixTBatch, TwixTCudaError, game_info  # noqa: F401
from .spiel import TwixTGame, TwixTState, load_game  # noqa: F401

__all__ = ["TwixTBatch", "TwixTGame", "TwixTState", "load_game", "game_info", "SpielFatalError", "TwixTCudaError"]
