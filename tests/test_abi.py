"""CPU-side checks of the drop-in boundary: libtwixt_b200.so loads, exports every symbol
include/twixt_b200.h declares, answers the GPU-free calls, and FAILS LOUDLY without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "twixt_b200.h")


def declared_functions():
    with open(HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(twixt_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from twixt_for_open_spiel_b200 import _lib
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 25
    raw = C.CDLL(_lib.library_path())
    for name in names:
        assert hasattr(raw, name), "missing export: " + name
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert bound == set(names), (bound ^ set(names))
    assert b"sm_100a" in lib.twixt_version()


def test_library_is_sm100a_only():
    """One architecture, no PTX-JIT fallbacks for other GPUs."""
    import subprocess
    from twixt_for_open_spiel_b200 import _lib
    _lib.load()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_game_info_and_range_errors():
    from twixt_for_open_spiel_b200 import SpielFatalError, game_info
    for n in range(5, 25):
        info = game_info(n)
        assert info.num_distinct_actions == n * n
        assert info.max_game_length == n * n - 3
        assert list(info.obs_shape) == [12, n, n - 2] and info.obs_size == 12 * n * (n - 2)
        assert info.max_legal_actions == n * (n - 2)
        assert info.record_words == (4 + 9 * n + 3) // 4 * 4
        assert (info.min_utility, info.max_utility, info.utility_sum) == (-1.0, 1.0, 0.0)
    for bad in (30, 3, 4, 25, -1):
        with pytest.raises(SpielFatalError) as e:
            game_info(bad)
        assert str(e.value) == "board_size out of range [5..24]: %d" % bad


def test_adapter_parameter_handling_needs_no_gpu():
    from twixt_for_open_spiel_b200 import SpielFatalError, load_game
    g = load_game("twixt(board_size=12,ansi_color_output=False)")
    assert g.board_size() == 12 and g.ansi_color_output() is False
    assert g.num_distinct_actions() == 144 and g.observation_tensor_shape() == [12, 12, 10]
    assert g.max_game_length() == 141 and g.num_players() == 2
    assert str(load_game("twixt")) == "twixt()"
    with pytest.raises(SpielFatalError) as e:
        load_game("twixt(board_size=30)")
    assert str(e.value) == "board_size out of range [5..24]: 30"
    with pytest.raises(SpielFatalError) as e:
        load_game("twixt", {"bad_param": 3})
    assert str(e.value) == "Unknown parameter 'bad_param'. Available parameters are: ansi_color_output, board_size"


def test_no_cpu_fallback():
    """Without a CUDA device creating a batch must fail, not silently compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from twixt_for_open_spiel_b200 import TwixTBatch, TwixTCudaError, load_game
    with pytest.raises(TwixTCudaError, match="no CPU fallback"):
        TwixTBatch(8, 4)
    with pytest.raises(TwixTCudaError):
        load_game("twixt").new_initial_state()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "twixt_for_open_spiel_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".inc")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "pyoracle" not in text and "twixt_oracle" not in text and "liboracle" not in text, fn
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn


def test_header_is_plain_c_and_links():
    """The FFI boundary: a C99 program including only include/twixt_b200.h builds, links and runs."""
    import subprocess
    from twixt_for_open_spiel_b200 import _lib
    _lib.load()
    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.join(ROOT, "twixt_for_open_spiel_b200")
    exe = os.path.join(here, "_build", "abi_c_smoke")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(here, "abi_c_smoke.c"), "-o", exe, "-L", pkg, "-ltwixt_b200",
                           "-Wl,-rpath," + pkg])
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.startswith("OK"), (res.returncode, res.stdout, res.stderr)


def test_observation_index_arithmetic_is_exact():
    """The multiply-shift division used by the observation kernel (bit / (n-2)) is exact on its whole range."""
    for w in range(3, 23):
        m = ((1 << 20) + w - 1) // w
        for j in range(0, 12 * 24 * 22 + 64):
            assert (j * m) >> 20 == j // w and j * m < 2 ** 32


def test_ctypes_structs_match_the_c_header(tmp_path):
    """_lib.py's ctypes mirrors of the ABI structs have the size and field offsets the C compiler gives them."""
    import subprocess
    from twixt_for_open_spiel_b200 import _lib
    src = tmp_path / "layout.c"
    fields = {"twixt_game_info": [f for f, _ in _lib.GameInfo._fields_],
              "twixt_stats": [f for f, _ in _lib.Stats._fields_],
              "twixt_step_result": [f for f, _ in _lib.StepResult._fields_]}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "twixt_b200.h"', 'int main(void) {']
    for name, fs in fields.items():
        lines.append('  printf("%s %%zu", sizeof(%s));' % (name, name))
        for f in fs:
            lines.append('  printf(" %%zu", offsetof(%s, %s));' % (name, f))
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout.strip().splitlines()
    mirrors = {"twixt_game_info": _lib.GameInfo, "twixt_stats": _lib.Stats, "twixt_step_result": _lib.StepResult}
    assert len(out) == 3
    for line in out:
        toks = line.split()
        cls = mirrors[toks[0]]
        assert C.sizeof(cls) == int(toks[1]), toks[0]
        assert [getattr(cls, f).offset for f, _ in cls._fields_] == [int(t) for t in toks[2:]], toks[0]


def test_shard_helpers_need_no_gpu():
    """twixt_shard_range / twixt_stats_accumulate (what sharding.py itself calls): shares are contiguous,
    cover the range and differ by at most one env."""
    from twixt_for_open_spiel_b200 import _lib
    from twixt_for_open_spiel_b200.sharding import shard_range
    for total, world in ((1 << 20, 8), (10, 4), (7, 7), (3, 5), (0, 2)):
        nxt, sizes = 0, []
        for r in range(world):
            first, count = shard_range(total, world, r)
            assert first == nxt
            nxt += count
            sizes.append(count)
        assert nxt == total and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 4, 4)
    lib = _lib.load()
    acc, part = _lib.Stats(), _lib.Stats()
    part.plies, part.games, part.max_length, part.kernel_launches = 11, 2, 9, 1
    for _ in range(3):
        assert lib.twixt_stats_accumulate(C.byref(acc), C.byref(part)) == 0
    assert (acc.plies, acc.games, acc.max_length, acc.kernel_launches) == (33, 6, 9, 3)
