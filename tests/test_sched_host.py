"""CPU: the static tile schedule of the synthesis kernel (csrc/siren_sched.h), compiled with g++ and enumerated on the
host for both tile-per-iteration counts (4: two tile slots, fp16/bf16; 2: one slot, fp16x3): every (patch, coordinate
block) of every cluster exactly once, for ragged patch counts, remainder blocks and more clusters than work."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("sched") / "sched_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "sched_check.cpp")], check=True)
    return exe


@pytest.mark.parametrize("tpi", [4, 2])
@pytest.mark.parametrize("n_act,C,clusters", [
    (1, 576, 1), (2, 576, 1), (3, 576, 74), (5, 576, 2), (223, 576, 55), (400, 576, 74), (94000, 576, 74),
    (20011, 576, 74), (129, 576, 1), (257, 576, 2), (37, 144, 9), (9, 256, 3), (23, 400, 6), (131, 576, 33),
    (6, 1024, 2), (1000, 128, 74), (77, 200, 5)])
def test_schedule_covers_every_tile_exactly_once(checker, tpi, n_act, C, clusters):
    r = subprocess.run([checker, str(n_act), str(C), str(clusters), str(tpi)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
