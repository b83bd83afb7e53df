"""CPU: the in-register DFT building blocks of the 320-point transform (csrc/fft320.cuh), compiled with g++: 16- and
20-point DFTs against naive fp64 DFTs for both signs, and the composed centred 320-point transform -- 16-point DFTs,
w320 twiddles, 20-point prime-factor DFTs, ifftshift / fftshift folded into the index maps -- against
fftshift(dft(ifftshift(x)))."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft320_building_blocks(tmp_path):
    exe = str(tmp_path / "fft320_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "fft320_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
