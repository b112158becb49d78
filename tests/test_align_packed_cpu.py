"""The packed (int16x2, DPX) Smith-Waterman forward pass of kaamer_b200/csrc/align_packed.cuh, run on the CPU:
tests/csrc/align_packed_host.cu compiles the kernel's own per-lane step as host code, emulates the 32 lanes of a
warp in lock-step and compares every traceback byte and the end cell of both pairs of a job with a scalar DP
(jobs of unequal pairs, all four column widths, multi-block subjects, a pair with illegal letters, other
gap-open scores).  No GPU needed: the DPX intrinsics have host definitions in the CUDA headers."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_forward_pass_against_a_scalar_dp(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    exe = str(tmp_path / "align_packed_host")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "csrc", "align_packed_host.cu")])
    out = subprocess.run([exe, "64"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.strip().endswith("0 mismatches"), out.stdout[-500:]


def test_alignment_schedule_invariants(tmp_path):
    """build_align_plan (kaamer_b200/csrc/align_plan.hpp): every pair scheduled exactly once as a long pair, a single
    pair or half of a packed job; jobs only from pairs the int16 lanes can hold; traceback regions of a chunk
    disjoint, aligned and inside the budget; chunking under a small budget; an oversized pair is reported"""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    exe = str(tmp_path / "align_plan_host")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "csrc", "align_plan_host.cu")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert out.stdout.strip().endswith("0 failures"), out.stdout[-500:]
