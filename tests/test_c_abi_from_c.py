"""include/aps_b200.h is a C header and libaps_b200.so a C library: compile a plain C99 caller with
gcc, link it against the in-tree library, run it (no GPU needed: argument errors only)."""
import os
import subprocess

from advancedps_b200 import _abi, _lib
import ctypes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_caller_links_and_runs(tmp_path):
    so = _lib.build()
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(so)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror=implicit-function-declaration",
                         "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                         "-o", exe, "-L" + libdir, "-laps_b200", "-Wl,-rpath," + libdir],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert cc.returncode == 0, cc.stdout
    run = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert run.returncode == 0, (run.returncode, run.stdout)
    # the C compiler's struct sizes are the ones the ctypes / Julia mirrors assume
    assert f"sizeof(aps_config)={ctypes.sizeof(_abi.ApsConfig)}" in run.stdout
    assert f"sizeof(aps_model)={ctypes.sizeof(_abi.ApsModel)}" in run.stdout


def _fnv(a):
    h = 1469598103934665603
    for b in a.tobytes():
        h = ((h ^ b) * 1099511628211) & (2**64 - 1)
    return h


import pytest  # noqa: E402


@pytest.mark.gpu
def test_plain_c_caller_runs_a_real_sweep(tmp_path):
    """tests/c/abi_sweep.c: create -> set_observations -> sweep -> accessors -> pick -> conditional
    sweep, from a gcc-compiled C99 program; every printed number equals the oracle's."""
    import numpy as np

    import oracle as O
    from advancedps_b200 import models

    so = _lib.build()
    exe = str(tmp_path / "abi_sweep")
    libdir = os.path.dirname(so)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror=implicit-function-declaration",
                         "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_sweep.c"),
                         "-o", exe, "-L" + libdir, "-laps_b200", "-Wl,-rpath," + libdir],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert cc.returncode == 0, cc.stdout
    run = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert run.returncode == 0, (run.returncode, run.stdout)
    out = dict(line.split(" ", 1) for line in run.stdout.strip().splitlines())
    N, T = 4096, 8
    Y = (0.35 + 0.01 * np.arange(T)).reshape(T, 1)
    cfg = _abi.make_config(models.linear_gaussian(), N, T, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
    ro = O.sweep(cfg, Y, 1234, mode=O.CANON)
    assert float(out["logevidence"]) == ro.logevidence
    assert int(out["anc_hash"]) == _fnv(ro.anc_hist[T])
    assert int(out["x_hash"]) == _fnv(np.ascontiguousarray(ro.x_hist[T - 1][:, 0]))
    assert int(out["w_hash"]) == _fnv(ro.final_w)
    slot, traj = O.pick_trajectory(cfg, 1234, ro, mode=O.CANON)
    assert int(out["slot"]) == slot
    assert np.array_equal(np.array([float(v) for v in out["traj"].split()]), traj[:, 0])
    ro2 = O.sweep(cfg, Y, 1235, ref_traj=traj, mode=O.CANON)
    assert float(out["logevidence2"]) == ro2.logevidence
