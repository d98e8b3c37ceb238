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
