"""The C++ drop-in shells (csrc/host: ORB_SLAM2::ORBextractor and the four GPU-backed ORBmatcher members) compiled against
the cv:: shim and driven by a Frame::ExtractORB look-alike and stub Frame/KeyFrame/MapPoint objects (tests/cpp/shell_test.cc).
CPU: everything compiles, links and fails loudly without a device.  GPU: every result equals the CPU oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "orb_slam2_with_comment_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "shell_test")


def build_exe():
    from orb_slam2_with_comment_b200 import build
    import oracle_lib
    build.build()
    oracle_lib.load_port()
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "csrc", "host")])
    src = os.path.join(ROOT, "tests", "cpp", "shell_test.cc")
    deps = [src, os.path.join(PKG, "liborbslam2_shell.so"), os.path.join(ROOT, "oracle", "_build", "liborboracle.so")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) > os.path.getmtime(d) for d in deps):
        return EXE
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wno-unused-function", f"-I{ROOT}/include", f"-I{ROOT}/shim",
                           f"-I{ROOT}/shim/orbslam2", f"-I{PKG}/csrc/host", "-o", EXE, src, f"-L{PKG}", "-lorbslam2_shell", "-lorbgpu",
                           f"-L{ROOT}/oracle/_build", "-lorboracle", "-pthread", f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{ROOT}/oracle/_build"])
    return EXE


def test_shells_compile_link_and_fail_loudly_without_a_device():
    import torch
    exe = build_exe()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3, r.stdout + r.stderr
    assert r.stdout.count("no CUDA device") == 2


@pytest.mark.gpu
def test_shells_equal_oracle_on_gpu():
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
