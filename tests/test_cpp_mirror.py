"""include/rbp.hpp — the C++ host-side mirror of the reference's trait surface over the C ABI (the reference is compiled
code; there is no Rust toolchain here) — compiles against the header and the library, refuses to compute without a
device, and drives a short run with one."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    if not shutil.which("g++"):
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "host_mirror")
    lib_dir = os.path.join(ROOT, "robopoker_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"), "-o", exe, "-L", lib_dir, "-l:librbp_b200.so", f"-Wl,-rpath,{lib_dir}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def _run(exe):
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    return r.stdout


def test_cpp_mirror_refuses_without_a_device(rbp, tmp_path):
    if rbp.load_library().rbp_device_count() > 0:
        pytest.skip("GPU present")
    assert "no device: 0 failures" in _run(_build(tmp_path))


@pytest.mark.gpu
def test_cpp_mirror_drives_the_library(rbp, tmp_path):
    assert "device: 0 failures" in _run(_build(tmp_path))


def test_header_is_a_plain_c_abi(rbp, tmp_path):
    if not shutil.which("gcc"):
        pytest.skip("no C compiler")
    exe = str(tmp_path / "abi_plain_c")
    lib_dir = os.path.join(ROOT, "robopoker_b200")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_plain_c.c"), "-o", exe, "-L", lib_dir, "-l:librbp_b200.so", f"-Wl,-rpath,{lib_dir}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    _run(exe)
