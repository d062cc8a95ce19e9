"""TEST INFRASTRUCTURE ONLY -- checks the CUDA emulator itself (see selftest_kernels.cu)."""
import ctypes as C
import os
import subprocess

from . import build_emu

HERE = os.path.dirname(os.path.abspath(__file__))


def run_selftest():
    os.makedirs(build_emu.GEN, exist_ok=True)
    src = os.path.join(HERE, "selftest_kernels.cu")
    gen = os.path.join(build_emu.BUILD, "selftest_kernels.cpp")
    so = os.path.join(build_emu.BUILD, "libemu_selftest.so")
    open(gen, "w").write(build_emu.transform(open(src).read(), "selftest_kernels.cu"))
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I",
                        os.path.join(HERE, "include"), "-o", so, gen, os.path.join(HERE, "engine.cpp"), "-lpthread"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(so)
    lib.selftest_error.restype = C.c_char_p
    out = (C.c_int * 192)()
    assert lib.selftest_run(0, out) == 0, lib.selftest_error()
    # v = 496 everywhere; seg = 8*(lane//8)+3 of the thread 32 further on; 11 lanes with lane % 3 == 0; all_sync true
    for b in range(3):
        for t in range(64):
            other = (t + 32) % 64
            assert out[b * 64 + t] == 496 + (8 * ((other % 32) // 8) + 3) + 11 + 1000, (b, t, out[b * 64 + t])
    for which, needle in ((1, b"different *_sync collectives"), (2, b"exited"), (3, b"deadlock")):
        assert lib.selftest_run(which, out) != 0, f"selftest kernel {which} must fail"
        msg = lib.selftest_error()
        assert needle in msg, msg
    assert lib.selftest_run(0, out) == 0     # the engine recovers after a reported error
