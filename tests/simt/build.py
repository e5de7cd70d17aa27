"""TEST INFRASTRUCTURE ONLY — builds tests/simt/build/libmccfr_simt.so: the MCCFR kernel SOURCE of robopoker_b200/csrc/mccfr.cu compiled as host
code under tests/simt/simt.hpp.  The kernels are taken verbatim from the .cu file (the text between the "device-side views" marker and the end of
`namespace rbp`); the only edits are the two spellings of shared memory, which have no host meaning:
    extern __shared__ ... smem_raw[];  ->  unsigned char* smem_raw = ::smem_raw;
    __shared__ T name[...];            ->  static T name[...];
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "robopoker_b200", "csrc")
OUT = os.path.join(HERE, "build")
LIB = os.path.join(OUT, "libmccfr_simt.so")


def extract():
    src = open(os.path.join(CSRC, "mccfr.cu")).read()
    a = src.index("// ───────────────────────────── device-side views")
    b = src.index("}  // namespace rbp", a) + len("}  // namespace rbp")
    body = src[a:b]
    body = re.sub(r"extern __shared__ __align__\(16\) unsigned char smem_raw\[\];", "unsigned char* smem_raw = ::smem_raw;", body)
    body = body.replace("__shared__ ", "static ")
    assert "__shared__" not in body
    return body


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    inc = os.path.join(OUT, "mccfr_kernels.inc")
    body = extract()
    srcs = [os.path.join(HERE, "mccfr_simt.cpp"), os.path.join(CSRC, "flat_game.cpp")]
    deps = srcs + [os.path.join(HERE, "simt.hpp"), os.path.join(CSRC, "mccfr.cu"), os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "flat_game.hpp")]
    stale = force or not os.path.exists(LIB) or not os.path.exists(inc) or open(inc).read() != body or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)
    if not stale:
        return LIB
    open(inc, "w").write(body)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", OUT, "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"),
           "-I", "/usr/local/cuda/include", "-o", LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("simt build failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
