"""In-tree build of librbp_b200.so (sm_100a only) and of the oracle's C++ restatement.

nvcc cross-compiles without a GPU; the .so files travel to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "robopoker_b200", "csrc")
LIB = os.path.join(ROOT, "robopoker_b200", "librbp_b200.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "build", "librbp_oracle.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]

# per-source flags: kernels whose results must be bit-identical to the oracle's strict f32 arithmetic are
# compiled without FMA contraction and with IEEE division/sqrt (nvcc defaults: -prec-div=true -prec-sqrt=true -ftz=false)
SOURCES = {
    "flat_game.cpp": ["-Xcompiler", "-ffp-contract=off"],
    "mccfr.cu": ["-fmad=false", "-Xcompiler", "-ffp-contract=off"],
    "nlhe.cu": ["-fmad=false", "-Xcompiler", "-ffp-contract=off"],
    "deuce.cu": ["-fmad=false"],
    "iso.cu": ["-fmad=false"],
    "lloyd_w1.cu": ["-fmad=false"],
    "lloyd_sk.cu": ["-fmad=false"],
    "kmeans_api.cu": ["-fmad=false"],
    "comm.cu": [],
}


def _nccl_include():
    """nccl.h for the types only: the library dlopens libnccl at run time (csrc/comm.cu), it does not link it."""
    try:
        import nvidia.nccl as n
        return ["-I" + os.path.join(list(n.__path__)[0], "include")]
    except Exception:
        return []


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r.stdout


def build_lib(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "rbp.h"))
    objs = []
    for src, flags in SOURCES.items():
        path = os.path.join(CSRC, src)
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        if force or not _newer(obj, [path] + headers):
            out = _run([NVCC] + ARCH + COMMON + _nccl_include() + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj])
            if verbose:
                print(out)
        objs.append(obj)
    if force or not _newer(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"])
    return LIB


def _host_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            return any(line.startswith("flags") and " fma " in line + " " for line in f)
    except OSError:
        return False


def _cpu_signature():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha256(line.encode()).hexdigest()[:16]
    except OSError:
        pass
    return "unknown"


def build_oracle(force=False, native=False):
    """The checker.  `native=False`: the portable build every test uses (it travels to the GPU box prebuilt, so it must not
    assume this container's CPU).  `native=True`: the CPU-baseline build BASELINE.md promises — `-O3 -march=native`, compiled
    on the host that is about to be timed (bench.py's cpu_baseline / --impl reference legs), keyed by that host's CPU flags.
    Both are `-ffp-contract=off -fno-fast-math`: same results bit for bit (tests/test_oracle_mccfr.py checks the native one)."""
    srcs = [os.path.join(ORACLE_DIR, f) for f in sorted(os.listdir(ORACLE_DIR)) if f.endswith(".cpp")]
    deps = srcs + [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith(".hpp")]
    os.makedirs(os.path.dirname(ORACLE_LIB), exist_ok=True)
    common = ["-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread"]
    if native:
        lib = os.path.join(os.path.dirname(ORACLE_LIB), "librbp_oracle_native.so")
        stamp = lib + ".cpu"
        sig = _cpu_signature()
        fresh = _newer(lib, deps) and os.path.exists(stamp) and open(stamp).read().strip() == sig
        if force or not fresh:
            _run(["g++", "-std=c++17", "-O3", "-march=native"] + common + ["-o", lib] + srcs)
            with open(stamp, "w") as f:
                f.write(sig)
        return lib
    if force or not _newer(ORACLE_LIB, deps):
        # explicit fmaf() of the exp/ln contract: inline vfmadd when this host has it (nothing else contracts:
        # -ffp-contract=off), else glibc's exact fmaf
        fma = ["-mfma"] if _host_has_fma() else []
        _run(["g++", "-std=c++17", "-O2", "-march=x86-64-v2"] + fma + common + ["-o", ORACLE_LIB] + srcs)
    return ORACLE_LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_oracle(force="--force" in sys.argv))
