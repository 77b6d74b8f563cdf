"""Build the CUDA library in-tree with nvcc for sm_100a (no torch extension machinery: the
boundary is a plain C ABI, include/alchemy_b200.h)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libalchemy_b200.so")
SOURCES = ["ab_plan.cpp", "ab_mixed_plan.cpp", "ab_abi.cpp", "ab_she.cpp", "ab_elementwise.cu", "ab_crt_generic.cu", "ab_crt_mixed.cu", "ab_crt_mixed_inv.cu", "ab_crt_mixed_mac.cu", "ab_crt_mixed_mac_smem.cu",
           "ab_ntt_pow2.cu", "ab_symm.cu", "ab_rng.cu", "ab_mulks_pow2.cu", "ab_mulks_n256.cu", "ab_wire.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-x", "cu"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant: build a copy with extra -D flags into build_variants/<variant>/ (kernel A/B experiments; load it with
    ALCHEMY_B200_LIB=<path>); the in-tree library is built without arguments"""
    global LIBDIR, LIB
    if variant:
        LIBDIR = os.path.join(os.path.dirname(HERE), "build_variants", variant)
        LIB = os.path.join(LIBDIR, "libalchemy_b200.so")
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "alchemy_b200.h"))
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out.decode()}")
    if force or procs or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-cudart", "static"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    print(build(force="--force" in sys.argv, verbose=True, variant=var[0] if var else None, defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
