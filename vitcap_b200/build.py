"""Builds libvitcap_b200.so (sm_100a) in-tree with nvcc. No torch involvement: the library is a plain C-ABI
shared object (include/vitcap_b200.h) that links only the static CUDA runtime."""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libvitcap_b200.so")
# the same sources with IEEE-half storage of every 16-bit operand (csrc/common.cuh, VC_STORE_F16; selected by VITCAP_STORE=fp16)
LIB_F16 = os.path.join(LIBDIR, "libvitcap_b200_f16.so")
OBJDIR_F16 = os.path.join(HERE, "build", "f16")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    for extra in ("common.cuh", "pair.cuh"):
        with open(os.path.join(CSRC, extra), "rb") as f:
            h.update(f.read())
    with open(os.path.join(os.path.dirname(HERE), "include", "vitcap_b200.h"), "rb") as f:
        h.update(f.read())
    with open(path, "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose, objdir=OBJDIR, defines=()):
    path = os.path.join(CSRC, src)
    obj = os.path.join(objdir, src[:-3] + ".o")
    stamp = obj + ".sha1"
    dg = _digest(path) + "".join(defines)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dg:
        return obj, False, ""
    cmd = [_nvcc()] + NVCC_FLAGS + list(defines) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dg)
    return obj, True, r.stderr


def build(verbose=False, force=False, half_store=False):
    """half_store: the VC_STORE_F16 build (libvitcap_b200_f16.so) instead of the default library."""
    objdir, lib, defines = (OBJDIR_F16, LIB_F16, ("-DVC_STORE_F16",)) if half_store else (OBJDIR, LIB, ())
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    if force:
        for f in os.listdir(objdir):
            if os.path.isfile(os.path.join(objdir, f)):
                os.remove(os.path.join(objdir, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose, objdir, defines), srcs))
    objs = [o for o, _, _ in results]
    changed = any(c for _, c, _ in results)
    log = "\n".join("== %s ==\n%s" % (s, l) for s, (_, c, l) in zip(srcs, results) if c and l)
    if log:
        with open(os.path.join(objdir, "ptxas.log"), "w") as f:
            f.write(log)
        if verbose:
            print(log)
    if changed or not os.path.exists(lib):
        cmd = [_nvcc(), "-shared", "-o", lib] + objs + ["-cudart", "static", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, half_store=True))
