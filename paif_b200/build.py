"""In-tree build of the C-ABI shared library ``paif_b200/libpaif_b200.so`` with nvcc for
sm_100a (the only target).  ``python -m paif_b200.build`` or ``__graft_entry__.build()``."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpaif_b200.so")
SOURCES = ["abi.cu", "stem.cu", "gf.cu", "gf_mix.cu", "conv_direct.cu", "conv_tcgen05.cu", "pointwise.cu", "glue.cu", "fusion_net.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC"]


def _nvcc():
    cand = os.environ.get("NVCC")
    if cand:
        return cand
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        return "/usr/local/cuda/bin/nvcc"
    return "nvcc"


def _deps():
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    deps.append(os.path.join(HERE, "..", "include", "paif_b200.h"))
    return deps


def _src_hash():
    """sha256 over the sources, the header and the compiler flags: what the library was built from."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    """The library is rebuilt when its sources changed.  Decided by content (a hash written next to the library), not by
    file times: a snapshot of the tree on another machine (gpurun) does not keep them, and the library that travelled
    with it would be rebuilt for nothing in every process."""
    if not os.path.exists(LIB):
        return True
    stamp = LIB + ".srchash"
    if os.path.exists(stamp):
        with open(stamp) as f:
            return f.read().strip() != _src_hash()
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False, profile=False, sanitize=False):
    """profile: -DPAIF_TC_PROFILE role-timeline counters (libpaif_b200_prof.so); sanitize: a build for compute-sanitizer
    runs (libpaif_b200_san.so: mbarrier waits poll without the suspend-time hint; the conv engine's block barrier is a
    non-inlined function so that synccheck sees one barrier instruction)."""
    if sanitize:
        return _build_variant("_san", ["-DPAIF_NO_SUSPEND_HINT", "-DPAIF_SANITIZER_BUILD"])
    lib = LIB.replace(".so", "_prof.so") if profile else LIB
    if profile:
        if os.path.exists(lib) and os.path.getmtime(lib) >= max(
                os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)):
            return lib
    elif not force and not needs_build():
        return LIB
    objs, procs = [], []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", "_prof.o" if profile else ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DPAIF_TC_PROFILE"] if profile else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("paif_b200: nvcc build failed")
    subprocess.check_call([_nvcc(), "-shared", "-o", lib] + objs + ["-lcudart"])
    if lib == LIB:
        with open(LIB + ".srchash", "w") as f:
            f.write(_src_hash() + "\n")
    return lib


def _build_variant(suffix, defines):
    lib = LIB.replace(".so", suffix + ".so")
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", suffix + ".o"))
        procs.append(subprocess.Popen([_nvcc()] + NVCC_FLAGS + defines + ["-c", os.path.join(CSRC, src), "-o", obj],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        objs.append(obj)
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("paif_b200: nvcc build failed:\n" + out)
    subprocess.check_call([_nvcc(), "-shared", "-o", lib] + objs + ["-lcudart"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, profile="--profile" in sys.argv,
                sanitize="--sanitize" in sys.argv))
