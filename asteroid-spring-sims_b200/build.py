"""In-tree build of libass_b200.so (sm_100a only) with nvcc.  No JIT cache, no torch extension machinery: the .so
lands next to the sources so it travels to the GPU box with the snapshot."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libass_b200.so")
# tuning only: ASS_LIB loads / builds a differently-named library, ASS_NVCC_EXTRA appends flags (e.g. -DASS_SPRING_LANES=4)
if os.environ.get("ASS_LIB"):
    LIB = os.path.abspath(os.environ["ASS_LIB"])

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=default",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "ass_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link libass_b200.so. Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libass_b200.so (there is no CPU fallback)")
    objdir = os.path.join(PKG, "build") if not os.environ.get("ASS_LIB") else LIB + ".obj"
    extra = os.environ.get("ASS_NVCC_EXTRA", "").split()
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if src.endswith(".cpp"):
            cmd = [nvcc] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-x", "c++", "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (os.path.basename(src), out))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
