"""Build libvplines_b200.so in-tree with nvcc for sm_100a.

-fmad=false is REQUIRED: the float32/float64 sequences of LSD and LBD mirror the
CPU code operation by operation and must not be contracted into FMAs
(vpl_create refuses to run a library built without it).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvplines_b200.so")
SOURCES = ["prims.cu", "preproc.cu", "lsd_front.cu", "lsd_engine.cu", "lsd_engine_spec.cu", "lsd_nfa.cu", "lbd_match.cu", "edlines.cu", "linematch.cu", "vp.cu", "capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--threads", "4",
] + os.environ.get("VPL_EXTRA_NVCC", "").split()


def nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.exists(p) or p == "nvcc"):
            return p
    return "nvcc"


FLAGS_STAMP = os.path.join(HERE, "build", "nvcc_flags.txt")


def needs_build():
    if not os.path.exists(OUT):
        return True
    try:  # built with other flags (VPL_EXTRA_NVCC)?
        if open(FLAGS_STAMP).read() != " ".join(NVCC_FLAGS):
            return True
    except OSError:
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "vpl_capi.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc(), "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.check_call(cmd)
    with open(FLAGS_STAMP, "w") as f:
        f.write(" ".join(NVCC_FLAGS))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
