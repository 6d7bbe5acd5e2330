"""Builds shannon_b200/libshannon_b200.so in-tree with nvcc for sm_100a (no JIT cache: the built
.so travels to the GPU box with the repo snapshot).  `python -m shannon_b200.build [--force]`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libshannon_b200.so")
SOURCES = ["api.cu", "table.cu", "l3.cu", "l4.cu", "selfjoin.cu", "synth.cu", "reads.cu", "route.cu",
           "shard.cu", "reps.cu", "condense.cu", "hostio.cpp"]
# translation units that depend on the K1-mer key width: built a second time with -DSHN_WIDE
# (128-bit keys, K1 = 33) into namespace `wide`; api.cu dispatches on k1
WIDE_SOURCES = ["table.cu", "l3.cu", "l4.cu", "synth.cu", "shard.cu"]
HEADERS = ["common.cuh", "table_dev.cuh", "selfjoin.cuh", "impls.h", "reads.cuh", "uf_dev.cuh",
           os.path.join("..", "..", "include", "shannon_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-O3,-pthread", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


# e.g. SHN_NVCC_DEFINES="-DSHN_SPEC_DEBUG" python -m shannon_b200.build  (debug builds only)
EXTRA_DEFINES = os.environ.get("SHN_NVCC_DEFINES", "").split()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    procs = []
    for src, extra, tag in [(x, [], "") for x in SOURCES] + [(x, ["-DSHN_WIDE"], ".wide") for x in WIDE_SOURCES]:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src + tag + ".o")
        objs.append(op)
        if force or _stale(op, [sp] + hdrs):
            cmd = [NVCC] + FLAGS + EXTRA_DEFINES + extra + ["-x", "cu", "-c", sp, "-o", op]
            procs.append((src + tag, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        else:
            with open(os.path.join(OBJ, src + ".ptxas.log"), "w") as f:
                f.write(out)
            if verbose:
                sys.stderr.write(out)
    if failed:
        raise RuntimeError("shannon_b200: CUDA build failed")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-pthread", "-lpthread"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
