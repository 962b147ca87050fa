"""Builds libtem_b200.so (sm_100a) in-tree with nvcc.  Usage: python -m transfer_em_b200.build [--force]"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtem_b200.so")
SOURCES = ["tem_runtime.cu", "conv_direct.cu", "elementwise.cu", "wgrad_mma.cu", "wgrad_c1.cu", "conv_mma.cu", "conv_c1.cu", "conv_tc3.cu", "conv_tc_s2.cu", "wgrad_tc.cu", "conv_tcw.cu", "wgrad_tc_s2.cu", "conv_small.cu", "wgrad_tcw.cu", "disc_tail.cu", "conv_c1tc.cu", "wgrad_c1tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _digest(srcs):
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/transfer_em_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and not f.endswith(".o"):     # sources and headers only: objects change with every build
            h.update(f.encode()); h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, ablation=False):
    """ablation=True builds libtem_b200_abl.so with -DTEM_ABLATION (stage-ablation knobs live; results wrong by design);
    it is only ever loaded when TEM_ABLATION_LIB=1 is set (profiling tools), never by the product or the tests."""
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    lib = LIB.replace(".so", "_abl.so") if ablation else LIB
    flags = NVCC_FLAGS + (["-DTEM_ABLATION"] if ablation else [])
    stamp = lib + ".stamp"
    dig = _digest(srcs) + ("-abl" if ablation else "")
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(CSRC, ("abl_" if ablation else "") + s.replace(".cu", ".o"))
        cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    open(stamp, "w").write(dig)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ablation="--ablation" in sys.argv))
