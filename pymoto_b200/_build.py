"""Build libpmb.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpmb.so")
SOURCES = ["pmb_api.cu", "pmb_assembly.cu", "pmb_spmv.cu", "pmb_symstore.cu", "pmb_multigrid.cu", "pmb_vector.cu", "pmb_filter.cu", "pmb_elem.cu", "pmb_optim.cu", "pmb_solver.cu", "pmb_probe.cu", "pmb_comm.cu", "pmb_peer.cu"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pmb.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # IEEE arithmetic everywhere (no --use_fast_math); FMA contraction is on except where kernels ask for separate
    # multiply/add with __dmul_rn/__dadd_rn to reproduce the reference's rounding bit for bit
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"]
    cmd += ["-Xcompiler", "-fPIC", "-shared", "--fmad=true", "-cudart", "shared"]
    cmd += ["--threads", "0"]  # one compilation job per source file and core: 100 s -> 37 s on 8 cores, the same SASS
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libpmb.so")
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
