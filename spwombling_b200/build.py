"""Build libspwombling.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

``python -m spwombling_b200.build`` or ``build()``; objects are rebuilt only when a source or header
is newer.  The shared library lands next to the sources (``spwombling_b200/csrc/libspwombling.so``) so
it travels with the repository snapshot to the GPU box.
"""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libspwombling.so")
SOURCES = ["api.cu", "cov.cu", "linalg.cu", "potrf_dag.cu", "potrf_la.cu", "wombling.cu", "gradient_tma.cu", "tma.cu", "summary.cu", "hlm.cu"]
# cov.cu is compiled without FMA contraction: its elementwise formulas then round exactly like the reference's
# R arithmetic (and K[i][j] == K[j][i] bit for bit, whichever thread computes it)
# wombling.cu likewise: its double quadrature relies on t2 - t1 being exactly zero on the diagonal node pairs (the
# reference switches to a zero-lag formula there); an FMA-contracted a2 * tval - a1 * tval would leave a rounding residue
EXTRA_FLAGS = {"cov.cu": ["-fmad=false"], "wombling.cu": ["-fmad=false"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr"]


def _newest_header():
    inc = os.path.join(os.path.dirname(os.path.dirname(CSRC)), "include", "spwombling.h")
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [inc]
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    hdr_time = _newest_header()
    objs, rebuilt = [], False
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = s[:-3] + ".o"
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_time):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + EXTRA_FLAGS.get(src, []) + ["-c", s, "-o", o]
            procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            rebuilt = True
    for src, cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s: %s" % (src, " ".join(cmd)))
    if rebuilt or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
