"""Build libucoslam_b200.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

    python ucoslam-cv3_b200/build.py [--force]

Every .cu under csrc/ is compiled with -gencode arch=compute_100a,code=sm_100a -lineinfo.  Floating-point contraction
is disabled for all device code (-fmad=false) and host code (-ffp-contract=off): the extractor restates OpenCV float
arithmetic that the reference compiles without FMA (cmake/compiler.cmake:11-22), and the BA path wants run-to-run
reproducible sums.
"""
import os, subprocess, sys, glob, hashlib

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
LIB = os.path.join(OUT, "libucoslam_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-O3", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _stamp(paths):
    h = hashlib.sha1(" ".join(FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every csrc/*.cu (in parallel; an object is reused when neither its source, nor any header, nor the flags changed)
    and link libucoslam_b200.so."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    stamp = _stamp(srcs + hdrs)
    stamp_file = os.path.join(OUT, "build.stamp")
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB

    def compile_one(s):
        o = os.path.join(OUT, os.path.basename(s)[:-3] + ".o")
        ostamp, olog = o + ".stamp", o + ".log"
        st = _stamp([s] + hdrs)
        if not force and os.path.exists(o) and os.path.exists(ostamp) and os.path.exists(olog) and open(ostamp).read() == st:
            return o, open(olog).read(), 0
        cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        text = "$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr
        if r.returncode == 0:
            open(ostamp, "w").write(st)
            open(olog, "w").write(text)
        return o, text, r.returncode

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, srcs))
    log, objs = [], []
    for (o, text, rc), s in zip(results, srcs):
        log.append(text)
        if rc != 0:
            sys.stderr.write(text)
            raise RuntimeError("nvcc failed for " + s)
        objs.append(o)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]   # -ldl: NVTX v3 loads its injection library lazily
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    with open(os.path.join(OUT, "build.log"), "w") as f:
        f.write("\n".join(log))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
