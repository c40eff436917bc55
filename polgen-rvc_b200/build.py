"""In-tree nvcc build of libpolgen_rvc.so (sm_100a only).

    python polgen-rvc_b200/build.py [--force]

The shared library is the C-ABI product boundary (include/polgen_rvc.h).  It
links cudart statically and resolves the one driver entry point it needs
(cuTensorMapEncodeTiled) at run time, so it loads on a GPU-less host too.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpolgen_rvc.so")
SOURCES = ["pg_api.cu", "pg_simt.cu", "pg_source.cu", "pg_conv_umma.cu", "pg_conv_planes.cu", "pg_planes.cu", "pg_attention.cu", "pg_pair_planes.cu", "pg_pipeline.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "polgen_rvc.h"))
    return hs


def _obj_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src, *_headers()])


def build(force: bool = False, verbose: bool = False) -> str:
    """compile the stale objects (in parallel: the big kernels take a minute each) and link"""
    from concurrent.futures import ThreadPoolExecutor
    jobs, objs = [], []
    for src in SOURCES:
        path, obj = os.path.join(CSRC, src), os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _obj_stale(path, obj):
            jobs.append((src, [NVCC, *FLAGS, "-c", path, "-o", obj]))
    if not jobs and os.path.exists(OUT) and all(os.path.getmtime(o) <= os.path.getmtime(OUT) for o in objs):
        return OUT

    def run(job):
        return job[0], subprocess.run(job[1], capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as ex:
        results = list(ex.map(run, jobs))
    for src, r in results:
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
