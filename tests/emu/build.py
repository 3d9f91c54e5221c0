"""TEST INFRASTRUCTURE ONLY -- builds tests/emu/libfcp_emu.so: the product's csrc/*.cu re-compiled by g++ against the CPU
emulation of the CUDA programming model in this directory (cuda_runtime.h, emu.cpp).

The only source transformation is the launch syntax: ``kernel<<<cfg>>>(args)`` becomes
``emu::launch(emu::Cfg(cfg), kernel, args)``.  Everything else is the product source, compiled with
``-ffp-contract=off`` (the counterpart of nvcc's ``-fmad=false``), so the kernels' arithmetic, loop order, reduction trees,
tickets, barriers and the peer-memory protocol run exactly as written -- on fibers instead of CUDA threads.

Nothing in the product loads this library; tests opt in with FCP_TEST_EMU=1 (tests/conftest.py) so that the `-m gpu`
parity suite can be exercised in a container without a GPU.  It says nothing about performance.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "freecappuccino-dev_b200", "csrc")
OUT = os.path.join(HERE, "_build")
SO = os.path.join(HERE, "libfcp_emu.so")
CXXFLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-DFCP_EMU=1", "-I", HERE, "-I", CSRC,
            "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-sign-compare", "-Wno-maybe-uninitialized", "-Wno-parentheses"]


def _match(src: str, i: int, open_c: str, close_c: str) -> int:
    depth = 0
    while True:
        c = src[i]
        if c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1


def rewrite_launches(src: str) -> str:
    out, i = [], 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            out.append(src[i:])
            return "".join(out)
        ls = src.rfind("\n", 0, j) + 1
        if "//" in src[ls:j]:                       # inside a line comment
            out.append(src[i:j + 3])
            i = j + 3
            continue
        k = j
        while src[k - 1].isspace():
            k -= 1
        if src[k - 1] == ">":                       # template arguments of the kernel
            depth = 0
            while True:
                k -= 1
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
            while src[k - 1].isspace():
                k -= 1
        while src[k - 1].isalnum() or src[k - 1] in "_:":
            k -= 1
        kern = src[k:j].strip()
        m = src.find(">>>", j)
        cfg = src[j + 3:m]
        p = m + 3
        while src[p].isspace():
            p += 1
        assert src[p] == "(", f"launch of {kern}: expected '(' after >>>"
        q = _match(src, p, "(", ")")
        args = src[p + 1:q].strip()
        out.append(src[i:k])
        out.append(f"emu::launch(emu::Cfg({cfg}), {kern}" + (f", {args})" if args else ")"))
        i = q + 1


def build(force: bool = False, verbose: bool = False) -> str:
    """Rebuild when a source is newer than the library; serialised with a file lock (the ranks of a multi-process test all call this)."""
    import fcntl
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build(force: bool, verbose: bool) -> str:
    cus = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
           [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cpp", ".py"))] + [os.path.join(ROOT, "include", "fcp.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    jobs = []
    for f in cus:
        with open(os.path.join(CSRC, f)) as fh:
            text = fh.read()
        cpp = os.path.join(OUT, f[:-3] + ".cpp")
        with open(cpp, "w") as fh:
            fh.write(f'#line 1 "{os.path.join(CSRC, f)}"\n' + rewrite_launches(text))
        jobs.append((cpp, cpp[:-4] + ".o"))
    jobs.append((os.path.join(HERE, "emu.cpp"), os.path.join(OUT, "emu.o")))

    def cc(job):
        src, obj = job
        r = subprocess.run(["g++"] + CXXFLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0 or (verbose and r.stderr):
            sys.stderr.write(r.stderr[-20000:])
        if r.returncode != 0:
            raise RuntimeError(f"emulation build failed: {src}")
    with ThreadPoolExecutor(8) as ex:
        list(ex.map(cc, jobs))
    subprocess.check_call(["g++", "-shared", "-o", SO] + [o for _, o in jobs] + ["-lpthread", "-ldl", "-lrt"])
    return SO


if __name__ == "__main__":
    print(build(force="-B" in sys.argv, verbose=True))
