"""TEST INFRASTRUCTURE ONLY: builds the C++ host side of curvine_b200 against the mock CUDA runtime in this directory
(+ CPU stand-ins for the cvk_* launchers) into a library under /tmp, so the ingest pipeline's host logic can run on a
machine without a GPU (and under ASan/TSan).  csrc/kernels.cu is NOT part of it.  Nothing under curvine_b200/ knows
about this library; the product library is built by curvine_b200/build.py with nvcc against the real runtime."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HOST = os.path.join(ROOT, "curvine_b200", "csrc", "host")


def sources():
    out = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith((".cc", ".cu")) and f != "gds.cc"]  # no cuFile here: mock_cuda.cc has the stub
    return out + [os.path.join(HERE, "mock_cuda.cc"), os.path.join(HERE, "mock_cvk.cc")]


def _digest(extra):
    h = hashlib.sha256(extra.encode())
    for root in (HOST, os.path.join(ROOT, "curvine_b200", "csrc"), os.path.join(ROOT, "include"), HERE):
        for f in sorted(os.listdir(root)):
            p = os.path.join(root, f)
            if os.path.isfile(p) and f.endswith((".cc", ".cu", ".h")):
                h.update(p.encode())
                h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def build(sanitize: str = "") -> str:
    """sanitize: "" | "address,undefined" | "thread".  Returns the path of the built library (cached by source digest)."""
    flags = ["-O1", "-g", "-std=c++17", "-fPIC", "-pthread", "-msse4.2", "-Wall", "-Wno-unknown-pragmas", "-I", HERE, "-I", os.path.join(ROOT, "include"),
             "-I", os.path.join(ROOT, "curvine_b200", "csrc")]
    if sanitize:
        flags += ["-fsanitize=" + sanitize, "-fno-omit-frame-pointer"]
    out_dir = os.path.join("/tmp", "cv_mock_" + _digest(" ".join(flags) + open(os.path.abspath(__file__)).read()))
    lib = os.path.join(out_dir, "libcurvine_b200_mock.so")
    if os.path.exists(lib):
        return lib
    os.makedirs(out_dir, exist_ok=True)
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(out_dir, os.path.basename(src) + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen(["g++"] + flags + ["-x", "c++", "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            raise RuntimeError("g++ failed for %s:\n%s" % (src, out.decode()))
    # -Bsymbolic: the mock's cuda*/cvk_* definitions must win inside this library even when the process has the real
    # libcudart (torch) or the product library loaded with RTLD_GLOBAL
    link = ["g++", "-shared", "-Wl,-Bsymbolic", "-o", lib + ".tmp"] + objs + ["-lpthread", "-ldl", "-lrt"] + (["-fsanitize=" + sanitize] if sanitize else [])
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode:
        raise RuntimeError("link failed:\n" + r.stdout.decode())
    os.replace(lib + ".tmp", lib)
    return lib


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else ""))
