"""Build liborbx.so (hand-written sm_100a CUDA kernels + C ABI) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "liborbx.so")
SOURCES = ["abi.cu", "pyramid.cu", "fast.cu", "octree.cu", "describe.cu", "knn.cu", "search_init.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--fmad=true", "-shared", "-cudart", "static", "-ldl"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "orbx.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def source_id():
    """sha256 over the kernel sources and the ABI header: the library reports it as orbx_build_id()."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "orbx.h")]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode()); h.update(fh.read())
    return h.hexdigest()[:16]


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ['-DORBX_BUILD_ID="%s"' % source_id()] + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", SO] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd, cwd=HERE)
    return SO


CPP_DIR = os.path.join(HERE, "cpp")
FRONTEND_SO = os.path.join(HERE, "liborbslam_frontend.so")


def build_cpp(force=False):
    """Host-side C++ mirror of the reference classes (ORBSlam::ORBextractor / ORBmatcher) over the C ABI."""
    srcs = [os.path.join(CPP_DIR, f) for f in ("ORBextractor.cc", "ORBmatcher.cc")]
    deps = srcs + [os.path.join(CPP_DIR, f) for f in ("ORBextractor.h", "ORBmatcher.h", "cv_compat.h")] + [SO]
    if not force and os.path.exists(FRONTEND_SO) and all(os.path.getmtime(FRONTEND_SO) >= os.path.getmtime(d) for d in deps):
        if not os.path.exists(os.path.join(os.path.dirname(HERE), "tools", "_build", "cpp_latency")):
            build_tools()
        return FRONTEND_SO
    cmd = ["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-Wall", "-I", CPP_DIR, "-o", FRONTEND_SO] + srcs + \
          ["-L", HERE, "-lorbx", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd, cwd=HERE)
    build_tools()
    return FRONTEND_SO


def build_tools():
    """tools/_build/cpp_latency: the C++ class timed one frame per call (bench.py's `latency.cpp_operator_call`)."""
    root = os.path.dirname(HERE)
    out = os.path.join(root, "tools", "_build")
    os.makedirs(out, exist_ok=True)
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-I", CPP_DIR, "-o", os.path.join(out, "cpp_latency"),
                           os.path.join(root, "tools", "cpp_latency.cpp"), "-L", HERE, "-lorbslam_frontend", "-lorbx", "-Wl,-rpath," + HERE])


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_cpp(force="--force" in sys.argv))
