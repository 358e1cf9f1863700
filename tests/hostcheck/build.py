"""Builds tests/hostcheck/libhostcheck.so (test fixture; see hostcheck.cu)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libhostcheck.so")
SRC = os.path.join(HERE, "hostcheck.cu")
DEPS = [SRC] + [os.path.join(HERE, "..", "..", "sidm-nbody_b200", "csrc", f) for f in ("build_logic.h", "tree_logic.h")]


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
                           "-gencode", "arch=compute_100a,code=sm_100a", SRC, "-o", OUT])
    return OUT


if __name__ == "__main__":
    print(build())
