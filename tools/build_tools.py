"""Build the measurement helpers under tools/ (sm_100a)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from jellyfysh_b200.build import nvcc_path  # noqa: E402


def build():
    src, lib = os.path.join(HERE, "fp64_peak.cu"), os.path.join(HERE, "libfp64_peak.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared",
                        "-Xcompiler", "-fPIC", "-cudart", "shared", "-o", lib, src], check=True)
    return lib


if __name__ == "__main__":
    build()
