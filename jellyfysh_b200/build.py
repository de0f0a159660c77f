"""Build libecmc_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m jellyfysh_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBRARY = os.path.join(HERE, "libecmc_b200.so")
SOURCES = [os.path.join(CSRC, "ecmc_engine.cu")]
HEADERS = [os.path.join(CSRC, name) for name in ("ecmc_math.cuh", "ecmc_program.cuh", "ecmc_kernels.cuh", "ecmc_molecules.cuh", "ecmc_spec.cuh", "ecmc_spec_cta.cuh", "ecmc_disks.cuh", "ecmc_log_table.cuh")] + \
          [os.path.join(ROOT, "include", "ecmc.h")]


def nvcc_path() -> str:
    for candidate in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if candidate and os.path.exists(candidate):
            return candidate
    raise RuntimeError("nvcc not found: libecmc_b200.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIBRARY):
        return True
    built = os.path.getmtime(LIBRARY)
    return any(os.path.getmtime(path) > built for path in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, output: str = None, defines=()) -> str:
    """Compile the library if it is missing or older than its sources; returns its path.
    `output` / `defines` build a tuning variant next to the default library."""
    if output is None and not force and not is_stale():
        return LIBRARY
    command = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-cudart", "shared",
               "-o", output or LIBRARY] + ["-D" + d for d in defines] + SOURCES
    if verbose:
        command.insert(1, "-Xptxas=-v")
    result = subprocess.run(command, capture_output=True, text=True)
    if verbose or result.returncode != 0:
        sys.stderr.write(result.stdout + result.stderr)
    if result.returncode != 0:
        raise RuntimeError("nvcc failed building libecmc_b200.so")
    return output or LIBRARY


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
