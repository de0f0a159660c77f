"""Resource usage of the built kernels (cuobjdump, no GPU needed): the one-wave property of the event kernels that
DESIGN.md §4 relies on — 72 registers, so that 28 warps (= chains) are resident per SM and 148 SMs hold the 4096 chains
of the bench workload at once — and shared memory that lets the CTAs of an SM be resident together."""
import os
import re
import shutil
import subprocess

import pytest

from jellyfysh_b200 import build

SMS, REGISTERS_PER_SM, SHARED_PER_SM, RESIDENT_WARPS = 148, 65536, 227 * 1024, 28


def resource_usage():
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    text = subprocess.run([tool, "-res-usage", build.LIBRARY], capture_output=True, text=True, check=True).stdout
    usage = {}
    for name, regs, shared in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:\d+ SHARED:(\d+)", text):
        usage[name] = (int(regs), int(shared))
    return usage


def test_event_kernels_keep_one_wave_of_the_bench_workload_resident():
    kernels = {name: use for name, use in resource_usage().items() if "12event_kernel" in name}
    assert kernels, "no event_kernel instantiation found in the library"
    for name, (registers, shared) in kernels.items():
        warps_per_cta = int(re.search(r"ELb[01]ELb[01]ELi(\d+)E", name).group(1))
        assert RESIDENT_WARPS % warps_per_cta == 0, name
        ctas = RESIDENT_WARPS // warps_per_cta
        assert registers * 32 * RESIDENT_WARPS <= REGISTERS_PER_SM, (name, registers)
        assert ctas * (shared + 1024) <= SHARED_PER_SM, (name, shared)   # 1 KB reserved per CTA
    assert SMS * RESIDENT_WARPS >= 4096


def test_every_kernel_fits_the_static_shared_memory_limit():
    for name, (_, shared) in resource_usage().items():
        assert shared <= 48 * 1024, (name, shared)
