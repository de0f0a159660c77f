"""The C-ABI boundary without a GPU: libecmc_b200.so loads, exports every symbol include/ecmc.h declares, its struct
layouts agree with the ctypes mirror, host-only entry points work, and compute entry points fail loudly (no CPU
fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from jellyfysh_b200 import abi, engine
from jellyfysh_b200.program import ProgramBuilder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ecmc.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ecmc_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = engine.library()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    assert lib.ecmc_abi_version() == abi.ECMC_ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof from the C compiler against the ctypes structures."""
    fields = {"EcmcPotential": ["kind", "params"], "EcmcWalkerTable": ["n_entries", "cell_a", "rate_a", "mean_rate"],
              "EcmcVetoTables": ["upper", "lower", "bounds"],
              "EcmcProgram": ["dimension", "system_length", "cells_per_side", "pair_handler", "pair_potential",
                              "pair_bounding_potential", "veto_enabled", "veto_potential", "veto_target_charge",
                              "veto_tables", "chain_time", "initial_active", "seed", "nodes_per_root", "bonds",
                              "bond_potential", "cell_level", "composite_lifting", "inter_factors", "inter_potential",
                              "bending_children", "bending_separations", "boundary_keeps_factors", "bending_potential",
                              "bending_offset", "bending_max_displacement"],
              "EcmcChainState": ["active", "time_q", "eoc_next_active", "event_counter", "stream", "pending_q",
                                 "pending_stamp_r", "pending_root_position", "kept_kind", "kept_q", "kept_rate",
                                 "kept_stamp_r"],
              "EcmcEventRecord": ["kind", "n_candidates", "time_q", "active_pos"],
              "EcmcStats": ["events", "candidates", "capacity_errors", "bond_events", "factor_pair_events", "pair_targets"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % HEADER, "int main(void) {"]
    for struct, names in fields.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (struct, struct))
        for name in names:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (struct, name, struct, name))
    lines.append("return 0; }")
    source = tmp_path / "layout.c"
    source.write_text("\n".join(lines))
    binary = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(binary), str(source)], check=True)
    out = subprocess.run([str(binary)], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        key, value = line.split()
        if "." in key:
            struct, name = key.split(".")
            assert getattr(getattr(abi, struct), name).offset == int(value), key
        else:
            assert C.sizeof(getattr(abi, key)) == int(value), key
    assert abi.record_dtype().itemsize == C.sizeof(abi.EcmcEventRecord)
    assert abi.chain_state_dtype().itemsize == C.sizeof(abi.EcmcChainState)


def test_host_random_stream_matches_oracle(oracle):
    for seed, stream, event, slot in [(0, 0, 0, 0), (7, 3, 12345, abi.slot(abi.SLOT_PAIR_TIME, 77)),
                                      (0xdeadbeef, 4095, (1 << 40) + 17, abi.slot(abi.SLOT_VETO_CHOICE))]:
        assert np.array_equal(engine.random_words(seed, stream, event, slot, 0, 13),
                              oracle.random_words(seed, stream, event, slot, 0, 13))
        assert np.array_equal(engine.random_doubles(seed, stream, event, slot, 3, 9),
                              oracle.random_doubles(seed, stream, event, slot, 3, 9))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="checks the behaviour WITHOUT a CUDA device")
def test_no_cpu_fallback_without_a_device():
    builder = ProgramBuilder(3, 8, 4.0, 1.0, [4, 4, 4], 1, chain_time=1.0)
    builder.set_pair(abi.PAIR_TWO_LEAF_UNIT, abi.EcmcPotential.make(abi.POT_LENNARD_JONES, 4.0, 1.0))
    with pytest.raises(engine.EcmcError) as error:
        engine.Engine(builder, n_chains=2)
    assert error.value.status == abi.ECMC_ERR_CUDA and "no CPU fallback" in str(error.value)
    potential = abi.EcmcPotential.make(abi.POT_LENNARD_JONES, 4.0, 1.0)
    with pytest.raises(engine.EcmcError) as error:
        engine.potential_derivative(potential, 3, 4.0, 0, [[1.0, 0.2, 0.1]])
    assert error.value.status == abi.ECMC_ERR_CUDA


def test_argument_errors_are_reported_before_any_device_work():
    lib = engine.library()
    handle = C.c_void_p()
    assert lib.ecmc_create(None, 0, 1, C.byref(handle)) == abi.ECMC_ERR_INVALID
    assert b"NULL" in lib.ecmc_last_error(None)
    builder = ProgramBuilder(3, 8, 4.0, 1.0, [4, 4, 4], 1, chain_time=1.0)
    assert lib.ecmc_create(C.byref(builder.program), 0, 0, C.byref(handle)) == abi.ECMC_ERR_INVALID
    potential = abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, 1.0, 3.45, 6, 2)
    with pytest.raises(engine.EcmcError) as error:  # a Coulomb derivative needs three dimensions
        engine.potential_derivative(potential, 2, 1.0, 0, [[0.1, 0.2]])
    assert error.value.status == abi.ECMC_ERR_INVALID
    hard = abi.EcmcPotential.make(abi.POT_HARD_SPHERE, 0.5)
    with pytest.raises(engine.EcmcError):  # hard spheres have no derivative
        engine.potential_derivative(hard, 3, 1.0, 0, [[0.1, 0.2, 0.3]])
