"""jellyfysh_b200 -- B200-native batched event-chain Monte Carlo engine behind JeLLyFysh's plugin surface.

engine      ctypes binding of libecmc_b200.so (include/ecmc.h), the CUDA library of the hot path
program     EcmcProgram assembly, tables: init-time cell-veto tables on the device, workloads: synthetic configs
compiler    JeLLyFysh object graph (factory-built from INI) -> EcmcProgram
mediator    CudaBatchedMediator, the drop-in for jellyfysh.mediator.single_process_mediator
"""
import importlib
import os
import sys


def install():
    """Make `[Run] mediator = cuda_batched_mediator` resolvable by the reference's factory
    (jellyfysh/base/factory.py:111-122 imports jellyfysh.mediator.<snake_case_name>) without copying a file into
    the jellyfysh tree: the module is registered under that name. Needs `jellyfysh` importable.
    Where MDAnalysis is not installed, a reader for .pdb start configurations takes its place (shims/MDAnalysis), so the
    shipped configuration files with a `pdb_input_handler` run unchanged."""
    try:
        importlib.import_module("MDAnalysis")
    except ImportError:
        shims = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
        if shims not in sys.path:
            sys.path.append(shims)
    module = importlib.import_module("jellyfysh_b200.mediator.cuda_batched_mediator")
    sys.modules["jellyfysh.mediator.cuda_batched_mediator"] = module
    import jellyfysh.mediator
    jellyfysh.mediator.cuda_batched_mediator = module
    # `state_handler = cuda_state_handler` (optional): the state-handler contract over the chains of the device engines
    handler = importlib.import_module("jellyfysh_b200.state_handler.cuda_state_handler")
    sys.modules["jellyfysh.state_handler.cuda_state_handler"] = handler
    import jellyfysh.state_handler
    jellyfysh.state_handler.cuda_state_handler = handler
    return module
