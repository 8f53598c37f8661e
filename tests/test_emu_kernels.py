"""CPU tests: the kernels' logic, executed by the host SIMT simulator build of the same .cu sources
(tests/emu, csrc/cpu_emu.h), against the reference-generated golden vectors.  This is a debugging aid for
the GPU-less build container; the GPU parity tests (test_gpu_parity.py, -m gpu) are the real gate."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

import gnf_b200 as G
from helpers import golden_names
import parity


@pytest.fixture(scope="module")
def emu():
    import build_emu
    path = build_emu.build()
    G._lib._install_simulator_for_tests(path)
    yield
    G._lib._uninstall_simulator_for_tests()


@pytest.mark.parametrize("name", golden_names())
def test_golden_case_in_simulator(emu, name):
    torch.set_num_threads(1)
    parity.run_case(name, "cpu")


@pytest.mark.parametrize("name", [n for n in golden_names() if "mono" in n])
def test_golden_monotonic_case_layerwise_engine_in_simulator(emu, name):
    """The layer-wise UMNN engine (umnn_lw.cu) with the FFMA GEMMs: per-layer passes, once-per-row conditioning half of
    the first layer, first-layer reductions, against the reference's golden vectors."""
    torch.set_num_threads(1)
    G.ops.UMNN_ENGINE = "layerwise"
    try:
        parity.run_case(name, "cpu")
    finally:
        G.ops.UMNN_ENGINE = "auto"
