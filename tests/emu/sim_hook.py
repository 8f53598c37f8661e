"""tests only: route the package's C-ABI calls to the host SIMT-simulator build of the same .cu sources (tests/emu/build_emu.py),
so that kernel indexing logic is checked in the GPU-less container.  Lives here, not in the product package."""
import ctypes as C

import gnf_b200 as G


def install(path):
    G._lib._lib = G._lib._bind(C.CDLL(path))
    G._lib._SIMULATOR = True


def uninstall():
    G._lib._lib = None
    G._lib._SIMULATOR = False
