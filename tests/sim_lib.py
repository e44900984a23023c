"""Loader for the host-simulator build of the kernel sources (TEST INFRASTRUCTURE, tests/hostsim).  It lets the kernel
logic be checked against the oracle on a machine without a GPU.  Never used by the product package."""
import os
import subprocess

from visioncortex_b200._ffi import Library
from visioncortex_b200.runtime import Context

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_DIR = os.path.join(ROOT, "tests", "hostsim")
SIM_SO = os.path.join(SIM_DIR, "libvcb200_sim.so")

_CTX = None


def load_sim() -> Context:
    global _CTX
    if _CTX is None:
        subprocess.check_call(["make", "-C", SIM_DIR, "-s"])
        _CTX = Context(0, lib=Library(SIM_SO))
    return _CTX
