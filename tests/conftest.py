import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Builds the native pieces once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge

    ge.build()
    return True


@pytest.fixture(scope="session")
def gsv(built):
    import gsv_b200

    return gsv_b200


@pytest.fixture(scope="session")
def orc(built):
    from oracle import oracle

    oracle.lib()
    return oracle


_streams = {}


@pytest.fixture(scope="session")
def circuit(gsv, orc):
    """circuit(name) -> (Program, oracle Stream), cached."""

    def get(name):
        if name not in _streams:
            p = gsv.Program(name)
            t, a, b, c, outs, nw = p.flat_stream()
            _streams[name] = (p, orc.Stream(t, a, b, c, outs, nw, p.n_inputs))
        return _streams[name]

    return get


_lane_programs = {}


@pytest.fixture(scope="session")
def lane_program(gsv):
    """lane_program(name) -> Program(name, lane_only=True), cached for the session (the verifier takes ~25 s to record
    and ~10 GB; two CPU tests look at it)."""

    def get(name):
        if name not in _lane_programs:
            _lane_programs[name] = gsv.Program(name, lane_only=True)
        return _lane_programs[name]

    return get
