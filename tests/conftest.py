import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_sessionstart(session):
    """The host side loads libxvec_b200.so even without a GPU (native ark index of the reader, exported-symbol checks):
    build it once when the in-tree library is absent or older than its sources (nvcc cross-compiles without a GPU;
    a no-op otherwise).  A failing build is reported by the tests that need the library, not here."""
    try:
        from xvector_b200 import _native
        _native.build_library()
    except Exception as err:                      # noqa: BLE001
        print("conftest: could not build libxvec_b200.so: %s" % err, file=sys.stderr)
