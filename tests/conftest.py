import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# virtual ranks (tests/test_virtual_ranks.py): the streams of different ranks must not share a hardware queue; the
# variable is read when CUDA initialises, so it has to be set before the first test touches the library
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and every kernel must be loaded up front: the lazy loading of a kernel on its first launch synchronises the
# context, which would wait for a peer rank's spinning kernel (cmb_vgroup_create refuses to run under lazy loading)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
