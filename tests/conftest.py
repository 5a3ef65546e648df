import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
# The in-process sharding tests drive up to 8 "ranks" x 2 streams on one GPU, and a rank's stream may sit in a kernel that
# waits for another rank's flag: with the default 8 hardware queues two such streams could share one and block each other.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and all of them are queued by ONE host thread: a lazily loaded kernel's first launch synchronises the context, which
# would wait for a flag wait whose producer this thread has not queued yet (ass_shard_import_local insists on this)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/ass_oracle.c). Test infrastructure only."""
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_available():
    so = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")
    return os.path.exists(so) or os.path.isdir("/root/reference/src_spring")


@pytest.fixture()
def ref(ref_available):
    """The unmodified reference compiled into oracle/_ref (skips where it is absent)."""
    if not ref_available:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    from oracle.bindings import Ref
    r = Ref()
    yield r
    r.close()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def kat3():
    return load_golden("kat3")


@pytest.fixture(scope="session")
def body015():
    return load_golden("ellipsoid_d015")


@pytest.fixture(scope="session")
def body010():
    return load_golden("c1_d010")


@pytest.fixture(scope="session")
def lattices():
    return load_golden("lattices")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same_bits(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def kat_particles():
    """SpringTest fixture, modules/tests/tests.cpp:60-118."""
    P = np.zeros((3, 11))
    P[0, :3] = (0, 0, 0)
    P[1, :3] = (1, 2, 2)
    P[2, :3] = (2, 3, 6)
    P[:, 3:6] = P[:, :3]
    P[:, 9] = (1, 2, 3)
    return P


@pytest.fixture(scope="session")
def gpu():
    """The CUDA library with a device selected. Fails loudly (no fallback) when either is missing."""
    import asteroid_spring_sims_b200 as A
    A.init(int(os.environ.get("LOCAL_RANK", "0")))
    return A


@pytest.fixture(scope="session")
def orb_golden():
    return load_golden("orb")
