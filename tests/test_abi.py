"""CPU-side checks of the product boundary: the C-ABI library loads, exports every symbol include/ass_b200.h declares,
refuses to compute without a GPU (no CPU fallback), and its host-side node generators reproduce the reference."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT, same_bits

A_, B_, C_ = 1.0, 0.7, 0.5


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ass_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(ass_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


@pytest.fixture(scope="module")
def A():
    import asteroid_spring_sims_b200 as pkg
    return pkg


def test_library_exports_every_declared_symbol(A):
    L = A.lib()
    names = declared_functions()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/ass_b200.h but not exported: %s" % missing
    assert L.ass_api_version() == 1


def test_struct_layouts_match_the_reference(A):
    # struct spring is 32 B {k, rs0, gamma, particle_1, particle_2} (src_spring/springs.h:16-22)
    assert A.SPRING_DTYPE.itemsize == 32
    assert [A.SPRING_DTYPE.fields[k][1] for k in ("k", "rs0", "gamma", "particle_1", "particle_2")] == [0, 8, 16, 24, 28]
    assert C.sizeof(A.Params) == 5 * 8 + 7 * 4 + 4  # padded to 8


def test_no_cpu_fallback(A):
    if A.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(A.AssError) as e:
        A.Sim()
    assert e.value.code == -3  # ASS_ENODEVICE
    with pytest.raises(A.AssError):
        A.probe_fp64_peak(16)


def test_host_generators_match_reference(A, body015, body010, lattices):
    """rand_ellipsoid / hcp / cubic are host code (sequential rand() stream): same seed -> same bits as the reference."""
    for g in (body015, body010):
        A.srand(12345)
        p = A.rand_ellipsoid(float(g["d"][0]), A_, B_, C_, 1.0)
        assert same_bits(p, g["gen"])
    assert same_bits(A.hcp_ellipsoid(0.2, A_, B_, C_, 1.0), lattices["hcp"])
    assert same_bits(A.cubic_ellipsoid(0.2, A_, B_, C_, 1.0), lattices["cubic"])


def test_host_generator_appends_and_rejects(A, body015):
    """A second call appends after i_0 and only checks distances against its own new particles (shapes.cpp:231-255)."""
    A.srand(12345)
    first = A.rand_ellipsoid(0.15, A_, B_, C_, 1.0)
    both = A.rand_ellipsoid(0.3, A_, B_, C_, 2.0, existing=first)
    assert same_bits(both[:len(first)], first)
    extra = both[len(first):]
    assert len(extra) > 0 and np.all(extra[:, 9] == 2.0 / len(extra))
    with pytest.raises(A.AssError):
        A.rand_ellipsoid(-1.0, A_, B_, C_, 1.0)


def test_host_generator_c2(A):
    path = os.path.join(ROOT, "tests", "golden", "c2_d00474.json")
    if not os.path.exists(path):
        pytest.skip("c2 fixture not generated")
    meta = json.load(open(path))
    A.srand(meta["seed"])
    p = A.rand_ellipsoid(meta["d"], A_, B_, C_, 1.0)
    assert len(p) == meta["N"]
    assert hashlib.sha256(p.tobytes()).hexdigest() == meta["sha256_gen_rows"]


def test_product_never_touches_the_oracle():
    """The product path may not import, link or call anything under oracle/ (task section 3)."""
    pkg = os.path.join(ROOT, "asteroid-spring-sims_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\boracle[/.]|ass_oracle|libref_oracle|\borc_[a-z]", text):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
