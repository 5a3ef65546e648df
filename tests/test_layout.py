"""CPU-side checks of the host logic behind the spring kernels: what ass_set_springs builds before anything runs on the
device -- spatial ranking, CSR rows, the lane-transposed copy of the half-edge list, the bank-group-aware row order and the
tile cut of the whole-body kernel (csrc/state.cu: ass_build_spring_layout, csrc/springs.cu: build_tiles_host).  The library
verifies its own invariants host-side (ass_selfcheck_spring_layout, no device needed); the GPU tests then only have to show
that the kernels that read these structures agree with the oracle and with each other."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def A():
    import asteroid_spring_sims_b200 as pkg
    return pkg


def springs_of(A, pairs, k=0.08, rs0=0.1, gamma=5.0):
    sp = np.zeros(len(pairs), dtype=A.SPRING_DTYPE)
    for q, (i, j) in enumerate(pairs):
        sp[q] = (k * (1 + q % 3), rs0 + 1e-3 * q, gamma, i, j)
    return sp


@pytest.mark.parametrize("grid,slots", [(148, 1 << 20), (1, 1 << 20), (2, 400), (3, 1000), (1, 64), (7, 200)])
def test_layout_invariants_on_reference_bodies(A, body015, body010, grid, slots):
    for g in (body015, body010):
        p, sp = g["start"], g["springs"]
        st = A.selfcheck_spring_layout(p, sp, grid=grid, slots=slots)
        assert st["violations"] == 0
        per_slice = 32 // A.spring_lanes()                            # nodes per slice (2 lanes per node: 16)
        assert st["slices"] == (len(p) + per_slice - 1) // per_slice
        assert st["records"] == 2 * len(sp) + st["padding"] and st["records"] % 64 == 0
        assert st["padding"] < 0.45 * 2 * len(sp)                     # small bodies pad most (C3: 11 %)
        assert 1 <= st["tiles"] <= st["slices"]
        if slots >= 1 << 20:
            assert st["tiles"] <= min(grid, st["slices"])             # one tile per CTA while tables fit


def test_row_order_spreads_gathers_over_the_bank_groups(A, body010):
    """Rows sorted by neighbour rank cost ~2.1 passes of the shared-memory / L1 data array per quarter-warp gather on this
    body (1 065 nodes, BASELINE C1); ranks refined inside blocks of eight plus the bank-group-aware row order: ~1.32."""
    st = A.selfcheck_spring_layout(body010["start"], body010["springs"])
    assert st["violations"] == 0
    assert st["gather_passes"] < 1.45 * st["gather_passes_min"]


def test_layout_edge_cases(A):
    rng = np.random.default_rng(5)
    def body(n):
        p = np.zeros((n, 11)); p[:, :3] = rng.uniform(-1, 1, (n, 3)); p[:, 9] = 1.0 / n
        return p
    # no springs at all; fewer nodes than one slice; a node count that is not a multiple of the slice size
    for n in (1, 3, 8, 9, 257):
        st = A.selfcheck_spring_layout(body(n), springs_of(A, []))
        assert st["violations"] == 0 and st["records"] == 0 and st["slices"] == (n + 32 // A.spring_lanes() - 1) // (32 // A.spring_lanes())
    # one spring; a chain; isolated nodes next to connected ones (point-mass perturbers carry no springs)
    assert A.selfcheck_spring_layout(body(2), springs_of(A, [(0, 1)]))["violations"] == 0
    assert A.selfcheck_spring_layout(body(50), springs_of(A, [(i, i + 1) for i in range(30)]))["violations"] == 0
    # a hub: one node of degree 299 (its slice alone touches 300 nodes, more than a small table holds), duplicates of a pair
    # (add_spring refuses them, ass_set_springs must still lay them out: both copies carry their own rest length)
    n = 300
    pairs = [(0, j) for j in range(1, n)] + [(5, 6), (5, 6), (6, 5)]
    for grid, slots in ((148, 1 << 20), (2, 64), (1, 128)):
        st = A.selfcheck_spring_layout(body(n), springs_of(A, pairs), grid=grid, slots=slots)
        assert st["violations"] == 0 and st["max_slots"] >= 300
    # complete graph on 40 nodes: every row longer than a slice's typical length
    pairs = [(i, j) for i in range(40) for j in range(i + 1, 40)]
    assert A.selfcheck_spring_layout(body(40), springs_of(A, pairs), grid=3, slots=64)["violations"] == 0


def test_layout_rejects_bad_springs(A):
    p = np.zeros((4, 11)); p[:, 0] = np.arange(4); p[:, 9] = 1.0
    for pairs in ([(1, 1)], [(0, 4)], [(-1, 2)]):                      # springs.cpp:57-59 throws on i == j
        with pytest.raises(A.AssError):
            A.selfcheck_spring_layout(p, springs_of(A, pairs))


def test_layout_invariants_on_random_graphs(A):
    """Random sparse and dense graphs on random point sets, several grids and table sizes: every invariant the kernels rely on
    (rows complete, transposed copy == CSR copy, padding points at the node itself, tiles partition the slices and fit in shared
    memory next to the record rings, renumbered records lead back to the same ranks)."""
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(2, 160), density=st.floats(0.01, 0.6), seed=st.integers(0, 10 ** 6), grid=st.sampled_from([1, 2, 5, 148]),
           slots=st.sampled_from([64, 200, 1000, 1 << 20]))
    def check(n, density, seed, grid, slots):
        rng = np.random.default_rng(seed)
        p = np.zeros((n, 11)); p[:, :3] = rng.uniform(-1, 1, (n, 3)); p[:, 9] = rng.uniform(0.5, 2.0, n)
        pairs = [(i, j) for i in range(n) for j in range(i + 1, n) if rng.random() < density]
        stt = A.selfcheck_spring_layout(p, springs_of(A, pairs), grid=grid, slots=slots)
        assert stt["violations"] == 0
        assert stt["records"] == 2 * len(pairs) + stt["padding"]

    check()
