"""CPU-side checks of the host logic behind the pair-once gravity kernel (csrc/gravity_sym.cu: build_sym_cut): how two
particle ranges are cut into blocks (the partial block leads, its padding in front), which kernel variant runs (three or
four residents per thread), how tiles are cut into pieces for small problems, and the CTA list.  The library checks its own
plan host-side (ass_selfcheck_gravity_plan, no device needed): every (resident block, REAL streamed particle) pair must be
covered by exactly one CTA.  The GPU tests then show that the kernels agree with the oracle."""
import pytest


@pytest.fixture(scope="module")
def A():
    import asteroid_spring_sims_b200 as pkg
    return pkg


# BASELINE sizes (C1, C2, C3, C4 + point mass), block boundaries, the slabs of C3 / 8 and C4 / 8
SIZES = [1, 31, 32, 33, 767, 768, 769, 1023, 1024, 1025, 1065, 9327, 12067, 12288, 10520, 96536, 118200, 945597]


@pytest.mark.parametrize("n", SIZES)
def test_triangle_plans_cover_every_pair_once(A, n):
    st = A.selfcheck_gravity_plan(n)
    assert st["violations"] == 0
    assert st["residents"] in (3, 4) and st["resident_blocks"] == st["streamed_blocks"] == -(-n // (256 * st["residents"]))
    assert st["ctas"] >= st["resident_blocks"]
    if n >= 90000:
        assert st["residents"] == 3          # large problems: blocks of 768, three CTAs per SM
    if n <= 20000:
        assert st["residents"] == 4 and st["tiles_per_cta"] == 1


@pytest.mark.parametrize("n_res,n_str", [(12288, 36864), (6144, 12288), (12288, 6144), (10520, 12288), (118200, 354600), (59100, 118200),
                                         (500, 700), (1, 1), (1025, 31), (96536, 1)])
def test_rectangle_plans_cover_every_pair_once(A, n_res, n_str):
    st = A.selfcheck_gravity_plan(n_res, n_str, rect=True)
    assert st["violations"] == 0
    p = 256 * st["residents"]
    assert st["resident_blocks"] == -(-n_res // p) and st["streamed_blocks"] == -(-n_str // p)


def test_pieces_in_the_leading_padding_get_no_cta(A):
    """A body of 9 327 particles is 9.1 blocks of 1024: block 0 holds 111 real particles behind 913 dummies, so of the eight
    128-particle pieces of each of its ten tiles as the streamed side seven have nothing to do."""
    st = A.selfcheck_gravity_plan(9327)
    assert st["cut"] == 8 and st["skipped"] == 70 and st["ctas"] == 55 * 8 - 70
    st = A.selfcheck_gravity_plan(12288)      # a slab of C3 / 8: whole blocks, nothing to skip
    assert st["skipped"] == 0 and st["ctas"] == 78 * 8


def test_forced_variants(A, monkeypatch):
    for r in ("3", "4"):
        monkeypatch.setenv("ASS_SYM_R", r)
        for n in (9327, 96536):
            st = A.selfcheck_gravity_plan(n)
            assert st["violations"] == 0 and st["residents"] == int(r)
    monkeypatch.delenv("ASS_SYM_R")
    for c in ("1", "2", "4", "8", "16", "32"):
        monkeypatch.setenv("ASS_SYM_CUT", c)
        st = A.selfcheck_gravity_plan(9327)
        assert st["violations"] == 0 and st["cut"] == int(c)


def test_random_sizes(A):
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(n_res=st.integers(1, 60000), n_str=st.integers(1, 60000), rect=st.booleans(), sms=st.sampled_from([4, 132, 148]))
    def check(n_res, n_str, rect, sms):
        assert A.selfcheck_gravity_plan(n_res, n_str, rect=rect, sm_count=sms)["violations"] == 0

    check()


# ---------------------------------------------------------------------------------------------------
# a sharded body: gravity divided by pairs of slabs (csrc/shard.cu: shard_pieces_of, partial_sources_of)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("n,n_perts", [(96536, 0), (945597, 1), (1066, 1), (300, 0), (7, 0), (1, 0)])
def test_every_pair_of_a_sharded_body_is_evaluated_once(A, nranks, n, n_perts):
    """Over all ranks: own triangles + rectangles against the next (P-1)/2 slabs + the halves of the antipodal rectangles cover
    every unordered pair of body particles exactly once, and the partial accelerations a rank adds are the ones its peers
    compute for its slab (a peer with an empty slab contributes the zeros its window was created with)."""
    for align in (1024, 256, 1):
        bounds = A.slab_bounds(n, nranks, align=align)
        st = A.selfcheck_shard_division(bounds, n - n_perts)
        assert st["violations"] == 0, (bounds, st)
        nb = n - n_perts
        assert st["most_pairs"] >= nb * (nb - 1) // 2 // nranks
    if n >= 90000 and nranks > 1:
        st = A.selfcheck_shard_division(A.slab_bounds(n, nranks), n - n_perts)
        assert st["fewest_pairs"] > 0.8 * st["most_pairs"]          # the same work on every rank (block-aligned slabs: the last one is shorter)


def test_shard_division_random_bounds(A):
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=80, deadline=None)
    @given(cuts=st.lists(st.integers(0, 5000), min_size=0, max_size=7), n=st.integers(0, 5000), perts=st.integers(0, 3))
    def check(cuts, n, perts):
        bounds = [0] + sorted(min(c, n) for c in cuts) + [n]
        assert A.selfcheck_shard_division(bounds, max(n - perts, 0))["violations"] == 0

    check()
