"""Score kernels (K6 cosine GEMM, K7 clarity, K8 2-means polysemanticity, K9 redundancy) through the public
``semanticlens_b200.scores`` / ``Lens`` API against fixtures recorded from the imported reference and against the
oracle on seeded inputs. Tolerance: 1e-4 relative (max-norm) for fp32 outputs as BASELINE.json's north_star states;
polysemanticity additionally to 1e-9 absolute on the k-means path (same f64 arithmetic on the Gram matrix)."""

import numpy as np
import pytest
import torch

from oracle import polysem as P
from oracle import ref_port as rp
from tests.polysem_cases import CASES, make_case

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def maxnorm_close(got, ref, rtol=RTOL):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max()
    assert err <= rtol * scale, f"max err {err:.3e} > {rtol} * {scale:.3e}"


@pytest.fixture(scope="module")
def S():
    from semanticlens_b200 import scores

    return scores


def test_golden_similarity_all_branches(golden, S):
    z = np.load(golden / "scores.npz")
    t = lambda k: torch.from_numpy(z[k])  # noqa: E731
    out = S.similarity_score(t("sim_x"), t("sim_y"))
    assert out.device.type == "cpu" and out.dtype == torch.float32
    maxnorm_close(out.numpy(), z["sim_xy"])
    maxnorm_close(S.similarity_score(t("sim_x"), t("sim_y2")).numpy(), z["sim_xy2"])  # C == D: no transpose upstream
    maxnorm_close(S.similarity_score(t("sim_x3"), t("sim_y")).numpy(), z["sim_x3y"])  # equal shapes: row-wise
    with pytest.raises(ValueError, match="same shape"):
        S.similarity_score(torch.randn(4, 8), torch.randn(5, 9))


def test_feature_widths_that_are_not_multiples_of_four(S):
    """The reference takes any D; the kernels read 16-byte vectors, so the wrappers zero-pad."""
    g = torch.Generator().manual_seed(4)
    for D in (5, 30, 511):
        V = torch.randn(7, 6, D, generator=g)
        maxnorm_close(S.clarity_score(V).numpy(), rp.clarity_score(V).numpy())
        x, y = torch.randn(3, D, generator=g), torch.randn(9, D, generator=g)
        maxnorm_close(S.similarity_score(x, y).numpy(), rp.similarity_score(x, y).numpy())
        maxnorm_close(S.redundancy_score(y).numpy(), rp.redundancy_score(y).numpy())
    with pytest.raises(ValueError, match="at most 2048"):
        S.clarity_score(torch.randn(2, 3, 2052))


def test_similarity_with_a_batched_left_operand(S):
    """x (A, Q, D) against y (C, D): the reference's second branch fires when x.shape[1] == y.shape[1] and matmul
    broadcasts over the leading dimension."""
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(3, 16, 16, generator=g), torch.randn(9, 16, generator=g)
    want = rp.similarity_score(x, y)
    got = S.similarity_score(x, y)
    assert got.shape == want.shape == (3, 16, 9)
    maxnorm_close(got.numpy(), want.numpy())
    with pytest.raises(ValueError, match="same shape"):
        S.similarity_score(torch.randn(3, 16, 16), torch.randn(2, 9, 16))


def test_embeddings_in_half_precision_are_gathered(S):
    from semanticlens_b200 import ops

    table = torch.randn(11, 32, device="cuda").half()
    idx = torch.tensor([[0, 5, -1], [10, 2, 2]])
    assert torch.equal(ops.gather_rows(table, idx).cpu(), table.float().cpu()[idx])


@pytest.mark.parametrize("n,D", [(1, 8), (2, 64), (131, 96), (1000, 512), (4099, 320)])
def test_redundancy_fused_rowmax_vs_reference_port(S, n, D):
    """K9: the row maxima come out of the GEMM epilogue (atomics over column tiles); sizes that are not multiples of the
    8-row padding or of the tile sizes, a single row (the diagonal alone: cos - 2 = -1), duplicates (maximum = 1)."""
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, D, generator=g)
    if n > 100:
        x[7] = 3 * x[90]  # a duplicate direction: row maximum exactly 1 for both
    want = rp.redundancy_score(x)
    got = S.redundancy_score(x.cuda())
    assert got.shape == () and got.dtype == torch.float32
    maxnorm_close(got.cpu().numpy(), want.numpy())


def test_redundancy_nan_and_batched_and_empty(S):
    x = torch.randn(40, 32, generator=torch.Generator().manual_seed(1))
    x[3, 5] = float("nan")
    assert torch.isnan(S.redundancy_score(x.cuda())).item() and torch.isnan(rp.redundancy_score(x)).item()
    xb = torch.randn(3, 17, 24, generator=torch.Generator().manual_seed(2))
    maxnorm_close(S.redundancy_score(xb).numpy(), rp.redundancy_score(xb).numpy())
    with pytest.raises(RuntimeError):
        S.redundancy_score(torch.empty(0, 16).cuda())


def test_golden_clarity_redundancy(golden, S):
    z = np.load(golden / "scores.npz")
    maxnorm_close(S.clarity_score(torch.from_numpy(z["V"]).cuda()).cpu().numpy(), z["clarity"])
    maxnorm_close(S.redundancy_score(torch.from_numpy(z["red_in"])).numpy(), z["red"])
    out = S.redundancy_score(torch.from_numpy(z["red2_in"]).cuda())
    assert out.ndim == 0 and out.is_cuda
    maxnorm_close(out.cpu().numpy(), z["red2"])


def test_golden_polysemanticity_small(golden, S):
    z = np.load(golden / "scores.npz")
    V = torch.from_numpy(z["P"])
    out = S.polysemanticity_score(V)
    assert out.dtype == torch.float64 and out.shape == (10,) and out.device.type == "cpu"
    np.testing.assert_allclose(out.numpy(), z["poly"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(S.polysemanticity_score(V, replace_empty_clusters=False).numpy(), z["poly_noreplace"], rtol=0,
                               atol=1e-9)


@pytest.mark.parametrize("name", sorted(CASES))
def test_polysemanticity_cases_match_reference(golden, S, name):
    z = np.load(golden / "scores_poly.npz")
    V = torch.from_numpy(make_case(name)).cuda()
    got = S.polysemanticity_score(V).cpu().numpy()
    ref = z[f"{name}.poly"]
    mism = np.abs(got - ref) > 2e-6
    assert not mism.any(), f"{mism.sum()} of {len(ref)} neurons differ: {got[mism]} vs {ref[mism]}"
    got_nr = S.polysemanticity_score(V, replace_empty_clusters=False).cpu().numpy()
    np.testing.assert_allclose(got_nr, z[f"{name}.poly_noreplace"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", sorted(CASES))
def test_clarity_cases_match_reference(golden, S, name):
    z = np.load(golden / "scores_poly.npz")
    if CASES[name][2] % 4:
        pytest.skip("K7 needs D % 4 == 0")
    got = S.clarity_score(torch.from_numpy(make_case(name)).cuda()).cpu().numpy()
    ref = z[f"{name}.clarity"]
    # clarity cancels to ~0 for unclear neurons: absolute floor of 1e-4 * the score's natural scale (1)
    assert np.abs(got - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1.0) * 0.1, np.abs(got - ref).max()


def test_polysemanticity_many_neurons_vs_oracle(S):
    """More neurons than resident CTAs (grid-stride path, workspace slot reuse), checked against the Gram-form oracle."""
    rng = np.random.default_rng(11)
    V = rng.standard_normal((700, 24, 40)).astype(np.float32)
    V[::3, ::2] += rng.standard_normal((234, 1, 40)).astype(np.float32) * 2
    got = S.polysemanticity_score(torch.from_numpy(V).cuda()).cpu().numpy()
    ref = P.polysemanticity_gram(V)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)


def test_polysemanticity_zeros_subnormals_and_non_finite_examples(S):
    """Zeros, signed zeros and subnormals are ordinary values; an Inf / NaN example makes that neuron's score NaN (sklearn
    raises on such input) and leaves the neighbouring neurons alone."""
    rng = np.random.default_rng(5)
    V = rng.standard_normal((6, 32, 48)).astype(np.float32)
    V[:, ::2] += 2 * rng.standard_normal((6, 1, 48)).astype(np.float32)
    V[0, :, ::3] = 0.0                    # exact zeros (a ReLU'd embedding)
    V[1, 3, 5] = np.float32(1e-41)        # a subnormal
    V[1, 4, :8] = -0.0
    ref = P.polysemanticity_gram(V)
    bad = V.copy()
    bad[2, 7, 1] = np.inf
    bad[4, 0, 0] = np.nan
    got = S.polysemanticity_score(torch.from_numpy(V).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)
    got_bad = S.polysemanticity_score(torch.from_numpy(bad).cuda()).cpu().numpy()
    assert np.isnan(got_bad[[2, 4]]).all()
    np.testing.assert_allclose(got_bad[[0, 1, 3, 5]], ref[[0, 1, 3, 5]], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", list(__import__("tests.polysem_cases", fromlist=["GENERAL_CASES"]).GENERAL_CASES))
def test_polysemanticity_general_cases_match_reference(golden, S, name):
    """n_clusters = 2..8 and more than 256 examples per neuron (K8g, slb_polysem_kmeans) against outputs recorded from the
    imported reference (sklearn KMeans per neuron), and against the sample-space oracle."""
    from tests.polysem_cases import GENERAL_CASES, make_general_case

    z = np.load(golden / "scores_poly_general.npz")
    m = GENERAL_CASES[name][4]
    V = make_general_case(name)
    got = S.polysemanticity_score(torch.from_numpy(V).cuda(), n_clusters=m).cpu().numpy()
    np.testing.assert_allclose(got, z[f"{name}.poly"], rtol=0, atol=2e-6)
    got_nr = S.polysemanticity_score(torch.from_numpy(V).cuda(), replace_empty_clusters=False, n_clusters=m).cpu().numpy()
    np.testing.assert_allclose(got_nr, z[f"{name}.poly_noreplace"], rtol=0, atol=1e-9)


def test_polysemanticity_general_kernel_agrees_with_the_fast_kernel(S):
    """Two clusters, <= 256 examples: both kernels implement the same fit."""
    from semanticlens_b200 import ops

    rng = np.random.default_rng(21)
    V = rng.standard_normal((40, 48, 32)).astype(np.float32)
    V[::2, ::2] += 2 * rng.standard_normal((20, 1, 32)).astype(np.float32)
    Vg = torch.from_numpy(V).cuda()
    np.testing.assert_allclose(ops.polysem_kmeans(Vg, 2).cpu().numpy(), ops.polysem_2means(Vg).cpu().numpy(), rtol=0, atol=1e-9)


def test_polysemanticity_general_errors(S):
    V = torch.randn(2, 3, 8).cuda()
    with pytest.raises(ValueError, match="n_samples=3 should be >= n_clusters=4"):
        S.polysemanticity_score(V, n_clusters=4)
    bad = torch.randn(3, 300, 8)
    bad[1, 5, 2] = float("nan")
    got = S.polysemanticity_score(bad.cuda()).cpu()
    assert torch.isnan(got[1]) and torch.isfinite(got[[0, 2]]).all()


def test_cosine_gemm_shapes_vs_port(S):
    g = torch.Generator().manual_seed(3)
    for (Q, C, D) in ((1, 10, 128), (7, 1000, 512), (300, 129, 768), (130, 260, 100)):
        x, y = torch.randn(Q, D, generator=g), torch.randn(C, D, generator=g) * 3
        ref = rp.similarity_score(x, y)
        got = S.similarity_score(x.cuda(), y.cuda())
        assert got.is_cuda and got.shape == (Q, C)
        maxnorm_close(got.cpu().numpy(), ref.numpy())
        # fp32-grade, not just 1e-4: the 3-pass split GEMM carries 22-bit operands
        assert (got.cpu() - ref).abs().max() < 5e-6


def test_clarity_full_size_property(S):
    """BASELINE cfg-5-shaped rows (k = 256, D = 512): clarity of identical rows is 1, of an orthonormal set is 0
    (size-independent properties), plus agreement with the port on a random slab."""
    k, D = 256, 512
    same = torch.randn(1, 1, D).expand(3, k, D).contiguous().cuda()
    np.testing.assert_allclose(S.clarity_score(same).cpu().numpy(), 1.0, atol=1e-5)
    eye = torch.eye(D)[:k].unsqueeze(0).contiguous().cuda()
    np.testing.assert_allclose(S.clarity_score(eye).cpu().numpy(), 0.0, atol=1e-6)
    V = torch.randn(64, k, D, generator=torch.Generator().manual_seed(4))
    maxnorm_close(S.clarity_score(V.cuda()).cpu().numpy(), rp.clarity_score(V).numpy(), rtol=1e-3)


def test_lens_eval_and_probe_dispatch():
    from semanticlens_b200.lens import Lens, _probe

    class FM:
        device = torch.device("cuda")
        name = "fm"

        def to(self, d):
            return self

    lens = Lens(FM())
    V = torch.randn(6, 12, 64, generator=torch.Generator().manual_seed(8))
    db = {"a": V, "b": V[:3]}
    cl = lens.eval_clarity(db)
    assert set(cl) == {"a", "b"} and cl["b"].shape == (3,)
    maxnorm_close(cl["a"].numpy(), rp.clarity_score(V).numpy())
    po = lens.eval_polysemanticity(V)
    np.testing.assert_allclose(po.numpy(), P.polysemanticity_gram(V.numpy()), atol=2e-6)
    agg = V.mean(1)
    q = torch.randn(2, 64, generator=torch.Generator().manual_seed(9))
    pr = _probe(q, {"a": agg})
    maxnorm_close(pr["a"].numpy(), rp.similarity_score(q, agg).numpy())
    maxnorm_close(lens.eval_redundancy(agg).numpy(), rp.redundancy_score(agg).numpy())
