"""oracle/polysem.py (the sklearn-free restatement of polysemanticity_score, in sample space and in the Gram form the
K8 kernel runs) against the fixtures recorded from the imported reference and against sklearn itself."""

import warnings

import numpy as np
import pytest

from oracle import polysem as P
from tests.polysem_cases import CASES, make_case


@pytest.mark.parametrize("name", sorted(CASES))
def test_gram_form_matches_reference_fixture(golden, name):
    z = np.load(golden / "scores_poly.npz")
    V = make_case(name)
    got = P.polysemanticity_gram(V)
    # the reference evaluates the "< 2 members" fallback in fp32: 1e-6 there, 1e-9 on the k-means path
    np.testing.assert_allclose(got, z[f"{name}.poly"], rtol=0, atol=2e-6)
    got_nr = P.polysemanticity_gram(V, replace_empty_clusters=False)
    np.testing.assert_allclose(got_nr, z[f"{name}.poly_noreplace"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", ["gauss_k10", "planted_k64", "unbalanced_k48", "dups_k16"])
def test_direct_form_matches_reference_fixture(golden, name):
    z = np.load(golden / "scores_poly.npz")
    got = P.polysemanticity_direct(make_case(name))
    np.testing.assert_allclose(got, z[f"{name}.poly"], rtol=0, atol=2e-6)


def test_small_golden(golden):
    z = np.load(golden / "scores.npz")
    np.testing.assert_allclose(P.polysemanticity_gram(z["P"]), z["poly"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(P.polysemanticity_gram(z["P"], replace_empty_clusters=False), z["poly_noreplace"], rtol=0,
                               atol=1e-9)


def test_labels_match_sklearn():
    from sklearn.cluster import KMeans

    rng = np.random.default_rng(5)
    for k, D in ((17, 9), (64, 32), (128, 48)):
        X = rng.standard_normal((k, D)).astype(np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            km = KMeans(n_clusters=2, n_init=10, random_state=123).fit(X.astype(np.float64))
        labels, centres, inertia = P.kmeans2_direct(X)
        assert (labels == km.labels_).all()
        np.testing.assert_allclose(centres, km.cluster_centers_, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(inertia, km.inertia_, rtol=1e-10)
        first, rand = P.kmeanspp_draws(k)
        G0 = X.astype(np.float64) @ X.astype(np.float64).T
        gl, gmask, gin = P.kmeans2_gram(G0, D, first, rand)
        assert (gl == km.labels_).all()
        np.testing.assert_allclose(gin, km.inertia_, rtol=1e-9)


def test_product_draws_equal_oracle_draws():
    from semanticlens_b200 import ops

    for k in (2, 10, 20, 256):
        a, b = P.kmeanspp_draws(k)
        c, d = ops.kmeanspp_draws(k, 123)
        assert (a == c).all() and (b == d).all()
