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


# ---- the general form (any n_clusters, any number of examples): what slb_polysem_kmeans runs --------------------
def test_general_oracle_matches_sklearn():
    """labels / centres / inertia of oracle.polysem.kmeans_direct against sklearn's KMeans itself, incl. duplicate points
    (clusters that stay empty are placed at the biggest cluster's row, _average_centers) and more than 256 examples."""
    import torch
    from sklearn.cluster import KMeans

    rng = np.random.default_rng(0)
    for k, D, m in ((40, 16, 3), (300, 24, 2), (64, 8, 4), (30, 5, 5), (257, 12, 3), (50, 4, 6)):
        for trial in range(4):
            X = rng.standard_normal((k, D)).astype(np.float32)
            if trial == 1:
                X[::3] += 3
            if trial == 2:
                X[1:] = X[:1]
                X[5:, 0] += 1
            if trial == 3:
                X[:] = X[:1]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                km = KMeans(n_clusters=m, n_init=10, random_state=123).fit(torch.from_numpy(X))  # as the reference calls it
            labels, centres, inertia = P.kmeans_direct(X, m)
            assert (labels == km.labels_).all(), (k, D, m, trial)
            np.testing.assert_allclose(centres, km.cluster_centers_, rtol=0, atol=1e-9)
            np.testing.assert_allclose(inertia, km.inertia_, rtol=1e-9, atol=1e-12)


def test_general_oracle_matches_reference_fixture(golden):
    from tests.polysem_cases import GENERAL_CASES, make_general_case

    z = np.load(golden / "scores_poly_general.npz")
    for name, spec in GENERAL_CASES.items():
        V = make_general_case(name)
        np.testing.assert_allclose(P.polysemanticity_general(V, n_clusters=spec[4]), z[f"{name}.poly"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(P.polysemanticity_general(V, n_clusters=spec[4], replace_empty_clusters=False),
                                   z[f"{name}.poly_noreplace"], rtol=0, atol=1e-9)


def test_general_draws_equal_oracle_draws():
    from semanticlens_b200 import ops

    for k, m in ((40, 3), (300, 2), (100, 8)):
        f1, r1, L = ops.kmeanspp_draws_general(k, m, 123)
        f2, r2 = P.kmeanspp_draws_general(k, m, 123)
        assert L == P.n_local_trials(m) and (f1 == f2).all() and (r1 == r2).all()
    # two clusters: the general stream is the fast kernel's stream
    f1, r1, _ = ops.kmeanspp_draws_general(64, 2, 7)
    f2, r2 = ops.kmeanspp_draws(64, 7)
    assert (f1 == f2).all() and (r1[:, 0] == r2).all()
