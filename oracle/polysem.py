"""ORACLE (test infrastructure, not product code) — polysemanticity_score restated without sklearn.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

The reference (semanticlens/scores.py:132-185) fits ``sklearn.cluster.KMeans(n_clusters=2, n_init=10,
random_state=123)`` per neuron (third-party arithmetic: scikit-learn, pinned 1.6.1/1.7.1 in the reference's uv.lock,
1.9.0 installed here). This module restates that published algorithm twice:

* :func:`kmeans2_direct` — in sample space, float64, following sklearn/cluster/_kmeans.py (``KMeans.fit`` :1436-1563,
  ``_kmeans_plusplus`` :219-275, ``_kmeans_single_lloyd`` :686-758, ``_tolerance`` :285-293) and
  _k_means_lloyd.pyx (``_update_chunk_dense``: argmin with strict ``<``; ``_relocate_empty_clusters_dense`` in
  _k_means_common.pyx :167-211);
* :func:`kmeans2_gram` — the same algorithm expressed on the neuron's Gram matrix G0 = X X^T only, which is the form
  the K8 kernel runs (semanticlens_b200/csrc/polysem.cu). Every quantity sklearn derives from X (distances to
  centres that are means of subsets, centre shifts, inertia, tolerance, the final cosine of the two un-centred centres)
  is a function of G0.

* :func:`kmeans_direct` / :func:`polysemanticity_general` — the sample-space form for any ``n_clusters`` and any number of
  examples (the general kernel ``slb_polysem_kmeans``), pinned against sklearn for 2-5 clusters and 300+ examples.

Pinned: tests/test_oracle_polysem.py checks both against sklearn itself (labels identical, centres / scores to 1e-9)
and against tests/golden/scores.npz recorded from the imported reference.
"""

from __future__ import annotations

import numpy as np

N_INIT = 10
MAX_ITER = 300
TOL = 1e-4


def kmeanspp_draws(k: int, seed: int = 123, n_init: int = N_INIT):
    """The data-independent part of sklearn's RandomState stream for n_clusters=2: per init the index of the first
    centre (``random_state.choice(k, p=1/k)``, _kmeans.py:231) and the two uniforms of the local trials (:249)."""
    rs = np.random.RandomState(seed)
    first = np.empty(n_init, dtype=np.int64)
    rand = np.empty((n_init, 2), dtype=np.float64)
    p = np.ones(k, dtype=np.float64)
    p = p / p.sum()
    for i in range(n_init):
        first[i] = rs.choice(k, p=p)
        rand[i] = rs.uniform(size=2)
    return first, rand


def _same_clustering(a, b):
    mapping = {}
    for x, y in zip(a, b):
        if x not in mapping:
            mapping[x] = y
        elif mapping[x] != y:
            return False
    return True


# ------------------------------------------------------------------------------------------------
# direct restatement (sample space)
# ------------------------------------------------------------------------------------------------
def kmeans2_direct(X, seed: int = 123):
    """-> (labels (k,), centres (2, D) un-centred, inertia). X (k, D); float64 like sklearn given a torch tensor."""
    X = np.array(X, dtype=np.float64)
    k, D = X.shape
    tol = np.mean(np.var(X, axis=0)) * TOL
    mean = X.mean(axis=0)
    X = X - mean
    xsq = (X * X).sum(1)
    first, rand = kmeanspp_draws(k, seed)

    def d2(c):
        return np.maximum(xsq - 2.0 * (X @ c) + c @ c, 0.0)

    best = None
    for it in range(N_INIT):
        # k-means++ (n_local_trials = 2 + int(log 2) = 2)
        c0 = X[first[it]]
        closest = d2(c0)
        pot = closest.sum()
        cand = np.searchsorted(np.cumsum(closest), rand[it] * pot)
        cand = np.minimum(cand, k - 1)
        dist_c = np.stack([np.minimum(closest, d2(X[c])) for c in cand])
        pots = dist_c.sum(1)
        centers = np.stack([c0, X[cand[int(np.argmin(pots))]]])
        # Lloyd
        labels_old = np.full(k, -1)
        strict = False
        for _ in range(MAX_ITER):
            pd = (centers * centers).sum(1)[None, :] - 2.0 * (X @ centers.T)
            labels = (pd[:, 1] < pd[:, 0]).astype(np.int64)
            sums = np.stack([X[labels == a].sum(0) for a in (0, 1)])
            w = np.array([(labels == a).sum() for a in (0, 1)], dtype=np.float64)
            if (w == 0).any():
                dist = ((X - centers[labels]) ** 2).sum(1)
                if dist.max() > 0:
                    far = int(np.argmax(dist))
                    e = int(np.where(w == 0)[0][0])
                    o = labels[far]
                    sums[o] -= X[far]
                    sums[e] = X[far]
                    w[e] = 1
                    w[o] -= 1
            new = centers.copy()
            for a in (0, 1):
                if w[a] > 0:
                    new[a] = sums[a] / w[a]
                else:
                    new[a] = sums[a]
            shift = ((new - centers) ** 2).sum()
            centers = new
            if np.array_equal(labels, labels_old):
                strict = True
                break
            if shift <= tol:
                break
            labels_old = labels
        if not strict:
            pd = (centers * centers).sum(1)[None, :] - 2.0 * (X @ centers.T)
            labels = (pd[:, 1] < pd[:, 0]).astype(np.int64)
        inertia = ((X - centers[labels]) ** 2).sum()
        if best is None or (inertia < best[2] and not _same_clustering(labels, best[0])):
            best = (labels, centers, inertia)
    return best[0], best[1] + mean, best[2]


# ------------------------------------------------------------------------------------------------
# Gram-matrix form (what the kernel runs)
# ------------------------------------------------------------------------------------------------
def kmeans2_gram(G0, D: int, first, rand):
    """2-means on the un-centred Gram matrix G0 (k, k) float64 -> (labels, member mask of the final M-step, inertia).

    Centres are never formed: a centre is the mean of the points in a member set A, so with the centred Gram
    Gc = G0 - r 1^T - 1 r^T + m   (r = row means, m = grand mean)
        x_i . c_A = (1/|A|) sum_{j in A} Gc_ij        |c_A|^2 = (1/|A|) sum_{i in A} x_i . c_A
    and, the rows of Gc summing to zero, sum_{j in B} Gc_ij = -sum_{j in A} Gc_ij for the complement B.
    """
    k = G0.shape[0]
    r = G0.mean(1)
    m = r.mean()
    Gc = G0 - r[:, None] - r[None, :] + m
    diag = np.diag(Gc).copy()
    tol = diag.sum() / (k * D) * TOL

    def dist_to_point(j):
        return np.maximum(diag - 2.0 * Gc[:, j] + diag[j], 0.0)

    best = None
    for it in range(len(first)):
        i0 = int(first[it])
        closest = dist_to_point(i0)
        pot = closest.sum()
        cum = np.cumsum(closest)
        cand = [min(int((cum < v).sum()), k - 1) for v in rand[it] * pot]
        pots = [np.minimum(closest, dist_to_point(c)).sum() for c in cand]
        i1 = cand[int(np.argmin(pots))]
        # centre a is described by s[a][i] = x_i . c_a and n[a] = |c_a|^2
        s = np.stack([Gc[:, i0], Gc[:, i1]])
        n = np.array([diag[i0], diag[i1]])
        mask = None
        labels_old = np.full(k, -1)
        strict = False
        for _ in range(MAX_ITER):
            labels = ((n[1] - 2.0 * s[1]) < (n[0] - 2.0 * s[0])).astype(np.int64)
            mask = labels.copy()
            cnt = np.array([(mask == 0).sum(), (mask == 1).sum()])
            if (cnt == 0).any():
                o = 0 if cnt[1] == 0 else 1  # every point carries label o
                dist = diag - 2.0 * s[o] + n[o]
                if dist.max() > 0:
                    far = int(np.argmax(dist))
                    mask[far] = 1 - o
                    cnt = np.array([(mask == 0).sum(), (mask == 1).sum()])
            if (cnt == 0).any():
                # all points identical to the centre: sklearn leaves the empty centre at the zero vector
                e = 0 if cnt[0] == 0 else 1
                t = Gc @ (mask == 1 - e).astype(np.float64)
                s_new = np.zeros_like(s)
                n_new = np.zeros(2)
                s_new[1 - e] = t / cnt[1 - e]
                n_new[1 - e] = s_new[1 - e][mask == 1 - e].sum() / cnt[1 - e]
                cross = np.array([0.0, 0.0])
                cross[1 - e] = s[1 - e][mask == 1 - e].sum() / cnt[1 - e]
            else:
                t = Gc @ (mask == 0).astype(np.float64)  # sum over members of cluster 0; cluster 1 gets -t
                s_new = np.stack([t / cnt[0], -t / cnt[1]])
                n_new = np.array([s_new[0][mask == 0].sum() / cnt[0], s_new[1][mask == 1].sum() / cnt[1]])
                cross = np.array([s[0][mask == 0].sum() / cnt[0], s[1][mask == 1].sum() / cnt[1]])  # c_new . c_old
            shift = (n_new - 2.0 * cross + n).sum()
            s, n = s_new, n_new
            if np.array_equal(labels, labels_old):
                strict = True
                break
            if shift <= tol:
                break
            labels_old = labels
        if not strict:
            labels = ((n[1] - 2.0 * s[1]) < (n[0] - 2.0 * s[0])).astype(np.int64)
        inertia = (diag - 2.0 * np.where(labels == 1, s[1], s[0]) + np.where(labels == 1, n[1], n[0])).sum()
        if best is None or (inertia < best[2] and not _same_clustering(labels, best[0])):
            best = (labels, mask, inertia)
    return best


def poly_from_gram(G0, D: int, first, rand, replace_empty_clusters: bool = True, k_fallback: int = 10):
    """polysemanticity of one neuron from its Gram matrix (scores.py:167-184)."""
    k = G0.shape[0]
    labels, mask, _ = kmeans2_gram(G0, D, first, rand)
    cnt_l = np.bincount(labels, minlength=2)
    if replace_empty_clusters and cnt_l.min() < 2:
        # 1 - mean_{i<10} clarity([mean(V), V[:, i]]); clarity of two vectors is their cosine
        r = G0.mean(1)
        m = r.mean()
        ns = min(k_fallback, k)
        cos = r[:ns] / (np.maximum(np.sqrt(m), 1e-12) * np.maximum(np.sqrt(np.diag(G0)[:ns]), 1e-12))
        return 1.0 - cos.sum() / ns
    a, b = mask == 0, mask == 1
    if not a.any() or not b.any():
        # degenerate (all points equal to one centre): sklearn leaves the empty centre at the zero vector of the
        # centred space, i.e. at X_mean after `best_centers += X_mean` (_kmeans.py:1546)
        o = a if a.any() else b
        r = G0.mean(1)
        soo = G0[np.ix_(o, o)].sum() / o.sum() ** 2
        som = r[o].sum() / o.sum()
        return 1.0 - som / (max(np.sqrt(soo), 1e-12) * max(np.sqrt(r.mean()), 1e-12))
    saa = G0[np.ix_(a, a)].sum() / a.sum() ** 2
    sbb = G0[np.ix_(b, b)].sum() / b.sum() ** 2
    sab = G0[np.ix_(a, b)].sum() / (a.sum() * b.sum())
    return 1.0 - sab / (max(np.sqrt(saa), 1e-12) * max(np.sqrt(sbb), 1e-12))


def polysemanticity_gram(V, seed: int = 123, replace_empty_clusters: bool = True):
    """(C, k, D) -> (C,) float64 through the Gram form."""
    V = np.asarray(V, dtype=np.float64)
    C, k, D = V.shape
    first, rand = kmeanspp_draws(k, seed)
    return np.array([poly_from_gram(v @ v.T, D, first, rand, replace_empty_clusters) for v in V])


def polysemanticity_direct(V, seed: int = 123, replace_empty_clusters: bool = True):
    """(C, k, D) -> (C,) float64 through the sample-space restatement."""
    V32 = np.asarray(V, dtype=np.float32)
    out = np.empty(len(V32))
    for c, v in enumerate(V32):
        labels, centers, _ = kmeans2_direct(v, seed)
        cn = centers / np.maximum(np.linalg.norm(centers, axis=1, keepdims=True), 1e-12)
        out[c] = 1.0 - (((cn.mean(0) ** 2).sum() - 0.5) / 1.0 * 2.0)
        cnt = np.bincount(labels, minlength=2)
        if replace_empty_clusters and cnt.min() < 2:
            ns = min(10, v.shape[0])
            mean = v.mean(0)
            acc = np.float32(0)
            for i in range(ns):
                pair = np.stack([mean, v[i]]).astype(np.float32)
                pn = pair / np.maximum(np.linalg.norm(pair, axis=1, keepdims=True), np.float32(1e-12))
                acc = acc + np.float32(((pn.mean(0) ** 2).sum() - np.float32(0.5)) * 2)
            out[c] = 1.0 - float(acc) / ns
    return out

# ------------------------------------------------------------------------------------------------
# general restatement: any number of clusters, any number of examples (what slb_polysem_kmeans runs)
# ------------------------------------------------------------------------------------------------
def n_local_trials(n_clusters: int) -> int:
    """sklearn _kmeans.py:226  (2 + int(log(n_clusters)))"""
    return 2 + int(np.log(n_clusters))


def kmeanspp_draws_general(k: int, n_clusters: int, seed: int = 123, n_init: int = N_INIT):
    """The data-independent part of sklearn's RandomState stream for any n_clusters: per init the first centre
    (``random_state.choice(k, p=1/k)``) and, for each further centre, ``n_local_trials`` uniforms (``_kmeans_plusplus``).
    -> first (n_init,) int64, rand (n_init, n_clusters - 1, n_local_trials) float64."""
    rs = np.random.RandomState(seed)
    L = n_local_trials(n_clusters)
    first = np.empty(n_init, dtype=np.int64)
    rand = np.empty((n_init, max(n_clusters - 1, 0), L), dtype=np.float64)
    p = np.ones(k, dtype=np.float64)
    p = p / p.sum()
    for i in range(n_init):
        first[i] = rs.choice(k, p=p)
        for c in range(1, n_clusters):
            rand[i, c - 1] = rs.uniform(size=L)
    return first, rand


def kmeans_direct(X, n_clusters: int, seed: int = 123):
    """sklearn KMeans(n_clusters, n_init=10, random_state=seed).fit(X) in float64, sample space.
    -> (labels (k,), centres (n_clusters, D) un-centred, inertia). Several empty clusters in one iteration take the
    farthest points in descending order of distance (sklearn: np.argpartition's order, unspecified among them)."""
    X = np.array(X, dtype=np.float64)
    k, D = X.shape
    m = n_clusters
    if k < m:
        raise ValueError(f"n_samples={k} should be >= n_clusters={m}.")
    tol = np.mean(np.var(X, axis=0)) * TOL
    mean = X.mean(axis=0)
    X = X - mean
    xsq = (X * X).sum(1)
    first, rand = kmeanspp_draws_general(k, m, seed)

    def d2(c):
        return np.maximum(xsq - 2.0 * (X @ c) + c @ c, 0.0)

    def e_step(centers):
        pd = (centers * centers).sum(1)[None, :] - 2.0 * (X @ centers.T)
        return np.argmin(pd, axis=1)  # first minimum, like the strict `<` scan of _update_chunk_dense

    best = None
    for it in range(N_INIT):
        idx = [int(first[it])]
        closest = d2(X[idx[0]])
        pot = closest.sum()
        for c in range(1, m):
            cand = np.minimum(np.searchsorted(np.cumsum(closest), rand[it, c - 1] * pot), k - 1)
            dist_c = np.stack([np.minimum(closest, d2(X[j])) for j in cand])
            pots = dist_c.sum(1)
            b = int(np.argmin(pots))
            pot, closest = pots[b], dist_c[b]
            idx.append(int(cand[b]))
        centers = X[idx].copy()
        labels_old = np.full(k, -1)
        strict = False
        for _ in range(MAX_ITER):
            labels = e_step(centers)
            sums = np.stack([X[labels == a].sum(0) for a in range(m)])
            w = np.array([(labels == a).sum() for a in range(m)], dtype=np.float64)
            empty = np.where(w == 0)[0]
            if len(empty):
                dist = ((X - centers[labels]) ** 2).sum(1)
                if dist.max() > 0:
                    order = np.lexsort((np.arange(k), -dist))[: len(empty)]  # farthest first, lower index on ties
                    for e, far in zip(empty, order):
                        o = labels[far]
                        sums[o] -= X[far]
                        sums[e] = X[far]
                        w[e] = 1
                        w[o] -= 1
            # _average_centers (_k_means_common.pyx, sklearn 1.9): a cluster that is still empty goes to "the location of the
            # biggest cluster" — read in index order, i.e. the biggest cluster's raw SUM if it has not been averaged yet
            new = sums.copy()
            amax = int(np.argmax(w))
            for j in range(m):
                if w[j] > 0:
                    new[j] = new[j] / w[j]
                else:
                    new[j] = new[amax]
            shift = ((new - centers) ** 2).sum()
            centers = new
            if np.array_equal(labels, labels_old):
                strict = True
                break
            if shift <= tol:
                break
            labels_old = labels
        if not strict:
            labels = e_step(centers)
        inertia = ((X - centers[labels]) ** 2).sum()
        if best is None or (inertia < best[2] and not _same_clustering(labels, best[0])):
            best = (labels, centers, inertia)
    return best[0], best[1] + mean, best[2]


def polysemanticity_general(V, n_clusters: int = 2, seed: int = 123, replace_empty_clusters: bool = True):
    """scores.py:132-185 for any n_clusters: 1 - clarity of the cluster centres (float64), with the reference's fallback
    for neurons whose smallest cluster has fewer than 2 members (or that miss a cluster altogether)."""
    V32 = np.asarray(V, dtype=np.float32)
    m = n_clusters
    out = np.empty(len(V32))
    for c, v in enumerate(V32):
        labels, centers, _ = kmeans_direct(v, m, seed)
        cn = centers / np.maximum(np.linalg.norm(centers, axis=1, keepdims=True), 1e-12)
        out[c] = 1.0 - (((cn.mean(0) ** 2).sum() - 1.0 / m) / (m - 1) * m)
        cnt = np.bincount(labels, minlength=m)
        if replace_empty_clusters and cnt.min() < 2:
            ns = min(10, v.shape[0])
            mean = v.mean(0)
            acc = np.float32(0)
            for i in range(ns):
                pair = np.stack([mean, v[i]]).astype(np.float32)
                pn = pair / np.maximum(np.linalg.norm(pair, axis=1, keepdims=True), np.float32(1e-12))
                acc = acc + np.float32(((pn.mean(0) ** 2).sum() - np.float32(0.5)) * 2)
            out[c] = 1.0 - float(acc) / ns
    return out
