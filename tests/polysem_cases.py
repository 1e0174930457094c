"""Seeded inputs for the polysemanticity / clarity parity tests. Only the reference's OUTPUTS for these cases are stored
(tests/golden/scores_poly.npz, written by oracle/make_golden.py); the inputs are regenerated here (numpy Generator,
bit-stable across platforms)."""

import numpy as np

CASES = {
    # name: (C, k, D, kind)
    "gauss_k10": (24, 10, 128, "gauss"),
    "gauss_k20": (32, 20, 512, "gauss"),
    "gauss_k64": (16, 64, 96, "gauss"),
    "gauss_k100_d7": (16, 100, 7, "gauss"),
    "gauss_k256": (8, 256, 512, "gauss"),
    "planted_k64": (24, 64, 32, "planted"),
    "planted_k256": (8, 256, 128, "planted"),
    "unbalanced_k48": (24, 48, 64, "unbalanced"),
    "dups_k16": (20, 16, 32, "dups"),
    "allsame_k12": (4, 12, 32, "allsame"),
    "scaled_k32": (12, 32, 64, "scaled"),
}


def make_case(name: str) -> np.ndarray:
    C, k, D, kind = CASES[name]
    rng = np.random.default_rng(abs(hash_name(name)))
    V = rng.standard_normal((C, k, D)).astype(np.float32)
    if kind == "planted":
        V[:, ::2] += 2.0 * rng.standard_normal((C, 1, D)).astype(np.float32)
    elif kind == "unbalanced":
        V[:, :3] += 4.0 * rng.standard_normal((C, 1, D)).astype(np.float32)
        V[C // 2 :, 1:3] = V[C // 2 :, 3:5]  # second half: a single far outlier -> "< 2 members" fallback
    elif kind == "dups":
        V[:, 1:] = V[:, :1]
        V[5:, 8] += 1.0
        V[10:, 9] += 2.0
    elif kind == "allsame":
        V[:] = V[:, :1]
    elif kind == "scaled":
        V *= (10.0 ** rng.uniform(-3, 3, size=(C, 1, 1))).astype(np.float32)
    return V


def hash_name(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


# Cases for the general kernel (K8g): (C, k, D, kind, n_clusters) — more than two clusters and / or more than 256 examples.
# Reference outputs: tests/golden/scores_poly_general.npz (oracle/make_golden.py::gen_scores_poly_general).
GENERAL_CASES = {
    "g3_gauss_k40": (10, 40, 24, "gauss", 3),
    "g4_planted_k64": (8, 64, 16, "planted3", 4),
    "g3_planted_k300": (6, 300, 32, "planted3", 3),
    "g2_gauss_k300": (6, 300, 48, "gauss", 2),
    "g2_planted_k520": (4, 520, 24, "planted", 2),
    "g5_k30_d5": (10, 30, 5, "gauss", 5),
    "g3_dups_k24": (12, 24, 8, "dups", 3),
    "g8_gauss_k100": (4, 100, 12, "gauss", 8),
    "g2_outlier_k300": (6, 300, 16, "outlier", 2),
}


def make_general_case(name: str) -> np.ndarray:
    C, k, D, kind, _ = GENERAL_CASES[name]
    rng = np.random.default_rng(abs(hash_name(name)))
    V = rng.standard_normal((C, k, D)).astype(np.float32)
    if kind == "planted":
        V[:, ::2] += 2.0 * rng.standard_normal((C, 1, D)).astype(np.float32)
    elif kind == "planted3":
        V[:, ::3] += 3.0 * rng.standard_normal((C, 1, D)).astype(np.float32)
        V[:, 1::3] -= 3.0 * rng.standard_normal((C, 1, D)).astype(np.float32)
    elif kind == "dups":
        V[:, 1:] = V[:, :1]          # two or three distinct points only: clusters stay empty / get relocated
        V[4:, 8] += 1.0
        V[8:, 9] += 2.0
    elif kind == "outlier":
        V[:, 0] += 40.0              # one far example: a single-member cluster -> the "< 2 members" fallback
    return V
