"""Synthetic Criteo-libsvm-shaped batches.

Shape follows the reference's CTR.parseFeature (CTR.java:47-68): F=23 categorical ids
("E", F x N), Xn=45 numeric values in [0,1) rounded to 2 decimals ("X", Xn x N), the wide
input "W" = E mod 100000 (CTR.java:36,65; MatrixUtil.java:27-33) and a Bernoulli(0.356)
label ("Y", 1 x N).  All matrices are column-major like jblas' FloatMatrix, i.e. numpy
arrays of shape (N, rows) in C order; ids are int64 (the reference carries them as floats,
exact only below 2^24 — SURVEY quirk 2).
"""
import numpy as np

# per-field unique counts of the reference's bundled sample (src/main/resources/train.txt)
FIELD_UNIQUES = [1, 21, 3522, 209, 58, 600, 1056, 2735, 485, 41, 104, 2, 35, 366, 2749, 11, 54, 2, 701, 89, 13, 3, 5]
WIDE_SIZE = 100000  # CTR.java:36

CONFIGS = {
    # name: (model kind, B, F, Xn, D, vocab, fc dims, embedding optimizer)
    "cfg2": dict(kind="widedeep", B=4096, F=23, Xn=45, D=16, V=1_000_000, fc=[256, 256, 256, 1], emb_opt="adam"),
    "cfg3": dict(kind="widedeep", B=8192, F=23, Xn=45, D=32, V=100_000_000, fc=[256, 256, 256, 1], emb_opt="ftrl"),
    "cfg4": dict(kind="dnn", B=16384, F=23, Xn=45, D=64, V=10_000_000, fc=[256, 256, 256, 1], emb_opt="adam"),
    "cfg5": dict(kind="fcnn", B=1024, F=0, Xn=784, D=0, V=0, fc=[150, 50, 10], emb_opt="adam"),
    "ctr": dict(kind="dnn", B=1000, F=23, Xn=45, D=10, V=12862, fc=[150, 10, 1], emb_opt="adam"),
}


def field_vocab(V, F=23, uniform=False):
    """Disjoint per-field id ranges [base_j, base_j + V_j) with sum V_j ~= V."""
    if uniform:
        sizes = np.full(F, max(1, V // F), np.int64)
    else:
        u = np.asarray((FIELD_UNIQUES * ((F + 22) // 23))[:F], np.float64)
        sizes = np.maximum(1, np.floor(u / u.sum() * V)).astype(np.int64)
    base = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    return base, sizes


class Synth:
    def __init__(self, F=23, Xn=45, V=1_000_000, dist="zipf", seed=20261017, zipf_s=1.05, uniform_fields=False, n_classes=0):
        self.F, self.Xn, self.V, self.dist, self.zipf_s = F, Xn, V, dist, zipf_s
        self.rng = np.random.Generator(np.random.Philox(seed))
        self.n_classes = n_classes
        if F:
            self.base, self.sizes = field_vocab(V, F, uniform_fields)

    def _ids(self, N):
        F = self.F
        E = np.empty((N, F), np.int64)
        for j in range(F):
            vj = int(self.sizes[j])
            if self.dist == "uniform" or vj == 1:
                r = self.rng.integers(0, vj, N)
            else:
                # bounded zipf by inverse-CDF on a continuous approximation: rank ~ u^(-1/(s-1)) clipped
                u = self.rng.random(N)
                s = self.zipf_s
                r = np.floor((1.0 + u * (float(vj + 1) ** (1.0 - s) - 1.0)) ** (1.0 / (1.0 - s))).astype(np.int64) - 1
                r = np.clip(r, 0, vj - 1)
            E[:, j] = self.base[j] + r
        return E

    def batch(self, N):
        """Returns dict E (N,F) int64, X (N,Xn) f32, W (N,F) int64, Y (N,) f32 — column-major F x N etc."""
        out = {}
        if self.F:
            E = self._ids(N)
            out["E"] = E
            out["W"] = E % WIDE_SIZE
        if self.n_classes:
            out["X"] = np.floor(self.rng.random((N, self.Xn)) * 256).astype(np.float32)
            out["Y"] = self.rng.integers(0, self.n_classes, N).astype(np.float32)
        else:
            out["X"] = (np.floor(self.rng.random((N, self.Xn)) * 100) / 100).astype(np.float32)
            out["Y"] = (self.rng.random(N) < 0.356).astype(np.float32)
        return out
