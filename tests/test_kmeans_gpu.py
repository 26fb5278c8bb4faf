"""GPU parity: MinMaxScaler, k-means assignment and mini-batch update vs the real scikit-learn
(the library the reference calls at datasetbuilder.py:369-376)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _t(engine, a, dtype=None):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), device=engine.device, dtype=dtype)


def test_minmax_scaler_bit_exact(engine):
    rng = np.random.default_rng(1)
    X = rng.normal(3.0, 40.0, size=(7013, 45))
    X[:, 3] = 7.25            # constant column -> scale 1
    X[:, 9] = 1.0 + rng.random(7013) * 1e-16   # range below 10 eps
    X[:, 11] *= 1e-9
    Xs, dmin, dmax, scale, min_ = O.minmax_scale(X)
    d = _t(engine, X)
    mn, mx = engine.minmax_fit(d)
    assert np.array_equal(mn.cpu().numpy(), dmin) and np.array_equal(mx.cpu().numpy(), dmax)
    engine.minmax_transform(d, mn.cpu().numpy(), mx.cpu().numpy())
    assert np.array_equal(d.cpu().numpy(), Xs), "scaled matrix differs from sklearn MinMaxScaler"


@pytest.mark.parametrize("rows,D,K", [(5000, 37, 1000), (1024, 20, 10000), (777, 121, 130), (300, 5, 3)])
def test_assign_labels_match_sklearn(engine, rows, D, K):
    rng = np.random.default_rng(rows + D)
    X = rng.random((rows, D))
    centers = X[rng.choice(rows, K, replace=K > rows)] + rng.normal(0, 0.01, (K, D))
    if K >= 20:
        centers[17] = centers[5]      # exact duplicate: the lowest index must win
        X[3] = centers[5]
    labels, inertia = O.kmeans_labels(X, centers)
    got, gin, mind = engine.kmeans_assign(_t(engine, X), _t(engine, centers), want_inertia=True, want_mindist=True)
    got = got.cpu().numpy()
    assert np.array_equal(got, labels), f"{(got != labels).sum()} labels differ from sklearn"
    assert abs(gin - inertia) <= 1e-12 * max(inertia, 1e-300)
    d2 = ((X - centers[labels]) ** 2).sum(axis=1)
    assert np.allclose(mind.cpu().numpy(), d2, rtol=1e-12, atol=0)


def test_assign_rowidx_and_weights(engine):
    rng = np.random.default_rng(5)
    X = rng.random((4000, 33))
    centers = rng.random((500, 33))
    idx = rng.integers(0, 4000, 1024).astype(np.int32)
    w = rng.random(1024)
    labels, inertia = O.kmeans_labels(X[idx], centers)
    from sklearn.cluster._kmeans import _labels_inertia

    _, inertia_w = _labels_inertia(np.ascontiguousarray(X[idx]), w, centers, n_threads=1)
    got, gin, _ = engine.kmeans_assign(_t(engine, X), _t(engine, centers), rowidx=_t(engine, idx), weights=_t(engine, w),
                                       want_inertia=True)
    assert np.array_equal(got.cpu().numpy(), labels)
    assert abs(gin - inertia_w) <= 1e-12 * inertia_w


@pytest.mark.parametrize("B,D,K", [(1024, 40, 10000), (1024, 17, 50), (6000, 23, 300)])
def test_minibatch_update_matches_sklearn(engine, B, D, K):
    rng = np.random.default_rng(B + K)
    X = rng.random((B, D))
    centers = rng.random((K, D))
    wsums = rng.integers(0, 50, K).astype(np.float64)
    sw = np.ones(B) if B != 6000 else rng.random(B) + 0.5
    labels, _ = O.kmeans_labels(X, centers)
    ref_c, ref_w = O.minibatch_update(X, sw, centers, wsums, labels)
    dX, dC, dW = _t(engine, X), _t(engine, centers), _t(engine, wsums)
    sums, wsum = engine.kmeans_partial_sums(dX, _t(engine, labels.astype(np.int32)), K, weights=_t(engine, sw))
    engine.kmeans_apply_update(dC, dW, sums, wsum)
    # tolerance on centroid coordinates stated for this path: 1e-12 relative
    assert np.allclose(dC.cpu().numpy(), ref_c, rtol=1e-12, atol=1e-15)
    assert np.allclose(dW.cpu().numpy(), ref_w, rtol=1e-14, atol=0)


@pytest.mark.parametrize("rows,D,K", [(5000, 37, 1000), (4096, 35, 10000), (777, 110, 130), (900, 100, 260), (300, 5, 3),
                                      (20000, 64, 2500)])
def test_tensor_core_assign_matches_sklearn(engine, rows, D, K):
    """tcgen05 screening pass + float64 re-check of near-ties: identical labels to sklearn."""
    rng = np.random.default_rng(rows * 7 + D)
    X = rng.random((rows, D))
    centers = X[rng.choice(rows, K, replace=K > rows)] + rng.normal(0, 0.01, (K, D))
    if K >= 20:
        centers[17] = centers[5]
        X[3] = centers[5]
    labels, _ = O.kmeans_labels(X, centers)
    got, nre = engine.kmeans_assign_tc(_t(engine, X), _t(engine, centers))
    got = got.cpu().numpy()
    assert np.array_equal(got, labels), f"{(got != labels).sum()} labels differ from sklearn ({nre} re-checked)"
    assert nre < 0.25 * rows + 64, f"too many rows fell back to float64: {nre} of {rows}"
    idx = rng.integers(0, rows, 1500).astype(np.int32)
    got2, _ = engine.kmeans_assign_tc(_t(engine, X), _t(engine, centers), rowidx=_t(engine, idx))
    assert np.array_equal(got2.cpu().numpy(), labels[idx])


def test_tensor_core_assign_clustered_data(engine):
    """Descriptor-like data: tight clusters, many near-duplicate rows, unscaled magnitudes."""
    rng = np.random.default_rng(11)
    proto = rng.normal(0, 30.0, (400, 40))
    X = np.concatenate([p + rng.normal(0, 1e-3, (60, 40)) for p in proto])
    centers = np.concatenate([proto + rng.normal(0, 1e-4, proto.shape), proto[:100] + rng.normal(0, 1e-4, (100, 40))])
    labels, _ = O.kmeans_labels(X, centers)
    got, nre = engine.kmeans_assign_tc(_t(engine, X), _t(engine, centers))
    assert np.array_equal(got.cpu().numpy(), labels), f"{(got.cpu().numpy() != labels).sum()} differ; {nre} re-checked"


@pytest.mark.parametrize("n,D,K,seed", [(500, 7, 20, 0), (3000, 35, 150, 1), (1200, 3, 1, 2), (64, 5, 64, 3)])
def test_device_kmeans_plusplus_matches_sklearn(engine, n, D, K, seed):
    """The device-resident seeding loop picks the centres scikit-learn's _kmeans_plusplus picks from the
    same RandomState stream (cluster/_kmeans.py:180-279)."""
    import torch
    from sklearn.cluster._kmeans import _kmeans_plusplus
    from sklearn.utils.extmath import row_norms

    rng = np.random.default_rng(100 + seed)
    X = rng.random((n, D))
    rs = np.random.RandomState(seed)
    _, want = _kmeans_plusplus(X, K, row_norms(X, squared=True), np.ones(n), rs, None)
    rs = np.random.RandomState(seed)
    T = 2 + int(np.log(K))
    first = int(rs.choice(n, p=np.full(n, 1.0 / n)))
    rand = rs.uniform(size=(max(K - 1, 0), T))
    got = engine.kmeans_plusplus(_t(engine, X), K, rand if K > 1 else np.zeros((1, T)), first).cpu().numpy()
    assert np.array_equal(got, want)
    # rows addressed through an index list give the same answer in positions of that list
    perm = rng.permutation(n).astype(np.int32)
    Xp = np.empty_like(X)
    Xp[perm] = X
    got2 = engine.kmeans_plusplus(_t(engine, Xp), K, rand if K > 1 else np.zeros((1, T)), first,
                                  rowidx=torch.as_tensor(perm, device=engine.device)).cpu().numpy()
    assert np.array_equal(got2, want)
    # a batch of independent seedings (MiniBatchKMeans' n_init) in one device loop
    wants, firsts, rands = [], [], []
    for b in range(3):
        rs = np.random.RandomState(seed + 10 * b)
        wants.append(_kmeans_plusplus(X, K, row_norms(X, squared=True), np.ones(n), rs, None)[1])
        rs = np.random.RandomState(seed + 10 * b)
        firsts.append(int(rs.choice(n, p=np.full(n, 1.0 / n))))
        rands.append(rs.uniform(size=(max(K - 1, 0), T)) if K > 1 else np.zeros((1, T)))
    got3 = engine.kmeans_plusplus(_t(engine, X), K, np.stack(rands), np.array(firsts)).cpu().numpy()
    assert np.array_equal(got3, np.stack(wants))


@pytest.mark.parametrize("n,D,K,seed", [(6000, 12, 40, 0), (3000, 30, 25, 7), (5000, 8, 100, 3)])
def test_seeded_fit_matches_sklearn(engine, n, D, K, seed):
    """`MiniBatchKMeansB200(seed=s)` draws from one RandomState in the order
    scikit-learn's MiniBatchKMeans(random_state=s) does (validation rows, n_init seedings, mini-batches,
    reassignments), so the whole fit -- number of steps, centres, labels -- reproduces sklearn's.  Centres to
    1e-9 relative, labels identical.  (What is not reproducible in principle: sklearn sums the k-means++ potentials with
    a BLAS matrix-vector product, whose summation order is the BLAS build's; a candidate whose potential ties with
    another's to the last bits can then be picked differently.  Degenerate cases such as K = n / 3 hit that.)"""
    from sklearn.cluster import MiniBatchKMeans

    from mddatasetbuilder_b200.kmeans import MiniBatchKMeansB200

    rng = np.random.default_rng(100 + seed)
    blobs = rng.random((K, D))
    X = blobs[rng.integers(0, K, n)] + rng.normal(0, 0.03, (n, D))
    ref = MiniBatchKMeans(n_clusters=K, init_size=min(3 * K, n), n_init=3, random_state=seed).fit(X)
    km = MiniBatchKMeansB200(engine, n_clusters=K, init_size=min(3 * K, n), n_init=3, seed=seed).fit(_t(engine, X))
    assert km.n_steps_ == ref.n_steps_, "the fits stopped at different steps"
    C = km.cluster_centers_.cpu().numpy()
    assert np.allclose(C, ref.cluster_centers_, rtol=1e-9, atol=1e-12), "centres differ from scikit-learn's"
    assert np.array_equal(km.labels_.cpu().numpy(), ref.labels_)
