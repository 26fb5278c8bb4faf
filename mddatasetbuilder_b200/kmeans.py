"""MinMaxScaler + MiniBatchKMeans + per-cluster selection on the GPU.

Drop-in for the arithmetic of ``DatasetBuilder._clusterdatas`` (reference
datasetbuilder.py:351-384), which calls scikit-learn's ``MinMaxScaler`` and
``MiniBatchKMeans(n_clusters, init_size=min(3K, rows), n_init=3).fit_predict``.  The control
flow below follows scikit-learn 1.9 step by step (``cluster/_kmeans.py``: ``_kmeans_plusplus``
:180-279, ``_mini_batch_step`` :1601-1697, ``MiniBatchKMeans.fit`` :2060-2226 and its helpers
``_random_reassign`` / ``_mini_batch_convergence``); the heavy steps run in the kernels of
``csrc/kmeans.cu``.  All random draws come from one ``numpy.random.RandomState(seed)`` in the
same order scikit-learn draws them.  The reference seeds nothing (its results differ from run
to run); this implementation is reproducible for a given ``seed``.

Multi-GPU: the training rows (the streaming selector's seeded pool) are replicated on every rank and so
is the RandomState, hence every rank draws the same mini-batches.  ``exchange="replicated"`` runs the
whole step on every rank (no collective, bit-identical for any number of ranks); ``exchange="allreduce"``
gives rank r the r-th slice of each mini-batch and combines centroid sums, weights and the batch inertia
with ONE packed all-reduce per step (NCCL), which changes the summation order only.
"""

from __future__ import annotations

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


class MinMaxScalerB200:
    """``sklearn.preprocessing.MinMaxScaler`` (feature_range (0, 1)) on device rows."""

    def __init__(self, engine):
        self.eng = engine

    def fit_transform(self, X):
        import torch

        mn, mx = self.eng.minmax_fit(X)
        self.data_min_ = mn.cpu().numpy()
        self.data_max_ = mx.cpu().numpy()
        self.eng.minmax_transform(X, self.data_min_, self.data_max_)
        torch.cuda.current_stream(self.eng.device).synchronize()
        return X


class MiniBatchKMeansB200:
    def __init__(self, engine, n_clusters=8, init_size=None, n_init=3, batch_size=1024, max_iter=100,
                 max_no_improvement=10, reassignment_ratio=0.01, seed=0, verbose=False, exchange="replicated"):
        self.eng = engine
        self.n_clusters = int(n_clusters)
        self.init_size = init_size
        self.n_init = n_init
        self.batch_size = batch_size
        self.max_iter = max_iter
        self.max_no_improvement = max_no_improvement
        self.reassignment_ratio = reassignment_ratio
        self.seed = seed
        self.verbose = verbose
        self.exchange = exchange    # "replicated": every rank runs the whole step; "allreduce": batch slices + one all-reduce

    # ------------------------------------------------------------------ seeding
    def _init_centroids_all(self, X, rs, init_size):
        """The n_init seedings of sklearn's MiniBatchKMeans.fit (init_indices, then _kmeans_plusplus on
        those rows), run as ONE batched device loop.  Each seeding consumes a fixed number of variates
        (randint for the rows, one choice, (K-1) x n_local_trials uniforms) and nothing else touches the
        RandomState in between, so drawing them up front reproduces sklearn's stream."""
        import torch

        eng = self.eng
        n = int(X.shape[0])
        K = self.n_clusters
        T = 2 + int(np.log(K))
        rows, firsts, rands = [], [], []
        for _ in range(self.n_init):
            if init_size is not None and init_size < n:
                init_indices = rs.randint(0, n, init_size)
            else:
                init_indices = np.arange(n)
            m = len(init_indices)
            rows.append(init_indices.astype(np.int32))
            firsts.append(int(rs.choice(m, p=np.full(m, 1.0 / m))))
            rands.append(rs.uniform(size=(max(K - 1, 0), T)) if K > 1 else np.zeros((0, T)))
        rowidx = torch.as_tensor(np.stack(rows), dtype=torch.int32, device=eng.device)
        rand = np.stack(rands) if K > 1 else np.zeros((self.n_init, 1, T))
        idx = eng.kmeans_plusplus(X, K, rand, np.array(firsts, dtype=np.int32), rowidx=rowidx)
        return [X[rowidx[b].long()[idx[b].long()]].clone() for b in range(self.n_init)]

    # ------------------------------------------------------------------ one mini-batch step
    def _mini_batch_step(self, X, batch_idx, centers, weight_sums, rs, random_reassign, idx_host=None):
        """sklearn _mini_batch_step; updates centers / weight_sums in place, returns (inertia, whether any centre
        still has zero weight -- None when not known without a read-back).

        With ``exchange == "allreduce"`` and torch.distributed initialised, every rank holds the same rows
        X and the same batch; rank r assigns and sums only its slice of the batch, and ONE packed
        all-reduce (sums | weights | inertia) makes the update identical everywhere."""
        import torch

        eng = self.eng
        K = self.n_clusters
        d = _dist() if self.exchange == "allreduce" else None
        B = int(idx_host.size) if batch_idx is None else int(batch_idx.numel())
        any_zero = None
        if d is not None and d.get_world_size() > 1:
            if batch_idx is None:
                batch_idx = torch.as_tensor(idx_host, dtype=torch.int32, device=eng.device)
            w, r = d.get_world_size(), d.get_rank()
            lo, hi = (B * r) // w, (B * (r + 1)) // w
            part = batch_idx[lo:hi]
            sums, wsum = self._sums_buf
            if hi > lo:
                labels, inertia, _ = eng.kmeans_assign(X, centers, rowidx=part, want_inertia=True)
                eng.kmeans_partial_sums(X, labels, K, rowidx=part, out=self._sums_buf)
            else:
                inertia = 0.0
                self._packed.zero_()
            self._packed[-1] = inertia
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            d.all_reduce(self._packed)
            t1.record()
            self._comm_events.append((t0, t1))
            inertia = float(self._packed[-1].item())
            eng.kmeans_apply_update(centers, weight_sums, sums, wsum)
        elif B <= 4096 and int(X.shape[1]) <= 256 and idx_host is not None:
            # the whole step as one library call (indices from the host, one synchronisation)
            inertia, nzero = eng.kmeans_minibatch_step(X, idx_host, centers, weight_sums)
            any_zero = nzero > 0
        else:
            if batch_idx is None:
                batch_idx = torch.as_tensor(idx_host, dtype=torch.int32, device=eng.device)
            labels, inertia, _ = eng.kmeans_assign(X, centers, rowidx=batch_idx, want_inertia=True)
            if B <= 4096 and int(X.shape[1]) <= 256:      # sklearn's own operation order: bit-identical centres
                eng.kmeans_update_exact(X, labels, centers, weight_sums, rowidx=batch_idx)
            else:
                sums, wsum = eng.kmeans_partial_sums(X, labels, K, rowidx=batch_idx, out=self._sums_buf)
                eng.kmeans_apply_update(centers, weight_sums, sums, wsum)
        if random_reassign and self.reassignment_ratio > 0:
            ws = weight_sums.cpu().numpy()
            to_reassign = ws < self.reassignment_ratio * ws.max()
            if to_reassign.sum() > 0.5 * B:
                indices_dont_reassign = np.argsort(ws)[int(0.5 * B):]
                to_reassign[indices_dont_reassign] = False
            n_reassigns = int(to_reassign.sum())
            if n_reassigns:
                new_centers = rs.choice(B, replace=False, size=n_reassigns)
                if idx_host is not None:
                    rows = torch.as_tensor(np.asarray(idx_host)[new_centers], dtype=torch.int64, device=eng.device)
                else:
                    rows = batch_idx.long()[torch.as_tensor(new_centers, device=eng.device)]
                centers[torch.as_tensor(np.nonzero(to_reassign)[0], device=eng.device)] = X[rows]
            if (~to_reassign).any():
                ws[to_reassign] = np.min(ws[~to_reassign])
                weight_sums.copy_(torch.as_tensor(ws, device=eng.device))
            any_zero = bool((ws == 0).any())
        return inertia, any_zero

    def seed_centers(self, X):
        """First half of ``MiniBatchKMeans.fit``: the validation draw and the n_init k-means++ seedings, best one
        kept.  Returns the state ``run_loop`` continues from (the RandomState goes along, so the split changes
        no draw).  The streaming selector runs this half for the next bond type on a second stream while the
        mini-batch loop of the current one is running."""
        import torch

        eng = self.eng
        n_samples = int(X.shape[0])
        K = self.n_clusters
        batch_size = min(self.batch_size, n_samples)
        init_size = self.init_size
        if init_size is None:
            init_size = 3 * batch_size
        if init_size < K:
            init_size = 3 * K
        init_size = min(init_size, n_samples)
        rs = np.random.RandomState(self.seed)
        validation_indices = rs.randint(0, n_samples, init_size)
        valid_idx = torch.as_tensor(validation_indices, dtype=torch.int32, device=eng.device)
        best_inertia, centers = None, None
        for c in self._init_centroids_all(X, rs, init_size):
            _, inertia, _ = eng.kmeans_assign(X, c, rowidx=valid_idx, want_inertia=True)
            if best_inertia is None or inertia < best_inertia:
                centers, best_inertia = c, inertia
        return {"rs": rs, "centers": centers.contiguous(), "batch_size": batch_size}

    def run_loop(self, X, state):
        """Second half of ``MiniBatchKMeans.fit``: the mini-batch steps until the smoothed inertia stops improving."""
        import torch

        eng = self.eng
        n_samples = int(X.shape[0])
        K = self.n_clusters
        D = int(X.shape[1])
        rs, centers, batch_size = state["rs"], state["centers"], state["batch_size"]
        self._packed = torch.zeros(K * D + K + 1, dtype=torch.float64, device=eng.device)
        self._sums_buf = (self._packed[: K * D].view(K, D), self._packed[K * D: K * D + K])
        self._comm_events = []
        weight_sums = torch.zeros(K, dtype=torch.float64, device=eng.device)
        ewa, ewa_min, no_improvement = None, None, 0
        n_since_last_reassign = 0
        counts_zero = True
        zero_unknown = False         # all weights are zero before the first step
        n_steps = (self.max_iter * n_samples) // batch_size
        # scikit-learn >= 1.4 draws the mini-batch with `random_state.choice(n, batch, p=sample_weight / sum, replace=True)`
        # = searchsorted(cumsum(p) / cumsum(p)[-1], random_sample(batch), side="right") (numpy RandomState.choice);
        # the cdf is built once, with the same arithmetic
        cdf = np.cumsum(np.ones(n_samples) / float(n_samples))
        cdf /= cdf[-1]
        last = n_samples - 1

        def draw():
            # = cdf.searchsorted(u, side="right"), without the 21-level binary search through a 16 MB table
            # (220 us per step at 2M rows): the cdf is the grid (i + 1) / n up to rounding, so floor(u * n) is
            # within one of the answer and two compare-and-step rounds against the table make it exact
            u = rs.random_sample(batch_size)
            c = np.minimum((u * n_samples).astype(np.int64), last)
            for _ in range(3):
                c = c - ((c > 0) & (cdf[np.maximum(c - 1, 0)] > u))
                c = c + ((c <= last) & (cdf[np.minimum(c, last)] <= u))
            return c

        i = -1
        for i in range(n_steps):
            idx = draw()
            n_since_last_reassign += batch_size
            # weights only grow, so once no centre has a zero count the test never fires again; the fused step
            # reports the count, other paths read it back while it is still needed
            if counts_zero and zero_unknown:
                counts_zero = bool((weight_sums == 0).any().item())
            random_reassign = counts_zero or n_since_last_reassign >= 10 * K
            if random_reassign:
                n_since_last_reassign = 0
            batch_inertia, any_zero = self._mini_batch_step(X, None, centers, weight_sums, rs, random_reassign,
                                                            idx_host=idx.astype(np.int32))
            if any_zero is None:
                zero_unknown = True
            else:
                zero_unknown = False
                counts_zero = counts_zero and any_zero
            # _mini_batch_convergence (tol = 0)
            batch_inertia /= batch_size
            step = i + 1
            if step == 1:
                continue
            if ewa is None:
                ewa = batch_inertia
            else:
                alpha = min(batch_size * 2.0 / (n_samples + 1), 1)
                ewa = ewa * (1 - alpha) + batch_inertia * alpha
            if ewa_min is None or ewa < ewa_min:
                no_improvement = 0
                ewa_min = ewa
            else:
                no_improvement += 1
            if self.max_no_improvement is not None and no_improvement >= self.max_no_improvement:
                break
        self.cluster_centers_ = centers
        self.counts_ = weight_sums
        self.n_steps_ = i + 1
        self.comm_ms_ = None
        if self._comm_events:
            torch.cuda.synchronize()
            self.comm_ms_ = float(sum(a.elapsed_time(b) for a, b in self._comm_events))
            self.comm_bytes_per_step_ = int(self._packed.numel() * 8)
        self._comm_events = []
        return self

    def fit(self, X, compute_labels=True):
        """``MiniBatchKMeans.fit`` on the device rows X [n, D] (every rank holds the same X when
        torch.distributed is initialised: the streaming selector replicates its training pool)."""
        self.run_loop(X, self.seed_centers(X))
        if compute_labels:
            self.labels_, self.inertia_ = self.predict(X, with_inertia=True)
        return self

    def predict(self, X, with_inertia=False):
        """Labels of rows X against the fitted centres: tensor-core screening + float64 re-check (same labels
        as float64) for large inputs, the float64 kernel otherwise."""
        eng = self.eng
        centers = self.cluster_centers_
        if int(X.shape[1]) <= 112 and int(X.shape[0]) >= 4096:
            labels, self.n_rechecked_ = eng.kmeans_assign_tc(X, centers)
            inertia = eng.kmeans_inertia(X, centers, labels) if with_inertia else None
        else:
            labels, inertia, _ = eng.kmeans_assign(X, centers, want_inertia=with_inertia)
        return (labels, inertia) if with_inertia else labels

    def fit_predict(self, X):
        return self.fit(X).labels_


def choose_per_cluster(labels_local, n_clusters, n_each, seed, rank=0, world=1, all_counts=None, with_tags=False):
    """Per non-empty cluster pick ``n_each`` rows uniformly WITH replacement
    (``default_rng().choice(idx, n_each)``, datasetbuilder.py:378-383), seeded.

    Distributed: ``all_counts`` [world, K] holds every rank's member count per cluster; each
    rank runs the same draws and keeps the picks that fall into its own shard.  Returns the local
    row indices picked on this rank, in ascending cluster order."""
    labels_local = np.asarray(labels_local)
    order = np.argsort(labels_local, kind="stable")
    counts = np.bincount(labels_local, minlength=n_clusters)[:n_clusters]
    starts = np.concatenate([[0], np.cumsum(counts)])
    if all_counts is None:
        all_counts = counts[None, :]
    all_counts = np.asarray(all_counts)
    totals = all_counts.sum(axis=0)
    offsets = np.cumsum(all_counts, axis=0) - all_counts  # first global member index of each rank
    rng = np.random.default_rng(seed)
    # all clusters in one call (every rank draws the same numbers): n_each global member positions per
    # non-empty cluster; a rank keeps the draws that fall into its own run of the cluster's members
    nz = np.nonzero(totals[:n_clusters])[0]
    draws = rng.integers(0, totals[nz][:, None], size=(len(nz), n_each))
    loc = draws - offsets[rank, nz][:, None]
    mine = (loc >= 0) & (loc < all_counts[rank, nz][:, None])
    kk, jj = np.nonzero(mine)  # row-major: ascending cluster, then draw number
    picks = np.asarray(order[starts[nz[kk]] + loc[kk, jj]], dtype=np.int64)
    if with_tags:
        return picks, np.stack([nz[kk], jj], axis=1).astype(np.int64).reshape(-1, 2)
    return picks


def clusterdatas(engine, X, n_clusters, n_each=1, seed=0, with_tags=False):
    """``DatasetBuilder._clusterdatas`` (datasetbuilder.py:351-384) on device rows X [rows, D] held by this
    process; X is scaled in place.  Returns the selected row indices (numpy) -- with_tags adds the
    (cluster, draw) pair of every pick.  (Row sets spread over ranks or larger than HBM go through
    ``streaming.StreamingSelector``.)"""
    MinMaxScalerB200(engine).fit_transform(X)
    n_total = int(X.shape[0])
    clus = MiniBatchKMeansB200(engine, n_clusters=n_clusters, init_size=min(3 * n_clusters, n_total), n_init=3,
                               seed=seed)
    labels = clus.fit_predict(X).cpu().numpy()
    return choose_per_cluster(labels, n_clusters, n_each, seed, with_tags=with_tags)
