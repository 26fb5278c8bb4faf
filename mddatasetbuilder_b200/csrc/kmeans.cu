// kmeans.cu -- K7 assignment, K8 mini-batch centroid update, and the MinMaxScaler arithmetic.
//
// Replaces, inside DatasetBuilder._clusterdatas (reference datasetbuilder.py:351-384), the
// scikit-learn kernels it drives:
//   MinMaxScaler.fit_transform           sklearn preprocessing/_data.py:127-132,537-575
//   assignment  argmin_j ||c_j||^2 - 2 x.c_j (strict <, lowest index wins ties)
//                                        sklearn cluster/_k_means_lloyd.pyx:168-213
//   mini-batch update                    sklearn cluster/_k_means_minibatch.pyx:59-108
// The assignment is a double-precision tiled contraction (128 rows x 128 centroids per block,
// 8x8 register tiles) with the argmin fused into the epilogue, so labels are decided on the same
// quantity scikit-learn compares (float64 ||c||^2 - 2 x.c).
#include <float.h>

#include <algorithm>

#include "common.cuh"

#define AS_BM 128
#define AS_BN 128
#define AS_KC 16
#define AS_THREADS 256
#define AS_PITCH (AS_BM + 2)

__global__ void __launch_bounds__(256) k_cnorm(const double *__restrict__ C, int K, int D, double *__restrict__ cn) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= K) return;
    double s = 0.0;
    for (int k = 0; k < D; ++k) {
        double v = C[(size_t)j * D + k];
        s += v * v;
    }
    cn[j] = s;
}

// ||c_j||^2 exactly as the assignment kernel below consumes it (tc_assign.cu re-scores its candidates with it)
int kmeans_cnorm_run(mdb_ctx *c, const double *d_C, int K, int D, double *d_cn, cudaStream_t stream) {
    k_cnorm<<<cdiv(K, 256), 256, 0, stream>>>(d_C, K, D, d_cn);
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// grid: (row tiles, centroid splits).  part_val/part_idx: [splits][rows]
__global__ void __launch_bounds__(AS_THREADS) k_assign(const double *__restrict__ X, long long ldx, long long rows,
                                                       const int *__restrict__ rowidx, const double *__restrict__ C, int K,
                                                       int D, const double *__restrict__ cn, int tiles_per_split,
                                                       double *__restrict__ part_val, int *__restrict__ part_idx) {
    __shared__ double xs[AS_KC][AS_PITCH];
    __shared__ double cs[AS_KC][AS_PITCH];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * AS_BM;
    const int ctile0 = blockIdx.y * tiles_per_split;
    const int ntiles = (K + AS_BN - 1) / AS_BN;
    const int ctile1 = min(ntiles, ctile0 + tiles_per_split);
    // loader mapping: 2 threads per row, 8 consecutive features each
    const int lr = tid >> 1, lk = (tid & 1) * 8;
    long long xrow = row0 + lr;
    const bool xok = xrow < rows;
    if (xok && rowidx) xrow = rowidx[xrow];
    const double *xptr = X + (xok ? xrow : 0) * ldx;

    double bestv[8];
    int besti[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        bestv[i] = DBL_MAX;
        besti[i] = 0x7fffffff;
    }
    for (int ct = ctile0; ct < ctile1; ++ct) {
        const int col0 = ct * AS_BN;
        double acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
        const int crow = col0 + lr;
        const bool cok = crow < K;
        const double *cptr = C + (size_t)(cok ? crow : 0) * D;
        for (int k0 = 0; k0 < D; k0 += AS_KC) {
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = k0 + lk + q;
                xs[lk + q][lr] = (xok && k < D) ? xptr[k] : 0.0;
                cs[lk + q][lr] = (cok && k < D) ? cptr[k] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < AS_KC; ++kk) {
                double a[8], b[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double2 va = *reinterpret_cast<const double2 *>(&xs[kk][ty * 2 + 32 * i]);
                    const double2 vb = *reinterpret_cast<const double2 *>(&cs[kk][tx * 2 + 32 * i]);
                    a[2 * i] = va.x;
                    a[2 * i + 1] = va.y;
                    b[2 * i] = vb.x;
                    b[2 * i + 1] = vb.y;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
        }
        // epilogue: dist = ||c||^2 - 2 x.c ; per-row argmin over this tile, lowest index on ties
        double cnv[8];
        int cidx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            cidx[j] = col0 + tx * 2 + (j & 1) + 32 * (j >> 1);
            cnv[j] = cidx[j] < K ? cn[cidx[j]] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double v = DBL_MAX;
            int ix = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double dj = fma(-2.0, acc[i][j], cnv[j]);
                if (cidx[j] < K && (dj < v || (dj == v && cidx[j] < ix))) {
                    v = dj;
                    ix = cidx[j];
                }
            }
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
                if (ov < v || (ov == v && oi < ix)) {
                    v = ov;
                    ix = oi;
                }
            }
            if (v < bestv[i] || (v == bestv[i] && ix < besti[i])) {
                bestv[i] = v;
                besti[i] = ix;
            }
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long r = row0 + ty * 2 + (i & 1) + 32 * (i >> 1);
            if (r < rows) {
                part_val[(size_t)blockIdx.y * rows + r] = bestv[i];
                part_idx[(size_t)blockIdx.y * rows + r] = besti[i];
            }
        }
    }
}

// combine the centroid splits (ascending split = ascending centroid index), emit labels and
// the squared distance to the chosen centre computed directly (sklearn _inertia_dense)
__global__ void __launch_bounds__(256) k_assign_finish(const double *__restrict__ X, long long ldx, long long rows,
                                                       const int *__restrict__ rowidx, const double *__restrict__ C, int D,
                                                       int splits, const double *__restrict__ part_val,
                                                       const int *__restrict__ part_idx, int *__restrict__ labels,
                                                       double *__restrict__ mind) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double v = part_val[r];
    int ix = part_idx[r];
    for (int s = 1; s < splits; ++s) {
        const double ov = part_val[(size_t)s * rows + r];
        const int oi = part_idx[(size_t)s * rows + r];
        if (ov < v || (ov == v && oi < ix)) {
            v = ov;
            ix = oi;
        }
    }
    labels[r] = ix;
    if (mind) {
        const double *x = X + (rowidx ? (long long)rowidx[r] : r) * ldx;
        const double *c = C + (size_t)ix * D;
        double s = 0.0;
        for (int k = 0; k < D; ++k) {
            const double t = x[k] - c[k];
            s += t * t;
        }
        mind[r] = s;
    }
}

__global__ void __launch_bounds__(256) k_sum(const double *__restrict__ v, const double *__restrict__ w, long long n,
                                             double *__restrict__ out) {
    __shared__ double s[256];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += w ? v[i] * w[i] : v[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, s[0]);
}

// ---------------------------------------------------------------------------------------------
// mini-batch update: partial sums per centre (members added in ascending sample order when the
// batch fits one block, so the result does not depend on scheduling), then the running-mean step
// ---------------------------------------------------------------------------------------------
#define MB_MAX 4096
__global__ void __launch_bounds__(1024) k_mb_sums_sorted(const double *__restrict__ X, long long ldx,
                                                         const int *__restrict__ rowidx, const double *__restrict__ w,
                                                         const int *__restrict__ labels, int B, int D,
                                                         double *__restrict__ sums, double *__restrict__ wsum) {
    __shared__ unsigned long long key[MB_MAX];
    int P = 1;
    while (P < B) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        key[i] = i < B ? (((unsigned long long)(unsigned)labels[i] << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = key[i], b = key[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        key[i] = b;
                        key[l] = a;
                    }
                }
            }
            __syncthreads();
        }
    // one warp per run of equal labels; lanes span the features, members are added in order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = warp; i < B; i += nw) {
        const unsigned lab = (unsigned)(key[i] >> 32);
        if (i > 0 && (unsigned)(key[i - 1] >> 32) == lab) continue;  // not a run head
        int e = i;
        while (e < B && (unsigned)(key[e] >> 32) == lab) ++e;
        double ws = 0.0;
        for (int m = i; m < e; ++m) {
            const int smp = (int)(unsigned)key[m];
            ws += w ? w[smp] : 1.0;
        }
        for (int f0 = 0; f0 < D; f0 += 32) {
            const int f = f0 + lane;
            if (f < D) {
                double acc = 0.0;
                for (int m = i; m < e; ++m) {
                    const int smp = (int)(unsigned)key[m];
                    const double *x = X + (rowidx ? (long long)rowidx[smp] : (long long)smp) * ldx;
                    acc = __dadd_rn(acc, __dmul_rn(x[f], w ? w[smp] : 1.0));
                }
                sums[(size_t)lab * D + f] = acc;
            }
        }
        if (lane == 0) wsum[lab] = ws;
    }
}

__global__ void __launch_bounds__(256) k_mb_sums_atomic(const double *__restrict__ X, long long ldx,
                                                        const int *__restrict__ rowidx, const double *__restrict__ w,
                                                        const int *__restrict__ labels, long long B, int D,
                                                        double *__restrict__ sums, double *__restrict__ wsum) {
    const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= B) return;
    const int lane = threadIdx.x & 31;
    const int lab = labels[i];
    const double wi = w ? w[i] : 1.0;
    const double *x = X + (rowidx ? (long long)rowidx[i] : i) * ldx;
    for (int f = lane; f < D; f += 32) atomicAdd(&sums[(size_t)lab * D + f], x[f] * wi);
    if (lane == 0) atomicAdd(&wsum[lab], wi);
}

// centers = (centers * weight_sums + sums) / (weight_sums + wsum) for touched centres
// (sklearn _k_means_minibatch.pyx:84-104); weight_sums += wsum
__global__ void __launch_bounds__(256) k_mb_apply(double *__restrict__ C, double *__restrict__ weight_sums,
                                                  const double *__restrict__ sums, const double *__restrict__ wsum, int K,
                                                  int D) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= K) return;
    const int lane = threadIdx.x & 31;
    const double ws = wsum[j];
    if (!(ws > 0.0)) return;
    const double w0 = weight_sums[j];
    const double alpha = 1.0 / (w0 + ws);
    for (int f = lane; f < D; f += 32) {
        const double v = __dadd_rn(__dmul_rn(C[(size_t)j * D + f], w0), sums[(size_t)j * D + f]);
        C[(size_t)j * D + f] = __dmul_rn(v, alpha);
    }
    __syncwarp();
    if (lane == 0) weight_sums[j] = w0 + ws;
}

// The whole M-step of one mini-batch in scikit-learn's own operation order (_k_means_minibatch.pyx
// update_center_dense): per touched centre  c = c * weight_sum;  c += x_m * w_m  for its members m in
// ascending sample order;  weight_sum += sum(w);  c *= 1 / weight_sum  -- so that, labels being equal,
// the centres are bit-identical to sklearn's.  One warp per sample: the warp of the FIRST member of a
// cluster (no earlier sample carries the label) owns that centre, walks the later samples in order (label
// chunks of 32 by ballot) and adds its members feature-parallel; every other warp leaves after the check.
// With K >> batch most clusters have one or two members, so the batch spreads over the whole GPU instead of
// one block sorting it (109 -> ~10 us for 1024 rows, K = 10,000, D = 70).
#define MBX_MAXD 8  // features per lane: D <= 256
__global__ void __launch_bounds__(256) k_mb_update_exact(const double *__restrict__ X, long long ldx,
                                                         const int *__restrict__ rowidx, const double *__restrict__ w,
                                                         const int *__restrict__ labels, int B, int D,
                                                         double *__restrict__ C, double *__restrict__ weight_sums) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= B) return;
    const int lane = threadIdx.x & 31;
    const int lab = labels[i];
    // head test: any earlier sample with the same label?
    for (int j0 = 0; j0 < i; j0 += 32) {
        const int j = j0 + lane;
        const bool same = j < i && labels[j] == lab;
        if (__any_sync(0xffffffffu, same)) return;
    }
    const double w0 = weight_sums[lab];
    double acc[MBX_MAXD];
#pragma unroll
    for (int q = 0; q < MBX_MAXD; ++q) {
        const int f = q * 32 + lane;
        acc[q] = f < D ? __dmul_rn(C[(size_t)lab * D + f], w0) : 0.0;
    }
    double ws = 0.0;
    for (int j0 = i & ~31; j0 < B; j0 += 32) {
        const int j = j0 + lane;
        const bool same = j >= i && j < B && labels[j] == lab;
        unsigned mask = __ballot_sync(0xffffffffu, same);
        while (mask) {
            const int m = j0 + __ffs(mask) - 1;
            mask &= mask - 1;
            const double wm = w ? w[m] : 1.0;
            ws = __dadd_rn(ws, wm);
            const double *x = X + (rowidx ? (long long)rowidx[m] : (long long)m) * ldx;
#pragma unroll
            for (int q = 0; q < MBX_MAXD; ++q) {
                const int f = q * 32 + lane;
                if (f < D) acc[q] = __dadd_rn(acc[q], __dmul_rn(x[f], wm));
            }
        }
    }
    if (!(ws > 0.0)) return;
    const double wn = __dadd_rn(w0, ws);
    const double alpha = __ddiv_rn(1.0, wn);
#pragma unroll
    for (int q = 0; q < MBX_MAXD; ++q) {
        const int f = q * 32 + lane;
        if (f < D) C[(size_t)lab * D + f] = __dmul_rn(acc[q], alpha);
    }
    if (lane == 0) weight_sums[lab] = wn;
}

// Fixed-order sum (one block): the batch inertia feeds an early-stopping comparison, so it must not depend on
// the order in which atomics land (k_sum's block partials do).
__global__ void __launch_bounds__(1024) k_sum_det(const double *__restrict__ v, long long n, double *__restrict__ out) {
    __shared__ double s[1024];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) acc += v[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0];
}

__global__ void __launch_bounds__(1024) k_count_zero(const double *__restrict__ w, int K, double *__restrict__ out) {
    __shared__ int s[32];
    int c = 0;
    for (int i = threadIdx.x; i < K; i += 1024) c += w[i] == 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int q = 0; q < 32; ++q) t += s[q];
        out[0] = (double)t;
    }
}

int kmeans_update_exact_run(mdb_ctx *c, long long B, int D, const double *d_X, long long ldx, const int *d_rowidx,
                            const double *d_w, const int *d_labels, double *d_C, double *d_weight_sums, cudaStream_t stream) {
    if (B <= 0) return 0;
    if (B > MB_MAX || D > 32 * MBX_MAXD) {
        mdb_set_error("mdb_kmeans_update_exact: at most 4096 rows per mini-batch and 256 features");
        return -1;
    }
    ProfScope ps(c, "k_mb_update_exact", stream);
    k_mb_update_exact<<<cdiv(B, 8), 256, 0, stream>>>(d_X, ldx, d_rowidx, d_w, d_labels, (int)B, D, d_C, d_weight_sums);
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// One whole mini-batch step of MiniBatchKMeans (sklearn _mini_batch_step without the reassignment) as ONE call:
// indices host -> device through a page-locked staging buffer, assignment, fixed-order batch inertia, the M-step in
// sklearn's operation order, the number of centres still without weight -- and one synchronisation for the two
// scalars the host needs (the Python loop around it had 120 us of host time per step in separate calls).
int kmeans_minibatch_step_run(mdb_ctx *c, long long B, int D, int K, const double *d_X, long long ldx, const int *h_idx,
                              double *d_C, double *d_weight_sums, double *h_out, cudaStream_t stream) {
    if (B <= 0 || B > MB_MAX || D > 32 * MBX_MAXD) {
        mdb_set_error("mdb_kmeans_minibatch_step: 1..4096 rows per mini-batch and at most 256 features");
        return -1;
    }
    if (c->mb_cap < (size_t)B) {
        if (c->mb_hidx) cudaFreeHost(c->mb_hidx);
        if (c->mb_didx) cudaFree(c->mb_didx);
        c->mb_hidx = nullptr;
        c->mb_didx = nullptr;
        const size_t cap = MB_MAX;
        MDB_CUDA(cudaMallocHost(&c->mb_hidx, cap * sizeof(int) + 64));
        // device block: idx | labels | mindist | out[2]
        MDB_CUDA(cudaMalloc(&c->mb_didx, cap * (2 * sizeof(int) + sizeof(double)) + 64));
        c->mb_cap = cap;
    }
    int *hidx = (int *)c->mb_hidx;
    double *hout = (double *)((char *)c->mb_hidx + MB_MAX * sizeof(int));
    int *didx = (int *)c->mb_didx;
    int *dlab = didx + MB_MAX;
    double *dmind = (double *)(dlab + MB_MAX);
    double *dout = (double *)((char *)c->mb_didx + MB_MAX * (2 * sizeof(int) + sizeof(double)));  // 64 spare bytes
    memcpy(hidx, h_idx, (size_t)B * sizeof(int));
    MDB_CUDA(cudaMemcpyAsync(didx, hidx, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (kmeans_assign_run(c, B, D, K, d_X, ldx, didx, d_C, dlab, dmind, nullptr, nullptr, stream)) return -1;
    k_sum_det<<<1, 1024, 0, stream>>>(dmind, B, dout);
    if (kmeans_update_exact_run(c, B, D, d_X, ldx, didx, nullptr, dlab, d_C, d_weight_sums, stream)) return -1;
    k_count_zero<<<1, 1024, 0, stream>>>(d_weight_sums, K, dout + 1);
    c->launches += 2;
    MDB_CUDA(cudaMemcpyAsync(hout, dout, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    MDB_CUDA(cudaStreamSynchronize(stream));
    h_out[0] = hout[0];
    h_out[1] = hout[1];
    MDB_CUDA(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------------------------------------
// MinMaxScaler
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long dbl_order(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dbl_unorder(unsigned long long u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}

__global__ void __launch_bounds__(256) k_colminmax(const double *__restrict__ X, long long ldx, long long rows, int D,
                                                   unsigned long long *__restrict__ omin,
                                                   unsigned long long *__restrict__ omax) {
    // thread -> column (threadIdx.x % D-tile), rows strided by blocks
    const int cols_per_block = 32;
    const int c = blockIdx.y * cols_per_block + (threadIdx.x & 31);
    const int rlane = threadIdx.x >> 5, nr = blockDim.x >> 5;
    double mn = DBL_MAX, mx = -DBL_MAX;
    if (c < D)
        for (long long r = (long long)blockIdx.x * nr + rlane; r < rows; r += (long long)gridDim.x * nr) {
            const double v = X[r * ldx + c];
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
    __shared__ double smn[8][32], smx[8][32];
    smn[rlane][threadIdx.x & 31] = mn;
    smx[rlane][threadIdx.x & 31] = mx;
    __syncthreads();
    if (rlane == 0 && c < D) {
        for (int k = 1; k < nr; ++k) {
            mn = fmin(mn, smn[k][threadIdx.x]);
            mx = fmax(mx, smx[k][threadIdx.x]);
        }
        if (mn <= mx) {
            atomicMin(&omin[c], dbl_order(mn));
            atomicMax(&omax[c], dbl_order(mx));
        }
    }
}

__global__ void __launch_bounds__(256) k_minmax_init(unsigned long long *omin, unsigned long long *omax, int D) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < D) {
        omin[c] = dbl_order(DBL_MAX);
        omax[c] = dbl_order(-DBL_MAX);
    }
}

__global__ void __launch_bounds__(256) k_minmax_decode(const unsigned long long *omin, const unsigned long long *omax,
                                                       double *mn, double *mx, int D, int accumulate) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < D) {
        const double a = dbl_unorder(omin[c]), b = dbl_unorder(omax[c]);
        if (!accumulate) {
            mn[c] = a;
            mx[c] = b;
        } else if (a <= b) {  // running bounds over several calls (streaming scaler fit)
            mn[c] = fmin(mn[c], a);
            mx[c] = fmax(mx[c], b);
        }
    }
}

// X = X * scale + min_ as two rounded operations (sklearn _data.py:574-575)
__global__ void __launch_bounds__(256) k_scale(double *__restrict__ X, long long ldx, long long rows, int D,
                                               const double *__restrict__ scale, const double *__restrict__ shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * D) return;
    const long long r = i / D;
    const int c = (int)(i - r * D);
    double *p = X + r * ldx + c;
    *p = __dadd_rn(__dmul_rn(*p, scale[c]), shift[c]);
}

// ---------------------------------------------------------------------------------------------
// k-means++ seeding step (sklearn cluster/_kmeans.py:180-279): squared distances from T candidate
// rows to every row, clipped against the current closest distance, and the weighted potential of
// each candidate.  One warp per row; lanes span the features.
// ---------------------------------------------------------------------------------------------
#define KPP_MAXT 32
__global__ void __launch_bounds__(256) k_kpp_step(const double *__restrict__ X, long long ldx, const int *__restrict__ rowidx,
                                                  long long n, int D, const int *__restrict__ cand, int T,
                                                  const double *__restrict__ closest, const double *__restrict__ w,
                                                  double *__restrict__ dist, double *__restrict__ pot) {
    __shared__ double s_pot[8][KPP_MAXT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < KPP_MAXT) s_pot[warp][lane] = 0.0;
    __syncwarp();
    for (long long i = (long long)blockIdx.x * 8 + warp; i < n; i += (long long)gridDim.x * 8) {
        const double *x = X + (rowidx ? (long long)rowidx[i] : i) * ldx;
        const double cl = closest ? closest[i] : DBL_MAX;
        const double wi = w ? w[i] : 1.0;
        for (int t = 0; t < T; ++t) {
            const int ci = cand[t];
            const double *y = X + (rowidx ? (long long)rowidx[ci] : (long long)ci) * ldx;
            double s = 0.0;
            for (int k = lane; k < D; k += 32) {
                const double df = x[k] - y[k];
                s += df * df;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            s = fmin(s, cl);
            if (lane == 0) {
                dist[(size_t)t * n + i] = s;
                s_pot[warp][t] += s * wi;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < T) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += s_pot[k][threadIdx.x];
        atomicAdd(&pot[threadIdx.x], s);
    }
}

int kpp_step_run(mdb_ctx *c, long long n, int D, const double *d_X, long long ldx, const int *d_rowidx, const int *d_cand,
                 int T, const double *d_closest, const double *d_w, double *d_dist, double *d_pot, cudaStream_t stream) {
    if (T < 1 || T > KPP_MAXT) {
        mdb_set_error("kpp_step: between 1 and 32 candidates");
        return -1;
    }
    MDB_CUDA(cudaMemsetAsync(d_pot, 0, (size_t)T * sizeof(double), stream));
    if (n > 0) {
        ProfScope ps(c, "k_kpp_step", stream);
        k_kpp_step<<<(unsigned)std::min<long long>((n + 7) / 8, 148 * 8), 256, 0, stream>>>(d_X, ldx, d_rowidx, n, D, d_cand, T,
                                                                                         d_closest, d_w, d_dist, d_pot);
        c->launches++;
    }
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// The whole k-means++ seeding loop on the device (sklearn cluster/_kmeans.py:180-279, unit sample
// weights): K - 1 steps of { draw T candidates in proportion to the closest squared distance,
// score each by the potential it would leave, keep the best }.  The uniform variates come from the
// caller ((K-1) x T of them, drawn in sklearn's order), everything else -- cumulative sum, the
// searchsorted draw, candidate distances, argmin, the update of `closest` -- stays on the GPU, two
// launches per step and no host round trip (the Python loop with its three .item() syncs per step
// took 230 us per step; K = 10,000 with n_init = 3 is 30,000 steps).
// Sums are taken in a fixed order (per-thread run, block scan, per-block partials combined serially),
// so the seeding is reproducible bit for bit.
// ---------------------------------------------------------------------------------------------
struct KppState {
    int best;  // winning candidate of the previous step: dist[best] is the current `closest`
    int pad;
};

// Finish step c-1 (potential of each candidate from the per-block partials, argmin with the lowest
// index on ties like np.argmin) and draw the T candidates of step c: searchsorted(cumsum(closest),
// u * pot), side = left, clipped to n - 1 (:241-245).  The cumulative sum is taken hierarchically --
// the per-block partial sums of dist[best] that the distance pass already produced, then the 256
// rows of the block that holds the target -- so nothing of length n is touched here.  One warp per
// candidate (in turns).  Everything another block wrote is read with __ldcg (the persistent kernel
// below runs this inside one launch, where L1 is not coherent across SMs).
__device__ __forceinline__ void kpp_select_pick_body(long long n, int T, int Tprev, int c, int K,
                                                     const double *__restrict__ rnd, const double *partial, int nblk,
                                                     const double *dist, int *cand, int *indices, KppState *st, double *s_dyn,
                                                     double *s_pot, int *s_best, int *s_cand, double *s_rnd) {
    double *s_pre = s_dyn;               // [nblk + 1] prefix of the block sums of dist[best]
    double *s_par = s_dyn + (nblk + 1);  // [Tprev][nblk] partials, one row per candidate
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    // the previous step's candidates and this step's variates travel with the first round trip
    // instead of costing one of their own later (s_cand: old candidates in, new candidates out)
    if (tid < KPP_MAXT) {
        const int oldc = tid < Tprev ? __ldcg(cand + tid) : 0;
        const double u = (tid < T && c < K) ? rnd[(size_t)(c - 1) * T + tid] : 0.0;
        s_cand[tid] = oldc;
        s_rnd[tid] = u;
    }
    for (int q = tid; q < nblk * Tprev; q += blockDim.x) {
        const int t = q / nblk, b = q - t * nblk;
        s_par[q] = __ldcg(partial + (size_t)b * KPP_MAXT + t);
    }
    __syncthreads();
    // potential of candidate t: lane-strided runs over the blocks, then a fixed shuffle tree
    for (int t = warp; t < Tprev; t += nwarps) {
        double p = 0.0;
        for (int b = lane; b < nblk; b += 32) p += s_par[t * nblk + b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if (lane == 0) s_pot[t] = p;
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int t = 1; t < Tprev; ++t)
            if (s_pot[t] < s_pot[best]) best = t;
        *s_best = best;
        st->best = best;
        indices[c - 1] = s_cand[best];
        s_pre[0] = 0.0;
    }
    __syncthreads();
    if (warp == 0) {  // inclusive prefix of the winner's block sums, 32 at a time in a fixed order
        const int best = *s_best;
        double base = 0.0;
        for (int b0 = 0; b0 < nblk; b0 += 32) {
            const int b = b0 + lane;
            double v = b < nblk ? s_par[best * nblk + b] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            if (b < nblk) s_pre[b + 1] = base + v;
            base += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    if (c >= K) return;  // the last call only selects
    const int best = *s_best;
    for (int t = warp; t < T; t += nwarps) {
        const double r = s_rnd[t] * s_pot[best];
        // block whose cumulative range holds r: first b with pre[b + 1] >= r
        int lo = 0, hi = nblk;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_pre[mid + 1] < r) lo = mid + 1;
            else hi = mid;
        }
        long long pick = n - 1;
        if (lo < nblk) {
            // the 256 rows of that block in one round trip: 8 consecutive rows per lane, a running sum
            // inside the lane, a warp scan over the lane totals, then the first row that reaches r
            const double *d = dist + (size_t)best * n;
            const long long i0 = (long long)lo * 256 + 8 * lane;
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = (i0 + q < n) ? __ldcg(d + i0 + q) : 0.0;
#pragma unroll
            for (int q = 1; q < 8; ++q) v[q] += v[q - 1];
            double incl = v[7];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            const double base = s_pre[lo] + (incl - v[7]);  // everything before this lane's rows
            int first = 8;
#pragma unroll
            for (int q = 7; q >= 0; --q)
                if (i0 + q < n && base + v[q] >= r) first = q;
            const unsigned hit = __ballot_sync(0xffffffffu, first < 8);
            if (hit) {
                const int src = __ffs(hit) - 1;
                pick = (long long)lo * 256 + 8 * src + __shfl_sync(0xffffffffu, first, src);
            } else {
                pick = min(n - 1, (long long)lo * 256 + 256);  // rounding left r just past this block's rows
            }
        }
        if (lane == 0) {
            cand[t] = (int)pick;
            s_cand[t] = (int)pick;
        }
    }
}

__global__ void __launch_bounds__(1024) k_kpp_select_pick(long long n, int T, int Tprev, int c, int K,
                                                          const double *__restrict__ rnd,  // [(K-1)][T]
                                                          const double *__restrict__ partial, int nblk,
                                                          const double *__restrict__ dist, int *__restrict__ cand,
                                                          int *__restrict__ indices, KppState *st) {
    // blockIdx.y = independent seeding problem (sklearn's n_init runs share nothing but X)
    rnd += (size_t)blockIdx.y * (size_t)max(K - 1, 1) * T;
    partial += (size_t)blockIdx.y * nblk * KPP_MAXT;
    dist += (size_t)blockIdx.y * T * n;
    cand += blockIdx.y * KPP_MAXT;
    indices += (size_t)blockIdx.y * K;
    st += blockIdx.y;
    extern __shared__ double s_dyn[];
    __shared__ double s_pot[KPP_MAXT], s_rnd[KPP_MAXT];
    __shared__ int s_best, s_cand[KPP_MAXT];
    kpp_select_pick_body(n, T, Tprev, c, K, rnd, partial, nblk, dist, cand, indices, st, s_dyn, s_pot, &s_best, s_cand, s_rnd);
}

// squared distances of the T candidate rows to every row as scikit-learn takes them in
// _kmeans_plusplus (euclidean_distances with precomputed norms: ||x||^2 - 2 x.c + ||c||^2, clipped at
// 0), then clipped against the current closest distance (= dist[best] of the previous step, read
// before it is overwritten; nothing to clip in step 0).  One thread per row, one DFMA per (candidate,
// feature); candidates staged in shared memory feature-major so that two of them come per LDS.128;
// ||x||^2 is computed in step 0 and kept in xn; per-block partial potentials in a fixed order.
template <int TMAX>
__device__ __forceinline__ void kpp_dist_body(const double *__restrict__ X, long long ldx, const int *__restrict__ rowidx,
                                              long long n, int D, const int *cand, int T, int first_step, const KppState *st,
                                              double *dist, double *partial_blk, double *s_c, int chunk, bool stage,
                                              double *xn, const double *xt, const double *crows = nullptr,
                                              long long *tr = nullptr) {
    const bool tron = tr && blockIdx.x == 1 && blockIdx.y == 0 && threadIdx.x == 0;
    long long tq = tron ? clock64() : 0;
    double *s_cn = s_c + (size_t)D * TMAX;  // [TMAX] squared norms of the candidates
    double *s_w = s_cn + TMAX;              // [8][KPP_MAXT] warp partials
    if (stage) {
        for (int q = threadIdx.x; q < TMAX * D; q += 256) {
            const int t = q / D, k = q - t * D;
            double v = 0.0;
            if (t < T) {
                if (crows) {  // rows already gathered by the block that picked them: one hop instead of two
                    v = __ldcg(crows + q);
                } else {
                    const int ci = __ldcg(cand + t);
                    v = X[(rowidx ? (long long)rowidx[ci] : (long long)ci) * ldx + k];
                }
            }
            s_c[k * TMAX + t] = v;
        }
        __syncthreads();
        if (threadIdx.x < TMAX) {
            double a2 = 0.0;
            for (int k = 0; k < D; ++k) a2 = fma(s_c[k * TMAX + threadIdx.x], s_c[k * TMAX + threadIdx.x], a2);
            s_cn[threadIdx.x] = a2;
        }
    }
    __syncthreads();
    if (tron) {
        const long long now = clock64();
        tr[3] += now - tq;  // staging the candidates
        tq = now;
    }
    const long long i = (long long)chunk * 256 + threadIdx.x;
    double acc[TMAX];
#pragma unroll
    for (int t = 0; t < TMAX; ++t) acc[t] = 0.0;
    if (i < n) {
        const double cl = first_step ? DBL_MAX : dist[(size_t)__ldcg(&st->best) * n + i];  // this thread's own write
        // the rows come from the feature-major copy made once per seeding (k_kpp_transpose): consecutive
        // threads read consecutive doubles, instead of one 32-byte sector per thread and feature
        const double *x = xt + i;
        double x2 = 0.0;
#pragma unroll 5
        for (int k = 0; k < D; ++k) {
            const double xv = x[(size_t)k * n];
            if (first_step) x2 = fma(xv, xv, x2);
            const double2 *cr = reinterpret_cast<const double2 *>(s_c + k * TMAX);
#pragma unroll
            for (int t = 0; t < TMAX; t += 2) {
                const double2 cv = cr[t >> 1];
                acc[t] = fma(xv, cv.x, acc[t]);
                acc[t + 1] = fma(xv, cv.y, acc[t + 1]);
            }
        }
        if (first_step) xn[i] = x2;
        else x2 = xn[i];
#pragma unroll
        for (int t = 0; t < TMAX; ++t) {
            const double dsq = fmax(fma(-2.0, acc[t], x2 + s_cn[t]), 0.0);
            acc[t] = fmin(dsq, cl);
            if (t < T) dist[(size_t)t * n + i] = acc[t];
        }
    }
    if (tron) {
        const long long now = clock64();
        tr[4] += now - tq;  // rows x candidates (thread 0's own)
        tq = now;
    }
    // block potentials in a fixed order: lanes by shuffle tree, warps serially
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
        double v = acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_w[warp * KPP_MAXT + t] = v;
    }
    __syncthreads();
    if (threadIdx.x < T) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += s_w[w * KPP_MAXT + threadIdx.x];
        partial_blk[threadIdx.x] = v;
    }
    __syncthreads();  // s_w is reused by the next chunk of a persistent block
    if (tron) tr[5] += clock64() - tq;  // potentials
}

// feature-major copy of the seeding rows of every problem: xt[b][k][i] = X[row(b, i)][k]
__global__ void __launch_bounds__(256) k_kpp_transpose(const double *__restrict__ X, long long ldx, const int *__restrict__ rowidx,
                                                       long long n, int D, double *__restrict__ xt) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double *x = X + (rowidx ? (long long)rowidx[(size_t)blockIdx.y * n + i] : i) * ldx;
    double *o = xt + (size_t)blockIdx.y * D * n + i;
    for (int k = 0; k < D; ++k) o[(size_t)k * n] = x[k];
}

template <int TMAX>
__global__ void __launch_bounds__(256) k_kpp_dist(const double *__restrict__ X, long long ldx, const int *__restrict__ rowidx,
                                                  long long n, int D, const int *__restrict__ cand, int T, int Tfull,
                                                  int first_step, const KppState *__restrict__ st,
                                                  double *__restrict__ dist, double *__restrict__ partial,
                                                  double *__restrict__ xn, const double *__restrict__ xt) {
    xn += (size_t)blockIdx.y * n;
    xt += (size_t)blockIdx.y * D * n;
    if (rowidx) rowidx += (size_t)blockIdx.y * n;
    cand += blockIdx.y * KPP_MAXT;
    st += blockIdx.y;
    dist += (size_t)blockIdx.y * Tfull * n;
    partial += ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * KPP_MAXT;
    extern __shared__ double s_c[];
    kpp_dist_body<TMAX>(X, ldx, rowidx, n, D, cand, T, first_step, st, dist, partial, s_c, blockIdx.x, true, xn, xt);
}

// All K - 1 steps in ONE launch.  After its distance pass every block of a seeding problem takes a
// ticket; the block that draws the last ticket of the step runs the select / pick, gathers the rows
// of the new candidates and raises the step flag the others spin on -- one wait per step instead of
// two full barriers, and the candidates' coordinates travel with the flag.  Launched co-operatively,
// so every block is resident; the spin is bounded (a protocol bug traps instead of hanging the GPU).
template <int TMAX>
__global__ void __launch_bounds__(256) k_kpp_persistent(const double *__restrict__ X, long long ldx,
                                                        const int *__restrict__ rowidx, long long n, int D, int T, int K,
                                                        const double *__restrict__ rnd, int *cand, KppState *st, double *dist,
                                                        double *partial, int *indices, unsigned *sync, int nblk,
                                                        double *crows, long long *trace, double *xn,
                                                        const double *__restrict__ xt) {
    xn += (size_t)blockIdx.y * n;
    xt += (size_t)blockIdx.y * D * n;
    const int nb = gridDim.x;  // blocks per problem; each takes the 256-row chunks blockIdx.x, + nb, ...
    if (rowidx) rowidx += (size_t)blockIdx.y * n;
    rnd += (size_t)blockIdx.y * (size_t)max(K - 1, 1) * T;
    cand += blockIdx.y * KPP_MAXT;
    st += blockIdx.y;
    dist += (size_t)blockIdx.y * T * n;
    partial += (size_t)blockIdx.y * nblk * KPP_MAXT;
    indices += (size_t)blockIdx.y * K;
    crows += (size_t)blockIdx.y * KPP_MAXT * D;
    unsigned *tickets = sync + blockIdx.y * 64;   // per problem: ticket counter and step flag, a cache line apart
    unsigned *flag = tickets + 32;
    extern __shared__ double s_dyn[];
    __shared__ double s_pot[KPP_MAXT], s_rnd[KPP_MAXT];
    __shared__ int s_best, s_cand[KPP_MAXT];
    __shared__ int s_last;
    for (int ch = blockIdx.x; ch < nblk; ch += nb)
        kpp_dist_body<TMAX>(X, ldx, rowidx, n, D, cand, 1, 1, st, dist, partial + (size_t)ch * KPP_MAXT, s_dyn, ch,
                            ch == (int)blockIdx.x, xn, xt);
    int Tprev = 1;
    long long tr0 = 0, tr1 = 0, tr2 = 0, trt = clock64();
    for (int step = 1; step <= K; ++step) {
        __syncthreads();
        if (trace && blockIdx.x == 1 && blockIdx.y == 0 && threadIdx.x == 0) {
            const long long now = clock64();
            tr0 += now - trt;  // distance pass
            trt = now;
        }
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(tickets, 1u) == (unsigned)nb * (unsigned)step - 1u;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            kpp_select_pick_body(n, T, Tprev, step, K, rnd, partial, nblk, dist, cand, indices, st, s_dyn, s_pot, &s_best,
                                 s_cand, s_rnd);
            __syncthreads();
            if (step < K)  // the new candidates' rows, from the feature-major copy (no row-index hop)
                for (int q = threadIdx.x; q < T * D; q += 256) {
                    const int t = q / D, k = q - t * D;
                    crows[q] = xt[(size_t)k * n + s_cand[t]];
                }
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicExch(flag, (unsigned)step);
            }
        }
        if (step == K) break;
        if (threadIdx.x == 0) {
            volatile unsigned *vf = flag;
            unsigned spins = 0;
            while (*vf < (unsigned)step) {
                if (++spins > (1u << 28)) __trap();
            }
            __threadfence();
        }
        __syncthreads();
        if (trace && blockIdx.x == 1 && blockIdx.y == 0 && threadIdx.x == 0) {
            const long long now = clock64();
            if (s_last) tr2 += now - trt; else tr1 += now - trt;  // select by this block / wait for the flag
            trt = now;
        }
        for (int ch = blockIdx.x; ch < nblk; ch += nb)
            kpp_dist_body<TMAX>(X, ldx, rowidx, n, D, cand, T, 0, st, dist, partial + (size_t)ch * KPP_MAXT, s_dyn, ch,
                                ch == (int)blockIdx.x, xn, xt, crows, trace);
        Tprev = T;
    }
    if (trace && blockIdx.x == 1 && blockIdx.y == 0 && threadIdx.x == 0) {
        trace[0] = tr0;
        trace[1] = tr1;
        trace[2] = tr2;
    }
}

int kmeans_plusplus_run(mdb_ctx *c, int B, long long n, int D, const double *d_X, long long ldx, const int *d_rowidx, int K,
                        int T, const double *h_rand, const int *h_first, int *d_indices, cudaStream_t stream) {
    if (B < 1 || B > 64 || n < 1 || K < 1 || K > n || T < 1 || T > KPP_MAXT) {
        mdb_set_error("kmeans_plusplus: need 1 <= K <= n, 1 <= T <= 32, 1 <= batch <= 64");
        return -1;
    }
    for (int b = 0; b < B; ++b)
        if (h_first[b] < 0 || h_first[b] >= n) {
            mdb_set_error("kmeans_plusplus: first centre out of range");
            return -1;
        }
    const int nblk = (int)cdiv(n, 256);
    const size_t nrand = (size_t)std::max(K - 1, 1) * T;
    size_t need = align256((size_t)B * T * n * sizeof(double)) + align256((size_t)B * n * sizeof(double)) +
                  align256((size_t)B * D * n * sizeof(double)) + align256((size_t)B * nblk * KPP_MAXT * sizeof(double)) +
                  align256((size_t)B * nrand * sizeof(double)) + align256((size_t)B * KPP_MAXT * sizeof(int)) +
                  align256((size_t)B * sizeof(KppState)) + align256((size_t)B * 64 * sizeof(unsigned)) +
                  align256((size_t)B * KPP_MAXT * D * sizeof(double)) + 256 + 1024;
    if (arena_reserve(c, need, stream)) return -1;
    arena_reset(c);
    double *dist = (double *)arena_alloc(c, (size_t)B * T * n * sizeof(double));
    double *xn = (double *)arena_alloc(c, (size_t)B * n * sizeof(double));
    double *xt = (double *)arena_alloc(c, (size_t)B * D * n * sizeof(double));
    double *partial = (double *)arena_alloc(c, (size_t)B * nblk * KPP_MAXT * sizeof(double));
    double *rnd = (double *)arena_alloc(c, (size_t)B * nrand * sizeof(double));
    int *cand = (int *)arena_alloc(c, (size_t)B * KPP_MAXT * sizeof(int));
    KppState *st = (KppState *)arena_alloc(c, (size_t)B * sizeof(KppState));
    if (!dist || !xn || !xt || !partial || !rnd || !cand || !st) {
        mdb_set_error("kmeans_plusplus: arena too small (internal sizing error)");
        return -1;
    }
    k_kpp_transpose<<<dim3((unsigned)nblk, (unsigned)B), 256, 0, stream>>>(d_X, ldx, d_rowidx, n, D, xt);
    c->launches += 1;
    if (K > 1) MDB_CUDA(cudaMemcpyAsync(rnd, h_rand, (size_t)B * nrand * sizeof(double), cudaMemcpyHostToDevice, stream));
    std::vector<int> hc((size_t)B * KPP_MAXT, 0);
    for (int b = 0; b < B; ++b) hc[(size_t)b * KPP_MAXT] = h_first[b];
    MDB_CUDA(cudaMemcpyAsync(cand, hc.data(), hc.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    // candidate width of the instantiation (multiple of 2: candidates are read two per LDS.128)
    const int TW = T <= 4 ? 4 : T <= 8 ? 8 : T <= 12 ? 12 : T <= 16 ? 16 : T <= 24 ? 24 : 32;
    const size_t smem = ((size_t)TW * D + TW + 8 * KPP_MAXT) * sizeof(double);
    const size_t smem_sel = ((size_t)(nblk + 1) + (size_t)nblk * T) * sizeof(double);
    if (smem > 200 * 1024 || smem_sel > 200 * 1024) {
        mdb_set_error("kmeans_plusplus: too many features or seeding rows for the shared-memory staging");
        return -1;
    }
#define MDB_KPP_WIDTHS(M) M(4) M(8) M(12) M(16) M(24) M(32)
#define MDB_KPP_ATTR(W) \
    if (TW == W) MDB_CUDA(cudaFuncSetAttribute(k_kpp_dist<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MDB_KPP_WIDTHS(MDB_KPP_ATTR)
#undef MDB_KPP_ATTR
    MDB_CUDA(cudaFuncSetAttribute(k_kpp_select_pick, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sel));
    const dim3 gd((unsigned)nblk, (unsigned)B), gs(1, (unsigned)B);
    // one co-operative launch for the whole loop when every block can be resident at once
    {
        const size_t smem_p = std::max(smem, smem_sel);
        void *fn = nullptr;
#define MDB_KPP_FN(W) \
    if (TW == W) fn = (void *)k_kpp_persistent<W>;
        MDB_KPP_WIDTHS(MDB_KPP_FN)
#undef MDB_KPP_FN
        int per_sm = 0, coop = 0;
        cudaDeviceProp prop;
        MDB_CUDA(cudaGetDeviceProperties(&prop, c->device));
        MDB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
        MDB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
        MDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, smem_p));
        // blocks per problem: as many as fit at once, each taking several 256-row chunks if need be
        const int nb = (int)std::min<long long>(nblk, (long long)per_sm * prop.multiProcessorCount / B);
        if (coop && nb >= 1 && nblk <= 16 * nb && !getenv("MDB_KPP_NO_PERSISTENT")) {
            unsigned *bar = (unsigned *)arena_alloc(c, (size_t)B * 64 * sizeof(unsigned));
            double *crows = (double *)arena_alloc(c, (size_t)B * KPP_MAXT * D * sizeof(double));
            if (!bar || !crows) {
                mdb_set_error("kmeans_plusplus: arena too small (internal sizing error)");
                return -1;
            }
            MDB_CUDA(cudaMemsetAsync(bar, 0, (size_t)B * 64 * sizeof(unsigned), stream));
            const double *aX = d_X;
            long long aldx = ldx, an = n;
            const int *arow = d_rowidx;
            int aD = D, aT = T, aK = K;
            const double *arnd = rnd;
            const double *axt = xt;
            int anblk = nblk;
            long long *trace = nullptr;
            if (getenv("MDB_KPP_TRACE")) {  // developer switch: cycles of block 1 in the distance pass / waiting / selecting
                trace = (long long *)arena_alloc(c, 256);
                if (trace) MDB_CUDA(cudaMemsetAsync(trace, 0, 256, stream));
            }
            void *args[] = {&aX,   &aldx, &arow,    &an,        &aD,  &aT,    &aK,    &arnd,
                            &cand, &st,   &dist,    &partial,   &d_indices, &bar, &anblk, &crows, &trace, &xn, &axt};
            {
                ProfScope ps(c, "k_kpp_persistent", stream);
                MDB_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)nb, (unsigned)B), dim3(256), args, smem_p, stream));
            }
            c->launches += 1;
            MDB_CUDA(cudaGetLastError());
            MDB_CUDA(cudaStreamSynchronize(stream));
            if (trace) {
                long long h[6];
                MDB_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
                fprintf(stderr, "k_kpp_persistent block 1 (nb %d, chunks %d), cycles per step: distance pass %.0f (staging %.0f, rows %.0f, "
                        "potentials %.0f), waiting %.0f, selecting %.0f\n",
                        nb, nblk, (double)h[0] / K, (double)h[3] / K, (double)h[4] / K, (double)h[5] / K, (double)h[1] / K, (double)h[2] / K);
            }
            return 0;
        }
    }
    auto launch_dist = [&](int Tn, int first_step) {
#define MDB_KPP_LAUNCH(W) \
    if (TW == W) k_kpp_dist<W><<<gd, 256, smem, stream>>>(d_X, ldx, d_rowidx, n, D, cand, Tn, T, first_step, st, dist, partial, xn, xt);
        MDB_KPP_WIDTHS(MDB_KPP_LAUNCH)
#undef MDB_KPP_LAUNCH
    };
    {
        // step 0: the first centre alone (T = 1, nothing to clip against)
        launch_dist(1, 1);
        int Tprev = 1;
        for (int step = 1; step <= K; ++step) {
            {
                ProfScope ps(c, "k_kpp_select_pick", stream);
                k_kpp_select_pick<<<gs, 1024, smem_sel, stream>>>(n, T, Tprev, step, K, rnd, partial, nblk, dist, cand, d_indices,
                                                                  st);
            }
            if (step < K) {
                ProfScope ps(c, "k_kpp_dist", stream);
                launch_dist(T, 0);
            }
            Tprev = T;
        }
    }
    c->launches += 2ull * K;
    MDB_CUDA(cudaGetLastError());
    MDB_CUDA(cudaStreamSynchronize(stream));  // the host arrays are the caller's
    return 0;
}

// ---------------------------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------------------------
int kmeans_assign_run(mdb_ctx *c, long long rows, int D, int K, const double *d_X, long long ldx, const int *d_rowidx,
                      const double *d_C, int *d_labels, double *d_mind, double *h_inertia, const double *d_w,
                      cudaStream_t stream) {
    if (rows <= 0) {
        if (h_inertia) *h_inertia = 0.0;
        return 0;
    }
    if (K <= 0 || D <= 0) {
        mdb_set_error("kmeans_assign: K and D must be positive");
        return -1;
    }
    const int ntiles = (K + AS_BN - 1) / AS_BN;
    const long long rtiles = (rows + AS_BM - 1) / AS_BM;
    // enough blocks for ~2 waves over 148 SMs even for a 1024-row mini-batch
    int splits = (int)std::min<long long>(ntiles, std::max<long long>(1, (296 + rtiles - 1) / rtiles));
    const int tps = (ntiles + splits - 1) / splits;
    splits = (ntiles + tps - 1) / tps;
    const bool need_mind = d_mind != nullptr || h_inertia != nullptr;
    size_t need = align256((size_t)K * sizeof(double)) + align256((size_t)splits * rows * sizeof(double)) +
                  align256((size_t)splits * rows * sizeof(int)) + (need_mind && !d_mind ? align256(rows * sizeof(double)) : 0) +
                  1024;
    if (arena_reserve(c, need, stream)) return -1;
    arena_reset(c);
    double *cn = (double *)arena_alloc(c, (size_t)K * sizeof(double));
    double *pv = (double *)arena_alloc(c, (size_t)splits * rows * sizeof(double));
    int *pi = (int *)arena_alloc(c, (size_t)splits * rows * sizeof(int));
    double *mind = d_mind;
    if (need_mind && !mind) mind = (double *)arena_alloc(c, rows * sizeof(double));
    double *acc = (double *)arena_alloc(c, 256);
    if (!cn || !pv || !pi || !acc || (need_mind && !mind)) {
        mdb_set_error("kmeans_assign: arena too small (internal sizing error)");
        return -1;
    }
    {
        ProfScope ps(c, "k_cnorm", stream);
        k_cnorm<<<cdiv(K, 256), 256, 0, stream>>>(d_C, K, D, cn);
    }
    {
        ProfScope ps(c, "k_assign", stream);
        dim3 grid((unsigned)rtiles, (unsigned)splits);
        k_assign<<<grid, AS_THREADS, 0, stream>>>(d_X, ldx, rows, d_rowidx, d_C, K, D, cn, tps, pv, pi);
    }
    {
        ProfScope ps(c, "k_assign_finish", stream);
        k_assign_finish<<<cdiv(rows, 256), 256, 0, stream>>>(d_X, ldx, rows, d_rowidx, d_C, D, splits, pv, pi, d_labels, mind);
    }
    c->launches += 3;
    if (h_inertia) {
        if (!d_w && rows <= (1LL << 20)) {
            k_sum_det<<<1, 1024, 0, stream>>>(mind, rows, acc);  // fixed order: reproducible to the last bit
        } else {
            MDB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), stream));
            k_sum<<<std::min<unsigned>(cdiv(rows, 256), 1024u), 256, 0, stream>>>(mind, d_w, rows, acc);
        }
        c->launches++;
        MDB_CUDA(cudaMemcpyAsync(h_inertia, acc, sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// inertia of given labels: sum_i w_i ||x_i - c_label(i)||^2 (sklearn _inertia_dense)
__global__ void __launch_bounds__(256) k_label_dist(const double *__restrict__ X, long long ldx, long long rows,
                                                    const int *__restrict__ rowidx, const double *__restrict__ C, int D,
                                                    const int *__restrict__ labels, double *__restrict__ mind) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double *x = X + (rowidx ? (long long)rowidx[r] : r) * ldx;
    const double *cc = C + (size_t)labels[r] * D;
    double s = 0.0;
    for (int k = 0; k < D; ++k) {
        const double t = x[k] - cc[k];
        s += t * t;
    }
    mind[r] = s;
}

int kmeans_inertia_run(mdb_ctx *c, long long rows, int D, const double *d_X, long long ldx, const int *d_rowidx,
                       const double *d_C, const int *d_labels, const double *d_w, double *h_inertia, cudaStream_t stream) {
    *h_inertia = 0.0;
    if (rows <= 0) return 0;
    if (arena_reserve(c, align256(rows * sizeof(double)) + 1024, stream)) return -1;
    arena_reset(c);
    double *mind = (double *)arena_alloc(c, rows * sizeof(double));
    double *acc = (double *)arena_alloc(c, 256);
    if (!mind || !acc) return -1;
    k_label_dist<<<cdiv(rows, 256), 256, 0, stream>>>(d_X, ldx, rows, d_rowidx, d_C, D, d_labels, mind);
    MDB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), stream));
    k_sum<<<std::min<unsigned>(cdiv(rows, 256), 1024u), 256, 0, stream>>>(mind, d_w, rows, acc);
    c->launches += 2;
    MDB_CUDA(cudaMemcpyAsync(h_inertia, acc, sizeof(double), cudaMemcpyDeviceToHost, stream));
    MDB_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

int kmeans_sums_run(mdb_ctx *c, long long B, int D, int K, const double *d_X, long long ldx, const int *d_rowidx,
                    const double *d_w, const int *d_labels, double *d_sums, double *d_wsum, cudaStream_t stream) {
    MDB_CUDA(cudaMemsetAsync(d_sums, 0, (size_t)K * D * sizeof(double), stream));
    MDB_CUDA(cudaMemsetAsync(d_wsum, 0, (size_t)K * sizeof(double), stream));
    if (B <= 0) return 0;
    ProfScope ps(c, "k_mb_sums", stream);
    if (B <= MB_MAX)
        k_mb_sums_sorted<<<1, 1024, 0, stream>>>(d_X, ldx, d_rowidx, d_w, d_labels, (int)B, D, d_sums, d_wsum);
    else
        k_mb_sums_atomic<<<cdiv(B, 8), 256, 0, stream>>>(d_X, ldx, d_rowidx, d_w, d_labels, B, D, d_sums, d_wsum);
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

int kmeans_apply_run(mdb_ctx *c, int K, int D, double *d_C, double *d_weight_sums, const double *d_sums,
                     const double *d_wsum, cudaStream_t stream) {
    ProfScope ps(c, "k_mb_apply", stream);
    k_mb_apply<<<cdiv(K, 8), 256, 0, stream>>>(d_C, d_weight_sums, d_sums, d_wsum, K, D);
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

int minmax_fit_run(mdb_ctx *c, long long rows, int D, const double *d_X, long long ldx, double *d_min, double *d_max,
                   int accumulate, cudaStream_t stream) {
    size_t need = 2 * align256((size_t)D * sizeof(unsigned long long)) + 1024;
    if (arena_reserve(c, need, stream)) return -1;
    arena_reset(c);
    unsigned long long *omin = (unsigned long long *)arena_alloc(c, (size_t)D * sizeof(unsigned long long));
    unsigned long long *omax = (unsigned long long *)arena_alloc(c, (size_t)D * sizeof(unsigned long long));
    if (!omin || !omax) return -1;
    k_minmax_init<<<cdiv(D, 256), 256, 0, stream>>>(omin, omax, D);
    if (rows > 0) {
        ProfScope ps(c, "k_colminmax", stream);
        dim3 grid((unsigned)std::min<long long>(2048, (rows + 7) / 8), cdiv(D, 32));
        k_colminmax<<<grid, 256, 0, stream>>>(d_X, ldx, rows, D, omin, omax);
    }
    c->launches += 2;
    k_minmax_decode<<<cdiv(D, 256), 256, 0, stream>>>(omin, omax, d_min, d_max, D, accumulate);
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

int minmax_transform_run(mdb_ctx *c, long long rows, int D, double *d_X, long long ldx, const double *h_min,
                         const double *h_max, cudaStream_t stream) {
    // scale_ = 1 / range with range < 10 eps -> 1 ; min_ = 0 - data_min * scale_   (_data.py:127-132,537-541)
    std::vector<double> sc(2 * (size_t)D);
    for (int k = 0; k < D; ++k) {
        double range = h_max[k] - h_min[k];
        if (range < 10 * DBL_EPSILON) range = 1.0;
        sc[k] = 1.0 / range;
        sc[D + k] = 0.0 - h_min[k] * sc[k];
    }
    size_t need = align256(2 * (size_t)D * sizeof(double)) + 1024;
    if (arena_reserve(c, need, stream)) return -1;
    arena_reset(c);
    double *d_sc = (double *)arena_alloc(c, 2 * (size_t)D * sizeof(double));
    if (!d_sc) return -1;
    MDB_CUDA(cudaMemcpyAsync(d_sc, sc.data(), 2 * (size_t)D * sizeof(double), cudaMemcpyHostToDevice, stream));
    MDB_CUDA(cudaStreamSynchronize(stream));  // sc is a stack-lifetime host buffer
    if (rows > 0) {
        ProfScope ps(c, "k_scale", stream);
        k_scale<<<cdiv(rows * D, 256), 256, 0, stream>>>(d_X, ldx, rows, D, d_sc, d_sc + D);
        c->launches++;
    }
    MDB_CUDA(cudaGetLastError());
    return 0;
}
