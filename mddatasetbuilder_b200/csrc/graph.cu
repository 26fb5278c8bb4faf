// graph.cu -- K3 bond-type keys and K4 connectivity (union-find with pointer jumping), plus the
// exact reference `dps` visiting order for one frame.
//
// K3 replaces the per-atom key of readatombondtype (reference detect.py:105-119, 219-225);
// K4 replaces the partition computed by the Cython/C++ `dps` (dps.pyx:17-46, c_stack.cpp).
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// keys
// ---------------------------------------------------------------------------------------------
template <int MAXB>
__global__ void __launch_bounds__(256) k_keys(const int32_t *__restrict__ ell, const uint8_t *__restrict__ types, int N,
                                              long long fn, uint64_t *__restrict__ keys, int32_t *__restrict__ parent) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const int i = (int)(a % N);
    const int4 *row = reinterpret_cast<const int4 *>(ell + a * MAXB);
    int deg = 0, first = -1;
#pragma unroll
    for (int k = 0; k < MAXB / 4; ++k) {
        int4 q = row[k];
        if (k == 0) first = q.x;
        deg += (q.x >= 0) + (q.y >= 0) + (q.z >= 0) + (q.w >= 0);
    }
    if (keys) {
        // dump-only tier 1: every surviving bond has order 1 -> levels = [1] * degree
        const int nl = deg > 12 ? 12 : deg;
        const uint64_t ones = nl == 0 ? 0ull : (0x111111111111ull >> (4 * (12 - nl)));
        keys[a] = ((uint64_t)(types[i] + 1) << 56) | ((uint64_t)deg << 48) | ones;
    }
    if (parent) parent[a] = (first >= 0 && first < i) ? first : i;  // rows are ascending: first = smallest
}

template <int MAXB>
__global__ void __launch_bounds__(256) k_keys_bo(const int32_t *__restrict__ ell, const double *__restrict__ bo,
                                                 const uint8_t *__restrict__ types, int N, long long fn,
                                                 uint64_t *__restrict__ keys) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const int i = (int)(a % N);
    int lv[MAXB];
    int c = 0;
#pragma unroll
    for (int k = 0; k < MAXB; ++k) {
        if (ell[a * MAXB + k] >= 0) {
            int l = (int)rint(bo[a * MAXB + k]);  // round-half-even, like Python's round()
            lv[c++] = l < 1 ? 1 : (l > 15 ? 15 : l);
        }
    }
    for (int x = 1; x < c; ++x) {
        int v = lv[x], b = x - 1;
        while (b >= 0 && lv[b] > v) {
            lv[b + 1] = lv[b];
            --b;
        }
        lv[b + 1] = v;
    }
    uint64_t key = ((uint64_t)(types[i] + 1) << 56) | ((uint64_t)c << 48);
    for (int x = 0; x < c && x < 12; ++x) key |= (uint64_t)lv[x] << (4 * x);
    keys[a] = key;
}

int keys_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, const uint8_t *d_types, uint64_t *d_keys,
             int32_t *d_parent, cudaStream_t stream) {
    long long fn = (long long)nframes * natoms;
    if (fn == 0) return 0;
    {
        ProfScope ps(ctx, "k_keys", stream);
        if (ctx->opt_max_bonds == 8)
            k_keys<8><<<cdiv(fn, 256), 256, 0, stream>>>(d_ell, d_types, natoms, fn, d_keys, d_parent);
        else
            k_keys<16><<<cdiv(fn, 256), 256, 0, stream>>>(d_ell, d_types, natoms, fn, d_keys, d_parent);
    }
    ctx->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

int keys_bo_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, const double *d_bo, const uint8_t *d_types,
                uint64_t *d_keys, cudaStream_t stream) {
    long long fn = (long long)nframes * natoms;
    if (fn == 0) return 0;
    if (ctx->opt_max_bonds == 8)
        k_keys_bo<8><<<cdiv(fn, 256), 256, 0, stream>>>(d_ell, d_bo, d_types, natoms, fn, d_keys);
    else
        k_keys_bo<16><<<cdiv(fn, 256), 256, 0, stream>>>(d_ell, d_bo, d_types, natoms, fn, d_keys);
    ctx->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// union-find: hooks always attach the larger root under the smaller one, so the final root of a
// component is its smallest atom index -- the atom dps starts that molecule from (dps.pyx:27-30).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_root(int *par, int x) {
    int p = par[x];
    while (p != x) {
        int gp = par[p];
        if (gp != p) par[x] = gp;  // path halving; racing writers only ever store ancestors
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void uf_union(int *par, int a, int b) {
    int ra = uf_root(par, a), rb = uf_root(par, b);
    while (ra != rb) {
        if (ra < rb) {
            int t = ra;
            ra = rb;
            rb = t;
        }
        // ra > rb: hook ra under rb if ra is still a root
        int old = atomicCAS(&par[ra], ra, rb);
        if (old == ra) break;
        ra = uf_root(par, old);
        rb = uf_root(par, rb);
    }
}

template <int MAXB>
__global__ void __launch_bounds__(256) k_uf_hook(const int32_t *__restrict__ ell, int N, long long fn, int *parent) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const int i = (int)(a % N);
    int *par = parent + (a - i);
    const int4 *row = reinterpret_cast<const int4 *>(ell + a * MAXB);
#pragma unroll
    for (int k = 0; k < MAXB / 4; ++k) {
        int4 q = row[k];
        if (q.x >= 0 && q.x < i) uf_union(par, i, q.x);
        if (q.y >= 0 && q.y < i) uf_union(par, i, q.y);
        if (q.z >= 0 && q.z < i) uf_union(par, i, q.z);
        if (q.w >= 0 && q.w < i) uf_union(par, i, q.w);
    }
}

// keys + hooks in one pass over the rows (parent already seeded by k_bonds / k_cleanup)
template <int MAXB>
__global__ void __launch_bounds__(256) k_keys_hook(const int32_t *__restrict__ ell, const uint8_t *__restrict__ types,
                                                   int N, long long fn, uint64_t *__restrict__ keys, int *parent) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const int i = (int)(a % N);
    int *par = parent + (a - i);
    const int4 *row = reinterpret_cast<const int4 *>(ell + a * MAXB);
    int nb[MAXB];
#pragma unroll
    for (int k = 0; k < MAXB / 4; ++k) {
        int4 q = row[k];
        nb[4 * k] = q.x;
        nb[4 * k + 1] = q.y;
        nb[4 * k + 2] = q.z;
        nb[4 * k + 3] = q.w;
    }
    int deg = 0;
#pragma unroll
    for (int k = 0; k < MAXB; ++k) deg += nb[k] >= 0;
    if (keys) {
        const int nl = deg > 12 ? 12 : deg;
        const uint64_t ones = nl == 0 ? 0ull : (0x111111111111ull >> (4 * (12 - nl)));
        keys[a] = ((uint64_t)(types[i] + 1) << 56) | ((uint64_t)deg << 48) | ones;
    }
    // slot 0 (the smallest neighbour) is already the seed; hook the remaining smaller neighbours
#pragma unroll
    for (int k = 1; k < MAXB; ++k)
        if (nb[k] >= 0 && nb[k] < i) uf_union(par, i, nb[k]);
}

__global__ void __launch_bounds__(256) k_uf_flatten(int N, long long fn, int *parent) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const int i = (int)(a % N);
    int *par = parent + (a - i);
    int r = par[i];
    while (true) {
        int g = par[r];
        if (g == r) break;
        r = g;
    }
    par[i] = r;
}

int components_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, int32_t *d_labels, int seeded,
                   cudaStream_t stream) {
    long long fn = (long long)nframes * natoms;
    if (fn == 0) return 0;
    if (!seeded) {
        if (keys_run(ctx, nframes, natoms, d_ell, nullptr, nullptr, d_labels, stream)) return -1;
    }
    {
        ProfScope ps(ctx, "k_uf_hook", stream);
        if (ctx->opt_max_bonds == 8)
            k_uf_hook<8><<<cdiv(fn, 256), 256, 0, stream>>>(d_ell, natoms, fn, d_labels);
        else
            k_uf_hook<16><<<cdiv(fn, 256), 256, 0, stream>>>(d_ell, natoms, fn, d_labels);
    }
    {
        ProfScope ps(ctx, "k_uf_flatten", stream);
        k_uf_flatten<<<cdiv(fn, 256), 256, 0, stream>>>(natoms, fn, d_labels);
    }
    ctx->launches += 2;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

int keys_components_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, const uint8_t *d_types,
                        uint64_t *d_keys, int32_t *d_labels, cudaStream_t stream) {
    long long fn = (long long)nframes * natoms;
    if (fn == 0) return 0;
    {
        ProfScope ps(ctx, "k_keys_hook", stream);
        if (ctx->opt_max_bonds == 8)
            k_keys_hook<8><<<cdiv(fn, 256), 256, 0, stream>>>(d_ell, d_types, natoms, fn, d_keys, d_labels);
        else
            k_keys_hook<16><<<cdiv(fn, 256), 256, 0, stream>>>(d_ell, d_types, natoms, fn, d_keys, d_labels);
    }
    {
        ProfScope ps(ctx, "k_uf_flatten", stream);
        k_uf_flatten<<<cdiv(fn, 256), 256, 0, stream>>>(natoms, fn, d_labels);
    }
    ctx->launches += 2;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// exact dps output for one frame from a CSR in adjacency order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_csr_hook(const int *__restrict__ rowptr, const int *__restrict__ col, int N,
                                                  int *parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
        int j = col[p];
        if (j >= 0 && j < N && j != i) uf_union(parent, i, j);
    }
}

__global__ void __launch_bounds__(256) k_iota(int N, int *a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) a[i] = i;
}

// size[root] += 1, work[root] += deg + 1 (stack demand of the LIFO search: one push per edge + root)
__global__ void __launch_bounds__(256) k_mol_tally(const int *__restrict__ rowptr, const int *__restrict__ label, int N,
                                                   int *size, int *work, int *isroot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    if (i == N) {
        isroot[N] = 0;
        return;
    }
    int r = label[i];
    atomicAdd(&size[r], 1);
    atomicAdd(&work[r], rowptr[i + 1] - rowptr[i] + 1);
    isroot[i] = (r == i);
}

// after the scans: size_ofs / work_ofs / rank hold exclusive prefix sums over atoms
__global__ void __launch_bounds__(128) k_dps_walk(const int *__restrict__ rowptr, const int *__restrict__ col,
                                                  const int *__restrict__ label, const int *__restrict__ rank,
                                                  const int *__restrict__ size_ofs, const int *__restrict__ work_ofs, int N,
                                                  int *stack, unsigned char *visited, int *mol_atoms, int *mol_ptr,
                                                  int *bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (label[i] != i) return;
    const int out0 = size_ofs[i];
    const int want = size_ofs[i + 1] - out0;
    mol_ptr[rank[i]] = out0;
    if (rank[i + 1] == rank[N]) mol_ptr[rank[N]] = N;  // last molecule closes the pointer array
    int *st = stack + work_ofs[i];
    int sp = 0, w = 0;
    st[sp++] = i;
    while (sp > 0) {
        int s = st[--sp];
        if (visited[s]) continue;
        if (w < want) mol_atoms[out0 + w] = s;
        ++w;
        for (int p = rowptr[s]; p < rowptr[s + 1]; ++p) {
            int j = col[p];
            if (!visited[j]) st[sp++] = j;
        }
        visited[s] = 1;
    }
    if (w != want) atomicExch(bad, 1);  // directed reachability != undirected component
}

int dps_run(mdb_ctx *ctx, int natoms, const int32_t *d_rowptr, const int32_t *d_col, int32_t *d_mol_atoms,
            int32_t *d_mol_ptr, int *h_nmol, cudaStream_t stream) {
    const int N = natoms;
    if (N == 0) {
        *h_nmol = 0;
        MDB_CUDA(cudaMemsetAsync(d_mol_ptr, 0, sizeof(int), stream));
        return 0;
    }
    int nnz = 0;
    MDB_CUDA(cudaMemcpyAsync(&nnz, d_rowptr + N, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MDB_CUDA(cudaStreamSynchronize(stream));
    size_t n16 = (size_t)N + 16;
    size_t need = 6 * align256(n16 * sizeof(int)) + align256((size_t)nnz * sizeof(int) + n16 * sizeof(int)) +
                  align256(n16) + align256(scan_scratch_bytes(N + 1)) + 4096;
    if (arena_reserve(ctx, need, stream)) return -1;
    arena_reset(ctx);
    int *label = (int *)arena_alloc(ctx, n16 * sizeof(int));
    int *size = (int *)arena_alloc(ctx, n16 * sizeof(int));
    int *work = (int *)arena_alloc(ctx, n16 * sizeof(int));
    int *rank = (int *)arena_alloc(ctx, n16 * sizeof(int));
    int *stack = (int *)arena_alloc(ctx, (size_t)nnz * sizeof(int) + n16 * sizeof(int));
    unsigned char *visited = (unsigned char *)arena_alloc(ctx, n16);
    void *tmp = arena_alloc(ctx, scan_scratch_bytes(N + 1));
    int *bad = (int *)arena_alloc(ctx, 256);
    if (!label || !size || !work || !rank || !stack || !visited || !tmp || !bad) {
        mdb_set_error("dps_run: arena too small (internal sizing error)");
        return -1;
    }
    MDB_CUDA(cudaMemsetAsync(size, 0, n16 * sizeof(int), stream));
    MDB_CUDA(cudaMemsetAsync(work, 0, n16 * sizeof(int), stream));
    MDB_CUDA(cudaMemsetAsync(visited, 0, n16, stream));
    MDB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    k_iota<<<cdiv(N, 256), 256, 0, stream>>>(N, label);
    k_csr_hook<<<cdiv(N, 256), 256, 0, stream>>>(d_rowptr, d_col, N, label);
    k_uf_flatten<<<cdiv(N, 256), 256, 0, stream>>>(N, N, label);
    k_mol_tally<<<cdiv(N + 1, 256), 256, 0, stream>>>(d_rowptr, label, N, size, work, rank);
    ctx->launches += 4;
    // size/work are indexed by root and zero elsewhere: exclusive scans give per-molecule offsets
    if (device_exclusive_scan(ctx, size, N + 1, tmp, stream)) return -1;
    if (device_exclusive_scan(ctx, work, N + 1, tmp, stream)) return -1;
    if (device_exclusive_scan(ctx, rank, N + 1, tmp, stream)) return -1;
    k_dps_walk<<<cdiv(N, 128), 128, 0, stream>>>(d_rowptr, d_col, label, rank, size, work, N, stack, visited, d_mol_atoms,
                                                 d_mol_ptr, bad);
    ctx->launches++;
    MDB_CUDA(cudaGetLastError());
    int h[2] = {0, 0};
    MDB_CUDA(cudaMemcpyAsync(&h[0], rank + N, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MDB_CUDA(cudaMemcpyAsync(&h[1], bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MDB_CUDA(cudaStreamSynchronize(stream));
    if (h[1]) {
        mdb_set_error("mdb_dps: adjacency is not symmetric (a directed search from the smallest index does not reach "
                      "every atom of its connected component)");
        return -1;
    }
    *h_nmol = h[0];
    return 0;
}
