// stream.cu -- device-side bookkeeping of the streaming passes (SURVEY.md §7 hard part 3):
//
//   * bond-type keys and molecule labels straight from a CSR bond table (bond-file mode),
//   * the group-by of DatasetBuilder._readtimestepsbond (reference datasetbuilder.py:185-214: per frame
//     {key -> [atom ids]}, appended per type) as a key dictionary + stable counting sort on the device,
//   * the per-type max_counter of the parent loop (datasetbuilder.py:262-272),
//   * running column bounds for the scaler, row gathers for the training pool,
//   * the per-cluster selection (datasetbuilder.py:378-383) on stored labels: member counts per
//     (segment, cluster) and "the p-th member of cluster k" without sorting the labels.
//
// Everything here is integer / byte work bound by HBM bandwidth; nothing is reshaped into a GEMM.
#include <float.h>

#include <algorithm>

#include "common.cuh"

#define MDB_KEY_EMPTY 0xffffffffffffffffull
#define MDB_CODE_NONE 0xffffu

// ---------------------------------------------------------------------------------------------
// keys / labels from CSR (DetectBond.readatombondtype, detect.py:105-119; readmolecule :137-144)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_keys_csr(const int32_t *__restrict__ rowptr, const double *__restrict__ bo,
                                                  const uint8_t *__restrict__ types, int N, long long fn, long long colcap,
                                                  uint64_t *__restrict__ keys) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const long long f = a / N;
    const int i = (int)(a - f * N);
    const int32_t *rp = rowptr + f * (N + 1);
    const double *b = bo + f * colcap;
    const int p0 = rp[i], p1 = rp[i + 1];
    const int c = p1 - p0;
    uint64_t key = ((uint64_t)(types[i] + 1) << 56) | ((uint64_t)(c & 0xff) << 48);
    // levels = sorted(max(1, round_half_even(bo))): insertion into a packed nibble list (<= 12 levels)
    int lv[12];
    int m = 0;
    for (int p = p0; p < p1 && m < 12; ++p) {
        int l = (int)rint(b[p]);
        l = l < 1 ? 1 : (l > 15 ? 15 : l);
        int q = m - 1;
        while (q >= 0 && lv[q] > l) {
            lv[q + 1] = lv[q];
            --q;
        }
        lv[q + 1] = l;
        ++m;
    }
    for (int x = 0; x < m; ++x) key |= (uint64_t)lv[x] << (4 * x);
    keys[a] = key;
}

__device__ __forceinline__ int uf_root2(int *par, int x) {
    int p = par[x];
    while (p != x) {
        int gp = par[p];
        if (gp != p) par[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union2(int *par, int a, int b) {
    int ra = uf_root2(par, a), rb = uf_root2(par, b);
    while (ra != rb) {
        if (ra < rb) {
            int t = ra;
            ra = rb;
            rb = t;
        }
        int old = atomicCAS(&par[ra], ra, rb);
        if (old == ra) break;
        ra = uf_root2(par, old);
        rb = uf_root2(par, rb);
    }
}

__global__ void __launch_bounds__(256) k_iota_frames(int N, long long fn, int *__restrict__ lab) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a < fn) lab[a] = (int)(a % N);
}

__global__ void __launch_bounds__(256) k_csr_hook_frames(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                         int N, long long fn, long long colcap, int *lab) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const long long f = a / N;
    const int i = (int)(a - f * N);
    const int32_t *rp = rowptr + f * (N + 1);
    const int32_t *cl = col + f * colcap;
    int *par = lab + f * N;
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
        const int j = cl[p];
        if (j >= 0 && j < i) uf_union2(par, i, j);  // every bond is listed from both ends
    }
}

__global__ void __launch_bounds__(256) k_flatten_frames(int N, long long fn, int *lab) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const int i = (int)(a % N);
    int *par = lab + (a - i);
    int r = par[i];
    while (true) {
        const int g = par[r];
        if (g == r) break;
        r = g;
    }
    par[i] = r;
}

extern "C" int mdb_bond_keys_csr(mdb_ctx *c, int nframes, int natoms, const int32_t *d_rowptr, const double *d_bo,
                                 long long colcap, const uint8_t *d_types, uint64_t *d_keys, void *stream) {
    if (!c) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    const long long fn = (long long)nframes * natoms;
    if (fn == 0) return 0;
    {
        ProfScope ps(c, "k_keys_csr", s);
        k_keys_csr<<<cdiv(fn, 256), 256, 0, s>>>(d_rowptr, d_bo, d_types, natoms, fn, colcap, d_keys);
    }
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mdb_components_csr(mdb_ctx *c, int nframes, int natoms, const int32_t *d_rowptr, const int32_t *d_col,
                                  long long colcap, int32_t *d_labels, void *stream) {
    if (!c) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    const long long fn = (long long)nframes * natoms;
    if (fn == 0) return 0;
    {
        ProfScope ps(c, "k_iota_frames", s);
        k_iota_frames<<<cdiv(fn, 256), 256, 0, s>>>(natoms, fn, d_labels);
    }
    {
        ProfScope ps(c, "k_csr_hook_frames", s);
        k_csr_hook_frames<<<cdiv(fn, 256), 256, 0, s>>>(d_rowptr, d_col, natoms, fn, colcap, d_labels);
    }
    {
        ProfScope ps(c, "k_flatten_frames", s);
        k_flatten_frames<<<cdiv(fn, 256), 256, 0, s>>>(natoms, fn, d_labels);
    }
    c->launches += 3;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// key dictionary: 64-bit key -> slot of an open-addressing table that lives for the whole run.
// Slot numbers depend on insertion races, so callers translate through the table; everything
// keyed by slot (counts, first occurrence, the order of rows inside a slot) is deterministic.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t key_hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}

__global__ void __launch_bounds__(256) k_key_codes(const uint64_t *__restrict__ keys, const uint8_t *__restrict__ keep,
                                                   long long fn, int N, long long step0, int cap,
                                                   unsigned long long *__restrict__ tab_keys,
                                                   unsigned long long *__restrict__ tab_count,
                                                   unsigned long long *__restrict__ tab_first,
                                                   uint16_t *__restrict__ codes, int *__restrict__ err) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = a < fn && (!keep || keep[a]);
    const unsigned active = __ballot_sync(0xffffffffu, live);
    if (a >= fn) return;
    if (!live) {
        codes[a] = MDB_CODE_NONE;
        return;
    }
    const uint64_t key = keys[a];
    const unsigned peers = __match_any_sync(active, key);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    int slot = -1;
    if (lane == leader) {
        uint32_t h = key_hash(key) & (uint32_t)(cap - 1);
        for (int probe = 0; probe < cap; ++probe) {
            const unsigned long long old = atomicCAS(&tab_keys[h], MDB_KEY_EMPTY, (unsigned long long)key);
            if (old == MDB_KEY_EMPTY || old == key) {
                slot = (int)h;
                break;
            }
            h = (h + 1) & (uint32_t)(cap - 1);
        }
        if (slot < 0)
            atomicExch(err, MDB_ERR_KEY_TABLE_FULL);
        else {
            atomicAdd(&tab_count[slot], (unsigned long long)__popc(peers));
            // first occurrence as (step, atom): the leader is the lowest lane = the smallest atom-frame
            const long long f = a / N;
            const unsigned long long pos = ((unsigned long long)(step0 + f) << 32) | (unsigned long long)(a - f * N);
            atomicMin(&tab_first[slot], pos);
        }
    }
    slot = __shfl_sync(peers, slot, leader);
    codes[a] = slot < 0 ? MDB_CODE_NONE : (uint16_t)slot;
}

// stable counting sort of the atom-frames of a batch by slot ---------------------------------
#define GB_TILE 8192
#define GB_WARPS 8

__global__ void __launch_bounds__(256) k_gb_hist(const uint16_t *__restrict__ codes, long long fn, int cap, int nblk,
                                                 int *__restrict__ hist /* [cap][nblk] */) {
    extern __shared__ int s_h[];  // [cap]
    for (int k = threadIdx.x; k < cap; k += blockDim.x) s_h[k] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * GB_TILE;
    for (int q = threadIdx.x; q < GB_TILE; q += blockDim.x) {
        const long long a = base + q;
        const bool live = a < fn;
        const uint16_t cd = live ? codes[a] : MDB_CODE_NONE;
        const bool cnt = live && cd != MDB_CODE_NONE;
        const unsigned act = __ballot_sync(0xffffffffu, cnt);
        if (cnt) {
            const unsigned peers = __match_any_sync(act, (unsigned)cd);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_h[cd], __popc(peers));
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < cap; k += blockDim.x) hist[(size_t)k * nblk + blockIdx.x] = s_h[k];
}

// hist has been scanned exclusively in (slot-major, block-minor) order: hist[k][b] = first output
// position of the rows of slot k that lie in tile b.
__global__ void __launch_bounds__(GB_WARPS * 32) k_gb_scatter(const uint16_t *__restrict__ codes, long long fn, int cap, int nblk,
                                                              const int *__restrict__ hist, int32_t *__restrict__ order) {
    extern __shared__ int s_w[];  // [GB_WARPS][cap]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < GB_WARPS * cap; k += blockDim.x) s_w[k] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * GB_TILE + (long long)w * (GB_TILE / GB_WARPS);
    int *mine = s_w + w * cap;
    // pass 1: this warp's counts per slot
    for (int q = lane; q < GB_TILE / GB_WARPS; q += 32) {
        const long long a = base + q;
        const bool live = a < fn;
        const uint16_t cd = live ? codes[a] : MDB_CODE_NONE;
        const bool cnt = live && cd != MDB_CODE_NONE;
        const unsigned act = __ballot_sync(0xffffffffu, cnt);
        if (cnt) {
            const unsigned peers = __match_any_sync(act, (unsigned)cd);
            if (lane == __ffs(peers) - 1) mine[cd] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over the warps of the block, offset by the tile's global position
    for (int k = threadIdx.x; k < cap; k += blockDim.x) {
        int run = hist[(size_t)k * nblk + blockIdx.x];
        for (int ww = 0; ww < GB_WARPS; ++ww) {
            const int v = s_w[ww * cap + k];
            s_w[ww * cap + k] = run;
            run += v;
        }
    }
    __syncthreads();
    // pass 2: ranks in atom-frame order
    for (int q = lane; q < GB_TILE / GB_WARPS; q += 32) {
        const long long a = base + q;
        const bool live = a < fn;
        const uint16_t cd = live ? codes[a] : MDB_CODE_NONE;
        const bool cnt = live && cd != MDB_CODE_NONE;
        const unsigned act = __ballot_sync(0xffffffffu, cnt);
        if (cnt) {
            const unsigned peers = __match_any_sync(act, (unsigned)cd);
            const int leader = __ffs(peers) - 1;
            int cur = 0;
            if (lane == leader) {
                cur = mine[cd];
                mine[cd] = cur + __popc(peers);
            }
            cur = __shfl_sync(peers, cur, leader);
            order[cur + __popc(peers & ((1u << lane) - 1u))] = (int32_t)a;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_gb_offsets(const int *__restrict__ hist, int cap, int nblk, int total_kept_pos,
                                                    long long *__restrict__ seg) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < cap) seg[k] = hist[(size_t)k * nblk];
    if (k == cap) seg[cap] = hist[(size_t)cap * nblk];  // one extra scanned element = number of kept rows
    (void)total_kept_pos;
}

extern "C" int mdb_group_keys(mdb_ctx *c, int nframes, int natoms, const uint64_t *d_keys, const uint8_t *d_keep,
                              long long step0, int cap, uint64_t *d_tab_keys, long long *d_tab_count,
                              long long *d_tab_first, uint16_t *d_codes, int32_t *d_order, long long *d_seg,
                              void *stream) {
    if (!c) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (cap < 32 || cap > 4096 || (cap & (cap - 1))) {
        mdb_set_error("mdb_group_keys: cap must be a power of two in 32..4096");
        return -1;
    }
    const long long fn = (long long)nframes * natoms;
    if (fn >= (1LL << 31) - 64) {
        mdb_set_error("mdb_group_keys: batch too large, split the frames");
        return -1;
    }
    if (fn == 0) {
        if (d_seg) MDB_CUDA(cudaMemsetAsync(d_seg, 0, (size_t)(cap + 1) * sizeof(long long), s));
        return 0;
    }
    {
        ProfScope ps(c, "k_key_codes", s);
        k_key_codes<<<cdiv(fn, 256), 256, 0, s>>>(d_keys, d_keep, fn, natoms, step0, cap,
                                                  reinterpret_cast<unsigned long long *>(d_tab_keys),
                                                  reinterpret_cast<unsigned long long *>(d_tab_count),
                                                  reinterpret_cast<unsigned long long *>(d_tab_first), d_codes,
                                                  &c->d_stats->err);
    }
    c->launches++;
    if (!d_order) {
        MDB_CUDA(cudaGetLastError());
        return 0;
    }
    const int nblk = (int)cdiv(fn, GB_TILE);
    const long long nh = (long long)cap * nblk + 1;
    size_t need = align256((size_t)nh * sizeof(int)) + align256(scan_scratch_bytes(nh)) + 4096;
    if (arena_reserve(c, need, s)) return -1;
    arena_reset(c);
    int *hist = (int *)arena_alloc(c, (size_t)nh * sizeof(int));
    void *tmp = arena_alloc(c, scan_scratch_bytes(nh));
    if (!hist || !tmp) {
        mdb_set_error("mdb_group_keys: arena too small (internal sizing error)");
        return -1;
    }
    MDB_CUDA(cudaMemsetAsync(hist + (nh - 1), 0, sizeof(int), s));
    {
        ProfScope ps(c, "k_gb_hist", s);
        k_gb_hist<<<nblk, 256, (size_t)cap * sizeof(int), s>>>(d_codes, fn, cap, nblk, hist);
    }
    c->launches++;
    if (device_exclusive_scan(c, hist, nh, tmp, s)) return -1;
    const size_t smem = (size_t)GB_WARPS * cap * sizeof(int);
    MDB_CUDA(cudaFuncSetAttribute(k_gb_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ProfScope ps(c, "k_gb_scatter", s);
        k_gb_scatter<<<nblk, GB_WARPS * 32, smem, s>>>(d_codes, fn, cap, nblk, hist, d_order);
    }
    c->launches++;
    if (d_seg) {
        k_gb_offsets<<<cdiv(cap + 1, 256), 256, 0, s>>>(hist, cap, nblk, 0, d_seg);
        c->launches++;
    }
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// per-type max_counter: maxc[slot][el] = max over the targets of that slot of counts[t][el]
// (datasetbuilder.py:262-272, `max_counter |= symbols`, per bond type)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_type_maxcount(const int32_t *__restrict__ targets, long long T,
                                                       const uint16_t *__restrict__ codes, const int *__restrict__ counts,
                                                       int nt, int *__restrict__ maxc /* [cap][MDB_MAXT] */) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < T;
    uint16_t cd = MDB_CODE_NONE;
    if (live) cd = codes[targets ? targets[t] : t];
    const bool ok = live && cd != MDB_CODE_NONE;
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    if (!ok) return;
    const unsigned peers = __match_any_sync(act, (unsigned)cd);
    const int leader = __ffs(peers) - 1;
    for (int e = 0; e < nt; ++e) {
        int v = counts[t * nt + e];
        // max over the peer group (tree over the lanes of `peers` is not worth it: groups are whole warps
        // when the targets are sorted by slot, scattered otherwise -- use the generic reduce)
        v = __reduce_max_sync(peers, v);
        if ((threadIdx.x & 31) == leader && v > 0) atomicMax(&maxc[(int)cd * MDB_MAXT + e], v);
    }
}

extern "C" int mdb_type_maxcount(mdb_ctx *c, const int32_t *d_targets, long long ntargets, const uint16_t *d_codes,
                                 const int32_t *d_counts, int32_t *d_maxc, void *stream) {
    if (!c || !c->have_elements) {
        mdb_set_error("mdb_type_maxcount: call mdb_set_elements first");
        return -1;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (ntargets <= 0) return 0;
    {
        ProfScope ps(c, "k_type_maxcount", s);
        k_type_maxcount<<<cdiv(ntargets, 256), 256, 0, s>>>(d_targets, ntargets, d_codes, d_counts, c->el.nt, d_maxc);
    }
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// rows: running column bounds, gathers
// ---------------------------------------------------------------------------------------------
extern "C" int mdb_minmax_accumulate(mdb_ctx *c, long long rows, int D, const double *d_X, long long ldx, double *d_min,
                                     double *d_max, void *stream) {
    if (!c) return -1;
    if (rows <= 0 || D <= 0) return 0;
    // column bounds of these rows (k_colminmax, kmeans.cu) folded into the caller's running bounds,
    // which start at (+DBL_MAX, -DBL_MAX)
    return minmax_fit_run(c, rows, D, d_X, ldx, d_min, d_max, 1, (cudaStream_t)stream);
}

__global__ void __launch_bounds__(256) k_gather_rows(const double *__restrict__ X, long long ldx, int D,
                                                     const long long *__restrict__ src, const long long *__restrict__ dst,
                                                     long long n, double *__restrict__ out, long long ldo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * D) return;
    const long long r = i / D;
    const int k = (int)(i - r * D);
    out[dst[r] * ldo + k] = X[src[r] * ldx + k];
}

extern "C" int mdb_gather_rows(mdb_ctx *c, long long n, int D, const double *d_X, long long ldx, const long long *d_src,
                               const long long *d_dst, double *d_out, long long ldo, void *stream) {
    if (!c) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (n <= 0 || D <= 0) return 0;
    {
        ProfScope ps(c, "k_gather_rows", s);
        k_gather_rows<<<cdiv(n * D, 256), 256, 0, s>>>(d_X, ldx, D, d_src, d_dst, n, d_out, ldo);
    }
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// selection: member counts per cluster, and the row of the p-th member of a cluster inside a
// row range (ascending row order) -- `idx = np.where(labels == k)[0]; idx[p]` without the where.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_label_hist(const int32_t *__restrict__ labels, long long rows, int K,
                                                    int *__restrict__ hist) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int l = labels[r];
    if (l >= 0 && l < K) atomicAdd(&hist[l], 1);
}

extern "C" int mdb_label_hist(mdb_ctx *c, long long rows, int K, const int32_t *d_labels, int32_t *d_hist, void *stream) {
    if (!c) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (rows <= 0) return 0;
    {
        ProfScope ps(c, "k_label_hist", s);
        k_label_hist<<<cdiv(rows, 256), 256, 0, s>>>(d_labels, rows, K, d_hist);
    }
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// one block per pick: rows [r0, r1) are walked in order in chunks of blockDim * 4 labels; the block
// counts the members of cluster k per chunk and stops in the chunk that holds member number p.
__global__ void __launch_bounds__(256) k_pick_rows(const int32_t *__restrict__ labels, const long long *__restrict__ picks,
                                                   int npicks, long long *__restrict__ out) {
    __shared__ int s_cnt[8];
    __shared__ long long s_found;
    const int q = blockIdx.x;
    if (q >= npicks) return;
    const long long r0 = picks[4 * q], r1 = picks[4 * q + 1];
    const int k = (int)picks[4 * q + 2];
    long long p = picks[4 * q + 3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_found = -1;
    __syncthreads();
    for (long long base = r0; base < r1; base += 1024) {
        // thread t looks at 4 consecutive rows: order inside the chunk = thread-major
        const long long a = base + (long long)threadIdx.x * 4;
        unsigned m = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x)
            if (a + x < r1 && labels[a + x] == k) m |= 1u << x;
        const int mine = __popc(m);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_cnt[w] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) {
            const int v = s_cnt[ww];
            if (ww < w) before += v;
            total += v;
        }
        if (p < total) {
            const long long excl = before + incl - mine;
            if (p >= excl && p < excl + mine) {
                int need = (int)(p - excl);
                for (int x = 0; x < 4; ++x)
                    if (m & (1u << x)) {
                        if (need == 0) {
                            s_found = a + x;
                            break;
                        }
                        --need;
                    }
            }
            __syncthreads();
            break;
        }
        p -= total;
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) out[q] = s_found;
}

extern "C" int mdb_pick_rows(mdb_ctx *c, const int32_t *d_labels, int npicks, const long long *d_picks, long long *d_rows,
                             void *stream) {
    if (!c) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (npicks <= 0) return 0;
    {
        ProfScope ps(c, "k_pick_rows", s);
        k_pick_rows<<<npicks, 256, 0, s>>>(d_labels, d_picks, npicks, d_rows);
    }
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// DFMA peak of this device, measured (bench.py's denominator for the FP64-bound descriptor kernels):
// every thread runs 8 independent fused multiply-add chains in registers.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

extern "C" int mdb_fp64_peak(mdb_ctx *c, double *h_tflops, void *stream) {
    if (!c || !h_tflops) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    cudaDeviceProp prop;
    MDB_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (arena_reserve(c, 4096, s)) return -1;
    arena_reset(c);
    double *d_out = (double *)arena_alloc(c, 256);
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t e0, e1;
    MDB_CUDA(cudaEventCreate(&e0));
    MDB_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        MDB_CUDA(cudaEventRecord(e0, s));
        k_fp64_peak<<<blocks, 256, 0, s>>>(d_out, iters, 0.999999, 1e-7);
        MDB_CUDA(cudaEventRecord(e1, s));
        MDB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MDB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
        if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        c->launches++;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    MDB_CUDA(cudaGetLastError());
    *h_tflops = best;
    return 0;
}
