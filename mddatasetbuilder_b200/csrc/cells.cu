// cells.cu -- K1: periodic cell list (bin -> count -> scan -> scatter) for a batch of frames.
//
// The reference has no spatial index: Open Babel's periodic ConnectTheDots is an O(N^2) pair
// loop (called at detect.py:275) and ASE's get_distances is 1 -> all (datasetbuilder.py:312).
// This builds, per frame, a counting sort of the atoms by cell.  Sorted records hold the atom's
// wrapped position in GRID UNITS (fractional coordinate x cells per axis) as float, so that the
// cell index is exactly (int)u and neighbour deltas are small numbers, plus (type << 27 | index).
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// grid coordinates of one atom; the SAME function is used by the count and scatter passes so the
// (cell, rank) pair computed in the first pass addresses the slot filled in the second.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_coords(const FrameGeom &G, int periodic, double x, double y, double z, float u[3],
                                            int c[3]) {
    x -= G.origin[0];
    y -= G.origin[1];
    z -= G.origin[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double s = x * G.frac[a * 3] + y * G.frac[a * 3 + 1] + z * G.frac[a * 3 + 2];
        if (periodic) {
            s -= floor(s);
            if (!(s < 1.0)) s = 0.0;
        } else {
            s = fmin(fmax(s, 0.0), 1.0);
        }
        float n = (float)G.n[a];
        float v = (float)(s * (double)G.n[a]);
        if (!(v < n)) v = __int_as_float(__float_as_int(n) - 1);  // largest float below n
        if (!(v >= 0.f)) v = 0.f;
        u[a] = v;
        c[a] = (int)v;
    }
}

__global__ void __launch_bounds__(256) k_bin(const double *__restrict__ pos, const FrameGeom *__restrict__ geom, int N,
                                             int periodic, int *__restrict__ cell_count, uint32_t *__restrict__ cr,
                                             int *__restrict__ err) {
    int f = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const FrameGeom &G = geom[f];
    size_t a = (size_t)f * N + i;
    float u[3];
    int c[3];
    grid_coords(G, periodic, pos[a * 3], pos[a * 3 + 1], pos[a * 3 + 2], u, c);
    int cid = (c[2] * G.n[1] + c[1]) * G.n[0] + c[0];
    int rank = atomicAdd(&cell_count[G.cell_base + cid], 1);
    const int rmax = (1 << G.rank_bits) - 1;
    if (rank > rmax) {
        atomicExch(err, MDB_ERR_CELL_OVERFLOW);
        rank = rmax;
    }
    cr[a] = ((uint32_t)cid << G.rank_bits) | (uint32_t)rank;
}

__global__ void __launch_bounds__(256) k_scatter(const double *__restrict__ pos, const uint8_t *__restrict__ types,
                                                 const FrameGeom *__restrict__ geom, int N, int periodic,
                                                 const int *__restrict__ cell_start, const uint32_t *__restrict__ cr,
                                                 float4 *__restrict__ rec, uint32_t *__restrict__ ccell, int rec_mode,
                                                 int nt, int *__restrict__ err) {
    int f = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const FrameGeom &G = geom[f];
    size_t a = (size_t)f * N + i;
    float u[3];
    int c[3];
    grid_coords(G, periodic, pos[a * 3], pos[a * 3 + 1], pos[a * 3 + 2], u, c);
    uint32_t v = cr[a];
    int t = types[i];
    if (t >= nt) {
        atomicExch(err, MDB_ERR_TYPE_RANGE);
        t = 0;
    }
    int slot = cell_start[G.cell_base + (v >> G.rank_bits)] + (int)(v & ((1u << G.rank_bits) - 1u));
    float4 r;
    if (rec_mode == 1) {
        // wrapped Angstrom coordinates (diagonal cell): u / n * L per axis
        r.x = u[0] * G.h[0];
        r.y = u[1] * G.h[4];
        r.z = u[2] * G.h[8];
    } else {
        r.x = u[0];
        r.y = u[1];
        r.z = u[2];
    }
    r.w = __int_as_float((int)(((uint32_t)t << MDB_IDX_BITS) | (uint32_t)i));
    rec[slot] = r;
    ccell[slot] = (uint32_t)c[0] | ((uint32_t)c[1] << 10) | ((uint32_t)c[2] << 20);
}

// ---------------------------------------------------------------------------------------------
// single-pass exclusive scan (decoupled look-back) over int32, in place.
// status word: bits 62..63 = state (0 invalid, 1 aggregate, 2 inclusive prefix), low 32 = value
// ---------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

struct ScanScratch {
    unsigned int tile_counter;
    unsigned int pad[63];
    unsigned long long status[1];
};

size_t scan_scratch_bytes(long long n) {
    long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return sizeof(ScanScratch) + (size_t)(tiles + 1) * sizeof(unsigned long long);
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(int *__restrict__ data, long long n, ScanScratch *sc) {
    __shared__ unsigned int s_tile;
    __shared__ int s_warp[SCAN_THREADS / 32];
    __shared__ int s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(&sc->tile_counter, 1u);
    __syncthreads();
    unsigned int tile = s_tile;
    long long base = (long long)tile * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
    if (base + SCAN_ITEMS <= n) {
        const int4 *p = reinterpret_cast<const int4 *>(data + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            int4 q = p[k];
            v[4 * k] = q.x;
            v[4 * k + 1] = q.y;
            v[4 * k + 2] = q.z;
            v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < n) ? data[base + k] : 0;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) sum += v[k];
    // block-wide exclusive scan of the per-thread sums
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int ws = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += t;
        }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = ws;  // inclusive warp totals
    }
    __syncthreads();
    int block_total = s_warp[SCAN_THREADS / 32 - 1];
    int thread_excl = inc - sum + (w > 0 ? s_warp[w - 1] : 0);
    // publish aggregate, look back for the exclusive prefix of this tile
    if (threadIdx.x == 0) {
        volatile unsigned long long *st = sc->status;
        if (tile == 0) {
            st[0] = (2ull << 62) | (unsigned int)block_total;
            s_prefix = 0;
        } else {
            st[tile] = (1ull << 62) | (unsigned int)block_total;
            __threadfence();
            int run = 0;
            long long j = (long long)tile - 1;
            for (;;) {
                unsigned long long s = st[j];
                unsigned int state = (unsigned int)(s >> 62);
                if (state == 0) continue;
                run += (int)(unsigned int)s;
                if (state == 2) break;
                --j;
            }
            s_prefix = run;
            st[tile] = (2ull << 62) | (unsigned int)(run + block_total);
        }
    }
    __syncthreads();
    int ex = s_prefix + thread_excl;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int t = v[k];
        v[k] = ex;
        ex += t;
    }
    if (base + SCAN_ITEMS <= n) {
        int4 *p = reinterpret_cast<int4 *>(data + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) p[k] = make_int4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < n) data[base + k] = v[k];
    }
}

int device_exclusive_scan(mdb_ctx *ctx, int *d_data, long long n, void *d_scratch, cudaStream_t stream) {
    if (n <= 0) return 0;
    long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    {
        ProfScope ps(ctx, "memset", stream);
        MDB_CUDA(cudaMemsetAsync(d_scratch, 0, scan_scratch_bytes(n), stream));
    }
    {
        ProfScope ps(ctx, "k_scan", stream);
        k_scan<<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(d_data, n, (ScanScratch *)d_scratch);
    }
    ctx->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
size_t cells_workspace_bytes(int nframes, int natoms, long long total_cells) {
    size_t fn = (size_t)nframes * natoms;
    // cell arrays are padded so the scan's int4 path stays aligned
    return align256((size_t)(total_cells + 16) * sizeof(int)) + 2 * align256(fn * sizeof(uint32_t)) +
           align256(fn * sizeof(float4)) + align256(scan_scratch_bytes(total_cells + 1)) + 1024;
}

int cells_build(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos, const uint8_t *d_types,
                const FrameGeom *d_geom, long long total_cells, int periodic, int rec_mode, CellList *out,
                int *d_err, cudaStream_t stream) {
    size_t fn = (size_t)nframes * natoms;
    int *cell_start = (int *)arena_alloc(ctx, (size_t)(total_cells + 16) * sizeof(int));
    uint32_t *cr = (uint32_t *)arena_alloc(ctx, fn * sizeof(uint32_t));
    float4 *rec = (float4 *)arena_alloc(ctx, fn * sizeof(float4));
    uint32_t *ccell = (uint32_t *)arena_alloc(ctx, fn * sizeof(uint32_t));
    void *scan_tmp = arena_alloc(ctx, scan_scratch_bytes(total_cells + 1));
    if (!cell_start || !cr || !rec || !ccell || !scan_tmp) {
        mdb_set_error("cells_build: arena too small (internal sizing error)");
        return -1;
    }
    {
        ProfScope ps(ctx, "memset", stream);
        MDB_CUDA(cudaMemsetAsync(cell_start, 0, (size_t)(total_cells + 16) * sizeof(int), stream));
    }
    dim3 grid(cdiv(natoms, 256), nframes);
    {
        ProfScope ps(ctx, "k_bin", stream);
        k_bin<<<grid, 256, 0, stream>>>(d_pos, d_geom, natoms, periodic, cell_start, cr, d_err);
    }
    ctx->launches++;
    MDB_CUDA(cudaGetLastError());
    // exclusive scan over every cell of the batch (+1 sentinel): frame f's cells start at f*N
    if (device_exclusive_scan(ctx, cell_start, total_cells + 1, scan_tmp, stream)) return -1;
    {
        ProfScope ps(ctx, "k_scatter", stream);
        k_scatter<<<grid, 256, 0, stream>>>(d_pos, d_types, d_geom, natoms, periodic, cell_start, cr, rec, ccell,
                                            rec_mode, ctx->el.nt, d_err);
    }
    ctx->launches++;
    MDB_CUDA(cudaGetLastError());
    out->d_geom = d_geom;
    out->cell_start = cell_start;
    out->rec = rec;
    out->ccell = ccell;
    out->cr = cr;
    out->rec_mode = rec_mode;
    out->total_cells = total_cells;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// bounding boxes (non-periodic frames only): one block per frame
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bbox(const double *__restrict__ pos, int N, double *__restrict__ bbox) {
    int f = blockIdx.x;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        size_t a = ((size_t)f * N + i) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = pos[a + c];
            lo[c] = fmin(lo[c], v);
            hi[c] = fmax(hi[c], v);
        }
    }
    __shared__ double s[6][256];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        s[c][threadIdx.x] = lo[c];
        s[3 + c][threadIdx.x] = hi[c];
    }
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                s[c][threadIdx.x] = fmin(s[c][threadIdx.x], s[c][threadIdx.x + o]);
                s[3 + c][threadIdx.x] = fmax(s[3 + c][threadIdx.x], s[3 + c][threadIdx.x + o]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 6) bbox[f * 6 + threadIdx.x] = s[threadIdx.x][0];
}

int bbox_frames(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos, double *h_bbox, cudaStream_t stream) {
    double *d_bbox = nullptr;
    MDB_CUDA(cudaMalloc(&d_bbox, (size_t)nframes * 6 * sizeof(double)));
    if (natoms > 0) {
        k_bbox<<<nframes, 256, 0, stream>>>(d_pos, natoms, d_bbox);
        ctx->launches++;
    }
    cudaError_t e = cudaMemcpyAsync(h_bbox, d_bbox, (size_t)nframes * 6 * sizeof(double), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_bbox);
    if (e != cudaSuccess) return mdb_cuda_fail(e, "bbox", __FILE__, __LINE__);
    if (natoms == 0)
        for (int i = 0; i < nframes * 6; ++i) h_bbox[i] = 0.0;
    return 0;
}
