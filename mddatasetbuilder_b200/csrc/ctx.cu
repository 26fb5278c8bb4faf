// ctx.cu -- context, scratch arena, element tables and per-frame geometry (host side).
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

static thread_local std::string g_err;

void mdb_set_error(const std::string &msg) { g_err = msg; }

int mdb_cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    g_err = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ") at " + file + ":" +
            std::to_string(line) + " in " + what;
    return (int)e ? (int)e : -1;
}

// Open Babel 3.1.x element table (covalent radius, Cordero 2008; MaxBonds), Z = 0..18.
// The bond rule of ConnectTheDots is d2 <= (Rcov_i + Rcov_j + 0.45)^2.
static const double kRcov[MDB_MAXZ + 1] = {0.0,  0.31, 0.28, 1.28, 0.96, 0.84, 0.76, 0.71, 0.66, 0.57,
                                           0.58, 1.66, 1.41, 1.21, 1.11, 1.07, 1.05, 1.02, 1.06};
static const int kMaxBonds[MDB_MAXZ + 1] = {0, 1, 0, 1, 2, 4, 4, 4, 2, 1, 0, 1, 2, 6, 6, 6, 6, 1, 0};

extern "C" int mdb_version(void) { return 100; }
extern "C" const char *mdb_last_error(void) { return g_err.c_str(); }

extern "C" int mdb_create(int device, mdb_ctx **out) {
    if (!out) {
        mdb_set_error("mdb_create: out is NULL");
        return -1;
    }
    int ndev = 0;
    MDB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        mdb_set_error("mdb_create: no such CUDA device " + std::to_string(device));
        return -1;
    }
    MDB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MDB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        mdb_set_error(std::string("mdb_create: this library is built for sm_100a only; device is ") + prop.name +
                      " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
        return -1;
    }
    mdb_ctx *c = new mdb_ctx();
    c->device = device;
    MDB_CUDA(cudaMalloc(&c->d_stats, sizeof(BondStats)));
    MDB_CUDA(cudaMemset(c->d_stats, 0, sizeof(BondStats)));
    for (int i = 0; i < 3; ++i) {
        MDB_CUDA(cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking));
        MDB_CUDA(cudaEventCreateWithFlags(&c->pipe_ev[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < mdb_ctx::kRing; ++i) MDB_CUDA(cudaEventCreateWithFlags(&c->h_geom_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        MDB_CUDA(cudaEventCreateWithFlags(&c->pipe_in[i], cudaEventDisableTiming));
        MDB_CUDA(cudaEventCreateWithFlags(&c->pipe_comp[i], cudaEventDisableTiming));
        MDB_CUDA(cudaEventCreateWithFlags(&c->pipe_out[i], cudaEventDisableTiming));
    }
    *out = c;
    return 0;
}

extern "C" void mdb_destroy(mdb_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->arena) cudaFree(c->arena);
    if (c->d_stats) cudaFree(c->d_stats);
    for (int i = 0; i < 3; ++i) {
        if (c->pipe[i]) cudaStreamDestroy(c->pipe[i]);
        if (c->pipe_ev[i]) cudaEventDestroy(c->pipe_ev[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (c->pipe_mem[i]) cudaFree(c->pipe_mem[i]);
        if (c->pipe_hofs[i]) cudaFreeHost(c->pipe_hofs[i]);
        if (c->pipe_in[i]) cudaEventDestroy(c->pipe_in[i]);
        if (c->pipe_comp[i]) cudaEventDestroy(c->pipe_comp[i]);
        if (c->pipe_out[i]) cudaEventDestroy(c->pipe_out[i]);
    }
    if (c->pipe_types) cudaFree(c->pipe_types);
    if (c->tc_mem) cudaFree(c->tc_mem);
    if (c->mb_hidx) cudaFreeHost(c->mb_hidx);
    if (c->mb_didx) cudaFree(c->mb_didx);
    for (int i = 0; i < mdb_ctx::kRing; ++i) {
        if (c->h_geom[i]) cudaFreeHost(c->h_geom[i]);
        if (c->h_geom_ev[i]) cudaEventDestroy(c->h_geom_ev[i]);
    }
    delete c;
}

extern "C" int mdb_set_elements(mdb_ctx *c, int ntypes, const int *Z) {
    if (!c || ntypes < 1 || ntypes > MDB_MAXT) {
        mdb_set_error("mdb_set_elements: between 1 and " + std::to_string(MDB_MAXT) + " element types are supported");
        return -1;
    }
    ElemTables &t = c->el;
    memset(&t, 0, sizeof(t));
    t.nt = ntypes;
    double rmax = 0;
    for (int i = 0; i < ntypes; ++i) {
        if (Z[i] < 1 || Z[i] > MDB_MAXZ) {
            mdb_set_error("mdb_set_elements: atomic number " + std::to_string(Z[i]) + " outside the built-in table (1..18)");
            return -1;
        }
        t.Z[i] = Z[i];
        t.maxbonds[i] = kMaxBonds[Z[i]];
        t.diag[i] = pow((double)Z[i], 2.4) / 2;
        rmax = std::max(rmax, kRcov[Z[i]]);
    }
    for (int i = 0; i < MDB_MAXT; ++i)
        for (int j = 0; j < MDB_MAXT; ++j) {
            double cut = 0;
            if (i < ntypes && j < ntypes) {
                cut = kRcov[Z[i]] + kRcov[Z[j]] + 0.45;
                cut = cut * cut;
            }
            t.cut2[i * MDB_MAXT + j] = cut;
            t.cut2f[i * MDB_MAXT + j] = (float)cut;
        }
    t.rc_max = (float)(2 * rmax + 0.45);
    c->have_elements = true;
    return 0;
}

extern "C" int mdb_set_option(mdb_ctx *c, const char *key, double v) {
    std::string k(key ? key : "");
    if (k == "cell_edge")
        c->opt_cell_edge = v;
    else if (k == "max_bonds") {
        if ((int)v != 8 && (int)v != 16) {
            mdb_set_error("max_bonds must be 8 or 16");
            return -1;
        }
        c->opt_max_bonds = (int)v;
    } else if (k == "frames_per_launch")
        c->opt_frames_per_launch = (int)v;
    else if (k == "perceive_bond_orders")
        c->opt_perceive = v != 0.0;
    else if (k == "ql_pwk")
        c->opt_ql_pwk = v != 0.0;
    else if (k == "desc_legacy")
        c->opt_desc_legacy = (int)v;  // 0 default mix, 1 round-1 kernels, 2 multi-column kernels everywhere
    else {
        mdb_set_error("unknown option " + k);
        return -1;
    }
    return 0;
}

extern "C" uint64_t mdb_launch_count(const mdb_ctx *c) { return c ? c->launches : 0; }
extern "C" size_t mdb_arena_bytes(const mdb_ctx *c) { return c ? c->arena_bytes : 0; }

int arena_reserve(mdb_ctx *c, size_t bytes, cudaStream_t stream) {
    if (bytes <= c->arena_bytes) return 0;
    // growth happens on the first call of a new batch shape only
    MDB_CUDA(cudaDeviceSynchronize());
    if (c->arena) MDB_CUDA(cudaFree(c->arena));
    c->arena = nullptr;
    c->arena_bytes = 0;
    size_t want = bytes + bytes / 8 + (1 << 20);
    MDB_CUDA(cudaMalloc(&c->arena, want));
    c->arena_bytes = want;
    (void)stream;
    return 0;
}
void arena_reset(mdb_ctx *c) { c->arena_off = 0; }
void *arena_alloc(mdb_ctx *c, size_t bytes) {
    size_t off = align256(c->arena_off);
    if (off + bytes > c->arena_bytes) return nullptr;
    c->arena_off = off + bytes;
    return c->arena + off;
}

// cart->frac matrix exactly as Open Babel builds it (matrix3x3::inverse = adjugate / det);
// the oracle (oracle.c: orc_ob_cell_setup) performs the same operations in the same order.
static void ob_cell_setup(const double *cell, double *ortho9, double *frac9) {
    double e[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) e[r][c] = cell[c * 3 + r];
    double x = e[0][0] * (e[1][1] * e[2][2] - e[1][2] * e[2][1]);
    double y = e[0][1] * (e[1][2] * e[2][0] - e[1][0] * e[2][2]);
    double z = e[0][2] * (e[1][0] * e[2][1] - e[1][1] * e[2][0]);
    double det = x + y + z;
    double inv[3][3];
    inv[0][0] = e[1][1] * e[2][2] - e[1][2] * e[2][1];
    inv[1][0] = e[1][2] * e[2][0] - e[1][0] * e[2][2];
    inv[2][0] = e[1][0] * e[2][1] - e[1][1] * e[2][0];
    inv[0][1] = e[2][1] * e[0][2] - e[2][2] * e[0][1];
    inv[1][1] = e[2][2] * e[0][0] - e[2][0] * e[0][2];
    inv[2][1] = e[2][0] * e[0][1] - e[2][1] * e[0][0];
    inv[0][2] = e[0][1] * e[1][2] - e[0][2] * e[1][1];
    inv[1][2] = e[0][2] * e[1][0] - e[0][0] * e[1][2];
    inv[2][2] = e[0][0] * e[1][1] - e[0][1] * e[1][0];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            ortho9[r * 3 + c] = e[r][c];
            frac9[r * 3 + c] = inv[r][c] / det;
        }
}

int build_geometry(mdb_ctx *ctx, int nframes, int natoms, const double *h_cell, int periodic, const double *h_bbox,
                   double edge_min, double edge_hint, std::vector<FrameGeom> &out, long long *total_cells) {
    out.resize(nframes);
    long long base = 0;
    for (int f = 0; f < nframes; ++f) {
        FrameGeom &g = out[f];
        memset(&g, 0, sizeof(g));
        double cell[9];
        if (periodic) {
            memcpy(cell, h_cell + (size_t)f * 9, sizeof(cell));
        } else {
            // pseudo-cell = bounding box (binning only; distances are plain differences)
            memset(cell, 0, sizeof(cell));
            for (int a = 0; a < 3; ++a) {
                double lo = h_bbox[f * 6 + a], hi = h_bbox[f * 6 + 3 + a];
                // at least 1 A thick, so that a single atom or a flat molecule still spans a proper box
                double span = std::max((hi - lo) * (1.0 + 1e-9) + 1e-6, 1.0);
                cell[a * 3 + a] = span;
                g.origin[a] = lo;
            }
        }
        ob_cell_setup(cell, g.ortho, g.frac);
        double det = g.ortho[0] * (g.ortho[4] * g.ortho[8] - g.ortho[5] * g.ortho[7]) -
                     g.ortho[1] * (g.ortho[3] * g.ortho[8] - g.ortho[5] * g.ortho[6]) +
                     g.ortho[2] * (g.ortho[3] * g.ortho[7] - g.ortho[4] * g.ortho[6]);
        double vol = fabs(det);
        if (!(vol > 1e-12) || !std::isfinite(vol)) {
            mdb_set_error("frame " + std::to_string(f) + ": singular cell (volume " + std::to_string(vol) + ")");
            return -1;
        }
        g.is_ortho = 1;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                if (r != c && g.ortho[r * 3 + c] != 0.0) g.is_ortho = 0;
        double height[3];
        for (int a = 0; a < 3; ++a) {
            const double *fr = g.frac + a * 3;
            height[a] = 1.0 / sqrt(fr[0] * fr[0] + fr[1] * fr[1] + fr[2] * fr[2]);
        }
        double edge = edge_hint;
        if (!(edge > 0)) {
            double density = natoms / vol;
            edge = cbrt(0.5 / std::max(density, 1e-9));
        }
        edge = std::max(edge, edge_min);
        for (;;) {
            long long prod = 1;
            for (int a = 0; a < 3; ++a) {
                long long n = (long long)floor(height[a] / edge);
                if (periodic) {
                    if (n < 3) {
                        n = 1;
                        g.reduce[a] = 1;
                    } else
                        g.reduce[a] = 0;
                } else {
                    if (n < 1) n = 1;
                    g.reduce[a] = 0;
                }
                if (n > 1023) n = 1023;  // 10 bits per axis keeps ids well inside 24 bits... checked below
                g.n[a] = (int)n;
                prod *= n;
            }
            if (prod <= (1LL << 24)) {
                g.ncell = (int)prod;
                break;
            }
            edge *= 1.26;
        }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) g.h[r * 3 + c] = (float)(g.ortho[r * 3 + c] / g.n[c]);
        g.cell_base = base;
        base += g.ncell;
        // (cell id, rank within the cell) share 32 bits per atom: frames with few cells (small boxes
        // under a large cutoff) get the spare bits as rank, i.e. room for more atoms per cell
        g.rank_bits = 8;
        while (g.rank_bits < 20 && ((long long)g.ncell << (g.rank_bits + 1)) <= (1LL << 31)) ++g.rank_bits;
    }
    *total_cells = base;
    (void)ctx;
    return 0;
}

__global__ void k_copy_words(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t nwords) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int upload_geometry(mdb_ctx *c, const std::vector<FrameGeom> &g, FrameGeom *d_geom, cudaStream_t stream) {
    int slot = c->ring_pos;
    c->ring_pos = (c->ring_pos + 1) % mdb_ctx::kRing;
    size_t bytes = g.size() * sizeof(FrameGeom);
    MDB_CUDA(cudaEventSynchronize(c->h_geom_ev[slot]));
    if (c->h_geom_cap[slot] < bytes) {
        if (c->h_geom[slot]) MDB_CUDA(cudaFreeHost(c->h_geom[slot]));
        c->h_geom[slot] = nullptr;
        MDB_CUDA(cudaMallocHost(&c->h_geom[slot], bytes + 4096));
        c->h_geom_cap[slot] = bytes + 4096;
    }
    memcpy(c->h_geom[slot], g.data(), bytes);
    // a kernel reads the page-locked staging copy (unified addressing) instead of a DMA copy: in the
    // host-buffer pipeline a small copy on the compute stream would queue behind the next chunk's 100 MB
    // input copy in the copy engine and hold this chunk's kernels back by that long
    k_copy_words<<<(unsigned)std::min<size_t>(cdiv((long long)(bytes / 4), 256), 128), 256, 0, stream>>>(
        reinterpret_cast<const uint32_t *>(c->h_geom[slot]), reinterpret_cast<uint32_t *>(d_geom), bytes / 4);
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    MDB_CUDA(cudaEventRecord(c->h_geom_ev[slot], stream));
    return 0;
}
