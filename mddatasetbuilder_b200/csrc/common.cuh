// common.cuh -- shared definitions of the sm_100a kernels behind include/mdb_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/mdb_b200.h"

#define MDB_IDX_BITS 27
#define MDB_IDX_MASK ((1u << MDB_IDX_BITS) - 1u)
// flag bits carried in ELL slots while the cleanup pass runs (never visible to callers)
#define MDB_TOMB 0x80000000u   // this neighbour entry was deleted
#define MDB_PEND 0x40000000u   // (slot 0) row owner is a pending valence/angle violator
#define MDB_DIRTY 0x20000000u  // (slot 0) row needs compaction
#define MDB_MAXT 16            // max element types per trajectory
#define MDB_MAXZ 18

#define MDB_ERR_CELL_OVERFLOW 1  // more atoms in one cell than the rank field holds (>= 255)
#define MDB_ERR_BOND_OVERFLOW 2  // more raw neighbours than max_bonds
#define MDB_ERR_CSR_OVERFLOW 3
#define MDB_ERR_TYPE_RANGE 4
#define MDB_ERR_EIG_NOCONV 5  // QL iteration did not converge
#define MDB_ERR_KEY_TABLE_FULL 6  // more than 256 distinct bond-type keys in one coded call

struct FrameGeom {
    double ortho[9];    // fractional -> cartesian (lattice vectors as columns), OB's _mOrtho
    double frac[9];     // cartesian -> fractional, adjugate/det like matrix3x3::inverse
    double origin[3];   // non-periodic frames: bounding-box origin (0 when periodic)
    float h[9];         // grid-unit delta -> cartesian: h[r*3+c] = ortho[r][c] / n[c]
    int n[3];           // cells per axis
    int reduce[3];      // 1: axis holds a single cell and deltas are min-image reduced
    long long cell_base;  // first cell of this frame in the batch-wide cell arrays
    int ncell;
    int is_ortho;       // cell matrix is diagonal
    int rank_bits;      // low bits of the per-atom (cell id, rank) word that hold the rank (>= 8)
};

struct ElemTables {
    double cut2[MDB_MAXT * MDB_MAXT];  // (Rcov_i + Rcov_j + 0.45)^2, exact double
    float cut2f[MDB_MAXT * MDB_MAXT];
    int maxbonds[MDB_MAXT];
    int Z[MDB_MAXT];
    double diag[MDB_MAXT];  // Z^2.4 / 2
    int nt;
    float margin;  // half-width (A^2) of the band re-tested in double
    float rc_max;
};

struct BondStats {
    unsigned long long violators, exact_tests, overflow;
    unsigned long long ring_molecules;  // perceive.cu: molecules with a cycle (ring passes not restated)
    int err;
};

struct mdb_ctx {
    int device = 0;
    ElemTables el{};
    bool have_elements = false;
    double opt_cell_edge = 0.0;
    int opt_max_bonds = 8;
    int opt_frames_per_launch = 0;
    int opt_perceive = 0;
    // staging of the fused mini-batch step (page-locked indices + two result scalars; device idx | labels | mindist | out)
    void *mb_hidx = nullptr;
    void *mb_didx = nullptr;
    size_t mb_cap = 0;
    const int *cur_eig_off = nullptr;  // compact-spectra offsets of the descriptor call in flight
    int opt_ql_pwk = 1;      // 1: square-root-free QL (dsterf recurrence) in k_tql, 0: the tqli form
    int opt_desc_legacy = 0;  // 1: round-1 descriptor kernels (one column per lane / warp pairs) for A/B runs  // dump-only keys carry perceived bond orders (perceive.cu)
    uint64_t launches = 0;
    // arena
    char *arena = nullptr;
    size_t arena_bytes = 0;
    size_t arena_off = 0;
    // pinned staging ring for per-call geometry
    static const int kRing = 8;
    FrameGeom *h_geom[kRing] = {};
    size_t h_geom_cap[kRing] = {};
    cudaEvent_t h_geom_ev[kRing] = {};
    int ring_pos = 0;
    // persistent small device state
    BondStats *d_stats = nullptr;
    // streams for the host-buffer pipeline
    cudaStream_t pipe[3] = {};
    cudaEvent_t pipe_ev[3] = {};
    // double-buffered device staging of the host-buffer pipeline (grow-only)
    char *pipe_mem[2] = {};
    size_t pipe_cap[2] = {};
    cudaEvent_t pipe_in[2] = {}, pipe_comp[2] = {}, pipe_out[2] = {};
    uint8_t *pipe_types = nullptr;
    size_t pipe_types_cap = 0;
    int *pipe_hofs[2] = {};  // page-locked per-frame edge offsets of the chunk in flight (edge-list host output)
    size_t pipe_hofs_cap[2] = {};
    // operands / results of the tensor-core assignment (grow-only)
    char *tc_mem = nullptr;
    size_t tc_bytes = 0;
    long long tc_fallbacks = 0;  // calls whose float64 self-check disagreed (whole call redone in float64)
    // optional per-launch CUDA-event timing (bench.py's live roofline measurement)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_free;
    struct ProfRec {
        const char *name;
        cudaEvent_t a, b;
    };
    std::vector<ProfRec> prof_recs;
};

// RAII timing scope around one launch: records events on the launching stream when profiling is on
struct ProfScope {
    mdb_ctx *c;
    cudaStream_t s;
    cudaEvent_t b = nullptr;
    ProfScope(mdb_ctx *ctx, const char *name, cudaStream_t stream) : c(ctx), s(stream) {
        if (!c->prof_on) return;
        cudaEvent_t a;
        auto get = [&](cudaEvent_t &e) {
            if (!c->prof_free.empty()) {
                e = c->prof_free.back();
                c->prof_free.pop_back();
            } else
                cudaEventCreate(&e);
        };
        get(a);
        get(b);
        cudaEventRecord(a, s);
        c->prof_recs.push_back({name, a, b});
    }
    ~ProfScope() {
        if (b) cudaEventRecord(b, s);
    }
};

void mdb_set_error(const std::string &msg);
int mdb_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define MDB_CUDA(x)                                                           \
    do {                                                                      \
        cudaError_t e__ = (x);                                                \
        if (e__ != cudaSuccess) return mdb_cuda_fail(e__, #x, __FILE__, __LINE__); \
    } while (0)

// bump allocation from the context arena (256-byte aligned); returns nullptr when
// the arena is too small -- callers size it first with arena_reserve().
int arena_reserve(mdb_ctx *ctx, size_t bytes, cudaStream_t stream);
void arena_reset(mdb_ctx *ctx);
void *arena_alloc(mdb_ctx *ctx, size_t bytes);
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// host geometry
int build_geometry(mdb_ctx *ctx, int nframes, int natoms, const double *h_cell, int periodic,
                   const double *h_bbox, double edge_min, double edge_hint, std::vector<FrameGeom> &out,
                   long long *total_cells);
int upload_geometry(mdb_ctx *ctx, const std::vector<FrameGeom> &g, FrameGeom *d_geom, cudaStream_t stream);

// cells.cu
struct CellList {
    const FrameGeom *d_geom;
    int *cell_start;   // [total_cells + 1] exclusive scan over the batch (global slot offsets)
    float4 *rec;       // [F*N] sorted records: coords + (type << 27 | index)
    uint32_t *ccell;   // [F*N] sorted: packed cell coordinates cx | cy << 10 | cz << 20
    uint32_t *cr;      // [F*N] per atom (original order): cell id << 8 | rank within the cell
    long long total_cells;
    int rec_mode;      // 1: rec holds wrapped Angstrom coords (all frames orthorhombic, >= 3 cells per
                       //    periodic axis); 0: rec holds grid units (general cells / reduced axes)
};
size_t cells_workspace_bytes(int nframes, int natoms, long long total_cells);
int cells_build(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos, const uint8_t *d_types,
                const FrameGeom *d_geom, long long total_cells, int periodic, int rec_mode, CellList *out,
                int *d_err, cudaStream_t stream);
int device_exclusive_scan(mdb_ctx *ctx, int *d_data, long long n, void *d_scratch, cudaStream_t stream);
size_t scan_scratch_bytes(long long n);
int bbox_frames(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos, double *h_bbox, cudaStream_t stream);

// bonds.cu
int bonds_run(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos, const uint8_t *d_types,
              const CellList &cl, int periodic, int32_t *d_ell, int32_t *d_parent, cudaStream_t stream);
size_t bonds_workspace_bytes(int nframes, int natoms);
int ell_to_csr_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, int order, const double *d_pos,
                   int32_t *d_rowptr, int32_t *d_col, int colcap, cudaStream_t stream);

// descriptor.cu
int compositions_run(mdb_ctx *c, int F, int N, const double *d_pos, const double *h_cell, const uint8_t *d_types,
                     int periodic, double cutoff, const int *d_targets, long long ntargets, int *d_counts,
                     int *h_maxcount, cudaStream_t stream, int memcap = 0, int *d_members_out = nullptr);
int descriptors_run(mdb_ctx *c, int F, int N, const double *d_pos, const double *h_cell, const uint8_t *d_types,
                    int periodic, double cutoff, const int *d_targets, long long ntargets, const int *d_counts,
                    const int *h_maxcount, double *d_X, double *d_eig, int eigcap, int *d_n, cudaStream_t stream,
                    const int *d_members = nullptr, int memcap = 0, int *d_eig_off = nullptr);
int feature_rows_run(mdb_ctx *c, long long T, const double *d_eig, int eigcap, const int *d_n, const int *d_counts,
                     const int *h_maxcount, double *d_X, cudaStream_t stream, const int *d_eig_off = nullptr);

// kmeans.cu
int kmeans_assign_run(mdb_ctx *c, long long rows, int D, int K, const double *d_X, long long ldx, const int *d_rowidx,
                      const double *d_C, int *d_labels, double *d_mind, double *h_inertia, const double *d_w,
                      cudaStream_t stream);
int kmeans_sums_run(mdb_ctx *c, long long B, int D, int K, const double *d_X, long long ldx, const int *d_rowidx,
                    const double *d_w, const int *d_labels, double *d_sums, double *d_wsum, cudaStream_t stream);
int kmeans_apply_run(mdb_ctx *c, int K, int D, double *d_C, double *d_weight_sums, const double *d_sums,
                     const double *d_wsum, cudaStream_t stream);
int minmax_fit_run(mdb_ctx *c, long long rows, int D, const double *d_X, long long ldx, double *d_min, double *d_max,
                   int accumulate, cudaStream_t stream);
int minmax_transform_run(mdb_ctx *c, long long rows, int D, double *d_X, long long ldx, const double *h_min,
                         const double *h_max, cudaStream_t stream);

// graph.cu
int keys_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, const uint8_t *d_types, uint64_t *d_keys,
             int32_t *d_parent, cudaStream_t stream);
int keys_bo_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, const double *d_bo,
                const uint8_t *d_types, uint64_t *d_keys, cudaStream_t stream);
int components_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, int32_t *d_labels, int seeded,
                   cudaStream_t stream);
// keys + union-find hooks in one pass over the ELL rows (d_labels must hold the seed), then flatten
// perceive.cu
size_t perceive_workspace_bytes(int nframes, int natoms, int maxb, bool own_orders);
int perceive_run(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos, const FrameGeom *d_geom, const uint8_t *d_types,
                 int periodic, const int32_t *d_ell, const int32_t *d_labels, uint8_t *d_orders, uint64_t *d_keys,
                 cudaStream_t stream);
int keys_components_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, const uint8_t *d_types,
                        uint64_t *d_keys, int32_t *d_labels, cudaStream_t stream);
int dps_run(mdb_ctx *ctx, int natoms, const int32_t *d_rowptr, const int32_t *d_col, int32_t *d_mol_atoms,
            int32_t *d_mol_ptr, int *h_nmol, cudaStream_t stream);

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }
