// api.cu -- extern "C" entry points declared in include/mdb_b200.h.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

static int check_ctx(mdb_ctx *c, const char *fn, bool need_elements = true) {
    if (!c) {
        mdb_set_error(std::string(fn) + ": context is NULL");
        return -1;
    }
    if (need_elements && !c->have_elements) {
        mdb_set_error(std::string(fn) + ": call mdb_set_elements first");
        return -1;
    }
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return mdb_cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
    return 0;
}

static const char *err_text(int e) {
    switch (e) {
        case MDB_ERR_CELL_OVERFLOW:
            return "more atoms fell into one cell than its rank field can index (at least 255; overlapping atoms?)";
        case MDB_ERR_BOND_OVERFLOW:
            return "an atom has more raw neighbours than max_bonds; set option max_bonds=16";
        case MDB_ERR_CSR_OVERFLOW:
            return "CSR column capacity exceeded";
        case MDB_ERR_TYPE_RANGE:
            return "atom type index outside the element table";
        case MDB_ERR_EIG_NOCONV:
            return "QL eigenvalue iteration did not converge";
        case MDB_ERR_KEY_TABLE_FULL:
            return "more than 256 distinct bond-type keys: use the 64-bit key output";
        default:
            return "unknown device-side error";
    }
}

static int frames_per_chunk(const mdb_ctx *c, int nframes, int natoms) {
    if (c->opt_frames_per_launch > 0) return std::min(nframes, c->opt_frames_per_launch);
    // ~16M atom-frames per launch: large enough that launch gaps and the per-frame cleanup blocks
    // amortise (measured on B200: 5 -> 160 frames of 100k atoms = 43.9 -> 13.3 ms per 1e8 atom-frames)
    long long per = std::max(1LL, (16LL << 20) / std::max(natoms, 1));
    per = std::min<long long>(per, ((1LL << 30) / std::max(natoms, 1)));
    return (int)std::max(1LL, std::min<long long>(per, nframes));
}

// one chunk of frames through cell list -> bonds -> cleanup [-> keys -> components]
static int analyze_chunk(mdb_ctx *c, int F, int N, const double *d_pos, const double *h_cell, const uint8_t *d_types,
                         int periodic, int32_t *d_ell, uint64_t *d_keys, int32_t *d_labels, cudaStream_t stream) {
    std::vector<double> bbox;
    if (!periodic) {
        bbox.resize((size_t)F * 6);
        if (bbox_frames(c, F, N, d_pos, bbox.data(), stream)) return -1;
    } else if (!h_cell) {
        mdb_set_error("periodic frames need a cell");
        return -1;
    }
    std::vector<FrameGeom> geom;
    long long total_cells = 0;
    const double edge_min = (double)c->el.rc_max * 1.0005;
    if (build_geometry(c, F, N, h_cell, periodic, bbox.data(), edge_min, c->opt_cell_edge, geom, &total_cells)) return -1;
    double lmax = 0;
    bool all_ortho = true;
    for (const FrameGeom &g : geom) {
        all_ortho = all_ortho && g.is_ortho && !g.reduce[0] && !g.reduce[1] && !g.reduce[2];
        for (int v = 0; v < 3; ++v) {
            double l = sqrt(g.ortho[v] * g.ortho[v] + g.ortho[3 + v] * g.ortho[3 + v] + g.ortho[6 + v] * g.ortho[6 + v]);
            lmax = std::max(lmax, l);
        }
    }
    // float screening error: |delta d2| <~ 8 * rc * 2^-24 * L * sqrt(3); x1.5 safety
    c->el.margin = (float)(1.25e-6 * c->el.rc_max * lmax + 1e-5);

    const int maxb = c->opt_max_bonds;
    size_t fn = (size_t)F * N;
    size_t need = align256(geom.size() * sizeof(FrameGeom)) + cells_workspace_bytes(F, N, total_cells) +
                  bonds_workspace_bytes(F, N) + (d_ell ? 0 : align256(fn * maxb * sizeof(int32_t))) +
                  (d_labels ? 0 : align256(fn * sizeof(int32_t))) + 4096;
    const bool perceive = c->opt_perceive && d_keys;
    if (perceive) need += perceive_workspace_bytes(F, N, maxb, true);
    if (arena_reserve(c, need, stream)) return -1;
    arena_reset(c);
    FrameGeom *d_geom = (FrameGeom *)arena_alloc(c, geom.size() * sizeof(FrameGeom));
    if (!d_ell) d_ell = (int32_t *)arena_alloc(c, fn * maxb * sizeof(int32_t));
    if (!d_geom || !d_ell) {
        mdb_set_error("analyze_chunk: arena too small (internal sizing error)");
        return -1;
    }
    if (upload_geometry(c, geom, d_geom, stream)) return -1;
    CellList cl;
    if (cells_build(c, F, N, d_pos, d_types, d_geom, total_cells, periodic, all_ortho ? 1 : 0, &cl, &c->d_stats->err,
                    stream))
        return -1;
    if (bonds_run(c, F, N, d_pos, d_types, cl, periodic, d_ell, d_labels, stream)) return -1;
    if (d_labels) {
        if (keys_components_run(c, F, N, d_ell, d_types, perceive ? nullptr : d_keys, d_labels, stream)) return -1;
    } else if (d_keys && !perceive) {
        if (keys_run(c, F, N, d_ell, d_types, d_keys, nullptr, stream)) return -1;
    }
    // dump-only keys with perceived bond orders (reference detect.py:279-280)
    if (perceive && perceive_run(c, F, N, d_pos, d_geom, d_types, periodic, d_ell, d_labels, nullptr, d_keys, stream)) return -1;
    return 0;
}

static int analyze_frames(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                          const uint8_t *d_types, int periodic, int32_t *d_ell, uint64_t *d_keys, int32_t *d_labels,
                          cudaStream_t stream) {
    if (nframes < 0 || natoms < 0 || (long long)natoms >= (1LL << MDB_IDX_BITS)) {
        mdb_set_error("bad batch shape (natoms must be < 2^27)");
        return -1;
    }
    MDB_CUDA(cudaMemsetAsync(c->d_stats, 0, sizeof(BondStats), stream));
    if (nframes == 0 || natoms == 0) return 0;
    const int per = frames_per_chunk(c, nframes, natoms);
    const int maxb = c->opt_max_bonds;
    for (int f0 = 0; f0 < nframes; f0 += per) {
        int F = std::min(per, nframes - f0);
        size_t o = (size_t)f0 * natoms;
        if (analyze_chunk(c, F, natoms, d_pos + o * 3, h_cell ? h_cell + (size_t)f0 * 9 : nullptr, d_types, periodic,
                          d_ell ? d_ell + o * maxb : nullptr, d_keys ? d_keys + o : nullptr,
                          d_labels ? d_labels + o : nullptr, stream))
            return -1;
    }
    return 0;
}

extern "C" int mdb_bonds_from_coords(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                                     const uint8_t *d_types, int periodic, int32_t *d_ell, void *stream) {
    if (check_ctx(c, "mdb_bonds_from_coords")) return -1;
    if (!d_ell) {
        mdb_set_error("mdb_bonds_from_coords: d_ell is NULL");
        return -1;
    }
    return analyze_frames(c, nframes, natoms, d_pos, h_cell, d_types, periodic, d_ell, nullptr, nullptr,
                          (cudaStream_t)stream);
}

extern "C" int mdb_analyze_frames(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                                  const uint8_t *d_types, int periodic, int32_t *d_ell, uint64_t *d_keys,
                                  int32_t *d_labels, void *stream) {
    if (check_ctx(c, "mdb_analyze_frames")) return -1;
    return analyze_frames(c, nframes, natoms, d_pos, h_cell, d_types, periodic, d_ell, d_keys, d_labels,
                          (cudaStream_t)stream);
}

extern "C" int mdb_ell_to_csr(mdb_ctx *c, int nframes, int natoms, const int32_t *d_ell, int order, const double *d_pos,
                              int32_t *d_rowptr, int32_t *d_col, int colcap, void *stream) {
    if (check_ctx(c, "mdb_ell_to_csr", false)) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if ((long long)nframes * natoms >= (1LL << 31) - 64) {
        mdb_set_error("mdb_ell_to_csr: batch too large, split the frames");
        return -1;
    }
    if (ell_to_csr_run(c, nframes, natoms, d_ell, order, d_pos, d_rowptr, d_col, colcap, s)) return -1;
    int err = 0;
    MDB_CUDA(cudaMemcpyAsync(&err, &c->d_stats->err, sizeof(int), cudaMemcpyDeviceToHost, s));
    MDB_CUDA(cudaStreamSynchronize(s));
    if (err) {
        mdb_set_error(std::string("mdb_ell_to_csr: ") + err_text(err));
        return err;
    }
    return 0;
}

extern "C" int mdb_bond_keys(mdb_ctx *c, int nframes, int natoms, const int32_t *d_ell, const uint8_t *d_types,
                             uint64_t *d_keys, int32_t *d_parent, void *stream) {
    if (check_ctx(c, "mdb_bond_keys", false)) return -1;
    return keys_run(c, nframes, natoms, d_ell, d_types, d_keys, d_parent, (cudaStream_t)stream);
}

extern "C" int mdb_bond_keys_bo(mdb_ctx *c, int nframes, int natoms, const int32_t *d_ell, const double *d_bo,
                                const uint8_t *d_types, uint64_t *d_keys, void *stream) {
    if (check_ctx(c, "mdb_bond_keys_bo", false)) return -1;
    return keys_bo_run(c, nframes, natoms, d_ell, d_bo, d_types, d_keys, (cudaStream_t)stream);
}

extern "C" int mdb_components(mdb_ctx *c, int nframes, int natoms, const int32_t *d_ell, int32_t *d_labels, int seeded,
                              void *stream) {
    if (check_ctx(c, "mdb_components", false)) return -1;
    return components_run(c, nframes, natoms, d_ell, d_labels, seeded, (cudaStream_t)stream);
}

extern "C" int mdb_dps(mdb_ctx *c, int natoms, const int32_t *d_rowptr, const int32_t *d_col, int32_t *d_mol_atoms,
                       int32_t *d_mol_ptr, int *h_nmol, void *stream) {
    if (check_ctx(c, "mdb_dps", false)) return -1;
    return dps_run(c, natoms, d_rowptr, d_col, d_mol_atoms, d_mol_ptr, h_nmol, (cudaStream_t)stream);
}

extern "C" int mdb_bond_stats(mdb_ctx *c, long long *n_violators, long long *n_exact, long long *n_overflow) {
    if (check_ctx(c, "mdb_bond_stats", false)) return -1;
    MDB_CUDA(cudaDeviceSynchronize());
    BondStats h;
    MDB_CUDA(cudaMemcpy(&h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    if (n_violators) *n_violators = (long long)h.violators;
    if (n_exact) *n_exact = (long long)h.exact_tests;
    if (n_overflow) *n_overflow = (long long)h.overflow;
    if (h.err) {
        // read and clear: the flag is reported once (a caller may retry with other capacities)
        MDB_CUDA(cudaMemset(&c->d_stats->err, 0, sizeof(int)));
        mdb_set_error(std::string("device-side error: ") + err_text(h.err));
        return h.err;
    }
    return 0;
}

extern "C" int mdb_perceive_bond_orders(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                                        const uint8_t *d_types, int periodic, const int32_t *d_ell,
                                        const int32_t *d_labels, uint8_t *d_orders, uint64_t *d_keys, void *stream) {
    if (check_ctx(c, "mdb_perceive_bond_orders")) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (!d_pos || !d_ell || !d_types || (!d_orders && !d_keys)) {
        mdb_set_error("mdb_perceive_bond_orders: positions, types, ELL rows and one output are required");
        return -1;
    }
    if (nframes <= 0 || natoms <= 0) return 0;
    std::vector<FrameGeom> geom((size_t)nframes);
    if (periodic) {
        if (!h_cell) {
            mdb_set_error("periodic frames need a cell");
            return -1;
        }
        long long total_cells = 0;
        if (build_geometry(c, nframes, natoms, h_cell, 1, nullptr, (double)c->el.rc_max * 1.0005, c->opt_cell_edge, geom,
                           &total_cells))
            return -1;
    } else {
        memset(geom.data(), 0, geom.size() * sizeof(FrameGeom));
    }
    const int maxb = c->opt_max_bonds;
    size_t need = align256(geom.size() * sizeof(FrameGeom)) + perceive_workspace_bytes(nframes, natoms, maxb, !d_orders) + 1024;
    if (arena_reserve(c, need, s)) return -1;
    arena_reset(c);
    FrameGeom *d_geom = (FrameGeom *)arena_alloc(c, geom.size() * sizeof(FrameGeom));
    if (upload_geometry(c, geom, d_geom, s)) return -1;
    MDB_CUDA(cudaMemsetAsync(&c->d_stats->ring_molecules, 0, sizeof(unsigned long long), s));
    return perceive_run(c, nframes, natoms, d_pos, d_geom, d_types, periodic, d_ell, d_labels, d_orders, d_keys, s);
}

extern "C" int mdb_perceive_stats(mdb_ctx *c, long long *n_rings) {
    if (check_ctx(c, "mdb_perceive_stats", false)) return -1;
    MDB_CUDA(cudaDeviceSynchronize());
    unsigned long long v = 0;
    MDB_CUDA(cudaMemcpy(&v, &c->d_stats->ring_molecules, sizeof(v), cudaMemcpyDeviceToHost));
    if (n_rings) *n_rings = (long long)v / 2;  // the kernel sums deg - 2 + 2 [root] = 2 (bonds - atoms + molecules)
    return 0;
}

extern "C" int mdb_compositions(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                                const uint8_t *d_types, int periodic, double cutoff, const int32_t *d_targets,
                                long long ntargets, int32_t *d_counts, int32_t *h_maxcount, void *stream) {
    if (check_ctx(c, "mdb_compositions")) return -1;
    if ((long long)nframes * natoms >= (1LL << 31) - 64) {
        mdb_set_error("mdb_compositions: batch too large, split the frames");
        return -1;
    }
    return compositions_run(c, nframes, natoms, d_pos, h_cell, d_types, periodic, cutoff, d_targets, ntargets, d_counts,
                            h_maxcount, (cudaStream_t)stream);
}

int members_run(mdb_ctx *c, int F, int N, const double *d_pos, const double *h_cell, const uint8_t *d_types,
                int periodic, double cutoff, const int *d_targets, long long ntargets, int cap, int *d_members,
                int *d_nmem, cudaStream_t stream);

extern "C" int mdb_sphere_members(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                                  const uint8_t *d_types, int periodic, double cutoff, const int32_t *d_targets,
                                  long long ntargets, int cap, int32_t *d_members, int32_t *d_nmem, void *stream) {
    if (check_ctx(c, "mdb_sphere_members")) return -1;
    return members_run(c, nframes, natoms, d_pos, h_cell, d_types, periodic, cutoff, d_targets, ntargets, cap, d_members,
                       d_nmem, (cudaStream_t)stream);
}

extern "C" int mdb_descriptors(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                               const uint8_t *d_types, int periodic, double cutoff, const int32_t *d_targets,
                               long long ntargets, const int32_t *d_counts, const int32_t *h_maxcount, double *d_X,
                               double *d_eig, int eigcap, int32_t *d_n, void *stream) {
    if (check_ctx(c, "mdb_descriptors")) return -1;
    if ((long long)nframes * natoms >= (1LL << 31) - 64) {
        mdb_set_error("mdb_descriptors: batch too large, split the frames");
        return -1;
    }
    return descriptors_run(c, nframes, natoms, d_pos, h_cell, d_types, periodic, cutoff, d_targets, ntargets, d_counts,
                           h_maxcount, d_X, d_eig, eigcap, d_n, (cudaStream_t)stream);
}

extern "C" int mdb_compositions_members(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                                        const uint8_t *d_types, int periodic, double cutoff, const int32_t *d_targets,
                                        long long ntargets, int32_t *d_counts, int32_t *h_maxcount, int cap,
                                        int32_t *d_members, void *stream) {
    if (check_ctx(c, "mdb_compositions_members")) return -1;
    if ((long long)nframes * natoms >= (1LL << 31) - 64) {
        mdb_set_error("mdb_compositions_members: batch too large, split the frames");
        return -1;
    }
    if (!d_members) {
        mdb_set_error("mdb_compositions_members: d_members is required");
        return -1;
    }
    return compositions_run(c, nframes, natoms, d_pos, h_cell, d_types, periodic, cutoff, d_targets, ntargets, d_counts,
                            h_maxcount, (cudaStream_t)stream, cap, d_members);
}

extern "C" int mdb_descriptors_members(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                                       const uint8_t *d_types, int periodic, const int32_t *d_targets, long long ntargets,
                                       const int32_t *d_counts, const int32_t *h_maxcount, const int32_t *d_members, int cap,
                                       double *d_X, double *d_eig, int eigcap, int32_t *d_n, void *stream) {
    if (check_ctx(c, "mdb_descriptors_members")) return -1;
    if ((long long)nframes * natoms >= (1LL << 31) - 64) {
        mdb_set_error("mdb_descriptors_members: batch too large, split the frames");
        return -1;
    }
    if (!d_members || cap < 1) {
        mdb_set_error("mdb_descriptors_members: d_members / cap are required");
        return -1;
    }
    // the cutoff only sizes the (unused) cell geometry here: membership comes from the lists
    return descriptors_run(c, nframes, natoms, d_pos, h_cell, d_types, periodic, 5.0, d_targets, ntargets, d_counts,
                           h_maxcount, d_X, d_eig, eigcap, d_n, (cudaStream_t)stream, d_members, cap);
}

extern "C" int mdb_descriptors_members_compact(mdb_ctx *c, int nframes, int natoms, const double *d_pos, const double *h_cell,
                                               const uint8_t *d_types, int periodic, const int32_t *d_targets,
                                               long long ntargets, const int32_t *d_counts, const int32_t *h_maxcount,
                                               const int32_t *d_members, int cap, double *d_X, double *d_eig_flat,
                                               int32_t *d_eig_off, int32_t *d_n, void *stream) {
    if (check_ctx(c, "mdb_descriptors_members_compact")) return -1;
    if ((long long)nframes * natoms >= (1LL << 31) - 64) {
        mdb_set_error("mdb_descriptors_members_compact: batch too large, split the frames");
        return -1;
    }
    if (!d_members || cap < 1 || !d_eig_flat || !d_eig_off) {
        mdb_set_error("mdb_descriptors_members_compact: d_members / cap / d_eig_flat / d_eig_off are required");
        return -1;
    }
    return descriptors_run(c, nframes, natoms, d_pos, h_cell, d_types, periodic, 5.0, d_targets, ntargets, d_counts,
                           h_maxcount, d_X, d_eig_flat, 0, d_n, (cudaStream_t)stream, d_members, cap, d_eig_off);
}

extern "C" int mdb_feature_rows_compact(mdb_ctx *c, long long ntargets, const double *d_eig_flat, const int32_t *d_eig_off,
                                        const int32_t *d_counts, const int32_t *h_maxcount, double *d_X, void *stream) {
    if (check_ctx(c, "mdb_feature_rows_compact")) return -1;
    if (!d_eig_off) {
        mdb_set_error("mdb_feature_rows_compact: d_eig_off is required");
        return -1;
    }
    return feature_rows_run(c, ntargets, d_eig_flat, 0, nullptr, d_counts, h_maxcount, d_X, (cudaStream_t)stream, d_eig_off);
}

extern "C" int mdb_feature_rows(mdb_ctx *c, long long ntargets, const double *d_eig, int eigcap, const int32_t *d_n,
                                const int32_t *d_counts, const int32_t *h_maxcount, double *d_X, void *stream) {
    if (check_ctx(c, "mdb_feature_rows")) return -1;
    return feature_rows_run(c, ntargets, d_eig, eigcap, d_n, d_counts, h_maxcount, d_X, (cudaStream_t)stream);
}

extern "C" int mdb_kmeans_assign(mdb_ctx *c, long long rows, int D, int K, const double *d_X, long long ldx,
                                 const int32_t *d_rowidx, const double *d_centers, int32_t *d_labels, double *d_mindist,
                                 const double *d_weights, double *h_inertia, void *stream) {
    if (check_ctx(c, "mdb_kmeans_assign", false)) return -1;
    return kmeans_assign_run(c, rows, D, K, d_X, ldx, d_rowidx, d_centers, d_labels, d_mindist, h_inertia, d_weights,
                             (cudaStream_t)stream);
}

int kmeans_assign_tc_run(mdb_ctx *c, long long rows, int D, int K, const double *d_X, long long ldx, const int *d_rowidx,
                         const double *d_C, int *d_labels, long long *h_namb, float *d_diag, cudaStream_t stream);

extern "C" int mdb_kmeans_assign_tc(mdb_ctx *c, long long rows, int D, int K, const double *d_X, long long ldx,
                                    const int32_t *d_rowidx, const double *d_centers, int32_t *d_labels,
                                    long long *h_nrecheck, float *d_diag, void *stream) {
    if (check_ctx(c, "mdb_kmeans_assign_tc", false)) return -1;
    return kmeans_assign_tc_run(c, rows, D, K, d_X, ldx, d_rowidx, d_centers, d_labels, h_nrecheck, d_diag,
                                (cudaStream_t)stream);
}

int kmeans_inertia_run(mdb_ctx *c, long long rows, int D, const double *d_X, long long ldx, const int *d_rowidx,
                       const double *d_C, const int *d_labels, const double *d_w, double *h_inertia, cudaStream_t stream);

extern "C" int mdb_kmeans_inertia(mdb_ctx *c, long long rows, int D, const double *d_X, long long ldx,
                                  const int32_t *d_rowidx, const double *d_centers, const int32_t *d_labels,
                                  const double *d_weights, double *h_inertia, void *stream) {
    if (check_ctx(c, "mdb_kmeans_inertia", false)) return -1;
    return kmeans_inertia_run(c, rows, D, d_X, ldx, d_rowidx, d_centers, d_labels, d_weights, h_inertia,
                              (cudaStream_t)stream);
}

extern "C" int mdb_kmeans_partial_sums(mdb_ctx *c, long long nbatch, int D, int K, const double *d_X, long long ldx,
                                       const int32_t *d_rowidx, const double *d_weights, const int32_t *d_labels,
                                       double *d_sums, double *d_wsum, void *stream) {
    if (check_ctx(c, "mdb_kmeans_partial_sums", false)) return -1;
    return kmeans_sums_run(c, nbatch, D, K, d_X, ldx, d_rowidx, d_weights, d_labels, d_sums, d_wsum, (cudaStream_t)stream);
}

int kmeans_update_exact_run(mdb_ctx *c, long long B, int D, const double *d_X, long long ldx, const int *d_rowidx,
                            const double *d_w, const int *d_labels, double *d_C, double *d_weight_sums, cudaStream_t stream);
extern "C" int mdb_kmeans_update_exact(mdb_ctx *c, long long nbatch, int D, const double *d_X, long long ldx,
                                       const int32_t *d_rowidx, const double *d_weights, const int32_t *d_labels,
                                       double *d_centers, double *d_weight_sums, void *stream) {
    if (check_ctx(c, "mdb_kmeans_update_exact", false)) return -1;
    return kmeans_update_exact_run(c, nbatch, D, d_X, ldx, d_rowidx, d_weights, d_labels, d_centers, d_weight_sums,
                                   (cudaStream_t)stream);
}

int kmeans_minibatch_step_run(mdb_ctx *c, long long B, int D, int K, const double *d_X, long long ldx, const int *h_idx,
                              double *d_C, double *d_weight_sums, double *h_out, cudaStream_t stream);
extern "C" int mdb_kmeans_minibatch_step(mdb_ctx *c, long long nbatch, int D, int K, const double *d_X, long long ldx,
                                         const int32_t *h_rowidx, double *d_centers, double *d_weight_sums,
                                         double *h_inertia, int *h_nzero, void *stream) {
    if (check_ctx(c, "mdb_kmeans_minibatch_step", false)) return -1;
    double out[2] = {0.0, 0.0};
    int rc = kmeans_minibatch_step_run(c, nbatch, D, K, d_X, ldx, h_rowidx, d_centers, d_weight_sums, out,
                                       (cudaStream_t)stream);
    if (h_inertia) *h_inertia = out[0];
    if (h_nzero) *h_nzero = (int)out[1];
    return rc;
}

extern "C" int mdb_kmeans_apply_update(mdb_ctx *c, int K, int D, double *d_centers, double *d_weight_sums,
                                       const double *d_sums, const double *d_wsum, void *stream) {
    if (check_ctx(c, "mdb_kmeans_apply_update", false)) return -1;
    return kmeans_apply_run(c, K, D, d_centers, d_weight_sums, d_sums, d_wsum, (cudaStream_t)stream);
}

int kpp_step_run(mdb_ctx *c, long long n, int D, const double *d_X, long long ldx, const int *d_rowidx, const int *d_cand,
                 int T, const double *d_closest, const double *d_w, double *d_dist, double *d_pot, cudaStream_t stream);

extern "C" int mdb_kpp_step(mdb_ctx *c, long long n, int D, const double *d_X, long long ldx, const int32_t *d_rowidx,
                            const int32_t *d_cand, int ncand, const double *d_closest, const double *d_weights,
                            double *d_dist, double *d_pot, void *stream) {
    if (check_ctx(c, "mdb_kpp_step", false)) return -1;
    return kpp_step_run(c, n, D, d_X, ldx, d_rowidx, d_cand, ncand, d_closest, d_weights, d_dist, d_pot,
                        (cudaStream_t)stream);
}

int kmeans_plusplus_run(mdb_ctx *c, int B, long long n, int D, const double *d_X, long long ldx, const int *d_rowidx, int K,
                        int T, const double *h_rand, const int *h_first, int *d_indices, cudaStream_t stream);

extern "C" int mdb_kmeans_plusplus(mdb_ctx *c, int nbatch, long long n, int D, const double *d_X, long long ldx,
                                   const int32_t *d_rowidx, int K, int n_local_trials, const double *h_rand,
                                   const int32_t *h_first_center, int32_t *d_indices, void *stream) {
    if (check_ctx(c, "mdb_kmeans_plusplus", false)) return -1;
    return kmeans_plusplus_run(c, nbatch, n, D, d_X, ldx, d_rowidx, K, n_local_trials, h_rand, h_first_center, d_indices,
                               (cudaStream_t)stream);
}

extern "C" int mdb_minmax_fit(mdb_ctx *c, long long rows, int D, const double *d_X, long long ldx, double *d_min,
                              double *d_max, void *stream) {
    if (check_ctx(c, "mdb_minmax_fit", false)) return -1;
    return minmax_fit_run(c, rows, D, d_X, ldx, d_min, d_max, 0, (cudaStream_t)stream);
}

extern "C" int mdb_minmax_transform(mdb_ctx *c, long long rows, int D, double *d_X, long long ldx, const double *h_min,
                                    const double *h_max, void *stream) {
    if (check_ctx(c, "mdb_minmax_transform", false)) return -1;
    return minmax_transform_run(c, rows, D, d_X, ldx, h_min, h_max, (cudaStream_t)stream);
}

// per-kernel timing: "name count total_ms" lines, accumulated since the last read
extern "C" int mdb_profile_enable(mdb_ctx *c, int on) {
    if (check_ctx(c, "mdb_profile_enable", false)) return -1;
    c->prof_on = on != 0;
    return 0;
}

extern "C" int mdb_profile_read(mdb_ctx *c, char *buf, size_t cap) {
    if (check_ctx(c, "mdb_profile_read", false)) return -1;
    MDB_CUDA(cudaDeviceSynchronize());
    std::vector<std::string> names;
    std::vector<double> ms;
    std::vector<long long> cnt;
    for (auto &r : c->prof_recs) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        size_t k = 0;
        for (; k < names.size(); ++k)
            if (names[k] == r.name) break;
        if (k == names.size()) {
            names.push_back(r.name);
            ms.push_back(0);
            cnt.push_back(0);
        }
        ms[k] += t;
        cnt[k] += 1;
        c->prof_free.push_back(r.a);
        c->prof_free.push_back(r.b);
    }
    c->prof_recs.clear();
    std::string out;
    for (size_t k = 0; k < names.size(); ++k)
        out += names[k] + " " + std::to_string(cnt[k]) + " " + std::to_string(ms[k]) + "\n";
    if (buf && cap) {
        size_t n = std::min(cap - 1, out.size());
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// host-buffer pipeline: H2D(k+1) | compute(k) | D2H(k-1) on three streams, double buffered.
// ---------------------------------------------------------------------------------------------
// ELL rows narrowed to the first `w` slots (post-cleanup degrees are bounded by the largest valence of
// the element table, e.g. 4 for C/H/O); flags an error if a row holds more
__global__ void __launch_bounds__(256) k_ell_narrow(const int32_t *__restrict__ ell, int maxb, int w, long long fn,
                                                    int32_t *__restrict__ out, BondStats *stats) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const int4 q = *reinterpret_cast<const int4 *>(ell + a * maxb);
    if (ell[a * maxb + w] >= 0) atomicExch(&stats->err, MDB_ERR_CSR_OVERFLOW);
    *reinterpret_cast<int4 *>(out + a * 4) = q;
}

// Bond list of a chunk as (i, j) pairs with i < j, by ascending i then j (frame-local indices): what
// iterating the bonds of the reference's OBMol yields (detect.py:282-291), each bond once.
__global__ void __launch_bounds__(256) k_edge_count(const int32_t *__restrict__ ell, int maxb, int N, long long fn,
                                                    int *__restrict__ cnt) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a > fn) return;
    int c = 0;
    if (a < fn) {
        const int i = (int)(a % N);
        for (int k = 0; k < maxb; ++k) c += ell[a * maxb + k] > i;
    }
    cnt[a] = c;
}

__global__ void __launch_bounds__(256) k_edge_fill(const int32_t *__restrict__ ell, int maxb, int N, long long fn,
                                                   const int *__restrict__ ofs, int2 *__restrict__ edges) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= fn) return;
    const int i = (int)(a % N);
    int o = ofs[a];
    for (int k = 0; k < maxb; ++k) {
        const int e = ell[a * maxb + k];
        if (e > i) edges[o++] = make_int2(i, e);
    }
}

__global__ void k_frame_offsets(const int *__restrict__ ofs, int N, int F, int *__restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f <= F) out[f] = ofs[(size_t)f * N];
}

struct PipeBuf {
    double *pos = nullptr;
    int32_t *ell = nullptr;
    int32_t *ell4 = nullptr;
    uint64_t *keys = nullptr;
    int32_t *labels = nullptr;
    uint8_t *code = nullptr;
    int *ecnt = nullptr;    // edge-list output: per atom-frame offsets (after the scan), [cn + 1]
    int2 *edges = nullptr;  // [cn * max_bonds / 2]
    void *escan = nullptr;  // scan scratch
    int *fofs = nullptr;    // per-frame offsets [per + 1]
};

// keys -> one-byte codes through a 256-slot open-addressing dictionary shared by the whole call
#define KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
__global__ void __launch_bounds__(256) k_key_encode(const uint64_t *__restrict__ keys, long long n,
                                                    unsigned long long *__restrict__ table, uint8_t *__restrict__ code,
                                                    BondStats *stats) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = keys[i];
    unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 56);
    for (int probe = 0; probe < 256; ++probe, h = (h + 1) & 255u) {
        unsigned long long cur = table[h];
        if (cur == KEY_EMPTY) cur = atomicCAS(&table[h], KEY_EMPTY, key);
        if (cur == KEY_EMPTY || cur == key) {
            code[i] = (uint8_t)h;
            return;
        }
    }
    atomicExch(&stats->err, MDB_ERR_KEY_TABLE_FULL);
}

static int analyze_host_impl(mdb_ctx *c, int nframes, int natoms, const double *h_pos, const double *h_cell,
                             const uint8_t *h_types, int periodic, int32_t *h_ell, int ell_width, uint64_t *h_keys,
                             uint8_t *h_keycode, uint64_t *h_keytable, int32_t *h_labels, int32_t *h_edges = nullptr,
                             long long edge_cap = 0, long long *h_edge_ptr = nullptr) {
    if (h_edge_ptr) h_edge_ptr[0] = 0;
    if (nframes <= 0 || natoms <= 0) return 0;
    const int maxb = c->opt_max_bonds;
    // the host path pipelines copies against compute, so it wants more, smaller chunks than the device
    // path (measured on B200, 100k-atom frames: 160 / 80 / 40 / 20 / 10 frames per chunk = 68.9 / 63.1 /
    // 61.6 / 63.1 / 67.5 ms per 1e8 atom-frames)
    int per = frames_per_chunk(c, nframes, natoms);
    if (c->opt_frames_per_launch <= 0) per = std::max(1, std::min(per, (4 << 20) / std::max(natoms, 1)));
    const bool want_keys = h_keys || h_keycode;
    const size_t cn = (size_t)per * natoms;
    const size_t sz_pos = align256(cn * 3 * sizeof(double)), sz_ell = align256(cn * maxb * sizeof(int32_t)),
                 sz_keys = align256(cn * sizeof(uint64_t)), sz_lab = align256(cn * sizeof(int32_t)),
                 sz_code = h_keycode ? align256(cn) + 256 * sizeof(uint64_t) : 0;
    if (ell_width != 0 && ell_width != 4 && ell_width != maxb) {
        mdb_set_error("mdb_analyze_frames_host: ell_width must be 0 (= max_bonds), 4 or max_bonds");
        return -1;
    }
    const bool narrow = ell_width == 4 && maxb != 4;
    const size_t sz_ell4 = narrow ? align256(cn * 4 * sizeof(int32_t)) : 0;
    // edge-list output: counts / offsets, the pairs, scan scratch and per-frame offsets ride in the same buffers
    const size_t sz_ecnt = h_edges ? align256((cn + 16) * sizeof(int)) : 0,
                 sz_edges = h_edges ? align256(cn * (size_t)(maxb / 2) * sizeof(int2)) : 0,
                 sz_escan = h_edges ? align256(scan_scratch_bytes((long long)cn + 1)) : 0,
                 sz_fofs = h_edges ? align256(((size_t)per + 1) * sizeof(int)) : 0;
    const size_t need = sz_pos + sz_ell + sz_keys + sz_lab + sz_ell4 + sz_code + sz_ecnt + sz_edges + sz_escan + sz_fofs;
    PipeBuf b[2];
    for (int k = 0; k < 2; ++k) {
        if (c->pipe_cap[k] < need) {
            MDB_CUDA(cudaDeviceSynchronize());
            if (c->pipe_mem[k]) MDB_CUDA(cudaFree(c->pipe_mem[k]));
            c->pipe_mem[k] = nullptr;
            c->pipe_cap[k] = 0;
            MDB_CUDA(cudaMalloc(&c->pipe_mem[k], need));
            c->pipe_cap[k] = need;
        }
        char *m = c->pipe_mem[k];
        b[k].pos = (double *)m;
        b[k].ell = (int32_t *)(m + sz_pos);
        b[k].keys = (uint64_t *)(m + sz_pos + sz_ell);
        b[k].labels = (int32_t *)(m + sz_pos + sz_ell + sz_keys);
        b[k].ell4 = (int32_t *)(m + sz_pos + sz_ell + sz_keys + sz_lab);
        b[k].code = (uint8_t *)(m + sz_pos + sz_ell + sz_keys + sz_lab + sz_ell4);
        char *e = m + sz_pos + sz_ell + sz_keys + sz_lab + sz_ell4 + sz_code;
        b[k].ecnt = (int *)e;
        b[k].edges = (int2 *)(e + sz_ecnt);
        b[k].escan = e + sz_ecnt + sz_edges;
        b[k].fofs = (int *)(e + sz_ecnt + sz_edges + sz_escan);
        if (h_edges && c->pipe_hofs_cap[k] < (size_t)per + 1) {
            if (c->pipe_hofs[k]) MDB_CUDA(cudaFreeHost(c->pipe_hofs[k]));
            c->pipe_hofs[k] = nullptr;
            c->pipe_hofs_cap[k] = 0;
            MDB_CUDA(cudaMallocHost(&c->pipe_hofs[k], ((size_t)per + 1) * sizeof(int)));
            c->pipe_hofs_cap[k] = (size_t)per + 1;
        }
    }
    // the dictionary lives behind the code bytes of buffer 0
    unsigned long long *d_table = h_keycode ? (unsigned long long *)(b[0].code + align256(cn)) : nullptr;
    if (c->pipe_types_cap < (size_t)natoms) {
        MDB_CUDA(cudaDeviceSynchronize());
        if (c->pipe_types) MDB_CUDA(cudaFree(c->pipe_types));
        c->pipe_types = nullptr;
        MDB_CUDA(cudaMalloc(&c->pipe_types, (size_t)natoms + 256));
        c->pipe_types_cap = (size_t)natoms + 256;
    }
    cudaStream_t s_in = c->pipe[0], s_c = c->pipe[1], s_out = c->pipe[2];
    MDB_CUDA(cudaMemcpyAsync(c->pipe_types, h_types, natoms, cudaMemcpyHostToDevice, s_c));
    MDB_CUDA(cudaMemsetAsync(c->d_stats, 0, sizeof(BondStats), s_c));
    if (d_table) MDB_CUDA(cudaMemsetAsync(d_table, 0xFF, 256 * sizeof(uint64_t), s_c));
    long long edges_total = 0;
    int emit_err = 0;
    // device -> host copies of one finished chunk.  With the edge list the size of the copy is only known
    // once the chunk has been computed, so that chunk's copies are issued one iteration later (the next
    // chunk's input copy and kernels are already queued by then: nothing idles while the host waits).
    auto emit = [&](int f0, int F, int k) -> int {
        PipeBuf &B = b[k];
        const size_t o = (size_t)f0 * natoms, n = (size_t)F * natoms;
        MDB_CUDA(cudaStreamWaitEvent(s_out, c->pipe_comp[k], 0));
        if (h_edges) {
            MDB_CUDA(cudaEventSynchronize(c->pipe_comp[k]));
            const int *hf = c->pipe_hofs[k];
            const long long E = hf[F];
            if (edges_total + E > edge_cap) {
                mdb_set_error("mdb_analyze_frames_host_edges: edge capacity " + std::to_string(edge_cap) + " is too small");
                emit_err = 1;
                return -1;
            }
            for (int f = 0; f < F; ++f) h_edge_ptr[f0 + f + 1] = edges_total + hf[f + 1];
            if (E) MDB_CUDA(cudaMemcpyAsync(h_edges + 2 * edges_total, B.edges, (size_t)E * sizeof(int2), cudaMemcpyDeviceToHost, s_out));
            edges_total += E;
        } else if (h_ell && narrow) {
            MDB_CUDA(cudaMemcpyAsync(h_ell + o * 4, B.ell4, n * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s_out));
        } else if (h_ell) {
            MDB_CUDA(cudaMemcpyAsync(h_ell + o * maxb, B.ell, n * maxb * sizeof(int32_t), cudaMemcpyDeviceToHost, s_out));
        }
        if (h_keys) MDB_CUDA(cudaMemcpyAsync(h_keys + o, B.keys, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s_out));
        if (h_keycode) MDB_CUDA(cudaMemcpyAsync(h_keycode + o, B.code, n, cudaMemcpyDeviceToHost, s_out));
        if (h_labels)
            MDB_CUDA(cudaMemcpyAsync(h_labels + o, B.labels, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s_out));
        MDB_CUDA(cudaEventRecord(c->pipe_out[k], s_out));
        return 0;
    };
    int chunk = 0, prev_f0 = -1, prev_F = 0;
    for (int f0 = 0; f0 < nframes; f0 += per, ++chunk) {
        const int F = std::min(per, nframes - f0);
        const size_t o = (size_t)f0 * natoms, n = (size_t)F * natoms;
        const int k = chunk & 1;
        PipeBuf &B = b[k];
        if (chunk >= 2) MDB_CUDA(cudaStreamWaitEvent(s_in, c->pipe_comp[k], 0));
        MDB_CUDA(cudaMemcpyAsync(B.pos, h_pos + o * 3, n * 3 * sizeof(double), cudaMemcpyHostToDevice, s_in));
        MDB_CUDA(cudaEventRecord(c->pipe_in[k], s_in));
        MDB_CUDA(cudaStreamWaitEvent(s_c, c->pipe_in[k], 0));
        if (chunk >= 2) MDB_CUDA(cudaStreamWaitEvent(s_c, c->pipe_out[k], 0));
        if (analyze_chunk(c, F, natoms, B.pos, h_cell ? h_cell + (size_t)f0 * 9 : nullptr, c->pipe_types, periodic, B.ell,
                          want_keys ? B.keys : nullptr, h_labels ? B.labels : nullptr, s_c)) {
            cudaDeviceSynchronize();
            return -1;
        }
        if (h_keycode) {
            k_key_encode<<<cdiv((long long)n, 256), 256, 0, s_c>>>(B.keys, (long long)n, d_table, B.code, c->d_stats);
            c->launches++;
        }
        if (h_ell && narrow && !h_edges) {
            k_ell_narrow<<<cdiv((long long)n, 256), 256, 0, s_c>>>(B.ell, maxb, 4, (long long)n, B.ell4, c->d_stats);
            c->launches++;
        }
        if (h_edges) {
            k_edge_count<<<cdiv((long long)n + 1, 256), 256, 0, s_c>>>(B.ell, maxb, natoms, (long long)n, B.ecnt);
            if (device_exclusive_scan(c, B.ecnt, (long long)n + 1, B.escan, s_c)) {
                cudaDeviceSynchronize();
                return -1;
            }
            k_edge_fill<<<cdiv((long long)n, 256), 256, 0, s_c>>>(B.ell, maxb, natoms, (long long)n, B.ecnt, B.edges);
            // written straight into page-locked host memory (unified addressing): no copy-engine queueing
            k_frame_offsets<<<cdiv(F + 1, 256), 256, 0, s_c>>>(B.ecnt, natoms, F, c->pipe_hofs[k]);
            c->launches += 3;
        }
        MDB_CUDA(cudaEventRecord(c->pipe_comp[k], s_c));
        if (h_edges) {
            if (prev_f0 >= 0 && emit(prev_f0, prev_F, k ^ 1)) break;
            prev_f0 = f0;
            prev_F = F;
        } else if (emit(f0, F, k)) {
            break;
        }
    }
    if (h_edges && !emit_err && prev_f0 >= 0) emit(prev_f0, prev_F, (chunk - 1) & 1);
    MDB_CUDA(cudaStreamSynchronize(s_in));
    MDB_CUDA(cudaStreamSynchronize(s_c));
    MDB_CUDA(cudaStreamSynchronize(s_out));
    if (h_keycode && h_keytable) MDB_CUDA(cudaMemcpy(h_keytable, d_table, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    BondStats hs;
    MDB_CUDA(cudaMemcpy(&hs, c->d_stats, sizeof(hs), cudaMemcpyDeviceToHost));
    if (emit_err) return -1;
    if (hs.err) {
        mdb_set_error(std::string("mdb_analyze_frames_host: ") + err_text(hs.err));
        return hs.err;
    }
    return 0;
}

extern "C" int mdb_analyze_frames_host(mdb_ctx *c, int nframes, int natoms, const double *h_pos, const double *h_cell,
                                       const uint8_t *h_types, int periodic, int32_t *h_ell, int ell_width,
                                       uint64_t *h_keys, int32_t *h_labels) {
    if (check_ctx(c, "mdb_analyze_frames_host")) return -1;
    return analyze_host_impl(c, nframes, natoms, h_pos, h_cell, h_types, periodic, h_ell, ell_width, h_keys, nullptr, nullptr,
                             h_labels);
}

extern "C" int mdb_analyze_frames_host_edges(mdb_ctx *c, int nframes, int natoms, const double *h_pos, const double *h_cell,
                                             const uint8_t *h_types, int periodic, int32_t *h_edges, long long edge_cap,
                                             long long *h_edge_ptr, uint8_t *h_keycode, uint64_t *h_keytable,
                                             int32_t *h_labels) {
    if (check_ctx(c, "mdb_analyze_frames_host_edges")) return -1;
    if (!h_edges || !h_edge_ptr || edge_cap < 0 || (h_keycode && !h_keytable)) {
        mdb_set_error("mdb_analyze_frames_host_edges: h_edges and h_edge_ptr are required (and h_keytable with h_keycode)");
        return -1;
    }
    return analyze_host_impl(c, nframes, natoms, h_pos, h_cell, h_types, periodic, nullptr, 0, nullptr, h_keycode, h_keytable,
                             h_labels, h_edges, edge_cap, h_edge_ptr);
}

extern "C" int mdb_analyze_frames_host_coded(mdb_ctx *c, int nframes, int natoms, const double *h_pos, const double *h_cell,
                                             const uint8_t *h_types, int periodic, int32_t *h_ell, int ell_width,
                                             uint8_t *h_keycode, uint64_t *h_keytable, int32_t *h_labels) {
    if (check_ctx(c, "mdb_analyze_frames_host_coded")) return -1;
    if (!h_keycode || !h_keytable) {
        mdb_set_error("mdb_analyze_frames_host_coded: h_keycode and h_keytable are required");
        return -1;
    }
    return analyze_host_impl(c, nframes, natoms, h_pos, h_cell, h_types, periodic, h_ell, ell_width, nullptr, h_keycode,
                             h_keytable, h_labels);
}
