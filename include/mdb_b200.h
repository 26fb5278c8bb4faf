/* mdb_b200.h -- C ABI of the B200-native per-frame analysis path of MDDatasetBuilder.
 *
 * This is the drop-in boundary.  The reference has no FFI registry
 * for this path; its per-frame work sits behind Python call sites that run inside
 * multiprocessing.Pool children.  Each entry point below names the reference
 * interface (file:line under /root/reference/mddatasetbuilder) whose arithmetic it
 * replaces.  Signatures use plain pointers and sizes only (no torch types); the
 * reference-side ctypes stubs are in INTEGRATION.md.
 *
 * Conventions
 *  - All functions return 0 on success, non-zero on error; mdb_last_error() gives
 *    the message (thread-local).  Python maps non-zero to RuntimeError, like the
 *    reference's own error behaviour for malformed frames (detect.py:49,83,174,189).
 *  - Pointers named d_* are DEVICE pointers owned by the caller; h_* are HOST
 *    pointers.  Work is enqueued on `stream` (a cudaStream_t passed as void*); the
 *    call returns without synchronising unless stated.
 *  - A batch holds `nframes` frames of `natoms` atoms (the reference requires a
 *    constant atom count, detect.py:182-191).  Positions are double [F][N][3] in
 *    atom-id order (id-1), cells are double [F][9] = three lattice-vector rows, as
 *    DetectDump.readcrd builds them (detect.py:326-345).
 *  - types[i] is the 0-based index into `atomname` (LAMMPS type - 1;
 *    detect.py:87,193,314).
 *  - Neighbour lists are exchanged as ELL rows: int32 [F][N][maxb], ascending
 *    partner index, padded with -1 (one 32-byte sector per atom at maxb = 8), and on
 *    request as CSR (rowptr int32 [F][N+1] frame-local, col int32 [F][colcap]).
 *  - Scratch memory comes from an arena owned by the context; it grows on first use
 *    of a batch shape and is reused afterwards (no allocation in steady state).
 */
#ifndef MDB_B200_H
#define MDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mdb_ctx mdb_ctx;

#define MDB_ORDER_INDEX 0 /* rows ascending by partner index (canonical) */
#define MDB_ORDER_OB 1    /* rows in Open Babel bond-iteration order = ascending partner (z, index) */

/* ---- library / context ------------------------------------------------------ */
int mdb_version(void);
const char *mdb_last_error(void);

/* One context per (process, GPU).  Replaces the per-process state the reference
 * keeps in a Pool child (utils.py:198-231): one context drives one GPU. */
int mdb_create(int device, mdb_ctx **out);
void mdb_destroy(mdb_ctx *ctx);

/* Element table: atomic number per type index (atomname -> Z, as
 * ase.data.atomic_numbers at datasetbuilder.py:145-147).  Fixes the covalent radii,
 * max-valence and Coulomb diagonal (Z^2.4/2) tables used by the kernels. */
int mdb_set_elements(mdb_ctx *ctx, int ntypes, const int *atomic_numbers);

/* Tunables: "cell_edge" (A; 0 = heuristic), "max_bonds" (8 or 16),
 * "frames_per_launch" (0 = heuristic).  Returns non-zero for an unknown key. */
int mdb_set_option(mdb_ctx *ctx, const char *key, double value);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t mdb_launch_count(const mdb_ctx *ctx);
size_t mdb_arena_bytes(const mdb_ctx *ctx);

/* ---- K1+K2: bond detection from coordinates ---------------------------------
 * Replaces DetectDump._crd2bond(step_atoms, readlevel=False) (detect.py:249-292):
 * Open Babel ConnectTheDots = pairs with 0.16 <= d2_MIC <= (Rcov_i+Rcov_j+0.45)^2
 * followed by the index-ordered valence / 45-degree-angle cleanup.
 * h_cell: host [F][9]; periodic: 0/1 (DatasetBuilder(pbc=...)).
 * d_ell: out int32 [F][N][max_bonds]. */
int mdb_bonds_from_coords(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos,
                          const double *h_cell, const uint8_t *d_types, int periodic,
                          int32_t *d_ell, void *stream);

/* ELL -> CSR (adjacency lists of detect.py:278-291).  order = MDB_ORDER_INDEX or
 * MDB_ORDER_OB (needs d_pos for the z sort).  d_rowptr [F][N+1], d_col [F][colcap].
 * Fails (after synchronising) if a frame holds more than colcap entries. */
int mdb_ell_to_csr(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, int order,
                   const double *d_pos, int32_t *d_rowptr, int32_t *d_col, int colcap,
                   void *stream);

/* ---- K3: bond-type keys ------------------------------------------------------
 * Replaces the per-atom key of DetectDump.readatombondtype (detect.py:219-225) in
 * dump-only mode: (element, sorted bond orders) with every order 1 (tier-1
 * deviation: PerceiveBondOrders is not restated).  Key packing: bits 56..63
 * element index + 1, bits 48..55 number of levels, nibble k = k-th smallest level.
 * Also writes the union-find seed (parent = min(self, smallest neighbour)) when
 * d_parent != NULL, so that mdb_components can skip its own init pass. */
int mdb_bond_keys(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell,
                  const uint8_t *d_types, uint64_t *d_keys, int32_t *d_parent, void *stream);

/* Bond-file mode keys, replaces DetectBond.readatombondtype (detect.py:105-119):
 * level = max(1, round_half_even(bo)), sorted ascending.  d_bo: double per ELL slot. */
int mdb_bond_keys_bo(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell,
                     const double *d_bo, const uint8_t *d_types, uint64_t *d_keys, void *stream);

/* Bond-file mode keys from a CSR bond table (one row per atom and frame): d_rowptr [F][N+1]
 * frame-local offsets, d_bo [F][colcap] the bond orders in file order -- the columns
 * s[4+nb : 4+2nb] of a `fix reaxff/bonds` line (detect.py:112-115).  Same key as mdb_bond_keys_bo. */
int mdb_bond_keys_csr(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_rowptr, const double *d_bo,
                      long long colcap, const uint8_t *d_types, uint64_t *d_keys, void *stream);

/* ---- group-by of DatasetBuilder._readtimestepsbond (datasetbuilder.py:185-214) on the device ----
 * The reference turns every frame into {key -> [atom ids]} and appends (step, ids) to one list per
 * bond type.  Here a key dictionary lives on the device for the whole run -- d_tab_keys [cap]
 * (initialise to all ones = empty), d_tab_count [cap] (0; atoms seen per key, all frames),
 * d_tab_first [cap] (all ones; smallest (step << 32 | atom index) per key = first-seen order,
 * datasetbuilder.py:205-206) -- and every call adds one batch of frames:
 *   d_codes [F][N]   out: table slot of each atom-frame (0xffff = dropped by d_keep);
 *   d_order [F*N]    out (optional): batch-local atom-frame indices f*N + i sorted by slot, ascending
 *                    inside a slot (frames, then atom ids: the row order of a bond type);
 *   d_seg  [cap+1]   out (optional): d_order[d_seg[s] .. d_seg[s+1]) belongs to slot s;
 *                    d_seg[cap] = number of kept atom-frames.
 * d_keep [F][N] (optional, 0 = drop) is the model-deviation filter of detect.py:215-223.  Slot numbers
 * depend on insertion order, translate them through d_tab_keys.  cap: power of two, 32..4096; more
 * than cap distinct keys raises the device error flag (mdb_bond_stats). */
int mdb_group_keys(mdb_ctx *ctx, int nframes, int natoms, const uint64_t *d_keys, const uint8_t *d_keep,
                   long long step0, int cap, uint64_t *d_tab_keys, long long *d_tab_count,
                   long long *d_tab_first, uint16_t *d_codes, int32_t *d_order, long long *d_seg,
                   void *stream);

/* Per-type max_counter (datasetbuilder.py:262-272): d_maxc [cap][16] (in/out; 16 = the element-type capacity of the library) takes, for the slot of every
 * target, the element-wise maximum of its composition d_counts [ntargets][ntypes].  d_targets: atom-frame
 * indices into d_codes, or NULL = 0 .. ntargets-1. */
int mdb_type_maxcount(mdb_ctx *ctx, const int32_t *d_targets, long long ntargets, const uint16_t *d_codes,
                      const int32_t *d_counts, int32_t *d_maxc, void *stream);

/* ---- K4: connectivity --------------------------------------------------------
 * Replaces the partition computed by dps(bonds) (dps.pyx:17-46): d_labels[f][i] =
 * smallest atom index of the molecule containing i (the DFS root the reference
 * starts each molecule from).  seeded != 0 means d_labels already holds the seed
 * written by mdb_bond_keys. */
int mdb_components(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell,
                   int32_t *d_labels, int seeded, void *stream);

/* The same partition from a CSR bond table (bond-file mode, DetectBond.readmolecule, detect.py:137-144):
 * d_rowptr [F][N+1], d_col [F][colcap] 0-based partners, every bond listed from both ends. */
int mdb_components_csr(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_rowptr, const int32_t *d_col,
                       long long colcap, int32_t *d_labels, void *stream);

/* Exact dps output (molecule order by smallest index, members in the reference's
 * LIFO-DFS visiting order, dps.pyx:27-43) for one frame, from a CSR whose row
 * order is the adjacency order.  d_mol_atoms [N], d_mol_ptr [N+1], *h_nmol out
 * (synchronises). */
int mdb_dps(mdb_ctx *ctx, int natoms, const int32_t *d_rowptr, const int32_t *d_col,
            int32_t *d_mol_atoms, int32_t *d_mol_ptr, int *h_nmol, void *stream);

/* ---- fused per-frame bond path (the benchmark "step") -------------------------
 * cell list -> bonds -> cleanup -> keys -> connectivity for a batch, i.e. what
 * readatombondtype + readmolecule produce per frame in dump-only mode.
 * Any output pointer may be NULL. */
int mdb_analyze_frames(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos,
                       const double *h_cell, const uint8_t *d_types, int periodic,
                       int32_t *d_ell, uint64_t *d_keys, int32_t *d_labels, void *stream);

/* Host-buffer variant (the e2e path): copies positions H2D, runs
 * mdb_analyze_frames, copies ELL / keys / labels D2H, pipelined over internal
 * streams in chunks of frames; synchronises before returning.  Buffers should be
 * page-locked for full PCIe bandwidth. */
int mdb_analyze_frames_host(mdb_ctx *ctx, int nframes, int natoms, const double *h_pos,
                            const double *h_cell, const uint8_t *h_types, int periodic,
                            int32_t *h_ell, int ell_width, uint64_t *h_keys, int32_t *h_labels);
/* ell_width: slots per row of h_ell -- 0 or max_bonds for the full rows, 4 to return
 * [F][N][4] (lossless whenever every element's max valence is <= 4, e.g. C/H/O/N: the cleanup
 * guarantees degree <= MaxBonds; a longer row is an error). */

/* Same, with the bond-type keys returned as one byte per atom: h_keycode[f][i] indexes h_keytable
 * (256 entries, unused = all ones), a dictionary of the distinct 64-bit keys of the whole call.  What
 * the caller of readatombondtype (reference detect.py:196-226) groups by is the key, of which a frame
 * has a few dozen at most; 8 -> 1 byte per atom takes the device-to-host copy off the critical path.
 * More than 256 distinct keys is an error (use mdb_analyze_frames_host). */
int mdb_analyze_frames_host_coded(mdb_ctx *ctx, int nframes, int natoms, const double *h_pos,
                                  const double *h_cell, const uint8_t *h_types, int periodic,
                                  int32_t *h_ell, int ell_width, uint8_t *h_keycode,
                                  uint64_t *h_keytable, int32_t *h_labels);

/* Same pipeline with the bonds returned as a list instead of rows: h_edges [edge_cap][2] receives every
 * bond once as (i, j), i < j (frame-local 0-based indices, ascending i then j), the bonds of frame f at
 * h_edge_ptr[f] .. h_edge_ptr[f+1] (h_edge_ptr [nframes + 1]) -- the shape in which the reference walks
 * them, `for b in OBMolBondIter(mol)` (detect.py:282-291) or the id / partner columns of a bond file.
 * About 8 bytes per bond instead of 16 per atom take the device-to-host copy off the PCIe critical path
 * (a chunk's copy is sized after its kernels finished; the next chunk is already queued by then).
 * h_keycode / h_keytable / h_labels as above, each may be NULL.  A too small edge_cap is an error. */
int mdb_analyze_frames_host_edges(mdb_ctx *ctx, int nframes, int natoms, const double *h_pos,
                                  const double *h_cell, const uint8_t *h_types, int periodic,
                                  int32_t *h_edges, long long edge_cap, long long *h_edge_ptr,
                                  uint8_t *h_keycode, uint64_t *h_keytable, int32_t *h_labels);

/* ---- K5: element composition of the cutoff sphere -----------------------------
 * Replaces, per target atom, `distances = get_distances(a, all, mic=True);
 * Counter(symbols[distances < cutoff])` (datasetbuilder.py:312-322): d_counts
 * [ntargets][ntypes] (the sphere includes the target itself), and updates the
 * running per-element maximum h_maxcount[ntypes] (in/out; the reference's
 * max_counter, datasetbuilder.py:271) -- synchronises.  d_targets: batch-local
 * atom-frame indices f*natoms + i (int32), or NULL for every atom of every frame (NULL with
 * ntargets == 0 is an empty list: nothing is computed).
 * Orthorhombic cells only. */
int mdb_compositions(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos,
                     const double *h_cell, const uint8_t *d_types, int periodic, double cutoff,
                     const int32_t *d_targets, long long ntargets, int32_t *d_counts,
                     int32_t *h_maxcount, void *stream);

/* Sphere membership lists: for every target the atoms with MIC distance < cutoff in ascending
 * index (the target included) -- `np.where(distances < self.cutoff)` of the output cut,
 * datasetbuilder.py:554-557.  d_members [ntargets][cap] (-1 padded), d_nmem [ntargets] (a
 * value above cap means the list was truncated). */
int mdb_sphere_members(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos,
                       const double *h_cell, const uint8_t *d_types, int periodic, double cutoff,
                       const int32_t *d_targets, long long ntargets, int cap, int32_t *d_members,
                       int32_t *d_nmem, void *stream);

/* ---- K6: Coulomb-matrix eigen-spectrum descriptor -------------------------------
 * Replaces _writestepmatrix + _calcoulumbmatrix (datasetbuilder.py:283-349) and the
 * feature assembly (datasetbuilder.py:236-276).  For every target:
 *   d_eig [ntargets][eigcap]  ascending eigenvalues (NaN padded), optional;
 *   d_n   [ntargets]          cluster size, optional;
 *   d_X   [ntargets][D]       sort(eigenvalues U pads), D = sum(h_maxcount), optional,
 * where d_counts are the compositions from mdb_compositions for the same targets and
 * h_maxcount the per-element maxima over ALL rows of the bond type. */
int mdb_descriptors(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos,
                    const double *h_cell, const uint8_t *d_types, int periodic, double cutoff,
                    const int32_t *d_targets, long long ntargets, const int32_t *d_counts,
                    const int32_t *h_maxcount, double *d_X, double *d_eig, int eigcap,
                    int32_t *d_n, void *stream);

/* One-pass variant for callers that visit every frame once (the reference visits them once per
 * bond type, datasetbuilder.py:236-276, but needs max_counter over ALL rows before it can pad):
 *   mdb_compositions_members   = mdb_compositions + the sphere membership of every target into
 *                                d_members [ntargets][cap] (-1 padded, any order);
 *   mdb_descriptors_members    = mdb_descriptors with the membership taken from those lists instead
 *                                of a second cell scan (no cutoff argument; d_X may be NULL and the
 *                                spectra kept in d_eig / d_n);
 *   mdb_feature_rows           = the padded sorted rows d_X [ntargets][D] from stored spectra once
 *                                the final h_maxcount is known (the same merge mdb_descriptors does). */
int mdb_compositions_members(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos,
                             const double *h_cell, const uint8_t *d_types, int periodic,
                             double cutoff, const int32_t *d_targets, long long ntargets,
                             int32_t *d_counts, int32_t *h_maxcount, int cap, int32_t *d_members,
                             void *stream);
int mdb_descriptors_members(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos,
                            const double *h_cell, const uint8_t *d_types, int periodic,
                            const int32_t *d_targets, long long ntargets, const int32_t *d_counts,
                            const int32_t *h_maxcount, const int32_t *d_members, int cap,
                            double *d_X, double *d_eig, int eigcap, int32_t *d_n, void *stream);
int mdb_feature_rows(mdb_ctx *ctx, long long ntargets, const double *d_eig, int eigcap,
                     const int32_t *d_n, const int32_t *d_counts, const int32_t *h_maxcount,
                     double *d_X, void *stream);

/* Compact spectra for callers that keep them across passes (the streaming job computes every row twice --
 * scaler bounds first, assignment later -- and keeps the spectra of as many frames as HBM allows):
 * mdb_descriptors_members_compact = mdb_descriptors_members with the eigenvalues of target t written to
 * d_eig_flat[d_eig_off[t] .. d_eig_off[t+1]) (ascending; d_eig_off int32 [ntargets + 1] is an output, the
 * caller sizes d_eig_flat with the sum of d_counts); mdb_feature_rows_compact = mdb_feature_rows from that
 * layout: the same padded sorted rows, bit for bit. */
int mdb_descriptors_members_compact(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos,
                                    const double *h_cell, const uint8_t *d_types, int periodic,
                                    const int32_t *d_targets, long long ntargets, const int32_t *d_counts,
                                    const int32_t *h_maxcount, const int32_t *d_members, int cap,
                                    double *d_X, double *d_eig_flat, int32_t *d_eig_off, int32_t *d_n,
                                    void *stream);
int mdb_feature_rows_compact(mdb_ctx *ctx, long long ntargets, const double *d_eig_flat,
                             const int32_t *d_eig_off, const int32_t *d_counts, const int32_t *h_maxcount,
                             double *d_X, void *stream);

/* ---- MinMaxScaler (datasetbuilder.py:369-370; sklearn preprocessing/_data.py) ------
 * fit: column minima / maxima of d_X [rows][D] (row stride ldx doubles) into d_min /
 * d_max [D] (device).  transform: X = X * scale + min_ in place with scale = 1/range
 * (range < 10 eps -> 1) and min_ = 0 - data_min * scale, as two rounded operations. */
int mdb_minmax_fit(mdb_ctx *ctx, long long rows, int D, const double *d_X, long long ldx,
                   double *d_min, double *d_max, void *stream);
int mdb_minmax_transform(mdb_ctx *ctx, long long rows, int D, double *d_X, long long ldx,
                         const double *h_min, const double *h_max, void *stream);
/* Streaming fit (MinMaxScaler.partial_fit): folds the column bounds of these rows into the running
 * d_min / d_max [D], which the caller starts at (+DBL_MAX, -DBL_MAX). */
int mdb_minmax_accumulate(mdb_ctx *ctx, long long rows, int D, const double *d_X, long long ldx,
                          double *d_min, double *d_max, void *stream);
/* d_out[d_dst[r]] = d_X[d_src[r]] for r < n (rows of D doubles): collects the rows of a batch that
 * belong to the seeded training pool of the streaming clustering. */
int mdb_gather_rows(mdb_ctx *ctx, long long n, int D, const double *d_X, long long ldx,
                    const long long *d_src, const long long *d_dst, double *d_out, long long ldo,
                    void *stream);

/* ---- K7: k-means assignment ---------------------------------------------------------
 * Replaces sklearn _labels_inertia / lloyd_iter_chunked_dense(update_centers=False)
 * (cluster/_k_means_lloyd.pyx:168-213) as driven by MiniBatchKMeans.fit_predict at
 * datasetbuilder.py:371-376: label = argmin_j ||c_j||^2 - 2 x.c_j in float64, strict
 * '<' so the lowest centroid index wins ties.  d_rowidx (optional) selects rows of X
 * (a mini-batch) without gathering them.  d_mindist (optional) = ||x - c_label||^2;
 * h_inertia (optional, synchronises) = sum(weight * mindist). */
int mdb_kmeans_assign(mdb_ctx *ctx, long long rows, int D, int K, const double *d_X,
                      long long ldx, const int32_t *d_rowidx, const double *d_centers,
                      int32_t *d_labels, double *d_mindist, const double *d_weights,
                      double *h_inertia, void *stream);

/* Same labels as mdb_kmeans_assign, computed with a tcgen05 (tensor-core) screening pass on
 * split-fp16 operands; rows whose best / second-best gap is inside the error bound are
 * re-decided in float64 (*h_nrecheck = how many).  D <= 112.  Synchronises. */
int mdb_kmeans_assign_tc(mdb_ctx *ctx, long long rows, int D, int K, const double *d_X,
                         long long ldx, const int32_t *d_rowidx, const double *d_centers,
                         int32_t *d_labels, long long *h_nrecheck, float *d_diag, void *stream);
/* d_diag (optional, float [4*rows+1]): per row the best and second-best approximate scores,
 * the best index (int bits) and ||x - mean||^2, then max_j ||c_j - mean||^2 -- what the
 * error bound of the screening pass is validated against in the tests. */

/* Inertia of given labels: sum_i weight_i ||x_i - c_label(i)||^2 (sklearn _inertia_dense).
 * Synchronises. */
int mdb_kmeans_inertia(mdb_ctx *ctx, long long rows, int D, const double *d_X, long long ldx,
                       const int32_t *d_rowidx, const double *d_centers, const int32_t *d_labels,
                       const double *d_weights, double *h_inertia, void *stream);

/* ---- K8: mini-batch centroid update -------------------------------------------------
 * Replaces _minibatch_update_dense (cluster/_k_means_minibatch.pyx:59-108) in two
 * halves so that ranks can all-reduce between them: partial_sums writes, per centre,
 * sum(weight * x) over the batch members (ascending sample order when the batch holds
 * <= 4096 rows) and the summed weight; apply_update performs
 * centre = (centre * weight_sums + sums) / (weight_sums + wsum); weight_sums += wsum
 * for every centre with wsum > 0. */
int mdb_kmeans_partial_sums(mdb_ctx *ctx, long long nbatch, int D, int K, const double *d_X,
                            long long ldx, const int32_t *d_rowidx, const double *d_weights,
                            const int32_t *d_labels, double *d_sums, double *d_wsum, void *stream);
int mdb_kmeans_apply_update(mdb_ctx *ctx, int K, int D, double *d_centers, double *d_weight_sums,
                            const double *d_sums, const double *d_wsum, void *stream);

/* Both halves in one call for a mini-batch of at most 4096 rows held by ONE process, in scikit-learn's own
 * operation order (update_center_dense, cluster/_k_means_minibatch.pyx:59-108: centre * weight_sum, members
 * added in ascending sample order, times 1 / new weight_sum): with equal labels the centres equal sklearn's
 * bit for bit. */
int mdb_kmeans_update_exact(mdb_ctx *ctx, long long nbatch, int D, const double *d_X, long long ldx,
                            const int32_t *d_rowidx, const double *d_weights, const int32_t *d_labels,
                            double *d_centers, double *d_weight_sums, void *stream);

/* One whole mini-batch step (sklearn _mini_batch_step, cluster/_kmeans.py:1601-1697, minus the random
 * reassignment, which stays with the caller): rows h_rowidx [nbatch] (HOST indices into d_X) are assigned to
 * the centres, the centres and weight sums are updated like mdb_kmeans_update_exact; *h_inertia = the batch
 * inertia before the update (summed in a fixed order), *h_nzero = centres whose weight sum is still zero after
 * it.  nbatch <= 4096, D <= 256.  Synchronises once. */
int mdb_kmeans_minibatch_step(mdb_ctx *ctx, long long nbatch, int D, int K, const double *d_X, long long ldx,
                              const int32_t *h_rowidx, double *d_centers, double *d_weight_sums,
                              double *h_inertia, int *h_nzero, void *stream);

/* k-means++ seeding step (sklearn cluster/_kmeans.py:180-279, run by MiniBatchKMeans on
 * its init subset): for ncand (<= 32) candidate rows, d_dist[t][i] = min(d_closest[i],
 * ||x_i - x_cand[t]||^2) and d_pot[t] = sum_i weight_i * d_dist[t][i].  d_closest may be
 * NULL for the first centre.  Rows are addressed through d_rowidx when given (the init
 * subset), and d_cand indexes the same row numbering. */
int mdb_kpp_step(mdb_ctx *ctx, long long n, int D, const double *d_X, long long ldx,
                 const int32_t *d_rowidx, const int32_t *d_cand, int ncand,
                 const double *d_closest, const double *d_weights, double *d_dist, double *d_pot,
                 void *stream);

/* The whole k-means++ seeding (sklearn cluster/_kmeans.py:180-279, unit sample weights) on the device,
 * for nbatch independent runs at once (MiniBatchKMeans' n_init seedings share nothing but X): K centres
 * out of the n rows X[rowidx[b]] (d_rowidx [nbatch][n], or NULL = X itself), n_local_trials candidates
 * per step.  h_rand [nbatch][(K-1) * n_local_trials] holds the uniform variates in the order sklearn
 * draws them, h_first_center [nbatch] the first centres; d_indices [nbatch][K] receives the chosen
 * positions (0..n-1).  Synchronises. */
int mdb_kmeans_plusplus(mdb_ctx *ctx, int nbatch, long long n, int D, const double *d_X, long long ldx,
                        const int32_t *d_rowidx, int K, int n_local_trials, const double *h_rand,
                        const int32_t *h_first_center, int32_t *d_indices, void *stream);

/* ---- per-cluster selection on stored labels (datasetbuilder.py:378-383) ------------------------
 * `idx = np.where(labels == k)[0]; choice(idx, n_each)` needs, per cluster, its member count and then
 * the row of member number p.  mdb_label_hist adds the member counts of d_labels [rows] to d_hist [K];
 * mdb_pick_rows resolves npicks queries d_picks [npicks][4] = (first row, end row, cluster, p): the row
 * (index into d_labels) of the p-th member of that cluster inside [first row, end row), ascending row
 * order, or -1 when the range holds fewer members. */
int mdb_label_hist(mdb_ctx *ctx, long long rows, int K, const int32_t *d_labels, int32_t *d_hist,
                   void *stream);
int mdb_pick_rows(mdb_ctx *ctx, const int32_t *d_labels, int npicks, const long long *d_picks,
                  long long *d_rows, void *stream);

/* ---- bond orders for the dump-only key ---------------------------------------------
 * Replaces `mol.PerceiveBondOrders()` (reference detect.py:279-280; Open Babel 3.1.x
 * OBMol::PerceiveBondOrders, third-party) for the C/H/O(/N) subset restated in oracle/perceive.py:
 * hybridisation from the average bond angle, the carboxylic / CO2 / allene patterns of bondtyp.txt
 * and the coded carbonyl rule, then the electronegativity-ordered double / triple assignment.  The
 * ring passes are not restated (independent cycles are counted, mdb_perceive_stats).  PARITY
 * UNPINNED against Open Babel itself.
 * d_ell [F][N][max_bonds] final rows (after the cleanup), d_labels [F][N] molecule roots or NULL (no
 * ring statistics); outputs, either may be NULL: d_orders [F][N][max_bonds] one byte per ELL slot
 * (0 = empty), d_keys [F][N] the 64-bit key (element, degree, sorted orders).  With option
 * "perceive_bond_orders" = 1, mdb_analyze_frames* return these keys instead of the all-single ones. */
int mdb_perceive_bond_orders(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos,
                             const double *h_cell, const uint8_t *d_types, int periodic,
                             const int32_t *d_ell, const int32_t *d_labels, uint8_t *d_orders,
                             uint64_t *d_keys, void *stream);
/* Independent cycles (bonds - atoms + molecules, summed over frames) seen by the perception since the
 * last analyze / perceive call, which needs d_labels for it (synchronises). */
int mdb_perceive_stats(mdb_ctx *ctx, long long *n_rings);

/* Diagnostics of the last bond batch: counts summed over frames (synchronises).  A device-side error
 * flag raised by any kernel since the last call is returned (non-zero) and cleared. */
int mdb_bond_stats(mdb_ctx *ctx, long long *n_violators, long long *n_exact_tests,
                   long long *n_overflow);

/* ---- LAMMPS text readers (host side) -------------------------------------------------
 * Replace the per-line Python parsing of DetectDump._readN / readcrd (detect.py:151-194,
 * 294-345), DetectBond._readN / readatombondtype / readmolecule (detect.py:56-88, 105-119,
 * 137-142) and the frame slicing of DatasetBuilder.lineiter (datasetbuilder.py:616-638).
 * kind 0 = dump, 1 = fix reaxff/bonds table.  Opening scans the first two frames like
 * _readN (atom count, lines per frame, id/type/x/y/z columns, per-id LAMMPS types). */
typedef struct mdb_reader mdb_reader;
int mdb_reader_open(int kind, const char *const *paths, int npaths, int nthreads, mdb_reader **out);
void mdb_reader_close(mdb_reader *r);
int mdb_reader_info(const mdb_reader *r, int *natoms, long *steplinenum, int *cols5);
int mdb_reader_types(const mdb_reader *r, int32_t *types /* [N] LAMMPS type by id-1 */);
int mdb_reader_rewind(mdb_reader *r);
/* Next up to max_frames kept frames (every stride-th frame of each file).  dump: h_pos
 * [max_frames][N][3] in id order, h_cell [max_frames][9].  bonds: h_nbr [max_frames][N][width]
 * neighbours in file order (0-based, -1 padded), h_bo same shape (optional). */
int mdb_reader_next(mdb_reader *r, int stride, int max_frames, double *h_pos, double *h_cell,
                    int32_t *h_nbr, double *h_bo, int width, int *nread);
/* Raw text of the next kept frames, packed back to back into h_buf (page-locked), for the device
 * parsers below: frame k occupies [frame_begin[k], frame_end[k]).  Dump readers: its atom lines
 * start at atoms_begin[k] (-1: no "ITEM: ATOMS" section found, parse it with
 * mdb_reader_parse_buffer) and h_cell [k][9] is its cell.  Bonds readers: atoms_begin = frame_begin,
 * h_cell is not written.  Stops early when the next frame would not fit cap. */
int mdb_reader_next_raw(mdb_reader *r, int stride, int max_frames, char *h_buf, size_t cap,
                        long long *frame_begin, long long *atoms_begin, long long *frame_end,
                        double *h_cell, int *nread);
/* Device-side parse of the atom lines delivered by mdb_reader_next_raw (after the caller copied
 * h_buf to d_text): what the per-line split()/float() of DetectDump.readcrd (reference
 * detect.py:294-345) produces, positions sorted by id into d_pos [F][N][3].  Decimal strings are
 * converted exactly (significand <= 2^53 and |exponent| <= 22, one correctly rounded multiply or
 * divide); a frame holding anything else -- longer numbers, inf/nan, a malformed line, a wrong
 * atom count, a type that differs from the first frame -- gets d_status[k] != 0 and must be parsed
 * on the host.  cols5 = column indices of id, type, x, y, z (mdb_reader_info); d_types = LAMMPS
 * types by id-1 (mdb_reader_types); d_count [F] is scratch. */
int mdb_parse_dump_text(mdb_ctx *ctx, const char *d_text, int nframes, int natoms,
                        const long long *d_atoms_begin, const long long *d_frame_end,
                        long long max_frame_bytes, const int *cols5, const int32_t *d_types,
                        double *d_pos, int *d_status, int *d_count, void *stream);
/* The same for fix reaxff/bonds tables (mdb_reader_next_raw on a bonds reader; frame k =
 * [frame_begin[k], frame_end[k])): neighbours in file order, 0-based, -1 padded, into d_nbr
 * [F][N][width], bond orders into d_bo (same shape, optional) -- DetectBond.readatombondtype /
 * readmolecule, reference detect.py:105-119, 137-142.  d_status[k] != 0: parse frame k on the host. */
int mdb_parse_bonds_text(mdb_ctx *ctx, const char *d_text, int nframes, int natoms,
                         const long long *d_frame_begin, const long long *d_frame_end,
                         long long max_frame_bytes, int width, int32_t *d_nbr, double *d_bo,
                         int *d_status, int *d_count, void *stream);
/* One frame already in memory (the lines a per-frame drop-in method receives). */
int mdb_reader_parse_buffer(const mdb_reader *r, const char *buf, size_t len, double *h_pos,
                            double *h_cell, int32_t *h_nbr, double *h_bo, int width);

/* Output side: many ASE "simple xyz" files in one call (the per-structure ase.io.write of reference
 * datasetbuilder.py:577-585): "<n>\n\n" then "%-2s %22.15f %22.15f %22.15f\n" per atom.  File k holds
 * atoms [offsets[k], offsets[k+1]) of the flat arrays; symbols4 = 4-byte NUL-padded strings. */
int mdb_write_xyz_batch(int nfiles, const char *const *paths, const long long *offsets,
                        const char *symbols4, const double *h_pos, int nthreads);

/* Gaussian input files of many structures at once: the reference's _convertgjf (datasetbuilder.py:468-521)
 * with detect_multiplicity (:446-466), byte for byte.  File k holds atoms [offsets[k], offsets[k+1]) of the
 * flat arrays (symbols4: 4-byte NUL-padded symbols, znum: atomic numbers, pos: [n][3]); its molecules are the
 * consecutive runs of sizes mol_sizes[mol_offsets[k] .. mol_offsets[k+1]).  keywords: the qmkeywords sections
 * (several -> %chk lines and --link1-- sections); fragment != 0 -> "guess=fragment=K" input for structures of
 * more than one molecule.  Written by nthreads threads (<= 0: hardware concurrency). */
int mdb_write_gjf_batch(int nfiles, const char *const *paths, const long long *offsets,
                        const char *symbols4, const int32_t *znum, const double *pos,
                        const long long *mol_offsets, const int32_t *mol_sizes, int nkeywords,
                        const char *const *keywords, int fragment, int nthreads);

/* DFMA throughput of the device in TFLOP/s, measured by a register-resident fused multiply-add kernel
 * (bench.py's denominator for the FP64-bound descriptor kernels).  Synchronises. */
int mdb_fp64_peak(mdb_ctx *ctx, double *h_tflops, void *stream);

/* Per-launch CUDA-event timing on the launching stream (bench.py's live roofline
 * measurement).  mdb_profile_read synchronises and writes "name count total_ms" lines
 * accumulated since the previous read into buf. */
int mdb_profile_enable(mdb_ctx *ctx, int on);
int mdb_profile_read(mdb_ctx *ctx, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* MDB_B200_H */
