// descriptor.cu -- K5 composition pass and K6 Coulomb-matrix eigen-spectrum descriptor.
//
// Replaces, per target atom (reference datasetbuilder.py:305-349 and the parent-side assembly
// :236-276):
//   sphere cut   distances = get_distances(a, all, mic=True); members = distances < cutoff
//                (ASE general_find_mic; for an orthorhombic cell the 28 images decouple per axis:
//                 p = ((D/L) % 1.0) * L, v = the smaller of p and p - L),
//   Counter      element composition of the sphere,
//   descriptor   ascending eigenvalues of M_ij = Z_i Z_j / r_ij (i != j), M_ii = Z_i^2.4 / 2,
//   feature row  sort(eigenvalues U {Z_el^2.4/2 repeated (max_el - count_el) times}).
//
// Membership is bit-exact (float screen on the 5 A cell list, borderline distances re-evaluated
// in double with the oracle's operation order); the spectrum is computed in double by Householder
// tridiagonalisation followed by implicit-shift QL (one thread per matrix), well inside the 1e-5
// relative tolerance of the task.  Targets are bucketed by cluster size; up to 64 atoms the matrix
// lives in registers (one column per lane: k_tridiag_reg for <= 32 atoms on LPM lanes of a warp,
// k_tridiag_reg2 for <= 64 on a warp pair), above that in shared memory (k_tridiag_blk, one block
// per matrix).
// General (triclinic) cells take the grid-unit cell list and the minimum image through the cell matrix
// (mic_general); diagonal cells keep the per-axis fast path.
#include <algorithm>

#include "common.cuh"

struct DescParams {
    const float4 *rec;
    const uint32_t *ccell;
    const uint32_t *cr;  // per atom (cell id << rank_bits | rank within cell)
    const int *cell_start;
    const FrameGeom *geom;
    const double *pos;
    const uint8_t *types;
    const int *targets;  // batch-local atom-frame index f*N + i, or nullptr = every atom-frame
    long long ntargets;
    int N;
    int periodic;
    double cutoff;
    float cut2f, margin;
    int nt;
    double diag[MDB_MAXT];
    int Z[MDB_MAXT];
    BondStats *stats;
    // optional sphere membership lists [ntargets][memcap] (-1 padded, any order): written by the
    // composition pass (members_out), read instead of a second cell scan by the descriptor pass
    const int *members;
    int *members_out;
    int memcap;
    int rec_mode;  // 1: records hold wrapped Angstrom coordinates (diagonal cells), 0: grid units (general cells)
};

// ASE general_find_mic along one axis of an orthorhombic cell (oracle.c: ase_axis_general)
__device__ __forceinline__ double ase_axis_general(double D, double L) {
    double f = __ddiv_rn(D, L);
    double m = fmod(f, 1.0);
    if (m != 0.0) {
        if (m < 0.0) m = __dadd_rn(m, 1.0);
    } else
        m = 0.0;
    double p = __dmul_rn(m, L);
    double q = __dsub_rn(p, L);
    return (__dmul_rn(q, q) < __dmul_rn(p, p)) ? q : p;
}

// True minimum image of (x, y, z) in a general cell: fractional rounding, then the 26 neighbour images of
// that candidate (enough for LAMMPS cells, whose tilt factors are bounded by half a box length).  This is
// what ASE's general_find_mic returns (Minkowski reduction + image search) up to the last
// bits of the length; only a distance within ~1e-13 A of the cutoff could be decided differently.
// ortho = lattice vectors as columns (frac -> cart), frac = its inverse (FrameGeom).
// (Kept inline on purpose: as a __noinline__ call from the unrolled register kernels it returned garbage on
// sm_100a / CUDA 12.9 -- tools/tric_debug.py, the stand-alone tools/micro/mic_test.cu is fine.)
__device__ __forceinline__ void mic_general(const double *__restrict__ ortho, const double *__restrict__ frac, double &x, double &y,
                                         double &z) {
    double f0 = frac[0] * x + frac[1] * y + frac[2] * z;
    double f1 = frac[3] * x + frac[4] * y + frac[5] * z;
    double f2 = frac[6] * x + frac[7] * y + frac[8] * z;
    f0 -= rint(f0);
    f1 -= rint(f1);
    f2 -= rint(f2);
    const double bx = ortho[0] * f0 + ortho[1] * f1 + ortho[2] * f2;
    const double by = ortho[3] * f0 + ortho[4] * f1 + ortho[5] * f2;
    const double bz = ortho[6] * f0 + ortho[7] * f1 + ortho[8] * f2;
    double best = bx * bx + by * by + bz * bz;
    x = bx;
    y = by;
    z = bz;
    for (int h = -1; h <= 1; ++h)
        for (int k = -1; k <= 1; ++k)
            for (int l = -1; l <= 1; ++l) {
                if (h == 0 && k == 0 && l == 0) continue;
                const double cx = bx + ortho[0] * h + ortho[1] * k + ortho[2] * l;
                const double cy = by + ortho[3] * h + ortho[4] * k + ortho[5] * l;
                const double cz = bz + ortho[6] * h + ortho[7] * k + ortho[8] * l;
                const double d2 = cx * cx + cy * cy + cz * cz;
                if (d2 < best) {
                    best = d2;
                    x = cx;
                    y = cy;
                    z = cz;
                }
            }
}

// displacement target -> member as the reference sees it (minimum image when periodic)
__device__ __forceinline__ void member_delta(int periodic, const FrameGeom &G, const double *__restrict__ pos, int j,
                                             const double pa[3], double out[3]) {
    out[0] = pos[(size_t)j * 3] - pa[0];
    out[1] = pos[(size_t)j * 3 + 1] - pa[1];
    out[2] = pos[(size_t)j * 3 + 2] - pa[2];
    if (!periodic) return;
    if (G.is_ortho) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double L = G.ortho[c * 4];
            out[c] -= L * rint(out[c] / L);
        }
    } else
        mic_general(G.ortho, G.frac, out[0], out[1], out[2]);
}

// exact (double) membership decision of a borderline pair: the diagonal-cell form follows ASE per axis, the
// general form takes the minimum image through the cell matrix; each scan carries only its own
__device__ __forceinline__ bool sphere_exact(const DescParams &P, const FrameGeom &G, const double *pos, int a, int j) {
    double v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double D = __dsub_rn(pos[(size_t)j * 3 + c], pos[(size_t)a * 3 + c]);
        v[c] = P.periodic ? ase_axis_general(D, G.ortho[c * 4]) : D;
    }
    double d2 = __dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2]));
    return __dsqrt_rn(d2) < P.cutoff;
}
__device__ __forceinline__ bool sphere_exact_general(const DescParams &P, const FrameGeom &G, const double *pos, int a, int j) {
    if (!P.periodic || G.is_ortho) return sphere_exact(P, G, pos, a, j);
    double v[3];
    const double pa[3] = {pos[(size_t)a * 3], pos[(size_t)a * 3 + 1], pos[(size_t)a * 3 + 2]};
    member_delta(1, G, pos, j, pa, v);
    const double d2g = __dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2]));
    return __dsqrt_rn(d2g) < P.cutoff;
}

// The same scan for general (triclinic) cells: the records hold grid units (fractional coordinate x cells per
// axis), a neighbour across the periodic boundary is shifted by the cell count of that axis, the float screen
// maps the grid delta to cartesian through G.h, and borderline pairs go through the double minimum image.
template <class Visit>
__device__ __forceinline__ void sphere_scan_general(const DescParams &P, int f, int a, Visit visit) {
    const int lane = threadIdx.x & 31;
    const FrameGeom &G = P.geom[f];
    const size_t fbase = (size_t)f * P.N;
    const int cbase = (int)G.cell_base;
    const uint32_t v = P.cr[fbase + a];
    const int slot = P.cell_start[cbase + (int)(v >> G.rank_bits)] + (int)(v & ((1u << G.rank_bits) - 1u));
    const float4 me = P.rec[slot];
    const uint32_t cc = P.ccell[slot];
    const int c3[3] = {(int)(cc & 1023u), (int)((cc >> 10) & 1023u), (int)(cc >> 20)};
    const int n3[3] = {G.n[0], G.n[1], G.n[2]};
    float h[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) h[k] = G.h[k];
    const double *pos = P.pos + fbase * 3;
    for (int dz = -1; dz <= 1; ++dz) {
        int z = c3[2] + dz;
        float sz = 0.f;
        if (G.reduce[2]) {
            if (dz != 0) continue;
            z = 0;
        } else if (z < 0 || z >= n3[2]) {
            if (!P.periodic) continue;
            sz = z < 0 ? -(float)n3[2] : (float)n3[2];
            z += z < 0 ? n3[2] : -n3[2];
        }
        for (int dy = -1; dy <= 1; ++dy) {
            int y = c3[1] + dy;
            float sy = 0.f;
            if (G.reduce[1]) {
                if (dy != 0) continue;
                y = 0;
            } else if (y < 0 || y >= n3[1]) {
                if (!P.periodic) continue;
                sy = y < 0 ? -(float)n3[1] : (float)n3[1];
                y += y < 0 ? n3[1] : -n3[1];
            }
            const int rowbase = cbase + (z * n3[1] + y) * n3[0];
            for (int dx = -1; dx <= 1; ++dx) {
                int x = c3[0] + dx;
                float sx = 0.f;
                if (G.reduce[0]) {
                    if (dx != 0) continue;
                    x = 0;
                } else if (x < 0 || x >= n3[0]) {
                    if (!P.periodic) continue;
                    sx = x < 0 ? -(float)n3[0] : (float)n3[0];
                    x += x < 0 ? n3[0] : -n3[0];
                }
                const int lo = P.cell_start[rowbase + x], hi = P.cell_start[rowbase + x + 1];
                for (int q0 = lo; q0 < hi; q0 += 32) {
                    const int q = q0 + lane;
                    if (q < hi) {
                        const float4 o = P.rec[q];
                        float du = o.x - me.x + sx, dv = o.y - me.y + sy, dw = o.z - me.z + sz;
                        if (G.reduce[0]) du -= rintf(du);
                        if (G.reduce[1]) dv -= rintf(dv);
                        if (G.reduce[2]) dw -= rintf(dw);
                        const float ex = h[0] * du + h[1] * dv + h[2] * dw;
                        const float ey = h[3] * du + h[4] * dv + h[5] * dw;
                        const float ez = h[6] * du + h[7] * dv + h[8] * dw;
                        const float d2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
                        if (d2 <= P.cut2f + P.margin) {
                            const unsigned ow = (unsigned)__float_as_int(o.w);
                            const int j = (int)(ow & MDB_IDX_MASK);
                            bool in = true;
                            if (!(d2 < P.cut2f - P.margin)) {
                                in = sphere_exact_general(P, G, pos, a, j);
                                atomicAdd(&P.stats->exact_tests, 1ull);
                            }
                            if (in) visit(j, (int)(ow >> MDB_IDX_BITS));
                        }
                    }
                }
            }
        }
    }
}

// Warp-cooperative scan of the 3x3x3 cell neighbourhood of target atom `a` of frame f.
// Calls visit(j, type_j) on the lanes that own a member; every lane must call this function.
template <class Visit>
__device__ __forceinline__ void sphere_scan(const DescParams &P, int f, int a, Visit visit) {
    const int lane = threadIdx.x & 31;
    const FrameGeom &G = P.geom[f];
    const size_t fbase = (size_t)f * P.N;
    const int cbase = (int)G.cell_base;
    const uint32_t v = P.cr[fbase + a];
    const int slot = P.cell_start[cbase + (int)(v >> G.rank_bits)] + (int)(v & ((1u << G.rank_bits) - 1u));
    const float4 me = P.rec[slot];
    const uint32_t cc = P.ccell[slot];
    const int c3[3] = {(int)(cc & 1023u), (int)((cc >> 10) & 1023u), (int)(cc >> 20)};
    const int n3[3] = {G.n[0], G.n[1], G.n[2]};
    const float per[3] = {(float)G.ortho[0], (float)G.ortho[4], (float)G.ortho[8]};
    const float hp[3] = {0.5f * per[0], 0.5f * per[1], 0.5f * per[2]};
    const double *pos = P.pos + fbase * 3;
    for (int dz = -1; dz <= 1; ++dz) {
        int z = c3[2] + dz;
        if (G.reduce[2]) {
            if (dz != 0) continue;
            z = 0;
        } else if (z < 0 || z >= n3[2]) {
            if (!P.periodic) continue;
            z += z < 0 ? n3[2] : -n3[2];
        }
        for (int dy = -1; dy <= 1; ++dy) {
            int y = c3[1] + dy;
            if (G.reduce[1]) {
                if (dy != 0) continue;
                y = 0;
            } else if (y < 0 || y >= n3[1]) {
                if (!P.periodic) continue;
                y += y < 0 ? n3[1] : -n3[1];
            }
            const int rowbase = cbase + (z * n3[1] + y) * n3[0];
            // x: contiguous run plus at most one wrapped cell
            int xa, xb, xw = -1;
            if (G.reduce[0]) {
                xa = xb = 0;
            } else {
                xa = c3[0] - 1;
                xb = c3[0] + 1;
                if (xa < 0) {
                    xa = 0;
                    if (P.periodic) xw = n3[0] - 1;
                }
                if (xb >= n3[0]) {
                    xb = n3[0] - 1;
                    if (P.periodic) xw = 0;
                }
            }
            for (int seg = 0; seg < 2; ++seg) {
                int lo, hi;
                if (seg == 0) {
                    lo = P.cell_start[rowbase + xa];
                    hi = P.cell_start[rowbase + xb + 1];
                } else {
                    if (xw < 0) break;
                    lo = P.cell_start[rowbase + xw];
                    hi = P.cell_start[rowbase + xw + 1];
                }
                for (int q0 = lo; q0 < hi; q0 += 32) {
                    const int q = q0 + lane;
                    if (q < hi) {
                        const float4 o = P.rec[q];
                        float dx = o.x - me.x, dy2 = o.y - me.y, dz2 = o.z - me.z;
                        if (P.periodic) {
                            dx += dx > hp[0] ? -per[0] : (dx < -hp[0] ? per[0] : 0.f);
                            dy2 += dy2 > hp[1] ? -per[1] : (dy2 < -hp[1] ? per[1] : 0.f);
                            dz2 += dz2 > hp[2] ? -per[2] : (dz2 < -hp[2] ? per[2] : 0.f);
                        }
                        const float d2 = fmaf(dz2, dz2, fmaf(dy2, dy2, dx * dx));
                        if (d2 <= P.cut2f + P.margin) {
                            const unsigned ow = (unsigned)__float_as_int(o.w);
                            const int j = (int)(ow & MDB_IDX_MASK);
                            bool in = true;
                            if (!(d2 < P.cut2f - P.margin)) {
                                in = sphere_exact(P, G, pos, a, j);
                                atomicAdd(&P.stats->exact_tests, 1ull);
                            }
                            if (in) visit(j, (int)(ow >> MDB_IDX_BITS));
                        }
                    }
                }
            }
        }
    }
}

// Flattened form of sphere_scan for the composition / member kernels (diagonal cells): ncu had the serial form at
// 1,224 warp instructions per target -- nine (dy, dz) rows handled one after the other, each with its own bounds
// arithmetic and a 32-lane pass over ~19 candidates.  Here lanes 0..17 work out the 18 possible segments (nine
// x-runs + nine wrapped cells) side by side, a warp prefix sum numbers all candidates of the target, and the lanes
// then walk ONE list of ~170 candidates, finding the segment of a candidate by a 5-step search over the prefix held
// in the lanes.  Same candidates, same float screen, same exact re-check: the membership is unchanged.
template <class Visit>
__device__ __forceinline__ void sphere_scan_flat(const DescParams &P, int f, int a, Visit visit) {
    const int lane = threadIdx.x & 31;
    const FrameGeom &G = P.geom[f];
    const size_t fbase = (size_t)f * P.N;
    const int cbase = (int)G.cell_base;
    const uint32_t v = P.cr[fbase + a];
    const int slot = P.cell_start[cbase + (int)(v >> G.rank_bits)] + (int)(v & ((1u << G.rank_bits) - 1u));
    const float4 me = P.rec[slot];
    const uint32_t cc = P.ccell[slot];
    const int c3[3] = {(int)(cc & 1023u), (int)((cc >> 10) & 1023u), (int)(cc >> 20)};
    const int n3[3] = {G.n[0], G.n[1], G.n[2]};
    const float per[3] = {(float)G.ortho[0], (float)G.ortho[4], (float)G.ortho[8]};
    const float hp[3] = {0.5f * per[0], 0.5f * per[1], 0.5f * per[2]};
    const double *pos = P.pos + fbase * 3;
    // segment of this lane: row r = lane / 2 (dy = r % 3 - 1, dz = r / 3 - 1), part 0 = contiguous x-run, 1 = wrapped cell
    int lo = 0, hi = 0;
    if (lane < 18) {
        const int r = lane >> 1, part = lane & 1;
        const int dz = r / 3 - 1, dy = r % 3 - 1;
        bool ok = true;
        int z = c3[2] + dz, y = c3[1] + dy;
        if (G.reduce[2]) {
            ok = dz == 0;
            z = 0;
        } else if (z < 0 || z >= n3[2]) {
            ok = P.periodic != 0;
            z += z < 0 ? n3[2] : -n3[2];
        }
        if (G.reduce[1]) {
            ok = ok && dy == 0;
            y = 0;
        } else if (y < 0 || y >= n3[1]) {
            ok = ok && P.periodic != 0;
            y += y < 0 ? n3[1] : -n3[1];
        }
        if (ok) {
            const int rowbase = cbase + (z * n3[1] + y) * n3[0];
            int xa, xb, xw = -1;
            if (G.reduce[0]) {
                xa = xb = 0;
            } else {
                xa = c3[0] - 1;
                xb = c3[0] + 1;
                if (xa < 0) {
                    xa = 0;
                    if (P.periodic) xw = n3[0] - 1;
                }
                if (xb >= n3[0]) {
                    xb = n3[0] - 1;
                    if (P.periodic) xw = 0;
                }
            }
            if (part == 0) {
                lo = P.cell_start[rowbase + xa];
                hi = P.cell_start[rowbase + xb + 1];
            } else if (xw >= 0) {
                lo = P.cell_start[rowbase + xw];
                hi = P.cell_start[rowbase + xw + 1];
            }
        }
    }
    // inclusive prefix of the segment lengths over the lanes
    int end = hi - lo;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, end, o);
        if (lane >= o) end += t;
    }
    const int total = __shfl_sync(0xffffffffu, end, 31);
    const int base = lo - (end - (hi - lo));  // candidate p of this segment sits at rec[base + p]
    for (int p0 = 0; p0 < total; p0 += 32) {
        const int p = p0 + lane;
        // smallest segment s with end[s] > p (18 segments: lanes 0..17; lanes >= 18 repeat end[17] = total)
        int sidx = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int e = __shfl_sync(0xffffffffu, end, sidx + step - 1);
            if (e <= p) sidx += step;
        }
        const int sb = __shfl_sync(0xffffffffu, base, sidx);
        if (p < total) {
            const float4 o = P.rec[sb + p];
            float dx = o.x - me.x, dy2 = o.y - me.y, dz2 = o.z - me.z;
            if (P.periodic) {
                dx += dx > hp[0] ? -per[0] : (dx < -hp[0] ? per[0] : 0.f);
                dy2 += dy2 > hp[1] ? -per[1] : (dy2 < -hp[1] ? per[1] : 0.f);
                dz2 += dz2 > hp[2] ? -per[2] : (dz2 < -hp[2] ? per[2] : 0.f);
            }
            const float d2 = fmaf(dz2, dz2, fmaf(dy2, dy2, dx * dx));
            if (d2 <= P.cut2f + P.margin) {
                const unsigned ow = (unsigned)__float_as_int(o.w);
                const int j = (int)(ow & MDB_IDX_MASK);
                bool in = true;
                if (!(d2 < P.cut2f - P.margin)) {
                    in = sphere_exact(P, G, pos, a, j);
                    atomicAdd(&P.stats->exact_tests, 1ull);
                }
                if (in) visit(j, (int)(ow >> MDB_IDX_BITS));
            }
        }
    }
}

// The same visit over a stored membership list (written by k_composition) instead of the cell scan.
template <class Visit>
__device__ __forceinline__ void member_scan(const DescParams &P, long long ord, Visit visit) {
    const int lane = threadIdx.x & 31;
    const int *m = P.members + (size_t)ord * P.memcap;
    for (int k0 = 0; k0 < P.memcap; k0 += 32) {
        const int k = k0 + lane;
        const int j = k < P.memcap ? m[k] : -1;
        if (j >= 0) visit(j, (int)P.types[j]);
        if (__all_sync(0xffffffffu, j < 0)) break;
    }
}
template <class Visit>
__device__ __forceinline__ void gather_scan(const DescParams &P, int f, int a, long long ord, Visit visit) {
    if (P.members)
        member_scan(P, ord, visit);
    else
        sphere_scan(P, f, a, visit);
}

// ---------------------------------------------------------------------------------------------
// K5: element composition of every target's sphere; running per-element maximum
// ---------------------------------------------------------------------------------------------
template <bool GENERAL>
__global__ void __launch_bounds__(256) k_composition(const __grid_constant__ DescParams P, int *__restrict__ counts,
                                                     int *__restrict__ maxcount) {
    __shared__ int s_cnt[8][MDB_MAXT];
    __shared__ int s_max[MDB_MAXT];
    __shared__ int s_n[8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < MDB_MAXT) s_max[threadIdx.x] = 0;
    __syncthreads();
    const long long t = (long long)blockIdx.x * 8 + w;
    if (t < P.ntargets) {
        const int g = P.targets ? P.targets[t] : (int)t;
        const int f = g / P.N, a = g - f * P.N;
        if (lane < MDB_MAXT) s_cnt[w][lane] = 0;
        if (lane == 0) s_n[w] = 0;
        __syncwarp();
        int *mo = P.members_out ? P.members_out + (size_t)t * P.memcap : nullptr;
        auto visit = [&](int j, int tj) {
            atomicAdd(&s_cnt[w][tj], 1);
            if (mo) {
                const int k = atomicAdd(&s_n[w], 1);
                if (k < P.memcap) mo[k] = j;
            }
        };
        if (GENERAL)
            sphere_scan_general(P, f, a, visit);
        else
            sphere_scan_flat(P, f, a, visit);
        __syncwarp();
        if (mo) {
            const int n = s_n[w];
            if (n > P.memcap && lane == 0) atomicExch(&P.stats->err, MDB_ERR_BOND_OVERFLOW);
            for (int k = min(n, P.memcap) + lane; k < P.memcap; k += 32) mo[k] = -1;
        }
        if (lane < P.nt) {
            const int c = s_cnt[w][lane];
            counts[t * P.nt + lane] = c;
            atomicMax(&s_max[lane], c);
        }
    }
    __syncthreads();
    if (threadIdx.x < P.nt && s_max[threadIdx.x] > 0) atomicMax(&maxcount[threadIdx.x], s_max[threadIdx.x]);
}

// sphere membership lists (ascending atom index, includes the target): the `np.where(distances <
// cutoff)` of the output cut, datasetbuilder.py:554-557.  One warp per target, cap <= 1024.
template <bool GENERAL>
__global__ void __launch_bounds__(256) k_members(const __grid_constant__ DescParams P, int cap, int *__restrict__ members,
                                                 int *__restrict__ nmem) {
    extern __shared__ int s_mem[];  // [8][cap]
    __shared__ int s_n[8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * 8 + w;
    if (t >= P.ntargets) return;
    const int g = P.targets ? P.targets[t] : (int)t;
    const int f = g / P.N, a = g - f * P.N;
    int *mine = s_mem + w * cap;
    if (lane == 0) s_n[w] = 0;
    __syncwarp();
    auto visit = [&](int j, int) {
        const int k = atomicAdd(&s_n[w], 1);
        if (k < cap) mine[k] = j;
    };
    if (GENERAL)
        sphere_scan_general(P, f, a, visit);
    else
        sphere_scan_flat(P, f, a, visit);
    __syncwarp();
    int n = s_n[w];
    if (lane == 0) nmem[t] = n;
    if (n > cap) n = cap;
    // rank sort: each element's position = number of smaller elements (indices are distinct)
    for (int k = lane; k < n; k += 32) {
        const int v = mine[k];
        int r = 0;
        for (int q = 0; q < n; ++q) r += mine[q] < v;
        members[t * cap + r] = v;
    }
    for (int k = n + lane; k < cap; k += 32) members[t * cap + k] = -1;
}

// ---------------------------------------------------------------------------------------------
// K6a, clusters of 65..160 atoms: one thread block per matrix, matrix in shared memory.  Two threads
// share a row (even / odd columns), so the matrix-vector product and the rank-2 update are split
// over 2 n threads; the row stride NMAX + 2 keeps the accesses of a half-warp on distinct banks.
// Same step as the register kernels below (one fused reduction of (sigma, x.T x) per step); output is
// the diagonal d[0..n) and the off-diagonal e[0..n-1) in a scratch laid out [k][slot] so that the QL
// kernel reads it coalesced.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x);
__device__ __forceinline__ double fast_rsqrt(double x);

// Members arrive in the order of the cell scan (atomic slots); the matrix is built in ascending
// atom index, the order of the reference's np.where (datasetbuilder.py:315), so that the spectrum
// does not depend on how the cell list happened to be filled.  Entry e (one per participating
// thread, e < n) moves to the rank of its atom index; `sync` separates the read and write phases.
template <class Sm, class Sync>
__device__ __forceinline__ void canonical_member_order(Sm *S, int n, int e, Sync sync) {
    const bool has = e < n;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    int zz = 0, tt = 0, rank = 0;
    if (has) {
        v0 = S->vec[e][0];
        v1 = S->vec[e][1];
        v2 = S->vec[e][2];
        zz = S->z[e];
        tt = S->t[e];
        const int ii = S->idx[e];
        for (int q = 0; q < n; ++q) rank += S->idx[q] < ii;
    }
    sync();
    if (has) {
        S->vec[rank][0] = v0;
        S->vec[rank][1] = v1;
        S->vec[rank][2] = v2;
        S->z[rank] = zz;
        S->t[rank] = tt;
    }
    sync();
}

// Coulomb matrix (datasetbuilder.py:341-348) of the n members held in S->vec / S->z / S->t into the staging copy
// S->A, every pair once, by `nl` lanes (this one is lane `gl`).  Rows are folded -- row i together with row
// n - i gives n pairs -- so that the lanes stay busy without any index arithmetic per pair.  Both vectors lie
// within +-L/2 of the target, so one conditional shift per axis gives the minimum image of a pair in a diagonal
// cell, and none is needed when the sphere is small against the box (4 r_max < L, decided once per matrix);
// general cells go through mic_general.  The reciprocal distance is a MUFU seed + two Newton steps.
template <class Sm>
__device__ __forceinline__ void coulomb_pairs(const DescParams &P, Sm *S, int n, int gl, int nl) {
    if (n < 2) return;
    const bool general = P.periodic && !S->gp->is_ortho;
    const double L0 = S->L[0], L1 = S->L[1], L2 = S->L[2];
    const double hL0 = 0.5 * L0, hL1 = 0.5 * L1, hL2 = 0.5 * L2;
    // largest member distance from the target (every lane scans the list: n broadcasts, no reduction needed)
    bool wrap = false;
    if (P.periodic && !general) {
        double r2max = 0.0;
        for (int k = 0; k < n; ++k) {
            const double a = S->vec[k][0], b = S->vec[k][1], c = S->vec[k][2];
            r2max = fmax(r2max, a * a + b * b + c * c);
        }
        const double lmin = fmin(L0, fmin(L1, L2));
        wrap = !(16.0 * r2max < lmin * lmin);
    }
    for (int i = 1; i <= n / 2; ++i) {
        const int rb = n - i;
        const int cnt = (rb == i) ? i : n;
        for (int t = gl; t < cnt; t += nl) {
            const int a = t < i ? i : rb;
            const int b = t < i ? t : t - i;
            double dx = S->vec[b][0] - S->vec[a][0], dy = S->vec[b][1] - S->vec[a][1], dz = S->vec[b][2] - S->vec[a][2];
            if (wrap) {
                dx += dx > hL0 ? -L0 : (dx < -hL0 ? L0 : 0.0);
                dy += dy > hL1 ? -L1 : (dy < -hL1 ? L1 : 0.0);
                dz += dz > hL2 ? -L2 : (dz < -hL2 ? L2 : 0.0);
            } else if (general)
                mic_general(S->gp->ortho, S->gp->frac, dx, dy, dz);
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double mv = r2 > 0.0 ? (double)(S->z[a] * S->z[b]) * fast_rsqrt(r2) : 0.0;  // inf -> 0 (:347)
            S->A[a][b] = mv;
            S->A[b][a] = mv;
        }
    }
}

template <int NMAX>
struct BlkSmem {
    double A[NMAX][NMAX + 2];
    double vec[NMAX][3];
    double xs[NMAX];
    double2 vw[NMAX];
    double2 red[2 * NMAX / 32];
    double2 bc;
    int z[NMAX];
    int t[NMAX];
    int idx[NMAX];  // atom index of each member (canonical order = ascending index)
    double L[3];
    const FrameGeom *gp;  // geometry of the target's frame (general cells: minimum image through the cell matrix)
    int n;
    int pad;
};

template <int NMAX>
__global__ void __launch_bounds__(2 * NMAX) k_tridiag_blk(const __grid_constant__ DescParams P, const int *__restrict__ list,
                                                          int nlist, int slot0, int nslots, double *__restrict__ dsc,
                                                          double *__restrict__ esc, int *__restrict__ nsc) {
    constexpr int T = 2 * NMAX, NW = T / 32;
    typedef BlkSmem<NMAX> Sm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sm *S = reinterpret_cast<Sm *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x;  // one slot per block
    if (tid == 0) S->n = 0;
    for (int q = tid; q < NMAX * (NMAX + 2); q += T) (&S->A[0][0])[q] = 0.0;
    __syncthreads();
    if (warp == 0) {
        const int ord = list[slot0 + s];
        const int gt = P.targets ? P.targets[ord] : ord;
        const int f = gt / P.N, a = gt - f * P.N;
        const FrameGeom &G0 = P.geom[f];
        const double *pos = P.pos + (size_t)f * P.N * 3;
        const double pa[3] = {pos[(size_t)a * 3], pos[(size_t)a * 3 + 1], pos[(size_t)a * 3 + 2]};
        const double Lg[3] = {G0.ortho[0], G0.ortho[4], G0.ortho[8]};
        gather_scan(P, f, a, ord, [&](int j, int tj) {
            const int k = atomicAdd(&S->n, 1);
            if (k < NMAX) {
double dv[3];
                member_delta(P.periodic, G0, pos, j, pa, dv);
                S->vec[k][0] = dv[0];
                S->vec[k][1] = dv[1];
                S->vec[k][2] = dv[2];
                S->z[k] = P.Z[tj];
                S->t[k] = tj;
                S->idx[k] = j;
            }
        });
        if (lane == 0) {
            S->L[0] = Lg[0];
            S->L[1] = Lg[1];
            S->L[2] = Lg[2];
            S->gp = &G0;
        }
    }
    __syncthreads();
    int n = S->n;
    if (n > NMAX) {
        if (tid == 0) atomicExch(&P.stats->err, MDB_ERR_BOND_OVERFLOW);
        n = NMAX;
    }
    canonical_member_order(S, n, tid, [] { __syncthreads(); });
    // Coulomb matrix (datasetbuilder.py:341-348), each pair once; see k_tridiag_reg for the minimum image
    for (int i = tid; i < n; i += T) S->A[i][i] = P.diag[S->t[i]];
    coulomb_pairs(P, S, n, tid, T);
    __syncthreads();
    const int row = tid >> 1, hf = tid & 1;
    double *arow = &S->A[row][0];
    for (int k = 0; k < n - 1; ++k) {
        const bool act = row > k && row < n;
        const double x = act ? arow[k] : 0.0;
        if (hf == 0) S->xs[row] = x;
        __syncthreads();
        // this thread's half of (T x)_row: columns k+1+hf, k+3+hf, ...
        double acc0 = 0.0, acc1 = 0.0;
        if (act) {
            int c = k + 1 + hf;
            for (; c + 2 < n; c += 4) {
                acc0 = fma(arow[c], S->xs[c], acc0);
                acc1 = fma(arow[c + 2], S->xs[c + 2], acc1);
            }
            if (c < n) acc0 = fma(arow[c], S->xs[c], acc0);
        }
        double pt = acc0 + acc1;
        pt += __shfl_xor_sync(0xffffffffu, pt, 1);  // both threads of the row hold (T x)_row
        double sig = hf == 0 ? x * x : 0.0, tau = hf == 0 ? x * pt : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sig += __shfl_xor_sync(0xffffffffu, sig, o);
            tau += __shfl_xor_sync(0xffffffffu, tau, o);
        }
        if (lane == 0) S->red[warp] = make_double2(sig, tau);
        if (row == k + 1 && hf == 0) S->bc = make_double2(pt, arow[k + 1]);
        __syncthreads();
        sig = 0.0;
        tau = 0.0;
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            const double2 r = S->red[q];
            sig += r.x;
            tau += r.y;
        }
        const double2 bc = S->bc;
        const double x1 = S->xs[k + 1];
        const bool doit = sig > 0.0;
        const double rs = doit ? sig * fast_rsqrt(sig) : 0.0;
        const double alpha = x1 >= 0.0 ? -rs : rs;
        const double H = sig - alpha * x1;
        const double iH = doit ? fast_rcp(H) : 0.0;
        const double Kc = (tau - 2.0 * alpha * bc.x + alpha * alpha * bc.y) * (0.5 * iH * iH);
        const double v = (row == k + 1) ? x - alpha : x;
        const double pg = act ? (pt - alpha * arow[k + 1]) * iH : 0.0;
        const double w = act ? pg - Kc * v : 0.0;
        if (tid == 0) {
            dsc[(size_t)k * nslots + s] = S->A[k][k];
            esc[(size_t)k * nslots + s] = alpha;
        }
        if (hf == 0) S->vw[row] = make_double2(v, w);
        __syncthreads();
        if (act && doit) {
            for (int c = k + 1 + hf; c < n; c += 2) {
                const double2 q = S->vw[c];
                arow[c] = fma(-q.x, w, fma(-q.y, v, arow[c]));
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (n > 0) dsc[(size_t)(n - 1) * nslots + s] = S->A[n - 1][n - 1];
        nsc[s] = n;
    }
}

// ---------------------------------------------------------------------------------------------
// K6a'', clusters of 161..512 atoms (cutoffs of 7 A and more in condensed phases): one block of 512 threads
// per matrix, the matrix in a global-memory scratch row-major with stride 512 (2 MB, L2-resident while a few
// hundred blocks are in flight), member list and the two vectors of a step in shared memory.  Textbook
// Householder step (EISPACK tred1): u = x - alpha e1, H = u.u / 2, p = T u / H, K = u.p / (2 H), q = p - K u,
// T -= u q^T + q u^T, with IEEE sqrt and divide -- this class is a fallback for completeness, not a hot path.
// One warp per row of T, the lanes along the row (coalesced).  Output as k_tridiag_blk.
// ---------------------------------------------------------------------------------------------
#define GLB_NMAX 512
#define GLB_THREADS 512
struct GlbMatrix {
    double *p;
    __device__ __forceinline__ double *operator[](int a) const { return p + (size_t)a * GLB_NMAX; }
};
struct GlbSmem {
    GlbMatrix A;
    double vec[GLB_NMAX][3];
    double u[GLB_NMAX];
    double q[GLB_NMAX];
    double red[GLB_THREADS / 32];
    double bc[2];
    int z[GLB_NMAX];
    int t[GLB_NMAX];
    int idx[GLB_NMAX];
    double L[3];
    const FrameGeom *gp;
    int n;
};

__device__ __forceinline__ double glb_block_sum(GlbSmem *S, double v, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // red[] of the previous reduction has been read by everyone
    if ((tid & 31) == 0) S->red[tid >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < GLB_THREADS / 32; ++w) r += S->red[w];
    return r;
}

__global__ void __launch_bounds__(GLB_THREADS) k_tridiag_glb(const __grid_constant__ DescParams P, const int *__restrict__ list,
                                                             int slot0, int sub0, int nslots, double *__restrict__ Ascr,
                                                             double *__restrict__ dsc, double *__restrict__ esc,
                                                             int *__restrict__ nsc) {
    __shared__ GlbSmem Sm;
    GlbSmem *S = &Sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = sub0 + blockIdx.x;  // slot of this block inside the [k][slot] scratch
    if (tid == 0) {
        S->n = 0;
        S->A.p = Ascr + (size_t)blockIdx.x * GLB_NMAX * GLB_NMAX;
    }
    __syncthreads();
    if (warp == 0) {
        const int ord = list[slot0 + s];
        const int gt = P.targets ? P.targets[ord] : ord;
        const int f = gt / P.N, a = gt - f * P.N;
        const FrameGeom &G0 = P.geom[f];
        const double *pos = P.pos + (size_t)f * P.N * 3;
        const double pa[3] = {pos[(size_t)a * 3], pos[(size_t)a * 3 + 1], pos[(size_t)a * 3 + 2]};
        gather_scan(P, f, a, ord, [&](int j, int tj) {
            const int k = atomicAdd(&S->n, 1);
            if (k < GLB_NMAX) {
                double dv[3];
                member_delta(P.periodic, G0, pos, j, pa, dv);
                S->vec[k][0] = dv[0];
                S->vec[k][1] = dv[1];
                S->vec[k][2] = dv[2];
                S->z[k] = P.Z[tj];
                S->t[k] = tj;
                S->idx[k] = j;
            }
        });
        if (lane == 0) {
            S->L[0] = G0.ortho[0];
            S->L[1] = G0.ortho[4];
            S->L[2] = G0.ortho[8];
            S->gp = &G0;
        }
    }
    __syncthreads();
    int n = S->n;
    if (n > GLB_NMAX) {
        if (tid == 0) atomicExch(&P.stats->err, MDB_ERR_BOND_OVERFLOW);
        n = GLB_NMAX;
    }
    canonical_member_order(S, n, tid, [] { __syncthreads(); });
    for (int i = tid; i < n; i += GLB_THREADS) S->A[i][i] = P.diag[S->t[i]];
    coulomb_pairs(P, S, n, tid, GLB_THREADS);
    __syncthreads();
    const GlbMatrix A = S->A;
    for (int k = 0; k + 1 < n; ++k) {
        const int m = n - k - 1;  // order of the trailing block T = A[k+1.., k+1..]; x = A[k][k+1..]
        const double x = tid < m ? A[k][k + 1 + tid] : 0.0;
        const double sig = glb_block_sum(S, x * x, tid);
        const double x0 = A[k][k + 1];
        const double alpha = sig > 0.0 ? (x0 >= 0.0 ? -sqrt(sig) : sqrt(sig)) : 0.0;
        if (tid == 0) {
            dsc[(size_t)k * nslots + s] = A[k][k];
            esc[(size_t)k * nslots + s] = alpha;
        }
        if (!(sig > 0.0)) continue;  // column already reduced (uniform over the block)
        const double H = sig - alpha * x0;
        const double uval = tid == 0 ? x - alpha : x;
        if (tid < m) S->u[tid] = uval;
        __syncthreads();
        // p = T u / H
        for (int i = warp; i < m; i += GLB_THREADS / 32) {
            const double *row = A[k + 1 + i] + k + 1;
            double acc = 0.0;
            for (int j = lane; j < m; j += 32) acc = fma(row[j], S->u[j], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) S->q[i] = acc / H;
        }
        __syncthreads();
        const double pv = tid < m ? S->q[tid] : 0.0;
        const double K = glb_block_sum(S, uval * pv, tid) / (2.0 * H);
        if (tid < m) S->q[tid] = pv - K * uval;
        __syncthreads();
        for (int i = warp; i < m; i += GLB_THREADS / 32) {
            double *row = A[k + 1 + i] + k + 1;
            const double ui = S->u[i], qi = S->q[i];
            for (int j = lane; j < m; j += 32) row[j] = fma(-ui, S->q[j], fma(-qi, S->u[j], row[j]));
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (n > 0) dsc[(size_t)(n - 1) * nslots + s] = A[n - 1][n - 1];
        nsc[s] = n;
    }
}

// ---------------------------------------------------------------------------------------------
// K6a', clusters of up to 32 atoms: the matrix lives in registers.  LPM lanes own one matrix, lane g
// holds column g (= row g) as NMAX doubles, so a warp works on 32 / LPM matrices; shared memory
// carries the member list, a staging copy of the matrix (built pair by pair, each pair once) and the
// two vectors every lane has to see.  The Householder loop is fully unrolled (every register index
// is a compile-time constant).  Step k, with x = A[k+1.., k] and T = A[k+1.., k+1..]:
//     sigma = x.x, alpha = -sign(x1) sqrt(sigma), v = x - alpha e1, H = sigma - alpha x1,
//     p = T v / H = (T x - alpha T[:, 0]) / H,
//     K = v.T v / (2 H^2) with v.T v = x.T x - 2 alpha (T x)_0 + alpha^2 T_00,
//     w = p - K v,  T -= v w^T + w v^T,  d_k = A_kk, e_k = alpha,
// so one pass over the columns (T x) and ONE fused reduction (sigma, x.T x) give everything.
// Reciprocal and reciprocal square root are MUFU seeds + Newton steps (a few ulp; the spectrum
// tolerance is 1e-5): the IEEE sequences would triple the size of the unrolled code.
// ---------------------------------------------------------------------------------------------
// the same with NS Newton steps: the MUFU seeds are good to ~2^-20, one step gives ~2^-40 (1e-12), which a
// Householder reflector tolerates (a relative error d in H perturbs the similarity transform by d ||A||; the
// spectrum tolerance is 1e-5 with a floor of 1e-8 ||M||)
#ifndef MDB_NEWTON
#define MDB_NEWTON 1
#endif
template <int NS>
__device__ __forceinline__ double fast_rcp_n(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < NS; ++it) {
        const double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
    }
    return y;
}
template <int NS>
__device__ __forceinline__ double fast_rsqrt_n(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
#pragma unroll
    for (int it = 0; it < NS; ++it) {
        const double e = fma(-h * y, y, 0.5);
        y = fma(y, e, y);
    }
    return y;
}
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // the seed is good to about 2^-20; two Newton steps square that twice
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double e = fma(-h * y, y, 0.5);
        y = fma(y, e, y);
    }
    return y;
}

template <int NMAX, int LPM>
struct RegSmem {
    double A[NMAX][NMAX + 1];
    double vec[NMAX][3];
    double xs[LPM];
    double2 vw[LPM];
    int z[NMAX];
    int t[NMAX];
    int idx[NMAX];  // atom index of each member (canonical order = ascending index)
    double L[3];
    const FrameGeom *gp;  // geometry of the target's frame (general cells: minimum image through the cell matrix)
    int n;
    int pad;
};

template <int NMAX, int LPM, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (NMAX > 16 ? 2 : 3)) k_tridiag_reg(const __grid_constant__ DescParams P, const int *__restrict__ list,
                                                            int nlist, int slot0, int nslots, double *__restrict__ dsc,
                                                            double *__restrict__ esc, int *__restrict__ nsc) {
    static_assert(NMAX <= LPM, "one column per lane");
    constexpr int G = 32 / LPM;
    typedef RegSmem<NMAX, LPM> Sm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sm *Sw = reinterpret_cast<Sm *>(smem_raw) + (threadIdx.x >> 5) * G;
    const int lane = threadIdx.x & 31;
    const int grp = lane / LPM, gl = lane % LPM;
    const int sw0 = (blockIdx.x * WARPS + (threadIdx.x >> 5)) * G;  // first slot of this warp
    for (int g = 0; g < G; ++g) {
        const int sg = sw0 + g;
        Sm *Sg = Sw + g;
        if (lane == 0) Sg->n = 0;
        if (sg >= nslots) continue;  // warp-uniform
        const int ord = list[slot0 + sg];
        const int gt = P.targets ? P.targets[ord] : ord;
        const int f = gt / P.N, a = gt - f * P.N;
        const FrameGeom &G0 = P.geom[f];
        const double *pos = P.pos + (size_t)f * P.N * 3;
        __syncwarp();
        const double pa[3] = {pos[(size_t)a * 3], pos[(size_t)a * 3 + 1], pos[(size_t)a * 3 + 2]};
        const double Lg[3] = {G0.ortho[0], G0.ortho[4], G0.ortho[8]};
        gather_scan(P, f, a, ord, [&](int j, int tj) {
            const int k = atomicAdd(&Sg->n, 1);
            if (k < NMAX) {
double dv[3];
                member_delta(P.periodic, G0, pos, j, pa, dv);
                Sg->vec[k][0] = dv[0];
                Sg->vec[k][1] = dv[1];
                Sg->vec[k][2] = dv[2];
                Sg->z[k] = P.Z[tj];
                Sg->t[k] = tj;
                Sg->idx[k] = j;
            }
        });
        if (lane == 0) {
            Sg->L[0] = Lg[0];
            Sg->L[1] = Lg[1];
            Sg->L[2] = Lg[2];
            Sg->gp = &G0;
        }
    }
    __syncwarp();
    Sm *S = Sw + grp;
    const int s = sw0 + grp;
    const bool valid = s < nslots;
    int n = S->n;
    if (n > NMAX) {
        if (gl == 0) atomicExch(&P.stats->err, MDB_ERR_BOND_OVERFLOW);
        n = NMAX;
    }
    canonical_member_order(S, n, gl, [] { __syncwarp(); });
    // Coulomb matrix (datasetbuilder.py:341-348) into the staging copy: the n(n-1)/2 pairs are dealt
    // evenly to the lanes of the group, pair p -> (i, j), j < i.  Both vectors lie within +-L/2 of the
    // target, so one conditional shift per axis gives the minimum image.  Rows / columns >= n are zero.
    for (int q = gl; q < NMAX * NMAX; q += LPM) S->A[q / NMAX][q % NMAX] = 0.0;
    __syncwarp();
    for (int i = gl; i < n; i += LPM) S->A[i][i] = P.diag[S->t[i]];
    coulomb_pairs(P, S, n, gl, LPM);
    __syncwarp();
    double a[NMAX];
    {
        const int col = gl < NMAX ? gl : 0;
#pragma unroll
        for (int i = 0; i < NMAX; ++i) a[i] = gl < NMAX ? S->A[i][col] : 0.0;
    }
    double *dptr = dsc + s + (size_t)gl * nslots;
    double *eptr = esc + s + (size_t)gl * nslots;
    const bool hi = (lane & (LPM >> 1)) != 0;  // upper half of the lane group
#pragma unroll
    for (int k = 0; k < NMAX - 1; ++k) {
        // keeps the warps of the block within a few steps of each other: the unrolled code is larger
        // than the instruction cache, and warps that run together share its lines
        if ((k & 3) == 0) __syncthreads();
        const double x = (gl > k && gl < n) ? a[k] : 0.0;
        S->xs[gl] = x;
        __syncwarp();
        // (T x)_gl over rows k+1 .. NMAX-1; independent chains hide the DFMA latency
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
        for (int i = k + 1; i < NMAX; ++i) {
            const double xi = S->xs[i];
            if (((i - k - 1) & 3) == 0) acc0 = fma(a[i], xi, acc0);
            if (((i - k - 1) & 3) == 1) acc1 = fma(a[i], xi, acc1);
            if (((i - k - 1) & 3) == 2) acc2 = fma(a[i], xi, acc2);
            if (((i - k - 1) & 3) == 3) acc3 = fma(a[i], xi, acc3);
        }
        const double pt = (acc0 + acc1) + (acc2 + acc3);
        // fused reduction of (sigma, tau): the halves of the group first swap one value each, then a
        // single butterfly runs on the kept value, and a final swap gives every lane both sums
        const double sg0 = x * x, ta0 = x * pt;
        double keep = hi ? ta0 : sg0;
        keep += __shfl_xor_sync(0xffffffffu, hi ? sg0 : ta0, LPM >> 1);
#pragma unroll
        for (int o = LPM >> 2; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
        const double other = __shfl_xor_sync(0xffffffffu, keep, LPM >> 1);
        const double sig = hi ? other : keep, tau = hi ? keep : other;
        const double x1 = S->xs[k + 1];
        const double pt1 = __shfl_sync(0xffffffffu, pt, k + 1, LPM);
        const double t00 = __shfl_sync(0xffffffffu, a[k + 1], k + 1, LPM);
        const bool doit = sig > 0.0;
        const double rs = doit ? sig * fast_rsqrt(sig) : 0.0;  // sqrt(sigma)
        const double alpha = x1 >= 0.0 ? -rs : rs;
        const double H = sig - alpha * x1;
        const double iH = doit ? fast_rcp(H) : 0.0;
        const double K = (tau - 2.0 * alpha * pt1 + alpha * alpha * t00) * (0.5 * iH * iH);
        const double v = (gl == k + 1) ? x - alpha : x;
        const double pg = (pt - alpha * a[k + 1]) * iH;
        const double w = (gl > k) ? pg - K * v : 0.0;
        if (gl == k && valid && k < n) {
            *dptr = a[k];
            if (k + 1 < n) *eptr = alpha;
        }
        S->vw[gl] = make_double2(v, w);
        __syncwarp();
#pragma unroll
        for (int i = k + 1; i < NMAX; ++i) {
            const double2 q = S->vw[i];
            a[i] = fma(-q.x, w, fma(-q.y, v, a[i]));
        }
    }
    if (valid) {
        if (gl == NMAX - 1 && n == NMAX) *dptr = a[NMAX - 1];
        if (gl == 0) nsc[s] = n;
    }
}

// ---------------------------------------------------------------------------------------------
// K6a-bord, clusters of 33..40 atoms (round 2): the first n - 32 Householder steps run on the staging copy in
// shared memory (one warp, two rows per lane), which leaves a trailing matrix of at most 32 columns -- and that
// one goes through the one-column register loop of k_tridiag_reg<32> unchanged.  The two-columns-per-lane kernel
// needed 32 ns per 33..36-atom matrix (a second register slot of which 4 lanes in 32 are used) against 13 ns
// for a 32-atom matrix; a handful of shared-memory steps costs far less than doubling the register work.
// ---------------------------------------------------------------------------------------------
template <int NS>
struct BordSmem {
    double A[NS][NS + 1];
    double vec[NS][3];
    double xs[NS < 32 ? 32 : NS];
    double2 vw[NS < 32 ? 32 : NS];
    int z[NS];
    int t[NS];
    int idx[NS];
    double L[3];
    const FrameGeom *gp;
    int n;
    int pad;
};

template <int NS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2) k_tridiag_bord(const __grid_constant__ DescParams P, const int *__restrict__ list,
                                                                int nlist, int slot0, int nslots, double *__restrict__ dsc,
                                                                double *__restrict__ esc, int *__restrict__ nsc) {
    static_assert(NS > 32 && NS <= 64, "two staged rows per lane");
    constexpr int NMAX = 32;
    typedef BordSmem<NS> Sm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sm *S = reinterpret_cast<Sm *>(smem_raw) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, gl = lane;
    const int s = blockIdx.x * WARPS + (threadIdx.x >> 5);
    const bool valid = s < nslots;
    if (lane == 0) S->n = 0;
    if (valid) {
        const int ord = list[slot0 + s];
        const int gt = P.targets ? P.targets[ord] : ord;
        const int f = gt / P.N, at = gt - f * P.N;
        const FrameGeom &G0 = P.geom[f];
        const double *pos = P.pos + (size_t)f * P.N * 3;
        __syncwarp();
        const double pa[3] = {pos[(size_t)at * 3], pos[(size_t)at * 3 + 1], pos[(size_t)at * 3 + 2]};
        gather_scan(P, f, at, ord, [&](int j, int tj) {
            const int k = atomicAdd(&S->n, 1);
            if (k < NS) {
                double dv[3];
                member_delta(P.periodic, G0, pos, j, pa, dv);
                S->vec[k][0] = dv[0];
                S->vec[k][1] = dv[1];
                S->vec[k][2] = dv[2];
                S->z[k] = P.Z[tj];
                S->t[k] = tj;
                S->idx[k] = j;
            }
        });
        if (lane == 0) {
            S->L[0] = G0.ortho[0];
            S->L[1] = G0.ortho[4];
            S->L[2] = G0.ortho[8];
            S->gp = &G0;
        }
    }
    __syncwarp();
    int n = S->n;
    if (n > NS) {
        if (gl == 0) atomicExch(&P.stats->err, MDB_ERR_BOND_OVERFLOW);
        n = NS;
    }
    {   // canonical member order (ascending atom index), two entries per lane
        double v0[2], v1[2], v2[2];
        int zz[2], tt[2], rank[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int e = c * 32 + gl;
            rank[c] = -1;
            if (e < n) {
                v0[c] = S->vec[e][0];
                v1[c] = S->vec[e][1];
                v2[c] = S->vec[e][2];
                zz[c] = S->z[e];
                tt[c] = S->t[e];
                const int ii = S->idx[e];
                int r = 0;
                for (int q = 0; q < n; ++q) r += S->idx[q] < ii;
                rank[c] = r;
            }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (rank[c] >= 0) {
                S->vec[rank[c]][0] = v0[c];
                S->vec[rank[c]][1] = v1[c];
                S->vec[rank[c]][2] = v2[c];
                S->z[rank[c]] = zz[c];
                S->t[rank[c]] = tt[c];
            }
        __syncwarp();
    }
    for (int q = gl; q < NS * NS; q += 32) S->A[q / NS][q % NS] = 0.0;
    __syncwarp();
    for (int i = gl; i < n; i += 32) S->A[i][i] = P.diag[S->t[i]];
    coulomb_pairs(P, S, n, gl, 32);
    __syncwarp();
    // ---- border: steps 0 .. nb-1 on the staged matrix, rows gl and gl + 32 of the trailing block per lane
    const int nb = n > NMAX ? n - NMAX : 0;
    for (int k = 0; k < nb; ++k) {
        const int r0 = k + 1 + gl, r1 = r0 + 32;
        const double x0 = r0 < n ? S->A[r0][k] : 0.0;
        const double x1r = r1 < n ? S->A[r1][k] : 0.0;
        if (r0 < NS) S->xs[r0] = x0;
        if (r1 < NS) S->xs[r1] = x1r;
        __syncwarp();
        double p0 = 0.0, p1 = 0.0;
        for (int j = k + 1; j < n; ++j) {
            const double xj = S->xs[j];
            if (r0 < n) p0 = fma(S->A[r0][j], xj, p0);
            if (r1 < n) p1 = fma(S->A[r1][j], xj, p1);
        }
        double sig = x0 * x0 + x1r * x1r, tau = x0 * p0 + x1r * p1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sig += __shfl_xor_sync(0xffffffffu, sig, o);
            tau += __shfl_xor_sync(0xffffffffu, tau, o);
        }
        const double xk1 = S->xs[k + 1];
        const double pt1 = __shfl_sync(0xffffffffu, p0, 0);  // row k + 1 is lane 0's first row
        const double t00 = S->A[k + 1][k + 1];
        const bool doit = sig > 0.0;
        const double rs = doit ? sig * fast_rsqrt(sig) : 0.0;
        const double alpha = xk1 >= 0.0 ? -rs : rs;
        const double H = sig - alpha * xk1;
        const double iH = doit ? fast_rcp(H) : 0.0;
        const double Kc = (tau - 2.0 * alpha * pt1 + alpha * alpha * t00) * (0.5 * iH * iH);
        const double va = (gl == 0) ? x0 - alpha : x0;  // r0 == k + 1 for lane 0
        const double vb = x1r;
        const double wa = r0 < n ? (p0 - alpha * S->A[r0][k + 1]) * iH - Kc * va : 0.0;
        const double wb = r1 < n ? (p1 - alpha * S->A[r1][k + 1]) * iH - Kc * vb : 0.0;
        if (gl == 0 && valid) {
            dsc[s + (size_t)k * nslots] = S->A[k][k];
            esc[s + (size_t)k * nslots] = alpha;
        }
        __syncwarp();
        if (r0 < NS) S->vw[r0] = make_double2(r0 < n ? va : 0.0, wa);
        if (r1 < NS) S->vw[r1] = make_double2(r1 < n ? vb : 0.0, wb);
        __syncwarp();
        if (doit) {
            for (int j = k + 1; j < n; ++j) {
                const double2 q = S->vw[j];
                if (r0 < n) S->A[r0][j] = fma(-q.x, wa, fma(-q.y, va, S->A[r0][j]));
                if (r1 < n) S->A[r1][j] = fma(-q.x, wb, fma(-q.y, vb, S->A[r1][j]));
            }
        }
        __syncwarp();
    }
    // ---- trailing matrix (at most 32 columns) into registers: lane g holds column nb + g
    const int nt2 = n - nb;
    double a[NMAX];
    {
#pragma unroll
        for (int i = 0; i < NMAX; ++i) a[i] = (nb + i < n && nb + gl < n) ? S->A[nb + i][nb + gl] : 0.0;
    }
    double *dptr = dsc + s + (size_t)(nb + gl) * nslots;
    double *eptr = esc + s + (size_t)(nb + gl) * nslots;
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < NMAX - 1; ++k) {
        if ((k & 3) == 0) __syncthreads();
        const double x = (gl > k && gl < nt2) ? a[k] : 0.0;
        S->xs[gl] = x;
        __syncwarp();
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
        for (int i = k + 1; i < NMAX; ++i) {
            const double xi = S->xs[i];
            if (((i - k - 1) & 3) == 0) acc0 = fma(a[i], xi, acc0);
            if (((i - k - 1) & 3) == 1) acc1 = fma(a[i], xi, acc1);
            if (((i - k - 1) & 3) == 2) acc2 = fma(a[i], xi, acc2);
            if (((i - k - 1) & 3) == 3) acc3 = fma(a[i], xi, acc3);
        }
        const double pt = (acc0 + acc1) + (acc2 + acc3);
        const double sg0 = x * x, ta0 = x * pt;
        double keep = hi ? ta0 : sg0;
        keep += __shfl_xor_sync(0xffffffffu, hi ? sg0 : ta0, 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
        const double other = __shfl_xor_sync(0xffffffffu, keep, 16);
        const double sig = hi ? other : keep, tau = hi ? keep : other;
        const double x1 = S->xs[k + 1];
        const double pt1 = __shfl_sync(0xffffffffu, pt, k + 1);
        const double t00 = __shfl_sync(0xffffffffu, a[k + 1], k + 1);
        const bool doit = sig > 0.0;
        const double rs = doit ? sig * fast_rsqrt(sig) : 0.0;
        const double alpha = x1 >= 0.0 ? -rs : rs;
        const double H = sig - alpha * x1;
        const double iH = doit ? fast_rcp(H) : 0.0;
        const double K = (tau - 2.0 * alpha * pt1 + alpha * alpha * t00) * (0.5 * iH * iH);
        const double v = (gl == k + 1) ? x - alpha : x;
        const double pg = (pt - alpha * a[k + 1]) * iH;
        const double w = (gl > k) ? pg - K * v : 0.0;
        if (gl == k && valid && k < nt2) {
            *dptr = a[k];
            if (k + 1 < nt2) *eptr = alpha;
        }
        S->vw[gl] = make_double2(v, w);
        __syncwarp();
#pragma unroll
        for (int i = k + 1; i < NMAX; ++i) {
            const double2 q = S->vw[i];
            a[i] = fma(-q.x, w, fma(-q.y, v, a[i]));
        }
    }
    if (valid) {
        if (gl == NMAX - 1 && nt2 == NMAX) *dptr = a[NMAX - 1];
        if (gl == 0) nsc[s] = n;
    }
}

// ---------------------------------------------------------------------------------------------
// K6a-mc, clusters of up to 48 atoms (round 2): the register-resident Householder scheme with SEVERAL
// columns per lane.  A group of LPM lanes owns one matrix, lane g holds columns g, g + LPM, ... (CPL
// "slots" of NMAX doubles each), so a warp works on 32 / LPM matrices and
//   * every shared-memory broadcast (x_i in the T x pass, (v_i, w_i) in the rank-2 update) feeds CPL
//     fused multiply-adds instead of one,
//   * the per-step scalar work (one butterfly, rsqrt / rcp, the (v, w) publication) is paid once per
//     32 / LPM matrices instead of once per matrix,
//   * a slot whose columns are all finished (slot c is done once k >= (c + 1) LPM - 1) drops out of the
//     unrolled code at compile time.
// ncu of the one-column kernel had the FP64 pipe 32 % active with ~8,900 warp instructions per 32-atom
// matrix, half of them per-step overhead.  The 33..48 classes use LPM = 32, CPL = 2 on ONE warp (the
// warp-pair kernel below needed three named barriers per step and ran 3.3 x slower per matrix than the
// 32-atom class for 1.7 x the arithmetic).
// ---------------------------------------------------------------------------------------------
template <int NMAX, int LPM, int CPL>
struct McSmem {
    double A[NMAX][NMAX + 1];  // staging copy of the Coulomb matrix (each pair evaluated once)
    double vec[NMAX][3];
    double xs[LPM * CPL];
    double2 vw[LPM * CPL];
    int z[NMAX];
    int t[NMAX];
    int idx[NMAX];
    double L[3];
    const FrameGeom *gp;  // geometry of the target's frame (general cells: minimum image through the cell matrix)
    int n;
    int pad;
};

template <int NMAX, int LPM, int CPL>
struct McCtx {
    McSmem<NMAX, LPM, CPL> *S;
    double *dsc, *esc;  // scratch of this matrix: element k at [k * nslots]
    size_t nslots;
    int gl, n;
    bool valid, hi;
};

template <int K, int NMAX, int LPM, int CPL>
__device__ __forceinline__ void mc_step(double (&a)[CPL][NMAX], const McCtx<NMAX, LPM, CPL> &cx) {
    constexpr int C0 = (K + 1) / LPM;  // slots below C0 hold finished columns only
    McSmem<NMAX, LPM, CPL> *S = cx.S;
    const int gl = cx.gl, n = cx.n;
    // the unrolled code is far larger than the instruction caches (ncu: `no_instruction` was the top stall
    // with free-running warps): the warps of the block stay within two steps of each other and share the lines
#ifndef MDB_MC_SYNC
#define MDB_MC_SYNC 7
#endif
    if ((K & MDB_MC_SYNC) == 0) __syncthreads();
    double x[CPL];
#pragma unroll
    for (int c = C0; c < CPL; ++c) {
        const int col = c * LPM + gl;
        x[c] = (col > K && col < n) ? a[c][K] : 0.0;
        S->xs[col] = x[c];
    }
    __syncwarp();
    double acc0[CPL], acc1[CPL];
#pragma unroll
    for (int c = C0; c < CPL; ++c) acc0[c] = acc1[c] = 0.0;
#pragma unroll
    for (int i = K + 1; i < NMAX; ++i) {
        const double xi = S->xs[i];
#pragma unroll
        for (int c = C0; c < CPL; ++c) {
            if (((i - K - 1) & 1) == 0) acc0[c] = fma(a[c][i], xi, acc0[c]);
            else acc1[c] = fma(a[c][i], xi, acc1[c]);
        }
    }
    double pt[CPL];
    double sg0 = 0.0, ta0 = 0.0;
#pragma unroll
    for (int c = C0; c < CPL; ++c) {
        pt[c] = acc0[c] + acc1[c];
        sg0 = fma(x[c], x[c], sg0);
        ta0 = fma(x[c], pt[c], ta0);
    }
    // fused reduction of (sigma, tau) over the LPM lanes of the group (see k_tridiag_reg)
    const bool hi = cx.hi;
    double keep = hi ? ta0 : sg0;
    keep += __shfl_xor_sync(0xffffffffu, hi ? sg0 : ta0, LPM >> 1);
#pragma unroll
    for (int o = LPM >> 2; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    const double other = __shfl_xor_sync(0xffffffffu, keep, LPM >> 1);
    const double sig = hi ? other : keep, tau = hi ? keep : other;
    constexpr int C1 = (K + 1) / LPM, L1 = (K + 1) % LPM;  // owner of column K + 1
    const double x1 = S->xs[K + 1];
    const double pt1 = __shfl_sync(0xffffffffu, pt[C1], L1, LPM);
    const double t00 = __shfl_sync(0xffffffffu, a[C1][K + 1], L1, LPM);
    const bool doit = sig > 0.0;
    const double rs = doit ? sig * fast_rsqrt_n<MDB_NEWTON>(sig) : 0.0;
    const double alpha = x1 >= 0.0 ? -rs : rs;
    const double H = sig - alpha * x1;
    const double iH = doit ? fast_rcp_n<MDB_NEWTON>(H) : 0.0;
    const double Kc = (tau - 2.0 * alpha * pt1 + alpha * alpha * t00) * (0.5 * iH * iH);
    constexpr int CK = K / LPM, LK = K % LPM;  // owner of column K
    if (gl == LK && cx.valid && K < n) {
        cx.dsc[(size_t)K * cx.nslots] = a[CK][K];
        if (K + 1 < n) cx.esc[(size_t)K * cx.nslots] = alpha;
    }
    double v[CPL], w[CPL];
#pragma unroll
    for (int c = C0; c < CPL; ++c) {
        const int col = c * LPM + gl;
        v[c] = (col == K + 1) ? x[c] - alpha : x[c];
        const double pg = (pt[c] - alpha * a[c][K + 1]) * iH;
        w[c] = (col > K) ? pg - Kc * v[c] : 0.0;
        S->vw[col] = make_double2(v[c], w[c]);
    }
    __syncwarp();
#pragma unroll
    for (int i = K + 1; i < NMAX; ++i) {
        const double2 q = S->vw[i];
#pragma unroll
        for (int c = C0; c < CPL; ++c) a[c][i] = fma(-q.x, w[c], fma(-q.y, v[c], a[c][i]));
    }
}

template <int K, int NMAX, int LPM, int CPL>
struct McSteps {
    static __device__ __forceinline__ void run(double (&a)[CPL][NMAX], const McCtx<NMAX, LPM, CPL> &cx) {
        mc_step<K, NMAX, LPM, CPL>(a, cx);
        McSteps<K + 1, NMAX, LPM, CPL>::run(a, cx);
    }
};
template <int NMAX, int LPM, int CPL>
struct McSteps<NMAX - 1, NMAX, LPM, CPL> {
    static __device__ __forceinline__ void run(double (&)[CPL][NMAX], const McCtx<NMAX, LPM, CPL> &) {}
};

template <int NMAX, int LPM, int CPL, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_tridiag_mc(const __grid_constant__ DescParams P, const int *__restrict__ list,
                                                                 int nlist, int slot0, int nslots, double *__restrict__ dsc,
                                                                 double *__restrict__ esc, int *__restrict__ nsc) {
    static_assert(NMAX <= LPM * CPL, "every column needs a lane slot");
    static_assert(LPM >= 4 && LPM <= 32 && (LPM & (LPM - 1)) == 0, "lane groups are powers of two");
    constexpr int G = 32 / LPM;
    typedef McSmem<NMAX, LPM, CPL> Sm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sm *Sw = reinterpret_cast<Sm *>(smem_raw) + (threadIdx.x >> 5) * G;
    const int lane = threadIdx.x & 31;
    const int grp = lane / LPM, gl = lane % LPM;
    const int sw0 = (blockIdx.x * WARPS + (threadIdx.x >> 5)) * G;  // first slot of this warp
    for (int g = 0; g < G; ++g) {
        const int sg = sw0 + g;
        Sm *Sg = Sw + g;
        if (lane == 0) Sg->n = 0;
        if (sg >= nslots) continue;  // warp-uniform
        const int ord = list[slot0 + sg];
        const int gt = P.targets ? P.targets[ord] : ord;
        const int f = gt / P.N, at = gt - f * P.N;
        const FrameGeom &G0 = P.geom[f];
        const double *pos = P.pos + (size_t)f * P.N * 3;
        __syncwarp();
        const double pa[3] = {pos[(size_t)at * 3], pos[(size_t)at * 3 + 1], pos[(size_t)at * 3 + 2]};
        const double Lg[3] = {G0.ortho[0], G0.ortho[4], G0.ortho[8]};
        gather_scan(P, f, at, ord, [&](int j, int tj) {
            const int k = atomicAdd(&Sg->n, 1);
            if (k < NMAX) {
double dv[3];
                member_delta(P.periodic, G0, pos, j, pa, dv);
                Sg->vec[k][0] = dv[0];
                Sg->vec[k][1] = dv[1];
                Sg->vec[k][2] = dv[2];
                Sg->z[k] = P.Z[tj];
                Sg->t[k] = tj;
                Sg->idx[k] = j;
            }
        });
        if (lane == 0) {
            Sg->L[0] = Lg[0];
            Sg->L[1] = Lg[1];
            Sg->L[2] = Lg[2];
            Sg->gp = &G0;
        }
    }
    __syncwarp();
    Sm *S = Sw + grp;
    const int s = sw0 + grp;
    const bool valid = s < nslots;
    int n = S->n;
    if (n > NMAX) {
        if (gl == 0) atomicExch(&P.stats->err, MDB_ERR_BOND_OVERFLOW);
        n = NMAX;
    }
    // canonical member order (ascending atom index, the order of np.where at datasetbuilder.py:315):
    // every lane moves the entries of its own columns to their ranks
    {
        double v0[CPL], v1[CPL], v2[CPL];
        int zz[CPL], tt[CPL], rank[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int e = c * LPM + gl;
            rank[c] = -1;
            if (e < n) {
                v0[c] = S->vec[e][0];
                v1[c] = S->vec[e][1];
                v2[c] = S->vec[e][2];
                zz[c] = S->z[e];
                tt[c] = S->t[e];
                const int ii = S->idx[e];
                int r = 0;
                for (int q = 0; q < n; ++q) r += S->idx[q] < ii;
                rank[c] = r;
            }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CPL; ++c)
            if (rank[c] >= 0) {
                S->vec[rank[c]][0] = v0[c];
                S->vec[rank[c]][1] = v1[c];
                S->vec[rank[c]][2] = v2[c];
                S->z[rank[c]] = zz[c];
                S->t[rank[c]] = tt[c];
            }
        __syncwarp();
    }
    // Coulomb matrix (datasetbuilder.py:341-348) into the staging copy: the n(n-1)/2 pairs are dealt evenly to
    // the lanes of the group, pair p -> (i, j), j < i.  Both vectors lie within +-L/2 of the target, so one
    // conditional shift per axis gives the minimum image.  Rows / columns >= n are zero.
    for (int q = gl; q < NMAX * NMAX; q += LPM) S->A[q / NMAX][q % NMAX] = 0.0;
    __syncwarp();
    for (int i = gl; i < n; i += LPM) S->A[i][i] = P.diag[S->t[i]];
    coulomb_pairs(P, S, n, gl, LPM);
    __syncwarp();
    double a[CPL][NMAX];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        const int col = c * LPM + gl;
        const int cc = col < NMAX ? col : 0;
#pragma unroll
        for (int i = 0; i < NMAX; ++i) a[c][i] = col < NMAX ? S->A[i][cc] : 0.0;
    }
    McCtx<NMAX, LPM, CPL> cx;
    cx.S = S;
    cx.dsc = dsc + s;
    cx.esc = esc + s;
    cx.nslots = (size_t)nslots;
    cx.gl = gl;
    cx.n = n;
    cx.valid = valid;
    cx.hi = (lane & (LPM >> 1)) != 0;
    McSteps<0, NMAX, LPM, CPL>::run(a, cx);
    if (valid) {
        constexpr int CL = (NMAX - 1) / LPM, LL = (NMAX - 1) % LPM;
        if (gl == LL && n == NMAX) dsc[s + (size_t)(NMAX - 1) * nslots] = a[CL][NMAX - 1];
        if (gl == 0) nsc[s] = n;
    }
}

// ---------------------------------------------------------------------------------------------
// K6a'', clusters of 33..64 atoms: the same register-resident scheme on a pair of warps (lane g of
// the pair, 0 <= g < 64, holds column g).  The pair meets at a named barrier three times per step:
// after publishing x, after publishing the per-warp partial sums, and after publishing (v, w).
// ---------------------------------------------------------------------------------------------
template <int NMAX>
struct Reg2Smem {
    double A[NMAX][NMAX + 1];
    double vec[NMAX][3];
    double xs[64];
    double2 vw[64];
    double2 red[2];
    double2 bc;
    int z[NMAX];
    int t[NMAX];
    int idx[NMAX];  // atom index of each member (canonical order = ascending index)
    double L[3];
    const FrameGeom *gp;  // geometry of the target's frame (general cells: minimum image through the cell matrix)
    int n;
    int pad;
};

template <int NMAX>
struct Reg2Ctx {
    Reg2Smem<NMAX> *S;
    double *dptr, *eptr;
    int gl, lane, wp, n, barid;
    bool valid;
};

// one Householder step with a compile-time column index (template recursion instead of `#pragma
// unroll`: the unroller gives up on 63 iterations of this size and would leave a[] in local memory)
template <int K, int NMAX>
__device__ __forceinline__ void reg2_step(double (&a)[NMAX], const Reg2Ctx<NMAX> &cx) {
    Reg2Smem<NMAX> *S = cx.S;
    const int gl = cx.gl, n = cx.n;
#define PAIR_SYNC() asm volatile("bar.sync %0, 64;" ::"r"(cx.barid) : "memory")
    if ((K & 3) == 0) __syncthreads();  // see k_tridiag_reg
    const double x = (gl > K && gl < n) ? a[K] : 0.0;
    S->xs[gl] = x;
    PAIR_SYNC();
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
    for (int i = K + 1; i < NMAX; ++i) {
        const double xi = S->xs[i];
        if (((i - K - 1) & 3) == 0) acc0 = fma(a[i], xi, acc0);
        if (((i - K - 1) & 3) == 1) acc1 = fma(a[i], xi, acc1);
        if (((i - K - 1) & 3) == 2) acc2 = fma(a[i], xi, acc2);
        if (((i - K - 1) & 3) == 3) acc3 = fma(a[i], xi, acc3);
    }
    const double pt = (acc0 + acc1) + (acc2 + acc3);
    const bool hi = (cx.lane & 16) != 0;
    const double sg0 = x * x, ta0 = x * pt;
    double keep = hi ? ta0 : sg0;
    keep += __shfl_xor_sync(0xffffffffu, hi ? sg0 : ta0, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    const double other = __shfl_xor_sync(0xffffffffu, keep, 16);
    if (cx.lane == 0) S->red[cx.wp] = make_double2(keep, other);  // lane 0 is in the lower half: (sigma, tau)
    if (gl == K + 1) S->bc = make_double2(pt, a[K + 1]);
    PAIR_SYNC();
    const double2 r0 = S->red[0], r1 = S->red[1], bc = S->bc;
    const double sig = r0.x + r1.x, tau = r0.y + r1.y;
    const double x1 = S->xs[K + 1];
    const double pt1 = bc.x, t00 = bc.y;
    const bool doit = sig > 0.0;
    const double rs = doit ? sig * fast_rsqrt(sig) : 0.0;
    const double alpha = x1 >= 0.0 ? -rs : rs;
    const double H = sig - alpha * x1;
    const double iH = doit ? fast_rcp(H) : 0.0;
    const double Kc = (tau - 2.0 * alpha * pt1 + alpha * alpha * t00) * (0.5 * iH * iH);
    const double v = (gl == K + 1) ? x - alpha : x;
    const double pg = (pt - alpha * a[K + 1]) * iH;
    const double w = (gl > K) ? pg - Kc * v : 0.0;
    if (gl == K && cx.valid && K < n) {
        *cx.dptr = a[K];
        if (K + 1 < n) *cx.eptr = alpha;
    }
    S->vw[gl] = make_double2(v, w);
    PAIR_SYNC();
#pragma unroll
    for (int i = K + 1; i < NMAX; ++i) {
        const double2 q = S->vw[i];
        a[i] = fma(-q.x, w, fma(-q.y, v, a[i]));
    }
#undef PAIR_SYNC
}

template <int K, int NMAX>
struct Reg2Steps {
    static __device__ __forceinline__ void run(double (&a)[NMAX], const Reg2Ctx<NMAX> &cx) {
        reg2_step<K, NMAX>(a, cx);
        Reg2Steps<K + 1, NMAX>::run(a, cx);
    }
};
template <int NMAX>
struct Reg2Steps<NMAX - 1, NMAX> {
    static __device__ __forceinline__ void run(double (&)[NMAX], const Reg2Ctx<NMAX> &) {}
};

template <int NMAX, int PAIRS>
__global__ void __launch_bounds__(PAIRS * 64, (NMAX <= 40 ? 4 : NMAX <= 48 ? 3 : 2)) k_tridiag_reg2(const __grid_constant__ DescParams P, const int *__restrict__ list,
                                                                int nlist, int slot0, int nslots, double *__restrict__ dsc,
                                                                double *__restrict__ esc, int *__restrict__ nsc) {
    static_assert(NMAX > 32 && NMAX <= 64, "one column per lane of a warp pair");
    typedef Reg2Smem<NMAX> Sm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int pair = threadIdx.x >> 6;
    Sm *S = reinterpret_cast<Sm *>(smem_raw) + pair;
    const int lane = threadIdx.x & 31;
    const int wp = (threadIdx.x >> 5) & 1;  // warp within the pair
    const int gl = threadIdx.x & 63;
    const int s = blockIdx.x * PAIRS + pair;
    const bool valid = s < nslots;
    const int barid = 1 + pair;
#define PAIR_SYNC() asm volatile("bar.sync %0, 64;" ::"r"(barid) : "memory")
    if (wp == 0) {
        if (lane == 0) S->n = 0;
        if (valid) {
            const int ord = list[slot0 + s];
            const int gt = P.targets ? P.targets[ord] : ord;
            const int f = gt / P.N, a = gt - f * P.N;
            const FrameGeom &G0 = P.geom[f];
            const double *pos = P.pos + (size_t)f * P.N * 3;
            __syncwarp();
            const double pa[3] = {pos[(size_t)a * 3], pos[(size_t)a * 3 + 1], pos[(size_t)a * 3 + 2]};
            const double Lg[3] = {G0.ortho[0], G0.ortho[4], G0.ortho[8]};
            gather_scan(P, f, a, ord, [&](int j, int tj) {
                const int k = atomicAdd(&S->n, 1);
                if (k < NMAX) {
double dv[3];
                    member_delta(P.periodic, G0, pos, j, pa, dv);
                    S->vec[k][0] = dv[0];
                    S->vec[k][1] = dv[1];
                    S->vec[k][2] = dv[2];
                    S->z[k] = P.Z[tj];
                    S->t[k] = tj;
                    S->idx[k] = j;
                }
            });
            if (lane == 0) {
                S->L[0] = Lg[0];
                S->L[1] = Lg[1];
                S->L[2] = Lg[2];
                S->gp = &G0;
            }
        }
    }
    for (int q = gl; q < NMAX * NMAX; q += 64) S->A[q / NMAX][q % NMAX] = 0.0;
    PAIR_SYNC();
    int n = S->n;
    if (n > NMAX) {
        if (gl == 0) atomicExch(&P.stats->err, MDB_ERR_BOND_OVERFLOW);
        n = NMAX;
    }
    canonical_member_order(S, n, gl, [barid] { asm volatile("bar.sync %0, 64;" ::"r"(barid) : "memory"); });
    for (int i = gl; i < n; i += 64) S->A[i][i] = P.diag[S->t[i]];
    coulomb_pairs(P, S, n, gl, 64);
    PAIR_SYNC();
    double a[NMAX];
    {
        const int col = gl < NMAX ? gl : 0;
#pragma unroll
        for (int i = 0; i < NMAX; ++i) a[i] = gl < NMAX ? S->A[i][col] : 0.0;
    }
    Reg2Ctx<NMAX> cx;
    cx.S = S;
    cx.dptr = dsc + s + (size_t)gl * nslots;
    cx.eptr = esc + s + (size_t)gl * nslots;
    cx.gl = gl;
    cx.lane = lane;
    cx.wp = wp;
    cx.n = n;
    cx.barid = barid;
    cx.valid = valid;
    Reg2Steps<0, NMAX>::run(a, cx);
    double *dptr = cx.dptr;
#undef PAIR_SYNC
    if (valid) {
        if (gl == NMAX - 1 && n == NMAX) *dptr = a[NMAX - 1];
        if (gl == 0) nsc[s] = n;
    }
}

// ---------------------------------------------------------------------------------------------
// K6b: implicit-shift QL on the tridiagonal (d, e), one thread per matrix (EISPACK tql1 / tqli
// without vectors), ascending sort, then the padded sorted feature row.
// ---------------------------------------------------------------------------------------------
template <int NMAX, int THREADS>
__global__ void __launch_bounds__(THREADS) k_tql(const int *__restrict__ list, int slot0, int nslots,
                                                 const double *__restrict__ dsc, const double *__restrict__ esc,
                                                 const int *__restrict__ nsc, const int *__restrict__ counts, int nt,
                                                 const int *__restrict__ padorder,   // element indices by ascending pad value
                                                 const double *__restrict__ padval,  // pad value per element
                                                 const int *__restrict__ maxcount, int D, double *__restrict__ X,
                                                 double *__restrict__ eig, int eigcap, int *__restrict__ nout,
                                                 BondStats *stats, const int *__restrict__ eig_off, int pwk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double(*d)[THREADS] = reinterpret_cast<double(*)[THREADS]>(smem_raw);
    double(*e)[THREADS] = d + NMAX;
    const int tid = threadIdx.x;
    const int s = blockIdx.x * THREADS + tid;
    if (s >= nslots) return;
    const int n = nsc[s];
    for (int k = 0; k < n; ++k) d[k][tid] = dsc[(size_t)k * nslots + s];
    for (int k = 0; k + 1 < n; ++k) e[k][tid] = esc[(size_t)k * nslots + s];
    if (n > 0) e[n - 1][tid] = 0.0;
    // off-diagonals below 1e-12 of their neighbours are dropped: that moves an eigenvalue by at most
    // 1e-12 ||M||, four orders below the floor of the stated tolerance; reciprocals and square roots
    // are MUFU seeds + Newton steps (fast_rcp / fast_rsqrt), a few ulp
    const double EPS = 1e-12;
    if (pwk) {
        // Square-root-free QL (Pal-Walker-Kahan, the recurrence of LAPACK dsterf): the off-diagonals are kept
        // squared, a rotation costs two reciprocals and no square root (~26 instructions against ~45 of the
        // tqli form below).  Same deflation threshold, applied to e^2.
        for (int k = 0; k + 1 < n; ++k) {
            const double v = e[k][tid];
            e[k][tid] = v * v;
        }
        const double EPS2 = EPS * EPS;
        int l = 0;
        int jtot = 0;
        // m = lowest index >= l with a negligible e2[m] (n - 1 when there is none).  The O(n) search for it runs
        // only when l has passed the previous m; otherwise e2[l] is tested at the top of every iteration and every
        // other off-diagonal right when a sweep rewrites it, which is the only time it changes (ncu: searching
        // from l in every iteration cost as many instructions as the rotations themselves).
        int m = -1;
        while (l < n) {
            if (m < l) {
                for (m = l; m < n - 1; ++m) {
                    const double dd = fabs(d[m][tid]) + fabs(d[m + 1][tid]);
                    if (e[m][tid] <= EPS2 * dd * dd) break;
                }
                if (m < n - 1) e[m][tid] = 0.0;
            }
            if (m == l) {
                ++l;
                continue;
            }
            {
                const double dd = fabs(d[l][tid]) + fabs(d[l + 1][tid]);
                if (e[l][tid] <= EPS2 * dd * dd) {  // l has converged; m stays valid for l + 1
                    e[l][tid] = 0.0;
                    ++l;
                    continue;
                }
            }
            if (jtot++ > 80 * NMAX) {
                atomicExch(&stats->err, MDB_ERR_EIG_NOCONV);
                break;
            }
            const double p0 = d[l][tid];
            const double e2l = e[l][tid];
            const double rte = e2l * fast_rsqrt(e2l);  // sqrt(e2[l]) > 0 here
            double sigma = (d[l + 1][tid] - p0) * (0.5 * fast_rcp(rte));
            const double t1 = fma(sigma, sigma, 1.0);
            const double r1 = t1 * fast_rsqrt(t1);
            sigma = p0 - rte * fast_rcp(sigma + (sigma >= 0.0 ? r1 : -r1));
            double c = 1.0, sn = 0.0;
            double gamma = d[m][tid] - sigma;
            double p = gamma * gamma;
            double dup = fabs(d[m][tid]);  // |d[i + 2]| as the sweep left it
            int msplit = -1;
            for (int i = m - 1; i >= l; --i) {
                const double bb = e[i][tid];
                const double r = p + bb;
                const double enew = sn * r;  // new e2[i + 1]
                if (i != m - 1) e[i + 1][tid] = enew;
                const double oldc = c;
                // 1 / c = r / p: both reciprocals depend on p only and run side by side (the chain of one
                // rotation holds one MUFU + Newton sequence instead of two)
                const double ir = fast_rcp(r);
                const double ip = fast_rcp(p);
                const bool pz = p == 0.0;
                c = p * ir;
                sn = bb * ir;
                const double oldgam = gamma;
                const double alpha = d[i][tid];
                gamma = c * (alpha - sigma) - sn * oldgam;
                const double dn = oldgam + (alpha - gamma);
                d[i + 1][tid] = dn;
                const double dd = fabs(dn) + dup;
                if (i != m - 1 && enew <= EPS2 * dd * dd) msplit = i + 1;
                dup = fabs(dn);
                p = pz ? oldc * bb : gamma * gamma * (r * ip);
            }
            e[l][tid] = sn * p;
            d[l][tid] = sigma + gamma;
            if (msplit >= 0) {
                e[msplit][tid] = 0.0;
                m = msplit;
            }
        }
    } else {
    // m = lowest index >= l of a negligible off-diagonal.  The O(n) search for it runs only at the
    // start (and after an exact-zero stop); afterwards e[l] is tested at the top of every iteration
    // and every other off-diagonal right when a sweep rewrites it, which is the only time it changes.
    int m = -1;
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        for (;;) {
            if (m < l) {
                for (m = l; m < n - 1; ++m) {
                    const double dd = fabs(d[m][tid]) + fabs(d[m + 1][tid]);
                    if (fabs(e[m][tid]) <= EPS * dd) break;
                }
            }
            if (m == l) break;
            // l < m <= n - 1: l has converged when e[l] is negligible; m stays valid for l + 1
            if (fabs(e[l][tid]) <= EPS * (fabs(d[l][tid]) + fabs(d[l + 1][tid]))) break;
            if (iter++ == 80) {
                atomicExch(&stats->err, MDB_ERR_EIG_NOCONV);
                break;
            }
            const double dl = d[l][tid], el = e[l][tid];
            double g = (d[l + 1][tid] - dl) * (0.5 * fast_rcp(el));
            const double t1 = fma(g, g, 1.0);
            double r = t1 * fast_rsqrt(t1);
            g = d[m][tid] - dl + el * fast_rcp(g + (g >= 0.0 ? r : -r));
            double sn = 1.0, c = 1.0, p = 0.0;
            double dup = 0.0;  // d[i + 2] as the sweep left it
            int i, msplit = -1;
            for (i = m - 1; i >= l; --i) {
                const double ei = e[i][tid];
                double f = sn * ei;
                const double b = c * ei;
                const double t2 = fma(f, f, g * g);
                if (t2 == 0.0) {
                    r = 0.0;
                    e[i + 1][tid] = 0.0;
                    d[i + 1][tid] -= p;
                    e[m][tid] = 0.0;
                    break;
                }
                const double ir = fast_rsqrt(t2);
                r = t2 * ir;
                e[i + 1][tid] = r;
                sn = f * ir;
                c = g * ir;
                g = d[i + 1][tid] - p;
                r = (d[i][tid] - g) * sn + 2.0 * c * b;
                p = sn * r;
                const double dn = g + p;
                d[i + 1][tid] = dn;
                if (i + 1 < m && t2 <= (EPS * EPS) * (fabs(dn) + fabs(dup)) * (fabs(dn) + fabs(dup))) msplit = i + 1;
                dup = dn;
                g = c * r - b;
            }
            if (r == 0.0 && i >= l) {
                m = -1;  // the sweep stopped early at an exact zero: search again
                continue;
            }
            d[l][tid] -= p;
            e[l][tid] = g;
            e[m][tid] = 0.0;
            if (msplit >= 0) m = msplit;
        }
    }
    }
    // ascending insertion sort
    for (int a = 1; a < n; ++a) {
        const double v = d[a][tid];
        int b = a - 1;
        while (b >= 0 && d[b][tid] > v) {
            d[b + 1][tid] = d[b][tid];
            --b;
        }
        d[b + 1][tid] = v;
    }
    const int ord = list[slot0 + s];
    if (nout) nout[ord] = n;
    if (eig && eig_off) {  // compact spectra: target t owns eig[eig_off[t] .. eig_off[t + 1])
        double *o = eig + eig_off[ord];
        for (int k = 0; k < n; ++k) o[k] = d[k][tid];
    } else if (eig) {
        double *o = eig + (size_t)ord * eigcap;
        for (int k = 0; k < eigcap; ++k) o[k] = k < n ? d[k][tid] : __longlong_as_double(0x7ff8000000000000LL);
    }
    if (X) {
        // merge the sorted spectrum with the sorted pad multiset (closed form of :236-276)
        double *o = X + (size_t)ord * D;
        int w = 0, k = 0;
        for (int q = 0; q < nt; ++q) {
            const int el = padorder[q];
            const int reps = maxcount[el] - counts[(size_t)ord * nt + el];
            const double pv = padval[el];
            for (int r = 0; r < reps; ++r) {
                while (k < n && d[k][tid] <= pv && w < D) o[w++] = d[k++][tid];
                if (w < D) o[w++] = pv;
            }
        }
        while (k < n && w < D) o[w++] = d[k++][tid];
        while (w < D) o[w++] = 0.0;
    }
}

// bucket id by cluster size: classes of 4 up to 48 atoms, of 8 up to 64 (register-resident kernels), then 96, 128 and 160
#define DESC_NB 16
__device__ __host__ __forceinline__ int desc_bucket(int n) {
    return n <= 8 ? 0 : n <= 48 ? (n - 5) >> 2 : n <= 64 ? 11 + ((n - 49) >> 3) : n <= 96 ? 13 : n <= 128 ? 14 : n <= 160 ? 15 : DESC_NB;
}

__global__ void __launch_bounds__(256) k_bucket_count(const int *__restrict__ counts, int nt, long long ntargets,
                                                      int *__restrict__ bucket_n, int *__restrict__ nsum) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntargets) return;
    int n = 0;
    for (int k = 0; k < nt; ++k) n += counts[t * nt + k];
    nsum[t] = n;
    atomicAdd(&bucket_n[n > GLB_NMAX ? 162 : n > 160 ? 161 : n], 1);  // by exact size; 161: the global-memory class, 162: too large
}

__global__ void __launch_bounds__(256) k_bucket_fill(const int *__restrict__ nsum, long long ntargets,
                                                     const int *__restrict__ bucket_ofs, int *__restrict__ cursor,
                                                     int *__restrict__ list) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntargets) return;
    // matrices of equal size sit together (ascending size, so a size class is still one contiguous range): the
    // threads of a QL warp and the warps of a bordered block then run loops of equal length
    const int b = nsum[t] > GLB_NMAX ? 162 : min(nsum[t], 161);
    const int k = atomicAdd(&cursor[b], 1);
    list[bucket_ofs[b] + k] = (int)t;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int desc_setup(mdb_ctx *c, int F, int N, const double *d_pos, const double *h_cell, const uint8_t *d_types,
                      int periodic, double cutoff, const int *d_targets, long long ntargets, cudaStream_t stream,
                      DescParams *P, size_t extra_bytes, bool need_cells = true) {
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
    if (build_geometry(c, F, N, h_cell, periodic, bbox.data(), cutoff * 1.0005, cutoff * 1.0005, geom, &total_cells))
        return -1;
    double lmax = 0;
    bool all_ortho = true;
    for (const FrameGeom &g : geom) {
        all_ortho = all_ortho && g.is_ortho;
        for (int v = 0; v < 3; ++v) {
            const double len = sqrt(g.ortho[v] * g.ortho[v] + g.ortho[3 + v] * g.ortho[3 + v] + g.ortho[6 + v] * g.ortho[6 + v]);
            lmax = std::max(lmax, len);
            // perpendicular height of the cell along axis v = 1 / |row v of the inverse|
            const double *fr = g.frac + v * 3;
            const double height = 1.0 / sqrt(fr[0] * fr[0] + fr[1] * fr[1] + fr[2] * fr[2]);
            // beyond half a cell height a sphere meets two images of one atom, which the reference
            // (ASE get_distances, mic=True) counts once at its nearer image: not what the scan does
            if (periodic && 2.0 * cutoff > height) {
                mdb_set_error("cutoff " + std::to_string(cutoff) + " A exceeds half of a cell height (" + std::to_string(height) +
                              " A): not supported on the descriptor path");
                return -1;
            }
        }
    }
    size_t need = align256(geom.size() * sizeof(FrameGeom)) + (need_cells ? cells_workspace_bytes(F, N, total_cells) : 0) +
                  extra_bytes + 4096;
    if (arena_reserve(c, need, stream)) return -1;
    arena_reset(c);
    FrameGeom *d_geom = (FrameGeom *)arena_alloc(c, geom.size() * sizeof(FrameGeom));
    if (!d_geom) return -1;
    if (upload_geometry(c, geom, d_geom, stream)) return -1;
    CellList cl;
    memset(&cl, 0, sizeof(cl));
    const int rec_mode = all_ortho ? 1 : 0;
    if (need_cells && cells_build(c, F, N, d_pos, d_types, d_geom, total_cells, periodic, rec_mode, &cl, &c->d_stats->err, stream))
        return -1;
    P->rec_mode = rec_mode;
    P->rec = cl.rec;
    P->ccell = cl.ccell;
    P->cr = cl.cr;
    P->cell_start = cl.cell_start;
    P->members = nullptr;
    P->members_out = nullptr;
    P->memcap = 0;
    P->geom = d_geom;
    P->pos = d_pos;
    P->types = d_types;
    P->targets = d_targets;
    // NULL = every atom of every frame; NULL with ntargets == 0 is an empty list (nothing to do)
    P->ntargets = d_targets ? ntargets : (ntargets == 0 ? 0 : (long long)F * N);
    P->N = N;
    P->periodic = periodic;
    P->cutoff = cutoff;
    P->cut2f = (float)(cutoff * cutoff);
    P->margin = (float)(1.25e-6 * cutoff * lmax + 1e-5) * 2.0f;
    P->nt = c->el.nt;
    for (int k = 0; k < MDB_MAXT; ++k) {
        P->diag[k] = c->el.diag[k];
        P->Z[k] = c->el.Z[k];
    }
    P->stats = c->d_stats;
    return 0;
}

// Padded sorted feature rows from stored spectra (the closed form of the parent loop,
// datasetbuilder.py:236-276): row = sort(eig U {pad_el repeated (max_el - count_el) times}).  One
// thread per row; the same merge as at the end of k_tql, for callers that keep the eigenvalues of a
// batch until the per-element maxima over ALL rows are known.
__global__ void __launch_bounds__(128) k_feature_rows(long long T, const double *__restrict__ eig, int eigcap,
                                                      const int *__restrict__ nn, const int *__restrict__ counts, int nt,
                                                      const int *__restrict__ padorder, const double *__restrict__ padval,
                                                      const int *__restrict__ maxcount, int D, double *__restrict__ X) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int n = min(nn[t], eigcap);
    const double *d = eig + (size_t)t * eigcap;
    double *o = X + (size_t)t * D;
    int w = 0, k = 0;
    for (int q = 0; q < nt; ++q) {
        const int el = padorder[q];
        const int reps = maxcount[el] - counts[(size_t)t * nt + el];
        const double pv = padval[el];
        for (int r = 0; r < reps; ++r) {
            while (k < n && d[k] <= pv && w < D) o[w++] = d[k++];
            if (w < D) o[w++] = pv;
        }
    }
    while (k < n && w < D) o[w++] = d[k++];
    while (w < D) o[w++] = 0.0;
}

// One warp per row: the row is the merge of the ascending spectrum with the ascending pad values, the spectrum first on
// ties (the sequential loop of k_feature_rows), so eigenvalue k lands at k + (pads below it) and the pads of class q start
// at (pads of earlier classes) + (eigenvalues <= their value).  Reads and writes are contiguous per row.
__global__ void __launch_bounds__(256) k_feature_rows_compact(long long T, const double *__restrict__ eig,
                                                              const int *__restrict__ off, const int *__restrict__ counts, int nt,
                                                              const int *__restrict__ padorder, const double *__restrict__ padval,
                                                              const int *__restrict__ maxcount, int D, double *__restrict__ X) {
    __shared__ double s_pv[MDB_MAXT];
    __shared__ int s_el[MDB_MAXT], s_mc[MDB_MAXT];
    if (threadIdx.x < nt) {
        const int el = padorder[threadIdx.x];
        s_el[threadIdx.x] = el;
        s_pv[threadIdx.x] = padval[el];
        s_mc[threadIdx.x] = maxcount[el];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= T) return;
    const int o0 = off[t], n = off[t + 1] - o0;
    const double *d = eig + o0;
    double *o = X + (size_t)t * D;
    // pads of class q (lane q < nt): how many, and how many pads precede the class
    int reps = lane < nt ? s_mc[lane] - counts[(size_t)t * nt + s_el[lane]] : 0;
    int before = reps;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, before, s);
        if (lane >= s) before += v;
    }
    before -= reps;
    int le = 0;  // lane q: eigenvalues <= the pad value of class q
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        const double v = k < n ? d[k] : 0.0;
        int pads_below = 0;
        for (int q = 0; q < nt; ++q) {
            const int rq = __shfl_sync(0xffffffffu, reps, q);
            const bool lei = k < n && v <= s_pv[q];
            if (k < n && !lei) pads_below += rq;
            const unsigned bl = __ballot_sync(0xffffffffu, lei);
            if (lane == q) le += __popc(bl);
        }
        if (k < n && k + pads_below < D) o[k + pads_below] = v;
    }
    for (int q = 0; q < nt; ++q) {
        const int rq = __shfl_sync(0xffffffffu, reps, q);
        const int start = __shfl_sync(0xffffffffu, before + le, q);
        const double pv = s_pv[q];
        for (int r = lane; r < rq; r += 32)
            if (start + r < D) o[start + r] = pv;
    }
    // (n + pads = D by construction: D = sum(maxcount), n = sum(counts))
    const int total = n + __shfl_sync(0xffffffffu, before + reps, 31);
    for (int w = total + lane; w < D; w += 32) o[w] = 0.0;
}

int feature_rows_run(mdb_ctx *c, long long T, const double *d_eig, int eigcap, const int *d_n, const int *d_counts,
                     const int *h_maxcount, double *d_X, cudaStream_t stream, const int *d_eig_off) {
    if (T <= 0) return 0;
    const int nt = c->el.nt;
    int D = 0;
    int order[MDB_MAXT];
    for (int k = 0; k < nt; ++k) {
        D += h_maxcount[k];
        order[k] = k;
    }
    for (int x = 1; x < nt; ++x) {  // elements by ascending pad value
        int v = order[x], y = x - 1;
        while (y >= 0 && c->el.diag[order[y]] > c->el.diag[v]) {
            order[y + 1] = order[y];
            --y;
        }
        order[y + 1] = v;
    }
    if (arena_reserve(c, 4096, stream)) return -1;
    arena_reset(c);
    int *d_maxcount = (int *)arena_alloc(c, 256);
    int *d_padorder = (int *)arena_alloc(c, 256);
    double *d_padval = (double *)arena_alloc(c, 256);
    if (!d_maxcount || !d_padorder || !d_padval) return -1;
    MDB_CUDA(cudaMemcpyAsync(d_maxcount, h_maxcount, nt * sizeof(int), cudaMemcpyHostToDevice, stream));
    MDB_CUDA(cudaMemcpyAsync(d_padorder, order, nt * sizeof(int), cudaMemcpyHostToDevice, stream));
    MDB_CUDA(cudaMemcpyAsync(d_padval, c->el.diag, nt * sizeof(double), cudaMemcpyHostToDevice, stream));
    {
        ProfScope ps(c, "k_feature_rows", stream);
        if (d_eig_off)
            k_feature_rows_compact<<<cdiv(T, 8), 256, 0, stream>>>(T, d_eig, d_eig_off, d_counts, nt, d_padorder, d_padval,
                                                                    d_maxcount, D, d_X);
        else
            k_feature_rows<<<cdiv(T, 128), 128, 0, stream>>>(T, d_eig, eigcap, d_n, d_counts, nt, d_padorder, d_padval,
                                                            d_maxcount, D, d_X);
    }
    c->launches++;
    MDB_CUDA(cudaGetLastError());
    MDB_CUDA(cudaStreamSynchronize(stream));  // the host arrays above live on the stack
    return 0;
}

int compositions_run(mdb_ctx *c, int F, int N, const double *d_pos, const double *h_cell, const uint8_t *d_types,
                     int periodic, double cutoff, const int *d_targets, long long ntargets, int *d_counts,
                     int *h_maxcount, cudaStream_t stream, int memcap, int *d_members_out) {
    DescParams P;
    if (desc_setup(c, F, N, d_pos, h_cell, d_types, periodic, cutoff, d_targets, ntargets, stream, &P, 1024)) return -1;
    if (d_members_out) {
        if (memcap < 1) {
            mdb_set_error("mdb_compositions_members: cap must be positive");
            return -1;
        }
        P.members_out = d_members_out;
        P.memcap = memcap;
    }
    int *d_max = (int *)arena_alloc(c, MDB_MAXT * sizeof(int));
    if (!d_max) return -1;
    MDB_CUDA(cudaMemcpyAsync(d_max, h_maxcount, c->el.nt * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (P.ntargets > 0) {
        ProfScope ps(c, "k_composition", stream);
        if (P.rec_mode == 0)
            k_composition<true><<<cdiv(P.ntargets, 8), 256, 0, stream>>>(P, d_counts, d_max);
        else
            k_composition<false><<<cdiv(P.ntargets, 8), 256, 0, stream>>>(P, d_counts, d_max);
        c->launches++;
    }
    MDB_CUDA(cudaGetLastError());
    MDB_CUDA(cudaMemcpyAsync(h_maxcount, d_max, c->el.nt * sizeof(int), cudaMemcpyDeviceToHost, stream));
    MDB_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

int members_run(mdb_ctx *c, int F, int N, const double *d_pos, const double *h_cell, const uint8_t *d_types,
                int periodic, double cutoff, const int *d_targets, long long ntargets, int cap, int *d_members,
                int *d_nmem, cudaStream_t stream) {
    if (cap < 1 || cap > 1024) {
        mdb_set_error("mdb_sphere_members: cap must be in 1..1024");
        return -1;
    }
    DescParams P;
    if (desc_setup(c, F, N, d_pos, h_cell, d_types, periodic, cutoff, d_targets, ntargets, stream, &P, 1024)) return -1;
    if (P.ntargets > 0) {
        ProfScope ps(c, "k_members", stream);
        if (P.rec_mode == 0)
            k_members<true><<<cdiv(P.ntargets, 8), 256, (size_t)8 * cap * sizeof(int), stream>>>(P, cap, d_members, d_nmem);
        else
            k_members<false><<<cdiv(P.ntargets, 8), 256, (size_t)8 * cap * sizeof(int), stream>>>(P, cap, d_members, d_nmem);
        c->launches++;
    }
    MDB_CUDA(cudaGetLastError());
    return 0;
}

static const char *bucket_name(int nmax, bool tql) {
    switch (nmax) {
        case 8: return tql ? "k_tql<8>" : "k_tridiag_reg<8>";
        case 12: return tql ? "k_tql<12>" : "k_tridiag_reg<12>";
        case 16: return tql ? "k_tql<16>" : "k_tridiag_reg<16>";
        case 20: return tql ? "k_tql<20>" : "k_tridiag_reg<20>";
        case 24: return tql ? "k_tql<24>" : "k_tridiag_reg<24>";
        case 28: return tql ? "k_tql<28>" : "k_tridiag_reg<28>";
        case 32: return tql ? "k_tql<32>" : "k_tridiag_reg<32>";
        case 36: return tql ? "k_tql<36>" : "k_tridiag_reg2<36>";
        case 40: return tql ? "k_tql<40>" : "k_tridiag_reg2<40>";
        case 44: return tql ? "k_tql<44>" : "k_tridiag_reg2<44>";
        case 48: return tql ? "k_tql<48>" : "k_tridiag_reg2<48>";
        case 56: return tql ? "k_tql<56>" : "k_tridiag_reg2<56>";
        case 64: return tql ? "k_tql<64>" : "k_tridiag_reg2<64>";
        case 96: return tql ? "k_tql<96>" : "k_tridiag_blk<96>";
        case 128: return tql ? "k_tql<128>" : "k_tridiag_blk<128>";
        default: return tql ? "k_tql<160>" : "k_tridiag_blk<160>";
    }
}

template <int NMAX, int WARPS, int LPM, int QTHREADS, int REG>
static int run_bucket(mdb_ctx *c, const DescParams &P, const int *d_list, int ofs, int cnt, double *dsc, double *esc,
                      int *nsc, int chunk, const int *d_counts, const int *d_padorder, const double *d_padval,
                      const int *d_maxcount, int D, double *d_X, double *d_eig, int eigcap, int *d_n, cudaStream_t stream) {
    constexpr int G = 32 / LPM;
    size_t smem_tri;
    if constexpr (REG == 1) smem_tri = sizeof(RegSmem<NMAX, LPM>) * WARPS * G;
    else if constexpr (REG == 2) smem_tri = sizeof(Reg2Smem<NMAX>) * (WARPS / 2);
    else smem_tri = sizeof(BlkSmem<NMAX>);
    const size_t smem_ql = (size_t)2 * NMAX * QTHREADS * sizeof(double);
    if constexpr (REG == 1) {
        MDB_CUDA(cudaFuncSetAttribute(k_tridiag_reg<NMAX, LPM, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tri));
    } else if constexpr (REG == 2) {
        MDB_CUDA(cudaFuncSetAttribute(k_tridiag_reg2<NMAX, WARPS / 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tri));
    } else {
        MDB_CUDA(cudaFuncSetAttribute(k_tridiag_blk<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tri));
    }
    MDB_CUDA(cudaFuncSetAttribute(k_tql<NMAX, QTHREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ql));
    for (int s0 = 0; s0 < cnt; s0 += chunk) {
        const int ns = std::min(chunk, cnt - s0);
        {
            ProfScope ps(c, bucket_name(NMAX, false), stream);
            if constexpr (REG == 1) {
                k_tridiag_reg<NMAX, LPM, WARPS><<<cdiv(ns, WARPS * G), WARPS * 32, smem_tri, stream>>>(P, d_list, cnt, ofs + s0, ns,
                                                                                                   dsc, esc, nsc);
            } else if constexpr (REG == 2) {
                k_tridiag_reg2<NMAX, WARPS / 2><<<cdiv(ns, WARPS / 2), WARPS * 32, smem_tri, stream>>>(P, d_list, cnt, ofs + s0, ns,
                                                                                                   dsc, esc, nsc);
            } else {
                k_tridiag_blk<NMAX><<<ns, 2 * NMAX, smem_tri, stream>>>(P, d_list, cnt, ofs + s0, ns, dsc, esc, nsc);
            }
        }
        {
            ProfScope ps(c, bucket_name(NMAX, true), stream);
            k_tql<NMAX, QTHREADS><<<cdiv(ns, QTHREADS), QTHREADS, smem_ql, stream>>>(
                d_list, ofs + s0, ns, dsc, esc, nsc, d_counts, P.nt, d_padorder, d_padval, d_maxcount, D, d_X, d_eig, eigcap,
                d_n, c->d_stats, c->cur_eig_off, c->opt_ql_pwk);
        }
        c->launches += 2;
    }
    MDB_CUDA(cudaGetLastError());
    return 0;
}

static const char *mc_name(int nmax) {
    switch (nmax) {
        case 8: return "k_tridiag_mc<8>";
        case 12: return "k_tridiag_mc<12>";
        case 16: return "k_tridiag_mc<16>";
        case 20: return "k_tridiag_mc<20>";
        case 24: return "k_tridiag_mc<24>";
        case 28: return "k_tridiag_mc<28>";
        case 32: return "k_tridiag_mc<32>";
        case 36: return "k_tridiag_mc<36>";
        case 40: return "k_tridiag_mc<40>";
        case 44: return "k_tridiag_mc<44>";
        default: return "k_tridiag_mc<48>";
    }
}

// multi-column register kernel + QL for one size class
template <int NMAX, int LPM, int CPL, int WARPS, int MINB, int QTHREADS>
static int run_bucket_mc(mdb_ctx *c, const DescParams &P, const int *d_list, int ofs, int cnt, double *dsc, double *esc,
                         int *nsc, int chunk, const int *d_counts, const int *d_padorder, const double *d_padval,
                         const int *d_maxcount, int D, double *d_X, double *d_eig, int eigcap, int *d_n, cudaStream_t stream) {
    constexpr int G = 32 / LPM;
    const size_t smem_tri = sizeof(McSmem<NMAX, LPM, CPL>) * WARPS * G;
    const size_t smem_ql = (size_t)2 * NMAX * QTHREADS * sizeof(double);
    MDB_CUDA(cudaFuncSetAttribute(k_tridiag_mc<NMAX, LPM, CPL, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_tri));
    MDB_CUDA(cudaFuncSetAttribute(k_tql<NMAX, QTHREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ql));
    for (int s0 = 0; s0 < cnt; s0 += chunk) {
        const int ns = std::min(chunk, cnt - s0);
        {
            ProfScope ps(c, mc_name(NMAX), stream);
            k_tridiag_mc<NMAX, LPM, CPL, WARPS, MINB><<<cdiv(ns, WARPS * G), WARPS * 32, smem_tri, stream>>>(
                P, d_list, cnt, ofs + s0, ns, dsc, esc, nsc);
        }
        {
            ProfScope ps(c, bucket_name(NMAX, true), stream);
            k_tql<NMAX, QTHREADS><<<cdiv(ns, QTHREADS), QTHREADS, smem_ql, stream>>>(
                d_list, ofs + s0, ns, dsc, esc, nsc, d_counts, P.nt, d_padorder, d_padval, d_maxcount, D, d_X, d_eig, eigcap,
                d_n, c->d_stats, c->cur_eig_off, c->opt_ql_pwk);
        }
        c->launches += 2;
    }
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// global-memory class (161..512 atoms): the matrix scratch is allocated for this call only (two blocks per SM in
// flight); the tridiagonals of up to `chunk * 160 / 512` matrices are collected in d/e before one QL launch
static int run_bucket_glb(mdb_ctx *c, const DescParams &P, const int *d_list, int ofs, int cnt, double *dsc, double *esc,
                          int *nsc, int chunk, const int *d_counts, const int *d_padorder, const double *d_padval,
                          const int *d_maxcount, int D, double *d_X, double *d_eig, int eigcap, int *d_n, cudaStream_t stream) {
    constexpr int QT = 8;
    const int super = (int)((size_t)chunk * 160 / GLB_NMAX);
    const int nb = std::min(cnt, 2 * 148);
    double *Ascr = nullptr;
    MDB_CUDA(cudaMalloc(&Ascr, (size_t)nb * GLB_NMAX * GLB_NMAX * sizeof(double)));
    const size_t smem_ql = (size_t)2 * GLB_NMAX * QT * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(k_tql<GLB_NMAX, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ql);
    for (int s0 = 0; s0 < cnt && e == cudaSuccess; s0 += super) {
        const int ns = std::min(super, cnt - s0);
        {
            ProfScope ps(c, "k_tridiag_glb", stream);
            for (int b0 = 0; b0 < ns; b0 += nb) {
                k_tridiag_glb<<<std::min(nb, ns - b0), GLB_THREADS, 0, stream>>>(P, d_list, ofs + s0, b0, ns, Ascr, dsc, esc, nsc);
                c->launches++;
            }
        }
        {
            ProfScope ps(c, "k_tql<512>", stream);
            k_tql<GLB_NMAX, QT><<<cdiv(ns, QT), QT, smem_ql, stream>>>(d_list, ofs + s0, ns, dsc, esc, nsc, d_counts, P.nt, d_padorder,
                                                                     d_padval, d_maxcount, D, d_X, d_eig, eigcap, d_n, c->d_stats,
                                                                     c->cur_eig_off, c->opt_ql_pwk);
            c->launches++;
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(Ascr);
    MDB_CUDA(e);
    return 0;
}

// bordered register kernel + QL for the 33..40-atom classes
template <int NS, int WARPS, int QTHREADS>
static int run_bucket_bord(mdb_ctx *c, const DescParams &P, const int *d_list, int ofs, int cnt, double *dsc, double *esc,
                           int *nsc, int chunk, const int *d_counts, const int *d_padorder, const double *d_padval,
                           const int *d_maxcount, int D, double *d_X, double *d_eig, int eigcap, int *d_n, cudaStream_t stream) {
    const size_t smem_tri = sizeof(BordSmem<NS>) * WARPS;
    const size_t smem_ql = (size_t)2 * NS * QTHREADS * sizeof(double);
    MDB_CUDA(cudaFuncSetAttribute(k_tridiag_bord<NS, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tri));
    MDB_CUDA(cudaFuncSetAttribute(k_tql<NS, QTHREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ql));
    for (int s0 = 0; s0 < cnt; s0 += chunk) {
        const int ns = std::min(chunk, cnt - s0);
        {
            ProfScope ps(c, NS == 36 ? "k_tridiag_bord<36>" : "k_tridiag_bord<40>", stream);
            k_tridiag_bord<NS, WARPS><<<cdiv(ns, WARPS), WARPS * 32, smem_tri, stream>>>(P, d_list, cnt, ofs + s0, ns, dsc, esc, nsc);
        }
        {
            ProfScope ps(c, bucket_name(NS, true), stream);
            k_tql<NS, QTHREADS><<<cdiv(ns, QTHREADS), QTHREADS, smem_ql, stream>>>(
                d_list, ofs + s0, ns, dsc, esc, nsc, d_counts, P.nt, d_padorder, d_padval, d_maxcount, D, d_X, d_eig, eigcap,
                d_n, c->d_stats, c->cur_eig_off, c->opt_ql_pwk);
        }
        c->launches += 2;
    }
    MDB_CUDA(cudaGetLastError());
    return 0;
}

int descriptors_run(mdb_ctx *c, int F, int N, const double *d_pos, const double *h_cell, const uint8_t *d_types,
                    int periodic, double cutoff, const int *d_targets, long long ntargets, const int *d_counts,
                    const int *h_maxcount, double *d_X, double *d_eig, int eigcap, int *d_n, cudaStream_t stream,
                    const int *d_members, int memcap, int *d_eig_off) {
    const int nt = c->el.nt;
    c->cur_eig_off = nullptr;
    const long long T = d_targets ? ntargets : (ntargets == 0 ? 0 : (long long)F * N);
    if (T == 0) return 0;
    if (T >= (1LL << 31)) {
        mdb_set_error("too many targets in one call");
        return -1;
    }
    if (!d_members && periodic && h_cell) {
        // general (triclinic) cells: the register kernels carry the diagonal-cell scan only; the membership comes
        // from the composition pass (k_composition<true>) through a scratch list, then the member route runs
        bool general = false;
        for (int f = 0; f < F && !general; ++f)
            for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 3; ++q)
                    if (r != q && h_cell[(size_t)f * 9 + r * 3 + q] != 0.0) general = true;
        if (general) {
            int cap = 0;
            for (int k = 0; k < nt; ++k) cap += h_maxcount[k];
            cap = std::max(cap, 1);
            int *tmp_members = nullptr, *tmp_counts = nullptr;
            MDB_CUDA(cudaMalloc(&tmp_members, (size_t)T * cap * sizeof(int)));
            MDB_CUDA(cudaMalloc(&tmp_counts, (size_t)T * nt * sizeof(int)));
            int hmax[MDB_MAXT] = {0};
            int rc = compositions_run(c, F, N, d_pos, h_cell, d_types, periodic, cutoff, d_targets, ntargets, tmp_counts, hmax,
                                      stream, cap, tmp_members);
            if (!rc)
                rc = descriptors_run(c, F, N, d_pos, h_cell, d_types, periodic, cutoff, d_targets, ntargets, d_counts,
                                     h_maxcount, d_X, d_eig, eigcap, d_n, stream, tmp_members, cap, d_eig_off);
            cudaStreamSynchronize(stream);
            cudaFree(tmp_members);
            cudaFree(tmp_counts);
            return rc;
        }
    }
    const int chunk = 1 << 16;
    // scratch: bucket bookkeeping + lists + d/e for one chunk at the largest NMAX in use
    size_t extra = align256(T * sizeof(int)) * 2 + align256((size_t)chunk * 160 * sizeof(double)) * 2 +
                   align256((size_t)chunk * sizeof(int)) + 32 * 256 + (d_eig_off ? align256(scan_scratch_bytes(T + 1)) : 0);
    DescParams P;
    if (desc_setup(c, F, N, d_pos, h_cell, d_types, periodic, cutoff, d_targets, ntargets, stream, &P, extra, d_members == nullptr))
        return -1;
    P.members = d_members;  // membership from the composition pass: no cell list, no second scan
    P.memcap = d_members ? memcap : 0;
    int *d_nsum = (int *)arena_alloc(c, T * sizeof(int));
    int *d_list = (int *)arena_alloc(c, T * sizeof(int));
    double *dsc = (double *)arena_alloc(c, (size_t)chunk * 160 * sizeof(double));
    double *esc = (double *)arena_alloc(c, (size_t)chunk * 160 * sizeof(double));
    int *nsc = (int *)arena_alloc(c, (size_t)chunk * sizeof(int));
    int *d_book = (int *)arena_alloc(c, 3 * 192 * sizeof(int));   // per sphere size: count[192] | cursor[192] | first position[192]
    int *d_maxcount = (int *)arena_alloc(c, 256);
    int *d_padorder = (int *)arena_alloc(c, 256);
    double *d_padval = (double *)arena_alloc(c, 256);
    if (!d_nsum || !d_list || !dsc || !esc || !nsc || !d_book || !d_maxcount || !d_padorder || !d_padval) {
        mdb_set_error("descriptors_run: arena too small (internal sizing error)");
        return -1;
    }
    int D = 0;
    int order[MDB_MAXT];
    for (int k = 0; k < nt; ++k) {
        D += h_maxcount[k];
        order[k] = k;
    }
    for (int x = 1; x < nt; ++x) {  // elements by ascending pad value
        int v = order[x], y = x - 1;
        while (y >= 0 && c->el.diag[order[y]] > c->el.diag[v]) {
            order[y + 1] = order[y];
            --y;
        }
        order[y + 1] = v;
    }
    MDB_CUDA(cudaMemsetAsync(d_book, 0, 3 * 192 * sizeof(int), stream));
    MDB_CUDA(cudaMemcpyAsync(d_maxcount, h_maxcount, nt * sizeof(int), cudaMemcpyHostToDevice, stream));
    MDB_CUDA(cudaMemcpyAsync(d_padorder, order, nt * sizeof(int), cudaMemcpyHostToDevice, stream));
    MDB_CUDA(cudaMemcpyAsync(d_padval, c->el.diag, nt * sizeof(double), cudaMemcpyHostToDevice, stream));
    k_bucket_count<<<cdiv(T, 256), 256, 0, stream>>>(d_counts, nt, T, d_book, d_nsum);
    c->launches++;
    if (d_eig_off) {
        // compact spectra: offsets = exclusive scan of the sphere sizes (caller sized d_eig with sum(counts))
        void *tmp = arena_alloc(c, scan_scratch_bytes(T + 1));
        if (!tmp) {
            mdb_set_error("descriptors_run: arena too small (internal sizing error)");
            return -1;
        }
        MDB_CUDA(cudaMemcpyAsync(d_eig_off, d_nsum, T * sizeof(int), cudaMemcpyDeviceToDevice, stream));
        MDB_CUDA(cudaMemsetAsync(d_eig_off + T, 0, sizeof(int), stream));
        if (device_exclusive_scan(c, d_eig_off, T + 1, tmp, stream)) return -1;
        c->cur_eig_off = d_eig_off;
    }
    int hn[192], hb[DESC_NB + 1];
    MDB_CUDA(cudaMemcpyAsync(hn, d_book, sizeof(hn), cudaMemcpyDeviceToHost, stream));
    MDB_CUDA(cudaStreamSynchronize(stream));
    int ofs_n[192];
    memset(hb, 0, sizeof(hb));
    {
        int run = 0;
        for (int q = 0; q < 192; ++q) {
            ofs_n[q] = run;
            run += hn[q];
            if (q <= 161) hb[q == 161 ? DESC_NB : desc_bucket(q)] += hn[q];
        }
    }
    if (hn[162] > 0) {
        mdb_set_error("a cutoff sphere holds more than 512 atoms; reduce the cutoff");
        return -1;
    }
    int ofs[DESC_NB + 1];
    ofs[0] = 0;
    for (int b = 0; b < DESC_NB; ++b) ofs[b + 1] = ofs[b] + hb[b];
    MDB_CUDA(cudaMemcpyAsync(d_book + 384, ofs_n, sizeof(ofs_n), cudaMemcpyHostToDevice, stream));
    k_bucket_fill<<<cdiv(T, 256), 256, 0, stream>>>(d_nsum, T, d_book + 384, d_book + 192, d_list);
    c->launches++;
#define MDB_BUCKET(B, NMAX, WARPS, LPM, QT, REG)                                                                              \
    if (hb[B] && run_bucket<NMAX, WARPS, LPM, QT, REG>(c, P, d_list, ofs[B], hb[B], dsc, esc, nsc, chunk, d_counts, d_padorder, \
                                                       d_padval, d_maxcount, D, d_X, d_eig, eigcap, d_n, stream))               \
        return -1;
#define MDB_BUCKET_MC(B, NMAX, LPM, CPL, WARPS, MINB, QT)                                                                  \
    if (hb[B] && run_bucket_mc<NMAX, LPM, CPL, WARPS, MINB, QT>(c, P, d_list, ofs[B], hb[B], dsc, esc, nsc, chunk, d_counts,  \
                                                                d_padorder, d_padval, d_maxcount, D, d_X, d_eig, eigcap, d_n, \
                                                                stream))                                                      \
        return -1;
    if (c->opt_desc_legacy == 1) {
        MDB_BUCKET(0, 8, 8, 8, 128, 1)
        MDB_BUCKET(1, 12, 8, 16, 128, 1)
        MDB_BUCKET(2, 16, 8, 16, 128, 1)
        MDB_BUCKET(3, 20, 8, 32, 128, 1)
        MDB_BUCKET(4, 24, 8, 32, 128, 1)
        MDB_BUCKET(5, 28, 8, 32, 128, 1)
        MDB_BUCKET(6, 32, 8, 32, 128, 1)
        MDB_BUCKET(7, 40, 4, 32, 64, 2)
        MDB_BUCKET(8, 40, 4, 32, 64, 2)
        MDB_BUCKET(9, 48, 4, 32, 64, 2)
        MDB_BUCKET(10, 48, 4, 32, 64, 2)
    } else {
        // measured per class on B200 (tools/desc_ab.py, 0.05 atoms/A^3): several columns per lane win up to 24
        // atoms (4 matrices per warp) and tie with the warp-pair kernel at 33..48 (no named barriers); at 25..32
        // the one-column kernel keeps 16 warps per SM against 8 and stays ahead (2.88 vs 3.05 ms)
        MDB_BUCKET_MC(0, 8, 4, 2, 8, 2, 128)
        MDB_BUCKET_MC(1, 12, 4, 3, 8, 2, 128)
        MDB_BUCKET_MC(2, 16, 8, 2, 8, 2, 128)
        MDB_BUCKET_MC(3, 20, 8, 3, 8, 1, 128)
        MDB_BUCKET_MC(4, 24, 8, 3, 8, 1, 128)
        if (c->opt_desc_legacy == 2) {
            MDB_BUCKET_MC(5, 28, 16, 2, 8, 1, 128)
            MDB_BUCKET_MC(6, 32, 16, 2, 8, 1, 128)
        } else {
            MDB_BUCKET(5, 28, 8, 32, 128, 1)
            MDB_BUCKET(6, 32, 8, 32, 128, 1)
        }
        if (c->opt_desc_legacy == 2) {
            MDB_BUCKET_MC(7, 36, 32, 2, 8, 1, 64)
            MDB_BUCKET_MC(8, 40, 32, 2, 8, 1, 64)
        } else {
            if (hb[7] && run_bucket_bord<36, 8, 64>(c, P, d_list, ofs[7], hb[7], dsc, esc, nsc, chunk, d_counts, d_padorder,
                                                    d_padval, d_maxcount, D, d_X, d_eig, eigcap, d_n, stream))
                return -1;
            if (hb[8] && run_bucket_bord<40, 6, 64>(c, P, d_list, ofs[8], hb[8], dsc, esc, nsc, chunk, d_counts, d_padorder,
                                                    d_padval, d_maxcount, D, d_X, d_eig, eigcap, d_n, stream))
                return -1;
        }
        MDB_BUCKET_MC(9, 44, 32, 2, 8, 1, 64)
        MDB_BUCKET_MC(10, 48, 32, 2, 8, 1, 64)
    }
#undef MDB_BUCKET_MC
    MDB_BUCKET(11, 56, 4, 32, 64, 2)
    MDB_BUCKET(12, 64, 4, 32, 64, 2)
    MDB_BUCKET(13, 96, 1, 32, 64, 0)
    MDB_BUCKET(14, 128, 1, 32, 64, 0)
    MDB_BUCKET(15, 160, 1, 32, 64, 0)
#undef MDB_BUCKET
    if (hb[DESC_NB] && run_bucket_glb(c, P, d_list, ofs[DESC_NB], hb[DESC_NB], dsc, esc, nsc, chunk, d_counts, d_padorder, d_padval,
                                      d_maxcount, D, d_X, d_eig, eigcap, d_n, stream))
        return -1;
    return 0;
}
