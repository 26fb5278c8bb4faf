// bonds.cu -- K2: covalent-radius bond detection on the cell list, the Open Babel valence / angle
// cleanup, and ELL -> CSR conversion.
//
// Replaces DetectDump._crd2bond (reference detect.py:249-292), i.e. Open Babel 3.1.x
// OBMol::ConnectTheDots: bond iff 0.16 <= d2 <= (Rcov_i + Rcov_j + 0.45)^2 with
// d = FracToCart(frac - round(frac)), then -- atoms in index order -- while an atom exceeds its
// max valence or has a bond angle < 45 deg: drop its first H-H bond if it is H, else its longest.
//
// Exactness: candidate pairs are screened in float; any pair whose float d2 lies within `margin`
// of either bound is re-evaluated in double with the oracle's operation order
// (__dmul_rn/__dadd_rn, no FMA contraction), so the bond set is bit-identical to the CPU
// restatement.  The cleanup runs entirely in double.
#include <algorithm>

#include "common.cuh"
#include "geom.cuh"

#define BOND_THREADS 128
#define EMPTY_SLOT 0xFFFFFFFFu
#define COS_FLAG 0.7040f  // cos(45 deg) = 0.70711; anything above this is re-checked in double

struct BondParams {
    const float4 *rec;
    const uint32_t *ccell;
    const int *cell_start;
    const FrameGeom *geom;
    const double *pos;
    const uint8_t *types;
    int32_t *ell;
    int32_t *parent;  // optional union-find seed (min(self, smallest neighbour))
    int *viol;        // [F][N] violator lists
    int *nviol;       // [F]
    int *dirty;       // [F][N]
    int *ndirty;      // [F]
    BondStats *stats;
    int N;
    int periodic;
    ElemTables el;
};

// one neighbour cell coordinate along an axis: wrapped index and the shift to add to the
// neighbour's coordinate; returns false when the cell does not exist (open boundary / reduced axis)
__device__ __forceinline__ bool nbr_cell(int c, int d, int n, int reduce, int periodic, float period, int &q, float &sh) {
    if (reduce) {
        q = 0;
        sh = 0.f;
        return d == 0;
    }
    q = c + d;
    sh = 0.f;
    if (q < 0) {
        if (!periodic) return false;
        q += n;
        sh = -period;
    } else if (q >= n) {
        if (!periodic) return false;
        q -= n;
        sh = period;
    }
    return true;
}

// MODE 0: every frame orthorhombic with >= 3 cells per axis -- records hold wrapped Angstrom
//         coordinates, the float distance is 3 adds + 3 fused multiply-adds, and neighbour rows /
//         cells farther from the atom than its largest cutoff are skipped.
// MODE 1: general cell (triclinic and/or single-cell "reduce" axes) -- records hold grid units.
// Phase 1 streams the candidates of the (pruned) 3x3 rows and keeps the few that pass the coarse
// float test in a per-thread pending list; phase 2 decides them (type-specific cutoff, borderline
// cases in double) and phase 3 sorts, flags and writes the row.
// pending candidates per atom (coarse float test passed): a few more than the row holds
#define BOND_PEND(MAXB) ((MAXB) + ((MAXB) >> 1))
template <int MAXB, int MODE>
__global__ void __launch_bounds__(BOND_THREADS) k_bonds(const __grid_constant__ BondParams P) {
    __shared__ float4 s_nb[MAXB][BOND_THREADS];
    constexpr int PEND = BOND_PEND(MAXB);
    __shared__ int s_pend[PEND][BOND_THREADS];
    __shared__ float s_cut2f[MDB_MAXT * MDB_MAXT];
    const int tid = threadIdx.x;
    for (int q = tid; q < MDB_MAXT * MDB_MAXT; q += BOND_THREADS) s_cut2f[q] = P.el.cut2f[q];
    __syncthreads();
    const int N = P.N;
    const int f = blockIdx.y;
    const int p = blockIdx.x * BOND_THREADS + tid;
    if (p >= N) return;
    const FrameGeom &G = P.geom[f];
    const size_t fbase = (size_t)f * N;
    const float4 *__restrict__ rec = P.rec;
    const int *__restrict__ cs = P.cell_start;
    const float4 me = rec[fbase + p];
    const uint32_t cc = P.ccell[fbase + p];
    const int cx = (int)(cc & 1023u), cy = (int)((cc >> 10) & 1023u), cz = (int)(cc >> 20);
    const unsigned mw = (unsigned)__float_as_int(me.w);
    const int i = (int)(mw & MDB_IDX_MASK);
    const int ti = (int)(mw >> MDB_IDX_BITS);
    const int nx = G.n[0], ny = G.n[1], nz = G.n[2];
    const int periodic = P.periodic;
    const float m = P.el.margin;
    const int cbase = (int)G.cell_base;
    const float *cutrow = s_cut2f + ti * MDB_MAXT;
    float cmax = 0.f;
#pragma unroll
    for (int k = 0; k < MDB_MAXT; ++k) cmax = fmaxf(cmax, cutrow[k]);
    cmax += m;
    const float dmin = 0.16f - m;

    // MODE 0: periods / cell widths in Angstrom; MODE 1: periods in grid units + grid->cartesian matrix
    float px, py, pz;
    float h[9];
    int rx = 0, ry = 0, rz = 0;
    // squared distance from the atom to the lower / upper face of its cell per axis (MODE 0 pruning)
    float flo[3] = {0.f, 0.f, 0.f}, fhi[3] = {0.f, 0.f, 0.f};
    if (MODE == 0) {
        px = (float)G.ortho[0];
        py = (float)G.ortho[4];
        pz = (float)G.ortho[8];
        const float w[3] = {G.h[0], G.h[4], G.h[8]};
        const float mc[3] = {me.x, me.y, me.z};
        const int c3[3] = {cx, cy, cz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            // conservative: shrink the face distances by the float error of the coordinates
            float lo = fmaxf(mc[a] - (float)c3[a] * w[a] - 2e-4f * w[a], 0.f);
            float hi = fmaxf((float)(c3[a] + 1) * w[a] - mc[a] - 2e-4f * w[a], 0.f);
            flo[a] = lo * lo;
            fhi[a] = hi * hi;
        }
    } else {
        px = (float)nx;
        py = (float)ny;
        pz = (float)nz;
        rx = G.reduce[0];
        ry = G.reduce[1];
        rz = G.reduce[2];
#pragma unroll
        for (int k = 0; k < 9; ++k) h[k] = G.h[k];
    }

    int np = 0;
    auto scan = [&](int lo, int hi, float ax, float ay, float az) {
#pragma unroll 2
        for (int q = lo; q < hi; ++q) {
            const float4 o = rec[q];
            float dx, dy, dz;
            if (MODE == 0) {
                dx = o.x + ax;
                dy = o.y + ay;
                dz = o.z + az;
            } else {
                float du = o.x + ax, dv = o.y + ay, dw = o.z + az;
                if (rx) du -= rintf(du);
                if (ry) dv -= rintf(dv);
                if (rz) dw -= rintf(dw);
                dx = h[0] * du + h[1] * dv + h[2] * dw;
                dy = h[3] * du + h[4] * dv + h[5] * dw;
                dz = h[6] * du + h[7] * dv + h[8] * dw;
            }
            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (d2 <= cmax && d2 >= dmin) {
                if (np < PEND) s_pend[np][tid] = q;
                ++np;
            }
        }
    };

    // per-axis neighbour cells for d = -1, 0, +1 (wrapped index pre-multiplied to a cell offset, the
    // shift to add to a neighbour's coordinate, and -- MODE 0 -- the squared distance to that slab)
    int yoff[3], zoff[3];
    float ysh[3], zsh[3], yb2[3], zb2[3];
    bool yok[3], zok[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int q;
        yok[d] = nbr_cell(cy, d - 1, ny, ry, periodic, py, q, ysh[d]);
        yoff[d] = q * nx;
        zok[d] = nbr_cell(cz, d - 1, nz, rz, periodic, pz, q, zsh[d]);
        zoff[d] = q * ny * nx;
        yb2[d] = d == 0 ? flo[1] : (d == 2 ? fhi[1] : 0.f);
        zb2[d] = d == 0 ? flo[2] : (d == 2 ? fhi[2] : 0.f);
    }
    // all 9 (dy, dz) rows: bounds of the x-contiguous run are loaded up front (independent loads),
    // then the rows are scanned; fully unrolled so every index above is a register
    int lo[9], hi[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int iz = r / 3, iy = r % 3;
        bool ok = yok[iy] && zok[iz];
        int xa = cx - 1, xb = cx + 1;
        if (MODE == 0) {
            const float base = yb2[iy] + zb2[iz];
            ok = ok && base <= cmax;
            if (base + flo[0] > cmax) xa = cx;
            if (base + fhi[0] > cmax) xb = cx;
        } else if (rx) {
            xa = xb = 0;
        }
        xa = max(xa, 0);
        xb = min(xb, nx - 1);
        const int rowbase = cbase + zoff[iz] + yoff[iy];
        lo[r] = ok ? cs[rowbase + xa] : 0;
        hi[r] = ok ? cs[rowbase + xb + 1] : 0;
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) scan(lo[r], hi[r], -me.x, ysh[r % 3] - me.y, zsh[r / 3] - me.z);
    // atoms in the first / last x cell also see one wrapped cell per row (rare: 2 of nx columns)
    if (periodic && !rx && (cx == 0 || cx == nx - 1)) {
        const int xw = cx == 0 ? nx - 1 : 0;
        const float xws = cx == 0 ? -px : px;
        const float fx = cx == 0 ? flo[0] : fhi[0];
        const bool both = nx == 1;  // cannot happen (rx would be set); kept for clarity
        (void)both;
#pragma unroll 1
        for (int r = 0; r < 9; ++r) {
            const int dzc = r / 3 - 1, dyc = r - (r / 3) * 3 - 1;
            int y, z;
            float sy, sz;
            if (!nbr_cell(cy, dyc, ny, ry, periodic, py, y, sy)) continue;
            if (!nbr_cell(cz, dzc, nz, rz, periodic, pz, z, sz)) continue;
            if (MODE == 0) {
                const float base = (dyc < 0 ? flo[1] : dyc > 0 ? fhi[1] : 0.f) + (dzc < 0 ? flo[2] : dzc > 0 ? fhi[2] : 0.f);
                if (base + fx > cmax) continue;
            }
            const int rowbase = cbase + (z * ny + y) * nx;
            scan(cs[rowbase + xw], cs[rowbase + xw + 1], xws - me.x, sy - me.y, sz - me.z);
        }
    }
    if (np > PEND) {
        atomicAdd(&P.stats->overflow, 1ull);
        atomicExch(&P.stats->err, MDB_ERR_BOND_OVERFLOW);
        np = PEND;
    }

    // phase 2: decide the pending candidates
    const float hpx = 0.5f * px, hpy = 0.5f * py, hpz = 0.5f * pz;
    int cnt = 0;
    for (int k = 0; k < np; ++k) {
        const float4 o = rec[s_pend[k][tid]];
        float dx, dy, dz;
        if (MODE == 0) {
            dx = o.x - me.x;
            dy = o.y - me.y;
            dz = o.z - me.z;
            if (periodic) {  // >= 3 cells per axis: the stencil image is the minimum image
                dx += dx > hpx ? -px : (dx < -hpx ? px : 0.f);
                dy += dy > hpy ? -py : (dy < -hpy ? py : 0.f);
                dz += dz > hpz ? -pz : (dz < -hpz ? pz : 0.f);
            }
        } else {
            float du = o.x - me.x, dv = o.y - me.y, dw = o.z - me.z;
            if (periodic) {
                du += du > hpx ? -px : (du < -hpx ? px : 0.f);
                dv += dv > hpy ? -py : (dv < -hpy ? py : 0.f);
                dw += dw > hpz ? -pz : (dw < -hpz ? pz : 0.f);
            }
            dx = h[0] * du + h[1] * dv + h[2] * dw;
            dy = h[3] * du + h[4] * dv + h[5] * dw;
            dz = h[6] * du + h[7] * dv + h[8] * dw;
        }
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const unsigned ow = (unsigned)__float_as_int(o.w);
        const int tj = (int)(ow >> MDB_IDX_BITS);
        const float c2 = cutrow[tj];
        if (d2 <= c2 + m) {
            const int j = (int)(ow & MDB_IDX_MASK);
            bool ok = true;
            if (!(d2 < c2 - m && d2 > 0.16f + m)) {
                // borderline: decide in double exactly like the oracle
                double v[3];
                delta_exact(P.pos + fbase * 3, G, periodic, i, j, v);
                const double e2 = len2_rn(v);
                ok = !(e2 > P.el.cut2[ti * MDB_MAXT + tj]) && !(e2 < 0.16) && (i != j);
                atomicAdd(&P.stats->exact_tests, 1ull);
            }
            if (ok) {
                if (cnt < MAXB) s_nb[cnt][tid] = make_float4(dx, dy, dz, __int_as_float(j));
                ++cnt;
            }
        }
    }
    if (cnt > MAXB) {
        atomicAdd(&P.stats->overflow, 1ull);
        atomicExch(&P.stats->err, MDB_ERR_BOND_OVERFLOW);
        cnt = MAXB;
    }
    // ascending partner index (the within-cell order of the counting sort is not deterministic)
    for (int a = 1; a < cnt; ++a) {
        float4 v = s_nb[a][tid];
        int key = __float_as_int(v.w);
        int b = a - 1;
        while (b >= 0 && __float_as_int(s_nb[b][tid].w) > key) {
            s_nb[b + 1][tid] = s_nb[b][tid];
            --b;
        }
        s_nb[b + 1][tid] = v;
    }
    // candidate violators of the valence / 45-degree rule (resolved exactly by k_cleanup)
    bool flag = cnt > P.el.maxbonds[ti];
    if (!flag && cnt >= 2) {
        // unit vectors once, then one dot product per pair
        for (int a = 0; a < cnt; ++a) {
            float4 v = s_nb[a][tid];
            const float inv = rsqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
            v.x *= inv;
            v.y *= inv;
            v.z *= inv;
            s_nb[a][tid] = v;
        }
        for (int a = 0; a < cnt && !flag; ++a) {
            const float4 va = s_nb[a][tid];
            for (int b = a + 1; b < cnt; ++b) {
                const float4 vb = s_nb[b][tid];
                if (va.x * vb.x + va.y * vb.y + va.z * vb.z > COS_FLAG) {
                    flag = true;
                    break;
                }
            }
        }
    }
    unsigned row[MAXB];
#pragma unroll
    for (int k = 0; k < MAXB; ++k) row[k] = k < cnt ? (unsigned)__float_as_int(s_nb[k][tid].w) : EMPTY_SLOT;
    if (P.parent) P.parent[fbase + i] = (cnt > 0 && (int)row[0] < i) ? (int)row[0] : i;
    if (flag) {
        row[0] |= MDB_PEND;
        const int k = atomicAdd(&P.nviol[f], 1);
        P.viol[fbase + k] = i;
        atomicAdd(&P.stats->violators, 1ull);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(P.ell + (fbase + i) * MAXB);
#pragma unroll
    for (int k = 0; k < MAXB / 4; ++k) dst[k] = make_uint4(row[4 * k], row[4 * k + 1], row[4 * k + 2], row[4 * k + 3]);
}

// ---------------------------------------------------------------------------------------------
// cleanup: one block per frame.  Violators are resolved in rounds; in a round an atom runs only
// if no still-pending violator with a smaller index is bonded to it, which reproduces the
// sequential index-ordered semantics of Open Babel (a deletion only ever affects the two end
// atoms, and only deletions made by lower-index atoms can precede an atom's own turn).
// ---------------------------------------------------------------------------------------------
template <int MAXB>
__device__ __forceinline__ void load_row(const unsigned *ell, int v, unsigned e[MAXB]) {
    const uint4 *src = reinterpret_cast<const uint4 *>(ell + (size_t)v * MAXB);
#pragma unroll
    for (int k = 0; k < MAXB / 4; ++k) {
        uint4 q = __ldcg(src + k);  // L2 view: sees the atomics of other threads of this kernel
        e[4 * k] = q.x;
        e[4 * k + 1] = q.y;
        e[4 * k + 2] = q.z;
        e[4 * k + 3] = q.w;
    }
}

template <int MAXB>
__device__ void kill_bond(unsigned *ell, int v, int slot_v, int w, int *dirty, int *ndirty) {
    atomicOr(&ell[(size_t)v * MAXB + slot_v], MDB_TOMB);
    unsigned e[MAXB];
    load_row<MAXB>(ell, w, e);
#pragma unroll
    for (int s = 0; s < MAXB; ++s)
        if (e[s] != EMPTY_SLOT && !(e[s] & MDB_TOMB) && (int)(e[s] & MDB_IDX_MASK) == v)
            atomicOr(&ell[(size_t)w * MAXB + s], MDB_TOMB);
    unsigned o = atomicOr(&ell[(size_t)v * MAXB], MDB_DIRTY);
    if (!(o & MDB_DIRTY)) dirty[atomicAdd(ndirty, 1)] = v;
    o = atomicOr(&ell[(size_t)w * MAXB], MDB_DIRTY);
    if (!(o & MDB_DIRTY)) dirty[atomicAdd(ndirty, 1)] = w;
}

template <int MAXB>
__device__ void resolve_atom(const BondParams &P, const FrameGeom &G, unsigned *ell, const double *pos, int v,
                             const unsigned e[MAXB], int *dirty, int *ndirty) {
    int nb[MAXB], slot[MAXB];
    double z[MAXB], vec[MAXB][3], len[MAXB];
    int cnt = 0;
#pragma unroll
    for (int s = 0; s < MAXB; ++s) {
        if (e[s] == EMPTY_SLOT || (e[s] & MDB_TOMB)) continue;
        nb[cnt] = (int)(e[s] & MDB_IDX_MASK);
        slot[cnt] = s;
        ++cnt;
    }
    double pv[3] = {pos[(size_t)v * 3], pos[(size_t)v * 3 + 1], pos[(size_t)v * 3 + 2]};
    for (int k = 0; k < cnt; ++k) {
        const double *q = pos + (size_t)nb[k] * 3;
        double d[3] = {__dsub_rn(q[0], pv[0]), __dsub_rn(q[1], pv[1]), __dsub_rn(q[2], pv[2])};
        z[k] = q[2];
        ob_mic(G, P.periodic, d, vec[k]);
        len[k] = sqrt(len2_rn(vec[k]));
    }
    // creation order of the atom's bonds = ascending partner (z, index)  
    int ord[MAXB];
    for (int a = 0; a < cnt; ++a) ord[a] = a;
    for (int a = 1; a < cnt; ++a) {
        int o0 = ord[a];
        int b = a - 1;
        while (b >= 0 && (z[ord[b]] > z[o0] || (z[ord[b]] == z[o0] && nb[ord[b]] > nb[o0]))) {
            ord[b + 1] = ord[b];
            --b;
        }
        ord[b + 1] = o0;
    }
    const int tv = P.types[v];
    const int maxb = P.el.maxbonds[tv];
    const bool isH = P.el.Z[tv] == 1;
    unsigned alive = (1u << cnt) - 1u;  // bit a <-> ord position a
    for (;;) {
        const int deg = __popc(alive);
        if (deg == 0) break;
        bool bad = deg > maxb;
        if (!bad) {
            double mindeg = 360.0;
            for (int a = 0; a < cnt; ++a) {
                if (!(alive >> a & 1u)) continue;
                const int ka = ord[a];
                for (int b = a + 1; b < cnt; ++b) {
                    if (!(alive >> b & 1u)) continue;
                    const int kb = ord[b];
                    double ang;
                    if (fabs(len[ka]) < 1.0e-3 || fabs(len[kb]) < 1.0e-3)
                        ang = 0.0;
                    else {
                        double dot = __dadd_rn(__dadd_rn(__dmul_rn(vec[ka][0], vec[kb][0]), __dmul_rn(vec[ka][1], vec[kb][1])),
                                               __dmul_rn(vec[ka][2], vec[kb][2]));
                        double dp = __ddiv_rn(dot, __dmul_rn(len[ka], len[kb]));
                        if (dp < -0.999999) dp = -0.9999999;
                        if (dp > 0.9999999) dp = 0.9999999;
                        ang = __dmul_rn(180.0 / 3.14159265358979323846, acos(dp));
                    }
                    if (ang < mindeg) mindeg = ang;
                }
            }
            bad = mindeg < 45.0;
        }
        if (!bad) break;
        int victim = -1;
        if (isH) {
            for (int a = 0; a < cnt; ++a)
                if ((alive >> a & 1u) && P.el.Z[P.types[nb[ord[a]]]] == 1) {
                    victim = a;
                    break;
                }
        }
        if (victim < 0) {
            double mx = -1.0;
            for (int a = 0; a < cnt; ++a)
                if ((alive >> a & 1u) && (victim < 0 || len[ord[a]] > mx)) {
                    victim = a;
                    mx = len[ord[a]];
                }
        }
        alive &= ~(1u << victim);
        kill_bond<MAXB>(ell, v, slot[ord[victim]], nb[ord[victim]], dirty, ndirty);
    }
}

// One round of the same rule with many blocks per frame (large frames: a single block per frame would
// leave the GPU idle).  Atoms that are still blocked stay pending for the next round / for k_cleanup.
template <int MAXB>
__global__ void __launch_bounds__(256) k_cleanup_round(const __grid_constant__ BondParams P) {
    const int f = blockIdx.y;
    const int nv = P.nviol[f];
    const int N = P.N;
    const FrameGeom &G = P.geom[f];
    int *viol = P.viol + (size_t)f * N;
    int *dirty = P.dirty + (size_t)f * N;
    int *ndirty = P.ndirty + f;
    unsigned *ell = reinterpret_cast<unsigned *>(P.ell) + (size_t)f * N * MAXB;
    const double *pos = P.pos + (size_t)f * N * 3;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nv; t += gridDim.x * blockDim.x) {
        const int v = viol[t];
        if (v < 0) continue;
        unsigned e[MAXB];
        load_row<MAXB>(ell, v, e);
        unsigned s0[MAXB];
#pragma unroll
        for (int s = 0; s < MAXB; ++s) {
            const int w = (int)(e[s] & MDB_IDX_MASK);
            const bool look = e[s] != EMPTY_SLOT && !(e[s] & MDB_TOMB) && w < v;
            s0[s] = look ? __ldcg(&ell[(size_t)w * MAXB]) : EMPTY_SLOT;
        }
        bool blocked = false;
#pragma unroll
        for (int s = 0; s < MAXB; ++s) blocked |= (s0[s] != EMPTY_SLOT && (s0[s] & MDB_PEND));
        if (blocked) continue;
        __threadfence();
        load_row<MAXB>(ell, v, e);
        resolve_atom<MAXB>(P, G, ell, pos, v, e, dirty, ndirty);
        __threadfence();
        atomicAnd(&ell[(size_t)v * MAXB], ~MDB_PEND);
        viol[t] = ~v;
    }
}

template <int MAXB>
__global__ void __launch_bounds__(256) k_cleanup(const __grid_constant__ BondParams P) {
    const int f = blockIdx.x;
    const int nv = P.nviol[f];
    if (nv == 0) return;
    const int N = P.N;
    const FrameGeom &G = P.geom[f];
    int *viol = P.viol + (size_t)f * N;
    int *dirty = P.dirty + (size_t)f * N;
    int *ndirty = P.ndirty + f;
    unsigned *ell = reinterpret_cast<unsigned *>(P.ell) + (size_t)f * N * MAXB;
    const double *pos = P.pos + (size_t)f * N * 3;
    for (;;) {
        int pending = 0;
        for (int t = threadIdx.x; t < nv; t += blockDim.x) {
            const int v = viol[t];
            if (v < 0) continue;
            unsigned e[MAXB];
            load_row<MAXB>(ell, v, e);
            // blocked while a still-pending violator with a smaller index is bonded to v
            unsigned s0[MAXB];
#pragma unroll
            for (int s = 0; s < MAXB; ++s) {
                const int w = (int)(e[s] & MDB_IDX_MASK);
                const bool look = e[s] != EMPTY_SLOT && !(e[s] & MDB_TOMB) && w < v;
                s0[s] = look ? __ldcg(&ell[(size_t)w * MAXB]) : EMPTY_SLOT;
            }
            bool blocked = false;
#pragma unroll
            for (int s = 0; s < MAXB; ++s) blocked |= (s0[s] != EMPTY_SLOT && (s0[s] & MDB_PEND));
            if (blocked) {
                ++pending;
                continue;
            }
            // a smaller neighbour may have finished between the two loads above: its deletions are
            // visible once its PEND flag reads clear, so take a fresh copy of the row before resolving
            __threadfence();
            load_row<MAXB>(ell, v, e);
            resolve_atom<MAXB>(P, G, ell, pos, v, e, dirty, ndirty);
            __threadfence();
            atomicAnd(&ell[(size_t)v * MAXB], ~MDB_PEND);
            viol[t] = ~v;
        }
        __threadfence();
        if (__syncthreads_count(pending > 0) == 0) break;
    }
    __syncthreads();
    const int nd = *((volatile int *)ndirty);
    for (int t = threadIdx.x; t < nd; t += blockDim.x) {
        const int r = dirty[t];
        unsigned e[MAXB], keep[MAXB];
        load_row<MAXB>(ell, r, e);
        int c = 0;
#pragma unroll
        for (int s = 0; s < MAXB; ++s) {
            keep[s] = EMPTY_SLOT;
        }
#pragma unroll
        for (int s = 0; s < MAXB; ++s)
            if (e[s] != EMPTY_SLOT && !(e[s] & MDB_TOMB)) {
#pragma unroll
                for (int k = 0; k < MAXB; ++k)
                    if (k == c) keep[k] = e[s] & MDB_IDX_MASK;
                ++c;
            }
        uint4 *dst = reinterpret_cast<uint4 *>(ell + (size_t)r * MAXB);
#pragma unroll
        for (int k = 0; k < MAXB / 4; ++k) dst[k] = make_uint4(keep[4 * k], keep[4 * k + 1], keep[4 * k + 2], keep[4 * k + 3]);
        if (P.parent) P.parent[(size_t)f * N + r] = (c > 0 && (int)keep[0] < r) ? (int)keep[0] : r;
    }
}

size_t bonds_workspace_bytes(int nframes, int natoms) {
    size_t fn = (size_t)nframes * natoms;
    return 2 * align256(fn * sizeof(int)) + 2 * align256((size_t)nframes * sizeof(int)) + 1024;
}

int bonds_run(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos, const uint8_t *d_types, const CellList &cl,
              int periodic, int32_t *d_ell, int32_t *d_parent, cudaStream_t stream) {
    size_t fn = (size_t)nframes * natoms;
    BondParams P;
    P.rec = cl.rec;
    P.ccell = cl.ccell;
    P.cell_start = cl.cell_start;
    P.geom = cl.d_geom;
    P.pos = d_pos;
    P.types = d_types;
    P.ell = d_ell;
    P.parent = d_parent;
    P.viol = (int *)arena_alloc(ctx, fn * sizeof(int));
    P.dirty = (int *)arena_alloc(ctx, fn * sizeof(int));
    // nviol and ndirty share one zero-initialised block
    P.nviol = (int *)arena_alloc(ctx, 2 * align256((size_t)nframes * sizeof(int)));
    if (!P.viol || !P.dirty || !P.nviol) {
        mdb_set_error("bonds_run: arena too small (internal sizing error)");
        return -1;
    }
    P.ndirty = P.nviol + align256((size_t)nframes * sizeof(int)) / sizeof(int);
    P.stats = ctx->d_stats;
    P.N = natoms;
    P.periodic = periodic;
    P.el = ctx->el;
    {
        ProfScope ps(ctx, "memset", stream);
        MDB_CUDA(cudaMemsetAsync(P.nviol, 0, 2 * align256((size_t)nframes * sizeof(int)), stream));
    }
    if (natoms == 0 || nframes == 0) return 0;
    dim3 grid(cdiv(natoms, BOND_THREADS), nframes);
    const int maxb = ctx->opt_max_bonds;
    {
        ProfScope ps(ctx, "k_bonds", stream);
        if (maxb == 8) {
            if (cl.rec_mode == 1)
                k_bonds<8, 0><<<grid, BOND_THREADS, 0, stream>>>(P);
            else
                k_bonds<8, 1><<<grid, BOND_THREADS, 0, stream>>>(P);
        } else {
            if (cl.rec_mode == 1)
                k_bonds<16, 0><<<grid, BOND_THREADS, 0, stream>>>(P);
            else
                k_bonds<16, 1><<<grid, BOND_THREADS, 0, stream>>>(P);
        }
    }
    {
        ProfScope ps(ctx, "k_cleanup", stream);
        // frames too large for one block each: a few wide rounds first, the single-block kernel then
        // finishes the (rare) longer dependency chains and compacts the dirty rows
        const long long per_sm = (long long)nframes * 1;
        if (per_sm < 148 && natoms >= 50000) {
            dim3 rg((unsigned)std::min<long long>(128, std::max<long long>(1, 296 / nframes)), nframes);
            for (int round = 0; round < 3; ++round) {
                if (maxb == 8)
                    k_cleanup_round<8><<<rg, 256, 0, stream>>>(P);
                else
                    k_cleanup_round<16><<<rg, 256, 0, stream>>>(P);
                ctx->launches++;
            }
        }
        if (maxb == 8)
            k_cleanup<8><<<nframes, 256, 0, stream>>>(P);
        else
            k_cleanup<16><<<nframes, 256, 0, stream>>>(P);
    }
    ctx->launches += 2;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// ELL -> CSR.  order = MDB_ORDER_OB sorts each row by partner (z, index): Open Babel creates
// bonds in (z-rank j, z-rank k) lexicographic order, so every atom's adjacency list as built at
// detect.py:282-291 is ascending in the partner's z-rank.
// ---------------------------------------------------------------------------------------------
template <int MAXB>
__global__ void __launch_bounds__(256) k_ell_degree(const int32_t *__restrict__ ell, long long fn, int *__restrict__ deg) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a > fn) return;
    int d = 0;
    if (a < fn) {
        const int4 *row = reinterpret_cast<const int4 *>(ell + a * MAXB);
#pragma unroll
        for (int k = 0; k < MAXB / 4; ++k) {
            int4 q = row[k];
            d += (q.x >= 0) + (q.y >= 0) + (q.z >= 0) + (q.w >= 0);
        }
    }
    deg[a] = d;
}

template <int MAXB>
__global__ void __launch_bounds__(256) k_ell_fill(const int32_t *__restrict__ ell, const int *__restrict__ gofs, int N,
                                                  int order, const double *__restrict__ pos, int32_t *__restrict__ rowptr,
                                                  int32_t *__restrict__ col, int colcap, BondStats *stats) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    const size_t fbase = (size_t)f * N;
    const int base = gofs[fbase];
    const int start = gofs[fbase + i] - base;
    rowptr[(size_t)f * (N + 1) + i] = start;
    if (i == N) {
        if (start > colcap) atomicExch(&stats->err, MDB_ERR_CSR_OVERFLOW);
        return;
    }
    int nb[MAXB];
    int c = 0;
#pragma unroll
    for (int k = 0; k < MAXB; ++k) {
        int e = ell[(fbase + i) * MAXB + k];
        if (e >= 0) nb[c++] = e;
    }
    if (order == MDB_ORDER_OB && c > 1) {
        double z[MAXB];
        for (int k = 0; k < c; ++k) z[k] = pos[(fbase + nb[k]) * 3 + 2];
        for (int a = 1; a < c; ++a) {
            int n0 = nb[a];
            double z0 = z[a];
            int b = a - 1;
            while (b >= 0 && (z[b] > z0 || (z[b] == z0 && nb[b] > n0))) {
                nb[b + 1] = nb[b];
                z[b + 1] = z[b];
                --b;
            }
            nb[b + 1] = n0;
            z[b + 1] = z0;
        }
    }
    if (start + c <= colcap)
        for (int k = 0; k < c; ++k) col[(size_t)f * colcap + start + k] = nb[k];
}

int ell_to_csr_run(mdb_ctx *ctx, int nframes, int natoms, const int32_t *d_ell, int order, const double *d_pos,
                   int32_t *d_rowptr, int32_t *d_col, int colcap, cudaStream_t stream) {
    long long fn = (long long)nframes * natoms;
    size_t need = align256((size_t)(fn + 16) * sizeof(int)) + align256(scan_scratch_bytes(fn + 1)) + 1024;
    if (arena_reserve(ctx, need, stream)) return -1;
    arena_reset(ctx);
    int *gofs = (int *)arena_alloc(ctx, (size_t)(fn + 16) * sizeof(int));
    void *tmp = arena_alloc(ctx, scan_scratch_bytes(fn + 1));
    if (order == MDB_ORDER_OB && !d_pos) {
        mdb_set_error("mdb_ell_to_csr: MDB_ORDER_OB needs positions");
        return -1;
    }
    const int maxb = ctx->opt_max_bonds;
    if (maxb == 8)
        k_ell_degree<8><<<cdiv(fn + 1, 256), 256, 0, stream>>>(d_ell, fn, gofs);
    else
        k_ell_degree<16><<<cdiv(fn + 1, 256), 256, 0, stream>>>(d_ell, fn, gofs);
    ctx->launches++;
    if (device_exclusive_scan(ctx, gofs, fn + 1, tmp, stream)) return -1;
    dim3 grid(cdiv(natoms + 1, 256), nframes);
    if (maxb == 8)
        k_ell_fill<8><<<grid, 256, 0, stream>>>(d_ell, gofs, natoms, order, d_pos, d_rowptr, d_col, colcap, ctx->d_stats);
    else
        k_ell_fill<16><<<grid, 256, 0, stream>>>(d_ell, gofs, natoms, order, d_pos, d_rowptr, d_col, colcap, ctx->d_stats);
    ctx->launches++;
    MDB_CUDA(cudaGetLastError());
    return 0;
}
