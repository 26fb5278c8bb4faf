// bonds.cu -- K2: covalent-radius bond detection on the cell list, the Open Babel valence / angle
// cleanup, and ELL -> CSR conversion.
//
// Replaces DetectDump._crd2bond (reference detect.py:249-292), i.e. Open Babel 3.1.x
// OBMol::ConnectTheDots: bond iff 0.16 <= d2 <= (Rcov_i + Rcov_j + 0.45)^2 with
// d = FracToCart(frac - round(frac)), then -- atoms in index order -- while an atom exceeds its
// max valence or has a bond angle < 45 deg: drop its first H-H bond if it is H, else its longest.
//
// Exactness: candidate pairs are screened in float on grid-unit deltas; any pair whose float d2
// lies within `margin` of either bound is re-evaluated in double with the oracle's operation
// order (__dmul_rn/__dadd_rn, no FMA contraction), so the bond set is bit-identical to the CPU
// restatement.  The cleanup runs entirely in double.
#include "common.cuh"

#define BOND_THREADS 128
#define EMPTY_SLOT 0xFFFFFFFFu
#define COS_FLAG 0.7040f  // cos(45 deg) = 0.70711; anything above this is re-checked in double

struct BondParams {
    const float4 *rec;
    const int *cell_start;
    const FrameGeom *geom;
    const double *pos;
    const uint8_t *types;
    int32_t *ell;
    int *viol;    // [F][N] violator lists
    int *nviol;   // [F]
    int *dirty;   // [F][N]
    int *ndirty;  // [F]
    BondStats *stats;
    int N;
    int periodic;
    ElemTables el;
};

// Open Babel MinimumImageCartesian in the oracle's operation order (oracle.c: ob_mic).
__device__ __forceinline__ void ob_mic(const FrameGeom &G, int periodic, const double d[3], double out[3]) {
    if (!periodic) {
        out[0] = d[0];
        out[1] = d[1];
        out[2] = d[2];
        return;
    }
    double fr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double t = __dadd_rn(__dadd_rn(__dmul_rn(d[0], G.frac[r * 3]), __dmul_rn(d[1], G.frac[r * 3 + 1])),
                             __dmul_rn(d[2], G.frac[r * 3 + 2]));
        fr[r] = __dsub_rn(t, round(t));
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
        out[r] = __dadd_rn(__dadd_rn(__dmul_rn(fr[0], G.ortho[r * 3]), __dmul_rn(fr[1], G.ortho[r * 3 + 1])),
                           __dmul_rn(fr[2], G.ortho[r * 3 + 2]));
}

__device__ __forceinline__ double len2_rn(const double v[3]) {
    return __dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2]));
}

// vector(a) - vector(b), minimum-imaged (oracle.c: sys_delta)
__device__ __forceinline__ void delta_exact(const double *pos, const FrameGeom &G, int periodic, int a, int b,
                                            double out[3]) {
    double d[3] = {__dsub_rn(pos[(size_t)a * 3], pos[(size_t)b * 3]),
                   __dsub_rn(pos[(size_t)a * 3 + 1], pos[(size_t)b * 3 + 1]),
                   __dsub_rn(pos[(size_t)a * 3 + 2], pos[(size_t)b * 3 + 2])};
    ob_mic(G, periodic, d, out);
}

__device__ __forceinline__ int axis_list(int c, int n, int reduce, int periodic, int q[3], float s[3]) {
    if (reduce) {
        q[0] = 0;
        s[0] = 0.f;
        return 1;
    }
    int cnt = 0;
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
        int v = c + d;
        float sh = 0.f;
        if (v < 0) {
            if (!periodic) continue;
            v += n;
            sh = -(float)n;
        } else if (v >= n) {
            if (!periodic) continue;
            v -= n;
            sh = (float)n;
        }
        q[cnt] = v;
        s[cnt] = sh;
        ++cnt;
    }
    return cnt;
}

template <int MAXB, bool ORTHO>
__global__ void __launch_bounds__(BOND_THREADS) k_bonds(const __grid_constant__ BondParams P) {
    __shared__ float4 s_nb[MAXB][BOND_THREADS];
    __shared__ float s_cut2f[MDB_MAXT * MDB_MAXT];
    const int tid = threadIdx.x;
    if (tid < MDB_MAXT * MDB_MAXT) s_cut2f[tid] = P.el.cut2f[tid];
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
    const unsigned mw = (unsigned)__float_as_int(me.w);
    const int i = (int)(mw & MDB_IDX_MASK);
    const int ti = (int)(mw >> MDB_IDX_BITS);
    const int nx = G.n[0], ny = G.n[1];
    const int rx = G.reduce[0], ry = G.reduce[1], rz = G.reduce[2];
    float h[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) h[k] = G.h[k];
    const float m = P.el.margin;
    const long long cbase = G.cell_base;

    int yq[3], zq[3];
    float ys[3], zs[3];
    const int nyq = axis_list((int)me.y, ny, ry, P.periodic, yq, ys);
    const int nzq = axis_list((int)me.z, G.n[2], rz, P.periodic, zq, zs);
    // x: contiguous run of cells plus at most one wrapped cell on either side
    int xa[3], xb[3];
    float xs[3];
    int nxs = 0;
    {
        const int cx = (int)me.x;
        if (rx) {
            xa[0] = 0;
            xb[0] = 0;
            xs[0] = 0.f;
            nxs = 1;
        } else {
            int lo = cx - 1, hi = cx + 1;
            xa[0] = lo < 0 ? 0 : lo;
            xb[0] = hi >= nx ? nx - 1 : hi;
            xs[0] = 0.f;
            nxs = 1;
            if (P.periodic && lo < 0) {
                xa[nxs] = nx - 1;
                xb[nxs] = nx - 1;
                xs[nxs] = -(float)nx;
                ++nxs;
            }
            if (P.periodic && hi >= nx) {
                xa[nxs] = 0;
                xb[nxs] = 0;
                xs[nxs] = (float)nx;
                ++nxs;
            }
        }
    }

    int cnt = 0;
    const float *cutrow = s_cut2f + ti * MDB_MAXT;
#pragma unroll
    for (int iz = 0; iz < 3; ++iz) {
        if (iz >= nzq) break;
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            if (iy >= nyq) break;
            const long long rowbase = cbase + ((long long)zq[iz] * ny + yq[iy]) * nx;
            const float sy = ys[iy] - me.y, sz = zs[iz] - me.z;
#pragma unroll
            for (int is = 0; is < 3; ++is) {
                if (is >= nxs) break;
                const int lo = cs[rowbase + xa[is]];
                const int hi = cs[rowbase + xb[is] + 1];
                const float sx = xs[is] - me.x;
                for (int q = lo; q < hi; ++q) {
                    const float4 o = rec[q];
                    float du = o.x + sx, dv = o.y + sy, dw = o.z + sz;
                    if (rx) du -= rintf(du);
                    if (ry) dv -= rintf(dv);
                    if (rz) dw -= rintf(dw);
                    float dx, dy, dz;
                    if (ORTHO) {
                        dx = h[0] * du;
                        dy = h[4] * dv;
                        dz = h[8] * dw;
                    } else {
                        dx = h[0] * du + h[1] * dv + h[2] * dw;
                        dy = h[3] * du + h[4] * dv + h[5] * dw;
                        dz = h[6] * du + h[7] * dv + h[8] * dw;
                    }
                    const float d2 = dx * dx + dy * dy + dz * dz;
                    const unsigned ow = (unsigned)__float_as_int(o.w);
                    const float c2 = cutrow[ow >> MDB_IDX_BITS];
                    if (d2 <= c2 + m && d2 >= 0.16f - m) {
                        const int j = (int)(ow & MDB_IDX_MASK);
                        bool ok = true;
                        if (!(d2 < c2 - m && d2 > 0.16f + m)) {
                            // borderline: decide in double exactly like the oracle
                            double v[3];
                            delta_exact(P.pos + fbase * 3, G, P.periodic, i, j, v);
                            const double e2 = len2_rn(v);
                            const double cut2 = P.el.cut2[ti * MDB_MAXT + (int)(ow >> MDB_IDX_BITS)];
                            ok = !(e2 > cut2) && !(e2 < 0.16) && (i != j);
                            atomicAdd(&P.stats->exact_tests, 1ull);
                        }
                        if (ok) {
                            if (cnt < MAXB) s_nb[cnt][tid] = make_float4(dx, dy, dz, __int_as_float(j));
                            ++cnt;
                        }
                    }
                }
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
        for (int a = 0; a < cnt && !flag; ++a) {
            const float4 va = s_nb[a][tid];
            const float la = va.x * va.x + va.y * va.y + va.z * va.z;
            for (int b = a + 1; b < cnt; ++b) {
                const float4 vb = s_nb[b][tid];
                const float lb = vb.x * vb.x + vb.y * vb.y + vb.z * vb.z;
                const float dp = (va.x * vb.x + va.y * vb.y + va.z * vb.z) * rsqrtf(la * lb);
                if (dp > COS_FLAG) {
                    flag = true;
                    break;
                }
            }
        }
    }
    unsigned row[MAXB];
#pragma unroll
    for (int k = 0; k < MAXB; ++k) row[k] = k < cnt ? (unsigned)__float_as_int(s_nb[k][tid].w) : EMPTY_SLOT;
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
__device__ __forceinline__ unsigned ld_vol(const unsigned *p) { return *((const volatile unsigned *)p); }

template <int MAXB>
__device__ void kill_bond(unsigned *ell, int v, int slot_v, int w, int *dirty, int *ndirty) {
    atomicOr(&ell[(size_t)v * MAXB + slot_v], MDB_TOMB);
    for (int s = 0; s < MAXB; ++s) {
        unsigned e = ld_vol(&ell[(size_t)w * MAXB + s]);
        if (e == EMPTY_SLOT) break;
        if (!(e & MDB_TOMB) && (int)(e & MDB_IDX_MASK) == v) {
            atomicOr(&ell[(size_t)w * MAXB + s], MDB_TOMB);
            break;
        }
    }
    unsigned o = atomicOr(&ell[(size_t)v * MAXB], MDB_DIRTY);
    if (!(o & MDB_DIRTY)) dirty[atomicAdd(ndirty, 1)] = v;
    o = atomicOr(&ell[(size_t)w * MAXB], MDB_DIRTY);
    if (!(o & MDB_DIRTY)) dirty[atomicAdd(ndirty, 1)] = w;
}

template <int MAXB>
__device__ void resolve_atom(const BondParams &P, const FrameGeom &G, unsigned *ell, const double *pos, int v,
                             int *dirty, int *ndirty) {
    int nb[MAXB], slot[MAXB];
    double z[MAXB], vec[MAXB][3], len[MAXB];
    int cnt = 0;
    for (int s = 0; s < MAXB; ++s) {
        unsigned e = ld_vol(&ell[(size_t)v * MAXB + s]);
        if (e == EMPTY_SLOT) break;
        if (e & MDB_TOMB) continue;
        nb[cnt] = (int)(e & MDB_IDX_MASK);
        slot[cnt] = s;
        z[cnt] = pos[(size_t)nb[cnt] * 3 + 2];
        ++cnt;
    }
    // creation order of the atom's bonds = ascending partner (z, index)  
    for (int a = 1; a < cnt; ++a) {
        int n0 = nb[a], s0 = slot[a];
        double z0 = z[a];
        int b = a - 1;
        while (b >= 0 && (z[b] > z0 || (z[b] == z0 && nb[b] > n0))) {
            nb[b + 1] = nb[b];
            slot[b + 1] = slot[b];
            z[b + 1] = z[b];
            --b;
        }
        nb[b + 1] = n0;
        slot[b + 1] = s0;
        z[b + 1] = z0;
    }
    for (int k = 0; k < cnt; ++k) {
        delta_exact(pos, G, P.periodic, nb[k], v, vec[k]);
        len[k] = sqrt(len2_rn(vec[k]));
    }
    const int tv = P.types[v];
    const int maxb = P.el.maxbonds[tv];
    const bool isH = P.el.Z[tv] == 1;
    unsigned alive = cnt >= 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u);
    for (;;) {
        const int deg = __popc(alive);
        if (deg == 0) break;
        bool bad = deg > maxb;
        if (!bad) {
            double mindeg = 360.0;
            for (int a = 0; a < cnt; ++a) {
                if (!(alive >> a & 1u)) continue;
                for (int b = a + 1; b < cnt; ++b) {
                    if (!(alive >> b & 1u)) continue;
                    double ang;
                    if (fabs(len[a]) < 1.0e-3 || fabs(len[b]) < 1.0e-3)
                        ang = 0.0;
                    else {
                        double dot = __dadd_rn(__dadd_rn(__dmul_rn(vec[a][0], vec[b][0]), __dmul_rn(vec[a][1], vec[b][1])),
                                               __dmul_rn(vec[a][2], vec[b][2]));
                        double dp = __ddiv_rn(dot, __dmul_rn(len[a], len[b]));
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
                if ((alive >> a & 1u) && P.el.Z[P.types[nb[a]]] == 1) {
                    victim = a;
                    break;
                }
        }
        if (victim < 0) {
            double mx = -1.0;
            for (int a = 0; a < cnt; ++a)
                if ((alive >> a & 1u) && (victim < 0 || len[a] > mx)) {
                    victim = a;
                    mx = len[a];
                }
        }
        alive &= ~(1u << victim);
        kill_bond<MAXB>(ell, v, slot[victim], nb[victim], dirty, ndirty);
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
            bool blocked = false;
            for (int s = 0; s < MAXB; ++s) {
                unsigned e = ld_vol(&ell[(size_t)v * MAXB + s]);
                if (e == EMPTY_SLOT) break;
                if (e & MDB_TOMB) continue;
                int w = (int)(e & MDB_IDX_MASK);
                if (w < v) {
                    unsigned s0 = ld_vol(&ell[(size_t)w * MAXB]);
                    if (s0 != EMPTY_SLOT && (s0 & MDB_PEND)) {
                        blocked = true;
                        break;
                    }
                }
            }
            if (blocked) {
                ++pending;
                continue;
            }
            resolve_atom<MAXB>(P, G, ell, pos, v, dirty, ndirty);
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
        unsigned keep[MAXB];
        int c = 0;
        for (int s = 0; s < MAXB; ++s) {
            unsigned e = ld_vol(&ell[(size_t)r * MAXB + s]);
            if (e == EMPTY_SLOT) break;
            if (e & MDB_TOMB) continue;
            keep[c++] = e & MDB_IDX_MASK;
        }
        for (int s = 0; s < MAXB; ++s) ell[(size_t)r * MAXB + s] = s < c ? keep[s] : EMPTY_SLOT;
    }
}

size_t bonds_workspace_bytes(int nframes, int natoms) {
    size_t fn = (size_t)nframes * natoms;
    return 2 * align256(fn * sizeof(int)) + 2 * align256((size_t)nframes * sizeof(int)) + 1024;
}

int bonds_run(mdb_ctx *ctx, int nframes, int natoms, const double *d_pos, const uint8_t *d_types, const CellList &cl,
              int periodic, int32_t *d_ell, cudaStream_t stream) {
    size_t fn = (size_t)nframes * natoms;
    BondParams P;
    P.rec = cl.rec;
    P.cell_start = cl.cell_start;
    P.geom = cl.d_geom;
    P.pos = d_pos;
    P.types = d_types;
    P.ell = d_ell;
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
    MDB_CUDA(cudaMemsetAsync(P.nviol, 0, 2 * align256((size_t)nframes * sizeof(int)), stream));
    if (natoms == 0 || nframes == 0) return 0;
    dim3 grid(cdiv(natoms, BOND_THREADS), nframes);
    const int maxb = ctx->opt_max_bonds;
    // the float screen takes the diagonal shortcut only when every frame of the batch is orthorhombic
    if (maxb == 8) {
        if (cl.all_ortho)
            k_bonds<8, true><<<grid, BOND_THREADS, 0, stream>>>(P);
        else
            k_bonds<8, false><<<grid, BOND_THREADS, 0, stream>>>(P);
        k_cleanup<8><<<nframes, 256, 0, stream>>>(P);
    } else {
        if (cl.all_ortho)
            k_bonds<16, true><<<grid, BOND_THREADS, 0, stream>>>(P);
        else
            k_bonds<16, false><<<grid, BOND_THREADS, 0, stream>>>(P);
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
