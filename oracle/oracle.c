/* oracle.c -- CPU restatement of MDDatasetBuilder's per-frame analysis path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / the timed CPU
 * baseline.  The product path (mddatasetbuilder_b200/) never links or calls it.
 *
 * Parity status of this file:
 *   - orc_dps              restates /root/reference/mddatasetbuilder/dps.pyx:17-46
 *                          (+ c_stack.cpp:18-33).  PINNED: checked against the
 *                          reference's own compiled dps (oracle/_ref) in tests.
 *   - orc_connect_the_dots restates Open Babel 3.1.x OBMol::ConnectTheDots
 *                          (src/mol.cpp), called at reference detect.py:255-292.
 *                          Open Babel (openbabel-wheel >= 3.1.0.0, pyproject.toml:38)
 *                          is NOT in /root/reference nor installable here, so this
 *                          is a restatement of its published algorithm.  PINNED only
 *                          by the reference's tests/test_bond.py:9-34 (O2 cases);
 *                          otherwise PARITY UNPINNED.
 *   - orc_ase_*            restate ASE geometry.find_mic / general_find_mic /
 *                          naive_find_mic (ase/geometry/geometry.py), called at
 *                          reference datasetbuilder.py:312-315,342,554-557.  ASE is
 *                          not available here: PARITY UNPINNED (no reference test
 *                          asserts a number for it).
 *   - orc_bond_keys        restates detect.py:112-116 (bond-file mode levels).
 *
 * All arithmetic is IEEE double with no FMA contraction (build with
 * -ffp-contract=off): the CUDA kernels reproduce the same operation order with
 * __dmul_rn/__dadd_rn so that integer results (bond lists, membership) are
 * bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXZ 18
/* Open Babel 3.1.x element table (src/elementtable.h): covalent radius (Cordero
 * 2008) and MaxBonds per atomic number.  Index = Z.  From memory of the
 * published table; C/H/N/O values are also listed in SURVEY.md App. A.2. */
static const double ORC_RCOV[ORC_MAXZ + 1] = {
    0.0,  0.31, 0.28, 1.28, 0.96, 0.84, 0.76, 0.71, 0.66, 0.57,
    0.58, 1.66, 1.41, 1.21, 1.11, 1.07, 1.05, 1.02, 1.06};
static const int ORC_MAXBONDS[ORC_MAXZ + 1] = {0, 1, 0, 1, 2, 4, 4, 4, 2, 1,
                                               0, 1, 2, 6, 6, 6, 6, 1, 0};

double orc_rcov(int Z) { return (Z >= 0 && Z <= ORC_MAXZ) ? ORC_RCOV[Z] : -1.0; }
int orc_maxbonds(int Z) { return (Z >= 0 && Z <= ORC_MAXZ) ? ORC_MAXBONDS[Z] : -1; }
void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------ */
/* Open Babel unit-cell arithmetic (OBUnitCell, src/generic.cpp).      */
/* cell[v*3+c] = component c of lattice vector v (rows = vectors, as   */
/* passed at detect.py:268-272).  _mOrtho has the vectors as COLUMNS.  */
/* ------------------------------------------------------------------ */
typedef struct {
    double o[3][3];   /* _mOrtho: frac -> cart */
    double f[3][3];   /* _mOrtho.inverse(): cart -> frac (adjugate / det) */
} ob_cell;

/* matrix3x3::determinant + matrix3x3::inverse (src/math/matrix3x3.cpp). */
void orc_ob_cell_setup(const double *cell, double *ortho9, double *frac9) {
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

static void ob_cell_init(ob_cell *uc, const double *cell) {
    orc_ob_cell_setup(cell, &uc->o[0][0], &uc->f[0][0]);
}

/* OBUnitCell::MinimumImageCartesian = FracToCart(frac - round(frac)). */
static inline void ob_mic(const ob_cell *uc, const double d[3], double out[3]) {
    double fr[3];
    for (int r = 0; r < 3; ++r)
        fr[r] = (d[0] * uc->f[r][0] + d[1] * uc->f[r][1]) + d[2] * uc->f[r][2];
    for (int r = 0; r < 3; ++r) fr[r] = fr[r] - round(fr[r]);
    for (int r = 0; r < 3; ++r)
        out[r] = (fr[0] * uc->o[r][0] + fr[1] * uc->o[r][1]) + fr[2] * uc->o[r][2];
}

void orc_ob_mic(const double *cell, const double *d, double *out) {
    ob_cell uc;
    ob_cell_init(&uc, cell);
    ob_mic(&uc, d, out);
}

/* ------------------------------------------------------------------ */
/* ConnectTheDots                                                      */
/* ------------------------------------------------------------------ */
typedef struct {
    int n;
    const int *Z;
    const double *pos;
    int periodic;
    ob_cell uc;
} ctd_sys;

static inline void sys_delta(const ctd_sys *s, int a, int b, double out[3]) {
    /* vector(a) - vector(b), minimum-imaged when periodic */
    double d[3] = {s->pos[a * 3] - s->pos[b * 3], s->pos[a * 3 + 1] - s->pos[b * 3 + 1],
                   s->pos[a * 3 + 2] - s->pos[b * 3 + 2]};
    if (s->periodic)
        ob_mic(&s->uc, d, out);
    else {
        out[0] = d[0];
        out[1] = d[1];
        out[2] = d[2];
    }
}

static inline double len2(const double v[3]) { return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]; }

typedef struct {
    int a, b;       /* begin (lower z-rank), end */
    int ra, rb;     /* z-ranks */
    double d2;
    int alive;
} ctd_bond;

typedef struct {
    double z;
    int idx;
} zrec;

static int zrec_cmp(const void *pa, const void *pb) {
    const zrec *a = (const zrec *)pa, *b = (const zrec *)pb;
    if (a->z < b->z) return -1;
    if (a->z > b->z) return 1;
    /* std::sort leaves ties unspecified in Open Babel; we define them by index */
    return (a->idx > b->idx) - (a->idx < b->idx);
}

static int bond_cmp(const void *pa, const void *pb) {
    const ctd_bond *a = (const ctd_bond *)pa, *b = (const ctd_bond *)pb;
    if (a->ra != b->ra) return (a->ra > b->ra) - (a->ra < b->ra);
    return (a->rb > b->rb) - (a->rb < b->rb);
}

typedef struct {
    ctd_bond *v;
    int n, cap;
} bondvec;

static void bv_push(bondvec *bv, ctd_bond b) {
    if (bv->n == bv->cap) {
        bv->cap = bv->cap ? bv->cap * 2 : 1024;
        bv->v = (ctd_bond *)realloc(bv->v, (size_t)bv->cap * sizeof(ctd_bond));
    }
    bv->v[bv->n++] = b;
}

/* test one z-ranked pair (rj < rk) exactly as the OB pair loop does */
static inline void ctd_try_pair(const ctd_sys *s, const zrec *zs, const double *rad, int rj, int rk,
                                bondvec *bv) {
    int a = zs[rj].idx, b = zs[rk].idx;
    double cut = rad[rj] + rad[rk] + 0.45;
    cut = cut * cut;
    double v[3];
    sys_delta(s, a, b, v);
    double d2 = len2(v);
    if (d2 > cut) return;
    if (d2 < 0.16) return;
    ctd_bond bd = {a, b, rj, rk, d2, 1};
    bv_push(bv, bd);
}

/* smallest bond angle at `center` over its alive bonds (OBAtom::SmallestBondAngle) */
static double smallest_angle(const ctd_sys *s, const ctd_bond *bonds, const int *lst, int cnt, int center) {
    double mindeg = 360.0;
    for (int i = 0; i < cnt; ++i) {
        if (!bonds[lst[i]].alive) continue;
        int b = bonds[lst[i]].a == center ? bonds[lst[i]].b : bonds[lst[i]].a;
        double v1[3];
        sys_delta(s, b, center, v1);
        for (int k = i + 1; k < cnt; ++k) {
            if (!bonds[lst[k]].alive) continue;
            int c = bonds[lst[k]].a == center ? bonds[lst[k]].b : bonds[lst[k]].a;
            double v2[3];
            sys_delta(s, c, center, v2);
            double l1 = sqrt(len2(v1)), l2 = sqrt(len2(v2));
            double deg;
            if (fabs(l1) < 1.0e-3 || fabs(l2) < 1.0e-3)
                deg = 0.0;
            else {
                double dp = ((v1[0] * v2[0] + v1[1] * v2[1]) + v1[2] * v2[2]) / (l1 * l2);
                if (dp < -0.999999) dp = -0.9999999;
                if (dp > 0.9999999) dp = 0.9999999;
                deg = (180.0 / M_PI) * acos(dp);
            }
            if (deg < mindeg) mindeg = deg;
        }
    }
    return mindeg;
}

/* Returns number of bonds; *bonds_out = malloc'd int[2*nb] (begin,end) in
 * OBMolBondIter order (creation order with deleted bonds removed).
 * mode 0: the O(N^2) pair loop of Open Babel (faithful cost model).
 * mode 1: cell-list candidate generation, identical result. */
int orc_connect_the_dots(int n, const int *Z, const double *pos, const double *cell, int periodic,
                         int mode, int **bonds_out, int *ndeleted_out) {
    ctd_sys s;
    s.n = n;
    s.Z = Z;
    s.pos = pos;
    s.periodic = periodic;
    if (periodic) ob_cell_init(&s.uc, cell);
    zrec *zs = (zrec *)malloc((size_t)(n > 0 ? n : 1) * sizeof(zrec));
    double *rad = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    int *rank = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    for (int i = 0; i < n; ++i) {
        zs[i].z = pos[i * 3 + 2];
        zs[i].idx = i;
    }
    qsort(zs, (size_t)n, sizeof(zrec), zrec_cmp);
    double maxrad = 0.0;
    for (int j = 0; j < n; ++j) {
        rad[j] = orc_rcov(Z[zs[j].idx]);
        if (rad[j] > maxrad) maxrad = rad[j];
        rank[zs[j].idx] = j;
    }
    bondvec bv = {0, 0, 0};

    if (mode == 0) {
        for (int j = 0; j < n; ++j) {
            double maxcut = rad[j] + maxrad + 0.45;
            maxcut = maxcut * maxcut;
            for (int k = j + 1; k < n; ++k) {
                if (!periodic) {
                    double dz = pos[zs[j].idx * 3 + 2] - pos[zs[k].idx * 3 + 2];
                    if (dz * dz > maxcut) break;
                }
                ctd_try_pair(&s, zs, rad, j, k, &bv);
            }
        }
    } else {
        /* candidate generation on a grid of (wrapped) fractional coordinates */
        double rc = 2.0 * maxrad + 0.45;
        double fo[9], fi[9];
        double lo[3] = {0, 0, 0}, span[3] = {1, 1, 1};
        int nb[3] = {1, 1, 1};
        double *g = (double *)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(double));
        if (periodic) {
            orc_ob_cell_setup(cell, fo, fi);
            /* perpendicular heights h_a = 1/|row a of frac matrix| */
            for (int a = 0; a < 3; ++a) {
                double nr = sqrt(fi[a * 3] * fi[a * 3] + fi[a * 3 + 1] * fi[a * 3 + 1] + fi[a * 3 + 2] * fi[a * 3 + 2]);
                double h = 1.0 / nr;
                int m = (int)floor(h / (rc * 1.0001));
                nb[a] = m < 1 ? 1 : (m > 512 ? 512 : m);
            }
            for (int i = 0; i < n; ++i)
                for (int a = 0; a < 3; ++a) {
                    double f = pos[i * 3] * fi[a * 3] + pos[i * 3 + 1] * fi[a * 3 + 1] + pos[i * 3 + 2] * fi[a * 3 + 2];
                    f -= floor(f);
                    if (f >= 1.0) f = 0.0;
                    g[i * 3 + a] = f;
                }
        } else {
            for (int a = 0; a < 3; ++a) {
                double mn = 1e300, mx = -1e300;
                for (int i = 0; i < n; ++i) {
                    if (pos[i * 3 + a] < mn) mn = pos[i * 3 + a];
                    if (pos[i * 3 + a] > mx) mx = pos[i * 3 + a];
                }
                if (n == 0) mn = mx = 0;
                lo[a] = mn;
                span[a] = (mx - mn) > 0 ? (mx - mn) * 1.000001 + 1e-9 : 1.0;
                int m = (int)floor(span[a] / (rc * 1.0001));
                nb[a] = m < 1 ? 1 : (m > 512 ? 512 : m);
            }
            for (int i = 0; i < n; ++i)
                for (int a = 0; a < 3; ++a) g[i * 3 + a] = (pos[i * 3 + a] - lo[a]) / span[a];
        }
        size_t ncell = (size_t)nb[0] * nb[1] * nb[2];
        int *cstart = (int *)calloc(ncell + 1, sizeof(int));
        int *cid = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
        int *order = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
        for (int i = 0; i < n; ++i) {
            int c[3];
            for (int a = 0; a < 3; ++a) {
                c[a] = (int)(g[i * 3 + a] * nb[a]);
                if (c[a] >= nb[a]) c[a] = nb[a] - 1;
                if (c[a] < 0) c[a] = 0;
            }
            cid[i] = (c[2] * nb[1] + c[1]) * nb[0] + c[0];
            cstart[cid[i] + 1]++;
        }
        for (size_t c = 0; c < ncell; ++c) cstart[c + 1] += cstart[c];
        int *fill = (int *)calloc(ncell, sizeof(int));
        for (int i = 0; i < n; ++i) order[cstart[cid[i]] + fill[cid[i]]++] = i;
        free(fill);
        for (int i = 0; i < n; ++i) {
            int ci = cid[i];
            int c0 = ci % nb[0], c1 = (ci / nb[0]) % nb[1], c2 = ci / (nb[0] * nb[1]);
            int cc[3] = {c0, c1, c2};
            /* distinct neighbour bins per axis */
            int lst[3][3], cnt[3];
            for (int a = 0; a < 3; ++a) {
                cnt[a] = 0;
                for (int d = -1; d <= 1; ++d) {
                    int q = cc[a] + d;
                    if (periodic) {
                        q = ((q % nb[a]) + nb[a]) % nb[a];
                    } else if (q < 0 || q >= nb[a])
                        continue;
                    int dup = 0;
                    for (int t = 0; t < cnt[a]; ++t) dup |= (lst[a][t] == q);
                    if (!dup) lst[a][cnt[a]++] = q;
                }
            }
            for (int iz = 0; iz < cnt[2]; ++iz)
                for (int iy = 0; iy < cnt[1]; ++iy)
                    for (int ix = 0; ix < cnt[0]; ++ix) {
                        int c = (lst[2][iz] * nb[1] + lst[1][iy]) * nb[0] + lst[0][ix];
                        for (int p = cstart[c]; p < cstart[c + 1]; ++p) {
                            int j = order[p];
                            if (rank[i] < rank[j]) ctd_try_pair(&s, zs, rad, rank[i], rank[j], &bv);
                        }
                    }
        }
        free(g);
        free(cstart);
        free(cid);
        free(order);
        qsort(bv.v, (size_t)bv.n, sizeof(ctd_bond), bond_cmp);
    }

    /* per-atom bond lists in creation order */
    int *ptr = (int *)calloc((size_t)n + 1, sizeof(int));
    for (int b = 0; b < bv.n; ++b) {
        ptr[bv.v[b].a + 1]++;
        ptr[bv.v[b].b + 1]++;
    }
    for (int i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
    int *lst = (int *)malloc((size_t)(2 * bv.n > 0 ? 2 * bv.n : 1) * sizeof(int));
    int *fl = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int b = 0; b < bv.n; ++b) {
        lst[ptr[bv.v[b].a] + fl[bv.v[b].a]++] = b;
        lst[ptr[bv.v[b].b] + fl[bv.v[b].b]++] = b;
    }
    free(fl);

    /* cleanup pass: atoms in index order (mol.cpp, "Cleanup -- delete long bonds") */
    int ndel = 0;
    for (int i = 0; i < n; ++i) {
        const int *L = lst + ptr[i];
        int cnt = ptr[i + 1] - ptr[i];
        int maxb = orc_maxbonds(Z[i]);
        for (;;) {
            int deg = 0, first = -1;
            for (int t = 0; t < cnt; ++t)
                if (bv.v[L[t]].alive) {
                    if (first < 0) first = t;
                    ++deg;
                }
            if (!(deg > maxb || smallest_angle(&s, bv.v, L, cnt, i) < 45.0)) break;
            if (first < 0) break; /* no bonds left */
            if (Z[i] == 1) {
                int hit = -1;
                for (int t = first; t < cnt; ++t) {
                    if (!bv.v[L[t]].alive) continue;
                    int o = bv.v[L[t]].a == i ? bv.v[L[t]].b : bv.v[L[t]].a;
                    if (Z[o] == 1) {
                        hit = t;
                        break;
                    }
                }
                if (hit >= 0) {
                    bv.v[L[hit]].alive = 0;
                    ++ndel;
                    continue;
                }
            }
            int mx = first;
            double mxlen = sqrt(bv.v[L[first]].d2);
            for (int t = first + 1; t < cnt; ++t) {
                if (!bv.v[L[t]].alive) continue;
                double l = sqrt(bv.v[L[t]].d2);
                if (l > mxlen) {
                    mx = t;
                    mxlen = l;
                }
            }
            bv.v[L[mx]].alive = 0;
            ++ndel;
        }
    }

    int nb_alive = 0;
    for (int b = 0; b < bv.n; ++b) nb_alive += bv.v[b].alive;
    int *out = (int *)malloc((size_t)(nb_alive > 0 ? nb_alive : 1) * 2 * sizeof(int));
    int w = 0;
    for (int b = 0; b < bv.n; ++b)
        if (bv.v[b].alive) {
            out[w * 2] = bv.v[b].a;
            out[w * 2 + 1] = bv.v[b].b;
            ++w;
        }
    *bonds_out = out;
    if (ndeleted_out) *ndeleted_out = ndel;
    free(zs);
    free(rad);
    free(rank);
    free(bv.v);
    free(ptr);
    free(lst);
    return nb_alive;
}

/* Bounded sample of the O(N^2) pair loop for the CPU baseline: runs the periodic pair loop of
 * mode 0 for z-ranks j = offset, offset+stride, ... only (no cleanup) and returns the number of
 * bonds found.  One full frame costs `stride` times the sampled time; the rows are strided so
 * the triangular loop (k > j) is sampled evenly. */
long orc_ctd_sample(int n, const int *Z, const double *pos, const double *cell, int periodic, int stride,
                    int offset) {
    ctd_sys s;
    s.n = n;
    s.Z = Z;
    s.pos = pos;
    s.periodic = periodic;
    if (periodic) ob_cell_init(&s.uc, cell);
    zrec *zs = (zrec *)malloc((size_t)(n > 0 ? n : 1) * sizeof(zrec));
    double *rad = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    for (int i = 0; i < n; ++i) {
        zs[i].z = pos[i * 3 + 2];
        zs[i].idx = i;
    }
    qsort(zs, (size_t)n, sizeof(zrec), zrec_cmp);
    double maxrad = 0.0;
    for (int j = 0; j < n; ++j) {
        rad[j] = orc_rcov(Z[zs[j].idx]);
        if (rad[j] > maxrad) maxrad = rad[j];
    }
    bondvec bv = {0, 0, 0};
    for (int j = offset; j < n; j += stride) {
        double maxcut = rad[j] + maxrad + 0.45;
        maxcut = maxcut * maxcut;
        for (int k = j + 1; k < n; ++k) {
            if (!periodic) {
                double dz = pos[zs[j].idx * 3 + 2] - pos[zs[k].idx * 3 + 2];
                if (dz * dz > maxcut) break;
            }
            ctd_try_pair(&s, zs, rad, j, k, &bv);
        }
    }
    long nb = bv.n;
    free(bv.v);
    free(zs);
    free(rad);
    return nb;
}

/* adjacency in the order detect.py:282-291 appends it: for each bond in
 * iteration order, bond[s1].append(s2); bond[s2].append(s1).  rowptr[n+1], col[2*nb]. */
void orc_bonds_to_csr(int n, int nb, const int *bonds, int *rowptr, int *col) {
    memset(rowptr, 0, (size_t)(n + 1) * sizeof(int));
    for (int b = 0; b < nb; ++b) {
        rowptr[bonds[2 * b] + 1]++;
        rowptr[bonds[2 * b + 1] + 1]++;
    }
    for (int i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
    int *fl = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int b = 0; b < nb; ++b) {
        int a = bonds[2 * b], c = bonds[2 * b + 1];
        col[rowptr[a] + fl[a]++] = c;
        col[rowptr[c] + fl[c]++] = a;
    }
    free(fl);
}

/* ------------------------------------------------------------------ */
/* dps: iterative LIFO depth-first search (dps.pyx:17-46).             */
/* mol_atoms[n] = atoms in emission order; mol_ptr[nmol+1].  Returns   */
/* the number of molecules.  labels[i] (optional) = min index of the   */
/* molecule containing i (== its DFS root).                            */
/* ------------------------------------------------------------------ */
int orc_dps(int n, const int *rowptr, const int *col, int *mol_atoms, int *mol_ptr, int *labels) {
    char *visited = (char *)calloc((size_t)(n > 0 ? n : 1), 1);
    size_t cap = (size_t)rowptr[n] + (size_t)n + 1;
    int *stack = (int *)malloc(cap * sizeof(int));
    int nmol = 0, w = 0;
    mol_ptr[0] = 0;
    for (int i = 0; i < n; ++i) {
        if (visited[i]) continue;
        size_t sp = 0;
        stack[sp++] = i;
        while (sp > 0) {
            int s = stack[--sp];
            if (visited[s]) continue;
            mol_atoms[w++] = s;
            if (labels) labels[s] = i;
            for (int p = rowptr[s]; p < rowptr[s + 1]; ++p)
                if (!visited[col[p]]) stack[sp++] = col[p];
            visited[s] = 1;
        }
        mol_ptr[++nmol] = w;
    }
    free(visited);
    free(stack);
    return nmol;
}

/* ------------------------------------------------------------------ */
/* Bond-type keys.  64-bit packing shared with the CUDA path:          */
/*   bits 56..63 element index + 1, bits 48..55 number of levels,      */
/*   bits 4k..4k+3 (k < 12) the k-th smallest level (clipped to 15).   */
/* Bond-file mode (detect.py:112-116): level = max(1, round(bo)) with  */
/* Python's round-half-even, sorted ascending.                         */
/* ------------------------------------------------------------------ */
uint64_t orc_pack_key(int elem, int nlev, const int *sorted_levels) {
    uint64_t k = ((uint64_t)(elem + 1) << 56) | ((uint64_t)nlev << 48);
    for (int i = 0; i < nlev && i < 12; ++i) {
        int l = sorted_levels[i] > 15 ? 15 : sorted_levels[i];
        k |= (uint64_t)l << (4 * i);
    }
    return k;
}

int orc_bond_keys(int n, const int *elem, const int *rowptr, const double *bo, uint64_t *keys) {
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        int lv[64];
        int m = rowptr[i + 1] - rowptr[i];
        if (m > 12) {
            bad++;
            m = 12;
        }
        for (int t = 0; t < m; ++t) {
            double r = nearbyint(bo[rowptr[i] + t]); /* half-even, like Python round() */
            int l = (int)r;
            lv[t] = l < 1 ? 1 : l;
        }
        for (int a = 1; a < m; ++a) {
            int v = lv[a], b = a - 1;
            while (b >= 0 && lv[b] > v) {
                lv[b + 1] = lv[b];
                --b;
            }
            lv[b + 1] = v;
        }
        keys[i] = orc_pack_key(elem[i], m, lv);
    }
    return bad;
}

/* ------------------------------------------------------------------ */
/* ASE minimum-image distances, orthorhombic cells.                    */
/* general_find_mic (geometry.py): wrap_positions(eps=0) then test the */
/* 28 neighbour images; for an orthorhombic cell the images decouple   */
/* per axis and the minimum is min(p, L-p) with p = ((D/L) % 1.0) * L. */
/* ------------------------------------------------------------------ */
static inline double np_remainder1(double f) {
    double m = fmod(f, 1.0);
    if (m != 0.0) {
        if (m < 0.0) m += 1.0;
    } else
        m = 0.0;
    return m;
}

static inline double ase_axis_general(double D, double L) {
    double p = np_remainder1(D / L) * L;
    double q = p - L;
    /* images tested in order 0, -1, 0, +1: the first minimum wins; only the
     * magnitude matters for the distance */
    return (q * q < p * p) ? q : p;
}

static inline double ase_axis_naive(double D, double L) {
    double f = D / L;
    f = f - floor(f + 0.5);
    return f * L;
}

/* distance from atom a to atom j as Atoms.get_distances(a, idx, mic=True)
 * computes it for a whole frame (general branch). */
double orc_ase_dist_general(const double *L, int periodic, const double *ra, const double *rj) {
    double v[3];
    for (int c = 0; c < 3; ++c) {
        double D = rj[c] - ra[c];
        v[c] = periodic ? ase_axis_general(D, L[c]) : D;
    }
    return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
}

/* members of the cutoff sphere of target a, ascending index, includes a itself
 * (datasetbuilder.py:312-315: distances < cutoff).  Returns the count.
 * naive_ok selects ASE's naive_find_mic branch, which find_mic takes when EVERY
 * distance of the query is < 0.5*min(cell lengths); for a 1->all query over a
 * whole frame that is decided by the caller (orc_sphere_auto below). */
static int sphere_impl(int n, const double *pos, const double *L, int periodic, int a, double cutoff,
                       int naive, int *out, int *all_short) {
    int cnt = 0;
    double half = 0.5 * fmin(L[0], fmin(L[1], L[2]));
    int ok = 1;
    for (int j = 0; j < n; ++j) {
        double v[3];
        for (int c = 0; c < 3; ++c) {
            double D = pos[j * 3 + c] - pos[a * 3 + c];
            v[c] = !periodic ? D : (naive ? ase_axis_naive(D, L[c]) : ase_axis_general(D, L[c]));
        }
        double d = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
        if (!(d < half)) ok = 0;
        if (d < cutoff) out[cnt++] = j;
    }
    if (all_short) *all_short = ok;
    return cnt;
}

int orc_sphere(int n, const double *pos, const double *L, int periodic, int a, double cutoff, int *out) {
    if (periodic) {
        int ok = 0;
        int cnt = sphere_impl(n, pos, L, 1, a, cutoff, 1, out, &ok);
        if (ok) return cnt; /* naive_find_mic was safe for the whole query */
        return sphere_impl(n, pos, L, 1, a, cutoff, 0, out, 0);
    }
    return sphere_impl(n, pos, L, 0, a, cutoff, 0, out, 0);
}

/* Coulomb matrix of a cluster (datasetbuilder.py:341-348) with ASE's
 * get_all_distances(mic=True): pairs i<j, D = R[j]-R[i]; naive branch when all
 * pair distances < 0.5*min(L), else the general branch.  M is m*m row-major. */
void orc_coulomb_matrix(int m, const int *Z, const double *diag, const double *pos, const double *L,
                        int periodic, double *M) {
    double half = 0.5 * fmin(L[0], fmin(L[1], L[2]));
    int naive = periodic;
    if (periodic) {
        for (int i = 0; i < m && naive; ++i)
            for (int j = i + 1; j < m; ++j) {
                double v[3];
                for (int c = 0; c < 3; ++c) v[c] = ase_axis_naive(pos[j * 3 + c] - pos[i * 3 + c], L[c]);
                double d = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
                if (!(d < half)) {
                    naive = 0;
                    break;
                }
            }
    }
    for (int i = 0; i < m; ++i) {
        M[i * m + i] = diag[i];
        for (int j = i + 1; j < m; ++j) {
            double v[3];
            for (int c = 0; c < 3; ++c) {
                double D = pos[j * 3 + c] - pos[i * 3 + c];
                v[c] = !periodic ? D : (naive ? ase_axis_naive(D, L[c]) : ase_axis_general(D, L[c]));
            }
            double r = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
            double t = (double)(Z[i] * Z[j]) / r;
            if (isinf(t) || isnan(t)) t = 0.0;
            M[i * m + j] = t;
            M[j * m + i] = t;
        }
    }
}

/* whole-frame helper used by the CPU baseline and the parity tests: for each
 * target, gather the sphere (ascending id) and emit the Coulomb matrix into a
 * ragged buffer.  counts[t] = cluster size; members/M are appended; returns
 * total matrix doubles written, or -1 if a capacity is exceeded. */
long orc_descriptor_gather(int n, const int *Z, const double *diag_by_atom, const double *pos,
                           const double *L, int periodic, double cutoff, int ntargets, const int *targets,
                           int *counts, int *members, long members_cap, double *mats, long mats_cap) {
    int *buf = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    long mw = 0, xw = 0;
    for (int t = 0; t < ntargets; ++t) {
        int m = orc_sphere(n, pos, L, periodic, targets[t], cutoff, buf);
        counts[t] = m;
        if (mw + m > members_cap || xw + (long)m * m > mats_cap) {
            free(buf);
            return -1;
        }
        int *zz = (int *)malloc((size_t)m * sizeof(int));
        double *dd = (double *)malloc((size_t)m * sizeof(double));
        double *pp = (double *)malloc((size_t)m * 3 * sizeof(double));
        for (int k = 0; k < m; ++k) {
            members[mw + k] = buf[k];
            zz[k] = Z[buf[k]];
            dd[k] = diag_by_atom[buf[k]];
            for (int c = 0; c < 3; ++c) pp[k * 3 + c] = pos[buf[k] * 3 + c];
        }
        orc_coulomb_matrix(m, zz, dd, pp, L, periodic, mats + xw);
        mw += m;
        xw += (long)m * m;
        free(zz);
        free(dd);
        free(pp);
    }
    free(buf);
    return xw;
}
