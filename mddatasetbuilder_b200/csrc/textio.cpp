// textio.cpp -- fast LAMMPS text readers (host C++; no CUDA).
//
// Replaces the per-line Python parsing of the reference:
//   DetectDump._readN / readcrd     detect.py:151-194, 294-345   (dump: "ITEM:" sections)
//   DetectBond._readN / readatombondtype / readmolecule
//                                   detect.py:56-88, 105-119, 137-142 (fix reaxff/bonds tables)
//   DatasetBuilder.lineiter         datasetbuilder.py:616-638    (fixed lines per frame, stride)
// Frames are fixed-size groups of `steplinenum` lines (exactly how lineiter slices the file);
// within a frame the dump is parsed by ITEM: section like readcrd.  Numbers are converted with
// std::from_chars (correctly rounded, identical to Python's float()).
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mdb_b200.h"

void mdb_set_error(const std::string &msg);

namespace {

struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) {
            p = "";
            return true;
        }
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, n, MADV_SEQUENTIAL);
        p = (const char *)m;
        return true;
    }
    void close() {
        if (p && n) munmap((void *)p, n);
        if (fd >= 0) ::close(fd);
        p = nullptr;
        n = 0;
        fd = -1;
    }
};

inline const char *skip_ws(const char *p, const char *e) {
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    return p;
}
inline const char *token_end(const char *p, const char *e) {
    while (p < e && !(*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
    return p;
}
inline bool parse_double(const char *a, const char *b, double &v) {
    if (a < b && *a == '+') ++a;
    auto r = std::from_chars(a, b, v);
    return r.ec == std::errc() && r.ptr == b;
}
inline bool parse_long(const char *a, const char *b, long &v) {
    if (a < b && *a == '+') ++a;
    auto r = std::from_chars(a, b, v);
    return r.ec == std::errc() && r.ptr == b;
}
inline bool starts(const char *p, const char *e, const char *lit) {
    size_t k = strlen(lit);
    return (size_t)(e - p) >= k && memcmp(p, lit, k) == 0;
}
// advance `lines` newline-terminated lines from p; returns pointer after them, or nullptr if the
// buffer ends first (a final line without '\n' counts as a line)
const char *advance_lines(const char *p, const char *e, long lines) {
    for (long k = 0; k < lines; ++k) {
        if (p >= e) return nullptr;
        const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
        p = q ? q + 1 : e;
    }
    return p;
}

// The same walk for `nframes` groups of `lines` lines at once, with the newline counting spread
// over `nthreads` threads: ends[k] = pointer after group k.  Returns the number of COMPLETE groups
// found (<= nframes); a trailing partial group is not reported.  est = a guess of the bytes one
// group takes (the window of the first counting round).
int advance_groups(const char *p, const char *e, long lines, int nframes, int nthreads, size_t est,
                   std::vector<const char *> &ends) {
    ends.clear();
    if (nframes <= 0 || p >= e) return 0;
    if (nthreads <= 1 || nframes * (size_t)lines < 4096) {
        for (int k = 0; k < nframes; ++k) {
            const char *nx = advance_lines(p, e, lines);
            if (!nx) break;
            ends.push_back(nx);
            p = nx;
        }
        return (int)ends.size();
    }
    const size_t BLK = 1 << 18;
    const bool tail_line = e[-1] != '\n';  // a final line without '\n' counts as a line
    std::vector<size_t> cnt;               // newlines per block, blocks start at p
    size_t covered = 0;                    // bytes of [p, e) counted so far
    const long long need = (long long)nframes * lines;
    long long have = 0;
    size_t window = std::max<size_t>(BLK, (size_t)((double)std::max<size_t>(est, 1) * nframes * 1.02));
    while (have < need && p + covered < e) {
        const size_t lim = std::min<size_t>((size_t)(e - p), covered + window);
        const size_t b0 = cnt.size(), b1 = (lim + BLK - 1) / BLK;
        cnt.resize(b1, 0);
        std::atomic<size_t> next(b0);
        auto work = [&]() {
            for (;;) {
                const size_t b = next.fetch_add(1);
                if (b >= b1) break;
                const char *a = p + b * BLK, *z = p + std::min(lim, (b + 1) * BLK);
                cnt[b] = (size_t)std::count(a, z, '\n');
            }
        };
        std::vector<std::thread> th;
        const int nt = (int)std::min<size_t>((size_t)nthreads, b1 - b0);
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        for (size_t b = b0; b < b1; ++b) have += (long long)cnt[b];
        covered = lim;
        // a window that stopped inside a block recounts that block next round
        if (covered < (size_t)(e - p) && covered % BLK != 0) {
            have -= (long long)cnt.back();
            cnt.pop_back();
            covered = cnt.size() * BLK;
        }
        window *= 2;
    }
    const bool at_end = p + covered >= e;
    long long total = have + ((at_end && tail_line) ? 1 : 0);
    const int nfull = (int)std::min<long long>(nframes, total / lines);
    ends.assign((size_t)nfull, nullptr);
    // block prefix sums, then every boundary is located inside its block independently
    std::vector<long long> pre(cnt.size() + 1, 0);
    for (size_t b = 0; b < cnt.size(); ++b) pre[b + 1] = pre[b] + (long long)cnt[b];
    std::atomic<int> nextk(0);
    auto locate = [&]() {
        for (;;) {
            const int k = nextk.fetch_add(1);
            if (k >= nfull) break;
            const long long target = (long long)(k + 1) * lines;  // the target-th newline ends group k
            if (target > have) {  // only possible for the unterminated last line
                ends[k] = e;
                continue;
            }
            const size_t b = (size_t)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin()) - 1;
            long long left = target - pre[b];
            const char *q = p + b * BLK;
            const char *z = p + std::min((size_t)(e - p), (b + 1) * BLK);
            while (left > 0) {
                q = (const char *)memchr(q, '\n', (size_t)(z - q)) + 1;
                --left;
            }
            ends[k] = q;
        }
    };
    std::vector<std::thread> th;
    const int nt = std::min(nthreads, nfull);
    for (int t = 1; t < nt; ++t) th.emplace_back(locate);
    locate();
    for (auto &t : th) t.join();
    return nfull;
}

enum Section { S_NONE, S_TIMESTEP, S_NUMBER, S_BOX, S_ATOMS, S_OTHER };
Section section_of(const char *p, const char *e) {
    // LineType.linecontent (detect.py:356-367)
    if (starts(p, e, "ITEM: TIMESTEP")) return S_TIMESTEP;
    if (starts(p, e, "ITEM: ATOMS")) return S_ATOMS;
    if (starts(p, e, "ITEM: NUMBER OF ATOMS")) return S_NUMBER;
    if (starts(p, e, "ITEM: BOX")) return S_BOX;
    return S_OTHER;
}

}  // namespace

struct mdb_reader {
    int kind = 0;  // 0 dump, 1 bonds
    std::vector<std::string> paths;
    std::vector<Mapped> files;
    int natoms = 0;
    long steplinenum = 0;
    std::vector<int> types;  // LAMMPS type (1-based) per id-1, from the first frame
    int id_idx = 0, tidx = 1, xidx = 2, yidx = 3, zidx = 4;
    // iteration state
    size_t cur_file = 0;
    const char *cur = nullptr;
    long frame_in_file = 0;
    int nthreads = 1;
    size_t est_frame_bytes = 0;  // size of the last frame seen (sizes the newline-counting window)
};

static int fail(const std::string &m) {
    mdb_set_error(m);
    return -1;
}

// ---- _readN ----------------------------------------------------------------------------------
static int dump_readN(mdb_reader *r) {
    // detect.py:151-194: scan until the second "NUMBER OF ATOMS" body line
    const Mapped &f = r->files[0];
    const char *p = f.p, *e = f.p + f.n;
    Section sec = S_NONE;
    bool completed = false;
    long index = 0, stepa = -1, stepb = -1;
    int N = -1;
    while (p < e) {
        const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = q ? q : e;
        if (starts(p, le, "ITEM:")) {
            sec = section_of(p, le);
            if (sec == S_ATOMS) {
                // column names follow "ITEM: ATOMS"
                std::vector<std::string> keys;
                const char *t = p;
                while (t < le) {
                    t = skip_ws(t, le);
                    const char *te = token_end(t, le);
                    if (te > t) keys.emplace_back(t, te);
                    t = te;
                }
                auto find = [&](const char *name) {
                    for (size_t k = 0; k < keys.size(); ++k)
                        if (keys[k] == name) return (int)k - 2;
                    return -1000;
                };
                r->id_idx = find("id");
                r->tidx = find("type");
                r->xidx = find("x");
                r->yidx = find("y");
                r->zidx = find("z");
                if (r->id_idx < 0 || r->tidx < 0 || r->xidx < 0 || r->yidx < 0 || r->zidx < 0)
                    return fail("dump file: ITEM: ATOMS must list id type x y z");
            }
        } else {
            if (sec == S_NONE) return fail("No ITEM: in the dump file");
            if (sec == S_NUMBER) {
                if (completed) {
                    stepb = index;
                    break;
                }
                completed = true;
                stepa = index;
                const char *t = skip_ws(p, le);
                long v = 0;
                if (!parse_long(t, token_end(t, le), v) || v < 0) return fail("dump file: bad NUMBER OF ATOMS");
                N = (int)v;
                r->types.assign((size_t)N, 0);
            } else if (sec == S_ATOMS) {
                const char *t = p;
                long id = -1, ty = -1;
                for (int c = 0; t < le; ++c) {
                    t = skip_ws(t, le);
                    const char *te = token_end(t, le);
                    if (te == t) break;
                    if (c == r->id_idx) parse_long(t, te, id);
                    if (c == r->tidx) parse_long(t, te, ty);
                    t = te;
                }
                if (id < 1 || id > N) return fail("dump file: atom id outside 1..N");
                r->types[(size_t)id - 1] = (int)ty;
            }
        }
        p = q ? q + 1 : e;
        ++index;
    }
    if (stepa < 0 || stepb < 0 || N < 0) return fail("The dump file is not completed");
    r->natoms = N;
    r->steplinenum = stepb - stepa;
    return 0;
}

static int bonds_readN(mdb_reader *r) {
    // detect.py:56-88
    const Mapped &f = r->files[0];
    const char *p = f.p, *e = f.p + f.n;
    bool completed = false;
    long index = 0, stepa = -1, stepb = -1;
    int N = -1;
    while (p < e) {
        const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = q ? q : e;
        if (p < le && *p == '#') {
            if (starts(p, le, "# Number of particles")) {
                if (completed) {
                    stepb = index;
                    break;
                }
                completed = true;
                stepa = index;
                // first all-digit token
                const char *t = p;
                long v = -1;
                while (t < le) {
                    t = skip_ws(t, le);
                    const char *te = token_end(t, le);
                    bool digits = te > t;
                    for (const char *c = t; c < te; ++c) digits = digits && (*c >= '0' && *c <= '9');
                    if (digits) {
                        parse_long(t, te, v);
                        break;
                    }
                    t = te;
                }
                if (v < 0) return fail("bond file: no particle count");
                N = (int)v;
                r->types.assign((size_t)N, 0);
            }
        } else if (le > p) {
            if (N < 0) return fail("The bond file is not completed");
            const char *t = skip_ws(p, le);
            const char *te = token_end(t, le);
            long id = -1, ty = -1;
            parse_long(t, te, id);
            t = skip_ws(te, le);
            te = token_end(t, le);
            parse_long(t, te, ty);
            if (id >= 1 && id <= N) r->types[(size_t)id - 1] = (int)ty;
        }
        p = q ? q + 1 : e;
        ++index;
    }
    if (stepa < 0 || stepb < 0 || N < 0) return fail("The bond file is not completed");
    r->natoms = N;
    r->steplinenum = stepb - stepa;
    return 0;
}

// ---- one frame ---------------------------------------------------------------------------------
// dump frame (readcrd, detect.py:294-345): positions placed at id-1, 3x3 cell incl. tilt
// lower-triangular cell from the `xlo xhi [tilt]` bounds, tilt correction included (detect.py:326-341)
static void box_to_cell(const double box[3][3], int boxcols, double *cell) {
    double xy = 0, xz = 0, yz = 0;
    if (boxcols > 2) {
        xy = box[0][2];
        xz = box[1][2];
        yz = box[2][2];
    }
    const double xlo = box[0][0] - std::min({0.0, xy, xz, xy + xz});
    const double xhi = box[0][1] - std::max({0.0, xy, xz, xy + xz});
    const double ylo = box[1][0] - std::min(0.0, yz);
    const double yhi = box[1][1] - std::max(0.0, yz);
    const double zlo = box[2][0], zhi = box[2][1];
    const double c9[9] = {xhi - xlo, 0.0, 0.0, xy, yhi - ylo, 0.0, xz, yz, zhi - zlo};
    memcpy(cell, c9, sizeof(c9));
}

static int parse_dump_frame(const mdb_reader *r, const char *p, const char *e, double *pos, double *cell, int *type_mismatch,
                            std::string *err) {
    const int N = r->natoms;
    Section sec = S_NONE;
    double box[3][3];
    int nbox = 0, boxcols = 0, natom_lines = 0;
    const int maxcol = std::max({r->id_idx, r->tidx, r->xidx, r->yidx, r->zidx});
    while (p < e) {
        const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = q ? q : e;
        const char *t = skip_ws(p, le);
        if (t < le) {
            if (starts(p, le, "ITEM:"))
                sec = section_of(p, le);
            else if (sec == S_NONE) {
                *err = "No ITEM: in the dump file";
                return -1;
            } else if (sec == S_ATOMS) {
                long id = -1, ty = -1;
                double x = 0, y = 0, z = 0;
                bool ok = true;
                int c = 0;
                for (; t < le && c <= maxcol; ++c) {
                    t = skip_ws(t, le);
                    const char *te = token_end(t, le);
                    if (te == t) break;
                    if (c == r->id_idx) ok = ok && parse_long(t, te, id);
                    else if (c == r->tidx) ok = ok && parse_long(t, te, ty);
                    else if (c == r->xidx) ok = ok && parse_double(t, te, x);
                    else if (c == r->yidx) ok = ok && parse_double(t, te, y);
                    else if (c == r->zidx) ok = ok && parse_double(t, te, z);
                    t = te;
                }
                if (!ok || c <= maxcol || id < 1 || id > N) {
                    *err = "dump frame: malformed atom line";
                    return -1;
                }
                double *o = pos + (size_t)(id - 1) * 3;
                o[0] = x;
                o[1] = y;
                o[2] = z;
                if (ty != r->types[(size_t)id - 1]) *type_mismatch = 1;
                ++natom_lines;
            } else if (sec == S_BOX) {
                if (nbox < 3) {
                    int c = 0;
                    while (t < le && c < 3) {
                        t = skip_ws(t, le);
                        const char *te = token_end(t, le);
                        if (te == t) break;
                        if (!parse_double(t, te, box[nbox][c])) {
                            *err = "dump frame: malformed box line";
                            return -1;
                        }
                        ++c;
                        t = te;
                    }
                    boxcols = std::max(boxcols, c);
                    ++nbox;
                }
            }
        }
        p = q ? q + 1 : e;
    }
    if (natom_lines != N || nbox != 3) {
        *err = "dump frame: expected " + std::to_string(N) + " atoms and 3 box lines, found " +
               std::to_string(natom_lines) + " and " + std::to_string(nbox);
        return -1;
    }
    box_to_cell(box, boxcols, cell);
    return 0;
}

// bonds frame (detect.py:105-119, 137-142): neighbours in file order (0-based) and bond orders
static int parse_bonds_frame(const mdb_reader *r, const char *p, const char *e, int width, int32_t *nbr, double *bo,
                             std::string *err) {
    const int N = r->natoms;
    for (size_t k = 0; k < (size_t)N * width; ++k) nbr[k] = -1;
    if (bo) memset(bo, 0, (size_t)N * width * sizeof(double));
    int seen = 0;
    while (p < e) {
        const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = q ? q : e;
        const char *t = skip_ws(p, le);
        if (t < le && *p != '#') {
            long id = -1, ty = -1, nb = -1;
            const char *te = token_end(t, le);
            bool ok = parse_long(t, te, id);
            t = skip_ws(te, le);
            te = token_end(t, le);
            ok = ok && parse_long(t, te, ty);
            t = skip_ws(te, le);
            te = token_end(t, le);
            ok = ok && parse_long(t, te, nb);
            if (!ok || id < 1 || id > N || nb < 0) {
                *err = "bond frame: malformed line";
                return -1;
            }
            if (nb > width) {
                *err = "bond frame: an atom lists " + std::to_string(nb) + " bonds; raise max_bonds";
                return -1;
            }
            int32_t *row = nbr + (size_t)(id - 1) * width;
            for (long k = 0; k < nb; ++k) {
                t = skip_ws(te, le);
                te = token_end(t, le);
                long j = -1;
                if (!parse_long(t, te, j) || j < 1 || j > N) {
                    *err = "bond frame: bad neighbour id";
                    return -1;
                }
                row[k] = (int32_t)(j - 1);
            }
            // s[3+nb] is the molecule id; bond orders follow at s[4+nb : 4+2nb]
            t = skip_ws(te, le);
            te = token_end(t, le);
            if (bo) {
                double *brow = bo + (size_t)(id - 1) * width;
                for (long k = 0; k < nb; ++k) {
                    t = skip_ws(te, le);
                    te = token_end(t, le);
                    if (!parse_double(t, te, brow[k])) {
                        *err = "bond frame: bad bond order";
                        return -1;
                    }
                }
            }
            ++seen;
        }
        p = q ? q + 1 : e;
    }
    if (seen != N) {
        *err = "bond frame: expected " + std::to_string(N) + " atom lines, found " + std::to_string(seen);
        return -1;
    }
    return 0;
}

// ---- C ABI --------------------------------------------------------------------------------------
extern "C" int mdb_reader_open(int kind, const char *const *paths, int npaths, int nthreads, mdb_reader **out) {
    if (!out || npaths < 1 || (kind != 0 && kind != 1)) return fail("mdb_reader_open: bad arguments");
    mdb_reader *r = new mdb_reader();
    r->kind = kind;
    r->nthreads = nthreads > 0 ? nthreads : (int)std::max(1u, std::thread::hardware_concurrency());
    for (int k = 0; k < npaths; ++k) {
        r->paths.emplace_back(paths[k]);
        Mapped m;
        if (!m.open(paths[k])) {
            for (auto &f : r->files) f.close();
            delete r;
            return fail(std::string("cannot open ") + paths[k]);
        }
        r->files.push_back(m);
    }
    int rc = kind == 0 ? dump_readN(r) : bonds_readN(r);
    if (rc) {
        for (auto &f : r->files) f.close();
        delete r;
        return rc;
    }
    r->cur_file = 0;
    r->cur = r->files[0].p;
    r->frame_in_file = 0;
    *out = r;
    return 0;
}

extern "C" void mdb_reader_close(mdb_reader *r) {
    if (!r) return;
    for (auto &f : r->files) f.close();
    delete r;
}

extern "C" int mdb_reader_info(const mdb_reader *r, int *natoms, long *steplinenum, int *cols5) {
    if (!r) return fail("reader is NULL");
    if (natoms) *natoms = r->natoms;
    if (steplinenum) *steplinenum = r->steplinenum;
    if (cols5) {
        cols5[0] = r->id_idx;
        cols5[1] = r->tidx;
        cols5[2] = r->xidx;
        cols5[3] = r->yidx;
        cols5[4] = r->zidx;
    }
    return 0;
}

extern "C" int mdb_reader_types(const mdb_reader *r, int32_t *types) {
    if (!r) return fail("reader is NULL");
    for (int k = 0; k < r->natoms; ++k) types[k] = r->types[(size_t)k];
    return 0;
}

extern "C" int mdb_reader_rewind(mdb_reader *r) {
    if (!r) return fail("reader is NULL");
    r->cur_file = 0;
    r->cur = r->files[0].p;
    r->frame_in_file = 0;
    return 0;
}

// Frame boundaries are found in batches (advance_groups counts newlines on all reader threads); this
// hands them out one by one as long as the caller consumes the file contiguously.
struct GroupCursor {
    std::vector<const char *> ends;
    size_t i = 0;
    const char *expect = nullptr;
    const char *next(mdb_reader *r, const char *cur, const char *e, int want) {
        if (i >= ends.size() || expect != cur) {
            advance_groups(cur, e, r->steplinenum, std::max(want, 1), r->nthreads, r->est_frame_bytes, ends);
            i = 0;
            if (ends.empty()) return nullptr;
        }
        const char *nx = ends[i++];
        expect = nx;
        r->est_frame_bytes = (size_t)(nx - cur);
        return nx;
    }
};

// Reads up to max_frames KEPT frames (every `stride`-th group of steplinenum lines of each file,
// like itertools.islice(zip_longest(*[f]*steplinenum), 0, None, stride) at datasetbuilder.py:632-637).
// dump: pos [max_frames][N][3], cell [max_frames][9].  bonds: nbr [max_frames][N][width] (file
// order, -1 padded), bo same shape (may be NULL).  *nread = frames delivered (0 at end of input).
extern "C" int mdb_reader_next(mdb_reader *r, int stride, int max_frames, double *pos, double *cell, int32_t *nbr,
                               double *bo, int width, int *nread) {
    if (!r || stride < 1 || max_frames < 0) return fail("mdb_reader_next: bad arguments");
    struct Span {
        const char *a, *b;
    };
    std::vector<Span> spans;
    GroupCursor gc;
    while ((int)spans.size() < max_frames && r->cur_file < r->files.size()) {
        const Mapped &f = r->files[r->cur_file];
        const char *e = f.p + f.n;
        if (r->cur >= e) {
            ++r->cur_file;
            if (r->cur_file < r->files.size()) {
                r->cur = r->files[r->cur_file].p;
                r->frame_in_file = 0;
            }
            continue;
        }
        const char *nx = gc.next(r, r->cur, e, (max_frames - (int)spans.size()) * stride);
        const bool complete = nx != nullptr;
        if (!complete) nx = e;  // trailing partial group
        if (r->frame_in_file % stride == 0) {
            if (!complete) {
                // a truncated final frame (zip_longest would pad it with None): dropped
                r->cur = e;
                continue;
            }
            spans.push_back({r->cur, nx});
        }
        r->cur = nx;
        ++r->frame_in_file;
    }
    const int nf = (int)spans.size();
    *nread = nf;
    if (nf == 0) return 0;
    // skip mode (no output buffers): the frames are only stepped over, e.g. batches owned by
    // another rank when frames are sharded across GPUs
    if ((r->kind == 0 && !pos) || (r->kind == 1 && !nbr)) return 0;
    const int N = r->natoms;
    std::atomic<int> next(0), bad(0), mismatch(0);
    std::vector<std::string> errs((size_t)nf);
    auto work = [&]() {
        for (;;) {
            int k = next.fetch_add(1);
            if (k >= nf) break;
            int rc;
            if (r->kind == 0) {
                int mm = 0;
                rc = parse_dump_frame(r, spans[k].a, spans[k].b, pos + (size_t)k * N * 3, cell + (size_t)k * 9, &mm, &errs[k]);
                if (mm) mismatch = 1;
            } else {
                rc = parse_bonds_frame(r, spans[k].a, spans[k].b, width, nbr + (size_t)k * N * width,
                                       bo ? bo + (size_t)k * N * width : nullptr, &errs[k]);
            }
            if (rc) bad = 1;
        }
    };
    const int nt = std::min(r->nthreads, nf);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    if (bad) {
        for (auto &s : errs)
            if (!s.empty()) return fail(s);
        return fail("frame parse error");
    }
    if (mismatch) return fail("dump file: atom types change between frames (not supported)");
    return 0;
}

// Header of one dump frame: the cell from the BOX section and the offset of the first line after
// "ITEM: ATOMS" (the atom lines run from there to the end of the frame).  Returns -1 if the frame
// does not have that shape (the caller then parses the frame on the host).
static long long dump_header(const char *p, const char *e, double *cell) {
    Section sec = S_NONE;
    double box[3][3];
    int nbox = 0, boxcols = 0;
    const char *p0 = p;
    while (p < e) {
        const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = q ? q : e;
        const char *t = skip_ws(p, le);
        const char *next = q ? q + 1 : e;
        if (t < le) {
            if (starts(p, le, "ITEM:")) {
                sec = section_of(p, le);
                if (sec == S_ATOMS) {
                    if (nbox != 3) return -1;
                    box_to_cell(box, boxcols, cell);
                    return (long long)(next - p0);
                }
            } else if (sec == S_BOX && nbox < 3) {
                int c = 0;
                while (t < le && c < 3) {
                    t = skip_ws(t, le);
                    const char *te = token_end(t, le);
                    if (te == t) break;
                    if (!parse_double(t, te, box[nbox][c])) return -1;
                    ++c;
                    t = te;
                }
                boxcols = std::max(boxcols, c);
                ++nbox;
            } else if (sec == S_NONE) {
                return -1;
            }
        }
        p = next;
    }
    return -1;
}

// Raw text of the next kept frames, packed back to back into buf (page-locked memory of the
// caller), for the device parsers (csrc/textparse.cu).  frame_begin[k] / frame_end[k] bracket frame
// k inside buf.  dump: atoms_begin[k] is its first atom line (or -1: parse that frame on the host)
// and cell [k][9] comes from its BOX section; bonds: atoms_begin[k] = frame_begin[k], cell unused.
// Stops early when the next frame would not fit `cap`.
extern "C" int mdb_reader_next_raw(mdb_reader *r, int stride, int max_frames, char *buf, size_t cap, long long *frame_begin,
                                   long long *atoms_begin, long long *frame_end, double *cell, int *nread) {
    if (!r || stride < 1 || max_frames < 0 || !buf) return fail("mdb_reader_next_raw: bad arguments");
    struct Span {
        const char *a, *b;
    };
    std::vector<Span> spans;
    size_t used = 0;
    GroupCursor gc;
    while ((int)spans.size() < max_frames && r->cur_file < r->files.size()) {
        const Mapped &f = r->files[r->cur_file];
        const char *e = f.p + f.n;
        if (r->cur >= e) {
            ++r->cur_file;
            if (r->cur_file < r->files.size()) {
                r->cur = r->files[r->cur_file].p;
                r->frame_in_file = 0;
            }
            continue;
        }
        const char *nx = gc.next(r, r->cur, e, (max_frames - (int)spans.size()) * stride);
        const bool complete = nx != nullptr;
        if (!complete) nx = e;
        if (r->frame_in_file % stride == 0) {
            if (!complete) {
                r->cur = e;  // truncated final frame: dropped (see mdb_reader_next)
                continue;
            }
            const size_t sz = (size_t)(nx - r->cur);
            if (used + sz > cap) {
                if (spans.empty()) return fail("mdb_reader_next_raw: one frame does not fit the staging buffer");
                break;  // this frame stays for the next call
            }
            used += sz;
            spans.push_back({r->cur, nx});
        }
        r->cur = nx;
        ++r->frame_in_file;
    }
    const int nf = (int)spans.size();
    *nread = nf;
    if (nf == 0) return 0;
    std::vector<size_t> off((size_t)nf + 1, 0);
    for (int k = 0; k < nf; ++k) off[k + 1] = off[k] + (size_t)(spans[k].b - spans[k].a);
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= nf) break;
            memcpy(buf + off[k], spans[k].a, off[k + 1] - off[k]);
            frame_begin[k] = (long long)off[k];
            frame_end[k] = (long long)off[k + 1];
            if (r->kind == 0) {
                const long long h = dump_header(spans[k].a, spans[k].b, cell + (size_t)k * 9);
                atoms_begin[k] = h < 0 ? -1 : (long long)off[k] + h;
            } else {
                atoms_begin[k] = (long long)off[k];  // bonds tables: '#' lines are skipped by the parser
            }
        }
    };
    const int nt = std::min(r->nthreads, nf);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    return 0;
}

// Parse one frame held in memory (the per-frame drop-in methods readcrd / readatombondtype /
// readmolecule receive the frame's lines): same parsers as above.
extern "C" int mdb_reader_parse_buffer(const mdb_reader *r, const char *buf, size_t len, double *pos, double *cell,
                                       int32_t *nbr, double *bo, int width) {
    if (!r) return fail("reader is NULL");
    std::string err;
    int rc;
    if (r->kind == 0) {
        int mm = 0;
        rc = parse_dump_frame(r, buf, buf + len, pos, cell, &mm, &err);
        if (!rc && mm) return fail("dump frame: atom types differ from the first frame (not supported)");
    } else
        rc = parse_bonds_frame(r, buf, buf + len, width, nbr, bo, &err);
    if (rc) return fail(err);
    return 0;
}

// ---- output: ASE "simple xyz" files, many at once ----------------------------------------------
// The per-structure writer of the reference (ase.io.write(..., format="xyz") at
// datasetbuilder.py:577-585): "<n>\n\n" then one "%-2s %22.15f %22.15f %22.15f\n" line per atom.
// File k holds atoms [offsets[k], offsets[k+1]) of the flat arrays; symbols are 4-byte NUL-padded
// strings.  Files are written by `nthreads` threads (<= 0: the hardware concurrency).
extern "C" int mdb_write_xyz_batch(int nfiles, const char *const *paths, const long long *offsets, const char *symbols4,
                                   const double *pos, int nthreads) {
    if (nfiles < 0 || (nfiles > 0 && (!paths || !offsets || !symbols4 || !pos))) return fail("mdb_write_xyz_batch: bad arguments");
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<int> next(0), bad(-1);
    auto work = [&]() {
        std::string buf;
        char line[128];
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= nfiles) break;
            const long long a = offsets[k], b = offsets[k + 1];
            buf.clear();
            int m = snprintf(line, sizeof(line), "%lld\n\n", b - a);
            buf.append(line, (size_t)m);
            for (long long i = a; i < b; ++i) {
                char sym[5] = {0, 0, 0, 0, 0};
                memcpy(sym, symbols4 + 4 * i, 4);
                m = snprintf(line, sizeof(line), "%-2s %22.15f %22.15f %22.15f\n", sym, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
                buf.append(line, (size_t)m);
            }
            FILE *f = fopen(paths[k], "wb");
            if (!f || fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
                int expect = -1;
                bad.compare_exchange_strong(expect, k);
            }
            if (f) fclose(f);
        }
    };
    std::vector<std::thread> th;
    const int nt = std::min(nthreads, std::max(nfiles, 1));
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    if (bad.load() >= 0) return fail(std::string("cannot write ") + paths[bad.load()]);
    return 0;
}

// ---- output: Gaussian input files, many at once ------------------------------------------------
// The reference's _convertgjf (datasetbuilder.py:468-521), byte for byte: lines joined by "\n" --
// [%chk=<name>.chk when there are several keyword sections,] keywords (+ " guess=fragment=K"), the title
// block, charge / multiplicity line ("0 M" or "0 M 0 m1 0 m2 ..." in fragment mode), one line per atom
// ("S x y z" or "S(Fragment=k) x y z", %.5f), then for every further keyword section "\n--link1--\n",
// [%chk,] keywords, title, "0 M", "\n", and a final "\n".  Multiplicities follow detect_multiplicity
// (:446-466): an O2 molecule is a triplet, otherwise (sum of Z) % 2 + 1; M = sum(m) - K + 1.
// File k holds atoms [offsets[k], offsets[k+1]); its molecules are mol_sizes[mol_offsets[k] ..
// mol_offsets[k+1]) (consecutive runs of the file's atoms); znum are the atomic numbers per atom.
extern "C" int mdb_write_gjf_batch(int nfiles, const char *const *paths, const long long *offsets, const char *symbols4,
                                   const int32_t *znum, const double *pos, const long long *mol_offsets,
                                   const int32_t *mol_sizes, int nkeywords, const char *const *keywords, int fragment,
                                   int nthreads) {
    if (nfiles < 0 || nkeywords < 1 || !keywords ||
        (nfiles > 0 && (!paths || !offsets || !symbols4 || !znum || !pos || !mol_offsets || !mol_sizes)))
        return fail("mdb_write_gjf_batch: bad arguments");
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    static const char *kTitle = "\nGenerated by MDDatasetMaker (Author: Jinzhe Zeng)\n";
    std::atomic<int> next(0), bad(-1);
    auto work = [&]() {
        std::string buf, chk;
        std::vector<int> mult;
        char line[160];
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= nfiles) break;
            const long long a = offsets[k];
            const long long m0 = mol_offsets[k], m1 = mol_offsets[k + 1];
            const int K = (int)(m1 - m0);
            mult.assign((size_t)K, 1);
            long long at = a;
            long long msum = 0;
            for (int q = 0; q < K; ++q) {
                const int sz = mol_sizes[m0 + q];
                long long zsum = 0;
                for (int i = 0; i < sz; ++i) zsum += znum[at + i];
                const bool o2 = sz == 2 && znum[at] == 8 && znum[at + 1] == 8;
                mult[(size_t)q] = o2 ? 3 : (int)(zsum % 2 + 1);
                msum += mult[(size_t)q];
                at += sz;
            }
            const long long whole = msum - K + 1;
            chk.clear();
            if (nkeywords > 1) {
                // %chk=<basename without its extension>.chk
                std::string path(paths[k]);
                const size_t slash = path.find_last_of('/');
                std::string base = slash == std::string::npos ? path : path.substr(slash + 1);
                const size_t dot = base.find_last_of('.');
                if (dot != std::string::npos && dot > 0) base = base.substr(0, dot);
                chk = "%chk=" + base + ".chk";
            }
            buf.clear();
            auto item = [&](const std::string &sitem) {  // an item of the "\n".join that is not the last one
                buf += sitem;
                buf += '\n';
            };
            const bool frag = fragment && K != 1;
            if (!chk.empty()) item(chk);
            if (frag)
                item(std::string(keywords[0]) + " guess=fragment=" + std::to_string(K));
            else
                item(keywords[0]);
            item(kTitle);
            {
                std::string ml = "0 " + std::to_string(whole);
                if (frag)
                    for (int q = 0; q < K; ++q) ml += " 0 " + std::to_string(mult[(size_t)q]);
                item(ml);
            }
            at = a;
            for (int q = 0; q < K; ++q) {
                const int sz = mol_sizes[m0 + q];
                for (int i = 0; i < sz; ++i, ++at) {
                    char sym[5] = {0, 0, 0, 0, 0};
                    memcpy(sym, symbols4 + 4 * at, 4);
                    int m;
                    if (frag)
                        m = snprintf(line, sizeof(line), "%s(Fragment=%d) %.5f %.5f %.5f\n", sym, q + 1, pos[3 * at], pos[3 * at + 1],
                                     pos[3 * at + 2]);
                    else
                        m = snprintf(line, sizeof(line), "%s %.5f %.5f %.5f\n", sym, pos[3 * at], pos[3 * at + 1], pos[3 * at + 2]);
                    buf.append(line, (size_t)m);
                }
            }
            for (int w = 1; w < nkeywords; ++w) {
                item("\n--link1--\n");
                if (!chk.empty()) item(chk);
                item(keywords[w]);
                item(kTitle);
                item("0 " + std::to_string(whole));
                item("\n");
            }
            buf += '\n';  // the last item of the join: "\n" itself, without a separator after it
            FILE *f = fopen(paths[k], "wb");
            if (!f || fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
                int expect = -1;
                bad.compare_exchange_strong(expect, k);
            }
            if (f) fclose(f);
        }
    };
    std::vector<std::thread> th;
    const int nt = std::min(nthreads, std::max(nfiles, 1));
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    if (bad.load() >= 0) return fail(std::string("cannot write ") + paths[bad.load()]);
    return 0;
}
