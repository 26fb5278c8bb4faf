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
    // detect.py:326-341
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
        const char *nx = advance_lines(r->cur, e, r->steplinenum);
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
