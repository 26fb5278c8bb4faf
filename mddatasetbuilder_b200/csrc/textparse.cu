// textparse.cu -- LAMMPS dump atom lines -> float64 positions on the device.
//
// Replaces the per-line `split()` / `float()` of DetectDump.readcrd (reference detect.py:294-345,
// column discovery :166-171) for the frames of a batch whose raw text has been copied to the GPU
// (mdb_reader_next_raw, textio.cpp).  One thread per 64-byte piece of a frame's atom section: it
// finds the lines that START in its piece and parses each of them to the end (which may lie in a
// later piece), writing x, y, z to pos[frame][id - 1], the reference's sort by id (:343).
//
// Decimal -> double is the classical exact fast path: a decimal significand w <= 2^53 and a power
// of ten 10^|e|, |e| <= 22, are both exactly representable, so ONE correctly rounded multiply or
// divide gives the correctly rounded result, the value Python's float() and std::from_chars
// return.  Anything else (more than 19 digits, larger exponents, inf / nan, a malformed line, a
// wrong atom count, a type that differs from the first frame) only sets the frame's status word:
// the caller parses that frame with the host parser, which either handles it or reports the error.
#include <algorithm>

#include "common.cuh"

#define TP_PIECE 64
#define TP_THREADS 256

struct TextParams {
    const char *text;
    const long long *abeg;  // [F] first atom line
    const long long *fend;  // [F] end of frame
    const int *types;       // [N] LAMMPS type by id-1 (first frame)
    double *pos;            // [F][N][3]
    int *status;            // [F] 0 = parsed, != 0 = parse this frame on the host
    int *count;             // [F] atom lines seen
    int N;
    int cid, cty, cx, cy, cz, maxcol;
};

__constant__ double c_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

__device__ __forceinline__ bool tp_space(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// [+-]digits[.digits][(e|E)[+-]digits] -> correctly rounded double; false = leave it to the host parser
__device__ __forceinline__ bool tp_double(const char *t, const char *te, double &v) {
    bool neg = false;
    if (t < te && *t == '+') {
        ++t;
    }
    if (t < te && *t == '-') {
        neg = true;
        ++t;
    }
    // w = significant digits without leading zeros and without the zeros that trail the last
    // non-zero digit: zeros are held back (zi before the point, zf after it) until a non-zero digit
    // follows; what is still held back at the end only moves the exponent (zi) or is dropped (zf),
    // so "9007199254740992.000000" and "1000000000000000000000" stay on the fast path
    unsigned long long w = 0;
    int nd = 0, frac = 0, ndig = 0, zi = 0, zf = 0;
    bool point = false;
    for (; t < te; ++t) {
        const char c = *t;
        if (c >= '0' && c <= '9') {
            ++ndig;
            if (c == '0') {
                if (w != 0) {
                    if (point) ++zf;
                    else ++zi;
                } else if (point) {
                    ++frac;  // leading zero after the point
                }
            } else {
                nd += zi + zf + 1;
                if (nd > 19) return false;
                for (int k = zi + zf; k > 0; --k) w *= 10ull;
                frac += zf + (point ? 1 : 0);
                zi = zf = 0;
                w = w * 10ull + (unsigned)(c - '0');
            }
        } else if (c == '.' && !point) {
            point = true;
        } else {
            break;
        }
    }
    if (ndig == 0) return false;
    int ex = 0;
    if (t < te) {
        if (*t != 'e' && *t != 'E') return false;
        ++t;
        bool eneg = false;
        if (t < te && (*t == '+' || *t == '-')) {
            eneg = *t == '-';
            ++t;
        }
        if (t >= te) return false;
        for (; t < te; ++t) {
            const char c = *t;
            if (c < '0' || c > '9') return false;
            ex = ex * 10 + (c - '0');
            if (ex > 9999) return false;
        }
        if (eneg) ex = -ex;
    }
    ex += zi - frac;
    double r;
    if (w == 0) {
        r = 0.0;
    } else {
        if (w > (1ull << 53)) return false;
        if (ex >= 0 && ex <= 22)
            r = __dmul_rn((double)w, c_pow10[ex]);
        else if (ex < 0 && ex >= -22)
            r = __ddiv_rn((double)w, c_pow10[-ex]);
        else
            return false;
    }
    v = neg ? -r : r;
    return true;
}

__device__ __forceinline__ bool tp_int(const char *t, const char *te, long long &v) {
    bool neg = false;
    if (t < te && *t == '+') ++t;
    if (t < te && *t == '-') {
        neg = true;
        ++t;
    }
    if (t >= te) return false;
    long long w = 0;
    for (; t < te; ++t) {
        const char c = *t;
        if (c < '0' || c > '9') return false;
        w = w * 10 + (c - '0');
        if (w > (1LL << 40)) return false;
    }
    v = neg ? -w : w;
    return true;
}

__global__ void __launch_bounds__(TP_THREADS) k_parse_dump(const __grid_constant__ TextParams P) {
    const int f = blockIdx.y;
    const long long begin = P.abeg[f], end = P.fend[f];
    if (begin < 0) return;  // no atom section found on the host: status is already set
    const long long a = begin + ((long long)blockIdx.x * TP_THREADS + threadIdx.x) * TP_PIECE;
    const long long b = min(a + (long long)TP_PIECE, end);
    const char *__restrict__ tx = P.text;
    __shared__ int s_lines[TP_THREADS / 32];
    int nlines = 0;
    bool bad = false;
    // first the line starts of the piece as a bit mask (every lane walks its 64 bytes in lock step),
    // then the lines themselves, k-th line of every lane at the same time: parsing a line inside the
    // byte walk would serialise the lanes (each finds its line starts at different offsets)
    unsigned long long starts = 0;
    for (long long p = a; p < b; ++p)
        if (p == begin || tx[p - 1] == '\n') starts |= 1ull << (int)(p - a);
    while (starts) {
        const long long p = a + (__ffsll((long long)starts) - 1);
        starts &= starts - 1;
        // one line, from p to its '\n' (or the end of the frame)
        const char *q = tx + p, *e = tx + end;
        long long id = -1, ty = -1;
        double x = 0.0, y = 0.0, z = 0.0;
        int col = 0;
        bool ok = true;
        while (col <= P.maxcol) {
            while (q < e && tp_space(*q)) ++q;
            if (q >= e || *q == '\n') break;
            const char *te = q;
            while (te < e && !tp_space(*te) && *te != '\n') ++te;
            if (col == P.cid) ok = ok && tp_int(q, te, id);
            else if (col == P.cty) ok = ok && tp_int(q, te, ty);
            else if (col == P.cx) ok = ok && tp_double(q, te, x);
            else if (col == P.cy) ok = ok && tp_double(q, te, y);
            else if (col == P.cz) ok = ok && tp_double(q, te, z);
            q = te;
            ++col;
        }
        if (col == 0) continue;  // blank line: skipped like the host parser does
        if (!ok || col <= P.maxcol || id < 1 || id > P.N) {
            bad = true;
            continue;
        }
        if (ty != P.types[id - 1]) bad = true;
        double *o = P.pos + ((size_t)f * P.N + (size_t)(id - 1)) * 3;
        o[0] = x;
        o[1] = y;
        o[2] = z;
        ++nlines;
    }
    // one atomic per block for the line count (every thread of the block is still here)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nlines += __shfl_xor_sync(0xffffffffu, nlines, o);
    if ((threadIdx.x & 31) == 0) s_lines[threadIdx.x >> 5] = nlines;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < TP_THREADS / 32; ++w) tot += s_lines[w];
        if (tot) atomicAdd(&P.count[f], tot);
    }
    if (bad) atomicOr(&P.status[f], 1);
}

__global__ void k_parse_finish(const long long *__restrict__ abeg, const int *__restrict__ count, int F, int N,
                               int *__restrict__ status) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    if (abeg[f] < 0 || count[f] != N) status[f] |= 1;
}

// ---------------------------------------------------------------------------------------------
// fix reaxff/bonds tables: "id type nb id_1..id_nb mol bo_1..bo_nb abo nlp q" (reference
// detect.py:105-119, 137-142) -> neighbours in file order (0-based) and bond orders, rows by id-1.
// Lines that start with '#' and blank lines are skipped; nbr / bo are pre-filled with -1 / 0.
// ---------------------------------------------------------------------------------------------
struct BondTextParams {
    const char *text;
    const long long *fbeg;
    const long long *fend;
    int32_t *nbr;  // [F][N][width]
    double *bo;    // [F][N][width] or nullptr
    int *status;
    int *count;
    int N, width;
};

__device__ __forceinline__ bool tp_token(const char *&q, const char *e, const char *&te) {
    while (q < e && tp_space(*q)) ++q;
    if (q >= e || *q == '\n') return false;
    te = q;
    while (te < e && !tp_space(*te) && *te != '\n') ++te;
    return true;
}

__global__ void __launch_bounds__(TP_THREADS) k_parse_bonds(const __grid_constant__ BondTextParams P) {
    const int f = blockIdx.y;
    const long long begin = P.fbeg[f], end = P.fend[f];
    const long long a = begin + ((long long)blockIdx.x * TP_THREADS + threadIdx.x) * TP_PIECE;
    const long long b = min(a + (long long)TP_PIECE, end);
    const char *__restrict__ tx = P.text;
    __shared__ int s_lines[TP_THREADS / 32];
    int nlines = 0;
    bool bad = false;
    unsigned long long starts = 0;
    for (long long p = a; p < b; ++p)
        if (p == begin || tx[p - 1] == '\n') starts |= 1ull << (int)(p - a);
    while (starts) {
        const long long p = a + (__ffsll((long long)starts) - 1);
        starts &= starts - 1;
        if (tx[p] == '#') continue;
        const char *q = tx + p, *e = tx + end, *te;
        if (!tp_token(q, e, te)) continue;  // blank line
        long long id = -1, ty = -1, nb = -1;
        bool ok = tp_int(q, te, id);
        q = te;
        ok = ok && tp_token(q, e, te) && tp_int(q, te, ty);
        q = te;
        ok = ok && tp_token(q, e, te) && tp_int(q, te, nb);
        q = te;
        if (!ok || id < 1 || id > P.N || nb < 0 || nb > P.width) {
            bad = true;
            continue;
        }
        int32_t *row = P.nbr + ((size_t)f * P.N + (size_t)(id - 1)) * P.width;
        for (int k = 0; k < (int)nb && ok; ++k) {
            long long j = -1;
            ok = tp_token(q, e, te) && tp_int(q, te, j) && j >= 1 && j <= P.N;
            q = te;
            if (ok) row[k] = (int32_t)(j - 1);
        }
        // s[3 + nb] is the molecule id; the bond orders follow at s[4 + nb : 4 + 2 nb]
        if (ok) {
            ok = tp_token(q, e, te);
            q = te;
        }
        if (ok && P.bo) {
            double *brow = P.bo + ((size_t)f * P.N + (size_t)(id - 1)) * P.width;
            for (int k = 0; k < (int)nb && ok; ++k) {
                double v = 0.0;
                ok = tp_token(q, e, te) && tp_double(q, te, v);
                q = te;
                if (ok) brow[k] = v;
            }
        }
        if (!ok) {
            bad = true;
            continue;
        }
        ++nlines;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nlines += __shfl_xor_sync(0xffffffffu, nlines, o);
    if ((threadIdx.x & 31) == 0) s_lines[threadIdx.x >> 5] = nlines;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < TP_THREADS / 32; ++w) tot += s_lines[w];
        if (tot) atomicAdd(&P.count[f], tot);
    }
    if (bad) atomicOr(&P.status[f], 1);
}

extern "C" int mdb_parse_bonds_text(mdb_ctx *c, const char *d_text, int nframes, int natoms, const long long *d_frame_begin,
                                    const long long *d_frame_end, long long max_frame_bytes, int width, int32_t *d_nbr,
                                    double *d_bo, int *d_status, int *d_count, void *stream_) {
    if (!c) {
        mdb_set_error("mdb_parse_bonds_text: context is NULL");
        return -1;
    }
    if (nframes <= 0 || natoms <= 0 || width <= 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    BondTextParams P;
    P.text = d_text;
    P.fbeg = d_frame_begin;
    P.fend = d_frame_end;
    P.nbr = d_nbr;
    P.bo = d_bo;
    P.status = d_status;
    P.count = d_count;
    P.N = natoms;
    P.width = width;
    const size_t cells = (size_t)nframes * natoms * width;
    MDB_CUDA(cudaMemsetAsync(d_nbr, 0xFF, cells * sizeof(int32_t), stream));
    if (d_bo) MDB_CUDA(cudaMemsetAsync(d_bo, 0, cells * sizeof(double), stream));
    MDB_CUDA(cudaMemsetAsync(d_status, 0, (size_t)nframes * sizeof(int), stream));
    MDB_CUDA(cudaMemsetAsync(d_count, 0, (size_t)nframes * sizeof(int), stream));
    const long long pieces = cdiv(std::max<long long>(max_frame_bytes, 1), TP_PIECE);
    dim3 grid((unsigned)cdiv(pieces, TP_THREADS), (unsigned)nframes);
    {
        ProfScope ps(c, "k_parse_bonds", stream);
        k_parse_bonds<<<grid, TP_THREADS, 0, stream>>>(P);
    }
    k_parse_finish<<<cdiv(nframes, 128), 128, 0, stream>>>(d_frame_begin, d_count, nframes, natoms, d_status);
    c->launches += 2;
    MDB_CUDA(cudaGetLastError());
    return 0;
}

// d_status [F] (int) and d_count [F] (int) are scratch of the caller; d_status is the result.
extern "C" int mdb_parse_dump_text(mdb_ctx *c, const char *d_text, int nframes, int natoms, const long long *d_atoms_begin,
                                   const long long *d_frame_end, long long max_frame_bytes, const int *cols5,
                                   const int32_t *d_types, double *d_pos, int *d_status, int *d_count, void *stream_) {
    if (!c) {
        mdb_set_error("mdb_parse_dump_text: context is NULL");
        return -1;
    }
    if (nframes <= 0 || natoms <= 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    TextParams P;
    P.text = d_text;
    P.abeg = d_atoms_begin;
    P.fend = d_frame_end;
    P.types = d_types;
    P.pos = d_pos;
    P.status = d_status;
    P.count = d_count;
    P.N = natoms;
    P.cid = cols5[0];
    P.cty = cols5[1];
    P.cx = cols5[2];
    P.cy = cols5[3];
    P.cz = cols5[4];
    P.maxcol = std::max(std::max(std::max(P.cid, P.cty), std::max(P.cx, P.cy)), P.cz);
    MDB_CUDA(cudaMemsetAsync(d_status, 0, (size_t)nframes * sizeof(int), stream));
    MDB_CUDA(cudaMemsetAsync(d_count, 0, (size_t)nframes * sizeof(int), stream));
    const long long pieces = cdiv(std::max<long long>(max_frame_bytes, 1), TP_PIECE);
    dim3 grid((unsigned)cdiv(pieces, TP_THREADS), (unsigned)nframes);
    {
        ProfScope ps(c, "k_parse_dump", stream);
        k_parse_dump<<<grid, TP_THREADS, 0, stream>>>(P);
    }
    k_parse_finish<<<cdiv(nframes, 128), 128, 0, stream>>>(d_atoms_begin, d_count, nframes, natoms, d_status);
    c->launches += 2;
    MDB_CUDA(cudaGetLastError());
    return 0;
}
