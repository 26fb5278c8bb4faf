// tc_assign.cu -- K7 on the 5th-generation tensor cores: a screening pass for the k-means assignment.
//
// The reference's assignment (scikit-learn lloyd_iter_chunked_dense, cluster/_k_means_lloyd.pyx:168-213,
// driven from datasetbuilder.py:371-376) compares float64 scores ||c_j||^2 - 2 x.c_j.  Here the
// equivalent quantity s_j = x.c_j - ||c_j||^2 / 2 (to be MAXIMISED) comes out of one tcgen05.mma GEMM
// (kind::f16, fp32 accumulation in TMEM) on split-fp16 operands
//     x ~ x_hi + x_lo,  c ~ c_hi + c_lo,   x.c ~ x_hi.c_hi + x_hi.c_lo + x_lo.c_hi,
// i.e. the inner dimension tripled (A' = [x_hi | x_hi | x_lo], B' = [c_hi | c_lo | c_hi]) after centring
// both operands on the centroid mean so that the products stay small, plus one more 16-wide block that
// carries the norm: A' gets (1, 1, 1, 0...), B' gets -||c||^2/2 split over three fp16 values.  The
// epilogue therefore only keeps, per row, the best and second-best accumulator value.  A row whose gap
// exceeds its error bound takes the tensor-core label; the others (near-ties) are re-decided in float64
// (see kmeans_assign_tc_run), so the final labels are those of the float64 comparison.
//
// One CTA = one 128-row tile against all centroids: A' stays in shared memory (TMA, 128B swizzle),
// B' tiles of 128 centroids stream through a TMA/mbarrier ring, accumulators are double-buffered
// in TMEM (2 x 128 columns), 8 epilogue warps (two per TMEM lane quarter, one column half each) read
// them back with tcgen05.ld.
#include <cuda.h>
#include <cuda_fp16.h>
#include <float.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

#define TC_BM 128
#define TC_BN 128
#define TC_KCH 64                      // fp16 elements per 128-byte swizzle row
#define TC_CHUNK_BYTES (TC_BM * 128)   // one [128 rows][64 elem] slab
#define TC_PARTS 2                     // column parts per TMEM lane quarter: 4 * TC_PARTS epilogue warps
#define TC_PCOLS (TC_BN / TC_PARTS)
#define TC_EPI_END (2 + 4 * TC_PARTS)   // warps 2 .. TC_EPI_END-1 are epilogue warps, warp TC_EPI_END the second MMA issuer
#define TC_THREADS (64 + 128 * TC_PARTS + 32)
#define TC_TAIL (128 + 1536 * (TC_PARTS - 1) + 64)
#define TC_MAXCH 6                     // 3*Kd + 16 <= 384

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    // bounded spin: a protocol bug traps instead of hanging the GPU
    for (uint32_t it = 0; it < (1u << 27); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, 128B-swizzled operand slab: rows are 128 B, 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TC_TMEM_LD32(R, ADDR)                                                                                         \
    asm volatile(                                                                                                    \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                    \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                    \
        : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]), \
          "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]),      \
          "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]),     \
          "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                   \
        : "r"(ADDR))

struct TcParams {
    int rows, K, stages;
    int two_issuers;   // 1: even / odd centroid tiles go to two MMA threads, each with its own half of the B ring
    float *best1, *best2;
    int *idx1;
    long long *trace;   // optional [ntiles][8] clock64 stamps of CTA 0 (tools/tc_trace.py), else nullptr
    // candidate mode: every (row, centroid) whose approximate score is <= thr[row] is appended to cand
    const float *thr;
    int2 *cand;
    int *ncand;
    int candcap;
};

// Ring stage and mbarrier phase of centroid tile t.  With two MMA threads every stage belongs to one of
// them (even tiles: stages 0, 2, ...; odd tiles: 1, 3, ...), so each thread meets the phases of its
// stages in order -- a thread that waited on a stage the OTHER thread consumed last could run a whole
// phase ahead, and a parity wait cannot tell phase k from phase k - 2.
__device__ __forceinline__ void tc_stage(int t, int S, int two, int &s, uint32_t &ph) {
    if (two) {
        const int h = S >> 1, u = t >> 1;
        s = (t & 1) + 2 * (u % h);
        ph = (uint32_t)(u / h) & 1u;
    } else {
        s = t % S;
        ph = (uint32_t)(t / S) & 1u;
    }
}

// KD16 = padded feature width / 16: the inner dimension is 3 * 16 * KD16 + 16, i.e. 3 * KD16 + 1 MMAs per tile
template <int KD16, bool CAND>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_assign_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // carve: A slabs | B ring | barriers | merge staging | tmem base
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    unsigned char *gen = smem + (sbase - smem_u32(smem));
    constexpr int nch = (3 * 16 * KD16 + 16 + TC_KCH - 1) / TC_KCH;
    constexpr int NMMA = 3 * KD16 + 1;
    const int S = P.stages;
    const uint32_t sA = sbase;
    const uint32_t sB = sA + (uint32_t)nch * TC_CHUNK_BYTES;
    const uint32_t tail = sB + (uint32_t)S * nch * TC_CHUNK_BYTES;
    unsigned char *gtail = gen + (tail - sbase);
    const uint32_t bar_afull = tail;                 // 8 B each
    const uint32_t bar_bfull = tail + 8;             // [S]
    const uint32_t bar_bempty = bar_bfull + 8 * 4;   // [S] (S <= 4)
    const uint32_t bar_tfull = bar_bempty + 8 * 4;   // [2]
    const uint32_t bar_tempty = bar_tfull + 16;      // [2]
    float *mrg_s = reinterpret_cast<float *>(gtail + 128);         // [TC_PARTS - 1][3][128]: merge of the column parts
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gtail + 128 + 1536 * (TC_PARTS - 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * TC_BM;
    const int ntiles = (P.K + TC_BN - 1) / TC_BN;

    if (threadIdx.x == 0) {
        mbar_init(bar_afull, 1);
        for (int s = 0; s < S; ++s) {
            mbar_init(bar_bfull + 8 * s, 1);
            mbar_init(bar_bempty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4 * TC_PARTS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(bar_afull, (uint32_t)nch * TC_CHUNK_BYTES);
            for (int c = 0; c < nch; ++c) tma_load_2d(sA + c * TC_CHUNK_BYTES, &tmA, c * TC_KCH, row0, bar_afull);
            for (int t = 0; t < ntiles; ++t) {
                int s;
                uint32_t ph;
                tc_stage(t, S, P.two_issuers, s, ph);
                mbar_wait(bar_bempty + 8 * s, ph ^ 1u);
                mbar_expect_tx(bar_bfull + 8 * s, (uint32_t)nch * TC_CHUNK_BYTES);
                for (int c = 0; c < nch; ++c)
                    tma_load_2d(sB + (s * nch + c) * TC_CHUNK_BYTES, &tmB, c * TC_KCH, t * TC_BN, bar_bfull + 8 * s);
            }
        }
    } else if (warp == 1 || warp == TC_EPI_END) {
        // ===== MMA issuers: warp 1 takes the even centroid tiles (TMEM buffer 0), warp TC_EPI_END the odd
        // ones (buffer 1).  While one thread issues its tile the other is already through the barrier
        // waits of the next one, so the tensor pipe does not idle across the per-tile handshake. =====
        // instruction descriptor: D = f32, A = B = f16, both K-major, N = 128, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        // one thread runs the whole role: no warp-collective work between the issues, and the operand
        // descriptors of MMA i differ from the slab bases only by compile-time offsets
        if (lane == 0) {
            constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SWIZZLE_128B
            const uint32_t alo = ((sA & 0x3FFFFu) >> 4) | (1u << 16);
            const int first = warp == 1 ? 0 : (P.two_issuers ? 1 : ntiles), step = P.two_issuers ? 2 : 1;
            mbar_wait(bar_afull, 0);
            for (int t = first; t < ntiles; t += step) {
                int s;
                uint32_t ph;
                tc_stage(t, S, P.two_issuers, s, ph);
                const int a = t & 1;
                const uint32_t aph = (uint32_t)(t >> 1) & 1u;
                const bool tr = P.trace && blockIdx.x == 0;
                if (tr) P.trace[t * 8 + 0] = clock64();
                mbar_wait(bar_tempty + 8 * a, aph ^ 1u);
                if (tr) P.trace[t * 8 + 1] = clock64();
                mbar_wait(bar_bfull + 8 * s, ph);
                if (tr) P.trace[t * 8 + 2] = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t blo = (((sB + (uint32_t)(s * nch) * TC_CHUNK_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
                const uint32_t tacc = tmem_base + (uint32_t)a * TC_BN;
#pragma unroll
                for (int i = 0; i < NMMA; ++i) {
                    // chunk i / 4, then 16 elements = 32 bytes (2 descriptor units) per step inside the swizzle row
                    const uint32_t off = (uint32_t)((i >> 2) * (TC_CHUNK_BYTES >> 4) + (i & 3) * 2);
                    const uint64_t da = ((uint64_t)DESC_HI << 32) | (uint64_t)(alo + off);
                    const uint64_t db = ((uint64_t)DESC_HI << 32) | (uint64_t)(blo + off);
                    umma_f16(tacc, da, db, idesc, (uint32_t)(i > 0));
                }
                umma_commit(bar_bempty + 8 * s);
                umma_commit(bar_tfull + 8 * a);
                if (tr) P.trace[t * 8 + 3] = clock64();
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..; warp % 4 selects the TMEM lane quarter, (warp - 2) / 4 the column part =====
        const int q = warp & 3;
        const int part = (warp - 2) >> 2;
        const int myrow = row0 + q * 32 + lane;
        // running best / second best (largest) score and the index of the best, in four independent sets
        // (value i goes to set i & 3) so that the updates do not serialise; all updates are branch-free
        float b1[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX}, b2[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
        int ib[4] = {0, 1, 2, 3};  // first of the eight columns (stride 4) that hold the best of the set
        const float mythr = (CAND && myrow < P.rows) ? P.thr[myrow] : FLT_MAX;  // rows past the end never qualify
        for (int t = 0; t < ntiles; ++t) {
            const int a = t & 1;
            const uint32_t aph = (uint32_t)(t >> 1) & 1u;
            const int col0 = t * TC_BN;
            const bool tr = P.trace && blockIdx.x == 0 && threadIdx.x == 64;
            if (tr) P.trace[t * 8 + 4] = clock64();
            mbar_wait(bar_tfull + 8 * a, aph);
            if (tr) P.trace[t * 8 + 5] = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // this warp's 32 lanes x TC_PCOLS columns leave TMEM in one go; the accumulator is handed back to
            // the MMA warp before any arithmetic
            uint32_t r[TC_PCOLS / 32][32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TC_BN + part * TC_PCOLS);
#pragma unroll
            for (int jj = 0; jj < TC_PCOLS / 32; ++jj) TC_TMEM_LD32(r[jj], taddr + 32 * jj);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * a);
            if (tr) P.trace[t * 8 + 6] = clock64();
#pragma unroll
            for (int jj = 0; jj < TC_PCOLS / 32; ++jj) {
                const int cbase = col0 + part * TC_PCOLS + jj * 32;
                if (CAND) {
                    // the same arithmetic as the screening pass, so a row's own best score qualifies again
                    unsigned hit = 0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) hit |= (__uint_as_float(r[jj][i]) >= mythr ? 1u : 0u) << i;
                    while (hit) {
                        const int i = __ffs(hit) - 1;
                        hit &= hit - 1;
                        if (cbase + i < P.K) {  // padding centroids never count
                            const int slot = atomicAdd(P.ncand, 1);
                            if (slot < P.candcap) P.cand[slot] = make_int2(myrow, cbase + i);
                        }
                    }
                } else {
                    // per score: three min/max and nothing else.  Where the best sits is noted per 32-column
                    // chunk and set only (one compare + select per set and chunk); the winner among the eight
                    // columns base + k, base + k + 4, ... is found later, in float64, by k_tc_decide
                    float o1[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) o1[k] = b1[k];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = __uint_as_float(r[jj][i]);
                        const int k = i & 3;
                        b2[k] = fmaxf(b2[k], fminf(b1[k], v));
                        b1[k] = fmaxf(b1[k], v);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) ib[k] = b1[k] > o1[k] ? cbase + k : ib[k];
                }
            }
            if (tr) P.trace[t * 8 + 7] = clock64();
        }
        // merge the four sets of this thread, then the column parts of the row; on equal scores the
        // lower centroid index wins (such rows are re-decided in float64 anyway: their gap is zero)
        float m1 = b1[0], m2 = b2[0];
        int mi = ib[0];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            const int ik = ib[k];
            if (b1[k] > m1 || (b1[k] == m1 && ik < mi)) {
                m2 = fmaxf(m1, b2[k]);
                m1 = b1[k];
                mi = ik;
            } else {
                m2 = fmaxf(m2, b1[k]);
            }
        }
        // parts 1.. hand their (best, second, index) to part 0 of the same lane quarter
        const int slot = q * 32 + lane;
        if (part > 0) {
            float *mg = mrg_s + (part - 1) * 384;
            mg[slot] = m1;
            mg[128 + slot] = m2;
            reinterpret_cast<int *>(mg)[256 + slot] = mi;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_PARTS) : "memory");
        if (!CAND && part == 0 && myrow < P.rows) {
#pragma unroll
            for (int pp = 0; pp < TC_PARTS - 1; ++pp) {
                const float *mg = mrg_s + pp * 384;
                const float c1 = mg[slot], c2 = mg[128 + slot];
                const int ci = reinterpret_cast<const int *>(mg)[256 + slot];
                if (c1 > m1 || (c1 == m1 && ci < mi)) {
                    m2 = fmaxf(m1, c2);
                    m1 = c1;
                    mi = ci;
                } else {
                    m2 = fmaxf(m2, c1);
                }
            }
            P.best1[myrow] = m1;
            P.best2[myrow] = m2;
            P.idx1[myrow] = mi;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
    }
}

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_col_mean(const double *__restrict__ C, int K, int D, double *__restrict__ mu) {
    const int k = blockIdx.x;
    __shared__ double s[256];
    double acc = 0.0;
    for (int j = threadIdx.x; j < K; j += 256) acc += C[(size_t)j * D + k];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) mu[k] = s[0] / K;
}

// power of two S with max||c||^2 / 2 <= 32768 S: the norm block holds -||c||^2 / (2 S) on the centroid
// side and S on the row side, so it stays inside the fp16 range whatever the magnitude of the data
__device__ __forceinline__ float norm_scale(float cmax2) {
    float s = 1.f;
    while (s < 32768.f && cmax2 > 65536.f * s) s *= 2.f;
    return s;
}

// rows of X (or of C when is_c) -> split fp16, tripled inner dimension + the 16-wide norm block
//   X:  [hi | hi | lo | S S S 0...]      C:  [hi | lo | hi | h0 h1 h2 0...],  S (h0 + h1 + h2) ~ -||c'||^2 / 2
// The C side is written in two steps (the scale needs the largest norm): features and norms here,
// the block by k_norm_block; rows_out > rows pads B' to whole 128-centroid tiles with centroids that
// never win (zero features, block -60000).
__global__ void __launch_bounds__(256) k_split_fp16(const double *__restrict__ X, long long ldx, const int *__restrict__ rowidx,
                                                    long long rows, long long rows_out, int D, int Kd,
                                                    const double *__restrict__ mu, int is_c, const float *__restrict__ cmax2,
                                                    __half *__restrict__ out, float *__restrict__ norm2) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows_out) return;
    const int lane = threadIdx.x & 31;
    __half *o = out + (size_t)r * (3 * Kd + 16);
    if (r >= rows) {  // padding centroid
        for (int k = lane; k < 3 * Kd + 16; k += 32) o[k] = __float2half(k == 3 * Kd ? -60000.f : 0.f);
        return;
    }
    const double *x = X + (rowidx ? (long long)rowidx[r] : r) * ldx;
    double n2 = 0.0;
    for (int k = lane; k < Kd; k += 32) {
        __half hi = __float2half(0.f), lo = hi;
        if (k < D) {
            const double v = x[k] - mu[k];
            hi = __double2half(v);
            lo = __double2half(v - (double)__half2float(hi));
            n2 += v * v;
        }
        o[k] = hi;
        o[Kd + k] = is_c ? lo : hi;
        o[2 * Kd + k] = is_c ? hi : lo;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, s);
    if (!is_c && lane < 16) o[3 * Kd + lane] = __float2half(lane < 3 ? norm_scale(*cmax2) : 0.f);
    if (lane == 0) norm2[r] = (float)n2;
}

__global__ void __launch_bounds__(256) k_norm_block(const double *__restrict__ C, int K, int D, int Kd,
                                                    const double *__restrict__ mu, const float *__restrict__ cmax2,
                                                    __half *__restrict__ out) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= K) return;
    const int lane = threadIdx.x & 31;
    double n2 = 0.0;
    for (int k = lane; k < D; k += 32) {
        const double v = C[(size_t)r * D + k] - mu[k];
        n2 += v * v;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, s);
    if (lane < 16) {
        const double h = -0.5 * n2 / (double)norm_scale(*cmax2);
        const __half h0 = __double2half(h);
        const double r1 = h - (double)__half2float(h0);
        const __half h1 = __double2half(r1);
        const __half h2 = __double2half(r1 - (double)__half2float(h1));
        out[(size_t)r * (3 * Kd + 16) + 3 * Kd + lane] = lane == 0 ? h0 : lane == 1 ? h1 : lane == 2 ? h2 : __float2half(0.f);
    }
}

__global__ void __launch_bounds__(256) k_fmax(const float *__restrict__ v, int n, float *__restrict__ out) {
    __shared__ float s[256];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, v[i]);
    s[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] = fmaxf(s[threadIdx.x], s[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s[0];
}

// accept the tensor-core label when the gap beats the error bound; list the rest for float64
// accept the tensor-core decision when the gap beats the error bound; list the rest for float64.
// i1[r] is the first of eight columns (stride 4) one of which holds the row's best approximate score:
// for an accepted row that one beats every other centroid by more than the error bound, so the
// float64 scores of just these eight name it (any summation order will do at that margin).
// One warp per row: 4 lanes per candidate column (features dealt round-robin), 8 candidates.
__global__ void __launch_bounds__(256) k_tc_decide(const float *__restrict__ b1, const float *__restrict__ b2,
                                                   const int *__restrict__ i1, const float *__restrict__ xn2,
                                                   const float *__restrict__ cmax2, long long rows, int K, float relerr,
                                                   const int *__restrict__ rowidx, const double *__restrict__ X, long long ldx,
                                                   const double *__restrict__ C, int D, const double *__restrict__ cn64,
                                                   int *__restrict__ labels, int *__restrict__ amb_pos,
                                                   int *__restrict__ amb_row, float *__restrict__ amb_thr,
                                                   int *__restrict__ namb) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    // |score error| <= relerr * ||x|| * max||c||  (Cauchy-Schwarz on the centred operands)
    //                  + 2^-22 max||c||^2 (the norm block is built from the unsplit centroid)
    //                  + one fp32 rounding of the final sum, 2^-23 (||x|| ||c|| + ||c||^2 / 2)
    const float xc = sqrtf(xn2[r] * (*cmax2));
    const float bound = (relerr + 1.2e-7f) * xc + 3.0e-7f * (*cmax2) + 1e-30f;
    const int first = i1[r];
    const long long xr = rowidx ? (long long)rowidx[r] : r;
    if (b1[r] - b2[r] > 2.f * bound && first < K) {
        const double *x = X + xr * ldx;
        const int j = first + 4 * (lane >> 2), sub = lane & 3;
        double acc = 0.0;
        if (j < K) {
            const double *c = C + (size_t)j * D;
            for (int k = sub; k < D; k += 4) acc = fma(x[k], c[k], acc);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        double sc = j < K ? fma(-2.0, acc, cn64[j]) : DBL_MAX;
        int lab = j;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, sc, o);
            const int ol = __shfl_xor_sync(0xffffffffu, lab, o);
            if (os < sc || (os == sc && ol < lab)) {
                sc = os;
                lab = ol;
            }
        }
        if (lane == 0) labels[r] = lab;
    } else if (lane == 0) {
        const int k = atomicAdd(namb, 1);
        amb_pos[k] = (int)r;
        amb_row[k] = (int)xr;
        // a centroid can only beat the approximate winner if its own approximate score is within
        // 2 * bound of it (each score is off by at most `bound`); the factor leaves room for the
        // float rounding of the threshold itself
        amb_thr[k] = b1[r] - 2.0001f * bound - 1e-30f;
    }
}

__global__ void __launch_bounds__(256) k_scatter_labels(const int *__restrict__ pos, const int *__restrict__ lab, int n,
                                                        int *__restrict__ labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) labels[pos[i]] = lab[i];
}

__global__ void __launch_bounds__(256) k_sample_rows(long long rows, int ns, const int *__restrict__ rowidx,
                                                     int *__restrict__ pos, int *__restrict__ row) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ns) return;
    const long long p = (long long)k * rows / ns;
    pos[k] = (int)p;
    row[k] = rowidx ? rowidx[p] : (int)p;
}

__global__ void __launch_bounds__(256) k_count_mismatch(const int *__restrict__ pos, const int *__restrict__ lab, int n,
                                                        const int *__restrict__ labels, int *__restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && labels[pos[i]] != lab[i]) atomicAdd(bad, 1);
}

// ambiguous rows of the split operand, gathered so that the candidate pass reads them through one map
__global__ void __launch_bounds__(256) k_gather_half_rows(const __half *__restrict__ A, const int *__restrict__ pos, int n,
                                                          int kin, __half *__restrict__ out) {
    const int vec = kin / 8;  // 16-byte pieces per row (kin is a multiple of 16)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * vec) return;
    const int r = (int)(i / vec), v = (int)(i % vec);
    reinterpret_cast<uint4 *>(out)[(size_t)r * vec + v] = reinterpret_cast<const uint4 *>(A)[(size_t)pos[r] * vec + v];
}

// doubles in an order-preserving unsigned form (for atomicMin)
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// exact score of one candidate pair, the arithmetic of k_assign (kmeans.cu): a forward chain of
// fused multiply-adds over the features, then fma(-2, x.c, ||c||^2)
__global__ void __launch_bounds__(256) k_cand_score(const int2 *__restrict__ cand, int ncand, const int *__restrict__ amb_row,
                                                    const double *__restrict__ X, long long ldx, const double *__restrict__ C,
                                                    int D, const double *__restrict__ cn, unsigned long long *__restrict__ key,
                                                    unsigned long long *__restrict__ best) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ncand) return;
    const int2 rc = cand[q];
    const double *x = X + (long long)amb_row[rc.x] * ldx;
    const double *c = C + (size_t)rc.y * D;
    double acc = 0.0;
    for (int k = 0; k < D; ++k) acc = fma(x[k], c[k], acc);
    const unsigned long long kb = ordered_bits(fma(-2.0, acc, cn[rc.y]));
    key[q] = kb;
    atomicMin(&best[rc.x], kb);
}

// lowest centroid index among the candidates that reach the row minimum (strict <, lowest index on ties)
__global__ void __launch_bounds__(256) k_cand_pick(const int2 *__restrict__ cand, int ncand,
                                                   const unsigned long long *__restrict__ key,
                                                   const unsigned long long *__restrict__ best, unsigned *__restrict__ lab) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ncand) return;
    const int2 rc = cand[q];
    if (key[q] == best[rc.x]) atomicMin(&lab[rc.x], (unsigned)rc.y);
}

__global__ void __launch_bounds__(256) k_count_unset(const unsigned *__restrict__ lab, int n, int *__restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && lab[i] == 0xFFFFFFFFu) atomicAdd(bad, 1);
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static int make_map(CUtensorMap *m, void *base, long long rows, int kin) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        mdb_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return -1;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)kin, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)kin * sizeof(__half)};
    cuuint32_t box[2] = {TC_KCH, TC_BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        mdb_set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")");
        return -1;
    }
    return 0;
}

int kmeans_assign_run(mdb_ctx *c, long long rows, int D, int K, const double *d_X, long long ldx, const int *d_rowidx,
                      const double *d_C, int *d_labels, double *d_mind, double *h_inertia, const double *d_w,
                      cudaStream_t stream);
int kmeans_cnorm_run(mdb_ctx *c, const double *d_C, int K, int D, double *d_cn, cudaStream_t stream);

template <bool CAND>
static int tc_launch_mode(int kd16, long long nblocks, size_t smem, const CUtensorMap &tmA, const CUtensorMap &tmB,
                          const TcParams &P, cudaStream_t stream) {
    const dim3 grid((unsigned)nblocks);
#define MDB_TC_LAUNCH(KD)                                                                                               \
    case KD:                                                                                                            \
        MDB_CUDA(cudaFuncSetAttribute(k_assign_tc<KD, CAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_assign_tc<KD, CAND><<<grid, TC_THREADS, smem, stream>>>(tmA, tmB, P);                                         \
        break;
    switch (kd16) {
        MDB_TC_LAUNCH(1)
        MDB_TC_LAUNCH(2)
        MDB_TC_LAUNCH(3)
        MDB_TC_LAUNCH(4)
        MDB_TC_LAUNCH(5)
        MDB_TC_LAUNCH(6)
        MDB_TC_LAUNCH(7)
    }
#undef MDB_TC_LAUNCH
    MDB_CUDA(cudaGetLastError());
    return 0;
}
static int tc_launch(bool cand, int kd16, long long nblocks, size_t smem, const CUtensorMap &tmA, const CUtensorMap &tmB,
                     const TcParams &P, cudaStream_t stream) {
    return cand ? tc_launch_mode<true>(kd16, nblocks, smem, tmA, tmB, P, stream)
                : tc_launch_mode<false>(kd16, nblocks, smem, tmA, tmB, P, stream);
}

// Tensor-core assignment with float64 resolution of near-ties.  *h_namb (optional) receives the
// number of rows that went through the float64 kernel.
int kmeans_assign_tc_run(mdb_ctx *c, long long rows, int D, int K, const double *d_X, long long ldx, const int *d_rowidx,
                         const double *d_C, int *d_labels, long long *h_namb, float *d_diag, cudaStream_t stream) {
    if (rows <= 0) {
        if (h_namb) *h_namb = 0;
        return 0;
    }
    const int Kd = (D + 15) / 16 * 16;
    const int Kin = 3 * Kd + 16;
    const int nch = (Kin + TC_KCH - 1) / TC_KCH;
    if (nch > TC_MAXCH || rows >= (1LL << 31) || K < 1) {
        mdb_set_error("kmeans_assign_tc: D must be <= 112 and rows < 2^31");
        return -1;
    }
    int stages = nch <= 2 ? 4 : (nch <= 3 ? 3 : (nch <= 4 ? 2 : 1));
    size_t smem = 1024 + (size_t)nch * TC_CHUNK_BYTES * (1 + stages) + TC_TAIL;
    while (smem > 227 * 1024 && stages > 1) {
        --stages;
        smem = 1024 + (size_t)nch * TC_CHUNK_BYTES * (1 + stages) + TC_TAIL;
    }
    if (smem > 227 * 1024) {
        mdb_set_error("kmeans_assign_tc: tile does not fit shared memory");
        return -1;
    }
    // the exact kernel below re-uses the arena, so everything that must survive it lives in its own
    // allocation (kept for the lifetime of the context, grown on demand)
    const long long Kpad = cdiv((long long)K, TC_BN) * TC_BN;
    const size_t needA = align256((size_t)rows * Kin * sizeof(__half)), needB = align256((size_t)Kpad * Kin * sizeof(__half));
    const size_t needf = align256((size_t)rows * sizeof(float));
    // near-ties are re-decided from a candidate list as long as at most a quarter of the rows are ambiguous
    const long long ambcap = std::max<long long>(rows / 4, 1024);
    const int candcap = (int)std::min<long long>(16 * ambcap, 1LL << 28);
    const size_t needAa = align256((size_t)ambcap * Kin * sizeof(__half));
    const size_t total = needA + needB + needAa + 5 * needf + 3 * align256((size_t)rows * sizeof(int)) +
                         align256((size_t)K * sizeof(float)) + align256((size_t)K * sizeof(double)) +
                         align256((size_t)D * sizeof(double)) + align256((size_t)candcap * sizeof(int2)) +
                         align256((size_t)candcap * sizeof(unsigned long long)) +
                         align256((size_t)ambcap * sizeof(unsigned long long)) + 8192;
    if (c->tc_bytes < total) {
        MDB_CUDA(cudaDeviceSynchronize());
        if (c->tc_mem) MDB_CUDA(cudaFree(c->tc_mem));
        c->tc_mem = nullptr;
        c->tc_bytes = 0;
        MDB_CUDA(cudaMalloc(&c->tc_mem, total + total / 8));
        c->tc_bytes = total + total / 8;
    }
    char *p = c->tc_mem;
    auto take = [&](size_t b) {
        char *q = p;
        p += align256(b);
        return q;
    };
    __half *A = (__half *)take(needA);
    __half *B = (__half *)take(needB);
    __half *Aamb = (__half *)take(needAa);
    float *b1 = (float *)take(needf), *b2 = (float *)take(needf), *xn2 = (float *)take(needf), *amb_thr = (float *)take(needf);
    int *i1 = (int *)take(rows * sizeof(int)), *amb_pos = (int *)take(rows * sizeof(int)), *amb_row = (int *)take(rows * sizeof(int));
    float *cn = (float *)take((size_t)K * sizeof(float));
    double *cn64 = (double *)take((size_t)K * sizeof(double));
    double *mu = (double *)take((size_t)D * sizeof(double));
    float *cmax2 = (float *)take(256);
    int *namb = (int *)(cmax2 + 8);
    int *ncand = namb + 1;
    int *amb_lab = (int *)take(needf);
    int2 *cand = (int2 *)take((size_t)candcap * sizeof(int2));
    unsigned long long *ckey = (unsigned long long *)take((size_t)candcap * sizeof(unsigned long long));
    unsigned long long *cbest = (unsigned long long *)take((size_t)ambcap * sizeof(unsigned long long));

    MDB_CUDA(cudaMemsetAsync(namb, 0, 2 * sizeof(int), stream));
    {
        ProfScope ps(c, "k_tc_prep", stream);
        k_col_mean<<<D, 256, 0, stream>>>(d_C, K, D, mu);
        k_split_fp16<<<cdiv(Kpad, 8), 256, 0, stream>>>(d_C, D, nullptr, K, Kpad, D, Kd, mu, 1, nullptr, B, cn);
        k_fmax<<<1, 256, 0, stream>>>(cn, K, cmax2);
        k_norm_block<<<cdiv(K, 8), 256, 0, stream>>>(d_C, K, D, Kd, mu, cmax2, B);
    }
    {
        ProfScope ps(c, "k_split_fp16", stream);
        k_split_fp16<<<cdiv(rows, 8), 256, 0, stream>>>(d_X, ldx, d_rowidx, rows, rows, D, Kd, mu, 0, cmax2, A, xn2);
    }
    c->launches += 5;
    CUtensorMap tmA, tmB;
    if (make_map(&tmA, A, rows, Kin) || make_map(&tmB, B, Kpad, Kin)) return -1;
    TcParams P;
    P.rows = (int)rows;
    P.K = K;
    P.two_issuers = stages >= 2 ? 1 : 0;
    if (P.two_issuers) stages &= ~1;  // an even ring: half of it per MMA thread
    P.stages = stages;
    P.trace = nullptr;
    P.thr = nullptr;
    P.cand = nullptr;
    P.ncand = nullptr;
    P.candcap = 0;
    if (getenv("MDB_TC_TRACE") && d_diag) {   // developer switch: the stamps go to the start of d_diag
        P.trace = reinterpret_cast<long long *>(d_diag);
    }
    P.best1 = b1;
    P.best2 = b2;
    P.idx1 = i1;
    {
        ProfScope ps(c, "k_assign_tc", stream);
        if (tc_launch(false, Kd / 16, cdiv(rows, TC_BM), smem, tmA, tmB, P, stream)) return -1;
    }
    c->launches++;
    // per-product relative error of the split-fp16 contraction with fp32 accumulation over Kin terms
    // error model of the split-fp16 contraction, relative to sum_k |x_k c_k| <= ||x|| ||c||:
    //   operand split residuals + dropped lo*lo term            4 * 2^-22
    //   one fp32 rounding of the running sum per MMA k-step    (Kin / 16) * 2^-23
    //   alignment truncation inside a 16-product MMA           16 * 2^-24
    // times a safety factor of 2.  Measured on B200 (tools/tc_error_probe.py): the largest observed error is
    // 1.5e-6 (Kin = 144) to 3.5e-6 (Kin = 384) in these units, i.e. 2.7x to 4x below this bound.
    const float relerr = 2.0f * (4.0f * 2.3841858e-7f + (float)(Kin / 16) * 1.1920929e-7f + 16.0f * 5.9604645e-8f);
    if (kmeans_cnorm_run(c, d_C, K, D, cn64, stream)) return -1;
    {
        ProfScope ps(c, "k_tc_decide", stream);
        k_tc_decide<<<cdiv(rows, 8), 256, 0, stream>>>(b1, b2, i1, xn2, cmax2, rows, K, relerr, d_rowidx, d_X, ldx, d_C, D, cn64,
                                                        d_labels, amb_pos, amb_row, amb_thr, namb);
    }
    c->launches++;
    if (d_diag && !P.trace) {  // diagnostics: best / second-best approximate score, index, ||x'||^2 per row, then max ||c'||^2
        MDB_CUDA(cudaMemcpyAsync(d_diag, b1, rows * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        MDB_CUDA(cudaMemcpyAsync(d_diag + rows, b2, rows * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        MDB_CUDA(cudaMemcpyAsync(d_diag + 2 * rows, i1, rows * sizeof(int), cudaMemcpyDeviceToDevice, stream));
        MDB_CUDA(cudaMemcpyAsync(d_diag + 3 * rows, xn2, rows * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        MDB_CUDA(cudaMemcpyAsync(d_diag + 4 * rows, cmax2, sizeof(float), cudaMemcpyDeviceToDevice, stream));
    }
    int h = 0;
    MDB_CUDA(cudaMemcpyAsync(&h, namb, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MDB_CUDA(cudaStreamSynchronize(stream));
    MDB_CUDA(cudaGetLastError());
    if (h_namb) *h_namb = h;
    if (h > 0) {
        // near-ties: a second tensor-core pass over just these rows lists every centroid whose
        // approximate score is within the error bound of the row's best; those few pairs are then
        // scored in float64 with the arithmetic of k_assign.  Too many ambiguous rows or candidates
        // (degenerate data) fall back to the full float64 kernel for the ambiguous rows.
        bool done = false;
        if (h <= ambcap) {
            k_gather_half_rows<<<cdiv((long long)h * (Kin / 8), 256), 256, 0, stream>>>(A, amb_pos, h, Kin, Aamb);
            CUtensorMap tmAa;
            if (make_map(&tmAa, Aamb, h, Kin)) return -1;
            TcParams Pc = P;
            Pc.rows = h;
            Pc.thr = amb_thr;
            Pc.cand = cand;
            Pc.ncand = ncand;
            Pc.candcap = candcap;
            Pc.trace = nullptr;
            {
                ProfScope ps(c, "k_assign_tc<cand>", stream);
                if (tc_launch(true, Kd / 16, cdiv(h, TC_BM), smem, tmAa, tmB, Pc, stream)) return -1;
            }
            c->launches += 2;
            int nc = 0;
            MDB_CUDA(cudaMemcpyAsync(&nc, ncand, sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
            if (nc >= h && nc <= candcap) {
                MDB_CUDA(cudaMemsetAsync(cbest, 0xFF, (size_t)h * sizeof(unsigned long long), stream));
                MDB_CUDA(cudaMemsetAsync(amb_lab, 0xFF, (size_t)h * sizeof(int), stream));
                MDB_CUDA(cudaMemsetAsync(ncand, 0, sizeof(int), stream));
                {
                    ProfScope ps(c, "k_cand_score", stream);
                    k_cand_score<<<cdiv(nc, 256), 256, 0, stream>>>(cand, nc, amb_row, d_X, ldx, d_C, D, cn64, ckey, cbest);
                }
                k_cand_pick<<<cdiv(nc, 256), 256, 0, stream>>>(cand, nc, ckey, cbest, (unsigned *)amb_lab);
                k_count_unset<<<cdiv(h, 256), 256, 0, stream>>>((const unsigned *)amb_lab, h, ncand);
                c->launches += 3;
                int unset = 0;
                MDB_CUDA(cudaMemcpyAsync(&unset, ncand, sizeof(int), cudaMemcpyDeviceToHost, stream));
                MDB_CUDA(cudaStreamSynchronize(stream));
                done = unset == 0;  // every ambiguous row must have found its own best again
            }
        }
        if (!done && kmeans_assign_run(c, h, D, K, d_X, ldx, amb_row, d_C, amb_lab, nullptr, nullptr, nullptr, stream)) return -1;
        k_scatter_labels<<<cdiv(h, 256), 256, 0, stream>>>(amb_pos, amb_lab, h, d_labels);
        c->launches++;
        MDB_CUDA(cudaGetLastError());
    }
    // self-check: a strided sample of the rows is re-decided in float64; any disagreement (a violated
    // error model) sends the whole call through the float64 kernel -- "used only if it matches"
    const int ns = (int)std::min<long long>(rows, 1024);
    k_sample_rows<<<cdiv(ns, 256), 256, 0, stream>>>(rows, ns, d_rowidx, amb_pos, amb_row);
    if (kmeans_assign_run(c, ns, D, K, d_X, ldx, amb_row, d_C, amb_lab, nullptr, nullptr, nullptr, stream)) return -1;
    MDB_CUDA(cudaMemsetAsync(namb, 0, sizeof(int), stream));
    k_count_mismatch<<<cdiv(ns, 256), 256, 0, stream>>>(amb_pos, amb_lab, ns, d_labels, namb);
    c->launches += 2;
    int bad = 0;
    MDB_CUDA(cudaMemcpyAsync(&bad, namb, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MDB_CUDA(cudaStreamSynchronize(stream));
    if (bad) {
        c->tc_fallbacks++;
        if (h_namb) *h_namb = rows;
        return kmeans_assign_run(c, rows, D, K, d_X, ldx, d_rowidx, d_C, d_labels, nullptr, nullptr, nullptr, stream);
    }
    return 0;
}
