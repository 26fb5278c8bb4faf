// geom.cuh -- exact (double, fixed operation order) minimum-image deltas shared by bonds.cu and
// perceive.cu.  Follows Open Babel's OBUnitCell::MinimumImageCartesian as restated in oracle/oracle.c.
#pragma once
#include "common.cuh"

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
