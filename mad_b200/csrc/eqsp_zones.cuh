// EQSP zone lookup shared by the orientation and description kernels.
//
// The reference tests a direction against every zone independently with STRICT inequalities
// (mad/Orientator.py:324-334, mad/Descriptor.py:176-187):
//     ((tmin < th < tmax) or (tmin < th+2pi < tmax)) and (pmin < ph < pmax)
// Zones of a belt share (pmin, pmax) and tile the circle in index order, so only the zones
// next to an analytic guess can pass; at most two do (the table's 6.2832 > 2*pi makes each
// belt's wrap-around zone overlap its successor by 1.469e-5 rad).
#pragma once
#include "common.cuh"

struct ZoneTab {
    const double* bounds;      // [n_zones][4] theta_min, phi_min, theta_max, phi_max
    const int* belt_first;     // [n_belts + 1]
    const double* belt_phi;    // [n_belts + 1]
    int n_zones;
    int n_belts;
};

#define MAD_TWO_PI 6.283185307179586

__device__ __forceinline__ bool zone_has(const ZoneTab& T, int a, double th, double sth) {
    const double tmin = T.bounds[4 * a + 0], tmax = T.bounds[4 * a + 2];
    return ((th < tmax) && (th > tmin)) || ((sth < tmax) && (sth > tmin));
}

// Writes the passing zones (ascending candidates order not guaranteed) to z[0..1]; returns how many.
__device__ __forceinline__ int zones_of(const ZoneTab& T, double th, double sth, double ph, int z[2]) {
    int b = -1;
    for (int k = 0; k < T.n_belts; ++k)
        if (ph > T.belt_phi[k] && ph < T.belt_phi[k + 1]) { b = k; break; }
    if (b < 0) return 0;
    const int first = T.belt_first[b];
    const int nb = T.belt_first[b + 1] - first;
    int n = 0;
    if (nb <= 3) {
        for (int k = 0; k < nb; ++k)
            if (zone_has(T, first + k, th, sth)) { if (n < 2) z[n] = first + k; ++n; }
        return n < 2 ? n : 2;
    }
    double u = th - T.bounds[4 * first + 0];
    while (u < 0.0) u += MAD_TWO_PI;
    int k = (int)(u * (double)nb / MAD_TWO_PI);
    if (k >= nb) k = nb - 1;
    const int c0 = first + (k + nb - 1) % nb, c1 = first + k, c2 = first + (k + 1) % nb;
    if (zone_has(T, c0, th, sth)) z[n++] = c0;
    if (zone_has(T, c1, th, sth)) { if (n < 2) z[n] = c1; ++n; }
    if (zone_has(T, c2, th, sth)) { if (n < 2) z[n] = c2; ++n; }
    return n < 2 ? n : 2;
}
