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
    const void* fast;          // prebuilt ZoneFast image (mad_zone_fast_build) or nullptr
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

// ------------------------------------------------------------------------------------------
// Fast classification.  The exact rule above needs atan2/acos in float64 (hundreds of FP64
// instructions per sample).  A direction that is farther than MAD_ZONE_EPS (2e-5 rad) from every
// zone edge gets the same zone from float32 arithmetic (float32 rotation + atan2f: error below
// 2e-6 rad away from the poles; the polar caps need no theta at all).  zone_fast returns that
// zone, or -1 for the ~0.1 % of directions near an edge, which then take the exact path.
// ------------------------------------------------------------------------------------------
#define MAD_ZONE_EPS 2e-5
#define MAD_ZONE_MAX 128
#define MAD_BELT_MAX 32

// Exact squared-magnitude forms of the reference's float32 cut-offs on m = sqrt(m2) (sqrt is
// monotone and correctly rounded):  m < 1e-5f  <=>  m2 < MAD_M2_LT;   m > 1e-5f  <=>  m2 >= MAD_M2_GT.
#define MAD_M2_LT __uint_as_float(0x2edbe6fdu)
#define MAD_M2_GT __uint_as_float(0x2edbe700u)

// atan2 mapped to [0, 2 pi), |error| < 1e-6 rad: degree-13 odd minimax polynomial on min/max (fitted
// and checked over 2e6 float32 arguments: 3.2e-7) + fast division; ~20 instructions instead of ~45.
__device__ __forceinline__ float mad_atan2_2pi(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = (mx > 0.f) ? __fdividef(mn, mx) : 0.f;
    const float s = a * a;
    float p = 0.006811789236962795f;
    p = fmaf(p, s, -0.0336042121052742f);
    p = fmaf(p, s, 0.07962366938591003f);
    p = fmaf(p, s, -0.1323334276676178f);
    p = fmaf(p, s, 0.19807815551757812f);
    p = fmaf(p, s, -0.3331736922264099f);
    p = fmaf(p, s, 0.9999961256980896f);
    p *= a;
    p = (ay > ax) ? 1.5707964f - p : p;
    p = (x < 0.f) ? 3.1415927f - p : p;
    p = (y < 0.f) ? 6.2831855f - p : p;
    return p;
}

// rsqrt.approx.ftz: one MUFU.RSQ (rsqrtf() adds a denormal-range fix-up branch; the callers clamp the argument to
// >= 1e-30 and the fast path's guard band covers the 2-ulp error of either form).
__device__ __forceinline__ float mad_rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct ZoneFast {                    // per-CTA copy in shared memory (zone_fast_init)
    float2 tb[MAD_ZONE_MAX];         // theta bounds shrunk by eps: (tmin + eps, tmax - eps)
    float4 belt[MAD_BELT_MAX];       // x: vz_lo (belt b certain iff vz_lo < vz < vz_hi), y: theta_min of the belt's first
                                     // zone, z: zones / 2 pi, w: bits = first zone | zones << 16
    float vz_hi[MAD_BELT_MAX];       // -inf beyond n_belts, so a fixed-length scan counts only real belts
    int n_belts;
};

static_assert(sizeof(ZoneFast) <= MAD_ZONE_FAST_BYTES && sizeof(ZoneFast) % 4 == 0, "MAD_ZONE_FAST_BYTES too small");

__device__ __forceinline__ void zone_fast_init(ZoneFast* F, const ZoneTab& T) {
    if (T.fast) {                                  // prebuilt once per table: a 1.7 KB copy instead of float64 cosines
        const uint32_t* src = reinterpret_cast<const uint32_t*>(T.fast);
        uint32_t* dst = reinterpret_cast<uint32_t*>(F);
        for (int i = threadIdx.x; i < (int)(sizeof(ZoneFast) / 4); i += blockDim.x) dst[i] = __ldg(src + i);
        return;
    }
    for (int a = threadIdx.x; a < T.n_zones; a += blockDim.x)
        F->tb[a] = make_float2((float)(T.bounds[4 * a + 0] + MAD_ZONE_EPS), (float)(T.bounds[4 * a + 2] - MAD_ZONE_EPS));
    for (int b = threadIdx.x; b < MAD_BELT_MAX; b += blockDim.x) {
        if (b < T.n_belts) {
            // cos is decreasing on [0, pi]; 1e-6 covers the float32 error of the rotated z component
            F->vz_hi[b] = (float)(cos(T.belt_phi[b] + MAD_ZONE_EPS) - 1e-6);
            const int f = T.belt_first[b], nb = T.belt_first[b + 1] - f;
            F->belt[b] = make_float4((float)(cos(T.belt_phi[b + 1] - MAD_ZONE_EPS) + 1e-6), (float)T.bounds[4 * f + 0],
                                     (float)((double)nb / MAD_TWO_PI), __int_as_float(f | (nb << 16)));
        } else {
            F->vz_hi[b] = __int_as_float(0xff800000);            // -inf
            F->belt[b] = make_float4(2.f, 0.f, 0.f, __int_as_float(1 << 16));
        }
    }
    if (threadIdx.x == 0) F->n_belts = T.n_belts;
}

// Upper vz bounds of the first NB belts, held in registers by the caller for the whole sample loop.
template <int NB>
__device__ __forceinline__ void zone_fast_hi(const ZoneFast& F, float (&hi)[NB]) {
#pragma unroll
    for (int k = 0; k < NB; ++k) hi[k] = F.vz_hi[k];
}

// (vx, vy, vz): unit direction in float32.  Returns the zone, or -1 if the exact test must decide.
// Straight-line code (selects, no branches, no loop: NB >= n_belts is a compile-time bound) so that several
// independent samples of one thread interleave in the pipeline; two shared-memory loads per sample.
template <int NB>
__device__ __forceinline__ int zone_fast(const ZoneFast& F, const float (&hi)[NB], float vx, float vy, float vz) {
    // belts are contiguous and descending in vz: the belt is the number of upper bounds above vz
    int cntb = 0;
#pragma unroll
    for (int k = 0; k < NB; ++k) cntb += (vz < hi[k]) ? 1 : 0;
    const int b = max(cntb - 1, 0);
    const float4 B = F.belt[b];
    const bool belt_ok = (cntb > 0) && (vz > B.x);
    const int fb = __float_as_int(B.w);
    const int first = fb & 0xFFFF, nb = fb >> 16;
    const float th = mad_atan2_2pi(vy, vx);
    float u = th - B.y;
    u += (u < 0.f) ? 6.2831855f : 0.f;
    const int z = first + min((int)(u * B.z), nb - 1);
    const float sth = th + 6.2831855f;
    const float2 t = F.tb[z];
    const bool in = ((th > t.x) & (th < t.y)) | ((sth > t.x) & (sth < t.y));
    // a polar cap (one zone in the belt) accepts every theta (theta = 0 passes through theta + 2 pi)
    return (belt_ok && (in || nb == 1)) ? z : -1;
}
