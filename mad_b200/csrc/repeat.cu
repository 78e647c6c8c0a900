// Per-pair repeatability of MaD._match_dsc (mad/MaD.py:426-453), the loop that follows the
// descriptor contraction: for every matched pair (hi, lo) the rigid transform
//     R = inv(Rfinal_lo) . Rfinal_hi,   q = (p - subv_hi) . R^T + subv_lo
// is applied to the cloud of matched hi anchors and the share of transformed anchors that have a
// matched lo anchor within `dist` (cKDTree.query with distance_upper_bound, `distances < dist`) is
// recorded together with the pair's bookkeeping (23 float64 per pair).
//
// One CTA per pair, threads over the hi cloud.  The lo cloud sits in a uniform grid of cell size
// `dist` (cell-sorted points + cell_start, built once per call); a bitmap of the cells whose 27-cell
// neighbourhood holds any lo point rejects most transformed anchors with one probe.
// All arithmetic is float64 like NumPy's.
#include "common.cuh"

namespace {

struct RepeatArgs {
    const int32_t* pair_hi;
    const int32_t* pair_lo;
    const double* pair_score;
    long long n_pairs;
    const double* hi_subv;      // [Dh][3]
    const double* lo_subv;      // [Dl][3]
    const int32_t* hi_meta;     // [Dh][4] index, oct_scale, main_bin, sec_bin
    const int32_t* lo_meta;     // [Dl][4]
    const double* rf;           // [zones*zones][9] Rfinal(main, sec)
    const double* rf_inv;       // [zones*zones][9] inv(Rfinal)
    int zones;
    const double* hi_cloud;     // [H][3]
    int H;
    const double* lo_sorted;    // [L][3] cell-sorted lo cloud
    const int32_t* cell_start;  // [ncell + 1]
    const uint32_t* near_bits;  // [ceil(ncell / 32)]: some lo point in the 27-cell neighbourhood
    double org[3];
    int dims[3];
    double cell;                // = dist
    double dist;
    double* results;            // [n_pairs][23]
};

__global__ void __launch_bounds__(128)
repeatability_kernel(RepeatArgs a) {
    __shared__ double sR[9], shi[3], slo[3];
    __shared__ int s_count;
    const long long pi = blockIdx.x;
    const int tid = threadIdx.x;
    const int h = a.pair_hi[pi], l = a.pair_lo[pi];
    if (tid < 9) {
        // R = inv(Rlo) . Rhi, row-major (np.dot of two 3x3)
        const double* A = a.rf_inv + ((long long)a.lo_meta[4 * l + 2] * a.zones + a.lo_meta[4 * l + 3]) * 9;
        const double* B = a.rf + ((long long)a.hi_meta[4 * h + 2] * a.zones + a.hi_meta[4 * h + 3]) * 9;
        const int i = tid / 3, j = tid % 3;
        sR[tid] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
    }
    if (tid < 3) { shi[tid] = a.hi_subv[3 * (long long)h + tid]; slo[tid] = a.lo_subv[3 * (long long)l + tid]; }
    if (tid == 0) s_count = 0;
    __syncthreads();
    int cnt = 0;
    for (int p = tid; p < a.H; p += blockDim.x) {
        const double dx = a.hi_cloud[3 * p] - shi[0], dy = a.hi_cloud[3 * p + 1] - shi[1], dz = a.hi_cloud[3 * p + 2] - shi[2];
        const double qx = ((dx * sR[0] + dy * sR[1]) + dz * sR[2]) + slo[0];
        const double qy = ((dx * sR[3] + dy * sR[4]) + dz * sR[5]) + slo[1];
        const double qz = ((dx * sR[6] + dy * sR[7]) + dz * sR[8]) + slo[2];
        const int cx = (int)floor((qx - a.org[0]) / a.cell), cy = (int)floor((qy - a.org[1]) / a.cell),
                  cz = (int)floor((qz - a.org[2]) / a.cell);
        if (cx < 0 || cy < 0 || cz < 0 || cx >= a.dims[0] || cy >= a.dims[1] || cz >= a.dims[2]) continue;
        const int c = (cx * a.dims[1] + cy) * a.dims[2] + cz;
        if (!((a.near_bits[c >> 5] >> (c & 31)) & 1u)) continue;
        bool found = false;
        for (int ix = max(cx - 1, 0); ix <= min(cx + 1, a.dims[0] - 1) && !found; ++ix)
            for (int iy = max(cy - 1, 0); iy <= min(cy + 1, a.dims[1] - 1) && !found; ++iy) {
                // the z neighbours are consecutive cells: one contiguous range of sorted points
                const int c0 = (ix * a.dims[1] + iy) * a.dims[2] + max(cz - 1, 0);
                const int c1 = (ix * a.dims[1] + iy) * a.dims[2] + min(cz + 1, a.dims[2] - 1);
                for (int s = a.cell_start[c0]; s < a.cell_start[c1 + 1]; ++s) {
                    const double ex = a.lo_sorted[3 * s] - qx, ey = a.lo_sorted[3 * s + 1] - qy, ez = a.lo_sorted[3 * s + 2] - qz;
                    if (sqrt((ex * ex + ey * ey) + ez * ez) < a.dist) { found = true; break; }
                }
            }
        cnt += found ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if ((tid & 31) == 0 && cnt) atomicAdd(&s_count, cnt);
    __syncthreads();
    double* out = a.results + pi * 23;
    if (tid == 0) {
        out[0] = a.pair_score[pi];
        out[1] = (double)(100LL * s_count) / (double)a.H;      // 100 * count / l
        out[2] = a.lo_meta[4 * l + 0]; out[3] = a.lo_meta[4 * l + 1]; out[4] = a.lo_meta[4 * l + 2];
        out[5] = a.hi_meta[4 * h + 0]; out[6] = a.hi_meta[4 * h + 1]; out[7] = a.hi_meta[4 * h + 2];
    }
    if (tid < 3) { out[8 + tid] = shi[tid]; out[11 + tid] = slo[tid]; }
    if (tid < 9) out[14 + tid] = sR[tid];
}

__global__ void mark_used_kernel(const int32_t* __restrict__ idx, long long n, uint8_t* __restrict__ used) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        used[idx[i]] = 1;
}

}  // namespace

extern "C" int mad_mark_used(const int32_t* idx, long long n, uint8_t* used, void* stream) {
    MAD_CHECK_ARG(n >= 0);
    if (n == 0) return MAD_OK;
    MAD_CHECK_ARG(idx && used);
    const int blocks = (int)std::min<long long>(mad_ceil_div(n, 256), (long long)mad_sm_count() * 8);
    MAD_PROF("mark_used_kernel", stream);
    mark_used_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(idx, n, used);
    MAD_LAUNCH_OK();
    return MAD_OK;
}

extern "C" int mad_repeatability(const int32_t* pair_hi, const int32_t* pair_lo, const double* pair_score, long long n_pairs,
                                 const double* hi_subv, const double* lo_subv, const int32_t* hi_meta, const int32_t* lo_meta,
                                 const double* rf_table, const double* rf_inv_table, int rf_zones,
                                 const double* hi_cloud, int n_hi_cloud, const double* lo_sorted, const int32_t* cell_start,
                                 const uint32_t* near_bits, const double* grid_org_host, const int* grid_dims_host,
                                 double dist, double* results, void* stream) {
    MAD_CHECK_ARG(n_pairs >= 0 && n_pairs < (1LL << 31));
    if (n_pairs == 0) return MAD_OK;
    MAD_CHECK_ARG(pair_hi && pair_lo && pair_score && hi_subv && lo_subv && hi_meta && lo_meta && rf_table && rf_inv_table);
    MAD_CHECK_ARG(hi_cloud && n_hi_cloud > 0 && lo_sorted && cell_start && near_bits && grid_org_host && grid_dims_host && results);
    MAD_CHECK_ARG(dist > 0.0 && rf_zones > 0);
    RepeatArgs a;
    a.pair_hi = pair_hi; a.pair_lo = pair_lo; a.pair_score = pair_score; a.n_pairs = n_pairs;
    a.hi_subv = hi_subv; a.lo_subv = lo_subv; a.hi_meta = hi_meta; a.lo_meta = lo_meta;
    a.rf = rf_table; a.rf_inv = rf_inv_table; a.zones = rf_zones;
    a.hi_cloud = hi_cloud; a.H = n_hi_cloud; a.lo_sorted = lo_sorted; a.cell_start = cell_start; a.near_bits = near_bits;
    for (int q = 0; q < 3; ++q) { a.org[q] = grid_org_host[q]; a.dims[q] = grid_dims_host[q]; }
    a.cell = dist; a.dist = dist; a.results = results;
    MAD_PROF("repeatability_kernel", stream);
    repeatability_kernel<<<(unsigned)n_pairs, 128, 0, (cudaStream_t)stream>>>(a);
    MAD_LAUNCH_OK();
    return MAD_OK;
}
