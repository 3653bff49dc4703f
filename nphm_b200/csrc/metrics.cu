// Nearest-neighbour distances between two point clouds on the GPU: the core of the reference's evaluation metrics
// (src/NPHM/evaluation/metrics.py:171-194 `distance_p2p`: scipy cKDTree.query on 250 k-point clouds, eval.py:111), from
// which completeness / accuracy / Chamfer-L1/L2 / F-score / normal consistency follow (metrics.py:46-145).
//
// Brute force, tiled: a CTA stages a tile of targets in shared memory (SoA), every thread owns one source point and keeps
// its running minimum of the SQUARED fp32 distance.  250 k x 250 k = 6.25e10 pairs at 4 FLOP each.  The winner's distance is
// then recomputed in fp64 (the KD-tree of the reference works on float64 points), so the result only differs from the
// reference where two targets tie to within fp32 round-off.
#include "common.cuh"

namespace nphm {
namespace metrics {

constexpr int kThreads = 256;
constexpr int kTile = 2048;

__global__ void __launch_bounds__(kThreads) nn_kernel(const float *__restrict__ src, long long n_src, const float *__restrict__ tgt,
                                                      long long n_tgt, double *__restrict__ dist, long long *__restrict__ idx)
{
    __shared__ float tx[kTile], ty[kTile], tz[kTile];
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < n_src) { x = src[i * 3]; y = src[i * 3 + 1]; z = src[i * 3 + 2]; }
    float best = 3.4e38f;
    long long best_j = 0;
    for (long long t0 = 0; t0 < n_tgt; t0 += kTile) {
        const int n = (int)min((long long)kTile, n_tgt - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += kThreads) {
            tx[k] = tgt[(t0 + k) * 3]; ty[k] = tgt[(t0 + k) * 3 + 1]; tz[k] = tgt[(t0 + k) * 3 + 2];
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < n; ++k) {
            const float dx = tx[k] - x, dy = ty[k] - y, dz = tz[k] - z;
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            if (d2 < best) { best = d2; best_j = t0 + k; }
        }
    }
    if (i < n_src) {
        const double dx = (double)tgt[best_j * 3] - (double)x, dy = (double)tgt[best_j * 3 + 1] - (double)y,
                     dz = (double)tgt[best_j * 3 + 2] - (double)z;
        dist[i] = sqrt(dx * dx + dy * dy + dz * dz);
        idx[i] = best_j;
    }
}

}  // namespace metrics
}  // namespace nphm

using namespace nphm;

// dist_dev[i] = min_j |src_i - tgt_j| (fp64), idx_dev[i] = argmin  == scipy.spatial.cKDTree(tgt).query(src) for fp32 clouds
extern "C" int nphm_nearest_neighbors(const float *src_dev, long long n_src, const float *tgt_dev, long long n_tgt,
                                      double *dist_dev, long long *idx_dev, void *stream_)
{
    NPHM_REQUIRE(n_src >= 0 && n_tgt > 0 && (n_src == 0 || (src_dev && dist_dev && idx_dev)) && tgt_dev,
                 "nphm_nearest_neighbors: bad arguments");
    if (n_src == 0) return NPHM_OK;
    metrics::nn_kernel<<<(unsigned)ceil_div(n_src, metrics::kThreads), metrics::kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
        src_dev, n_src, tgt_dev, n_tgt, dist_dev, idx_dev);
    NPHM_CUDA_CHECK(cudaGetLastError());
    return NPHM_OK;
}
