// Marching cubes over an SDF slab on sm_100a, bit-exact (triangle ids, vertex order, double-precision vertex
// positions) with the sequential PyMCubes algorithm restated in oracle/mc_oracle.c.
//
// Reference call site: src/NPHM/utils/reconstruction.py:30  mcubes.marching_cubes(logits, 0.0)
//
// The sequential algorithm numbers vertices in creation order while sweeping cells x-major.  A vertex lives
// on a grid edge and is created by the first cell (in sweep order) that contains that edge, so the id of any
// vertex is   (number of vertices created by earlier cells) + (rank among the vertices its owner cell creates),
// both of which are pure functions of the 8-corner case index of the cells involved:
//   K1 classify : case index per cell (u8) + per-block sums of created vertices / triangles   [HBM bound]
//   K2 scan     : exclusive scan of the block sums (one CTA)
//   K3 vertices : in-block scan -> per-cell vertex base (u32), emit vertex positions (fp64)
//   K4 triangles: per triangle corner, find the owner cell of the grid edge, id = base[owner] + rank
// Algorithmic HBM traffic: 4 B/voxel read + 24 B/vertex + 24 B/triangle written; the scratch adds 1 B (case)
// + 4 B (base) per cell, written once and read sparsely.
#include "common.cuh"
#include "mc_tables.h"

namespace nphm {

namespace {

constexpr int kThreads = 256;
constexpr int kCellsPerThread = 4;
constexpr int kCellsPerBlock = kThreads * kCellsPerThread;

__constant__ unsigned short c_edge_table[256];
__constant__ signed char c_tri_table[256][16];
__constant__ unsigned char c_num_tris[256];
// creation order inside a cell: edges 6, 5, 10, then 0, 1, 2, 3, 4, 7, 8, 9, 11
__constant__ unsigned short c_before_mask[12];     // edges created before edge e in that order
__constant__ unsigned char c_create_order[12];
bool g_tables_loaded[64] = {false};

const int h_create_order[12] = {6, 5, 10, 0, 1, 2, 3, 4, 7, 8, 9, 11};
// corner offsets and edge end corners (Bourke order: edge e runs from corner A[e] to corner B[e])
__constant__ unsigned char c_corner[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
__constant__ unsigned char c_edge_a[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
__constant__ unsigned char c_edge_b[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
// edge id from (axis, offset of its low end point along the two other axes): [axis][d_first][d_second]
// x edges: (dj,dk); y edges: (di,dk); z edges: (di,dj)
__constant__ unsigned char c_edge_of[3][2][2] = {{{0, 4}, {2, 6}}, {{3, 7}, {1, 5}}, {{8, 11}, {9, 10}}};

struct Dims {
    int nx, ny, nz;          // planes
    int cx, cy, cz;          // cells = planes - 1
    long long ncells;
    int x_global0, ghost_lo, negate;
    double iso;
};

__device__ __forceinline__ void cell_coords(const Dims &d, long long c, int &ci, int &cj, int &ck)
{
    ck = (int)(c % d.cz);
    const long long r = c / d.cz;
    cj = (int)(r % d.cy);
    ci = (int)(r / d.cy);
}

// edges whose vertex this cell creates (it is the first cell in sweep order containing the edge)
__device__ __forceinline__ unsigned own_mask(int gi, int cj, int ck)
{
    unsigned m = (1u << 6) | (1u << 5) | (1u << 10);
    const bool i0 = gi == 0, j0 = cj == 0, k0 = ck == 0;
    if (j0 && k0) m |= 1u << 0;
    if (k0) m |= (1u << 2) | (1u << 1);
    if (j0) m |= (1u << 4) | (1u << 9);
    if (i0) m |= (1u << 7) | (1u << 11);
    if (i0 && k0) m |= 1u << 3;
    if (i0 && j0) m |= 1u << 8;
    return m;
}

__device__ __forceinline__ unsigned classify_cell(const float *__restrict__ vol, const Dims &d, int ci, int cj, int ck)
{
    const size_t sx = (size_t)d.ny * d.nz, sy = d.nz;
    const float *p = vol + (size_t)ci * sx + (size_t)cj * sy + ck;
    float v[8];
    v[0] = __ldg(p);           v[1] = __ldg(p + sx);
    v[2] = __ldg(p + sx + sy); v[3] = __ldg(p + sy);
    v[4] = __ldg(p + 1);           v[5] = __ldg(p + sx + 1);
    v[6] = __ldg(p + sx + sy + 1); v[7] = __ldg(p + sy + 1);
    unsigned cube = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const double f = d.negate ? -(double)v[m] : (double)v[m];
        if (f <= d.iso) cube |= 1u << m;
    }
    return cube;
}

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned *warp_sums, unsigned &block_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < (kThreads / 32) ? warp_sums[lane] : 0;
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < kThreads / 32) warp_sums[lane] = winc - w;
        if (lane == kThreads / 32 - 1) warp_sums[kThreads / 32] = winc;
    }
    __syncthreads();
    block_total = warp_sums[kThreads / 32];
    const unsigned res = inc - v + warp_sums[warp];
    __syncthreads();
    return res;
}

// K1 ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) mc_classify_kernel(const float *__restrict__ vol, const Dims d,
                                                               unsigned char *__restrict__ cube_out,
                                                               unsigned *__restrict__ block_v, unsigned *__restrict__ block_t)
{
    __shared__ unsigned warp_sums[kThreads / 32 + 1];
    const long long c0 = ((long long)blockIdx.x * kThreads + threadIdx.x) * kCellsPerThread;
    unsigned nv = 0, nt = 0;
    unsigned char cubes[kCellsPerThread];
#pragma unroll
    for (int u = 0; u < kCellsPerThread; ++u) {
        const long long c = c0 + u;
        cubes[u] = 0;
        if (c < d.ncells) {
            int ci, cj, ck;
            cell_coords(d, c, ci, cj, ck);
            const unsigned cube = classify_cell(vol, d, ci, cj, ck);
            cubes[u] = (unsigned char)cube;
            const unsigned edges = c_edge_table[cube];
            nv += __popc(edges & own_mask(d.x_global0 + ci, cj, ck));
            nt += c_num_tris[cube];
        }
    }
    if (c0 + kCellsPerThread <= d.ncells && (c0 & 3) == 0) {
        *reinterpret_cast<uchar4 *>(cube_out + c0) = make_uchar4(cubes[0], cubes[1], cubes[2], cubes[3]);
    } else {
#pragma unroll
        for (int u = 0; u < kCellsPerThread; ++u)
            if (c0 + u < d.ncells) cube_out[c0 + u] = cubes[u];
    }
    unsigned tot_v, tot_t;
    block_exclusive_scan(nv, warp_sums, tot_v);
    block_exclusive_scan(nt, warp_sums, tot_t);
    if (threadIdx.x == 0) { block_v[blockIdx.x] = tot_v; block_t[blockIdx.x] = tot_t; }
}

// K2: exclusive scan of the block sums in place; totals[0..1] = grand totals, totals[2..3] = ghost layer part
__global__ void __launch_bounds__(1024) mc_scan_blocks_kernel(unsigned *__restrict__ block_v, unsigned *__restrict__ block_t,
                                                              int nblocks, unsigned long long *__restrict__ totals)
{
    __shared__ unsigned long long carry[2];
    __shared__ unsigned wsum[2][33];
    if (threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        unsigned v[2] = {i < nblocks ? block_v[i] : 0u, i < nblocks ? block_t[i] : 0u};
        unsigned inc[2] = {v[0], v[1]};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc[a], o);
                if (lane >= o) inc[a] += t;
            }
            if (lane == 31) wsum[a][warp] = inc[a];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                unsigned w = wsum[a][lane], winc = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
                    if (lane >= o) winc += t;
                }
                wsum[a][lane] = winc - w;
                if (lane == 31) wsum[a][32] = winc;
            }
        }
        __syncthreads();
        if (i < nblocks) {
            block_v[i] = (unsigned)(carry[0] + wsum[0][warp] + inc[0] - v[0]);
            block_t[i] = (unsigned)(carry[1] + wsum[1][warp] + inc[1] - v[1]);
        }
        __syncthreads();
        if (threadIdx.x == 0) { carry[0] += wsum[0][32]; carry[1] += wsum[1][32]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry[0]; totals[1] = carry[1]; }
}

// counts of the ghost layer (cells with ci == 0): number of created vertices / triangles
__global__ void __launch_bounds__(kThreads) mc_ghost_count_kernel(const unsigned char *__restrict__ cube, const Dims d,
                                                                  unsigned long long *__restrict__ totals)
{
    const long long layer = (long long)d.cy * d.cz;
    unsigned nv = 0, nt = 0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < layer; c += (long long)gridDim.x * blockDim.x) {
        int ci, cj, ck;
        cell_coords(d, c, ci, cj, ck);
        const unsigned cb = cube[c];
        nv += __popc(c_edge_table[cb] & own_mask(d.x_global0 + ci, cj, ck));
        nt += c_num_tris[cb];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        nt += __shfl_xor_sync(0xffffffffu, nt, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&totals[2], (unsigned long long)nv);
        atomicAdd(&totals[3], (unsigned long long)nt);
    }
}

// K3 ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double mc_interp(double iso, double fa, double fb, double xa, double xb)
{
    if (fb == fa) return __ddiv_rn(__dadd_rn(xb, xa), 2.0);
    // (xb - xa) * (iso - fa) / (fb - fa) + xa, evaluated left to right without contraction
    const double num = __dmul_rn(__dsub_rn(xb, xa), __dsub_rn(iso, fa));
    return __dadd_rn(__ddiv_rn(num, __dsub_rn(fb, fa)), xa);
}

__global__ void __launch_bounds__(kThreads) mc_vertices_kernel(const float *__restrict__ vol, const Dims d,
                                                               const unsigned char *__restrict__ cube_in,
                                                               const unsigned *__restrict__ block_v,
                                                               unsigned *__restrict__ vbase_out,
                                                               const unsigned long long *__restrict__ totals,
                                                               double *__restrict__ verts)
{
    __shared__ unsigned warp_sums[kThreads / 32 + 1];
    const long long c0 = ((long long)blockIdx.x * kThreads + threadIdx.x) * kCellsPerThread;
    unsigned created[kCellsPerThread];
    unsigned nv = 0;
#pragma unroll
    for (int u = 0; u < kCellsPerThread; ++u) {
        const long long c = c0 + u;
        created[u] = 0;
        if (c < d.ncells) {
            int ci, cj, ck;
            cell_coords(d, c, ci, cj, ck);
            created[u] = c_edge_table[cube_in[c]] & own_mask(d.x_global0 + ci, cj, ck);
            nv += __popc(created[u]);
        }
    }
    unsigned tot;
    unsigned base = block_exclusive_scan(nv, warp_sums, tot) + block_v[blockIdx.x];
    const unsigned ghost_v = d.ghost_lo ? (unsigned)totals[2] : 0u;
    const size_t sx = (size_t)d.ny * d.nz, sy = d.nz;
#pragma unroll
    for (int u = 0; u < kCellsPerThread; ++u) {
        const long long c = c0 + u;
        if (c >= d.ncells) break;
        vbase_out[c] = base;
        const unsigned cm = created[u];
        if (cm) {
            int ci, cj, ck;
            cell_coords(d, c, ci, cj, ck);
            const bool ghost = d.ghost_lo && ci == 0;
            if (!ghost) {
                const float *p = vol + (size_t)ci * sx + (size_t)cj * sy + ck;
                unsigned rank = 0;
#pragma unroll 1
                for (int o = 0; o < 12; ++o) {
                    const int e = c_create_order[o];
                    if (!(cm & (1u << e))) continue;
                    const int a = c_edge_a[e], b = c_edge_b[e];
                    double fa = (double)__ldg(p + c_corner[a][0] * sx + c_corner[a][1] * sy + c_corner[a][2]);
                    double fb = (double)__ldg(p + c_corner[b][0] * sx + c_corner[b][1] * sy + c_corner[b][2]);
                    if (d.negate) { fa = -fa; fb = -fb; }
                    double pos[3] = {(double)(d.x_global0 + ci + c_corner[a][0]), (double)(cj + c_corner[a][1]),
                                     (double)(ck + c_corner[a][2])};
                    const int axis = c_corner[a][0] != c_corner[b][0] ? 0 : (c_corner[a][1] != c_corner[b][1] ? 1 : 2);
                    const double qb = pos[axis] + ((double)c_corner[b][axis] - (double)c_corner[a][axis]);
                    pos[axis] = mc_interp(d.iso, fa, fb, pos[axis], qb);
                    double *dst = verts + (size_t)(base - ghost_v + rank) * 3;
                    dst[0] = pos[0]; dst[1] = pos[1]; dst[2] = pos[2];
                    ++rank;
                }
            }
        }
        base += __popc(cm);
    }
}

// K4 ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) mc_triangles_kernel(const Dims d, const unsigned char *__restrict__ cube_in,
                                                                const unsigned *__restrict__ block_t,
                                                                const unsigned *__restrict__ vbase,
                                                                const unsigned long long *__restrict__ totals,
                                                                long long vert_id_base, long long *__restrict__ tris)
{
    __shared__ unsigned warp_sums[kThreads / 32 + 1];
    const long long c0 = ((long long)blockIdx.x * kThreads + threadIdx.x) * kCellsPerThread;
    unsigned char cubes[kCellsPerThread];
    unsigned nt = 0;
#pragma unroll
    for (int u = 0; u < kCellsPerThread; ++u) {
        const long long c = c0 + u;
        cubes[u] = c < d.ncells ? cube_in[c] : 0;
        nt += c_num_tris[cubes[u]];
    }
    unsigned tot;
    unsigned tbase = block_exclusive_scan(nt, warp_sums, tot) + block_t[blockIdx.x];
    if (nt == 0) return;
    const unsigned ghost_v = d.ghost_lo ? (unsigned)totals[2] : 0u;
    const unsigned ghost_t = d.ghost_lo ? (unsigned)totals[3] : 0u;
#pragma unroll 1
    for (int u = 0; u < kCellsPerThread; ++u) {
        const unsigned cube = cubes[u];
        const int n = c_num_tris[cube];
        if (n == 0) continue;
        const long long c = c0 + u;
        int ci, cj, ck;
        cell_coords(d, c, ci, cj, ck);
        const bool ghost = d.ghost_lo && ci == 0;
        if (!ghost) {
            long long ids[12];
            const unsigned edges = c_edge_table[cube];
#pragma unroll 1
            for (int e = 0; e < 12; ++e) {
                if (!(edges & (1u << e))) continue;
                const int a = c_edge_a[e], b = c_edge_b[e];
                const int axis = c_corner[a][0] != c_corner[b][0] ? 0 : (c_corner[a][1] != c_corner[b][1] ? 1 : 2);
                const int lo = c_corner[a][axis] == 0 ? a : b;
                // low end point of the grid edge (slab-local cell coordinates)
                const int gi = ci + c_corner[lo][0], gj = cj + c_corner[lo][1], gk = ck + c_corner[lo][2];
                // owner: first cell in sweep order that contains the edge
                int fi = gi, fj = gj, fk = gk;
                if (axis != 0) fi = (d.x_global0 + gi) > 0 ? gi - 1 : gi;
                if (axis != 1) fj = gj > 0 ? gj - 1 : 0;
                if (axis != 2) fk = gk > 0 ? gk - 1 : 0;
                int d1, d2;
                if (axis == 0)      { d1 = gj - fj; d2 = gk - fk; }
                else if (axis == 1) { d1 = gi - fi; d2 = gk - fk; }
                else                { d1 = gi - fi; d2 = gj - fj; }
                const int eo = c_edge_of[axis][d1][d2];
                const long long oc = ((long long)fi * d.cy + fj) * d.cz + fk;
                const unsigned ocreated = c_edge_table[cube_in[oc]] & own_mask(d.x_global0 + fi, fj, fk);
                const unsigned rank = __popc(ocreated & c_before_mask[eo]);
                ids[e] = vert_id_base + (long long)(vbase[oc] + rank) - (long long)ghost_v;
            }
            long long *dst = tris + (size_t)(tbase - ghost_t) * 3;
            const signed char *row = c_tri_table[cube];
            for (int m = 0; m < 3 * n; ++m) dst[m] = ids[row[m]];
        }
        tbase += n;
    }
}

int load_tables()
{
    int dev = 0;
    NPHM_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 64 && g_tables_loaded[dev]) return NPHM_OK;
    NPHM_CUDA_CHECK(cudaMemcpyToSymbol(c_edge_table, MC_EDGE_TABLE, sizeof(MC_EDGE_TABLE)));
    NPHM_CUDA_CHECK(cudaMemcpyToSymbol(c_tri_table, MC_TRI_TABLE, sizeof(MC_TRI_TABLE)));
    NPHM_CUDA_CHECK(cudaMemcpyToSymbol(c_num_tris, MC_NUM_TRIS, sizeof(MC_NUM_TRIS)));
    unsigned short before[12];
    unsigned char order[12];
    unsigned short acc = 0;
    for (int o = 0; o < 12; ++o) {
        order[o] = (unsigned char)h_create_order[o];
        before[h_create_order[o]] = acc;
        acc |= (unsigned short)(1u << h_create_order[o]);
    }
    NPHM_CUDA_CHECK(cudaMemcpyToSymbol(c_before_mask, before, sizeof(before)));
    NPHM_CUDA_CHECK(cudaMemcpyToSymbol(c_create_order, order, sizeof(order)));
    if (dev < 64) g_tables_loaded[dev] = true;
    return NPHM_OK;
}

int make_dims(const nphm_mc_params *p, Dims &d)
{
    NPHM_REQUIRE(p != nullptr, "nphm_mc: params is NULL");
    NPHM_REQUIRE(p->nx >= 0 && p->ny >= 0 && p->nz >= 0, "nphm_mc: negative dimensions");
    // a slab that does not start at the global x = 0 plane must bring the cell layer below it (its owner cells)
    NPHM_REQUIRE(p->x_global0 >= 0 && (p->x_global0 == 0 || p->ghost_lo), "nphm_mc: x_global0 > 0 needs ghost_lo = 1");
    d.nx = p->nx; d.ny = p->ny; d.nz = p->nz;
    d.cx = p->nx - 1; d.cy = p->ny - 1; d.cz = p->nz - 1;
    d.ncells = (d.cx > 0 && d.cy > 0 && d.cz > 0) ? (long long)d.cx * d.cy * d.cz : 0;
    d.x_global0 = p->x_global0; d.ghost_lo = p->ghost_lo ? 1 : 0; d.negate = p->negate ? 1 : 0;
    d.iso = p->iso;
    NPHM_REQUIRE(d.ncells < (1LL << 31) * (long long)kCellsPerBlock, "nphm_mc: volume too large");
    return NPHM_OK;
}

struct Workspace {
    unsigned char *cube;
    unsigned *vbase, *block_v, *block_t;
    unsigned long long *totals;
    long long bytes;
};

Workspace carve(void *ws, const Dims &d)
{
    const long long nblocks = ceil_div(d.ncells > 0 ? d.ncells : 1, kCellsPerBlock);
    auto align = [](long long x) { return (x + 255) / 256 * 256; };
    Workspace w;
    char *p = reinterpret_cast<char *>(ws);
    long long off = 0;
    w.totals = reinterpret_cast<unsigned long long *>(p + off); off += 256;
    w.block_v = reinterpret_cast<unsigned *>(p + off); off += align(nblocks * 4);
    w.block_t = reinterpret_cast<unsigned *>(p + off); off += align(nblocks * 4);
    w.vbase = reinterpret_cast<unsigned *>(p + off); off += align(d.ncells * 4);
    w.cube = reinterpret_cast<unsigned char *>(p + off); off += align(d.ncells);
    w.bytes = off;
    return w;
}

}  // namespace
}  // namespace nphm

using namespace nphm;

extern "C" long long nphm_mc_workspace_bytes(const nphm_mc_params *p)
{
    Dims d;
    if (make_dims(p, d) != NPHM_OK) return -1;
    return carve(nullptr, d).bytes;
}

extern "C" int nphm_mc_count(const float *vol_dev, const nphm_mc_params *p, void *workspace_dev,
                             long long *n_verts_host, long long *n_tris_host, void *stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Dims d;
    int rc = make_dims(p, d);
    if (rc) return rc;
    NPHM_REQUIRE(n_verts_host && n_tris_host, "nphm_mc_count: NULL count pointers");
    *n_verts_host = 0; *n_tris_host = 0;
    if (d.ncells == 0) return NPHM_OK;
    NPHM_REQUIRE(vol_dev && workspace_dev, "nphm_mc_count: NULL volume / workspace");
    rc = load_tables();
    if (rc) return rc;
    Workspace w = carve(workspace_dev, d);
    const int nblocks = (int)ceil_div(d.ncells, kCellsPerBlock);
    NPHM_CUDA_CHECK(cudaMemsetAsync(w.totals, 0, 256, stream));
    mc_classify_kernel<<<nblocks, kThreads, 0, stream>>>(vol_dev, d, w.cube, w.block_v, w.block_t);
    NPHM_CUDA_CHECK(cudaGetLastError());
    mc_scan_blocks_kernel<<<1, 1024, 0, stream>>>(w.block_v, w.block_t, nblocks, w.totals);
    NPHM_CUDA_CHECK(cudaGetLastError());
    if (d.ghost_lo) {
        mc_ghost_count_kernel<<<64, kThreads, 0, stream>>>(w.cube, d, w.totals);
        NPHM_CUDA_CHECK(cudaGetLastError());
    }
    unsigned long long totals[4];
    NPHM_CUDA_CHECK(cudaMemcpyAsync(totals, w.totals, sizeof(totals), cudaMemcpyDeviceToHost, stream));
    NPHM_CUDA_CHECK(cudaStreamSynchronize(stream));
    NPHM_REQUIRE(totals[0] < 0xffffffffull && totals[1] < 0xffffffffull, "nphm_mc_count: mesh too large for 32-bit scratch");
    *n_verts_host = (long long)(totals[0] - (d.ghost_lo ? totals[2] : 0));
    *n_tris_host = (long long)(totals[1] - (d.ghost_lo ? totals[3] : 0));
    return NPHM_OK;
}

extern "C" int nphm_mc_emit(const float *vol_dev, const nphm_mc_params *p, void *workspace_dev,
                            long long vert_id_base, double *verts_dev, long long *tris_dev, void *stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Dims d;
    int rc = make_dims(p, d);
    if (rc) return rc;
    if (d.ncells == 0) return NPHM_OK;
    NPHM_REQUIRE(vol_dev && workspace_dev, "nphm_mc_emit: NULL volume / workspace");
    Workspace w = carve(workspace_dev, d);
    const int nblocks = (int)ceil_div(d.ncells, kCellsPerBlock);
    mc_vertices_kernel<<<nblocks, kThreads, 0, stream>>>(vol_dev, d, w.cube, w.block_v, w.vbase, w.totals, verts_dev);
    NPHM_CUDA_CHECK(cudaGetLastError());
    mc_triangles_kernel<<<nblocks, kThreads, 0, stream>>>(d, w.cube, w.block_t, w.vbase, w.totals, vert_id_base, tris_dev);
    NPHM_CUDA_CHECK(cudaGetLastError());
    return NPHM_OK;
}

extern "C" int nphm_marching_cubes_host(const float *vol_host, int nx, int ny, int nz, double iso, int negate,
                                        double *verts_host, long long *tris_host,
                                        long long *n_verts, long long *n_tris)
{
    NPHM_REQUIRE(vol_host && n_verts && n_tris, "nphm_marching_cubes_host: NULL argument");
    nphm_mc_params p{nx, ny, nz, 0, 0, negate, iso};
    const long long ws_bytes = nphm_mc_workspace_bytes(&p);
    NPHM_REQUIRE(ws_bytes >= 0, "nphm_marching_cubes_host: bad dimensions");
    const size_t vol_bytes = (size_t)nx * ny * nz * sizeof(float);
    float *vol_dev = nullptr;
    void *ws = nullptr;
    double *verts_dev = nullptr;
    long long *tris_dev = nullptr;
    int rc = NPHM_OK;
    auto cleanup = [&]() { cudaFree(vol_dev); cudaFree(ws); cudaFree(verts_dev); cudaFree(tris_dev); };
    if (vol_bytes == 0) { *n_verts = 0; *n_tris = 0; return NPHM_OK; }
    if (cudaMalloc(&vol_dev, vol_bytes) != cudaSuccess || cudaMalloc(&ws, ws_bytes) != cudaSuccess) {
        cleanup(); set_error("nphm_marching_cubes_host: cudaMalloc failed"); return NPHM_ERR_CUDA;
    }
    if (cudaMemcpy(vol_dev, vol_host, vol_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        cleanup(); set_error("nphm_marching_cubes_host: H2D copy failed"); return NPHM_ERR_CUDA;
    }
    rc = nphm_mc_count(vol_dev, &p, ws, n_verts, n_tris, nullptr);
    if (rc == NPHM_OK && verts_host && tris_host && (*n_verts > 0 || *n_tris > 0)) {
        if (cudaMalloc(&verts_dev, (size_t)(*n_verts + 1) * 24) != cudaSuccess ||
            cudaMalloc(&tris_dev, (size_t)(*n_tris + 1) * 24) != cudaSuccess) {
            cleanup(); set_error("nphm_marching_cubes_host: cudaMalloc failed"); return NPHM_ERR_CUDA;
        }
        rc = nphm_mc_emit(vol_dev, &p, ws, 0, verts_dev, tris_dev, nullptr);
        if (rc == NPHM_OK) {
            if (cudaMemcpy(verts_host, verts_dev, (size_t)*n_verts * 24, cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(tris_host, tris_dev, (size_t)*n_tris * 24, cudaMemcpyDeviceToHost) != cudaSuccess) {
                cleanup(); set_error("nphm_marching_cubes_host: D2H copy failed"); return NPHM_ERR_CUDA;
            }
        }
    }
    cleanup();
    return rc;
}
