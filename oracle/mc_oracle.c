/* ORACLE - test infrastructure only.  Nothing under nphm_b200/ may include, link or call this file.
 *
 * Sequential CPU restatement of the marching-cubes pass the reference runs at
 *   /root/reference/src/NPHM/utils/reconstruction.py:30   vertices, triangles = mcubes.marching_cubes(logits, 0.0)
 * `mcubes` is PyMCubes (PyPI `pymcubes`, listed un-pinned in the reference's pyproject.toml:24); its source
 * is NOT part of the reference checkout and the module is not installed in this image, so this restates
 * PyMCubes' published algorithm (mcubes/src/marchingcubes.h, `mc::marching_cubes` +
 * `mc_isovalue_interpolation`) from its documented behaviour:
 *
 *   - cells are visited x-major: for i (x) for j (y) for k (z), z fastest;
 *   - corner m of cell (i,j,k): v0(i,j,k) v1(i+1,j,k) v2(i+1,j+1,k) v3(i,j+1,k), v4..v7 the same at k+1;
 *   - case index bit m is set when v[m] <= isovalue;
 *   - one vertex per crossed grid edge, created by the FIRST cell (in visiting order) that contains the
 *     edge; inside a cell new vertices are appended in the order edge 6, 5, 10 (the three edges meeting at
 *     corner 6, always new) and then 0, 1, 2, 3, 4, 7, 8, 9, 11 (new only on the low boundary planes);
 *   - vertex position along the edge from its first corner a to its second corner b (Bourke's edge order):
 *     xa + (xb - xa) * (iso - fa) / (fb - fa) in double, the midpoint if fa == fb; index units;
 *   - triangles: classic 256-case table, vertex ids in table order.
 *
 * PARITY UNPINNED against PyMCubes itself (no copy of it and no reference test vectors exist here); the
 * bit-exactness claim of the CUDA kernel is against this file.  The case table is validated
 * topologically by tools/gen_mc_tables.py.
 */
#include <stdlib.h>
#include <string.h>
#include "mc_tables_oracle.h"

static const int EDGE_A[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
static const int EDGE_B[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
static const int CORNER[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
static const int CREATE_ORDER[12] = {6, 5, 10, 0, 1, 2, 3, 4, 7, 8, 9, 11};

static double interp(double iso, double fa, double fb, double xa, double xb)
{
    if (fb == fa) return (xb + xa) / 2;
    return (xb - xa) * (iso - fa) / (fb - fa) + xa;
}

/* Two-call protocol: call with verts == NULL / tris == NULL to obtain the counts, then with buffers of
 * 3*n_verts doubles and 3*n_tris uint64.  `negate` != 0 runs on -vol (the reference negates the SDF
 * volume in place before the call, utils/reconstruction.py:25).  Returns 0. */
int mc_oracle(const float *vol, int nx, int ny, int nz, double iso, int negate,
              double *verts, unsigned long long *tris, long long *n_verts, long long *n_tris)
{
    long long nv = 0, nt = 0;
    *n_verts = 0; *n_tris = 0;
    if (nx < 2 || ny < 2 || nz < 2) return 0;
    /* rolling per-axis id slabs, indexed by the grid point at the low end of the edge */
    size_t slab = (size_t)ny * nz;
    long long *idx = (long long *)malloc(sizeof(long long) * 3 * 2 * slab);
    if (!idx) return -1;
#define ID(axis, gi, gj, gk) idx[((size_t)(axis) * 2 + ((gi) & 1)) * slab + (size_t)(gj) * nz + (gk)]
#define VOL(a, b, c) ((double)vol[((size_t)(a) * ny + (b)) * nz + (c)])
    for (int i = 0; i < nx - 1; ++i)
        for (int j = 0; j < ny - 1; ++j)
            for (int k = 0; k < nz - 1; ++k) {
                double v[8];
                for (int m = 0; m < 8; ++m) {
                    double f = VOL(i + CORNER[m][0], j + CORNER[m][1], k + CORNER[m][2]);
                    v[m] = negate ? -f : f;
                }
                unsigned cube = 0;
                for (int m = 0; m < 8; ++m)
                    if (v[m] <= iso) cube |= 1u << m;
                unsigned edges = MC_EDGE_TABLE[cube];
                if (!edges) continue;
                long long id[12];
                for (int o = 0; o < 12; ++o) {
                    int e = CREATE_ORDER[o];
                    if (!(edges & (1u << e))) continue;
                    int a = EDGE_A[e], b = EDGE_B[e];
                    int axis = CORNER[a][0] != CORNER[b][0] ? 0 : (CORNER[a][1] != CORNER[b][1] ? 1 : 2);
                    /* grid point at the low end of the edge */
                    int lo = (CORNER[a][axis] == 0) ? a : b;
                    int gi = i + CORNER[lo][0], gj = j + CORNER[lo][1], gk = k + CORNER[lo][2];
                    /* first cell that contains this grid edge */
                    int fi = i, fj = j, fk = k;
                    if (axis != 0) fi = gi > 0 ? gi - 1 : 0;
                    if (axis != 1) fj = gj > 0 ? gj - 1 : 0;
                    if (axis != 2) fk = gk > 0 ? gk - 1 : 0;
                    if (fi == i && fj == j && fk == k) {
                        if (verts) {
                            double p[3];
                            p[0] = (double)(i + CORNER[a][0]);
                            p[1] = (double)(j + CORNER[a][1]);
                            p[2] = (double)(k + CORNER[a][2]);
                            double q = (double)((axis == 0 ? i : axis == 1 ? j : k) + CORNER[b][axis]);
                            p[axis] = interp(iso, v[a], v[b], p[axis], q);
                            verts[3 * nv + 0] = p[0]; verts[3 * nv + 1] = p[1]; verts[3 * nv + 2] = p[2];
                        }
                        ID(axis, gi, gj, gk) = nv;
                        id[e] = nv++;
                    } else {
                        id[e] = ID(axis, gi, gj, gk);
                    }
                }
                const signed char *row = MC_TRI_TABLE[cube];
                for (int m = 0; row[m] != -1; m += 3) {
                    if (tris) {
                        tris[3 * nt + 0] = (unsigned long long)id[(int)row[m]];
                        tris[3 * nt + 1] = (unsigned long long)id[(int)row[m + 1]];
                        tris[3 * nt + 2] = (unsigned long long)id[(int)row[m + 2]];
                    }
                    ++nt;
                }
            }
#undef ID
#undef VOL
    free(idx);
    *n_verts = nv; *n_tris = nt;
    return 0;
}
