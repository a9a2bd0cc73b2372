/*
 * ORACLE (test infrastructure, not product code): CPU restatement of the
 * marching-cubes call the reference makes at sdf/core.py:16-18,
 *     verts, faces, _, _ = skimage.measure.marching_cubes(volume, 0)
 *     return verts[faces].reshape((-1, 3))
 * i.e. scikit-image's Lewiner marching cubes with its defaults
 * (method='lewiner', step_size=1, allow_degenerate=True,
 * gradient_direction='descent', spacing=(1,1,1)).
 *
 * PARITY UNPINNED: scikit-image (setup.py:14 "scikit-image>=0.17", no lock
 * file) is a third-party dependency that is not under /root/reference, not
 * installed in this image and not in the wheelhouse; the reference holds no
 * tests, goldens or fixtures at this boundary.  What is restated here from the
 * published algorithm (skimage/measure/_marching_cubes_lewiner.py and
 * _marching_cubes_lewiner_cy.pyx; Lewiner et al. 2003):
 *   - volume cast to float32; ValueError if a dim < 2 or level outside
 *     [min,max]; RuntimeError if no vertex is produced   (-> status 1 / 2)
 *   - cube loop: slowest array axis outermost; corner k of Lewiner's cube is
 *     (x,y,z)_L with x_L = FASTEST array axis; index bit k <=> v_k - level > 0
 *   - test_face: sign of A*C - B*D with the +-FLT_EPSILON dead band
 *   - test_interior: Lewiner's interior test (z-sweep for cases 4/10,
 *     reference-edge form for 6/7/12/13)
 *   - vertex on an edge: inverse-|value| weighted mean of the two corners in
 *     double, w = 1/(FLT_EPSILON+|v|), stored as float32; centre vertex: same
 *     over the 8 corners
 *   - output: vertex columns flipped back to array-axis order, face winding
 *     flipped ('descent'), degenerate triangles kept, soup = verts[faces]
 * The tiling tables come from tools/gen_mc_tables.py (see its docstring).
 *
 * Build: make -C oracle   ->  oracle/libmc_oracle.so
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MC_TABLE_QUAL static const
#include "mc_tables.inc"

static const int CORNER_XYZ[8][3] = {
    {0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int EDGE_CORNERS[12][2] = {
    {0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
/* corners A,B,C,D read by test_face for |face| = 1..6 */
static const int FACE_CORNERS[7][4] = {
    {0, 0, 0, 0}, {0, 4, 5, 1}, {1, 5, 6, 2}, {2, 6, 7, 3}, {3, 7, 4, 0}, {0, 3, 2, 1}, {4, 7, 6, 5}};
/* test_interior, reference-edge form: t is taken on edge e = (p,q); the other
 * three parallel edges, in the order B, C, D of the original code */
static const int INTERIOR_EDGE[12][8] = {
    /* e: p q | B: from to | C: from to | D: from to */
    {0, 1, 3, 2, 7, 6, 4, 5}, {1, 2, 0, 3, 4, 7, 5, 6}, {2, 3, 1, 0, 5, 4, 6, 7}, {3, 0, 2, 1, 6, 5, 7, 4},
    {4, 5, 7, 6, 3, 2, 0, 1}, {5, 6, 4, 7, 0, 3, 1, 2}, {6, 7, 5, 4, 1, 0, 2, 3}, {7, 4, 6, 5, 2, 1, 3, 0},
    {0, 4, 3, 7, 2, 6, 1, 5}, {1, 5, 0, 4, 3, 7, 2, 6}, {2, 6, 1, 5, 0, 4, 3, 7}, {3, 7, 2, 6, 1, 5, 0, 4}};

static int test_face(const double v[8], int face) {
    int af = face < 0 ? -face : face;
    double A = v[FACE_CORNERS[af][0]], B = v[FACE_CORNERS[af][1]];
    double C = v[FACE_CORNERS[af][2]], D = v[FACE_CORNERS[af][3]];
    double acbd = A * C - B * D;
    if (acbd > -FLT_EPSILON && acbd < FLT_EPSILON) return face >= 0;
    return (double)face * A * acbd >= 0;
}

/* mode: 12 = cases 4/10 (sweep along Lewiner z), 0..11 = reference edge */
static int test_interior(const double v[8], int mode, int s) {
    double t, At = 0, Bt = 0, Ct = 0, Dt = 0;
    if (mode == 12) {
        double a = (v[4] - v[0]) * (v[6] - v[2]) - (v[7] - v[3]) * (v[5] - v[1]);
        double b = v[2] * (v[4] - v[0]) + v[0] * (v[6] - v[2]) - v[1] * (v[7] - v[3]) - v[3] * (v[5] - v[1]);
        t = -b / (2 * a + FLT_EPSILON);
        if (t < 0 || t > 1) return s > 0;
        At = v[0] + (v[4] - v[0]) * t;
        Bt = v[3] + (v[7] - v[3]) * t;
        Ct = v[2] + (v[6] - v[2]) * t;
        Dt = v[1] + (v[5] - v[1]) * t;
    } else {
        const int *e = INTERIOR_EDGE[mode];
        t = v[e[0]] / (v[e[0]] - v[e[1]] + FLT_EPSILON);
        At = 0;
        Bt = v[e[2]] + (v[e[3]] - v[e[2]]) * t;
        Ct = v[e[4]] + (v[e[5]] - v[e[4]]) * t;
        Dt = v[e[6]] + (v[e[7]] - v[e[6]]) * t;
    }
    int test = 0;
    if (At >= 0) test += 1;
    if (Bt >= 0) test += 2;
    if (Ct >= 0) test += 4;
    if (Dt >= 0) test += 8;
    switch (test) {
        case 0: case 1: case 2: case 3: case 4: case 6: case 8: case 9: case 12:
            return s > 0;
        case 5:
            if (At * Ct - Bt * Dt < FLT_EPSILON) return s > 0;
            break;
        case 10:
            if (At * Ct - Bt * Dt >= FLT_EPSILON) return s > 0;
            break;
        default:
            return s < 0;
    }
    return s < 0;
}

/* tiling id for one cube, or -1 for "emit nothing" */
static int select_tiling(int index, const double v[8]) {
    int nf = MC_NFACE[index];
    int npos = __builtin_popcount((unsigned)index);
    int sgn = npos <= 4 ? 1 : -1;
    int mask = 0;
    for (int j = 0; j < nf; ++j)
        if (test_face(v, sgn * (int)MC_FACES[index * 6 + j])) mask |= 1 << j;
    int sub = MC_SUB_BASE[index] + mask;
    int mode = MC_SUB_INTERIOR[sub];
    if (mode < 0) return MC_SUB_TRUE[sub];
    return test_interior(v, mode, 7 * sgn) ? MC_SUB_TRUE[sub] : MC_SUB_FALSE[sub];
}

static void edge_vertex(const double v[8], int e, int x, int y, int z, float out[3]) {
    double fx = 0, fy = 0, fz = 0, ff = 0;
    if (e == 12) {
        for (int k = 0; k < 8; ++k) {
            double w = 1.0 / (FLT_EPSILON + fabs(v[k]));
            fx += CORNER_XYZ[k][0] * w; fy += CORNER_XYZ[k][1] * w; fz += CORNER_XYZ[k][2] * w;
            ff += w;
        }
    } else {
        for (int s = 0; s < 2; ++s) {
            int k = EDGE_CORNERS[e][s];
            double w = 1.0 / (FLT_EPSILON + fabs(v[k]));
            fx += CORNER_XYZ[k][0] * w; fy += CORNER_XYZ[k][1] * w; fz += CORNER_XYZ[k][2] * w;
            ff += w;
        }
    }
    /* Lewiner (x,y,z); caller stores flipped (z,y,x) = array-axis order */
    out[0] = (float)((double)x + 1.0 * fx / ff);
    out[1] = (float)((double)y + 1.0 * fy / ff);
    out[2] = (float)((double)z + 1.0 * fz / ff);
}

/*
 * vol: float32, C-contiguous, shape (n0, n1, n2).  tris: 9 floats per triangle,
 * vertex coordinates in array-axis order (axis0, axis1, axis2), triangles in
 * skimage's emission order, vertices in verts[faces] order.  If tris == NULL
 * only counts.  Returns the triangle count, or -1 on capacity overflow.
 * *status: 0 ok, 1 = skimage ValueError, 2 = skimage RuntimeError (reference
 * maps both to an empty batch, core.py:53-56).
 */
long mc_lewiner_f32(const float *vol, int n0, int n1, int n2, double level,
                    float *tris, long cap, int *status, long *n_active_cubes) {
    *status = 0;
    if (n_active_cubes) *n_active_cubes = 0;
    if (n0 < 2 || n1 < 2 || n2 < 2) { *status = 1; return 0; }
    size_t n = (size_t)n0 * n1 * n2;
    float mn = vol[0], mx = vol[0];
    for (size_t i = 1; i < n; ++i) { if (vol[i] < mn) mn = vol[i]; if (vol[i] > mx) mx = vol[i]; }
    if (level < mn || level > mx) { *status = 1; return 0; }
    long count = 0, active = 0;
    const int Nx = n2, Ny = n1, Nz = n0;
    for (int z = 0; z < Nz - 1; ++z)
        for (int y = 0; y < Ny - 1; ++y)
            for (int x = 0; x < Nx - 1; ++x) {
                double v[8];
                int index = 0;
                for (int k = 0; k < 8; ++k) {
                    size_t off = ((size_t)(z + CORNER_XYZ[k][2]) * Ny + (y + CORNER_XYZ[k][1])) * Nx + (x + CORNER_XYZ[k][0]);
                    v[k] = (double)vol[off] - level;
                    if (v[k] > 0.0) index |= 1 << k;
                }
                if (index == 0 || index == 255) continue;
                ++active;
                int tile = select_tiling(index, v);
                if (tile < 0) continue;
                int nt = MC_TILE_NTRI[tile];
                const unsigned char *ed = MC_TILE_DATA + MC_TILE_OFF[tile];
                if (tris) {
                    if (count + nt > cap) return -1;
                    for (int t = 0; t < nt; ++t)
                        for (int k = 0; k < 3; ++k) {
                            float p[3];
                            edge_vertex(v, ed[3 * t + k], x, y, z, p);
                            /* faces flipped: (a,b,c) -> (c,b,a); columns flipped: (x,y,z)_L -> (z,y,x) */
                            float *dst = tris + (size_t)(count + t) * 9 + (size_t)(2 - k) * 3;
                            dst[0] = p[2]; dst[1] = p[1]; dst[2] = p[0];
                        }
                }
                count += nt;
            }
    if (n_active_cubes) *n_active_cubes = active;
    if (count == 0) { *status = 2; return 0; }
    return count;
}

/* Expose the selection for table-level tests: tiling id chosen for 8 corner values. */
int mc_select_tiling(const double v[8]) {
    int index = 0;
    for (int k = 0; k < 8; ++k) if (v[k] > 0.0) index |= 1 << k;
    if (index == 0 || index == 255) return -1;
    return select_tiling(index, v);
}
