// fp32 primitive narrowphase for sphere-swept cores (point / segment / box) and upright
// cylinders.  Same semantics as the fp64 oracle (oracle/oracle_scene.c): signed distance
// d = dist(core_a, core_b) - r_a - r_b; intersecting cores give d = -(r_a + r_b) - depth.
// Restates the reference's rai collision query (P/problems/rai_base_env.py:442-477): a pair
// contributes max(0, -d) to the total penetration.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "scene_blob.h"

namespace mrb {

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fmaf(ax, bx, fmaf(ay, by, az * bz));
}
__device__ __forceinline__ float clamp01(float x) { return __saturatef(x); }

// ---- point / segment cores -------------------------------------------------------------
__device__ __forceinline__ float d_point_point(const float* a, const float* b, float rsum) {
    float x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrtf(dot3(x, y, z, x, y, z)) - rsum;
}

__device__ __forceinline__ float d_point_seg(const float* p, const float* s, float rsum) {
    float abx = s[3] - s[0], aby = s[4] - s[1], abz = s[5] - s[2];
    float apx = p[0] - s[0], apy = p[1] - s[1], apz = p[2] - s[2];
    float den = dot3(abx, aby, abz, abx, aby, abz);
    float t = den > 0.f ? clamp01(__fdividef(dot3(apx, apy, apz, abx, aby, abz), den)) : 0.f;
    float x = fmaf(-t, abx, apx), y = fmaf(-t, aby, apy), z = fmaf(-t, abz, apz);
    return sqrtf(dot3(x, y, z, x, y, z)) - rsum;
}

// closest points of two segments (Ericson 5.1.9), branch-light
__device__ __forceinline__ float d_seg_seg(const float* s1, const float* s2, float rsum) {
    float d1x = s1[3] - s1[0], d1y = s1[4] - s1[1], d1z = s1[5] - s1[2];
    float d2x = s2[3] - s2[0], d2y = s2[4] - s2[1], d2z = s2[5] - s2[2];
    float rx = s1[0] - s2[0], ry = s1[1] - s2[1], rz = s1[2] - s2[2];
    float a = dot3(d1x, d1y, d1z, d1x, d1y, d1z);
    float e = dot3(d2x, d2y, d2z, d2x, d2y, d2z);
    float f = dot3(d2x, d2y, d2z, rx, ry, rz);
    float c = dot3(d1x, d1y, d1z, rx, ry, rz);
    float b = dot3(d1x, d1y, d1z, d2x, d2y, d2z);
    const float EPS = 1e-12f;
    float s, t;
    if (a <= EPS && e <= EPS) {
        s = t = 0.f;
    } else if (a <= EPS) {
        s = 0.f;
        t = clamp01(__fdividef(f, e));
    } else if (e <= EPS) {
        t = 0.f;
        s = clamp01(__fdividef(-c, a));
    } else {
        float den = fmaf(a, e, -b * b);
        s = den > 1e-7f * a * e ? clamp01(__fdividef(fmaf(b, f, -c * e), den)) : 0.f;
        t = __fdividef(fmaf(b, s, f), e);
        if (t < 0.f) {
            t = 0.f;
            s = clamp01(__fdividef(-c, a));
        } else if (t > 1.f) {
            t = 1.f;
            s = clamp01(__fdividef(b - c, a));
        }
    }
    float x = fmaf(s, d1x, rx) - t * d2x, y = fmaf(s, d1y, ry) - t * d2y, z = fmaf(s, d1z, rz) - t * d2z;
    return sqrtf(dot3(x, y, z, x, y, z)) - rsum;
}

// ---- box cores: c[3], R[9] row-major (columns = axes), half[3] ---------------------------
__device__ __forceinline__ void to_box_local(const float* c, const float* R, const float* p, float* o) {
    float x = p[0] - c[0], y = p[1] - c[1], z = p[2] - c[2];
    o[0] = dot3(R[0], R[3], R[6], x, y, z);
    o[1] = dot3(R[1], R[4], R[7], x, y, z);
    o[2] = dot3(R[2], R[5], R[8], x, y, z);
}

__device__ __forceinline__ float d_point_box(const float* p, const float* c, const float* R, const float* h, float rsum) {
    float l[3];
    to_box_local(c, R, p, l);
    float d2 = 0.f, inside = 3.0e38f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float ex = fabsf(l[k]) - h[k];
        float m = fmaxf(ex, 0.f);
        d2 = fmaf(m, m, d2);
        inside = fminf(inside, -ex);
    }
    return d2 > 0.f ? sqrtf(d2) - rsum : -inside - rsum;
}

__device__ __forceinline__ void segbox_eval(const float* a, const float* d, const float* h, float t, float& F, float& G) {
    F = 0.f;
    G = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float p = fmaf(t, d[k], a[k]);
        float ex = fabsf(p) - h[k];
        if (ex > 0.f) {
            F = fmaf(ex, ex, F);
            G = fmaf(2.f * ex, p > 0.f ? d[k] : -d[k], G);
        }
    }
}

// exact squared distance between segment a + t d (box-local, t in [0,1]) and box [-h,h]:
// f is convex piecewise quadratic; f' is piecewise linear, bracket its root over the <= 8
// candidate parameters (0, 1 and the slab crossings), then one linear interpolation.
__device__ __forceinline__ float segbox_dist2_local(const float* a, const float* d, const float* h) {
    float flo, glo, fhi, ghi;
    segbox_eval(a, d, h, 0.f, flo, glo);
    if (glo >= 0.f) return flo;
    segbox_eval(a, d, h, 1.f, fhi, ghi);
    if (ghi <= 0.f) return fhi;
    float lo = 0.f, hi = 1.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (d[k] != 0.f) {
            float inv = __fdividef(1.f, d[k]);
#pragma unroll
            for (int sgn = 0; sgn < 2; sgn++) {
                float t = ((sgn ? -h[k] : h[k]) - a[k]) * inv;
                if (t > 0.f && t < 1.f) {
                    float f, g;
                    segbox_eval(a, d, h, t, f, g);
                    if (g < 0.f) {
                        if (t > lo) { lo = t; glo = g; }
                    } else {
                        if (t < hi) { hi = t; ghi = g; }
                    }
                }
            }
        }
    }
    float t = fmaf(hi - lo, __fdividef(-glo, ghi - glo), lo);
    float f, g;
    segbox_eval(a, d, h, t, f, g);
    return f;
}

__device__ __forceinline__ float d_seg_box(const float* seg, const float* c, const float* R, const float* h, float rsum) {
    float a[3], b[3], d[3];
    to_box_local(c, R, seg, a);
    to_box_local(c, R, seg + 3, b);
    d[0] = b[0] - a[0];
    d[1] = b[1] - a[1];
    d[2] = b[2] - a[2];
    float f = segbox_dist2_local(a, d, h);
    return f > 0.f ? sqrtf(f) - rsum : -rsum;
}

// relative pose of box B in the frame of box A: R[i][j] = a_i . b_j, t = RA^T (cB - cA)
__device__ __forceinline__ void box_box_rel(const float* cA, const float* RA, const float* cB, const float* RB,
                                            float (&R)[3][3], float (&t)[3]) {
    const float twx = cB[0] - cA[0], twy = cB[1] - cA[1], twz = cB[2] - cA[2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t[i] = dot3(RA[i], RA[3 + i], RA[6 + i], twx, twy, twz);
#pragma unroll
        for (int j = 0; j < 3; j++) R[i][j] = dot3(RA[i], RA[3 + i], RA[6 + i], RB[j], RB[3 + j], RB[6 + j]);
    }
}

// 15-axis separating-axis test: max over axes of the gap (<= 0: cores overlap)
__device__ __forceinline__ float box_box_sat(const float (&R)[3][3], const float (&t)[3], const float* hA, const float* hB) {
    float AR[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) AR[i][j] = fabsf(R[i][j]);
    float s = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 3; i++)
        s = fmaxf(s, fabsf(t[i]) - (hA[i] + dot3(hB[0], hB[1], hB[2], AR[i][0], AR[i][1], AR[i][2])));
#pragma unroll
    for (int j = 0; j < 3; j++)
        s = fmaxf(s, fabsf(dot3(t[0], t[1], t[2], R[0][j], R[1][j], R[2][j])) -
                         (hB[j] + dot3(hA[0], hA[1], hA[2], AR[0][j], AR[1][j], AR[2][j])));
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            float l2 = fmaf(R[i1][j], R[i1][j], R[i2][j] * R[i2][j]);  // |a_i x b_j|^2 without cancellation
            if (l2 >= (float)MRB_SAT_PARALLEL_EPS2) {
                float ra = fmaf(hA[i1], AR[i2][j], hA[i2] * AR[i1][j]);
                float rb = fmaf(hB[j1], AR[i][j2], hB[j2] * AR[i][j1]);
                float g = (fabsf(fmaf(t[i2], R[i1][j], -t[i1] * R[i2][j])) - (ra + rb)) * rsqrtf(l2);
                s = fmaxf(s, g);
            }
        }
    }
    return s;
}

// one out-of-line copy of the segment-box routine with everything passed in registers
__device__ __noinline__ float segbox_dist2_regs(float a0, float a1, float a2, float d0, float d1, float d2, float h0,
                                                float h1, float h2) {
    const float a[3] = {a0, a1, a2}, d[3] = {d0, d1, d2}, h[3] = {h0, h1, h2};
    return segbox_dist2_local(a, d, h);
}

// min squared distance of the 12 edges of box X (centre c, axes x0 x1 x2 given in the frame of box Y, half
// extents hX) to box Y = [-hY, hY]
__device__ __forceinline__ float box_edges_vs_box(const float* c, const float (&x)[3][3], const float* hX, const float* hY) {
    float best = 3.0e38f;
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
        const int u = (ax + 1) % 3, v = (ax + 2) % 3;
#pragma unroll 1
        for (int sgn = 0; sgn < 4; sgn++) {
            const float su = (sgn & 1) ? hX[u] : -hX[u], sv = (sgn & 2) ? hX[v] : -hX[v];
            float a[3], d[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float base = fmaf(su, x[u][k], fmaf(sv, x[v][k], c[k]));
                a[k] = fmaf(-hX[ax], x[ax][k], base);
                d[k] = 2.f * hX[ax] * x[ax][k];
            }
            best = fminf(best, segbox_dist2_regs(a[0], a[1], a[2], d[0], d[1], d[2], hY[0], hY[1], hY[2]));
        }
    }
    return best;
}

// rare path: rounded boxes whose cores are separated by less than r_a + r_b along every SAT axis.
// The closest features of two disjoint boxes always include an edge (a vertex is the end of one), so the
// distance is the minimum over the 24 edge-versus-box distances.
__device__ __forceinline__ float box_box_exact_dist(const float (&R)[3][3], const float (&t)[3], const float* hA,
                                                    const float* hB) {
    // B's axes in A's frame are the columns of R; A's axes in B's frame are its rows; cA in B's frame = -R^T t
    float xb[3][3], tb[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        tb[j] = -dot3(t[0], t[1], t[2], R[0][j], R[1][j], R[2][j]);
#pragma unroll
        for (int k = 0; k < 3; k++) xb[j][k] = R[k][j];
    }
    const float eb = box_edges_vs_box(t, xb, hB, hA);  // edges of B against A, in A's frame
    const float ea = box_edges_vs_box(tb, R, hA, hB);  // edges of A against B, in B's frame
    return sqrtf(fminf(ea, eb));
}

__device__ __forceinline__ float d_box_box(const float* cA, const float* RA, const float* hA, const float* cB,
                                           const float* RB, const float* hB, float rsum) {
    float R[3][3], t[3];
    box_box_rel(cA, RA, cB, RB, R, t);
    const float s = box_box_sat(R, t, hA, hB);
    if (s <= 0.f || s >= rsum) return s - rsum;
    return box_box_exact_dist(R, t, hA, hB) - rsum;
}

// ---- upright cylinders (z-prisms) --------------------------------------------------------
__device__ __forceinline__ float prism_combine(float s2, float sz) {
    if (s2 > 0.f && sz > 0.f) return sqrtf(fmaf(s2, s2, sz * sz));
    if (s2 > 0.f) return s2;
    if (sz > 0.f) return sz;
    return fmaxf(s2, sz);
}

__device__ __forceinline__ float d_cylz_cylz(const float* a, float ra, float ha, const float* b, float rb, float hb) {
    float dx = a[0] - b[0], dy = a[1] - b[1];
    return prism_combine(sqrtf(fmaf(dx, dx, dy * dy)) - ra - rb, fabsf(a[2] - b[2]) - ha - hb);
}

__device__ __forceinline__ float d_box_cylz(const float* c, const float* R, const float* h, const float* cyl, float r,
                                            float hc) {
    float vx = cyl[0] - c[0], vy = cyl[1] - c[1], vz = cyl[2] - c[2];
    float px = fmaf(R[0], vx, R[3] * vy), py = fmaf(R[1], vx, R[4] * vy);
    float ex = fabsf(px) - h[0], ey = fabsf(py) - h[1], s2;
    if (ex <= 0.f && ey <= 0.f) {
        s2 = fmaxf(ex, ey) - r;
    } else {
        float mx = fmaxf(ex, 0.f), my = fmaxf(ey, 0.f);
        s2 = sqrtf(fmaf(mx, mx, my * my)) - r;
    }
    return prism_combine(s2, fabsf(vz) - h[2] - hc);
}

}  // namespace mrb
