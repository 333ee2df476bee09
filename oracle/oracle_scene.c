/* CPU ORACLE (test infrastructure, NOT product code).
 *
 * fp64 restatement of the reference's rai collision path for primitive scenes.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product package never does.
 *
 * What it restates (paths relative to /root/reference,
 * P/ = src/multi_robot_multi_goal_planning/):
 *   - rai_env.is_collision_free / is_collision_free_np  P/problems/rai_base_env.py:442-513
 *       setJointState (forward kinematics of every frame), getCollisionFree, and when not
 *       free getCollisionsTotalPenetration; collision <=> total penetration > tolerance
 *       (:460-477).  Restated as: free <=> sum over collidable pairs of max(0,-d) <= tol.
 *   - rai_env.is_collision_free_for_robot              P/problems/rai_base_env.py:515-615
 *   - rai_env.is_edge_collision_free                    P/problems/rai_base_env.py:618-676
 *       with generate_binary_search_indices             P/problems/planning_env.py:34-51
 * The arithmetic itself lives in the un-vendored third-party wheel `robotic`
 * (>=0.2.2,<0.3.0, pyproject.toml:24; rai C++: FCL broadphase + GJK/libccd narrowphase).
 * Its published model is restated here analytically: every primitive is a sphere-swept
 * core (point / segment / box) and d = dist(core_a, core_b) - r_a - r_b; when two cores
 * intersect, d = -(r_a + r_b) - depth with depth = exact interior depth (point in box),
 * SAT minimum overlap (box-box) or 0 (segment through box: lower bound).
 *
 * PARITY STATUS: UNPINNED against rai.  The reference holds no golden vector for any rai
 * collision flag and `robotic` cannot be installed here.  What IS pinned:
 * the geometric primitives below against independent numerical minimisation
 * (tests/test_oracle_scene.py), the edge discretisation against the reference's golden
 * vectors (tests/golden/abstract_golden.npz), and FK against the host model.
 *
 * Input is the fp64 twin of the device scene blob (csrc/scene_blob.h): 8-byte words.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../multirobot_pathplanning_benchmark_b200/csrc/scene_blob.h"

#ifdef _OPENMP
#include <omp.h>
#endif

typedef const uint64_t* blob_t;
static inline int64_t BI(blob_t b, int64_t i) { return (int64_t)b[i]; }
static inline double BF(blob_t b, int64_t i) { double d; memcpy(&d, &b[i], 8); return d; }

/* ------------------------------------------------------------------ vector helpers */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ------------------------------------------------------------------ core distances */
/* each returns the signed distance between two sphere-swept shapes (radii rsum) */

static double d_point_point(const double* a, const double* b, double rsum) {
    double v[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    return sqrt(dot3(v, v)) - rsum;
}

static double d_point_seg(const double* p, const double* s, double rsum) {
    const double* a = s; const double* b = s + 3;
    double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    double ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
    double den = dot3(ab, ab);
    double t = den > 0 ? clampd(dot3(ap, ab) / den, 0.0, 1.0) : 0.0;
    double v[3] = {ap[0] - t * ab[0], ap[1] - t * ab[1], ap[2] - t * ab[2]};
    return sqrt(dot3(v, v)) - rsum;
}

/* closest points of two segments (Ericson, Real-Time Collision Detection 5.1.9) */
static double d_seg_seg(const double* s1, const double* s2, double rsum) {
    const double *p1 = s1, *q1 = s1 + 3, *p2 = s2, *q2 = s2 + 3;
    double d1[3] = {q1[0] - p1[0], q1[1] - p1[1], q1[2] - p1[2]};
    double d2[3] = {q2[0] - p2[0], q2[1] - p2[1], q2[2] - p2[2]};
    double r[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
    double s, t;
    const double EPS = 1e-18;
    if (a <= EPS && e <= EPS) { s = t = 0; }
    else if (a <= EPS) { s = 0; t = clampd(f / e, 0, 1); }
    else {
        double c = dot3(d1, r);
        if (e <= EPS) { t = 0; s = clampd(-c / a, 0, 1); }
        else {
            double b = dot3(d1, d2), den = a * e - b * b;
            s = den > 1e-14 * a * e ? clampd((b * f - c * e) / den, 0, 1) : 0.0;
            t = (b * s + f) / e;
            if (t < 0) { t = 0; s = clampd(-c / a, 0, 1); }
            else if (t > 1) { t = 1; s = clampd((b - c) / a, 0, 1); }
        }
    }
    double v[3];
    for (int k = 0; k < 3; k++) v[k] = (p1[k] + s * d1[k]) - (p2[k] + t * d2[k]);
    return sqrt(dot3(v, v)) - rsum;
}

/* box data: c[3], R[9] row-major (columns = box axes in world), half[3] */
static void to_box_local(const double* box, const double* p, double* out) {
    const double* c = box; const double* R = box + 3;
    double v[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    for (int j = 0; j < 3; j++) out[j] = R[0 * 3 + j] * v[0] + R[1 * 3 + j] * v[1] + R[2 * 3 + j] * v[2];
}

static double d_point_box(const double* p, const double* box, double rsum) {
    const double* h = box + 12;
    double l[3]; to_box_local(box, p, l);
    double d2 = 0, inside = 1e300;
    for (int k = 0; k < 3; k++) {
        double ex = fabs(l[k]) - h[k];
        if (ex > 0) d2 += ex * ex;
        if (-ex < inside) inside = -ex;
    }
    if (d2 > 0) return sqrt(d2) - rsum;
    return -inside - rsum; /* centre inside the core: exact interior depth */
}

/* squared distance from local point to the box [-h,h] and its t-derivative along d */
static void segbox_eval(const double* a, const double* d, const double* h, double t, double* f, double* df) {
    double F = 0, G = 0;
    for (int k = 0; k < 3; k++) {
        double p = a[k] + t * d[k];
        double ex = fabs(p) - h[k];
        if (ex > 0) { F += ex * ex; G += 2 * ex * (p > 0 ? d[k] : -d[k]); }
    }
    *f = F; *df = G;
}

/* exact squared distance between segment (local coords a + t d, t in [0,1]) and box [-h,h]:
 * f(t) is convex piecewise quadratic with breakpoints where a coordinate crosses +-h; f' is
 * piecewise linear and non-decreasing, so its root is found by bracketing over the <= 8
 * candidate parameters and one linear interpolation. */
static double segbox_dist2_local(const double* a, const double* d, const double* h) {
    double cand[8]; int n = 0;
    cand[n++] = 0; cand[n++] = 1;
    for (int k = 0; k < 3; k++) {
        if (d[k] != 0) {
            double t1 = (h[k] - a[k]) / d[k], t2 = (-h[k] - a[k]) / d[k];
            if (t1 > 0 && t1 < 1) cand[n++] = t1;
            if (t2 > 0 && t2 < 1) cand[n++] = t2;
        }
    }
    double lo = 0, hi = 1, flo, fhi, glo, ghi;
    segbox_eval(a, d, h, 0, &flo, &glo);
    if (glo >= 0) return flo;
    segbox_eval(a, d, h, 1, &fhi, &ghi);
    if (ghi <= 0) return fhi;
    for (int i = 2; i < n; i++) {
        double f, g; segbox_eval(a, d, h, cand[i], &f, &g);
        if (g < 0) { if (cand[i] > lo) { lo = cand[i]; glo = g; } }
        else { if (cand[i] < hi) { hi = cand[i]; ghi = g; } }
    }
    double t = lo + (hi - lo) * (-glo) / (ghi - glo);
    double f, g; segbox_eval(a, d, h, t, &f, &g);
    return f;
}

static double d_seg_box(const double* seg, const double* box, double rsum) {
    double a[3], b[3], d[3];
    to_box_local(box, seg, a); to_box_local(box, seg + 3, b);
    for (int k = 0; k < 3; k++) d[k] = b[k] - a[k];
    double f = segbox_dist2_local(a, d, box + 12);
    if (f > 0) return sqrt(f) - rsum;
    return -rsum; /* segment touches / crosses the core: penetration >= r_a + r_b (lower bound) */
}

/* world corner / edge enumeration for the rare exact box-box distance */
static void box_edge(const double* box, int e, double* seg) {
    /* edge e: axis = e / 4, the two other coordinates take signs from bits of e % 4 */
    const double* c = box; const double* R = box + 3; const double* h = box + 12;
    int ax = e / 4, u = (ax + 1) % 3, v = (ax + 2) % 3;
    double su = (e & 1) ? 1 : -1, sv = (e & 2) ? 1 : -1;
    for (int k = 0; k < 3; k++) {
        double base = c[k] + su * h[u] * R[k * 3 + u] + sv * h[v] * R[k * 3 + v];
        seg[k] = base - h[ax] * R[k * 3 + ax];
        seg[3 + k] = base + h[ax] * R[k * 3 + ax];
    }
}

static double box_box_exact_dist(const double* A, const double* B) {
    double best = 1e300, seg[6];
    for (int e = 0; e < 12; e++) {
        box_edge(A, e, seg);
        double a[3], b[3], d[3];
        to_box_local(B, seg, a); to_box_local(B, seg + 3, b);
        for (int k = 0; k < 3; k++) d[k] = b[k] - a[k];
        double f = segbox_dist2_local(a, d, B + 12);
        if (f < best) best = f;
        box_edge(B, e, seg);
        to_box_local(A, seg, a); to_box_local(A, seg + 3, b);
        for (int k = 0; k < 3; k++) d[k] = b[k] - a[k];
        f = segbox_dist2_local(a, d, A + 12);
        if (f < best) best = f;
    }
    return sqrt(best);
}

/* separating-axis test, 15 axes (Gottschalk et al.): returns max over axes of the gap
 * (>0: separated by at least that much; <=0: cores overlap, -value = minimum overlap) */
static double box_box_sat(const double* A, const double* B) {
    const double *cA = A, *RA = A + 3, *hA = A + 12, *cB = B, *RB = B + 3, *hB = B + 12;
    double R[3][3], AR[3][3], t[3], tw[3] = {cB[0] - cA[0], cB[1] - cA[1], cB[2] - cA[2]};
    for (int i = 0; i < 3; i++) {
        t[i] = RA[0 * 3 + i] * tw[0] + RA[1 * 3 + i] * tw[1] + RA[2 * 3 + i] * tw[2];
        for (int j = 0; j < 3; j++) {
            R[i][j] = RA[0 * 3 + i] * RB[0 * 3 + j] + RA[1 * 3 + i] * RB[1 * 3 + j] + RA[2 * 3 + i] * RB[2 * 3 + j];
            AR[i][j] = fabs(R[i][j]);
        }
    }
    double s = -1e300, g;
    for (int i = 0; i < 3; i++) {
        g = fabs(t[i]) - (hA[i] + hB[0] * AR[i][0] + hB[1] * AR[i][1] + hB[2] * AR[i][2]);
        if (g > s) s = g;
    }
    for (int j = 0; j < 3; j++) {
        g = fabs(t[0] * R[0][j] + t[1] * R[1][j] + t[2] * R[2][j]) - (hB[j] + hA[0] * AR[0][j] + hA[1] * AR[1][j] + hA[2] * AR[2][j]);
        if (g > s) s = g;
    }
    for (int i = 0; i < 3; i++) {
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        for (int j = 0; j < 3; j++) {
            int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            double l2 = R[i1][j] * R[i1][j] + R[i2][j] * R[i2][j]; /* |a_i x b_j|^2 */
            if (l2 < MRB_SAT_PARALLEL_EPS2) continue;
            double ra = hA[i1] * AR[i2][j] + hA[i2] * AR[i1][j];
            double rb = hB[j1] * AR[i][j2] + hB[j2] * AR[i][j1];
            g = (fabs(t[i2] * R[i1][j] - t[i1] * R[i2][j]) - (ra + rb)) / sqrt(l2);
            if (g > s) s = g;
        }
    }
    return s;
}

/* z-prisms (disc x interval, rectangle x interval; axes parallel to world z): exact signed
 * distance from the planar signed distance s2 and the signed z gap sz */
static double prism_combine(double s2, double sz) {
    if (s2 > 0 && sz > 0) return sqrt(s2 * s2 + sz * sz);
    if (s2 > 0) return s2;
    if (sz > 0) return sz;
    return s2 > sz ? s2 : sz; /* both overlap: penetration depth = smaller of the two */
}

static double d_cylz_cylz(const double* a, const double* b) {
    double dx = a[0] - b[0], dy = a[1] - b[1];
    double s2 = sqrt(dx * dx + dy * dy) - a[3] - b[3];
    double sz = fabs(a[2] - b[2]) - a[4] - b[4];
    return prism_combine(s2, sz);
}

static double d_box_cylz(const double* box, const double* cyl) {
    const double* c = box; const double* R = box + 3; const double* h = box + 12;
    double v[3] = {cyl[0] - c[0], cyl[1] - c[1], cyl[2] - c[2]};
    double px = R[0] * v[0] + R[3] * v[1], py = R[1] * v[0] + R[4] * v[1]; /* box z axis = world z */
    double ex = fabs(px) - h[0], ey = fabs(py) - h[1], s2;
    if (ex <= 0 && ey <= 0) s2 = (ex > ey ? ex : ey) - cyl[3];
    else { double mx = ex > 0 ? ex : 0, my = ey > 0 ? ey : 0; s2 = sqrt(mx * mx + my * my) - cyl[3]; }
    double sz = fabs(v[2]) - h[2] - cyl[4];
    return prism_combine(s2, sz);
}

static double d_box_box(const double* A, const double* B, double rsum) {
    double s = box_box_sat(A, B);
    if (s <= 0 || s >= rsum) return s - rsum;
    return box_box_exact_dist(A, B) - rsum; /* rounded boxes closer than r_a + r_b along every axis */
}

/* ------------------------------------------------------------------ FK */
typedef struct { double R[9]; double t[3]; } xf_t;

static void xf_mul(const xf_t* X, const double* AR, const double* At, xf_t* out) {
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            out->R[i * 3 + j] = X->R[i * 3 + 0] * AR[0 * 3 + j] + X->R[i * 3 + 1] * AR[1 * 3 + j] + X->R[i * 3 + 2] * AR[2 * 3 + j];
        out->t[i] = X->R[i * 3 + 0] * At[0] + X->R[i * 3 + 1] * At[1] + X->R[i * 3 + 2] * At[2] + X->t[i];
    }
}

static void joint_xf(int code, const double* q, double* R, double* t) {
    for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    t[0] = t[1] = t[2] = 0;
    double c, s;
    switch (code) {
        case MRB_J_HINGE_X: c = cos(q[0]); s = sin(q[0]); R[4] = c; R[5] = -s; R[7] = s; R[8] = c; break;
        case MRB_J_HINGE_Y: c = cos(q[0]); s = sin(q[0]); R[0] = c; R[2] = s; R[6] = -s; R[8] = c; break;
        case MRB_J_HINGE_Z: c = cos(q[0]); s = sin(q[0]); R[0] = c; R[1] = -s; R[3] = s; R[4] = c; break;
        case MRB_J_TRANS_XY_PHI: c = cos(q[2]); s = sin(q[2]); R[0] = c; R[1] = -s; R[3] = s; R[4] = c; t[0] = q[0]; t[1] = q[1]; break;
        case MRB_J_TRANS_X: t[0] = q[0]; break;
        case MRB_J_TRANS_Y: t[1] = q[0]; break;
        case MRB_J_TRANS_Z: t[2] = q[0]; break;
        default: break;
    }
}

#define MAX_FRAMES 128
#define MAX_SHAPES 512

/* world data of every shape (moving: from FK at q; static: from the blob): 16 doubles each */
static void world_shapes(blob_t b, const double* q, double* W /* [n_shapes*16] */) {
    int64_t nf = BI(b, MRB_H_NFRAMES), nmov = BI(b, MRB_H_NMOV), nsta = BI(b, MRB_H_NSTA);
    int64_t offF = BI(b, MRB_H_OFF_FRAMES), offS = BI(b, MRB_H_OFF_SHAPES);
    xf_t X[MAX_FRAMES];
    for (int64_t f = 0; f < nf; f++) {
        int64_t base = offF + f * MRB_FRAME_WORDS;
        int64_t par = BI(b, base), code = BI(b, base + 1), qi = BI(b, base + 2);
        double AR[9], At[3], JR[9], Jt[3];
        for (int k = 0; k < 9; k++) AR[k] = BF(b, base + 4 + k);
        for (int k = 0; k < 3; k++) At[k] = BF(b, base + 13 + k);
        xf_t P, T;
        if (par < 0) { for (int k = 0; k < 9; k++) P.R[k] = (k % 4 == 0); P.t[0] = P.t[1] = P.t[2] = 0; }
        else P = X[par];
        xf_mul(&P, AR, At, &T);
        joint_xf((int)code, q + qi, JR, Jt);
        xf_mul(&T, JR, Jt, &X[f]);
    }
    for (int64_t s = 0; s < nmov + nsta; s++) {
        int64_t base = offS + s * MRB_SHAPE_WORDS;
        int64_t core = BI(b, base), fr = BI(b, base + 1);
        double* w = W + s * 16;
        double L[16];
        for (int k = 0; k < 16; k++) L[k] = BF(b, base + 4 + k);
        if (fr < 0) { memcpy(w, L, sizeof(L)); continue; }
        const xf_t* T = &X[fr];
        int npts = core == MRB_CORE_SEG ? 2 : 1;
        for (int p = 0; p < npts; p++)
            for (int i = 0; i < 3; i++)
                w[p * 3 + i] = T->R[i * 3] * L[p * 3] + T->R[i * 3 + 1] * L[p * 3 + 1] + T->R[i * 3 + 2] * L[p * 3 + 2] + T->t[i];
        if (core == MRB_CORE_CYLZ) { w[3] = L[3]; w[4] = L[4]; }
        if (core == MRB_CORE_BOX) {
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                    w[3 + i * 3 + j] = T->R[i * 3] * L[3 + j] + T->R[i * 3 + 1] * L[3 + 3 + j] + T->R[i * 3 + 2] * L[3 + 6 + j];
            for (int k = 0; k < 3; k++) w[12 + k] = L[12 + k];
        }
    }
}

static double pair_distance(int type, const double* wa, const double* wb, double rsum) {
    switch (type) {
        case MRB_PT_POINT_POINT: return d_point_point(wa, wb, rsum);
        case MRB_PT_POINT_SEG: return d_point_seg(wa, wb, rsum);
        case MRB_PT_SEG_SEG: return d_seg_seg(wa, wb, rsum);
        case MRB_PT_POINT_BOX: return d_point_box(wa, wb, rsum);
        case MRB_PT_SEG_BOX: return d_seg_box(wa, wb, rsum);
        case MRB_PT_CYLZ_CYLZ: return d_cylz_cylz(wa, wb);
        case MRB_PT_BOX_CYLZ: return d_box_cylz(wa, wb);
        default: return d_box_box(wa, wb, rsum);
    }
}

static double shape_radius(blob_t b, int64_t s) { return BF(b, BI(b, MRB_H_OFF_SHAPES) + s * MRB_SHAPE_WORDS + 3); }

/* constant contribution of static-static pairs */
double orc_static_penetration(blob_t b) {
    int64_t n = BI(b, MRB_H_N_STATIC_PAIRS), off = BI(b, MRB_H_OFF_STATIC_PAIRS);
    int64_t ns = BI(b, MRB_H_NMOV) + BI(b, MRB_H_NSTA);
    double* W = (double*)malloc(sizeof(double) * 16 * ns);
    double q0[64] = {0};
    world_shapes(b, q0, W);
    double pen = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t t = BI(b, off + 3 * i), a = BI(b, off + 3 * i + 1), c = BI(b, off + 3 * i + 2);
        double d = pair_distance((int)t, W + a * 16, W + c * 16, shape_radius(b, a) + shape_radius(b, c));
        if (d < 0) pen -= d;
    }
    free(W);
    return pen;
}

/* one configuration: total penetration over ALL collidable pairs, minimum pair distance, and
 * (A6, is_collision_free_for_robot, rai_base_env.py:556-578) whether some penetrating pair
 * (d < 0) involves a "relevant" shape (a shape of one of the queried robots, or one of their
 * active task's frames) and no shape of another robot.  rel/oth: per-shape 0/1, nullable. */
static void eval_config(blob_t b, const double* q, double static_pen, const uint8_t* rel, const uint8_t* oth,
                        double* pen_out, double* mind_out, int* relpen_out, double* W) {
    world_shapes(b, q, W);
    double pen = static_pen, mind = 1e300;
    int relpen = 0;
    for (int t = 0; t < MRB_NUM_PAIR_TYPES; t++) {
        int64_t n = BI(b, MRB_H_N_PAIRS + t), off = BI(b, MRB_H_OFF_PAIRS + t);
        for (int64_t i = 0; i < n; i++) {
            int64_t pk = BI(b, off + i), a = pk & 0xffff, c = (pk >> 16) & 0xfff;
            double d = pair_distance(t, W + a * 16, W + c * 16, shape_radius(b, a) + shape_radius(b, c));
            if (d < 0) {
                pen -= d;
                if (rel && (rel[a] | rel[c]) && !(oth[a] | oth[c])) relpen = 1;
            }
            if (d < mind) mind = d;
        }
    }
    *pen_out = pen; *mind_out = mind; *relpen_out = relpen;
}

/* flags[i] = 1 iff configuration i is collision free:  total penetration <= tol
 * (with rel/oth given: the A6 rule  free <=> !(total penetration > tol && relevant pair penetrates)).
 * pen / mind (nullable) receive total penetration and min pair distance for margin tests.
 * tol < 0 -> use the blob's tolerance.  nthreads <= 1 -> scalar. */
int orc_check_configs(blob_t b, const double* q, int64_t B, double tol, const uint8_t* rel, const uint8_t* oth,
                      uint8_t* flags, double* pen, double* mind, int nthreads) {
    if (BI(b, MRB_H_MAGIC) != MRB_BLOB_MAGIC || BI(b, MRB_H_VERSION) != MRB_BLOB_VERSION) return -1;
    int64_t D = BI(b, MRB_H_DOF), ns = BI(b, MRB_H_NMOV) + BI(b, MRB_H_NSTA);
    if (ns > MAX_SHAPES || BI(b, MRB_H_NFRAMES) > MAX_FRAMES) return -2;
    if (tol < 0) tol = BF(b, MRB_H_TOL);
    double sp = orc_static_penetration(b);
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 1 ? nthreads : 1)
#endif
    {
        double* W = (double*)malloc(sizeof(double) * 16 * ns);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < B; i++) {
            double p, m; int rp;
            eval_config(b, q + i * D, sp, rel, oth, &p, &m, &rp, W);
            flags[i] = rel ? !(p > tol && rp) : !(p > tol);
            if (pen) pen[i] = p;
            if (mind) mind[i] = m;
        }
        free(W);
    }
    return 0;
}

/* p-th element of generate_binary_search_indices(N) by BFS, cached per N by the caller */
static void binary_indices(int N, int* seq) {
    int* qs = (int*)malloc(sizeof(int) * 2 * (N + 1));
    int head = 0, tail = 0, k = 0;
    qs[0] = 0; qs[1] = N - 1; tail = 1;
    while (head < tail) {
        int s = qs[2 * head], e = qs[2 * head + 1]; head++;
        int mid = (s + e) / 2;
        seq[k++] = mid;
        if (s <= mid - 1) { qs[2 * tail] = s; qs[2 * tail + 1] = mid - 1; tail++; }
        if (mid + 1 <= e) { qs[2 * tail] = mid + 1; qs[2 * tail + 1] = e; tail++; }
    }
    free(qs);
}

/* edges: sequential, early exit, exactly the reference's loop (rai_base_env.py:636-676).
 * q1,q2: [E,D] fp64.  Ns nullable (then N = max(2, int(|dq|_inf / resolution) + 1)).
 * first_pos[e] = position in binary order of the first colliding sample, -1 if free.
 * checks[e] (nullable) = number of configuration checks performed. */
int orc_check_edges(blob_t b, const double* q1, const double* q2, int64_t E, double resolution, const int32_t* Ns,
                    int n_start, int n_max, int include_endpoints, double tol, uint8_t* flags, int32_t* first_pos,
                    int32_t* checks, int nthreads) {
    if (BI(b, MRB_H_MAGIC) != MRB_BLOB_MAGIC) return -1;
    int64_t D = BI(b, MRB_H_DOF), ns = BI(b, MRB_H_NMOV) + BI(b, MRB_H_NSTA);
    if (tol < 0) tol = BF(b, MRB_H_TOL);
    double sp = orc_static_penetration(b);
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 1 ? nthreads : 1)
#endif
    {
        double* W = (double*)malloc(sizeof(double) * 16 * ns);
        double* q = (double*)malloc(sizeof(double) * D);
        double* dir = (double*)malloc(sizeof(double) * D);
        int* seq = NULL; int seqN = -1;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int64_t e = 0; e < E; e++) {
            const double *a = q1 + e * D, *c = q2 + e * D;
            int N;
            if (Ns) N = Ns[e];
            else {
                double m = 0;
                for (int64_t k = 0; k < D; k++) { double v = fabs(a[k] - c[k]); if (v > m) m = v; }
                N = (int)(m / resolution) + 1;
                if (N < 2) N = 2;
            }
            int nmax = (n_max < 0 || n_max > N) ? N : n_max;
            if (N != seqN) { free(seq); seq = (int*)malloc(sizeof(int) * N); binary_indices(N, seq); seqN = N; }
            for (int64_t k = 0; k < D; k++) dir[k] = (c[k] - a[k]) / (N - 1);
            int fp = -1, cnt = 0;
            for (int p = n_start; p < nmax; p++) {
                int i = seq[p];
                if (!include_endpoints && (i == 0 || i == N - 1)) continue;
                for (int64_t k = 0; k < D; k++) q[k] = a[k] + dir[k] * i;
                double pen, mind; int rp;
                eval_config(b, q, sp, NULL, NULL, &pen, &mind, &rp, W);
                cnt++;
                if (pen > tol) { fp = p; break; }
            }
            flags[e] = fp < 0;
            if (first_pos) first_pos[e] = fp;
            if (checks) checks[e] = cnt;
        }
        free(W); free(q); free(dir); free(seq);
    }
    return 0;
}

/* test hooks */
void orc_world_shapes(blob_t b, const double* q, double* W) { world_shapes(b, q, W); }
double orc_pair_distance(int type, const double* wa, const double* wb, double rsum) { return pair_distance(type, wa, wb, rsum); }
double orc_segbox_dist2_local(const double* a, const double* d, const double* h) { return segbox_dist2_local(a, d, h); }
double orc_box_box_sat(const double* A, const double* B) { return box_box_sat(A, B); }
double orc_box_box_exact_dist(const double* A, const double* B) { return box_box_exact_dist(A, B); }
void orc_binary_indices(int N, int* seq) { binary_indices(N, seq); }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
