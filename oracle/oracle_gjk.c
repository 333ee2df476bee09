/* SECOND CPU ORACLE (test infrastructure, NOT product code): generic convex narrowphase.
 *
 * The first oracle (oracle_scene.c) restates rai's collision arithmetic with ANALYTIC primitive routines and a few
 * conventions of its own (cylinders in general position as capsules, a segment passing through a box clamped at
 * -(r_a + r_b)).  rai itself (the un-vendored `robotic` wheel; P/problems/rai_base_env.py:211-213 names its
 * `pairCollision` machinery) runs GJK on convex cores and an expanding-polytope step for penetrating ones (libccd), with
 * sphere-swept shapes handled as core + radius and cylinders as convex bodies.  This file does THAT, independently of
 * the analytic routines: Gilbert-Johnson-Keerthi distance between support-mapped convex cores (point, segment, box, true
 * cylinder), Expanding Polytope Algorithm for the penetration depth of intersecting cores, then
 *      d = dist(core_a, core_b) - r_a - r_b        (separated cores)
 *      d = -(depth(core_a, core_b) + r_a + r_b)    (intersecting cores)
 * and the same flag rule (free <=> sum of max(0, -d) <= tol).  It shares forward kinematics and the collidable pair
 * lists with oracle_scene.c (geometry is pinned separately against the reference's own model files, tests/test_gfile.py);
 * what it adds is a bound on what the analytic conventions can cost: tests/test_oracle_gjk.py reports, per scene, how many
 * flags of the two oracles differ and the margins of those samples.
 *
 * Still no substitute for rai's own answers (scripts/dump_rai_flags.py): parity stays UNPINNED until those exist.
 */
#include "oracle_scene.c"

typedef struct {
    int kind;          /* 0 point, 1 segment, 2 box, 3 cylinder (true convex cylinder) */
    double a[3], b[3]; /* point: a; segment / cylinder axis: a -> b; box: a = centre */
    double R[9], h[3]; /* box axes (columns of R) and half extents */
    double rc;         /* cylinder radius */
    double c[3];       /* some interior point */
} gshape_t;

static void support(const gshape_t* s, const double* d, double* out) {
    if (s->kind == 0) { memcpy(out, s->a, 24); return; }
    if (s->kind == 1 || s->kind == 3) {
        double ab[3] = {s->b[0] - s->a[0], s->b[1] - s->a[1], s->b[2] - s->a[2]};
        const double* e = dot3(d, ab) > 0 ? s->b : s->a;
        memcpy(out, e, 24);
        if (s->kind == 3) {   /* rim point: radius times the unit radial part of d (none if d is parallel to the axis) */
            double l = sqrt(dot3(ab, ab)), u[3] = {0, 0, 1};
            if (l > 0) for (int k = 0; k < 3; k++) u[k] = ab[k] / l;
            double dp[3], t = dot3(d, u);
            for (int k = 0; k < 3; k++) dp[k] = d[k] - t * u[k];
            double n = sqrt(dot3(dp, dp));
            if (n > 1e-9 * sqrt(dot3(d, d))) {
                for (int k = 0; k < 3; k++) dp[k] /= n;
                t = dot3(dp, u);                      /* rounding residue along the axis must not leak into the rim point */
                for (int k = 0; k < 3; k++) dp[k] -= t * u[k];
                for (int k = 0; k < 3; k++) out[k] += s->rc * dp[k];
            }
        }
        return;
    }
    memcpy(out, s->a, 24);
    for (int j = 0; j < 3; j++) {
        double ax[3] = {s->R[0 * 3 + j], s->R[1 * 3 + j], s->R[2 * 3 + j]};
        double sg = dot3(d, ax) >= 0 ? s->h[j] : -s->h[j];
        for (int k = 0; k < 3; k++) out[k] += sg * ax[k];
    }
}

static void msupport(const gshape_t* A, const gshape_t* B, const double* d, double* w) {
    double pa[3], pb[3], nd[3] = {-d[0], -d[1], -d[2]};
    support(A, d, pa);
    support(B, nd, pb);
    for (int k = 0; k < 3; k++) w[k] = pa[k] - pb[k];
}

static inline void sub3(const double* a, const double* b, double* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

/* closest point to the origin on segment / triangle (Ericson 5.1.2, 5.1.5); returns the sub-simplex that carries it */
static int closest_seg(const double* A, const double* B, double* v, int* keep) {
    double ab[3]; sub3(B, A, ab);
    double t = -dot3(A, ab), den = dot3(ab, ab);
    if (t <= 0 || den <= 0) { memcpy(v, A, 24); keep[0] = 0; return 1; }
    if (t >= den) { memcpy(v, B, 24); keep[0] = 1; return 1; }
    t /= den;
    for (int k = 0; k < 3; k++) v[k] = A[k] + t * ab[k];
    keep[0] = 0; keep[1] = 1;
    return 2;
}

static int closest_tri(const double* A, const double* B, const double* C, double* v, int* keep) {
    double ab[3], ac[3], ap[3] = {-A[0], -A[1], -A[2]};
    sub3(B, A, ab); sub3(C, A, ac);
    {   /* (nearly) collinear points -- support points along one edge of a box: the best of the three edges */
        double nn[3]; cross3(ab, ac, nn);
        if (dot3(nn, nn) <= 1e-22 * dot3(ab, ab) * dot3(ac, ac)) {
            const double* P[3] = {A, B, C};
            static const int E[3][2] = {{0, 1}, {0, 2}, {1, 2}};
            double best = 1e300;
            int bm = 1;
            for (int e = 0; e < 3; e++) {
                double tv[3]; int tk[2];
                int tm = closest_seg(P[E[e][0]], P[E[e][1]], tv, tk);
                double d2 = dot3(tv, tv);
                if (d2 < best) { best = d2; memcpy(v, tv, 24); bm = tm; for (int k = 0; k < tm; k++) keep[k] = E[e][tk[k]]; }
            }
            return bm;
        }
    }
    double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0 && d2 <= 0) { memcpy(v, A, 24); keep[0] = 0; return 1; }
    double bp[3] = {-B[0], -B[1], -B[2]};
    double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0 && d4 <= d3) { memcpy(v, B, 24); keep[0] = 1; return 1; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        double t = d1 / (d1 - d3);
        for (int k = 0; k < 3; k++) v[k] = A[k] + t * ab[k];
        keep[0] = 0; keep[1] = 1; return 2;
    }
    double cp[3] = {-C[0], -C[1], -C[2]};
    double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0 && d5 <= d6) { memcpy(v, C, 24); keep[0] = 2; return 1; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        double t = d2 / (d2 - d6);
        for (int k = 0; k < 3; k++) v[k] = A[k] + t * ac[k];
        keep[0] = 0; keep[1] = 2; return 2;
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int k = 0; k < 3; k++) v[k] = B[k] + t * (C[k] - B[k]);
        keep[0] = 1; keep[1] = 2; return 2;
    }
    double den = 1.0 / (va + vb + vc), s = vb * den, t = vc * den;
    for (int k = 0; k < 3; k++) v[k] = A[k] + ab[k] * s + ac[k] * t;
    keep[0] = 0; keep[1] = 1; keep[2] = 2;
    return 3;
}

/* simplex S (n points, newest last): closest point to the origin, reduced to its carrier; returns new n, or -1 if the
 * origin lies inside a tetrahedron */
static int closest_simplex(double S[4][3], int n, double* v) {
    int keep[3], m;
    if (n == 1) { memcpy(v, S[0], 24); return 1; }
    if (n == 2) { m = closest_seg(S[0], S[1], v, keep); }
    else if (n == 3) { m = closest_tri(S[0], S[1], S[2], v, keep); }
    else {
        /* tetrahedron: origin outside a face plane -> closest point on that face; keep the best */
        static const int F[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
        static const int O[4] = {3, 2, 1, 0};
        double best = 1e300, bv[3] = {0, 0, 0};
        int bk[3] = {0, 0, 0}, bm = 0, bf = -1, outside_any = 0;
        for (int f = 0; f < 4; f++) {
            const double *A = S[F[f][0]], *B = S[F[f][1]], *C = S[F[f][2]], *Dp = S[O[f]];
            double ab[3], ac[3], nrm[3], ad[3];
            sub3(B, A, ab); sub3(C, A, ac); cross3(ab, ac, nrm); sub3(Dp, A, ad);
            double sd = dot3(nrm, ad), so = -dot3(nrm, A);
            /* relative tolerance: flat simplices (support points of parallel faces are coplanar) must not be read as
             * "origin inside" from the sign of a rounding error */
            double tolf = 1e-11 * sqrt(dot3(nrm, nrm)) * (sqrt(dot3(ad, ad)) + sqrt(dot3(A, A)) + 1e-300);
            if (fabs(sd) > tolf && fabs(so) > tolf && so * sd > 0) continue;   /* origin strictly on the inner side of this face */
            outside_any = 1;
            double tv[3]; int tk[3];
            int tm = closest_tri(A, B, C, tv, tk);
            double d2 = dot3(tv, tv);
            if (d2 < best) { best = d2; memcpy(bv, tv, 24); bm = tm; bf = f; for (int k = 0; k < tm; k++) bk[k] = tk[k]; }
        }
        if (!outside_any || bf < 0) return -1;
        memcpy(v, bv, 24);
        double T[3][3];
        for (int k = 0; k < bm; k++) memcpy(T[k], S[F[bf][bk[k]]], 24);
        for (int k = 0; k < bm; k++) memcpy(S[k], T[k], 24);
        return bm;
    }
    double T[3][3];
    for (int k = 0; k < m; k++) memcpy(T[k], S[keep[k]], 24);
    for (int k = 0; k < m; k++) memcpy(S[k], T[k], 24);
    return m;
}

/* ---- EPA ---- */
#define EPA_MAXV 160
#define EPA_MAXF 320
typedef struct { int v[3]; double n[3], d; int alive; } eface_t;

static int epa_add_face(eface_t* F, int* nf, double V[][3], int a, int b, int c, const double* inside) {
    if (*nf >= EPA_MAXF) return -1;
    eface_t* f = &F[*nf];
    double ab[3], ac[3];
    sub3(V[b], V[a], ab); sub3(V[c], V[a], ac); cross3(ab, ac, f->n);
    double l = sqrt(dot3(f->n, f->n));
    if (l < 1e-300) return -1;
    for (int k = 0; k < 3; k++) f->n[k] /= l;
    f->d = dot3(f->n, V[a]);
    f->v[0] = a; f->v[1] = b; f->v[2] = c;
    {   /* outward = away from a point known to lie inside the polytope (the sign of d is useless when the origin is on the face) */
        double t[3]; sub3(inside, V[a], t);
        if (dot3(f->n, t) > 0) { f->d = -f->d; for (int k = 0; k < 3; k++) f->n[k] = -f->n[k]; f->v[1] = c; f->v[2] = b; }
    }
    if (f->d < 0) f->d = 0;
    f->alive = 1;
    return (*nf)++;
}

/* penetration depth of the origin inside A - B, starting from a tetrahedron S that contains it */
static double epa_depth(const gshape_t* A, const gshape_t* B, double S[4][3]) {
    static const int F0[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
    double V[EPA_MAXV][3];
    eface_t F[EPA_MAXF];
    int nv = 4, nf = 0;
    double inside[3] = {0, 0, 0};
    for (int k = 0; k < 4; k++) { memcpy(V[k], S[k], 24); for (int j = 0; j < 3; j++) inside[j] += 0.25 * S[k][j]; }
    for (int f = 0; f < 4; f++) if (epa_add_face(F, &nf, V, F0[f][0], F0[f][1], F0[f][2], inside) < 0) return 0.0;
    double best_depth = 0.0;
    for (int it = 0; it < 120; it++) {
        int bf = -1;
        for (int f = 0; f < nf; f++) if (F[f].alive && (bf < 0 || F[f].d < F[bf].d)) bf = f;
        if (bf < 0) break;
        double w[3];
        msupport(A, B, F[bf].n, w);
        double dw = dot3(w, F[bf].n);
        best_depth = F[bf].d;
        if (dw - F[bf].d < 1e-10 || nv >= EPA_MAXV) return dw > F[bf].d ? 0.5 * (dw + F[bf].d) : F[bf].d;
        int wi = nv++;
        memcpy(V[wi], w, 24);
        /* remove the faces that see w, collect the horizon */
        int E[EPA_MAXF * 3][2], ne = 0;
        for (int f = 0; f < nf; f++) {
            if (!F[f].alive) continue;
            double t[3]; sub3(w, V[F[f].v[0]], t);
            if (dot3(F[f].n, t) > 1e-14) {
                F[f].alive = 0;
                for (int e = 0; e < 3; e++) {
                    int a = F[f].v[e], b = F[f].v[(e + 1) % 3], found = -1;
                    for (int j = 0; j < ne; j++) if (E[j][0] == b && E[j][1] == a) { found = j; break; }
                    if (found >= 0) { E[found][0] = E[ne - 1][0]; E[found][1] = E[ne - 1][1]; ne--; }
                    else { E[ne][0] = a; E[ne][1] = b; ne++; }
                }
            }
        }
        if (ne == 0) return best_depth;
        for (int j = 0; j < ne; j++) if (epa_add_face(F, &nf, V, E[j][0], E[j][1], wi, inside) == -1 && nf >= EPA_MAXF) return best_depth;
    }
    return best_depth;
}

/* grow a degenerate GJK simplex (n < 4 points, origin on it) into a tetrahedron around the origin */
static int blow_up(const gshape_t* A, const gshape_t* B, double S[4][3], int n) {
    static const double AX[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    if (n == 1) {
        for (int k = 0; k < 6 && n < 2; k++) {
            double w[3], t[3];
            msupport(A, B, AX[k], w); sub3(w, S[0], t);
            if (dot3(t, t) > 1e-20) { memcpy(S[1], w, 24); n = 2; }
        }
        if (n < 2) return 0;
    }
    if (n == 2) {
        double ab[3]; sub3(S[1], S[0], ab);
        for (int k = 0; k < 6 && n < 3; k++) {
            double dir[3], w[3], t[3], c[3];
            cross3(ab, AX[k], dir);
            if (dot3(dir, dir) < 1e-20) continue;
            msupport(A, B, dir, w); sub3(w, S[0], t); cross3(ab, t, c);
            if (dot3(c, c) > 1e-24) { memcpy(S[2], w, 24); n = 3; }
        }
        if (n < 3) return 0;
    }
    if (n == 3) {
        double ab[3], ac[3], nrm[3], w[3], t[3];
        sub3(S[1], S[0], ab); sub3(S[2], S[0], ac); cross3(ab, ac, nrm);
        for (int sgn = 0; sgn < 2 && n < 4; sgn++) {
            double dir[3] = {sgn ? -nrm[0] : nrm[0], sgn ? -nrm[1] : nrm[1], sgn ? -nrm[2] : nrm[2]};
            msupport(A, B, dir, w); sub3(w, S[0], t);
            if (fabs(dot3(t, nrm)) > 1e-14 * (1 + dot3(nrm, nrm))) { memcpy(S[3], w, 24); n = 4; }
        }
        if (n < 4) return 0;
    }
    return 1;
}

/* signed distance between two convex cores: > 0 separated, < 0 = -(penetration depth) */
static double gjk_epa(const gshape_t* A, const gshape_t* B) {
    double S[4][3], v[3], w[3];
    int n = 1;
    double d0[3]; sub3(A->c, B->c, d0);
    if (dot3(d0, d0) < 1e-24) { d0[0] = 1; d0[1] = 0; d0[2] = 0; }
    msupport(A, B, d0, S[0]);
    memcpy(v, S[0], 24);
    for (int it = 0; it < 100; it++) {
        double vv = dot3(v, v);
        if (vv < 1e-22) break;                        /* origin on the simplex: touching or intersecting */
        double nv[3] = {-v[0], -v[1], -v[2]};
        msupport(A, B, nv, w);
        double vw = dot3(v, w);
        if (vv - vw <= 1e-14 * vv + 1e-18) return sqrt(vv);   /* no progress possible: |v| is the distance */
        int dup = 0;
        for (int k = 0; k < n; k++) { double t[3]; sub3(w, S[k], t); if (dot3(t, t) < 1e-26) dup = 1; }
        if (dup) return sqrt(vv);
        memcpy(S[n], w, 24);
        n++;
        n = closest_simplex(S, n, v);
        if (n < 0) { n = 4; double z[3] = {0, 0, 0}; memcpy(v, z, 24); break; }
    }
    if (dot3(v, v) >= 1e-22) return sqrt(dot3(v, v));
    if (n < 4 && !blow_up(A, B, S, n)) return 0.0;      /* cores touch in a degenerate way */
    /* the blown-up tetrahedron must contain the origin for EPA; if it does not, the cores merely touch */
    {
        static const int F[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
        static const int O[4] = {3, 2, 1, 0};
        for (int f = 0; f < 4; f++) {
            double ab[3], ac[3], nrm[3], ad[3];
            sub3(S[F[f][1]], S[F[f][0]], ab); sub3(S[F[f][2]], S[F[f][0]], ac); cross3(ab, ac, nrm); sub3(S[O[f]], S[F[f][0]], ad);
            double sd = dot3(nrm, ad), so = -dot3(nrm, S[F[f][0]]);
            if (sd * so < -1e-18) return 0.0;
        }
    }
    return -epa_depth(A, B, S);
}

static void make_gshape(blob_t b, int64_t s, const double* w, int is_cyl, gshape_t* g) {
    int64_t base = BI(b, MRB_H_OFF_SHAPES) + s * MRB_SHAPE_WORDS;
    int64_t core = BI(b, base);
    memset(g, 0, sizeof(*g));
    if (core == MRB_CORE_POINT) { g->kind = 0; memcpy(g->a, w, 24); memcpy(g->c, w, 24); }
    else if (core == MRB_CORE_SEG) {
        g->kind = is_cyl ? 3 : 1;
        memcpy(g->a, w, 24); memcpy(g->b, w + 3, 24);
        g->rc = is_cyl ? shape_radius(b, s) : 0.0;
        for (int k = 0; k < 3; k++) g->c[k] = 0.5 * (w[k] + w[3 + k]);
    } else if (core == MRB_CORE_BOX) {
        g->kind = 2; memcpy(g->a, w, 24); memcpy(g->R, w + 3, 72); memcpy(g->h, w + 12, 24); memcpy(g->c, w, 24);
    } else { /* upright cylinder c, r, half height */
        g->kind = 3; g->rc = w[3];
        memcpy(g->a, w, 24); memcpy(g->b, w, 24); g->a[2] -= w[4]; g->b[2] += w[4]; memcpy(g->c, w, 24);
    }
}

/* same contract as orc_check_configs; is_cyl[n_shapes] (nullable): shapes that are rai cylinders (true convex cylinders
 * here; capsules of the same radius and length in the first oracle and on the device) */
int orc_gjk_check_configs(blob_t b, const double* q, int64_t B, double tol, const uint8_t* is_cyl, uint8_t* flags, double* pen,
                          double* mind, int nthreads) {
    if (BI(b, MRB_H_MAGIC) != MRB_BLOB_MAGIC || BI(b, MRB_H_VERSION) != MRB_BLOB_VERSION) return -1;
    int64_t D = BI(b, MRB_H_DOF), ns = BI(b, MRB_H_NMOV) + BI(b, MRB_H_NSTA);
    if (ns > MAX_SHAPES || BI(b, MRB_H_NFRAMES) > MAX_FRAMES) return -2;
    if (tol < 0) tol = BF(b, MRB_H_TOL);
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 1 ? nthreads : 1)
#endif
    {
        double* W = (double*)malloc(sizeof(double) * 16 * ns);
        gshape_t* G = (gshape_t*)malloc(sizeof(gshape_t) * ns);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < B; i++) {
            world_shapes(b, q + i * D, W);
            for (int64_t s = 0; s < ns; s++) make_gshape(b, s, W + s * 16, is_cyl ? is_cyl[s] : 0, &G[s]);
            double p = 0, m = 1e300;
            for (int pass = 0; pass < 2; pass++) {   /* pass 0: moving pairs, pass 1: static-static pairs (constant per mode) */
                int ntypes = pass == 0 ? MRB_NUM_PAIR_TYPES : 1;
                for (int t = 0; t < ntypes; t++) {
                    int64_t n = pass == 0 ? BI(b, MRB_H_N_PAIRS + t) : BI(b, MRB_H_N_STATIC_PAIRS);
                    int64_t off = pass == 0 ? BI(b, MRB_H_OFF_PAIRS + t) : BI(b, MRB_H_OFF_STATIC_PAIRS);
                    for (int64_t j = 0; j < n; j++) {
                        int64_t a, c;
                        if (pass == 0) { int64_t pk = BI(b, off + j); a = pk & 0xffff; c = (pk >> 16) & 0xfff; }
                        else { a = BI(b, off + 3 * j + 1); c = BI(b, off + 3 * j + 2); }
                        double ra = (is_cyl && is_cyl[a]) || BI(b, BI(b, MRB_H_OFF_SHAPES) + a * MRB_SHAPE_WORDS) == MRB_CORE_CYLZ ? 0.0 : shape_radius(b, a);
                        double rc = (is_cyl && is_cyl[c]) || BI(b, BI(b, MRB_H_OFF_SHAPES) + c * MRB_SHAPE_WORDS) == MRB_CORE_CYLZ ? 0.0 : shape_radius(b, c);
                        double dc = gjk_epa(&G[a], &G[c]);
                        double d = dc - ra - rc;
                        if (d < 0) p -= d;
                        if (pass == 0 && d < m) m = d;
                    }
                }
            }
            flags[i] = !(p > tol);
            if (pen) pen[i] = p;
            if (mind) mind[i] = m;
        }
        free(W); free(G);
    }
    return 0;
}

/* test hook: signed distance of one pair of cores given as world data rows (16 doubles each), core types as in the blob */
double orc_gjk_pair(int core_a, const double* wa, double cyl_ra, int core_b, const double* wb, double cyl_rb) {
    gshape_t A, B;
    const double* w[2] = {wa, wb};
    int core[2] = {core_a, core_b};
    double cr[2] = {cyl_ra, cyl_rb};
    gshape_t* G[2] = {&A, &B};
    for (int i = 0; i < 2; i++) {
        gshape_t* g = G[i];
        memset(g, 0, sizeof(*g));
        if (core[i] == MRB_CORE_POINT) { g->kind = 0; memcpy(g->a, w[i], 24); memcpy(g->c, w[i], 24); }
        else if (core[i] == MRB_CORE_SEG) {
            g->kind = cr[i] > 0 ? 3 : 1; g->rc = cr[i];
            memcpy(g->a, w[i], 24); memcpy(g->b, w[i] + 3, 24);
            for (int k = 0; k < 3; k++) g->c[k] = 0.5 * (w[i][k] + w[i][3 + k]);
        } else { g->kind = 2; memcpy(g->a, w[i], 24); memcpy(g->R, w[i] + 3, 72); memcpy(g->h, w[i] + 12, 24); memcpy(g->c, w[i], 24); }
    }
    return gjk_epa(&A, &B);
}
