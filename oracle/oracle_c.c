/* CPU ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the C helpers on the reference's NanoWrap hot path.  Built by
 * oracle/build.py into oracle/liboracle.so and called through ctypes by
 * oracle/nanowrap_oracle.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load it.
 *
 * Pinning: tests/test_oracle_vs_ref.py (run wherever /root/reference exists) compares every
 * function here bit-for-bit with the reference's own compiled C (oracle/_ref/), and
 * tests/test_oracle_golden.py replays fixtures produced by that reference build.
 *
 * Arithmetic notes: the reference is built without -march flags, so gcc emits no FMA; this file
 * must be built the same way (oracle/build.py passes -ffp-contract=off to make that explicit).
 * float/double mixing below deliberately follows the reference expression by expression.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define NBR 20 /* NEIGHBORSIZE, membrane_mesh_utils.h:29 */

/* record layouts, membrane_mesh_utils.h:31-65 */
typedef struct { int32_t vertex, face, twin, next, prev; float length; int32_t component; } he_rec;
typedef struct { int32_t halfedge; float normal[3]; float area; int32_t component; } face_rec;
typedef struct { float position[3]; float normal[3]; int32_t halfedge, valence; int32_t neighbors[NBR];
                 int32_t component, locally_manifold; } vert_rec;

/* ---- adjoint scatter: conj_grad_utils.c:153-162 ------------------------------------------- */
void orc_ah_scatter(const int32_t *v_idx, const float *w, const float *fv, float *out, int64_t n_points)
{
    for (int64_t p = 0; p < n_points; ++p)
        for (int c = 0; c < 3; ++c) {
            float *dst = out + 3 * (int64_t)v_idx[3 * p + c];
            const float wc = w[3 * p + c];
            for (int a = 0; a < 3; ++a) dst[a] += wc * fv[3 * p + a];
        }
}

/* ---- umbrella Laplacian: conj_grad_utils.c:286-302 ---------------------------------------- */
void orc_l_func(const float *f, const int32_t *nb, const float *ref, float *d, int M, int N)
{
    (void)ref;
    for (int i = 0; i < M; ++i) {
        const int32_t *row = nb + (int64_t)i * N;
        if (row[0] == -1) continue;
        for (int a = 0; a < 3; ++a) {
            int cnt = 0;
            for (int k = 0; k < N && row[k] != -1; ++k, ++cnt)
                d[3 * i + a] += f[3 * row[k] + a] - f[3 * i + a];
            d[3 * i + a] /= cnt;
        }
    }
}

/* ---- its sequential "transpose": conj_grad_utils.c:344-364 -------------------------------- */
void orc_lh_func(const float *f, const int32_t *nb, const float *ref, float *d, int M, int N)
{
    (void)ref;
    for (int i = 0; i < M; ++i) {
        const int32_t *row = nb + (int64_t)i * N;
        if (row[0] == -1) continue;
        for (int a = 0; a < 3; ++a) {
            int cnt = 0;
            for (int k = 0; k < N && row[k] != -1; ++k, ++cnt)
                d[3 * row[k] + a] += f[3 * i + a] - f[3 * row[k] + a];
            for (int k = 0; k < cnt; ++k) d[3 * row[k] + a] /= cnt;
        }
    }
}

/* sum of squared 1-ring edge lengths on the reference geometry (conj_grad_utils.c:412-440) */
static float ring_sq(const float *g, const int32_t *row, int i, int N, int *cnt, int flip)
{
    float s = 0;
    int c = 0;
    for (int k = 0; k < N && row[k] != -1; ++k, ++c) {
        float d2 = 0;
        for (int a = 0; a < 3; ++a) {
            float dd = flip ? (g[3 * i + a] - g[3 * row[k] + a]) : (g[3 * row[k] + a] - g[3 * i + a]);
            d2 += dd * dd;
        }
        s += d2;
    }
    *cnt = c;
    return s;
}

/* ---- edge-length-normalised Laplacian: conj_grad_utils.c:412-491 (the acosf angle sum there is
 * computed and discarded, so it is not restated) ----------------------------------------------- */
void orc_lw_func(const float *f, const int32_t *nb, const float *g, float *d, int M, int N)
{
    for (int i = 0; i < M; ++i) {
        const int32_t *row = nb + (int64_t)i * N;
        int cnt;
        if (row[0] == -1) continue;
        float s = ring_sq(g, row, i, N, &cnt, 0);
        if (!(s > 0)) continue;
        for (int k = 0; k < cnt; ++k)
            for (int a = 0; a < 3; ++a) d[3 * i + a] += (f[3 * row[k] + a] - f[3 * i + a]) / sqrtf(s);
    }
}

/* ---- scatter form: conj_grad_utils.c:628-704 ------------------------------------------------ */
void orc_lhw_func(const float *f, const int32_t *nb, const float *g, float *d, int M, int N)
{
    for (int i = 0; i < M; ++i) {
        const int32_t *row = nb + (int64_t)i * N;
        int cnt;
        if (row[0] == -1) continue;
        float s = ring_sq(g, row, i, N, &cnt, 1);
        if (!(s > 0)) continue;
        for (int k = 0; k < cnt; ++k)
            for (int a = 0; a < 3; ++a) d[3 * row[k] + a] += (f[3 * i + a] - f[3 * row[k] + a]) / sqrtf(s);
    }
}

/* ---- 1/sqrt(sum|e|^2 + 1): conj_grad_utils.c:500-549 ---------------------------------------- */
void orc_vertex_area_weights(const float *g, const int32_t *nb, const float *unused, float *out, int M, int N)
{
    (void)unused;
    for (int i = 0; i < M; ++i) {
        const int32_t *row = nb + (int64_t)i * N;
        int cnt;
        if (row[0] == -1) continue;
        float s = ring_sq(g, row, i, N, &cnt, 0);
        float w = (s > 0) ? (float)(1.0 / sqrtf(s + 1)) : 0.0f;
        out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = w;
    }
}

/* ================= curvature: membrane_mesh_utils.c:915-1250 ================================= */
#define TINY 1e-15
typedef struct { double x, y, z; } d3;

static inline double nrm(d3 a) { double n = 0.0; n += a.x * a.x; n += a.y * a.y; n += a.z * a.z; return sqrt(n); }
static inline float nrmf(const float *a) { float n = 0.0f; n += a[0] * a[0]; n += a[1] * a[1]; n += a[2] * a[2]; return (float)sqrt(n); }
static inline double sdiv(double x, double y) { return ((y < 0 ? -y : y) < TINY) ? 0.0 : x / y; }   /* :63-68 */
static inline d3 fsub(const float *a, const float *b) { d3 r = { (double)a[0] - (double)b[0], (double)a[1] - (double)b[1], (double)a[2] - (double)b[2] }; return r; }
static inline double fdot(const float *a, d3 b) { double c = 0.0; c += (double)a[0] * b.x; c += (double)a[1] * b.y; c += (double)a[2] * b.z; return c; }
/* s(a) of SURVEY A.4: chord between unit normals from the squared cosine (:1086-1103) */
static inline double chord(double c) { double q = c * c; return (q > 1.0) ? sqrt(2.0) : sqrt(2.0 - 2.0 * sqrt(1.0 - q)); }

/* I - coef v v^T with the off-diagonals rounded through float (:231-253) */
static void projector(const float *v, double coef, double *m)
{
    const double a = (double)v[0], b = (double)v[1], c = (double)v[2];
    const float ab = (float)(-1.0 * coef * a * b), ac = (float)(-1.0 * coef * a * c), bc = (float)(-1.0 * coef * b * c);
    m[0] = 1.0 - coef * a * a; m[1] = ab; m[2] = ac;
    m[3] = ab; m[4] = 1.0 - coef * b * b; m[5] = bc;
    m[6] = ac; m[7] = bc; m[8] = 1.0 - coef * c * c;
}

static void mm3(const double *a, const double *b, double *c)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * k + j];
            c[3 * i + j] = s;
        }
}

/* Householder + one Givens rotation on the tangent 2x2 block (:618-720) */
static void tensor_eig(const double *Mv, const float *N, double *l1, double *l2, double *v1, double *v2)
{
    float dm[3] = { 1.0f - N[0], 0.0f - N[1], 0.0f - N[2] };
    float dp[3] = { 1.0f + N[0], 0.0f + N[1], 0.0f + N[2] };
    float nm = nrmf(dm), np_ = nrmf(dp), W[3];
    if (nm > np_) { W[0] = dm[0] / nm; W[1] = dm[1] / nm; W[2] = dm[2] / nm; }
    else { W[0] = dp[0] / np_; W[1] = dp[1] / np_; W[2] = dp[2] / np_; }
    double Q[9], QT[9], QM[9], B[9];
    projector(W, 2.0, Q);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) QT[3 * j + i] = Q[3 * i + j];
    mm3(Q, Mv, QM);
    mm3(QM, QT, B);
    double tau = sdiv(B[8] - B[4], 2.0 * B[5]);
    double t = ((tau < 0) ? -1 : 1) / (fabs(tau) + sqrt(1 + tau * tau));
    double a = B[4] - t * B[5], b = B[8] + t * B[5];
    double cs = 1.0 / sqrt(1 + t * t), sn = t * cs;
    double p[3] = { cs * QT[1] - sn * QT[2], cs * QT[4] - sn * QT[5], cs * QT[7] - sn * QT[8] };
    double q[3] = { sn * QT[1] + cs * QT[2], sn * QT[4] + cs * QT[5], sn * QT[7] + cs * QT[8] };
    if (a > b) { *l1 = a; *l2 = b; for (int i = 0; i < 3; ++i) { v1[i] = p[i]; v2[i] = q[i]; } }
    else { *l1 = b; *l2 = a; for (int i = 0; i < 3; ++i) { v2[i] = p[i]; v1[i] = q[i]; } }
}

/* closed-form 2x2 pseudo-inverse through the SVD angles (:841-890) */
static void pinv2(const double *A, double *Ai)
{
    const double a = A[0], b = A[1], c = A[2], d = A[3];
    const double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d;
    const double ab2 = a2 + b2, cd2 = c2 + d2, diff = ab2 - cd2, cross = 2 * (a * c + b * d);
    const double th = 0.5 * atan2(2 * (a * b + c * d), a2 + c2 - b2 - d2), ph = 0.5 * atan2(cross, diff);
    const double ct = cos(th), cp = cos(ph), st = sin(th), sp = sin(ph);
    const double ctcp = ct * cp, ctsp = ct * sp, stcp = st * cp, stsp = st * sp;
    const int sg0 = ((ctcp * a + ctsp * c + stcp * b + stsp * d) < 0) ? -1 : 1;
    const int sg1 = ((stsp * a - stcp * c - ctsp * b + ctcp * d) < 0) ? -1 : 1;
    const double ss = ab2 + cd2, sd = sqrt(diff * diff + cross * cross);
    const double s0 = sqrt((ss + sd) / 2.0), rem = ss - sd, s1 = (rem > 0) ? sqrt(rem / 2.0) : 0.0;
    const double thr = (1e-8) * 0.5 * sqrt(5.0) * s0;
    const double i0 = (s0 < thr) ? 0.0 : (1.0 / s0), i1 = (s1 < thr) ? 0.0 : (1.0 / s1);
    const double u = sg0 * i0, v = sg1 * i1;
    Ai[0] = ctcp * u + stsp * v; Ai[1] = ctsp * u - stcp * v;
    Ai[2] = stcp * u - ctsp * v; Ai[3] = stsp * u + ctcp * v;
}

void orc_curvature_grad(const void *vertices_, const void *faces_, const void *halfedges_,
                        float dN, float skip_prob, int n_vertices,
                        float *k_0, float *k_1, float *e_0, float *e_1, float *H, float *K,
                        float *dH, float *dK, float *E, float *pE, float *dE_nb,
                        float kc, float kg, float c0, float *dEdN, const double *jitter_u)
{
    const vert_rec *V = (const vert_rec *)vertices_;
    const face_rec *F = (const face_rec *)faces_;
    const he_rec *HE = (const he_rec *)halfedges_;
    const double kcd = (double)kc, kgd = (double)kg, c0d = (double)c0, dNd = (double)dN;
    /* unit edge vectors persist across neighbours AND vertices when an edge is degenerate (:1059-1062) */
    d3 e_hat = { 0, 0, 0 }, e1_hat = { 0, 0, 0 };
    int64_t ju = 0;
    (void)skip_prob; /* Monte-Carlo skipping (:962) is not used on the recipe path (skip_prob = 0) */

    for (int i = 0; i < n_vertices; ++i) {
        const vert_rec *cv = &V[i];
        if (cv->halfedge == -1) {                                             /* :962-973 */
            H[i] = K[i] = dH[i] = dK[i] = dE_nb[i] = E[i] = pE[i] = 0.0f;
            dEdN[3 * i] = dEdN[3 * i + 1] = dEdN[3 * i + 2] = 0.0f;
            continue;
        }
        const float *vi = cv->position, *Ni = cv->normal;
        float cen[3] = { 0, 0, 0 };
        double r_sum = 0.0, jw = 10000000000000000.0;
        int n = 0;
        /* pass 1 (:986-1009); the jitter-width update is unconditional (missing braces, :1002-1005) */
        while (n < NBR && cv->neighbors[n] != -1) {
            const float *vj = V[HE[cv->neighbors[n]].vertex].position;
            for (int a = 0; a < 3; ++a) cen[a] += vj[a];
            double len = nrm(fsub(vj, vi));
            if (len > TINY) r_sum += 1.0 / len;
            if (len < jw) jw = len;
            ++n;
        }
        for (int a = 0; a < 3; ++a) cen[a] /= n;
        for (int a = 0; a < 3; ++a) {                                         /* :1016-1017 */
            double u = jitter_u ? jitter_u[ju++] : 0.5;
            cen[a] += jw * (u - 0.5);
        }
        float dir[3] = { cen[0] - vi[0], cen[1] - vi[1], cen[2] - vi[2] };
        float dir_n = nrmf(dir);
        for (int a = 0; a < 3; ++a) dir[a] = (dir_n > 0.0f) ? dir[a] / dir_n : 0.0f;
        d3 sh = { (double)dir[0] * dNd, (double)dir[1] * dNd, (double)dir[2] * dNd };   /* NvidN */
        d3 vsh = { (double)vi[0] - sh.x, (double)vi[1] - sh.y, (double)vi[2] - sh.z }; /* viNvidN */
        double P[9], Mv[9] = { 0 };
        projector(Ni, 1.0, P);
        double areas = 0.0, dareas = 0.0;
        dE_nb[i] = 0.0f;
        /* pass 2 (:1046-1122) */
        for (int j = 0; j < n; ++j) {
            const he_rec *h = &HE[cv->neighbors[j]];
            const vert_rec *nv = &V[h->vertex];
            d3 e = fsub(nv->position, vi);
            d3 e1 = { e.x - sh.x, e.y - sh.y, e.z - sh.z };
            double len = nrm(e), len1 = nrm(e1);
            if (len > TINY) { e_hat.x = e.x / len; e_hat.y = e.y / len; e_hat.z = e.z / len; }
            if (len1 > TINY) { e1_hat.x = e1.x / len1; e1_hat.y = e1.y / len1; e1_hat.z = e1.z / len1; }
            d3 me = { e.x * -1.0, e.y * -1.0, e.z * -1.0 };
            d3 T = { P[0] * me.x + P[1] * me.y + P[2] * me.z,
                     P[3] * me.x + P[4] * me.y + P[5] * me.z,
                     P[6] * me.x + P[7] * me.y + P[8] * me.z };
            double Tn = nrm(T);
            double Tij[3] = { 0, 0, 0 };
            if (Tn > TINY) { Tij[0] = T.x / Tn; Tij[1] = T.y / Tn; Tij[2] = T.z / Tn; }
            double ci = chord(fdot(Ni, e_hat));
            double cj = chord(fdot(nv->normal, e_hat));
            double cj1 = chord(fdot(nv->normal, e1_hat));
            double kj = sdiv(2.0 * cj, len), kj1 = sdiv(2.0 * cj1, len1);
            double w = sdiv(sdiv(1.0, len), r_sum);
            double k = sdiv(2.0 * ((fdot(Ni, me) < 0) ? -1 : 1) * ci, len);
            double Aj = F[h->face].area;
            const float *vn = V[HE[h->next].vertex].position;
            d3 en = { (double)vn[0] - vsh.x, (double)vn[1] - vsh.y, (double)vn[2] - vsh.z };
            d3 cr = { e1.y * en.z - e1.z * en.y, e1.z * en.x - e1.x * en.z, e1.x * en.y - e1.y * en.x };
            double dAj = 0.5 * nrm(cr);
            dareas += dAj;
            areas += Aj;
            double t0 = 2.0 * kj - c0d, t1 = 2.0 * kj1 - c0d;
            dE_nb[i] += ((float)(Aj * w * 0.5 * kcd * (t0 * t0) - dAj * w * 0.5 * kcd * (t1 * t1))) / dN;   /* :1115 */
            const double wk = w * k;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Mv[3 * r + c] += (Tij[r] * Tij[c]) * wk;
        }
        double l1, l2, v1[3], v2[3];
        tensor_eig(Mv, Ni, &l1, &l2, v1, v2);
        if (isnan(l1)) {                                                      /* :1129-1139 */
            k_0[i] = k_1[i] = 0.0f;
            for (int a = 0; a < 3; ++a) v1[a] = v2[a] = 0.0;
        } else {
            k_0[i] = (float)(3.0 * l1 - l2);
            k_1[i] = (float)(3.0 * l2 - l1);
        }
        for (int a = 0; a < 3; ++a) { e_0[3 * i + a] = (float)v1[a]; e_1[3 * i + a] = (float)v2[a]; }
        H[i] = (float)(0.5 * (k_0[i] + k_1[i]));                              /* :1151 (float add) */
        K[i] = (float)(k_0[i] * k_1[i]);                                      /* :1152 (float mul) */
        /* pass 3 (:1161-1192): quadric in the principal frame displaced by dN; the zero rows the
         * reference pads to 20 contribute exact zeros, so the sums run over the real ring only */
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
        double Ar[2 * NBR], br[NBR];
        for (int j = 0; j < n; ++j) {
            d3 e = fsub(V[HE[cv->neighbors[j]].vertex].position, vi);
            double p = e.x * v1[0] + e.y * v1[1] + e.z * v1[2];
            double q = e.x * v2[0] + e.y * v2[1] + e.z * v2[2];
            Ar[2 * j] = p * p; Ar[2 * j + 1] = q * q;
            br[j] = Ar[2 * j] * k_0[i] + Ar[2 * j + 1] * k_1[i] - dNd;
        }
        for (int j = 0; j < n; ++j) {
            a00 += Ar[2 * j] * Ar[2 * j]; a01 += Ar[2 * j] * Ar[2 * j + 1];
            a10 += Ar[2 * j + 1] * Ar[2 * j]; a11 += Ar[2 * j + 1] * Ar[2 * j + 1];
        }
        double AtA[4] = { a00, a01, a10, a11 }, Ai[4], kp0 = 0.0, kp1 = 0.0;
        pinv2(AtA, Ai);
        for (int j = 0; j < n; ++j) {
            double r0 = 0.0, r1 = 0.0;
            r0 += Ai[0] * Ar[2 * j]; r0 += Ai[1] * Ar[2 * j + 1];
            r1 += Ai[2] * Ar[2 * j]; r1 += Ai[3] * Ar[2 * j + 1];
            kp0 += r0 * br[j]; kp1 += r1 * br[j];
        }
        dH[i] = (float)(0.5 * (kp0 + kp1));
        dK[i] = (float)(kp0 * kp1);
        double th = 2.0 * (double)H[i] - c0d;
        E[i] = (float)(areas * ((0.5 * kcd * (th * th) + kgd * (double)K[i])));            /* :1195 */
        pE[i] = (float)exp(-(1.0 / 0.0257) * (double)E[i]);                                /* :1197 */
        double td = 2.0 * (double)dH[i] - c0d;
        double dE_H = dareas * ((0.5 * kcd * (td * td) + kgd * (double)dK[i]));
        double dsum = ((double)E[i] - dE_H) / dNd + (double)dE_nb[i];
        double lim = 0.5 * (double)dir_n;
        double cl = (dsum > lim) ? lim : ((dsum < -lim) ? -lim : dsum);
        float g = (float)(-1.0 * ((float)cl) * (1.0 - pE[i]));                             /* :1213 */
        for (int a = 0; a < 3; ++a) dEdN[3 * i + a] = g * dir[a];
    }
}

/* ---- hole-punch candidate pairing: membrane_mesh_utils.c:1261-1285 (face centroid) and :1301-1379 ---------------- */
typedef struct { int32_t vertex, face, twin, next, prev; float length; int32_t component; } orc_he_t;        /* membrane_mesh_utils.h:31-39 */
typedef struct { int32_t halfedge; float normal[3]; float area; int32_t component; } orc_face_t;             /* :41-46 */
typedef struct { float position[3]; float normal[3]; int32_t halfedge, valence; int32_t neighbors[20];
                 int32_t component, locally_manifold; } orc_vert_t;                                          /* :57-65 */

static float orc_fdot(const float *a, const float *b) { float c = 0.0; int i; for (i = 0; i < 3; ++i) c += a[i] * b[i]; return c; }
static float orc_fnorm(const float *p) { float n = 0.0; int i; for (i = 0; i < 3; ++i) n += p[i] * p[i]; return sqrt(n); }
static void orc_centroid(const orc_face_t *f, const orc_vert_t *V, const orc_he_t *HE, float *c)
{
    const int32_t he = f->halfedge;
    const float *p0 = V[HE[HE[he].prev].vertex].position, *p1 = V[HE[he].vertex].position, *p2 = V[HE[HE[he].next].vertex].position;
    const float third = 0.3333333333333333;
    int k;
    for (k = 0; k < 3; ++k) { float s = p0[k] + p1[k]; s = s + p2[k]; c[k] = s * third; }
}

void orc_holepunch_pair(const void *vertices, const void *faces, const void *halfedges, const int32_t *candidates,
                        int n_candidates, int32_t *pairs)
{
    const orc_vert_t *V = (const orc_vert_t *)vertices;
    const orc_face_t *F = (const orc_face_t *)faces;
    const orc_he_t *HE = (const orc_he_t *)halfedges;
    int i, j, k;
    for (i = 0; i < n_candidates; ++i) {
        const orc_face_t *fi = &F[candidates[i]];
        float ci[3], cj[3], n_hat[3], sh[3], s[3];
        float min_shift = 1e6;
        orc_centroid(fi, V, HE, ci);
        for (j = i + 1; j < n_candidates; ++j) {
            const orc_face_t *fj = &F[candidates[j]];
            float nd, ndi, ndj, shn, dotn, b, a;
            if (pairs[j] != -1) continue;                              /* :1334 */
            nd = orc_fdot(fi->normal, fj->normal);
            if (nd > -0.6) continue;                                   /* :1342 */
            orc_centroid(fj, V, HE, cj);
            for (k = 0; k < 3; ++k) n_hat[k] = (fi->normal[k] + fj->normal[k]) * 0.5f;    /* :1348-1349 */
            for (k = 0; k < 3; ++k) sh[k] = ci[k] - cj[k];                                /* :1352 */
            ndi = orc_fdot(fi->normal, sh);
            ndj = orc_fdot(fj->normal, sh);
            if ((ndi < 0) && (ndj > 0)) continue;                      /* :1359 */
            shn = orc_fnorm(sh);
            dotn = orc_fdot(n_hat, sh);
            b = dotn * shn;                                            /* :1364 */
            for (k = 0; k < 3; ++k) s[k] = sh[k] - n_hat[k] * b;
            a = orc_fdot(s, s);
            if (a < min_shift) { min_shift = a; pairs[i] = j; }        /* :1370-1374 */
        }
    }
}
