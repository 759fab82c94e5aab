// quality.cu -- the host hot spots either side of the solver that SURVEY section 8(f) ranks next:
//
//   nw_points_from_mesh              : uniform grid sampling of every triangle in its own plane
//                                      (evaluation_utils.py:35-145, the per-triangle Python loop)
//   nw_holepunch_pair_candidate_faces: O(n^2) pairing of opposing candidate faces
//                                      (membrane_mesh_utils.c:1301-1379; binding _membrane_mesh.pyx:890-907)
//
// (The third one, nearest-point queries against a raw point set for average_squared_distance and the hole-punch
// candidate search, is nw_set_point_targets in tree.cu: it reuses the solver's search hierarchy.)
//
// THIS FILE IS COMPILED WITH -fmad=false: every +,-,*,/ and sqrt is one IEEE operation in the order numpy / the
// reference's C evaluates it, so samples and pairings are bit-identical to the reference's.
#include <cub/cub.cuh>
#include <cfloat>
#include <cmath>
#include <cstring>
#include "common.cuh"

namespace {

struct HeRec { int vertex, face, twin, next, prev; float length; int component; };                 // membrane_mesh_utils.h:31-39
struct FaceRec { int halfedge; float normal[3]; float area; int component; };                      // :41-46
struct VertRec { float position[3]; float normal[3]; int halfedge, valence; int neighbors[NW_NEIGHBORSIZE];
                 int component, locally_manifold; };                                               // :57-65
static_assert(sizeof(HeRec) == 28 && sizeof(FaceRec) == 24 && sizeof(VertRec) == 120, "record layouts");

// ---- points_from_mesh ------------------------------------------------------------------------------------------------
struct f3 { float x, y, z; };
__device__ __forceinline__ f3 sub3(const f3 a, const f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
// numpy.cross for 3-vectors: c0 = a1*b2 - a2*b1, ... (float32 products, one subtraction)
__device__ __forceinline__ f3 cross3(const f3 a, const f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// numpy.linalg.norm(axis=1) on float32: sqrt((x0^2 + x1^2) + x2^2)
__device__ __forceinline__ float norm3(const f3 a) { return sqrtf((a.x * a.x + a.y * a.y) + a.z * a.z); }
// (a*b).sum(1) on float32 rows of three
__device__ __forceinline__ float dot3(const f3 a, const f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float sign32(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : (v == 0.f ? 0.f : v)); }   // numpy.sign: nan -> nan

// numpy.arange(start, stop, step) called with numpy.float32 scalars for start/stop and a Python float step (NEP 50:
// the scalar arithmetic inside arange stays float32, the result array is float64): length from float32 arithmetic,
// first two values start and float32(start + step), the rest start + i * (second - first) in float64.
struct Arange {
    double start, delta;
    int n;
    __device__ __forceinline__ Arange(float start32, float stop32, float step32) {
        const float q = (stop32 - start32) / step32;
        const double c = ceil((double)q);
        n = (c > 0.0 && c < 2.0e9) ? (int)c : 0;                       // NaN / non-positive: empty
        start = (double)start32;
        delta = (double)(start32 + step32) - start;
    }
    __device__ __forceinline__ double at(int i) const { return i == 0 ? start : start + (double)i * delta; }
};

// Walks the sample grid of one face in numpy's order (y outer, x inner) and hands every accepted sample to `emit`.
// Returns the number of samples.  evaluation_utils.py:59-133.
template <typename Emit>
__device__ __forceinline__ int face_samples(const float4 p0, const float4 p1, const float4 p2, float dx32, float half32, Emit emit) {
    const f3 t0 = {p0.x, p0.y, p0.z}, t1 = {p1.x, p1.y, p1.z}, t2 = {p2.x, p2.y, p2.z};
    f3 nrm = cross3(sub3(t2, t1), sub3(t0, t1));                                    // :60
    const float nn = norm3(nrm);                                                    // :61
    if (!(nn != 0.f)) return 0;                                                     // :66 zero-area triangles are dropped
    nrm = {nrm.x / nn, nrm.y / nn, nrm.z / nn};                                     // :72
    const f3 v0 = sub3(t1, t0);                                                     // :78
    const float e0n = norm3(v0);
    const f3 e0 = {v0.x / e0n, v0.y / e0n, v0.z / e0n};                             // :80
    const f3 e1 = cross3(nrm, e0);                                                  // :81
    const float x0 = dot3(t0, e0), y0 = dot3(t0, e1), x1 = dot3(t1, e0), y1 = dot3(t1, e1), x2 = dot3(t2, e0), y2 = dot3(t2, e1);   // :84-89
    const float xl = fminf(fminf(x0, x1), x2), xu = fmaxf(fmaxf(x0, x1), x2);       // :94-97 (numpy min/max propagate NaN; a NaN here empties the grid either way)
    const float yl = fminf(fminf(y0, y1), y2), yu = fmaxf(fmaxf(y0, y1), y2);
    const float x1x0 = x1 - x0, x2x1 = x2 - x1, x0x2 = x0 - x2;                     // :100-102
    float m0 = (y1 - y0) / x1x0, m1 = (y2 - y1) / x2x1, m2 = (y0 - y2) / x0x2;      // :103-108
    if (x1x0 == 0.f) m0 = 0.f;
    if (x2x1 == 0.f) m1 = 0.f;
    if (x0x2 == 0.f) m2 = 0.f;
    const double s1 = (double)sign32(m1), s2 = (double)sign32(m2);                  // :109-110
    if (x0 != x0 || y0 != y0 || xl != xl || xu != xu || yl != yl || yu != yu) return 0;
    const Arange ax((xl - x0) - half32, xu - x0, dx32), ay((yl - y0) - half32, yu - y0, dx32);   // :122-123
    const double m0d = m0, m1d = m1, m2d = m2, x0d = x0, x1d = x1, x2d = x2;
    const double y10 = (double)(y1 - y0), y20 = (double)(y2 - y0);                  // float32 scalar differences, then promoted
    int n = 0;
    for (int iy = 0; iy < ay.n; ++iy) {
        const double Y = ay.at(iy);
        for (int ix = 0; ix < ax.n; ++ix) {
            const double X = ax.at(ix);
            // :128  (Y > X*m0) & (s1*Y > s1*(y1-y0 + (X-x1+x0)*m1)) & (s2*Y < s2*(y2-y0 + (X-x2+x0)*m2))
            const bool in = (Y > X * m0d) && (s1 * Y > s1 * (y10 + ((X - x1d) + x0d) * m1d)) && (s2 * Y < s2 * (y20 + ((X - x2d) + x0d) * m2d));
            if (in) {
                // :131  X*e0 + Y*e1 + tris[i,0,:]  in float64
                emit(n, (X * (double)e0.x + Y * (double)e1.x) + (double)t0.x, (X * (double)e0.y + Y * (double)e1.y) + (double)t0.y,
                     (X * (double)e0.z + Y * (double)e1.z) + (double)t0.z);
                ++n;
            }
        }
    }
    return n;
}

struct NoEmit { __device__ __forceinline__ void operator()(int, double, double, double) const {} };
struct WriteEmit {
    double *out;
    __device__ __forceinline__ void operator()(int k, double x, double y, double z) const { out[3 * (size_t)k] = x; out[3 * (size_t)k + 1] = y; out[3 * (size_t)k + 2] = z; }
};

__global__ void __launch_bounds__(128) k_sample_count(const float4 *__restrict__ pos, const int *__restrict__ faces, int F, float dx32, float half32,
                                                      int *__restrict__ count) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    count[f] = face_samples(pos[faces[3 * (size_t)f]], pos[faces[3 * (size_t)f + 1]], pos[faces[3 * (size_t)f + 2]], dx32, half32, NoEmit());
}
__global__ void __launch_bounds__(128) k_sample_emit(const float4 *__restrict__ pos, const int *__restrict__ faces, int F, float dx32, float half32,
                                                     const long long *__restrict__ offset, double *__restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    WriteEmit w;
    w.out = out + 3 * (size_t)offset[f];
    face_samples(pos[faces[3 * (size_t)f]], pos[faces[3 * (size_t)f + 1]], pos[faces[3 * (size_t)f + 2]], dx32, half32, w);
}
__global__ void k_pack_pos(const float *__restrict__ src, int M, float4 *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) dst[i] = make_float4(src[3 * (size_t)i], src[3 * (size_t)i + 1], src[3 * (size_t)i + 2], 0.f);
}

// ---- hole-punch candidate pairing ---------------------------------------------------------------------------------------
__device__ __forceinline__ float fdot(const float *a, const float *b) { float c = 0.0f; c += a[0] * b[0]; c += a[1] * b[1]; c += a[2] * b[2]; return c; }   // ffdot3f, :392
__device__ __forceinline__ float fnorm(const float *p) { float n = 0.0f; n += p[0] * p[0]; n += p[1] * p[1]; n += p[2] * p[2]; return (float)sqrt((double)n); }   // fnorm3f, :47

// calculate_face_centroid, :1261-1285 (vertex order prev.vertex, he.vertex, next.vertex; (p0+p1)+p2 times float(1/3))
__global__ void k_candidate_faces(const VertRec *__restrict__ V, const FaceRec *__restrict__ Fc, const HeRec *__restrict__ HE,
                                  const int *__restrict__ cand, int n, float4 *__restrict__ cen, float4 *__restrict__ nrm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FaceRec *f = &Fc[cand[i]];
    const int he = f->halfedge;
    const float *p0 = V[HE[HE[he].prev].vertex].position, *p1 = V[HE[he].vertex].position, *p2 = V[HE[HE[he].next].vertex].position;
    const float third = (float)0.3333333333333333;
    cen[i] = make_float4(((p0[0] + p1[0]) + p2[0]) * third, ((p0[1] + p1[1]) + p2[1]) * third, ((p0[2] + p1[2]) + p2[2]) * third, 0.f);
    nrm[i] = make_float4(f->normal[0], f->normal[1], f->normal[2], 0.f);
}

// One warp per candidate i; lanes stride over j > i; the warp keeps (smallest abs_shift, then smallest j): the serial
// loop (:1327-1376) takes the FIRST j that is strictly smaller than everything before it, i.e. exactly that.  The
// `pairs[j] != -1` test of the serial loop (:1334) never fires: pairs[j] is only ever written when the outer loop reaches
// i == j, which is after every i < j has finished, and the caller initialises pairs to -1 (_membrane_mesh.pyx:899).
__global__ void __launch_bounds__(256) k_pair_candidates(const float4 *__restrict__ cen, const float4 *__restrict__ nrm, int n, int *__restrict__ pairs) {
    const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const float4 ci4 = cen[i], ni4 = nrm[i];
    const float ci[3] = {ci4.x, ci4.y, ci4.z}, ni[3] = {ni4.x, ni4.y, ni4.z};
    float best = 1e6f;                                                   // min_shift, :1325
    int best_j = 0x7fffffff;
    for (int j = i + 1 + lane; j < n; j += 32) {
        const float4 cj4 = cen[j], nj4 = nrm[j];
        const float cj[3] = {cj4.x, cj4.y, cj4.z}, nj[3] = {nj4.x, nj4.y, nj4.z};
        const float nd = fdot(ni, nj);
        if ((double)nd > -0.6) continue;                                 // :1342 (float against a double constant)
        const float n_hat[3] = {(ni[0] + nj[0]) * 0.5f, (ni[1] + nj[1]) * 0.5f, (ni[2] + nj[2]) * 0.5f};   // :1348-1349
        const float sh[3] = {ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2]};                                  // :1352
        const float ndi = fdot(ni, sh), ndj = fdot(nj, sh);
        if ((ndi < 0) && (ndj > 0)) continue;                            // :1359
        const float shn = fnorm(sh);
        const float k = fdot(n_hat, sh) * shn;                           // :1363-1364
        const float s[3] = {sh[0] - n_hat[0] * k, sh[1] - n_hat[1] * k, sh[2] - n_hat[2] * k};              // :1364-1365
        const float a = fdot(s, s);                                      // :1369
        if (a < best) { best = a; best_j = j; }                          // lanes see their j in ascending order: first minimum kept
    }
    // warp argmin on (abs_shift, j); NaN never wins (a < best is false for NaN, as in the serial loop)
    for (int o = 16; o; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (ob < best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
    }
    if (lane == 0) pairs[i] = best_j == 0x7fffffff ? -1 : best_j;
}

}  // namespace

static int scan_counts(nw_ctx *h, const int *in, long long *out, int n) {
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, h->stream);
    if (tmp > h->cub_tmp_bytes) { NW_CHECK(nw_alloc(h, (char **)&h->cub_tmp, tmp)); h->cub_tmp_bytes = tmp; }
    NW_CUDA(cub::DeviceScan::ExclusiveSum(h->cub_tmp, tmp, in, out, n, h->stream));
    h->launches += 2;
    return NW_OK;
}

// Two-step protocol: out == NULL (or capacity too small) -> only *n_out is set; otherwise the samples are written, (n,3)
// float64, in the reference's generation order (face by face, rows of the face's grid y-outer / x-inner).  The
// reference then shuffles them (np.random.choice without replacement, :137); that part stays with the caller.
extern "C" int nw_points_from_mesh(nw_ctx *h, const float *pos, const int32_t *faces, int M, int F, double dx_min, double *out,
                                   int64_t capacity, int64_t *n_out) {
    if (!h) return NW_ERR_ARG;
    NW_ARG(pos && faces && M > 0 && F > 0 && n_out, "nw_points_from_mesh: empty mesh");
    NW_ARG(dx_min > 0.0 && dx_min <= FLT_MAX, "nw_points_from_mesh: dx_min must be positive");
    NW_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int B = 128;
    float4 *d_pos = nullptr;
    float *d_raw = nullptr;
    int *d_faces = nullptr, *d_cnt = nullptr;
    long long *d_off = nullptr;
    double *d_out = nullptr;
    int rc = NW_OK;
    auto done = [&](int r) { nw_free(&d_pos); nw_free(&d_raw); nw_free(&d_faces); nw_free(&d_cnt); nw_free(&d_off); nw_free(&d_out); return r; };
#define QX(x) do { rc = (x); if (rc != NW_OK) return done(rc); } while (0)
    QX(nw_alloc(h, &d_pos, (size_t)M)); QX(nw_alloc(h, &d_raw, (size_t)3 * M)); QX(nw_alloc(h, &d_faces, (size_t)3 * F));
    QX(nw_alloc(h, &d_cnt, (size_t)F + 1)); QX(nw_alloc(h, &d_off, (size_t)F + 1));
    QX(nw_h2d(h, d_raw, pos, sizeof(float) * 3 * (size_t)M));
    QX(nw_h2d(h, d_faces, faces, sizeof(int) * 3 * (size_t)F));
    k_pack_pos<<<nw_grid(M, 256), 256, 0, s>>>(d_raw, M, d_pos);
    // the Python scalars of the reference become float32 where they meet float32 numpy scalars (NEP 50)
    const float dx32 = (float)dx_min, half32 = (float)(dx_min / 2);
    cudaMemsetAsync(d_cnt + F, 0, sizeof(int), s);
    k_sample_count<<<nw_grid(F, B), B, 0, s>>>(d_pos, d_faces, F, dx32, half32, d_cnt);
    h->launches += 2;
    QX(scan_counts(h, d_cnt, d_off, F + 1));
    long long total = 0;
    if (cudaMemcpyAsync(&total, d_off + F, sizeof(long long), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        h->err = "nw_points_from_mesh: count read-back failed";
        return done(NW_ERR_CUDA);
    }
    *n_out = total;
    if (!out || capacity < total || total == 0) return done(NW_OK);
    QX(nw_alloc(h, &d_out, (size_t)3 * total));
    k_sample_emit<<<nw_grid(F, B), B, 0, s>>>(d_pos, d_faces, F, dx32, half32, d_off, d_out);
    h->launches += 1;
    if (cudaMemcpyAsync(out, d_out, sizeof(double) * 3 * (size_t)total, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        h->err = "nw_points_from_mesh: sample read-back failed";
        return done(NW_ERR_CUDA);
    }
    return done(NW_OK);
#undef QX
}

// Same argument list as the reference's C function plus the array lengths; pairs (n_candidates int32) is overwritten:
// pairs[i] = index INTO candidates of the face paired with candidates[i], or -1.
extern "C" int nw_holepunch_pair_candidate_faces(nw_ctx *h, const void *vertices, const void *faces, const void *halfedges, int n_vertices,
                                                 int n_faces, int n_halfedges, const int32_t *candidates, int n_candidates, int32_t *pairs) {
    if (!h) return NW_ERR_ARG;
    NW_ARG(vertices && faces && halfedges && n_vertices > 0 && n_faces > 0 && n_halfedges > 0, "nw_holepunch_pair_candidate_faces: empty mesh");
    NW_ARG(n_candidates >= 0 && (n_candidates == 0 || (candidates && pairs)), "nw_holepunch_pair_candidate_faces: bad candidate arrays");
    if (n_candidates == 0) return NW_OK;
    for (int i = 0; i < n_candidates; ++i)
        if (candidates[i] < 0 || candidates[i] >= n_faces) { h->err = "nw_holepunch_pair_candidate_faces: candidate index out of range"; return NW_ERR_ARG; }
    NW_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    // the record arrays share the curvature call's grow-only device copies (same three arrays, same caller)
    NW_CHECK(nw_alloc(h, (VertRec **)&h->cvV, (size_t)n_vertices)); NW_CHECK(nw_alloc(h, (FaceRec **)&h->cvF, (size_t)n_faces));
    NW_CHECK(nw_alloc(h, (HeRec **)&h->cvH, (size_t)n_halfedges));
    h->curvM = 0; h->curvK = nullptr;                     // the curvature outputs no longer describe these records
    const nw_h2d_job up[3] = {{h->cvV, vertices, sizeof(VertRec) * (size_t)n_vertices, 4}, {h->cvF, faces, sizeof(FaceRec) * (size_t)n_faces, 4},
                              {h->cvH, halfedges, sizeof(HeRec) * (size_t)n_halfedges, 4}};
    NW_CHECK(nw_h2d_many(h, up, 3));
    int *d_cand = nullptr, *d_pairs = nullptr;
    float4 *d_cen = nullptr, *d_nrm = nullptr;
    int rc = NW_OK;
    auto done = [&](int r) { nw_free(&d_cand); nw_free(&d_pairs); nw_free(&d_cen); nw_free(&d_nrm); return r; };
#define QX(x) do { rc = (x); if (rc != NW_OK) return done(rc); } while (0)
    QX(nw_alloc(h, &d_cand, (size_t)n_candidates)); QX(nw_alloc(h, &d_pairs, (size_t)n_candidates));
    QX(nw_alloc(h, &d_cen, (size_t)n_candidates)); QX(nw_alloc(h, &d_nrm, (size_t)n_candidates));
    QX(nw_h2d(h, d_cand, candidates, sizeof(int) * (size_t)n_candidates));
    k_candidate_faces<<<nw_grid(n_candidates, 256), 256, 0, s>>>((const VertRec *)h->cvV, (const FaceRec *)h->cvF, (const HeRec *)h->cvH, d_cand,
                                                                 n_candidates, d_cen, d_nrm);
    k_pair_candidates<<<nw_grid((int64_t)n_candidates * 32, 256), 256, 0, s>>>(d_cen, d_nrm, n_candidates, d_pairs);
    h->launches += 2;
    if (cudaMemcpyAsync(pairs, d_pairs, sizeof(int) * (size_t)n_candidates, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        h->err = std::string("nw_holepunch_pair_candidate_faces: ") + cudaGetErrorString(cudaGetLastError());
        return done(NW_ERR_CUDA);
    }
    return done(NW_OK);
#undef QX
}
