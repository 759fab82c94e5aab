// points.cu -- nw_set_points: upload one shard of localisations, Hilbert-sort it once (points never
// move during a fit) and keep every per-point stream as SoA in that order, so a warp of
// consecutive points queries neighbouring faces and every per-point load/store is a unit-stride
// 128 B line per warp.
#include <cub/cub.cuh>
#include <cfloat>
#include <cmath>
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {
    v &= 0x1fffffULL;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}

template <typename T>
__global__ void k_bbox(const T *__restrict__ pts, int64_t P, float *__restrict__ out6) {
    // out6 = {minx,miny,minz,maxx,maxy,maxz} as ordered-int atomics on floats
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x)
        for (int a = 0; a < 3; ++a) {
            float v = (float)pts[3 * i + a];
            if (v == v) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
        }
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int a = 0; a < 3; ++a) {
            // monotone float -> int mapping
            int l = __float_as_int(lo[a]); l = l >= 0 ? l : l ^ 0x7fffffff;
            int u = __float_as_int(hi[a]); u = u >= 0 ? u : u ^ 0x7fffffff;
            atomicMin((int *)out6 + a, l);
            atomicMax((int *)out6 + 3 + a, u);
        }
    }
}

template <typename T>
__global__ void k_point_keys(const T *__restrict__ pts, int64_t P, float3 lo, float3 inv,
                             unsigned long long *__restrict__ keys, int *__restrict__ idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    float x = ((float)pts[3 * i] - lo.x) * inv.x, y = ((float)pts[3 * i + 1] - lo.y) * inv.y,
          z = ((float)pts[3 * i + 2] - lo.z) * inv.z;
    const float top = 2097151.f;
    unsigned long long qx = (unsigned long long)fminf(fmaxf(x, 0.f), top);
    unsigned long long qy = (unsigned long long)fminf(fmaxf(y, 0.f), top);
    unsigned long long qz = (unsigned long long)fminf(fmaxf(z, 0.f), top);
    hilbert_axes_to_transpose(qx, qy, qz, 21);
    keys[i] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    idx[i] = (int)i;
}

template <typename T>
__global__ void k_permute_points(const T *__restrict__ pts, const int *__restrict__ perm, int64_t P,
                                 float *__restrict__ px, float *__restrict__ py, float *__restrict__ pz,
                                 double *__restrict__ px64, double *__restrict__ py64, double *__restrict__ pz64) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    int64_t s = perm[i];
    T x = pts[3 * s], y = pts[3 * s + 1], z = pts[3 * s + 2];
    px[i] = (float)x; py[i] = (float)y; pz[i] = (float)z;
    if (px64) { px64[i] = (double)x; py64[i] = (double)y; pz64[i] = (double)z; }
}

__global__ void k_permute_f3(const float *__restrict__ src, const int *__restrict__ perm, int64_t P,
                             float *__restrict__ ox, float *__restrict__ oy, float *__restrict__ oz) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    int64_t s = perm[i];
    ox[i] = src[3 * s]; oy[i] = src[3 * s + 1]; oz[i] = src[3 * s + 2];
}

// sum, max and count(<=0) of a weight array: double partials per block, folded on the host
__global__ void k_weight_stats(const float *__restrict__ w, int64_t n, double *__restrict__ part) {
    double s = 0.0, mx = 0.0, nonpos = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = w[i];
        s += (double)v;
        mx = fmax(mx, fabs((double)v));
        nonpos += (v > 0.f) ? 0.0 : 1.0;
    }
    __shared__ double sh[3][32];
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        nonpos += __shfl_xor_sync(0xffffffffu, nonpos, o);
    }
    int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][wid] = s; sh[1][wid] = mx; sh[2][wid] = nonpos; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a += sh[0][k]; b = fmax(b, sh[1][k]); c += sh[2][k]; }
        part[3 * blockIdx.x] = a; part[3 * blockIdx.x + 1] = b; part[3 * blockIdx.x + 2] = c;
    }
}

__global__ void k_mask_bits(const float *__restrict__ wx, const float *__restrict__ wy,
                            const float *__restrict__ wz, int64_t P, uint8_t *__restrict__ m) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    m[i] = (uint8_t)((wx[i] > 0.f ? 1 : 0) | (wy[i] > 0.f ? 2 : 0) | (wz[i] > 0.f ? 4 : 0));
}

__global__ void k_fill_int(int *p, int64_t n, int v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

inline float ordered_to_float(int v) {
    v = v >= 0 ? v : v ^ 0x7fffffff;
    float f;
    memcpy(&f, &v, 4);
    return f;
}

}  // namespace

int nw_comm_allreduce_host_doubles(nw_ctx *h, double *vals, int n);   // comm.cu

template <typename T>
static int set_points_impl(nw_ctx *h, const T *pts_host, int64_t P, const float *sigma_inv, float sinv_scalar,
                           const float *weights) {
    cudaStream_t s = h->stream;
    const int B = 256;
    unsigned long long *&keys = h->sp_k0, *&keys2 = h->sp_k1;
    int *&idx = h->sp_idx;
    float *d_bbox = nullptr, *&d_tmp3 = h->sp_tmp3;
    double *d_part = nullptr;
    int rc = NW_OK;
    auto cleanup = [&]() { nw_free(&d_bbox); nw_free(&d_part); };
#define NWX(x) do { rc = (x); if (rc != NW_OK) { cleanup(); return rc; } } while (0)
#define NWC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { h->err = std::string(#x) + ": " + cudaGetErrorString(e_); cleanup(); return NW_ERR_CUDA; } } while (0)

    h->P = P;
    h->weights_valid = false;
    h->seeds_cold = true;
    h->order_stale = true;
    h->feet_valid = false;
    h->epoch++;
    NWX(nw_alloc(h, &h->sp_pts, sizeof(T) * 3 * (size_t)P));
    T *d_pts = (T *)h->sp_pts;
    NWX(nw_h2d(h, d_pts, pts_host, sizeof(T) * 3 * (size_t)P));
    NWX(nw_alloc(h, &d_bbox, 6));
    {
        int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
        NWC(cudaMemcpyAsync(d_bbox, init, sizeof(init), cudaMemcpyHostToDevice, s));
    }
    k_bbox<T><<<std::min(nw_grid(P, B), 148 * 8), B, 0, s>>>(d_pts, P, d_bbox);
    h->launches++;
    int bb[6];
    NWC(cudaMemcpyAsync(bb, d_bbox, sizeof(bb), cudaMemcpyDeviceToHost, s));
    NWC(cudaStreamSynchronize(s));
    for (int a = 0; a < 6; ++a) h->bbox_pts[a] = ordered_to_float(bb[a]);
    if (P == 0) for (int a = 0; a < 6; ++a) h->bbox_pts[a] = 0.f;

    // Hilbert keys (21 bits per axis) + sort
    float3 lo = make_float3(h->bbox_pts[0], h->bbox_pts[1], h->bbox_pts[2]);
    float ext = fmaxf(fmaxf(h->bbox_pts[3] - lo.x, h->bbox_pts[4] - lo.y), h->bbox_pts[5] - lo.z);
    float iv = ext > 0.f ? 2097151.f / ext : 0.f;
    float3 inv = make_float3(iv, iv, iv);
    NWX(nw_alloc(h, &keys, (size_t)P)); NWX(nw_alloc(h, &keys2, (size_t)P));
    NWX(nw_alloc(h, &idx, (size_t)P)); NWX(nw_alloc(h, &h->perm, (size_t)P));
    if (P) {
        k_point_keys<T><<<nw_grid(P, B), B, 0, s>>>(d_pts, P, lo, inv, keys, idx);
        h->launches++;
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys, keys2, idx, h->perm, (int64_t)P, 0, 63, s);
        if (tmp > h->cub_tmp_bytes) { NWX(nw_alloc(h, (char **)&h->cub_tmp, tmp)); h->cub_tmp_bytes = tmp; }
        NWC(cub::DeviceRadixSort::SortPairs(h->cub_tmp, tmp, keys, keys2, idx, h->perm, (int64_t)P, 0, 63, s));
        h->launches += 4;
    }

    NWX(nw_alloc(h, &h->px, (size_t)P)); NWX(nw_alloc(h, &h->py, (size_t)P)); NWX(nw_alloc(h, &h->pz, (size_t)P));
    if (sizeof(T) == 8) {
        NWX(nw_alloc(h, &h->px64, (size_t)P)); NWX(nw_alloc(h, &h->py64, (size_t)P)); NWX(nw_alloc(h, &h->pz64, (size_t)P));
    } else {
        nw_free(&h->px64); nw_free(&h->py64); nw_free(&h->pz64);
    }
    if (P) {
        k_permute_points<T><<<nw_grid(P, B), B, 0, s>>>(d_pts, h->perm, P, h->px, h->py, h->pz, h->px64, h->py64, h->pz64);
        h->launches++;
    }
    NWC(cudaStreamSynchronize(s));

    // sigma_inv / weights
    h->sinv_scalar = sinv_scalar;
    // non-NULL arrays double as mode flags for the kernels: release the ones this call does not provide, keep (grow-only)
    // the ones it does
    if (!sigma_inv) { nw_free(&h->sx); nw_free(&h->sy); nw_free(&h->sz); }
    if (!weights) { nw_free(&h->wx); nw_free(&h->wy); nw_free(&h->wz); }
    h->has_mask = 0;
    h->wmean = 1.f;
    h->wn_max = fabs((double)sinv_scalar);
    h->weights_mode = 0;
    if ((sigma_inv || weights) && P) NWX(nw_alloc(h, &d_tmp3, (size_t)3 * P));
    if (sigma_inv) {
        NWX(nw_alloc(h, &h->sx, (size_t)P)); NWX(nw_alloc(h, &h->sy, (size_t)P)); NWX(nw_alloc(h, &h->sz, (size_t)P));
        if (P) {
            NWX(nw_h2d(h, d_tmp3, sigma_inv, sizeof(float) * 3 * (size_t)P));
            k_permute_f3<<<nw_grid(P, B), B, 0, s>>>(d_tmp3, h->perm, P, h->sx, h->sy, h->sz);
            h->launches++;
        }
        h->weights_mode = 1;
    }
    if (weights) {
        NWX(nw_alloc(h, &h->wx, (size_t)P)); NWX(nw_alloc(h, &h->wy, (size_t)P)); NWX(nw_alloc(h, &h->wz, (size_t)P));
        if (P) {
            NWC(cudaStreamSynchronize(s));
            NWX(nw_h2d(h, d_tmp3, weights, sizeof(float) * 3 * (size_t)P));
            k_permute_f3<<<nw_grid(P, B), B, 0, s>>>(d_tmp3, h->perm, P, h->wx, h->wy, h->wz);
            h->launches++;
        }
        h->weights_mode = 2;
    }
    // global weight statistics: weights / weights.mean() (mesh_conj_grad.py:162) over ALL ranks
    double stats[4] = {0.0, 0.0, 0.0, (double)(3 * P)};   // sum, max, nonpositive, count
    if (h->weights_mode != 0 && P) {
        const int G = 296;
        NWX(nw_alloc(h, &d_part, (size_t)3 * G * 3));
        std::vector<double> hp(3 * G * 3);
        const float *arr[3] = {h->weights_mode == 2 ? h->wx : h->sx, h->weights_mode == 2 ? h->wy : h->sy,
                               h->weights_mode == 2 ? h->wz : h->sz};
        for (int a = 0; a < 3; ++a) {
            k_weight_stats<<<G, B, 0, s>>>(arr[a], P, d_part + (size_t)a * 3 * G);
            h->launches++;
        }
        NWC(cudaMemcpyAsync(hp.data(), d_part, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost, s));
        NWC(cudaStreamSynchronize(s));
        for (int a = 0; a < 3; ++a)
            for (int g = 0; g < G; ++g) {
                stats[0] += hp[(size_t)a * 3 * G + 3 * g];
                stats[1] = std::max(stats[1], hp[(size_t)a * 3 * G + 3 * g + 1]);
                stats[2] += hp[(size_t)a * 3 * G + 3 * g + 2];
            }
    }
    double pg = (double)P;
    if (h->nranks > 1) {
        double v[3] = {stats[0], stats[2], stats[3]};
        NWX(nw_comm_allreduce_host_doubles(h, v, 3));      // sums
        stats[0] = v[0]; stats[2] = v[1]; stats[3] = v[2];
        double mx[2] = {stats[1], 0};
        // max via sum of one-hot is not available: gather maxima by allreducing per-rank slots
        std::vector<double> slots(h->nranks, 0.0);
        slots[h->rank] = stats[1];
        NWX(nw_comm_allreduce_host_doubles(h, slots.data(), h->nranks));
        for (double m : slots) mx[0] = std::max(mx[0], m);
        stats[1] = mx[0];
        double pp = pg;
        NWX(nw_comm_allreduce_host_doubles(h, &pp, 1));
        pg = pp;
    }
    h->P_global = (int64_t)pg;
    if (h->weights_mode != 0) {
        h->wmean = stats[3] > 0 ? (float)(stats[0] / stats[3]) : 1.f;
        h->wn_max = h->wmean != 0.f ? stats[1] / fabs((double)h->wmean) : stats[1];
        h->has_mask = stats[2] > 0 ? 1 : 0;
        if (!h->has_mask) nw_free(&h->pmask);
        if (h->has_mask && P) {
            NWX(nw_alloc(h, &h->pmask, (size_t)P));
            const float *ax = h->weights_mode == 2 ? h->wx : h->sx, *ay = h->weights_mode == 2 ? h->wy : h->sy,
                        *az = h->weights_mode == 2 ? h->wz : h->sz;
            k_mask_bits<<<nw_grid(P, B), B, 0, s>>>(ax, ay, az, P, h->pmask);
            h->launches++;
        }
    }
    if (h->weights_mode == 0) nw_free(&h->pmask);
    // per-point outputs
    NWX(nw_alloc(h, &h->slot, (size_t)P));
    NWX(nw_alloc(h, &h->w0, (size_t)P)); NWX(nw_alloc(h, &h->w1, (size_t)P)); NWX(nw_alloc(h, &h->w2, (size_t)P));
    NWX(nw_alloc(h, &h->rx, (size_t)P)); NWX(nw_alloc(h, &h->ry, (size_t)P)); NWX(nw_alloc(h, &h->rz, (size_t)P));
    NWX(nw_alloc(h, &h->fx, (size_t)P)); NWX(nw_alloc(h, &h->fy, (size_t)P)); NWX(nw_alloc(h, &h->fz, (size_t)P));   // foot points (seeds across uploads)
    if (P) {
        k_fill_int<<<nw_grid(P, B), B, 0, s>>>(h->slot, P, -1);
        h->launches++;
        NWC(cudaMemsetAsync(h->rx, 0, sizeof(float) * P, s));
        NWC(cudaMemsetAsync(h->ry, 0, sizeof(float) * P, s));
        NWC(cudaMemsetAsync(h->rz, 0, sizeof(float) * P, s));
    }
    NWC(cudaStreamSynchronize(s));
    cleanup();
    nw_free(&h->scratchP);
    h->scratchP_elems = 0;
    return NW_OK;
#undef NWX
#undef NWC
}

extern "C" int nw_set_points(nw_ctx *h, const void *pts, int pts_is_f64, int64_t P, const float *sigma_inv,
                             float sigma_inv_scalar, const float *weights) {
    if (!h) return NW_ERR_ARG;
    NW_ARG(P >= 0 && P < 2147483647LL, "nw_set_points: P must be in [0, 2^31)");
    NW_ARG(pts != nullptr || P == 0, "nw_set_points: pts is NULL");
    NW_CUDA(cudaSetDevice(h->device));
    if (pts_is_f64) return set_points_impl<double>(h, (const double *)pts, P, sigma_inv, sigma_inv_scalar, weights);
    return set_points_impl<float>(h, (const float *)pts, P, sigma_inv, sigma_inv_scalar, weights);
}
