// ring.cu -- the selectable 1-ring regularisers of conj_grad_utils.c:249-710 (Lfunc/Lhfunc/Lfunc3/
// Lhfunc3/wfunc in mesh_conj_grad.py:590-736).  Inactive with the shipped Lfuncs = ["I"], but part of
// the solver's operator surface.  All accumulate into the caller's output vector like the C originals.
//
// The "transpose" forms scatter into neighbours while sweeping the vertices in index order, i.e. every
// target vertex sees an ORDERED fold over its in-neighbours (ascending source index).  Here each target
// owns that fold: a transposed adjacency (CSR, sources ascending) is built once per topology with a
// radix sort, and one thread per target replays its fold in the same order -- bit-identical, no atomics.
#include <cub/cub.cuh>
#include "common.cuh"

namespace {

__global__ void k_edge_keys(const int *__restrict__ nbrT, const int *__restrict__ valence, int M,
                            const int *__restrict__ off, unsigned long long *__restrict__ keys) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    int n = valence[v], o = off[v];
    for (int k = 0; k < n; ++k) {
        unsigned long long t = (unsigned)nbrT[(size_t)k * M + v];
        keys[o + k] = (t << 32) | (unsigned)v;      // sort by (target, source)
    }
}
__global__ void k_split_keys(const unsigned long long *__restrict__ keys, int E, int *__restrict__ tgt, int *__restrict__ src) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) { tgt[e] = (int)(keys[e] >> 32); src[e] = (int)(keys[e] & 0xffffffffu); }
}
__global__ void k_count_targets(const int *__restrict__ tgt, int E, int *__restrict__ cnt) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) atomicAdd(&cnt[tgt[e]], 1);
}

// d_i = (d_i + sum_k (f_nk - f_i)) / N_i     conj_grad_utils.c:286-302
__global__ void k_l_func(const float *__restrict__ f, const int *__restrict__ nbrT, const int *__restrict__ valence, int M,
                         float *__restrict__ d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int n = valence[i];
    if (n == 0) return;
    for (int a = 0; a < 3; ++a) {
        float acc = d[3 * i + a];
        const float fi = f[3 * i + a];
        for (int k = 0; k < n; ++k) acc = __fadd_rn(acc, __fsub_rn(f[3 * nbrT[(size_t)k * M + i] + a], fi));
        d[3 * i + a] = __fdiv_rn(acc, (float)n);
    }
}

// for i ascending: d_t = (d_t + f_i - f_t) / N_i  for every neighbour t of i     :344-364
__global__ void k_lh_func(const float *__restrict__ f, const int *__restrict__ toff, const int *__restrict__ tsrc,
                          const int *__restrict__ valence, int M, float *__restrict__ d) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    for (int a = 0; a < 3; ++a) {
        float acc = d[3 * t + a];
        const float ft = f[3 * t + a];
        for (int e = toff[t]; e < toff[t + 1]; ++e) {
            const int i = tsrc[e];
            acc = __fdiv_rn(__fadd_rn(acc, __fsub_rn(f[3 * i + a], ft)), (float)valence[i]);
        }
        d[3 * t + a] = acc;
    }
}

// s_i = sum_k |g_nk - g_i|^2 in float32, reference accumulation order     :412-440
__global__ void k_ring_sq(const float *__restrict__ g, const int *__restrict__ nbrT, const int *__restrict__ valence, int M,
                          float *__restrict__ s) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int n = valence[i];
    float w = 0.f;
    for (int k = 0; k < n; ++k) {
        const int nb = nbrT[(size_t)k * M + i];
        float d2 = 0.f;
        for (int a = 0; a < 3; ++a) {
            const float dd = __fsub_rn(g[3 * nb + a], g[3 * i + a]);
            d2 = __fadd_rn(d2, __fmul_rn(dd, dd));
        }
        w = __fadd_rn(w, d2);
    }
    s[i] = w;
}

// d_i += sum_k (f_nk - f_i)/sqrt(s_i)     :474-491
__global__ void k_lw_func(const float *__restrict__ f, const float *__restrict__ s, const int *__restrict__ nbrT,
                          const int *__restrict__ valence, int M, float *__restrict__ d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int n = valence[i];
    const float w = s[i];
    if (n == 0 || !(w > 0.f)) return;
    const float r = __fsqrt_rn(w);
    float acc[3] = {d[3 * i], d[3 * i + 1], d[3 * i + 2]};
    for (int k = 0; k < n; ++k) {
        const int nb = nbrT[(size_t)k * M + i];
        for (int a = 0; a < 3; ++a) acc[a] = __fadd_rn(acc[a], __fdiv_rn(__fsub_rn(f[3 * nb + a], f[3 * i + a]), r));
    }
    d[3 * i] = acc[0]; d[3 * i + 1] = acc[1]; d[3 * i + 2] = acc[2];
}

// d_t += (f_i - f_t)/sqrt(s_i), sources ascending     :685-704
__global__ void k_lhw_func(const float *__restrict__ f, const float *__restrict__ s, const int *__restrict__ toff,
                           const int *__restrict__ tsrc, int M, float *__restrict__ d) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    float acc[3] = {d[3 * t], d[3 * t + 1], d[3 * t + 2]};
    for (int e = toff[t]; e < toff[t + 1]; ++e) {
        const int i = tsrc[e];
        const float w = s[i];
        if (!(w > 0.f)) continue;
        const float r = __fsqrt_rn(w);
        for (int a = 0; a < 3; ++a) acc[a] = __fadd_rn(acc[a], __fdiv_rn(__fsub_rn(f[3 * i + a], f[3 * t + a]), r));
    }
    d[3 * t] = acc[0]; d[3 * t + 1] = acc[1]; d[3 * t + 2] = acc[2];
}

// out_i = 1/sqrt(s_i + 1) (0 if s_i == 0), x3; rows of invalid vertices untouched     :500-549
__global__ void k_area_weights(const float *__restrict__ s, const int *__restrict__ valence, int M, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M || valence[i] == 0) return;
    const float w = s[i];
    const float r = (w > 0.f) ? (float)(1.0 / (double)__fsqrt_rn(__fadd_rn(w, 1.0f))) : 0.f;
    out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = r;
}

struct RingTmp {
    float *f = nullptr, *g = nullptr, *d = nullptr, *s = nullptr;
    int *toff = nullptr, *tsrc = nullptr;
};

}  // namespace

static int build_transpose(nw_ctx *h, int **toff_out, int **tsrc_out) {
    // in-neighbour CSR with ascending sources; rebuilt per call (these operators are off the default path)
    cudaStream_t s = h->stream;
    const int M = h->M;
    int *off = nullptr, *tgt = nullptr, *src = nullptr, *cnt = nullptr, *toff = nullptr;
    unsigned long long *keys = nullptr, *keys2 = nullptr;
    NW_CHECK(nw_alloc(h, &off, (size_t)M + 1));
    NW_CUDA(cudaMemsetAsync(off, 0, sizeof(int) * (M + 1), s));
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, h->valence, off, M, s);
    if (tmp > h->cub_tmp_bytes) { NW_CHECK(nw_alloc(h, (char **)&h->cub_tmp, tmp)); h->cub_tmp_bytes = tmp; }
    NW_CUDA(cub::DeviceScan::ExclusiveSum(h->cub_tmp, tmp, h->valence, off, M, s));
    int lo = 0, lv = 0;
    NW_CUDA(cudaMemcpyAsync(&lo, off + M - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    NW_CUDA(cudaMemcpyAsync(&lv, h->valence + M - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    NW_CUDA(cudaStreamSynchronize(s));
    const int E = lo + lv;
    NW_CHECK(nw_alloc(h, &keys, (size_t)E + 1)); NW_CHECK(nw_alloc(h, &keys2, (size_t)E + 1));
    NW_CHECK(nw_alloc(h, &tgt, (size_t)E + 1)); NW_CHECK(nw_alloc(h, &src, (size_t)E + 1));
    NW_CHECK(nw_alloc(h, &cnt, (size_t)M + 1)); NW_CHECK(nw_alloc(h, &toff, (size_t)M + 1));
    k_edge_keys<<<nw_grid(M, 256), 256, 0, s>>>(h->nbrT, h->valence, M, off, keys);
    NW_LAUNCH_CHECK();
    if (E) {
        tmp = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys, keys2, E, 0, 64, s);
        if (tmp > h->cub_tmp_bytes) { NW_CHECK(nw_alloc(h, (char **)&h->cub_tmp, tmp)); h->cub_tmp_bytes = tmp; }
        NW_CUDA(cub::DeviceRadixSort::SortKeys(h->cub_tmp, tmp, keys, keys2, E, 0, 64, s));
        k_split_keys<<<nw_grid(E, 256), 256, 0, s>>>(keys2, E, tgt, src);
        NW_LAUNCH_CHECK();
    }
    NW_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (M + 1), s));
    if (E) { k_count_targets<<<nw_grid(E, 256), 256, 0, s>>>(tgt, E, cnt); NW_LAUNCH_CHECK(); }
    tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, toff, M + 1, s);
    if (tmp > h->cub_tmp_bytes) { NW_CHECK(nw_alloc(h, (char **)&h->cub_tmp, tmp)); h->cub_tmp_bytes = tmp; }
    NW_CUDA(cub::DeviceScan::ExclusiveSum(h->cub_tmp, tmp, cnt, toff, M + 1, s));
    NW_CUDA(cudaStreamSynchronize(s));
    nw_free(&off); nw_free(&tgt); nw_free(&cnt); nw_free(&keys); nw_free(&keys2);
    *toff_out = toff; *tsrc_out = src;
    h->launches += 6;
    return NW_OK;
}

// kind: 0 l, 1 lh, 2 lw, 3 lhw, 4 area weights
static int ring_op(nw_ctx *h, int kind, const float *f, const float *ref, float *out) {
    if (!h) return NW_ERR_ARG;
    NW_ARG(h->M > 0, "ring operator: no topology");
    NW_ARG(out && (f || kind == 4) && (ref || kind < 2), "ring operator: NULL array");
    NW_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int M = h->M, B = 256, G = nw_grid(M, 256);
    RingTmp t;
    int rc = NW_OK;
    auto cleanup = [&]() { nw_free(&t.f); nw_free(&t.g); nw_free(&t.d); nw_free(&t.s); nw_free(&t.toff); nw_free(&t.tsrc); };
#define NWX(x) do { rc = (x); if (rc != NW_OK) { cleanup(); return rc; } } while (0)
#define NWC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { h->err = std::string(#x) + ": " + cudaGetErrorString(e_); cleanup(); return NW_ERR_CUDA; } } while (0)
    NWX(nw_alloc(h, &t.d, (size_t)3 * M));
    NWC(cudaMemcpyAsync(t.d, out, sizeof(float) * 3 * M, cudaMemcpyHostToDevice, s));
    if (f) { NWX(nw_alloc(h, &t.f, (size_t)3 * M)); NWC(cudaMemcpyAsync(t.f, f, sizeof(float) * 3 * M, cudaMemcpyHostToDevice, s)); }
    if (kind >= 2) {
        NWX(nw_alloc(h, &t.g, (size_t)3 * M)); NWX(nw_alloc(h, &t.s, (size_t)M));
        NWC(cudaMemcpyAsync(t.g, ref, sizeof(float) * 3 * M, cudaMemcpyHostToDevice, s));
        k_ring_sq<<<G, B, 0, s>>>(t.g, h->nbrT, h->valence, M, t.s);
        h->launches++;
    }
    if (kind == 1 || kind == 3) NWX(build_transpose(h, &t.toff, &t.tsrc));
    switch (kind) {
        case 0: k_l_func<<<G, B, 0, s>>>(t.f, h->nbrT, h->valence, M, t.d); break;
        case 1: k_lh_func<<<G, B, 0, s>>>(t.f, t.toff, t.tsrc, h->valence, M, t.d); break;
        case 2: k_lw_func<<<G, B, 0, s>>>(t.f, t.s, h->nbrT, h->valence, M, t.d); break;
        case 3: k_lhw_func<<<G, B, 0, s>>>(t.f, t.s, t.toff, t.tsrc, M, t.d); break;
        default: k_area_weights<<<G, B, 0, s>>>(t.s, h->valence, M, t.d); break;
    }
    h->launches++;
    NWC(cudaGetLastError());
    NWC(cudaMemcpyAsync(out, t.d, sizeof(float) * 3 * M, cudaMemcpyDeviceToHost, s));
    NWC(cudaStreamSynchronize(s));
    cleanup();
    return NW_OK;
#undef NWX
#undef NWC
}

extern "C" int nw_l_func(nw_ctx *h, const float *f, float *out) { return ring_op(h, 0, f, nullptr, out); }
extern "C" int nw_lh_func(nw_ctx *h, const float *f, float *out) { return ring_op(h, 1, f, nullptr, out); }
extern "C" int nw_lw_func(nw_ctx *h, const float *f, const float *ref, float *out) { return ring_op(h, 2, f, ref, out); }
extern "C" int nw_lhw_func(nw_ctx *h, const float *f, const float *ref, float *out) { return ring_op(h, 3, f, ref, out); }
extern "C" int nw_vertex_area_weights(nw_ctx *h, const float *ref, float *out) { return ring_op(h, 4, nullptr, ref, out); }
