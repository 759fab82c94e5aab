// benchhook.cu -- device-resident single-kernel launches for CUDA-event timing (bench.py roofline leg).
// Not part of the reference surface.  Each named kernel is launched `reps` times on the handle's stream
// with the data of the last nw_search / nw_compute_weights call; events bracket the launches only.
#include <cstring>
#include "common.cuh"

int nw_bench_launch(nw_ctx *h, const char *name);   // sweep.cu
int nw_launch_influence(nw_ctx *h);

extern "C" int nw_bench_kernel(nw_ctx *h, const char *name, int reps, float *ms_per_launch) {
    if (!h) return NW_ERR_ARG;
    NW_ARG(name && reps > 0 && ms_per_launch, "nw_bench_kernel: bad arguments");
    NW_CUDA(cudaSetDevice(h->device));
    if (std::string(name) == "curvature") {
        cudaEvent_t c0, c1;
        NW_CUDA(cudaEventCreate(&c0)); NW_CUDA(cudaEventCreate(&c1));
        int rc = nw_curvature_relaunch(h);
        if (rc == NW_OK) {
            cudaEventRecord(c0, h->stream);
            for (int r = 0; r < reps && rc == NW_OK; ++r) rc = nw_curvature_relaunch(h);
            cudaEventRecord(c1, h->stream);
            cudaEventSynchronize(c1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c0, c1);
            *ms_per_launch = ms / reps;
        }
        cudaEventDestroy(c0); cudaEventDestroy(c1);
        return rc;
    }
    NW_ARG(h->M > 0 && h->weights_valid, "nw_bench_kernel: run nw_search or nw_compute_weights first");
    SolverState s;
    NW_CUDA(cudaMemcpy(&s, h->st, sizeof(s), cudaMemcpyDeviceToHost));
    s.stop = 0; s.nan_flag = 0;
    for (int a = 0; a < 3; ++a) { s.bbox[a] = 0x7fffffff; s.bbox[3 + a] = (int)0x80000000; }
    if (s.n_search < 2) s.n_search = 2;
    NW_CUDA(cudaMemcpy(h->st, &s, sizeof(s), cudaMemcpyHostToDevice));
    NW_CHECK(nw_set_acc_shifts(h));
    cudaEvent_t e0, e1;
    NW_CUDA(cudaEventCreate(&e0)); NW_CUDA(cudaEventCreate(&e1));
    int rc = nw_bench_launch(h, name);   // warm-up + validates the name
    if (rc == NW_OK) {
        cudaEventRecord(e0, h->stream);
        for (int r = 0; r < reps && rc == NW_OK; ++r) rc = nw_bench_launch(h, name);
        cudaEventRecord(e1, h->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_per_launch = ms / reps;
    }
    cudaMemsetAsync(h->acc, 0, sizeof(unsigned long long) * 4 * h->M, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}
