// Small data-format kernels either side of the hot path: kNN lists -> CSR, the unfused SNGNN++ backward, segmented means.
#include "sng_common.cuh"
#include <cub/device/device_scan.cuh>

namespace sng {

__global__ void __launch_bounds__(256) clamp_counts_kernel(const int* __restrict__ cnt, int64_t nq, int k, int* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= nq; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < nq ? min(max(__ldg(cnt + i), 0), k) : 0;
}

// one thread per list slot: slot (r, t) with t < cnt[r] lands at rowptr[r] + t (rank order is kept)
__global__ void __launch_bounds__(256) knn_compact_kernel(const int* __restrict__ rowptr, const int* __restrict__ idx, const float* __restrict__ sim,
                                                         int64_t nq, int k, int* __restrict__ col, float* __restrict__ val) {
    const int64_t total = nq * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / k;
        const int t = (int)(e - r * k);
        const int b = __ldg(rowptr + r);
        if (t < __ldg(rowptr + r + 1) - b) {
            col[b + t] = __ldg(idx + e);
            if (val) val[b + t] = __ldg(sim + e);
        }
    }
}

// g0 = beta g, g1 = (1 - beta) g
__global__ void __launch_bounds__(256) beta_split_kernel(const float* __restrict__ g, const float* __restrict__ beta_p, int64_t n4,
                                                        float* __restrict__ g0, float* __restrict__ g1) {
    const float beta = __ldg(beta_p);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = ldg4(g + 4 * i);
        reinterpret_cast<float4*>(g0)[i] = scale4(v, beta);
        reinterpret_cast<float4*>(g1)[i] = scale4(v, 1.0f - beta);
    }
}

__global__ void __launch_bounds__(256) segment_accum_kernel(const float* __restrict__ val, const int* __restrict__ seg, int64_t ne, int n_seg,
                                                           double* __restrict__ sums, int* __restrict__ counts) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        const int s = __ldg(seg + e);
        if (s >= 0 && s < n_seg) { atomicAdd(sums + s, (double)__ldg(val + e)); atomicAdd(counts + s, 1); }
    }
}

__global__ void __launch_bounds__(256) segment_finish_kernel(const double* __restrict__ sums, const int* __restrict__ counts, int n_seg,
                                                            float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_seg; i += gridDim.x * blockDim.x)
        out[i] = (float)(sums[i] / (double)max(counts[i], 1));
}

static int grid_1d(int64_t items) {
    const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
    const int64_t need = (items + 255) / 256;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace sng

using namespace sng;

extern "C" size_t sng_knn_to_csr_workspace_bytes(int64_t nq) {
    if (nq < 0 || nq >= (1ll << 31) - 1) return 0;
    size_t t = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t, (const int*)nullptr, (int*)nullptr, (int)(nq + 1));
    return t + (size_t)(nq + 1) * 4 + 512;
}

extern "C" int sng_knn_to_csr(const int32_t* idx, const float* sim, const int32_t* cnt, int64_t nq, int top_k, int32_t* rowptr, int32_t* col,
                              float* val, void* workspace, size_t workspace_bytes, void* stream) {
    SNG_REQUIRE(idx && cnt && rowptr && col && workspace && nq >= 0 && top_k >= 1 && nq * (int64_t)top_k < (1ll << 31), "sng_knn_to_csr: bad arguments");
    SNG_REQUIRE(!val || sim, "sng_knn_to_csr: val needs sim");
    if (workspace_bytes < sng_knn_to_csr_workspace_bytes(nq)) { set_error("sng_knn_to_csr: workspace too small"); return SNG_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    int* clamped = reinterpret_cast<int*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    void* temp = reinterpret_cast<uint8_t*>(clamped) + (((size_t)(nq + 1) * 4 + 255) & ~(size_t)255);
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, (const int*)nullptr, (int*)nullptr, (int)(nq + 1));
    clamp_counts_kernel<<<grid_1d(nq + 1), 256, 0, st>>>(cnt, nq, top_k, clamped);
    if (cub::DeviceScan::ExclusiveSum(temp, temp_bytes, clamped, rowptr, (int)(nq + 1), st) != cudaSuccess) return check_launch("sng_knn_to_csr scan");
    if (nq > 0) knn_compact_kernel<<<grid_1d(nq * top_k), 256, 0, st>>>(rowptr, idx, sim, nq, top_k, col, val);
    return check_launch("sng_knn_to_csr");
}

extern "C" int sng_pp_fuse_bwd(const float* out0, const float* out1, const float* g, const float* beta, int64_t n, int64_t c, int64_t ld,
                               const int32_t* rowptr_in, const int32_t* col_in_shift, float* dbeta, float* partials, float* g0, float* dout1,
                               float* dwt, void* stream) {
    SNG_REQUIRE(out0 && out1 && g && beta && rowptr_in && col_in_shift && dbeta && partials && g0 && dout1 && dwt, "sng_pp_fuse_bwd: null pointer");
    SNG_REQUIRE(n >= 0 && c > 0 && c % 4 == 0 && ld == c, "sng_pp_fuse_bwd: rows must be contiguous with a multiple of 4 channels");
    if (n == 0) return SNG_OK;
    if (int rc = sng_pp_beta_grad(out0, out1, g, n * c, dbeta, partials, stream)) return rc;
    beta_split_kernel<<<grid_1d(n * c / 4), 256, 0, (cudaStream_t)stream>>>(g, beta, n * c / 4, g0, dout1);
    if (int rc = check_launch("sng_pp_fuse_bwd split")) return rc;
    return sng_spmm_fwd(g0, n, c, c, rowptr_in, col_in_shift, nullptr, nullptr, nullptr, dwt, c, stream);   // dL/dW^T = A^T g0
}

extern "C" int sng_segment_mean(const float* val, const int32_t* seg, int64_t num, int64_t n_seg, float* out, void* workspace,
                                size_t workspace_bytes, void* stream) {
    SNG_REQUIRE(val && seg && out && workspace && num >= 0 && n_seg > 0 && n_seg < (1ll << 31), "sng_segment_mean: bad arguments");
    SNG_REQUIRE(workspace_bytes >= (size_t)n_seg * 12 + 256, "sng_segment_mean: workspace needs 12 bytes per segment + 256");
    cudaStream_t st = (cudaStream_t)stream;
    double* sums = reinterpret_cast<double*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int* counts = reinterpret_cast<int*>(sums + n_seg);
    if (cudaMemsetAsync(sums, 0, (size_t)n_seg * 12, st) != cudaSuccess) return check_launch("sng_segment_mean memset");
    if (num > 0) segment_accum_kernel<<<grid_1d(num), 256, 0, st>>>(val, seg, num, (int)n_seg, sums, counts);
    segment_finish_kernel<<<grid_1d(n_seg), 256, 0, st>>>(sums, counts, (int)n_seg, out);
    return check_launch("sng_segment_mean");
}
