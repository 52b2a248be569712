/*
 * lbad_frame.cu — the two computing functions of the reference's Frame API on the GPU, for frames of any shape
 * (file:line into /root/reference/LBAudioDetective/LBAudioDetectiveFrame.m):
 *   FrameDecompose          Frame.m:113-132   every row, then every column, through DecomposeArray
 *   DecomposeArray          Frame.m:134-153   x /= sqrtf(n); while (n > 1) { n /= 2; (a + b) / sqrtf(2), (a - b) / sqrtf(2) } — integer
 *                                             halving, so for lengths that are not powers of two the tail elements stay as they are
 *   FrameExtractFingerprint Frame.m:165-191   ranks by |value| descending, ties in ascending flat-index order (the stable reading of
 *                                             -sortUsingComparator:); rank i < t sets out[2i] (value > 0) or out[2i+1] (< 0)
 * The extraction pipeline has its own kernels for its 128 x B images (lbad_extract.cu: warp-local Haar, radix select); these serve
 * callers of the Frame API itself (include/LBAudioDetectiveFrame.h), where shapes are arbitrary and calls are single frames: one
 * thread per row / column for the transform (true IEEE divisions, the reference's order of operations), rank by counting over
 * shared-memory tiles for the fingerprint.
 */
#include "lbad_common.cuh"

namespace lbad {

/* Frame.m:134-153 on a strided vector of n elements; tmp: n floats with the same stride */
__device__ __forceinline__ void frame_decompose_array(float* a, const size_t stride, uint32_t n, float* tmp) {
    const float sn = sqrtf((float)n), s2 = sqrtf(2.0f);                         /* sqrtf(inCount): UInt32 -> float, correctly rounded root */
    for (uint32_t i = 0; i < n; i++) a[i * stride] = __fdiv_rn(a[i * stride], sn);
    while (n > 1) {
        n /= 2;
        for (uint32_t i = 0; i < n; i++) {
            const float x0 = a[(2 * i) * stride], x1 = a[(2 * i + 1) * stride];
            tmp[i * stride] = __fdiv_rn(__fadd_rn(x0, x1), s2);
            tmp[(n + i) * stride] = __fdiv_rn(__fsub_rn(x0, x1), s2);
        }
        for (uint32_t i = 0; i < 2 * n; i++) a[i * stride] = tmp[i * stride];
    }
}

/* a, tmp: [rows][cols].  by_columns = 0: thread r transforms row r (Frame.m:114-116); 1: thread c transforms column c (Frame.m:118-131) */
__global__ void frame_decompose_kernel(float* __restrict__ a, float* __restrict__ tmp, const uint32_t rows, const uint32_t cols, const int by_columns) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (!by_columns) { if (i < rows) frame_decompose_array(a + (size_t)i * cols, 1, cols, tmp + (size_t)i * cols); }
    else             { if (i < cols) frame_decompose_array(a + i, cols, rows, tmp + i); }
}

constexpr int RANK_THREADS = 256;

/* One thread per coefficient i: rank = number of coefficients ahead of it in the reference's order — larger |value|, or equal |value|
 * and a lower flat index.  The others stream through shared memory in tiles.  out: 2t bytes, zeroed by the caller. */
__global__ void __launch_bounds__(RANK_THREADS)
frame_rank_kernel(const float* __restrict__ a, const uint32_t n, const uint32_t t, unsigned char* __restrict__ out) {
    __shared__ uint32_t tile[RANK_THREADS];
    const uint32_t i = blockIdx.x * RANK_THREADS + threadIdx.x;
    const float v = i < n ? a[i] : 0.0f;
    const uint32_t key = __float_as_uint(v) & 0x7fffffffu;                      /* |v| as an integer: the order of fabs() for everything but NaN */
    uint32_t rank = 0;
    for (uint32_t j0 = 0; j0 < n; j0 += RANK_THREADS) {
        const uint32_t j = j0 + threadIdx.x;
        tile[threadIdx.x] = j < n ? (__float_as_uint(a[j]) & 0x7fffffffu) : 0u;
        __syncthreads();
        const uint32_t m = min((uint32_t)RANK_THREADS, n - j0);
        for (uint32_t u = 0; u < m; u++) {
            const uint32_t kj = tile[u];
            rank += (kj > key || (kj == key && j0 + u < i)) ? 1u : 0u;
        }
        __syncthreads();
    }
    if (i < n && rank < t) {
        if (v > 0.0f) out[2 * rank] = 1;                                        /* Frame.m:184-186 */
        else if (v < 0.0f) out[2 * rank + 1] = 1;                               /* Frame.m:187-189 */
    }
}

}  // namespace lbad

using namespace lbad;

/* h_a: [rows][cols] in host memory, transformed in place */
extern "C" int lbadcu_frame_decompose_host(float* h_a, uint32_t rows, uint32_t cols) {
    if (!h_a && rows && cols) return LBAD_ERR_ARG;
    if (lbadcu_device_available() != LBAD_OK) { set_error("no CUDA device available (this library has no CPU fallback)"); return LBAD_ERR_NODEVICE; }
    if (rows == 0 || cols == 0) return LBAD_OK;
    const size_t n = (size_t)rows * cols;
    DevBuf<float> d_a, d_tmp;
    LBAD_CUDA_TRY(d_a.alloc(n)); LBAD_CUDA_TRY(d_tmp.alloc(n));
    LBAD_CUDA_TRY(cudaMemcpy(d_a, h_a, n * sizeof(float), cudaMemcpyHostToDevice));
    frame_decompose_kernel<<<(rows + 127) / 128, 128>>>(d_a, d_tmp, rows, cols, 0);
    frame_decompose_kernel<<<(cols + 127) / 128, 128>>>(d_a, d_tmp, rows, cols, 1);
    LBAD_CUDA_TRY(cudaGetLastError());
    LBAD_CUDA_TRY(cudaMemcpy(h_a, d_a, n * sizeof(float), cudaMemcpyDeviceToHost));
    return LBAD_OK;
}

/* h_a: [n] coefficients in flat (row-major) order; h_out: 2t bytes that receive 1 where the reference sets TRUE, 0 elsewhere */
extern "C" int lbadcu_frame_extract_host(const float* h_a, uint32_t n, uint32_t t, unsigned char* h_out) {
    if ((!h_a && n) || (!h_out && t)) return LBAD_ERR_ARG;
    if (lbadcu_device_available() != LBAD_OK) { set_error("no CUDA device available (this library has no CPU fallback)"); return LBAD_ERR_NODEVICE; }
    if (t == 0) return LBAD_OK;
    DevBuf<float> d_a; DevBuf<unsigned char> d_out;
    LBAD_CUDA_TRY(d_a.alloc(n)); LBAD_CUDA_TRY(d_out.alloc((size_t)2 * t));
    LBAD_CUDA_TRY(cudaMemset(d_out, 0, (size_t)2 * t));
    if (n) {
        LBAD_CUDA_TRY(cudaMemcpy(d_a, h_a, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
        frame_rank_kernel<<<(n + RANK_THREADS - 1) / RANK_THREADS, RANK_THREADS>>>(d_a, n, t, d_out);
        LBAD_CUDA_TRY(cudaGetLastError());
    }
    LBAD_CUDA_TRY(cudaMemcpy(h_out, d_out, (size_t)2 * t, cudaMemcpyDeviceToHost));
    return LBAD_OK;
}
