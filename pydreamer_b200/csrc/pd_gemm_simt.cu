// pd_gemm_simt.cu — plain fp32 CUDA-core tile GEMM with the same contract as the tcgen05 kernel.
// It exists as the validation arm for tests (PD_GEMM_SIMT): every composite test can be run with
// either implementation, which separates "is the tensor-core pipeline right" from "is the model
// math right".  Not used by the product path (pd_create selects PD_GEMM_TCGEN05).
#include "pd_common.cuh"

namespace {
constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
pd_gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, long sAm, long sAk,
                    const float* __restrict__ B, long sBn, long sBk, PdEpilogue e, int kchunk) {
    __shared__ float As[TK][TM + 1];
    __shared__ float Bs[TK][TN + 1];
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = kbeg; k0 < kend; k0 += TK) {
        for (int i = threadIdx.x; i < TM * TK; i += 256) {
            int mm, kk;
            if (sAk == 1) { kk = i % TK; mm = i / TK; } else { mm = i % TM; kk = i / TM; }
            int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < kend) ? A[(long)m * sAm + (long)k * sAk] : 0.f;
        }
        for (int i = threadIdx.x; i < TN * TK; i += 256) {
            int nn, kk;
            if (sBk == 1) { kk = i % TK; nn = i / TK; } else { nn = i % TN; kk = i / TN; }
            int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < kend) ? B[(long)n * sBn + (long)k * sBk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int row = m0 + ty * 4 + i;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int col = n0 + tx * 4 + j;
            if (col >= N) continue;
            float* c = e.C + (long)row * e.ldc + col;
            if (e.accumulate) atomicAdd(c, acc[i][j]);
            else *c = pd_epi_value(e, row, col, acc[i][j]);
        }
    }
}
}  // namespace

int pd_gemm_simt_launch(pd_handle* h, int M, int N, int K, const float* A, long lda, int a_mn, const float* B,
                        long ldb, int b_mn, const PdEpilogue& epi, cudaStream_t stream) {
    long sAm = a_mn ? 1 : lda, sAk = a_mn ? lda : 1;
    long sBn = b_mn ? 1 : ldb, sBk = b_mn ? ldb : 1;
    int gx = pd_cdiv(N, TN), gy = pd_cdiv(M, TM), gz = 1;
    if (epi.accumulate) {
        int tiles = gx * gy;
        gz = (2 * h->num_sms) / tiles;
        int maxz = pd_cdiv(K, 256);
        if (gz > maxz) gz = maxz;
        if (gz < 1) gz = 1;
    }
    int kchunk = pd_cdiv(pd_cdiv(K, gz), TK) * TK;
    gz = pd_cdiv(K, kchunk);
    dim3 grid(gx, gy, gz);
    pd_gemm_simt_kernel<<<grid, 256, 0, stream>>>(M, N, K, A, sAm, sAk, B, sBn, sBk, epi, kchunk);
    PD_CHECK_LAUNCH(h, "pd_gemm_simt_kernel");
    return PD_OK;
}
