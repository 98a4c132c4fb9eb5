// pd_dbg_im2col.cu — bring-up probe for TMA im2col-mode loads (cuTensorMapEncodeIm2col + cp.async.bulk.tensor.4d.im2col):
// loads ONE [pixels x 32 channels] tile of an NHWC fp32 tensor for filter tap (kh,kw) and writes it (de-swizzled) to global.
#include "pd_common.cuh"

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

namespace {
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void dbg_im2col_kernel(const __grid_constant__ CUtensorMap tm, int c0, int w0, int h0, int n0, int kw, int kh,
                                  int pixels, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + pixels * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(pixels * 128) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
            ::"r"(s32(smem)), "l"((uint64_t)&tm), "r"(s32(bar)), "r"(c0), "r"(w0), "r"(h0), "r"(n0),
              "h"((uint16_t)kw), "h"((uint16_t)kh)
            : "memory");
    }
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s32(bar)) : "memory");
        if (++spins > (1u << 22)) __trap();
    }
    // de-swizzle (SWIZZLE_128B): 16-byte chunk j of row r lives at chunk j ^ (r & 7)
    for (int i = threadIdx.x; i < pixels * 32; i += blockDim.x) {
        int r = i / 32, c = i % 32;
        int chunk = (c / 4) ^ (r & 7);
        out[i] = *reinterpret_cast<float*>(smem + r * 128 + chunk * 16 + (c % 4) * 4);
    }
}
}  // namespace

extern "C" int pd_dbg_im2col_load(pd_handle* h, const float* in, int NB, int H, int W, int C, int k, int lower, int upper,
                                  int c0, int w0, int h0, int n0, int kw, int kh, int pixels, float* out, void* stream) {
    static EncodeIm2colFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            PD_FAIL(h, PD_ERR_DEVICE, "cuTensorMapEncodeIm2col entry point not found");
        fn = (EncodeIm2colFn)p;
    }
    CUtensorMap tm;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    int lo[2] = {lower, lower}, up[2] = {upper, upper};
    cuuint32_t estr[4] = {1, 2, 2, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, gdim, gstr, lo, up, 32, (cuuint32_t)pixels, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PD_FAIL(h, PD_ERR_ARG, "cuTensorMapEncodeIm2col failed (%d)", (int)r);
    int smem = pixels * 128 + 1024 + 64;
    cudaFuncSetAttribute(dbg_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    dbg_im2col_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tm, c0, w0, h0, n0, kw, kh, pixels, out);
    PD_CHECK_LAUNCH(h, "dbg_im2col_kernel");
    return PD_OK;
}
