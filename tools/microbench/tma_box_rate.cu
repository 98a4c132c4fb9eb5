// tma_box_rate.cu — how fast does one SM's TMA unit deliver 2-D tiled boxes of different shapes / swizzle modes?
// Every CTA (one per SM) streams `iters` stages of `nbox` boxes (box = 32 fp32 columns x `rows` rows) from its own slice of a
// [R, 32*G] fp32 matrix through a 4-stage mbarrier ring; nobody reads the data.  Prints bytes/clk/SM for each variant.
// Motivation: the MN-major operands of pd_gemm (weight gradients, pd_conv_gemm modes 2/3) are staged as eight 4 KB boxes per
// k-block and run at ~1500 clk per k-block where the K-major path (two 16 KB boxes) runs at ~400.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/tma_box_rate tools/microbench/tma_box_rate.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int rows, int nbox, int iters,
                                                       int col_groups, long rows_total, long long* clk_out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    constexpr int ST = 4;
    uint64_t* full = (uint64_t*)(smem + ST * 49152);
    const uint32_t box_bytes = (uint32_t)rows * 128u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < ST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    const long slice = rows_total / gridDim.x / rows * rows;          // my row range
    const long r0 = slice * blockIdx.x;
    long pos = 0;
    for (int it = 0; it < iters + ST; ++it) {
        const int st = it % ST;
        if (it >= ST) {                                                // wait for the stage issued ST iterations ago
            uint32_t done = 0, par = ((it - ST) / ST) & 1;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(s32(full + st)), "r"(par) : "memory");
        }
        if (it < iters) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full + st)), "r"(box_bytes * nbox) : "memory");
            for (int b = 0; b < nbox; ++b) {
                const int col = (b % col_groups) * 32;
                const long row = r0 + (pos % (slice / rows)) * rows;
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(s32(smem + st * 49152 + b * box_bytes)), "l"((uint64_t)&tm), "r"(s32(full + st)), "r"(col), "r"((int)row)
                             : "memory");
                if (b % col_groups == col_groups - 1 || b == nbox - 1) ++pos;
            }
        }
    }
    clk_out[blockIdx.x] = clock64() - t0;
}


// 3-D view of the same matrix: (32 columns of a group, rows, column groups) with strides (row pitch, 128 B): ONE box
// {32, rows, 4} lands in shared memory exactly like four 2-D boxes {32, rows} at 4096-byte (rows = 32) offsets.
__global__ void __launch_bounds__(64, 1) stream3d_kernel(const __grid_constant__ CUtensorMap tm, int rows, int nbox, int iters,
                                                         long rows_total, long long* clk_out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    constexpr int ST = 4;
    uint64_t* full = (uint64_t*)(smem + ST * 49152);
    const uint32_t box_bytes = (uint32_t)rows * 128u * 4u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < ST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    const long slice = rows_total / gridDim.x / rows * rows;
    const long r0 = slice * blockIdx.x;
    long pos = 0;
    for (int it = 0; it < iters + ST; ++it) {
        const int st = it % ST;
        if (it >= ST) {
            uint32_t done = 0, par = ((it - ST) / ST) & 1;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(s32(full + st)), "r"(par) : "memory");
        }
        if (it < iters) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full + st)), "r"(box_bytes * nbox) : "memory");
            for (int b = 0; b < nbox; ++b) {
                const long row = r0 + (pos % (slice / rows)) * rows;
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(s32(smem + st * 49152 + b * box_bytes)), "l"((uint64_t)&tm), "r"(s32(full + st)), "r"(0), "r"((int)row), "r"(0)
                             : "memory");
                ++pos;
            }
        }
    }
    clk_out[blockIdx.x] = clock64() - t0;
}

// correctness of the 3-D box: load rows [row0, row0+32) x 128 columns once as four 2-D boxes and once as one 3-D box, compare
__global__ void check3d_kernel(const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3, int row0, int g0, int* diff) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + 2 * 16384);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(32768) : "memory");
        for (int j = 0; j < 4; ++j)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(smem + j * 4096)), "l"((uint64_t)&tm2), "r"(s32(bar)), "r"((g0 + j) * 32), "r"(row0) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(s32(smem + 16384)), "l"((uint64_t)&tm3), "r"(s32(bar)), "r"(0), "r"(row0), "r"(g0) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(s32(bar)), "r"(0) : "memory");
    }
    __syncthreads();
    int d = 0;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) d += ((uint32_t*)smem)[i] != ((uint32_t*)(smem + 16384))[i];
    atomicAdd(diff, d);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    const int G = 4;                                                  // 128 columns = 4 groups of 32
    struct V { const char* name; int rows, nbox; CUtensorMapSwizzle sw; long total_mb; };
    const V vs[] = {
        {"8 boxes 32x32 fp32 (4 KB), SWIZZLE_128B_ATOM_32B  [pd_gemm MN-major today]", 32, 8, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 2048},
        {"4 boxes 32x64 fp32 (8 KB), SWIZZLE_128B_ATOM_32B", 64, 4, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 2048},
        {"2 boxes 32x128 fp32 (16 KB), SWIZZLE_128B_ATOM_32B", 128, 2, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 2048},
        {"8 boxes 32x32 fp32 (4 KB), SWIZZLE_128B", 32, 8, CU_TENSOR_MAP_SWIZZLE_128B, 2048},
        {"2 boxes 32x128 fp32 (16 KB), SWIZZLE_128B  [pd_gemm K-major]", 128, 2, CU_TENSOR_MAP_SWIZZLE_128B, 2048},
        {"8 boxes 32x32 fp32 (4 KB), ATOM_32B, L2-resident 64 MB", 32, 8, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 64},
        {"2 boxes 32x128 fp32 (16 KB), SWIZZLE_128B, L2-resident 64 MB", 128, 2, CU_TENSOR_MAP_SWIZZLE_128B, 64},
    };
    long long* clk; cudaMalloc(&clk, sms * sizeof(long long));
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152 + 2048);
    for (const V& v : vs) {
        const long rows_total = v.total_mb * 1024L * 1024L / (128 * 4);
        float* buf; cudaMalloc(&buf, rows_total * 128 * 4); cudaMemset(buf, 0, rows_total * 128 * 4);
        CUtensorMap tm;
        cuuint64_t gdim[2] = {128, (cuuint64_t)rows_total}, gstr[1] = {128 * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)v.rows}, es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", v.name, (int)r); cudaFree(buf); continue; }
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) {
            stream_kernel<<<sms, 64, 4 * 49152 + 2048>>>(tm, v.rows, v.nbox, iters, G, rows_total, clk);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
        }
        long long* h = (long long*)malloc(sms * sizeof(long long));
        cudaMemcpy(h, clk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
        const double bytes = (double)iters * v.nbox * v.rows * 128.0;
        printf("%-78s %7.1f clk/stage (32 KB)  %6.2f B/clk/SM  %6.2f TB/s at 1.965 GHz x %d SMs\n", v.name, (double)mx / iters,
               bytes / mx, bytes / mx * 1.965e9 * sms / 1e12, sms);
        free(h); cudaFree(buf);
    }

    {   // 3-D boxes
        const long total_mb = 2048;
        const long rows_total = total_mb * 1024L * 1024L / (128 * 4);
        float* buf; cudaMalloc(&buf, rows_total * 128 * 4);
        // fill with a position code so that a wrong layout shows
        { float* hb = (float*)malloc(1 << 20); for (int i = 0; i < (1 << 18); ++i) hb[i] = (float)(i % 9973) + 0.25f * (i % 7);
          cudaMemset(buf, 0, rows_total * 128 * 4); cudaMemcpy(buf, hb, 1 << 20, cudaMemcpyHostToDevice); free(hb); }
        for (int cols = 128; cols >= 72; cols -= 56) {                // 128 columns (4 full groups) and 72 (2 full groups + 8 columns)
            const int G = cols / 32;
            CUtensorMap tm3, tm2;
            cuuint64_t gdim3[3] = {32, (cuuint64_t)rows_total, (cuuint64_t)G}, gstr3[2] = {128 * 4, 128};
            cuuint32_t box3[3] = {32, 32, 4}, es3[3] = {1, 1, 1};
            CUresult r = enc(&tm3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, gdim3, gstr3, box3, es3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            printf("3-D map (32, rows, %d groups) strides (512 B, 128 B): encode rc=%d\n", G, (int)r);
            if (r != CUDA_SUCCESS) continue;
            cuuint64_t gdim[2] = {(cuuint64_t)(G * 32), (cuuint64_t)rows_total}, gstr[1] = {128 * 4};
            cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
            enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            int* diff; cudaMalloc(&diff, 4); cudaMemset(diff, 0, 4);
            cudaFuncSetAttribute(check3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
            check3d_kernel<<<1, 128, 40000>>>(tm2, tm3, 64, 0, diff);
            check3d_kernel<<<1, 128, 40000>>>(tm2, tm3, 96, G == 4 ? 0 : 0, diff);
            int hd = -1; cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&hd, diff, 4, cudaMemcpyDeviceToHost);
            printf("  one 3-D box vs four 2-D boxes (groups beyond %d zero-filled by both): %d differing words (%s)\n", G, hd, cudaGetErrorString(e));
            if (G == 4) {
                cudaFuncSetAttribute(stream3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152 + 2048);
                const int iters = 2000;
                for (int rep = 0; rep < 2; ++rep) { stream3d_kernel<<<sms, 64, 4 * 49152 + 2048>>>(tm3, 32, 2, iters, rows_total, clk); cudaDeviceSynchronize(); }
                long long* h = (long long*)malloc(sms * sizeof(long long));
                cudaMemcpy(h, clk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
                const double bytes = (double)iters * 2 * 16384.0;
                printf("%-78s %7.1f clk/stage (32 KB)  %6.2f B/clk/SM  %6.2f TB/s\n", "2 boxes (32 x 32 rows x 4 groups) fp32 (16 KB) 3-D, ATOM_32B", (double)mx / iters,
                       bytes / mx, bytes / mx * 1.965e9 * sms / 1e12);
                free(h);
            }
        }
        cudaFree(buf);
    }
    return 0;
}
