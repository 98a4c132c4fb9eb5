// pd_mlp_layer.cu — one hidden layer of the reference's MLP (pydreamer/models/common.py:37-65: Linear -> LayerNorm(eps 1e-3)
// -> ELU) as ONE kernel for the 400-wide layers: Y = ELU(LN(A . W^T + b)) with fp16 operands.
//
// Why: the 400-wide MLP layers (reward / terminal / critic / target / actor heads, dreamer.py:207-213, a2c.py:81-113) ran as a
// tcgen05 GEMM that wrote the pre-norm activations to HBM followed by a LayerNorm kernel that read them back and wrote fp32 +
// fp16 outputs (r02 profile: 115 LayerNorm launches, 2.3 ms, 23 % of HBM peak; [40000,400,400] GEMM 38 us + LayerNorm 53 us).
// LayerNorm needs whole rows, so a CTA owns ROW-COMPLETE tiles: 128 rows x all N <= 512 columns of the accumulator = the
// whole TMEM (512 columns x 128 lanes x fp32).  Pipeline (persistent, one CTA per SM):
//   warp 0     TMA producer: A tile [128 x 64 halfs] + the full weight slab [N x 64 halfs] per k-block, 128-byte swizzle
//   warp 1     tcgen05.mma.kind::f16, M = 128, N in one or two instructions (256 + N - 256), accumulator in TMEM
//   warps 2-9  epilogue, thread = row, two warps per TMEM lane quarter split the columns: pass 1 reads the row from TMEM
//              (+bias) for mean / variance (partial sums exchanged through shared memory), pass 2 re-reads it, normalises,
//              applies ELU and stores fp16 (and, for layers whose backward needs them, fp32 x / y / mean / rstd);
//              bias / gamma / beta live in shared memory
#include "pd_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int BM = 128;
constexpr int KB = 64;                          // halfs per k-block (128-byte rows)
constexpr int A_BYTES = BM * 128;               // 16 KB
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;

struct MlpArgs {
    int M, N, Npad, K, nstage, stage_bytes;
    float eps;
    const float *bias, *gamma, *beta;
    __half* y16; long ldy16;
    float* y; long ldy;
    float* x; long ldx;
    float *mean, *rstd;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 24)) __trap();     // watchdog: a broken pipeline must not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint64_t* bar, void* smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B tile descriptor (same bit layout as pd_gemm_tcgen05.cu: SBO = 1024 B between 8-row groups)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((16u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }

__global__ void __launch_bounds__(NUM_THREADS, 1)
pd_mlp_layer_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmB2, const MlpArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + g.nstage * g.stage_bytes);
    uint64_t* full = bars;                       // [nstage]
    uint64_t* empty = bars + 4;                  // [nstage]
    uint64_t* tfull = bars + 8;
    uint64_t* tempty = bars + 9;
    uint32_t* tmem_slot = (uint32_t*)(bars + 10);
    float* prm = (float*)(bars + 16);             // [3][N]: bias, gamma, beta
    float* xch = prm + 3 * g.N;                   // [2 halves][128 rows][2]: partial (sum, sumsq) of a row
    for (int i = threadIdx.x; i < g.N; i += NUM_THREADS) { prm[i] = g.bias[i]; prm[g.N + i] = g.gamma[i]; prm[2 * g.N + i] = g.beta[i]; }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB1) : "memory");
        for (int i = 0; i < g.nstage; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tfull, 1);
        mbar_init(tempty, EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int tiles = (g.M + BM - 1) / BM;
    const int kbs = (g.K + KB - 1) / KB;
    const int N1 = g.Npad < 256 ? g.Npad : 256, N2 = g.Npad - N1;
    const uint32_t stage_tx = (uint32_t)A_BYTES + (uint32_t)g.Npad * 128u;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                for (int kb = 0; kb < kbs; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * g.stage_bytes;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_expect_tx(&full[stage], stage_tx);
                    tma_load_2d(&tmA, &full[stage], sa, kb * KB, tile * BM);
                    tma_load_2d(&tmB1, &full[stage], sb, kb * KB, 0);
                    if (N2 > 0) tma_load_2d(&tmB2, &full[stage], sb + N1 * 128, kb * KB, N1);
                    if (++stage == g.nstage) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id1 = idesc_f16(N1), id2 = idesc_f16(N2 > 0 ? N2 : 16);
            int stage = 0; uint32_t phase = 0, tphase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait(tempty, tphase ^ 1);                 // the epilogue has drained the previous tile
                tc_fence_after();
                for (int kb = 0; kb < kbs; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * g.stage_bytes), sb = sa + A_BYTES;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const uint32_t acc = (kb > 0 || s > 0) ? 1u : 0u;
                        tc_mma_f16(tmem, make_desc(sa + s * 32), make_desc(sb + s * 32), id1, acc);
                        if (N2 > 0) tc_mma_f16(tmem + (uint32_t)N1, make_desc(sa + s * 32), make_desc(sb + N1 * 128 + s * 32), id2, acc);
                    }
                    tc_commit(&empty[stage]);
                    if (++stage == g.nstage) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull);
                tphase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3, half = (warp - 2) >> 2;       // two warps per lane quarter: column halves
        uint32_t tphase = 0;
        const int nchunk = (g.N + 31) / 32;
        const int c_lo = half == 0 ? 0 : (nchunk + 1) / 2, c_hi = half == 0 ? (nchunk + 1) / 2 : nchunk;
        const float invN = 1.f / (float)g.N;
        const float *sbias = prm, *sgam = prm + g.N, *sbet = prm + 2 * g.N;
        const int rloc = quarter * 32 + lane;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            mbar_wait(tfull, tphase);
            tc_fence_after();
            const long row = (long)tile * BM + rloc;
            const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16);
            // ---- pass 1: partial row statistics of (acc + bias) over my columns
            float s = 0.f, q = 0.f;
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
                uint32_t r[32];
                tc_ld_32x32b_x32(tbase + (uint32_t)(c * 32), r);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = c * 32 + j;
                    if (col < g.N) {
                        const float v = __uint_as_float(r[j]) + sbias[col];
                        s += v; q += v * v;
                    }
                }
            }
            xch[(half * BM + rloc) * 2] = s;
            xch[(half * BM + rloc) * 2 + 1] = q;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");      // the two warps of this quarter
            s += xch[((half ^ 1) * BM + rloc) * 2];
            q += xch[((half ^ 1) * BM + rloc) * 2 + 1];
            const float mean = s * invN;
            const float var = fmaxf(q * invN - mean * mean, 0.f);
            const float rstd = 1.0f / sqrtf(var + g.eps);
            const bool live = row < g.M;
            if (live && g.mean && half == 0) { g.mean[row] = mean; g.rstd[row] = rstd; }
            // ---- pass 2: normalise, ELU, store
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
                uint32_t r[32];
                tc_ld_32x32b_x32(tbase + (uint32_t)(c * 32), r);
                const int col0 = c * 32;
                if (!live) continue;                            // (after the warp-collective load)
                float xv[32], yv[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = col0 + j;
                    if (col < g.N) {
                        xv[j] = __uint_as_float(r[j]) + sbias[col];
                        const __half hv = __float2half_rn(pd_elu((xv[j] - mean) * rstd * sgam[col] + sbet[col]));
                        yv[j] = __half2float(hv);
                    } else { xv[j] = 0.f; yv[j] = 0.f; }
                }
                const bool fullc = col0 + 32 <= g.N;
                if (g.y16) {
                    __half* dst = g.y16 + row * g.ldy16 + col0;
                    if (fullc && ((g.ldy16 & 7) == 0)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __half2 h2[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(yv[8 * j + 2 * e], yv[8 * j + 2 * e + 1]);
                            *reinterpret_cast<uint4*>(dst + 8 * j) = *reinterpret_cast<const uint4*>(h2);
                        }
                    } else {
                        for (int j = 0; j < 32; ++j) if (col0 + j < g.N) dst[j] = __float2half_rn(yv[j]);
                    }
                }
                if (g.y) {
                    float* dst = g.y + row * g.ldy + col0;
                    if (fullc && ((g.ldy & 3) == 0)) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(yv[4 * j], yv[4 * j + 1], yv[4 * j + 2], yv[4 * j + 3]);
                    } else {
                        for (int j = 0; j < 32; ++j) if (col0 + j < g.N) dst[j] = yv[j];
                    }
                }
                if (g.x) {
                    float* dst = g.x + row * g.ldx + col0;
                    if (fullc && ((g.ldx & 3) == 0)) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(xv[4 * j], xv[4 * j + 1], xv[4 * j + 2], xv[4 * j + 3]);
                    } else {
                        for (int j = 0; j < 32; ++j) if (col0 + j < g.N) dst[j] = xv[j];
                    }
                }
            }
            tc_fence_before();
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");      // xch is reused by the next tile
            if (lane == 0) mbar_arrive(tempty);
            tphase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int f16_map(pd_handle* h, CUtensorMap* tm, const void* base, long rows, int K, long ld, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)h->encode_tiled)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, gdim, gstride, box, estr,
                                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PD_FAIL(h, PD_ERR_ARG, "pd_mlp_layer_f16: cuTensorMapEncodeTiled failed (%d) for [%ld, %d]", (int)r, rows, K);
    return PD_OK;
}

}  // namespace

extern "C" int pd_mlp_layer_f16(pd_handle* h, int M, int N, int K, const void* A16, long lda, const void* W16, long ldw,
                                const float* bias, const float* gamma, const float* beta, float eps, void* y16, long ldy16,
                                float* y, long ldy, float* x, long ldx, float* mean, float* rstd, void* stream) {
    if (!h) return PD_ERR_ARG;
    PD_REQUIRE(h, M > 0 && K >= 8 && N >= 16 && N <= TMEM_COLS && (N % 16) == 0, "pd_mlp_layer_f16: shape %d %d %d (16 <= N <= 512, N %% 16 == 0)", M, N, K);
    PD_REQUIRE(h, A16 && W16 && bias && gamma && beta && (y16 || y), "pd_mlp_layer_f16: null operand");
    PD_REQUIRE(h, (lda % 8) == 0 && (ldw % 8) == 0 && ((((uintptr_t)A16) | ((uintptr_t)W16)) & 15) == 0, "pd_mlp_layer_f16: operand alignment");
    PD_REQUIRE(h, !mean == !rstd, "pd_mlp_layer_f16: mean and rstd come together");
    MlpArgs g;
    memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.Npad = N; g.K = K; g.eps = eps; g.bias = bias; g.gamma = gamma; g.beta = beta;
    g.y16 = (__half*)y16; g.ldy16 = ldy16; g.y = y; g.ldy = ldy; g.x = x; g.ldx = ldx; g.mean = mean; g.rstd = rstd;
    g.stage_bytes = (A_BYTES + N * 128 + 1023) / 1024 * 1024;
    g.nstage = (h->max_smem_optin - 2048) / g.stage_bytes;
    if (g.nstage > 4) g.nstage = 4;
    PD_REQUIRE(h, g.nstage >= 2, "pd_mlp_layer_f16: N = %d needs %d bytes per stage", N, g.stage_bytes);
    const int N1 = N < 256 ? N : 256, N2 = N - N1;
    CUtensorMap tmA, tmB1, tmB2;
    int rc = f16_map(h, &tmA, A16, M, K, lda, BM);
    if (!rc) rc = f16_map(h, &tmB1, W16, N, K, ldw, N1);
    if (!rc && N2 > 0) rc = f16_map(h, &tmB2, W16, N, K, ldw, N2);
    if (rc) return rc;
    if (N2 == 0) tmB2 = tmB1;
    const int extra = 128 + 3 * N * 4 + 2 * BM * 2 * 4;         // barriers, bias / gamma / beta, row-statistics exchange
    g.nstage = (h->max_smem_optin - 2048 - extra) / g.stage_bytes;
    if (g.nstage > 4) g.nstage = 4;
    PD_REQUIRE(h, g.nstage >= 2, "pd_mlp_layer_f16: N = %d needs %d bytes per stage", N, g.stage_bytes);
    const int smem = g.nstage * g.stage_bytes + 1024 + extra;
    if (!h->mlp_smem_configured) {                            // once per handle: opt in to the device's full shared memory
        cudaError_t e = cudaFuncSetAttribute(pd_mlp_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin);
        if (e != cudaSuccess) PD_FAIL(h, PD_ERR_DEVICE, "pd_mlp_layer_f16: cudaFuncSetAttribute(smem=%d): %s", h->max_smem_optin, cudaGetErrorString(e));
        h->mlp_smem_configured = 1;
    }
    const int tiles = (M + BM - 1) / BM;
    const int grid = tiles < h->num_sms ? tiles : h->num_sms;
    pd_mlp_layer_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB1, tmB2, g);
    PD_CHECK_LAUNCH(h, "pd_mlp_layer_f16");
    return PD_OK;
}
