// pydreamer_b200 — shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/pd_b200.h"

// ---------------------------------------------------------------------------
// Handle: host-side state only (no device allocations, see include/pd_b200.h)
// ---------------------------------------------------------------------------
struct pd_handle {
    int device;
    int num_sms;
    int gemm_impl;            // PD_GEMM_TCGEN05 / PD_GEMM_SIMT
    int max_smem_optin;
    long launches;            // kernels launched through this handle
    char err[512];
    void* encode_tiled;       // cuTensorMapEncodeTiled entry point
    void* encode_im2col;      // cuTensorMapEncodeIm2col entry point (lazy)
    int gemm_smem_configured;
    int gemm_2cta;            // allow the cta_group::2 256x256 kernel for large problems
    int gemm_2cta_min_m;      // smallest M that goes to the 2-CTA kernel (PD_GEMM_2CTA_MINM, default 384: three of four 128-row tiles real)
    int fuse_actbwd;          // ELU backward + bias gradient inside the producing GEMM / col2im (PD_B200_FUSE_ACTBWD=0: separate pass)
    int gemm_2cta_k2;         // 2-CTA kernel, K-major operands: two k-chunks per 3-D TMA box (opt-in: PD_GEMM_2CTA_K2=1)
    int gemm_plain_m2;        // tall plain GEMMs with N <= 128 on the M2 instantiation (opt-in PD_GEMM_PLAIN_M2=1 / 2: faster alone, step 23.99 vs 23.90 ms)
    int gemm_conv_m2;         // pd_conv_gemm mode 1: two 128-pixel tiles per weight box (PD_GEMM_CONV_M2=0 disables)
    int gemm_conv_2cta;       // pd_conv_gemm mode 1 on the 2-CTA kernel (PD_GEMM_CONV_2CTA=0: 1-CTA 128x128 tiles)
    int gemm_conv_k64;        // pd_conv_gemm modes 2 / 3 with 64-pixel k-blocks (PD_GEMM_CONV_K64=0: 32)
    int gemm_mn3;             // MN-major operands as one 3-D TMA box per tile (PD_GEMM_MN3=0: four 2-D boxes, the round-1 form)
    int gemm2_smem_configured;
    int round_ops;            // round tensor-core operands to tf32 (rna) where they are produced
    int k1_configured;        // persistent RSSM kernels: shared-memory opt-in done on THIS handle's device
    int k1_ctas;              // ... and the co-resident grid they launch (one CTA per SM)
    int k1b_configured;
    int k1b_ctas;
};

// Launch wrappers run on the handle's device whatever the caller's current device is (and put it back).
struct PdDeviceGuard {
    int prev;
    bool switched;
    explicit PdDeviceGuard(const pd_handle* h) : prev(-1), switched(false) {
        if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->device) {
            cudaSetDevice(h->device);
            switched = true;
        }
    }
    ~PdDeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

#define PD_FAIL(h, code, ...)                                        \
    do {                                                             \
        if (h) snprintf((h)->err, sizeof((h)->err), __VA_ARGS__);    \
        return (code);                                               \
    } while (0)

#define PD_CHECK_LAUNCH(h, name)                                                  \
    do {                                                                          \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess)                                                   \
            PD_FAIL(h, PD_ERR_LAUNCH, "%s: %s", name, cudaGetErrorString(e__));   \
        (h)->launches++;                                                          \
    } while (0)

#define PD_REQUIRE(h, cond, ...)                                     \
    do {                                                             \
        if (!(cond)) PD_FAIL(h, PD_ERR_ARG, __VA_ARGS__);            \
    } while (0)

static inline int pd_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------
// Device math helpers
// ---------------------------------------------------------------------------
// Round-to-nearest(-away) to TF32 precision (10 explicit mantissa bits).  Producers of
// tensor-core operands apply this so the hardware's operand truncation is exact
// (no systematic shrink of every product) — see DESIGN.md "precision".
__device__ __forceinline__ float pd_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ float pd_round_if(float x, int on) { return on ? pd_tf32(x) : x; }

// ELU(alpha = 1).  exp(x) - 1 for x <= 0 without libm's expm1f (~45 instructions with branches — the GEMM epilogues that fuse
// the activation were bound by it, r02 ncu of the conv1 GEMM): a degree-7 Taylor polynomial where exp(x) - 1 would cancel
// (|x| < 0.25, truncation error < 4e-10) and ex2.approx elsewhere (result magnitude >= 0.22, relative error < 5e-7).
__device__ __forceinline__ float pd_selp(float a, float b, bool p) {   // p ? a : b as ONE select: the compiler otherwise turns
    float r;                                                            // the two-sided expressions below into a branch per element
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\tselp.f32 %0, %1, %2, q;\n\t}" : "=f"(r) : "f"(a), "f"(b), "r"((int)p));
    return r;
}
__device__ __forceinline__ float pd_expm1_nonpos(float x) {
    float p = fmaf(x, 1.f / 5040.f, 1.f / 720.f);
    p = fmaf(p, x, 1.f / 120.f);
    p = fmaf(p, x, 1.f / 24.f);
    p = fmaf(p, x, 1.f / 6.f);
    p = fmaf(p, x, 0.5f);
    p = fmaf(p, x, 1.f);
    p *= x;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return pd_selp(p, e - 1.f, x > -0.25f);
}
// both sides are evaluated on min(x, 0) and selected (r02 ncu of the conv1 GEMM: one BSSY / BSYNC region per element)
__device__ __forceinline__ float pd_elu(float x) { return pd_selp(x, pd_expm1_nonpos(fminf(x, 0.f)), x > 0.f); }
// d ELU / dx expressed through the ELU *output* y (alpha = 1): x>0 -> 1, else exp(x) = y + 1
__device__ __forceinline__ float pd_elu_grad_from_out(float y) { return y > 0.f ? 1.f : y + 1.f; }
__device__ __forceinline__ float pd_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float pd_softplus(float x) {
    // torch.nn.functional.softplus(beta=1, threshold=20)
    return x > 20.f ? x : log1pf(expf(x));
}

__device__ __forceinline__ float pd_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float pd_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (result valid in all threads).
__device__ __forceinline__ float pd_block_sum(float v, float* sh /* >= 33 floats */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = pd_warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
    if (w == 0) {
        r = pd_warp_sum(r);
        if (lane == 0) sh[32] = r;
    }
    __syncthreads();
    return sh[32];
}

// ---------------------------------------------------------------------------
// GEMM epilogue shared by the tcgen05 and the SIMT kernels
// ---------------------------------------------------------------------------
struct PdEpilogue {
    float* C;
    long ldc;
    const float* bias;   // [N] or nullptr
    const float* R;      // residual, row (m / r_div), or nullptr
    long ldr;
    int r_div;
    int act;             // PD_ACT_NONE / PD_ACT_ELU
    int round_out;       // round result to tf32 precision
    int accumulate;      // 0: C = v ; 1: atomicAdd(C, v) (split-K safe, no bias/act)
    int c_zeroed;        // caller cleared C already (lets a skinny-M split-K launch skip its memset)
    int c_f16;           // C is an fp16 matrix (ldc in halfs): the epilogue converts and stores rows with vector stores
    // backward through the ELU that FOLLOWED the layer whose input gradient this GEMM produces (pd_gemm_actbwd):
    // v *= elu'(dact[m, n]) (dact = that layer's saved output), then dbias[n] += sum_m v — what pd_bias_act_bwd does in a
    // separate pass over C
    const float* dact;
    long lddact;
    float* dbias;
};

__device__ __forceinline__ float pd_epi_value(const PdEpilogue& e, int row, int col, float acc) {
    float v = acc;
    if (e.bias) v += __ldg(e.bias + col);
    if (e.R) v += __ldg(e.R + (long)(row / e.r_div) * e.ldr + col);
    if (e.act == PD_ACT_ELU) v = pd_elu(v);
    if (e.round_out) v = pd_tf32(v);
    return v;
}
