// pd_gemm_tcgen05.cu — persistent, warp-specialised TF32 GEMM for sm_100a.
//
//   C[M,N] (=|+=) sum_k A(m,k) * B(n,k) (+bias)(+residual) -> act
//
// Pipeline (one CTA per SM, 192 threads):
//   warp 0      : TMA producer  — cp.async.bulk.tensor 2-D tiles (SWIZZLE_128B) into a 6-stage smem ring
//   warp 1      : MMA issuer    — one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (128x128x8)
//                                 with the fp32 accumulator tile in TMEM (double buffered, 2 x 128 columns)
//   warps 2..9  : epilogue      — tcgen05.ld 32x32b.x32 -> registers -> bias/residual/ELU -> swizzled smem box -> TMA store
//                                 (two warps per TMEM lane quarter, alternate 32-column chunks: with a single warp per
//                                  scheduler the dependent FP chains of bias+ELU+rounding had no latency hiding and
//                                  small-K tiles were epilogue-bound, profiles/README.md r02)
// Work units are (m-tile, n-tile, k-split); split-K units add into C with red.global.add.f32.
//
// Operand layouts: both operands may be K-major ([rows][K], K contiguous) or MN-major ([K][rows]);
// the second form lets the backward contractions dX = dY*W and dW = dY^T*X read the forward tensors
// in place (no transposes in HBM).  Smem tiles follow the canonical UMMA layouts
// (K-major:  8-row x 128 B SWIZZLE_128B atoms, SBO = 1024 B;
//  MN-major: 32-element x 4-k SWIZZLE_128B_BASE32B atoms (TMA SWIZZLE_128B_ATOM_32B), LBO = 4096 B between
//            32-wide MN groups, SBO = 512 B between 4-row k groups).
#include "pd_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128;           // UMMA M
constexpr int BN = 128;           // UMMA N
constexpr int BK = 32;            // fp32 elements per k-block = 128 B = one swizzle row
constexpr int UMMA_K = 8;         // tf32: 32 B per instruction
constexpr int STAGES = 5;
constexpr int A_BYTES = BM * BK * 4;   // 16 KB
constexpr int B_BYTES = BN * BK * 4;   // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;   // 256 (power of two)
constexpr int EPI_WARPS = 8;                   // two warps per TMEM lane quarter: they take alternate 32-column chunks
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_STAGING = EPI_WARPS * 2 * 4096;   // per epilogue warp: two 32x32 fp32 swizzled TMA-store boxes
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGING;
// Ring depth of the plain 128x128 instantiation.  -DPD_GEMM_STAGES0=4 (197 KB instead of 229 KB of shared memory, room for
// co-resident CTAs of the step's other branch) was measured against 5 on one box: 23.94 / 23.92 vs 23.94 / 23.94 ms per step —
// no difference either way, so the deeper ring stays.
#ifndef PD_GEMM_STAGES0
#define PD_GEMM_STAGES0 5
#endif
constexpr int STAGES0 = PD_GEMM_STAGES0;
constexpr int SMEM0_BYTES = STAGES0 * STAGE_BYTES + 1024 + 256 + EPI_STAGING;

struct GemmArgs {
    int M, N, K;
    int a_mn, b_mn;
    int num_m, num_n, splits, kb_total, kb_per_split;
    uint32_t mn_lbo, mn_sbo;   // MN-major descriptor strides (bytes)
    // MN-major tiled operands: one 3-D box {32 columns, 32 k-rows, 4 column groups} (16 KB) instead of four 2-D boxes of 4 KB
    // (tools/microbench/tma_box_rate.cu: 4 KB boxes stream at 9.9 B/clk/SM, 16 KB boxes at 25).  a3_on / b3_on: the 3-D map is
    // valid; a3_part / b3_part: index of the one partial column group (MN % 32 != 0; tiles holding it keep the 2-D boxes), or -1.
    int a3_on, b3_on, a3_part, b3_part;
    int k2_full;               // 2-CTA K2 instantiation: number of FULL 128-byte k-chunks (K / 32 or K / 64); a stage that reaches
                               // past them is fetched with the 2-D boxes, which clip at K
    // implicit-GEMM convolution operands (TMA im2col mode on an NHWC tensor, k x k taps, stride 2, no padding):
    //   a_mode 1: A rows = output pixels, K = (tap, channel)            (conv forward / deconv input-gradient)
    //   a_mode 2: A' rows = (tap, channel padded to 32), K = output pixels  (deconv weight gradient)
    //   b_mode 2: B' rows = (tap, channel padded to 32), K = output pixels  (conv weight gradient)
    int a_mode, b_mode;
    int cv_PQ, cv_Q, cv_C, cv_k, cv_cblocks, cv_cpad;
    int f16;                   // operands are fp16 (kind::f16, 64 elements per 128-byte k-block) instead of tf32
    int tma_store;             // C is TMA-addressable: epilogue uses cp.async.bulk.tensor store / reduce
    int extras_on_split0;      // split-K of a plain (non-accumulate) GEMM: split 0 adds bias/residual, C pre-zeroed
    PdEpilogue epi;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1u << 24)) __trap();   // watchdog: a broken pipeline must not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint64_t* bar, void* smem, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(const void* tmap, uint64_t* bar, void* smem, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// MN-major tile whose first column group is gi: may it be fetched as one 3-D box (no partial column group inside)?
__device__ __forceinline__ bool mn3_ok(int on, int part, int gi) { return on && !(part >= gi && part < gi + 4); }
__device__ __forceinline__ void tma_load_im2col(const void* tmap, uint64_t* bar, void* smem, int c, int w, int h, int n,
                                                int off_w, int off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(smem)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n),
          "h"((uint16_t)off_w), "h"((uint16_t)off_h)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
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

// UMMA shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
// layout: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B — the only layout tcgen05 accepts for
// MN-major 32-bit (tf32) operands (cutlass sm100_common.inl:88-93): 32-byte chunks swizzled over 4 k-rows.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

// fp16 output (PD_GEMM_C_F16): one epilogue warp converts PAIRS of 32-column chunks (64 halfs = one 128-byte row per thread),
// stages them in the same 128B-swizzled 32-row box the fp32 path uses and hands the box to the copy engine (tensor map over
// the fp16 matrix, box {64 halfs, 32 rows}).  Called warp-uniformly; `chalf` selects this warp's pairs.
__device__ __forceinline__ void epi_f16_tile(const PdEpilogue& e, const CUtensorMap* tmC, int N, uint32_t tbase, int n0, int rbase,
                                             int nchunk, int chalf, bool extras, uint8_t* stg0, int& sbuf) {
    const int lane = threadIdx.x & 31;
    const int npair = (nchunk + 1) >> 1;
#pragma unroll 1
    for (int cp = chalf; cp < npair; cp += 2) {
        __half2 h[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[32];
            tc_ld_32x32b_x32(tbase + (uint32_t)(cp * 64 + hh * 32), r);
            const int col0 = n0 + cp * 64 + hh * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float x0 = __uint_as_float(r[j]), x1 = __uint_as_float(r[j + 1]);
                if (extras && e.bias) {
                    x0 += (col0 + j < N) ? __ldg(e.bias + col0 + j) : 0.f;
                    x1 += (col0 + j + 1 < N) ? __ldg(e.bias + col0 + j + 1) : 0.f;
                }
                if (e.act == PD_ACT_ELU) { x0 = pd_elu(x0); x1 = pd_elu(x1); }
                h[hh * 16 + (j >> 1)] = __floats2half2_rn(x0, x1);
            }
        }
        uint8_t* buf = stg0 + sbuf * 4096;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        const uint32_t rowaddr = smem_u32(buf) + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                         // SWIZZLE_128B: 16-byte chunk j of row r at j ^ (r & 7)
            const uint4 q = *reinterpret_cast<const uint4*>(&h[4 * j]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((j ^ (lane & 7)) << 4)), "r"(q.x),
                         "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"((uint64_t)tmC), "r"(smem_u32(buf)), "r"(n0 + cp * 64), "r"(rbase) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        sbuf ^= 1;
    }
}

// BIG: the instantiation for pd_conv_gemm modes 2 / 3 (K = output pixels, both operands MN-major): k-blocks of 64 pixels
// (three 64 KB stages, one 4 KB store box per epilogue warp) so that the im2col operand arrives as four 8 KB boxes per
// 64 pixels instead of eight 4 KB boxes — its TMA boxes cost ~415 clk each whatever their size up to 8 KB
// (tools/microbench/tma_box_rate.cu; the r02 ncu of mode 2 showed exactly 4 x 415 clk per k-block).
// VAR 2 (M2): pd_conv_gemm mode 1 with TWO 128-pixel tiles per weight box: one 256-pixel im2col box (32 KB) and one weight box
// per k-block feed two MMAs (two accumulator tiles, all 512 TMEM columns): a TMA box costs ~400 clk + ~16 clk/KB, and with
// N = 96-192 output channels the 128x128 form spent one such box per operand per 262 clk of MMA.
template <int VAR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pd_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmA3,
                    const __grid_constant__ CUtensorMap tmB3, const GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    constexpr bool BIG = VAR == 1, M2 = VAR == 2;
    constexpr bool ONEBUF = BIG || M2;                          // one 4 KB store box per epilogue warp
    constexpr int NST = BIG ? 3 : (M2 ? 4 : STAGES0);           // BIG / M2: NST * STB + staging = STAGES * STAGE_BYTES + EPI_STAGING
    constexpr int STB = BIG ? 2 * STAGE_BYTES : (M2 ? 2 * A_BYTES + B_BYTES : STAGE_BYTES);
    constexpr int ABY = (BIG || M2) ? 2 * A_BYTES : A_BYTES;
    constexpr int MROWS = M2 ? 2 * BM : BM;                     // output rows of one unit
    constexpr int ACOLS = M2 ? 2 * BN : BN;                     // TMEM columns of one accumulator stage
    constexpr int TCOLS = ACC_STAGES * ACOLS;
    constexpr int KSTEPS = (BIG ? 2 : 1) * (BK / UMMA_K);
    constexpr int GSTR = BIG ? 8192 : 4096;                     // bytes between the 32-column groups of an MN-major operand
    constexpr int KROWS = BIG ? 64 : 32;                        // k-rows (pixels) of one MN-major k-block
    static_assert(VAR == 0 || NST * STB + EPI_WARPS * (ONEBUF ? 1 : 2) * 4096 == STAGES * STAGE_BYTES + EPI_STAGING, "shared-memory footprint");
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + NST * STB + EPI_WARPS * (ONEBUF ? 1 : 2) * 4096);
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + STAGES;             // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;         // [ACC_STAGES]
    uint64_t* tempty = bars + 2 * STAGES + ACC_STAGES;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 2 * ACC_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
        if (g.a3_on) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA3) : "memory");
        if (g.b3_on) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB3) : "memory");
        if (g.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmC) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int units = g.num_m * g.num_n * g.splits;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int split = u % g.splits;
                const int tile = u / g.splits;
                const int m0 = (tile / g.num_n) * MROWS;
                const int n0 = (tile % g.num_n) * BN;
                const int kb0 = split * g.kb_per_split;
                const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STB;
                    uint8_t* sb = sa + ABY;
                    mbar_expect_tx(&full[stage], STB);
                    int k0 = kb * ((g.f16 || BIG) ? 2 * BK : BK);
                    if (g.a_mode == 1) {
                        // implicit im2col rows: k-block = 32 channels of one filter tap; pixel tile starts at m0
                        const int tap = kb / g.cv_cblocks, c0 = (kb - tap * g.cv_cblocks) * 32;
                        const int kh = tap / g.cv_k, kw = tap - kh * g.cv_k;
                        const int n_ = m0 / g.cv_PQ, r_ = m0 - n_ * g.cv_PQ;
                        const int p_ = r_ / g.cv_Q, q_ = r_ - p_ * g.cv_Q;
                        tma_load_im2col(&tmA, &full[stage], sa, c0, 2 * q_, 2 * p_, n_, kw, kh);   // 128 pixels x 32 ch
                        k0 = tap * g.cv_C + c0;                       // matching rows of the (tap, channel)-major weight
                    } else if (g.a_mode == 2 || g.b_mode == 2) {
                        ;                                             // handled below (pixel k-blocks)
                    }
                    if (g.a_mode == 2 || g.b_mode == 2) {
                        // K = output pixels: k-block = 32 consecutive pixels starting at kb*32
                        const int pix = kb * KROWS;
                        const int n_ = pix / g.cv_PQ, r_ = pix - n_ * g.cv_PQ;
                        const int p_ = r_ / g.cv_Q, q_ = r_ - p_ * g.cv_Q;
                        const void* tm = g.a_mode == 2 ? (const void*)&tmA : (const void*)&tmB;
                        uint8_t* dst = g.a_mode == 2 ? sa : sb;
                        const int base = g.a_mode == 2 ? m0 : n0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {                 // 4 boxes of 32 (tap, channel) rows x 32 pixels
                            const int idx = base + 32 * j;
                            int tap = idx / g.cv_cpad, c0 = idx - tap * g.cv_cpad;
                            if (tap >= g.cv_k * g.cv_k) { tap = 0; c0 = g.cv_cpad + 32; }        // past the last tap: all-OOB -> zeros
                            const int kh = tap / g.cv_k, kw = tap - kh * g.cv_k;
                            tma_load_im2col(tm, &full[stage], dst + j * GSTR, c0, 2 * q_, 2 * p_, n_, kw, kh);
                        }
                    }
                    if (g.a_mode == 0) {
                        if (!g.a_mn) {
                            tma_load_2d(&tmA, &full[stage], sa, k0, m0);              // box {32 k, 128 m}
                        } else if (mn3_ok(g.a3_on, g.a3_part, m0 >> 5)) {
                            tma_load_3d(&tmA3, &full[stage], sa, 0, k0, m0 >> 5);     // box {32 m, 32 k, 4 groups}
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 32; ++j)                        // box {32 m, 32 k} x 4
                                tma_load_2d(&tmA, &full[stage], sa + j * GSTR, m0 + j * 32, k0);
                        }
                    }
                    if (g.b_mode == 0) {
                        if (!g.b_mn) {
                            tma_load_2d(&tmB, &full[stage], sb, k0, n0);
                        } else if (mn3_ok(g.b3_on, g.b3_part, n0 >> 5)) {
                            tma_load_3d(&tmB3, &full[stage], sb, 0, k0, n0 >> 5);
                        } else {
#pragma unroll
                            for (int j = 0; j < BN / 32; ++j)
                                tma_load_2d(&tmB, &full[stage], sb + j * GSTR, n0 + j * 32, k0);
                        }
                    }
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor (cute InstrDescriptor): c=F32, a=b=TF32, majors, N>>3, M>>4
            const uint32_t fmt = g.f16 ? 0u : 2u;                   // InstrDescriptor a/b format: 0 = F16, 2 = TF32
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)g.a_mn << 15) |
                                   ((uint32_t)g.b_mn << 16) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int split = u % g.splits;
                const int kb0 = split * g.kb_per_split;
                const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(as * ACOLS);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STB);
                    const uint32_t sb = sa + ABY;
#pragma unroll
                    for (int s = 0; s < KSTEPS; ++s) {
                        // MN-major: 8 k-rows per MMA = two 4-row (512 B) swizzle atoms, SBO apart;
                        //           LBO = 4096 B between 32-element MN groups (one TMA box each).
                        const uint64_t ad = g.a_mn ? make_desc(sa + s * 1024, g.mn_lbo, g.mn_sbo, 1)
                                                   : make_desc(sa + s * 32, 16, 1024, 2);
                        const uint64_t bd = g.b_mn ? make_desc(sb + s * 1024, g.mn_lbo, g.mn_sbo, 1)
                                                   : make_desc(sb + s * 32, 16, 1024, 2);
                        if (g.f16) tc_mma_f16(tacc, ad, bd, idesc, (kb > kb0 || s > 0) ? 1u : 0u);
                        else       tc_mma_tf32(tacc, ad, bd, idesc, (kb > kb0 || s > 0) ? 1u : 0u);
                        if (M2)    tc_mma_tf32(tacc + BN, make_desc(sa + A_BYTES + s * 32, 16, 1024, 2), bd, idesc,
                                               (kb > kb0 || s > 0) ? 1u : 0u);     // second 128-pixel tile, same weights
                    }
                    tc_commit(&empty[stage]);          // frees the smem slot when these MMAs retire
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull[as]);                 // accumulator tile complete
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        // TMEM -> registers (thread = accumulator row, 32 consecutive columns) -> bias / residual / ELU / tf32 round
        // in registers -> 128B-swizzled smem box (32 rows x 32 cols) -> TMA store (cp.async.bulk.tensor) or, for
        // accumulate / split-K, TMA reduce-add (cp.reduce.async.bulk.tensor .add): the copy engine does the
        // addressing, coalescing and M/N edge clipping; the warp issues ~60 instructions per 4 KB of output.
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int chalf = (warp - 2) >> 2;             // which of the quarter's two warps: chunks chalf, chalf + 2, ...
        int as = 0; uint32_t aphase = 0;
        const PdEpilogue& e = g.epi;
        uint8_t* stg0 = smem + NST * STB + (warp - 2) * ((ONEBUF ? 1 : 2) * 4096);   // BIG / M2: one store box per warp
        const bool b_vec = e.bias && ((((uintptr_t)e.bias) & 15) == 0);
        float* stgf = reinterpret_cast<float*>(stg0);     // scalar fallback view (pitch 33 floats fits in 8 KB)
        int sbuf = 0;
        const bool r_vec = e.R && ((e.ldr & 3) == 0) && ((((uintptr_t)e.R) & 15) == 0);
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int split = u % g.splits;
            const int tile = u / g.splits;
            const int m0 = (tile / g.num_n) * MROWS;
            const int n0 = (tile % g.num_n) * BN;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
          for (int sub = 0; sub < (M2 ? 2 : 1); ++sub) {                 // M2: the unit's two 128-row accumulator tiles
            const int rbase = m0 + sub * BM + quarter * 32;
            const int row = rbase + lane;
            const uint32_t tbase = tmem_base + (uint32_t)(as * ACOLS + sub * BN) + ((uint32_t)(quarter * 32) << 16);
            const bool extras = !e.accumulate || (g.extras_on_split0 && split == 0);
            const int nchunk = rbase >= g.M ? 0 : min(BN / 32, (g.N - n0 + 31) / 32);   // chunks with real columns (warp-uniform)
            if (e.c_f16) epi_f16_tile(e, &tmC, g.N, tbase, n0, rbase, nchunk, chalf, extras, stg0, sbuf);
#pragma unroll 1
            for (int c = chalf; c < (e.c_f16 ? 0 : nchunk); c += 2) {
                uint32_t r[32];
                tc_ld_32x32b_x32(tbase + (uint32_t)(c * 32), r);     // warp-collective: no divergence above
                const int col0 = n0 + c * 32;
                if (g.tma_store) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (extras) {
                        if (e.bias) {
                            if (b_vec && col0 + 32 <= g.N) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 q = __ldg(reinterpret_cast<const float4*>(e.bias + col0) + j);
                                    v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] += (col0 + j < g.N) ? __ldg(e.bias + col0 + j) : 0.f;
                            }
                        }
                        if (e.R && row < g.M) {
                            const float* rp = e.R + (long)(row / e.r_div) * e.ldr + col0;
                            if (r_vec && col0 + 32 <= g.N) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 q = __ldg(reinterpret_cast<const float4*>(rp) + j);
                                    v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (col0 + j < g.N) v[j] += __ldg(rp + j);
                            }
                        }
                    }
                    if (e.act == PD_ACT_ELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = pd_elu(v[j]);
                    }
                    if (e.dact && row < g.M) {            // ELU backward of the layer below: v *= elu'(saved output)
                        const float* yp = e.dact + (long)row * e.lddact + col0;
                        if (((e.lddact & 3) == 0) && ((((uintptr_t)e.dact) & 15) == 0) && col0 + 32 <= g.N) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 q = __ldg(reinterpret_cast<const float4*>(yp) + j);
                                v[4 * j] *= pd_elu_grad_from_out(q.x); v[4 * j + 1] *= pd_elu_grad_from_out(q.y);
                                v[4 * j + 2] *= pd_elu_grad_from_out(q.z); v[4 * j + 3] *= pd_elu_grad_from_out(q.w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (col0 + j < g.N) v[j] *= pd_elu_grad_from_out(__ldg(yp + j));
                        }
                    }
                    if (e.round_out) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = pd_tf32(v[j]);
                    }
                    uint8_t* buf = stg0 + sbuf * 4096;
                    if (lane == 0) {                                                                 // buffer free again?
                        if (ONEBUF) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        else     asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    }
                    __syncwarp();
                    const uint32_t rowaddr = smem_u32(buf) + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                     // SWIZZLE_128B: 16-byte chunk j of row r at j ^ (r & 7)
                        const uint32_t a = rowaddr + (uint32_t)((j ^ (lane & 7)) << 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                                     "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                     : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        if (e.accumulate)
                            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"((uint64_t)&tmC), "r"(smem_u32(buf)), "r"(col0), "r"(rbase) : "memory");
                        else
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"((uint64_t)&tmC), "r"(smem_u32(buf)), "r"(col0), "r"(rbase) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (e.dbias) {                        // bias gradient of the layer below: column sums of the staged box
                        // (rows past M hold zeros; lane = column: 16-byte chunk (lane >> 2) of row r sits at chunk
                        //  (lane >> 2) ^ (r & 7) of the swizzled box: conflict-free)
                        float sacc = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; ++r)
                            sacc += *reinterpret_cast<const float*>(buf + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + (lane & 3) * 4);
                        if (col0 + lane < g.N) atomicAdd(e.dbias + col0 + lane, sacc);
                    }
                    if (!ONEBUF) sbuf ^= 1;
                } else {
                    // generic fallback (C not TMA-addressable: ldc % 4 != 0, e.g. N = 1 / 18 outputs)
#pragma unroll
                    for (int j = 0; j < 32; ++j) stgf[lane * 33 + j] = __uint_as_float(r[j]);
                    __syncwarp();
                    const int col = col0 + lane;
                    const int rows = min(32, g.M - rbase);
                    if (col < g.N) {
                        const float bv = (extras && e.bias) ? __ldg(e.bias + col) : 0.f;
                        for (int rr = 0; rr < rows; ++rr) {
                            const int rw = rbase + rr;
                            float x = stgf[rr * 33 + lane];
                            float* cp = e.C + (long)rw * e.ldc + col;
                            if (extras) {
                                x += bv;
                                if (e.R) x += __ldg(e.R + (long)(rw / e.r_div) * e.ldr + col);
                            }
                            if (e.accumulate) {
                                atomicAdd(cp, x);
                            } else {
                                if (e.act == PD_ACT_ELU) x = pd_elu(x);
                                if (e.round_out) x = pd_tf32(x);
                                *cp = x;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
          }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // all stores/reductions landed
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS)
                     : "memory");
    }
}

// =====================================================================================================================
// 2-CTA variant (tcgen05 cta_group::2): a thread-block cluster of two CTAs (one TPC) computes a 256 x 256 tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 n-rows); the leader CTA's elected
// thread issues one tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 8) that reads both CTAs' shared memory and writes
// 128 x 256 accumulators into each CTA's TMEM.  Per CTA and k-block the pair moves 32 KB from L2 for 128 x 256 outputs
// — half the L2 traffic per FLOP of the 1-CTA 128 x 128 kernel, which is L2-bandwidth-bound at TF32 operand width.
// =====================================================================================================================
constexpr int BN2 = 256;                       // pair tile N (UMMA N)
constexpr int BNH = 128;                       // B rows staged per CTA
constexpr int STAGES2 = 4;
constexpr int STAGE2_BYTES = A_BYTES + BNH * BK * 4;        // 32 KB
constexpr int TMEM_COLS2 = ACC_STAGES * BN2;   // 512: the whole TMEM
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256 + EPI_STAGING;
constexpr int SMEM2K_BYTES = 3 * 2 * STAGE2_BYTES + 1024 + 256 + EPI_STAGING / 2;   // K2: three 64 KB stages, one store box per warp
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;    // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const void* tmap, uint32_t leader_bar, void* smem, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"((uint64_t)tmap), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_2sm(const void* tmap, uint32_t leader_bar, void* smem, int c, int w, int h, int n,
                                                    int off_w, int off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(smem)), "l"((uint64_t)tmap), "r"(leader_bar), "r"(c), "r"(w), "r"(h), "r"(n),
          "h"((uint16_t)off_w), "h"((uint16_t)off_h)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(const void* tmap, uint32_t leader_bar, void* smem, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem)), "l"((uint64_t)tmap), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {     // arrives on this barrier offset in BOTH CTAs
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2sm(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// K2: both operands K-major and fetched TWO k-blocks at a time by one 3-D box each ({128-byte k-chunk, 128 rows, 2 k-chunks} =
// 32 KB): three 64 KB stages, one store box per epilogue warp.  OPT-IN (PD_GEMM_2CTA_K2=1): alone, the fp16 [2500,6144,2048]
// GEMM goes from 741 to 891 TFLOP/s and [40000,400,3072] from 691 to 827, but the overlapped step LOSES 0.4 ms (24.30 vs 23.89,
// same box, twice) although its serialised phases get shorter (dream 6.79 vs 6.96 ms); with two 64 KB stages in the old 198 KB
// footprint the kernel-level gain is gone (787 vs 749 TFLOP/s, step 24.0 vs 23.9).  Not understood; left off.  r02 ncu of the fp16 [2500,6144,2048] GEMM: tensor pipe 48 %
// active with two 16 KB boxes per 64-wide k-block — the boxes, not the MMAs, set the pace (tools/microbench/tma_box_rate.cu).
template <bool K2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
pd_gemm_tf32_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmA3,
                         const __grid_constant__ CUtensorMap tmB3, const GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int NST2 = K2 ? 3 : STAGES2;
    constexpr int STB2 = K2 ? 2 * STAGE2_BYTES : STAGE2_BYTES;
    constexpr int ABY2 = K2 ? 2 * A_BYTES : A_BYTES;
    constexpr int NBUF = K2 ? 1 : 2;                            // store boxes per epilogue warp
    uint64_t* bars = (uint64_t*)(smem + NST2 * STB2 + EPI_WARPS * NBUF * 4096);
    uint64_t* full = bars;                       // [STAGES2]   (only the leader's are waited on)
    uint64_t* empty = bars + STAGES2;            // [STAGES2]
    uint64_t* tfull = bars + 2 * STAGES2;        // [ACC_STAGES]
    uint64_t* tempty = bars + 2 * STAGES2 + ACC_STAGES;   // (only the leader's are waited on)
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES2 + 2 * ACC_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
        if (g.a3_on) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA3) : "memory");
        if (g.b3_on) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB3) : "memory");
        if (g.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmC) : "memory");
        for (int i = 0; i < STAGES2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS2)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                          // barrier inits + TMEM of both CTAs visible before any cross-CTA traffic
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_m2 = (g.M + 255) / 256;
    const int num_n2 = (g.N + BN2 - 1) / BN2;
    const int units = num_m2 * num_n2 * g.splits;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = pair; u < units; u += npairs) {
                const int split = u % g.splits;
                const int tile = u / g.splits;
                const int m0 = (tile / num_n2) * 256 + (int)rank * BM;        // this CTA's 128 rows of A / C
                const int n0 = (tile % num_n2) * BN2 + (int)rank * BNH;       // this CTA's half of the B tile
                const int kb0 = split * g.kb_per_split;
                const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STB2;
                    uint8_t* sb = sa + ABY2;
                    const uint32_t lbar = smem_u32(&full[stage]) & PEER_MASK;  // the leader's barrier collects both CTAs' bytes
                    if (leader) mbar_expect_tx(&full[stage], 2 * STB2);
                    if (K2) {                                                   // two k-chunks of both operands per box
                        if (2 * kb + 2 <= g.k2_full) {
                            tma_load_3d_2sm(&tmA3, lbar, sa, 0, m0, 2 * kb);
                            tma_load_3d_2sm(&tmB3, lbar, sb, 0, n0, 2 * kb);
                        } else {                                                // the stage with the partial last chunk: 2-D boxes clip at K
                            const int kc = g.f16 ? 2 * BK : BK;
                            tma_load_2d_2sm(&tmA, lbar, sa, 2 * kb * kc, m0);
                            tma_load_2d_2sm(&tmA, lbar, sa + A_BYTES, (2 * kb + 1) * kc, m0);
                            tma_load_2d_2sm(&tmB, lbar, sb, 2 * kb * kc, n0);
                            tma_load_2d_2sm(&tmB, lbar, sb + A_BYTES, (2 * kb + 1) * kc, n0);
                        }
                        if (++stage == NST2) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    int k0 = kb * (g.f16 ? 2 * BK : BK);
                    if (g.a_mode == 1) {
                        // implicit im2col rows (as in the 1-CTA kernel): this CTA's 128 pixels x 32 channels of one filter tap
                        const int tap = kb / g.cv_cblocks, c0 = (kb - tap * g.cv_cblocks) * 32;
                        const int kh = tap / g.cv_k, kw = tap - kh * g.cv_k;
                        const int n_ = m0 / g.cv_PQ, r_ = m0 - n_ * g.cv_PQ;
                        const int p_ = r_ / g.cv_Q, q_ = r_ - p_ * g.cv_Q;
                        tma_load_im2col_2sm(&tmA, lbar, sa, c0, 2 * q_, 2 * p_, n_, kw, kh);
                        k0 = tap * g.cv_C + c0;
                    } else if (!g.a_mn) {
                        tma_load_2d_2sm(&tmA, lbar, sa, k0, m0);
                    } else if (mn3_ok(g.a3_on, g.a3_part, m0 >> 5)) {
                        tma_load_3d_2sm(&tmA3, lbar, sa, 0, k0, m0 >> 5);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 32; ++j) tma_load_2d_2sm(&tmA, lbar, sa + j * 4096, m0 + j * 32, k0);
                    }
                    if (!g.b_mn) {
                        tma_load_2d_2sm(&tmB, lbar, sb, k0, n0);
                    } else if (BNH == 128 && mn3_ok(g.b3_on, g.b3_part, n0 >> 5)) {
                        tma_load_3d_2sm(&tmB3, lbar, sb, 0, k0, n0 >> 5);
                    } else {
#pragma unroll
                        for (int j = 0; j < BNH / 32; ++j) tma_load_2d_2sm(&tmB, lbar, sb + j * 4096, n0 + j * 32, k0);
                    }
                    if (++stage == NST2) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            const uint32_t fmt = g.f16 ? 0u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)g.a_mn << 15) |
                                   ((uint32_t)g.b_mn << 16) | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int u = pair; u < units; u += npairs) {
                const int split = u % g.splits;
                const int kb0 = split * g.kb_per_split;
                const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(as * BN2);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STB2);
                    const uint32_t sb = sa + ABY2;
#pragma unroll
                    for (int s = 0; s < (K2 ? 2 : 1) * (BK / UMMA_K); ++s) {
                        // K2: steps 0-3 read the first k-chunk tile (16 KB), steps 4-7 the second
                        const uint32_t ko = K2 ? (uint32_t)((s >> 2) * A_BYTES + (s & 3) * 32) : (uint32_t)(s * 32);
                        const uint64_t ad = (!K2 && g.a_mn) ? make_desc(sa + s * 1024, g.mn_lbo, g.mn_sbo, 1)
                                                            : make_desc(sa + ko, 16, 1024, 2);
                        const uint64_t bd = (!K2 && g.b_mn) ? make_desc(sb + s * 1024, g.mn_lbo, g.mn_sbo, 1)
                                                            : make_desc(sb + ko, 16, 1024, 2);
                        if (g.f16) tc_mma_f16_2sm(tacc, ad, bd, idesc, (kb > kb0 || s > 0) ? 1u : 0u);
                        else       tc_mma_tf32_2sm(tacc, ad, bd, idesc, (kb > kb0 || s > 0) ? 1u : 0u);
                    }
                    tc_commit_2sm(&empty[stage]);      // frees the stage in BOTH CTAs
                    if (++stage == NST2) { stage = 0; phase ^= 1; }
                }
                tc_commit_2sm(&tfull[as]);             // both CTAs' epilogues may read their accumulator halves
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5, both CTAs; same scheme as the 1-CTA kernel) =====================
        const int quarter = warp & 3;
        const int chalf = (warp - 2) >> 2;
        int as = 0; uint32_t aphase = 0;
        const PdEpilogue& e = g.epi;
        uint8_t* stg0 = smem + NST2 * STB2 + (warp - 2) * (NBUF * 4096);
        const bool b_vec = e.bias && ((((uintptr_t)e.bias) & 15) == 0);
        int sbuf = 0;
        const bool r_vec = e.R && ((e.ldr & 3) == 0) && ((((uintptr_t)e.R) & 15) == 0);
        uint32_t leader_tempty[ACC_STAGES];
#pragma unroll
        for (int i = 0; i < ACC_STAGES; ++i) leader_tempty[i] = smem_u32(&tempty[i]) & PEER_MASK;
        for (int u = pair; u < units; u += npairs) {
            const int split = u % g.splits;
            const int tile = u / g.splits;
            const int m0 = (tile / num_n2) * 256 + (int)rank * BM;
            const int n0 = (tile % num_n2) * BN2;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const int rbase = m0 + quarter * 32;
            const int row = rbase + lane;
            const uint32_t tbase = tmem_base + (uint32_t)(as * BN2) + ((uint32_t)(quarter * 32) << 16);
            const bool extras = !e.accumulate || (g.extras_on_split0 && split == 0);
            const int nchunk = rbase >= g.M ? 0 : min(BN2 / 32, (g.N - n0 + 31) / 32);
            if (e.c_f16) epi_f16_tile(e, &tmC, g.N, tbase, n0, rbase, nchunk, chalf, extras, stg0, sbuf);
#pragma unroll 1
            for (int c = chalf; c < (e.c_f16 ? 0 : nchunk); c += 2) {
                uint32_t r[32];
                tc_ld_32x32b_x32(tbase + (uint32_t)(c * 32), r);
                const int col0 = n0 + c * 32;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (extras) {
                    if (e.bias) {
                        if (b_vec && col0 + 32 <= g.N) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 q = __ldg(reinterpret_cast<const float4*>(e.bias + col0) + j);
                                v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += (col0 + j < g.N) ? __ldg(e.bias + col0 + j) : 0.f;
                        }
                    }
                    if (e.R && row < g.M) {
                        const float* rp = e.R + (long)(row / e.r_div) * e.ldr + col0;
                        if (r_vec && col0 + 32 <= g.N) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 q = __ldg(reinterpret_cast<const float4*>(rp) + j);
                                v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (col0 + j < g.N) v[j] += __ldg(rp + j);
                        }
                    }
                }
                if (e.act == PD_ACT_ELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = pd_elu(v[j]);
                }
                if (e.dact && row < g.M) {                // ELU backward of the layer below (as in the 1-CTA kernel)
                    const float* yp = e.dact + (long)row * e.lddact + col0;
                    if (((e.lddact & 3) == 0) && ((((uintptr_t)e.dact) & 15) == 0) && col0 + 32 <= g.N) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 q = __ldg(reinterpret_cast<const float4*>(yp) + j);
                            v[4 * j] *= pd_elu_grad_from_out(q.x); v[4 * j + 1] *= pd_elu_grad_from_out(q.y);
                            v[4 * j + 2] *= pd_elu_grad_from_out(q.z); v[4 * j + 3] *= pd_elu_grad_from_out(q.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (col0 + j < g.N) v[j] *= pd_elu_grad_from_out(__ldg(yp + j));
                    }
                }
                if (e.round_out) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = pd_tf32(v[j]);
                }
                uint8_t* buf = stg0 + sbuf * 4096;
                if (lane == 0) {
                    if (K2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
                __syncwarp();
                const uint32_t rowaddr = smem_u32(buf) + lane * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t a = rowaddr + (uint32_t)((j ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                                 "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                 : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (e.accumulate)
                        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"((uint64_t)&tmC), "r"(smem_u32(buf)), "r"(col0), "r"(rbase) : "memory");
                    else
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"((uint64_t)&tmC), "r"(smem_u32(buf)), "r"(col0), "r"(rbase) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (e.dbias) {                            // bias gradient: column sums of the staged box (see the 1-CTA kernel)
                    float sacc = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        sacc += *reinterpret_cast<const float*>(buf + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + (lane & 3) * 4);
                    if (col0 + lane < g.N) atomicAdd(e.dbias + col0 + lane, sacc);
                }
                if (!K2) sbuf ^= 1;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_tempty[as]);      // 8 warps x 2 CTAs -> the leader's MMA issuer
            if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    cluster_sync_all();                          // nobody frees TMEM / exits while the peer may still touch it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS2)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D fp32 tensor map: dim0 = contiguous dimension.
int make_map(pd_handle* h, CUtensorMap* tm, const void* base, uint64_t dim0, uint64_t dim1, uint64_t ld_elems,
             uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz, int elt_bytes = 4) {
    cuuint64_t gdim[2] = {dim0, dim1};
    cuuint64_t gstride[1] = {ld_elems * (uint64_t)elt_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)h->encode_tiled)(tm, elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride,
                                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PD_FAIL(h, PD_ERR_ARG, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu ld %llu", (int)r,
                                   (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)ld_elems);
    return PD_OK;
}

// 3-D view of an MN-major fp32 operand [K rows][MN columns, ld]: (32 columns of a group, k, column group) with strides
// (ld * 4 B, 128 B), box {32, 32, 4} — in shared memory the same bytes as four 2-D {32, 32} boxes 4096 B apart.  Only the
// MN / 32 FULL column groups are addressable (a partial last group would read past the row); *on = 0 if there is none.
int make_map3(pd_handle* h, CUtensorMap* tm, const void* base, uint64_t mn, uint64_t k, uint64_t ld_elems, int* on, int* part,
              int krows = BK) {
    const uint64_t groups = mn / 32;
    *part = (mn % 32) ? (int)groups : -1;
    *on = 0;
    if (groups == 0) return PD_OK;
    cuuint64_t gdim[3] = {32, k, groups};
    cuuint64_t gstride[2] = {ld_elems * 4, 128};
    cuuint32_t box[3] = {32, (cuuint32_t)krows, 4};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)h->encode_tiled)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstride, box, estr,
                                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PD_FAIL(h, PD_ERR_ARG, "cuTensorMapEncodeTiled(3-D MN-major) failed (%d): mn %llu k %llu ld %llu", (int)r,
                                   (unsigned long long)mn, (unsigned long long)k, (unsigned long long)ld_elems);
    *on = 1;
    return PD_OK;
}

// 3-D view of a K-major operand [rows][K, ld]: (one 128-byte k-chunk, rows, k-chunks) with strides (ld, 128 B), box
// {chunk, 128 rows, 2 chunks} = two consecutive k-block tiles of the canonical SWIZZLE_128B layout in ONE TMA operation.
int make_map3k(pd_handle* h, CUtensorMap* tm, const void* base, uint64_t rows, uint64_t k, uint64_t ld_elems, int f16) {
    const uint64_t chunk = f16 ? 64 : 32, esz = f16 ? 2 : 4;
    cuuint64_t gdim[3] = {chunk, rows, k / chunk};                    // FULL chunks only (the kernel takes the tail with 2-D boxes)
    cuuint64_t gstride[2] = {ld_elems * esz, 128};
    cuuint32_t box[3] = {(cuuint32_t)chunk, 128, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)h->encode_tiled)(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                                                   (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PD_FAIL(h, PD_ERR_ARG, "cuTensorMapEncodeTiled(3-D K-major) failed (%d): rows %llu k %llu ld %llu", (int)r,
                                   (unsigned long long)rows, (unsigned long long)k, (unsigned long long)ld_elems);
    return PD_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// TMA im2col-mode map over an NHWC fp32 tensor (k x k taps, stride 2, no padding): lower corner 0, upper corner -(k-1).
int make_im2col_map(pd_handle* h, CUtensorMap* tm, const float* base, int NB, int H, int W, int C, int k, int pixels,
                    CUtensorMapSwizzle swz) {
    if (!h->encode_im2col) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            PD_FAIL(h, PD_ERR_DEVICE, "cuTensorMapEncodeIm2col entry point not found");
        h->encode_im2col = p;
    }
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    int lo[2] = {0, 0}, up[2] = {-(k - 1), -(k - 1)};
    cuuint32_t estr[4] = {1, 2, 2, 1};
    CUresult r = ((EncodeIm2colFn)h->encode_im2col)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, gdim, gstr, lo, up, 32,
                                                   (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PD_FAIL(h, PD_ERR_ARG, "cuTensorMapEncodeIm2col failed (%d): %dx%dx%dx%d k=%d", (int)r, NB, H, W, C, k);
    return PD_OK;
}

// Split-K factor for an accumulating GEMM (weight gradients: few output tiles, K = rows of the batch): the factor that
// minimises waves x (k-blocks per unit + per-unit epilogue), waves = ceil(tiles * splits / slots).  Filling the chip once
// (splits = slots / tiles) is not the same thing: 96 tiles on 148 SMs leave a third of the chip idle unsplit, 3 splits
// give 288 units = 1.95 waves of a third of the work each.
int pick_splits(int tiles, int kb_total, int slots, int min_kb) {
    int maxs = kb_total / min_kb;
    if (maxs < 1) maxs = 1;
    if (maxs > 4 * slots) maxs = 4 * slots;
    const double epi_kb = 6.0;                 // a unit's TMEM drain + TMA reduce-add, in k-block times
    int best = 1;
    double best_cost = 1e300;
    for (int sp = 1; sp <= maxs; ++sp) {
        const int kbs = pd_cdiv(kb_total, sp);
        const int eff = pd_cdiv(kb_total, kbs);               // units really created (no empty ones)
        if (eff != sp) continue;
        const long units = (long)tiles * eff;
        const long waves = (units + slots - 1) / slots;
        const double cost = (double)waves * ((double)kbs + epi_kb);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
    }
    return best;
}

}  // namespace

int configure_2cta(pd_handle* h) {
    if (h->gemm2_smem_configured) return PD_OK;
    cudaError_t e2 = cudaFuncSetAttribute(pd_gemm_tf32_2cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(pd_gemm_tf32_2cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2K_BYTES);
    if (e2 != cudaSuccess) PD_FAIL(h, PD_ERR_DEVICE, "cudaFuncSetAttribute(2cta smem=%d/%d): %s", SMEM2_BYTES, SMEM2K_BYTES, cudaGetErrorString(e2));
    h->gemm2_smem_configured = 1;
    return PD_OK;
}

// Implicit-GEMM convolution launcher.  mode 1: C[pixels, N] = im2col(X) * B   (B: [N][K] or, b_mn, [K][N]; K = (tap, c))
//                                      mode 2: C[(tap,cpad), N] += im2col(X)^T * Bt   (Bt stored [pixels][N])
//                                      mode 3: C[M, (tap,cpad)] += At^T * im2col(X)   (At stored [pixels][M])
int pd_conv_gemm_launch(pd_handle* h, int mode, int NB, int H, int W, int C, int k, const float* X, const float* O, long ldo,
                        int o_mn, int ODIM, const PdEpilogue& epi, cudaStream_t stream) {
    PD_REQUIRE(h, (C % 4) == 0 && ((((uintptr_t)X) & 15) == 0), "pd_conv_gemm: C %% 4 and 16-byte alignment required");
    PD_REQUIRE(h, (ldo % 4) == 0 && ((((uintptr_t)O) & 15) == 0), "pd_conv_gemm: operand alignment");
    PD_REQUIRE(h, (epi.ldc % 4) == 0 && ((((uintptr_t)epi.C) & 15) == 0), "pd_conv_gemm: C must be TMA-addressable");
    if (!h->gemm_smem_configured) {
        cudaError_t e = cudaFuncSetAttribute(pd_gemm_tf32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM0_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pd_gemm_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pd_gemm_tf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) PD_FAIL(h, PD_ERR_DEVICE, "cudaFuncSetAttribute(smem=%d): %s", SMEM_BYTES, cudaGetErrorString(e));
        h->gemm_smem_configured = 1;
    }
    const int P = (H - k) / 2 + 1, Q = (W - k) / 2 + 1;
    const long pixels = (long)NB * P * Q;
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.cv_PQ = P * Q; g.cv_Q = Q; g.cv_C = C; g.cv_k = k; g.cv_cblocks = pd_cdiv(C, 32); g.cv_cpad = g.cv_cblocks * 32;
    // modes 2 / 3 (K = pixels): 64-pixel k-blocks, the BIG instantiation (PD_GEMM_CONV_K64=0: the 32-pixel form)
    const bool big = mode != 1 && h->gemm_conv_k64;
    // mode 1: two 128-pixel tiles per weight box (M2 instantiation; PD_GEMM_CONV_M2=0: one, or the 2-CTA kernel)
    const bool m2 = mode == 1 && h->gemm_conv_m2 && pixels > BM;
    const int kpix = big ? 64 : 32;
    g.mn_lbo = big ? 8192 : 4096; g.mn_sbo = 512;
    g.epi = epi; g.tma_store = 1;
    CUtensorMap tmA, tmB, tmC, tmA3, tmB3;
    memset(&tmA3, 0, sizeof(tmA3)); memset(&tmB3, 0, sizeof(tmB3));
    g.a3_part = g.b3_part = -1;
    int rc, M, N;
    if (mode == 1) {
        M = (int)pixels; N = ODIM;
        g.a_mode = 1; g.a_mn = 0; g.b_mode = 0; g.b_mn = o_mn;
        g.kb_total = k * k * g.cv_cblocks;
        rc = make_im2col_map(h, &tmA, X, NB, H, W, C, k, m2 ? 2 * BM : BM, CU_TENSOR_MAP_SWIZZLE_128B); if (rc) return rc;
        const long Ktot = (long)k * k * C;
        if (!o_mn) rc = make_map(h, &tmB, O, (uint64_t)Ktot, (uint64_t)N, (uint64_t)ldo, BK, BN, CU_TENSOR_MAP_SWIZZLE_128B);
        else       rc = make_map(h, &tmB, O, (uint64_t)N, (uint64_t)Ktot, (uint64_t)ldo, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        if (o_mn && h->gemm_mn3) { rc = make_map3(h, &tmB3, O, (uint64_t)N, (uint64_t)Ktot, (uint64_t)ldo, &g.b3_on, &g.b3_part); if (rc) return rc; }
    } else if (mode == 2) {
        M = k * k * g.cv_cpad; N = ODIM;
        g.a_mode = 2; g.a_mn = 1; g.b_mode = 0; g.b_mn = 1;
        g.kb_total = pd_cdiv(pixels, kpix);
        rc = make_im2col_map(h, &tmA, X, NB, H, W, C, k, kpix, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); if (rc) return rc;
        rc = make_map(h, &tmB, O, (uint64_t)N, (uint64_t)pixels, (uint64_t)ldo, 32, kpix, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        if (h->gemm_mn3) { rc = make_map3(h, &tmB3, O, (uint64_t)N, (uint64_t)pixels, (uint64_t)ldo, &g.b3_on, &g.b3_part, kpix); if (rc) return rc; }
    } else {
        M = ODIM; N = k * k * g.cv_cpad;
        g.a_mode = 0; g.a_mn = 1; g.b_mode = 2; g.b_mn = 1;
        g.kb_total = pd_cdiv(pixels, kpix);
        rc = make_map(h, &tmA, O, (uint64_t)M, (uint64_t)pixels, (uint64_t)ldo, 32, kpix, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        if (h->gemm_mn3) { rc = make_map3(h, &tmA3, O, (uint64_t)M, (uint64_t)pixels, (uint64_t)ldo, &g.a3_on, &g.a3_part, kpix); if (rc) return rc; }
        rc = make_im2col_map(h, &tmB, X, NB, H, W, C, k, kpix, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); if (rc) return rc;
    }
    rc = make_map(h, &tmC, epi.C, (uint64_t)N, (uint64_t)M, (uint64_t)epi.ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    g.M = M; g.N = N; g.K = 0;
    g.num_m = pd_cdiv(M, m2 ? 2 * BM : BM); g.num_n = pd_cdiv(N, BN);
    int tiles = g.num_m * g.num_n, splits = 1;
    // mode 1 (K-major im2col rows) on the 2-CTA kernel: a pair covers 256 pixels x 256 output channels, each CTA gathers its own
    // 128 pixels and half of the weight rows — fewer operand bytes per SM than two independent 128x128 tiles (a 128x128 TF32
    // k-block is fed at ~1045 clk by one SM's TMA path against 262 clk of MMA)
    if (mode == 1 && !m2 && h->gemm_2cta && h->gemm_conv_2cta && M >= h->gemm_2cta_min_m) {
        const int tiles2 = pd_cdiv(M, 256) * pd_cdiv(N, BN2);
        const int pairs_avail = h->num_sms / 2;
        g.kb_per_split = g.kb_total; g.splits = 1;
        const int gridp = tiles2 < pairs_avail ? tiles2 : pairs_avail;
        rc = configure_2cta(h); if (rc) return rc;
        pd_gemm_tf32_2cta_kernel<false><<<gridp * 2, NUM_THREADS, SMEM2_BYTES, stream>>>(tmA, tmB, tmC, tmA3, tmB3, g);
        PD_CHECK_LAUNCH(h, "pd_gemm_tf32_2cta_kernel(im2col)");
        return PD_OK;
    }
    if (epi.accumulate) splits = pick_splits(tiles, g.kb_total, h->num_sms, big ? 4 : 8);
    g.kb_per_split = pd_cdiv(g.kb_total, splits);
    g.splits = pd_cdiv(g.kb_total, g.kb_per_split);
    int units = tiles * g.splits;
    int grid = units < h->num_sms ? units : h->num_sms;
    if (big)     pd_gemm_tf32_kernel<1><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, tmC, tmA3, tmB3, g);
    else if (m2) pd_gemm_tf32_kernel<2><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, tmC, tmA3, tmB3, g);
    else         pd_gemm_tf32_kernel<0><<<grid, NUM_THREADS, SMEM0_BYTES, stream>>>(tmA, tmB, tmC, tmA3, tmB3, g);
    PD_CHECK_LAUNCH(h, "pd_gemm_tf32_kernel(im2col)");
    return PD_OK;
}

int pd_gemm_tcgen05_launch(pd_handle* h, int M, int N, int K, const void* A, long lda, int a_mn, const void* B,
                           long ldb, int b_mn, const PdEpilogue& epi, cudaStream_t stream, int f16) {
    PD_REQUIRE(h, !f16 || (!a_mn && !b_mn && (lda % 8) == 0 && (ldb % 8) == 0), "pd_gemm_f16: K-major operands with ld %% 8 == 0 only");
    PD_REQUIRE(h, (lda % 4) == 0 && (ldb % 4) == 0, "pd_gemm(tcgen05): lda/ldb must be multiples of 4 (got %ld, %ld)",
               lda, ldb);
    PD_REQUIRE(h, (((uintptr_t)A) & 15) == 0 && (((uintptr_t)B) & 15) == 0, "pd_gemm(tcgen05): A/B must be 16B aligned");
    if (!h->gemm_smem_configured) {
        cudaError_t e = cudaFuncSetAttribute(pd_gemm_tf32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM0_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pd_gemm_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pd_gemm_tf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) PD_FAIL(h, PD_ERR_DEVICE, "cudaFuncSetAttribute(smem=%d): %s", SMEM_BYTES, cudaGetErrorString(e));
        h->gemm_smem_configured = 1;
    }
    CUtensorMap tmA, tmB, tmC, tmA3, tmB3;
    memset(&tmA3, 0, sizeof(tmA3)); memset(&tmB3, 0, sizeof(tmB3));
    int rc;
    if (f16)        rc = make_map(h, &tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 2 * BK, BM, CU_TENSOR_MAP_SWIZZLE_128B, 2);
    else if (!a_mn) rc = make_map(h, &tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    else       rc = make_map(h, &tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    if (f16)        rc = make_map(h, &tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 2 * BK, BN, CU_TENSOR_MAP_SWIZZLE_128B, 2);
    else if (!b_mn) rc = make_map(h, &tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN, CU_TENSOR_MAP_SWIZZLE_128B);
    else       rc = make_map(h, &tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;

    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K; g.a_mn = a_mn; g.b_mn = b_mn;
    g.num_m = pd_cdiv(M, BM); g.num_n = pd_cdiv(N, BN);
    g.f16 = f16;
    g.kb_total = pd_cdiv(K, f16 ? 2 * BK : BK);
    g.epi = epi;
    g.mn_lbo = 4096; g.mn_sbo = 512;
    g.a3_part = g.b3_part = -1;
    if (h->gemm_mn3 && !f16 && a_mn) { rc = make_map3(h, &tmA3, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, &g.a3_on, &g.a3_part); if (rc) return rc; }
    if (h->gemm_mn3 && !f16 && b_mn) { rc = make_map3(h, &tmB3, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, &g.b3_on, &g.b3_part); if (rc) return rc; }
    if (const char* dbg = getenv("PD_GEMM_MN_DESC")) {   // bring-up aid: "lbo,sbo" in bytes
        unsigned a = 0, b = 0;
        if (sscanf(dbg, "%u,%u", &a, &b) == 2) { g.mn_lbo = a; g.mn_sbo = b; }
    }
    g.tma_store = ((epi.ldc % 4) == 0) && ((((uintptr_t)epi.C) & 15) == 0);
    if (epi.c_f16) {
        PD_REQUIRE(h, !epi.accumulate && !epi.R && !epi.round_out, "pd_gemm: an fp16 output takes bias / activation only");
        PD_REQUIRE(h, (epi.ldc % 8) == 0 && ((((uintptr_t)epi.C) & 15) == 0), "pd_gemm: fp16 output needs ldc %% 8 == 0 (16-byte rows)");
        g.tma_store = 1;
        rc = make_map(h, &tmC, epi.C, (uint64_t)N, (uint64_t)M, (uint64_t)epi.ldc, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, 2);
        if (rc) return rc;
    } else if (g.tma_store) {
        rc = make_map(h, &tmC, epi.C, (uint64_t)N, (uint64_t)M, (uint64_t)epi.ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    } else {
        tmC = tmA;
    }
    int tiles = g.num_m * g.num_n;
    int splits = 1;
    g.extras_on_split0 = 0;
    // Skinny-M layers (the per-timestep RSSM GEMMs, M = B*I = 50) have too few output tiles to pull their
    // weights through more than a handful of SMs: split K over the idle SMs.  C is zeroed, every split adds its
    // partial product with red.global.add, split 0 also adds bias + residual.
    if (!epi.accumulate && !epi.c_f16 && !epi.dact && g.num_m == 1 && epi.act == PD_ACT_NONE && !epi.round_out && tiles * 2 <= h->num_sms &&
        g.kb_total >= 8) {
        int want = h->num_sms / tiles;
        if (const char* ms = getenv("PD_GEMM_SKINNY_MAXSPLIT")) { int v = atoi(ms); if (want > v) want = v; }   // tuning aid
        int max_splits = g.kb_total / 4;
        if (want > max_splits) want = max_splits;
        if (want > 1) {
            splits = want;
            g.extras_on_split0 = 1;
            g.epi.accumulate = 1;
            if (epi.R == epi.C) {
                g.epi.R = nullptr;                       // in-place residual: C already holds it, do not clear
            } else if (!epi.c_zeroed) {
                cudaError_t me = cudaMemset2DAsync(epi.C, (size_t)epi.ldc * 4, 0, (size_t)N * 4, (size_t)M, stream);
                if (me != cudaSuccess) PD_FAIL(h, PD_ERR_LAUNCH, "cudaMemset2DAsync: %s", cudaGetErrorString(me));
            }
        }
    }
    if (epi.accumulate) splits = pick_splits(tiles, g.kb_total, h->num_sms, 8);   // split-K over idle SMs / partial waves
    // Large tiles-rich problems go to the 2-CTA (cta_group::2) 256x256 kernel; it needs a TMA-addressable C.
    PD_REQUIRE(h, !epi.dact || (g.tma_store && !epi.c_f16 && !epi.accumulate), "pd_gemm(actbwd): needs a TMA-addressable fp32 C");
    int use2 = h->gemm_2cta && g.tma_store && M >= h->gemm_2cta_min_m && N >= 256 && !g.extras_on_split0;
    if (use2) {
        int tiles2 = pd_cdiv(M, 256) * pd_cdiv(N, BN2);
        int pairs_avail = h->num_sms / 2;
        if (tiles2 * 2 < pairs_avail && !epi.accumulate) use2 = 0;          // too few pair-tiles to fill the chip
        if (use2) {
            // re-box B for the half tile (128 rows per CTA) — same box as the 1-CTA kernel, so tmB is reused as is
            // K2: both operands K-major -> two k-chunks per 3-D box (the maps travel in the tmA3 / tmB3 slots)
            const int kchunk = f16 ? 2 * BK : BK;                          // elements per 128-byte k-chunk
            const bool k2 = h->gemm_2cta_k2 && !a_mn && !b_mn && !epi.c_f16 && K >= 4 * kchunk;
            if (k2) {
                rc = make_map3k(h, &tmA3, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, f16); if (rc) return rc;
                rc = make_map3k(h, &tmB3, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, f16); if (rc) return rc;
                g.kb_total = pd_cdiv(K, 2 * kchunk);
                g.k2_full = K / kchunk;
            }
            int splits2 = 1;
            if (epi.accumulate) splits2 = pick_splits(tiles2, g.kb_total, pairs_avail, k2 ? 4 : 8);
            g.kb_per_split = pd_cdiv(g.kb_total, splits2);
            g.splits = pd_cdiv(g.kb_total, g.kb_per_split);
            int units2 = tiles2 * g.splits;
            int gridp = units2 < pairs_avail ? units2 : pairs_avail;
            rc = configure_2cta(h); if (rc) return rc;
            if (k2) pd_gemm_tf32_2cta_kernel<true><<<gridp * 2, NUM_THREADS, SMEM2K_BYTES, stream>>>(tmA, tmB, tmC, tmA3, tmB3, g);
            else    pd_gemm_tf32_2cta_kernel<false><<<gridp * 2, NUM_THREADS, SMEM2_BYTES, stream>>>(tmA, tmB, tmC, tmA3, tmB3, g);
            PD_CHECK_LAUNCH(h, "pd_gemm_tf32_2cta_kernel");
            return PD_OK;
        }
    }
    // Tall GEMMs with one n-tile (N <= 128: the 3-channel conv layers' column forms): two 128-row tiles per B box, like mode 1
    // of the convolutions (M2 instantiation: one 256-row A box + one B box feed two MMAs).
    // OPT-IN: alone the MN-major-B case gains, inside the step it does not (23.99 / 23.99 vs 23.90 / 23.92 ms, same box).
    // (gemm_plain_m2: 1 = only with MN-major B — [2250000,48,108] 0.434 -> 0.312 ms; 2 = also K-major B, where the single store
    //  box per warp costs the ELU-heavy [2402500,48,48] conv1 GEMM more than the shared B box saves: 0.197 -> 0.248 ms)
    if (h->gemm_plain_m2 && (b_mn || h->gemm_plain_m2 > 1) && !f16 && !a_mn && N <= BN && M >= 32 * BM && splits == 1 &&
        !epi.accumulate && g.tma_store && !epi.c_f16) {
        rc = make_map(h, &tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, 2 * BM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        g.num_m = pd_cdiv(M, 2 * BM);
        g.kb_per_split = g.kb_total; g.splits = 1;
        const int units2 = g.num_m * g.num_n;
        const int grid2 = units2 < h->num_sms ? units2 : h->num_sms;
        pd_gemm_tf32_kernel<2><<<grid2, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, tmC, tmA3, tmB3, g);
        PD_CHECK_LAUNCH(h, "pd_gemm_tf32_kernel(M2)");
        return PD_OK;
    }
    g.kb_per_split = pd_cdiv(g.kb_total, splits);
    g.splits = pd_cdiv(g.kb_total, g.kb_per_split);   // no empty units
    int units = tiles * g.splits;
    int grid = units < h->num_sms ? units : h->num_sms;
    pd_gemm_tf32_kernel<0><<<grid, NUM_THREADS, SMEM0_BYTES, stream>>>(tmA, tmB, tmC, tmA3, tmB3, g);
    PD_CHECK_LAUNCH(h, "pd_gemm_tf32_kernel");
    return PD_OK;
}
