// pd_rssm_bptt.cu — back-propagation through time of the RSSM posterior unroll as ONE persistent cooperative kernel
// (pd_rssm_unroll_bwd).
//
// Reference semantics: autograd of pydreamer/models/rssm.py:21-78 (time loop) and :125-153 (RSSMCell.forward:
// z_mlp + a_mlp -> in_norm -> ELU -> GRUCell (rnn.py:60-67 -> nn.GRUCell) -> post_mlp_h + post_mlp_e -> post_norm -> ELU ->
// post_mlp -> straight-through one-hot sample), seeded by the gradients of the losses w.r.t. features / posterior logits.
// It replaces, per timestep, the launch chain  cat_st_bwd -> gemm -> ln_elu_bwd -> gemm -> gru_bwd -> gemm x2 -> ln_elu_bwd
// -> gemm  of pydreamer_b200/dreamer.py (_wm_backward) and writes exactly the tensors that chain writes (dpost, dy2, dgi, dgh,
// dx1: the operands of the batched weight-gradient GEMMs that follow) plus the LayerNorm / bias gradients it accumulates.
//
// Structure (B200: 148 SMs, one CTA per SM, cooperative launch):
//   * 8 consumer warps + 1 producer warp.  The producer streams operands with TMA (cp.async.bulk.tensor.2d, 128-byte
//     swizzle) into a 4-stage shared-memory ring guarded by full / empty mbarriers; a stage = one 64-wide k-block of up to
//     six 16-row weight tiles (fp16) and up to four 64-row x 32-float boxes of the gradient operand (fp32).
//     WEIGHT tiles of the next phase are requested BEFORE the grid barrier that separates the phases (they do not depend
//     on it); only the gradient boxes wait for the barrier, so a phase starts with its weights already in shared memory.
//   * contractions: out[rows, batch] = W^T[rows, K] . X[batch, K]^T on the legacy tensor path
//     (mma.sync.m16n8k8 tf32): weight fragments come from fp16 tiles (ldmatrix, exact fp16 -> tf32 unpack), the gradient
//     operand stays fp32 / tf32-rounded — gradients need fp32's exponent range, so fp16 operands are not an option here,
//     and both operands carry 10 mantissa bits exactly like the TF32 tcgen05 GEMMs of the launch chain.
//   * per timestep six dependent phases, separated by grid barriers (one atomic + one polled word in L2):
//       P9+P1  latent-group owners : dz_{t} = dx1_{t+1} W_z (kept in smem) -> straight-through softmax backward + KL term -> dpost_t
//       P2     (row-group, k-slice): dpin = dpost_t W_pm          partial sums over 4 k-slices -> global
//       P3     batch-row owners    : LayerNorm+ELU backward (post_norm)  -> dy2_t, accumulates dgamma / dbeta / db
//       P4     hidden-unit owners  : dh = dy2_t W_ph + dfeat_h + carry ; GRU gate backward -> dgi_t, dgh_t (dh*u kept in smem)
//       P6/7   (row-group, k-slice): dh_{t-1} partials = dgh_t W_hh ; dza partials = dgi_t W_ih   (K = 3D split in 4)
//       P8     batch-row owners    : LayerNorm+ELU backward (in_norm)    -> dx1_t
//     K-split partial sums are added by their consumers (row owners / unit owners), which costs no extra barrier.
#include "pd_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int NCW = 8;                         // consumer warps
constexpr int NCT = 32 * NCW;                  // consumer threads
constexpr int NT = NCT + 32;                   // + producer warp
constexpr int BROWS = 64;                      // batch rows staged per box (B*I <= 64)
constexpr int KB = 64;                         // k per stage: 64 halfs of weights (128 B rows) = two 32-float boxes of X
constexpr int MAXT = 6;                        // weight tiles (16 rows) per stage
constexpr int A_TILE = 16 * 128;               // 2 KB
constexpr int X_BOX = BROWS * 128;             // 8 KB: 64 rows x 32 floats
constexpr int STAGE_BYTES = MAXT * A_TILE + 4 * X_BOX;     // 44 KB
constexpr int NSTAGE = 4;
constexpr int OFF_BAR = NSTAGE * STAGE_BYTES;               // full[NSTAGE], empty[NSTAGE]
constexpr int OFF_SH = OFF_BAR + 128;                       // 64 floats: block reductions
constexpr int OFF_DHC = OFF_SH + 256;                       // [16][BROWS] floats: dh*u of my hidden units (carry to t-1)
constexpr int OFF_DZ = OFF_DHC + 16 * BROWS * 4;            // [16][32] floats: dzin of my (rows, latent group)
constexpr int SMEM_BYTES = OFF_DZ + 16 * 32 * 4;
constexpr int KSPLIT = 4;

struct BwdMaps {
    CUtensorMap wpmT, wphT, whhT, wihT, wzT;   // fp16 [rows][K], box {64 halfs, 16 rows}, SWIZZLE_128B
    CUtensorMap dpost, dy2, dgh, dgi, dx1;     // fp32 [(T*BI)][K], box {32 floats, 64 rows}, SWIZZLE_128B
    CUtensorMap dx1_16;                        // same tensor, box {32 floats, 16 rows}
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();            // a broken pipeline must not hang the GPU
    }
}
__device__ __forceinline__ void tma_box(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s_u32(dst)), "l"((uint64_t)map), "r"(s_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t h_lo(uint32_t v) { return __float_as_uint(__half2float(__ushort_as_half((unsigned short)(v & 0xffffu)))); }
__device__ __forceinline__ uint32_t h_hi(uint32_t v) { return __float_as_uint(__half2float(__ushort_as_half((unsigned short)(v >> 16)))); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

// Block-wide sum over the 256 consumer threads (result valid in all of them).
__device__ __forceinline__ float cons_sum(float v, float* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = pd_warp_sum(v);
    cons_sync();
    if (lane == 0) sh[w] = v;
    cons_sync();
    float r = lane < NCW ? sh[lane] : 0.f;
    return pd_warp_sum(r);
}

// Grid-wide barrier among the consumer threads of all CTAs (monotonic counter, cleared by the host before the launch).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& epoch) {
    cons_sync();
    epoch += 1;
    if (threadIdx.x == 0) {
        const unsigned target = epoch * gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned spins = 0;
        while (ld_acquire(ctr) < target) {
            if (++spins > (1u << 24)) __trap();
        }
        __threadfence();
    }
    cons_sync();
}
// Producer side: wait until barrier number `epoch` has completed, then make what the other CTAs published through the
// generic proxy visible to this thread's TMA (async proxy) reads.
__device__ __forceinline__ void producer_wait_barrier(const unsigned* ctr, unsigned epoch) {
    const unsigned target = epoch * gridDim.x;
    unsigned spins = 0;
    while (ld_acquire(ctr) < target) {
        if (++spins > (1u << 24)) __trap();
    }
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
}

struct Ring {                       // both sides count stages identically: slot = n % NSTAGE, parity = (n / NSTAGE) & 1
    uint8_t* smem;
    uint64_t* full;
    uint64_t* empty;
    uint32_t n;
    __device__ __forceinline__ uint8_t* stage(uint32_t i) const { return smem + (i % NSTAGE) * STAGE_BYTES; }
};

// One contraction job of this CTA for one phase: `ntile` weight tiles out of `wmap` (tile i = rows [row0[i], row0[i]+16)),
// `nkb` k-blocks of 64 starting at column kcol0; gradient operand boxes: X map 0 for tiles with xsel == 0, X map 1 otherwise.
struct Job {
    const CUtensorMap* wmap[MAXT];
    int row0[MAXT];
    int ntile;
    const CUtensorMap* xmap[2];
    int nx;                          // 1 or 2 gradient operands
    int xrow0;                       // row coordinate of the box (t * BI [+ sub-range start])
    int xrows;                       // 64 or 16 rows per box
    int kcol0, nkb;
    int x2_from;                     // second operand only differs from the first for k >= x2_from (else the first is reused)
};

__device__ __forceinline__ bool job_needs_x2(const Job& j, int kb) { return j.nx == 2 && j.kcol0 + (kb + 1) * KB > j.x2_from; }
__device__ __forceinline__ uint32_t job_bytes(const Job& j, int kb) {
    const uint32_t xb = (uint32_t)j.xrows * 128u * 2u;
    return (uint32_t)j.ntile * A_TILE + xb * (job_needs_x2(j, kb) ? 2u : 1u);
}

// Producer: weights of the first stages are requested before the grid barrier `wait_epoch` (0 = no barrier to wait for),
// gradient boxes after it.
__device__ void produce(Ring& ring, const Job& j, const unsigned* ctr, unsigned wait_epoch) {
    const int npre = j.nkb < NSTAGE ? j.nkb : NSTAGE;
    auto weights = [&](int kb) {
        const uint32_t n = ring.n + kb;
        mbar_wait(ring.empty + n % NSTAGE, ((n / NSTAGE) & 1) ^ 1);
        mbar_expect_tx(ring.full + n % NSTAGE, job_bytes(j, kb));
        uint8_t* st = ring.stage(n);
        for (int i = 0; i < j.ntile; ++i) tma_box(j.wmap[i], ring.full + n % NSTAGE, st + i * A_TILE, j.kcol0 + kb * KB, j.row0[i]);
    };
    auto xboxes = [&](int kb) {
        const uint32_t n = ring.n + kb;
        uint8_t* st = ring.stage(n) + MAXT * A_TILE;
        const int kf = j.kcol0 + kb * KB;                           // column (floats) of this k-block
        tma_box(j.xmap[0], ring.full + n % NSTAGE, st, kf, j.xrow0);
        tma_box(j.xmap[0], ring.full + n % NSTAGE, st + X_BOX, kf + 32, j.xrow0);
        if (job_needs_x2(j, kb)) {
            tma_box(j.xmap[1], ring.full + n % NSTAGE, st + 2 * X_BOX, kf, j.xrow0);
            tma_box(j.xmap[1], ring.full + n % NSTAGE, st + 3 * X_BOX, kf + 32, j.xrow0);
        }
    };
    for (int kb = 0; kb < npre; ++kb) weights(kb);
    if (wait_epoch) producer_wait_barrier(ctr, wait_epoch);
    for (int kb = 0; kb < npre; ++kb) xboxes(kb);
    for (int kb = npre; kb < j.nkb; ++kb) { weights(kb); xboxes(kb); }
    ring.n += j.nkb;
}

// Consumer: this warp accumulates TW weight tiles (tile0 ..) x NW8 n8-tiles of batch rows (n8_0 ..) over all k-blocks of
// the job; xsel = which gradient operand its tiles contract with.  Warps without work pass TW = 0 (they still walk the ring).
// acc[i][j][4]: mma C fragment of (tile i, n8-tile j): rows g, g+8 of the tile, batch columns 2t, 2t+1 of the n8-tile.
template <int TW, int NW8>
__device__ void consume(Ring& ring, const Job& j, int tile0, int n8_0, int xsel, bool active, float (&acc)[TW > 0 ? TW : 1][NW8 > 0 ? NW8 : 1][4]) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < (TW > 0 ? TW : 1); ++i)
#pragma unroll
        for (int jn = 0; jn < (NW8 > 0 ? NW8 : 1); ++jn)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][jn][e] = 0.f;
    for (int kb = 0; kb < j.nkb; ++kb) {
        const uint32_t n = ring.n + kb;
        mbar_wait(ring.full + n % NSTAGE, (n / NSTAGE) & 1);
        if (TW > 0 && active) {
            const uint8_t* st = ring.stage(n);
            const bool x2 = job_needs_x2(j, kb);
            const uint8_t* xb = st + MAXT * A_TILE + ((xsel && x2) ? 2 * X_BOX : 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {                    // four k16 steps of the 64-wide block
                uint32_t a[TW > 0 ? TW : 1][4];
#pragma unroll
                for (int i = 0; i < TW; ++i) {
                    const int r = lane & 15;
                    ldsm_x4(s_u32(st + (tile0 + i) * A_TILE + r * 128 + (((ks * 2 + (lane >> 4)) ^ (r & 7)) << 4)), a[i][0],
                            a[i][1], a[i][2], a[i][3]);
                }
                const uint8_t* box = xb + (ks >> 1) * X_BOX;    // two k16 steps per 32-float box
                const int kk0 = (ks & 1) * 16;
#pragma unroll
                for (int jn = 0; jn < NW8; ++jn) {
                    const int row = (n8_0 + jn) * 8 + g;
                    const uint8_t* rp = box + row * 128;
                    const float2 fa = *reinterpret_cast<const float2*>(rp + ((((kk0 + 2 * t) >> 2) ^ (row & 7)) << 4) + ((2 * t) & 3) * 4);
                    const float2 fb = *reinterpret_cast<const float2*>(rp + ((((kk0 + 2 * t + 8) >> 2) ^ (row & 7)) << 4) + ((2 * t) & 3) * 4);
#pragma unroll
                    for (int i = 0; i < TW; ++i) {
                        // even k (2t, 2t+8) from the low halves, odd k (2t+1, 2t+9) from the high halves
                        mma_tf32(acc[i][jn], h_lo(a[i][0]), h_lo(a[i][1]), h_lo(a[i][2]), h_lo(a[i][3]), __float_as_uint(fa.x),
                                 __float_as_uint(fb.x));
                        mma_tf32(acc[i][jn], h_hi(a[i][0]), h_hi(a[i][1]), h_hi(a[i][2]), h_hi(a[i][3]), __float_as_uint(fa.y),
                                 __float_as_uint(fb.y));
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ring.empty + n % NSTAGE);
    }
    ring.n += j.nkb;
}

__device__ __forceinline__ float rnd(float x, int on) { return on ? pd_tf32(x) : x; }

__global__ void __launch_bounds__(NT, 1) rssm_unroll_bwd_kernel(const pd_rssm_bwd_args a, const __grid_constant__ BwdMaps maps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* sh = (float*)(smem + OFF_SH);
    float* dhc = (float*)(smem + OFF_DHC);                  // dhc[r * BROWS + b]
    float* dzs = (float*)(smem + OFF_DZ);                   // dzs[rb * 32 + class]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = warp == NCW;
    const int P = gridDim.x, c = blockIdx.x;
    const int T = a.T, BI = a.BI, D = a.D, Hd = a.Hd, G = a.G, C = a.C, Z = G * C, F = D + Z, D3 = 3 * D;
    const int rnd_on = a.round_out;

    Ring ring;
    ring.smem = smem;
    ring.full = (uint64_t*)(smem + OFF_BAR);
    ring.empty = ring.full + NSTAGE;
    ring.n = 0;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(ring.full + i, 1); mbar_init(ring.empty + i, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 16 * BROWS; i += NT) dhc[i] = 0.f;
    for (int i = tid; i < 16 * 32; i += NT) dzs[i] = 0.f;
    __syncthreads();

    // ---- static ownership
    // k-split phases: CTA = (row group rg, k slice ks); KS2 for P2 (K = Z), KS6 for P6/7 (K = 3D)
    const int KS2 = a.ks2, KS6 = a.ks6;
    const int RG2 = P / KS2, RG6 = P / KS6;
    const int rg2 = c / KS2, ks2 = c % KS2, rg6 = c / KS6, ks6 = c % KS6;
    const bool in2 = rg2 < RG2, in6 = rg6 < RG6;
    const int f2_0 = (int)((long)rg2 * Hd / RG2), f2_1 = (int)((long)(rg2 + 1) * Hd / RG2);          // P2: dpin features
    const int u6_0 = (int)((long)rg6 * D / RG6), u6_1 = (int)((long)(rg6 + 1) * D / RG6);            // P6: dh rows (units)
    const int f6_0 = (int)((long)rg6 * Hd / RG6), f6_1 = (int)((long)(rg6 + 1) * Hd / RG6);          // P7: dza features
    const int u4_0 = (int)((long)c * D / P), u4_1 = (int)((long)(c + 1) * D / P), nu4 = u4_1 - u4_0;  // P4: my hidden units
    const int R = max(1, min(4, P / G));                                                               // P9/P1: CTAs per group
    const int RB = (BI + R - 1) / R;                                                                   // rows per such CTA (<= 16)
    const bool in9 = c < G * R;
    const int g9 = c / R, sub9 = c % R, b9_0 = sub9 * RB, b9_1 = min(BI, b9_0 + RB);
    const int nt2 = in2 ? (f2_1 - f2_0 + 15) / 16 : 0;
    const int nt6h = in6 ? (u6_1 - u6_0 + 15) / 16 : 0, nt6z = in6 ? (f6_1 - f6_0 + 15) / 16 : 0;
    const int kslice2 = Z / KS2, kslice6 = D3 / KS6;

    auto job_p2 = [&](int t) {
        Job j; j.ntile = nt2; j.nx = 1; j.xmap[0] = &maps.dpost; j.xmap[1] = &maps.dpost; j.xrow0 = t * BI; j.xrows = BROWS;
        j.kcol0 = ks2 * kslice2; j.nkb = nt2 ? (kslice2 + KB - 1) / KB : 0; j.x2_from = 1 << 30;
        for (int i = 0; i < MAXT; ++i) { j.wmap[i] = &maps.wpmT; j.row0[i] = f2_0 + 16 * i; }
        return j;
    };
    auto job_p4 = [&](int t) {
        Job j; j.ntile = nu4 > 0 ? 1 : 0; j.nx = 1; j.xmap[0] = &maps.dy2; j.xmap[1] = &maps.dy2; j.xrow0 = t * BI; j.xrows = BROWS;
        j.kcol0 = 0; j.nkb = j.ntile ? (Hd + KB - 1) / KB : 0; j.x2_from = 1 << 30;
        for (int i = 0; i < MAXT; ++i) { j.wmap[i] = &maps.wphT; j.row0[i] = u4_0; }
        return j;
    };
    auto job_p6 = [&](int t) {
        Job j; j.ntile = (nt6h || nt6z) ? 6 : 0; j.nx = 2; j.xmap[0] = &maps.dgh; j.xmap[1] = &maps.dgi; j.xrow0 = t * BI; j.xrows = BROWS;
        j.kcol0 = ks6 * kslice6; j.nkb = j.ntile ? (kslice6 + KB - 1) / KB : 0; j.x2_from = 2 * D;   // dgi == dgh for the r, u gates
        for (int i = 0; i < 4; ++i) { j.wmap[i] = &maps.whhT; j.row0[i] = u6_0 + 16 * i; }
        for (int i = 0; i < 2; ++i) { j.wmap[4 + i] = &maps.wihT; j.row0[4 + i] = f6_0 + 16 * i; }
        return j;
    };
    auto job_p9 = [&](int t) {
        Job j; j.ntile = in9 ? (C + 15) / 16 : 0; j.nx = 1; j.xmap[0] = &maps.dx1_16; j.xmap[1] = &maps.dx1_16;
        j.xrow0 = t * BI + b9_0; j.xrows = 16;
        j.kcol0 = 0; j.nkb = j.ntile ? (Hd + KB - 1) / KB : 0; j.x2_from = 1 << 30;
        for (int i = 0; i < MAXT; ++i) { j.wmap[i] = &maps.wzT; j.row0[i] = g9 * C + 16 * i; }
        return j;
    };

    // ================================================= producer warp =================================================
    if (producer) {
        if (lane == 0) {
            unsigned epoch = 0;
            for (int t = T - 1; t >= 0; --t) {
                // barriers of a step, in order: after P1 (1), after P2 (2), after P3 (3), after P4 (4), after P6/7 (5), after P8 (6)
                { const Job j = job_p2(t); if (j.nkb) produce(ring, j, a.ws_barrier, epoch + 1); }
                { const Job j = job_p4(t); if (j.nkb) produce(ring, j, a.ws_barrier, epoch + 3); }
                { const Job j = job_p6(t); if (j.nkb) produce(ring, j, a.ws_barrier, epoch + 4); }
                if (t > 0) { const Job j = job_p9(t); if (j.nkb) produce(ring, j, a.ws_barrier, epoch + 6); }
                epoch += 6;
            }
        }
        return;
    }

    // ================================================= consumer warps =================================================
    unsigned epoch = 0;
    // LayerNorm / bias gradient accumulators of the batch-row owners (4 features per thread), kept over all timesteps
    float ag2[4] = {0, 0, 0, 0}, ab2[4] = {0, 0, 0, 0}, ax2[4] = {0, 0, 0, 0};
    float ag1[4] = {0, 0, 0, 0}, ab1[4] = {0, 0, 0, 0}, ax1[4] = {0, 0, 0, 0};

    // LayerNorm+ELU backward of one row held 4 features per thread (same formulas as ln_elu_bwd_row_kernel, pd_rowwise.cu)
    auto ln_bwd_row = [&](const float (&dyv)[4], const float* xrow, const float* yrow, const float* gamma, float mean, float rstd,
                          float* dxrow, float (&ag)[4], float (&ab)[4], float (&ax)[4]) {
        float gg[4], xh[4], dxh[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = tid + NCT * i;
            if (f < Hd) {
                gg[i] = dyv[i] * pd_elu_grad_from_out(yrow[f]);
                xh[i] = (xrow[f] - mean) * rstd;
                dxh[i] = gg[i] * gamma[f];
                s1 += dxh[i]; s2 += dxh[i] * xh[i];
            } else { gg[i] = xh[i] = dxh[i] = 0.f; }
        }
        const float c1 = cons_sum(s1, sh) / (float)Hd;
        const float c2 = cons_sum(s2, sh) / (float)Hd;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = tid + NCT * i;
            if (f < Hd) {
                const float d = rstd * (dxh[i] - c1 - xh[i] * c2);
                dxrow[f] = rnd(d, rnd_on);
                ag[i] += gg[i] * xh[i]; ab[i] += gg[i]; ax[i] += d;
            }
        }
    };

    for (int t = T - 1; t >= 0; --t) {
        const bool nxt = t + 1 < T;
        // ---------------- P1 (latent-group owners): straight-through softmax backward + KL term -> dpost_t
        if (in9) {
            for (int rb = warp; rb < b9_1 - b9_0; rb += NCW) {
                const int b = b9_0 + rb;
                const long row = (long)t * BI + b;
                const bool valid = lane < C;
                const long off = (long)g9 * C + lane;
                const float l = valid ? a.post[row * Z + off] : 0.f;
                // group softmax exactly as cat_st_bwd_kernel / cat_sample_kernel (pd_rowwise.cu)
                const float mx = pd_warp_max(valid ? l : -INFINITY);
                const float e = valid ? expf(l - mx) : 0.f;
                const float lse = mx + logf(pd_warp_sum(e));
                const float ln = valid ? l - lse : -INFINITY;
                const float mx2 = pd_warp_max(ln);
                const float e2 = valid ? expf(ln - mx2) : 0.f;
                const float p = e2 / pd_warp_sum(e2);
                float dz = 0.f;
                if (valid) {
                    dz = a.dfeat[row * F + D + off];
                    if (nxt) dz += dzs[rb * 32 + lane] * a.mask[row + BI];
                }
                const float s = pd_warp_sum(valid ? p * dz : 0.f);
                if (valid) {
                    float d = p * (dz - s);
                    d += a.kl_weight * a.w[row] * a.dpost_u[row * Z + off];
                    a.dpost[row * Z + off] = rnd(d, rnd_on);
                }
            }
        }
        grid_barrier(a.ws_barrier, epoch);                                      // (1) dpost_t complete

        // ---------------- P2 (row group x k slice): dpin partials = dpost_t . W_pm
        {
            const Job j = job_p2(t);
            float acc[2][1][4];
            const bool act = warp * 8 < BI;
            consume<2, 1>(ring, j, 0, warp, 0, act && j.nkb > 0, acc);
            if (act && j.nkb > 0) {
                const int g = lane >> 2, tq = lane & 3;
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int f = f2_0 + 16 * i + g + 8 * (e >> 1), b = warp * 8 + 2 * tq + (e & 1);
                        if (f < f2_1 && b < BI) a.ws_part2[((long)ks2 * BI + b) * Hd + f] = acc[i][0][e];
                    }
            }
        }
        grid_barrier(a.ws_barrier, epoch);                                      // (2) dpin partials complete

        // ---------------- P3 (batch-row owners): post_norm LayerNorm+ELU backward -> dy2_t
        if (c < BI) {
            const int b = c;
            const long row = (long)t * BI + b;
            float dyv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int f = tid + NCT * i;
                float s = 0.f;
                if (f < Hd)
                    for (int k = 0; k < KS2; ++k) s += __ldcg(a.ws_part2 + ((long)k * BI + b) * Hd + f);
                dyv[i] = s;
            }
            ln_bwd_row(dyv, a.y2 + row * Hd, a.pin + row * Hd, a.ln2_g, a.m2[row], a.r2[row], a.dy2 + row * Hd, ag2, ab2, ax2);
        }
        grid_barrier(a.ws_barrier, epoch);                                      // (3) dy2_t complete

        // ---------------- P4 (hidden-unit owners): dh = dy2_t . W_ph + dfeat_h + carry ; GRU gate backward -> dgi_t, dgh_t
        {
            const Job j = job_p4(t);
            float acc[1][1][4];
            const bool act = warp * 8 < BI;
            consume<1, 1>(ring, j, 0, warp, 0, act && j.nkb > 0, acc);
            if (act && j.nkb > 0) {
                const int g = lane >> 2, tq = lane & 3;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int r = g + 8 * (e >> 1), b = warp * 8 + 2 * tq + (e & 1), u = u4_0 + r;
                    if (r < nu4 && b < BI) {
                        const long row = (long)t * BI + b;
                        float dh = acc[0][0][e] + a.dfeat[row * F + u];
                        if (nxt) {
                            float carry = dhc[r * BROWS + b];
                            for (int k = 0; k < KS6; ++k) carry += __ldcg(a.ws_part6 + ((long)k * BI + b) * D + u);
                            dh += carry * a.mask[row + BI];
                        }
                        const float* gt = a.gates + row * 4 * D;
                        const float rg = gt[u], ug = gt[D + u], ng = gt[2 * D + u], ghn = gt[3 * D + u];
                        const float hp = a.hin[row * D + u];
                        const float dn_pre = dh * (1.f - ug) * (1.f - ng * ng);
                        const float du_pre = dh * (hp - ng) * ug * (1.f - ug);
                        const float dr_pre = dn_pre * ghn * rg * (1.f - rg);
                        float* gi = a.dgi + row * D3;
                        float* gh = a.dgh + row * D3;
                        const float v0 = rnd(dr_pre, rnd_on), v1 = rnd(du_pre, rnd_on);
                        gi[u] = v0; gi[D + u] = v1; gi[2 * D + u] = rnd(dn_pre, rnd_on);
                        gh[u] = v0; gh[D + u] = v1; gh[2 * D + u] = rnd(dn_pre * rg, rnd_on);
                        dhc[r * BROWS + b] = dh * ug;
                    }
                }
            }
        }
        grid_barrier(a.ws_barrier, epoch);                                      // (4) dgi_t, dgh_t complete

        // ---------------- P6/7 (row group x k slice): dh_{t-1} partials = dgh_t . W_hh ; dza partials = dgi_t . W_ih
        {
            const Job j = job_p6(t);
            float acc[2][4][4];
            // warps 0..5: tile pair (warp % 3) x batch half (warp / 3); tile pairs 0,1 = W_hh^T rows, pair 2 = W_ih^T rows
            const int pair = warp % 3, half = warp / 3;
            const bool act = warp < 6 && j.nkb > 0 && half * 32 < BI;
            consume<2, 4>(ring, j, 2 * pair, 4 * half, pair == 2 ? 1 : 0, act, acc);
            if (act) {
                const int g = lane >> 2, tq = lane & 3;
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int r = 16 * (2 * pair + i) + g + 8 * (e >> 1), b = (4 * half + jn) * 8 + 2 * tq + (e & 1);
                            if (b >= BI) continue;
                            if (pair < 2) {
                                const int u = u6_0 + r;
                                if (u < u6_1) a.ws_part6[((long)ks6 * BI + b) * D + u] = acc[i][jn][e];
                            } else {
                                const int f = f6_0 + r - 64;
                                if (f < f6_1) a.ws_part7[((long)ks6 * BI + b) * Hd + f] = acc[i][jn][e];
                            }
                        }
            }
        }
        grid_barrier(a.ws_barrier, epoch);                                      // (5) partials complete

        // ---------------- P8 (batch-row owners): in_norm LayerNorm+ELU backward -> dx1_t
        if (c < BI) {
            const int b = c;
            const long row = (long)t * BI + b;
            float dyv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int f = tid + NCT * i;
                float s = 0.f;
                if (f < Hd)
                    for (int k = 0; k < KS6; ++k) s += __ldcg(a.ws_part7 + ((long)k * BI + b) * Hd + f);
                dyv[i] = s;
            }
            ln_bwd_row(dyv, a.x1 + row * Hd, a.za + row * Hd, a.ln1_g, a.m1[row], a.r1[row], a.dx1 + row * Hd, ag1, ab1, ax1);
        }
        grid_barrier(a.ws_barrier, epoch);                                      // (6) dx1_t complete

        // ---------------- P9 (latent-group owners): dz of my (rows, group) = dx1_t . W_z, kept in smem for P1 of step t-1
        if (t > 0) {
            const Job j = job_p9(t);
            float acc[1][1][4];
            // warps 0..3: class half (warp & 1) x row octet (warp >> 1)
            const bool act = warp < 4 && j.nkb > 0 && (warp & 1) * 16 < C && (warp >> 1) * 8 < b9_1 - b9_0;
            consume<1, 1>(ring, j, warp & 1, warp >> 1, 0, act, acc);
            if (act) {
                const int g = lane >> 2, tq = lane & 3;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int cls = 16 * (warp & 1) + g + 8 * (e >> 1), rb = (warp >> 1) * 8 + 2 * tq + (e & 1);
                    if (cls < C && rb < 16) dzs[rb * 32 + cls] = acc[0][0][e];
                }
            }
            cons_sync();
        }
    }

    // ---- flush the LayerNorm / bias gradient accumulators (batch-row owners)
    if (c < BI) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = tid + NCT * i;
            if (f < Hd) {
                atomicAdd(a.g_ln2_g + f, ag2[i]); atomicAdd(a.g_ln2_b + f, ab2[i]); atomicAdd(a.g_b_ph + f, ax2[i]);
                atomicAdd(a.g_ln1_g + f, ag1[i]); atomicAdd(a.g_ln1_b + f, ab1[i]); atomicAdd(a.g_b_z + f, ax1[i]);
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// row-major [rows, K] matrix as a 2-D tensor map, boxes of 128 bytes x box_rows, 128-byte swizzle, zero OOB fill
int bmap(pd_handle* h, CUtensorMap* tm, const void* base, long rows, int K, int box_rows, bool f16) {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)K * (f16 ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)(f16 ? 64 : 32), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)h->encode_tiled)(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                                   (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PD_FAIL(h, PD_ERR_ARG, "pd_rssm_unroll_bwd: cuTensorMapEncodeTiled failed (%d) for [%ld, %d]", (int)r, rows, K);
    return PD_OK;
}

}  // namespace

extern "C" int pd_rssm_unroll_bwd(pd_handle* h, const pd_rssm_bwd_args* a_in, void* stream) {
    if (!h || !a_in) return PD_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    PdDeviceGuard guard(h);
    constexpr size_t SMEM_REQ = (size_t)SMEM_BYTES + 1024;
    if (!h->k1b_configured) {
        if (cudaFuncSetAttribute(rssm_unroll_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_REQ) != cudaSuccess)
            PD_FAIL(h, PD_ERR_LAUNCH, "pd_rssm_unroll_bwd: cannot reserve %d bytes of shared memory", (int)SMEM_REQ);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rssm_unroll_bwd_kernel, NT, SMEM_REQ);
        h->k1b_ctas = per_sm > 0 ? h->num_sms : 0;
        h->k1b_configured = 1;
    }
    const int P = h->k1b_ctas;
    PD_REQUIRE(h, P > 0, "pd_rssm_unroll_bwd: kernel does not fit an SM");
    pd_rssm_bwd_args a = *a_in;
    const int Z = a.G * a.C, D3 = 3 * a.D;
    // k-split factors: slices must be whole 64-wide k-blocks
    a.ks2 = (Z % (KSPLIT * KB) == 0 && P >= KSPLIT) ? KSPLIT : 1;
    a.ks6 = (D3 % (KSPLIT * KB) == 0 && P >= KSPLIT) ? KSPLIT : 1;
    const int RG2 = P / a.ks2, RG6 = P / a.ks6;
    const int R = P / a.G < 4 ? (P / a.G < 1 ? 1 : P / a.G) : 4;
    const bool ok = a.T >= 1 && a.BI >= 1 && a.BI <= BROWS && a.BI <= P && a.Hd <= 4 * NCT && a.Hd % 4 == 0 && a.D % 4 == 0 &&
                    Z % 4 == 0 && a.C >= 1 && a.C <= 32 && a.G >= 1 && a.G <= P && (a.BI + R - 1) / R <= 16 &&
                    (a.D + P - 1) / P <= 16 && (a.Hd + RG2 - 1) / RG2 <= 32 && (a.D + RG6 - 1) / RG6 <= 64 &&
                    (a.Hd + RG6 - 1) / RG6 <= 32 && a.Hd % 8 == 0 && D3 % 8 == 0 && Z % 8 == 0;
    if (!ok)
        PD_FAIL(h, PD_ERR_UNSUPPORTED, "pd_rssm_unroll_bwd: shape T=%d BI=%d D=%d Hd=%d G=%d C=%d outside the kernel's limits",
                a.T, a.BI, a.D, a.Hd, a.G, a.C);
    BwdMaps maps;
    memset(&maps, 0, sizeof(maps));
    const long rows = (long)a.T * a.BI;
    int rc = bmap(h, &maps.wpmT, a.w_pmT16, a.Hd, Z, 16, true);
    if (!rc) rc = bmap(h, &maps.wphT, a.w_phT16, a.D, a.Hd, 16, true);
    if (!rc) rc = bmap(h, &maps.whhT, a.w_hhT16, a.D, D3, 16, true);
    if (!rc) rc = bmap(h, &maps.wihT, a.w_ihT16, a.Hd, D3, 16, true);
    if (!rc) rc = bmap(h, &maps.wzT, a.w_zT16, Z, a.Hd, 16, true);
    if (!rc) rc = bmap(h, &maps.dpost, a.dpost, rows, Z, BROWS, false);
    if (!rc) rc = bmap(h, &maps.dy2, a.dy2, rows, a.Hd, BROWS, false);
    if (!rc) rc = bmap(h, &maps.dgh, a.dgh, rows, D3, BROWS, false);
    if (!rc) rc = bmap(h, &maps.dgi, a.dgi, rows, D3, BROWS, false);
    if (!rc) rc = bmap(h, &maps.dx1, a.dx1, rows, a.Hd, BROWS, false);
    if (!rc) rc = bmap(h, &maps.dx1_16, a.dx1, rows, a.Hd, 16, false);
    if (rc) return rc;
    if (cudaMemsetAsync(a.ws_barrier, 0, 16 * sizeof(unsigned), s) != cudaSuccess)
        PD_FAIL(h, PD_ERR_LAUNCH, "pd_rssm_unroll_bwd: memset failed");
    void* kargs[] = {(void*)&a, (void*)&maps};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)rssm_unroll_bwd_kernel, dim3(P), dim3(NT), kargs, SMEM_REQ, s);
    if (e != cudaSuccess) PD_FAIL(h, PD_ERR_LAUNCH, "pd_rssm_unroll_bwd: %s", cudaGetErrorString(e));
    PD_CHECK_LAUNCH(h, "pd_rssm_unroll_bwd");
    return PD_OK;
}
