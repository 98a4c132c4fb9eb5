// pd_rssm_persistent.cu — the posterior unroll of the RSSM as ONE cooperative kernel (pd_rssm_unroll_fwd).
//
// Reference semantics: pydreamer/models/rssm.py:21-78 (RSSMCore.forward time loop) and :125-153
// (RSSMCell.forward: z_mlp + a_mlp -> in_norm -> ELU -> GRUCell -> post_mlp_h + post_mlp_e -> post_norm -> ELU ->
// post_mlp -> OneHotCategoricalStraightThrough sample).
//
// Why one kernel: with B = 50 rows a timestep is ~60 µs of nine latency-bound launches; the arithmetic is 2.8 GFLOP
// and 45 MB of fp16 weights.  Here every CTA keeps a fixed slice of the problem for all T steps
//   * hidden units  [u0,u1)  of the GRU   (3 gate rows each, W_ih and W_hh)
//   * features      [f0,f1)  of post_mlp_h
//   * one latent group g (CTAs 0..4G-1, four per group sharing the batch rows) of post_mlp + its categorical sample
//   * one batch row b   (CTAs 0..BI-1) of the two LayerNorms and of the z_mlp gather
// and the five dependent phases of a step are separated by grid barriers (one atomic + one polled word in L2).
// Contractions: out[rows, batch] = W[rows, K] · X[batch, K]^T as mma.sync.m16n8k16 (fp16 operands, fp32 accumulate);
// a K-chunk of 128 is staged with 16-byte cp.async.cg into padded smem rows (3 stages) and the 8 warps take one
// 16-wide k-step each, partial accumulators are reduced through smem.  z is one-hot, so z_mlp is a 32-row gather of
// the transposed weight instead of a 1024-deep contraction.  Activations cross CTAs through L2 (written with plain
// stores + __threadfence, read with cp.async.cg / ld.global.cg, never through L1).
#include "pd_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int NT = 256;                       // threads per CTA
constexpr int NW = 8;                         // warps = k-steps per chunk
constexpr int KC = 256;                       // halfs per staged chunk (two 16-wide k-steps per warp)
constexpr int ROWB = KC * 2 + 16;             // padded smem row pitch in bytes (conflict-free ldmatrix)
constexpr int MAXMT = 4;                      // m16 tiles of weight rows per phase
constexpr int AROWS = MAXMT * 16;
constexpr int BROWS = 64;                     // batch rows (8 n8 tiles)
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = (AROWS + BROWS) * ROWB;
constexpr int REDP = 65;                      // padded row of the reduction scratch [NW][MAXMT][16][REDP]
constexpr int RED_BYTES = NW * MAXMT * 16 * REDP * 4;
constexpr int SMEM_MAIN = (STAGES * STAGE_BYTES > RED_BYTES) ? STAGES * STAGE_BYTES : RED_BYTES;
constexpr int OFF_SH = SMEM_MAIN;              // 64 floats: block reductions
constexpr int OFF_SIDX = OFF_SH + 256;        // 64 ints: sampled classes of one row
constexpr int OFF_GH = OFF_SIDX + 256;        // [3][16][BROWS] floats: h·W_hh^T of my units (lives across phases)
constexpr int OFF_HC = OFF_GH + 3 * 16 * BROWS * 4;   // [16][BROWS] floats: masked h of my units, input of the next step
constexpr int SMEM_BYTES = OFF_HC + 16 * BROWS * 4;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
    const int n = valid ? 16 : 0;             // src-size 0 => the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_u32(dst_smem)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid-wide barrier on a monotonically increasing counter (cleared by the host before the launch).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += 1;
        const unsigned target = epoch * gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned spins = 0;
        while (ld_acquire(ctr) < target) {
            if (++spins > (1u << 24)) __trap();          // ~10 s: a lost CTA must not hang the device
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Diagnostic: CTA 0 accumulates the nanoseconds between consecutive grid barriers per phase into ws_barrier[2 + 2*slot]
// (read by tools/k1_time.py); one clock read per phase, no effect on the result.
struct PhaseClock {
    unsigned long long last;
    unsigned long long* acc;
    __device__ void start(unsigned* ws) { acc = (unsigned long long*)(ws + 2); last = gtimer(); }
    __device__ void lap(int slot) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const unsigned long long now = gtimer();
            acc[slot] += now - last;
            last = now;
        }
    }
};

struct Tile {                  // one m16 tile of weight rows: `rows` valid rows starting at `base`, row pitch K halfs
    const __half* base;
    int rows;
};

// red[w][tile][r][b] (+)= sum over this warp's k-steps of W[tile row r, k] * X[b, k];  MT tiles, batch rows < BI.
// On return red holds the NW partial sums (the caller adds them up in its epilogue).
template <int MT>
__device__ void contract(uint8_t* smem, const Tile (&tiles)[MT], const __half* X, int BI, int K, float* red) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (K + KC - 1) / KC;
    const int ntile8 = (BI + 7) >> 3;                      // n8 tiles that hold real batch rows
    float acc[MT][8][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    // Staging: a warp-wide cp.async moves one whole 512-byte row slice (lane = 16-byte piece), so the per-chunk
    // address work is one add per row: rows warp, warp+8, ... of the A tiles (tile index is a compile-time constant
    // after unrolling) and of X.
    const __half* asrc[2 * MT];
    bool aval[2 * MT];
#pragma unroll
    for (int i = 0; i < 2 * MT; ++i) {
        const int r = warp + 8 * (i & 1);
        aval[i] = r < tiles[i >> 1].rows;
        asrc[i] = tiles[i >> 1].base + (aval[i] ? (long)r * K : 0) + lane * 8;
    }
    auto issue = [&](int chunk) {
        if (chunk < nchunks) {
            uint8_t* st = smem + (chunk % STAGES) * STAGE_BYTES + lane * 16;
            const int k0 = chunk * KC;
            const bool kin = k0 + lane * 8 < K;
#pragma unroll
            for (int i = 0; i < 2 * MT; ++i)
                if (aval[i]) cp_async16(st + (warp + 8 * i) * ROWB, asrc[i] + (kin ? k0 : 0), kin);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int b = warp + 8 * j;
                if (b < BI) cp_async16(st + (AROWS + b) * ROWB, X + (long)b * K + (kin ? k0 + lane * 8 : 0), kin);
            }
        }
        cp_commit();
    };

#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i) issue(i);
    for (int c = 0; c < nchunks; ++c) {
        cp_wait<STAGES - 2>();                                  // chunk c has landed (this thread's pieces) ...
        __syncthreads();                                        // ... everyone's, and chunk c-1 is fully consumed
        issue(c + STAGES - 1);                                  // refills the stage chunk c-1 used
        const uint8_t* st = smem + (c % STAGES) * STAGE_BYTES;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int ks = warp + 8 * half;                     // this warp's two k-steps of the chunk
            if (c * KC + ks * 16 < K) {
                uint32_t a[MT][4];
#pragma unroll
                for (int i = 0; i < MT; ++i)
                    ldsm_x4(s_u32(st + (i * 16 + (lane & 15)) * ROWB + ks * 32 + (lane >> 4) * 16), a[i][0], a[i][1],
                            a[i][2], a[i][3]);
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {
                    if (jp * 2 < ntile8) {
                        uint32_t b0, b1, b2, b3;
                        ldsm_x4(s_u32(st + (AROWS + (jp * 2 + (lane >> 4)) * 8 + (lane & 7)) * ROWB + ks * 32 +
                                      ((lane >> 3) & 1) * 16), b0, b1, b2, b3);
#pragma unroll
                        for (int i = 0; i < MT; ++i) {
                            mma16816(acc[i][jp * 2], a[i], b0, b1);
                            if (jp * 2 + 1 < ntile8) mma16816(acc[i][jp * 2 + 1], a[i], b2, b3);
                        }
                    }
                }
            }
        }
    }
    cp_wait<0>();
    __syncthreads();                                            // red aliases the staging buffers
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float* r0 = red + ((warp * MAXMT + i) * 16 + gq) * REDP + j * 8 + tq * 2;
            r0[0] = acc[i][j][0]; r0[1] = acc[i][j][1];
            r0[8 * REDP] = acc[i][j][2]; r0[8 * REDP + 1] = acc[i][j][3];
        }
    __syncthreads();
}

__device__ __forceinline__ float red_sum(const float* red, int tile, int r, int b) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[((w * MAXMT + tile) * 16 + r) * REDP + b];
    return s;
}

// LayerNorm + ELU of one row held as v[4] per thread (features tid + 256 i); writes fp32 (fp16-representable)
// and fp16 copies, mean / rstd.  Same formulas as ln_elu_fwd_kernel (pd_rowwise.cu).
__device__ void ln_elu_row(float (&v)[4], int N, const float* __restrict__ gamma, const float* __restrict__ beta,
                           float eps, float* yrow, __half* y16row, float* mean_out, float* rstd_out, float* sh) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += (threadIdx.x + NT * i < N) ? v[i] : 0.f;
    const float mean = pd_block_sum(s, sh) / (float)N;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float d = (threadIdx.x + NT * i < N) ? v[i] - mean : 0.f;
        q += d * d;
    }
    const float var = pd_block_sum(q, sh) / (float)N;
    const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + NT * i;
        if (c < N) {
            const __half hv = __float2half_rn(pd_elu((v[i] - mean) * rstd * gamma[c] + beta[c]));
            yrow[c] = __half2float(hv);
            y16row[c] = hv;
        }
    }
    if (threadIdx.x == 0) { *mean_out = mean; *rstd_out = rstd; }
}

__global__ void __launch_bounds__(NT, 1) rssm_unroll_fwd_kernel(const pd_rssm_fwd_args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    float* red = (float*)smem;
    float* sh = (float*)(smem + OFF_SH);
    int* sidx = (int*)(smem + OFF_SIDX);
    float* ghs = (float*)(smem + OFF_GH);                   // ghs[(g*16 + r)*BROWS + b]
    float* hcs = (float*)(smem + OFF_HC);                   // hcs[r*BROWS + b]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P = gridDim.x, c = blockIdx.x;
    const int T = a.T, BI = a.BI, D = a.D, Hd = a.Hd, G = a.G, C = a.C, Z = G * C, F = D + Z;
    const int Bq = BI / a.I;                                // sequences (rows of aa / ea per timestep)
    const int u0 = (int)((long)c * D / P), u1 = (int)((long)(c + 1) * D / P), nu = u1 - u0;
    const int f0 = (int)((long)c * Hd / P), f1 = (int)((long)(c + 1) * Hd / P), nf = f1 - f0;
    const __half* Wz = (const __half*)a.w_z16;
    const __half* Wih = (const __half*)a.w_ih16;
    const __half* Whh = (const __half*)a.w_hh16;
    const __half* Wph = (const __half*)a.w_ph16;
    const __half* Wpm = (const __half*)a.w_pm16;
    __half* wzT = (__half*)a.ws_wzT16;
    __half* za16 = (__half*)a.ws_za16;
    __half* h16 = (__half*)a.ws_h16;
    __half* pin16 = (__half*)a.ws_pin16;
    unsigned epoch = 0;
    PhaseClock clk;
    clk.start(a.ws_barrier);

    // ---- prologue: transposed z_mlp weight (rows = latent classes, contiguous over features) and fp16 h_0
    {
        __half (*tile)[33] = (__half (*)[33])smem;
        const int tx = (Z + 31) / 32, ty = (Hd + 31) / 32;
        for (int t = c; t < tx * ty; t += P) {
            const int j0 = (t % tx) * 32, i0 = (t / tx) * 32;       // j over Z (input), i over Hd (output feature)
            for (int r = warp; r < 32; r += NW)
                tile[r][lane] = (i0 + r < Hd && j0 + lane < Z) ? Wz[(long)(i0 + r) * Z + j0 + lane] : __float2half(0.f);
            __syncthreads();
            for (int r = warp; r < 32; r += NW)
                if (j0 + r < Z && i0 + lane < Hd) wzT[(long)(j0 + r) * Hd + i0 + lane] = tile[lane][r];
            __syncthreads();
        }
        for (long i = (long)c * NT + tid; i < (long)BI * D; i += (long)P * NT) h16[i] = __float2half_rn(__ldcg(a.hin + i));
        for (int o = tid; o < nu * BI; o += NT) hcs[(o % nu) * BROWS + o / nu] = __ldcg(a.hin + (long)(o / nu) * D + u0 + o % nu);
    }
    grid_barrier(a.ws_barrier, epoch);
    clk.lap(0);

    // gh_0 = h_0 · W_hh^T (raw product; bias and the step mask are applied where it is consumed)
    auto phase_hidden = [&](int t, bool want_gh, bool want_y2) {
        Tile tiles[4];
#pragma unroll
        for (int g = 0; g < 3; ++g) { tiles[g].base = Whh + ((long)g * D + u0) * D; tiles[g].rows = want_gh ? nu : 0; }
        tiles[3].base = Wph + (long)f0 * D;
        tiles[3].rows = want_y2 ? nf : 0;
        float pea[4];                                     // epilogue operands are fetched before the contraction
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int o = tid + NT * i;
            pea[i] = 0.f;
            if (want_y2 && o < nf * BI) {
                const int f = f0 + o % nf, b = o / nf;
                pea[i] = a.b_ph[f] + (a.ea ? a.ea[((long)t * Bq + b / a.I) * Hd + f] : 0.f);
            }
        }
        contract<4>(smem, tiles, h16, BI, D, red);
        if (want_gh)
            for (int o = tid; o < 3 * nu * BI; o += NT) {
                const int r = o % nu, b = (o / nu) % BI, g = o / (nu * BI);
                ghs[(g * 16 + r) * BROWS + b] = red_sum(red, g, r, b);
            }
        if (want_y2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = tid + NT * i;
                if (o < nf * BI) {
                    const int r = o % nf, b = o / nf;
                    a.y2[((long)t * BI + b) * Hd + f0 + r] = red_sum(red, 3, r, b) + pea[i];
                }
            }
        }
    };
    phase_hidden(0, true, false);
    grid_barrier(a.ws_barrier, epoch);
    clk.lap(1);

    for (int t = 0; t < T; ++t) {
        // ---- phase A (CTA b < BI): x1 = mask * gather(WzT, idx_{t-1}) + b_z + aa_t ; LayerNorm + ELU -> za
        if (c < BI) {
            const int b = c;
            const long row = (long)t * BI + b;
            float v[4];
            if (t > 0) {
                const float m = a.mask[row];
                for (int g = tid; g < G; g += NT) sidx[g] = __ldcg(a.idx + ((long)(t - 1) * BI + b) * G + g);
                float pa[4];                                              // bias + action term, in flight during the gather
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int f = tid + NT * i;
                    pa[i] = f < Hd ? a.b_z[f] + a.aa[((long)t * Bq + b / a.I) * Hd + f] : 0.f;
                }
                __syncthreads();
                // gather-sum of G rows of WzT: thread = (8 features, half of the groups), 16-byte loads, all independent
                float* part = (float*)smem;                                // [2][Hd]
                {
                    const int fg = tid & 127, gh = tid >> 7;
                    if (fg * 8 < Hd) {
                        float s8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) s8[e] = 0.f;
                        const int g0 = gh * ((G + 1) / 2), g1 = min(G, g0 + (G + 1) / 2);
#pragma unroll 4
                        for (int g = g0; g < g1; ++g) {
                            const uint4 w = *(const uint4*)(wzT + (long)(g * C + sidx[g]) * Hd + fg * 8);
                            const __half2* h2 = (const __half2*)&w;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f2 = __half22float2(h2[e]);
                                s8[2 * e] += f2.x; s8[2 * e + 1] += f2.y;
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) part[gh * Hd + fg * 8 + e] = s8[e];
                    }
                }
                __syncthreads();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int f = tid + NT * i;
                    v[i] = 0.f;
                    if (f < Hd) {
                        v[i] = m * (part[f] + part[Hd + f]) + pa[i];
                        a.x1[row * Hd + f] = v[i];
                    }
                }
                for (int j = tid; j < Z; j += NT) a.zin[row * Z + j] = (sidx[j / C] == j % C) ? m : 0.f;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int f = tid + NT * i;
                    v[i] = f < Hd ? __ldcg(a.x1 + row * Hd + f) : 0.f;
                }
            }
            ln_elu_row(v, Hd, a.ln1_g, a.ln1_b, a.eps, a.za + row * Hd, za16 + (long)b * Hd, a.m1 + row, a.r1 + row, sh);
        }
        grid_barrier(a.ws_barrier, epoch);
        clk.lap(2);

        // ---- phase B (all CTAs): gi = za · W_ih^T for my units, GRU gate math, h' -> feat / hin[t+1] / h16
        {
            Tile tiles[3];
#pragma unroll
            for (int g = 0; g < 3; ++g) { tiles[g].base = Wih + ((long)g * D + u0) * Hd; tiles[g].rows = nu; }
            float pm[4], pmn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = tid + NT * i;
                pm[i] = 1.f; pmn[i] = 0.f;
                if (o < nu * BI) {
                    const long row = (long)t * BI + o / nu;
                    if (t > 0) pm[i] = a.mask[row];                  // h_0 arrives already masked
                    if (t + 1 < T) pmn[i] = a.mask[row + BI];
                }
            }
            contract<3>(smem, tiles, za16, BI, Hd, red);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = tid + NT * i;
                if (o < nu * BI) {
                    const int r = o % nu, b = o / nu, u = u0 + r;
                    const long row = (long)t * BI + b;
                    const float m = pm[i];
                    const float ghr = m * ghs[(0 * 16 + r) * BROWS + b] + a.b_hh[u];
                    const float ghu = m * ghs[(1 * 16 + r) * BROWS + b] + a.b_hh[D + u];
                    const float ghn = m * ghs[(2 * 16 + r) * BROWS + b] + a.b_hh[2 * D + u];
                    const float rg = pd_sigmoid(red_sum(red, 0, r, b) + a.b_ih[u] + ghr);
                    const float ug = pd_sigmoid(red_sum(red, 1, r, b) + a.b_ih[D + u] + ghu);
                    const float ng = tanhf(red_sum(red, 2, r, b) + a.b_ih[2 * D + u] + rg * ghn);
                    const float hp = hcs[r * BROWS + b];
                    const __half hh = __float2half_rn((1.f - ug) * ng + ug * hp);
                    const float hn = __half2float(hh);
                    a.feat[row * F + u] = hn;
                    h16[(long)b * D + u] = hh;
                    hcs[r * BROWS + b] = hn * pmn[i];
                    if (t + 1 < T) a.hin[(row + BI) * D + u] = hn * pmn[i];
                    float* gt = a.gates + row * 4 * D;
                    gt[u] = rg; gt[D + u] = ug; gt[2 * D + u] = ng; gt[3 * D + u] = ghn;
                }
            }
        }
        grid_barrier(a.ws_barrier, epoch);
        clk.lap(3);

        // ---- phase C (all CTAs): y2 = h' · W_ph^T + b + ea_t for my features; gh_{t+1} = h' · W_hh^T for my units
        phase_hidden(t, t + 1 < T, true);
        grid_barrier(a.ws_barrier, epoch);
        clk.lap(4);

        // ---- phase C' (CTA b < BI): LayerNorm + ELU of y2 -> pin
        if (c < BI) {
            const int b = c;
            const long row = (long)t * BI + b;
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int f = tid + NT * i;
                v[i] = f < Hd ? __ldcg(a.y2 + row * Hd + f) : 0.f;
            }
            ln_elu_row(v, Hd, a.ln2_g, a.ln2_b, a.eps, a.pin + row * Hd, pin16 + (long)b * Hd, a.m2 + row, a.r2 + row, sh);
        }
        grid_barrier(a.ws_barrier, epoch);
        clk.lap(5);

        // ---- phase D (CTA g < G): logits of latent group g, softmax, argmax(p / q) -> post, idx, z
        // (R CTAs repeat the small contraction of a group and share its batch rows: the softmax / argmax chain is the
        //  serial part of this phase)
        const int R = max(1, min(4, P / G));
        if (c < G * R) {
            const int g = c / R, sub = c % R;
            Tile tiles[2];
            tiles[0].base = Wpm + (long)g * C * Hd;
            tiles[0].rows = C < 16 ? C : 16;
            tiles[1].base = Wpm + ((long)g * C + 16) * Hd;
            tiles[1].rows = C > 16 ? C - 16 : 0;
            float pq[8], pbias = lane < C ? a.b_pm[g * C + lane] : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int b = warp + NW * i;
                pq[i] = (lane < C && b < BI && i % R == sub) ? a.noise[((long)t * BI + b) * Z + g * C + lane] : 1.f;
            }
            contract<2>(smem, tiles, pin16, BI, Hd, red);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int b = warp + NW * i;
                if (b >= BI) break;
                if (i % R != sub) continue;
                const long row = (long)t * BI + b;
                const bool valid = lane < C;
                float l = 0.f;
                if (valid) {
                    l = red_sum(red, lane >> 4, lane & 15, b) + pbias;
                    a.post[row * Z + g * C + lane] = l;
                }
                // same arithmetic as cat_sample_kernel (pd_rowwise.cu): logits - logsumexp, softmax, argmax(p / q)
                const float mx = pd_warp_max(valid ? l : -INFINITY);
                const float e = valid ? expf(l - mx) : 0.f;
                const float lse = mx + logf(pd_warp_sum(e));
                const float ln = valid ? l - lse : -INFINITY;
                const float mx2 = pd_warp_max(ln);
                const float e2 = valid ? expf(ln - mx2) : 0.f;
                const float p = e2 / pd_warp_sum(e2);
                const float q = pq[i];
                float val = valid ? p / q : -INFINITY;
                int k = lane;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, val, o);
                    const int ok = __shfl_xor_sync(0xffffffffu, k, o);
                    if (ov > val || (ov == val && ok < k)) { val = ov; k = ok; }
                }
                if (valid) a.feat[row * F + D + g * C + lane] = (lane == k) ? 1.f : 0.f;
                if (lane == 0) a.idx[row * G + g] = k;
            }
        }
        if (t + 1 < T) grid_barrier(a.ws_barrier, epoch);
        clk.lap(6);
    }
}

}  // namespace

extern "C" int pd_rssm_unroll_fwd(pd_handle* h, const pd_rssm_fwd_args* a, void* stream) {
    if (!h || !a) return PD_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    PdDeviceGuard guard(h);
    if (!h->k1_configured) {                                   // per handle = per device (the attribute is per device)
        if (cudaFuncSetAttribute(rssm_unroll_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 128) !=
            cudaSuccess)
            PD_FAIL(h, PD_ERR_LAUNCH, "pd_rssm_unroll_fwd: cannot reserve %d bytes of shared memory", SMEM_BYTES);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rssm_unroll_fwd_kernel, NT, SMEM_BYTES + 128);
        h->k1_ctas = per_sm > 0 ? h->num_sms : 0;             // one CTA per SM
        h->k1_configured = 1;
    }
    const int P = h->k1_ctas;
    PD_REQUIRE(h, P > 0, "pd_rssm_unroll_fwd: kernel does not fit an SM");
    const int Z = a->G * a->C;
    const bool ok = a->T >= 1 && a->BI >= 1 && a->BI <= BROWS && a->BI <= P && a->I >= 1 && a->BI % a->I == 0 &&
                    a->Hd <= 4 * NT && a->Hd % 8 == 0 && a->D % 8 == 0 && a->C >= 1 && a->C <= 32 && a->G >= 1 &&
                    a->G <= P && a->G <= 64 && (a->D + P - 1) / P <= 16 && (a->Hd + P - 1) / P <= 16 && Z >= 1;
    if (!ok)
        PD_FAIL(h, PD_ERR_UNSUPPORTED, "pd_rssm_unroll_fwd: shape T=%d BI=%d D=%d Hd=%d G=%d C=%d outside the kernel's limits",
                a->T, a->BI, a->D, a->Hd, a->G, a->C);
    if (cudaMemsetAsync(a->ws_barrier, 0, 16 * sizeof(unsigned), s) != cudaSuccess)
        PD_FAIL(h, PD_ERR_LAUNCH, "pd_rssm_unroll_fwd: memset failed");
    pd_rssm_fwd_args args = *a;
    void* kargs[] = {(void*)&args};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)rssm_unroll_fwd_kernel, dim3(P), dim3(NT), kargs,
                                                (size_t)SMEM_BYTES + 128, s);
    if (e != cudaSuccess) PD_FAIL(h, PD_ERR_LAUNCH, "pd_rssm_unroll_fwd: %s", cudaGetErrorString(e));
    PD_CHECK_LAUNCH(h, "pd_rssm_unroll_fwd");
    return PD_OK;
}
