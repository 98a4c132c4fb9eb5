/*
 * pd_b200.h — C ABI of libpd_b200.so: hand-written sm_100a kernels behind the PyDreamer
 * world-model training step + imagination rollout (BASELINE.json north_star).
 *
 * The reference (jurgisp/pydreamer) has no FFI: its boundary is the Python class
 * pydreamer.models.Dreamer (pydreamer/models/dreamer.py:19).  These entry points are the
 * arithmetic that class performs, one per fused device op; pydreamer_b200/dreamer.py composes
 * them behind the reference's Dreamer API.  Each declaration cites the reference lines whose
 * math it replaces.
 *
 * Conventions:
 *   - plain pointers/sizes only; every pointer is a DEVICE pointer to fp32 unless noted;
 *   - `ld*` are row strides in ELEMENTS, so callers can pass views into concatenated buffers;
 *   - never allocates device memory, never synchronises, enqueues on `stream` (a cudaStream_t);
 *   - returns 0 on success, a negative PD_ERR_* otherwise; pd_last_error() gives the message;
 *   - one handle per (process, device); a handle is not re-entrant.
 */
#ifndef PD_B200_H
#define PD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pd_handle pd_handle;

#define PD_OK 0
#define PD_ERR_ARG (-1)
#define PD_ERR_LAUNCH (-2)
#define PD_ERR_DEVICE (-3)
#define PD_ERR_UNSUPPORTED (-4)

#define PD_ACT_NONE 0
#define PD_ACT_ELU 1

#define PD_GEMM_TCGEN05 0 /* tcgen05.mma kind::tf32 + TMA + TMEM (default, the product path) */
#define PD_GEMM_SIMT 1    /* plain fp32 CUDA-core tile kernel: validation arm for the tests  */
#define PD_GEMM_C_ZEROED 1 /* pd_gemm flags bit */
#define PD_GEMM_C_F16 2    /* pd_gemm flags bit: C is an fp16 matrix (ldc in halfs); not with accumulate */

/* ---- lifetime ---------------------------------------------------------------------------- */
int pd_create(int device_ordinal, pd_handle** out);
void pd_destroy(pd_handle* h);
const char* pd_last_error(const pd_handle* h);
const char* pd_version(void);
int pd_set_gemm_impl(pd_handle* h, int impl);
long pd_launch_count(const pd_handle* h); /* kernels enqueued through this handle so far */
/* 1 (default): kernels that produce tensor-core operands round them to TF32 (rna) so the MMA's
 * operand truncation is exact; 0: keep full fp32 outputs (used by exactness tests). */
int pd_set_round_operands(pd_handle* h, int on);

/* ---- dense contraction ------------------------------------------------------------------- */
/* C[M,N] (=|+=) sum_k A(m,k) * B(n,k)  [+ bias[n]] [+ R[m / r_div, n]] -> act -> (tf32 round)
 *   a_mn = 0: A stored [M][K] (K contiguous, row stride lda); a_mn = 1: A stored [K][M] (M contiguous).
 *   b_mn = 0: B stored [N][K] (nn.Linear weight layout);      b_mn = 1: B stored [K][N].
 *   accumulate = 1: atomically adds into C (split-K over all SMs; bias/R/act must be off).
 * Replaces every nn.Linear / nn.GRUCell matmul and the conv/deconv contractions:
 * common.py:47-55, rssm.py:103-116,138-146, rnn.py:60-67, encoders.py:80-90, decoders.py:128-155.
 * flags: PD_GEMM_C_ZEROED = the caller has already cleared C (skinny-M launches that split K skip their own clear);
 *        PD_GEMM_C_F16 = C points to an fp16 matrix (the deconvolution column matrices of the decoder forward: they are
 *        written once and read once, in fp16 they cost half the HBM traffic; decoders.py:149-155).
 * TMA constraints (tcgen05 impl): lda/ldb multiples of 4 elements, base pointers 16-byte aligned. */
int pd_gemm(pd_handle* h, int M, int N, int K,
            const float* A, long lda, int a_mn,
            const float* B, long ldb, int b_mn,
            float* C, long ldc,
            const float* bias, const float* R, long ldr, int r_div,
            int act, int round_out, int accumulate, int flags, void* stream);

/* Forward-only contraction with fp16 operands (tcgen05 kind::f16, fp32 accumulate / output): same 10-bit mantissa as
 * TF32 at twice the tensor rate and half the operand bytes.  Used where no gradient flows through the GEMM (the
 * imagination rollout and the heads evaluated on dreamed features, dreamer.py:188-216, a2c.py:88,112).
 * A: [M][K] fp16, B: [N][K] fp16 (both K-major, ld % 8 == 0), C fp32. */
int pd_gemm_f16(pd_handle* h, int M, int N, int K, const void* A, long lda, const void* B, long ldb,
                float* C, long ldc, const float* bias, const float* R, long ldr, int r_div,
                int act, int round_out, void* stream);
/* Implicit-GEMM convolution contractions: one operand is gathered on the fly from an NHWC fp32 tensor X[NB,H,W,C] by TMA
 * im2col-mode loads (k x k taps, stride 2, no padding; P,Q = (H-k)/2+1) instead of a materialised im2col matrix.
 *   mode 1: Cmat[NB*P*Q, odim] = im2col(X) * O        O: [odim][k*k*C] (o_mn=0) or [k*k*C][odim] (o_mn=1)
 *           -> Conv2d forward (encoders.py:80-88) and ConvTranspose2d input gradient
 *   mode 2: Cmat[k*k*cpad, odim] += im2col(X)^T * O   O: [NB*P*Q][odim]; rows (tap, channel) with channels padded to 32
 *           -> ConvTranspose2d weight gradient
 *   mode 3: Cmat[odim, k*k*cpad] += O^T * im2col(X)   O: [NB*P*Q][odim]
 *           -> Conv2d weight gradient */
int pd_conv_gemm(pd_handle* h, int mode, int NB, int H, int W, int C, int k, const float* X, const float* O, long ldo, int o_mn,
                 int odim, float* Cmat, long ldc, const float* bias, int act, int round_out, int accumulate, void* stream);
/* dst(fp16)[m, n] = src(fp32)[m, n] */
int pd_to_half(pd_handle* h, long M, long N, const float* src, long lds, void* dst, long ldd, void* stream);

/* ---- LayerNorm(eps, biased var, affine) + ELU -------------------------------------------- */
/* y = ELU(LN(x)); saves per-row mean / rstd.  common.py:45-51, rssm.py:105,110,115,139-140,144-145. */
int pd_ln_elu_fwd(pd_handle* h, int M, int N, const float* x, long ldx,
                  const float* gamma, const float* beta, float eps,
                  float* y, long ldy, float* mean, float* rstd,
                  void* y16 /* optional fp16 copy of y (operand of a forward-only pd_gemm_f16) */, long ldy16, void* stream);
/* dx from dy; ACCUMULATES dgamma, dbeta and (optional) dbias (= column sums of dx, the grad of the
 * bias of the Linear that produced x) with atomics. */
int pd_ln_elu_bwd(pd_handle* h, int M, int N, const float* dy, long lddy,
                  const float* x, long ldx, const float* y, long ldy,
                  const float* gamma, const float* mean, const float* rstd,
                  float* dx, long lddx, float* dgamma, float* dbeta, float* dbias, void* stream);

/* ---- GRU cell pointwise (torch.nn.GRUCell gate order r|u|n) -------------------------------- */
/* rnn.py:48-49,60-67 -> nn.GRUCell: r=s(gi_r+gh_r) u=s(gi_u+gh_u) n=tanh(gi_n+r*gh_n) h'=(1-u)n+u*h.
 * gates[M,4,D] (optional) saves r,u,n,gh_n.  hmask (optional): also writes h' * mask_next[m] (the next step's
 * reset-masked input, rssm.py:134). */
int pd_gru_fwd(pd_handle* h, int M, int D, const float* gi, long ldgi, const float* gh, long ldgh,
               const float* hprev, long ldh, float* hout, long ldho,
               float* hmask, long ldhm, const float* mask_next,
               float* gates, void* h16 /* optional fp16 copy of h' */, long ldh16, void* stream);
/* dh_out = dh_a + dh_b * mask_b (either optional);  outputs dgi[M,3D], dgh[M,3D] and dh_carry = dh_out*u. */
int pd_gru_bwd(pd_handle* h, int M, int D, const float* dh_a, long ldda, const float* dh_b, long lddb,
               const float* mask_b, const float* gates, const float* hprev, long ldh,
               float* dgi, long lddgi, float* dgh, long lddgh, float* dh_carry, long lddc,
               void* stream);

/* ---- categorical straight-through latent -------------------------------------------------- */
/* logits[M, G*C] -> per (m,g): l = logits - logsumexp; p = softmax(l); k = argmax_c p_c / q_c
 * (== torch.multinomial's sampling); z = onehot(k).  C <= 32.
 * rssm.py:147-148,178-179,195-201; a2c.py:47-48 with G = 1 for the one-hot actor.
 * zmask (optional) = z * mask_next[m]; idx (optional) int32 [M,G]. */
int pd_cat_sample(pd_handle* h, int M, int G, int C, const float* logits, long ldl,
                  const float* noise, long ldn, float* z, long ldz,
                  float* zmask, long ldzm, const float* mask_next, int32_t* idx,
                  void* z16 /* optional fp16 copy of z */, long ldz16, void* stream);
/* straight-through backward: dlogits = p * (dz - sum_c p dz) + alpha * rowscale[m] * extra;
 * dz = dz_a + dz_b * mask_b (each optional). */
int pd_cat_st_bwd(pd_handle* h, int M, int G, int C, const float* logits, long ldl,
                  const float* dz_a, long ldda, const float* dz_b, long lddb, const float* mask_b,
                  const float* extra, long ldex, const float* rowscale, float alpha,
                  float* dlogits, long lddl, void* stream);

/* ---- persistent RSSM posterior unroll (rssm.py:21-78 RSSMCore.forward, 125-153 RSSMCell.forward) -------------
 * ONE cooperative kernel walks all T timesteps (csrc/pd_rssm_fwd3.cu): per step { z_mlp as a gather over the sampled
 * one-hot + a_mlp term -> LayerNorm+ELU -> GRU gates -> post_mlp_h + embed term -> LayerNorm+ELU -> post_mlp ->
 * categorical sample }, with grid-wide barriers between the dependent phases instead of kernel boundaries; every CTA owns
 * a fixed slice of hidden units / features / latent groups for the whole sequence.  Operands are staged by a producer
 * warp with TMA into an mbarrier ring (weights of the next phase are prefetched across the grid barrier), contractions are
 * fp16 mma.sync with fp32 accumulation, the recurrent products over K = D are split in four k-slices.  Writes exactly the
 * buffers the chain of per-step kernels writes (the backward reads them).
 * Caller prepares: hin[0] / zin[0] (masked in_state), x1[0] (pre-norm input of step 0 incl. bias and action term),
 * aa / ea (action and embed projections hoisted over T), fp16 weight copies incl. the TRANSPOSED z_mlp weight ws_wzT16
 * (pd_transpose_to_half).  Limits: BI <= 64, Hd <= 1024, C <= 32, G <= #CTAs, D <= 16 * #CTAs, D and Hd multiples of 8;
 * otherwise PD_ERR_UNSUPPORTED (use the chain). */
typedef struct pd_rssm_fwd_args {
    int T, BI, I, D, Hd, G, C;                 /* rows BI = B*I; Z = G*C; feat row pitch F = D + Z */
    const void *w_z16, *w_ih16, *w_hh16, *w_ph16, *w_pm16;   /* fp16 [Hd,Z] [3D,Hd] [3D,D] [Hd,D] [Z,Hd] */
    const float *b_z, *ln1_g, *ln1_b, *b_ih, *b_hh, *b_ph, *ln2_g, *ln2_b, *b_pm;
    float eps;
    const float *aa;                           /* [T*B, Hd] */
    const float *ea;                           /* [T*B, Hd] or NULL (open loop: no embed term) */
    const float *mask;                         /* [T, BI]  1 - reset */
    const float *noise;                        /* [T, BI, Z] Exp(1) */
    float *x1, *za, *m1, *r1;                  /* [T,BI,Hd] x2, [T,BI] x2 */
    float *gates;                              /* [T,BI,4D]  r,u,n,gh_n */
    float *feat;                               /* [T,BI,D+Z] h' | z */
    float *hin, *zin;                          /* [T,BI,D], [T,BI,Z] masked step inputs */
    float *y2, *pin, *m2, *r2;                 /* [T,BI,Hd] x2, [T,BI] x2 */
    float *post;                               /* [T,BI,Z] posterior logits */
    int32_t *idx;                              /* [T,BI,G] sampled classes */
    void *ws_wzT16;                            /* in: fp16 [Z,Hd] = z_mlp.weight^T (pd_transpose_to_half) */
    void *ws_za16, *ws_h16, *ws_pin16;         /* workspace fp16 [BI,Hd] [BI,D] [BI,Hd] */
    unsigned int *ws_barrier;                  /* workspace, 16 words, 8-byte aligned, cleared by the call: [0] barrier counter */
    float *ws_ghpart, *ws_y2part;              /* workspace 4*BI*3D and 4*BI*Hd floats: k-slice partial sums of the recurrent products */
} pd_rssm_fwd_args;
int pd_rssm_unroll_fwd(pd_handle* h, const pd_rssm_fwd_args* a, void* stream);

/* ---- Back-propagation through time of the posterior unroll, one persistent cooperative kernel ------------------------ */
/* Autograd of rssm.py:21-78,125-153 / rnn.py:60-67 for all T timesteps (csrc/pd_rssm_bptt.cu): per step, descending t,
 *   dz     = dfeat[t][:, D:] + mask[t+1] * (dx1[t+1] . W_z)                       straight-through sample, rssm.py:147-148
 *   dpost  = softmax'(post[t]) dz + kl_weight * w[t] * dpost_u[t]                 (+ the KL gradient of dreamer.py:328-343)
 *   dy2    = LN+ELU backward(post_norm)(dpost . W_pm)                             rssm.py:115-116
 *   dh     = dy2 . W_ph + dfeat[t][:, :D] + mask[t+1] * (dgh[t+1] . W_hh + dh[t+1] * u[t+1])
 *   dgi, dgh = GRU gate backward(dh; gates[t], hin[t])                            nn.GRUCell, gate order r|u|n
 *   dx1    = LN+ELU backward(in_norm)(dgi . W_ih)                                 rssm.py:138-140
 * Inputs are what pd_rssm_unroll_fwd (or the per-timestep chain) saved; outputs dpost, dy2, dgi, dgh, dx1 [T, BI, .] are the
 * operands of the batched weight-gradient GEMMs, tf32-rounded when round_out != 0.  LayerNorm affine and the two bias
 * gradients fed by LayerNorm inputs are ACCUMULATED (+=) into g_*.  Weights come as TRANSPOSED fp16 copies (pd_transpose_to_half):
 * contraction operands carry 10 mantissa bits on both sides (fp16 weights are exact in tf32; gradients are tf32-rounded fp32),
 * like the TF32 GEMMs of the launch chain this replaces.  Limits: BI <= 64, Hd <= 1024, C <= 32, D/#SMs <= 16. */
typedef struct pd_rssm_bwd_args {
    int T, BI, D, Hd, G, C;
    int round_out;
    int ks2, ks6;                              /* set by the library (k-split factors) */
    float kl_weight;
    const void *w_pmT16, *w_phT16, *w_hhT16, *w_ihT16, *w_zT16;   /* fp16: [Hd,Z] [D,Hd] [D,3D] [Hd,3D] [Z,Hd] */
    const float *ln2_g, *ln1_g;                /* post_norm / in_norm weight [Hd] */
    const float *post, *pin, *y2, *m2, *r2;    /* [T,BI,Z] [T,BI,Hd] [T,BI,Hd] [T,BI] [T,BI] */
    const float *x1, *za, *m1, *r1;            /* [T,BI,Hd] [T,BI,Hd] [T,BI] [T,BI] */
    const float *gates, *hin, *mask;           /* [T,BI,4D] (r,u,n,gh_n) [T,BI,D] [T,BI] */
    const float *dfeat, *dpost_u, *w;          /* [T,BI,D+Z] seeds, [T,BI,Z] unweighted KL gradient, [T,BI] row weights */
    float *dpost, *dy2, *dgi, *dgh, *dx1;      /* out [T,BI,Z] [T,BI,Hd] [T,BI,3D] [T,BI,3D] [T,BI,Hd] */
    float *g_ln2_g, *g_ln2_b, *g_b_ph, *g_ln1_g, *g_ln1_b, *g_b_z;   /* += [Hd] each */
    float *ws_part2, *ws_part6, *ws_part7;     /* workspace [4,BI,Hd] [4,BI,D] [4,BI,Hd] */
    unsigned int *ws_barrier;                  /* workspace, 16 words, cleared by the call */
} pd_rssm_bwd_args;
int pd_rssm_unroll_bwd(pd_handle* h, const pd_rssm_bwd_args* a, void* stream);
/* dst[n, m] (fp16) = src[m, n] (fp32): the transposed fp16 weight copies pd_rssm_unroll_bwd contracts with. */
int pd_transpose_to_half(pd_handle* h, int M, int N, const float* src, long lds, void* dst, long ldd, void* stream);

/* ---- KL(post || prior) with balancing, entropies, unweighted grads ------------------------ */
/* dreamer.py:328-343,369-379.  mode 0 (I == 1): value KL, grads (1-bal)*dKL/dpost and bal*dKL/dprior
 * (bal < 0 => plain KL, kl_balance == 0.5 case dreamer.py:241).  mode 1 (I > 1): sampled
 * log q(z) - log p(z) with idx from pd_cat_sample.  kl_exact / entropies are for metrics. */
int pd_kl(pd_handle* h, int M, int G, int C, const float* post, long ldpo, const float* prior, long ldpr,
          const int32_t* idx, int mode, float balance,
          float* loss_kl, float* kl_exact, float* ent_post, float* ent_prior,
          float* dpost, long lddpo, float* dprior, long lddpr, void* stream);

/* ---- convolution data movement (k x k, stride 2, no padding) ------------------------------ */
/* col[(n,oy,ox), kidx] = in[n, 2oy+kh, 2ox+kw, c]; korder 0: kidx=(kh,kw,c); 1: kidx=(c,kh,kw).
 * Input addressed by element strides (sN,sY,sX,sC) so NCHW and NHWC both work.
 * Forward operand of Conv2d (encoders.py:80-88) and backward operand of ConvTranspose2d. */
int pd_im2col(pd_handle* h, int NB, int Hin, int Win, int Cc, int k, int korder,
              const float* in, long sN, long sY, long sX, long sC,
              float* col, long ldcol, int round_out, void* stream);
/* out[n,y,x,c] = act(bias[c] + sum_{kh,kw: (y-kh)%2==0...} col[(n,(y-kh)/2,(x-kw)/2), (kh,kw,c)]).
 * Forward of ConvTranspose2d (decoders.py:149-155) and input-gradient of Conv2d. */
int pd_col2im(pd_handle* h, int NB, int Hin, int Win, int Hout, int Wout, int Cc, int k,
              const float* col, long ldcol, const float* bias, int act, int round_out,
              float* out, long sN, long sY, long sX, long sC, void* stream);
/* Last decoder layer fused with the image loss (decoders.py:163-167): dec NCHW, target NCHW of
 * image row n / tgt_div; loss[n] = 0.5*sum diff^2; diff stored NHWC for the backward gather. */
int pd_col2im_imgloss(pd_handle* h, int NB, int Hin, int Win, int Cc, int k,
                      const float* col, long ldcol, const float* bias,
                      const float* target, int tgt_div,
                      float* dec, float* diff, float* loss, float* csum /* [NB,Cc] per-image channel sums of diff */,
                      void* stream);
/* The same two folds over an fp16 (col_f16 = 1) or fp32 (0) column matrix; ldcol in elements. */
int pd_col2im_t(pd_handle* h, int NB, int Hin, int Win, int Hout, int Wout, int Cc, int k, const void* col, long ldcol,
                int col_f16, const float* bias, int act, int round_out, float* out, long sN, long sY, long sX, long sC,
                void* stream);
int pd_col2im_imgloss_t(pd_handle* h, int NB, int Hin, int Win, int Cc, int k, const void* col, long ldcol, int col_f16,
                        const float* bias, const float* target, int tgt_div, float* dec, float* diff, float* loss,
                        float* csum, void* stream);
/* dy <- dy * act'(y) in place (act from output y), db[c] += column sums of the result. */
int pd_bias_act_bwd(pd_handle* h, long M, int N, float* dy, long lddy, const float* y, long ldy,
                    int act, float* db, void* stream);
/* The ELU backward + bias gradient of the layer BELOW fused into the kernel that produces that layer's output gradient
 * (instead of a separate pd_bias_act_bwd pass over the gradient image; PD_B200_FUSE_ACTBWD=0 composes the two launches):
 *   pd_gemm_actbwd:      C = (A B^T) .* elu'(dact), dbias[n] += sum_m C[m, n]      (Linear / explicit-column deconv dX;
 *                        decoders.py:128-155 backward, autograd of nn.ELU + bias)
 *   pd_conv_gemm_actbwd: the same for pd_conv_gemm mode 1 (ConvTranspose2d input gradient gathered by TMA im2col)
 *   pd_col2im_actbwd:    out = fold(col) .* elu'(dact), dbias[c] += sum over pixels  (Conv2d input gradient; encoders.py:80-90
 *                        backward); out / dact contiguous NHWC [NB, Hout, Wout, Cc], Hout >= 2(Hin-1)+k (rows a stride-2 conv never read get 0)
 * dact is the saved forward output of the layer below (ELU derivative from the output: y > 0 ? 1 : y + 1). */
int pd_gemm_actbwd(pd_handle* h, int M, int N, int K, const float* A, long lda, int a_mn, const float* B, long ldb, int b_mn,
                   float* C, long ldc, const float* dact, long lddact, float* dbias, void* stream);
int pd_conv_gemm_actbwd(pd_handle* h, int NB, int H, int W, int C, int k, const float* X, const float* O, long ldo, int o_mn,
                        int odim, float* Cmat, long ldc, const float* dact, long lddact, float* dbias, void* stream);
int pd_col2im_actbwd(pd_handle* h, int NB, int Hin, int Win, int Hout, int Wout, int Cc, int k, const float* col, long ldcol,
                     const float* dact, float* dbias, float* out, void* stream);
/* generic 4-D permutation copy out[perm(i)] (+)= in[i];  dims (HOST int[4]) of `in`, perm[j] (HOST) = source axis of out axis j. */
int pd_permute4(pd_handle* h, const float* in, float* out, const int* dims, const int* perm,
                const long* in_strides /* HOST long[4] element strides of `in`, or NULL = contiguous */,
                int accumulate, int round_out, void* stream);

/* ---- small pointwise / reductions ---------------------------------------------------------- */
int pd_round_copy(pd_handle* h, const float* src, float* dst, long n, int round_out, void* stream);
int pd_pad_cols(pd_handle* h, long M, int C, int Cp, const float* src, long lds, float* dst, long ldd, void* stream);
int pd_mask_rows(pd_handle* h, int M, int N, const float* x, long ldx, const float* mask,
                 float* out, long ldo, void* stream);
/* x[m,:] *= alpha * scale[m / scale_div] */
int pd_rowscale(pd_handle* h, long M, long N, float* x, long ldx, const float* scale, int scale_div, float alpha, void* stream);
/* out[m, :] = W[idx[m], :]: `a_mlp` (rssm.py:104, Linear without bias) applied to a one-hot action is a row gather of the
 * transposed weight W = a_mlp.weight^T [A, Hd] (dreamer.py:197-205: the imagination rollout samples one-hot actions). */
int pd_gather_rows(pd_handle* h, long M, int N, const int32_t* idx, const float* W, long ldw, float* out, long ldo,
                   void* stream);
/* out[r, :] = sum_{i<I} x[r*I + i, :]  (undo the IWAE row expansion, rssm.py:35-41) */
int pd_group_sum(pd_handle* h, long R, int I, int W, const float* x, long ldx, float* out, long ldo, void* stream);
int pd_colsum(pd_handle* h, long M, int N, const float* x, long ldx, float* out, void* stream); /* out += */
int pd_fill(pd_handle* h, float* x, long n, float v, void* stream);
/* reset (uint8/bool [T,B]) -> mask f32 [T, B*I] = !reset   (rssm.py:41) */
int pd_reset_mask(pd_handle* h, int T, int B, int I, const uint8_t* reset, float* mask, void* stream);

/* ---- scalar-head losses (decoders.py:257-319) ---------------------------------------------- */
/* kind 0: DenseNormalDecoder  loss = 0.5 (t-y)^2, dy = (y-t);  kind 1: DenseBernoulliDecoder
 * loss = softplus(y) - t*y, dy = sigmoid(y)-t, rec = sigmoid(y).  target row = m / tgt_div. */
int pd_scalar_head_loss(pd_handle* h, long M, int kind, const float* y, const float* target, int tgt_div,
                        float* loss, float* dy, float* rec, void* stream);

/* World-model loss assembly (dreamer.py:362-379, decoders.py:50-108): per (t,b) over I samples.
 * in: per-row losses [TB*I]; out: w[TB*I] = softmax_i(-L)/(TB) (grad of loss_model wrt L_tbi),
 * tb[TB,8] = {loss_model, loss_image, loss_reward, loss_terminal, loss_kl(exact), ent_prior, ent_post, 0}. */
int pd_wm_loss(pd_handle* h, int TB, int I, float kl_weight, float w_img, float w_rew, float w_term,
               const float* l_img, const float* l_rew, const float* l_term, const float* l_kl,
               const float* kl_exact, const float* ent_prior, const float* ent_post,
               float* w, float* tb, void* stream);
/* out[c] = mean over rows of x[M,N]  (N <= 32) */
int pd_colmean(pd_handle* h, long M, int N, const float* x, float* out, void* stream);

/* ---- actor-critic (a2c.py:61-149) ---------------------------------------------------------- */
/* Column-wise GAE(lambda) scan + reality weights + critic loss/grad.  J = H+1 rows of Md columns.
 * vt = critic_target values, v = critic values, rew = reward head output, term_logit = terminal logits.
 * outputs: adv, agae, target, weight [H,Md]; dv [H,Md] = d loss_critic / d v; term [J,Md] = sigmoid;
 * sums[5] (double) += {loss_critic*HMd, sum v0[0], sum v0, sum r1, sum r1^2}. */
int pd_gae_critic(pd_handle* h, int H, int Md, float gamma, float lambda,
                  const float* vt, const float* v, const float* rew, const float* term_logit,
                  float* term, float* adv, float* agae, float* target, float* weight, float* dv,
                  double* sums, void* stream);
/* reinforce actor loss, one-hot policy (a2c.py:119-130): rows = H*Md.
 * sums[2] (double) += {sum (loss_policy - eta*ent)*w, sum ent};  dlogits = d mean(.)/d logits. */
int pd_actor_loss_onehot(pd_handle* h, long rows, int A, float eta, const float* logits, long ldl,
                         const float* actions, long lda, const float* agae, const float* weight,
                         float* dlogits, long lddl, double* sums, void* stream);
/* tanh_normal policy (functions.py:69-78): out[rows,2A] = (mean_, std_) raw, action in (-1,1). */
int pd_actor_loss_tanh_normal(pd_handle* h, long rows, int A, float eta, const float* out, long ldo,
                              const float* actions, long lda, const float* agae, const float* weight,
                              float* dout, long lddo, double* sums, void* stream);
/* a = tanh(5 tanh(m/5) + (softplus(s)+0.1) * eps)   (dreamer.py:197-200 with tanh_normal) */
int pd_tanh_normal_sample(pd_handle* h, long rows, int A, const float* out, long ldo,
                          const float* eps, float* action, long lda, void* stream);

/* ---- replay preprocessing on the device (pydreamer/preprocessing.py:91-188) ------------ */
/* uint8 image (NB,H,W,C) -> fp32 (NB,C,H,W) = x/255 - 0.5 (preprocessing.py:21-29): the batch crosses PCIe as bytes. */
int pd_image_u8_to_f32(pd_handle* h, long NB, int H, int W, int C, const uint8_t* src, float* dst, void* stream);
/* int64 action index -> one-hot fp32 (preprocessing.py:135-138) */
int pd_onehot_i64(pd_handle* h, long rows, int A, const int64_t* idx, float* out, void* stream);
/* y = tanh(x): clip_rewards 'tanh' (functions.py:153-160) */
int pd_tanh(pd_handle* h, long n, const float* x, float* y, void* stream);

/* ---- optimizer (dreamer.py:60-87, train.py:193-198) ---------------------------------------- */
/* out += sum x^2 in a fixed summation order (bit-identical on every data-parallel replica); ws: pd_sumsq_ws_floats(h) floats */
int pd_sumsq(pd_handle* h, const float* x, long n, float* out /* += */, float* ws, void* stream);
int pd_sumsq_ws_floats(const pd_handle* h);
/* norm = sqrt(*sumsq); coef = min(1, max_norm/(norm+1e-6)); x *= coef; *norm_out = norm
 * (== torch.nn.utils.clip_grad_norm_). */
int pd_clip_scale(pd_handle* h, float* x, long n, const float* sumsq, float max_norm, float* norm_out, void* stream);
/* x *= alpha * (*scale) without operand rounding (scale may be NULL = 1); exits without touching memory when the factor is
 * exactly 1.  Applies the grad_output of `loss.backward(g)` (train.py:184-187, GradScaler under amp) and the 1/world of the
 * data-parallel mean to the gradient arena. */
int pd_scale_by(pd_handle* h, float* x, long n, const float* scale, float alpha, void* stream);
/* torch.optim.AdamW (decoupled wd, no amsgrad); *step is a device int32 already incremented. */
int pd_adamw(pd_handle* h, float* p, const float* g, float* m, float* v, long n,
             float lr, float beta1, float beta2, float eps, float wd, const int32_t* step, void* stream);
int pd_inc(pd_handle* h, int32_t* counter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PD_B200_H */
