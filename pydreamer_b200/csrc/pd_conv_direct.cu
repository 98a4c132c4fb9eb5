// pd_conv_direct.cu — direct (no column matrix) kernel for the FIRST encoder convolution (encoders.py:80-81:
// Conv2d(image_channels, d, kernel 4, stride 2) + ELU on the NCHW image).
//
// With K = 16·IC = 48 contraction terms and N = 48 output channels the layer is far below the tensor-core ridge:
// as a GEMM it needs a 461 MB column matrix written and read back and 128-wide tiles that are 62 % empty.  Here one
// block owns ROWS output rows of one image: the 2·ROWS+2 input rows of every channel are staged once in shared
// memory (tf32-rounded, like every tensor-core operand of the other layers), each thread keeps the 16·IC weights of
// its output channel in registers and walks output pixels; the four kw taps of a (channel, kh) row are two 8-byte
// shared loads (broadcast across the channel threads) feeding four FMAs.  HBM traffic = image in + activation out.
#include "pd_common.cuh"

namespace {

constexpr int ROWS = 4;            // output rows per block
constexpr int NTHR = 256;

template <int IC>
__global__ void __launch_bounds__(NTHR)
conv1_direct_fwd_kernel(int Hin, int Win, int Hout, int Wout, int Cout, const float* __restrict__ img,
                        const float* __restrict__ W /* [Cout, IC*16] (c,kh,kw) */, const float* __restrict__ bias,
                        float* __restrict__ out /* NHWC */, int round_ops, int round_out) {
    constexpr int K = IC * 16;
    extern __shared__ float sm[];
    const int in_rows = 2 * ROWS + 2;
    float* win = sm;                                   // [IC][in_rows][Win]
    const int tiles_y = (Hout + ROWS - 1) / ROWS;
    const int n = blockIdx.x / tiles_y, oy0 = (blockIdx.x % tiles_y) * ROWS;
    const int tid = threadIdx.x;

    // stage the input rows 2*oy0 .. 2*oy0 + in_rows - 1 of every channel (rows past the image are never read)
    for (int i = tid; i < IC * in_rows * Win; i += NTHR) {
        const int x = i % Win, r = (i / Win) % in_rows, c = i / (Win * in_rows);
        const int y = 2 * oy0 + r;
        float v = 0.f;
        if (y < Hin) v = img[(((long)n * IC + c) * Hin + y) * Win + x];
        win[i] = pd_round_if(v, round_ops);
    }
    const int groups = NTHR / Cout;                    // pixel groups that work side by side
    const int co = tid % Cout, grp = tid / Cout;
    float w[K];
    if (grp < groups) {
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = W[(long)co * K + k];
    }
    __syncthreads();
    if (grp >= groups) return;
    const float b = bias ? bias[co] : 0.f;
    const int rows_here = min(ROWS, Hout - oy0);
    for (int p = grp; p < rows_here * Wout; p += groups) {
        const int oyl = p / Wout, ox = p - oyl * Wout;
        float acc = b;
#pragma unroll
        for (int c = 0; c < IC; ++c)
#pragma unroll
            for (int kh = 0; kh < 4; ++kh) {
                const float* src = win + (c * in_rows + 2 * oyl + kh) * Win + 2 * ox;     // 8-byte aligned (2*ox even, Win even)
                const float2 p0 = *reinterpret_cast<const float2*>(src);
                const float2 p1 = *reinterpret_cast<const float2*>(src + 2);
                const int k = (c * 4 + kh) * 4;
                acc = fmaf(p0.x, w[k], acc);
                acc = fmaf(p0.y, w[k + 1], acc);
                acc = fmaf(p1.x, w[k + 2], acc);
                acc = fmaf(p1.y, w[k + 3], acc);
            }
        out[(((long)n * Hout + oy0 + oyl) * Wout + ox) * Cout + co] = pd_round_if(pd_elu(acc), round_out && round_ops);
    }
}

}  // namespace

extern "C" int pd_conv1_direct_fwd(pd_handle* h, int NB, int IC, int Hin, int Win, int Cout, const float* img,
                                   const float* W, const float* bias, float* out, int round_out, void* stream) {
    if (!h) return PD_ERR_ARG;
    PD_REQUIRE(h, NB >= 1 && Hin >= 4 && Win >= 4 && (Win % 2) == 0 && Cout >= 1 && Cout <= NTHR,
               "pd_conv1_direct_fwd: bad shape NB=%d H=%d W=%d Cout=%d", NB, Hin, Win, Cout);
    const int Hout = (Hin - 4) / 2 + 1, Wout = (Win - 4) / 2 + 1;
    const int tiles_y = (Hout + ROWS - 1) / ROWS;
    const size_t smem = (size_t)IC * (2 * ROWS + 2) * Win * sizeof(float);
    PD_REQUIRE(h, smem <= 48 * 1024, "pd_conv1_direct_fwd: image rows do not fit shared memory (W=%d)", Win);
    const unsigned grid = (unsigned)((long)NB * tiles_y);
    cudaStream_t s = (cudaStream_t)stream;
    if (IC == 3)
        conv1_direct_fwd_kernel<3><<<grid, NTHR, smem, s>>>(Hin, Win, Hout, Wout, Cout, img, W, bias, out, h->round_ops, round_out);
    else if (IC == 1)
        conv1_direct_fwd_kernel<1><<<grid, NTHR, smem, s>>>(Hin, Win, Hout, Wout, Cout, img, W, bias, out, h->round_ops, round_out);
    else
        PD_FAIL(h, PD_ERR_UNSUPPORTED, "pd_conv1_direct_fwd: image_channels=%d (1 or 3)", IC);
    PD_CHECK_LAUNCH(h, "conv1_direct_fwd");
    return PD_OK;
}
