// Per-pixel cross entropy over the channel axis with a mean over the pixels: `cross_entropy2d` of the reference's
// training/loss_utils.py:4-18 (called from training/loss.py:611-616 on the 512^2 semantic image and the 128^2 raw one).
// The reference transposes the logits to [N*H*W, C] (a full copy) and calls F.cross_entropy(reduction='mean'); here one pass
// reads the NCHW logits in place (online softmax per pixel, coalesced along W), block partials are summed in a fixed order
// in double precision, and the backward writes softmax - onehot in one more pass.
// HBM streaming: forward N*C*H*W*4 + N*H*W*8 bytes, backward twice the logits.
#include "p3d_common.cuh"

namespace p3d {

constexpr int kCeBlock = 256;
constexpr int kCeMaxBlocks = P3D_CE_WORKSPACE_DOUBLES / 2;

struct CeParams {
    const float* x; const int64_t* t; const float* w;
    int C; long long HW, P; long long ignore;
};

// log-sum-exp over the channels of pixel (n, p) and the target logit; returns false for an ignored pixel
__device__ __forceinline__ bool ce_pixel(const CeParams& a, long long idx, float& lse, float& xt, float& wt, int& tgt) {
    const long long t = a.t[idx];
    if (t == a.ignore) return false;
    const long long n = idx / a.HW, p = idx - n * a.HW;
    const float* px = a.x + n * a.C * a.HW + p;
    float m = -INFINITY, s = 0.f;
    xt = __int_as_float(0x7fc00000);                      // stays NaN for a target outside [0, C)
    for (int c = 0; c < a.C; ++c) {
        const float v = __ldg(px + (long long)c * a.HW);
        if (c == t) xt = v;
        const float mn = fmaxf(m, v);
        s = s * expf(m - mn) + expf(v - mn);
        m = mn;
    }
    lse = m + logf(s);
    tgt = (int)t;
    wt = (a.w && t >= 0 && t < a.C) ? __ldg(a.w + t) : 1.f;
    return true;
}

__global__ void __launch_bounds__(kCeBlock) ce2d_fwd_kernel(const CeParams a, double* __restrict__ partial) {
    double sl = 0.0, sw = 0.0;
    for (long long idx = (long long)blockIdx.x * kCeBlock + threadIdx.x; idx < a.P; idx += (long long)gridDim.x * kCeBlock) {
        float lse, xt, wt; int tg;
        if (ce_pixel(a, idx, lse, xt, wt, tg)) { sl += (double)(wt * (lse - xt)); sw += (double)wt; }
    }
    __shared__ double red[2][kCeBlock / 32];
    for (int o = 16; o > 0; o >>= 1) { sl += __shfl_xor_sync(0xffffffffu, sl, o); sw += __shfl_xor_sync(0xffffffffu, sw, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sl; red[1][threadIdx.x >> 5] = sw; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i < kCeBlock / 32; ++i) { a0 += red[0][i]; a1 += red[1][i]; }
        partial[2 * blockIdx.x] = a0; partial[2 * blockIdx.x + 1] = a1;
    }
}

__global__ void ce2d_finish_kernel(const double* __restrict__ partial, int blocks, float* __restrict__ loss, float* __restrict__ sum_w) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double a0 = 0.0, a1 = 0.0;
    for (int i = 0; i < blocks; ++i) { a0 += partial[2 * i]; a1 += partial[2 * i + 1]; }
    *loss = (float)(a0 / a1);           // 0 / 0 = NaN when every pixel is ignored, as F.cross_entropy
    *sum_w = (float)a1;
}

__global__ void __launch_bounds__(kCeBlock) ce2d_bwd_kernel(const CeParams a, const float* __restrict__ grad_loss,
                                                            const float* __restrict__ sum_w, float* __restrict__ gx) {
    const float g = __ldg(grad_loss) / __ldg(sum_w);
    for (long long idx = (long long)blockIdx.x * kCeBlock + threadIdx.x; idx < a.P; idx += (long long)gridDim.x * kCeBlock) {
        const long long n = idx / a.HW, p = idx - n * a.HW;
        const float* px = a.x + n * a.C * a.HW + p;
        float* pg = gx + n * a.C * a.HW + p;
        float lse, xt, wt; int tg;
        if (!ce_pixel(a, idx, lse, xt, wt, tg)) {
            for (int c = 0; c < a.C; ++c) pg[(long long)c * a.HW] = 0.f;
            continue;
        }
        const float gw = g * wt;
        for (int c = 0; c < a.C; ++c) {
            const float sm = expf(__ldg(px + (long long)c * a.HW) - lse);
            pg[(long long)c * a.HW] = gw * (sm - (c == tg ? 1.f : 0.f));
        }
    }
}

static int ce_grid(long long P) {
    long long b = (P + kCeBlock - 1) / kCeBlock;
    const long long cap = (long long)sm_count() * 6 < kCeMaxBlocks ? (long long)sm_count() * 6 : kCeMaxBlocks;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace p3d

using namespace p3d;

extern "C" int p3d_cross_entropy2d_fwd(const float* logits, const int64_t* target, const float* class_weight, int N, int C,
                                       int64_t HW, int64_t ignore_index, float* loss, float* sum_w, double* workspace,
                                       p3d_stream_t stream) {
    if (!logits || !target || !loss || !sum_w || !workspace || N <= 0 || C <= 0 || HW <= 0) return P3D_BAD_ARG;
    CeParams a{logits, target, class_weight, C, (long long)HW, (long long)N * HW, (long long)ignore_index};
    const int grid = ce_grid(a.P);
    ce2d_fwd_kernel<<<grid, kCeBlock, 0, (cudaStream_t)stream>>>(a, workspace);
    P3D_LAUNCH_CHECK();
    ce2d_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(workspace, grid, loss, sum_w);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

extern "C" int p3d_cross_entropy2d_bwd(const float* logits, const int64_t* target, const float* class_weight,
                                       const float* grad_loss, const float* sum_w, int N, int C, int64_t HW,
                                       int64_t ignore_index, float* grad_logits, p3d_stream_t stream) {
    if (!logits || !target || !grad_loss || !sum_w || !grad_logits || N <= 0 || C <= 0 || HW <= 0) return P3D_BAD_ARG;
    CeParams a{logits, target, class_weight, C, (long long)HW, (long long)N * HW, (long long)ignore_index};
    ce2d_bwd_kernel<<<ce_grid(a.P), kCeBlock, 0, (cudaStream_t)stream>>>(a, grad_loss, sum_w, grad_logits);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}
