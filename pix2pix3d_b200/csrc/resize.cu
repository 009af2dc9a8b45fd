// Bilinear resize with optional anti-aliasing: F.interpolate(x, size, mode='bilinear', align_corners=False, antialias=...)
// as the reference uses it in Superresolution*.forward (training/superresolution.py:315-319) and `filtered_resizing`
// (training/dual_discriminator.py:86-102; both directions: raw render 128 -> 512 for D, real image 512 -> 128 for the loss).
// The op is separable and banded: output o of an axis reads the source indices [start(o), start(o) + count(o)) with
// normalised triangle weights (ATen's `_compute_indices_min_size_weights_aa`: support = max(scale, 1), centre
// scale * (o + 0.5)), or the two clamped taps of plain bilinear interpolation when antialias is off. The adjoint (the
// op's backward, and through self-recursion every higher order) is the same kernel with the bands of the transposed
// matrix: source index i gathers from the contiguous range of outputs whose window contains it.
// HBM streaming: numel_in * sizeof + numel_out * sizeof per launch (SURVEY 8d "resize").
#include "p3d_common.cuh"

namespace p3d {

struct ResizeAxis {
    int in_size, out_size;      // sizes of the FORWARD op along this axis
    int antialias, K;           // K: band capacity (shared-memory stride) of this launch
    float scale, support, invscale;
};

struct ResizeParams {
    ResizeAxis ay, ax;
    int transposed;
    int src_h, src_w, dst_h, dst_w;
    int tiles_x, tiles_y;
    int plane_chunks;           // blocks per tile: a block walks planes plane0, plane0 + plane_chunks, ... with ONE set of bands
    long long planes;
};

constexpr int kRzTX = 32, kRzTY = 8;

__host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ float tri(float v) { v = fabsf(v); return v < 1.f ? 1.f - v : 0.f; }

// Window of forward output o: [lo, lo + n) in the source, and the normaliser of its weights.
__host__ __device__ __forceinline__ void fwd_window(const ResizeAxis& a, int o, int& lo, int& n, float& centre, float& total) {
    if (a.antialias) {
        centre = a.scale * ((float)o + 0.5f);
        lo = imax((int)(centre - a.support + 0.5f), 0);
        n = imin((int)(centre + a.support + 0.5f), a.in_size) - lo;
        total = 0.f;
        for (int j = 0; j < n; ++j) total += tri(((float)(j + lo) - centre + 0.5f) * a.invscale);
    } else {
        // area_pixel_compute_source_index + guard_index_and_lambda (ATen UpSample.h)
        const float src = fmaxf(a.scale * ((float)o + 0.5f) - 0.5f, 0.f);
        lo = imin((int)src, a.in_size - 1);
        centre = fminf(fmaxf(src - (float)lo, 0.f), 1.f);      // lambda
        n = lo + 1 < a.in_size ? 2 : 1;
        total = 1.f;
    }
}
__host__ __device__ __forceinline__ float fwd_raw(const ResizeAxis& a, int lo, int n, float centre, int i) {
    if (a.antialias) return tri(((float)i - centre + 0.5f) * a.invscale);
    if (n == 1) return 1.f;
    return i == lo ? 1.f - centre : centre;
}

// Band of destination index r for this launch: forward -> r is an output index, band over the input;
// transposed -> r is an input index, band over the outputs whose window contains r.
__host__ __device__ inline void make_band(const ResizeAxis& a, int transposed, int r, int& start, int& count, float* w) {
    if (!transposed) {
        int lo, n; float c, tot;
        fwd_window(a, r, lo, n, c, tot);
        n = imin(n, a.K);
        const float inv = tot != 0.f ? 1.f / tot : 0.f;
        for (int j = 0; j < n; ++j) w[j] = fwd_raw(a, lo, n, c, lo + j) * inv;
        start = lo; count = n;
        return;
    }
    const float sup = a.antialias ? a.support : 1.f;
    int o_lo = (int)floorf(((float)r - sup - 0.5f) / a.scale - 1.5f);
    int o_hi = (int)ceilf(((float)r + sup + 1.5f) / a.scale + 0.5f);
    o_lo = imax(o_lo, 0); o_hi = imin(o_hi, a.out_size - 1);
    start = 0; count = 0;
    for (int o = o_lo; o <= o_hi && count < a.K; ++o) {
        int lo, n; float c, tot;
        fwd_window(a, o, lo, n, c, tot);
        const bool inside = r >= lo && r < lo + n;
        if (!inside) { if (count > 0) break; continue; }
        if (count == 0) start = o;
        w[count++] = tot != 0.f ? fwd_raw(a, lo, n, c, r) / tot : 0.f;
    }
}

template <typename T, typename Acc>
__global__ void __launch_bounds__(kRzTX * kRzTY) resize_bilinear_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                          const ResizeParams p) {
    extern __shared__ __align__(16) unsigned char rz_smem[];
    int* sx = reinterpret_cast<int*>(rz_smem);
    int* cx = sx + kRzTX;
    int* sy = cx + kRzTX;
    int* cy = sy + kRzTY;
    float* wx = reinterpret_cast<float*>(cy + kRzTY);
    float* wy = wx + kRzTX * p.ax.K;
    const int tid = threadIdx.x;
    const long long blk = blockIdx.x;
    const int tx = (int)(blk % p.tiles_x), ty = (int)((blk / p.tiles_x) % p.tiles_y);
    const long long plane0 = blk / ((long long)p.tiles_x * p.tiles_y);
    if (tid < kRzTX) {
        const int r = tx * kRzTX + tid;
        int s = 0, c = 0;
        if (r < p.dst_w) make_band(p.ax, p.transposed, r, s, c, wx + tid * p.ax.K);
        sx[tid] = s; cx[tid] = c;
    } else if (tid < kRzTX + kRzTY) {
        const int t = tid - kRzTX, r = ty * kRzTY + t;
        int s = 0, c = 0;
        if (r < p.dst_h) make_band(p.ay, p.transposed, r, s, c, wy + t * p.ay.K);
        sy[t] = s; cy[t] = c;
    }
    __syncthreads();
    const int lx = tid % kRzTX, ly = tid / kRzTX;
    const int ox = tx * kRzTX + lx, oy = ty * kRzTY + ly;
    if (ox >= p.dst_w || oy >= p.dst_h) return;
    const float* wyr = wy + ly * p.ay.K;
    const float* wxr = wx + lx * p.ax.K;
    const int ny = cy[ly], nx = cx[lx];
    const long long src_off = (long long)sy[ly] * p.src_w + sx[lx], dst_off = (long long)oy * p.dst_w + ox;
    const long long src_plane = (long long)p.src_h * p.src_w, dst_plane = (long long)p.dst_h * p.dst_w;
    // the bands depend on the axis sizes only: every plane of this tile reuses them
    for (long long plane = plane0; plane < p.planes; plane += p.plane_chunks) {
        const T* src = x + plane * src_plane + src_off;
        Acc acc = (Acc)0;
        for (int ky = 0; ky < ny; ++ky) {
            Acc row = (Acc)0;
            for (int kx = 0; kx < nx; ++kx) row += (Acc)wxr[kx] * (Acc)src[kx];
            acc += (Acc)wyr[ky] * row;
            src += p.src_w;
        }
        y[plane * dst_plane + dst_off] = (T)acc;
    }
}

static ResizeAxis make_axis(int in_size, int out_size, int antialias, int transposed) {
    ResizeAxis a;
    a.in_size = in_size; a.out_size = out_size; a.antialias = antialias;
    a.scale = (float)in_size / (float)out_size;
    a.support = a.scale >= 1.f ? a.scale : 1.f;
    a.invscale = a.scale >= 1.f ? 1.f / a.scale : 1.f;
    const float sup = antialias ? a.support : 1.f;
    if (!transposed) a.K = antialias ? 2 * (int)ceilf(a.support) + 2 : 2;
    else a.K = (int)ceilf((2.f * sup + 2.f) / a.scale) + 3;
    return a;
}

template <typename T, typename Acc>
static int launch_resize(const void* x, void* y, const ResizeParams& p, cudaStream_t stream) {
    const size_t smem = (size_t)(2 * kRzTX + 2 * kRzTY) * sizeof(int) + (size_t)(kRzTX * p.ax.K + kRzTY * p.ay.K) * sizeof(float);
    if (smem > 200 * 1024) return P3D_UNSUPPORTED;
    if (smem > 48 * 1024)
        P3D_CUDA_TRY(cudaFuncSetAttribute(resize_bilinear_kernel<T, Acc>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (long long)p.plane_chunks * p.tiles_x * p.tiles_y;
    if (blocks > 0x7fffffffLL) return P3D_UNSUPPORTED;
    resize_bilinear_kernel<T, Acc><<<(unsigned)blocks, kRzTX * kRzTY, smem, stream>>>((const T*)x, (T*)y, p);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

}  // namespace p3d

using namespace p3d;

extern "C" int p3d_resize_bilinear(const void* x, void* y, int dtype, int64_t planes, int in_h, int in_w, int out_h, int out_w,
                                   int antialias, int transposed, p3d_stream_t stream) {
    if (planes < 0 || in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0) return P3D_BAD_ARG;
    if (planes == 0) return P3D_OK;
    if (!x || !y) return P3D_BAD_ARG;
    ResizeParams p;
    p.ay = make_axis(in_h, out_h, antialias ? 1 : 0, transposed ? 1 : 0);
    p.ax = make_axis(in_w, out_w, antialias ? 1 : 0, transposed ? 1 : 0);
    p.transposed = transposed ? 1 : 0;
    p.src_h = transposed ? out_h : in_h; p.src_w = transposed ? out_w : in_w;
    p.dst_h = transposed ? in_h : out_h; p.dst_w = transposed ? in_w : out_w;
    p.tiles_x = ceil_div(p.dst_w, kRzTX); p.tiles_y = ceil_div(p.dst_h, kRzTY);
    p.planes = planes;
    {
        // enough blocks for ~4 waves of 8 resident blocks per SM, the rest of the planes inside the block
        const long long tiles = (long long)p.tiles_x * p.tiles_y, want = (long long)sm_count() * 32;
        long long chunks = (want + tiles - 1) / tiles;
        if (chunks < 1) chunks = 1;
        if (chunks > planes) chunks = planes;
        p.plane_chunks = (int)chunks;
    }
    switch (dtype) {
        case P3D_F32: return launch_resize<float, float>(x, y, p, (cudaStream_t)stream);
        case P3D_F16: return launch_resize<__half, float>(x, y, p, (cudaStream_t)stream);
        case P3D_F64: return launch_resize<double, double>(x, y, p, (cudaStream_t)stream);
        default: return P3D_BAD_ARG;
    }
}
