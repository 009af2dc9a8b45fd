// Memory-bound helpers around the tensor-core convolution (conv_gemm.cu): per-sample weight modulation,
// NCHW<->NHWC converters, the NHWC FIR + noise + bias + lrelu tail of an up=2 layer and the skip-image upsampler.
// All are HBM-streaming kernels with 8/16-byte vector accesses along the (contiguous) channel axis.
#include "p3d_common.cuh"

namespace p3d {

__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// ---------------------------------------------------------------------------------------------
// modulated_conv2d weight preparation (networks_stylegan2.py:58-67)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) modulate_weights_kernel(const float* __restrict__ weight, const float* __restrict__ styles,
                                                               int Cout, int Cin, int ktaps, int Cout_p, int Cin_p, int cin_off, int demod,
                                                               float pre_scale, float out_scale, int planes, int B,
                                                               __half* __restrict__ out) {
    const int o = blockIdx.x, b = blockIdx.y;
    const size_t row = ((size_t)b * Cout_p + o) * (size_t)ktaps * Cin_p;
    const size_t plane_stride = (size_t)B * Cout_p * ktaps * Cin_p;
    __shared__ float red[8];
    __shared__ float dcoef;
    const int n = Cin * ktaps;
    if (o >= Cout) {
        for (int idx = threadIdx.x; idx < ktaps * Cin_p; idx += blockDim.x) {
            out[row + idx] = __float2half_rn(0.f);
            if (planes == 2) out[plane_stride + row + idx] = __float2half_rn(0.f);
        }
        return;
    }
    const float* w = weight + (size_t)o * n;   // [Cin][ktaps]
    const float* s = styles + (size_t)b * Cin;
    float acc = 0.f;
    if (demod) {
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            float v = __ldg(w + idx) * (__ldg(s + idx / ktaps) * pre_scale);
            acc = fmaf(v, v, acc);
        }
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
            dcoef = rsqrtf(t + 1e-8f);
        }
        __syncthreads();
    }
    const float d = demod ? dcoef : 1.f;
    for (int idx = threadIdx.x; idx < ktaps * Cin_p; idx += blockDim.x) {
        const int t = idx / Cin_p, i = idx % Cin_p - cin_off;
        float v = 0.f;
        if (i >= 0 && i < Cin) v = __ldg(w + (size_t)i * ktaps + t) * (__ldg(s + i) * pre_scale) * d * out_scale;
        __half hi, lo;
        split_half(v, hi, lo);
        out[row + idx] = hi;
        if (planes == 2) out[plane_stride + row + idx] = lo;
    }
}

// ---------------------------------------------------------------------------------------------
// NCHW (fp32 / fp16) -> NHWC fp16 (1 or 2 planes, channel-padded); NHWC fp32 -> NCHW fp32
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void nchw_to_nhwc_f16_kernel(const T* __restrict__ in, __half* __restrict__ out, int C, int HW, int Cp, int planes,
                                        size_t plane_stride) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const T* src = in + (size_t)n * C * HW;
    for (int j = ty; j < 32; j += 8) {
        int c = c0 + j, p = p0 + tx;
        tile[j][tx] = (c < C && p < HW) ? (float)src[(size_t)c * HW + p] : 0.f;
    }
    __syncthreads();
    __half* dst = out + (size_t)n * HW * Cp;
    for (int j = ty; j < 32; j += 8) {
        int p = p0 + j, c = c0 + tx;
        if (c < Cp && p < HW) {
            __half hi, lo;
            split_half(tile[tx][j], hi, lo);
            dst[(size_t)p * Cp + c] = hi;
            if (planes == 2) dst[plane_stride + (size_t)p * Cp + c] = lo;
        }
    }
}

__global__ void nhwc_to_nchw_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW, int cstride, int coff) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const float* src = in + (size_t)n * HW * cstride + coff;
    for (int j = ty; j < 32; j += 8) {
        int p = p0 + j, c = c0 + tx;
        tile[j][tx] = (c < C && p < HW) ? src[(size_t)p * cstride + c] : 0.f;
    }
    __syncthreads();
    float* dst = out + (size_t)n * C * HW;
    for (int j = ty; j < 32; j += 8) {
        int c = c0 + j, p = p0 + tx;
        if (c < C && p < HW) dst[(size_t)c * HW + p] = tile[tx][j];
    }
}

// ---------------------------------------------------------------------------------------------
// NHWC 4x4 FIR + noise + bias + lrelu + gain + clamp (tail of an up=2 SynthesisLayer)
// ---------------------------------------------------------------------------------------------
template <class TIn, int VEC>
__device__ __forceinline__ void load_vec(const TIn* p, float (&v)[VEC]);
template <>
__device__ __forceinline__ void load_vec<float, 4>(const float* p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load_vec<__half, 8>(const __half* p, float (&v)[8]) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// Shared-memory tiled version: a CTA produces an 8 x 16 pixel tile for a 128-byte channel block (64 fp16 or 32 fp32
// channels). The (8+3) x (16+3) input halo is staged once with fully coalesced 16-byte loads (all issued up front),
// then every thread computes a 2x2 output block for one 16-byte channel vector from 25 conflict-free LDS.128.
constexpr int kFirTH = 8, kFirTW = 16, kFirIH = kFirTH + 3, kFirIW = kFirTW + 3;

template <class TIn, int VEC>
__global__ void __launch_bounds__(256) fir_act_nhwc_kernel(const TIn* __restrict__ x, const float* __restrict__ f,
                                                           const float* __restrict__ noise, const float* __restrict__ bias,
                                                           __half* __restrict__ y, int out_planes, size_t out_plane_stride,
                                                           int B, int inH, int inW, int outH, int outW, int C, int padx0,
                                                           int pady0, float fir_gain, int act, float alpha, float act_gain,
                                                           float clamp) {
    __shared__ uint4 tile[kFirIH * kFirIW * 8];      // 209 pixels x 128 bytes
    float ft[4][4];                                  // mirrored taps: true convolution (flip_filter=False)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) ft[j][i] = __ldg(f + (3 - j) * 4 + (3 - i));
    constexpr int CB = 8 * VEC;                      // channels per block (8 vectors of 16 bytes)
    const int tiles_x = (outW + kFirTW - 1) / kFirTW;
    const int tx0 = (blockIdx.x % tiles_x) * kFirTW, ty0 = (blockIdx.x / tiles_x) * kFirTH;
    const int c0 = blockIdx.y * CB, b = blockIdx.z;
    const bool round16 = (sizeof(TIn) == 2);
    // ---- stage the input halo (zero outside the image) ----
    const TIn* xb = x + (size_t)b * inH * inW * C + c0;
    for (int i = threadIdx.x; i < kFirIH * kFirIW * 8; i += 256) {
        const int v = i & 7, p = i >> 3;
        const int iy = ty0 - pady0 + p / kFirIW, ix = tx0 - padx0 + p % kFirIW;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < inH && ix >= 0 && ix < inW)
            val = __ldg(reinterpret_cast<const uint4*>(xb + ((size_t)iy * inW + ix) * C) + v);
        tile[i] = val;
    }
    __syncthreads();
    // ---- 2x2 outputs per thread ----
    const int v = threadIdx.x & 7, blk = threadIdx.x >> 3;          // 32 blocks: 4 rows x 8 cols of 2x2
    const int by = (blk >> 3) * 2, bx = (blk & 7) * 2;
    float acc[2][2][VEC];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[i][j][k] = 0.f;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        float win[5][VEC];
#pragma unroll
        for (int cc = 0; cc < 5; ++cc) {
            const uint4 raw = tile[((by + r) * kFirIW + bx + cc) * 8 + v];
            if (sizeof(TIn) == 2) {
                const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                for (int k = 0; k < VEC / 2; ++k) { float2 t = __half22float2(h[k]); win[cc][2 * k] = t.x; win[cc][2 * k + 1] = t.y; }
            } else {
                const float* fp = reinterpret_cast<const float*>(&raw);
#pragma unroll
                for (int k = 0; k < VEC; ++k) win[cc][k] = fp[k];
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int ty = r - i;
            if (ty < 0 || ty > 3) continue;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int tx = 0; tx < 4; ++tx)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[i][j][k] = fmaf(ft[ty][tx], win[tx + j][k], acc[i][j][k]);
        }
    }
    const int c = c0 + v * VEC;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int py = ty0 + by + i, px = tx0 + bx + j;
            if (py >= outH || px >= outW) continue;
            const float nz = noise ? __ldg(noise + (size_t)py * outW + px) : 0.f;
            const size_t o = (((size_t)b * outH + py) * outW + px) * C + c;
            __align__(16) __half hv[VEC], lv[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float val = acc[i][j][k] * fir_gain;
                if (round16) val = __half2float(__float2half_rn(val));
                val += nz;
                if (round16 && noise) val = __half2float(__float2half_rn(val));
                if (bias) val += __ldg(bias + c + k);
                if (act == 3) val = val > 0.f ? val : val * alpha;
                val *= act_gain;
                if (clamp >= 0.f) val = fminf(fmaxf(val, -clamp), clamp);
                split_half(val, hv[k], lv[k]);
            }
            if (VEC == 8) {
                *reinterpret_cast<uint4*>(y + o) = *reinterpret_cast<const uint4*>(hv);
                if (out_planes == 2) *reinterpret_cast<uint4*>(y + out_plane_stride + o) = *reinterpret_cast<const uint4*>(lv);
            } else {
                *reinterpret_cast<uint2*>(y + o) = *reinterpret_cast<const uint2*>(hv);
                if (out_planes == 2) *reinterpret_cast<uint2*>(y + out_plane_stride + o) = *reinterpret_cast<const uint2*>(lv);
            }
        }
    }
}

// upsample2d(img, [1,3,3,1]) on fp32 NHWC: zero-insert x2, pad (2,1), 4x4 FIR, gain 4 (upfirdn2d.py:344-350)
__global__ void __launch_bounds__(256) upsample2x_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ f,
                                                              float* __restrict__ y, int B, int H, int W, int C) {
    const int64_t total = (int64_t)B * 2 * H * 2 * W * C;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        int64_t t = idx / C;
        const int ox = (int)(t % (2 * W)); t /= 2 * W;
        const int oy = (int)(t % (2 * H));
        const int b = (int)(t / (2 * H));
        const float* xb = x + (size_t)b * H * W * C + c;
        float acc = 0.f;
        // padded/upsampled position p = o + tap - 2 holds x[p/2] when p is even
#pragma unroll
        for (int ty = 0; ty < 4; ++ty) {
            const int py = oy + ty - 2;
            if (py < 0 || (py & 1) || (py >> 1) >= H) continue;
#pragma unroll
            for (int tx = 0; tx < 4; ++tx) {
                const int px = ox + tx - 2;
                if (px < 0 || (px & 1) || (px >> 1) >= W) continue;
                acc = fmaf(__ldg(f + (3 - ty) * 4 + (3 - tx)), __ldg(xb + ((size_t)(py >> 1) * W + (px >> 1)) * C), acc);
            }
        }
        y[idx] = acc * 4.f;
    }
}

static unsigned grid1d(int64_t items, int block) {
    int64_t blocks = ceil_div64(items, block);
    int64_t cap = (int64_t)sm_count() * 32;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace p3d

using namespace p3d;

extern "C" int p3d_modulate_weights(const float* weight, const float* styles, int B, int Cout, int Cin, int ktaps,
                                    int Cout_padded, int Cin_padded, int cin_offset, int demodulate, float pre_scale,
                                    float out_scale, int planes, void* out, p3d_stream_t stream) {
    if (!weight || !styles || !out || B <= 0 || Cout <= 0 || Cin <= 0 || ktaps <= 0) return P3D_BAD_ARG;
    if (Cout_padded < Cout || cin_offset < 0 || Cin_padded < Cin + cin_offset || planes < 1 || planes > 2 || B > 65535) return P3D_BAD_ARG;
    dim3 grid(Cout_padded, B);
    modulate_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(weight, styles, Cout, Cin, ktaps, Cout_padded, Cin_padded,
                                                                    cin_offset, demodulate, pre_scale, out_scale, planes, B, (__half*)out);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

extern "C" int p3d_nchw_to_nhwc_f16(const void* x, int src_dtype, int N, int C, int H, int W, int C_padded, int planes,
                                    void* out, p3d_stream_t stream) {
    if (!x || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0 || C_padded < C || planes < 1 || planes > 2 || N > 65535)
        return P3D_BAD_ARG;
    dim3 grid(ceil_div(H * W, 32), ceil_div(C_padded, 32), N), block(32, 8);
    size_t ps = (size_t)N * H * W * C_padded;
    if (src_dtype == P3D_F32)
        nchw_to_nhwc_f16_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)x, (__half*)out, C, H * W, C_padded, planes, ps);
    else if (src_dtype == P3D_F16)
        nchw_to_nhwc_f16_kernel<__half><<<grid, block, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)out, C, H * W, C_padded, planes, ps);
    else
        return P3D_BAD_ARG;
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

extern "C" int p3d_nhwc_to_nchw_f32(const float* x, int N, int C, int H, int W, int c_stride, int c_offset, float* out,
                                    p3d_stream_t stream) {
    if (!x || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0 || c_stride < C + c_offset || N > 65535) return P3D_BAD_ARG;
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N), block(32, 8);
    nhwc_to_nchw_f32_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, out, C, H * W, c_stride, c_offset);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

extern "C" int p3d_fir_act_nhwc(const void* x, int in_dtype, const float* f, const float* noise, const float* bias, void* y,
                                int out_planes, int B, int inH, int inW, int outH, int outW, int C, int padx0, int pady0,
                                float fir_gain, int act, float alpha, float act_gain, float clamp, p3d_stream_t stream) {
    if (!x || !f || !y || B <= 0 || C <= 0 || out_planes < 1 || out_planes > 2) return P3D_BAD_ARG;
    if (act != 1 && act != 3) return P3D_UNSUPPORTED;
    const size_t ps = (size_t)B * outH * outW * C;
    const int tiles = ceil_div(outW, kFirTW) * ceil_div(outH, kFirTH);
    if (B > 65535) return P3D_UNSUPPORTED;
    if (in_dtype == P3D_F32) {
        if (C % 32) return P3D_UNSUPPORTED;
        dim3 grid(tiles, C / 32, B);
        fir_act_nhwc_kernel<float, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const float*)x, f, noise, bias, (__half*)y, out_planes, ps, B, inH, inW, outH, outW, C, padx0, pady0, fir_gain, act,
            alpha, act_gain, clamp);
    } else if (in_dtype == P3D_F16) {
        if (C % 64) return P3D_UNSUPPORTED;
        dim3 grid(tiles, C / 64, B);
        fir_act_nhwc_kernel<__half, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const __half*)x, f, noise, bias, (__half*)y, out_planes, ps, B, inH, inW, outH, outW, C, padx0, pady0, fir_gain, act,
            alpha, act_gain, clamp);
    } else {
        return P3D_BAD_ARG;
    }
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

extern "C" int p3d_upsample2x_nhwc(const float* x, const float* f, float* y, int B, int H, int W, int C, p3d_stream_t stream) {
    if (!x || !f || !y || B <= 0 || H <= 0 || W <= 0 || C <= 0) return P3D_BAD_ARG;
    int64_t items = (int64_t)B * 4 * H * W * C;
    upsample2x_nhwc_kernel<<<grid1d(items, 256), 256, 0, (cudaStream_t)stream>>>(x, f, y, B, H, W, C);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}
