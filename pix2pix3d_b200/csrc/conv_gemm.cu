// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05, sm_100a) for the modulated 3x3 / 1x1
// convolutions of the synthesis and super-resolution blocks (reference: training/networks_stylegan2.py:34-91,
// torch_utils/ops/conv2d_resample.py:48-143, which bottom out in cuDNN).
//
//   D[128 pixels x BN channels] (TMEM, fp32) += A[128 x 64] (smem) * B[BN x 64]^T (smem), fp16 operands
//
// * Activations are NHWC fp16, so the im2col operand of one filter tap is a plain 4-D TMA box
//   {64 channels, BW, BH, 1 image} shifted by the tap offset; out-of-bounds coordinates are zero-filled by the
//   TMA unit, which implements the convolution padding (and the borders of the transposed-conv phases) for free.
// * Weights are pre-modulated per sample (w * style * demod, fused_modconv semantics) and stored K-major
//   [B][Cout][taps*Cin] fp16.
// * fp32 layers (the tri-plane backbone) run as three fp16 passes over split operands, x = hi + lo,
//   hi*hi + hi*lo + lo*hi, accumulated in the same fp32 TMEM tile (error ~2^-22, products are exact in fp32).
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer, warps 2-5 = epilogue
//   (tcgen05.ld -> scale/noise/bias/activation/clamp -> NHWC store). mbarrier ring of kStages smem stages.
#include <stdlib.h>
#include <string.h>
#include "p3d_common.cuh"
#include "tc05.cuh"

namespace p3d {

constexpr int kBM = 128;        // pixels per tile (= TMEM lanes)
constexpr int kBK = 64;         // fp16 elements per k-step (128-byte swizzle span)
constexpr int kMaxGroups = 27;  // 9 taps x 3 precision passes

struct ConvKernelArgs {
    // k-loop: groups of (tap, A plane, B plane); each group covers Cin channels in kc steps
    int n_groups, kc_steps;
    int8_t dy[kMaxGroups], dx[kMaxGroups], a_plane[kMaxGroups], b_plane[kMaxGroups], tap[kMaxGroups];
    int Cin;
    // tiling
    int BW, BH, tiles_x, tiles_y, BN, w_per_sample;
    int MT, NT;                       // 128-row sub-tiles per CTA along pixels / output channels (operand reuse in smem)
    uint32_t idesc, tmem_cols;
    // epilogue
    int gH, gW;                       // size of the computed grid (phase grid for transposed conv)
    int oH, oW, sy, oy, sx, ox;       // output tensor size and affine map (Y = y*sy + oy)
    int Cout, y_cstride, y_coff;      // valid channels, channel stride of the output tensor, channel offset
    void* y; void* y_lo;
    int out_mode;                     // 0: f16, 1: f16 hi/lo split, 2: f32, 3: f32 accumulate (+=)
    const float* bias; const float* noise; const float* dscale;
    int act; float alpha, gain, clamp, acc_scale;
};

template <int kStages>
__global__ void __launch_bounds__(192, 2) conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const ConvKernelArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages][A 16 KB][B BN*128 B] | barriers | tmem ptr
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = (uint32_t)a.MT * kBM * 128, b_bytes = (uint32_t)a.NT * a.BN * 128;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x, tile_n = blockIdx.y, b = blockIdx.z;
    const int ty = tile_m / a.tiles_x, tx = tile_m % a.tiles_x;
    const int n0 = tile_n * a.BN * a.NT;
    const int total_k = a.n_groups * a.kc_steps;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
        for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
        tc::mbar_init(tmem_full_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr_smem, a.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int g = 0; g < a.n_groups; ++g) {
                const int x0 = tx * a.BW + a.dx[g], y0 = ty * a.BH * a.MT + a.dy[g];
                const int kb = a.tap[g] * a.Cin;
                for (int kc = 0; kc < a.kc_steps; ++kc) {
                    tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    tc::mbar_expect_tx(&full_bar[stage], stage_bytes);
                    tc::tma_load_5d(sa, &tmA, &full_bar[stage], kc * kBK, x0, y0, b, a.a_plane[g]);
                    tc::tma_load_4d(sb, &tmB, &full_bar[stage], kb + kc * kBK, n0, a.w_per_sample ? b : 0, a.b_plane[g]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int k = 0; k < total_k; ++k) {
                tc::mbar_wait(&full_bar[stage], phase);
                tc::tc_fence_after();
                const uint32_t sa = tc::smem_u32(smem + stage * stage_bytes);
                const uint64_t da = tc::umma_desc_k128(sa), db = tc::umma_desc_k128(sa + a_bytes);
                // every (pixel sub-tile, channel sub-tile) pair has its own 128-column accumulator; operands are shared
                for (int mt = 0; mt < a.MT; ++mt)
                    for (int nt = 0; nt < a.NT; ++nt) {
                        const uint64_t dam = da + (uint64_t)(mt * (kBM * 128 >> 4)), dbn = db + (uint64_t)(nt * (a.BN * 128 >> 4));
                        const uint32_t acc = tmem_base + (uint32_t)((mt * a.NT + nt) * 128);
#pragma unroll
                        for (int j = 0; j < kBK / 16; ++j)   // 4 MMAs of K=16: advance 32 bytes inside the swizzle span
                            tc::umma_f16(acc, dam + (uint64_t)(j * 2), dbn + (uint64_t)(j * 2), a.idesc, (k | j) != 0);
                    }
                tc::umma_commit(&empty_bar[stage]);                 // frees the smem stage when these MMAs retire
                if (k == total_k - 1) tc::umma_commit(tmem_full_bar);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
        const int q = warp & 3;
        const int m = q * 32 + lane;                    // tile row = TMEM lane
        tc::mbar_wait(tmem_full_bar, 0);
        tc::tc_fence_after();
        for (int mt = 0; mt < a.MT; ++mt) {
        const int gy = (ty * a.MT + mt) * a.BH + m / a.BW, gx = tx * a.BW + m % a.BW;
        const bool pix_ok = (gy < a.gH) && (gx < a.gW);
        const int Y = gy * a.sy + a.oy, X = gx * a.sx + a.ox;
        const size_t pix = ((size_t)b * a.oH + Y) * a.oW + X;
        const float nz = (a.noise && pix_ok) ? __ldg(a.noise + (size_t)Y * a.oW + X) : 0.f;
        for (int nt = 0; nt < a.NT; ++nt)
        for (int c0 = 0; c0 < a.BN; c0 += 32) {
            uint32_t v[32];
            tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((mt * a.NT + nt) * 128 + c0), v);
            tc::tmem_ld_wait();
            if (!pix_ok) continue;
            const int ch0 = n0 + nt * a.BN + c0;
            if (ch0 >= a.Cout) continue;
            float r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int ch = ch0 + i;
                float x = __uint_as_float(v[i]) * a.acc_scale;
                if (ch < a.Cout) {
                    if (a.dscale) x *= __ldg(a.dscale + (size_t)b * a.Cout + ch);
                    x += nz;
                    if (a.bias) x += __ldg(a.bias + ch);
                    if (a.act == 3) x = x > 0.f ? x : x * a.alpha;
                    x *= a.gain;
                    if (a.clamp >= 0.f) x = fminf(fmaxf(x, -a.clamp), a.clamp);
                } else {
                    x = 0.f;
                }
                r[i] = x;
            }
            const int nvalid = min(32, a.Cout - ch0);
            const size_t off = pix * a.y_cstride + a.y_coff + ch0;
            if (a.out_mode == 0 || a.out_mode == 1) {
                __half* yh = reinterpret_cast<__half*>(a.y) + off;
                __half* yl = a.out_mode == 1 ? reinterpret_cast<__half*>(a.y_lo) + off : nullptr;
                if (nvalid == 32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __align__(16) __half2 hv[4], lv[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            float f0 = r[j * 8 + 2 * t], f1 = r[j * 8 + 2 * t + 1];
                            __half h0 = __float2half_rn(f0), h1 = __float2half_rn(f1);
                            hv[t] = __halves2half2(h0, h1);
                            lv[t] = __halves2half2(__float2half_rn(f0 - __half2float(h0)), __float2half_rn(f1 - __half2float(h1)));
                        }
                        *reinterpret_cast<uint4*>(yh + j * 8) = *reinterpret_cast<uint4*>(hv);
                        if (yl) *reinterpret_cast<uint4*>(yl + j * 8) = *reinterpret_cast<uint4*>(lv);
                    }
                } else {
                    for (int i = 0; i < nvalid; ++i) {
                        __half h = __float2half_rn(r[i]);
                        yh[i] = h;
                        if (yl) yl[i] = __float2half_rn(r[i] - __half2float(h));
                    }
                }
            } else {
                float* yf = reinterpret_cast<float*>(a.y) + off;
                if (nvalid == 32 && (a.y_cstride & 3) == 0 && ((a.y_coff + ch0) & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                        if (a.out_mode == 3) {
                            float4 p = *reinterpret_cast<float4*>(yf + 4 * j);
                            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                        }
                        *reinterpret_cast<float4*>(yf + 4 * j) = o;
                    }
                } else {
                    for (int i = 0; i < nvalid; ++i) yf[i] = (a.out_mode == 3 ? yf[i] : 0.f) + r[i];
                }
            }
        }
        }   // mt
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return P3D_UNSUPPORTED;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? P3D_OK : P3D_BAD_ARG;
}

}  // namespace p3d

using namespace p3d;

extern "C" int p3d_conv_gemm(const p3d_conv_args_t* p, p3d_stream_t stream) {
    if (!p || !p->x || !p->w || !p->y) return P3D_BAD_ARG;
    if (p->C % kBK != 0 || p->C <= 0 || p->n_taps < 1 || p->n_taps > 9) return P3D_UNSUPPORTED;
    if (p->x_planes < 1 || p->x_planes > 2 || p->w_planes < 1 || p->w_planes > 2) return P3D_BAD_ARG;
    if (p->split && (p->x_planes != 2 || p->w_planes != 2)) return P3D_BAD_ARG;
    if (p->out_mode < 0 || p->out_mode > 3 || (p->out_mode == 1 && !p->y_lo)) return P3D_BAD_ARG;
    if (p->Cout_padded % 16 != 0 || p->Cout_padded < p->Cout) return P3D_BAD_ARG;
    if (p->gH <= 0 || p->gW <= 0 || p->B <= 0) return P3D_BAD_ARG;

    // spatial tile BW x BH = 128 pixels: the power-of-two split with the least padded area (ties -> wider rows)
    int BW = 1;
    {
        long best = -1;
        for (int bw = 1; bw <= 64; bw *= 2) {
            int bh = kBM / bw;
            long area = (long)ceil_div(p->gW, bw) * bw * ceil_div(p->gH, bh) * bh;
            if (best < 0 || area <= best) { best = area; BW = bw; }
        }
    }
    const int BH = kBM / BW;
    const int BN = p->Cout_padded > 128 ? 128 : p->Cout_padded;
    // operand reuse: up to 2 x 2 sub-tiles per CTA (A and B stages are shared by the accumulators), as long as the
    // grid still fills the machine; it halves the L2 -> shared-memory traffic per FLOP, which is what bounds this kernel
    // (measured on B200, profiles/r01_conv_subtiles.md: 1x1 with two co-resident CTAs per SM beats 2x1 / 2x2 with one
    //  CTA per SM by ~1.5x, because the second CTA hides the epilogue and the TMA latency; so 1x1 is the default and the
    //  larger shapes stay selectable through P3D_CONV_MT / P3D_CONV_NT for experiments.)
    int MT = 1, NT = 1;
    {
        const char* emt = getenv("P3D_CONV_MT");
        const char* ent = getenv("P3D_CONV_NT");
        const long ctas1 = (long)ceil_div(p->gW, BW) * ceil_div(p->gH, BH) * ceil_div(p->Cout_padded, BN) * p->B;
        const int nsm = sm_count();
        if (ent && atoi(ent) == 2 && BN == 128 && p->Cout_padded >= 256 && ctas1 / 2 >= nsm) NT = 2;
        if (emt && atoi(emt) == 2 && 2 * BH <= 256 && p->gH >= 2 * BH && ctas1 / (2 * NT) >= nsm) MT = 2;
    }
    const int K = p->n_kblocks * p->C;
    if (p->n_kblocks < 1) return P3D_BAD_ARG;

    CUtensorMap tmA, tmB;
    {
        uint64_t dims[5] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B, (uint64_t)p->x_planes};
        uint64_t str[4] = {(uint64_t)p->C * 2, (uint64_t)p->W * p->C * 2, (uint64_t)p->H * p->W * p->C * 2,
                           (uint64_t)p->B * p->H * p->W * p->C * 2};
        uint32_t box[5] = {(uint32_t)kBK, (uint32_t)BW, (uint32_t)(BH * MT), 1, 1};
        int rc = make_tmap(&tmA, p->x, 5, dims, str, box);
        if (rc != P3D_OK) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)K, (uint64_t)p->Cout_padded, (uint64_t)p->Bw, (uint64_t)p->w_planes};
        uint64_t str[3] = {(uint64_t)K * 2, (uint64_t)p->Cout_padded * K * 2, (uint64_t)p->Bw * p->Cout_padded * K * 2};
        uint32_t box[4] = {(uint32_t)kBK, (uint32_t)(BN * NT), 1, 1};
        int rc = make_tmap(&tmB, p->w, 4, dims, str, box);
        if (rc != P3D_OK) return rc;
    }

    ConvKernelArgs a;
    memset(&a, 0, sizeof(a));
    int g = 0;
    const int passes = p->split ? 3 : 1;
    static const int pa[3] = {0, 0, 1}, pb[3] = {0, 1, 0};     // hi*hi, hi*lo, lo*hi
    for (int pass = 0; pass < passes; ++pass)
        for (int t = 0; t < p->n_taps; ++t) {
            a.dy[g] = p->tap_dy[t]; a.dx[g] = p->tap_dx[t]; a.tap[g] = p->tap_k[t];
            a.a_plane[g] = (int8_t)pa[pass]; a.b_plane[g] = (int8_t)pb[pass];
            ++g;
        }
    a.n_groups = g;
    a.kc_steps = p->C / kBK;
    a.Cin = p->C;
    a.BW = BW; a.BH = BH;
    a.tiles_x = ceil_div(p->gW, BW); a.tiles_y = ceil_div(p->gH, BH * MT);
    a.MT = MT; a.NT = NT;
    a.BN = BN; a.w_per_sample = p->Bw > 1 ? 1 : 0;
    a.idesc = tc::umma_idesc_f16(kBM, BN, 0);
    a.tmem_cols = (MT * NT > 1) ? (uint32_t)(MT * NT * 128) : (BN <= 32 ? 32u : BN <= 64 ? 64u : 128u);
    if (a.tmem_cols == 384) a.tmem_cols = 512;
    a.gH = p->gH; a.gW = p->gW; a.oH = p->oH; a.oW = p->oW; a.sy = p->sy; a.oy = p->oy; a.sx = p->sx; a.ox = p->ox;
    a.Cout = p->Cout; a.y_cstride = p->y_cstride; a.y_coff = p->y_coff;
    a.y = p->y; a.y_lo = p->y_lo; a.out_mode = p->out_mode;
    a.bias = p->bias; a.noise = p->noise; a.dscale = p->dscale;
    a.act = p->act; a.alpha = p->alpha; a.gain = p->gain; a.clamp = p->clamp; a.acc_scale = p->acc_scale;

    constexpr int kStages = 3;   // 3 stages: with 1x1 sub-tiles (96 KB) two CTAs per SM overlap epilogue and MMA phases
    const size_t smem = (size_t)kStages * ((size_t)MT * kBM * 128 + (size_t)NT * BN * 128) + (2 * kStages + 1) * 8 + 16 + 1024;
    P3D_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_kernel<kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(a.tiles_x * a.tiles_y, ceil_div(p->Cout_padded, BN * NT), p->B);
    conv_gemm_kernel<kStages><<<grid, 192, smem, (cudaStream_t)stream>>>(tmA, tmB, a);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}
