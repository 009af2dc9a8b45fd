// Tensor-core point query (sm_100a): ImportanceRenderer.run_model at free 3-D points -- tri-plane lookup, mean over the
// planes and the OSG decoder MLP, without a ray marcher (reference training/volumetric_rendering/renderer.py:142-148,
// entered through TriPlane*Generator.sample / sample_mixed, training/triplane_cond.py:1063-1074, e.g. the 512^3 sigma
// grid of applications/extract_mesh.py:60-81).
//   * one CTA per SM, 384 threads = 3 groups of 128; a group walks 128-point tiles on its own (own mbarrier, own named
//     barrier, own TMEM columns), so one group's gather overlaps another group's MMAs and epilogue;
//   * gather and decoder are the renderer's (render_tc.cuh): fp16 hi|lo feature rows in the SWIZZLE_128B K-major layout
//     are the A operand of layer 1, hidden activations return to TMEM as packed fp16 hi/lo and feed layer 2 from there;
//   * sigma is the fp32 dot of the hidden row with the sigma weights (as in the renderer's density passes);
//   * colours are staged through the (now free) feature rows and leave as full 128-byte lines.
#include "render_tc.cuh"

namespace p3d {

struct TcQueryParams {
    const float* planes;
    const float* coords;                // [total, 3]
    const void* decoder_packed;         // p3d_pack_decoder_tc image
    float* out_rgb;                     // [total, 32 * n_nets]
    float* out_sigma;                   // [total]
    long long total;
    long long n_tiles;
    int M, H, W, n_nets, sigma_net;
    uint32_t mask[2];
    float coord_scale;
    uint32_t img_stride, plane_stride, pix_stride;
};

constexpr int kQThreads = 384;
constexpr int kQColsPerGroup = 160;     // D1: [0,128)  D2: [128,160)
constexpr int kQOffFeat = kTcTail;                       // weight tiles first (1024-aligned base)
constexpr int kQOffTail = kQOffFeat + 3 * 16384;
constexpr int kQOffBar = kQOffTail + kTcTailFloats * 4;  // 3 mbarriers + the TMEM base
constexpr int kQSmemBytes = kQOffBar + 3 * 8 + 16;
static_assert(kQOffBar % 8 == 0, "mbarrier alignment");

__global__ void __launch_bounds__(kQThreads, 1) run_model_tc_kernel(const TcQueryParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = tid >> 7, m = tid & 127, q = warp & 3;
    const int n_nets = P.n_nets, sig = P.sigma_net, cout = n_nets * kOut;

    uint8_t* tile = smem + kQOffFeat + grp * 16384;
    const float* tail = reinterpret_cast<const float*>(smem + kQOffTail);
    const float* b1 = tail, *b2c = tail + 128, *b2s = tail + 192, *w2s = tail + 196;
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + kQOffBar);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 3);

    {
        const uint4* src = reinterpret_cast<const uint4*>(P.decoder_packed);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < kTcTail / 16; i += kQThreads) dst[i] = __ldg(src + i);
        const float* tsrc = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(P.decoder_packed) + kTcTail);
        float* tdst = reinterpret_cast<float*>(smem + kQOffTail);
        for (int i = tid; i < kTcTailFloats; i += kQThreads) tdst[i] = __ldg(tsrc + i);
    }
    if (tid == 0) {
        for (int g = 0; g < 3; ++g) tc::mbar_init(&mma_bar[g], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_ptr_smem, 512);
    tc::fence_proxy_async();          // weight tiles were written through the generic proxy
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_grp = *tmem_ptr_smem + (uint32_t)(grp * kQColsPerGroup);
    const uint32_t tmem_row = tmem_grp + ((uint32_t)(q * 32) << 16);
    uint32_t bar_phase = 0;
    const uint32_t w1a = tc::smem_u32(smem + kTcW1A), w1b = tc::smem_u32(smem + kTcW1B);
    const bool colours = P.out_rgb != nullptr;      // sigma-only query (mesh extraction): layer 1 of the sigma net + one dot
    const uint32_t idesc_l1 = (colours && n_nets == 2) ? tc::umma_idesc_f16(128, 128, 0) : tc::umma_idesc_f16(128, 64, 0);
    const int l1_row0 = colours ? 0 : sig * 64;     // first W1 row (= hidden unit) layer 1 computes
    const uint32_t idesc_n32 = tc::umma_idesc_f16(128, 32, 0);
    const uint32_t tile32 = tc::smem_u32(tile);
    const TcPlaneView pv{P.planes, P.H, P.W, P.img_stride, P.plane_stride, P.pix_stride};

    auto group_sync = [&]() { tc::tc_fence_before(); tc::named_bar_sync(1 + grp, 128); tc::tc_fence_after(); };
    auto wait_mma = [&]() { tc::mbar_wait(&mma_bar[grp], bar_phase); bar_phase ^= 1; tc::tc_fence_after(); };

    for (long long t = (long long)blockIdx.x * 3 + grp; t < P.n_tiles; t += (long long)gridDim.x * 3) {
        const long long p = t * 128 + m;
        const bool v = p < P.total;
        float px = 0.f, py = 0.f, pz = 0.f;
        int b = 0;
        if (v) {
            b = (int)(p / P.M);
            const float* c = P.coords + p * 3;
            px = __fmul_rn(P.coord_scale, __ldg(c + 0));
            py = __fmul_rn(P.coord_scale, __ldg(c + 1));
            pz = __fmul_rn(P.coord_scale, __ldg(c + 2));
        }
        tc_gather_rows(pv, tile, q, lane, v, b, px, py, pz);
        tc::fence_proxy_async();
        group_sync();                  // all 128 rows written; every warp is done with D1 / D2 of the previous tile
        if (m == 0) {
            // layer 1: D1[:, 0:64*n_nets) = [hi|lo] x [Whi|Whi]^T + [hi|lo] x [Wlo|0]^T
            const uint64_t da = tc::umma_desc_k128(tile32);
            const uint64_t dba = tc::umma_desc_k128(w1a + l1_row0 * 128), dbb = tc::umma_desc_k128(w1b + l1_row0 * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::umma_f16(tmem_grp, da + 2 * k, dba + 2 * k, idesc_l1, k != 0);
#pragma unroll
            for (int k = 0; k < 2; ++k) tc::umma_f16(tmem_grp, da + 2 * k, dbb + 2 * k, idesc_l1, 1);
            tc::umma_commit(&mma_bar[grp]);
        }
        wait_mma();                    // the feature rows are free from here on (used as the colour staging rows below)
        float sg0 = b2s[sig], sg1 = 0.f;
        if (!colours) {
            // D1[:, 0:64) holds the sigma net's hidden pre-activations: sigma = w2[0] . softplus(h) + b2[0]
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t vv[32];
                tc::tmem_ld_32x32(tmem_row + half * 32, vv);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const int jj = sig * 64 + half * 32 + j;
                    sg0 = fmaf(w2s[jj], softplus2(__uint_as_float(vv[j]) + b1[jj]), sg0);
                    sg1 = fmaf(w2s[jj + 1], softplus2(__uint_as_float(vv[j + 1]) + b1[jj + 1]), sg1);
                }
            }
        }
#pragma unroll 1
        for (int net = 0; colours && net < n_nets; ++net) {
            // hidden: D1[:, net*64 .. +64) -> softplus -> packed fp16 hi (32 cols) | lo (32 cols), in place
            uint32_t ph[32], pl[32];
            const bool is_sig = net == sig;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t vv[32];
                tc::tmem_ld_32x32(tmem_row + net * 64 + half * 32, vv);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const int jj = net * 64 + half * 32 + j;
                    const float h0 = softplus2(__uint_as_float(vv[j]) + b1[jj]);
                    const float h1 = softplus2(__uint_as_float(vv[j + 1]) + b1[jj + 1]);
                    if (is_sig) { sg0 = fmaf(w2s[jj], h0, sg0); sg1 = fmaf(w2s[jj + 1], h1, sg1); }
                    const __half2 hh = __floats2half2_rn(h0, h1);
                    const float2 back = __half22float2(hh);
                    const __half2 ll = __floats2half2_rn(h0 - back.x, h1 - back.y);
                    ph[half * 16 + j / 2] = *reinterpret_cast<const uint32_t*>(&hh);
                    pl[half * 16 + j / 2] = *reinterpret_cast<const uint32_t*>(&ll);
                }
            }
            tc::tmem_st_32x32(tmem_row + net * 64, ph);
            tc::tmem_st_32x32(tmem_row + net * 64 + 32, pl);
            tc::tmem_st_wait();
            group_sync();
            if (m == 0) {
                // layer 2: D2 = A2hi*W2hi + A2lo*W2hi + A2hi*W2lo, A2 packed fp16 in TMEM cols [net*64, net*64+64)
                const uint32_t ahi = tmem_grp + net * 64, alo = ahi + 32, d2 = tmem_grp + 128;
                const uint64_t bh = tc::umma_desc_k128(tc::smem_u32(smem + kTcW2H + net * 8192));
                const uint64_t bl = tc::umma_desc_k128(tc::smem_u32(smem + kTcW2L + net * 8192));
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16_ts(d2, ahi + 8 * k, bh + 2 * k, idesc_n32, k != 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16_ts(d2, alo + 8 * k, bh + 2 * k, idesc_n32, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16_ts(d2, ahi + 8 * k, bl + 2 * k, idesc_n32, 1);
                tc::umma_commit(&mma_bar[grp]);
            }
            wait_mma();
            uint32_t cv[32];
            tc::tmem_ld_32x32(tmem_row + 128, cv);
            tc::tmem_ld_wait();
            const uint32_t smask = net == 0 ? P.mask[0] : P.mask[1];
            // stage this row's 32 colours in its feature row (16-byte chunks XOR-swizzled by the row), then let the warp
            // write its 32 rows as 128-byte lines: 4 rows per instruction, lane = 4 consecutive channels
            const uint32_t rb = tile32 + (uint32_t)m * 128, swx = (uint32_t)(m & 7) << 4;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float c[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x = __uint_as_float(cv[i + e]) + b2c[net * 32 + i + e];
                    c[e] = ((smask >> (i + e)) & 1u) ? sigmoid_clamp_f(x) : x;
                }
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rb | (((uint32_t)i << 2) ^ swx)), "f"(c[0]),
                             "f"(c[1]), "f"(c[2]), "f"(c[3]) : "memory");
            }
            __syncwarp();
            const int sub = lane >> 3, cq = lane & 7;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int row = q * 32 + it * 4 + sub;
                const long long pr = t * 128 + row;
                const uint4 val = lds_u4((tile32 + (uint32_t)row * 128) | (((uint32_t)cq << 4) ^ ((uint32_t)(row & 7) << 4)));
                if (pr < P.total) *reinterpret_cast<uint4*>(P.out_rgb + pr * cout + net * kOut + cq * 4) = val;
            }
            __syncwarp();              // rows are rewritten by the next net / the next tile's taps
        }
        if (v) P.out_sigma[p] = sg0 + sg1;
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::tc_fence_after();
        tc::tmem_dealloc(*tmem_ptr_smem, 512);
    }
}

}  // namespace p3d

using namespace p3d;

// ImportanceRenderer.run_model on the tensor cores; same contract as p3d_run_model (render.cu) with the decoder image of
// p3d_pack_decoder_tc and optional plane strides (elements; all zero = dense [B,3,H,W,32]). out_rgb == NULL asks for the
// densities only (applications/extract_mesh.py discards the colours): layer 2, the sigmoids and the colour writes are skipped.
extern "C" int p3d_run_model_tc(const float* planes_nhwc, const int64_t plane_strides[3], const float* coords,
                                const void* decoder_tc_packed, int n_nets, int sigma_net, const uint32_t sigmoid_mask[2],
                                int B, int64_t M, int H, int W, float coord_scale, float* out_rgb, float* out_sigma,
                                p3d_stream_t stream) {
    if (!planes_nhwc || !coords || !decoder_tc_packed || !out_sigma || !sigmoid_mask) return P3D_BAD_ARG;
    if (B <= 0 || M <= 0 || H <= 0 || W <= 0 || n_nets < 1 || n_nets > 2 || sigma_net < 0 || sigma_net >= n_nets) return P3D_BAD_ARG;
    if (M > INT32_MAX) return P3D_UNSUPPORTED;
    TcQueryParams P;
    P.planes = planes_nhwc; P.coords = coords; P.decoder_packed = decoder_tc_packed;
    P.out_rgb = out_rgb; P.out_sigma = out_sigma;
    P.total = (long long)B * M;
    P.n_tiles = (P.total + 127) / 128;
    P.M = (int)M; P.H = H; P.W = W; P.n_nets = n_nets; P.sigma_net = sigma_net;
    P.mask[0] = sigmoid_mask[0]; P.mask[1] = sigmoid_mask[1];
    P.coord_scale = coord_scale;
    {
        const int64_t psz = (int64_t)H * W * kC;
        const bool dense = !plane_strides || !(plane_strides[0] || plane_strides[1] || plane_strides[2]);
        const int64_t is = dense ? 3 * psz : plane_strides[0], pls = dense ? psz : plane_strides[1], pxs = dense ? kC : plane_strides[2];
        if (is <= 0 || pls <= 0 || pxs < kC) return P3D_BAD_ARG;
        if ((((uintptr_t)planes_nhwc) & 15) != 0 || (is & 3) || (pls & 3) || (pxs & 3)) return P3D_UNSUPPORTED;
        if (out_rgb && (((uintptr_t)out_rgb) & 15) != 0) return P3D_UNSUPPORTED;
        const int64_t max_off = (int64_t)(B - 1) * is + 2 * pls + ((int64_t)H * W - 1) * pxs + kC;
        if (max_off >= ((int64_t)1 << 32)) return P3D_UNSUPPORTED;
        P.img_stride = (uint32_t)is; P.plane_stride = (uint32_t)pls; P.pix_stride = (uint32_t)pxs;
    }
    const size_t smem = (size_t)kQSmemBytes + 1024;
    P3D_CUDA_TRY(cudaFuncSetAttribute(run_model_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (P.n_tiles + 2) / 3;
    if (grid > sm_count()) grid = sm_count();
    run_model_tc_kernel<<<(int)grid, kQThreads, smem, (cudaStream_t)stream>>>(P);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}
