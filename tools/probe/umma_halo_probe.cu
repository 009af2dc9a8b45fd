// Probe: can a K-major SWIZZLE_128B A operand be read from an arbitrary 128-byte row offset inside a larger TMA-written
// patch, with a stride between 8-row groups that is NOT a multiple of 1024 bytes? (needed to reuse one haloed pixel patch
// for all 9 taps of a 3x3 convolution instead of re-loading the im2col box per tap)
//   patch: 256 rows x 64 fp16 (128 B per row), written by one TMA box with SWIZZLE_128B at a 1024-aligned address
//   A operand of an M=128, K=64 MMA: row m = patch row  s + (m / 8) * G + (m % 8)   (G rows between 8-row groups)
// For each (s, G) and each candidate value of the descriptor's base_offset field it prints the number of mismatching
// outputs against the host result.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../../pix2pix3d_b200/csrc/tc05.cuh"
#include "../../pix2pix3d_b200/csrc/tmap.cuh"
using namespace p3d;

constexpr int kRows = 256, kK = 64, kN = 64;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                       int s, int G, int base_off, float* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem;                       // 256 x 128 B = 32 KB
    uint8_t* sW = smem + kRows * 128;         // 64 x 128 B = 8 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(sW + kN * 128);
    uint64_t* mma_bar = bar + 1;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(mma_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc::mbar_init(bar, 1); tc::mbar_init(mma_bar, 1); tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tptr, 64);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(bar, kRows * 128 + kN * 128);
        tc::tma_load_4d(sX, &tmX, bar, 0, 0, 0, 0);
        tc::tma_load_4d(sW, &tmW, bar, 0, 0, 0, 0);
        tc::mbar_wait(bar, 0);
        tc::tc_fence_after();
        uint64_t da = 0;
        const uint32_t a_addr = tc::smem_u32(sX) + (uint32_t)s * 128u;
        da |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
        da |= (uint64_t)1 << 16;
        da |= (uint64_t)((G * 128) >> 4) << 32;          // stride between 8-row groups
        da |= (uint64_t)1 << 46;
        da |= (uint64_t)(base_off & 7) << 49;             // matrix base offset
        da |= (uint64_t)2 << 61;
        const uint64_t db = tc::umma_desc_k128(tc::smem_u32(sW));
        const uint32_t idesc = tc::umma_idesc_f16(128, kN, 0);
        for (int j = 0; j < 4; ++j) tc::umma_f16(tmem, da + (uint64_t)(j * 2), db + (uint64_t)(j * 2), idesc, j != 0);
        tc::umma_commit(mma_bar);
    }
    tc::mbar_wait(mma_bar, 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < kN; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tc::tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(size_t)(warp * 32 + lane) * kN + c0 + i] = __uint_as_float(v[i]);
    }
    tc::tc_fence_before(); __syncthreads();
    if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 64); }
}

int main() {
    std::vector<__half> hx((size_t)kRows * kK), hw((size_t)kN * kK);
    std::vector<float> fx(hx.size()), fw(hw.size());
    srand(1);
    for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 17 - 8); hx[i] = __float2half(fx[i]); }
    for (size_t i = 0; i < hw.size(); ++i) { fw[i] = (float)(rand() % 9 - 4); hw[i] = __float2half(fw[i]); }
    __half *dx, *dw; float* dout;
    cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dw, hw.size() * 2); cudaMalloc(&dout, 128 * kN * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tmX, tmW;
    {
        uint64_t dims[4] = {kK, kRows, 1, 1}; uint64_t str[3] = {kK * 2, (uint64_t)kRows * kK * 2, (uint64_t)kRows * kK * 2};
        uint32_t box[4] = {kK, kRows, 1, 1};
        if (make_tmap(&tmX, dx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, 4, dims, str, box)) { printf("tmap X failed\n"); return 1; }
        uint64_t dimw[4] = {kK, kN, 1, 1}; uint64_t strw[3] = {kK * 2, (uint64_t)kN * kK * 2, (uint64_t)kN * kK * 2};
        uint32_t boxw[4] = {kK, kN, 1, 1};
        if (make_tmap(&tmW, dw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, 4, dimw, strw, boxw)) { printf("tmap W failed\n"); return 1; }
    }
    const size_t smem = kRows * 128 + kN * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> got(128 * kN);
    const int cases[][2] = {{0, 8}, {8, 8}, {1, 8}, {3, 8}, {0, 10}, {1, 10}, {11, 10}, {21, 10}, {5, 18}, {0, 16}, {2, 12}};
    for (auto& cs : cases) {
        const int s = cs[0], G = cs[1];
        if (s + 15 * G + 8 > kRows) continue;
        std::vector<float> ref(128 * kN);
        for (int m = 0; m < 128; ++m) {
            const int row = s + (m / 8) * G + (m % 8);
            for (int n = 0; n < kN; ++n) { float acc = 0; for (int k = 0; k < kK; ++k) acc += fx[(size_t)row * kK + k] * fw[(size_t)n * kK + k]; ref[m * kN + n] = acc; }
        }
        printf("s=%2d G=%2d :", s, G);
        for (int bo = 0; bo < 8; ++bo) {
            cudaMemset(dout, 0, 128 * kN * 4);
            probe_kernel<<<1, 128, smem>>>(tmX, tmW, s, G, bo, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf(" bo%d:ERR(%s)", bo, cudaGetErrorString(e)); return 0; }
            cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0; for (size_t i = 0; i < got.size(); ++i) bad += got[i] != ref[i];
            printf(" bo%d:%d", bo, bad);
        }
        printf("\n");
    }
    return 0;
}
