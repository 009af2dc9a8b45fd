// Broadcasting fused multiply-add y = a * b + c -- torch_utils/ops/fma.py:17-34 of the reference (forward = torch.addcmul(c, a, b);
// its gradients are the same op with other operands plus reductions over the broadcast dimensions). The op sits on the
// non-fused modulated convolution of the training passes (networks_stylegan2.py:79-82: x * dcoefs + noise with dcoefs [N,C,1,1]
// and noise [N,1,H,W]): HBM streaming, numel_out * sizeof * (1 read + 1 write) + the broadcast operands.
#include "p3d_common.cuh"

namespace p3d {

struct FmaParams {
    const void* a; const void* b; const void* c; void* y;
    int64_t shape[4];
    int64_t sa[4], sb[4], sc[4];      // element strides, 0 on broadcast dimensions; y is dense
    int64_t total;                    // elements of y (scalar kernel) or groups of VEC along the last dimension (vector kernel)
};

template <class T> struct FmaAcc { using type = float; };
template <> struct FmaAcc<double> { using type = double; };

template <class T>
__device__ __forceinline__ T fma_one(T a, T b, T c) {
    using A = typename FmaAcc<T>::type;
    return (T)fma((A)a, (A)b, (A)c);
}

template <class T>
__global__ void __launch_bounds__(256) fma_scalar_kernel(const FmaParams p) {
    const T* a = (const T*)p.a; const T* b = (const T*)p.b; const T* c = (const T*)p.c; T* y = (T*)p.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int64_t i3 = r % p.shape[3]; r /= p.shape[3];
        const int64_t i2 = r % p.shape[2]; r /= p.shape[2];
        const int64_t i1 = r % p.shape[1];
        const int64_t i0 = r / p.shape[1];
        y[i] = fma_one<T>(a[i0 * p.sa[0] + i1 * p.sa[1] + i2 * p.sa[2] + i3 * p.sa[3]], b[i0 * p.sb[0] + i1 * p.sb[1] + i2 * p.sb[2] + i3 * p.sb[3]],
                          c[i0 * p.sc[0] + i1 * p.sc[1] + i2 * p.sc[2] + i3 * p.sc[3]]);
    }
}

// 16 bytes of y per thread and step; every operand is either dense along the last dimension (stride 1, 16-byte aligned rows) or
// broadcast along it (stride 0)
template <class T, int VEC>
__global__ void __launch_bounds__(256) fma_vec_kernel(const FmaParams p) {
    const T* a = (const T*)p.a; const T* b = (const T*)p.b; const T* c = (const T*)p.c; T* y = (T*)p.y;
    const int64_t groups = p.shape[3] / VEC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int64_t g3 = r % groups; r /= groups;
        const int64_t i2 = r % p.shape[2]; r /= p.shape[2];
        const int64_t i1 = r % p.shape[1];
        const int64_t i0 = r / p.shape[1];
        const int64_t i3 = g3 * VEC;
        struct alignas(16) Pack { T v[VEC]; };
        Pack va, vb, vc, vy;
        const T* pa = a + i0 * p.sa[0] + i1 * p.sa[1] + i2 * p.sa[2] + i3 * p.sa[3];
        const T* pb = b + i0 * p.sb[0] + i1 * p.sb[1] + i2 * p.sb[2] + i3 * p.sb[3];
        const T* pc = c + i0 * p.sc[0] + i1 * p.sc[1] + i2 * p.sc[2] + i3 * p.sc[3];
        if (p.sa[3]) va = *reinterpret_cast<const Pack*>(pa);
        else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) va.v[k] = pa[0];
        }
        if (p.sb[3]) vb = *reinterpret_cast<const Pack*>(pb);
        else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) vb.v[k] = pb[0];
        }
        if (p.sc[3]) vc = *reinterpret_cast<const Pack*>(pc);
        else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) vc.v[k] = pc[0];
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) vy.v[k] = fma_one<T>(va.v[k], vb.v[k], vc.v[k]);
        *reinterpret_cast<Pack*>(y + i * VEC) = vy;
    }
}

template <class T, int VEC>
static int fma_launch(FmaParams p, bool vec, cudaStream_t stream) {
    int64_t n = p.shape[0] * p.shape[1] * p.shape[2] * p.shape[3];
    if (vec) n /= VEC;
    p.total = n;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return P3D_OK;
    if (vec) fma_vec_kernel<T, VEC><<<(unsigned)blocks, 256, 0, stream>>>(p);
    else fma_scalar_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(p);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

}  // namespace p3d

using namespace p3d;

extern "C" int p3d_fma(const void* a, const void* b, const void* c, void* y, int dtype, const int64_t shape[4], const int64_t a_stride[4],
                       const int64_t b_stride[4], const int64_t c_stride[4], p3d_stream_t stream) {
    if (!a || !b || !c || !y || !shape || !a_stride || !b_stride || !c_stride) return P3D_BAD_ARG;
    FmaParams p;
    p.a = a; p.b = b; p.c = c; p.y = y;
    for (int i = 0; i < 4; ++i) {
        if (shape[i] < 0 || a_stride[i] < 0 || b_stride[i] < 0 || c_stride[i] < 0) return P3D_BAD_ARG;
        p.shape[i] = shape[i]; p.sa[i] = a_stride[i]; p.sb[i] = b_stride[i]; p.sc[i] = c_stride[i];
    }
    if (shape[0] * shape[1] * shape[2] * shape[3] == 0) return P3D_OK;
    const int es = dtype == P3D_F16 ? 2 : dtype == P3D_F32 ? 4 : dtype == P3D_F64 ? 8 : 0;
    if (!es) return P3D_BAD_ARG;
    const int vec = 16 / es;
    // vector path: last dimension a multiple of the vector, every operand dense (stride 1) or broadcast (stride 0) along it, and
    // all row starts 16-byte aligned (base pointers and the outer strides of the dense operands)
    bool ok = shape[3] % vec == 0 && (((uintptr_t)y) & 15) == 0;
    const void* ptrs[3] = {a, b, c};
    const int64_t* strides[3] = {a_stride, b_stride, c_stride};
    for (int o = 0; o < 3 && ok; ++o) {
        const int64_t* s = strides[o];
        if (s[3] == 0) continue;
        if (s[3] != 1 || (((uintptr_t)ptrs[o]) & 15) != 0) { ok = false; break; }
        for (int i = 0; i < 3; ++i)
            if ((s[i] * es) % 16 != 0) ok = false;
    }
    if (dtype == P3D_F16) return fma_launch<__half, 8>(p, ok, (cudaStream_t)stream);
    if (dtype == P3D_F32) return fma_launch<float, 4>(p, ok, (cudaStream_t)stream);
    return fma_launch<double, 2>(p, ok, (cudaStream_t)stream);
}
