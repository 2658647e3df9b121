// Helpers for the row-sharded multi-GPU path (contiguous node-range split, one process per GPU).
// The halo exchange itself is an NCCL all-to-all issued by the host shim; this kernel packs the
// rows each peer needs into the send buffer.  The propagation step of the reference
// (gnntf/core/gnn/architectures/filter.py:19-22) is otherwise unchanged: the local CSR simply
// addresses [owned rows | halo rows].
#include <algorithm>

#include "common.cuh"

namespace gnntf {

// out[i, 0:F] = H[send_idx[i], 0:F]; one warp per row, float4 when the layout allows.
template <int VEC>
__global__ void halo_pack_kernel(const float* __restrict__ H, int64_t ld, const int32_t* __restrict__ send_idx,
                                 int64_t n_send, float* __restrict__ out, int64_t ldo, int F) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n_send; i += nwarps) {
        const float* src = H + (int64_t)__ldg(send_idx + i) * ld;
        float* dst = out + i * ldo;
        for (int f = lane * VEC; f < F; f += 32 * VEC) Vec<VEC>::gather(src + f).store(dst + f);
    }
}

}  // namespace gnntf

using namespace gnntf;

extern "C" int gnntf_halo_pack_f32(const float* H, int64_t ld, const int32_t* send_idx, int64_t n_send,
                                   float* out, int64_t ldo, int64_t F, void* stream) {
    if (n_send < 0 || F < 0 || F > 0x7fffffff || ld < F || ldo < F) return GNNTF_E_SIZE;
    if (n_send == 0 || F == 0) return GNNTF_OK;
    if (H == nullptr || send_idx == nullptr || out == nullptr) return GNNTF_E_NULL;
    const int grid = (int)std::min<int64_t>(ceil_div(n_send, 8), (int64_t)kNumSMs * 16);
    const bool v4 = (F % 4 == 0) && (ld % 4 == 0) && (ldo % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    if (v4)
        halo_pack_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(H, ld, send_idx, n_send, out, ldo, (int)F);
    else
        halo_pack_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(H, ld, send_idx, n_send, out, ldo, (int)F);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}
