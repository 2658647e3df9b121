// Helpers for the row-sharded multi-GPU path (contiguous node-range split, one process per GPU).
// The halo exchange itself is an NCCL all-to-all issued by the host shim; this kernel packs the
// rows each peer needs into the send buffer.  The propagation step of the reference
// (gnntf/core/gnn/architectures/filter.py:19-22) is otherwise unchanged: the local CSR simply
// addresses [owned rows | halo rows].
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "push.cuh"
#include "spmm.cuh"

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

// Fused pack + send: the destination is a peer GPU's halo buffer (NVLink stores).  G lanes own a
// row (G sized to the row width), a warp runs 32/G rows side by side and every group keeps U rows
// in flight (index loads, then all row loads, then all remote stores): the first version — one warp
// per row, one row at a time — was latency-bound at 233 GB/s (profiles/r1, N=8 phase timing).
template <int VEC, int G, int U>
__global__ void halo_push_kernel(const float* __restrict__ H, int64_t ld, const int32_t* __restrict__ send_idx,
                                 const int64_t* __restrict__ send_off, float* const* __restrict__ peer_base,
                                 const int64_t* __restrict__ peer_row0, int n_peers, int64_t n_send,
                                 int64_t rotate, int64_t ldo, int F, PushSignal sg) {
    halo_push_body<VEC, G, U>(H, ld, send_idx, send_off, peer_base, peer_row0, n_peers, n_send, rotate, ldo, F,
                              (int)blockIdx.x, (int)gridDim.x);
    push_signal_tail(sg, n_peers, (int)gridDim.x);
}

// Spin (one warp) until flags[i] >= *epoch_base + epoch_delta for every i != skip.  Enqueued on the
// consumer's stream in front of the kernel that reads the pushed rows: the acquire pairs with the
// producers' release stores; the kernel boundary orders it before the consumer's loads.  A peer that
// never arrives traps after ~4 s instead of hanging the GPU.
__global__ void wait_flags_kernel(const int32_t* __restrict__ flags, int n, int skip, const int32_t* __restrict__ epoch_base,
                                  int32_t epoch_delta) {
    const int32_t epoch = *epoch_base + epoch_delta;
    const long long t0 = clock64();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (i == skip) continue;
        while (ld_acquire_sys(flags + i) - epoch < 0) {     // wrap-safe comparison
            if (clock64() - t0 > 8000000000LL) __trap();
            __nanosleep(100);
        }
    }
}

// flags_of_peer[d][my_slot] = *epoch_base + epoch_delta for every mapped peer (release): "this rank is
// done reading its halo buffers of the propagation that just ended".
__global__ void signal_flags_kernel(int32_t* const* __restrict__ peer_flags, int n_peers, int my_slot,
                                    const int32_t* __restrict__ epoch_base, int32_t epoch_delta) {
    const int32_t epoch = *epoch_base + epoch_delta;
    for (int d = threadIdx.x; d < n_peers; d += blockDim.x)
        if (peer_flags[d] != nullptr) st_release_sys(peer_flags[d] + my_slot, epoch);
}

// *flag = *value at system scope, after everything this GPU wrote before (used behind a copy-engine transfer: the
// kernel starts only when the copy has completed, the fence + release make the order visible to the peer's acquire)
__global__ void signal_value_kernel(int32_t* __restrict__ flag, const int32_t* __restrict__ value) {
    __threadfence_system();
    st_release_sys(flag, *value);
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

static int halo_push_impl(const float* H, int64_t ld, const int32_t* send_idx, const int64_t* send_off,
                          float* const* peer_base, const int64_t* peer_row0, int n_peers, int64_t n_send,
                          int64_t rotate, int64_t ldo, int64_t F, PushSignal sg, cudaStream_t st) {
    if (n_send < 0 || F < 0 || F > 0x7fffffff || ld < F || ldo < F || n_peers < 1 || n_peers > 64) return GNNTF_E_SIZE;
    if (rotate < 0 || (n_send > 0 && rotate >= n_send)) return GNNTF_E_SIZE;
    if (F == 0 || (n_send == 0 && sg.done_counter == nullptr)) return GNNTF_OK;
    if (n_send > 0 && (H == nullptr || send_idx == nullptr || send_off == nullptr || peer_base == nullptr || peer_row0 == nullptr))
        return GNNTF_E_NULL;
    // The push overlaps the owned-column SpMM pass: a small grid leaves the SMs to that kernel
    // (NVLink needs far fewer warps in flight than HBM does).
    const int max_ctas = kNumSMs;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_send, 8), (int64_t)max_ctas));  // >= 1: the signal must go out
    const bool v4 = (F % 4 == 0) && (ld % 4 == 0) && (ldo % 4 == 0) && (reinterpret_cast<uintptr_t>(H) & 15u) == 0;
#define GNNTF_PUSH(VEC, G)                                                                                     \
    halo_push_kernel<VEC, G, 4><<<grid, 256, 0, st>>>(H, ld, send_idx, send_off, peer_base, peer_row0, n_peers, \
                                                      n_send, rotate, ldo, (int)F, sg)
    if (v4) {
        if (F <= 16) GNNTF_PUSH(4, 4);
        else if (F <= 32) GNNTF_PUSH(4, 8);
        else if (F <= 64) GNNTF_PUSH(4, 16);
        else GNNTF_PUSH(4, 32);
    } else {
        if (F <= 8) GNNTF_PUSH(1, 8);
        else if (F <= 16) GNNTF_PUSH(1, 16);
        else GNNTF_PUSH(1, 32);
    }
#undef GNNTF_PUSH
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_halo_push_f32(const float* H, int64_t ld, const int32_t* send_idx, const int64_t* send_off,
                                   float* const* peer_base, const int64_t* peer_row0, int n_peers,
                                   int64_t n_send, int64_t rotate, int64_t ldo, int64_t F, void* stream) {
    return halo_push_impl(H, ld, send_idx, send_off, peer_base, peer_row0, n_peers, n_send, rotate, ldo, F, PushSignal{},
                          (cudaStream_t)stream);
}

extern "C" int gnntf_halo_push_signal_f32(const float* H, int64_t ld, const int32_t* send_idx, const int64_t* send_off,
                                          float* const* peer_base, const int64_t* peer_row0, int n_peers,
                                          int64_t n_send, int64_t rotate, int64_t ldo, int64_t F,
                                          int32_t* done_counter, int32_t* const* peer_flags, int my_slot,
                                          const int32_t* epoch_base, int32_t epoch_delta, void* stream) {
    if (done_counter == nullptr || peer_flags == nullptr || epoch_base == nullptr) return GNNTF_E_NULL;
    if (my_slot < 0) return GNNTF_E_SIZE;
    return halo_push_impl(H, ld, send_idx, send_off, peer_base, peer_row0, n_peers, n_send, rotate, ldo, F,
                          PushSignal{done_counter, peer_flags, epoch_base, epoch_delta, my_slot}, (cudaStream_t)stream);
}

// One sharded step over the OWNED columns with the halo push of its input riding in the same launch.
extern "C" int gnntf_step_push_f32(const gnntf_csr_t* A, const float* H_in, const float* H0, float* H_out, int64_t ld,
                                   int64_t F, double alpha, const int32_t* send_idx, const int64_t* send_off,
                                   float* const* peer_base, const int64_t* peer_row0, int n_peers, int64_t n_send,
                                   int64_t rotate, int32_t* done_counter, int32_t* const* peer_flags, int my_slot,
                                   const int32_t* epoch_base, int32_t epoch_delta, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (A == nullptr || done_counter == nullptr || peer_flags == nullptr || epoch_base == nullptr) return GNNTF_E_NULL;
    if (F < 0 || F > 0x7fffffff || ld < F || n_send < 0 || n_peers < 1 || n_peers > 64 || my_slot < 0) return GNNTF_E_SIZE;
    if (rotate < 0 || (n_send > 0 && rotate >= n_send)) return GNNTF_E_SIZE;
    if (A->n_rows > 0 && F > 0 && (H_in == nullptr || H_out == nullptr)) return GNNTF_E_NULL;
    if (n_send > 0 && (send_idx == nullptr || send_off == nullptr || peer_base == nullptr || peer_row0 == nullptr))
        return GNNTF_E_NULL;
    Epilogue e{};
    e.s = (H0 != nullptr) ? (float)(1.0 - alpha) : 1.0f;   // H0 == NULL: plain SpMM (the R-MAT sweep)
    e.H0 = H0;
    e.ldh = ld;
    e.t = (float)alpha;
    e.act = GNNTF_ACT_IDENTITY;
    e.C = H_out;
    e.ldc = ld;
    e.F = (int)F;
    PushArgs pa{};
    pa.n_ctas = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_send, 8), (int64_t)kNumSMs));
    pa.send_idx = send_idx;
    pa.send_off = send_off;
    pa.peer_base = peer_base;
    pa.peer_row0 = peer_row0;
    pa.n_peers = n_peers;
    pa.n_send = n_send;
    pa.rotate = rotate;
    pa.ldo = ld;
    pa.sg = PushSignal{done_counter, peer_flags, epoch_base, epoch_delta, my_slot};
    bool pushed = false;
    int rc = spmm_dispatch(A, H_in, ld, e, st, &pa, &pushed);
    if (rc != GNNTF_OK) return rc;
    if (!pushed)  // scalar layout / empty shard: the push (or at least its signal) goes out as its own launch
        rc = halo_push_impl(H_in, ld, send_idx, send_off, peer_base, peer_row0, n_peers, n_send, rotate, ld, F, pa.sg, st);
    return rc;
}

extern "C" int gnntf_flags_wait(const int32_t* flags, int n, int skip, const int32_t* epoch_base, int32_t epoch_delta,
                                void* stream) {
    if (n < 0 || n > 1024) return GNNTF_E_SIZE;
    if (n == 0) return GNNTF_OK;
    if (flags == nullptr || epoch_base == nullptr) return GNNTF_E_NULL;
    wait_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, n, skip, epoch_base, epoch_delta);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_flags_signal(int32_t* const* peer_flags, int n_peers, int my_slot, const int32_t* epoch_base,
                                  int32_t epoch_delta, void* stream) {
    if (n_peers < 0 || n_peers > 1024 || my_slot < 0) return GNNTF_E_SIZE;
    if (n_peers == 0) return GNNTF_OK;
    if (peer_flags == nullptr || epoch_base == nullptr) return GNNTF_E_NULL;
    signal_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peer_flags, n_peers, my_slot, epoch_base, epoch_delta);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

// Copy-engine exchange: this rank's block of rows goes to a peer's buffer as ONE DMA copy over NVLink — no SM, no
// L1TEX and no shared-memory traffic: the SpMM running next to it keeps the whole gather path.  The epoch flag is
// written by a one-thread kernel enqueued behind the copy (st.release.sys): a kernel behind a copy starts only when
// the copy has completed.  (The first version wrote the flag with a 4-byte DMA behind the block DMA; it passed at
// N = 2 and FAILED the parity check at N = 8 (4 x 2, three peers sending at once, errors ~1e-2): stream order between
// two copies evidently does not make the first one's bytes visible to a polling peer before the second one's.  The
// kernel form is validated at N = 2, which is where the copy mode is used by default.)
extern "C" int gnntf_peer_copy_signal(void* dst, const void* src, size_t bytes, int32_t* peer_flag,
                                      const int32_t* epoch_value, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (bytes > 0) {
        if (dst == nullptr || src == nullptr) return GNNTF_E_NULL;
        GNNTF_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
    }
    if (peer_flag != nullptr) {
        if (epoch_value == nullptr) return GNNTF_E_NULL;
        signal_value_kernel<<<1, 1, 0, st>>>(peer_flag, epoch_value);
        GNNTF_LAUNCH_CHECK();
    }
    return GNNTF_OK;
}

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");

extern "C" int gnntf_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char handle[64]) {
    if (dev_ptr == nullptr || handle == nullptr) return GNNTF_E_NULL;
    if (bytes == 0) return GNNTF_E_SIZE;
    GNNTF_CUDA_TRY(cudaMalloc(dev_ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
    if (e != cudaSuccess) {
        cudaFree(*dev_ptr);
        *dev_ptr = nullptr;
        return (int)e;
    }
    memcpy(handle, &h, 64);
    return GNNTF_OK;
}

extern "C" int gnntf_ipc_open(const unsigned char handle[64], void** dev_ptr) {
    if (dev_ptr == nullptr || handle == nullptr) return GNNTF_E_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    GNNTF_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GNNTF_OK;
}

extern "C" int gnntf_ipc_close(void* dev_ptr) {
    if (dev_ptr == nullptr) return GNNTF_OK;
    GNNTF_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return GNNTF_OK;
}

extern "C" int gnntf_ipc_free(void* dev_ptr) {
    if (dev_ptr == nullptr) return GNNTF_OK;
    GNNTF_CUDA_TRY(cudaFree(dev_ptr));
    return GNNTF_OK;
}
