// Fused pack + send of halo rows over NVLink peer memory, with a completion signal (sm_100a).
// Shared by the stand-alone push kernel (dist.cu) and the SpMM row kernel, whose LEADING CTAs run the
// push of the step's input rows while the rest of the grid computes (spmm.cu: one launch, the push is
// dispatched first by construction — no second stream, no stream priorities).
#pragma once
#include "common.cuh"

namespace gnntf {

__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Completion signal of a push (all arguments NULL/0 = no signalling).  Every CTA fences its remote
// stores at system scope and counts itself in; the LAST one publishes `*epoch_base + epoch_delta` into
// slot `my_slot` of every destination's flag array with a release store over NVLink, so a consumer
// that acquires the flag also sees every row this rank pushed (and the counter is left at zero for
// the next launch).  This replaces the host-issued NCCL all-reduce round 1 used as the barrier.
struct PushSignal {
    int32_t* done_counter;              // local, zero between launches
    int32_t* const* peer_flags;         // [n_peers] flag arrays of the peers (peer memory), NULL entries skipped
    const int32_t* epoch_base;          // device scalar advanced by the host side once per propagation
    int32_t epoch_delta;
    int32_t my_slot;
};

__device__ __forceinline__ void push_signal_tail(const PushSignal& sg, int n_peers, int n_ctas) {
    if (sg.done_counter == nullptr) return;
    __threadfence_system();             // this thread's remote stores are performed before its CTA is counted
    __syncthreads();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(sg.done_counter, 1);
        if (prev == n_ctas - 1) {
            __threadfence_system();     // order the other CTAs' (fenced, counted) stores before the flags
            const int32_t epoch = *sg.epoch_base + sg.epoch_delta;
            for (int d = 0; d < n_peers; ++d)
                if (sg.peer_flags[d] != nullptr) st_release_sys(sg.peer_flags[d] + sg.my_slot, epoch);
            *sg.done_counter = 0;
        }
    }
}

// CTA `cta` of `n_ctas` (either a whole grid, or the leading CTAs of the SpMM grid).
template <int VEC, int G, int U>
__device__ __forceinline__ void halo_push_body(const float* __restrict__ H, int64_t ld, const int32_t* __restrict__ send_idx,
                                               const int64_t* __restrict__ send_off, float* const* __restrict__ peer_base,
                                               const int64_t* __restrict__ peer_row0, int n_peers, int64_t n_send,
                                               int64_t rotate, int64_t ldo, int F, int cta, int n_ctas) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const int64_t group = (((int64_t)cta * blockDim.x + threadIdx.x) >> 5) * RPW + lane / G;
    const int64_t n_groups = (((int64_t)n_ctas * blockDim.x) >> 5) * RPW;
    for (int64_t k0 = group; k0 < n_send; k0 += n_groups * U) {
        const float* src[U];
        float* dst[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t k = k0 + (int64_t)u * n_groups;
            ok[u] = k < n_send;
            // every rank starts with a different destination, so at any moment each receiver is
            // the target of (about) one sender instead of all of them
            int64_t i = ok[u] ? k + rotate : 0;
            if (i >= n_send) i -= n_send;
            int d = 0;
            while (d + 1 < n_peers && i >= send_off[d + 1]) ++d;   // n_peers <= 8: linear scan
            src[u] = H + (int64_t)__ldg(send_idx + i) * ld;
            dst[u] = ok[u] ? peer_base[d] + (peer_row0[d] + (i - send_off[d])) * ldo : nullptr;
        }
        if (F <= G * VEC) {  // one slot per lane: all U loads first, then all U stores
            Vec<VEC> x[U];
            const bool mine = gl * VEC < F;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u] && mine) x[u] = Vec<VEC>::gather(src[u] + gl * VEC);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u] && mine) x[u].store(dst[u] + gl * VEC);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u])
                    for (int f = gl * VEC; f < F; f += G * VEC) Vec<VEC>::gather(src[u] + f).store(dst[u] + f);
        }
    }
}


// Everything the leading CTAs of an SpMM launch need to push the launch's dense operand (n_ctas == 0: no push).
struct PushArgs {
    int n_ctas;
    const int32_t* send_idx;
    const int64_t* send_off;
    float* const* peer_base;
    const int64_t* peer_row0;
    int n_peers;
    int64_t n_send;
    int64_t rotate;
    int64_t ldo;
    PushSignal sg;
};

}  // namespace gnntf
