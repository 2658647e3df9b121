// Cluster-resident K-step APPNP propagation (sm_100a): the whole problem lives in the shared memory of ONE
// thread-block cluster for all K steps.
//
// The K PPRIteration layers (gnntf/core/gnn/architectures/filter.py:34-35, driven by Layered.__call__,
// gnntf/core/nn/layered.py:52-55) on a Cora / PubMed-sized graph move a few hundred KB per step: the step is
// bound by launch latency and by the L2 round trips of a dependent chain (row_ptr -> (col,val) -> gather),
// not by bandwidth.  Here every CTA of a cluster of C CTAs owns n/C consecutive rows and keeps, in its own
// shared memory and for the whole launch,
//   * its rows of the two ping-pong feature matrices and of the teleport matrix H0,
//   * its slice of the CSR ((col,val) pairs with the column pre-split into owner CTA and local row),
// so a step is: gather neighbour rows through distributed shared memory (mapa + ld.shared::cluster),
// accumulate in CSR slot order with the same separately rounded multiply and add as spmm_rows4_kernel
// (bit-identical results), apply the teleport epilogue, store into the local slice of the other buffer, and
// meet the other CTAs at a hardware cluster barrier (barrier.cluster.arrive.release / wait.acquire).  Global
// memory is read once at the start and written once by the last step.
#include <cooperative_groups.h>

#include "spmm.cuh"

namespace gnntf {

namespace {

constexpr int kOwnerShift = 24;  // packed column: owner CTA << 24 | local row
constexpr size_t kMaxDynSmem = 227 * 1024;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
    return out;
}
__device__ __forceinline__ float4 ld_cluster4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float mac2(float v, float x, float acc) { return __fadd_rn(__fmul_rn(v, x), acc); }

struct ClusterArgs {
    const int* row_ptr;
    const int* col_idx;
    const float* val;
    const float* H0;
    float* H_out;
    int64_t ld;      // pitch of H0 / H_out in floats
    int n_rows;
    int rows_per_cta;
    int K;
    int F;
    int cap;         // (col,val) pairs of this CTA's slice that fit its shared memory; the rest is read from global
    int h0_resident; // teleport rows in shared memory (else re-read from global each step)
    float s, t;
};

// GROUP lanes own one row (float4 slot each); P = 4*GROUP floats is the row pitch in shared memory.
// THREADS = 1024: 64 registers per thread, 4 gathers in flight per lane; 512: 8 in flight.
template <int GROUP, int kClusterThreads>
__global__ void __launch_bounds__(kClusterThreads, 1) appnp_cluster_kernel(ClusterArgs a) {
    constexpr int UNROLL = kClusterThreads >= 1024 ? 4 : 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int P = 4 * GROUP;
    constexpr int NG = 32 / GROUP;
    const uint32_t rank = cooperative_groups::this_cluster().block_rank();
    const int R = a.rows_per_cta;
    const int row0 = (int)rank * R;
    const int nrows = max(0, min(R, a.n_rows - row0));
    float* const Hs = reinterpret_cast<float*>(smem_raw);  // two ping-pong matrices of R rows, back to back
    const size_t mat = (size_t)R * P;
    float* H0s = Hs + 2 * mat;
    int* rp = reinterpret_cast<int*>(H0s + (a.h0_resident ? (size_t)R * P : 0));
    int2* cv = reinterpret_cast<int2*>(rp + ((R + 1 + 3) & ~3));

    const int tid = threadIdx.x;
    // ---- load phase: row_ptr slice, CSR slice (columns split into owner / local row), H0 rows ----
    for (int i = tid; i <= nrows; i += kClusterThreads) rp[i] = __ldg(a.row_ptr + row0 + i);
    const int e0 = nrows > 0 ? __ldg(a.row_ptr + row0) : 0;
    const int entries = nrows > 0 ? __ldg(a.row_ptr + row0 + nrows) - e0 : 0;
    const int resident = min(entries, a.cap);
    for (int j = tid; j < resident; j += kClusterThreads) {
        const int c = ld_stream(a.col_idx + e0 + j);
        const int owner = c / R;
        cv[j] = make_int2((owner << kOwnerShift) | (c - owner * R), __float_as_int(ld_stream(a.val + e0 + j)));
    }
    for (int i = tid; i < nrows * GROUP; i += kClusterThreads) {
        const int r = i / GROUP, gl = i % GROUP;
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gl * 4 < a.F) h = ld_stream4(a.H0 + (int64_t)(row0 + r) * a.ld + gl * 4);
        *reinterpret_cast<float4*>(Hs + (size_t)r * P + gl * 4) = h;
        if (a.h0_resident) *reinterpret_cast<float4*>(H0s + (size_t)r * P + gl * 4) = h;
    }
    cluster_barrier();

    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane / GROUP, gl = lane % GROUP;
    const bool live = gl * 4 < a.F;
    for (int k = 0; k < a.K; ++k) {
        const float* src = Hs + (k & 1) * mat;
        float* dst = Hs + ((k & 1) ^ 1) * mat;
        const uint32_t src_lane = smem_addr(src) + gl * 16;
        const bool last = (k + 1 == a.K);
        for (int r = warp * NG + g; r < nrows; r += (kClusterThreads / 32) * NG) {
            const int begin = rp[r] - e0, end = rp[r + 1] - e0;
            float4 h;
            if (a.h0_resident) h = *reinterpret_cast<const float4*>(H0s + (size_t)r * P + gl * 4);
            else h = live ? ld_stream4(a.H0 + (int64_t)(row0 + r) * a.ld + gl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = begin; e < end; e += UNROLL) {
                float4 x[UNROLL];
                float v[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    v[u] = 0.f;
                    x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e + u < end) {
                        int2 p;
                        if (e + u < resident) {
                            p = cv[e + u];
                        } else {  // slice larger than the shared-memory budget: the tail stays in global memory
                            const int c = __ldg(a.col_idx + e0 + e + u);
                            const int owner = c / R;
                            p = make_int2((owner << kOwnerShift) | (c - owner * R), __float_as_int(__ldg(a.val + e0 + e + u)));
                        }
                        v[u] = __int_as_float(p.y);
                        const uint32_t local = (uint32_t)(p.x & ((1 << kOwnerShift) - 1));
                        x[u] = ld_cluster4(map_to_rank(src_lane + local * (P * 4), (uint32_t)(p.x >> kOwnerShift)));
                    }
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if (e + u < end) {
                        acc.x = mac2(v[u], x[u].x, acc.x); acc.y = mac2(v[u], x[u].y, acc.y);
                        acc.z = mac2(v[u], x[u].z, acc.z); acc.w = mac2(v[u], x[u].w, acc.w);
                    }
                }
            }
            // teleport epilogue, same roundings as apply_epilogue_h (spmm.cu): (acc*s) + (h*t)
            float4 o;
            o.x = __fadd_rn(__fmul_rn(acc.x, a.s), __fmul_rn(h.x, a.t));
            o.y = __fadd_rn(__fmul_rn(acc.y, a.s), __fmul_rn(h.y, a.t));
            o.z = __fadd_rn(__fmul_rn(acc.z, a.s), __fmul_rn(h.z, a.t));
            o.w = __fadd_rn(__fmul_rn(acc.w, a.s), __fmul_rn(h.w, a.t));
            if (last) {
                if (live) st_stream4(a.H_out + (int64_t)(row0 + r) * a.ld + gl * 4, o);
            } else {
                *reinterpret_cast<float4*>(dst + (size_t)r * P + gl * 4) = o;
            }
        }
        // every CTA is done reading `src` and writing `dst`; also keeps the shared memory of a CTA that
        // finished early alive until its peers' last gathers have landed
        __syncwarp();
        cluster_barrier();
    }
}

struct Layout {
    int rows_per_cta;
    int cap;
    int h0_resident;
    size_t smem;
};

// Shared-memory layout for a cluster of C CTAs, or smem == 0 when the features do not fit.
static Layout plan_layout(int64_t n, int64_t nnz, int P, int C) {
    Layout L{};
    const int64_t R = ceil_div(n, C);
    if (R >= (1 << kOwnerShift)) return L;
    const size_t mat = (size_t)R * P * sizeof(float);
    const size_t rp = (size_t)((R + 1 + 3) & ~3) * sizeof(int);
    const size_t avg_cv = (size_t)ceil_div(nnz, C) * sizeof(int2);
    size_t fixed = 3 * mat + rp;
    L.h0_resident = 1;
    if (fixed + avg_cv > kMaxDynSmem) {  // the teleport rows go first: they are streamed, not gathered
        fixed = 2 * mat + rp;
        L.h0_resident = 0;
    }
    if (fixed + avg_cv > kMaxDynSmem) return L;
    // CTAs own equal ROW counts, so their entry counts differ: give the slice all the room there is (a slice
    // that still does not fit reads its tail from global memory)
    const size_t room = (kMaxDynSmem - fixed) / sizeof(int2);
    L.cap = (int)std::min<size_t>(room, (size_t)nnz);
    L.rows_per_cta = (int)R;
    L.smem = fixed + (size_t)L.cap * sizeof(int2);
    return L;
}

template <int GROUP, int THREADS>
static int launch_cluster_t(const ClusterArgs& args, int C, size_t smem, cudaStream_t st, bool* taken) {
    auto kern = appnp_cluster_kernel<GROUP, THREADS>;
    static bool configured = false;  // per instantiation: opt in to 227 KB of dynamic shared memory and 16-CTA clusters
    if (!configured) {
        GNNTF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
        GNNTF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)C);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess || clusters < 1) {
        (void)cudaGetLastError();  // this cluster shape cannot be scheduled here: the caller takes another path
        return GNNTF_OK;
    }
    GNNTF_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args));
    *taken = true;
    return GNNTF_OK;
}

template <int GROUP>
static int launch_cluster(const ClusterArgs& args, int C, int threads, size_t smem, cudaStream_t st, bool* taken) {
    if (threads == 1024) return launch_cluster_t<GROUP, 1024>(args, C, smem, st, taken);
    return launch_cluster_t<GROUP, 512>(args, C, smem, st, taken);
}

}  // namespace

// cluster_size: 0 = choose, else 1/2/4/8/16.  *taken false: shape not eligible (the caller falls back).
int appnp_cluster_propagate(const gnntf_csr_t* A, const float* H0, float* H_out, int64_t ld, int64_t F, double alpha,
                            int K, int cluster_size, int threads, cudaStream_t st, bool* taken) {
    *taken = false;
    if (K < 1 || A->n_long > 0 || A->row_map != nullptr || A->n_rows <= 0 || F <= 0 || F > 128) return GNNTF_OK;
    if (F % 4 != 0 || ld % 4 != 0) return GNNTF_OK;
    if ((reinterpret_cast<uintptr_t>(H0) & 15u) || (reinterpret_cast<uintptr_t>(H_out) & 15u)) return GNNTF_OK;
    int group = 1;
    while (group * 4 < F) group *= 2;
    const int P = 4 * group;
    Layout L{};
    int C = cluster_size;
    if (C == 0) {
        // Measured (profiles/r2/17): gathers of 16-32 bytes through DSMEM run at only ~5 bytes per clock per SM, so
        // the only shapes where the resident form is competitive are Cora-sized ones spread over 16 CTAs
        // (K=10, F=8: 43 us); PubMed-sized ones are 3x slower than through L2.  Everything else declines.
        if (A->n_rows > 4096 || group > 2) return GNNTF_OK;
        L = plan_layout(A->n_rows, A->nnz, P, 16);
        if (L.smem == 0 || !L.h0_resident || (int64_t)L.cap * 16 < A->nnz + A->nnz / 4) return GNNTF_OK;
        C = 16;
    } else {
        if (C != 1 && C != 2 && C != 4 && C != 8 && C != 16) return GNNTF_E_SIZE;
        L = plan_layout(A->n_rows, A->nnz, P, C);
        if (L.smem == 0) return GNNTF_OK;
    }
    if (threads == 0) threads = 512;
    if (threads != 512 && threads != 1024) return GNNTF_E_SIZE;
    ClusterArgs args{A->row_ptr, A->col_idx, A->val, H0, H_out, ld, (int)A->n_rows, L.rows_per_cta, K, (int)F,
                     L.cap, L.h0_resident, (float)(1.0 - alpha), (float)alpha};
    switch (group) {
        case 1: return launch_cluster<1>(args, C, threads, L.smem, st, taken);
        case 2: return launch_cluster<2>(args, C, threads, L.smem, st, taken);
        case 4: return launch_cluster<4>(args, C, threads, L.smem, st, taken);
        case 8: return launch_cluster<8>(args, C, threads, L.smem, st, taken);
        case 16: return launch_cluster<16>(args, C, threads, L.smem, st, taken);
        default: return launch_cluster<32>(args, C, threads, L.smem, st, taken);
    }
}

}  // namespace gnntf

// Explicit entry: the K steps in one cluster launch, or GNNTF_E_SHAPE when the shape is not eligible.
extern "C" int gnntf_appnp_propagate_cluster_f32(const gnntf_csr_t* A, const float* H0, float* H_out, int64_t ld,
                                                 int64_t F, double alpha, int K, int cluster_size, int threads,
                                                 void* stream) {
    using namespace gnntf;
    int rc = validate_csr(A);
    if (rc != GNNTF_OK) return rc;
    if (K < 1 || F < 0 || ld < F) return GNNTF_E_SIZE;
    if (A->n_rows > 0 && F > 0 && (H0 == nullptr || H_out == nullptr)) return GNNTF_E_NULL;
    bool taken = false;
    rc = appnp_cluster_propagate(A, H0, H_out, ld, F, alpha, K, cluster_size, threads, (cudaStream_t)stream, &taken);
    if (rc != GNNTF_OK) return rc;
    return taken ? GNNTF_OK : GNNTF_E_SHAPE;
}
