// GPU CSR builder and adjacency normaliser (sm_100a).
//
// gnntf_csr_build   : gnntf/core/gnn/graph_manipulation.py:24-31 (graph2adj) — symmetrise by
//                     appending the reversed edge list with duplicated values (no dedup, no
//                     coalescing, no self loops added), then derive the stable-by-row CSR view.
// gnntf_normalize_f32: gnntf/core/nn/layered.py:47-50 (sparse_dropout with an explicit mask)
//                     + gnntf/core/gnn/gnn.py:36-50 (get_adjacency: column sums -> sqrt ->
//                     divide_no_nan -> row scale -> column scale).
//
// Pipeline (all HBM-bound integer/scan work; DESIGN.md gives the byte counts):
//   1. keys[q] = row of COO slot q, vals[q] = q               (one pass over the edge list)
//   2. stable LSD radix sort of (key, q) pairs on ceil(log2 n) key bits (cub::DeviceRadixSort —
//      stable, deterministic) -> coo_pos
//   3. row_ptr from the sorted keys (boundary fill, handles empty rows)
//   4. col_idx / raw_val gathered through coo_pos
// Because the symmetrised list holds every entry (r,c) at slot q together with its mirror (c,r) at
// slot q±E, the column sum of column j equals the sum over ROW j of the mirrors' values, so the
// degree is a deterministic segmented row reduction — no float atomics, no second sort (TF's
// SparseReduceSum re-sorts all nnz on every call).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace gnntf {

__device__ __forceinline__ void coo_entry(const int64_t* __restrict__ edges,
                                          const float* __restrict__ weights, int64_t E,
                                          int64_t n_graph, int64_t q, int64_t& r, int64_t& c, float& w) {
    if (q < n_graph) {
        if (q < E) {
            r = edges[2 * q];
            c = edges[2 * q + 1];
            w = weights ? weights[q] : 1.0f;
        } else {  // appended reversed list, graph_manipulation.py:29-30
            r = edges[2 * (q - E) + 1];
            c = edges[2 * (q - E)];
            w = weights ? weights[q - E] : 1.0f;
        }
    } else {  // tf.sparse.eye entries, gnn.py:39,49
        r = c = q - n_graph;
        w = 1.0f;
    }
}

__global__ void emit_coo_kernel(const int64_t* __restrict__ edges, const float* __restrict__ weights,
                                int64_t E, int64_t n_graph, int64_t nnz, int by_column,
                                int64_t* __restrict__ coo_indices, float* __restrict__ coo_values,
                                uint32_t* __restrict__ keys, int32_t* __restrict__ pos) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz;
         q += (int64_t)gridDim.x * blockDim.x) {
        int64_t r, c;
        float w;
        coo_entry(edges, weights, E, n_graph, q, r, c, w);
        if (coo_indices) {
            reinterpret_cast<longlong2*>(coo_indices)[q] = make_longlong2(r, c);
        }
        if (coo_values) coo_values[q] = w;
        keys[q] = (uint32_t)(by_column ? c : r);
        pos[q] = (int32_t)q;
    }
}

// row_ptr[r] = first sorted slot whose key >= r.  Slot p fills (key[p-1], key[p]]; the last slot
// also fills (key[nnz-1], n].
__global__ void row_ptr_kernel(const uint32_t* __restrict__ sorted_keys, int64_t nnz, int64_t n,
                               int32_t* __restrict__ row_ptr) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cur = sorted_keys[p];
        const int64_t prev = (p == 0) ? -1 : (int64_t)sorted_keys[p - 1];
        for (int64_t r = prev + 1; r <= cur; ++r) row_ptr[r] = (int32_t)p;
        if (p == nnz - 1)
            for (int64_t r = cur + 1; r <= n; ++r) row_ptr[r] = (int32_t)nnz;
    }
}

__global__ void gather_csr_kernel(const int64_t* __restrict__ edges, const float* __restrict__ weights,
                                  int64_t E, int64_t n_graph, int64_t nnz, int by_column,
                                  const int32_t* __restrict__ coo_pos, int32_t* __restrict__ col_idx,
                                  float* __restrict__ raw_val) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
         p += (int64_t)gridDim.x * blockDim.x) {
        int64_t r, c;
        float w;
        coo_entry(edges, weights, E, n_graph, coo_pos[p], r, c, w);
        col_idx[p] = (int32_t)(by_column ? r : c);
        raw_val[p] = w;
    }
}

static int key_bits(int64_t n) {
    int b = 1;
    while (b < 32 && (1LL << b) < n) ++b;
    return b;
}

struct BuildWs {
    size_t keys_in, keys_out, pos_in, cub, total, cub_bytes;
};

static int build_ws_layout(int64_t nnz, int64_t n, BuildWs& L) {
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs<uint32_t, int32_t>(
        nullptr, cub_bytes, nullptr, nullptr, nullptr, nullptr, (int)nnz, 0, key_bits(n));
    if (e != cudaSuccess) return (int)e;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 0;
    L.keys_in = off;  off += up((size_t)nnz * 4);
    L.keys_out = off; off += up((size_t)nnz * 4);
    L.pos_in = off;   off += up((size_t)nnz * 4);
    L.cub = off;      off += up(cub_bytes);
    L.cub_bytes = cub_bytes;
    L.total = off + 256;
    return GNNTF_OK;
}

static int coo_nnz(int64_t n, int64_t n_edges, int directed, int add_eye, int64_t& n_graph, int64_t& nnz) {
    if (n < 0 || n_edges < 0 || n > 0x7ffffffeLL) return GNNTF_E_SIZE;
    n_graph = directed ? n_edges : 2 * n_edges;
    nnz = n_graph + (add_eye ? n : 0);
    if (nnz > 0x7fffffffLL) return GNNTF_E_SIZE;
    return GNNTF_OK;
}

// ---------------------------------------------------------------------------------------------
// Normalisation
// ---------------------------------------------------------------------------------------------

// Masked raw value of COO slot q: (x*scale)*keep, layered.py:50 / tf.nn.dropout.
__device__ __forceinline__ float masked(float raw, int64_t q, int64_t n_graph,
                                        const uint8_t* __restrict__ keep, float scale) {
    if (keep == nullptr || q >= n_graph) return raw;  // eye entries are added after dropout
    return keep[q] ? raw * scale : 0.0f;
}

// Undirected: deg[j] = sum over row j of the MIRROR entries' masked values (= column sum of j).
// One warp per row, lanes stride the row, fixed shuffle tree: deterministic.
__global__ void degree_rows_kernel(const int32_t* __restrict__ row_ptr, const float* __restrict__ raw_val,
                                   const int32_t* __restrict__ coo_pos, int64_t n, int64_t n_graph,
                                   int64_t E, const uint8_t* __restrict__ keep, float scale,
                                   int eye_mode, float* __restrict__ deg) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int s = row_ptr[r], e = row_ptr[r + 1];
        float acc = 0.0f;
        for (int p = s + lane; p < e; p += 32) {
            const int64_t q = coo_pos[p];
            float m;
            if (q >= n_graph) {
                m = (eye_mode == GNNTF_EYE_AFTER) ? 0.0f : raw_val[p];  // diagonal: its own mirror
            } else {
                const int64_t mirror = (q < E) ? q + E : q - E;
                m = masked(raw_val[p], mirror, n_graph, keep, scale);  // mirror carries the same raw weight
            }
            acc += m;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) deg[r] = acc;
    }
}

// Directed: no mirror half, so scatter into the column sums.  Exact (order-independent) for the
// unit / dyadic weights graph2adj produces by default; otherwise last-bit nondeterministic.
__global__ void degree_atomic_kernel(const int32_t* __restrict__ col_idx, const float* __restrict__ raw_val,
                                     const int32_t* __restrict__ coo_pos, int64_t nnz, int64_t n_graph,
                                     const uint8_t* __restrict__ keep, float scale, int eye_mode,
                                     float* __restrict__ deg) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = coo_pos[p];
        float m = masked(raw_val[p], q, n_graph, keep, scale);
        if (q >= n_graph && eye_mode == GNNTF_EYE_AFTER) m = 0.0f;
        if (m != 0.0f) atomicAdd(deg + col_idx[p], m);
    }
}

__global__ void dinv_kernel(const float* __restrict__ deg, int64_t n, int mode, float* __restrict__ dinv) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x) {
        const float d = (mode == GNNTF_NORM_SYMMETRIC) ? __fsqrt_rn(deg[j]) : deg[j];  // gnn.py:41 / :44
        dinv[j] = (d == 0.0f) ? 0.0f : __fdiv_rn(1.0f, d);                              // divide_no_nan
    }
}

// One warp per row so that the row index is known without a search; lanes stride the row.
__global__ void scale_values_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                    const float* __restrict__ raw_val, const int32_t* __restrict__ coo_pos,
                                    int64_t n, int64_t n_graph, int64_t E, int directed,
                                    const uint8_t* __restrict__ keep, float scale, int mode, int eye_mode,
                                    const float* __restrict__ dinv, float* __restrict__ norm_val,
                                    float* __restrict__ norm_val_T, float* __restrict__ norm_val_coo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int s = row_ptr[r], e = row_ptr[r + 1];
        const float dr = (mode == GNNTF_NORM_NONE) ? 1.0f : dinv[r];
        for (int p = s + lane; p < e; p += 32) {
            const int64_t q = coo_pos[p];
            const int c = col_idx[p];
            const float raw = raw_val[p];
            float v, vT;
            if (q >= n_graph && eye_mode == GNNTF_EYE_AFTER) {
                v = vT = raw;  // identity added after normalisation, gnn.py:48-49
            } else {
                const float m = masked(raw, q, n_graph, keep, scale);
                float mT = m;
                if (!directed && q < n_graph) {
                    const int64_t mirror = (q < E) ? q + E : q - E;
                    mT = masked(raw, mirror, n_graph, keep, scale);
                }
                if (mode == GNNTF_NORM_SYMMETRIC) {
                    const float dc = dinv[c];
                    v = __fmul_rn(__fmul_rn(m, dr), dc);    // (v*D[row])*D[col], gnn.py:42
                    vT = __fmul_rn(__fmul_rn(mT, dc), dr);  // mirror entry: row c, col r
                } else if (mode == GNNTF_NORM_BIPARTITE) {
                    v = __fmul_rn(m, dr);                   // v*D[row], gnn.py:45
                    vT = __fmul_rn(mT, dinv[c]);
                } else {
                    v = m;
                    vT = mT;
                }
            }
            if (norm_val) norm_val[p] = v;
            if (norm_val_T) norm_val_T[p] = vT;
            if (norm_val_coo) norm_val_coo[q] = v;
        }
    }
}

// One warp per node; positions (9.8 MB at products scale) stay L2-resident.
__global__ void arrange_sweep_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                     const float* __restrict__ theta_in, float eps, float* __restrict__ theta_out,
                                     int64_t n) {
    const float two_pi = 6.283185307179586f;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int s = row_ptr[i], e = row_ptr[i + 1];
        const float ti = theta_in[i];
        float num = 0.0f, den = 0.0f;
        for (int p = s + lane; p < e; p += 32) {
            float d = __ldg(theta_in + col_idx[p]) - ti;
            d -= two_pi * rintf(d * (1.0f / two_pi));  // signed circular offset in (-pi, pi]
            const float w = 1.0f / (fabsf(d) + eps);
            num += w * d;
            den += w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            num += __shfl_xor_sync(0xffffffffu, num, o);
            den += __shfl_xor_sync(0xffffffffu, den, o);
        }
        if (lane == 0) {
            float t = (den > 0.0f) ? ti + num / den : ti;
            t -= two_pi * floorf(t * (1.0f / two_pi));
            theta_out[i] = t;
        }
    }
}

}  // namespace gnntf

using namespace gnntf;

extern "C" int gnntf_csr_build_ws_bytes(int64_t n, int64_t n_edges, int directed, int add_eye,
                                        size_t* bytes) {
    if (bytes == nullptr) return GNNTF_E_NULL;
    int64_t n_graph, nnz;
    int rc = coo_nnz(n, n_edges, directed, add_eye, n_graph, nnz);
    if (rc != GNNTF_OK) return rc;
    BuildWs L;
    rc = build_ws_layout(nnz, n, L);
    if (rc != GNNTF_OK) return rc;
    *bytes = L.total;
    return GNNTF_OK;
}

extern "C" int gnntf_csr_build(const int64_t* edges, const float* weights, int64_t n, int64_t n_edges,
                               int directed, int add_eye, int by_column, int64_t* coo_indices,
                               float* coo_values, int32_t* row_ptr, int32_t* col_idx, float* raw_val,
                               int32_t* coo_pos, void* ws, size_t ws_bytes, void* stream) {
    int64_t n_graph, nnz;
    int rc = coo_nnz(n, n_edges, directed, add_eye, n_graph, nnz);
    if (rc != GNNTF_OK) return rc;
    if (row_ptr == nullptr) return GNNTF_E_NULL;
    if (n_edges > 0 && edges == nullptr) return GNNTF_E_NULL;
    if (nnz > 0 && (col_idx == nullptr || raw_val == nullptr || coo_pos == nullptr || ws == nullptr))
        return GNNTF_E_NULL;
    // the COO index pairs are written as one 16-byte store each
    if (coo_indices != nullptr && (reinterpret_cast<uintptr_t>(coo_indices) & 15u) != 0) return GNNTF_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    if (nnz == 0) {
        GNNTF_CUDA_TRY(cudaMemsetAsync(row_ptr, 0, (size_t)(n + 1) * sizeof(int32_t), st));
        return GNNTF_OK;
    }
    BuildWs L;
    rc = build_ws_layout(nnz, n, L);
    if (rc != GNNTF_OK) return rc;
    const uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
    if (base + L.total - 256 > reinterpret_cast<uintptr_t>(ws) + ws_bytes) return GNNTF_E_WORKSPACE;
    uint32_t* keys_in = reinterpret_cast<uint32_t*>(base + L.keys_in);
    uint32_t* keys_out = reinterpret_cast<uint32_t*>(base + L.keys_out);
    int32_t* pos_in = reinterpret_cast<int32_t*>(base + L.pos_in);
    void* cub_ws = reinterpret_cast<void*>(base + L.cub);

    const int threads = 256;
    const int grid = (int)std::min<int64_t>(ceil_div(nnz, threads), (int64_t)kNumSMs * 16);
    emit_coo_kernel<<<grid, threads, 0, st>>>(edges, weights, n_edges, n_graph, nnz, by_column,
                                              coo_indices, coo_values, keys_in, pos_in);
    GNNTF_LAUNCH_CHECK();
    size_t cub_bytes = L.cub_bytes;
    GNNTF_CUDA_TRY((cub::DeviceRadixSort::SortPairs<uint32_t, int32_t>(
        cub_ws, cub_bytes, keys_in, keys_out, pos_in, coo_pos, (int)nnz, 0, key_bits(n), st)));
    row_ptr_kernel<<<grid, threads, 0, st>>>(keys_out, nnz, n, row_ptr);
    GNNTF_LAUNCH_CHECK();
    gather_csr_kernel<<<grid, threads, 0, st>>>(edges, weights, n_edges, n_graph, nnz, by_column,
                                                coo_pos, col_idx, raw_val);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_normalize_f32(const int32_t* row_ptr, const int32_t* col_idx, const float* raw_val,
                                   const int32_t* coo_pos, int64_t n, int64_t nnz, int64_t n_graph,
                                   int directed, const uint8_t* keep_mask_coo, float scale, int mode,
                                   int eye_mode, float* deg, float* dinv, float* norm_val,
                                   float* norm_val_T, float* norm_val_coo, void* stream) {
    if (mode != GNNTF_NORM_SYMMETRIC && mode != GNNTF_NORM_BIPARTITE && mode != GNNTF_NORM_NONE)
        return GNNTF_E_MODE;  // "Invalid matrix normalization", gnn.py:46-47
    if (eye_mode != GNNTF_EYE_NONE && eye_mode != GNNTF_EYE_BEFORE && eye_mode != GNNTF_EYE_AFTER)
        return GNNTF_E_MODE;
    if (n < 0 || nnz < 0 || n_graph < 0 || n_graph > nnz || nnz > 0x7fffffffLL || n > 0x7ffffffeLL)
        return GNNTF_E_SIZE;
    if ((eye_mode == GNNTF_EYE_NONE) != (nnz == n_graph) && n > 0) return GNNTF_E_SIZE;
    if (n == 0) return GNNTF_OK;
    if (row_ptr == nullptr || deg == nullptr || dinv == nullptr) return GNNTF_E_NULL;
    if (nnz > 0 && (col_idx == nullptr || raw_val == nullptr || coo_pos == nullptr)) return GNNTF_E_NULL;
    if (nnz > 0 && norm_val == nullptr && norm_val_T == nullptr && norm_val_coo == nullptr) return GNNTF_E_NULL;  // nothing to produce
    if (directed && norm_val_T != nullptr) return GNNTF_E_MODE;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    const int64_t E = directed ? n_graph : n_graph / 2;
    const int grid_rows = (int)std::min<int64_t>(ceil_div(n, threads / 32), (int64_t)kNumSMs * 16);
    const int grid_n = (int)std::min<int64_t>(ceil_div(n, threads), (int64_t)kNumSMs * 16);
    if (mode != GNNTF_NORM_NONE) {
        if (directed) {
            GNNTF_CUDA_TRY(cudaMemsetAsync(deg, 0, (size_t)n * sizeof(float), st));
            if (nnz > 0) {
                const int grid_nnz = (int)std::min<int64_t>(ceil_div(nnz, threads), (int64_t)kNumSMs * 16);
                degree_atomic_kernel<<<grid_nnz, threads, 0, st>>>(col_idx, raw_val, coo_pos, nnz, n_graph,
                                                                   keep_mask_coo, scale, eye_mode, deg);
                GNNTF_LAUNCH_CHECK();
            }
        } else {
            degree_rows_kernel<<<grid_rows, threads, 0, st>>>(row_ptr, raw_val, coo_pos, n, n_graph, E,
                                                              keep_mask_coo, scale, eye_mode, deg);
            GNNTF_LAUNCH_CHECK();
        }
        dinv_kernel<<<grid_n, threads, 0, st>>>(deg, n, mode, dinv);
        GNNTF_LAUNCH_CHECK();
    }
    if (nnz > 0) {
        scale_values_kernel<<<grid_rows, threads, 0, st>>>(row_ptr, col_idx, raw_val, coo_pos, n, n_graph,
                                                           E, directed, keep_mask_coo, scale, mode,
                                                           eye_mode, dinv, norm_val, norm_val_T,
                                                           norm_val_coo);
        GNNTF_LAUNCH_CHECK();
    }
    return GNNTF_OK;
}

extern "C" int gnntf_arrange_sweep_f32(const int32_t* row_ptr, const int32_t* col_idx, const float* theta_in,
                                       float eps, float* theta_out, int64_t n, void* stream) {
    if (n < 0 || n > 0x7ffffffeLL || !(eps > 0.0f)) return GNNTF_E_SIZE;
    if (n == 0) return GNNTF_OK;
    if (row_ptr == nullptr || theta_in == nullptr || theta_out == nullptr) return GNNTF_E_NULL;
    const int grid = (int)std::min<int64_t>(ceil_div(n, 8), (int64_t)kNumSMs * 16);
    arrange_sweep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(row_ptr, col_idx, theta_in, eps, theta_out, n);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}
