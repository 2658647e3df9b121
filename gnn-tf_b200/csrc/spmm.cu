// CSR SpMM for sm_100a: row-split main kernel + split long rows (nnz-chunked, fixed-order reduce).
//
// Replaces tf.sparse.sparse_dense_matmul at gnntf/core/gnn/architectures/filter.py:19 and
// gcn.py:24,48,88,104,131, fused with the teleport axpy / dropout / activation of
// filter.py:21-22.  The path is bandwidth-bound (AI < 7 flop/B): no tensor cores.
//
// Lane mapping.  A feature row of F floats is cut into VEC-wide slots (VEC = 4 -> LDG.128 when
// every leading dimension and base pointer allows it, else scalar).  GROUP lanes (a power of two
// <= 32) own one sparse row; lane `gl` of the group owns slots gl, gl+GROUP, ... (NSLOT of them),
// so the lanes of a group read one contiguous GROUP*VEC*4-byte span of each gathered row.  A warp
// therefore runs 32/GROUP sparse rows at once.  The group's lanes fetch GROUP (col,val) pairs with
// one coalesced streaming load each and broadcast them with width-GROUP shuffles; the gathers of
// UNROLL consecutive entries are issued back to back before the FMAs (memory-level parallelism).
//
// Accumulation order inside a row is CSR slot order = COO storage order (the builder's sort is
// stable), i.e. the order TF's CPU kernel uses.  Rows longer than A->long_threshold are skipped
// here and run as fixed-size pieces on separate warps (spmm_chunk_kernel), whose partial sums are
// combined in piece order by spmm_long_reduce_kernel: deterministic, no float atomics.
#include <algorithm>

#include "spmm.cuh"

namespace gnntf {


template <int VEC>
__device__ __forceinline__ void apply_epilogue(const Epilogue& e, int64_t row, int f, Vec<VEC> acc) {
    Vec<VEC> out;
#pragma unroll
    for (int i = 0; i < VEC; ++i) out.v[i] = acc.v[i] * e.s;
    if (e.H0 != nullptr) {
        Vec<VEC> h = Vec<VEC>::stream(e.H0 + row * e.ldh + f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = __fadd_rn(out.v[i], __fmul_rn(h.v[i], e.t));
    }
    if (e.keep != nullptr) {
        const uint8_t* k = e.keep + row * (int64_t)e.F + f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = k[i] ? out.v[i] * e.p_scale : 0.0f;
    }
    if (e.act == GNNTF_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = fmaxf(out.v[i], 0.0f);
    }
    if (e.C != nullptr) out.store(e.C + row * e.ldc + f);
    if (e.ACC != nullptr) {
        float* a = e.ACC + row * e.ldacc + f;
        Vec<VEC> b = Vec<VEC>::stream(e.B + row * e.ldb + f);
        Vec<VEC> r;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float prev = e.acc_init ? 0.0f : a[i];
            r.v[i] = prev + e.u * b.v[i] + e.w * out.v[i];
        }
        r.store(a);
    }
}

// ---------------------------------------------------------------------------------------------
// Main kernel: GROUP lanes per sparse row, rows with deg > long_threshold are left to the
// chunk path.  grid.x walks row blocks, grid.y walks feature tiles of GROUP*NSLOT*VEC floats.
// ---------------------------------------------------------------------------------------------
template <int VEC, int NSLOT, int GROUP, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS)
spmm_rows_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                 const float* __restrict__ val, const int* __restrict__ row_map,
                 const float* __restrict__ B, int64_t ldb, int n_rows, int long_threshold, Epilogue epi) {
    constexpr int ROWS_PER_WARP = 32 / GROUP;
    constexpr int WARPS = THREADS / 32;
    const int lane = threadIdx.x & 31;
    const int g = lane / GROUP;
    const int gl = lane % GROUP;
    const int64_t row = ((int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * ROWS_PER_WARP + g;
    const int f_base = blockIdx.y * (GROUP * NSLOT * VEC);
    const int F = epi.F;

    int start = 0, deg = 0;
    bool mine = false;
    if (row < n_rows) {
        start = __ldg(row_ptr + row);
        deg = __ldg(row_ptr + row + 1) - start;
        mine = !(long_threshold > 0 && deg > long_threshold);
        if (!mine) deg = 0;
    }
    int maxdeg = deg;
    if (ROWS_PER_WARP > 1) {
#pragma unroll
        for (int o = GROUP; o < 32; o <<= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
    }

    int fo[NSLOT];
    bool fok[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        fo[s] = f_base + (s * GROUP + gl) * VEC;
        fok[s] = fo[s] < F;  // VEC == 4 implies F % 4 == 0, so the whole slot is in range
    }
    Vec<VEC> acc[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[s].v[i] = 0.0f;

    for (int off = 0; off < maxdeg; off += GROUP) {
        int c = 0;
        float v = 0.0f;
        if (off + gl < deg) {
            c = ld_stream(col_idx + start + off + gl);
            v = ld_stream(val + start + off + gl);
        }
        const int lim = min(GROUP, maxdeg - off);
        for (int j = 0; j < lim; j += UNROLL) {
            int cj[UNROLL];
            float vj[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                cj[u] = __shfl_sync(0xffffffffu, c, j + u, GROUP);
                vj[u] = __shfl_sync(0xffffffffu, v, j + u, GROUP);
            }
            Vec<VEC> x[UNROLL][NSLOT];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const bool live = (j + u < GROUP) && (off + j + u < deg);
                const float* src = B + (int64_t)cj[u] * ldb;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    if (live && fok[s]) {
                        x[u][s] = Vec<VEC>::gather(src + fo[s]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) x[u][s].v[i] = 0.0f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[s].v[i] = fmaf(vj[u], x[u][s].v[i], acc[s].v[i]);
        }
    }
    if (mine) {
        const int64_t out_row = row_map ? (int64_t)__ldg(row_map + row) : row;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if (fok[s]) apply_epilogue<VEC>(epi, out_row, fo[s], acc[s]);
    }
}

// ---------------------------------------------------------------------------------------------
// Long rows, phase 1: one warp per piece.  The 32/GROUP lane groups take the piece's entries
// round-robin and are combined with xor-shuffles in a fixed pattern.
// partials[piece, ldp] (ldp = round_up(F,4)).
// ---------------------------------------------------------------------------------------------
template <int VEC, int NSLOT, int GROUP, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS)
spmm_chunk_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                  const float* __restrict__ val, const float* __restrict__ B, int64_t ldb,
                  const int* __restrict__ chunk_row, const int* __restrict__ chunk_begin,
                  int n_chunks, int chunk, float* __restrict__ partials, int ldp, int F) {
    constexpr int NGROUPS = 32 / GROUP;
    constexpr int WARPS = THREADS / 32;
    const int lane = threadIdx.x & 31;
    const int g = lane / GROUP;
    const int gl = lane % GROUP;
    const int piece = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (piece >= n_chunks) return;  // warp-uniform
    const int f_base = blockIdx.y * (GROUP * NSLOT * VEC);
    const int row = __ldg(chunk_row + piece);
    const int begin = __ldg(chunk_begin + piece);
    const int end = min(begin + chunk, __ldg(row_ptr + row + 1));

    int fo[NSLOT];
    bool fok[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        fo[s] = f_base + (s * GROUP + gl) * VEC;
        fok[s] = fo[s] < F;
    }
    Vec<VEC> acc[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[s].v[i] = 0.0f;

    // 32 entries per outer step: lane l holds entry base+l; group g consumes entries with
    // (index % NGROUPS) == g.
    for (int base = begin; base < end; base += 32) {
        int c = 0;
        float v = 0.0f;
        if (base + lane < end) {
            c = ld_stream(col_idx + base + lane);
            v = ld_stream(val + base + lane);
        }
        const int cnt = min(32, end - base);
        for (int j = 0; j < cnt; j += NGROUPS * UNROLL) {
            int cj[UNROLL];
            float vj[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int src = j + u * NGROUPS + g;
                cj[u] = __shfl_sync(0xffffffffu, c, src & 31);
                vj[u] = __shfl_sync(0xffffffffu, v, src & 31);
                if (src >= cnt) vj[u] = 0.0f, cj[u] = -1;
            }
            Vec<VEC> x[UNROLL][NSLOT];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float* src = B + (int64_t)cj[u] * ldb;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    if (cj[u] >= 0 && fok[s]) {
                        x[u][s] = Vec<VEC>::gather(src + fo[s]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) x[u][s].v[i] = 0.0f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[s].v[i] = fmaf(vj[u], x[u][s].v[i], acc[s].v[i]);
        }
    }
    if (NGROUPS > 1) {
#pragma unroll
        for (int o = GROUP; o < 32; o <<= 1)
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[s].v[i] += __shfl_xor_sync(0xffffffffu, acc[s].v[i], o);
    }
    if (g == 0) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if (fok[s]) acc[s].store(partials + (int64_t)piece * ldp + fo[s]);
    }
}

// Long rows, phase 2: one CTA per long row; thread t owns feature t (strided), sums the row's
// pieces in piece order, then applies the epilogue.
__global__ void __launch_bounds__(128)
spmm_long_reduce_kernel(const int* __restrict__ long_row, const int* __restrict__ long_first_chunk,
                        const int* __restrict__ long_n_chunks, const int* __restrict__ row_map,
                        const float* __restrict__ partials, int ldp, Epilogue epi) {
    const int li = blockIdx.x;
    const int csr_row = __ldg(long_row + li);
    const int64_t row = row_map ? (int64_t)__ldg(row_map + csr_row) : (int64_t)csr_row;
    const int first = __ldg(long_first_chunk + li);
    const int cnt = __ldg(long_n_chunks + li);
    for (int f = threadIdx.x; f < epi.F; f += blockDim.x) {
        const float* p = partials + (int64_t)first * ldp + f;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // 4 independent chains, fixed assignment
        int k = 0;
        for (; k + 4 <= cnt; k += 4) {
            a0 += ld_stream(p + (int64_t)(k + 0) * ldp);
            a1 += ld_stream(p + (int64_t)(k + 1) * ldp);
            a2 += ld_stream(p + (int64_t)(k + 2) * ldp);
            a3 += ld_stream(p + (int64_t)(k + 3) * ldp);
        }
        for (; k < cnt; ++k) a0 += ld_stream(p + (int64_t)k * ldp);
        Vec<1> acc{{(a0 + a1) + (a2 + a3)}};
        apply_epilogue<1>(epi, row, f, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// Host-side dispatch
// ---------------------------------------------------------------------------------------------
template <int VEC, int NSLOT, int GROUP>
static int launch_cfg(const gnntf_csr_t* A, const float* B, int64_t ldb, const Epilogue& epi,
                      cudaStream_t st) {
    constexpr int THREADS = 256;
    constexpr int UNROLL = (NSLOT >= 4) ? 2 : (NSLOT == 2 ? 4 : (GROUP >= 8 ? 8 : 4));
    constexpr int ROWS_PER_CTA = (THREADS / 32) * (32 / GROUP);
    const int F = epi.F;
    const int tile = GROUP * NSLOT * VEC;
    const unsigned gy = (unsigned)ceil_div(F, tile);
    const int thr = (A->n_long > 0) ? A->long_threshold : 0;
    if (A->n_rows > 0) {
        dim3 grid((unsigned)ceil_div(A->n_rows, ROWS_PER_CTA), gy);
        spmm_rows_kernel<VEC, NSLOT, GROUP, UNROLL, THREADS><<<grid, THREADS, 0, st>>>(
            A->row_ptr, A->col_idx, A->val, A->row_map, B, ldb, (int)A->n_rows, thr, epi);
        GNNTF_LAUNCH_CHECK();
    }
    if (A->n_long > 0) {
        const int ldp = (int)round_up(F, 4);
        constexpr int CU = (NSLOT >= 4) ? 2 : 4;
        dim3 grid((unsigned)ceil_div(A->n_chunks, THREADS / 32), gy);
        spmm_chunk_kernel<VEC, NSLOT, GROUP, CU, THREADS><<<grid, THREADS, 0, st>>>(
            A->row_ptr, A->col_idx, A->val, B, ldb, A->chunk_row, A->chunk_begin, A->n_chunks,
            A->chunk, A->partials, ldp, F);
        GNNTF_LAUNCH_CHECK();
        spmm_long_reduce_kernel<<<A->n_long, 128, 0, st>>>(A->long_row, A->long_first_chunk,
                                                          A->long_n_chunks, A->row_map, A->partials,
                                                          ldp, epi);
        GNNTF_LAUNCH_CHECK();
    }
    return GNNTF_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int validate_csr(const gnntf_csr_t* A) {
    if (A == nullptr) return GNNTF_E_NULL;
    if (A->n_rows < 0 || A->nnz < 0 || A->nnz > 0x7fffffffLL || A->n_rows > 0x7ffffffeLL) return GNNTF_E_SIZE;
    if (A->n_rows > 0 && A->row_ptr == nullptr) return GNNTF_E_NULL;
    if (A->nnz > 0 && (A->col_idx == nullptr || A->val == nullptr)) return GNNTF_E_NULL;
    if (A->n_long < 0 || A->n_chunks < 0) return GNNTF_E_SIZE;
    if (A->n_long > 0) {
        if (A->long_row == nullptr || A->long_first_chunk == nullptr || A->long_n_chunks == nullptr ||
            A->chunk_row == nullptr || A->chunk_begin == nullptr || A->partials == nullptr)
            return GNNTF_E_NULL;
        if (A->long_threshold <= 0 || A->chunk <= 0) return GNNTF_E_SIZE;
    }
    return GNNTF_OK;
}

// C (and ACC) = epilogue(A·B).  Chooses the lane mapping from F and the alignment of every
// operand; see the file header.
int spmm_dispatch(const gnntf_csr_t* A, const float* B, int64_t ldb, Epilogue epi, cudaStream_t st) {
    int rc = validate_csr(A);
    if (rc != GNNTF_OK) return rc;
    const int64_t F = epi.F;
    if (F < 0 || F > 0x7fffffff) return GNNTF_E_SIZE;
    if (A->n_rows == 0 || F == 0) return GNNTF_OK;
    if (B == nullptr) return GNNTF_E_NULL;
    if (ldb < F || (epi.C && epi.ldc < F) || (epi.H0 && epi.ldh < F) || (epi.ACC && epi.ldacc < F))
        return GNNTF_E_SIZE;
    epi.B = B;
    epi.ldb = ldb;

    bool v4 = (F % 4 == 0) && (ldb % 4 == 0) && aligned16(B);
    if (epi.C) v4 = v4 && (epi.ldc % 4 == 0) && aligned16(epi.C);
    if (epi.H0) v4 = v4 && (epi.ldh % 4 == 0) && aligned16(epi.H0);
    if (epi.ACC) v4 = v4 && (epi.ldacc % 4 == 0) && aligned16(epi.ACC);
    if (epi.keep) v4 = false;  // byte mask rows are F-strided: keep the scalar path

    if (v4) {
        const int64_t slots = F / 4;
        if (slots <= 4) return launch_cfg<4, 1, 4>(A, B, ldb, epi, st);
        if (slots <= 8) return launch_cfg<4, 1, 8>(A, B, ldb, epi, st);
        if (slots <= 16) return launch_cfg<4, 1, 16>(A, B, ldb, epi, st);
        if (slots <= 32) return launch_cfg<4, 1, 32>(A, B, ldb, epi, st);
        if (slots <= 64) return launch_cfg<4, 2, 32>(A, B, ldb, epi, st);
        return launch_cfg<4, 4, 32>(A, B, ldb, epi, st);  // tiles of 512 floats over grid.y
    }
    if (F <= 4) return launch_cfg<1, 1, 4>(A, B, ldb, epi, st);
    if (F <= 8) return launch_cfg<1, 1, 8>(A, B, ldb, epi, st);
    if (F <= 16) return launch_cfg<1, 1, 16>(A, B, ldb, epi, st);
    if (F <= 32) return launch_cfg<1, 1, 32>(A, B, ldb, epi, st);
    if (F <= 64) return launch_cfg<1, 2, 32>(A, B, ldb, epi, st);
    return launch_cfg<1, 4, 32>(A, B, ldb, epi, st);  // tiles of 128 floats over grid.y
}

// ---------------------------------------------------------------------------------------------
// Long-row plan
// ---------------------------------------------------------------------------------------------
__global__ void plan_count_kernel(const int* __restrict__ row_ptr, int n_rows, int thr, int chunk,
                                  int* __restrict__ counts) {
    int nl = 0, nc = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int deg = row_ptr[r + 1] - row_ptr[r];
        if (deg > thr) {
            nl += 1;
            nc += (deg + chunk - 1) / chunk;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nl += __shfl_xor_sync(0xffffffffu, nl, o);
        nc += __shfl_xor_sync(0xffffffffu, nc, o);
    }
    if ((threadIdx.x & 31) == 0 && nl > 0) {
        atomicAdd(counts + 0, nl);
        atomicAdd(counts + 1, nc);
    }
}

// Long rows are rare; one thread claims a contiguous range of pieces for its row.  The order in
// which rows claim ranges is arbitrary but never changes a result (pieces of one row stay
// contiguous and ordered).
__global__ void plan_fill_kernel(const int* __restrict__ row_ptr, int n_rows, int thr, int chunk,
                                 int* __restrict__ counters, int* __restrict__ long_row,
                                 int* __restrict__ long_first_chunk, int* __restrict__ long_n_chunks,
                                 int* __restrict__ chunk_row, int* __restrict__ chunk_begin) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int start = row_ptr[r];
        const int deg = row_ptr[r + 1] - start;
        if (deg > thr) {
            const int nch = (deg + chunk - 1) / chunk;
            const int li = atomicAdd(counters + 0, 1);
            const int first = atomicAdd(counters + 1, nch);
            long_row[li] = (int)r;
            long_first_chunk[li] = first;
            long_n_chunks[li] = nch;
            for (int j = 0; j < nch; ++j) {
                chunk_row[first + j] = (int)r;
                chunk_begin[first + j] = start + j * chunk;
            }
        }
    }
}

}  // namespace gnntf

using namespace gnntf;

extern "C" int gnntf_spmm_plan_count(const int32_t* row_ptr, int64_t n_rows, int32_t long_threshold,
                                     int32_t chunk, int32_t* counts, void* stream) {
    if (counts == nullptr || (n_rows > 0 && row_ptr == nullptr)) return GNNTF_E_NULL;
    if (n_rows < 0 || n_rows > 0x7ffffffeLL || long_threshold <= 0 || chunk <= 0) return GNNTF_E_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    GNNTF_CUDA_TRY(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st));
    if (n_rows == 0) return GNNTF_OK;
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, 256), (int64_t)kNumSMs * 8);
    plan_count_kernel<<<grid, 256, 0, st>>>(row_ptr, (int)n_rows, long_threshold, chunk, counts);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_spmm_plan_fill(const int32_t* row_ptr, int64_t n_rows, int32_t long_threshold,
                                    int32_t chunk, int32_t* counters_ws, int32_t* long_row,
                                    int32_t* long_first_chunk, int32_t* long_n_chunks,
                                    int32_t* chunk_row, int32_t* chunk_begin, void* stream) {
    if (counters_ws == nullptr || (n_rows > 0 && row_ptr == nullptr)) return GNNTF_E_NULL;
    if (n_rows < 0 || n_rows > 0x7ffffffeLL || long_threshold <= 0 || chunk <= 0) return GNNTF_E_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    GNNTF_CUDA_TRY(cudaMemsetAsync(counters_ws, 0, 2 * sizeof(int32_t), st));
    if (n_rows == 0) return GNNTF_OK;
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, 256), (int64_t)kNumSMs * 8);
    plan_fill_kernel<<<grid, 256, 0, st>>>(row_ptr, (int)n_rows, long_threshold, chunk, counters_ws,
                                           long_row, long_first_chunk, long_n_chunks, chunk_row,
                                           chunk_begin);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_spmm_f32(const gnntf_csr_t* A, const float* B, int64_t ldb, float* C,
                              int64_t ldc, int64_t F, void* stream) {
    if (C == nullptr && A != nullptr && A->n_rows > 0 && F > 0) return GNNTF_E_NULL;
    Epilogue e{};
    e.s = 1.0f;
    e.C = C;
    e.ldc = ldc;
    e.F = (int)F;
    e.act = GNNTF_ACT_IDENTITY;
    if (F < 0 || F > 0x7fffffff) return GNNTF_E_SIZE;
    return spmm_dispatch(A, B, ldb, e, (cudaStream_t)stream);
}
