/* CPU oracle (C) for gnntf's sparse adjacency propagation path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference algorithm, used (a) by tests/ as the checker at sizes
 * where the NumPy oracle is too slow and (b) by bench.py's cpu_baseline / --impl reference
 * legs as the timed CPU arm.  Nothing in the product (gnn-tf_b200/) links or calls it.
 *
 * PARITY UNPINNED: the reference's arithmetic lives in TensorFlow (un-vendored, un-pinned,
 * not installable here), and the reference has no tests or golden vectors for this path; see
 * oracle/gnntf_oracle.py for the assumptions.  This file is checked against that NumPy
 * oracle (tests/test_oracle.py), which is itself pinned only by SURVEY.md's KAT-1/KAT-2.
 *
 * Threading mirrors TF-CPU: the SparseTensorDenseMatMul CPU functor is one sequential loop
 * over the COO entries (kept single-threaded here); the element-wise ops around it run on
 * Eigen's thread pool in TF (OpenMP here when compiled with -fopenmp).
 *
 * Reference lines followed (relative to the reference root):
 *   gnntf/core/gnn/gnn.py:40-42                     column sums, D = divide_no_nan(1, sqrt(deg)), row then col scale
 *   gnntf/core/gnn/architectures/filter.py:19       propagated = sparse_dense_matmul(G, features)
 *   gnntf/core/gnn/architectures/filter.py:21       propagated*(1-a) + H0.value*a
 *   gnntf/core/nn/layered.py:52-55                  the K-iteration driver loop
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* tf.sparse.reduce_sum(A, axis=0): one sum per COLUMN (gnn.py:41). */
void oracle_colsum_f32(const int64_t* idx, const float* val, int64_t nnz, int64_t n, float* deg) {
    memset(deg, 0, (size_t)n * sizeof(float));
    for (int64_t i = 0; i < nnz; ++i) deg[idx[2 * i + 1]] += val[i];
}

/* gnn.py:41-42: D = divide_no_nan(1, sqrt(deg)); v_i = (v_i * D[row_i]) * D[col_i]. */
void oracle_normalize_sym_f32(const int64_t* idx, const float* val, int64_t nnz, int64_t n,
                              float* D, float* out) {
    oracle_colsum_f32(idx, val, nnz, n, D);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) {
        float s = sqrtf(D[j]);
        D[j] = (s == 0.0f) ? 0.0f : 1.0f / s;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nnz; ++i) {
        float r = val[i] * D[idx[2 * i]];
        out[i] = r * D[idx[2 * i + 1]];
    }
}

/* tf.sparse.sparse_dense_matmul on CPU (filter.py:19): out[row_i,:] += val_i * H[col_i,:] for the
 * COO entries [lo, hi) in storage order; `out` must be zeroed by the caller for lo == 0. */
void oracle_spmm_coo_f32(const int64_t* idx, const float* val, int64_t lo, int64_t hi,
                         const float* H, int64_t F, float* out) {
    for (int64_t i = lo; i < hi; ++i) {
        const float v = val[i];
        float* __restrict__ o = out + idx[2 * i] * F;
        const float* __restrict__ h = H + idx[2 * i + 1] * F;
        for (int64_t f = 0; f < F; ++f) o[f] += v * h[f];
    }
}

/* filter.py:21: act = propagated*(1-a) + H0*a  (two multiplies, one add; no fused contraction). */
void oracle_teleport_f32(const float* P, const float* H0, int64_t count, float a, float* out) {
    const float one_minus_a = (float)(1 - (double)a), af = a;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i) {
        volatile float x = P[i] * one_minus_a;
        volatile float y = H0[i] * af;
        out[i] = x + y;
    }
}

/* K PPRIteration layers in eval mode (filter.py:17-22 under layered.py:52-55).  As in the
 * reference, every iteration re-runs get_adjacency (filter.py:18) when renormalize != 0;
 * with renormalize == 0 the normalised values are computed once (same result in eval mode).
 * scratch: nnz floats (normalised values) + n floats (D) + n*F floats (propagated). */
void oracle_appnp_propagate_f32(const int64_t* idx, const float* raw_val, int64_t nnz, int64_t n,
                                const float* H0, int64_t F, float a, int K, int renormalize,
                                float* scratch, float* H_out) {
    float* norm = scratch;
    float* D = norm + nnz;
    float* P = D + n;
    const float* H = H0;
    for (int k = 0; k < K; ++k) {
        if (k == 0 || renormalize) oracle_normalize_sym_f32(idx, raw_val, nnz, n, D, norm);
        memset(P, 0, (size_t)n * F * sizeof(float));
        oracle_spmm_coo_f32(idx, norm, 0, nnz, H, F, P);
        oracle_teleport_f32(P, H0, n * F, a, H_out);
        H = H_out;
    }
}

/* Context only, NOT what the reference executes: one PPR iteration as a row-parallel CSR SpMM with
 * the teleport fused, on all OpenMP threads — the best a CPU port could reasonably do with the
 * host's cores.  bench.py reports it beside the faithful single-threaded COO loop above
 * (cpu_baseline.mt_value) so the GPU/CPU ratio is not only a ratio against one core.
 * row_ptr int64 [n+1], col int32 [nnz], val fp32 [nnz] (normalised), rows [r_lo, r_hi). */
void oracle_appnp_step_csr_omp_f32(const int64_t* row_ptr, const int32_t* col, const float* val,
                                   const float* H, const float* H0, int64_t F, float a,
                                   int64_t r_lo, int64_t r_hi, float* out) {
    const float one_minus_a = (float)(1 - (double)a);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = r_lo; r < r_hi; ++r) {
        float* __restrict__ o = out + r * F;
        for (int64_t f = 0; f < F; ++f) o[f] = 0.0f;
        for (int64_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
            const float v = val[p];
            const float* __restrict__ h = H + (int64_t)col[p] * F;
            for (int64_t f = 0; f < F; ++f) o[f] += v * h[f];
        }
        const float* __restrict__ h0 = H0 + r * F;
        for (int64_t f = 0; f < F; ++f) o[f] = o[f] * one_minus_a + h0[f] * a;
    }
}

/* Stable-by-row CSR view of a COO list, in O(nnz) (counting sort).  The reference never builds a
 * CSR: tf.sparse.sparse_dense_matmul (filter.py:19) walks the COO entries in storage order, so each
 * output row accumulates ITS entries in storage order.  A stable sort by row keeps exactly that
 * order inside every row, hence a row-wise loop over this CSR performs the same fp32 operations on
 * every output element as the sequential COO loop (oracle_spmm_coo_f32) — bit-identical results —
 * while letting rows run on different threads.  This is what lets the tests and bench.py check
 * full-size configs in seconds.  idx: int64 [nnz,2] (graph_manipulation.py:31 order);
 * row_ptr int64 [n+1], col int32 [nnz], coo_pos int64 [nnz] (COO slot of each CSR slot). */
void oracle_csr_from_coo(const int64_t* idx, int64_t nnz, int64_t n, int64_t* row_ptr, int32_t* col,
                         int64_t* coo_pos) {
    memset(row_ptr, 0, (size_t)(n + 1) * sizeof(int64_t));
    for (int64_t i = 0; i < nnz; ++i) row_ptr[idx[2 * i] + 1] += 1;
    for (int64_t r = 0; r < n; ++r) row_ptr[r + 1] += row_ptr[r];
    int64_t* cursor = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    memcpy(cursor, row_ptr, (size_t)n * sizeof(int64_t));
    for (int64_t i = 0; i < nnz; ++i) {
        const int64_t p = cursor[idx[2 * i]]++;
        col[p] = (int32_t)idx[2 * i + 1];
        coo_pos[p] = i;
    }
    free(cursor);
}

/* out[p] = src[pos[p]] (values from COO order into CSR order). */
void oracle_gather_f32(const float* src, const int64_t* pos, int64_t count, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < count; ++p) out[p] = src[pos[p]];
}

/* Row-wise SpMM over the stable CSR (same per-element operation sequence as oracle_spmm_coo_f32,
 * see oracle_csr_from_coo), rows on OpenMP threads.  filter.py:19 / gcn.py:88. */
void oracle_spmm_csr_omp_f32(const int64_t* row_ptr, const int32_t* col, const float* val,
                             const float* H, int64_t F, int64_t r_lo, int64_t r_hi, float* out) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = r_lo; r < r_hi; ++r) {
        float* __restrict__ o = out + r * F;
        for (int64_t f = 0; f < F; ++f) o[f] = 0.0f;
        for (int64_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
            const float v = val[p];
            const float* __restrict__ h = H + (int64_t)col[p] * F;
            for (int64_t f = 0; f < F; ++f) o[f] += v * h[f];
        }
    }
}

/* K PPRIteration layers (filter.py:17-22 under layered.py:52-55), eval mode, over the stable CSR:
 * bit-identical to oracle_appnp_propagate_f32 on the same normalised values, but rows run on all
 * OpenMP threads.  scratch: n*F floats.  The result ends in H_out. */
void oracle_appnp_propagate_csr_omp_f32(const int64_t* row_ptr, const int32_t* col, const float* val,
                                        int64_t n, const float* H0, int64_t F, float a, int K,
                                        float* scratch, float* H_out) {
    if (K == 0) {
        memcpy(H_out, H0, (size_t)n * F * sizeof(float));
        return;
    }
    const float* src = H0;
    for (int k = 0; k < K; ++k) {
        float* dst = ((K - 1 - k) % 2 == 0) ? H_out : scratch;
        oracle_spmm_csr_omp_f32(row_ptr, col, val, src, F, 0, n, dst);
        oracle_teleport_f32(dst, H0, n * F, a, dst);
        src = dst;
    }
}

/* Same row-wise SpMM with the row sum accumulated in DOUBLE and rounded to fp32 once per element:
 * the error-budget twin for rows where a sequential fp32 sum is itself the dominant error (hub rows
 * with 10^4 entries: the fp32 sequential sum — TF-CPU's and oracle_spmm_coo_f32's — sits ~5e-6·‖y‖∞
 * from this; an implementation that sums such a row in pieces is CLOSER to it than TF is).
 * Not what the reference computes; used only to judge rows the GPU splits. */
void oracle_spmm_csr_omp_acc64_f32(const int64_t* row_ptr, const int32_t* col, const float* val,
                                   const float* H, int64_t F, int64_t r_lo, int64_t r_hi, float* out) {
#pragma omp parallel
    {
        double* acc = (double*)malloc((size_t)(F > 0 ? F : 1) * sizeof(double));
#pragma omp for schedule(dynamic, 256)
        for (int64_t r = r_lo; r < r_hi; ++r) {
            for (int64_t f = 0; f < F; ++f) acc[f] = 0.0;
            for (int64_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
                const double v = val[p];
                const float* __restrict__ h = H + (int64_t)col[p] * F;
                for (int64_t f = 0; f < F; ++f) acc[f] += v * (double)h[f];
            }
            float* __restrict__ o = out + r * F;
            for (int64_t f = 0; f < F; ++f) o[f] = (float)acc[f];
        }
        free(acc);
    }
}

/* K steps with the double-accumulating SpMM above (state kept in fp32 between steps, teleport as
 * filter.py:21 in fp32). */
void oracle_appnp_propagate_csr_omp_acc64_f32(const int64_t* row_ptr, const int32_t* col, const float* val,
                                              int64_t n, const float* H0, int64_t F, float a, int K,
                                              float* scratch, float* H_out) {
    if (K == 0) {
        memcpy(H_out, H0, (size_t)n * F * sizeof(float));
        return;
    }
    const float* src = H0;
    for (int k = 0; k < K; ++k) {
        float* dst = ((K - 1 - k) % 2 == 0) ? H_out : scratch;
        oracle_spmm_csr_omp_acc64_f32(row_ptr, col, val, src, F, 0, n, dst);
        oracle_teleport_f32(dst, H0, n * F, a, dst);
        src = dst;
    }
}
