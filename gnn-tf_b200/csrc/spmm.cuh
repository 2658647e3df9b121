// Epilogue descriptor shared by the SpMM kernels and the APPNP drivers.
#pragma once
#include "common.cuh"
#include "push.cuh"

namespace gnntf {

struct Epilogue {
    float s;             // scale on (A·B)[m]
    const float* H0;     // + t * H0[m]            (NULL = none)
    int64_t ldh;
    float t;
    const uint8_t* keep; // feature keep-mask, dense [n_rows, F] (NULL = none)
    float p_scale;
    int act;             // GNNTF_ACT_*
    float* C;            // output (NULL = not written)
    int64_t ldc;
    float* ACC;          // ACC[m] = (acc_init ? 0 : ACC[m]) + u*B[m] + w*out[m]   (NULL = none)
    int64_t ldacc;
    float u, w;
    int acc_init;
    const float* B;      // the dense operand (for the u*B[m] term)
    int64_t ldb;
    int F;
};

// `push` (sharded runs): the launch's leading CTAs also send rows of B to the peers; *push_done tells
// the caller whether that happened (only the float4 path can host it; else the caller pushes separately).
int spmm_dispatch(const gnntf_csr_t* A, const float* B, int64_t ldb, Epilogue epi, cudaStream_t st,
                  const PushArgs* push = nullptr, bool* push_done = nullptr);
int validate_csr(const gnntf_csr_t* A);
// K fused APPNP steps in ONE cooperative launch when the whole step is a single wave of CTAs and no
// row is split; *taken says whether it was enqueued (otherwise the caller launches step by step).
int spmm_persistent_propagate(const gnntf_csr_t* A, const float* H0, float* H_out, float* scratch, int64_t ld,
                              int64_t F, double alpha, int K, cudaStream_t st, bool* taken);

// The same K steps with the whole problem resident in the shared memory of one thread-block cluster
// (cluster.cu); cluster_size / threads 0 = choose.  *taken false: the shape does not qualify.
int appnp_cluster_propagate(const gnntf_csr_t* A, const float* H0, float* H_out, int64_t ld, int64_t F, double alpha,
                            int K, int cluster_size, int threads, cudaStream_t st, bool* taken);

}  // namespace gnntf
