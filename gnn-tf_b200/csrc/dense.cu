// The element-wise stages on either side of the propagation path, fused:
//
// gnntf_bias_act_dropout_f32 (+ _bwd)  the tail of Dense.__forward__ (gnntf/core/nn/layers.py:135-136)
//                                      and of GCNLayer.__forward__ (gnntf/core/gnn/architectures/gcn.py:89):
//                                      dropout(activation(Z + b)) in ONE pass over Z (TF issues a bias add,
//                                      an activation, a RandomUniform, a compare, a scale and a multiply);
// gnntf_node_xent_f32 (+ _bwd)         NodeClassification.loss (gnntf/core/gnn/graph_predictor.py:19-25):
//                                      embedding_lookup + log_softmax + SparseCategoricalCrossentropy
//                                      (from_logits) over the selected nodes, mean-reduced, one kernel.
//
// The GEMMs themselves (X·W) stay on cuBLAS through the host framework: a plain library GEMM.
#include <algorithm>

#include "common.cuh"

namespace gnntf {

__device__ __forceinline__ float act_apply(float x, int act, float slope) {
    if (act == GNNTF_ACT_RELU) return fmaxf(x, 0.0f);
    if (act == GNNTF_ACT_LEAKY_RELU) return x > 0.0f ? x : x * slope;
    return x;
}

// out = keep ? act(Z + b) * p_scale : 0      (keep == NULL: no dropout)
template <int VEC>
__global__ void bias_act_dropout_kernel(const float* __restrict__ Z, int64_t ldz, const float* __restrict__ bias,
                                        const uint8_t* __restrict__ keep, float p_scale, int act, float slope,
                                        float* __restrict__ out, int64_t ldo, int64_t n, int F) {
    const int fv = F / VEC;
    const int64_t total = n * fv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / fv;
        const int f = (int)(i - r * fv) * VEC;
        Vec<VEC> z = Vec<VEC>::stream(Z + r * ldz + f);
        Vec<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float x = z.v[k] + (bias ? __ldg(bias + f + k) : 0.0f);
            x = act_apply(x, act, slope);
            if (keep) x = keep[r * (int64_t)F + f + k] ? x * p_scale : 0.0f;
            o.v[k] = x;
        }
        o.store(out + r * ldo + f);
    }
}

// dZ = g * keep * p_scale * act'(pre)   with act' evaluated from the forward OUTPUT y:
//   relu:       y > 0          (dropped elements have y == 0 and keep == 0: either test zeroes them)
//   leaky relu: y > 0 ? 1 : slope   (p_scale > 0 keeps the sign of the pre-activation in y where kept)
template <int VEC>
__global__ void bias_act_dropout_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ y,
                                            int64_t ldy, const uint8_t* __restrict__ keep, float p_scale, int act,
                                            float slope, float* __restrict__ dZ, int64_t ldd, int64_t n, int F) {
    const int fv = F / VEC;
    const int64_t total = n * fv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / fv;
        const int f = (int)(i - r * fv) * VEC;
        Vec<VEC> gv = Vec<VEC>::stream(g + r * ldg + f);
        Vec<VEC> yv;
#pragma unroll
        for (int k = 0; k < VEC; ++k) yv.v[k] = 1.0f;
        if (act != GNNTF_ACT_IDENTITY) yv = Vec<VEC>::stream(y + r * ldy + f);
        Vec<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float d = gv.v[k];
            if (keep) d = keep[r * (int64_t)F + f + k] ? d * p_scale : 0.0f;
            if (act == GNNTF_ACT_RELU) d = yv.v[k] > 0.0f ? d : 0.0f;
            if (act == GNNTF_ACT_LEAKY_RELU) d = yv.v[k] > 0.0f ? d : d * slope;
            o.v[k] = d;
        }
        o.store(dZ + r * ldd + f);
    }
}

// One warp per selected node i: row = logits[nodes[i], 0:C];  loss_i = logsumexp(row) - row[label_i].
// The per-node losses are written out; a second tiny kernel sums them in a fixed order (deterministic).
__global__ void node_xent_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ nodes,
                                 const int64_t* __restrict__ labels, int64_t m, int C, float* __restrict__ per_node) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < m; i += nwarps) {
        const float* row = logits + nodes[i] * ld;
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.0f;
        for (int c = lane; c < C; c += 32) sum += expf(row[c] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) per_node[i] = (logf(sum) + mx) - row[labels[i]];
    }
}

__global__ void mean_reduce_kernel(const float* __restrict__ x, int64_t m, float* __restrict__ out) {
    __shared__ float part[32];
    float acc = 0.0f;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) acc += x[i];  // fixed assignment, fixed tree: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) out[0] = acc / (float)m;
    }
}

// dlogits[nodes[i], c] += gscale * (softmax(row)_c - [c == label_i]) / m      (dlogits zeroed by the caller)
__global__ void node_xent_bwd_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ nodes,
                                     const int64_t* __restrict__ labels, int64_t m, int C, const float* __restrict__ gscale,
                                     float* __restrict__ dlogits, int64_t ldd) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float gs = gscale[0] / (float)m;
    for (int64_t i = warp; i < m; i += nwarps) {
        const int64_t node = nodes[i];
        const float* row = logits + node * ld;
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.0f;
        for (int c = lane; c < C; c += 32) sum += expf(row[c] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        const int lab = (int)labels[i];
        for (int c = lane; c < C; c += 32) {
            const float p = expf(row[c] - mx) * inv;
            atomicAdd(dlogits + node * ldd + c, gs * (p - (c == lab ? 1.0f : 0.0f)));  // one add per element unless a node repeats
        }
    }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace gnntf

using namespace gnntf;

extern "C" int gnntf_bias_act_dropout_f32(const float* Z, int64_t ldz, const float* bias, const uint8_t* keep,
                                          float p_scale, int activation, float slope, float* out, int64_t ldo,
                                          int64_t n, int64_t F, void* stream) {
    if (n < 0 || F < 0 || F > 0x7fffffff || ldz < F || ldo < F) return GNNTF_E_SIZE;
    if (activation < GNNTF_ACT_IDENTITY || activation > GNNTF_ACT_LEAKY_RELU) return GNNTF_E_MODE;
    if (n == 0 || F == 0) return GNNTF_OK;
    if (Z == nullptr || out == nullptr) return GNNTF_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = (F % 4 == 0) && (ldz % 4 == 0) && (ldo % 4 == 0) && al16(Z) && al16(out);
    const int64_t work = n * (v4 ? F / 4 : F);
    const int grid = (int)std::min<int64_t>(ceil_div(work, 256), (int64_t)kNumSMs * 8);
    if (v4)
        bias_act_dropout_kernel<4><<<grid, 256, 0, st>>>(Z, ldz, bias, keep, p_scale, activation, slope, out, ldo, n, (int)F);
    else
        bias_act_dropout_kernel<1><<<grid, 256, 0, st>>>(Z, ldz, bias, keep, p_scale, activation, slope, out, ldo, n, (int)F);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_bias_act_dropout_bwd_f32(const float* g, int64_t ldg, const float* y, int64_t ldy,
                                              const uint8_t* keep, float p_scale, int activation, float slope,
                                              float* dZ, int64_t ldd, int64_t n, int64_t F, void* stream) {
    if (n < 0 || F < 0 || F > 0x7fffffff || ldg < F || ldd < F) return GNNTF_E_SIZE;
    if (activation < GNNTF_ACT_IDENTITY || activation > GNNTF_ACT_LEAKY_RELU) return GNNTF_E_MODE;
    if (n == 0 || F == 0) return GNNTF_OK;
    if (g == nullptr || dZ == nullptr) return GNNTF_E_NULL;
    if (activation != GNNTF_ACT_IDENTITY && (y == nullptr || ldy < F)) return GNNTF_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = (F % 4 == 0) && (ldg % 4 == 0) && (ldd % 4 == 0) && al16(g) && al16(dZ) &&
                    (activation == GNNTF_ACT_IDENTITY || (ldy % 4 == 0 && al16(y)));
    const int64_t work = n * (v4 ? F / 4 : F);
    const int grid = (int)std::min<int64_t>(ceil_div(work, 256), (int64_t)kNumSMs * 8);
    if (v4)
        bias_act_dropout_bwd_kernel<4><<<grid, 256, 0, st>>>(g, ldg, y, ldy, keep, p_scale, activation, slope, dZ, ldd, n, (int)F);
    else
        bias_act_dropout_bwd_kernel<1><<<grid, 256, 0, st>>>(g, ldg, y, ldy, keep, p_scale, activation, slope, dZ, ldd, n, (int)F);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_node_xent_f32(const float* logits, int64_t ld, const int64_t* nodes, const int64_t* labels,
                                   int64_t m, int64_t C, float* per_node_ws, float* loss, void* stream) {
    if (m < 0 || C <= 0 || C > 0x7fffffff || ld < C) return GNNTF_E_SIZE;
    if (loss == nullptr) return GNNTF_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) {
        GNNTF_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
        return GNNTF_OK;
    }
    if (logits == nullptr || nodes == nullptr || labels == nullptr || per_node_ws == nullptr) return GNNTF_E_NULL;
    const int grid = (int)std::min<int64_t>(ceil_div(m, 8), (int64_t)kNumSMs * 8);
    node_xent_kernel<<<grid, 256, 0, st>>>(logits, ld, nodes, labels, m, (int)C, per_node_ws);
    GNNTF_LAUNCH_CHECK();
    mean_reduce_kernel<<<1, 1024, 0, st>>>(per_node_ws, m, loss);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_node_xent_bwd_f32(const float* logits, int64_t ld, const int64_t* nodes, const int64_t* labels,
                                       int64_t m, int64_t C, const float* grad_loss, float* dlogits, int64_t ldd,
                                       void* stream) {
    if (m < 0 || C <= 0 || C > 0x7fffffff || ld < C || ldd < C) return GNNTF_E_SIZE;
    if (m == 0) return GNNTF_OK;
    if (logits == nullptr || nodes == nullptr || labels == nullptr || grad_loss == nullptr || dlogits == nullptr)
        return GNNTF_E_NULL;
    const int grid = (int)std::min<int64_t>(ceil_div(m, 8), (int64_t)kNumSMs * 8);
    node_xent_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, ld, nodes, labels, m, (int)C, grad_loss, dlogits, ldd);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}
