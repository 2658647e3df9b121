// Fused APPNP personalised-PageRank steps and the K-step drivers (sm_100a).
//
// gnntf_appnp_step_f32        : PPRIteration.__forward__, gnntf/core/gnn/architectures/filter.py:19-22
//                               — SpMM, teleport axpy, feature dropout and activation in ONE pass
//                               (the reference issues an SpMM + 3 element-wise TF ops + dropout).
// gnntf_appnp_propagate*_f32  : the K PPRIteration layers of filter.py:34-35 as driven by
//                               Layered.__call__ (gnntf/core/nn/layered.py:52-55); ping-pong buffers,
//                               the last step writes the caller's output directly.
// gnntf_appnp_propagate_bwd_f32: VJP of that loop (what tape.gradient computes at
//                               gnntf/core/nn/trainable.py:78), without the dÂ branch TF evaluates
//                               and discards.
#include "spmm.cuh"


using namespace gnntf;

static int step_impl(const gnntf_csr_t* A, const float* H_in, const float* H0, float* H_out, int64_t ld,
                     int64_t F, double alpha, const uint8_t* feat_keep, float p_scale, int activation,
                     cudaStream_t st) {
    if (F < 0 || F > 0x7fffffff || ld < F) return GNNTF_E_SIZE;
    if (activation != GNNTF_ACT_IDENTITY && activation != GNNTF_ACT_RELU) return GNNTF_E_MODE;
    if (A != nullptr && A->n_rows > 0 && F > 0 && (H_in == nullptr || H0 == nullptr || H_out == nullptr))
        return GNNTF_E_NULL;
    Epilogue e{};
    e.s = (float)(1.0 - alpha);  // filter.py:21  propagated*(1-a)
    e.H0 = H0;
    e.ldh = ld;
    e.t = (float)alpha;          //               + H0.value*a
    e.keep = feat_keep;
    e.p_scale = p_scale;
    e.act = activation;
    e.C = H_out;
    e.ldc = ld;
    e.F = (int)F;
    return spmm_dispatch(A, H_in, ld, e, st);
}

extern "C" int gnntf_appnp_step_f32(const gnntf_csr_t* A, const float* H_in, const float* H0,
                                    float* H_out, int64_t ld, int64_t F, double alpha,
                                    const uint8_t* feat_keep, float p_scale, int activation,
                                    void* stream) {
    return step_impl(A, H_in, H0, H_out, ld, F, alpha, feat_keep, p_scale, activation,
                     (cudaStream_t)stream);
}

// Step k reads src_k and writes dst_k; dst_{K-1} = H_out and the buffers alternate backwards
// from there, so no final copy is needed.
static int propagate_impl(const gnntf_csr_t* A_k, int n_adj, int K, const float* H0, float* H_out,
                          float* scratch, int64_t ld, int64_t F, double alpha, cudaStream_t st) {
    if (K < 0) return GNNTF_E_SIZE;
    if (A_k == nullptr) return GNNTF_E_NULL;
    const int64_t n = A_k[0].n_rows;
    if (F < 0 || ld < F) return GNNTF_E_SIZE;
    if (n > 0 && F > 0 && (H0 == nullptr || H_out == nullptr)) return GNNTF_E_NULL;
    if (K == 0) {  // zero iterations: the stack is the identity on H0
        if (n > 0 && F > 0)
            GNNTF_CUDA_TRY(cudaMemcpy2DAsync(H_out, ld * sizeof(float), H0, ld * sizeof(float),
                                             F * sizeof(float), n, cudaMemcpyDeviceToDevice, st));
        return GNNTF_OK;
    }
    if (K > 1 && scratch == nullptr && n > 0 && F > 0) return GNNTF_E_NULL;
    if (n_adj == 1 && K > 1) {  // launch-bound shapes: all K steps in one cooperative launch
        // (the cluster-resident form, gnntf_appnp_propagate_cluster_f32, is NOT taken here: with the wide lane groups
        // the cooperative kernel runs Cora in 42.5 us, the cluster kernel in 43.1 us, and everything larger is
        // slower through DSMEM than through L2 — profiles/r2/17)
        int rc = validate_csr(&A_k[0]);
        if (rc != GNNTF_OK) return rc;
        bool taken = false;
        rc = spmm_persistent_propagate(&A_k[0], H0, H_out, scratch, ld, F, alpha, K, st, &taken);
        if (rc != GNNTF_OK || taken) return rc;
    }
    const float* src = H0;
    for (int k = 0; k < K; ++k) {
        float* dst = ((K - 1 - k) % 2 == 0) ? H_out : scratch;
        const gnntf_csr_t* A = &A_k[n_adj == 1 ? 0 : k];
        if (A->n_rows != n) return GNNTF_E_SIZE;
        int rc = step_impl(A, src, H0, dst, ld, F, alpha, nullptr, 1.0f, GNNTF_ACT_IDENTITY, st);
        if (rc != GNNTF_OK) return rc;
        src = dst;
    }
    return GNNTF_OK;
}

extern "C" int gnntf_appnp_propagate_f32(const gnntf_csr_t* A, const float* H0, float* H_out,
                                         float* scratch, int64_t ld, int64_t F, double alpha, int K,
                                         void* stream) {
    return propagate_impl(A, 1, K, H0, H_out, scratch, ld, F, alpha, (cudaStream_t)stream);
}

extern "C" int gnntf_appnp_propagate_multi_f32(const gnntf_csr_t* A_k, int K, const float* H0,
                                               float* H_out, float* scratch, int64_t ld, int64_t F,
                                               double alpha, void* stream) {
    return propagate_impl(A_k, K, K, H0, H_out, scratch, ld, F, alpha, (cudaStream_t)stream);
}

// g_K = dH_K.  For k = K-1..0:  dH0 += alpha*g_{k+1};  g_k = (1-alpha)*AT_k*g_{k+1}.  dH0 += g_0.
// Each step is one fused launch: the SpMM epilogue writes g_k and folds alpha*g_{k+1}[m] (the
// dense operand's own row) into dH0[m]; the last step also folds g_0 in and skips the g write.
extern "C" int gnntf_appnp_propagate_bwd_f32(const gnntf_csr_t* AT_k, int K, const float* dHK,
                                             float* dH0, float* scratch, int64_t ld, int64_t F,
                                             double alpha, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (K < 0) return GNNTF_E_SIZE;
    if (AT_k == nullptr) return GNNTF_E_NULL;
    const int64_t n = AT_k[0].n_rows;
    if (F < 0 || F > 0x7fffffff || ld < F) return GNNTF_E_SIZE;
    if (n == 0 || F == 0) return GNNTF_OK;
    if (dHK == nullptr || dH0 == nullptr) return GNNTF_E_NULL;
    if (K == 0) {
        GNNTF_CUDA_TRY(cudaMemcpy2DAsync(dH0, ld * sizeof(float), dHK, ld * sizeof(float),
                                         F * sizeof(float), n, cudaMemcpyDeviceToDevice, st));
        return GNNTF_OK;
    }
    if (K > 1 && scratch == nullptr) return GNNTF_E_NULL;
    float* buf[2] = {scratch, scratch ? scratch + n * ld : nullptr};
    const float* g = dHK;
    for (int k = K - 1; k >= 0; --k) {
        const gnntf_csr_t* AT = &AT_k[k];
        if (AT->n_rows != n) return GNNTF_E_SIZE;
        Epilogue e{};
        e.s = (float)(1.0 - alpha);
        e.act = GNNTF_ACT_IDENTITY;
        e.F = (int)F;
        e.ACC = dH0;
        e.ldacc = ld;
        e.u = (float)alpha;
        e.acc_init = (k == K - 1);
        if (k == 0) {
            e.C = nullptr;  // g_0 only feeds dH0
            e.w = 1.0f;
        } else {
            e.C = buf[k & 1];
            e.ldc = ld;
            e.w = 0.0f;
        }
        int rc = spmm_dispatch(AT, g, ld, e, st);
        if (rc != GNNTF_OK) return rc;
        g = e.C;
    }
    return GNNTF_OK;
}

// Host <-> device copies of a dense [n, F] host matrix and an [n, ld] device matrix.  Dense on both
// sides: one linear DMA (2-D copies are issued row by row).
static cudaError_t copy_in(float* dev, const float* host, int64_t n, int64_t ld, int64_t F, cudaStream_t st) {
    if (ld == F) return cudaMemcpyAsync(dev, host, (size_t)n * F * sizeof(float), cudaMemcpyHostToDevice, st);
    return cudaMemcpy2DAsync(dev, ld * sizeof(float), host, F * sizeof(float), F * sizeof(float), n,
                             cudaMemcpyHostToDevice, st);
}

static cudaError_t copy_out(float* host, const float* dev, int64_t n, int64_t ld, int64_t F, cudaStream_t st) {
    if (ld == F) return cudaMemcpyAsync(host, dev, (size_t)n * F * sizeof(float), cudaMemcpyDeviceToHost, st);
    return cudaMemcpy2DAsync(host, F * sizeof(float), dev, ld * sizeof(float), F * sizeof(float), n,
                             cudaMemcpyDeviceToHost, st);
}

extern "C" int gnntf_appnp_propagate_host_f32(const gnntf_csr_t* A, const float* H0_host,
                                              float* out_host, float* dev_H0, float* dev_out,
                                              float* dev_scratch, int64_t ld, int64_t F, double alpha,
                                              int K, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (A == nullptr) return GNNTF_E_NULL;
    const int64_t n = A->n_rows;
    if (F < 0 || ld < F) return GNNTF_E_SIZE;
    if (n == 0 || F == 0) return GNNTF_OK;
    if (H0_host == nullptr || out_host == nullptr || dev_H0 == nullptr || dev_out == nullptr)
        return GNNTF_E_NULL;
    GNNTF_CUDA_TRY(copy_in(dev_H0, H0_host, n, ld, F, st));
    int rc = propagate_impl(A, 1, K, dev_H0, dev_out, dev_scratch, ld, F, alpha, st);
    if (rc != GNNTF_OK) return rc;
    GNNTF_CUDA_TRY(copy_out(out_host, dev_out, n, ld, F, st));
    return GNNTF_OK;
}

// Several feature matrices through the same adjacency, software-pipelined over three streams: while the
// K steps of matrix b run on the caller's stream, matrix b+1 travels host -> device on a copy stream and
// the result of matrix b-1 travels device -> host on another (PCIe is full duplex, and the copy engines
// do not take SMs).  Two device slots {H0, out} alternate; the scratch is shared (the K-step chains are
// serialised on the caller's stream).  Ordering per slot s = b & 1:
//   H2D(b)     after compute(b-2)   (the slot's H0 is free)
//   compute(b) after H2D(b) and D2H(b-2)  (the slot's out is free)
//   D2H(b)     after compute(b)
// The caller's stream finally waits for the last two read-backs, so synchronising it is enough.
namespace {
struct HostPipeline {
    cudaStream_t up = nullptr, down = nullptr;
    cudaEvent_t start = nullptr, in[2] = {nullptr, nullptr}, comp[2] = {nullptr, nullptr}, out[2] = {nullptr, nullptr};
    cudaError_t create() {
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking)) != cudaSuccess) return e;
        cudaEvent_t* all[] = {&start, &in[0], &in[1], &comp[0], &comp[1], &out[0], &out[1]};
        for (cudaEvent_t* ev : all)
            if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    ~HostPipeline() {  // destroying a stream / event with work in flight is deferred by the runtime
        cudaEvent_t all[] = {start, in[0], in[1], comp[0], comp[1], out[0], out[1]};
        for (cudaEvent_t ev : all)
            if (ev) cudaEventDestroy(ev);
        if (up) cudaStreamDestroy(up);
        if (down) cudaStreamDestroy(down);
    }
};
}  // namespace

extern "C" int gnntf_appnp_propagate_host_batched_f32(const gnntf_csr_t* A, const float* const* H0_host,
                                                      float* const* out_host, int n_batches, float* dev_work,
                                                      int64_t ld, int64_t F, double alpha, int K, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (A == nullptr) return GNNTF_E_NULL;
    const int64_t n = A->n_rows;
    if (F < 0 || ld < F || n_batches < 0) return GNNTF_E_SIZE;
    if (n == 0 || F == 0 || n_batches == 0) return GNNTF_OK;
    if (H0_host == nullptr || out_host == nullptr || dev_work == nullptr) return GNNTF_E_NULL;
    for (int b = 0; b < n_batches; ++b)
        if (H0_host[b] == nullptr || out_host[b] == nullptr) return GNNTF_E_NULL;
    const size_t mat = (size_t)n * (size_t)ld;
    float* dev_H0[2] = {dev_work, dev_work + 2 * mat};
    float* dev_out[2] = {dev_work + mat, dev_work + 3 * mat};
    float* dev_scratch = dev_work + 4 * mat;
    HostPipeline p;
    GNNTF_CUDA_TRY(p.create());
    GNNTF_CUDA_TRY(cudaEventRecord(p.start, st));  // the copies start after what the caller enqueued before
    GNNTF_CUDA_TRY(cudaStreamWaitEvent(p.up, p.start, 0));
    GNNTF_CUDA_TRY(cudaStreamWaitEvent(p.down, p.start, 0));
    for (int b = 0; b < n_batches; ++b) {
        const int s = b & 1;
        if (b >= 2) GNNTF_CUDA_TRY(cudaStreamWaitEvent(p.up, p.comp[s], 0));
        GNNTF_CUDA_TRY(copy_in(dev_H0[s], H0_host[b], n, ld, F, p.up));
        GNNTF_CUDA_TRY(cudaEventRecord(p.in[s], p.up));
        GNNTF_CUDA_TRY(cudaStreamWaitEvent(st, p.in[s], 0));
        if (b >= 2) GNNTF_CUDA_TRY(cudaStreamWaitEvent(st, p.out[s], 0));
        int rc = propagate_impl(A, 1, K, dev_H0[s], dev_out[s], dev_scratch, ld, F, alpha, st);
        if (rc != GNNTF_OK) return rc;
        GNNTF_CUDA_TRY(cudaEventRecord(p.comp[s], st));
        GNNTF_CUDA_TRY(cudaStreamWaitEvent(p.down, p.comp[s], 0));
        GNNTF_CUDA_TRY(copy_out(out_host[b], dev_out[s], n, ld, F, p.down));
        GNNTF_CUDA_TRY(cudaEventRecord(p.out[s], p.down));
    }
    GNNTF_CUDA_TRY(cudaStreamWaitEvent(st, p.out[(n_batches - 1) & 1], 0));
    if (n_batches >= 2) GNNTF_CUDA_TRY(cudaStreamWaitEvent(st, p.out[n_batches & 1], 0));
    return GNNTF_OK;
}

extern "C" int gnntf_abi_version(void) { return GNNTF_ABI_VERSION; }

extern "C" const char* gnntf_status_str(int code) {
    switch (code) {
        case GNNTF_OK: return "ok";
        case GNNTF_E_NULL: return "a required pointer is NULL";
        case GNNTF_E_SIZE: return "invalid size (negative, nnz >= 2^31, or leading dimension < F)";
        case GNNTF_E_MODE: return "Invalid matrix normalization";
        case GNNTF_E_WORKSPACE: return "workspace too small";
        case GNNTF_E_ALIGN: return "pointer misaligned (coo_indices must be 16-byte aligned)";
        case GNNTF_E_SHAPE: return "shape does not qualify for this specialised entry point";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown gnntf status";
}
