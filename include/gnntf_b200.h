/* gnntf_b200.h — C-ABI of the B200-native sparse adjacency propagation path of gnntf.
 *
 * The reference (MKLab-ITI/gnn-tf) has no FFI of its own: its hot path is a chain of stock
 * TensorFlow ops issued from Python.  Each entry point below replaces one of those call sites
 * (file:line relative to the reference root) and is what a TF custom op (tf_op/gnntf_ops.cc,
 * loaded with tf.load_op_library) or the ctypes host shim (gnntf/_native.py) binds:
 *
 *   gnntf_csr_build*        gnntf/core/gnn/graph_manipulation.py:24-31  graph2adj (symmetrise by
 *                           appending the reversed list, duplicates kept) -> COO .indices/.values
 *                           + the stable-by-row CSR view the kernels consume
 *   gnntf_normalize_f32     gnntf/core/nn/layered.py:47-50   sparse_dropout  (explicit keep-mask)
 *                           gnntf/core/gnn/gnn.py:36-50      get_adjacency   (colsum, D, row*col scale)
 *   gnntf_spmm_f32          tf.sparse.sparse_dense_matmul at filter.py:19, gcn.py:24,48,88,104,131
 *                           (and its adjoint_a=True gradient, trainable.py:78)
 *   gnntf_appnp_step_f32    gnntf/core/gnn/architectures/filter.py:19-22  SpMM + teleport axpy
 *                           (+ feature dropout + relu) in one pass
 *   gnntf_appnp_propagate*  the K PPRIteration layers, filter.py:34-35 under layered.py:52-55
 *   gnntf_appnp_propagate_bwd_f32   their VJP (tape.gradient, trainable.py:78)
 *
 * Conventions.  Every pointer is a DEVICE pointer unless its name ends in _host.  The library
 * is stateless and re-entrant: the caller owns every buffer, workspace sizes are queried first,
 * nothing is allocated and nothing synchronises (all work is enqueued on `stream`, a
 * cudaStream_t passed as void*; 0 = legacy default stream).  Return value: 0 = OK, < 0 = one of
 * the GNNTF_E_* argument errors, > 0 = a cudaError_t.  No exception crosses the ABI.
 * Index widths: caller-facing COO indices are int64 (tf.SparseTensor.indices); the internal CSR
 * is int32, so nnz must be < 2^31 (the largest BASELINE config, R-MAT 1e9 edges, is 2.0e9).
 * Dense matrices are row-major fp32 with an explicit leading dimension (in floats).
 */
#ifndef GNNTF_B200_H
#define GNNTF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNNTF_ABI_VERSION 1

#define GNNTF_OK 0
#define GNNTF_E_NULL (-1)      /* a required pointer is NULL */
#define GNNTF_E_SIZE (-2)      /* negative size, nnz >= 2^31, or ld < F */
#define GNNTF_E_MODE (-3)      /* "Invalid matrix normalization" (gnn.py:46-47) or bad enum */
#define GNNTF_E_WORKSPACE (-4) /* workspace too small */
#define GNNTF_E_ALIGN (-5)     /* pointer misaligned (coo_indices of gnntf_csr_build: 16 bytes) */
#define GNNTF_E_SHAPE (-6)     /* the shape does not qualify for this specialised entry (use the general one) */

/* normalized= of GNN.get_adjacency (gnn.py:36) */
#define GNNTF_NORM_SYMMETRIC 0 /* D = 1/sqrt(colsum); v*D[row]*D[col]   gnn.py:40-42 */
#define GNNTF_NORM_BIPARTITE 1 /* D = 1/colsum;       v*D[row]          gnn.py:43-45 */
#define GNNTF_NORM_NONE 2

/* add_eye= of GNN.get_adjacency (gnn.py:38-39, 48-49) */
#define GNNTF_EYE_NONE 0
#define GNNTF_EYE_BEFORE 1
#define GNNTF_EYE_AFTER 2

/* activation fused in the step epilogue (PPRIteration activation, filter.py:22) */
#define GNNTF_ACT_IDENTITY 0
#define GNNTF_ACT_RELU 1
#define GNNTF_ACT_LEAKY_RELU 2 /* gnntf_bias_act_dropout* only (NGCFLayer, gcn.py:117) */

int gnntf_abi_version(void);
const char* gnntf_status_str(int code);

/* ------------------------------------------------------------------------------------------
 * CSR view of a (normalised) adjacency.  Built by gnntf_csr_build + gnntf_spmm_plan_*.
 * Rows with more than `long_threshold` entries are split into `chunk`-sized pieces that run on
 * separate warps and are reduced in a fixed order (deterministic, no float atomics).
 * n_long == 0 means no row is split (then the plan pointers may be NULL).
 * ---------------------------------------------------------------------------------------- */
typedef struct gnntf_csr {
    int64_t n_rows;
    int64_t nnz;
    const int32_t* row_ptr; /* [n_rows+1] */
    const int32_t* col_idx; /* [nnz]      */
    const float* val;       /* [nnz]      */
    const int32_t* row_map; /* [n_rows] dense-matrix row that CSR row i reads its teleport term
                               from and writes to, or NULL = identity.  Lets a row SUBSET (interior /
                               boundary rows of a shard) run as its own compact CSR. */
    int32_t long_threshold; /* rows with deg > long_threshold are split; <= 0 disables */
    int32_t chunk;          /* entries per piece */
    int32_t n_long;
    int32_t n_chunks;
    const int32_t* long_row;         /* [n_long]   row index                       */
    const int32_t* long_first_chunk; /* [n_long]   first piece of that row         */
    const int32_t* long_n_chunks;    /* [n_long]   number of pieces                */
    const int32_t* chunk_row;        /* [n_chunks] row index of the piece          */
    const int32_t* chunk_begin;      /* [n_chunks] first CSR slot of the piece     */
    float* partials;                 /* [n_chunks * round_up(F,4)] workspace, or NULL if n_long==0 */
} gnntf_csr_t;

/* ------------------------------------------------------------------------------------------
 * (a) graph2adj -> COO + CSR                                   graph_manipulation.py:24-31
 * edges: int64 [n_edges,2] in graph2indices order (:19-21); weights fp32 [n_edges] or NULL (=1., :27).
 * nnz = (directed ? 1 : 2)*n_edges + (add_eye ? n : 0).  Slot q of the COO list is
 *   q <  E          : (u_q, v_q, w_q)
 *   E <= q < 2E     : (v_{q-E}, u_{q-E}, w_{q-E})      [undirected only, :28-30]
 *   then n diagonal entries (i,i,1.)                    [add_eye != 0; tf.sparse.eye, gnn.py:39,49]
 * Outputs: coo_indices int64 [nnz,2] (16-byte aligned, else GNNTF_E_ALIGN) / coo_values fp32 [nnz]
 * (the SparseTensor fields; either may be NULL), and the CSR obtained by a STABLE sort of that list by row: row_ptr int32 [n+1],
 * col_idx int32 [nnz], raw_val fp32 [nnz], coo_pos int32 [nnz] (COO slot of each CSR slot).
 * Indices outside [0,n) are not detected here (the host shim validates).
 * ---------------------------------------------------------------------------------------- */
int gnntf_csr_build_ws_bytes(int64_t n, int64_t n_edges, int directed, int add_eye, size_t* bytes);
int gnntf_csr_build(const int64_t* edges, const float* weights, int64_t n, int64_t n_edges,
                    int directed, int add_eye, int by_column,
                    int64_t* coo_indices, float* coo_values,
                    int32_t* row_ptr, int32_t* col_idx, float* raw_val, int32_t* coo_pos,
                    void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * sparse_dropout + get_adjacency on the CSR view            layered.py:47-50, gnn.py:36-50
 * keep_mask_coo: uint8 [n_graph] in COO order (1 = keep) or NULL (eval mode / p == 0);
 * scale = fp32(1/(1-p)).  n_graph = nnz without the eye entries; eye_mode says how the trailing
 * n eye entries (if the CSR was built with add_eye) take part (BEFORE: normalised like any
 * entry; AFTER: excluded from the degree, value stays 1).  directed != 0 means the COO list has
 * no appended reverse half, so column sums use float atomics instead of the partner trick.
 * Outputs: deg [n], dinv [n] (0 where deg == 0, divide_no_nan), norm_val [nnz] (CSR order) or NULL,
 * norm_val_T [nnz] or NULL (values of the TRANSPOSED matrix laid on the same CSR structure —
 * valid for undirected graphs only; equals norm_val when no mask is given),
 * norm_val_coo [nnz] or NULL (COO order: the .values of the SparseTensor get_adjacency returns).
 * ---------------------------------------------------------------------------------------- */
int gnntf_normalize_f32(const int32_t* row_ptr, const int32_t* col_idx, const float* raw_val,
                        const int32_t* coo_pos, int64_t n, int64_t nnz, int64_t n_graph,
                        int directed, const uint8_t* keep_mask_coo, float scale, int mode,
                        int eye_mode, float* deg, float* dinv, float* norm_val, float* norm_val_T,
                        float* norm_val_coo, void* stream);

/* ------------------------------------------------------------------------------------------
 * Locality-restoring node order (builder-side; external indices are never changed).
 * One sweep of a robust 1-D circular arrangement: every node moves to the weighted circular mean
 * of its neighbours' positions, weights 1/(|offset| + eps) (iteratively re-weighted least
 * absolute deviation: near neighbours dominate, far ones are ignored):
 *   theta_out[i] = theta_in[i] + sum_j w_ij * wrap(theta_in[j] - theta_in[i]) / sum_j w_ij   (mod 2*pi)
 * theta in [0, 2*pi).  The host re-ranks the angles between sweeps and seeds them with the angle
 * of the two leading non-trivial eigenvectors of the normalised adjacency (gnntf/reorder.py).
 * ---------------------------------------------------------------------------------------- */
int gnntf_arrange_sweep_f32(const int32_t* row_ptr, const int32_t* col_idx, const float* theta_in,
                            float eps, float* theta_out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Long-row plan for the SpMM kernels (two calls so the caller can size the arrays).
 * counts: int32 [2] device = {n_long, n_chunks}.
 * ---------------------------------------------------------------------------------------- */
int gnntf_spmm_plan_count(const int32_t* row_ptr, int64_t n_rows, int32_t long_threshold,
                          int32_t chunk, int32_t* counts, void* stream);
int gnntf_spmm_plan_fill(const int32_t* row_ptr, int64_t n_rows, int32_t long_threshold,
                         int32_t chunk, int32_t* counters_ws /* int32[2], zeroed by the call */,
                         int32_t* long_row, int32_t* long_first_chunk, int32_t* long_n_chunks,
                         int32_t* chunk_row, int32_t* chunk_begin, void* stream);

/* ------------------------------------------------------------------------------------------
 * (b) SpMM   C[n_rows,F] = A · B                  tf.sparse.sparse_dense_matmul (filter.py:19, gcn.py:88)
 * The backward dB = Aᵀ·dC is the same call on the transposed values (gnntf_normalize_f32's
 * norm_val_T; identical to A when no edge dropout was applied).  B and C must not alias.
 * ---------------------------------------------------------------------------------------- */
int gnntf_spmm_f32(const gnntf_csr_t* A, const float* B, int64_t ldb, float* C, int64_t ldc,
                   int64_t F, void* stream);
/* C[row_map[i],:] += scale · (A·B)[i,:] — the second pass of a sharded step: the entries whose
 * columns are halo rows are accumulated onto the result of the owned-column pass once the halo has
 * arrived.  Rows are the CSR's own (row_map selects the output rows); B and C must not alias. */
int gnntf_spmm_acc_f32(const gnntf_csr_t* A, const float* B, int64_t ldb, float* C, int64_t ldc,
                       int64_t F, double scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * (c) one PPRIteration, fused                                          filter.py:19-22
 *   H_out = act( keep ∘ p_scale ∘ ( (1-alpha)·(A·H_in) + alpha·H0 ) )
 * feat_keep: uint8 [n_rows, F] (dense, ld = F) or NULL (p_feat == 0, the APPNP default, filter.py:8).
 * ---------------------------------------------------------------------------------------- */
int gnntf_appnp_step_f32(const gnntf_csr_t* A, const float* H_in, const float* H0, float* H_out,
                         int64_t ld, int64_t F, double alpha, const uint8_t* feat_keep,
                         float p_scale, int activation, void* stream);

/* K fused steps, eval-mode semantics (one adjacency for all steps): H_1 = step(H0) … H_K -> H_out.
 * scratch: n_rows*ld floats (ping-pong partner of H_out).  H0, H_out, scratch must be distinct. */
int gnntf_appnp_propagate_f32(const gnntf_csr_t* A, const float* H0, float* H_out, float* scratch,
                              int64_t ld, int64_t F, double alpha, int K, void* stream);

/* The same K steps inside ONE thread-block-cluster launch: the cluster's CTAs keep their rows of the two
 * ping-pong feature matrices, of H0 and of the CSR in shared memory for the whole call, gather neighbour
 * rows through distributed shared memory and meet at a hardware cluster barrier between steps — global
 * memory is read once and written once.  For graphs the size of Cora / PubMed (n_rows*F*12 bytes plus the
 * CSR must fit the cluster's shared memory, F <= 128, F % 4 == 0, ld % 4 == 0, no split rows, 16-byte
 * aligned bases); results are bit-identical to gnntf_appnp_propagate_f32.  An explicit alternative, not
 * a default: measured no faster than the cooperative launch the general entry uses.  cluster_size: 0 = choose, else 1, 2, 4, 8 or 16 CTAs; threads: 0 = choose,
 * else 512 or 1024 per CTA.
 * Returns GNNTF_E_SHAPE when the shape does not qualify (nothing was enqueued).  No scratch needed. */
int gnntf_appnp_propagate_cluster_f32(const gnntf_csr_t* A, const float* H0, float* H_out, int64_t ld,
                                      int64_t F, double alpha, int K, int cluster_size, int threads, void* stream);

/* Training-mode forward: step k uses its own adjacency A_k[k] (own edge mask, filter.py:18). */
int gnntf_appnp_propagate_multi_f32(const gnntf_csr_t* A_k, int K, const float* H0, float* H_out,
                                    float* scratch, int64_t ld, int64_t F, double alpha, void* stream);

/* VJP of the K-step loop (p_feat = 0, act = identity): given dH_K returns dH0.
 * AT_k[k] = TRANSPOSED adjacency of step k (pass the same struct K times, or K == n_adj == 1
 * semantics via gnntf_appnp_propagate_f32 on the transposed matrix, when all steps share one).
 * For k = K-1..0:  dH0 += alpha·g;  g = (1-alpha)·AT_k[k]·g;  finally dH0 += g.
 * scratch: 2*n_rows*ld floats.  dHK, dH0, scratch distinct. */
int gnntf_appnp_propagate_bwd_f32(const gnntf_csr_t* AT_k, int K, const float* dHK, float* dH0,
                                  float* scratch, int64_t ld, int64_t F, double alpha, void* stream);

/* Same as gnntf_appnp_propagate_f32 but with HOST feature buffers: H0_host (n_rows*F floats,
 * dense, pinned for full speed) is copied to dev_H0, the K steps run, and dev_out is copied back
 * to out_host — all enqueued on `stream` (the caller synchronises).  dev_* are device buffers of
 * n_rows*ld floats each.  This is the end-to-end entry the benchmark's `e2e` leg times. */
int gnntf_appnp_propagate_host_f32(const gnntf_csr_t* A, const float* H0_host, float* out_host,
                                   float* dev_H0, float* dev_out, float* dev_scratch, int64_t ld,
                                   int64_t F, double alpha, int K, void* stream);

/* n_batches feature matrices through the same adjacency, HOST buffers on both sides, software-pipelined:
 * while the K steps of matrix b run on `stream`, matrix b+1 is copied host -> device and the result of
 * matrix b-1 device -> host on two internal copy streams (PCIe is full duplex).  H0_host / out_host: HOST
 * arrays of n_batches host pointers (n_rows*F floats each, dense, pinned for full speed; the same pointer
 * may appear several times in H0_host, and in out_host if the caller does not need every result).
 * dev_work: 5*n_rows*ld floats of device memory (two {H0, out} slots and one scratch).  On return `stream`
 * waits for the last read-back: synchronising it is enough.  Results are those of n_batches calls of
 * gnntf_appnp_propagate_host_f32 (the serving form of Layered.__call__, gnntf/core/nn/layered.py:52-55,
 * applied to a sequence of inputs). */
int gnntf_appnp_propagate_host_batched_f32(const gnntf_csr_t* A, const float* const* H0_host,
                                           float* const* out_host, int n_batches, float* dev_work,
                                           int64_t ld, int64_t F, double alpha, int K, void* stream);

/* ------------------------------------------------------------------------------------------
 * The element-wise stages on either side of the path, fused (callers of the propagation).
 *
 * gnntf_bias_act_dropout_f32: out = keep ∘ p_scale ∘ act(Z + bias) — the tail of Dense.__forward__
 * (gnntf/core/nn/layers.py:135-136) and of GCNLayer.__forward__ (gcn.py:89) after the GEMM, one pass.
 * bias: [F] or NULL; keep: uint8 [n,F] dense or NULL; slope: leaky-relu negative slope.  out may alias Z.
 * gnntf_bias_act_dropout_bwd_f32: dZ = g ∘ keep ∘ p_scale ∘ act'(·) with act' taken from the forward
 * OUTPUT y (not needed for GNNTF_ACT_IDENTITY); the bias gradient is the column sum of dZ (caller).
 *
 * gnntf_node_xent_f32: NodeClassification.loss (gnntf/core/gnn/graph_predictor.py:19-25) —
 * mean_i [ logsumexp(logits[nodes[i],:]) − logits[nodes[i], labels[i]] ]; tf.nn.embedding_lookup,
 * log_softmax and SparseCategoricalCrossentropy(from_logits=True) in one kernel (log_softmax is
 * idempotent, so applying the loss to log-probabilities as the reference does gives the same value).
 * per_node_ws: m floats of workspace; loss: 1 float.  The reduction order is fixed (deterministic).
 * gnntf_node_xent_bwd_f32: dlogits[nodes[i],:] += grad_loss·(softmax − onehot)/m into a caller-zeroed
 * [N,C] matrix (float atomics: bitwise reproducible when no node is listed twice).
 * ---------------------------------------------------------------------------------------- */
int gnntf_bias_act_dropout_f32(const float* Z, int64_t ldz, const float* bias, const uint8_t* keep,
                               float p_scale, int activation, float slope, float* out, int64_t ldo,
                               int64_t n, int64_t F, void* stream);
int gnntf_bias_act_dropout_bwd_f32(const float* g, int64_t ldg, const float* y, int64_t ldy,
                                   const uint8_t* keep, float p_scale, int activation, float slope,
                                   float* dZ, int64_t ldd, int64_t n, int64_t F, void* stream);
int gnntf_node_xent_f32(const float* logits, int64_t ld, const int64_t* nodes, const int64_t* labels,
                        int64_t m, int64_t C, float* per_node_ws, float* loss, void* stream);
int gnntf_node_xent_bwd_f32(const float* logits, int64_t ld, const int64_t* nodes, const int64_t* labels,
                            int64_t m, int64_t C, const float* grad_loss, float* dlogits, int64_t ldd,
                            void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-sharded multi-GPU helper (contiguous node-range split; BASELINE north_star).
 * Rank r owns rows [lo,hi).  Its local CSR addresses an EXTENDED feature matrix
 *   H_ext = [ n_local owned rows | n_halo rows received from peers ]
 * (the column remapping and the per-peer send lists are host-side index logic, gnntf/dist.py).
 * gnntf_halo_pack_f32 gathers the rows the peers need (send_idx: local row ids, grouped by
 * destination rank) into the dense send buffer of the NCCL all-to-all: out[i,:] = H[send_idx[i],:].
 * ---------------------------------------------------------------------------------------- */
int gnntf_halo_pack_f32(const float* H, int64_t ld, const int32_t* send_idx, int64_t n_send,
                        float* out, int64_t ldo, int64_t F, void* stream);
/* Fused pack + send over NVLink peer memory: row send_idx[i] of H, for i in
 * [send_off[d], send_off[d+1]), is stored straight into peer d's halo buffer
 *   peer_base[d] + (peer_row0[d] + i - send_off[d]) * ldo
 * (peer_base: device array of `n_peers` pointers mapped with CUDA IPC; entries whose send range is
 * empty may be NULL).  send_off / peer_base / peer_row0 are DEVICE arrays.  The list is walked
 * starting at row `rotate` (wrapping), so ranks can start with different destinations and avoid
 * all hitting the same receiver at once.  The caller follows it
 * with a cross-rank barrier (a one-element NCCL all-reduce) before reading its own halo rows. */
int gnntf_halo_push_f32(const float* H, int64_t ld, const int32_t* send_idx, const int64_t* send_off,
                        float* const* peer_base, const int64_t* peer_row0, int n_peers,
                        int64_t n_send, int64_t rotate, int64_t ldo, int64_t F, void* stream);

/* The same push, followed by a COMPLETION SIGNAL over peer memory: after all of this rank's rows have
 * been stored (fenced at system scope), the last CTA writes  *epoch_base + epoch_delta  with
 * st.release.sys into slot `my_slot` of every peer's flag array (peer_flags: DEVICE array of n_peers
 * pointers, NULL = skip; each flag array is int32 [n_ranks] in that peer's memory, mapped with CUDA
 * IPC).  done_counter: one zeroed int32 in local memory (left at zero).  epoch_base is a DEVICE scalar
 * so that the whole K-step sequence can be captured in a CUDA graph and replayed with new epochs.
 * gnntf_flags_wait: enqueue a one-warp kernel that spins (ld.acquire.sys) until flags[i] >= *epoch_base +
 * epoch_delta for every i != skip — placed on the consumer's stream in front of the kernel that reads
 * the halo rows; traps after ~4 s if a peer never arrives.  gnntf_flags_signal: publish an epoch to the
 * peers without pushing rows (end-of-propagation acknowledgement: "my halo buffers may be overwritten").
 * Each rank must run on its own GPU: a spinning kernel and the kernel it waits for must be co-resident. */
int gnntf_halo_push_signal_f32(const float* H, int64_t ld, const int32_t* send_idx, const int64_t* send_off,
                               float* const* peer_base, const int64_t* peer_row0, int n_peers,
                               int64_t n_send, int64_t rotate, int64_t ldo, int64_t F,
                               int32_t* done_counter, int32_t* const* peer_flags, int my_slot,
                               const int32_t* epoch_base, int32_t epoch_delta, void* stream);
/* One step of a shard over its OWNED columns — H_out = (1-alpha)·A·H_in + alpha·H0, or plain A·H_in when
 * H0 == NULL — with the halo push of H_in's rows (and its completion signal, as above) riding in the SAME
 * launch: the leading CTAs of the SpMM grid do the push, so it is dispatched first and overlaps the
 * whole pass without a second stream.  Peer halo buffers have leading dimension ld.  Layouts the float4
 * path cannot take fall back to two launches on `stream` (same result). */
int gnntf_step_push_f32(const gnntf_csr_t* A, const float* H_in, const float* H0, float* H_out, int64_t ld,
                        int64_t F, double alpha, const int32_t* send_idx, const int64_t* send_off,
                        float* const* peer_base, const int64_t* peer_row0, int n_peers, int64_t n_send,
                        int64_t rotate, int32_t* done_counter, int32_t* const* peer_flags, int my_slot,
                        const int32_t* epoch_base, int32_t epoch_delta, void* stream);
int gnntf_flags_wait(const int32_t* flags, int n, int skip, const int32_t* epoch_base, int32_t epoch_delta,
                     void* stream);
int gnntf_flags_signal(int32_t* const* peer_flags, int n_peers, int my_slot, const int32_t* epoch_base,
                       int32_t epoch_delta, void* stream);

/* Copy-engine form of the exchange (all-gather layout: every rank's buffer holds ALL rows, a rank's own block
 * at its global row offset): `bytes` from src (local) to dst (a peer's mapped buffer) as one DMA copy on
 * `stream`, then — behind it on the same stream — a one-thread kernel that publishes *epoch_value (DEVICE int32)
 * into peer_flag with st.release.sys (the peer's flag slot for this rank; NULL = no signal).  The consumer
 * acquires the flag with gnntf_flags_wait.  The transfer takes no SM: the SpMM that runs beside it keeps the
 * whole gather path. */
int gnntf_peer_copy_signal(void* dst, const void* src, size_t bytes, int32_t* peer_flag,
                           const int32_t* epoch_value, void* stream);

/* Peer-memory plumbing for gnntf_halo_push_f32 (the only entry points that allocate; used once at
 * shard set-up).  gnntf_ipc_alloc: cudaMalloc + cudaIpcGetMemHandle (handle = 64 bytes, shipped to
 * the peers through the host's own channel, e.g. torch.distributed.all_gather_object).
 * gnntf_ipc_open: map a peer's allocation into this process (peer access enabled lazily). */
int gnntf_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char handle[64]);
int gnntf_ipc_open(const unsigned char handle[64], void** dev_ptr);
int gnntf_ipc_close(void* dev_ptr);
int gnntf_ipc_free(void* dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* GNNTF_B200_H */
