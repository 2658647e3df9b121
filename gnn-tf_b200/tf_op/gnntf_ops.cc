// TensorFlow custom-op shim over the C-ABI (include/gnntf_b200.h).
//
// SOURCE ONLY in this repository: TensorFlow (headers, libtensorflow_framework) is not installable
// in the build image, so this file is compile-gated (see tf_op/Makefile: it builds only where
// `python -c "import tensorflow"` succeeds) and has not been compiled here.  It shows exactly what
// a maintainer of the reference adds to run its hot path on the sm_100a kernels:
//
//   GnntfSpmm             <- tf.sparse.sparse_dense_matmul(adj, H)        filter.py:19, gcn.py:88 ...
//   GnntfAppnpPropagate   <- the K PPRIteration layers                    filter.py:17-22,34-35
//   (graph2adj / get_adjacency run once per graph through the ctypes host shim, gnntf_tf.py; this file
//    registers no op for them)
//
// Kernels are stateless and re-entrant; every launch goes to the op's own GPU stream
// (ctx->eigen_gpu_device().stream()), temporaries come from ctx->allocate_temp, so TF may call
// Compute() from several inter-op threads.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#define EIGEN_USE_GPU
#include "gnntf_b200.h"

namespace tf = tensorflow;

namespace {

tf::Status FromGnntf(int rc, const char* what) {
  if (rc == GNNTF_OK) return tf::OkStatus();
  return tf::errors::InvalidArgument(what, ": ", gnntf_status_str(rc));
}

// The CSR travels as five tensors (row_ptr, col_idx, val, + the long-row plan packed in one int32
// tensor `plan` = [n_long, n_chunks, threshold, chunk, long_row..., long_first..., long_n...,
// chunk_row..., chunk_begin...]) so the op stays a pure function of its inputs.
gnntf_csr_t MakeCsr(const tf::Tensor& row_ptr, const tf::Tensor& col_idx, const tf::Tensor& val,
                    const tf::Tensor& plan_host, const tf::Tensor& plan_dev, float* partials) {
  gnntf_csr_t A{};
  A.n_rows = row_ptr.dim_size(0) - 1;
  A.nnz = col_idx.dim_size(0);
  A.row_ptr = row_ptr.flat<tf::int32>().data();
  A.col_idx = col_idx.flat<tf::int32>().data();
  A.val = val.flat<float>().data();
  const auto h = plan_host.flat<tf::int32>();  // host-memory copy of the 4 header ints
  A.n_long = h(0);
  A.n_chunks = h(1);
  A.long_threshold = h(2);
  A.chunk = h(3);
  if (A.n_long > 0) {
    const tf::int32* p = plan_dev.flat<tf::int32>().data();
    A.long_row = p;
    A.long_first_chunk = p + A.n_long;
    A.long_n_chunks = p + 2 * A.n_long;
    A.chunk_row = p + 3 * A.n_long;
    A.chunk_begin = p + 3 * A.n_long + A.n_chunks;
    A.partials = partials;
  }
  return A;
}

}  // namespace

REGISTER_OP("GnntfSpmm")
    .Input("row_ptr: int32").Input("col_idx: int32").Input("val: float")
    .Input("plan_header: int32").Input("plan: int32").Input("dense: float")
    .Output("product: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(5));
      return tf::OkStatus();
    });

REGISTER_OP("GnntfAppnpPropagate")
    .Input("row_ptr: int32").Input("col_idx: int32").Input("val: float")
    .Input("plan_header: int32").Input("plan: int32").Input("h0: float")
    .Attr("alpha: float = 0.1").Attr("iterations: int = 10")
    .Output("h_k: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(5));
      return tf::OkStatus();
    });

class GnntfSpmmOp : public tf::OpKernel {
 public:
  explicit GnntfSpmmOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {}
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& B = ctx->input(5);
    OP_REQUIRES(ctx, B.dims() == 2, tf::errors::InvalidArgument("dense operand must be [nodes, features]"));
    tf::Tensor* C = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, B.shape(), &C));
    const int64_t F = B.dim_size(1);
    tf::Tensor partials;
    const int n_chunks = ctx->input(3).flat<tf::int32>()(1);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({std::max<int64_t>(1, n_chunks * ((F + 3) / 4 * 4))}), &partials));
    gnntf_csr_t A = MakeCsr(ctx->input(0), ctx->input(1), ctx->input(2), ctx->input(3), ctx->input(4),
                            partials.flat<float>().data());
    auto stream = ctx->eigen_gpu_device().stream();
    OP_REQUIRES_OK(ctx, FromGnntf(gnntf_spmm_f32(&A, B.flat<float>().data(), F, C->flat<float>().data(), F, F, stream), "gnntf_spmm_f32"));
  }
};

class GnntfAppnpPropagateOp : public tf::OpKernel {
 public:
  explicit GnntfAppnpPropagateOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("alpha", &alpha_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("iterations", &iterations_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& H0 = ctx->input(5);
    OP_REQUIRES(ctx, H0.dims() == 2, tf::errors::InvalidArgument("h0 must be [nodes, features]"));
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, H0.shape(), &out));
    const int64_t F = H0.dim_size(1);
    tf::Tensor scratch, partials;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, H0.shape(), &scratch));
    const int n_chunks = ctx->input(3).flat<tf::int32>()(1);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({std::max<int64_t>(1, n_chunks * ((F + 3) / 4 * 4))}), &partials));
    gnntf_csr_t A = MakeCsr(ctx->input(0), ctx->input(1), ctx->input(2), ctx->input(3), ctx->input(4),
                            partials.flat<float>().data());
    auto stream = ctx->eigen_gpu_device().stream();
    OP_REQUIRES_OK(ctx, FromGnntf(gnntf_appnp_propagate_f32(&A, H0.flat<float>().data(), out->flat<float>().data(),
                                                            scratch.flat<float>().data(), F, F, alpha_, iterations_, stream),
                                  "gnntf_appnp_propagate_f32"));
  }

 private:
  float alpha_;
  int iterations_;
};

// plan_header lives in host memory so Compute can size temporaries without a device sync.
REGISTER_KERNEL_BUILDER(Name("GnntfSpmm").Device(tf::DEVICE_GPU).HostMemory("plan_header"), GnntfSpmmOp);
REGISTER_KERNEL_BUILDER(Name("GnntfAppnpPropagate").Device(tf::DEVICE_GPU).HostMemory("plan_header"), GnntfAppnpPropagateOp);
