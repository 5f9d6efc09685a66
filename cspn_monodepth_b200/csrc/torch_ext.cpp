// PyTorch operator layer over the C ABI (include/cspn_b200.h): TORCH_LIBRARY(cspn, ...) with a CUDA implementation and an
// autograd formula, so that an eager training step calls straight into libcspn_b200.so from C++ (no ctypes, no Python
// allocation logic) and the operator is visible to torch.library consumers.  This file contains no arithmetic: it checks
// arguments the way the reference's ATen calls would (RuntimeError via TORCH_CHECK), allocates outputs / workspace with
// torch's caching allocator on the current stream and forwards to cspn_fwd_* / cspn_bwd_*.
//
//   cspn::propagate(Tensor guidance, Tensor depth, Tensor? sparse, int iters, int ksize, int mode) -> Tensor      (differentiable)
//   cspn::forward  (same)                                                                        -> Tensor      (no autograd)
//   cspn::backward (Tensor grad_out, Tensor guidance, Tensor depth, Tensor? sparse, int iters, int ksize, int mode) -> (Tensor, Tensor)
//
// Reference call sites this serves: network/unet_cspn_nyu.py:386 (mode 0), network/unet_ours.py:333 (mode 1).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/csrc/autograd/custom_function.h>
#include <torch/library.h>

#include "../../include/cspn_b200.h"

namespace {

using at::Tensor;

void check_inputs(const Tensor& guidance, const Tensor& depth, const c10::optional<Tensor>& sparse, int64_t ksize)
{
    TORCH_CHECK(guidance.is_cuda() && depth.is_cuda(), "cspn: guidance and depth must be CUDA tensors (the B200 CSPN operator has no CPU fallback)");
    TORCH_CHECK(guidance.scalar_type() == at::kFloat || guidance.scalar_type() == at::kHalf, "cspn: unsupported dtype ", guidance.scalar_type(), " (float32 and float16 are supported)");
    TORCH_CHECK(depth.scalar_type() == guidance.scalar_type(), "cspn: dtype mismatch between guidance and depth");
    TORCH_CHECK(guidance.dim() == 4 && depth.dim() == 4, "cspn: guidance and depth must be 4-D NCHW tensors");
    TORCH_CHECK(guidance.size(0) == depth.size(0) && guidance.size(2) == depth.size(2) && guidance.size(3) == depth.size(3),
                "cspn: guidance ", guidance.sizes(), " does not match depth ", depth.sizes());
    TORCH_CHECK(guidance.size(1) >= ksize * ksize - 1, "cspn: guidance has ", guidance.size(1), " channels, the propagation kernel needs ", ksize * ksize - 1);
    TORCH_CHECK(guidance.device() == depth.device(), "cspn: guidance and depth must be on the same device");
    if (sparse.has_value() && sparse->defined()) {
        const Tensor& s = *sparse;
        TORCH_CHECK(s.is_cuda() && s.device() == depth.device() && s.scalar_type() == depth.scalar_type(), "cspn: sparse_depth must match depth in device and dtype");
        TORCH_CHECK(s.dim() == 4 && s.size(0) == depth.size(0) && s.size(2) == depth.size(2) && s.size(3) == depth.size(3) && (s.size(1) == 1 || s.size(1) == depth.size(1)),
                    "cspn: sparse_depth ", s.sizes(), " does not match depth ", depth.sizes());
    }
}

// guidance may be a channel-narrowed view of a wider tensor: only its batch stride is free
Tensor guidance_view(const Tensor& g, int64_t* batch_stride)
{
    const int64_t h = g.size(2), w = g.size(3);
    if (g.stride(3) == 1 && g.stride(2) == w && g.stride(1) == h * w && g.stride(0) >= g.size(1) * h * w) { *batch_stride = g.stride(0); return g; }
    Tensor c = g.contiguous();
    *batch_stride = c.size(1) * h * w;
    return c;
}

void raise_on(int rc)
{
    TORCH_CHECK(rc == 0, cspn_error_string(rc), " (code ", rc, ")");
}

Tensor forward_cuda(const Tensor& guidance, const Tensor& depth, const c10::optional<Tensor>& sparse, int64_t iters, int64_t ksize, int64_t mode)
{
    check_inputs(guidance, depth, sparse, ksize);
    if (iters == 0) return depth;
    const c10::cuda::CUDAGuard guard(depth.device());
    int64_t gbs = 0;
    const Tensor g = guidance_view(guidance, &gbs);
    const Tensor d = depth.contiguous();
    Tensor s;
    if (sparse.has_value() && sparse->defined()) s = sparse->contiguous();
    Tensor out = at::empty_like(d);
    const int B = (int)d.size(0), C = (int)d.size(1), H = (int)d.size(2), W = (int)d.size(3);
    const size_t nws = cspn_fwd_workspace_bytes(B, C, H, W, (int)iters, (int)ksize, (int)mode);
    Tensor ws = at::empty({(int64_t)nws}, d.options().dtype(at::kByte));
    void* stream = (void*)at::cuda::getCurrentCUDAStream().stream();
    const void* sp = s.defined() ? s.data_ptr() : nullptr;
    const int sc = s.defined() ? (int)s.size(1) : 1;
    if (d.scalar_type() == at::kFloat)
        raise_on(cspn_fwd_f32((const float*)g.data_ptr(), gbs, (const float*)d.data_ptr(), (const float*)sp, sc, (float*)out.data_ptr(), B, C, H, W,
                              (int)iters, (int)ksize, (int)mode, nws ? ws.data_ptr() : nullptr, nws, stream));
    else
        raise_on(cspn_fwd_f16(g.data_ptr(), gbs, d.data_ptr(), sp, sc, out.data_ptr(), B, C, H, W, (int)iters, (int)ksize, (int)mode,
                              nws ? ws.data_ptr() : nullptr, nws, stream));
    return out;
}

std::tuple<Tensor, Tensor> backward_cuda(const Tensor& grad_out, const Tensor& guidance, const Tensor& depth, const c10::optional<Tensor>& sparse,
                                         int64_t iters, int64_t ksize, int64_t mode)
{
    check_inputs(guidance, depth, sparse, ksize);
    TORCH_CHECK(grad_out.is_cuda() && grad_out.sizes() == depth.sizes() && grad_out.scalar_type() == depth.scalar_type(), "cspn: grad_out must match depth");
    const c10::cuda::CUDAGuard guard(depth.device());
    int64_t gbs = 0;
    const Tensor g = guidance_view(guidance, &gbs);
    const Tensor d = depth.contiguous();
    const Tensor go = grad_out.contiguous();
    Tensor s;
    if (sparse.has_value() && sparse->defined()) s = sparse->contiguous();
    Tensor gg = at::empty(guidance.sizes(), guidance.options());
    Tensor gd = at::empty_like(d);
    const int B = (int)d.size(0), C = (int)d.size(1), H = (int)d.size(2), W = (int)d.size(3);
    const size_t nws = cspn_bwd_workspace_bytes(B, C, H, W, (int)iters, (int)ksize, (int)mode);
    Tensor ws = at::empty({(int64_t)nws}, d.options().dtype(at::kByte));
    void* stream = (void*)at::cuda::getCurrentCUDAStream().stream();
    const void* sp = s.defined() ? s.data_ptr() : nullptr;
    const int sc = s.defined() ? (int)s.size(1) : 1;
    if (d.scalar_type() == at::kFloat)
        raise_on(cspn_bwd_f32((const float*)go.data_ptr(), (const float*)g.data_ptr(), gbs, (int)guidance.size(1), (const float*)d.data_ptr(), (const float*)sp, sc,
                              (float*)gg.data_ptr(), (float*)gd.data_ptr(), B, C, H, W, (int)iters, (int)ksize, (int)mode, nws ? ws.data_ptr() : nullptr, nws, stream));
    else
        raise_on(cspn_bwd_f16(go.data_ptr(), g.data_ptr(), gbs, (int)guidance.size(1), d.data_ptr(), sp, sc, gg.data_ptr(), gd.data_ptr(), B, C, H, W,
                              (int)iters, (int)ksize, (int)mode, nws ? ws.data_ptr() : nullptr, nws, stream));
    return std::make_tuple(gg, gd);
}

// autograd: saves the INPUTS only (no intermediates: the backward recomputes the recurrence), like the ctypes path
class Propagate : public torch::autograd::Function<Propagate> {
public:
    static Tensor forward(torch::autograd::AutogradContext* ctx, const Tensor& guidance, const Tensor& depth, const c10::optional<Tensor>& sparse,
                          int64_t iters, int64_t ksize, int64_t mode)
    {
        at::AutoDispatchBelowADInplaceOrView below;
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("cspn::forward", "").typed<Tensor(const Tensor&, const Tensor&, const c10::optional<Tensor>&, int64_t, int64_t, int64_t)>();
        Tensor out = op.call(guidance, depth, sparse, iters, ksize, mode);
        ctx->save_for_backward({guidance, depth, sparse.has_value() && sparse->defined() ? *sparse : Tensor()});
        ctx->saved_data["iters"] = iters; ctx->saved_data["ksize"] = ksize; ctx->saved_data["mode"] = mode;
        return out;
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::variable_list grads)
    {
        const auto saved = ctx->get_saved_variables();
        c10::optional<Tensor> sparse;
        if (saved[2].defined()) sparse = saved[2];
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("cspn::backward", "").typed<std::tuple<Tensor, Tensor>(const Tensor&, const Tensor&, const Tensor&, const c10::optional<Tensor>&, int64_t, int64_t, int64_t)>();
        auto r = op.call(grads[0], saved[0], saved[1], sparse, ctx->saved_data["iters"].toInt(), ctx->saved_data["ksize"].toInt(), ctx->saved_data["mode"].toInt());
        return {std::get<0>(r), std::get<1>(r), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

Tensor propagate_autograd(const Tensor& guidance, const Tensor& depth, const c10::optional<Tensor>& sparse, int64_t iters, int64_t ksize, int64_t mode)
{
    if (iters == 0) return depth;
    return Propagate::apply(guidance, depth, sparse, iters, ksize, mode);
}

}  // namespace

TORCH_LIBRARY(cspn, m)
{
    m.def("propagate(Tensor guidance, Tensor depth, Tensor? sparse, int iters, int ksize, int mode) -> Tensor");
    m.def("forward(Tensor guidance, Tensor depth, Tensor? sparse, int iters, int ksize, int mode) -> Tensor");
    m.def("backward(Tensor grad_out, Tensor guidance, Tensor depth, Tensor? sparse, int iters, int ksize, int mode) -> (Tensor, Tensor)");
}
TORCH_LIBRARY_IMPL(cspn, CUDA, m)
{
    m.impl("forward", &forward_cuda);
    m.impl("backward", &backward_cuda);
    m.impl("propagate", &forward_cuda);          // inference tensors / no_grad: straight to the kernel
}
// CPU tensors fail loudly, like the ctypes path: there is no CPU implementation of this operator anywhere in the product
namespace {
Tensor no_cpu_forward(const Tensor&, const Tensor& depth, const c10::optional<Tensor>&, int64_t, int64_t, int64_t)
{
    TORCH_CHECK(false, "cspn: depth is on ", depth.device(), ": the B200 CSPN operator is CUDA-only and has no CPU fallback");
}
std::tuple<Tensor, Tensor> no_cpu_backward(const Tensor&, const Tensor&, const Tensor& depth, const c10::optional<Tensor>&, int64_t, int64_t, int64_t)
{
    TORCH_CHECK(false, "cspn: depth is on ", depth.device(), ": the B200 CSPN operator is CUDA-only and has no CPU fallback");
}
}  // namespace
TORCH_LIBRARY_IMPL(cspn, CPU, m)
{
    m.impl("forward", &no_cpu_forward);
    m.impl("propagate", &no_cpu_forward);
    m.impl("backward", &no_cpu_backward);
}
TORCH_LIBRARY_IMPL(cspn, Autograd, m)
{
    m.impl("propagate", &propagate_autograd);
}
