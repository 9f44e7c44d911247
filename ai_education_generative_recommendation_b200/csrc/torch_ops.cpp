// torch_ops.cpp — the thin PyTorch C++ extension over the C ABI ("extension ops a replacement must export,
// registered with TORCH_LIBRARY").  Every op is a few lines: check the borrowed inputs (CUDA, fp32 / int64, contiguous —
// never copied silently), allocate the outputs from the caller's caching allocator, pass raw pointers and the CURRENT
// stream to librqvae_b200.so, and turn an error code into a c10::Error (Python RuntimeError) carrying rqb200_last_error().
// No kernel lives here and nothing synchronises that the C ABI call itself does not.
//
//   torch.ops.rqvae_b200.encode_indices(handle, x[N,in], mode)      -> idx[N,L] int64     RQVAE.get_indices (rqvae.py:67-71)
//   torch.ops.rqvae_b200.encode_latents(handle, x[N,in])            -> z[N,e]             encoder MLP, exact (layers.py:42-43)
//   torch.ops.rqvae_b200.quantize(handle, z[N,e])                   -> (idx, x_q, sumsq)  rq.py:39-56, vq.py:63-99
//   torch.ops.rqvae_b200.sinkhorn_assign(d[B,K], eps, iters)        -> idx[B] int64       vq.py:74-83, layers.py:85-108
//   torch.ops.rqvae_b200.resolve_collisions(handle, codes[N,L], Ks) -> ids[N,L+1] int64   infer.py:152-163
//   torch.ops.rqvae_b200.collision_rate(handle, codes[N,L], Ks)     -> (distinct, largest group)
// `handle` is the rqb200_model* of the Python RQVAE (model._handle) as an integer, so the ops are TorchScript-friendly.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include <vector>

#include "../../include/rqvae_b200.h"

namespace {

rqb200_model *as_model(int64_t handle) {
    TORCH_CHECK(handle != 0, "rqvae_b200: model handle is NULL");
    return reinterpret_cast<rqb200_model *>(static_cast<intptr_t>(handle));
}

void ok(int rc, const char *what) {
    TORCH_CHECK(rc == 0, "rqvae_b200::", what, ": ", rqb200_last_error());
}

void check_input(const at::Tensor &t, at::ScalarType dtype, int64_t dims, const char *name) {
    TORCH_CHECK(t.is_cuda(), "rqvae_b200: ", name, " must be a CUDA tensor (no CPU fallback)");
    TORCH_CHECK(t.scalar_type() == dtype, "rqvae_b200: ", name, " has the wrong dtype");
    TORCH_CHECK(t.dim() == dims, "rqvae_b200: ", name, " must have ", dims, " dimensions");
    TORCH_CHECK(t.is_contiguous(), "rqvae_b200: ", name, " must be contiguous");
}

void *cur_stream(const at::Tensor &t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

at::Tensor encode_indices(int64_t handle, const at::Tensor &x, int64_t mode) {
    check_input(x, at::kFloat, 2, "x");
    const c10::cuda::CUDAGuard guard(x.device());
    const int64_t n = x.size(0);
    const int L = rqb200_model_levels(as_model(handle));
    TORCH_CHECK(L > 0, "rqvae_b200::encode_indices: ", rqb200_last_error());
    at::Tensor idx = at::empty({n, L}, x.options().dtype(at::kLong));
    ok(rqb200_get_indices(as_model(handle), (int)mode, x.data_ptr<float>(), n, idx.data_ptr<int64_t>(), nullptr, nullptr, cur_stream(x)),
       "encode_indices");
    return idx;
}

at::Tensor encode_latents(int64_t handle, const at::Tensor &x) {
    check_input(x, at::kFloat, 2, "x");
    const c10::cuda::CUDAGuard guard(x.device());
    const int e = rqb200_model_e_dim(as_model(handle));
    TORCH_CHECK(e > 0, "rqvae_b200::encode_latents: ", rqb200_last_error());
    at::Tensor z = at::empty({x.size(0), e}, x.options());
    ok(rqb200_mlp_exact(as_model(handle), 0, x.data_ptr<float>(), nullptr, x.size(0), z.data_ptr<float>(), cur_stream(x)),
       "encode_latents");
    return z;
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> quantize(int64_t handle, const at::Tensor &z) {
    check_input(z, at::kFloat, 2, "z");
    const c10::cuda::CUDAGuard guard(z.device());
    const int L = rqb200_model_levels(as_model(handle));
    TORCH_CHECK(L > 0, "rqvae_b200::quantize: ", rqb200_last_error());
    const int64_t n = z.size(0);
    at::Tensor idx = at::empty({n, L}, z.options().dtype(at::kLong));
    at::Tensor xq = at::empty_like(z);
    at::Tensor sumsq = at::zeros({L}, z.options().dtype(at::kDouble));
    ok(rqb200_quantize(as_model(handle), z.data_ptr<float>(), n, idx.data_ptr<int64_t>(), nullptr, xq.data_ptr<float>(),
                       sumsq.data_ptr<double>(), nullptr, cur_stream(z)),
       "quantize");
    return {idx, xq, sumsq};
}

at::Tensor sinkhorn_assign(const at::Tensor &d, double epsilon, int64_t iters) {
    check_input(d, at::kFloat, 2, "d");
    const c10::cuda::CUDAGuard guard(d.device());
    const int64_t B = d.size(0), K = d.size(1);
    at::Tensor scratch = at::empty({B, K}, d.options().dtype(at::kDouble));
    at::Tensor idx = at::empty({B}, d.options().dtype(at::kLong));
    ok(rqb200_sinkhorn_assign(d.data_ptr<float>(), B, (int)K, epsilon, (int)iters, scratch.data_ptr<double>(), idx.data_ptr<int64_t>(),
                              cur_stream(d)),
       "sinkhorn_assign");
    return idx;
}

at::Tensor resolve_collisions(int64_t handle, const at::Tensor &codes, at::IntArrayRef num_emb) {
    check_input(codes, at::kLong, 2, "codes");
    const c10::cuda::CUDAGuard guard(codes.device());
    const int64_t n = codes.size(0), L = codes.size(1);
    TORCH_CHECK(num_emb.empty() || (int64_t)num_emb.size() == L, "rqvae_b200: one codebook size per level expected");
    std::vector<int> Ks(num_emb.begin(), num_emb.end());
    at::Tensor out = at::empty({n, L + 1}, codes.options());
    ok(rqb200_suffix_dedup(as_model(handle), codes.data_ptr<int64_t>(), n, (int)L, Ks.empty() ? nullptr : Ks.data(), out.data_ptr<int64_t>(),
                           nullptr, nullptr, cur_stream(codes)),
       "resolve_collisions");
    return out;
}

std::tuple<int64_t, int64_t> collision_rate(int64_t handle, const at::Tensor &codes, at::IntArrayRef num_emb) {
    check_input(codes, at::kLong, 2, "codes");
    const c10::cuda::CUDAGuard guard(codes.device());
    const int64_t n = codes.size(0), L = codes.size(1);
    std::vector<int> Ks(num_emb.begin(), num_emb.end());
    at::Tensor out = at::empty({n, L + 1}, codes.options());
    int64_t distinct = 0, largest = 0;
    ok(rqb200_suffix_dedup(as_model(handle), codes.data_ptr<int64_t>(), n, (int)L, Ks.empty() ? nullptr : Ks.data(), out.data_ptr<int64_t>(),
                           &distinct, &largest, cur_stream(codes)),
       "collision_rate");
    return {distinct, largest};
}

}  // namespace

TORCH_LIBRARY(rqvae_b200, m) {
    m.def("encode_indices(int handle, Tensor x, int mode) -> Tensor");
    m.def("encode_latents(int handle, Tensor x) -> Tensor");
    m.def("quantize(int handle, Tensor z) -> (Tensor, Tensor, Tensor)");
    m.def("sinkhorn_assign(Tensor d, float epsilon, int iters) -> Tensor");
    m.def("resolve_collisions(int handle, Tensor codes, int[] num_emb) -> Tensor");
    m.def("collision_rate(int handle, Tensor codes, int[] num_emb) -> (int, int)");
}

TORCH_LIBRARY_IMPL(rqvae_b200, CUDA, m) {
    m.impl("encode_indices", &encode_indices);
    m.impl("encode_latents", &encode_latents);
    m.impl("quantize", &quantize);
    m.impl("sinkhorn_assign", &sinkhorn_assign);
    m.impl("resolve_collisions", &resolve_collisions);
    m.impl("collision_rate", &collision_rate);
}
