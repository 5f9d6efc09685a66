/* cspn_b200.h - C ABI of the B200-native CSPN affinity-propagation library.
 *
 * This is the drop-in boundary for ONE hot path of dontLoveBugs/CSPN_monodepth: the
 * AffinityPropagate module.  The reference has no native interface for this path (it is
 * ~20 ATen calls per iteration issued from Python); the entry points below are what a
 * binding of that module's forward/backward needs, one call per nn.Module.forward /
 * autograd backward.  Reference interfaces replaced (paths relative to the reference):
 *
 *   cspn_fwd_*  mode CSPN_MODE_NEW   network/libs/post_process/CSPN_new.py:26-92
 *               AffinityPropagate(prop_time, prop_kernel=3).forward(guidance, blur_depth, sparse_depth)
 *               called from network/unet_cspn_nyu.py:386
 *   cspn_fwd_*  mode CSPN_MODE_OURS  network/libs/post_process/CSPN_ours.py:24-54
 *               AffinityPropagate(prop_time).forward(x, guided, sparse_depth)
 *               (pixel-adaptive conv network/libs/base/pac.py:124-144, Conv2dFn.forward :75-94)
 *               called from network/unet_ours.py:333
 *   cspn_bwd_*  autograd through the loops at CSPN_new.py:80-90 / CSPN_ours.py:47-53
 *               (Conv2dFn.backward, pac.py:96-121)
 *
 * Conventions
 *   - All tensors are NCHW, innermost dimension contiguous, plane stride H*W, channel
 *     stride H*W.  `guidance` may carry more channels than the K*K-1 taps that are read
 *     (the NYU UNet hands over 12, unet_cspn_nyu.py:332): it is addressed through
 *     `guidance_batch_stride` (in ELEMENTS).  Depth has C >= 1 channels that share the
 *     affinity of their image.  `sparse` may be NULL (no re-injection), or have 1 channel
 *     (broadcast over C) or C channels (`sparse_channels`).
 *   - *_f32: every tensor is float.  *_f16: guidance/depth/sparse/out/grads are IEEE half
 *     (2 bytes), arithmetic is fp32 inside the kernels, the workspace is fp32.
 *   - Device entry points take DEVICE pointers, enqueue on `stream` (a cudaStream_t passed
 *     as void*; NULL = legacy default stream), never synchronise, never allocate: scratch
 *     comes from the caller (`workspace`, size from cspn_*_workspace_bytes).  They are
 *     re-entrant, usable from several host threads on different devices (the
 *     reference's DataParallel replicas, network/libs/base/encoding.py:102-105) and
 *     capturable in CUDA graphs.  The only state they consult is the process-wide
 *     diagnostic override cspn_set_path() (tests / benchmarks; leave it at CSPN_PATH_AUTO
 *     in production) and per-device capability caches; cspn_last_*() are per host thread.
 *     Inputs are never modified; `out` must not alias inputs.
 *   - Numerics: fp32 arithmetic with IEEE FMAs everywhere; the softmax of mode OURS uses expf
 *     in the 3x3 kernels and `ex2.approx.ftz` (2 ulp, inputs pre-scaled by log2 e) in the
 *     blocked 5x5 kernels - both far inside the 1e-4 parity bound, neither bit-identical to
 *     torch.softmax.
 *   - Host entry points (cspn_fwd_host_*) take HOST pointers (pinned or pageable), do
 *     H2D copy -> kernels -> D2H copy on `stream` with stream-ordered device allocations
 *     and return after the result is in `out` (they synchronise the stream).
 *   - Return value: 0 = success; CSPN_ERR_* (negative) = rejected arguments, nothing was
 *     launched; positive = a cudaError_t from the CUDA runtime.  cspn_error_string() maps
 *     all three.  (The reference's only native convention is "int, checked by _check",
 *     network/libs/inplace_abn/functions.py:13-16.)
 *   - There is no CPU fallback: without a CUDA device the calls return a cudaError_t.
 */
#ifndef CSPN_B200_H
#define CSPN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSPN_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define CSPN_API __attribute__((visibility("default")))
#else
#define CSPN_API
#endif

/* mode */
#define CSPN_MODE_NEW 0  /* abs-normalised, weights indexed at the NEIGHBOUR, border renormalisation (CSPN_new.py) */
#define CSPN_MODE_OURS 1 /* softmax-normalised, weights indexed at the CENTRE, zero padding (CSPN_ours.py + pac.py) */

/* errors */
#define CSPN_OK 0
#define CSPN_ERR_NULL_POINTER (-1)
#define CSPN_ERR_BAD_SHAPE (-2)      /* B,C,H,W < 1 (B == 0 is a no-op success), iters < 0 */
#define CSPN_ERR_BAD_KERNEL_SIZE (-3)/* mode NEW: ksize != 3 (CSPN_new.py:122 only works for 3); mode OURS: ksize not odd in [3,7] */
#define CSPN_ERR_BAD_MODE (-4)
#define CSPN_ERR_BAD_STRIDE (-5)     /* guidance_batch_stride < (K*K-1)*H*W */
#define CSPN_ERR_WORKSPACE (-6)      /* workspace NULL or too small */
#define CSPN_ERR_BAD_SPARSE_CHANNELS (-7)
#define CSPN_ERR_ALIAS (-8)          /* out aliases an input */
#define CSPN_ERR_EXCHANGE_TIMEOUT (-9) /* host entry points only: a tile of a fused kernel never received its neighbours' halo ring
                                          (bounded spin expired; the output is NaN-filled).  Device entry points are asynchronous:
                                          a fused forward / backward whose plan exchanges halos through global memory (any fused
                                          3x3 call with a non-zero cspn_*_workspace_bytes) reports it through the first int of
                                          `workspace` (0 = fine, 1 = timeout), valid once the stream has reached the end of the call. */

/* path selection (cspn_set_path): which CUDA implementation the forward uses */
#define CSPN_PATH_AUTO 0    /* fused single-launch kernel whenever the configuration is supported */
#define CSPN_PATH_GENERIC 1 /* one launch per iteration (any K, any shape); for debugging and A/B timing */
#define CSPN_PATH_FUSED 2   /* fused kernel or CSPN_ERR_BAD_KERNEL_SIZE if unsupported */
#define CSPN_PATH_BLOCKED 3 /* reported by cspn_last_path only: temporally blocked 5x5 forward (4 steps per launch), chosen by AUTO */

CSPN_API int cspn_abi_version(void);
CSPN_API const char* cspn_error_string(int code);

/* Process-wide override of the forward path, mainly for tests and benchmarks. Returns the previous value. */
CSPN_API int cspn_set_path(int path);
/* Which path the last successful cspn_fwd_* / cspn_bwd_* call on this host thread took (CSPN_PATH_GENERIC, _FUSED or _BLOCKED). */
CSPN_API int cspn_last_path(void);
/* Number of kernel launches enqueued by the last successful call on this host thread. */
CSPN_API int cspn_last_launch_count(void);

/* Scratch sizes in bytes (0 is possible). */
CSPN_API size_t cspn_fwd_workspace_bytes(int B, int C, int H, int W, int iters, int ksize, int mode);
CSPN_API size_t cspn_bwd_workspace_bytes(int B, int C, int H, int W, int iters, int ksize, int mode);

/* Diagnostics: which forward kernel the planner picks for a problem on the current device (what the launch will use
 * for TMA-addressable guidance).  plan10[0] = CSPN_KERNEL_*; for CSPN_KERNEL_DUAL the rest is {rows per warp P (tile =
 * 64 x 8P pixels), tiles per unit cx, cy, units per image plane ntx, nty, CTAs, rounds, units per class and round,
 * units}; for CSPN_KERNEL_SINGLE {halo transport CSPN_TRANSPORT_*, cluster / image tiling cx, cy, cluster tiles per image plane
 * ntx, nty, CTAs per plane}; zeros otherwise.  Not needed to call the operators. */
#define CSPN_TRANSPORT_CLUSTER 0 /* hardware clusters of cx x cy CTAs, halo ring through DSMEM, decaying margins between cluster tiles */
#define CSPN_TRANSPORT_STREAM 1  /* an image = cx x cy tiles, halo ring through global-memory inboxes, persistent cooperative grid */
#define CSPN_TRANSPORT_HYBRID 2  /* an image = cy hardware clusters of (cx, 1): left / right through DSMEM, up / down through the inboxes */
#define CSPN_KERNEL_GENERIC 0
#define CSPN_KERNEL_SINGLE 1  /* one 64 x 80 register tile per CTA, hardware clusters (DSMEM) or persistent stream (cspn_fused3x3.cuh) */
#define CSPN_KERNEL_DUAL 2    /* two register tiles per CTA worked on alternately, halo messages through L2 (cspn_dual3x3.cu) */
#define CSPN_KERNEL_BLOCKED 3 /* temporally blocked 5x5 (cspn_blocked5x5.cu) */
CSPN_API int cspn_fwd_plan(int B, int C, int H, int W, int iters, int ksize, int mode, int* plan10);

/* Forward: out[B,C,H,W] = r^iters. iters == 0 copies depth to out. */
CSPN_API int cspn_fwd_f32(const float* guidance, int64_t guidance_batch_stride,
                 const float* depth, const float* sparse, int sparse_channels, float* out,
                 int B, int C, int H, int W, int iters, int ksize, int mode,
                 void* workspace, size_t workspace_bytes, void* stream);
CSPN_API int cspn_fwd_f16(const void* guidance, int64_t guidance_batch_stride,
                 const void* depth, const void* sparse, int sparse_channels, void* out,
                 int B, int C, int H, int W, int iters, int ksize, int mode,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Backward: given grad_out[B,C,H,W] writes grad_guidance[B,Cg,H,W] completely (channels that
 * the forward does not read get exact zeros; batch stride Cg*H*W) and grad_depth[B,C,H,W].
 * There is no gradient for `sparse` (sign() has zero derivative, CSPN_new.py:78). */
CSPN_API int cspn_bwd_f32(const float* grad_out, const float* guidance, int64_t guidance_batch_stride, int Cg,
                 const float* depth, const float* sparse, int sparse_channels,
                 float* grad_guidance, float* grad_depth,
                 int B, int C, int H, int W, int iters, int ksize, int mode,
                 void* workspace, size_t workspace_bytes, void* stream);
CSPN_API int cspn_bwd_f16(const void* grad_out, const void* guidance, int64_t guidance_batch_stride, int Cg,
                 const void* depth, const void* sparse, int sparse_channels,
                 void* grad_guidance, void* grad_depth,
                 int B, int C, int H, int W, int iters, int ksize, int mode,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Host-buffer forward (end-to-end path: H2D, kernels, D2H inside the call). */
CSPN_API int cspn_fwd_host_f32(const float* guidance, int64_t guidance_batch_stride,
                      const float* depth, const float* sparse, int sparse_channels, float* out,
                      int B, int C, int H, int W, int iters, int ksize, int mode, void* stream);
CSPN_API int cspn_fwd_host_f16(const void* guidance, int64_t guidance_batch_stride,
                      const void* depth, const void* sparse, int sparse_channels, void* out,
                      int B, int C, int H, int W, int iters, int ksize, int mode, void* stream);

/* Pipelined host-buffer forward: cspn_fwd_host_submit_* enqueues H2D copy -> kernels -> D2H copy on a library-owned
 * stream and returns at once with a ticket (> 0; 0 for an empty batch); cspn_host_wait(ticket) blocks until `out` of
 * that call is valid and returns its status.  Up to cspn_host_pipeline_depth() calls of a host thread are in flight per
 * device: the H2D copy of one call overlaps the kernel and the D2H copy of the previous one.  A further submit first waits
 * for the oldest call in flight.  The host buffers (ideally pinned) must stay valid and unmodified until the call has been
 * waited for.  Tickets belong to the submitting host thread and device. */
CSPN_API int cspn_fwd_host_submit_f32(const float* guidance, int64_t guidance_batch_stride,
                      const float* depth, const float* sparse, int sparse_channels, float* out,
                      int B, int C, int H, int W, int iters, int ksize, int mode, int* ticket);
CSPN_API int cspn_fwd_host_submit_f16(const void* guidance, int64_t guidance_batch_stride,
                      const void* depth, const void* sparse, int sparse_channels, void* out,
                      int B, int C, int H, int W, int iters, int ksize, int mode, int* ticket);
CSPN_API int cspn_host_wait(int ticket);
CSPN_API int cspn_host_pipeline_depth(void);

/* ---- the two output heads directly upstream of the module ---------------------------------------------
 * Simple_Gudi_UpConv_Block_Last_Layer, network/unet_cspn_nyu.py:195-218 (instances :331-332, applied to the same x at :383-384)
 * and network/unet_ours.py:194-202 (:278-279, :331-332): 2x unpooling by zero insertion, cropped to H x W (H <= 2h, W <= 2w),
 * then conv3x3 (padding 1, no bias).  One pass over x [B,Cin,h,w] for BOTH heads: weights w1 [n1,Cin,3,3] -> out1 [B,n1,H,W]
 * (the blur depth, n1 = 1) and w2 [n2,Cin,3,3] -> out2 [B,n2,H,W] (the guidance, n2 = 8 / 12); n2 = 0 with w2 = out2 = NULL runs
 * a single head.  n1 >= 1, n1 + n2 <= 16, Cin <= 256.  The zeros of the unpooled tensor are never materialised or multiplied.
 * Backward: go1 / go2 are the gradients of out1 / out2; writes gx [B,Cin,h,w] (skipped when NULL) and gw1 / gw2 (skipped when
 * gw1 is NULL; deterministic split-K reduction through `workspace`, cspn_heads_workspace_bytes() bytes, 16-byte aligned).
 * Same conventions as cspn_fwd_* (device pointers, stream-ordered, no allocation, graph-capturable). */
CSPN_API size_t cspn_heads_workspace_bytes(void);
CSPN_API int cspn_heads_fwd_f32(const float* x, const float* w1, const float* w2, float* out1, float* out2,
                                int B, int Cin, int h, int w, int H, int W, int n1, int n2, void* stream);
CSPN_API int cspn_heads_fwd_f16(const void* x, const void* w1, const void* w2, void* out1, void* out2,
                                int B, int Cin, int h, int w, int H, int W, int n1, int n2, void* stream);
CSPN_API int cspn_heads_bwd_f32(const float* x, const float* w1, const float* w2, const float* go1, const float* go2,
                                float* gx, float* gw1, float* gw2, int B, int Cin, int h, int w, int H, int W, int n1, int n2,
                                void* workspace, size_t workspace_bytes, void* stream);
CSPN_API int cspn_heads_bwd_f16(const void* x, const void* w1, const void* w2, const void* go1, const void* go2,
                                void* gx, void* gw1, void* gw2, int B, int Cin, int h, int w, int H, int W, int n1, int n2,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- in-place activated batch normalisation ------------------------------------------------------------
 * The reference's own native interface for this is network/libs/inplace_abn/src/lib_cffi.h / bn.h:
 *   bn_mean_var_cuda (bn.cu:235-249)  -> cspn_abn_stats_f32 + cspn_abn_finalize_f32
 *   bn_forward_cuda (:251-266) + leaky_relu_cuda / elu_cuda (:299-335)  -> cspn_abn_forward_f32 (one pass, in place)
 *   leaky_relu_backward_cuda / elu_backward_cuda / elu_inv_cuda (:312-377) + bn_edz_eydz_cuda (:268-283)  -> cspn_abn_bwd_reduce_f32
 *   bn_backard_cuda (:285-297)  -> cspn_abn_bwd_apply_f32
 * called from functions.py:70-163 (InPlaceABN) and :166-297 (InPlaceABNSync).  fp32, x [N,C,S] contiguous (S = H*W).
 * activation: 0 none, 1 leaky_relu(slope), 2 elu.  weight / bias may be NULL (affine=False): gamma = |weight| + eps, beta = bias.
 * sums: DEVICE vector of 2*C doubles - {sum x, sum x^2} per channel for the statistics, {sum dz, sum y*dz} for the backward.
 * The synchronised variant all-reduces (SUM) exactly this vector across ranks between the reduce and the finalize / apply call
 * and passes count = N*S*world.  workspace: cspn_abn_workspace_bytes(C) bytes, 8-byte aligned.
 *   cspn_abn_finalize_f32: mean, var (biased) from sums / count; running_mean / running_var (may be NULL) updated with momentum and
 *     the n/(n-1) correction (functions.py:90-92).
 *   cspn_abn_forward_f32: x <- act((x - mean) / sqrt(var + eps) * gamma + beta) IN PLACE.
 *   cspn_abn_bwd_reduce_f32: z is the saved OUTPUT, dz its gradient; neither is modified (the activation is undone on the fly).
 *   cspn_abn_bwd_apply_f32: dx (may alias dz; NULL = skip) = (dz' - edz - y*eydz) * gamma / sqrt(var + eps); sums == NULL is the
 *     inference-mode backward (edz = eydz = 0); dweight / dbias (may be NULL) = sign(weight) * eydz * count_local, edz * count_local. */
CSPN_API size_t cspn_abn_workspace_bytes(int C);
CSPN_API int cspn_abn_stats_f32(const float* x, int N, int C, int S, double* sums, void* workspace, size_t workspace_bytes, void* stream);
CSPN_API int cspn_abn_finalize_f32(const double* sums, double count, float* mean, float* var, float* running_mean, float* running_var,
                                   float momentum, int C, void* stream);
CSPN_API int cspn_abn_forward_f32(float* x, const float* mean, const float* var, const float* weight, const float* bias,
                                  int N, int C, int S, float eps, int activation, float slope, void* stream);
CSPN_API int cspn_abn_bwd_reduce_f32(const float* z, const float* dz, const float* weight, const float* bias, int N, int C, int S,
                                     float eps, int activation, float slope, double* sums, void* workspace, size_t workspace_bytes, void* stream);
CSPN_API int cspn_abn_bwd_apply_f32(const float* z, const float* dz, float* dx, const float* var, const float* weight, const float* bias,
                                    const double* sums, double count_total, double count_local, float* dweight, float* dbias,
                                    int N, int C, int S, float eps, int activation, float slope, void* stream);

/* ---- legacy max-of-8 CSPN ---------------------------------------------------------------------------
 * network/libs/post_process/CSPN.py:19-56, AffinityPropagate().forward(guidance, blur_depth, sparse_depth), and :132-164,
 * AffinityPropagate_prediction().forward(guidance, blur_depth) (sparse == NULL).  Per step and gate k = 0..7 (|guidance[:,k]|,
 * not shifted): out_k = box3x3(g_k * r) / box3x3(g_k) (zero padded, centre included), r = max_k out_k (NaN propagates like
 * torch.max), r = (1 - m) r + m * sparse with m = sign(sparse): the SPARSE SAMPLE is re-injected and also seeds r^0.  The
 * reference runs a fixed 16 steps (:35); `iters` >= 1 is a parameter here.  One depth channel, one sparse channel; guidance
 * may carry more than 8 channels (batch stride in elements).  Forward only: 4 steps per launch, ceil(iters / 4) launches
 * (cspn_legacy.cu); workspace = cspn_legacy_workspace_bytes, 16-byte aligned.  Same conventions as cspn_fwd_*. */
CSPN_API size_t cspn_legacy_workspace_bytes(int B, int H, int W, int iters);
CSPN_API int cspn_legacy_fwd_f32(const float* guidance, int64_t guidance_batch_stride, const float* depth, const float* sparse, float* out,
                                 int B, int H, int W, int iters, void* workspace, size_t workspace_bytes, void* stream);
CSPN_API int cspn_legacy_fwd_f16(const void* guidance, int64_t guidance_batch_stride, const void* depth, const void* sparse, void* out,
                                 int B, int H, int W, int iters, void* workspace, size_t workspace_bytes, void* stream);

/* ---- directly downstream of the module -----------------------------------------------------------
 * Masked L1 loss, libs/criterion/criteria.py:27-39 (MaskedL1Loss; reached through Criterion_No_DSN :170-188):
 *   loss = mean |target - pred| over the n elements with target > 0 (NaN when there is none, like the reference's mean of an
 *   empty tensor).  One streaming pass, deterministic two-stage reduction (no floating-point atomics).
 *   loss2 (device, 2 floats) = {loss, number of valid elements}; the backward reads it back:
 *   grad_pred[i] = grad_loss[0] * -sign(target[i] - pred[i]) / count on valid elements, 0 elsewhere (grad_loss NULL = 1).
 * Depth metrics, libs/metrics.py:49-83 (Result.evaluate), same pass structure:
 *   out11 (device, 11 floats) = {irmse, imae, mse, rmse, mae, absrel, lg10, delta1, delta2, delta3, valid count}.
 * workspace: cspn_loss_workspace_bytes() bytes, 8-byte aligned, ZEROED ONCE before its first use (the kernels leave it ready
 * for the next call); pred / target are device pointers, fp32 or fp16 (accumulation in double). */
CSPN_API size_t cspn_loss_workspace_bytes(void);
CSPN_API int cspn_masked_l1_fwd_f32(const float* pred, const float* target, int64_t n, float* loss2, void* workspace, size_t workspace_bytes, void* stream);
CSPN_API int cspn_masked_l1_fwd_f16(const void* pred, const void* target, int64_t n, float* loss2, void* workspace, size_t workspace_bytes, void* stream);
CSPN_API int cspn_masked_l1_bwd_f32(const float* pred, const float* target, int64_t n, const float* loss2, const float* grad_loss, float* grad_pred, void* stream);
CSPN_API int cspn_masked_l1_bwd_f16(const void* pred, const void* target, int64_t n, const float* loss2, const float* grad_loss, void* grad_pred, void* stream);
CSPN_API int cspn_depth_metrics_f32(const float* pred, const float* target, int64_t n, float* out11, void* workspace, size_t workspace_bytes, void* stream);
CSPN_API int cspn_depth_metrics_f16(const void* pred, const void* target, int64_t n, float* out11, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSPN_B200_H */
