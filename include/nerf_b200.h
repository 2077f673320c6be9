/*
 * nerf_b200.h -- C ABI of libnerf_b200.so: the B200 (sm_100a) NeRF hot path.
 *
 * The reference (dgsmith7/nerf-mlp) has no FFI layer: its operator API is the Python class
 * surface of `nerfmlp`.  The Python drop-in classes in nerf_mlp_b200/
 * bind exactly these entry points through ctypes; each one replaces the stock-PyTorch op
 * sequence cited next to it (file:line into the reference).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (torch tensors kept alive by the
 *    Python side) unless its name ends in `_host`;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing inside
 *    allocates device memory or synchronises;
 *  - return 0 on success; non-zero = failure (negative: argument check, positive: cudaError_t),
 *    message available from nerf_last_error() (thread-local);
 *  - there is no CPU fallback: a call on a machine without an sm_100 device fails.
 *  - all tensors are dense row-major float32 unless stated.
 */
#ifndef NERF_B200_H
#define NERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- fixed network geometry (nerfmlp/model.py:29-55, default ctor arguments) --------------- */
#define NERF_N_PARAMS      595844   /* 593 408 weights + 2 436 biases, state_dict order          */
#define NERF_XYZ_CH        63       /* 3 + 3*2*10, model.py:20-26 with L=10 (renderer.py:20)     */
#define NERF_DIR_CH        27       /* 3 + 3*2*4,  L=4 (renderer.py:21)                          */
#define NERF_W             256
#define NERF_MAX_SAMPLES   512      /* per-ray samples supported by the warp-per-ray kernels     */

/* precision modes of the MLP kernels */
#define NERF_PREC_BF16     0        /* tcgen05 bf16 tensor-core path (product path)              */
#define NERF_PREC_FP32     1        /* fp32 check mode (CUDA-core GEMMs), 1e-4 parity gate       */

/* Library / device probe: returns 0 and fills sm (e.g. 100) and the SM count. */
int nerf_device_info(int* sm_major_minor, int* sm_count);
const char* nerf_last_error(void);
const char* nerf_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long nerf_launch_count(void);

/* Bytes of the packed bf16 weight image consumed by the tcgen05 kernels. */
size_t nerf_packed_weight_bytes(void);

/* flat fp32 parameters (24 tensors, state_dict order: pts_linears.{0..7}.{weight,bias},
 * sigma_linear, bottleneck_linear, view_linear, rgb_linear) -> packed/swizzled bf16 image.
 * Replaces nothing in the reference (its weights stay fp32, model.py:39-53); must be re-run after
 * every parameter update (optimizer.step, load_state_dict, load_from_numpy model.py:83-127). */
int nerf_pack_weights(const float* flat_params, void* packed, void* stream);

/* PositionalEncoding.forward (model.py:20-26): out[n, d*(include_input+2L)] =
 * [x, sin(f0 x), cos(f0 x), ..., sin(f_{L-1} x), cos(f_{L-1} x)], freqs[L] on the device. */
int nerf_positional_encoding(const float* x, int64_t n, int d, const float* freqs, int L, int include_input,
                             float* out, void* stream);

/* z_vals[R,S] = near*(1-t)+far*t, optionally stratified-jittered with t_rand[R,S] (nullable).
 * Replaces renderer.py:52-61.  t_vals[S] comes from torch.linspace on the same device. */
int nerf_stratified_z(const float* t_vals, const float* t_rand, int R, int S, float near_, float far_,
                      float* z_vals, void* stream);

/* `save` argument of the MLP forwards:
 *   0                     inference
 *   1                     keep the activations the backward needs in the workspace
 *   NERF_FWD_DENSITY_ONLY inference that stops after layer 7 + sigma_linear: raw[...,3] (sigma) is the
 *                         reference's value, raw[...,0:3] = 0.  For the COARSE pass of render() and of a
 *                         training step, whose only consumer is the hierarchical resampling
 *                         (renderer.py:79-87 reads `weights`, a function of sigma alone; the coarse colour maps
 *                         are dropped at renderer.py:44 and never reach the loss, scripts/train.py:374-376):
 *                         skips bottleneck_linear, view_linear and rgb_linear (17 % of the pass).
 *                         (fp32 check mode evaluates the whole network in every mode.) */
#define NERF_FWD_DENSITY_ONLY 2

/* Workspace size for the MLP kernels: M sample rows, `save` as above. */
size_t nerf_mlp_workspace_bytes(int64_t M, int precision, int save);

/* MLP forward from rays: points o + d*z (renderer.py:63), *coord_scale (:67-68), positional
 * encoding L=10 (:70, model.py:20-26), view-direction normalisation d/(|d|+1e-8) + encoding L=4
 * (:72-74), NeRFMLP.forward (model.py:57-81).  raw[R,S,4] = [r,g,b,sigma] (pre-activation).
 * Nothing of the encodings is materialised in HBM in bf16 mode.
 * `params` = flat fp32 parameters; `packed` = bf16 image from nerf_pack_weights (bf16 mode). */
int nerf_mlp_fwd_rays(const float* rays_o, const float* rays_d, const float* z_vals, int R, int S,
                      float coord_scale, const float* params, const void* packed, float* raw,
                      void* workspace, size_t workspace_bytes, int precision, int save, void* stream);

/* Drop-in NeRFMLP.forward(x[M,63], viewdirs[M,27]) -> [M,4] on pre-encoded inputs
 * (model.py:57-81). */
int nerf_mlp_fwd_encoded(const float* x_enc, const float* d_enc, int64_t M, const float* params,
                         const void* packed, float* out, void* workspace, size_t workspace_bytes,
                         int precision, int save, void* stream);

/* Backward of either forward (the implicit autograd at scripts/train.py:382): given d_raw[M,4]
 * and the workspace written by a forward with save=1, ACCUMULATES into flat_grads[NERF_N_PARAMS]
 * (same layout as params).  Inputs receive no gradient.
 * rows_per_dir = samples per ray S for nerf_mlp_fwd_rays (rows sharing one view direction),
 * 1 for nerf_mlp_fwd_encoded. */
int nerf_mlp_bwd(const float* d_raw, int64_t M, int rows_per_dir, const float* params, const void* packed,
                 float* flat_grads, void* workspace, size_t workspace_bytes, int precision,
                 void* stream);
/* The same backward, one stage at a time (for per-stage timing / scheduling by the caller):
 * NERF_BWD_DGRAD writes d(pre-activation) of every layer into the workspace, NERF_BWD_WGRAD
 * accumulates the weight / bias gradients from it; NERF_BWD_ALL = both = nerf_mlp_bwd.
 * (fp32 check mode has a single fused backward: it runs on NERF_BWD_DGRAD | NERF_BWD_ALL and
 * NERF_BWD_WGRAD is a no-op.) */
#define NERF_BWD_ALL   0
#define NERF_BWD_DGRAD 1
#define NERF_BWD_WGRAD 2
int nerf_mlp_bwd_stage(const float* d_raw, int64_t M, int rows_per_dir, const float* params, const void* packed,
                       float* flat_grads, void* workspace, size_t workspace_bytes, int precision, int stage,
                       void* stream);

/* Volume rendering integral, NeRFRenderer._raw2outputs (renderer.py:114-163).
 * noise[R,S] nullable (already scaled by raw_noise_std, :134-136); weights[R,S] nullable. */
int nerf_composite_fwd(const float* raw, const float* z_vals, const float* rays_d, const float* noise,
                       int R, int S, int white_bkgd, float* rgb_map, float* depth_map, float* acc_map,
                       float* weights, void* stream);

/* Analytic backward of the above w.r.t. raw.  d_depth / d_acc / d_weights nullable. */
int nerf_composite_bwd(const float* raw, const float* z_vals, const float* rays_d, const float* noise,
                       int R, int S, int white_bkgd, const float* d_rgb_map, const float* d_depth,
                       const float* d_acc, const float* d_weights, float* d_raw, void* stream);

/* Hierarchical sampling, NeRFRenderer._sample_pdf (renderer.py:165-199): pdf -> cdf ->
 * searchsorted(right) -> inverse-cdf lerp.
 *   bins[R,NB], weights[R,NB-1]  (row strides in floats given explicitly so that callers can pass
 *   views);  u: [N_imp] if u_shared else [R,N_imp];  samples[R,N_imp].
 *   inds (int64 [R,N_imp]) and cdf ([R,NB]) are optional check-mode exports. */
int nerf_sample_pdf(const float* bins, int64_t bins_stride, const float* weights, int64_t weights_stride,
                    const float* u, int u_shared, int R, int NB, int N_imp, float* samples,
                    int64_t* inds, float* cdf, void* stream);

/* Fused resampling step of _render_rays (renderer.py:86-90): z_mid, _sample_pdf on
 * weights[:,1:-1], and z_fine = sort(cat[z_coarse, z_samples]).  z_samples/inds/cdf nullable. */
int nerf_resample_merge(const float* z_coarse, const float* weights, const float* u, int u_shared,
                        int R, int S_c, int N_imp, float* z_fine, float* z_samples, int64_t* inds,
                        float* cdf, void* stream);

/* torch.optim.Adam step (defaults: no amsgrad / weight decay; scripts/train.py:258,387) on flat
 * fp32 buffers; grad_scale multiplies the gradient first (1/world_size after the all-reduce).
 * `step` is 1-based. */
int nerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                   double lr, double beta1, double beta2, double eps, int64_t step, float grad_scale,
                   void* stream);

/* mean((pred-target)^2) over [R,3] and its gradient 2(pred-target)/(3R) (scripts/train.py:376);
 * loss is a single device float (no host sync). d_pred nullable. */
int nerf_mse_loss(const float* pred, const float* target, int64_t n, float* loss, float* d_pred,
                  void* stream);

/* ---- training-step glue kept on the device ------------------------------
 * The reference reads loss / PSNR / gradient norm back every step (scripts/train.py:376-390:
 * loss.item(), calculate_psnr -> .cpu().numpy() :33-37, get_gradient_norm -> 24x .item() :60-67)
 * and computes the Adam scalars in Python (torch/optim/adam.py).  These two entry points keep all
 * of that in a device-resident state block of NERF_TRAIN_STATE_DOUBLES doubles, so that one
 * training step is a fixed launch sequence (capturable as a CUDA graph) with no host sync:
 *   state[0..5]  = lr, beta1, beta2, eps, grad_scale, step      (written by the host; step 0-based)
 *   state[10..12] = loss, psnr (10 log10(1/loss), data_range 1), grad_norm (|grad_scale * g|_2)
 * nerf_train_prepare: step += 1, derives that step's Adam scalars, fills the metrics
 *   (loss nullable -> metrics 0/inf);  deterministic (block-ordered fp64 reduction).
 * nerf_adam_step_dev: nerf_adam_step with every scalar read from the prepared state block. */
#define NERF_TRAIN_STATE_DOUBLES 96
int nerf_train_prepare(double* state, const float* loss, const float* flat_grads, int64_t n, void* stream);
int nerf_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const double* state, void* stream);
/* nerf_train_prepare + nerf_adam_step_dev as ONE launch (same arithmetic, same state block, same metrics);
 * `scratch`: nerf_adam_fused_scratch_bytes(n) bytes, zero-initialised once by the caller. */
size_t nerf_adam_fused_scratch_bytes(int64_t n);
int nerf_adam_step_fused(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                         double* state, const float* loss, void* scratch, void* stream);
/* Data-parallel variant: gradient exchange + nerf_adam_step_fused in ONE kernel over NVLink peer memory (replaces
 * torch.distributed.all_reduce(flat_grad) + the update; reference: scripts/train.py has no multi-GPU path, the
 * data-parallel contract is SURVEY.md section 8e).  peer_grads[r] / peer_flags[r] (HOST arrays of `world` DEVICE
 * pointers, r = 0 .. world-1, r == rank is this process) point into every rank's symmetric (peer-mapped) block:
 * its flat fp32 gradient (n floats, 16-byte aligned) and 2 * NERF_PEER_MAX + 1 u32 flags, zeroed once by the caller
 * before the first step (with a host barrier after the zeroing).  Every rank sums the `world` gradients in rank
 * order, so all ranks compute bit-identical parameters; the gradient buffers keep the rank-LOCAL gradients.  The
 * kernel returns only when no peer reads this rank's gradient any more.  `scratch` as for nerf_adam_step_fused
 * (sized for n); state[4] (grad_scale) is 1 / world.
 * peer_red == NULL: one-shot (every rank reads all `world` gradients).  peer_red[r] != NULL (n floats in every rank's
 * symmetric block, 16-byte aligned): two-shot -- rank r reduces slice r and stores it into every rank's `red` buffer,
 * then every rank updates from its own copy; same results on every rank, (world-1)/world x 2n floats of NVLink traffic
 * per rank instead of (world-1) x n.  The flag block needs 2 * NERF_PEER_MAX + 2 u32 then. */
#define NERF_PEER_MAX 8
int nerf_adam_step_fused_peer(float* params, const float* const* peer_grads, float* const* peer_red,
                              uint32_t* const* peer_flags, int rank, int world, float* exp_avg, float* exp_avg_sq,
                              int64_t n, double* state, const float* loss, void* scratch, void* stream);

/* The fine-pass compositing of a training step in ONE launch (scripts/train.py:374-382 around
 * renderer.py:106-107): nerf_composite_fwd -> rgb/depth/acc maps; loss = mean((rgb_map - target)^2) (:376) and
 * its gradient 2 (rgb_map - target) / (3R); nerf_composite_bwd of that gradient -> d_raw[R,S,4]; and, because
 * it is the last launch before the weight-gradient kernels, optimizer.zero_grad(): zero_buf[zero_n] (nullable)
 * is cleared.  `scratch`: nerf_composite_train_scratch_bytes(R) bytes, zero-initialised once by the caller.
 * Bit-identical to the three separate entry points (loss: same fp64 sum, block-ordered). */
size_t nerf_composite_train_scratch_bytes(int R);
int nerf_composite_train(const float* raw, const float* z_vals, const float* rays_d, const float* noise, int R,
                         int S, int white_bkgd, const float* target, float* rgb_map, float* depth_map,
                         float* acc_map, float* d_raw, float* loss, void* scratch, float* zero_buf,
                         int64_t zero_n, void* stream);

/* ---- the callers either side of the path, on the device ----------
 * nerf_generate_rays: rays (and optionally target colours) of n flat ray ids
 *   id = img*H*W + j*W + i  (the row order of NeRFDataset.all_rays_*, nerfmlp/data.py:76-97):
 *   rays_d = [(i-W/2)/focal, -(j-H/2)/focal, -1] @ pose[:3,:3].T in float64, rounded once to float32
 *   (data.py:80,86 followed by .float() at :101); rays_o = pose[:3,3] (:87).
 *   idx != NULL: gather of a training batch (replaces Dataset.__getitem__ data.py:99-104 + DataLoader
 *   collate, scripts/train.py:219,368-371); idx == NULL: the contiguous ids first..first+n-1, e.g. all
 *   pixels of one view for rendering (scripts/render_example.py:245-250), or one rank's shard of it.
 *   Target colour (rgb_out nullable): from rgb_lin[n_poses,H,W,3] (already preprocessed linear RGB) or
 *   from raw rgba[n_poses,H,W,4] bytes with the reference's preprocessing fused: /255, white-background
 *   alpha composite in float64 (data.py:47-55), sRGB->linear in float32 (:8-22,62).
 *   poses: [n_poses,4,4] row-major float32 on the device. */
int nerf_generate_rays(const float* poses, int n_poses, int H, int W, double focal, const int64_t* idx,
                       int64_t first, int64_t n, float* rays_o, float* rays_d, const uint8_t* rgba,
                       const float* rgb_lin, int white_bkgd, float* rgb_out, void* stream);

/* Output post-processing of scripts/render_example.py:256-271 on n floats: * brightness,
 * optional linear->sRGB (:12-26, float32), clip to [0,1], * 255, truncate to uint8. */
int nerf_postprocess_rgb8(const float* rgb, int64_t n, float brightness, int to_srgb, uint8_t* out,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H */
