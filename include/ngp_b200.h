/*
 * ngp_b200.h — C ABI of the B200-native (sm_100a) torch-ngp hot path.
 *
 * Plain pointers + sizes + a CUDA stream; no torch types.  All pointers are DEVICE
 * pointers unless the name ends in _host.  Every function is asynchronous on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream) and returns 0 on
 * success or a non-zero NGP_E* code (message via ngp_last_error()).  The callee never
 * allocates, frees or retains caller memory: the caller owns every buffer, exactly like
 * the reference's pybind ABI (SURVEY §8b "Ownership").
 *
 * Each entry point replaces one function of the reference's native extension tables:
 *   gridencoder/src/gridencoder.h:12-15   (bindings.cpp:6-8)
 *   ffmlp/src/ffmlp.h:8-14                (bindings.cpp:6-10)
 *   shencoder/src/shencoder.h:9-10        (bindings.cpp:6-7)
 *   raymarching/src/raymarching.h:7-18    (bindings.cpp:7-18)
 * Argument order and meaning follow those declarations; at::Tensor arguments become raw
 * pointers, at::optional<Tensor> becomes a nullable pointer, and a trailing stream is
 * added (the reference launches on the legacy default stream only).
 */
#ifndef NGP_B200_H
#define NGP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGP_OK            0
#define NGP_EINVAL        1   /* unsupported template value / bad argument (reference: std::runtime_error) */
#define NGP_ECUDA         2   /* CUDA launch / runtime error (reference: unchecked)                          */
#define NGP_EUNSUPPORTED  3   /* feature outside the built configuration                                    */

/* dtype codes for `void*` tensor arguments */
#define NGP_F32 0
#define NGP_F16 1

typedef void* ngp_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------ */
const char* ngp_last_error(void);        /* thread-local message of the last failing call       */
int         ngp_version(void);           /* ABI version (this header: 1)                        */
const char* ngp_build_arch(void);        /* "sm_100a"                                           */
/* number of kernels launched by this library since load / last reset (bench `gpu_launches`)    */
uint64_t    ngp_launch_count(void);
void        ngp_reset_launch_count(void);

/* ---- gridencoder  (gridencoder/src/gridencoder.h:12-15) ------------------------------------- */
/* inputs [B,D] f32 in [0,1]; embeddings [sO,C] (dtype); offsets [L+1] i32;
 * outputs: level_major!=0 -> [L,B,C] (reference-native layout, gridencoder.cu:388)
 *          level_major==0 -> [B,L*C] (what grid.py:57 permutes to; the permute copy is fused away)
 * dy_dx (nullable) [B,L,D,C].  S = log2(per_level_scale), H = base resolution. */
int ngp_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                            void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                            float S, uint32_t H, void* dy_dx, uint32_t gridtype,
                            int align_corners, uint32_t interp, int dtype, int level_major,
                            ngp_stream_t stream);
/* grad: level_major!=0 -> [L,B,C], else [B,L*C].  grad_embeddings [sO,C] must be zeroed by the
 * caller (grid.py:77) and is accumulated with scatter-add.  grad_inputs (nullable) [B,D] (dtype),
 * requires dy_dx. */
int ngp_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                             const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D,
                             uint32_t C, uint32_t L, float S, uint32_t H, const void* dy_dx,
                             void* grad_inputs, uint32_t gridtype, int align_corners,
                             uint32_t interp, int dtype, int level_major, ngp_stream_t stream);
/* gridencoder.h:15 — inputs [B,D] (dtype), adds weight * TV-gradient into `grad` in place. */
int ngp_grad_total_variation(const void* inputs, const void* embeddings, void* grad,
                             const int32_t* offsets, float weight, uint32_t B, uint32_t D,
                             uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                             int align_corners, int dtype, ngp_stream_t stream);

/* ---- shencoder  (shencoder/src/shencoder.h:9-10) -------------------------------------------- */
/* inputs [B,3] f32; outputs [B,degree^2] f32; dy_dx (nullable) [B,3,degree^2] f32. */
int ngp_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D,
                          uint32_t degree, float* dy_dx, ngp_stream_t stream);
int ngp_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D,
                           uint32_t degree, const float* dy_dx, float* grad_inputs,
                           ngp_stream_t stream);

/* ---- ffmlp  (ffmlp/src/ffmlp.h:8-14) ------------------------------------------------------- */
/* All tensors fp16, row-major.  inputs [B,input_dim]; weights = [hidden,input_dim] then
 * (num_layers-1) x [hidden,hidden] then [output_dim,hidden], each [out,in] row-major
 * (ffmlp.cu:631-634).  Any B >= 0 (the ragged last 128-row tile is masked in-kernel; the reference pads, ffmlp.py:157-159).  hidden_dim == 64,
 * input_dim in {16,32,48,64}, output_dim == 16 in this build; activation codes as
 * ffmlp.py:89-96 (0 = ReLU ... 6 = None); output_activation must be 6 (ffmlp.py:108).
 * forward_buffer [num_layers,B,hidden] receives every hidden activation (training). */
int ngp_ffmlp_forward(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim,
                      uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                      uint32_t activation, uint32_t output_activation, void* forward_buffer,
                      void* outputs, ngp_stream_t stream);
/* inference_buffer is accepted for ABI parity (ffmlp.h:10) and never touched: activations stay
 * in shared/tensor memory. */
int ngp_ffmlp_inference(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim,
                        uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                        uint32_t activation, uint32_t output_activation, void* inference_buffer,
                        void* outputs, ngp_stream_t stream);
/* grad [B,output_dim]; backward_buffer [num_layers,B,hidden] scratch, NULLABLE when num_layers <= 5 (written: dL/d pre-activation,
 * deepest layer first, as ffmlp.cu:749-895); grad_inputs [B,input_dim] iff calc_grad_inputs;
 * grad_weights same layout as weights (fp16, overwritten).  workspace: at least
 * ngp_ffmlp_backward_workspace_bytes() bytes of device scratch (fp32 partial weight grads). */
size_t ngp_ffmlp_backward_workspace_bytes(uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                          uint32_t hidden_dim, uint32_t num_layers);
int ngp_ffmlp_backward(const void* grad, const void* inputs, const void* weights,
                       const void* forward_buffer, uint32_t B, uint32_t input_dim,
                       uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                       uint32_t activation, uint32_t output_activation, int calc_grad_inputs,
                       void* backward_buffer, void* grad_inputs, void* grad_weights,
                       void* workspace, size_t workspace_bytes, ngp_stream_t stream);
/* Pipelined variant of ngp_ffmlp_backward (extension): a caller that splits one batch into row chunks — to overlap this kernel
 * with the hash-grid scatter of the previous chunk on another stream — accumulates all chunks into ONE fp32 workspace. */
#define NGP_WGRAD_ACCUMULATE  1u  /* the workspace already holds partial sums: do not zero it                              */
#define NGP_WGRAD_NO_FINALIZE 2u  /* leave the fp32 sums in the workspace (grad_weights may be NULL)                        */
int ngp_ffmlp_backward_ex(const void* grad, const void* inputs, const void* weights,
                          const void* forward_buffer, uint32_t B, uint32_t input_dim,
                          uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                          uint32_t activation, uint32_t output_activation, int calc_grad_inputs,
                          void* backward_buffer, void* grad_inputs, void* grad_weights,
                          void* workspace, size_t workspace_bytes, uint32_t flags, ngp_stream_t stream);
/* zero_first > 0: clear the n_params-float workspace (start of a chunked pass); 0: convert it to fp16 grad_weights; < 0: convert and
 * leave the workspace cleared (a caller that keeps one persistent workspace then needs no memset per step). */
int ngp_ffmlp_wgrad_finalize(void* workspace, void* grad_weights, uint32_t n_params, int zero_first,
                             ngp_stream_t stream);
/* ffmlp.h:13-14 — the reference (re)creates global side streams here; this build needs none. */
int ngp_ffmlp_allocate_splitk(size_t size);
int ngp_ffmlp_free_splitk(void);

/* ---- raymarching  (raymarching/src/raymarching.h:7-18) -------------------------------------- */
int ngp_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                           float min_near, float* nears, float* fars, ngp_stream_t stream);
int ngp_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N,
                     float* coords, ngp_stream_t stream);
int ngp_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, ngp_stream_t stream);
int ngp_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, ngp_stream_t stream);
int ngp_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield,
                 ngp_stream_t stream);
/* rays [N,3] i32 = (ray id, offset, count); counter [2] i32 (points, rays) accumulated atomically.
 * xyzs/dirs [M,3], deltas [M,2] must be zero-initialised by the caller (raymarching.py:205-207). */
int ngp_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                         float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                         uint32_t M, const float* nears, const float* fars, float* xyzs,
                         float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                         const float* noises, ngp_stream_t stream);
int ngp_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                     const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                     float* weights_sum, float* depth, float* image,
                                     ngp_stream_t stream);
int ngp_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                      const float* sigmas, const float* rgbs, const float* deltas,
                                      const int32_t* rays, const float* weights_sum,
                                      const float* image, uint32_t M, uint32_t N, float T_thresh,
                                      float* grad_sigmas, float* grad_rgbs, ngp_stream_t stream);
/* Step-driver extensions (no reference counterpart).  _forward_mse = compositor + the training loss head in one launch: also forms
 * pred = image + (1 - weights_sum) * bg against target [N,3] (indexed like image), the squared error per row of `rays` (sqerr [N]) and
 * g_image [N,3] / g_ws [N] = d(loss * *scale)/d(image, weights_sum) for loss = inv_norm / 2 * sum(sqerr) (inv_norm = 2 / (3 R) for the mean
 * of nerf/utils.py:866); replaces the elementwise / reduction launches between ngp_composite_rays_train_forward and _backward.
 * step_counter_push: step_counter[ring] = counter; ring = (ring + 1) % 16; ++nsteps (renderer.py:281-283, on the device). */
int ngp_composite_rays_train_forward_mse(const float* sigmas, const float* rgbs, const float* deltas,
                                         const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                         const float* target, float bg, float inv_norm, const float* scale,
                                         float* weights_sum, float* depth, float* image, float* g_image,
                                         float* g_ws, float* sqerr, ngp_stream_t stream);
int ngp_step_counter_push(int32_t* ring, const int32_t* counter, int32_t* nsteps, int32_t* step_counter,
                          ngp_stream_t stream);
int ngp_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive,
                   const float* rays_t, const float* rays_o, const float* rays_d, float bound,
                   float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                   const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                   const float* noises, ngp_stream_t stream);
int ngp_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive,
                       float* rays_t, const float* sigmas, const float* rgbs, const float* deltas,
                       float* weights_sum, float* depth, float* image, ngp_stream_t stream);

/* ---- fused NeRF-field extensions (no reference counterpart: they fuse GridEncoder -> FFMLP -> trunc_exp and
 * SHEncoder -> cat -> FFMLP -> sigmoid of nerf/network_ff.py:51-74 so that encoder features, SH features and the
 * concatenated color input never exist in HBM).  Optional fast path; the reference-shaped ops above remain. ---- */
/* x01 [M,3] f32 in [0,1]; table fp16; writes h_out [M,16] fp16 (sigma-net output, nullable), sigma_out [M] f32 = exp(h[:,0])
 * (nullable), and when train != 0 the stashes feat_out [M,2L] fp16 and forward_buffer [num_layers,M,64] fp16. */
int ngp_field_sigma_forward(const float* x01, const void* table_f16, const int32_t* offsets, uint32_t L, float S,
                            uint32_t H, uint32_t gridtype, int align_corners, const void* weights,
                            uint32_t num_layers, uint32_t M, int train, void* feat_out, void* forward_buffer,
                            void* h_out, float* sigma_out, ngp_stream_t stream);
/* dirs [M,3] f32, h_sigma [M,16] fp16 -> rgb_out [M,3] f32 = sigmoid(color_net([SH4(dir) | h_sigma[:,1:] | 0])[:, :3]) */
int ngp_field_color_forward(const float* dirs, const void* h_sigma, const void* weights, uint32_t num_layers,
                            uint32_t M, int train, void* forward_buffer, float* rgb_out, ngp_stream_t stream);
/* General form: pad (nullable) [M] fp16 = last input column of the color net (NULL = the zero pad of network_ff.py:67);
 * rgb_out NULL -> plain FFMLP output (pre-sigmoid) into h_out [M,16] fp16 instead (what ffmlp.FFMLP returns: the drop-in modules
 * use this when they recognise the SHEncoder -> cat -> FFMLP pattern). */
int ngp_field_color_forward_ex(const float* dirs, const void* h_sigma, const void* pad, const void* weights,
                               uint32_t num_layers, uint32_t M, int train, void* forward_buffer, float* rgb_out,
                               void* h_out, ngp_stream_t stream);
/* color-net backward incl. sigmoid / cat / trunc_exp gradients: writes dys_out [M,16] fp16 = dL/d(sigma-net output)
 * and grad_weights (color net).  The sigma net then uses ngp_ffmlp_backward(grad = dys_out, inputs = feat). */
int ngp_field_color_backward(const float* d_rgb, const float* rgb, const float* d_sigma, const void* h_sigma,
                             const float* dirs, const void* weights, const void* forward_buffer,
                             uint32_t num_layers, uint32_t M, void* dys_out, void* grad_weights, void* workspace,
                             size_t workspace_bytes, ngp_stream_t stream);

/* ---- device-driven inference loop (extension).  The reference's eval loop (nerf/renderer.py:337-367) sizes every iteration on the
 * host from `rays_alive[rays_alive >= 0]` (a boolean-index copy + synchronisation per iteration).  Here the loop state is a device
 * control block ctrl[8] u32 = {n_alive, n_step = max(min(N / n_alive, 8), 1), M = n_alive * n_step padded to 128, samples marched so
 * far, scratch, ...}; every kernel is launched for the worst case (N rays / M_max rows) and reads its real extent from ctrl, so an
 * iteration needs no host round trip and a block of iterations can be captured in a CUDA graph.  Per-ray arithmetic, n_step schedule
 * and termination rule are those of ngp_march_rays / ngp_composite_rays; only the order of the compacted ray list differs. ---- */
int ngp_infer_init(uint32_t N, int32_t* rays_alive, uint32_t* ctrl, ngp_stream_t stream);
/* as ngp_march_rays with (n_alive, n_step) = ctrl[0..1]; writes a delta == 0 sentinel after a ray's last sample instead of requiring
 * zero-filled buffers; noises (nullable) are applied in the first iteration only (renderer.py:353). */
int ngp_march_rays_dev(const uint32_t* ctrl, uint32_t N, const int32_t* rays_alive, const float* rays_t,
                       const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                       uint32_t C, uint32_t H, const uint8_t* grid, const float* fars, float* xyzs, float* dirs,
                       float* deltas, const float* noises, ngp_stream_t stream);
int ngp_composite_rays_dev(const uint32_t* ctrl, uint32_t N, float T_thresh, int32_t* rays_alive, float* rays_t,
                           const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                           float* depth, float* image, ngp_stream_t stream);
/* survivors (rays_in[n] >= 0) -> rays_out, then ctrl for the next iteration (n_alive = 0 once max_steps samples were marched) */
int ngp_compact_rays_dev(uint32_t* ctrl, uint32_t N, uint32_t max_steps, const int32_t* rays_in, int32_t* rays_out,
                         ngp_stream_t stream);
/* fused field kernels reading their row count on the device (rows = min(M_max, *rows_dev)); xyz in WORLD coordinates */
int ngp_field_sigma_forward_dev(const float* xyz, float bound, const uint32_t* rows_dev, const void* table_f16,
                                const int32_t* offsets, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                int align_corners, const void* weights, uint32_t num_layers, uint32_t M_max, void* h_out,
                                float* sigma_out, ngp_stream_t stream);
int ngp_field_color_forward_dev(const float* dirs, const void* h_sigma, const uint32_t* rows_dev, const void* weights,
                                uint32_t num_layers, uint32_t M_max, float* rgb_out, ngp_stream_t stream);

/* General form: the output gradient either as (d_rgb, rgb) [sigmoid gradient formed in the kernel] or as grad_h [M,3] fp16 =
 * dL/d(network output); d_sigma nullable (column 0 of dys_out is then 0: the caller's autograd handles trunc_exp); pad as in the
 * forward; flags = NGP_WGRAD_* as ngp_ffmlp_backward_ex. */
int ngp_field_color_backward_ex(const float* d_rgb, const float* rgb, const void* grad_h, const float* d_sigma,
                                const void* h_sigma, const float* dirs, const void* pad, const void* weights,
                                const void* forward_buffer, uint32_t num_layers, uint32_t M, void* dys_out,
                                void* grad_weights, void* workspace, size_t workspace_bytes, uint32_t flags,
                                ngp_stream_t stream);

/* ---- fused optimizer step (replaces GradScaler.unscale_/inf-check + torch.optim.Adam +
 * the per-forward fp32->fp16 table cast + gradient zeroing of nerf/utils.py:866-868, main_nerf.py:132, grid.py:43-44).
 * `state` = 8 device words {float scale, int growth_tracker, int found_inf, int step, float lr_scale, 3 reserved}; the step size
 * is lr * lr_scale (lr_scale lives on the device so that a schedule can change it between replays of a captured CUDA graph);
 * no host synchronisation. ---- */
int ngp_optim_check_finite(const void* grads, int dtype, uint64_t n, void* state, ngp_stream_t stream);
int ngp_optim_adam_step(float* params, float* exp_avg, float* exp_avg_sq, void* grads, int dtype, void* shadow_f16,
                        uint64_t n, float lr, float beta1, float beta2, float eps, const void* state, int zero_grad,
                        ngp_stream_t stream);
int ngp_optim_scaler_update(void* state, float growth, float backoff, int growth_interval, ngp_stream_t stream);
/* ---- gradient exchange over NVLink peer memory, fused with the optimizer (the step's single exchange; the
 * baseline it replaces is an NCCL all-reduce of the fp16 gradient bucket followed by ngp_optim_* over all parameters on every rank).
 * One process per GPU; every rank allocates one peer-visible block with ngp_peer_alloc, ships the 64-byte CUDA IPC handle to the other
 * ranks of the node and maps theirs with ngp_peer_open.  `*_host` arguments are HOST arrays of `world` device pointers (entry r = rank
 * r's buffer as mapped in this process).  The flat parameter space is split into `world` shards (multiples of 8 elements):
 *   ngp_exchange_barrier(slot 0)                      all local gradient buckets are complete
 *   ngp_exchange_reduce(lo, count = my shard)         my bucket[shard] = fp16(sum over ranks, fp32 accumulate); inf/nan -> state.found_inf
 *   ngp_exchange_barrier(slot 1, &found_inf, &found_inf)   everybody has read my bucket; found_inf OR-ed over ranks (GradScaler skip rule)
 *   ngp_exchange_adam(...) per parameter piece        Adam on my shard's fp32 masters / moments; the updated fp16 operand copy is
 *                                                     stored into EVERY rank's shadow (the all-gather half of the exchange)
 *   ngp_exchange_zero                                 clear my whole bucket for the next step
 *   ngp_exchange_barrier(slot 2)                      all shadows complete
 *   ngp_optim_scaler_update
 * Barriers are one-warp kernels on monotonic device-resident epochs (graph-replay safe); a peer that does not arrive within
 * timeout_ms (0 = 2000) raises the pad's error word (ngp_exchange_error) instead of hanging the device. ---- */
int    ngp_peer_alloc(size_t bytes, void** ptr_out, void* handle_out_host /* 64 bytes */);
int    ngp_peer_open(const void* handle_host /* 64 bytes */, void** ptr_out);
int    ngp_peer_close(void* ptr);
int    ngp_peer_free(void* ptr);
size_t ngp_exchange_pad_bytes(void);           /* size of a signal pad (zero-initialised peer memory) */
int    ngp_exchange_error(const void* my_pad, uint32_t* error_out_host);
int    ngp_exchange_barrier(void* const* pads_host, uint32_t rank, uint32_t world, uint32_t slot, const int32_t* flag_in,
                            int32_t* flag_out, uint32_t timeout_ms, ngp_stream_t stream);
int    ngp_exchange_reduce(void* const* sinks_host, uint32_t rank, uint32_t world, uint64_t lo, uint64_t count, void* state,
                           ngp_stream_t stream);
/* params = fp32 master of ONE parameter tensor whose element 0 has flat index seg_off; exp_avg / exp_avg_sq are the flat moment
 * arrays; [lo, lo + count) = the piece of this rank's shard that lies inside that tensor. */
int    ngp_exchange_adam(float* params, float* exp_avg_flat, float* exp_avg_sq_flat, void* my_sink, void* const* shadows_host,
                         uint32_t world, uint64_t seg_off, uint64_t lo, uint64_t count, float lr, float beta1, float beta2,
                         float eps, const void* state, ngp_stream_t stream);
int    ngp_exchange_zero(void* my_sink, uint64_t n, ngp_stream_t stream);
/* Fused form of the same exchange (what the optimizer uses): the flag barriers ride inside the data kernels, three launches per step.
 *   ngp_exchange_reduce_fused   signals "my bucket is complete", every CTA waits for all ranks, then reduces [lo, lo + count)
 *   ngp_exchange_adam_fused     exchanges the non-finite flags, then Adam on the (<= 4) parameter pieces of the shard [shard_lo, shard_hi)
 *                               (params_host[i] = fp32 master of the tensor whose element 0 has flat index seg_off_host[i]; piece =
 *                               flat [lo_host[i], lo_host[i] + count_host[i])), stores the fp16 operand copies into every rank's shadow,
 *                               clears the whole local bucket (n_total elements) and signals "my stores are done"
 *   ngp_exchange_finish         waits for every rank's "done", then the GradScaler update (as ngp_optim_scaler_update) */
/* mc_sink / mc_shadow (nullable): the same buckets / shadows addressed through ONE NVSwitch multicast mapping of all ranks' blocks (NVLS).
 * When given, the reduction is a single multimem.ld_reduce per 16 bytes (summed inside the switch, fp32 accumulation) and the operand copies
 * are a single multimem.st per 16 bytes (delivered to every replica) instead of `world` peer loads / stores. */
int    ngp_exchange_reduce_fused(void* const* pads_host, void* const* sinks_host, const void* mc_sink, uint32_t rank, uint32_t world,
                                 uint64_t lo, uint64_t count, void* state, uint32_t timeout_ms, ngp_stream_t stream);
int    ngp_exchange_adam_fused(void* const* pads_host, void* const* shadows_host, void* mc_shadow, uint32_t rank, uint32_t world,
                               float* const* params_host, const uint64_t* seg_off_host, const uint64_t* lo_host,
                               const uint64_t* count_host, uint32_t n_pieces, float* exp_avg_flat, float* exp_avg_sq_flat,
                               void* my_sink, uint64_t shard_lo, uint64_t shard_hi, uint64_t n_total, float lr, float beta1,
                               float beta2, float eps, void* state, uint32_t timeout_ms, ngp_stream_t stream);
int    ngp_exchange_finish(void* const* pads_host, uint32_t rank, uint32_t world, void* state, float growth, float backoff,
                           int growth_interval, uint32_t timeout_ms, ngp_stream_t stream);
/* ---- occupancy-grid maintenance (replaces the Python/torch op sequences of
 * nerf/renderer.py:380-442 mark_untrained_grid and :445-538 update_extra_state).  density_grid is float [C, H^3] in Morton
 * order, bitfield uint8 [C*H^3/8] (the marcher's format, raymarching.cu:279-288).  All random numbers are caller-provided
 * device arrays (uniform [0,1) floats / integer cell coordinates), so every entry point is a deterministic function;
 * nothing synchronises with the host. ---- */
/* poses [B,4,4] f32 camera-to-world; cells covered by no camera get density -1 (renderer.py:440).  count_out [C,H^3] u32
 * (nullable) = number of covering cameras, n_marked [1] u32 (nullable) += number of cells marked. */
int ngp_density_grid_mark_untrained(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, float bound,
                                    uint32_t C, uint32_t H, float* density_grid, uint32_t* count_out, uint32_t* n_marked,
                                    ngp_stream_t stream);
/* occ_list [C,H^3] u32: ascending Morton indices of cells with density > 0 (torch.nonzero, renderer.py:495), occ_count [C]. */
size_t ngp_density_grid_occupied_scratch_bytes(uint32_t C, uint32_t H);
int ngp_density_grid_occupied(const float* density_grid, uint32_t C, uint32_t H, uint32_t* occ_list, uint32_t* occ_count,
                              void* scratch, ngp_stream_t stream);
/* full update (renderer.py:456-483): xyzs [C,H^3,3] = jittered world position of every cell, row = Morton index.
 * noise [C,H^3,3] (nullable = cell centres) is indexed by the cell's (x*H+y)*H+z, the reference's meshgrid order. */
int ngp_density_grid_sample_full(uint32_t C, uint32_t H, float bound, const float* noise, float* xyzs, ngp_stream_t stream);
/* partial update (renderer.py:487-509): per cascade N cells at coords_rand [C,N,3] i32 followed by N cells picked from the
 * occupied list — by occ_pick_idx [C,N] i64 (the reference's torch.randint(0, Nz)) or, when that is NULL, by
 * floor(occ_pick_u [C,N] * Nz) with Nz read on the device.  noise [C,2N,3]; xyzs [C,2N,3]; indices [C,2N] u32 (0xffffffff
 * = no sample: empty occupied list or out-of-range input). */
int ngp_density_grid_sample_partial(uint32_t C, uint32_t H, float bound, uint32_t N, const int32_t* coords_rand,
                                    const int64_t* occ_pick_idx, const float* occ_pick_u, const uint32_t* occ_list,
                                    const uint32_t* occ_count, const float* noise, float* xyzs, uint32_t* indices,
                                    ngp_stream_t stream);
/* renderer.py:480-482,511-530: tmp = -1; tmp[cas, indices] = sigmas * density_scale (duplicates keep the largest);
 * grid = max(grid*decay, tmp) where both >= 0; state[0] = mean(clamp(grid, 0)); state[1] = min(state[0], density_thresh);
 * bitfield = packbits(grid, state[1]).  sigmas [C,N] f32, indices [C,N] u32 (NULL = identity, N == H^3); tmp_grid [C,H^3]
 * and scratch (ngp_density_grid_update_scratch_bytes) are workspaces; state [2] f32 stays on the device. */
size_t ngp_density_grid_update_scratch_bytes(uint32_t C, uint32_t H);
int ngp_density_grid_update(float* density_grid, float* tmp_grid, const uint32_t* indices, const float* sigmas, uint32_t N,
                            float density_scale, float decay, float density_thresh, uint32_t C, uint32_t H,
                            uint8_t* bitfield, float* state, void* scratch, ngp_stream_t stream);
/* ---- ray generation + training-pixel gather (replaces the torch op sequences of
 * nerf/utils.py:54-137 get_rays, nerf/provider.py:308-312 and nerf/utils.py:494-508). ---- */
/* poses [B,4,4] f32 camera-to-world; inds [*,N] i64 pixel indices (row stride inds_stride: 0 = one list shared by all
 * cameras; NULL = all H*W pixels in order, N == H*W); rays_o / rays_d [B,N,3] f32 (directions normalised). */
int ngp_get_rays(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, uint32_t N,
                 const int64_t* inds, uint32_t inds_stride, float* rays_o, float* rays_d, ngp_stream_t stream);
/* images [n_img,H,W,C] (dtype 0 = f32, 2 = u8 scaled by 1/255; C = 3 or 4), image_index [B] i64 (NULL = camera b uses image
 * b).  pixels_out [B,N,C] (nullable) = raw gather; gt_out [B,N,3] (nullable) = target colour: sRGB->linear when linear != 0,
 * then rgb*a + bg*(1-a) when C == 4, bg = bg_pixels [B,N,3] or, when that is NULL, bg_scalar. */
int ngp_gather_pixels(const void* images, int dtype, const int64_t* image_index, uint32_t H, uint32_t W, uint32_t C, uint32_t B,
                      uint32_t N, const int64_t* inds, uint32_t inds_stride, int linear, const float* bg_pixels,
                      float bg_scalar, float* pixels_out, float* gt_out, ngp_stream_t stream);
/* ---- frequency encoding (freqencoder/src/freqencoder.h:6-9; not on the --ff path, built for breadth) ----
 * inputs [B,D] f32 -> outputs [B,C] f32, C = D + 2*deg*D: [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] in D-wide blocks. */
int ngp_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs,
                            ngp_stream_t stream);
int ngp_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                             float* grad_inputs, ngp_stream_t stream);
/* test hooks (not part of the reference ABI) */
int ngp_grid_level_scales(float* out_device, uint32_t L, float S, uint32_t H, ngp_stream_t stream);
int ngp_debug_umma(const void* A, const void* Bm, float* D, int mode, ngp_stream_t stream);
/* L2 reduction-rate probe: blocks x 256 threads x ops_per_thread reductions at pseudo-random entries of table[entries] (f16x2 entries);
 * mode 0 = 4-byte f16x2, 1 = 8-byte v2.f16x2, 2 = 16-byte v4.f16x2.  The ceiling the table scatter is reported against. */
int ngp_debug_red_probe(void* table_f16x2, uint32_t entries, uint32_t blocks, uint32_t ops_per_thread, int mode, ngp_stream_t stream);
/* MLP backward kernel selection: 1 = two tile contexts per CTA with a deep activation ring (default, where it fits), 0 = the
 * single-context kernel */
int ngp_debug_set_mlp_backward(int dual);
/* gather variant of the fused encoder -> sigma kernel (measured experiments): -1 = environment (default 0), 0 = 4-byte loads,
 * 1 = aligned x-pair 8-byte loads, 2 = the first tma_levels table levels staged in shared memory by one cp.async.bulk per CTA */
int ngp_debug_set_sigma_gather(int variant, int tma_levels);

#ifdef __cplusplus
}
#endif
#endif /* NGP_B200_H */
