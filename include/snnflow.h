/*
 * snnflow.h - C ABI of the B200-native spiking-FireNet hot path (libsnnflow.so).
 *
 * This is the drop-in boundary: plain device pointers, sizes and a CUDA stream; no torch types.
 * Every buffer (inputs, outputs, workspace) is owned and allocated by the caller; the library never
 * allocates device memory, never synchronises the stream and keeps no mutable global state apart from
 * the thread-local last-error string.  All tensors are fp32, contiguous, NCHW unless stated.  All
 * functions return 0 on success and a negative SNNFLOW_E* code on failure (the message is available
 * from snnflow_last_error()).  There is no CPU fallback.
 *
 * The reference (LSquarzoni/SNN_Event-based_Optical_Flow) has no FFI for this path - it is pure
 * PyTorch - so each entry point cites the Python function it replaces (paths relative to the
 * reference root).  The reference's only native-plugin convention is the libtorch custom op of
 * ONNX_LIF_operator/src/lif_op.cpp:71-83; INTEGRATION.md shows the ctypes binding a maintainer adds.
 */
#ifndef SNNFLOW_H_
#define SNNFLOW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNNFLOW_ABI_VERSION 1

/* error codes */
#define SNNFLOW_OK 0
#define SNNFLOW_EINVAL (-1)   /* bad argument (null pointer, unsupported size) */
#define SNNFLOW_ECUDA (-2)    /* a CUDA runtime call / kernel launch failed    */
#define SNNFLOW_EWORKSPACE (-3) /* workspace too small                         */

/* flags for the ConvLIF entry points */
#define SNNFLOW_HARD_RESET 1u   /* v' = v*lam*(1-z) + (1-lam)*I ; else soft: v' = v*lam + (1-lam)*I - z*theta */
#define SNNFLOW_DETACH_RESET 2u /* reset path does not carry gradient (spiking_submodules.py:139-140)      */
#define SNNFLOW_NO_TENSOR_CORES 4u /* force the exact-fp32 CUDA-core convolution                           */
#define SNNFLOW_STATE_INTERNAL 16u /* snnflow_window_forward (save = 0) streaming mode: the layer states stay inside the arena in
                                     the engine's own layout between calls (see snnflow_window_export_state)             */
#define SNNFLOW_STREAM_PHASE 32u   /* streaming mode: ping-pong phase of the recurrent layers' membrane slots (caller toggles
                                     it by T & 1 after every call, starting from 0 after an import)                      */
#define SNNFLOW_REUSE_PACKED 64u   /* snnflow_window_forward: the weights have not changed since the previous call on this
                                     arena: skip the weight packing launch                                               */
#define SNNFLOW_INPUT_EXACT16 8u /* caller guarantees x holds spikes / small integers (exact in fp16 AND bf16):
                                    lets the backward use the tensor-core weight-gradient kernel               */

/* surrogate gradient kinds (models/spiking_util.py) */
#define SNNFLOW_SG_ARCTAN 0     /* 1/(1+w*u^2)      spiking_util.py:92 */
#define SNNFLOW_SG_SUPERSPIKE 1 /* 1/(1+w*|u|)^2    spiking_util.py:42 */
#define SNNFLOW_SG_TRIANGLE 2   /* relu(1-w*|u|)    spiking_util.py:78 */
#define SNNFLOW_SG_MULTIGAUSS 3 /* 1.15 G(u;0,w) - 0.15 G(u;w,6w) - 0.15 G(u;-w,6w)   spiking_util.py:46-65 (per-step engine only) */

typedef void* snnflow_stream_t; /* cudaStream_t */

int snnflow_abi_version(void);
const char* snnflow_last_error(void);
/* number of kernels this library has launched in the calling process (for bench accounting) */
uint64_t snnflow_launch_count(void);
/* Per-launch profiler: enable(1) clears old records and brackets every subsequent kernel launch with CUDA
 * events on its own stream; summary() synchronises the device and writes one text line per kernel name,
 * "name launches total_ms algorithmic_bytes algorithmic_flops".  Off by default (no events recorded). */
int snnflow_profile_enable(int on);
int snnflow_profile_summary(char* buf, size_t capacity);

/* ---------------------------------------------------------------------------------------------
 * ConvLIF / ConvLIFRecurrent forward, one layer-step.
 * Replaces ConvLIF.forward (models/spiking_submodules.py:121-151) and ConvLIFRecurrent.forward
 * (:265-300): 3x3 conv (stride 1, pad 1, no bias) of x with w_ff [+ 3x3 conv of z_in with w_rec],
 * leak, delayed reset, threshold, spike - fused in one kernel.
 *   x [B,Cin,H,W]; w_ff [C,Cin,3,3]; w_rec [C,C,3,3] or NULL (feed-forward cell)
 *   v_in, z_in [B,C,H,W] or both NULL (= zeros, :128-129)
 *   lam [C] = sigmoid(leak), theta [C] = clamp_min(thresh, 0.01)      (:133,:136; computed by caller)
 *   residual [B,C,H,W] or NULL; out [B,C,H,W] or NULL: out = z_out + residual (:151)
 *   v_out, z_out [B,C,H,W]; cur_out [B,C,H,W] or NULL: the input current I, saved for the backward
 * Rounding order follows the reference exactly (separately rounded mul/add, strict '>' compare).
 * --------------------------------------------------------------------------------------------- */
int snnflow_convlif_fwd(const float* x, const float* w_ff, const float* w_rec, const float* v_in,
                        const float* z_in, const float* lam, const float* theta, const float* residual,
                        float* v_out, float* z_out, float* out, float* cur_out, int B, int Cin, int C, int H,
                        int W, unsigned flags, snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core variant of the forward (tcgen05 implicit GEMM, accumulators in TMEM) for Cin, C multiples of
 * 16 up to 64.  Same semantics and outputs as snnflow_convlif_fwd.  Precondition: every element of x (and of
 * z_in) is exactly representable in fp16 - true for spikes {0,1}, spikes + residual and event counts; the
 * kernel counts violations (snnflow_tc_inexact_count).  Weights are passed pre-packed:
 *   packed_bytes(Cin, C, recurrent)  size of the packed blob (0 = shape not covered, use snnflow_convlif_fwd)
 *   pack(w_ff, w_rec|NULL, packed)   fp32 [C,Cin,3,3] (+ [C,C,3,3]) -> power-of-two scaled fp16 hi+lo split in
 *                                    the UMMA shared-memory layout; re-run whenever the weights change
 * The conv result equals the fp32 conv to ~2^-22 relative per weight (bit-exact for 2^-12-grid weights).
 * --------------------------------------------------------------------------------------------- */
size_t snnflow_convlif_packed_bytes(int Cin, int C, int recurrent);
int snnflow_convlif_pack(const float* w_ff, const float* w_rec, void* packed, int Cin, int C,
                         snnflow_stream_t stream);
int snnflow_convlif_fwd_tc(const float* x, const void* packed, int recurrent, const float* v_in,
                           const float* z_in, const float* lam, const float* theta, const float* residual,
                           float* v_out, float* z_out, float* out, float* cur_out, int B, int Cin, int C, int H,
                           int W, unsigned flags, snnflow_stream_t stream);
/* number of input elements the tensor-core path saw that were NOT fp16-exact (synchronises); reset != 0 clears */
unsigned int snnflow_tc_inexact_count(int reset);

/* ---------------------------------------------------------------------------------------------
 * ConvLIF / ConvLIFRecurrent backward, one layer-step of BPTT.
 * Replaces the autograd graph of the forward above incl. ArctanSpike/SuperSpike/TriangleSpike
 * .backward (models/spiking_util.py:38-43,74-79,88-93).  Saved tensors: x, v_in, z_in (may be NULL =
 * zeros), v_out, cur (from cur_out).  Incoming: g_out (grad of `out`, may be NULL), g_v_out, g_z_out
 * (grads of the returned state, may be NULL).  Outgoing: g_x (NULL = not needed), g_v_in, g_z_in
 * (NULL allowed when w_rec == NULL and DETACH_RESET).  dw_ff / dw_rec / dlam / dtheta are ACCUMULATED
 * into (+=), in a fixed order (run-to-run deterministic).  dlam/dtheta are w.r.t. the effective
 * lam/theta; the caller applies the sigmoid / clamp chain rule.
 * workspace: snnflow_convlif_bwd_workspace_bytes() bytes, 256-byte aligned.
 * --------------------------------------------------------------------------------------------- */
size_t snnflow_convlif_bwd_workspace_bytes(int B, int Cin, int C, int H, int W, int recurrent);
int snnflow_convlif_bwd(const float* x, const float* w_ff, const float* w_rec, const float* v_in,
                        const float* z_in, const float* v_out, const float* cur, const float* lam,
                        const float* theta, const float* g_out, const float* g_v_out, const float* g_z_out,
                        float* g_x, float* g_v_in, float* g_z_in, float* dw_ff, float* dw_rec, float* dlam,
                        float* dtheta, void* workspace, size_t workspace_bytes, int B, int Cin, int C, int H,
                        int W, unsigned flags, int surrogate, float act_width, snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Flow prediction head: flow = tanh(conv1x1(x, w) + b)   (models/submodules.py:96-113 with
 * kernel_size=1, activation="tanh"; LIFFireNet.pred, models/model.py:105-107,182).
 *   x [B,C,H,W]; w [2,C]; b [2]; flow [B,2,H,W]
 * Backward: g_x [B,C,H,W] (overwritten), dw [2,C] and db [2] accumulated (+=).
 * workspace for the backward: snnflow_pred_bwd_workspace_bytes().
 * --------------------------------------------------------------------------------------------- */
int snnflow_pred_fwd(const float* x, const float* w, const float* b, float* flow, int B, int C, int H, int W,
                     snnflow_stream_t stream);
size_t snnflow_pred_bwd_workspace_bytes(int B, int C, int H, int W);
int snnflow_pred_bwd(const float* x, const float* w, const float* flow, const float* g_flow, float* g_x,
                     float* dw, float* db, void* workspace, size_t workspace_bytes, int B, int C, int H, int W,
                     snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-network window: LIFFireNet / LIFFireFlowNet forward over T time bins and the matching BPTT, as ONE
 * host call each (no Python between the ~80 + ~300 kernel launches; capturable in a CUDA graph).
 * Replaces T calls of LIFFireNet.forward (models/model.py:172-182: head -> G1 -> R1a -> R1b -> G2 -> R2a -> R2b
 * -> 1x1 tanh pred) plus the autograd graph torch would build for them (train_flow.py:232,262).
 *
 * Layers are indexed 0..6 in that order; bit l of recurrent_mask marks a ConvLIFRecurrent (FireNet: 0x12).
 * acts: activation arena of snnflow_net_acts_floats() floats.  With save != 0 it keeps, for every layer l and bin
 *   t, the block [v | z | I] (3 * B*C*H*W floats) at offset ((l*T + t) * 3) * B*C*H*W: the [v | z] pair of (l, T-1)
 *   is the layer's state [2,B,C,H,W] after the window.  With save == 0 (inference) only two bins are kept
 *   (ping-pong) and I is not stored; the final state of layer l is at slot (T-1) % 2.
 * state_in: NULL (zero state) or 7 pointers to [2,B,C,H,W] states from the previous window.
 * input [T,B,num_bins,H,W]; flow [T,B,2,H,W].  exact_input != 0: the caller guarantees the input holds small
 *   integers (event counts), which lets layer 0's weight gradient use bf16 operands (it does not today).
 * Backward: g_flow [T,B,2,H,W] -> accumulates (+=) into the d* pointers of every layer and d_pred_w / d_pred_b;
 *   gradients w.r.t. state_in and input are not produced (the reference detaches states at window boundaries,
 *   train_flow.py:278, and the event counts need no gradient).  model.residual (default False) is not covered.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  int B, C, H, W, T, num_bins;
  unsigned recurrent_mask;
  unsigned flags;       /* SNNFLOW_HARD_RESET | SNNFLOW_DETACH_RESET | SNNFLOW_NO_TENSOR_CORES */
  int surrogate;
  float act_width;
} snnflow_net_desc;

typedef struct {
  const float *w_ff, *w_rec, *lam, *theta; /* parameters (lam/theta already sigmoid'ed / clamped)        */
  const void* packed;                      /* snnflow_convlif_pack blob or NULL (CUDA-core forward)      */
  float *dw_ff, *dw_rec, *dlam, *dtheta;   /* gradient accumulators (backward only; may be NULL forward) */
  /* snnflow_window_backward only, optional (NULL: off): the chain rule through lam = sigmoid(leak) and theta =
   * clamp_min(thresh, 0.01) (spiking_submodules.py:133,136) applied inside the final reduction launch,
   *   d_leak[c] += dlam_c * lam_c * (1 - lam_c),   d_thresh[c] += dtheta_c * (thresh_raw[c] >= 0.01),
   * so that the caller's gradient buffers of the RAW parameters are written directly (dlam / dtheta may then be NULL). */
  const float* thresh_raw;
  float *d_leak, *d_thresh;
} snnflow_layer_ptrs;

size_t snnflow_net_acts_floats(const snnflow_net_desc* d, int save);
size_t snnflow_net_bwd_workspace_bytes(const snnflow_net_desc* d);
int snnflow_net_forward(const snnflow_net_desc* d, const snnflow_layer_ptrs* layers, const float* pred_w,
                        const float* pred_b, const float* input, const float* const* state_in, float* acts,
                        float* flow, int save, snnflow_stream_t stream);
int snnflow_net_backward(const snnflow_net_desc* d, const snnflow_layer_ptrs* layers, const float* pred_w,
                         const float* input, const float* const* state_in, const float* acts, const float* flow,
                         const float* g_flow, float* d_pred_w, float* d_pred_b, void* workspace,
                         size_t workspace_bytes, snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Layer-major window engine: the same window (same arguments, same results) executed layer by layer instead of
 * bin by bin.  A layer at bin t depends only on the layer below at bin t and on its own state, so a feed-forward
 * ConvLIF runs all T bins in ONE launch with its membrane in registers, a ConvLIFRecurrent takes one launch per
 * bin, and the backward pass computes the data gradient and the weight gradient of a layer in one launch each over
 * all T*B images.  Spikes travel between layers as bf16 planes with a zero border (TMA-staged straight into the
 * tensor-core layout); weights are split into three bf16 terms (every product exact, fp32 accumulation), gradients
 * into bf16 hi + lo.  Covers C in {16, 32, 64}, num_bins <= 16 with bf16-exact input values (event counts),
 * SNNFLOW_DETACH_RESET set, feed-forward head; snnflow_window_supported() tells, everything else stays on
 * snnflow_net_forward / snnflow_net_backward.
 *
 * arena: snnflow_window_arena_bytes() bytes, 256-byte aligned, ZERO-FILLED ONCE by the caller when it is allocated
 *   (or when the descriptor changes): the plane borders are never written by the kernels and must read as zero.
 *   The state [2,B,C,H,W] of layer l after the window lives inside the arena at the byte offset reported by
 *   snnflow_window_state_offsets() (zero copy).  With save != 0 the arena also keeps what the backward needs,
 *   including a copy of state_in - so state_in[l] may be that very state block from the previous window (one arena,
 *   fixed addresses: the whole training step can be captured in a CUDA graph).  With save == 0 the state is updated
 *   in place when state_in[l] aliases the state block.  state_in of the backward call is only tested for NULL.
 * workspace (backward): snnflow_window_workspace_bytes() bytes, 256-byte aligned, zero-filled once as well.
 * state_in, layers, pred_*, input, flow, g_flow and the gradient accumulation semantics are those of
 * snnflow_net_forward / snnflow_net_backward (layers[l].packed is ignored: the engine packs per window).
 * snnflow_window_flags_offset: byte offset inside the arena of four uint32 status words owned by the engine;
 *   word 0 = number of input values seen so far (by windows run in THIS arena) that were not bf16-exact - sticky until
 *   the caller clears it.  The caller reads it (a device-to-host copy at a moment of its choosing) and may hand its
 *   address to snnflow_clip_adam as the update gate.
 * --------------------------------------------------------------------------------------------- */
int snnflow_window_supported(const snnflow_net_desc* d, int backward /* 0: forward (save == 0) only */);
size_t snnflow_window_arena_bytes(const snnflow_net_desc* d, int save);
size_t snnflow_window_workspace_bytes(const snnflow_net_desc* d);
int snnflow_window_state_offsets(const snnflow_net_desc* d, int save, size_t* offsets_bytes /* [7] */);
size_t snnflow_window_flags_offset(const snnflow_net_desc* d, int save);
int snnflow_window_forward(const snnflow_net_desc* d, const snnflow_layer_ptrs* layers, const float* pred_w,
                           const float* pred_b, const float* input, const float* const* state_in, void* arena,
                           float* flow, int save, snnflow_stream_t stream);
/* Streaming inference (flags & SNNFLOW_STATE_INTERNAL, save = 0): the reference's eval loop calls the network once per time
 * bin (eval_flow.py:220) and hands the 7 states [2,B,C,H,W] back in each time (models/model.py:172-182) - 16 B per neuron
 * and layer-step of fp32 NCHW traffic that the arithmetic does not need.  In this mode snnflow_window_forward ignores
 * state_in and continues from the state the previous call left INSIDE the arena in the engine's layout (membranes c8,
 * spikes = the layers' bf16 planes); the two calls below convert to / from the reference's layout only when somebody
 * looks (the network's `states` getter) or assigns (reset_states, a caller-provided state).  All calls of a stream use the
 * same descriptor (T included) and arena.
 *   import: state_in[l] [2,B,C,H,W] or NULL (zeros) -> arena; the next forward call must pass phase 0
 *   export: arena -> state_out[l] [2,B,C,H,W]; `phase` = the SNNFLOW_STREAM_PHASE bit the NEXT forward call would pass */
int snnflow_window_import_state(const snnflow_net_desc* d, const float* const* state_in, void* arena, snnflow_stream_t stream);
int snnflow_window_export_state(const snnflow_net_desc* d, const void* arena, int phase, float* const* state_out,
                                snnflow_stream_t stream);
int snnflow_window_backward(const snnflow_net_desc* d, const snnflow_layer_ptrs* layers, const float* pred_w,
                            const float* const* state_in, const void* arena, const float* flow, const float* g_flow,
                            float* d_pred_w, float* d_pred_b, void* workspace, size_t workspace_bytes,
                            snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Event encodings (dataloader/encodings.py).  xs, ys, ts, ps: [N] fp32 device arrays (integer-valued
 * coordinates, truncated like .long()); events outside the sensor are ignored.
 * encode_cnt   : events_to_channels (:70-85)  -> out [2,H,W]  per-polarity counts (exact integers)
 * encode_image : events_to_image (:30-45)     -> out [H,W];   accumulate=0 keeps the LAST event's value
 *                per pixel (CPU index_put_ order); scratch = H*W int32 (only for accumulate=0)
 * encode_voxel : events_to_voxel (:48-67)     -> out [nb,H,W]; deterministic: accumulated in 64-bit
 *                fixed point (2^-32 resolution); scratch = nb*H*W int64
 * All outputs are overwritten (zero-initialised inside).  batch variants take B independent windows
 * of equal length N laid out [B,N] -> out [B,...].
 * --------------------------------------------------------------------------------------------- */
int snnflow_encode_cnt(const float* xs, const float* ys, const float* ps, float* out, int64_t N, int B, int H,
                       int W, snnflow_stream_t stream);
int snnflow_encode_image(const float* xs, const float* ys, const float* ps, float* out, int32_t* scratch,
                         int64_t N, int H, int W, int accumulate, snnflow_stream_t stream);
int snnflow_encode_voxel(const float* xs, const float* ys, const float* ts, const float* ps, float* out,
                         int64_t* scratch, int64_t N, int num_bins, int H, int W, int round_ts,
                         snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Loader window formatter: one raw window of events per batch slot -> the tensors of one batch
 * item, the event branch of H5Loader.__getitem__ without the HDF5 file handling.  Replaces, per call:
 *   BaseDataLoader.event_formatting   dataloader/base.py:71-99    (fp32 cast, p*2-1, min-max normalised timestamps)
 *   BaseDataLoader.augment_events     dataloader/base.py:101-126  (flips[b] = {horizontal, vertical, polarity})
 *   create_{cnt,mask,voxel,list}_encoding, create_polarity_mask   dataloader/base.py:160-235
 *   create_hot_mask / get_hot_event_mask   dataloader/base.py:237-256, dataloader/encodings.py:88-103
 *   hot-pixel application, avg_pool2d down-sampling, event-list rescaling + clamp   dataloader/h5.py:323-331,375-410
 *   custom_collate                    dataloader/base.py:261-278  (lists come out as [B,N,4] / [B,N,2])
 * Inputs (device): xs, ys, ps [B,N] fp32 - sensor coordinates and RAW polarity in {0,1}; ts [B,N] fp32, or fp64 with
 * ts_is_f64 = 1 (t0 [B] fp64, may be NULL, is subtracted in fp64 before the fp32 cast like h5.py:129); flips [B,3]
 * int32 or NULL.  hot_events [B,H,W] fp32 / hot_idx [B] int32: the filter's running state (read and updated), required
 * when hot_enabled.  Outputs: event_cnt [B,2,h,w], event_voxel [B,num_bins,h,w], event_mask [B,1,h,w] with
 * h = H / pool_h, w = W / pool_w; event_list [B,N,4] = (ts, y, x, p); event_pol [B,N,2].  pool_h = pool_w = 1: no
 * down-sampling; otherwise the list coordinates are scaled by target/H, target/W and clamped to [0, target-1].
 * Counts, masks and lists are bit-identical to the reference; the voxel grid is accumulated in 64-bit fixed point.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B;            /* batch slots                                   */
  int32_t H, W;         /* encoding (sensor) resolution                  */
  int64_t N;            /* events per slot                               */
  int32_t num_bins;     /* voxel-grid bins                               */
  int32_t round_ts;     /* round_encoding (encodings.py:58-59)           */
  int32_t pool_h, pool_w;       /* original // target (h5.py:381-382)    */
  int32_t target_h, target_w;   /* loader.resolution (list rescaling)    */
  int32_t hot_enabled, hot_max_px, hot_min_obvs;
  float hot_max_rate;
} snnflow_loader_desc;
size_t snnflow_format_window_workspace_bytes(const snnflow_loader_desc* d);
int snnflow_format_window(const snnflow_loader_desc* d, const float* xs, const float* ys, const void* ts, int ts_is_f64,
                          const double* t0, const float* ps, const int32_t* flips, float* hot_events, int32_t* hot_idx,
                          float* event_cnt, float* event_voxel, float* event_mask, float* event_list, float* event_pol,
                          void* workspace, size_t workspace_bytes, snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Image of warped events (utils/iwe.py, loss/flow.py).
 * events [B,N,4] = (ts, y, x, p); flow [B,2,H,W] (channel 0 = x, 1 = y); ev_flow [B,N,2] = (fy, fx).
 *
 * flow_gather_fwd : per-event flow lookup of loss/flow.py:66-81 / utils/iwe.py:110-120; the flat
 *   index is formed in fp32 as y*W + x and truncated, exactly like the reference.
 * flow_gather_bwd : its adjoint, g_flow [B,2,H,W] += scatter(g_ev_flow)  (g_flow must be initialised).
 *
 * iwe_splat_fwd : get_interpolation (utils/iwe.py:20-71) + interpolate (:74-93) fused: warps every
 *   event to tref, computes the 4 bilinear corners (or the rounded location when round_idx != 0),
 *   purges out-of-range corners and accumulates into `n_img` images per sample:
 *     img 0: sum w * pol_mask[..,0]          img 1: sum w * pol_mask[..,1]
 *     img 2: sum w * tsw * pol_mask[..,0]    img 3: sum w * tsw * pol_mask[..,1]     (n_img == 4)
 *   with tsw = ts (ts_mode 1) or ts_ref - ts (ts_mode 2): the four images of one direction of the
 *   contrast loss (loss/flow.py:199-213, :232-246); n_img == 2 is compute_pol_iwe (utils/iwe.py:133-154).
 *   out [B,n_img,H,W] is overwritten.  Accumulation is in 64-bit fixed point (deterministic);
 *   scratch = B*n_img*H*W int64.
 * iwe_splat_bwd : gradient w.r.t. ev_flow given g_img [B,n_img,H,W], with the reference's autograd
 *   semantics (abs'(0) = 0; the max(0, 0) tie passes half the gradient).  g_ev_flow [B,N,2] overwritten.
 * --------------------------------------------------------------------------------------------- */
int snnflow_flow_gather_fwd(const float* flow, const float* events, float* ev_flow, int B, int64_t N, int H,
                            int W, snnflow_stream_t stream);
int snnflow_flow_gather_bwd(const float* g_ev_flow, const float* events, float* g_flow, int B, int64_t N, int H,
                            int W, snnflow_stream_t stream);
int snnflow_iwe_splat_fwd(const float* events, const float* ev_flow, const float* pol_mask, float* out,
                          int64_t* scratch, int B, int64_t N, int H, int W, float tref, float flow_scaling,
                          int n_img, int ts_mode, float ts_ref, int round_idx, snnflow_stream_t stream);
int snnflow_iwe_splat_bwd(const float* events, const float* ev_flow, const float* pol_mask, const float* g_img,
                          float* g_ev_flow, int B, int64_t N, int H, int W, float tref, float flow_scaling,
                          int n_img, int ts_mode, float ts_ref, snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Contrast-maximisation loss of one window and its gradient w.r.t. the flow maps, in one call.
 * Replaces T calls of EventWarping.event_flow_association (loss/flow.py:58-121), EventWarping.forward (:178-303:
 * forward- and backward-warped event images, per-pixel average timestamps, squared sums scaled by the number of
 * pixels with events, Charbonnier smoothness over dx, dy, both diagonals and dt) and the autograd graph behind them.
 *   flow [T,B,2,H,W] (channel 0 = x, 1 = y); events [T,B,N,4] = (ts in [0,1], y, x, p) per bin - the timestamp shift
 *   ts += bin index of loss/flow.py:91 is applied internally, the input is not modified; pol_mask [T,B,N,2];
 *   event_mask [T,B,1,H,W] or NULL (config model.mask_output: masks the smoothness terms)
 *   loss [1]; g_flow [T,B,2,H,W] = d loss / d flow (overwritten); workspace: snnflow_window_loss_workspace_bytes()
 * The per-event gradients carry the reference's autograd semantics (abs'(0) = 0, max(0,0) passes half, zero-count
 * pixels keep the gradient path of the non-zero-pixel normaliser).  Reductions run in a fixed order; the scatter of
 * per-event gradients onto the flow maps uses fp32 atomics.
 * --------------------------------------------------------------------------------------------- */
size_t snnflow_window_loss_workspace_bytes(int T, int B, int64_t N, int H, int W);
int snnflow_window_loss(const float* flow, const float* events, const float* pol_mask, const float* event_mask, float* loss,
                        float* g_flow, void* workspace, size_t workspace_bytes, int T, int B, int64_t N, int H, int W,
                        float flow_scaling, float regul_weight, int loss_scaling, snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * snntorch-style Leaky neuron step: the LIF of SNNtorch_ConvLIF / SNNtorch_ConvLIFRecurrent
 * (models/SNNtorch_spiking_submodules.py:124-322, :324-567 - what models/model.py:37-39 wires into LIFFireNet by default),
 * i.e. snn.Leaky(beta, threshold, reset_mechanism, reset_delay=False) applied to the batch-normalised input current.
 * PARITY UNPINNED: snntorch 0.9.4 is absent offline; the arithmetic restates its published Leaky.forward (csrc/leaky.cu
 * header).  The convolution(s) run through snnflow_convlif_fwd / _bwd (lam = 0 turns the cell
 * into a plain fused conv [+ recurrent conv]); BatchNorm2d stays with the caller.
 *   cur [B,C,H,W] batch-normalised current; mem_in [B,C,H,W] or NULL (zeros); beta, theta [C] raw parameters (beta is
 *   clamped to [0,1] inside, theta is already >= 0.01); subtract: 0 = reset to zero, 1 = reset by subtraction
 *   mem_out, spk [B,C,H,W]; m_pre [B,C,H,W] or NULL: the membrane before the reset, saved for the backward
 * backward (the cell detaches mem_out, so spk is the only differentiable output):
 *   g_cur = g_spk / (1 + (pi (m_pre - theta))^2)  (ATan surrogate, alpha = 2);  d_beta, d_theta [C] are overwritten
 * --------------------------------------------------------------------------------------------- */
int snnflow_leaky_fwd(const float* cur, const float* mem_in, const float* beta, const float* theta, float* mem_out,
                      float* spk, float* m_pre, int B, int C, int H, int W, int subtract, snnflow_stream_t stream);
size_t snnflow_leaky_bwd_workspace_bytes(int B, int C, int H, int W);
int snnflow_leaky_bwd(const float* g_spk, const float* m_pre, const float* mem_in, const float* beta, const float* theta,
                      float* g_cur, float* d_beta, float* d_theta, void* workspace, size_t workspace_bytes, int B, int C,
                      int H, int W, int subtract, snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer update of the training step (train_flow.py:264-271): torch.nn.utils.clip_grad.clip_grad_norm_(parameters,
 * max_norm) followed by torch.optim.Adam.step() (no amsgrad, no weight decay) over ONE flat fp32 parameter buffer, as
 * two launches (fixed-order sum of squares; clip coefficient + Adam per slice) instead of ~20 small PyTorch kernels.
 *   params, grads, exp_avg, exp_avg_sq : n floats each (grads is read only: the clipped gradient is not written back)
 *   hyper    device array {lr, beta1, beta2, eps, max_norm}; max_norm <= 0 disables clipping
 *   step     device int64, Adam's step counter: incremented by the call (so a CUDA-graph replay advances it)
 *   state    4 device doubles owned by the optimizer {beta1^t, beta2^t, lr/(1-beta1^t), 1/sqrt(1-beta2^t)}: the powers
 *            are running products, (re)initialised by the call whenever step == 0
 *   partials snnflow_clip_adam_partials(n) floats of scratch;  grad_norm: device float or NULL, receives the total norm
 *   gate     device uint32 or NULL: when *gate != 0 at execution time the whole update is skipped (parameters, moments
 *            and the step counter stay as they are).  The training window passes the window engine's sticky
 *            "input was not bf16-exact" flag (snnflow_window_flags_offset), so that a CUDA-graph replay can never
 *            apply an update computed from rounded inputs; the host polls the flag and raises.
 * A NaN gradient norm propagates into the parameters, as clip_grad_norm_ does.
 * --------------------------------------------------------------------------------------------- */
int snnflow_clip_adam_partials(int64_t n);
int snnflow_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* hyper,
                      int64_t* step, double* state, float* partials, float* grad_norm, const unsigned int* gate,
                      snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange (the reference has no distributed code): SUM all-reduce
 * of a flat fp32 buffer across the ranks of one NVLink domain as ONE kernel per rank, replayable inside a CUDA graph.
 *   peer_bufs  device array of `world` pointers: every rank's symmetric buffer (n floats) mapped into this process
 *   peer_pads  device array of `world` pointers: every rank's symmetric signal pad, >= 256 + ctas * world uint32, zero
 *              beyond slot 256 (the first 256 slots are left to the pad's owner, e.g. torch's own barriers)
 *   out        local result, n floats (must not alias a symmetric buffer that peers read)
 *   counter    snnflow_dp_allreduce_ctas() uint32 in local device memory, zeroed once; counts launches
 * All ranks must launch the kernel the same number of times.  The terms are added in rank order (deterministic,
 * identical on every rank).  A peer that never arrives traps the kernel instead of hanging it.
 * --------------------------------------------------------------------------------------------- */
int snnflow_dp_allreduce_ctas(void);
int snnflow_dp_allreduce_sum(const void* peer_bufs, const void* peer_pads, float* out, unsigned int* counter, int rank, int world,
                             size_t n, snnflow_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The whole parameter update of a (data-parallel) step as ONE kernel per rank: gradient SUM over the ranks through NVLink
 * peer memory + clip_grad_norm_ + Adam (train_flow.py:262-271 behind the exchange of SURVEY.md section 8e).  The peers'
 * gradients are read once, the reduced gradient never goes back to memory, and there is ONE cross-rank barrier per step:
 * the symmetric buffers hold two slots that alternate with the launch parity (a device-side counter, so the launch replays
 * from a CUDA graph), which makes the trailing "everybody has finished reading" barrier of snnflow_dp_allreduce_sum
 * unnecessary.  world == 1: the single-GPU optimizer step in one launch (no symmetric memory needed).
 *   grad_local  this rank's flat gradient, n floats (plain device memory; staged into the symmetric slot by the kernel)
 *   peer_bufs   device array of `world` pointers to the ranks' symmetric buffers, 2 * slot_floats floats each,
 *               slot_floats >= n + snnflow_dp_clip_adam_ctas();  peer_pads, counter: as for snnflow_dp_allreduce_sum, with
 *               counter holding snnflow_dp_clip_adam_ctas() words (do not share pads / counters between the two kernels)
 *   params, exp_avg, exp_avg_sq, hyper, step, state, grad_norm, gate: as for snnflow_clip_adam; a raised gate on ANY rank
 *               vetoes the update on all ranks (replicas stay identical);  partials: snnflow_dp_clip_adam_ctas() floats
 *   grid_counter one uint32, zeroed once (grid barrier of the kernel's CTAs; the launch is cooperative)
 *   reduced     NULL, or n floats that receive the summed gradient (tests)
 * n <= snnflow_dp_clip_adam_max_n() (131072; LIFFireNet C = 32 has 74 818 parameters).
 * --------------------------------------------------------------------------------------------- */
int snnflow_dp_clip_adam_ctas(void);
int64_t snnflow_dp_clip_adam_max_n(void);
int snnflow_dp_clip_adam(const float* grad_local, const void* peer_bufs, const void* peer_pads, unsigned int* counter, int rank,
                         int world, int64_t n, int64_t slot_floats, float* params, float* exp_avg, float* exp_avg_sq,
                         const float* hyper, int64_t* step, double* state, float* partials, unsigned int* grid_counter,
                         float* grad_norm, const unsigned int* gate, float* reduced, snnflow_stream_t stream);
/* Test hook: the same kernel body with all `world` ranks emulated on ONE GPU by one cooperative launch (a single-GPU box
 * cannot run ranks that wait for each other as separate launches).  rank_ptrs: device array [world][13] of the per-rank
 * pointers of snnflow_dp_clip_adam in argument order (grad_local, counter, params, exp_avg, exp_avg_sq, hyper, step, state,
 * partials, grid_counter, grad_norm, gate, reduced); peer_bufs / peer_pads: ordinary device allocations. */
int snnflow_dp_clip_adam_emulated(const void* rank_ptrs, const void* peer_bufs, const void* peer_pads, int world, int64_t n,
                                  int64_t slot_floats, snnflow_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SNNFLOW_H_ */
