// Layer-major window engine: host-side sequencing of one loss window of LIFFireNet / LIFFireFlowNet.
//
// The reference calls the network once per time bin and lets autograd replay the graph (models/model.py:172-182,
// train_flow.py:232-279).  A layer at bin t only depends on the layer below at bin t and on its own state, so the
// window can be executed LAYER BY LAYER: a feed-forward ConvLIF processes all T bins in one launch with its
// membrane in registers, a ConvLIFRecurrent takes one launch per bin, and in the backward pass the data gradient
// and the weight gradient of a layer are single launches over all T*B images.  ~30 + ~50 launches per window instead
// of ~80 + ~380.  Layouts are described in window.cuh.
#include "window.cuh"

namespace snnflow {

struct WinLayout {
  PlaneGeom g;
  size_t n;                       // B*C*H*W
  int Kin[WIN_LAYERS];            // allocated input channels (multiple of 16)
  int Cin[WIN_LAYERS];            // real input channels
  bool rec[WIN_LAYERS];
  size_t off_inplanes, off_flags;
  size_t off_zp[WIN_LAYERS], zp_img_stride;   // bf16 planes; recurrent layers have B leading images (initial spikes)
  size_t off_v[WIN_LAYERS], off_cur[WIN_LAYERS], off_state[WIN_LAYERS], off_init[WIN_LAYERS];
  size_t off_fwd_blob[WIN_LAYERS], off_dg_blob[WIN_LAYERS], off_rb_blob[WIN_LAYERS], off_par[WIN_LAYERS], off_gridbar[WIN_LAYERS];
  uint32_t fwd_blob_bytes[WIN_LAYERS], dg_blob_bytes[WIN_LAYERS], rb_blob_bytes[WIN_LAYERS], rec_w_off[WIN_LAYERS];
  size_t total;
};

static WinLayout win_layout(const snnflow_net_desc* d, int save) {
  WinLayout L{};
  L.g = plane_geom(d->H, d->W);
  const int C = d->C, T = d->T, B = d->B;
  L.n = (size_t)B * C * d->H * d->W;
  L.zp_img_stride = (size_t)(C / 8) * L.g.plane_bytes;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  L.off_flags = take(256);   // status words of the engine (word 0: sticky count of inputs that were not bf16-exact)
  L.off_inplanes = take((size_t)T * B * 2 * L.g.plane_bytes);
  for (int l = 0; l < WIN_LAYERS; ++l) {
    L.rec[l] = (d->recurrent_mask >> l) & 1u;
    L.Cin[l] = l == 0 ? d->num_bins : C;
    L.Kin[l] = l == 0 ? 16 : C;
    L.off_zp[l] = take((size_t)(L.rec[l] ? T + 1 : T) * B * L.zp_img_stride);
    L.off_state[l] = take((size_t)2 * L.n * sizeof(float));       // [v | z] after the window, NCHW (the caller's view)
    // training keeps the window's INITIAL state for the backward pass: the caller may hand in the previous window's
    // state block of this very arena, which the forward pass overwrites
    L.off_init[l] = save ? take((size_t)2 * L.n * sizeof(float)) : 0;
    if (save) {
      L.off_v[l] = take((size_t)T * L.n * sizeof(float));          // membranes of all bins, c8 layout
      L.off_cur[l] = 0;   // the input currents are not stored: the backward recovers them from consecutive membranes
    } else {
      L.off_v[l] = L.rec[l] ? take((size_t)2 * L.n * sizeof(float)) : 0;   // ping-pong membranes of a recurrent layer
      L.off_cur[l] = 0;
    }
    const size_t ffb = (size_t)9 * 3 * L.Kin[l] * C * 2, recb = L.rec[l] ? (size_t)9 * 3 * C * C * 2 : 0;
    L.rec_w_off[l] = (uint32_t)ffb;
    L.fwd_blob_bytes[l] = (uint32_t)(ffb + recb);
    L.off_fwd_blob[l] = take(L.fwd_blob_bytes[l]);
    L.dg_blob_bytes[l] = l > 0 ? (uint32_t)((size_t)9 * 2 * L.Kin[l] * C * 2) : 0;
    L.off_dg_blob[l] = take(L.dg_blob_bytes[l]);
    L.rb_blob_bytes[l] = L.rec[l] ? (uint32_t)((size_t)9 * 2 * C * C * 2) : 0;
    L.off_rb_blob[l] = take(L.rb_blob_bytes[l]);
    L.off_par[l] = take((size_t)C * 4 * sizeof(float));
    L.off_gridbar[l] = take((size_t)B * d->H * sizeof(unsigned int));   // per-tile progress flags of the layer's time-fused launches
  }
  L.total = o;
  return L;
}

struct WinPlan {   // tile plans of the tensor-core kernels for this shape
  int R_ff, S_ff, R_head, S_head, R_rec, S_rec, R_dg, S_dg, R_rb, S_rb, R_dp, S_dp;
  uint32_t dp_aux_bytes;   // shared memory set aside for the staged epilogue inputs of the fused dgrad + pointwise kernel
  int n_col;               // forward kernels: 128-pixel column tiles per row (1: whole-row tiles)
  uint32_t sub_ff, cs_ff, st_ff, sub_head, cs_head, st_head, sub_rec, cs_rec, st_rec, sub_dg, cs_dg, st_dg, sub_rb, cs_rb, st_rb, sub_dp, cs_dp, st_dp;
  bool ok;
};

// SNNFLOW_RB_FUSE=0: the data gradient through W_ff of the layer above a recurrent layer runs as its own launch again
// (read per call, not cached: the tests flip these switches inside one process)
static bool win_fuse_dgrad() { return wt_env_int("SNNFLOW_RB_FUSE", 1) != 0; }

static WinPlan win_plan(const snnflow_net_desc* d, const WinLayout& L, bool backward = true) {
  WinPlan P{};
  const int C = d->C;
  bool any_rec = false;
  for (int l = 0; l < WIN_LAYERS; ++l) any_rec |= L.rec[l];
  uint32_t max_fwd_rec_blob = 0;
  for (int l = 0; l < WIN_LAYERS; ++l)
    if (L.rec[l] && L.fwd_blob_bytes[l] > max_fwd_rec_blob) max_fwd_rec_blob = L.fwd_blob_bytes[l];
  // Rows wider than one 128-pixel MMA segment: the forward kernels walk 128-pixel COLUMN tiles (one accumulator segment
  // per item, as at W = 128) instead of whole rows with two segments per item - the two-segment epilogues spill at the
  // 96-register cap and measured 25-45 % slower per segment (profiles/r2_experiments.md).  Not for the time-fused forward
  // (its progress flags are per row tile).  SNNFLOW_COL_TILES=0 restores whole-row tiles.
  const bool col = d->W > 128 && d->W % 128 == 0 && wt_env_int("SNNFLOW_COL_TILES", 1) != 0 &&
                   !(any_rec && wt_env_int("SNNFLOW_FWD_PERSIST", 0) != 0);
  P.n_col = col ? d->W / 128 : 1;
  P.ok = wt_plan(d->H, d->W, C / 8, C, (uint32_t)((size_t)9 * 3 * C * C * 2), true, 3, false, &P.R_ff, &P.S_ff, &P.sub_ff, &P.cs_ff, &P.st_ff, 0, col);
  P.ok = P.ok && wt_plan(d->H, d->W, 2, C, (uint32_t)((size_t)9 * 3 * 16 * C * 2), true, 3, false, &P.R_head, &P.S_head, &P.sub_head,
                         &P.cs_head, &P.st_head, 0, col);
  if (any_rec)
    P.ok = P.ok && wt_plan(d->H, d->W, C / 8, C, max_fwd_rec_blob, true, 3, false, &P.R_rec, &P.S_rec, &P.sub_rec, &P.cs_rec, &P.st_rec, 0, col);
  if (!backward) return P;
  if (any_rec) {
    // two-row tiles when they fit (measured faster than one-row tiles at 128x128); SNNFLOW_RB_R overrides
    const int rb_R = wt_env_int("SNNFLOW_RB_R", 0);
    // (the fused kernel also stages the data-gradient weights of the layer above: two blobs of this size)
    const uint32_t rb_blob = (uint32_t)((size_t)9 * 2 * C * C * 2) * (win_fuse_dgrad() ? 2u : 1u);
    bool have = false;
    if (rb_R == 0 || rb_R == 2)
      have = wt_plan(d->H, d->W, C / 8, C, rb_blob, false, 2, false, &P.R_rb, &P.S_rb, &P.sub_rb, &P.cs_rb, &P.st_rb, 2);
    if (!have) have = wt_plan(d->H, d->W, C / 8, C, rb_blob, false, 2, false, &P.R_rb, &P.S_rb, &P.sub_rb, &P.cs_rb, &P.st_rb, rb_R);
    P.ok = P.ok && have;
  }
  P.ok = P.ok && wt_plan(d->H, d->W, C / 8, C, (uint32_t)((size_t)9 * 2 * C * C * 2), false, 2, true, &P.R_dg, &P.S_dg, &P.sub_dg,
                         &P.cs_dg, &P.st_dg);
  // data gradient fused with the time-fused pointwise chain of the layer below (state in registers: short tiles)
  // ... whose epilogue input v[t-1] can be staged by the producer into a ring of three shared-memory buffers (one-row
  // tiles only; SNNFLOW_DP_AUX=1).  Off by default: measured equal (3.012 vs 3.018 ms per step) - the producer's bulk copies
  // are throttled by the memory system either way (66 KB per item and SM; 3x halo re-reads of one-row tiles from L2), so
  // the epilogue's own global load was not the limiter (profiles/r2_experiments.md).
  const int dp_aux = wt_env_int("SNNFLOW_DP_AUX", 0);
  P.dp_aux_bytes = dp_aux ? 3u * (uint32_t)align_up((size_t)(C / 8) * d->W * 32, 128) : 0u;
  bool dp_ok = P.dp_aux_bytes && wt_plan(d->H, d->W, C / 8, C, (uint32_t)((size_t)9 * 2 * C * C * 2) + P.dp_aux_bytes, true, 2, false,
                                         &P.R_dp, &P.S_dp, &P.sub_dp, &P.cs_dp, &P.st_dp, 1);
  if (!dp_ok || P.S_dp < 3) {
    P.dp_aux_bytes = 0;
    dp_ok = wt_plan(d->H, d->W, C / 8, C, (uint32_t)((size_t)9 * 2 * C * C * 2), true, 2, false, &P.R_dp, &P.S_dp, &P.sub_dp,
                    &P.cs_dp, &P.st_dp, wt_env_int("SNNFLOW_DP_R", 0));
  }
  P.ok = P.ok && dp_ok;
  if (P.ok && P.R_dp * ceil_div(d->W, 128) > 2) P.ok = false;   // the fused kernel is instantiated for 1 or 2 segments per item
  return P;
}

static bool win_supported(const snnflow_net_desc* d, bool backward = true) {
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0 || d->T <= 0) return false;
  if (d->C != 16 && d->C != 32 && d->C != 64) return false;
  if (d->num_bins <= 0 || d->num_bins > 16) return false;
  if (!(d->flags & SNNFLOW_DETACH_RESET)) return false;        // the non-detached reset path stays on the per-step engine
  if (d->flags & SNNFLOW_NO_TENSOR_CORES) return false;
  if (d->recurrent_mask & 1u) return false;
  if (d->surrogate < 0 || d->surrogate > 2) return false;
  const WinLayout L = win_layout(d, 1);
  if (!win_plan(d, L, backward).ok) return false;
  if (!backward) return true;
  for (int l = 0; l < WIN_LAYERS; ++l)
    if (!wg_supported(d->C, L.Kin[l] / 8, L.rec[l] ? d->C / 8 : 0, d->H, d->W)) return false;
  return true;
}

struct WinWorkspace {
  size_t off_g[2], off_gp[2], gp_term_stride, off_gv, off_wpart[WIN_LAYERS][2], off_cpart[WIN_LAYERS], off_ppart, total;
  int wg_grid_max, rb_grid, dp_grid, pw_parts, pred_parts;
};

static WinWorkspace win_workspace(const snnflow_net_desc* d, const WinLayout& L, const WinPlan& P) {
  WinWorkspace W{};
  const int C = d->C, T = d->T, B = d->B;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  W.off_g[0] = take((size_t)T * L.n * sizeof(float));
  W.off_g[1] = take((size_t)T * L.n * sizeof(float));
  W.gp_term_stride = align_up((size_t)T * B * L.zp_img_stride, 256);
  W.off_gp[0] = take(2 * W.gp_term_stride);   // g_I planes (hi | lo) of even layers ...
  W.off_gp[1] = take(2 * W.gp_term_stride);   // ... and of odd layers: a layer's planes are read while the next ones are written
  W.off_gv = take(L.n * sizeof(float));
  W.wg_grid_max = 0;
  for (int l = 0; l < WIN_LAYERS; ++l) {
    const int g = wg_grid(T * B, d->H, d->W, C, L.Kin[l] / 8, L.rec[l] ? C / 8 : 0);
    if (g > W.wg_grid_max) W.wg_grid_max = g;
  }
  // per-layer partial blocks: the reductions of all layers run as ONE launch at the end of the backward pass
  for (int l = 0; l < WIN_LAYERS; ++l) {
    const int g = wg_parts(T * B, d->H, d->W, C, L.Kin[l] / 8, L.rec[l] ? C / 8 : 0);
    W.off_wpart[l][0] = take((size_t)g * 9 * L.Kin[l] * C * sizeof(float));
    W.off_wpart[l][1] = L.rec[l] ? take((size_t)g * 9 * C * C * sizeof(float)) : W.off_wpart[l][0];
  }
  W.rb_grid = P.R_rb ? wt_grid(B * (d->H / P.R_rb)) : 0;
  W.pw_parts = B * ceil_div(d->H * d->W, 256);
  W.dp_grid = wt_grid(B * (d->H / P.R_dp));
  size_t c1 = (size_t)2 * C * W.pw_parts;
  const size_t c2 = (size_t)T * W.rb_grid * 2 * C, c3 = (size_t)W.dp_grid * 2 * C;
  c1 = c1 > c2 ? c1 : c2;
  for (int l = 0; l < WIN_LAYERS; ++l) W.off_cpart[l] = take((c1 > c3 ? c1 : c3) * sizeof(float));
  W.pred_parts = pred_planes_parts(T * B, d->H, d->W);
  const int pp = W.pred_parts > W.pw_parts ? W.pred_parts : W.pw_parts;
  W.off_ppart = take((size_t)pp * (2 * C + 2) * sizeof(float));
  W.total = o;
  return W;
}

}  // namespace snnflow
using namespace snnflow;

// Planes of a recurrent layer: block 0 = the spikes entering the window, block 1 + t = the spikes of bin t.  Streaming
// with T == 1 swaps the two blocks from call to call (no copy): the call with phase p reads block 1 - p and writes block p.
static inline int win_rec_out_block(bool stream_mode, int T, int phase) { return (stream_mode && T == 1) ? phase : 1; }

extern "C" int snnflow_window_supported(const snnflow_net_desc* d, int backward) { return win_supported(d, backward != 0) ? 1 : 0; }

extern "C" size_t snnflow_window_arena_bytes(const snnflow_net_desc* d, int save) {
  if (!win_supported(d, save != 0)) return 0;
  return win_layout(d, save).total;
}

extern "C" size_t snnflow_window_workspace_bytes(const snnflow_net_desc* d) {
  if (!win_supported(d)) return 0;
  const WinLayout L = win_layout(d, 1);
  return win_workspace(d, L, win_plan(d, L)).total;
}

extern "C" size_t snnflow_window_flags_offset(const snnflow_net_desc* d, int save) {
  if (!win_supported(d, save != 0)) return 0;
  return win_layout(d, save).off_flags;
}

extern "C" int snnflow_window_state_offsets(const snnflow_net_desc* d, int save, size_t* offsets_bytes) {
  SNNFLOW_REQUIRE(win_supported(d, save != 0) && offsets_bytes, "unsupported shape or null pointer");
  const WinLayout L = win_layout(d, save);
  for (int l = 0; l < WIN_LAYERS; ++l) offsets_bytes[l] = L.off_state[l];
  return SNNFLOW_OK;
}

extern "C" int snnflow_window_forward(const snnflow_net_desc* d, const snnflow_layer_ptrs* layers, const float* pred_w,
                                      const float* pred_b, const float* input, const float* const* state_in, void* arena,
                                      float* flow, int save, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(win_supported(d, save != 0), "shape / options not covered by the window engine (use snnflow_net_forward)");
  SNNFLOW_REQUIRE(layers && pred_w && input && arena && flow, "null pointer");
  SNNFLOW_REQUIRE(((uintptr_t)arena & 255) == 0, "arena must be 256-byte aligned");
  // streaming mode: the states stay inside the arena in the engine's layout between calls (snnflow.h)
  const bool stream_mode = (d->flags & SNNFLOW_STATE_INTERNAL) != 0;
  SNNFLOW_REQUIRE(!stream_mode || !save, "SNNFLOW_STATE_INTERNAL is an inference mode (save = 0)");
  const int phase = (d->flags & SNNFLOW_STREAM_PHASE) ? 1 : 0;
  if (stream_mode) state_in = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const WinLayout L = win_layout(d, save);
  const WinPlan P = win_plan(d, L, save != 0);
  unsigned char* A = (unsigned char*)arena;
  const int C = d->C, T = d->T, B = d->B, H = d->H, W = d->W;
  const size_t n = L.n;
  const double px = (double)B * H * W;

  PackArgs pk{};
  for (int l = 0; l < WIN_LAYERS; ++l) {
    const snnflow_layer_ptrs& P_ = layers[l];
    SNNFLOW_REQUIRE(P_.w_ff && P_.lam && P_.theta && (!L.rec[l] || P_.w_rec), "null layer parameter");
    pk.L[l].w_ff = P_.w_ff; pk.L[l].w_rec = L.rec[l] ? P_.w_rec : nullptr;
    pk.L[l].fwd_blob = A + L.off_fwd_blob[l];
    pk.L[l].dg_blob = (save && l > 0) ? A + L.off_dg_blob[l] : nullptr;
    pk.L[l].rb_blob = (save && L.rec[l]) ? A + L.off_rb_blob[l] : nullptr;
    pk.L[l].Cin = L.Cin[l]; pk.L[l].Kin = L.Kin[l]; pk.L[l].C = C;
    pk.L[l].leak_lam = P_.lam; pk.L[l].theta = P_.theta;
    pk.L[l].par = (float*)(A + L.off_par[l]);
  }
  int rc = SNNFLOW_OK;
  if (!(d->flags & SNNFLOW_REUSE_PACKED)) {
    rc = launch_pack_weights(pk, st);
    if (rc) return rc;
  }
  rc = launch_pack_input(input, A + L.off_inplanes, T * B, d->num_bins, 2, H, W, (unsigned int*)(A + L.off_flags), st);
  if (rc) return rc;

  if (save && state_in) {
    for (int l = 0; l < WIN_LAYERS; ++l)
      if (state_in[l])
        SNNFLOW_CUDA(cudaMemcpyAsync(A + L.off_init[l], state_in[l], 2 * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  for (int l = 0; l < WIN_LAYERS; ++l) {
    const float* v_init = (state_in && state_in[l]) ? (save ? (const float*)(A + L.off_init[l]) : state_in[l]) : nullptr;
    const float* z_init = v_init ? v_init + n : nullptr;
    float* vbase = (float*)(A + L.off_v[l]);
    float* state = (float*)(A + L.off_state[l]);
    // input planes of this layer: the packed event counts, or the spikes of the layer below
    const unsigned char* xin;
    size_t x_img_stride;
    if (l == 0) { xin = A + L.off_inplanes; x_img_stride = 2 * L.g.plane_bytes; }
    else { xin = A + L.off_zp[l - 1] + (L.rec[l - 1] ? (size_t)win_rec_out_block(stream_mode, T, phase) * B * L.zp_img_stride : 0); x_img_stride = L.zp_img_stride; }
    WtArgs a{};
    a.src[0].planes = xin; a.src[0].img_stride = x_img_stride;
    a.src[0].n_chunks = (uint32_t)(L.Kin[l] / 8); a.src[0].w_off = 0; a.src[0].w_terms = 3; a.src[0].w_used = 3;
    a.n_src = 1;
    a.wblob = A + L.off_fwd_blob[l]; a.wblob_bytes = L.fwd_blob_bytes[l];
    a.B = B; a.H = H; a.W = W; a.Wp = W + 2; a.n_seg = P.n_col > 1 ? 1 : ceil_div(W, 128); a.n_col = P.n_col; a.N = C;
    a.hard_reset = (d->flags & SNNFLOW_HARD_RESET) ? 1 : 0;
    a.par = (const float*)(A + L.off_par[l]);
    a.zp_img_stride = L.zp_img_stride;
    if (stream_mode) a.state_c8 = 1;
    if (!L.rec[l]) {
      a.n_outer = B; a.T = T;
      if (stream_mode) {
        // membrane: c8, in place in the layer's state block; spikes entering the window: the layer's own planes of the
        // previous call's last bin (the slots this launch overwrites at ITS last bin - read first by the same thread)
        v_init = state; z_init = nullptr;
        a.zin_planes = A + L.off_zp[l] + (size_t)(T - 1) * B * L.zp_img_stride; a.zin_img_stride = L.zp_img_stride;
      }
      if (l == 0) { a.R = P.R_head; a.S = P.S_head; a.sub_bytes = P.sub_head; a.chunk_stride = P.cs_head; a.stage_bytes = P.st_head; }
      else { a.R = P.R_ff; a.S = P.S_ff; a.sub_bytes = P.sub_ff; a.chunk_stride = P.cs_ff; a.stage_bytes = P.st_ff; }
      if (stream_mode && T == 1 && wt_env_int("SNNFLOW_STREAM_STEP", 1)) {
        // one bin per call: the step-mode epilogue requests the membrane and the spikes of the NEXT tile before it waits for
        // the current accumulator (the sequence-mode epilogue loads them at the top of the tile and stalls on them: 59 % of
        // its stall samples at T = 1; 121 vs 128 us per layer at 256x256 batch 16, profiles/r2_experiments.md); the state
        // is read and written in place (a pixel is read by the thread that later writes it)
        a.v_prev = state; a.v_prev_nchw = 0; a.v_out = state; a.cur_out = nullptr;
        a.zp_out = A + L.off_zp[l];
        a.v_last = nullptr; a.z_last = nullptr;
        rc = launch_wt_fwd(a, false, st, "win_fwd_seq", px * (2.0 * L.Kin[l] + 12.0 * C), 18.0 * px * C * L.Cin[l]);
        if (rc) return rc;
        continue;
      }
      a.v_init = v_init; a.z_init = z_init;
      a.v_out = save ? vbase : nullptr; a.cur_out = nullptr;
      a.zp_out = A + L.off_zp[l];
      a.v_last = state; a.z_last = state + n;
      rc = launch_wt_fwd(a, true, st, "win_fwd_seq", (double)T * px * (2.0 * L.Kin[l] + 2.0 * C + (save ? 4.0 * C : 0.0)),   /* x planes (bf16) in ; z planes (bf16) [, v c8 fp32] out */
                         18.0 * T * px * C * L.Cin[l]);
      if (rc) return rc;
    } else {
      // initial spikes of the window -> image block 0 of this layer's planes (zeros when there is no state)
      if (stream_mode) {   // ... which are the planes the previous call's last bin wrote (block T)
        if (T > 1)         // (T == 1: the two blocks swap roles from call to call instead, win_rec_out_block)
          SNNFLOW_CUDA(cudaMemcpyAsync(A + L.off_zp[l], A + L.off_zp[l] + (size_t)T * B * L.zp_img_stride,
                                       (size_t)B * L.zp_img_stride, cudaMemcpyDeviceToDevice, st));
      } else if (z_init == nullptr) {
        SNNFLOW_CUDA(cudaMemsetAsync(A + L.off_zp[l], 0, (size_t)B * L.zp_img_stride, st));
      } else {
        rc = launch_pack_spikes(z_init, A + L.off_zp[l], B, C, H, W, st);
        if (rc) return rc;
      }
      a.n_outer = B; a.T = 1;
      a.R = P.R_rec; a.S = P.S_rec; a.sub_bytes = P.sub_rec; a.chunk_stride = P.cs_rec; a.stage_bytes = P.st_rec;
      a.n_src = 2;
      a.src[1] = a.src[0];
      a.src[1].img_stride = L.zp_img_stride; a.src[1].n_chunks = (uint32_t)(C / 8); a.src[1].w_off = L.rec_w_off[l];
      const int persistent = wt_env_int("SNNFLOW_FWD_PERSIST", 0);
      if (persistent && T > 1 && !stream_mode) {
        // SNNFLOW_FWD_PERSIST=1: ONE cooperative launch walks the T bins (time-fused ConvLIFRecurrent forward): weights,
        // barriers and TMEM stay set up and the pipeline never drains; the spike planes of bin t are the recurrent operand
        // of bin t + 1, and a tile only waits for the per-tile progress flags of itself and its two row neighbours
        // (WtArgs.tile_flags, raised by the publisher warp).  Timed alone the window of a layer costs 183 us against
        // 10 x 24.6 us of per-bin launches; inside the captured training step, where consecutive per-bin launches overlap
        // their set-up with the previous bin's tail (programmatic dependent launch), the two are equal within noise
        // (3.005 vs 2.984 ms per step, profiles/r2_experiments.md), so one launch per bin stays the default here.  The
        // recurrent BACKWARD uses the same machinery by default (it gains: 3.053 -> 3.005 ms).
        a.src[0].planes = xin;
        a.src[1].planes = A + L.off_zp[l];
        a.zin_planes = a.src[1].planes; a.zin_img_stride = L.zp_img_stride;
        a.zp_out = A + L.off_zp[l] + (size_t)B * L.zp_img_stride;
        a.v_prev = v_init; a.v_prev_nchw = 1;
        a.v_out = vbase; a.cur_out = nullptr;
        a.v_last = state; a.z_last = state + n;
        a.n_bins = T; a.bin_dep_mask = 2;
        a.bin_src_stride[0] = (long long)B * (long long)x_img_stride;
        a.bin_src_stride[1] = (long long)B * (long long)L.zp_img_stride;
        a.bin_zp_stride = (long long)B * (long long)L.zp_img_stride;
        a.bin_v_stride = (long long)n; a.bin_v_mask = save ? -1 : 1;
        a.tile_flags = (unsigned int*)(A + L.off_gridbar[l]);
        rc = launch_wt_fwd(a, false, st, "win_fwd_rec", (double)T * px * (2.0 * L.Kin[l] + 14.0 * C),   /* x, z planes (bf16), v (fp32), z again in the epilogue in ; v, z out */
                           18.0 * T * px * C * (L.Cin[l] + C));
        if (rc) return rc;
      } else
      for (int t = 0; t < T; ++t) {
        const int blk_out = win_rec_out_block(stream_mode, T, phase) + t, blk_in = (stream_mode && T == 1) ? 1 - blk_out : t;
        a.src[0].planes = xin + (size_t)t * B * x_img_stride;
        a.src[1].planes = A + L.off_zp[l] + (size_t)blk_in * B * L.zp_img_stride;
        a.zin_planes = a.src[1].planes; a.zin_img_stride = L.zp_img_stride;
        const bool last = t == T - 1;
        a.v_prev_nchw = t == 0;   // the window's initial state is the caller's NCHW tensor
        if (save) {
          a.v_prev = t > 0 ? vbase + (size_t)(t - 1) * n : v_init;
          a.v_out = vbase + (size_t)t * n; a.cur_out = nullptr;
        } else if (stream_mode) {
          // two membrane slots, ping-pong across bins AND calls: bin t of this call writes slot (t + phase) & 1 and reads the
          // other one, which is where the previous call's last bin wrote (the caller toggles the phase by T & 1)
          a.v_prev_nchw = 0;
          a.v_prev = vbase + (size_t)((t + phase + 1) & 1) * n;
          a.v_out = vbase + (size_t)((t + phase) & 1) * n; a.cur_out = nullptr;
        } else {
          a.v_prev = t > 0 ? vbase + (size_t)((t - 1) & 1) * n : v_init;
          a.v_out = vbase + (size_t)(t & 1) * n; a.cur_out = nullptr;
        }
        a.v_last = last ? state : nullptr; a.z_last = last ? state + n : nullptr;
        a.zp_out = A + L.off_zp[l] + (size_t)blk_out * B * L.zp_img_stride;
        rc = launch_wt_fwd(a, false, st, "win_fwd_rec", (double)px * (2.0 * L.Kin[l] + 14.0 * C),   /* x, z planes (bf16), v (fp32), z again in the epilogue in ; v, z out */
                           18.0 * px * C * (L.Cin[l] + C));
        if (rc) return rc;
      }
    }
  }
  const int top = WIN_LAYERS - 1;
  return launch_pred_fwd_planes(A + L.off_zp[top] + (L.rec[top] ? (size_t)win_rec_out_block(stream_mode, T, phase) * B * L.zp_img_stride : 0),
                                L.zp_img_stride, pred_w,
                                pred_b, flow, T * B, C, H, W, st);
}

// where the streaming state of layer l lives: membrane (c8) and the planes image block that holds the spikes
static void win_stream_state(const WinLayout& L, const snnflow_net_desc* d, int l, int phase_next, unsigned char* A, float** v,
                             unsigned char** z_planes) {
  const int T = d->T, B = d->B;
  if (!L.rec[l]) {
    *v = (float*)(A + L.off_state[l]);
    *z_planes = A + L.off_zp[l] + (size_t)(T - 1) * B * L.zp_img_stride;
  } else {
    // the next call reads slot (0 + phase_next + 1) & 1 and copies planes block T to block 0 (T == 1: reads block 1 - phase_next)
    *v = (float*)(A + L.off_v[l]) + (size_t)((phase_next + 1) & 1) * L.n;
    *z_planes = A + L.off_zp[l] + (size_t)(T == 1 ? 1 - (phase_next & 1) : T) * B * L.zp_img_stride;
  }
}

extern "C" int snnflow_window_import_state(const snnflow_net_desc* d, const float* const* state_in, void* arena,
                                           snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(win_supported(d, false) && arena, "unsupported shape or null arena");
  const WinLayout L = win_layout(d, 0);
  for (int l = 0; l < WIN_LAYERS; ++l) {
    float* v;
    unsigned char* zp;
    win_stream_state(L, d, l, /*phase_next=*/0, (unsigned char*)arena, &v, &zp);
    const float* s = state_in ? state_in[l] : nullptr;
    int rc = launch_state_import(s, s ? s + L.n : nullptr, v, zp, L.zp_img_stride, d->B, d->C, d->H, d->W, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return SNNFLOW_OK;
}

extern "C" int snnflow_window_export_state(const snnflow_net_desc* d, const void* arena, int phase, float* const* state_out,
                                           snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(win_supported(d, false) && arena && state_out, "unsupported shape or null pointer");
  const WinLayout L = win_layout(d, 0);
  for (int l = 0; l < WIN_LAYERS; ++l) {
    if (!state_out[l]) continue;
    float* v;
    unsigned char* zp;
    win_stream_state(L, d, l, phase & 1, const_cast<unsigned char*>((const unsigned char*)arena), &v, &zp);
    int rc = launch_state_export(v, zp, L.zp_img_stride, state_out[l], state_out[l] + L.n, d->B, d->C, d->H, d->W, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return SNNFLOW_OK;
}

extern "C" int snnflow_window_backward(const snnflow_net_desc* d, const snnflow_layer_ptrs* layers, const float* pred_w,
                                       const float* const* state_in, const void* arena, const float* flow,
                                       const float* g_flow, float* d_pred_w, float* d_pred_b, void* workspace,
                                       size_t workspace_bytes, snnflow_stream_t stream) {
  SNNFLOW_REQUIRE(win_supported(d), "shape / options not covered by the window engine (use snnflow_net_backward)");
  SNNFLOW_REQUIRE(layers && pred_w && arena && flow && g_flow && workspace, "null pointer");
  SNNFLOW_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)arena & 255) == 0, "arena / workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const WinLayout L = win_layout(d, 1);
  const WinPlan P = win_plan(d, L);
  const WinWorkspace WS = win_workspace(d, L, P);
  if (workspace_bytes < WS.total) {
    set_error("snnflow_window_backward: workspace %zu < %zu", workspace_bytes, WS.total);
    return SNNFLOW_EWORKSPACE;
  }
  const unsigned char* A = (const unsigned char*)arena;
  unsigned char* Wk = (unsigned char*)workspace;
  const int C = d->C, T = d->T, B = d->B, H = d->H, W = d->W;
  const size_t n = L.n;
  const double px = (double)B * H * W;
  float* gbuf = (float*)(Wk + WS.off_g[0]);   // spike gradient of a recurrent layer (from the layer above), c8
  unsigned char* gplanes[2] = {Wk + WS.off_gp[0], Wk + WS.off_gp[1]};   // g_I planes of layer l live in gplanes[l & 1]
  float* g_v = (float*)(Wk + WS.off_gv);
  WinReduceArgs reduces[WIN_LAYERS];   // queued per layer, executed by one launch after the loop
  float* ppart = (float*)(Wk + WS.off_ppart);
  const int hard = (d->flags & SNNFLOW_HARD_RESET) ? 1 : 0;
  const int top = WIN_LAYERS - 1;
  int rc = SNNFLOW_OK;
  // A feed-forward top layer takes its spike gradient straight from the flow head inside its time-fused pointwise
  // kernel (no [T*B,C,H,W] gradient tensor is written or read); a recurrent top layer goes through the head's own kernel.
  const bool fuse_head = !L.rec[top];
  if (!fuse_head) {
    rc = launch_pred_bwd_planes(A + L.off_zp[top] + (size_t)B * L.zp_img_stride, L.zp_img_stride, pred_w, flow, g_flow, gbuf, ppart, T * B,
                                C, H, W, st);
    if (rc) return rc;
    rc = launch_pred_reduce_planes(ppart, d_pred_w, d_pred_b, C, WS.pred_parts, st);
    if (rc) return rc;
  }
  // Per layer, top down:  (1) its g_I planes - recurrent: one fused conv^T(W_rec) + pointwise launch per bin; feed-forward:
  // already produced by step (3) of the layer above (top layer: its own time-fused pointwise kernel);  (2) weight gradient
  // over all T*B images + reductions;  (3) data gradient through W_ff for the layer below - fused with that layer's
  // time-fused pointwise chain when it is feed-forward, a plain launch into the c8 gradient buffer when it is recurrent.
  int n_cpart = 0, cpart_layout = 0;   // d lam / d theta partials of the CURRENT layer
  for (int l = top; l >= 0; --l) {
    const snnflow_layer_ptrs& P_ = layers[l];
    const float* v_init = (state_in && state_in[l]) ? (const float*)(A + L.off_init[l]) : nullptr;   // forward's copy
    const float* z_init = v_init ? v_init + n : nullptr;
    const float* vbase = (const float*)(A + L.off_v[l]);
    const float* par = (const float*)(A + L.off_par[l]);
    unsigned char* gp = gplanes[l & 1];
    float* wpart[2] = {(float*)(Wk + WS.off_wpart[l][0]), (float*)(Wk + WS.off_wpart[l][1])};
    float* cpart = (float*)(Wk + WS.off_cpart[l]);
    if (L.rec[l]) {
      WtArgs a{};
      // Recurrent BPTT step t, one launch per bin.  Tensor cores: g_z = conv^T(g_I[t+1], W_rec) (hi planes x (w_hi, w_lo),
      // then lo planes x w_hi: two pipeline units per tile) and - fused, when there is a layer above - the spike gradient
      // coming down through that layer's W_ff, g_out[t] = conv^T(g_I^{l+1}[t], W_ff^{l+1}) (two more units into the SAME
      // accumulator), so that g_out [T*B,C,H,W] is never written or read and the separate data-gradient launch is gone.
      const bool fuse = win_fuse_dgrad() && l < top;
      WtSrc rec_hi{}, rec_lo{};
      rec_hi.img_stride = L.zp_img_stride; rec_hi.n_chunks = (uint32_t)(C / 8);
      rec_hi.w_off = fuse ? L.dg_blob_bytes[l + 1] : 0; rec_hi.w_terms = 2; rec_hi.w_used = 2;
      rec_lo = rec_hi; rec_lo.w_used = 1;
      if (fuse) {
        a.src[0] = rec_hi; a.src[0].w_off = 0;            // g_I planes of layer l + 1 (hi ; lo), weights: its data-gradient blob
        a.src[1] = rec_lo; a.src[1].w_off = 0;
        a.wblob = A + L.off_dg_blob[l + 1]; a.wblob_bytes = L.dg_blob_bytes[l + 1];
        a.wblob2 = A + L.off_rb_blob[l]; a.wblob2_bytes = L.rb_blob_bytes[l];
      } else {
        a.wblob = A + L.off_rb_blob[l]; a.wblob_bytes = L.rb_blob_bytes[l];
      }
      a.n_outer = B; a.T = 1; a.B = B; a.H = H; a.W = W; a.Wp = W + 2; a.n_seg = ceil_div(W, 128); a.N = C;
      a.R = P.R_rb; a.S = P.S_rb; a.sub_bytes = P.sub_rb; a.chunk_stride = P.cs_rb; a.stage_bytes = P.st_rb;
      a.hard_reset = hard; a.surrogate = d->surrogate; a.width = d->act_width;
      a.par = par;
      a.g_v = g_v; a.gp_img_stride = L.zp_img_stride; a.gp_term_stride = WS.gp_term_stride;
      const unsigned char* gp_above = gplanes[(l + 1) & 1];
      const int rb_persist = wt_env_int("SNNFLOW_RB_PERSIST", 1);
      const bool fused_time = fuse && rb_persist && T > 1;
      if (fused_time) {
        // ONE cooperative launch walks the window backwards (bin j of the launch = time bin T-1-j): weights, barriers and
        // TMEM stay set up, d lam / d theta partial sums stay in registers over all bins, and the g_I planes a tile needs
        // for its next bin are awaited through per-tile progress flags (WtArgs.tile_flags) instead of a launch boundary.
        const long long bin_planes = (long long)B * (long long)L.zp_img_stride;
        a.src[0].planes = gp_above + (size_t)(T - 1) * B * L.zp_img_stride;
        a.src[1].planes = a.src[0].planes + WS.gp_term_stride;
        a.src[2] = rec_hi; a.src[3] = rec_lo;
        a.src[2].planes = gp + (size_t)T * B * L.zp_img_stride;   // bin j reads the planes of t + 1 = T - j (bin 0: unused)
        a.src[3].planes = a.src[2].planes + WS.gp_term_stride;
        a.n_src = 4; a.n_src_bin0 = 2; a.has_gz = 1;
        a.n_bins = T; a.bin_dep_mask = 0xC;
        for (int i = 0; i < 4; ++i) a.bin_src_stride[i] = -bin_planes;
        a.gp_out = gp + (size_t)(T - 1) * B * L.zp_img_stride; a.bin_zp_stride = -bin_planes;
        a.g_out = nullptr;
        a.v_t = vbase + (size_t)(T - 1) * n; a.v_in = vbase + (size_t)(T - 2) * n; a.bin_v_stride = -(long long)n;
        a.v_init = v_init; a.z_init = z_init; a.z_from_v = 1;
        a.part = cpart;
        a.tile_flags = (unsigned int*)(const_cast<unsigned char*>(A) + L.off_gridbar[l]);   // scratch shared with the forward launch
        rc = launch_wt_recbwd(a, st, (double)T * px * C * 28.0 /* 2 x (hi + lo) planes in, v[t], v[t-1], g_v in/out, hi + lo planes out */,
                              18.0 * px * C * C * (2 * T - 1));
        if (rc) return rc;
      } else
      for (int t = T - 1; t >= 0; --t) {
        const bool have_rec = t < T - 1;
        a.first_step = t == T - 1;
        int ns = 0;
        if (fuse) {
          a.src[0].planes = gp_above + (size_t)t * B * L.zp_img_stride;
          a.src[1].planes = a.src[0].planes + WS.gp_term_stride;
          ns = 2;
        }
        if (have_rec) {
          a.src[ns] = rec_hi; a.src[ns + 1] = rec_lo;
          a.src[ns].planes = gp + (size_t)(t + 1) * B * L.zp_img_stride;
          a.src[ns + 1].planes = a.src[ns].planes + WS.gp_term_stride;
          ns += 2;
        }
        a.n_src = ns; a.has_gz = ns > 0;
        a.g_out = fuse ? nullptr : gbuf + (size_t)t * n; a.v_t = vbase + (size_t)t * n;
        a.v_in = t > 0 ? vbase + (size_t)(t - 1) * n : v_init;
        a.v_in_nchw = t == 0;
        a.z_from_v = t > 0; a.z_init = z_init;
        a.gp_out = gp + (size_t)t * B * L.zp_img_stride;
        a.part = cpart + (size_t)t * WS.rb_grid * 2 * C;
        rc = launch_wt_recbwd(a, st, (double)px * C * (20.0 + (fuse ? 4.0 : 4.0) + (have_rec ? 4.0 : 0.0)),   /* planes bf16, the rest fp32 */
                              18.0 * px * C * C * (ns / 2));
        if (rc) return rc;
      }
      n_cpart = (fused_time ? 1 : T) * WS.rb_grid; cpart_layout = 1;
    } else if (l == top) {
      PwSeqArgs a{};
      a.v = vbase; a.g_out = fuse_head ? nullptr : gbuf; a.v_init = v_init; a.z_init = z_init; a.par = par;
      a.flow = flow; a.g_flow = g_flow; a.pred_w = pred_w; a.pred_part = ppart;
      a.gp = gp; a.gp_img_stride = L.zp_img_stride; a.gp_term_stride = WS.gp_term_stride;
      a.part = cpart; a.T = T; a.B = B; a.C = C; a.H = H; a.W = W; a.hard_reset = hard; a.surrogate = d->surrogate;
      a.n_part = WS.pw_parts; a.width = d->act_width;
      rc = launch_pw_seq(a, st);
      if (rc) return rc;
      if (fuse_head) {
        rc = launch_pred_reduce_rows(ppart, d_pred_w, d_pred_b, C, WS.pw_parts, st);
        if (rc) return rc;
      }
      n_cpart = WS.pw_parts; cpart_layout = 0;
    }   // else: planes and partials of this feed-forward layer were produced by step (3) of layer l + 1

    // (2) weight gradients over all T*B images, then the fixed-order reduction into the caller's accumulators
    {
      WgArgs a{};
      if (l == 0) { a.xp[0] = A + L.off_inplanes; a.x_img_stride[0] = 2 * L.g.plane_bytes; }
      else { a.xp[0] = A + L.off_zp[l - 1] + (L.rec[l - 1] ? (size_t)B * L.zp_img_stride : 0); a.x_img_stride[0] = L.zp_img_stride; }
      a.x_chunks[0] = L.Kin[l] / 8; a.cin_alloc[0] = L.Kin[l]; a.cin_real[0] = L.Cin[l];
      a.n_xsrc = 1;
      if (L.rec[l]) {
        a.xp[1] = A + L.off_zp[l]; a.x_img_stride[1] = L.zp_img_stride;   // image t*B + b = spikes before bin t
        a.x_chunks[1] = C / 8; a.cin_alloc[1] = C; a.cin_real[1] = C;
        a.n_xsrc = 2;
      }
      a.gp = gp; a.g_img_stride = L.zp_img_stride; a.g_term_stride = WS.gp_term_stride;
      a.part[0] = wpart[0]; a.part[1] = wpart[1];
      a.n_img = T * B; a.H = H; a.W = W; a.C = C;
      rc = launch_wgrad_planes(a, st, (double)T * px * (4.0 * C + 2.0 * L.Kin[l] + (L.rec[l] ? 2.0 * C : 0.0)),   /* g hi + lo, x [, z] planes (bf16) */
                               18.0 * T * px * C * (L.Cin[l] + (L.rec[l] ? C : 0)));
      if (rc) return rc;
      WinReduceArgs r{};
      r.wpart[0] = wpart[0]; r.wdst[0] = P_.dw_ff; r.cin_alloc[0] = L.Kin[l]; r.cin_real[0] = L.Cin[l];
      r.wpart[1] = wpart[1]; r.wdst[1] = L.rec[l] ? P_.dw_rec : nullptr; r.cin_alloc[1] = C; r.cin_real[1] = C;
      r.n_wpart = wg_parts(T * B, H, W, C, L.Kin[l] / 8, L.rec[l] ? C / 8 : 0);
      r.cpart = cpart; r.n_cpart = n_cpart; r.cpart_layout = cpart_layout;
      r.dlam = P_.dlam; r.dtheta = P_.dtheta; r.C = C;
      r.lam = P_.lam; r.thresh_raw = P_.thresh_raw; r.d_leak = P_.d_leak; r.d_thresh = P_.thresh_raw ? P_.d_thresh : nullptr;
      reduces[l] = r;
    }

    // (3) data gradient through W_ff: the spike gradient of the layer below
    if (l > 0) {
      WtArgs a{};
      a.src[0].planes = gp; a.src[0].img_stride = L.zp_img_stride; a.src[0].n_chunks = (uint32_t)(C / 8);
      a.src[0].w_off = 0; a.src[0].w_terms = 2; a.src[0].w_used = 2;
      a.src[1] = a.src[0]; a.src[1].planes = gp + WS.gp_term_stride; a.src[1].w_used = 1;
      a.n_src = 2;
      a.wblob = A + L.off_dg_blob[l]; a.wblob_bytes = L.dg_blob_bytes[l];
      a.B = B; a.H = H; a.W = W; a.Wp = W + 2; a.n_seg = ceil_div(W, 128); a.N = L.Kin[l];
      if (!L.rec[l - 1]) {
        // ... fused with the time-fused pointwise chain of the feed-forward layer below (writes ITS g_I planes and partials)
        const float* v_init_b = (state_in && state_in[l - 1]) ? (const float*)(A + L.off_init[l - 1]) : nullptr;
        a.n_outer = B; a.T = T; a.t_reverse = 1;
        a.R = P.R_dp; a.S = P.S_dp; a.sub_bytes = P.sub_dp; a.chunk_stride = P.cs_dp; a.stage_bytes = P.st_dp;
        a.hard_reset = hard; a.surrogate = d->surrogate; a.width = d->act_width;
        a.par = (const float*)(A + L.off_par[l - 1]);
        a.v_t = (const float*)(A + L.off_v[l - 1]);
        a.v_init = v_init_b; a.z_init = v_init_b ? v_init_b + n : nullptr;
        a.gp_out = gplanes[(l - 1) & 1]; a.gp_img_stride = L.zp_img_stride; a.gp_term_stride = WS.gp_term_stride;
        a.part = (float*)(Wk + WS.off_cpart[l - 1]);
        if (P.dp_aux_bytes && P.R_dp == 1) {
          a.aux = a.v_t; a.aux_chunk_bytes = (uint32_t)W * 32u;
          a.aux_stage_bytes = P.dp_aux_bytes / 3u; a.aux_slots = 3;
        }
        rc = launch_wt_dgpw(a, st, (double)T * px * (4.0 * C + 8.0 * L.Kin[l]) /* g_I hi + lo in ; v in, g_I hi + lo out */, 18.0 * T * px * C * L.Kin[l]);
        if (rc) return rc;
        n_cpart = WS.dp_grid; cpart_layout = 1;
      } else if (!win_fuse_dgrad()) {   // (fused into the recurrent layer's own BPTT launches otherwise, see above)
        a.n_outer = T * B; a.T = 1;
        a.R = P.R_dg; a.S = P.S_dg; a.sub_bytes = P.sub_dg; a.chunk_stride = P.cs_dg; a.stage_bytes = P.st_dg;
        a.g_x = gbuf;
        rc = launch_wt_dgrad(a, st, (double)T * px * (4.0 * C + 4.0 * L.Kin[l]), 18.0 * T * px * C * L.Kin[l]);
        if (rc) return rc;
      }
    }
  }
  return launch_win_reduce(reduces, WIN_LAYERS, st);
}
