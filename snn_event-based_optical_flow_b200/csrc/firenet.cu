// Whole-network window driver: LIFFireNet / LIFFireFlowNet forward over T bins and the matching BPTT.
// Host-side sequencing only - every kernel is one of the per-layer entry points of this library; this file
// removes the Python / autograd round trip between them (models/model.py:172-182 called T times, plus the
// autograd graph of train_flow.py:262) and makes the whole window capturable in a CUDA graph.
#include "common.cuh"

namespace snnflow {

constexpr int NET_LAYERS = 7;

struct NetDims {
  size_t n;        // B*C*H*W
  size_t n_in;     // B*num_bins*H*W
  size_t n_flow;   // B*2*H*W
};

static NetDims dims(const snnflow_net_desc* d) {
  NetDims m;
  m.n = (size_t)d->B * d->C * d->H * d->W;
  m.n_in = (size_t)d->B * d->num_bins * d->H * d->W;
  m.n_flow = (size_t)d->B * 2 * d->H * d->W;
  return m;
}

static int check_desc(const snnflow_net_desc* d) {
  SNNFLOW_REQUIRE(d != nullptr, "null descriptor");
  SNNFLOW_REQUIRE(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0 && d->T > 0 && d->num_bins > 0, "bad dims");
  SNNFLOW_REQUIRE(!(d->recurrent_mask & 1u), "the head layer cannot be recurrent");
  return SNNFLOW_OK;
}

// block of (layer l, bin t) in the activation arena: [v | z | I]
static inline float* act_block(float* acts, const snnflow_net_desc* d, int save, int l, int t, size_t n) {
  if (save) return acts + ((size_t)l * d->T + t) * 3 * n;
  return acts + ((size_t)l * 2 + (t & 1)) * 2 * n;
}

}  // namespace snnflow
using namespace snnflow;

extern "C" size_t snnflow_net_acts_floats(const snnflow_net_desc* d, int save) {
  if (!d || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0 || d->T <= 0) return 0;
  const size_t n = (size_t)d->B * d->C * d->H * d->W;
  return save ? (size_t)NET_LAYERS * d->T * 3 * n : (size_t)NET_LAYERS * 2 * 2 * n;
}

extern "C" int snnflow_net_forward(const snnflow_net_desc* d, const snnflow_layer_ptrs* layers, const float* pred_w,
                                   const float* pred_b, const float* input, const float* const* state_in,
                                   float* acts, float* flow, int save, snnflow_stream_t stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  SNNFLOW_REQUIRE(layers && pred_w && input && acts && flow, "null pointer");
  const NetDims m = dims(d);
  const unsigned fwd_flags = d->flags & (SNNFLOW_HARD_RESET | SNNFLOW_DETACH_RESET);
  for (int t = 0; t < d->T; ++t) {
    const float* x = input + (size_t)t * m.n_in;
    int cin = d->num_bins;
    for (int l = 0; l < NET_LAYERS; ++l) {
      const snnflow_layer_ptrs& L = layers[l];
      const bool rec = (d->recurrent_mask >> l) & 1u;
      SNNFLOW_REQUIRE(L.w_ff && L.lam && L.theta && (!rec || L.w_rec), "null layer parameter");
      float* blk = act_block(acts, d, save, l, t, m.n);
      const float *v_in = nullptr, *z_in = nullptr;
      if (t > 0) {
        const float* prev = act_block(acts, d, save, l, t - 1, m.n);
        v_in = prev; z_in = prev + m.n;
      } else if (state_in && state_in[l]) {
        v_in = state_in[l]; z_in = state_in[l] + m.n;
      }
      float* cur = save ? blk + 2 * m.n : nullptr;
      // layer 0 sees event counts (CUDA-core exact path); deeper layers see this library's own spikes
      const bool tc = l > 0 && L.packed && !(d->flags & SNNFLOW_NO_TENSOR_CORES);
      if (tc)
        rc = snnflow_convlif_fwd_tc(x, L.packed, rec ? 1 : 0, v_in, z_in, L.lam, L.theta, nullptr, blk, blk + m.n, nullptr,
                                    cur, d->B, cin, d->C, d->H, d->W, fwd_flags, stream);
      else
        rc = snnflow_convlif_fwd(x, L.w_ff, rec ? L.w_rec : nullptr, v_in, z_in, L.lam, L.theta, nullptr, blk, blk + m.n,
                                 nullptr, cur, d->B, cin, d->C, d->H, d->W, fwd_flags, stream);
      if (rc) return rc;
      x = blk + m.n;   // spikes feed the next layer
      cin = d->C;
    }
    rc = snnflow_pred_fwd(x, pred_w, pred_b, flow + (size_t)t * m.n_flow, d->B, d->C, d->H, d->W, stream);
    if (rc) return rc;
  }
  return SNNFLOW_OK;
}

namespace snnflow {
struct NetBwdLayout {
  size_t off_gx[2], off_gstate, off_layer_ws, layer_ws_bytes, pred_ws_bytes, total;
};
static NetBwdLayout net_bwd_layout(const snnflow_net_desc* d) {
  NetBwdLayout L;
  const size_t n = (size_t)d->B * d->C * d->H * d->W;
  size_t o = 0;
  L.off_gx[0] = o; o += align_up(n * sizeof(float), 256);       // ping-pong: gradient w.r.t. a layer's input spikes
  L.off_gx[1] = o; o += align_up(n * sizeof(float), 256);
  L.off_gstate = o; o += align_up((size_t)NET_LAYERS * 2 * 2 * n * sizeof(float), 256);   // [layer][parity][g_v | g_z]
  size_t lw = 0;
  for (int l = 0; l < NET_LAYERS; ++l) {
    const size_t b = snnflow_convlif_bwd_workspace_bytes(d->B, l == 0 ? d->num_bins : d->C, d->C, d->H, d->W,
                                                         (d->recurrent_mask >> l) & 1u);
    lw = b > lw ? b : lw;
  }
  L.pred_ws_bytes = snnflow_pred_bwd_workspace_bytes(d->B, d->C, d->H, d->W);
  L.layer_ws_bytes = lw > L.pred_ws_bytes ? lw : L.pred_ws_bytes;
  L.off_layer_ws = o; o += align_up(L.layer_ws_bytes, 256);
  L.total = o;
  return L;
}
}  // namespace snnflow

extern "C" size_t snnflow_net_bwd_workspace_bytes(const snnflow_net_desc* d) {
  if (!d || d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0 || d->T <= 0 || d->num_bins <= 0) return 0;
  return net_bwd_layout(d).total;
}

extern "C" int snnflow_net_backward(const snnflow_net_desc* d, const snnflow_layer_ptrs* layers, const float* pred_w,
                                    const float* input, const float* const* state_in, const float* acts,
                                    const float* flow, const float* g_flow, float* d_pred_w, float* d_pred_b,
                                    void* workspace, size_t workspace_bytes, snnflow_stream_t stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  SNNFLOW_REQUIRE(layers && pred_w && input && acts && flow && g_flow && workspace, "null pointer");
  SNNFLOW_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  const NetBwdLayout W = net_bwd_layout(d);
  if (workspace_bytes < W.total) {
    set_error("snnflow_net_backward: workspace %zu < %zu", workspace_bytes, W.total);
    return SNNFLOW_EWORKSPACE;
  }
  const NetDims m = dims(d);
  char* ws = (char*)workspace;
  float* gx[2] = {(float*)(ws + W.off_gx[0]), (float*)(ws + W.off_gx[1])};
  float* gstate = (float*)(ws + W.off_gstate);
  void* lws = ws + W.off_layer_ws;
  const bool detach = d->flags & SNNFLOW_DETACH_RESET;
  float* acts_m = const_cast<float*>(acts);

  for (int t = d->T - 1; t >= 0; --t) {
    // flow head: g_flow[t] -> gradient w.r.t. the last layer's spikes
    const float* z_top = act_block(acts_m, d, 1, NET_LAYERS - 1, t, m.n) + m.n;
    int cur_gx = 0;
    rc = snnflow_pred_bwd(z_top, pred_w, flow + (size_t)t * m.n_flow, g_flow + (size_t)t * m.n_flow, gx[cur_gx], d_pred_w,
                          d_pred_b, lws, W.layer_ws_bytes, d->B, d->C, d->H, d->W, stream);
    if (rc) return rc;
    for (int l = NET_LAYERS - 1; l >= 0; --l) {
      const snnflow_layer_ptrs& L = layers[l];
      const bool rec = (d->recurrent_mask >> l) & 1u;
      const int cin = l == 0 ? d->num_bins : d->C;
      const float* blk = act_block(acts_m, d, 1, l, t, m.n);
      const float* x = l == 0 ? input + (size_t)t * m.n_in : act_block(acts_m, d, 1, l - 1, t, m.n) + m.n;
      const float *v_in = nullptr, *z_in = nullptr;
      if (t > 0) {
        const float* prev = act_block(acts_m, d, 1, l, t - 1, m.n);
        v_in = prev; z_in = prev + m.n;
      } else if (state_in && state_in[l]) {
        v_in = state_in[l]; z_in = state_in[l] + m.n;
      }
      // gradients of the state this step RETURNED come from step t+1 (none at the end of the window)
      float* gs_next = gstate + ((size_t)l * 2 + ((t + 1) & 1)) * 2 * m.n;
      float* gs_this = gstate + ((size_t)l * 2 + (t & 1)) * 2 * m.n;
      const bool has_gz = rec || !detach;   // otherwise the state's z carries no gradient at all
      const float* g_v_out = t + 1 < d->T ? gs_next : nullptr;
      const float* g_z_out = (t + 1 < d->T && has_gz) ? gs_next + m.n : nullptr;
      float* g_x = l > 0 ? gx[cur_gx ^ 1] : nullptr;
      unsigned flags = d->flags | (l > 0 ? SNNFLOW_INPUT_EXACT16 : 0u);
      rc = snnflow_convlif_bwd(x, L.w_ff, rec ? L.w_rec : nullptr, v_in, z_in, blk, blk + 2 * m.n, L.lam, L.theta, gx[cur_gx],
                               g_v_out, g_z_out, g_x, gs_this, has_gz ? gs_this + m.n : nullptr, L.dw_ff,
                               rec ? L.dw_rec : nullptr, L.dlam, L.dtheta, lws, W.layer_ws_bytes, d->B, cin, d->C, d->H, d->W,
                               flags, d->surrogate, d->act_width, stream);
      if (rc) return rc;
      cur_gx ^= 1;
    }
  }
  return SNNFLOW_OK;
}
