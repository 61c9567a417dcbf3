// Shared declarations of the layer-major window engine (window.cu, window_tc.cu, window_wgrad.cu, window_simt.cu).
//
// Data layout of the engine (all library-private, inside the caller-owned arena / workspace):
//   * spike planes: bf16 [image][chunk = C/8][H+2][W+2][8 channels]  - one 16-byte "slot" per pixel and 8-channel
//     chunk, with a one-pixel zero border around every image.  A row tile with its halo is therefore ONE contiguous
//     byte range per chunk: it is staged by a TMA bulk copy straight into the canonical no-swizzle UMMA layout
//     (K-major for the forward / data gradient, MN-major for the weight gradient) with no register staging, and the
//     zero border supplies the convolution padding and isolates images from each other.  Spikes {0,1} and event
//     counts are exact in bf16.  The borders are zeroed once (the arena is allocated zero-filled) and never written.
//   * gradient planes: the input-current gradient g_I as bf16 hi + lo planes of the same geometry (16 mantissa bits).
//   * membranes v, input currents I, gradients w.r.t. spikes: fp32 NCHW like the reference tensors.
#pragma once
#include "common.cuh"

namespace snnflow {

constexpr int WIN_LAYERS = 7;

struct PlaneGeom {
  int H, W, Hp, Wp;
  size_t plane_bytes;   // one chunk of one image
};
inline PlaneGeom plane_geom(int H, int W) {
  PlaneGeom g;
  g.H = H; g.W = W; g.Hp = H + 2; g.Wp = W + 2;
  g.plane_bytes = (size_t)g.Hp * g.Wp * 16;
  return g;
}

// ---- tensor-core tile pipeline (window_tc.cu) ------------------------------------------------------
struct WtSrc {
  const unsigned char* planes;   // image 0, term 0, chunk 0, padded row 0
  unsigned long long img_stride;   // bytes
  uint32_t n_chunks;             // K/8
  uint32_t w_off, w_terms, w_used;   // byte offset of this source's weights in the blob ; weight terms stored side by side
                                     // along N (3 forward, 2 gradients: ONE MMA of N = terms * channels per tap and k-step,
                                     // summed by the epilogue) ; terms multiplied with this source (gradients: hi planes x
                                     // {hi, lo}, lo planes x {hi})
};

constexpr int WT_MAX_SRC = 4;

struct WtArgs {
  WtSrc src[WT_MAX_SRC];
  int n_src;
  int n_src_bin0;                 // time-fused launches: sources of the first bin (0: n_src); later bins use all n_src
  const unsigned char* wblob;     // weight images, staged once per CTA; a second blob (wblob2) is placed right behind the
  uint32_t wblob_bytes;           // first in shared memory (WtSrc.w_off of its sources counts from the start of the first)
  const unsigned char* wblob2;
  uint32_t wblob2_bytes;
  int n_outer;   // sequence mode: B sequences (image = t*B + b) ; otherwise the number of images
  int T, B;
  int H, W, Wp, R, S, n_seg, N;   // R rows per tile, S pipeline stages, n_seg 128-pixel segments per row, N = MMA N
  int n_prod;                     // TMA producer warps (1 or 2)
  int n_col, Wsm;                 // forward kernels: 128-pixel column tiles per row (1 = whole rows); pitch of a tile row in shared memory
  uint32_t sub_bytes, chunk_stride, stage_bytes;
  int hard_reset, surrogate;
  float width;
  const float* par;   // [N][4] = (lam, 1 - lam, theta, 1 / (1 - lam))
  // forward
  const float *v_init, *z_init;   // sequence mode, t == 0: [B][N][H][W] or NULL (zeros)
  const float* v_prev;            // step mode: membrane before this step (c8, or NCHW when v_prev_nchw) or NULL (zeros)
  int v_prev_nchw;
  const unsigned char* zin_planes;   // step mode: spikes before this step (planes, image b) or NULL (zeros)
  unsigned long long zin_img_stride;
  float *v_out, *cur_out;         // c8 layout [images][N/8][H*W][8] or NULL
  unsigned char* zp_out;          // spike planes written by this launch (image 0 = first image of the launch)
  unsigned long long zp_img_stride;
  float *v_last, *z_last;         // [B][N][H][W] or NULL: state after the last step
  int state_c8;                   // streaming mode: v_init / v_last are c8 tensors, the spikes entering the window are read from
                                  // zin_planes (sequence mode too), no NCHW state is written
  // data gradient
  float* g_x;                     // c8 layout [images][N/8][H*W][8]
  // recurrent backward step
  const float *g_out, *v_t, *v_in;   // c8 ; v_in: membrane before the step (c8, or NCHW when v_in_nchw) or NULL
  int v_in_nchw;                  // z_init (NCHW): spikes before step 0
  float* g_v;                     // c8 [B][N/8][H*W][8] in/out: gradient w.r.t. the membrane carried to the previous step
  unsigned char* gp_out;          // g_I planes of this step (hi ; lo at + gp_term_stride)
  unsigned long long gp_img_stride, gp_term_stride;
  float* part;                    // [grid][2][N] partial sums of dlam, dtheta
  int t_reverse;                  // sequence mode: walk the bins backwards (t = T-1 .. 0)
  uint32_t acc_lg;                // log2 of the accumulator ring length, 1 or 2 (set by the launcher)
  // Time-fused mode of the per-bin (recurrent) kernels: ONE cooperative launch walks n_bins dependent bins.  The operand
  // planes of (tile, bin j) are written by the epilogues of bin j-1 of the tile and of its two row neighbours (halo rows),
  // possibly in other CTAs: the producer acquires their per-tile progress flags (common.cuh: tile_flag_*) before the bulk
  // copy - a point-to-point wavefront, no grid barrier, and the pipeline never drains between bins.  Pointers above
  // describe bin 0; the strides below step them per bin.
  int n_bins;                     // <= 1: single bin (plain launch)
  int bin_dep_mask;               // bit i: src[i] of bin j > 0 is produced by bin j-1 of this launch
  long long bin_src_stride[WT_MAX_SRC];   // bytes per bin for src[i].planes
  long long bin_zp_stride;        // bytes per bin for zin_planes / zp_out
  long long bin_v_stride;         // floats per bin for the membrane arena (v_out; v_prev of bin j is v_out of bin j-1)
  int bin_v_mask;                 // membrane slot of bin j = j & bin_v_mask (eval keeps two slots, training all bins)
  unsigned int* tile_flags;       // [n_outer * H / R] progress flags, zeroed by the launcher
  // Epilogue inputs staged by the producer ("aux ring", sequence mode of wt_dgpw_kernel): the c8 row(s) of `aux` the epilogue
  // of an item needs - v[t-1] of the tile - arrive by bulk copy in a ring of aux_slots buffers behind the operand stages, so
  // the epilogue reads them from shared memory instead of waiting for a global load at the top of every item.
  const float* aux;               // c8 tensor [T*B][N/8][H*W][8] or NULL (epilogue loads from global memory itself)
  uint32_t aux_chunk_bytes;       // R * W * 32: one 8-channel chunk of a tile's rows
  uint32_t aux_stage_bytes;       // (N/8) chunks, 128-byte aligned
  int aux_slots;
  int exp;                        // experiment switches (SNNFLOW_EXP), 0 in production
  int l2_prefetch;                // epilogues issue prefetch.global.L2 for the next item's streamed inputs (set by the launcher)
  int has_gz, first_step, z_from_v;   // first_step: g_v starts at zero ; z_from_v: z_in = spike(v_in) else from z_init
  long long* dbg;                 // optional [grid][8] cycle counters (SNNFLOW_WT_TIMING=1): where each role waits
};

int launch_wt_fwd(const WtArgs& a, bool seq, cudaStream_t st, const char* prof_name, double bytes, double flops);
int launch_wt_dgrad(const WtArgs& a, cudaStream_t st, double bytes, double flops);
int launch_wt_recbwd(const WtArgs& a, cudaStream_t st, double bytes, double flops);
// data gradient of layer l+1 fused with the time-fused pointwise BPTT of the feed-forward layer l below it
int launch_wt_dgpw(const WtArgs& a, cudaStream_t st, double bytes, double flops);
// picks rows-per-tile / stages for the given shapes; returns false when the shape does not fit
bool wt_plan(int H, int W, int max_chunks_per_stage, int N, uint32_t wblob_bytes, bool seq_state, int w_terms, bool tall, int* R, int* S,
             uint32_t* sub_bytes, uint32_t* chunk_stride, uint32_t* stage_bytes, int only_R = 0, bool column_tiles = false);
int wt_env_int(const char* name, int dflt);
int wt_grid(int n_tiles);

// ---- weight gradient (window_wgrad.cu) --------------------------------------------------------------
struct WgArgs {
  const unsigned char* xp[2];
  unsigned long long x_img_stride[2];
  int x_chunks[2], cin_alloc[2], cin_real[2], n_xsrc;
  const unsigned char* gp;
  unsigned long long g_img_stride, g_term_stride;
  float* part[2];   // per source: [wg_parts][9][cin_alloc][C]
  int n_img, H, W, Wp, R, S, C, P, n_cg, rpm, n_kyg, ksteps, x_rows;
  int n_prod;       // TMA producer warps (1 or 2)
  int pair;         // both gradient rows of a two-row tile in ONE MMA (N = 4C); two partial blocks per CTA
  uint32_t stage_bytes, g_off;
  long long* dbg;   // optional [grid][8] cycle counters (SNNFLOW_WT_TIMING=1)
};
bool wg_supported(int C, int cin_chunks, int rec_chunks, int H, int W);
int wg_grid(int n_img, int H, int W, int C, int cin_chunks, int rec_chunks);
int wg_parts(int n_img, int H, int W, int C, int cin_chunks, int rec_chunks);   // partial blocks the launch writes per source
int launch_wgrad_planes(WgArgs a, cudaStream_t st, double bytes, double flops);

// ---- CUDA-core helpers (window_simt.cu) ---------------------------------------------------------------
int launch_pack_input(const float* in, unsigned char* planes, int n_img, int nb, int n_chunks, int H, int W,
                      unsigned int* inexact, cudaStream_t st);
int launch_pack_spikes(const float* z, unsigned char* planes, int n_img, int C, int H, int W, cudaStream_t st);
// streaming state conversion: NCHW fp32 (v, z) <-> c8 membranes + bf16 spike planes, one layer
int launch_state_import(const float* v_nchw, const float* z_nchw, float* v_c8, unsigned char* planes, unsigned long long img_stride,
                        int B, int C, int H, int W, cudaStream_t st);
int launch_state_export(const float* v_c8, const unsigned char* planes, unsigned long long img_stride, float* v_nchw, float* z_nchw,
                        int B, int C, int H, int W, cudaStream_t st);
struct PackLayer {
  const float *w_ff, *w_rec;
  unsigned char *fwd_blob, *dg_blob, *rb_blob;   // forward (ff [+ rec]) ; data gradient through W_ff ; through W_rec
  int Cin, Kin, C;   // Cin real input channels, Kin = allocated (multiple of 16), C output channels
  const float *leak_lam, *theta;   // effective lam / theta (already sigmoid'ed / clamped)
  float* par;                      // [C][4]
};
struct PackArgs { PackLayer L[WIN_LAYERS]; };
int launch_pack_weights(const PackArgs& p, cudaStream_t st);
struct PwSeqArgs {
  const float *v, *g_out;          // c8 layout [T*B][C/8][H*W][8]
  const float *v_init, *z_init;    // [B][C][H][W] or NULL
  const float* par;                // [C][4]
  unsigned char* gp;               // g_I planes (hi ; lo at + term_stride), image t*B + b
  unsigned long long gp_img_stride, gp_term_stride;
  float* part;                     // [2][C][n_part]
  int T, B, C, H, W, hard_reset, surrogate, n_part;
  float width;
  // top layer only (g_out == NULL): the spike gradient comes straight from the flow head,
  //   g_out[c] = g_pre_x * w[0][c] + g_pre_y * w[1][c],  g_pre = g_flow * (1 - flow^2)      (models/submodules.py:96-113)
  // and the head's own gradients dw [2][C], db [2] are accumulated as per-block partials pred_part [2C + 2][n_part]
  const float *flow, *g_flow, *pred_w;   // [T*B][2][H*W], [T*B][2][H*W], [2][C]
  float* pred_part;
};
int launch_pw_seq(const PwSeqArgs& a, cudaStream_t st);
int launch_pred_fwd_planes(const unsigned char* zp, unsigned long long img_stride, const float* w, const float* b,
                           float* flow, int n_img, int C, int H, int W, cudaStream_t st);
int launch_pred_bwd_planes(const unsigned char* zp, unsigned long long img_stride, const float* w, const float* flow,
                           const float* g_flow, float* g_x, float* part, int n_img, int C, int H, int W,
                           cudaStream_t st);
int pred_planes_parts(int n_img, int H, int W);
struct WinReduceArgs {
  const float* wpart[2];   // [n_wpart][9][cin_alloc][C]
  float* wdst[2];          // [C][cin_real][9]  (+=)
  int cin_alloc[2], cin_real[2], n_wpart;
  const float* cpart;      // channel partials, rows of length n_cpart: row r < 2C
  int n_cpart, cpart_layout;   // layout 0: [2][C][n_cpart] ; 1: [n_cpart][2][C]
  float *dlam, *dtheta;    // (+=)
  // optional chain rule to the raw parameters (snnflow_layer_ptrs): d_leak += dlam * lam (1 - lam), d_thresh += dtheta [thresh >= 0.01]
  const float *lam, *thresh_raw;
  float *d_leak, *d_thresh;
  int C;
};
int launch_win_reduce(const WinReduceArgs* layers, int n_layers, cudaStream_t st);   // all layers in one launch
int launch_pred_reduce_planes(const float* part, float* dw, float* db, int C, int n_part, cudaStream_t st);
int launch_pred_reduce_rows(const float* part, float* dw, float* db, int C, int n_part, cudaStream_t st);

}  // namespace snnflow
