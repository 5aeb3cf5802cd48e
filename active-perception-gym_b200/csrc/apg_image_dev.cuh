// Device-side pieces shared by the image step kernels (internal, not part of the C-ABI): the kernel parameter block,
// the scipy-compatible tap, the exact uint8 -> float64 pixel conversion and the two per-environment scalar chains
// (env_move / env_loss).  Reference file:line citations are in the functions below and in apg_image.cu.
#pragma once
#include <cuda_runtime.h>

#include <cmath>

#include "apg_common.h"

namespace apg_img {

constexpr int kThreads = 256;
constexpr int kThreadBlendMaxP = 96;  // glimpses up to this many floats are blended by one or two threads per env
constexpr int kMaxK = 4096;

struct EnvHdr;
struct EnvCtr;
struct ImgParams {
  int N, H, W, C, S0, S1, P, K, task, invert, track, words, T;
  int E;             // environments per block
  int bank_pitch;    // bytes per bank image row (W*C*elem rounded up to 16: every row is 16-byte aligned)
  int pitch;         // bytes per staged tile row
  int tile_rows;     // staged rows per environment
  int tile_bytes;    // pitch * tile_rows
  int thread_blend;  // 1: thread-per-environment blend (small glimpses), 0: warp-per-environment
  EnvHdr* g_hdr;     // split == 2: tile headers / centres written by image_head_kernel, (n_sets, N)
  EnvCtr* g_ctr;
  int n0, n1;        // environments [n0, n1) of the shard are stepped by this launch (special-purpose kernels)
  int force_slow;    // TMA path (tests): send every environment through the general per-output path
  int pdl;           // launched with programmatic stream serialization: wait for the previous grid explicitly
  int tb_consec;     // thread-blend path: use the consecutive-row variants
  int ctr_off;       // byte offset of the EnvCtr array in dynamic shared memory (split phase A)
  int split;         //  phase A split into per-env scalars (one thread) + taps / staging (all threads)
  int sparse;        // 1: the loss only enters the reward of the step that ends the episode (`-sparse` ids)
  int fast;          // 1: float32 blend (within the 1e-5 bar), 0: the reference's float64 blend (bit-exact)
  int tb;            // thread-blend path: threads that share one environment's sensor columns
  int mode;          // 0 = move step, 1 = autoreset step, 2 = explicit reset
  int t_stamp;       // overlay stamp slot of this step (pre-move time index), mode 0 only
  float time_obs;    // obs["time_step"] value after this step
  unsigned char terminated;
  double scale, msl_x, msl_y, lim_x, lim_y, offy, offx, half0, half1;
  double rs, half_wr, half_hr;
  double ce_scale, ce_offset;
  float mse_scale, mse_offset;
  // inputs
  const void* bank;
  const int* bank_labels;
  const float* action;
  const float* pred;
  // reset payload (modes 1, 2)
  const long long* new_idx;
  const double* new_pos;
  const float* new_target;
  const unsigned char* new_inv;
  // state
  double* pos;
  int* st_idx;
  int* st_label;
  float* st_target;
  unsigned char* st_inv;
  // outputs
  float* glimpse;
  float* target_glimpse;
  float* glimpse_pos;
  float* time_step;
  int* inverted_label;
  float* reward;
  float* base_reward;
  float* loss;
  int* target_label;
  float* target_xy;
  unsigned char* terminated_out;
  unsigned char* truncated_out;
  long long* index;
  float* reward2;           // optional per-step rollout outputs
  unsigned char* done2;
  // episode statistics
  float* acc1; float* acc2; float* acc_ret;   // per-env running sums of the current episode
  float* ep_avg1; float* ep_final1; float* ep_avg2; float* ep_final2; float* ep_return;
  int* first_correct; int* last_incorrect;
  double* stats;            // (APG_NUM_STATS) shard-wide sums
  float inv_k;              // float32(1 / num_classes)
  // overlay
  int* stamp_xy;
  float* stamp_q;
  unsigned* visited;
  unsigned* err;
};

// What changes from step to step inside a multi-step launch (same field names as in ImgParams).
struct StepCtx {
  int mode, t_stamp;
  unsigned char terminated;
  float time_obs;
  const float* action;
  const float* pred;
  const long long* new_idx;
  const double* new_pos;
  const float* new_target;
  const unsigned char* new_inv;
  float* reward2;
  unsigned char* done2;
};

constexpr int kRolloutMaxSteps = 128;  // steps per launch
constexpr int kRolloutMaxSlots = 8;    // autoresets per launch

// What one multi-step launch steps through.  The episode clock is shared by all environments, so the mode of every step
// follows from (t0, need_reset0): a step is an autoreset step exactly when the previous one terminated.  p.action /
// p.pred / p.reward2 / p.done2 point at the FIRST step of the window (stacked (n_steps, N, ...) arrays).
struct RolloutPlan {
  int n_steps;
  int t0;           // episode clock before the first step
  int need_reset0;  // the first step is an autoreset step
  int first_mode;   // 2: the first (only) step is an explicit reset from slot 0
  unsigned char reset_slot[kRolloutMaxSteps];  // for autoreset steps: which payload slot below
  float time_obs[kRolloutMaxSteps];            // obs["time_step"] after each step (host: the reference's float64 expression)
  const long long* slot_idx[kRolloutMaxSlots];
  const double* slot_pos[kRolloutMaxSlots];
  const float* slot_target[kRolloutMaxSlots];
  const unsigned char* slot_inv[kRolloutMaxSlots];
};

// scipy find_indices on the grid g[k] = k - off (see oracle/image_oracle.py:tap_indices)
__device__ __forceinline__ void tap(double coord, double off, int len, int& k_out, double& w_out) {
  int k = (int)floor(coord + off);
  if ((double)k - off > coord) k -= 1;
  else if ((double)(k + 1) - off <= coord) k += 1;
  k = max(0, min(k, len - 2));
  k_out = k;
  w_out = coord - ((double)k - off);
}

// Row taps of uint8 banks keep the unclamped index len-1 when the coordinate sits exactly on the last grid point (the
// sensor touches the bottom edge): scipy's clamped tap (len-2, w = 1) and this one (len-1, w = 0) blend the same row
// with the same products, the other row gets weight +0 (finite uint8 pixels, exact), and the tap rows stay
// CONSECUTIVE, which is what the fast blend variants need.  Readers of row k+1 must not leave the image.
__device__ __forceinline__ void tap_row(double coord, double off, int len, bool keep_last, int& k_out, double& w_out) {
  int k = (int)floor(coord + off);
  if ((double)k - off > coord) k -= 1;
  else if ((double)(k + 1) - off <= coord) k += 1;
  k = max(0, min(k, keep_last ? len - 1 : len - 2));
  k_out = k;
  w_out = coord - ((double)k - off);
}

// bank pixel -> the float64 value the reference blends with: float64(float32(u8) / 255) or float64(f32 pixel).
// float32(u8)/255 is reproduced exactly for all 256 values by q0 = u*r, q = fma(fma(-255, q0, u), r, q0) with
// r = float32(1/255): no table, no division.
__device__ __forceinline__ double px_to_f64(unsigned char u) {
  const float r = 0.003921568859368563f;  // float32(1)/float32(255)
#ifndef APG_ALU_CONVERSIONS
  const float uf = (float)u;
#else
  // integer -> float without the conversion unit: 2^23 + u is exact in float32
  const float uf = __fsub_rn(__uint_as_float(0x4B000000u | (unsigned)u), 8388608.0f);
#endif
  const float q0 = __fmul_rn(uf, r);
  const float e = __fmaf_rn(-255.0f, q0, uf);
  const float q = __fmaf_rn(e, r, q0);
#ifndef APG_ALU_CONVERSIONS
  return (double)q;
#else
  // exact float32 -> float64 widening by re-biasing the exponent (q is 0 or a normal float >= 1/255).  Measured on
  // B200: trading the 3 conversion-pipe instructions for 8 ALU/FMA ones is 3 % SLOWER (the loop is issue-bound), so
  // this variant is compiled only with -DAPG_ALU_CONVERSIONS
  const unsigned b = __float_as_uint(q);
  const unsigned hi = b ? ((b >> 3) + 0x38000000u) : 0u;
  return __hiloint2double((int)hi, (int)(b << 29));
#endif
}
__device__ __forceinline__ double px_to_f64(float v) { return (double)v; }

struct __align__(16) Tap {
  union {
    double w;  // exact blend: normalised distance to the lower grid point (float64, like scipy)
    struct {
      float wf, omwf;  // float32 blend: w and 1 - w, each rounded once from the float64 values
    };
  };
  int k;  // tap index relative to the staged tile (row) / absolute column
  int pad;
};

struct __align__(16) EnvHdr {  // 16 bytes: the per-environment shared-memory footprint decides blocks per SM
  unsigned short lead;   // bits 0-3: byte offset of the tile's first needed column inside the staged (aligned-down)
                         // row; bit 15 (kNonConsec): the row taps are NOT k0, k0+1, ..., k0+S0-1
  unsigned short nrows;  // staged rows
  unsigned short cmin;   // first tap column of the tile
  unsigned short rmin;   // first tap row of the tile
  unsigned colbits[2];   // visited-mask column bits of this glimpse, words 0 and 1 (columns {j0, j0+1} of every tap)
};

struct __align__(16) EnvCtr {  // two 16-byte halves: what the tap items read / what the copy items read
  double ccx, ccy;   // sensor centre (pixels)
  long long src0;    // byte offset of the tile's first copied byte inside the bank
  int span, nrows;   // bytes per tile row to copy (multiple of 16), tile rows
};

constexpr unsigned kNonConsec = 0x8000u;
constexpr unsigned kLeadMask = 15u;

// float32 blend (opt-in): pixel -> float32 without the exact-division emulation; the result differs from the reference's
// float64 blend by a few float32 ulps of the (non-negative) sum, far inside the 1e-5 relative bar.
__device__ __forceinline__ float px_to_f32(unsigned char u) { return __fmul_rn((float)u, 0.003921568859368563f); }
__device__ __forceinline__ float px_to_f32(float v) { return v; }

// A block owns E consecutive environments and works in three phases per glimpse set:
//   A  thread e: the per-environment scalars (move, loss, reward, flags, overlay stamp), the S0+S1 bilinear
//      taps (-> shared memory) and the asynchronous staging (cp.async, LDGSTS) of the glimpse's bank footprint
//      -- rows [i0[0], i0[S0-1]+1] x the aligned byte range covering columns [j0[0], j0[S1-1]+1] -- into the
//      block's shared-memory tile.  All E footprints of the block are in flight at once (no registers held), which
//      is what hides the HBM/L2 latency of the random bank gather.
//   B  warp w walks environments w, w+4, ...; lane = (row-in-pass, column, channel) of the sensor window.  Taps come
//      from the tile (LDS.U8), with one sensor row per pass the bottom taps of row i are the top taps of row i+1;
//      the blend is the reference's 4-term float64 sum in the same order (bit-exact pixels); each output row is one
//      contiguous run of stores.  The visited bitmask rows of the environment are updated by the warp.
// The per-environment scalar work of one step is two INDEPENDENT chains with disjoint state, so that two threads can run
// them side by side (and neither waits for the other's memory round trips):
//   env_move   position state: pre-move overlay stamp, float64 position update, bank index; leaves the glimpse position
//              (px, py), the target centre (tcx, tcy; localization reset steps) and the bank index behind -- everything
//              the taps / staging need;
//   env_loss   labels / targets, loss, reward, flags, overlay colour, episode statistics.
// Both read the (read-only) action; env_move alone touches pos / st_idx, env_loss alone st_label / st_target / acc*.
// `c` carries what changes from step to step (mode, t_stamp, terminated, time_obs, action, pred, the reset payload new_*,
// reward2 / done2): the kernel parameter block itself for one-step launches, a StepCtx inside multi-step launches.
template <class Ctx>
__device__ __forceinline__ void env_move(const ImgParams& p, const Ctx& c, int n, double& px, double& py, double& tcx, double& tcy,
                                         int& idx, unsigned& err) {
  if (c.mode == 0) {
    const float2 aa = reinterpret_cast<const float2*>(c.action)[n];
    const double2 pp = reinterpret_cast<const double2*>(p.pos)[n];
    idx = p.st_idx[n];
    px = pp.x;
    py = pp.y;
    const float a0 = aa.x, a1 = aa.y;
    if (isnan(a0) || isnan(a1)) err |= APG_FLAG_NAN_ACTION;
    if (p.track) {  // overlay stamp of the PRE-move position
      p.stamp_xy[(size_t)(c.t_stamp * 2 + 0) * p.N + n] = (int)rint((px * p.lim_x) * p.rs + p.half_wr);
      p.stamp_xy[(size_t)(c.t_stamp * 2 + 1) * p.N + n] = (int)rint((py * p.lim_y) * p.rs + p.half_hr);
    }
    const float m = __fsqrt_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)));
    float q0 = a0, q1 = a1;
    if (m > 1.0f) {
      const float den = fmaxf(m, 1.0f);
      q0 = __fmul_rn(__fdiv_rn(a0, den), 1.0f);
      q1 = __fmul_rn(__fdiv_rn(a1, den), 1.0f);
    }
    px = fmin(fmax(px + p.msl_x * (double)q0, -1.0), 1.0);
    py = fmin(fmax(py + p.msl_y * (double)q1, -1.0), 1.0);
  } else {
    const long long ni = c.new_idx[n];
    const double2 pp = reinterpret_cast<const double2*>(c.new_pos)[n];
    idx = (int)ni;
    px = pp.x;
    py = pp.y;
    p.st_idx[n] = idx;
    p.index[n] = ni;
    if (p.task != APG_TASK_CLASSIFICATION) {
      const float2 nt = reinterpret_cast<const float2*>(c.new_target)[n];
      tcx = (double)nt.x * p.lim_x;  // float32 target promoted to float64
      tcy = (double)nt.y * p.lim_y;
    }
  }
  reinterpret_cast<double2*>(p.pos)[n] = make_double2(px, py);
  reinterpret_cast<float2*>(p.glimpse_pos)[n] = make_float2((float)px, (float)py);
}

template <class Ctx>
__device__ __forceinline__ void env_loss(const ImgParams& p, const Ctx& c, int n, double* s_stats, bool ep_end, unsigned& err) {
  float base = 0.0f, pq = 0.0f, loss_f = 0.0f;
  float m1 = 0.0f, m2 = 0.0f;  // LogWrapper metrics of this step
  double neg_loss_d = 0.0;
  const bool cls = p.task == APG_TASK_CLASSIFICATION;
  // Everything an ordinary step reads is requested here, back to back, before the first store: loads that sit behind
  // stores they might alias cost one memory round trip each.
  float2 in_action = make_float2(0.f, 0.f), in_pred2 = make_float2(0.f, 0.f), in_target = make_float2(0.f, 0.f);
  float in_acc1 = 0.0f, in_acc2 = 0.0f, in_accr = 0.0f, x_first = 0.0f, x_last = 0.0f;
  int in_fc = -1, in_lic = -1, label = 0;
  if (c.mode == 0) {
    in_action = reinterpret_cast<const float2*>(c.action)[n];
    if (c.t_stamp != 0) {
      in_acc1 = p.acc1[n];
      in_acc2 = p.acc2[n];
      in_accr = p.acc_ret[n];
      if (cls) {
        in_fc = p.first_correct[n];
        in_lic = p.last_incorrect[n];
      }
    }
    if (cls) label = p.st_label[n];
  }
  if (cls && c.mode != 2) {  // pulls the environment's logits into L1 with the other inputs
    x_first = c.pred[(size_t)n * p.K];
    x_last = c.pred[(size_t)n * p.K + p.K - 1];
  }
  if (!cls && c.mode != 2) {
    in_target = reinterpret_cast<const float2*>(p.st_target)[n];
    in_pred2 = reinterpret_cast<const float2*>(c.pred)[n];
  }
  // ---- labels / targets and the loss (the reset step still charges a loss)
  if (cls) {
    unsigned char inv = 0;
    if (c.mode != 0) {
      label = p.bank_labels[(int)c.new_idx[n]];
      if (p.invert) {
        inv = c.new_inv[n];
        if (inv) label = p.K - label - 1;
      }
      p.st_label[n] = label;
    }
    if (p.invert) p.inverted_label[n] = (c.mode == 0) ? 2 : (int)inv;
    p.target_label[n] = label;
    if (c.mode != 2) {
      const float* x = c.pred + (size_t)n * p.K;
      float mx = fmaxf(x_first, x_last);
      bool has_nan = isnan(x_first) || isnan(x_last);
      for (int k = 1; k < p.K - 1; ++k) {
        const float v = x[k];
        has_nan |= isnan(v);
        mx = fmaxf(mx, v);
      }
      if (has_nan) err |= APG_FLAG_NAN_PRED;
      if (!isfinite(mx)) mx = 0.0f;
      float s = 0.0f;
      for (int k = 0; k < p.K; ++k) s = __fadd_rn(s, expf(__fsub_rn(x[k], mx)));
      const float tl = __fsub_rn(x[label], mx);
      const float ce = -__fsub_rn(tl, logf(s));  // -log_softmax(x)[label], float32 like scipy
      pq = __fdiv_rn(expf(tl), s);               // softmax(x)[label], overlay colour only
      const double loss_d = (double)ce * p.ce_scale + p.ce_offset;  // np.float64 scale -> float64 loss
      loss_f = (float)loss_d;
      neg_loss_d = -loss_d;
      m1 = pq;                            // correct_label_prob (active_classification_env.py:157-161)
      m2 = (pq > p.inv_k) ? 1.0f : 0.0f;  // accuracy (:166-168)
    }
  } else {
    if (c.mode != 2) {
      const float t0 = in_target.x, t1 = in_target.y;
      const float2 qq = in_pred2;
      if (isnan(qq.x) || isnan(qq.y)) err |= APG_FLAG_NAN_PRED;
      const float d0 = __fsub_rn(qq.x, t0), d1 = __fsub_rn(qq.y, t1);
      const float ss = __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1));
      loss_f = __fadd_rn(__fmul_rn(__fdiv_rn(ss, 2.0f), p.mse_scale), p.mse_offset);
      m1 = __fsqrt_rn(ss);       // euclidean_distance (active_regression_env.py:211)
      m2 = __fdiv_rn(ss, 2.0f);  // mse (:212)
      if (c.mode == 0 && p.track)  // 1 - ||pred - target|| / sqrt(4) in float64 (np.sqrt(4) is np.float64)
        pq = (float)fmin(fmax(1.0 - (double)__fsqrt_rn(ss) / 2.0, 0.0), 1.0);
      reinterpret_cast<float2*>(p.target_xy)[n] = make_float2(t0, t1);
    }
    if (c.mode != 0) {
      const float2 nt = reinterpret_cast<const float2*>(c.new_target)[n];
      reinterpret_cast<float2*>(p.st_target)[n] = nt;
      if (c.mode == 2) reinterpret_cast<float2*>(p.target_xy)[n] = nt;
    }
  }
  if (c.mode == 0) {
    const float a0 = in_action.x, a1 = in_action.y;
    if (p.track) p.stamp_q[(size_t)c.t_stamp * p.N + n] = fminf(fmaxf(pq, 0.0f), 1.0f);
    const float m = __fsqrt_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)));
    base = __fmul_rn(-m, 1e-3f);
  }
  if (c.mode != 2) {
    float reward_f;
    // classification: float32 base - float64 loss -> float64; localization stays float32 except on the reset
    // step, whose base reward is np.zeros(N) (float64)
    if (cls) reward_f = (float)((double)base + neg_loss_d);
    else reward_f = (c.mode == 0) ? __fsub_rn(base, loss_f) : (float)(0.0 - (double)loss_f);
    float reward_out = reward_f;
    if (p.sparse) {  // base_reward - loss * weight, weight = float32(terminated)  (sparsify_wrapper.py:60-83)
      const bool w1 = (c.mode == 0) && c.terminated;
      if (cls) reward_out = (float)((double)base - (-neg_loss_d) * (w1 ? 1.0 : 0.0));
      else if (c.mode == 0) reward_out = __fsub_rn(base, __fmul_rn(loss_f, w1 ? 1.0f : 0.0f));
      else reward_out = (float)(0.0 - (double)__fmul_rn(loss_f, 0.0f));
    }
    p.reward[n] = reward_out;
    p.base_reward[n] = base;
    p.loss[n] = loss_f;
    if (c.reward2) c.reward2[n] = reward_out;
    if (c.done2) c.done2[n] = c.terminated;
    // ---- episode statistics (the *VectorLogWrapper loops): real steps only, cleared by the reset step
    if (c.mode == 0) {
      const float a1 = __fadd_rn(in_acc1, m1);
      const float a2 = __fadd_rn(in_acc2, m2);
      const float ar = __fadd_rn(in_accr, reward_f);
      if (cls) {
        int fc = in_fc, lic = in_lic;
        if (m2 > 0.5f) { if (fc < 0) fc = c.t_stamp; } else lic = c.t_stamp;
        p.first_correct[n] = fc;
        p.last_incorrect[n] = lic;
      }
      if (ep_end) {
        const float nt = (float)(c.t_stamp + 1);
        const float avg1 = __fdiv_rn(a1, nt), avg2 = __fdiv_rn(a2, nt);
        p.ep_avg1[n] = avg1; p.ep_final1[n] = m1; p.ep_avg2[n] = avg2; p.ep_final2[n] = m2; p.ep_return[n] = ar;
        atomicAdd(&s_stats[0], 1.0); atomicAdd(&s_stats[1], (double)avg1); atomicAdd(&s_stats[2], (double)m1);
        atomicAdd(&s_stats[3], (double)avg2); atomicAdd(&s_stats[4], (double)m2); atomicAdd(&s_stats[5], (double)ar);
      } else {
        p.acc1[n] = a1; p.acc2[n] = a2; p.acc_ret[n] = ar;
      }
    }
  }
  p.terminated_out[n] = c.terminated;
  p.truncated_out[n] = 0;
  p.time_step[n] = c.time_obs;
}

// Programmatic dependent launch: consecutive vector steps are launched back to back on one stream; with the launch
// attribute set, the blocks of step k+1 are scheduled while step k drains and park here until step k's memory is
// visible.  Every block lets its own dependents go at once (they wait at the same place).
__device__ __forceinline__ void grid_dependency_wait(int pdl) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  if (pdl) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
  }
#endif
}

}  // namespace apg_img
