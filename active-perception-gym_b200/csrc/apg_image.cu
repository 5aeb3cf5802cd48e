// Fused image-environment step for sm_100a.
//
// One launch performs, for every environment of the shard (reference file:line in /root/reference/ap_gym):
//   * project_sphere + float64 glimpse position update + clip         util.py:94-97, envs/image/image_perception_module.py:206-208
//   * base reward -||a||*1e-3                                          envs/image/image_perception_module.py:209
//   * prediction quality + overlay stamp (pre-move position)          envs/image_classification.py:114-116, envs/image_localization.py:157-159,
//                                                                      envs/image/image_perception_module.py:219-234
//   * bilinear resample of the sensor window from the image bank      envs/image/image_perception_module.py:294-331 (scipy RGI linear, f64 blend)
//   * normalised cross-entropy / MSE loss, reward = base - loss       loss_fn.py:100-110,212-220,261-267, active_perception_vector_env.py:84-111
//   * terminated / truncated, time_step obs, NEXT_STEP all-or-nothing autoreset from a queued payload
//                                                                      envs/image/image_perception_module.py:197-217,236-251
//   * visited bitmask (image resolution) update
//
// Numerics: the file is compiled with -fmad=false, every float op below is a separately rounded IEEE op in
// the order numpy performs it (see oracle/image_oracle.py).  uint8 banks are converted with the reference's
// f32(u8)/255, reproduced exactly by a multiply + two explicit fma corrections (px_to_f64).
//
// Thread mapping: a group of g lanes owns one environment (see image_step_kernel): scalars are spread over /
// replicated across the group, then lane = (column, channel) of the sensor window walks the sensor rows with a
// two-deep software pipeline on the bank loads; loads are contiguous row segments, stores contiguous runs.
#include <cmath>
#include <cstring>
#include <vector>

#include "apg_common.h"

namespace {

constexpr int kThreads = 128;
constexpr int kMaxK = 4096;

struct ImgParams {
  int N, H, W, C, S0, S1, P, K, task, invert, track, words, T;
  int g, log2g;      // lanes per environment (power of two, <= 32)
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
  // overlay
  int* stamp_xy;
  float* stamp_q;
  unsigned* visited;
  unsigned* err;
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

// bank pixel -> the float64 value the reference blends with: float64(float32(u8) / 255) or float64(f32 pixel).
// float32(u8)/255 is reproduced exactly (all 256 values, tests/test_capi_symbols.py) by q0 = u*r,
// q = fma(fma(-255, q0, u), r, q0) with r = float32(1/255): no table, no division.
__device__ __forceinline__ double px_to_f64(unsigned char u) {
  const float r = 0.003921568859368563f;  // float32(1)/float32(255)
  const float uf = (float)u;
  const float q0 = __fmul_rn(uf, r);
  const float e = __fmaf_rn(-255.0f, q0, uf);
  return (double)__fmaf_rn(e, r, q0);
}
__device__ __forceinline__ double px_to_f64(float v) { return (double)v; }

// An environment is owned by a group of g = min(32, pow2 >= S1*C) lanes (a whole warp for a 10x10x3 sensor, 8 lanes
// for a 5x5x1 one); a warp steps 32/g environments in lock-step.
//   phase A  per-environment scalars, spread over the group: lane k reduces logit k (shuffle max/sum), lane i
//            computes the tap of sensor row i, every lane its own column tap; the short scalar chain (move,
//            reward, flags) is evaluated redundantly by all lanes of the group, the leader stores it.
//   phase B  lane = (column, channel) of the sensor window; the group walks the S0 sensor rows.  Every bank load is
//            one contiguous row segment per group, the bottom taps of row i are the top taps of row i+1, and the
//            bottom row is fetched two iterations ahead (software pipeline) so the L2/HBM latency overlaps the
//            float64 blend.  Output rows are stored as contiguous runs.  The blend is the reference's 4-term
//            float64 sum in the same order (bit-exact pixels).
template <typename BankT, int S0T, int CT>
__global__ void __launch_bounds__(kThreads) image_step_kernel(const ImgParams p) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = p.g;
  const int li = lane & (g - 1);
  const int gbase = lane - li;  // first lane of this group inside the warp
  const int grp_in_block = tid >> p.log2g;
  const int n = blockIdx.x * (kThreads >> p.log2g) + grp_in_block;
  const bool act = n < p.N;
  const int nn = act ? n : p.N - 1;  // inactive groups shadow the last environment but store nothing
  const unsigned FULL = 0xffffffffu;
  const bool lead = act && (li == 0);
  const bool two_sets = (p.task == APG_TASK_LOCALIZATION) && (p.mode != 0);

  // ------------------------------------------------------------------ phase A
  unsigned err = 0;
  double px, py;
  int idx;
  float base = 0.0f, pq = 0.0f, loss_f = 0.0f;
  double neg_loss_d = 0.0;
  double tcx = 0.0, tcy = 0.0;
  if (p.mode == 0) {
    const double2 pp = reinterpret_cast<const double2*>(p.pos)[nn];
    px = pp.x;
    py = pp.y;
    idx = p.st_idx[nn];
  } else {
    const long long ni = p.new_idx[nn];
    idx = (int)ni;
    const double2 pp = reinterpret_cast<const double2*>(p.new_pos)[nn];
    px = pp.x;
    py = pp.y;
    if (lead) {
      p.st_idx[nn] = idx;
      p.index[nn] = ni;
    }
  }

  if (p.task == APG_TASK_CLASSIFICATION) {
    int label;
    unsigned char inv = 0;
    if (p.mode == 0) {
      label = p.st_label[nn];
    } else {
      label = p.bank_labels[idx];
      if (p.invert) {
        inv = p.new_inv[nn];
        if (inv) label = p.K - label - 1;
        if (lead) p.st_inv[nn] = inv;
      }
      if (lead) p.st_label[nn] = label;
    }
    if (lead) {
      if (p.invert) p.inverted_label[nn] = (p.mode == 0) ? 2 : (int)inv;
      p.target_label[nn] = label;
    }
    if (p.mode != 2) {
      // log-softmax over K logits spread over the group (float32 like scipy.special.log_softmax)
      const float* x = p.pred + (size_t)nn * p.K;
      float mx = -INFINITY;
      bool has_nan = false;
      for (int k = li; k < p.K; k += g) {
        const float v = x[k];
        has_nan |= isnan(v);
        mx = fmaxf(mx, v);
      }
      for (int o = g >> 1; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
      if (has_nan) err |= APG_FLAG_NAN_PRED;
      if (!isfinite(mx)) mx = 0.0f;
      float s = 0.0f, tl = 0.0f;
      for (int k = li; k < p.K; k += g) {
        const float z = __fsub_rn(x[k], mx);
        if (k == label) tl = z;
        s = __fadd_rn(s, expf(z));
      }
      for (int o = g >> 1; o > 0; o >>= 1) {
        s = __fadd_rn(s, __shfl_xor_sync(FULL, s, o));
        tl = __fadd_rn(tl, __shfl_xor_sync(FULL, tl, o));  // exactly one lane holds a non-zero value
      }
      const float ce = -__fsub_rn(tl, logf(s));
      pq = __fdiv_rn(expf(tl), s);  // softmax(x)[label], overlay colour only
      const double loss_d = (double)ce * p.ce_scale + p.ce_offset;  // np.float64 scale -> float64 loss
      loss_f = (float)loss_d;
      neg_loss_d = -loss_d;
    }
  } else {
    float t0 = 0.f, t1 = 0.f;
    if (p.mode != 2) {
      const float2 tt = reinterpret_cast<const float2*>(p.st_target)[nn];
      t0 = tt.x;
      t1 = tt.y;
      const float2 qq = reinterpret_cast<const float2*>(p.pred)[nn];
      if (isnan(qq.x) || isnan(qq.y)) err |= APG_FLAG_NAN_PRED;
      const float d0 = __fsub_rn(qq.x, t0), d1 = __fsub_rn(qq.y, t1);
      const float ss = __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1));
      loss_f = __fadd_rn(__fmul_rn(__fdiv_rn(ss, 2.0f), p.mse_scale), p.mse_offset);
      if (p.mode == 0)  // 1 - ||pred - target|| / sqrt(4) in float64 (np.sqrt(4) is np.float64), overlay colour only
        pq = (float)fmin(fmax(1.0 - (double)__fsqrt_rn(ss) / 2.0, 0.0), 1.0);
      if (lead) reinterpret_cast<float2*>(p.target_xy)[nn] = make_float2(t0, t1);
    }
    if (p.mode != 0) {
      const float2 nt = reinterpret_cast<const float2*>(p.new_target)[nn];
      if (lead) {
        reinterpret_cast<float2*>(p.st_target)[nn] = nt;
        if (p.mode == 2) reinterpret_cast<float2*>(p.target_xy)[nn] = nt;
      }
      tcx = (double)nt.x * p.lim_x;  // float32 target promoted to float64
      tcy = (double)nt.y * p.lim_y;
    }
  }

  if (p.mode == 0) {
    const float2 aa = reinterpret_cast<const float2*>(p.action)[nn];
    const float a0 = aa.x, a1 = aa.y;
    if (isnan(a0) || isnan(a1)) err |= APG_FLAG_NAN_ACTION;
    if (p.track && lead) {  // overlay stamp of the PRE-move position (image_perception_module.py:196,219-234)
      p.stamp_xy[(size_t)(p.t_stamp * 2 + 0) * p.N + nn] = (int)rint((px * p.lim_x) * p.rs + p.half_wr);
      p.stamp_xy[(size_t)(p.t_stamp * 2 + 1) * p.N + nn] = (int)rint((py * p.lim_y) * p.rs + p.half_hr);
      p.stamp_q[(size_t)p.t_stamp * p.N + nn] = fminf(fmaxf(pq, 0.0f), 1.0f);
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
    base = __fmul_rn(-m, 1e-3f);
  }
  if (lead) {
    reinterpret_cast<double2*>(p.pos)[nn] = make_double2(px, py);
    if (p.mode != 2) {
      float reward_f;
      // classification: float32 base - float64 loss -> float64; localization stays float32 except on the reset
      // step, whose base reward is np.zeros(N) (float64)
      if (p.task == APG_TASK_CLASSIFICATION) reward_f = (float)((double)base + neg_loss_d);
      else reward_f = (p.mode == 0) ? __fsub_rn(base, loss_f) : (float)(0.0 - (double)loss_f);
      p.reward[nn] = reward_f;
      p.base_reward[nn] = base;
      p.loss[nn] = loss_f;
      if (p.reward2) p.reward2[nn] = reward_f;
      if (p.done2) p.done2[nn] = p.terminated;
    }
    p.terminated_out[nn] = p.terminated;
    p.truncated_out[nn] = 0;
    reinterpret_cast<float2*>(p.glimpse_pos)[nn] = make_float2((float)px, (float)py);
    p.time_step[nn] = p.time_obs;
  }

  // ------------------------------------------------------------------ phase B
  const int S0 = S0T ? S0T : p.S0;
  const int C = CT ? CT : p.C;
  const BankT* img = static_cast<const BankT*>(p.bank) + (long long)idx * p.H * p.W * C;
  const int G = p.S1 * C;
  const int WC = p.W * C;
  const int n_chunks = (G + 31) >> 5;
  const int n_sets = two_sets ? 2 : 1;
  unsigned colbits0 = 0u, colbits1 = 0u;
  int vis_row = 0;  // lane li < S0 keeps the tap row of sensor row li (observation glimpse) for the visited mask

  for (int set = 0; set < n_sets; ++set) {
    const double ccx = (set == 0) ? px * p.lim_x : tcx;
    const double ccy = (set == 0) ? py * p.lim_y : tcy;
    float* out = (set == 0 ? p.glimpse : p.target_glimpse) + (size_t)nn * p.P;
    // lane li owns the tap of sensor row li (rows >= g are recomputed on the fly below)
    int my_i0 = 0;
    double my_wy = 0.0;
    if (li < S0) {
      const double y = ccy + ((double)li - p.half0) * p.scale;
      if (y < -p.offy || y > p.offy) err |= APG_FLAG_OUT_OF_GRID;
      tap(y, p.offy, p.H, my_i0, my_wy);
    }
    if (set == 0) vis_row = my_i0;
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const int jc = (chunk << 5) + li;
      const bool col_ok = jc < G;
      const int j = (C == 1) ? jc : ((C == 3) ? (int)(((unsigned)jc * 43691u) >> 17) : jc / C);
      const int c = jc - j * C;
      int j0 = 0;
      double wx = 0.0;
      {
        const double x = ccx + ((double)min(j, p.S1 - 1) - p.half1) * p.scale;
        if (col_ok && (x < -p.offx || x > p.offx)) err |= APG_FLAG_OUT_OF_GRID;
        tap(x, p.offx, p.W, j0, wx);
      }
      if (set == 0 && col_ok) {
        if ((j0 >> 5) == 0) colbits0 |= 1u << (j0 & 31);
        else if ((j0 >> 5) == 1) colbits1 |= 1u << (j0 & 31);
        if (((j0 + 1) >> 5) == 0) colbits0 |= 1u << ((j0 + 1) & 31);
        else if (((j0 + 1) >> 5) == 1) colbits1 |= 1u << ((j0 + 1) & 31);
      }
      const double omx = 1.0 - wx;
      const BankT* colp = img + (j0 * C + (col_ok ? c : 0));
      float* outc = out + jc;
      const bool st_ok = act && col_ok;

      if (S0T > 0) {
        // ---- compile-time sensor height: every bank load of the glimpse is issued before the first blend
        int i0s[S0T > 0 ? S0T : 1];
#pragma unroll
        for (int i = 0; i < S0T; ++i) i0s[i] = __shfl_sync(FULL, my_i0, gbase + i);
        BankT bot0[S0T > 0 ? S0T : 1], bot1[S0T > 0 ? S0T : 1];
        const BankT* r0 = colp + i0s[0] * WC;
        const BankT top0 = __ldg(r0), top1 = __ldg(r0 + C);
#pragma unroll
        for (int i = 0; i < S0T; ++i) {
          const BankT* rp = colp + (i0s[i] + 1) * WC;
          bot0[i] = __ldg(rp);
          bot1[i] = __ldg(rp + C);
        }
        double p00 = px_to_f64(top0), p01 = px_to_f64(top1);
#pragma unroll
        for (int i = 0; i < S0T; ++i) {
          const double wy = __shfl_sync(FULL, my_wy, gbase + i);
          if (i > 0 && i0s[i] != i0s[i - 1] + 1) {  // clamped border: previous bottom row is not this top row
            const BankT* rp = colp + i0s[i] * WC;
            p00 = px_to_f64(__ldg(rp));
            p01 = px_to_f64(__ldg(rp + C));
          }
          const double p10 = px_to_f64(bot0[i]), p11 = px_to_f64(bot1[i]);
          const double omy = 1.0 - wy;
          double acc = p00 * (omy * omx);  // 0.0 + x == x for x >= 0
          acc = acc + p01 * (omy * wx);
          acc = acc + p10 * (wy * omx);
          acc = acc + p11 * (wy * wx);
          // uint8 pixels are in [0,1] and the weights sum to 1 (+ a few ulp): float32(acc) is already clipped
          const float v = (sizeof(BankT) == 1) ? (float)acc : (float)fmin(fmax(acc, 0.0), 1.0);
          if (st_ok) outc[i * G] = v;
          p00 = p10;
          p01 = p11;
        }
      } else {
        // ---- generic sensor height: two-deep software pipeline over the rows
        auto row_tap = [&](int i, int& i0, double& wy) {
          if (S0 <= g) {
            i0 = __shfl_sync(FULL, my_i0, gbase + min(i, g - 1));
            wy = __shfl_sync(FULL, my_wy, gbase + min(i, g - 1));
          } else {
            tap(ccy + ((double)i - p.half0) * p.scale, p.offy, p.H, i0, wy);
          }
        };
        int i0_cur, i0_n1 = 0, i0_n2 = 0;
        double wy_cur, wy_n1 = 0.0, wy_n2 = 0.0;
        row_tap(0, i0_cur, wy_cur);
        row_tap(1, i0_n1, wy_n1);
        BankT top0 = __ldg(colp + i0_cur * WC);
        BankT top1 = __ldg(colp + i0_cur * WC + C);
        BankT bot0_a = __ldg(colp + (i0_cur + 1) * WC);
        BankT bot1_a = __ldg(colp + (i0_cur + 1) * WC + C);
        BankT bot0_b = bot0_a, bot1_b = bot1_a;
        if (1 < S0) {
          bot0_b = __ldg(colp + (i0_n1 + 1) * WC);
          bot1_b = __ldg(colp + (i0_n1 + 1) * WC + C);
        }
        double p00 = px_to_f64(top0), p01 = px_to_f64(top1);
        int prev_i0 = i0_cur - 1;
        for (int i = 0; i < S0; ++i) {
          BankT bot0_c = bot0_b, bot1_c = bot1_b;
          if (i + 2 < S0) {
            row_tap(i + 2, i0_n2, wy_n2);
            bot0_c = __ldg(colp + (i0_n2 + 1) * WC);
            bot1_c = __ldg(colp + (i0_n2 + 1) * WC + C);
          }
          if (i0_cur != prev_i0 + 1) {
            p00 = px_to_f64(__ldg(colp + i0_cur * WC));
            p01 = px_to_f64(__ldg(colp + i0_cur * WC + C));
          }
          const double p10 = px_to_f64(bot0_a), p11 = px_to_f64(bot1_a);
          const double omy = 1.0 - wy_cur;
          double acc = 0.0 + p00 * (omy * omx);
          acc = acc + p01 * (omy * wx);
          acc = acc + p10 * (wy_cur * omx);
          acc = acc + p11 * (wy_cur * wx);
          if (st_ok) outc[i * G] = (float)fmin(fmax(acc, 0.0), 1.0);
          p00 = p10;
          p01 = p11;
          prev_i0 = i0_cur;
          i0_cur = i0_n1; wy_cur = wy_n1;
          i0_n1 = i0_n2; wy_n1 = wy_n2;
          bot0_a = bot0_b; bot1_a = bot1_b;
          bot0_b = bot0_c; bot1_b = bot1_c;
        }
      }
    }
  }

  // ------------------------------------------------------------------ visited bitmask (observation glimpse)
  if (p.track) {
    unsigned* vm = p.visited + (size_t)nn * p.H * p.words;
    if (p.mode != 0) {
      if (act)
        for (int h = li; h < p.H * p.words; h += g) vm[h] = 0u;
      __syncwarp(FULL);
    }
    for (int w = 0; w < p.words; ++w) {
      unsigned my = (w == 0) ? colbits0 : ((w == 1) ? colbits1 : 0u);
      if (w >= 2) {  // wide images: recompute this word's column bits
        for (int jj = li; jj < p.S1; jj += g) {
          int j0;
          double wx;
          tap(px * p.lim_x + ((double)jj - p.half1) * p.scale, p.offx, p.W, j0, wx);
          if ((j0 >> 5) == w) my |= 1u << (j0 & 31);
          if (((j0 + 1) >> 5) == w) my |= 1u << ((j0 + 1) & 31);
        }
      }
      unsigned bits = my;
      if (g == 32) bits = __reduce_or_sync(FULL, my);
      else
        for (int o = g >> 1; o > 0; o >>= 1) bits |= __shfl_xor_sync(FULL, bits, o);
      // rows {i0, i0+1} of every sensor row; duplicate rows store identical values (benign)
      for (int r0 = 0; r0 < 2 * S0; r0 += g) {
        const int rr = r0 + li;
        const int i = (rr < S0) ? rr : rr - S0;
        int row;
        if (S0 <= g) {
          row = __shfl_sync(FULL, vis_row, gbase + min(max(i, 0), g - 1));
        } else {
          double wy;
          tap(py * p.lim_y + ((double)min(i, S0 - 1) - p.half0) * p.scale, p.offy, p.H, row, wy);
        }
        if (rr >= S0) row += 1;
        if (act && bits && rr < 2 * S0) vm[row * p.words + w] |= bits;
      }
    }
  }
  if (err && act) atomicOr(p.err, err);
}

}  // namespace

// ======================================================================================================
// host side
// ======================================================================================================
struct apg_image_env {
  apg_image_cfg cfg{};
  cudaStream_t stream = nullptr;
  int P = 0, words = 0, g = 0;
  size_t smem = 0;
  // bank
  void* d_bank = nullptr;
  int32_t* d_bank_labels = nullptr;
  int64_t n_items = 0;
  bool own_bank = false;
  // outputs / state
  apg_image_buffers b{};
  int32_t* d_st_idx = nullptr;
  int32_t* d_st_label = nullptr;
  uint8_t* d_st_inv = nullptr;
  unsigned* d_err = nullptr;
  // payload ring
  struct Slot {
    int64_t* d_idx = nullptr;
    double* d_pos = nullptr;
    float* d_target = nullptr;
    uint8_t* d_inv = nullptr;
    unsigned char* h_pin = nullptr;  // pinned staging: idx | pos | target | inv
    cudaEvent_t copied = nullptr;
  };
  std::vector<Slot> ring;
  int ring_head = 0, ring_count = 0;
  // staging for *_host
  float* d_in_action = nullptr;
  float* d_in_pred = nullptr;
  // episode clock (all-or-nothing, shared by every env of the vector env)
  int t = 0;
  bool has_reset = false;
  bool needs_reset = false;
  int pred_width = 0;
};

namespace {

size_t slot_bytes(const apg_image_env* e) {
  const size_t N = e->cfg.num_envs;
  return N * 8 + N * 16 + N * 8 + N;
}

int stage_payload(apg_image_env* e, apg_image_env::Slot& s, const int64_t* h_idx, const double* h_pos,
                  const float* h_target, const uint8_t* h_invert) {
  const size_t N = e->cfg.num_envs;
  APG_CHECK(h_idx && h_pos, APG_ERR_INVALID, "reset payload needs h_idx and h_pos");
  APG_CHECK(e->cfg.task != APG_TASK_LOCALIZATION || h_target, APG_ERR_INVALID,
            "localization reset payload needs h_target");
  APG_CHECK(!e->cfg.randomly_invert_labels || h_invert, APG_ERR_INVALID,
            "randomly_invert_labels reset payload needs h_invert");
  for (size_t i = 0; i < N; ++i)
    APG_CHECK(h_idx[i] >= 0 && h_idx[i] < e->n_items, APG_ERR_INVALID, "bank index %lld out of range [0, %lld)",
              (long long)h_idx[i], (long long)e->n_items);
  APG_CUDA(cudaEventSynchronize(s.copied));  // previous async copy out of the pinned slot is done
  unsigned char* h = s.h_pin;
  memcpy(h, h_idx, N * 8);
  memcpy(h + N * 8, h_pos, N * 16);
  if (h_target) memcpy(h + N * 24, h_target, N * 8);
  if (h_invert) memcpy(h + N * 32, h_invert, N);
  APG_CUDA(cudaMemcpyAsync(s.d_idx, h, N * 8, cudaMemcpyHostToDevice, e->stream));
  APG_CUDA(cudaMemcpyAsync(s.d_pos, h + N * 8, N * 16, cudaMemcpyHostToDevice, e->stream));
  if (h_target) APG_CUDA(cudaMemcpyAsync(s.d_target, h + N * 24, N * 8, cudaMemcpyHostToDevice, e->stream));
  if (h_invert) APG_CUDA(cudaMemcpyAsync(s.d_inv, h + N * 32, N, cudaMemcpyHostToDevice, e->stream));
  APG_CUDA(cudaEventRecord(s.copied, e->stream));
  return APG_OK;
}

// mode: 0 move, 1 autoreset, 2 explicit reset
int launch(apg_image_env* e, int mode, const float* d_action, const float* d_pred,
           const apg_image_env::Slot* slot, float* reward2, uint8_t* done2) {
  const apg_image_cfg& c = e->cfg;
  ImgParams p{};
  p.N = c.num_envs; p.H = c.height; p.W = c.width; p.C = c.channels; p.S0 = c.sensor_h; p.S1 = c.sensor_w;
  p.P = e->P; p.K = c.num_classes; p.task = c.task; p.invert = c.randomly_invert_labels; p.track = c.track_visited;
  p.words = e->words; p.T = c.step_limit; p.g = e->g; p.log2g = 31 - __builtin_clz(e->g); p.mode = mode;
  int t_after;
  if (mode == 0) {
    p.t_stamp = e->t;
    t_after = e->t + 1;
    p.terminated = (t_after >= c.step_limit) ? 1 : 0;
  } else {
    p.t_stamp = 0;
    t_after = 0;
    p.terminated = 0;
  }
  p.time_obs = (float)(((double)t_after / (double)c.step_limit) * 2.0 - 1.0);
  p.scale = c.sensor_scale;
  p.msl_x = c.max_step_length[0]; p.msl_y = c.max_step_length[1];
  // sensor_pos_lim_pixels (image_perception_module.py:419-423): x uses sensor_size[0], y uses sensor_size[1]
  p.lim_x = ((double)c.width - 1.0) / 2.0 - ((double)c.sensor_h * c.sensor_scale - 1.0) / 2.0;
  p.lim_y = ((double)c.height - 1.0) / 2.0 - ((double)c.sensor_w * c.sensor_scale - 1.0) / 2.0;
  p.offy = ((double)c.height - 1.0) / 2.0; p.offx = ((double)c.width - 1.0) / 2.0;
  p.half0 = ((double)c.sensor_h - 1.0) / 2.0; p.half1 = ((double)c.sensor_w - 1.0) / 2.0;
  // render geometry of the overlay (image_perception_module.py:167-171)
  const int render_w = c.sensor_w > 128 ? c.sensor_w : 128;
  p.rs = (double)render_w / (double)c.width;
  const int render_h = (int)nearbyint(p.rs * (double)c.height);
  p.half_wr = (double)render_w / 2.0; p.half_hr = (double)render_h / 2.0;
  if (c.task == APG_TASK_CLASSIFICATION) {
    p.ce_scale = 1.0 / (std::log((double)c.num_classes) - 0.0);
    p.ce_offset = -0.0 * p.ce_scale;
  } else {
    const double std_ = 2.0 / std::sqrt(12.0);
    const double blind = std_ * std_;
    const double s = 1.0 / (blind - 0.0);
    p.mse_scale = (float)s;   // python-float scale applied to a float32 array stays float32
    p.mse_offset = (float)(-0.0 * s);
  }
  p.bank = e->d_bank; p.bank_labels = e->d_bank_labels; p.action = d_action; p.pred = d_pred;
  if (slot) {
    p.new_idx = reinterpret_cast<const long long*>(slot->d_idx); p.new_pos = slot->d_pos;
    p.new_target = slot->d_target; p.new_inv = slot->d_inv;
  }
  p.pos = e->b.d_pos; p.st_idx = e->d_st_idx; p.st_label = e->d_st_label; p.st_target = e->b.d_cur_target_xy;
  p.st_inv = e->d_st_inv;
  p.glimpse = e->b.d_glimpse; p.target_glimpse = e->b.d_target_glimpse; p.glimpse_pos = e->b.d_glimpse_pos;
  p.time_step = e->b.d_time_step; p.inverted_label = e->b.d_inverted_label; p.reward = e->b.d_reward;
  p.base_reward = e->b.d_base_reward; p.loss = e->b.d_loss; p.target_label = e->b.d_target_label;
  p.target_xy = e->b.d_target_xy; p.terminated_out = e->b.d_terminated; p.truncated_out = e->b.d_truncated;
  p.index = reinterpret_cast<long long*>(e->b.d_index); p.reward2 = reward2; p.done2 = done2;
  p.stamp_xy = e->b.d_stamp_xy; p.stamp_q = e->b.d_stamp_q; p.visited = e->b.d_visited; p.err = e->d_err;

  const int grid = apg::ceil_div(c.num_envs, kThreads / e->g);
#define APG_IMG_LAUNCH(S0T, CT)                                                                   \
  do {                                                                                            \
    if (c.bank_dtype == APG_BANK_U8)                                                              \
      image_step_kernel<unsigned char, S0T, CT><<<grid, kThreads, 0, e->stream>>>(p);             \
    else                                                                                          \
      image_step_kernel<float, S0T, CT><<<grid, kThreads, 0, e->stream>>>(p);                     \
  } while (0)
  // compile-time sensor heights for the registered environments (5x5 and 10x10 sensors); the unrolled path
  // needs every sensor row's tap in a lane of the group (S0 <= g)
  const bool fast = c.sensor_h <= e->g;
  if (fast && c.sensor_h == 5 && c.channels == 1) APG_IMG_LAUNCH(5, 1);
  else if (fast && c.sensor_h == 5 && c.channels == 3) APG_IMG_LAUNCH(5, 3);
  else if (fast && c.sensor_h == 10 && c.channels == 3) APG_IMG_LAUNCH(10, 3);
  else if (fast && c.sensor_h == 10 && c.channels == 1) APG_IMG_LAUNCH(10, 1);
  else APG_IMG_LAUNCH(0, 0);
#undef APG_IMG_LAUNCH
  APG_CUDA(cudaGetLastError());
  apg::count_launch();
  e->t = t_after;
  e->needs_reset = (mode == 0) && p.terminated;
  return APG_OK;
}

int step_once(apg_image_env* e, const float* d_action, const float* d_pred, float* reward2, uint8_t* done2) {
  if (e->needs_reset) {
    APG_CHECK(e->ring_count > 0, APG_ERR_NEED_PAYLOAD,
              "autoreset step is due but no reset payload is queued (apg_image_queue_reset_payload)");
    apg_image_env::Slot& s = e->ring[e->ring_head];
    int rc = launch(e, 1, d_action, d_pred, &s, reward2, done2);
    if (rc != APG_OK) return rc;
    e->ring_head = (e->ring_head + 1) % (int)e->ring.size();
    e->ring_count -= 1;
    return APG_OK;
  }
  return launch(e, 0, d_action, d_pred, nullptr, reward2, done2);
}

}  // namespace

extern "C" {

int apg_image_create(const apg_image_cfg* cfg, apg_image_env** out) {
  APG_CHECK(cfg && out, APG_ERR_INVALID, "apg_image_create: null argument");
  *out = nullptr;
  APG_CHECK(cfg->num_envs > 0, APG_ERR_INVALID, "num_envs must be positive");
  APG_CHECK(cfg->task == APG_TASK_CLASSIFICATION || cfg->task == APG_TASK_LOCALIZATION, APG_ERR_INVALID, "bad task");
  APG_CHECK(cfg->channels == 1 || cfg->channels == 3, APG_ERR_INVALID,
            "Target channels must be either 1 or 3 but is %d.", cfg->channels);
  APG_CHECK(cfg->sensor_h > 0 && cfg->sensor_w > 0 && cfg->sensor_scale > 0, APG_ERR_INVALID, "bad sensor geometry");
  APG_CHECK((double)cfg->height >= cfg->sensor_h * cfg->sensor_scale &&
                (double)cfg->width >= cfg->sensor_w * cfg->sensor_scale,
            APG_ERR_INVALID, "Image size (%d, %d) cannot be smaller than effective sensor size (%g, %g).",
            cfg->height, cfg->width, cfg->sensor_h * cfg->sensor_scale, cfg->sensor_w * cfg->sensor_scale);
  APG_CHECK(cfg->height >= 2 && cfg->width >= 2, APG_ERR_INVALID, "image must be at least 2x2");
  APG_CHECK(cfg->step_limit > 0, APG_ERR_INVALID, "step_limit must be positive");
  APG_CHECK(cfg->bank_dtype == APG_BANK_U8 || cfg->bank_dtype == APG_BANK_F32, APG_ERR_INVALID, "bad bank dtype");
  if (cfg->task == APG_TASK_CLASSIFICATION)
    APG_CHECK(cfg->num_classes >= 2 && cfg->num_classes <= kMaxK, APG_ERR_INVALID, "num_classes must be in [2, %d]", kMaxK);
  int ndev = 0;
  cudaError_t de = cudaGetDeviceCount(&ndev);
  if (de != cudaSuccess || ndev <= 0)
    return apg::fail(APG_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                     cudaGetErrorString(de));
  APG_CHECK(cfg->device >= 0 && cfg->device < ndev, APG_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
  APG_CUDA(cudaSetDevice(cfg->device));

  apg_image_env* e = new apg_image_env();
  e->cfg = *cfg;
  const size_t N = cfg->num_envs;
  e->P = cfg->sensor_h * cfg->sensor_w * cfg->channels;
  e->words = (cfg->width + 31) / 32;
  {
    int g = 4;
    while (g < cfg->sensor_w * cfg->channels && g < 32) g <<= 1;
    e->g = g;
  }
  e->smem = 0;
  e->pred_width = cfg->task == APG_TASK_CLASSIFICATION ? cfg->num_classes : 2;
  const bool loc = cfg->task == APG_TASK_LOCALIZATION;
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
  A(apg::dev_alloc(&e->b.d_glimpse, N * e->P));
  if (loc) A(apg::dev_alloc(&e->b.d_target_glimpse, N * e->P));
  A(apg::dev_alloc(&e->b.d_glimpse_pos, N * 2));
  A(apg::dev_alloc(&e->b.d_time_step, N));
  if (cfg->randomly_invert_labels) A(apg::dev_alloc(&e->b.d_inverted_label, N));
  A(apg::dev_alloc(&e->b.d_reward, N));
  A(apg::dev_alloc(&e->b.d_base_reward, N));
  A(apg::dev_alloc(&e->b.d_loss, N));
  if (!loc) A(apg::dev_alloc(&e->b.d_target_label, N));
  if (loc) A(apg::dev_alloc(&e->b.d_target_xy, N * 2));
  A(apg::dev_alloc(&e->b.d_terminated, N));
  A(apg::dev_alloc(&e->b.d_truncated, N));
  A(apg::dev_alloc(&e->b.d_index, N));
  A(apg::dev_alloc(&e->b.d_pos, N * 2));
  if (loc) A(apg::dev_alloc(&e->b.d_cur_target_xy, N * 2));
  if (cfg->track_visited) {
    A(apg::dev_alloc(&e->b.d_visited, N * cfg->height * e->words));
    A(apg::dev_alloc(&e->b.d_stamp_xy, (size_t)cfg->step_limit * 2 * N));
    A(apg::dev_alloc(&e->b.d_stamp_q, (size_t)cfg->step_limit * N));
  }
  A(apg::dev_alloc(&e->d_st_idx, N));
  A(apg::dev_alloc(&e->d_st_label, N));
  A(apg::dev_alloc(&e->d_st_inv, N));
  A(apg::dev_alloc(&e->d_err, 1));
  A(apg::dev_alloc(&e->d_in_action, N * 2));
  A(apg::dev_alloc(&e->d_in_pred, N * e->pred_width));
  int depth = cfg->payload_ring <= 0 ? APG_PAYLOAD_RING_DEFAULT : cfg->payload_ring;
  if (depth > APG_PAYLOAD_RING_MAX) depth = APG_PAYLOAD_RING_MAX;
  e->ring.resize(depth);
  for (auto& s : e->ring) {
    A(apg::dev_alloc(&s.d_idx, N));
    A(apg::dev_alloc(&s.d_pos, N * 2));
    A(apg::dev_alloc(&s.d_target, N * 2));
    A(apg::dev_alloc(&s.d_inv, N));
    A(cudaMallocHost(reinterpret_cast<void**>(&s.h_pin), slot_bytes(e)));
    A(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
  }
  if (err != cudaSuccess) {
    apg_image_destroy(e);
    return apg::fail(APG_ERR_CUDA, "apg_image_create: allocation failed: %s", cudaGetErrorString(err));
  }
  *out = e;
  return APG_OK;
}

void apg_image_destroy(apg_image_env* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  if (e->own_bank) cudaFree(e->d_bank);
  if (e->own_bank) cudaFree(e->d_bank_labels);
  apg::dev_free(e->b.d_glimpse); apg::dev_free(e->b.d_target_glimpse); apg::dev_free(e->b.d_glimpse_pos);
  apg::dev_free(e->b.d_time_step); apg::dev_free(e->b.d_inverted_label); apg::dev_free(e->b.d_reward);
  apg::dev_free(e->b.d_base_reward); apg::dev_free(e->b.d_loss); apg::dev_free(e->b.d_target_label);
  apg::dev_free(e->b.d_target_xy); apg::dev_free(e->b.d_terminated); apg::dev_free(e->b.d_truncated);
  apg::dev_free(e->b.d_index); apg::dev_free(e->b.d_pos); apg::dev_free(e->b.d_cur_target_xy);
  apg::dev_free(e->b.d_visited); apg::dev_free(e->b.d_stamp_xy); apg::dev_free(e->b.d_stamp_q);
  apg::dev_free(e->d_st_idx); apg::dev_free(e->d_st_label); apg::dev_free(e->d_st_inv); apg::dev_free(e->d_err);
  apg::dev_free(e->d_in_action); apg::dev_free(e->d_in_pred);
  for (auto& s : e->ring) {
    apg::dev_free(s.d_idx); apg::dev_free(s.d_pos); apg::dev_free(s.d_target); apg::dev_free(s.d_inv);
    if (s.h_pin) cudaFreeHost(s.h_pin);
    if (s.copied) cudaEventDestroy(s.copied);
  }
  delete e;
}

int apg_image_set_stream(apg_image_env* e, void* cuda_stream) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  e->stream = static_cast<cudaStream_t>(cuda_stream);
  return APG_OK;
}

int apg_image_set_bank(apg_image_env* e, const void* bank, const int32_t* labels, int64_t n_items, int on_device) {
  APG_CHECK(e && bank && labels && n_items > 0, APG_ERR_INVALID, "apg_image_set_bank: bad argument");
  APG_CHECK(n_items < (1ll << 31), APG_ERR_INVALID, "bank too large");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  if (e->own_bank) {
    cudaFree(e->d_bank);
    cudaFree(e->d_bank_labels);
  }
  e->d_bank = nullptr; e->d_bank_labels = nullptr; e->own_bank = false;
  const size_t px = (size_t)e->cfg.height * e->cfg.width * e->cfg.channels;
  const size_t bytes = (size_t)n_items * px * (e->cfg.bank_dtype == APG_BANK_U8 ? 1 : 4);
  if (on_device) {
    e->d_bank = const_cast<void*>(bank);
    e->d_bank_labels = const_cast<int32_t*>(labels);
  } else {
    APG_CUDA(cudaMalloc(&e->d_bank, bytes));
    e->own_bank = true;
    APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_bank_labels), (size_t)n_items * 4));
    APG_CUDA(cudaMemcpy(e->d_bank, bank, bytes, cudaMemcpyHostToDevice));
    APG_CUDA(cudaMemcpy(e->d_bank_labels, labels, (size_t)n_items * 4, cudaMemcpyHostToDevice));
  }
  e->n_items = n_items;
  return APG_OK;
}

int apg_image_reset(apg_image_env* e, const int64_t* h_idx, const double* h_pos, const float* h_target,
                    const uint8_t* h_invert) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CHECK(e->d_bank, APG_ERR_STATE, "apg_image_reset before apg_image_set_bank");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  // an explicit reset drops queued autoreset payloads (the reference re-seeds its streams on reset(seed))
  e->ring_head = 0;
  e->ring_count = 0;
  apg_image_env::Slot& s = e->ring[0];
  int rc = stage_payload(e, s, h_idx, h_pos, h_target, h_invert);
  if (rc != APG_OK) return rc;
  rc = launch(e, 2, nullptr, nullptr, &s, nullptr, nullptr);
  if (rc != APG_OK) return rc;
  e->has_reset = true;
  e->needs_reset = false;
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

int apg_image_queue_reset_payload(apg_image_env* e, const int64_t* h_idx, const double* h_pos, const float* h_target,
                                  const uint8_t* h_invert) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CHECK(e->d_bank, APG_ERR_STATE, "payload queued before apg_image_set_bank");
  APG_CHECK(e->ring_count < (int)e->ring.size(), APG_ERR_STATE, "reset payload ring is full (%d pending)", e->ring_count);
  APG_CUDA(cudaSetDevice(e->cfg.device));
  apg_image_env::Slot& s = e->ring[(e->ring_head + e->ring_count) % (int)e->ring.size()];
  int rc = stage_payload(e, s, h_idx, h_pos, h_target, h_invert);
  if (rc != APG_OK) return rc;
  e->ring_count += 1;
  return APG_OK;
}

int apg_image_pending_payloads(apg_image_env* e) { return e ? e->ring_count : APG_ERR_INVALID; }

int apg_image_step(apg_image_env* e, const float* d_action, const float* d_pred) {
  APG_CHECK(e && d_action && d_pred, APG_ERR_INVALID, "apg_image_step: null argument");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "step before reset");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  return step_once(e, d_action, d_pred, nullptr, nullptr);
}

int apg_image_rollout(apg_image_env* e, const float* d_actions, const float* d_preds, int32_t n_steps,
                      float* d_rewards, uint8_t* d_dones) {
  APG_CHECK(e && d_actions && d_preds && n_steps >= 0, APG_ERR_INVALID, "apg_image_rollout: bad argument");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "rollout before reset");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  // count the autoreset steps inside the window up front so that a rollout never stops half way
  {
    int t = e->t, resets = 0;
    bool need = e->needs_reset;
    for (int k = 0; k < n_steps; ++k) {
      if (need) { resets++; t = 0; need = false; }
      else { t++; need = t >= e->cfg.step_limit; }
    }
    APG_CHECK(resets <= e->ring_count, APG_ERR_NEED_PAYLOAD, "rollout of %d steps needs %d reset payloads, %d queued",
              n_steps, resets, e->ring_count);
  }
  const size_t N = e->cfg.num_envs;
  for (int k = 0; k < n_steps; ++k) {
    int rc = step_once(e, d_actions + (size_t)k * N * 2, d_preds + (size_t)k * N * e->pred_width,
                       d_rewards ? d_rewards + (size_t)k * N : nullptr, d_dones ? d_dones + (size_t)k * N : nullptr);
    if (rc != APG_OK) return rc;
  }
  return APG_OK;
}

int apg_image_step_host(apg_image_env* e, const float* h_action, const float* h_pred, float* h_reward,
                        uint8_t* h_terminated, uint8_t* h_truncated, float* h_glimpse) {
  APG_CHECK(e && h_action && h_pred, APG_ERR_INVALID, "apg_image_step_host: null argument");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "step before reset");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  const size_t N = e->cfg.num_envs;
  APG_CUDA(cudaMemcpyAsync(e->d_in_action, h_action, N * 2 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  APG_CUDA(cudaMemcpyAsync(e->d_in_pred, h_pred, N * e->pred_width * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  int rc = step_once(e, e->d_in_action, e->d_in_pred, nullptr, nullptr);
  if (rc != APG_OK) return rc;
  if (h_reward) APG_CUDA(cudaMemcpyAsync(h_reward, e->b.d_reward, N * 4, cudaMemcpyDeviceToHost, e->stream));
  if (h_terminated) APG_CUDA(cudaMemcpyAsync(h_terminated, e->b.d_terminated, N, cudaMemcpyDeviceToHost, e->stream));
  if (h_truncated) APG_CUDA(cudaMemcpyAsync(h_truncated, e->b.d_truncated, N, cudaMemcpyDeviceToHost, e->stream));
  if (h_glimpse) APG_CUDA(cudaMemcpyAsync(h_glimpse, e->b.d_glimpse, N * e->P * 4, cudaMemcpyDeviceToHost, e->stream));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

int apg_image_get_buffers(apg_image_env* e, apg_image_buffers* out) {
  APG_CHECK(e && out, APG_ERR_INVALID, "null argument");
  *out = e->b;
  out->stamp_count = e->needs_reset ? e->cfg.step_limit : e->t;
  out->elapsed = e->t;
  out->needs_reset = e->needs_reset ? 1 : 0;
  return APG_OK;
}

int apg_image_poll_error(apg_image_env* e, uint32_t* flags, int clear) {
  APG_CHECK(e && flags, APG_ERR_INVALID, "null argument");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaMemcpyAsync(flags, e->d_err, 4, cudaMemcpyDeviceToHost, e->stream));
  if (clear) APG_CUDA(cudaMemsetAsync(e->d_err, 0, 4, e->stream));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

int apg_image_sync(apg_image_env* e) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

}  // extern "C"
