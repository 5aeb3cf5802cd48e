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
// f32(u8)/255 through a 256-entry shared-memory table filled with __fdiv_rn.
//
// Thread mapping: a block owns E consecutive environments.  Phase A: thread e (< E) does the per-environment
// scalar work and leaves tap rows/columns + float64 weights in shared memory.  Phase B: all threads sweep the
// block's contiguous (E * S0*S1*C) output floats, 4 consecutive floats per thread -> one 128-bit coalesced
// store each; bank taps are byte/float gathers served by L1/L2 (footprint (S0+1) x (S1+1) x C per glimpse).
#include <cmath>
#include <cstring>
#include <vector>

#include "apg_common.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxK = 4096;

struct ImgParams {
  int N, H, W, C, S0, S1, P, K, task, invert, track, words, T;
  int E;             // environments per block
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

__device__ __forceinline__ float load_px(const unsigned char* img, long long o, const float* lut) {
  return lut[__ldg(img + o)];
}
__device__ __forceinline__ float load_px(const float* img, long long o, const float*) { return __ldg(img + o); }

template <typename BankT>
__global__ void __launch_bounds__(kThreads) image_step_kernel(const ImgParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // shared layout: lut[256] | base[2E] (long long) | wy[2][E*S0] wx[2][E*S1] (double) | ti[2][E*S0] tj[2][E*S1] (int)
  float* lut = reinterpret_cast<float*>(smem_raw);
  long long* s_base = reinterpret_cast<long long*>(smem_raw + 1024);
  double* s_wy = reinterpret_cast<double*>(s_base + p.E);
  double* s_wx = s_wy + 2 * p.E * p.S0;
  int* s_ti = reinterpret_cast<int*>(s_wx + 2 * p.E * p.S1);
  int* s_tj = s_ti + 2 * p.E * p.S0;

  const int tid = threadIdx.x;
  const int env0 = blockIdx.x * p.E;
  const int nE = min(p.E, p.N - env0);
  const bool two_sets = (p.task == APG_TASK_LOCALIZATION) && (p.mode != 0);

  if (sizeof(BankT) == 1) lut[tid] = __fdiv_rn((float)tid, 255.0f);

  // ------------------------------------------------------------------ phase A: per-environment scalars
  if (tid < nE) {
    const int n = env0 + tid;
    unsigned err = 0;
    double px, py;
    int idx;
    float base = 0.0f;
    float pq = 0.0f;
    if (p.mode == 0) {
      px = p.pos[2 * n];
      py = p.pos[2 * n + 1];
      idx = p.st_idx[n];
    } else {
      const long long ni = p.new_idx[n];
      idx = (int)ni;
      px = p.new_pos[2 * n];
      py = p.new_pos[2 * n + 1];
      p.st_idx[n] = idx;
      p.index[n] = ni;
    }

    // ---- labels / targets and the loss (the reset step still charges a loss)
    float loss_f = 0.0f;
    double reward_d = 0.0;
    if (p.task == APG_TASK_CLASSIFICATION) {
      int label;
      if (p.mode == 0) {
        label = p.st_label[n];
      } else {
        label = p.bank_labels[idx];
        unsigned char inv = 0;
        if (p.invert) {
          inv = p.new_inv[n];
          if (inv) label = p.K - label - 1;
          p.st_inv[n] = inv;
        }
        p.st_label[n] = label;
      }
      if (p.invert) p.inverted_label[n] = (p.mode == 0) ? 2 : (int)p.st_inv[n];
      p.target_label[n] = label;
      if (p.mode != 2) {
        const float* x = p.pred + (size_t)n * p.K;
        float mx = -INFINITY;
        bool has_nan = false;
        for (int k = 0; k < p.K; ++k) {
          const float v = x[k];
          has_nan |= isnan(v);
          mx = fmaxf(mx, v);
        }
        if (has_nan) err |= APG_FLAG_NAN_PRED;
        if (!isfinite(mx)) mx = 0.0f;
        float s = 0.0f;
        for (int k = 0; k < p.K; ++k) s = __fadd_rn(s, expf(__fsub_rn(x[k], mx)));
        const float tl = __fsub_rn(x[label], mx);
        const float ce = -__fsub_rn(tl, logf(s));           // -log_softmax(x)[label], float32 like scipy
        pq = __fdiv_rn(expf(tl), s);                        // softmax(x)[label]
        const double loss_d = (double)ce * p.ce_scale + p.ce_offset;  // np.float64 scale -> float64 loss
        loss_f = (float)loss_d;
        // base (float32) - loss (float64) -> float64; on the reset step base is 0.0
        reward_d = -loss_d;  // base added below once known
      }
    } else {
      float t0 = 0.f, t1 = 0.f;
      if (p.mode != 2) {
        t0 = p.st_target[2 * n];
        t1 = p.st_target[2 * n + 1];
        const float q0 = p.pred[2 * n], q1 = p.pred[2 * n + 1];
        if (isnan(q0) || isnan(q1)) err |= APG_FLAG_NAN_PRED;
        const float d0 = __fsub_rn(q0, t0), d1 = __fsub_rn(q1, t1);
        const float mse = __fdiv_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), 2.0f);
        loss_f = __fadd_rn(__fmul_rn(mse, p.mse_scale), p.mse_offset);
        if (p.mode == 0) {
          // prediction quality 1 - ||pred - target|| / sqrt(4) (float64 because np.sqrt(4) is np.float64)
          const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)));
          pq = (float)fmin(fmax(1.0 - (double)nrm / 2.0, 0.0), 1.0);
        }
        p.target_xy[2 * n] = t0;
        p.target_xy[2 * n + 1] = t1;
      }
      if (p.mode != 0) {
        t0 = p.new_target[2 * n];
        t1 = p.new_target[2 * n + 1];
        p.st_target[2 * n] = t0;
        p.st_target[2 * n + 1] = t1;
        if (p.mode == 2) {
          p.target_xy[2 * n] = t0;
          p.target_xy[2 * n + 1] = t1;
        }
        // taps of the target glimpse (float32 target promoted to float64)
        const double cx = (double)t0 * p.lim_x, cy = (double)t1 * p.lim_y;
        for (int i = 0; i < p.S0; ++i) {
          const double y = cy + ((double)i - p.half0) * p.scale;
          if (y < -p.offy || y > p.offy) err |= APG_FLAG_OUT_OF_GRID;
          tap(y, p.offy, p.H, s_ti[(p.E + tid) * p.S0 + i], s_wy[(p.E + tid) * p.S0 + i]);
        }
        for (int j = 0; j < p.S1; ++j) {
          const double x = cx + ((double)j - p.half1) * p.scale;
          if (x < -p.offx || x > p.offx) err |= APG_FLAG_OUT_OF_GRID;
          tap(x, p.offx, p.W, s_tj[(p.E + tid) * p.S1 + j], s_wx[(p.E + tid) * p.S1 + j]);
        }
      }
    }

    // ---- move (mode 0): overlay stamp of the PRE-move position, then the float64 position update
    if (p.mode == 0) {
      const float a0 = p.action[2 * n], a1 = p.action[2 * n + 1];
      if (isnan(a0) || isnan(a1)) err |= APG_FLAG_NAN_ACTION;
      if (p.track) {
        const int cxr = (int)rint((px * p.lim_x) * p.rs + p.half_wr);
        const int cyr = (int)rint((py * p.lim_y) * p.rs + p.half_hr);
        p.stamp_xy[(size_t)(p.t_stamp * 2 + 0) * p.N + n] = cxr;
        p.stamp_xy[(size_t)(p.t_stamp * 2 + 1) * p.N + n] = cyr;
        p.stamp_q[(size_t)p.t_stamp * p.N + n] = fminf(fmaxf(pq, 0.0f), 1.0f);
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
    p.pos[2 * n] = px;
    p.pos[2 * n + 1] = py;

    if (p.mode != 2) {
      float reward_f;
      if (p.task == APG_TASK_CLASSIFICATION) reward_f = (float)((double)base + reward_d);
      else reward_f = (p.mode == 0) ? __fsub_rn(base, loss_f) : (float)(0.0 - (double)loss_f);
      p.reward[n] = reward_f;
      p.base_reward[n] = base;
      p.loss[n] = loss_f;
      if (p.reward2) p.reward2[n] = reward_f;
      if (p.done2) p.done2[n] = p.terminated;
    }
    p.terminated_out[n] = p.terminated;
    p.truncated_out[n] = 0;
    p.glimpse_pos[2 * n] = (float)px;
    p.glimpse_pos[2 * n + 1] = (float)py;
    p.time_step[n] = p.time_obs;

    // ---- taps of the observation glimpse at the (new) position
    {
      const double cx = px * p.lim_x, cy = py * p.lim_y;
      for (int i = 0; i < p.S0; ++i) {
        const double y = cy + ((double)i - p.half0) * p.scale;
        if (y < -p.offy || y > p.offy) err |= APG_FLAG_OUT_OF_GRID;
        tap(y, p.offy, p.H, s_ti[tid * p.S0 + i], s_wy[tid * p.S0 + i]);
      }
      for (int j = 0; j < p.S1; ++j) {
        const double x = cx + ((double)j - p.half1) * p.scale;
        if (x < -p.offx || x > p.offx) err |= APG_FLAG_OUT_OF_GRID;
        tap(x, p.offx, p.W, s_tj[tid * p.S1 + j], s_wx[tid * p.S1 + j]);
      }
      s_base[tid] = (long long)idx * p.H * p.W * p.C;
    }

    // ---- visited bitmask: every tap row/column of the glimpse just observed
    if (p.track) {
      unsigned* vm = p.visited + (size_t)n * p.H * p.words;
      if (p.mode != 0)
        for (int r = 0; r < p.H * p.words; ++r) vm[r] = 0u;
      for (int w = 0; w < p.words; ++w) {
        unsigned bits = 0u;
        for (int j = 0; j < p.S1; ++j) {
          const int c0 = s_tj[tid * p.S1 + j];
          if ((c0 >> 5) == w) bits |= 1u << (c0 & 31);
          if (((c0 + 1) >> 5) == w) bits |= 1u << ((c0 + 1) & 31);
        }
        if (bits) {
          int last = -1;  // tap rows are non-decreasing: skip rows already OR-ed
          for (int i = 0; i < p.S0; ++i) {
            const int r0 = s_ti[tid * p.S0 + i];
            if (r0 > last) vm[r0 * p.words + w] |= bits;
            if (r0 + 1 > last) vm[(r0 + 1) * p.words + w] |= bits;
            last = r0 + 1;
          }
        }
      }
    }
    if (err) atomicOr(p.err, err);
  }
  __syncthreads();

  // ------------------------------------------------------------------ phase B: resample, 4 floats / thread
  const BankT* bank = static_cast<const BankT*>(p.bank);
  const int WC = p.W * p.C;
  const int S1C = p.S1 * p.C;
  const int total = nE * p.P;
  const int groups = (total + 3) >> 2;
  const int n_sets = two_sets ? 2 : 1;
  for (int set = 0; set < n_sets; ++set) {
    float* out = (set == 0 ? p.glimpse : p.target_glimpse) + (size_t)env0 * p.P;
    const int so = set * p.E;
    for (int q = tid; q < groups; q += kThreads) {
      int e = q << 2;
      int el = e / p.P;
      int r = e - el * p.P;
      int i = r / S1C;
      int r2 = r - i * S1C;
      int j = r2 / p.C;
      int c = r2 - j * p.C;
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (e + u < total) {
          const BankT* img = bank + s_base[el];
          const int i0 = s_ti[(so + el) * p.S0 + i];
          const int j0 = s_tj[(so + el) * p.S1 + j];
          const double wy = s_wy[(so + el) * p.S0 + i];
          const double wx = s_wx[(so + el) * p.S1 + j];
          const long long o00 = (long long)i0 * WC + j0 * p.C + c;
          const double p00 = (double)load_px(img, o00, lut);
          const double p01 = (double)load_px(img, o00 + p.C, lut);
          const double p10 = (double)load_px(img, o00 + WC, lut);
          const double p11 = (double)load_px(img, o00 + WC + p.C, lut);
          const double omy = 1.0 - wy, omx = 1.0 - wx;
          double acc = 0.0 + p00 * (omy * omx);
          acc = acc + p01 * (omy * wx);
          acc = acc + p10 * (wy * omx);
          acc = acc + p11 * (wy * wx);
          v[u] = (float)fmin(fmax(acc, 0.0), 1.0);
          // advance (c, j, i, el) to the next output float
          if (++c == p.C) {
            c = 0;
            if (++j == p.S1) {
              j = 0;
              if (++i == p.S0) {
                i = 0;
                ++el;
              }
            }
          }
        } else {
          v[u] = 0.0f;
        }
      }
      if (e + 3 < total) {
        *reinterpret_cast<float4*>(out + e) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        for (int u = 0; u < 4 && e + u < total; ++u) out[e + u] = v[u];
      }
    }
  }
}

}  // namespace

// ======================================================================================================
// host side
// ======================================================================================================
struct apg_image_env {
  apg_image_cfg cfg{};
  cudaStream_t stream = nullptr;
  int P = 0, words = 0, E = 0;
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
  p.words = e->words; p.T = c.step_limit; p.E = e->E; p.mode = mode;
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

  const int grid = apg::ceil_div(c.num_envs, e->E);
  if (c.bank_dtype == APG_BANK_U8)
    image_step_kernel<unsigned char><<<grid, kThreads, e->smem, e->stream>>>(p);
  else
    image_step_kernel<float><<<grid, kThreads, e->smem, e->stream>>>(p);
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
  e->E = e->P <= 32 ? 128 : (e->P <= 128 ? 64 : 32);
  e->smem = 1024 + (size_t)e->E * 8 + 2 * (size_t)e->E * (cfg->sensor_h + cfg->sensor_w) * (8 + 4);
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
  if (err == cudaSuccess && e->smem > 48 * 1024) {
    A(cudaFuncSetAttribute(image_step_kernel<unsigned char>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
    A(cudaFuncSetAttribute(image_step_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
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
