// Image step for SMALL glimpses on uint8 single-channel banks at sensor_scale == 1 -- the MNIST-v0 shape of
// BASELINE.json (28x28x1 bank, 5x5 sensor, classification) -- as ONE persistent launch per rollout.  Reference path:
//   ImagePerceptionModule.step / get_glimpse      ap_gym/envs/image/image_perception_module.py:191-217,294-331
//   ImageClassificationVectorEnv._step            ap_gym/envs/image_classification.py:113-127
//   CrossEntropyLossFn(.normalized).numpy         ap_gym/loss_fn.py:100-110,212-220
//
// thread = environment for the WHOLE rollout: the float64 position, bank index, label and the episode accumulators live
// in registers across the K steps of the launch (loaded once, written back once), the visited bitmask lives in
// thread-private shared-memory words.  Environments are independent, so nothing in the kernel synchronises more than one
// warp: every warp free-runs through its K steps, different warps sit in different phases of a step and hide each
// other's memory latency.  Per step and thread:
//   1. action -> projected move -> float64 position; the S0 + S1 bilinear taps (one full scipy find_indices per axis, the
//      others verified against the consecutive-index assumption);
//   2. the (S0+1) x (S1+1) bank footprint is requested as 3 aligned 32-bit words per row straight into registers
//      (MNIST rows are one 32-byte sector each) -- the loads are in flight during
//   3. the cross-entropy loss, reward, flags, overlay stamp and episode statistics of the step;
//   4. the reference's 4-term float64 blend in scipy's order from registers (bit-exact pixels), every footprint pixel
//      converted exactly once;
//   5. the 25 floats are parked in the warp's shared-memory staging area and the warp streams its 32 glimpses out as
//      one contiguous run of 128-bit stores (3200 bytes per warp and step).
// The payloads of the autoresets inside the window are drawn (in-stream, csrc/apg_payload.cu) before the launch.
#include "apg_image_small.h"

namespace apg_img {
namespace {

constexpr int kSmallThreads = 64;  // 2 warps per block: 65536 envs = 1024 blocks = 6.9 per SM, all resident (one wave)

__device__ __forceinline__ float norm2f(float a, float b) {
  return __fsqrt_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
}

// Taps of one axis under the consecutive-index assumption: tap 0 is scipy's find_indices in full (tap_row with the
// unclamped last index, see apg_image_dev.cuh), tap t > 0 is assumed to sit at index k0 + t and VERIFIED: its weight
// w = coord - grid[k0 + t] is computed exactly like the reference's and must lie in [0, 1) -- precisely the condition
// under which find_indices returns that index (the float64 subtraction of two nearby numbers is exact, so the sign test
// is the reference's comparison).  Returns false when a tap breaks the pattern (a rounding accident at a grid line).
template <int S>
__device__ __forceinline__ bool taps_consecutive(double cc, double half, double scale, double off, int len, int& k0,
                                                 double (&w)[S], unsigned& err) {
  const double c0 = cc + (0.0 - half) * scale;
  const double cl = cc + ((double)(S - 1) - half) * scale;
  if (c0 < -off || c0 > off || cl < -off || cl > off) err |= APG_FLAG_OUT_OF_GRID;
  tap_row(c0, off, len, true, k0, w[0]);
  const double g0 = (double)k0 - off;  // grid value of index k0 (exact: small integers and halves)
  bool ok = true;
#pragma unroll
  for (int t = 1; t < S; ++t) {
    const double coord = cc + ((double)t - half) * scale;
    w[t] = coord - (g0 + (double)t);
    ok = ok && (w[t] >= 0.0) && (w[t] < 1.0) && (k0 + t <= len - 1);
  }
  return ok;
}

// General (rare) path of one environment: per-point taps, pixels straight from the bank.
template <int S0, int S1>
__device__ __noinline__ void blend_general(const unsigned char* img, int H, int W, int pitch, double ccx, double ccy,
                                           double half0, double half1, double scale, double offy, double offx, float* out,
                                           unsigned* vis_rows /* thread-private, stride kSmallThreads */) {
  unsigned colbits = 0u;
  int kc[S1];
  double wc[S1];
#pragma unroll
  for (int j = 0; j < S1; ++j) {
    tap(ccx + ((double)j - half1) * scale, offx, W, kc[j], wc[j]);
    colbits |= (1u << kc[j]) | (1u << (kc[j] + 1));
  }
#pragma unroll 1
  for (int i = 0; i < S0; ++i) {
    int kr;
    double wr;
    tap(ccy + ((double)i - half0) * scale, offy, H, kr, wr);
    if (vis_rows) {
      vis_rows[kr * kSmallThreads] |= colbits;
      vis_rows[(kr + 1) * kSmallThreads] |= colbits;
    }
    const unsigned char* a = img + (size_t)kr * pitch;
    const unsigned char* b = a + pitch;
#pragma unroll
    for (int j = 0; j < S1; ++j) {
      const double p00 = px_to_f64(a[kc[j]]), p01 = px_to_f64(a[kc[j] + 1]);
      const double p10 = px_to_f64(b[kc[j]]), p11 = px_to_f64(b[kc[j] + 1]);
      const double wy = wr, omy = 1.0 - wr, wx = wc[j], omx = 1.0 - wc[j];
      double acc = p00 * (omy * omx);
      acc = acc + p01 * (omy * wx);
      acc = acc + p10 * (wy * omx);
      acc = acc + p11 * (wy * wx);
      out[i * S1 + j] = (float)acc;
    }
  }
}

template <int S0, int S1>
__global__ void __launch_bounds__(kSmallThreads, 7) image_rollout_kernel(const ImgParams p, const SmallRollout ro) {
  constexpr int P = S0 * S1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_out = reinterpret_cast<float*>(smem_raw);                           // (warps, 32, P) staging
  unsigned* s_vis = reinterpret_cast<unsigned*>(smem_raw + (kSmallThreads / 32) * 32 * P * 4);  // (H, kSmallThreads)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x * kSmallThreads + tid;
  const bool live = n < p.N;
  const int warp_env0 = blockIdx.x * kSmallThreads + warp * 32;
  float* my_out = s_out + (warp * 32 + lane) * P;
  unsigned* my_vis = s_vis + tid;
  const unsigned char* bank = static_cast<const unsigned char*>(p.bank);
  const unsigned wmask = (p.W >= 32) ? 0xffffffffu : ((1u << p.W) - 1u);

  grid_dependency_wait(p.pdl);

  // ---- state -> registers
  double px = 0.0, py = 0.0;
  int idx = 0, label = 0, fc = -1, lic = -1;
  float acc1 = 0.0f, acc2 = 0.0f, accr = 0.0f;
  unsigned char inv = 0;
  unsigned err = 0;
  if (live && ro.first_mode != 2) {
    const double2 pp = reinterpret_cast<const double2*>(p.pos)[n];
    px = pp.x; py = pp.y;
    idx = p.st_idx[n]; label = p.st_label[n];
    acc1 = p.acc1[n]; acc2 = p.acc2[n]; accr = p.acc_ret[n];
    fc = p.first_correct[n]; lic = p.last_incorrect[n];
    if (p.invert) inv = p.st_inv[n];
    if (p.track)
      for (int r = 0; r < p.H; ++r) my_vis[r * kSmallThreads] = p.visited[(size_t)n * p.H + r];
  }
  double st1 = 0.0, st2 = 0.0, st3 = 0.0, st4 = 0.0, st5 = 0.0, st0 = 0.0;  // this thread's share of the shard statistics

  int t = ro.t0;
  bool need_reset = ro.need_reset0 != 0;
#pragma unroll 1
  for (int k = 0; k < ro.n_steps; ++k) {
    const int mode = (k == 0 && ro.first_mode == 2) ? 2 : (need_reset ? 1 : 0);
    const int t_stamp = (mode == 0) ? t : 0;
    const int t_after = (mode == 0) ? t + 1 : 0;
    const bool terminated = (mode == 0) && (t_after >= p.T);
    const float time_obs = (float)(((double)t_after / (double)p.T) * 2.0 - 1.0);
    const size_t koff = (size_t)k * p.N;
    const float* x = p.pred + (koff + n) * p.K;
    float* gout = p.glimpse + (size_t)warp_env0 * P;

    if (live) {
      // ---- inputs of the step, requested back to back
      float a0 = 0.0f, a1 = 0.0f;
      int slot = 0;
      if (mode == 0) {
        const float2 aa = reinterpret_cast<const float2*>(p.action)[koff + n];
        a0 = aa.x; a1 = aa.y;
      } else {
        slot = (mode == 2) ? 0 : ro.reset_slot[k];
        const long long ni = ro.slot_idx[slot][n];
        const double2 pp = reinterpret_cast<const double2*>(ro.slot_pos[slot])[n];
        idx = (int)ni;
        p.index[n] = ni;
        if (p.invert) inv = ro.slot_inv[slot][n];
        // ---- overlay / visited state of the new episode
        px = pp.x; py = pp.y;
      }
      if (k + 1 < ro.n_steps) {  // next step's inputs -> L2 while this step computes (one 128-byte line per lane)
        const char* nx = reinterpret_cast<const char*>(p.pred + ((koff + p.N) + warp_env0) * p.K);
        const int bytes = 32 * p.K * 4;
        if (warp_env0 + 32 <= p.N && lane * 128 < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + lane * 128));
        if (warp_env0 + 32 <= p.N && lane < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.action + ((koff + p.N) + warp_env0) * 2) + lane * 128));
      }

      // ---- move (util.py:94-97, image_perception_module.py:206-209), overlay stamp of the PRE-move position
      float base = 0.0f;
      if (mode == 0) {
        if (isnan(a0) || isnan(a1)) err |= APG_FLAG_NAN_ACTION;
        if (p.track) {
          p.stamp_xy[(size_t)(t_stamp * 2 + 0) * p.N + n] = (int)rint((px * p.lim_x) * p.rs + p.half_wr);
          p.stamp_xy[(size_t)(t_stamp * 2 + 1) * p.N + n] = (int)rint((py * p.lim_y) * p.rs + p.half_hr);
        }
        const float m = norm2f(a0, a1);
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
      reinterpret_cast<float2*>(p.glimpse_pos)[n] = make_float2((float)px, (float)py);

      // ---- taps, then the footprint words are requested: they land while the loss is computed
      const double ccx = px * p.lim_x, ccy = py * p.lim_y;
      int rmin, cmin;
      double wr[S0], wc[S1];
      bool consec = taps_consecutive<S0>(ccy, p.half0, p.scale, p.offy, p.H, rmin, wr, err);
      consec = taps_consecutive<S1>(ccx, p.half1, p.scale, p.offx, p.W, cmin, wc, err) && consec;
      const unsigned char* img = bank + (size_t)idx * p.H * p.bank_pitch;
      unsigned fw[S0 + 1][3];
      if (consec) {
        const unsigned* src = reinterpret_cast<const unsigned*>(img + (size_t)rmin * p.bank_pitch + (cmin & ~3));
        const int pitch_w = p.bank_pitch >> 2;
#pragma unroll
        for (int r = 0; r <= S0; ++r) {
          fw[r][0] = __ldg(src + r * pitch_w);
          fw[r][1] = __ldg(src + r * pitch_w + 1);
          fw[r][2] = __ldg(src + r * pitch_w + 2);
        }
      }

      // ---- label, cross-entropy, reward, flags (image_classification.py:113-127, loss_fn.py:100-110,212-220)
      if (mode != 0) {
        label = p.bank_labels[idx];
        if (p.invert && inv) label = p.K - label - 1;
      }
      if (p.invert) p.inverted_label[n] = (mode == 0) ? 2 : (int)inv;
      p.target_label[n] = label;
      float pq = 0.0f, loss_f = 0.0f, m1 = 0.0f, m2 = 0.0f;
      if (mode != 2) {
        float mx = x[0];
        bool has_nan = isnan(mx);
        for (int c = 1; c < p.K; ++c) {
          const float v = x[c];
          has_nan |= isnan(v);
          mx = fmaxf(mx, v);
        }
        if (has_nan) err |= APG_FLAG_NAN_PRED;
        if (!isfinite(mx)) mx = 0.0f;
        float s = 0.0f;
        for (int c = 0; c < p.K; ++c) s = __fadd_rn(s, expf(__fsub_rn(x[c], mx)));
        const float tl = __fsub_rn(x[label], mx);
        const float ce = -__fsub_rn(tl, logf(s));  // -log_softmax(x)[label], float32 like scipy
        pq = __fdiv_rn(expf(tl), s);               // softmax(x)[label], overlay colour only
        const double loss_d = (double)ce * p.ce_scale + p.ce_offset;  // np.float64 scale -> float64 loss
        loss_f = (float)loss_d;
        m1 = pq;                            // correct_label_prob (active_classification_env.py:157-161)
        m2 = (pq > p.inv_k) ? 1.0f : 0.0f;  // accuracy (:166-168)
        const float reward_f = (float)((double)base + (-loss_d));
        float reward_out = reward_f;
        if (p.sparse) reward_out = (float)((double)base - loss_d * (terminated ? 1.0 : 0.0));
        p.reward[n] = reward_out;
        p.base_reward[n] = base;
        p.loss[n] = loss_f;
        if (p.reward2) p.reward2[koff + n] = reward_out;
        if (p.done2) p.done2[koff + n] = terminated ? 1 : 0;
        if (mode == 0) {  // overlay colour + episode statistics: real steps only
          if (p.track) p.stamp_q[(size_t)t_stamp * p.N + n] = fminf(fmaxf(pq, 0.0f), 1.0f);
          if (t_stamp == 0) { acc1 = 0.0f; acc2 = 0.0f; accr = 0.0f; fc = -1; lic = -1; }
          acc1 = __fadd_rn(acc1, m1);
          acc2 = __fadd_rn(acc2, m2);
          accr = __fadd_rn(accr, reward_f);
          if (m2 > 0.5f) { if (fc < 0) fc = t_stamp; } else lic = t_stamp;
          if (terminated) {
            const float nt = (float)(t_stamp + 1);
            const float avg1 = __fdiv_rn(acc1, nt), avg2 = __fdiv_rn(acc2, nt);
            p.ep_avg1[n] = avg1; p.ep_final1[n] = m1; p.ep_avg2[n] = avg2; p.ep_final2[n] = m2; p.ep_return[n] = accr;
            p.first_correct[n] = fc; p.last_incorrect[n] = lic;
            st0 += 1.0; st1 += (double)avg1; st2 += (double)m1; st3 += (double)avg2; st4 += (double)m2; st5 += (double)accr;
          }
        }
      }
      p.terminated_out[n] = terminated ? 1 : 0;
      p.truncated_out[n] = 0;
      p.time_step[n] = time_obs;

      // ---- visited bitmask (thread-private shared-memory words) + blend
      if (p.track && mode != 0)
        for (int r = 0; r < p.H; ++r) my_vis[r * kSmallThreads] = 0u;
      if (consec) {
        if (p.track) {
          const unsigned colbits = (((1u << (S1 + 1)) - 1u) << cmin) & wmask;
#pragma unroll
          for (int r = 0; r <= S0; ++r)
            if (rmin + r < p.H) my_vis[(rmin + r) * kSmallThreads] |= colbits;
        }
        // every footprint pixel is converted once; row r + 1 of the window is the bottom row of sensor row r
        const int sh = (cmin & 3) * 8;
        double top[S1 + 1], bot[S1 + 1];
        {
          const unsigned x0 = __funnelshift_r(fw[0][0], fw[0][1], sh), x1 = __funnelshift_r(fw[0][1], fw[0][2], sh);
#pragma unroll
          for (int j = 0; j <= S1; ++j) top[j] = px_to_f64((unsigned char)(((j < 4) ? (x0 >> (8 * j)) : (x1 >> (8 * (j - 4)))) & 0xffu));
        }
#pragma unroll
        for (int i = 0; i < S0; ++i) {
          const unsigned x0 = __funnelshift_r(fw[i + 1][0], fw[i + 1][1], sh), x1 = __funnelshift_r(fw[i + 1][1], fw[i + 1][2], sh);
#pragma unroll
          for (int j = 0; j <= S1; ++j) bot[j] = px_to_f64((unsigned char)(((j < 4) ? (x0 >> (8 * j)) : (x1 >> (8 * (j - 4)))) & 0xffu));
          const double wy = wr[i], omy = 1.0 - wr[i];
#pragma unroll
          for (int j = 0; j < S1; ++j) {
            const double wx = wc[j], omx = 1.0 - wc[j];
            double acc = top[j] * (omy * omx);  // the reference's leading `0.0 +` is an exact no-op for products >= +0
            acc = acc + top[j + 1] * (omy * wx);
            acc = acc + bot[j] * (wy * omx);
            acc = acc + bot[j + 1] * (wy * wx);
            my_out[i * S1 + j] = (float)acc;  // uint8 pixels in [0,1], weights sum to 1: already clipped
          }
#pragma unroll
          for (int j = 0; j <= S1; ++j) top[j] = bot[j];
        }
      } else {
        blend_general<S0, S1>(img, p.H, p.W, p.bank_pitch, ccx, ccy, p.half0, p.half1, p.scale, p.offy, p.offx, my_out,
                              p.track ? my_vis : nullptr);
      }
    }
    // ---- the warp's 32 glimpses leave as one contiguous run of 128-bit stores
    __syncwarp();
    {
      const int n_env = min(32, p.N - warp_env0);
      if (n_env > 0) {
        const int total = n_env * P;  // floats; the run starts 16-byte aligned (warp_env0 is a multiple of 32)
        const float4* s4 = reinterpret_cast<const float4*>(s_out + warp * 32 * P);
        float4* g4 = reinterpret_cast<float4*>(gout);
        const int n4 = total >> 2;
        for (int q = lane; q < n4; q += 32) g4[q] = s4[q];
        for (int q = (n4 << 2) + lane; q < total; q += 32) gout[q] = s_out[warp * 32 * P + q];
      }
    }
    __syncwarp();
    t = t_after;
    need_reset = terminated;
  }

  // ---- registers -> state
  if (live) {
    reinterpret_cast<double2*>(p.pos)[n] = make_double2(px, py);
    p.st_idx[n] = idx; p.st_label[n] = label;
    p.acc1[n] = acc1; p.acc2[n] = acc2; p.acc_ret[n] = accr;
    p.first_correct[n] = fc; p.last_incorrect[n] = lic;
    if (p.invert) p.st_inv[n] = inv;
    if (p.track)
      for (int r = 0; r < p.H; ++r) p.visited[(size_t)n * p.H + r] = my_vis[r * kSmallThreads];
    if (err) atomicOr(p.err, err);
  }
  // shard-wide episode statistics: warp sums, one atomic per warp and statistic
  if (__any_sync(0xffffffffu, st0 != 0.0)) {
    double v[6] = {st0, st1, st2, st3, st4, st5};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
      if (lane == 0) atomicAdd(p.stats + i, v[i]);
    }
  }
}

}  // namespace

bool small_plan(SmallPlan& plan, const apg_image_cfg& cfg, int words) {
  plan = SmallPlan{};
  if (cfg.bank_dtype != APG_BANK_U8 || cfg.sensor_scale != 1.0 || cfg.blend_f32) return false;
  if (cfg.task != APG_TASK_CLASSIFICATION) return false;
  if (!(cfg.sensor_h == 5 && cfg.sensor_w == 5 && cfg.channels == 1)) return false;
  if (words != 1 || cfg.height > 32 || cfg.height < 6 || cfg.width < 6) return false;
  plan.spec = 1;
  plan.threads = kSmallThreads;
  plan.smem = (size_t)(kSmallThreads / 32) * 32 * 25 * 4 + (size_t)cfg.height * kSmallThreads * 4;
  return true;
}

cudaError_t small_preload(const SmallPlan& plan) {
  if (!plan.spec) return cudaSuccess;
  cudaFuncAttributes a;
  return cudaFuncGetAttributes(&a, image_rollout_kernel<5, 5>);
}

int small_launch(const SmallPlan& plan, const ImgParams& p, const SmallRollout& ro, cudaStream_t stream) {
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg{};
  cfg.stream = stream;
  cfg.attrs = at;
  cfg.numAttrs = p.pdl ? 1 : 0;
  cfg.gridDim = dim3((unsigned)((p.N + kSmallThreads - 1) / kSmallThreads));
  cfg.blockDim = dim3(kSmallThreads);
  cfg.dynamicSmemBytes = plan.smem;
  cudaError_t e = cudaLaunchKernelEx(&cfg, image_rollout_kernel<5, 5>, p, ro);
  if (e != cudaSuccess) return -(int)e;
  return 1;
}

}  // namespace apg_img
