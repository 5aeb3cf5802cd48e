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

#include <cuda_pipeline.h>

#include <cstdint>

namespace apg_img {
namespace {

#ifndef APG_SMALL_COOP
#define APG_SMALL_COOP 1  // 1: footprints copied warp-cooperatively through shared memory, 0: per-thread gathers (__ldg)
#endif
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
                                           unsigned* vis_rows /* thread-private words, one per image row */) {
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
      vis_rows[kr] |= colbits;
      vis_rows[kr + 1] |= colbits;
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

// asynchronous warp-cooperative copy global -> shared (cp.async, SASS LDGSTS): 16-byte pieces where both sides are
// aligned, 4-byte pieces otherwise / for the tail.  `bytes` is a multiple of 4.
__device__ __forceinline__ void warp_copy_async(unsigned char* dst, const unsigned char* src, int bytes, int lane) {
  int done = 0;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int n16 = bytes >> 4;
    for (int q = lane; q < n16; q += 32) __pipeline_memcpy_async(dst + q * 16, src + q * 16, 16);
    done = n16 << 4;
  }
  for (int q = done + lane * 4; q < bytes; q += 128) __pipeline_memcpy_async(dst + q, src + q, 4);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// 16-byte cp.async (SASS LDGSTS.128) on a shared-window address; joins the thread's current cp.async group like
// __pipeline_memcpy_async
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

struct SmallSmem {  // per-block layout of the dynamic shared memory
  int fp_stride;    // bytes per environment footprint slot: (S0+1) rows of the bank pitch + 16 (bank-conflict padding)
  int fp_bytes;     // per warp: 32 slots; the output staging (32 x P floats) aliases this area
  int lg_bytes;     // per warp and stage: 32 x K floats
  int warp_bytes;   // fp_bytes + 2 * lg_bytes
  int vis_stride;   // words per environment of the visited mask: H | 1 (odd: conflict-free thread-private access)
};

__host__ __device__ inline SmallSmem small_smem(int S0, int P, int K, int H, int pitch) {
  SmallSmem m;
  m.fp_stride = (S0 + 1) * pitch + 16;
  m.fp_bytes = 32 * m.fp_stride;
  if (m.fp_bytes < 32 * P * 4) m.fp_bytes = 32 * P * 4;
  m.lg_bytes = (32 * K * 4 + 15) & ~15;
  m.warp_bytes = m.fp_bytes + 2 * m.lg_bytes;
  m.vis_stride = H | 1;
  return m;
}

// PITCH = bytes per bank row (a multiple of 16); KC = number of classes when known at compile time (the loops over the
// logits unroll and their exp chains overlap), 0 = p.K
// FULL: every warp of the launch owns 32 live environments ((n1 - n0) % 64 == 0) -- no per-thread `live` predicates, and
// the three warp-cooperative copies of a step (footprints in, next logits in, glimpses out) are unrolled with per-lane
// constant offsets instead of index arithmetic per piece.
template <int S0, int S1, int PITCH, int KC, bool FULL>
__global__ void __launch_bounds__(kSmallThreads, 7) image_rollout_kernel(const ImgParams p, const SmallRollout ro) {
  constexpr int P = S0 * S1;
  const int K = KC ? KC : p.K;
  constexpr int kWarpsPerBlock = kSmallThreads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SmallSmem sm = small_smem(S0, P, K, p.H, PITCH);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n0 + blockIdx.x * kSmallThreads + tid;  // [n0, n1): the environments this launch steps (n0 % 32 == 0)
  const bool live = FULL ? true : (n < p.n1);
  const int warp_env0 = p.n0 + blockIdx.x * kSmallThreads + warp * 32;
  const int n_env = FULL ? 32 : max(0, min(32, p.n1 - warp_env0));  // environments of this warp
  unsigned char* wbase = smem_raw + warp * sm.warp_bytes;
  unsigned char* s_fp = wbase;                                   // footprint slots of the warp's environments
  float* s_out = reinterpret_cast<float*>(wbase);                // ... reused as the output staging (32 x P floats)
  float* s_lg = reinterpret_cast<float*>(wbase + sm.fp_bytes);   // two stages of the warp's logits
  unsigned* s_vis = reinterpret_cast<unsigned*>(smem_raw + kWarpsPerBlock * sm.warp_bytes);
  float* my_out = s_out + lane * P;
  unsigned* my_vis = s_vis + tid * sm.vis_stride;
  const unsigned char* bank = static_cast<const unsigned char*>(p.bank);
  const unsigned wmask = (p.W >= 32) ? 0xffffffffu : ((1u << p.W) - 1u);
  const int lg_words = sm.lg_bytes >> 2;
  constexpr int ch_per_env = (S0 + 1) * (PITCH >> 4);            // 16-byte pieces of one footprint (divisions by it fold)
  constexpr int kFpStride = (S0 + 1) * PITCH + 16;               // == sm.fp_stride
  // FULL copies: lane = (environment of a batch of 8, piece of a group of 4); an environment's ch_per_env pieces are
  // contiguous in the bank ((S0+1) whole rows) and in its slot, so one source / destination address per batch serves
  // its ch_per_env / 4 groups with immediate offsets
  static_assert(!FULL || ch_per_env % 4 == 0, "groups of 4 pieces");
  const unsigned fp_dst_u = smem_u32(s_fp) + (lane >> 2) * kFpStride + (lane & 3) * 16;
  const bool lg_fast = FULL && KC > 0 && ((reinterpret_cast<uintptr_t>(p.pred) | ((size_t)p.N * KC * 4)) & 15) == 0;

  grid_dependency_wait(p.pdl);

  // ---- state -> registers; the visited masks of the warp's environments are one contiguous run
  double px = 0.0, py = 0.0;
  int idx = 0, label = 0, fc = -1, lic = -1;
  float acc1 = 0.0f, acc2 = 0.0f, accr = 0.0f;
  unsigned char inv = 0;
  unsigned err = 0;
  if (ro.first_mode != 2) {
    if (live) {
      const double2 pp = reinterpret_cast<const double2*>(p.pos)[n];
      px = pp.x; py = pp.y;
      idx = p.st_idx[n]; label = p.st_label[n];
      acc1 = p.acc1[n]; acc2 = p.acc2[n]; accr = p.acc_ret[n];
      fc = p.first_correct[n]; lic = p.last_incorrect[n];
      if (p.invert) inv = p.st_inv[n];
    }
    if (p.track) {
      const unsigned* gv = p.visited + (size_t)warp_env0 * p.H;
      unsigned* wv = s_vis + warp * 32 * sm.vis_stride;
      // H <= 32 words per environment: one coalesced run per iteration, all of them in flight at once (cp.async)
      for (int e = 0; e < n_env; ++e)
        if (lane < p.H) __pipeline_memcpy_async(wv + e * sm.vis_stride + lane, gv + e * p.H + lane, 4);
    }
  }
  // logits of the first step on their way (stage 0)
  const unsigned char* pred_w = reinterpret_cast<const unsigned char*>(p.pred + (size_t)warp_env0 * K);
  const size_t pred_step = (size_t)p.N * K * 4;
  // the warp's logits of one step: 32 * K floats, contiguous
  auto copy_logits = [&](float* dst, const unsigned char* src) {
    if (lg_fast) {
      constexpr int n16 = (32 * KC * 4) / 16;
      const unsigned d = smem_u32(dst) + lane * 16;
#pragma unroll
      for (int q = 0; q < n16; q += 32)
        if (q + 32 <= n16 || q + lane < n16) cp_async16(d + q * 16, src + (q + lane) * 16);
    } else if (n_env > 0) {
      warp_copy_async(reinterpret_cast<unsigned char*>(dst), src, n_env * K * 4, lane);
    }
  };
  if (ro.first_mode != 2) copy_logits(s_lg, pred_w);
  __pipeline_commit();
  __syncwarp();
  double st1 = 0.0, st2 = 0.0, st3 = 0.0, st4 = 0.0, st5 = 0.0, st0 = 0.0;  // this thread's share of the shard statistics

  int t = ro.t0;
  bool need_reset = ro.need_reset0 != 0;
  const float2* act = reinterpret_cast<const float2*>(p.action) + n;  // running pointers: + N per step
  // the action of step k + 1 is requested during step k (registers): its latency is off the step's critical path
  float2 act_next = make_float2(0.0f, 0.0f);
  const bool have_act = live && p.action != nullptr;  // (an explicit reset launches without actions)
  if (have_act && ro.n_steps > 0) act_next = __ldcs(act);
  const unsigned char* pred_next = pred_w;
  float* rew2 = p.reward2 ? p.reward2 + n : nullptr;
  unsigned char* don2 = p.done2 ? p.done2 + n : nullptr;
  float* gout = p.glimpse + (size_t)warp_env0 * P;
#pragma unroll 1
  for (int k = 0; k < ro.n_steps; ++k) {
    const int mode = (k == 0 && ro.first_mode == 2) ? 2 : (need_reset ? 1 : 0);
    const int t_stamp = (mode == 0) ? t : 0;
    const int t_after = (mode == 0) ? t + 1 : 0;
    const bool terminated = (mode == 0) && (t_after >= p.T);
    const float time_obs = ro.time_obs[k];
    const float* x = s_lg + (k & 1) * lg_words + lane * K;

    // ---- inputs of the step
    float a0 = 0.0f, a1 = 0.0f;
    const float2 aa = act_next;
    act += p.N;
    if (have_act && k + 1 < ro.n_steps) act_next = __ldcs(act);
    if (live) {
      if (mode == 0) {
        a0 = aa.x; a1 = aa.y;
      } else {
        const int slot = (mode == 2) ? 0 : ro.reset_slot[k];
        const long long ni = ro.slot_idx[slot][n];
        const double2 pp = reinterpret_cast<const double2*>(ro.slot_pos[slot])[n];
        idx = (int)ni;
        p.index[n] = ni;
        if (p.invert) inv = ro.slot_inv[slot][n];
        px = pp.x; py = pp.y;
      }
    }
    // ---- move (util.py:94-97, image_perception_module.py:206-209), overlay stamp of the PRE-move position
    float base = 0.0f;
    if (live && mode == 0) {
      if (isnan(a0) || isnan(a1)) err |= APG_FLAG_NAN_ACTION;
      if (p.track) {
        int* sxy = p.stamp_xy + ((size_t)(t_stamp * 2) * p.N + n);
        sxy[0] = (int)rint((px * p.lim_x) * p.rs + p.half_wr);
        sxy[p.N] = (int)rint((py * p.lim_y) * p.rs + p.half_hr);
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
    if (live) reinterpret_cast<float2*>(p.glimpse_pos)[n] = make_float2((float)px, (float)py);

    // ---- taps; then the warp copies its 32 footprints cooperatively (consecutive lanes -> consecutive 16-byte pieces of
    // one environment's (S0+1) contiguous bank rows): every request covers whole lines instead of 32 scattered sectors
    const double ccx = px * p.lim_x, ccy = py * p.lim_y;
    int rmin = 0, cmin = 0;
    double wr[S0], wc[S1];
    bool consec = true;
    if (live) {
      consec = taps_consecutive<S0>(ccy, p.half0, p.scale, p.offy, p.H, rmin, wr, err);
      consec = taps_consecutive<S1>(ccx, p.half1, p.scale, p.offx, p.W, cmin, wc, err) && consec;
    }
    unsigned fw[S0 + 1][3];
#if APG_SMALL_COOP
    {
      const int my_row = idx * p.H + rmin;  // first bank row of this thread's footprint
      if (FULL) {
#pragma unroll
        for (int eb = 0; eb < 4; ++eb) {
          const int row_e = __shfl_sync(0xffffffffu, my_row, eb * 8 + (lane >> 2));
          const unsigned char* src = bank + (size_t)row_e * PITCH + (lane & 3) * 16;
#pragma unroll
          for (int g = 0; g < ch_per_env / 4; ++g) cp_async16(fp_dst_u + eb * 8 * kFpStride + g * 64, src + g * 64);
        }
      }
      const int total = FULL ? 0 : n_env * ch_per_env;
      for (int q = lane; q < ((total + 31) & ~31); q += 32) {
        const int e = q / ch_per_env, c = q - e * ch_per_env;
        const int row_e = __shfl_sync(0xffffffffu, my_row, e & 31);
        if (q < total)
          __pipeline_memcpy_async(s_fp + e * sm.fp_stride + c * 16, bank + (size_t)row_e * PITCH + c * 16, 16);
      }
      __pipeline_commit();
    }
#else
    if (live && consec) {  // three aligned words per footprint row straight into registers
      const unsigned* src = reinterpret_cast<const unsigned*>(bank + ((size_t)idx * p.H + rmin) * PITCH + (cmin & ~3));
      constexpr int pitch_w = PITCH >> 2;
#pragma unroll
      for (int r = 0; r <= S0; ++r) {
        fw[r][0] = __ldg(src + r * pitch_w);
        fw[r][1] = __ldg(src + r * pitch_w + 1);
        fw[r][2] = __ldg(src + r * pitch_w + 2);
      }
    }
    __pipeline_commit();  // (keeps the group count of the cooperative variant)
#endif
    // ---- next step's logits -> the other stage (they have the rest of the step to arrive); committed AFTER the
    // footprints so that waiting for the footprints does not wait for them
    pred_next += pred_step;
    if (k + 1 < ro.n_steps) copy_logits(s_lg + ((k + 1) & 1) * lg_words, pred_next);
    __pipeline_commit();

    // ---- label, cross-entropy, reward, flags (image_classification.py:113-127, loss_fn.py:100-110,212-220); the
    // logits of this step were requested one step ago
    __pipeline_wait_prior(2);  // everything but the two newest groups (footprints, next logits) has landed
    __syncwarp();
    if (live) {
      if (mode != 0) {
        label = p.bank_labels[idx];
        if (p.invert && inv) label = K - label - 1;
      }
      if (p.invert) p.inverted_label[n] = (mode == 0) ? 2 : (int)inv;
      p.target_label[n] = label;
      if (mode != 2) {
        float mx = x[0];
        bool has_nan = isnan(mx);
#pragma unroll
        for (int c = 1; c < K; ++c) {
          const float v = x[c];
          has_nan |= isnan(v);
          mx = fmaxf(mx, v);
        }
        if (has_nan) err |= APG_FLAG_NAN_PRED;
        if (!isfinite(mx)) mx = 0.0f;
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < K; ++c) s = __fadd_rn(s, expf(__fsub_rn(x[c], mx)));
        const float tl = __fsub_rn(x[label], mx);
        const float ce = -__fsub_rn(tl, logf(s));  // -log_softmax(x)[label], float32 like scipy
        const float pq = __fdiv_rn(expf(tl), s);   // softmax(x)[label], overlay colour only
        const double loss_d = (double)ce * p.ce_scale + p.ce_offset;  // np.float64 scale -> float64 loss
        const float m1 = pq;                            // correct_label_prob (active_classification_env.py:157-161)
        const float m2 = (pq > p.inv_k) ? 1.0f : 0.0f;  // accuracy (:166-168)
        const float reward_f = (float)((double)base + (-loss_d));
        float reward_out = reward_f;
        if (p.sparse) reward_out = (float)((double)base - loss_d * (terminated ? 1.0 : 0.0));
        p.reward[n] = reward_out;
        p.base_reward[n] = base;
        p.loss[n] = (float)loss_d;
        if (rew2) *rew2 = reward_out;
        if (don2) *don2 = terminated ? 1 : 0;
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
      // visited bitmask (thread-private shared-memory words)
      if (p.track && mode != 0)
        for (int r = 0; r < p.H; ++r) my_vis[r] = 0u;
      if (p.track && consec) {
        const unsigned colbits = (((1u << (S1 + 1)) - 1u) << cmin) & wmask;
#pragma unroll
        for (int r = 0; r <= S0; ++r)
          if (rmin + r < p.H) my_vis[rmin + r] |= colbits;
      }
    }

    // ---- the footprints have landed: 3 aligned words per row -> registers, then the slots are free (the output staging
    // aliases them)
#if APG_SMALL_COOP
    __pipeline_wait_prior(1);
    __syncwarp();
    if (live && consec) {
      const unsigned* src = reinterpret_cast<const unsigned*>(s_fp + lane * sm.fp_stride + (cmin & ~3));
      constexpr int pitch_w = PITCH >> 2;
#pragma unroll
      for (int r = 0; r <= S0; ++r) {
        fw[r][0] = src[r * pitch_w];
        fw[r][1] = src[r * pitch_w + 1];
        fw[r][2] = src[r * pitch_w + 2];
      }
    }
    __syncwarp();
#endif
    if (live) {
      if (consec) {
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
        blend_general<S0, S1>(bank + (size_t)idx * p.H * PITCH, p.H, p.W, PITCH, ccx, ccy, p.half0, p.half1,
                              p.scale, p.offy, p.offx, my_out, p.track ? my_vis : nullptr);
      }
    }
    // ---- the warp's 32 glimpses leave as one contiguous run of 128-bit stores
    __syncwarp();
    if (FULL) {
      constexpr int n4 = (32 * P) / 4;
      static_assert((32 * P) % 4 == 0, "whole float4s");
      const float4* s4 = reinterpret_cast<const float4*>(s_out) + lane;
      float4* g4 = reinterpret_cast<float4*>(gout) + lane;
#pragma unroll
      for (int q = 0; q < n4; q += 32)
        if (q + 32 <= n4 || q + lane < n4) g4[q] = s4[q];
    } else if (n_env > 0) {
      const int total = n_env * P;  // floats; the run starts 16-byte aligned (warp_env0 is a multiple of 32)
      const float4* s4 = reinterpret_cast<const float4*>(s_out);
      float4* g4 = reinterpret_cast<float4*>(gout);
      const int n4 = total >> 2;
      for (int q = lane; q < n4; q += 32) g4[q] = s4[q];
      for (int q = (n4 << 2) + lane; q < total; q += 32) gout[q] = s_out[q];
    }
    __syncwarp();
    t = t_after;
    need_reset = terminated;
    if (rew2) rew2 += p.N;
    if (don2) don2 += p.N;
  }
  __pipeline_wait_prior(0);

  // ---- registers -> state
  if (live) {
    reinterpret_cast<double2*>(p.pos)[n] = make_double2(px, py);
    p.st_idx[n] = idx; p.st_label[n] = label;
    p.acc1[n] = acc1; p.acc2[n] = acc2; p.acc_ret[n] = accr;
    p.first_correct[n] = fc; p.last_incorrect[n] = lic;
    if (p.invert) p.st_inv[n] = inv;
    if (err) atomicOr(p.err, err);
  }
  if (p.track) {
    unsigned* gv = p.visited + (size_t)warp_env0 * p.H;
    const unsigned* wv = s_vis + warp * 32 * sm.vis_stride;
    for (int e = 0; e < n_env; ++e)
      if (lane < p.H) gv[e * p.H + lane] = wv[e * sm.vis_stride + lane];
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
  if (words != 1 || cfg.height > 32 || cfg.height < 6 || cfg.width < 17) return false;  // bank rows of 32 bytes
  plan.threads = kSmallThreads;
  const int pitch = (cfg.width * cfg.channels + 15) & ~15;
  const SmallSmem sm = small_smem(5, 25, cfg.num_classes, cfg.height, pitch);
  plan.smem = (size_t)(kSmallThreads / 32) * sm.warp_bytes + (size_t)kSmallThreads * sm.vis_stride * 4;
  if (plan.smem > 30 * 1024) return false;  // 7 blocks per SM must fit (one wave at 65536 envs)
  plan.spec = 1;
  return true;
}

cudaError_t small_preload(const SmallPlan& plan) {
  if (!plan.spec) return cudaSuccess;
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, image_rollout_kernel<5, 5, 32, 10, true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, image_rollout_kernel<5, 5, 32, 10, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, image_rollout_kernel<5, 5, 32, 0, true>);
  if (e != cudaSuccess) return e;
  return cudaFuncGetAttributes(&a, image_rollout_kernel<5, 5, 32, 0, false>);
}

int small_launch(const SmallPlan& plan, const ImgParams& p, const SmallRollout& ro, cudaStream_t stream) {
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg{};
  cfg.stream = stream;
  cfg.attrs = at;
  cfg.numAttrs = p.pdl ? 1 : 0;
  cfg.gridDim = dim3((unsigned)((p.n1 - p.n0 + kSmallThreads - 1) / kSmallThreads));
  cfg.blockDim = dim3(kSmallThreads);
  cfg.dynamicSmemBytes = plan.smem;
  const bool full = (p.n1 - p.n0) % kSmallThreads == 0;
  cudaError_t e;
  if (p.K == 10)
    e = full ? cudaLaunchKernelEx(&cfg, image_rollout_kernel<5, 5, 32, 10, true>, p, ro)
             : cudaLaunchKernelEx(&cfg, image_rollout_kernel<5, 5, 32, 10, false>, p, ro);
  else
    e = full ? cudaLaunchKernelEx(&cfg, image_rollout_kernel<5, 5, 32, 0, true>, p, ro)
             : cudaLaunchKernelEx(&cfg, image_rollout_kernel<5, 5, 32, 0, false>, p, ro);
  if (e != cudaSuccess) return -(int)e;
  return 1;
}

}  // namespace apg_img
