// Persistent thread-per-environment rollout kernel for small glimpses on uint8 single-channel banks (internal, not
// part of the C-ABI).  See apg_image_small.cu.
#pragma once

#include <cuda_runtime.h>

#include "apg_image_dev.cuh"

namespace apg_img {

constexpr int kSmallMaxSteps = 128;  // steps per launch
constexpr int kSmallMaxSlots = 8;    // autoresets per launch

struct SmallPlan {
  int spec = 0;  // 1: 5x5x1 sensor, classification (the MNIST shape); 0 = not covered by this path
  int threads = 0;
  size_t smem = 0;
};

// What one launch steps through.  The episode clock is shared by all environments, so the mode of every step follows
// from (t0, need_reset0): a step is an autoreset step exactly when the previous one terminated.
struct SmallRollout {
  int n_steps;
  int t0;           // episode clock before the first step
  int need_reset0;  // the first step is an autoreset step
  int first_mode;   // 2: the first (only) step is an explicit reset from slot 0
  unsigned char reset_slot[kSmallMaxSteps];  // for autoreset steps: which payload slot below
  const long long* slot_idx[kSmallMaxSlots];
  const double* slot_pos[kSmallMaxSlots];
  const unsigned char* slot_inv[kSmallMaxSlots];
};

bool small_plan(SmallPlan& plan, const apg_image_cfg& cfg, int words);
cudaError_t small_preload(const SmallPlan& plan);
// p.action / p.pred / p.reward2 / p.done2 point at the FIRST step of the window (stacked (n_steps, N, ...) arrays);
// returns the number of launches or a negative cudaError
int small_launch(const SmallPlan& plan, const ImgParams& p, const SmallRollout& ro, cudaStream_t stream);

}  // namespace apg_img
