// Persistent thread-per-environment rollout kernel for small glimpses on uint8 single-channel banks (internal, not
// part of the C-ABI).  See apg_image_small.cu.
#pragma once

#include <cuda_runtime.h>

#include "apg_image_dev.cuh"

namespace apg_img {

struct SmallPlan {
  int spec = 0;  // 1: 5x5x1 sensor, classification (the MNIST shape); 0 = not covered by this path
  int threads = 0;
  size_t smem = 0;
};

typedef RolloutPlan SmallRollout;  // see apg_image_dev.cuh
constexpr int kSmallMaxSteps = kRolloutMaxSteps;
constexpr int kSmallMaxSlots = kRolloutMaxSlots;

bool small_plan(SmallPlan& plan, const apg_image_cfg& cfg, int words);
cudaError_t small_preload(const SmallPlan& plan);
// p.action / p.pred / p.reward2 / p.done2 point at the FIRST step of the window (stacked (n_steps, N, ...) arrays);
// returns the number of launches or a negative cudaError
int small_launch(const SmallPlan& plan, const ImgParams& p, const SmallRollout& ro, cudaStream_t stream);

}  // namespace apg_img
