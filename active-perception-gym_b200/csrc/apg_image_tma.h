// Large-glimpse image step for uint8 banks at sensor_scale == 1: TMA-staged tiles (internal, not part of the C-ABI).
// See apg_image_tma.cu.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "apg_image_dev.cuh"

namespace apg_img {

struct TmaPlan {
  int spec = 0;        // 1: S0 = S1 = 10, C = 3 (the TinyImageNet shape); 0 = geometry not covered by this path
  int grid = 0;        // persistent grid of the step kernel (blocks)
  int warps = 0;       // grid * warps per block
  void* d_table = nullptr;    // uint8 -> pixel value table (interleaved copies, float32 or float64 entries), device
  size_t smem = 0;     // dynamic shared memory of the blend kernel
  int box_w = 0, box_h = 0;
  CUtensorMap tmap[3]; // the bank as a (rows, row bytes) uint8 tensor, box = one glimpse footprint (+ q rows, slot q)
  bool have_map = false;
  void* d_bank_blk = nullptr;   // blocked copy of the bank (8 rows x 16 bytes per 128-byte line), engine-owned
  int bank_rows = 0;            // rows of the bank tensor (items x height): row-copy variant clamps to it
};

// geometry / device check + kernel attributes + persistent grid size; returns false when the shape is not covered
bool tma_plan(TmaPlan& plan, const apg_image_cfg& cfg, int bank_pitch, int words, cudaError_t* err);
// (re-)encodes the tensor map of the bank; false when the driver entry point is unavailable or the bank is too large
bool tma_encode(TmaPlan& plan, const void* d_bank, int bank_pitch, long long n_items, int height);
void tma_release(TmaPlan& plan);
// one launch = the steps of `ro` (p.action / p.pred / p.reward2 / p.done2 point at the first one); returns the number of
// launches or a negative cudaError
int tma_launch(const TmaPlan& plan, const ImgParams& p, const RolloutPlan& ro, cudaStream_t stream);
// loads the kernels (first-use module load off the step path)
cudaError_t tma_preload(const TmaPlan& plan);

}  // namespace apg_img
