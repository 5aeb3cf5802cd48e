// Device-side generation of the image environments' autoreset payloads (internal, not part of the C-ABI).
//
// The reference draws every autoreset payload from three numpy Generators shared by the whole vector env
// (image_localization.py:153-156 -> dataset_iterator.py:57 -> image_perception_module.py:131-133, :159-161):
//     g0.uniform(-1, 1, (N, 2)).astype(float32)      localization targets
//     g2.integers(0, len(dataset), N)                bank indices (Lemire, rare rejections shift the stream)
//     g1.integers(0, 2, N) == 1                      label inversion (optional)
//     g1.uniform(-1, 1, (N, 2))                      float64 start positions
// A PCG64 stream is an LCG and can be jumped to any draw in O(log n), so every value of a payload is drawn by its
// own thread, bit-identical to numpy; the three stream states live in device memory and are advanced by the
// generator itself (no host round trip, the payload is drawn in-stream right before the step that consumes it).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "apg_rng.h"

namespace apg {

constexpr int kPayloadSlack = 1024;  // rejected bank-index draws one payload may hold (see payload_device_ok)

struct PayloadArgs {
  const Pcg64* streams;  // device, [3] = g0, g1, g2 before the payload
  Pcg64* streams_out;    // device, [3]: the states after the payload (a different buffer: written while others read)
  long long Ng, off, N;  // global batch, this shard's first env, this shard's envs
  unsigned long long M;  // len(dataset)
  int loc, inv;
  long long* d_idx;      // (N)   int64
  double* d_pos;         // (N,2) float64
  float* d_target;       // (N,2) float32 (loc)
  unsigned char* d_inv;  // (N)   (inv)
  unsigned* words;       // scratch: this shard's window of the g2 word stream, words [off, off + N + kPayloadSlack)
  unsigned* rej;         // scratch: [0] = number of rejected words, [1 .. kPayloadSlack] their positions; zero on entry
  unsigned* rej_next;    // the other rejection list, cleared here for the next payload
};

// true when the device generator may be used: 2 <= M <= 2^32 and so few expected rejections that kPayloadSlack
// cannot overflow (otherwise the host generator draws the payload)
bool payload_device_ok(unsigned long long M, long long Ng);
size_t payload_scratch_words(long long N);  // N = this shard's envs
// loads the two kernels (module load off the step path)
cudaError_t payload_preload();
// two launches on `stream`; returns the number of launches or a negative cudaError
int launch_device_payload(const PayloadArgs& a, cudaStream_t stream);

}  // namespace apg
