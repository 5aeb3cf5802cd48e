// Device-side autoreset payloads of the image environments: see apg_payload.h.
#include "apg_payload.h"

#include <vector>

#include "apg_common.h"

namespace apg {
namespace {

constexpr int kDrawsPerThread = 4;  // consecutive uniform draws per thread: one O(log n) jump, then plain steps

// word t of the 32-bit word stream Generator.integers consumes from state g: the buffered half-word first (if any),
// then the low and the high half of every raw draw
__device__ __forceinline__ uint32_t stream_word(const Pcg64& g0, long long t) {
  if (g0.has_uint32) {
    if (t == 0) return g0.uinteger;
    t -= 1;
  }
  Pcg64 g = g0;
  pcg64_advance(g, (uint64_t)(t >> 1));
  const uint64_t n = pcg64_next64(g);
  return (t & 1) ? (uint32_t)(n >> 32) : (uint32_t)n;
}

// state after W words have been consumed through next32 (pcg64_next32 semantics, including the buffered half)
__device__ __forceinline__ void consume_words(Pcg64& g, long long W) {
  if (W <= 0) return;
  if (g.has_uint32) {
    g.has_uint32 = 0;
    W -= 1;
    if (W == 0) return;
  }
  const long long raws = (W + 1) >> 1;
  pcg64_advance(g, (uint64_t)(raws - 1));
  const uint64_t n = pcg64_next64(g);
  g.uinteger = (uint32_t)(n >> 32);
  g.has_uint32 = (uint32_t)(W & 1);
}

// raw draws g1 spends on integers(0, 2, Ng) before the uniform positions
__device__ __forceinline__ long long inv_raws(const Pcg64& g1, long long Ng) {
  const long long W = Ng - (g1.has_uint32 ? 1 : 0);
  return W <= 0 ? 0 : (W + 1) >> 1;
}

__device__ __forceinline__ bool lemire_rejects(uint32_t word, unsigned long long M, uint32_t threshold) {
  return (uint32_t)((uint64_t)word * M) < threshold;  // (leftover >= M can never be below the threshold < M)
}

// position of the accepted word of rank j: the least fixed point of t = j + #{rejected positions <= t}
__device__ __forceinline__ long long accepted_word(const unsigned* rej, long long j) {
  const int R = min((int)rej[0], kPayloadSlack);
  long long t = j;
  for (;;) {
    int c = 0;
    for (int i = 0; i < R; ++i) c += ((long long)rej[1 + i] <= t) ? 1 : 0;
    const long long t2 = j + c;
    if (t2 == t) return t;
    t = t2;
  }
}

// (a) the g2 word stream of the whole global batch (+ slack) and the positions Lemire's method rejects; every thread
// jumps once and then steps through kRawsPerThread consecutive raw draws (two words each)
constexpr int kRawsPerThread = 8;

__device__ __forceinline__ void note_word(const PayloadArgs& a, long long t, uint32_t w, uint32_t threshold) {
  // only this shard's window [off, off + N + slack) is kept; the rejections of the WHOLE global stream are listed (the
  // accepted word of rank j and the final stream state depend on every rejection before it)
  if (t >= a.off && t < a.off + a.N + kPayloadSlack) a.words[t - a.off] = w;
  if (lemire_rejects(w, a.M, threshold)) {
    const unsigned slot = atomicAdd(&a.rej[0], 1u);
    if (slot < (unsigned)kPayloadSlack) a.rej[1 + slot] = (unsigned)t;
  }
}

__global__ void payload_words_kernel(const PayloadArgs a) {
  const long long n_words = a.Ng + kPayloadSlack;
  const long long r0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kRawsPerThread;  // first raw draw
  const Pcg64 g2 = a.streams[2];
  const int has = g2.has_uint32 ? 1 : 0;
  const uint32_t threshold = (a.M >= 0x100000000ull) ? 0u : (uint32_t)((0x100000000ull - a.M) % a.M);
  if (has + 2 * r0 < n_words) {
    Pcg64 g = g2;
    pcg64_advance(g, (uint64_t)r0);
    for (int i = 0; i < kRawsPerThread; ++i) {
      const long long t0 = has + 2 * (r0 + i);
      if (t0 >= n_words) break;
      const uint64_t n = pcg64_next64(g);
      note_word(a, t0, (uint32_t)n, threshold);
      if (t0 + 1 < n_words) note_word(a, t0 + 1, (uint32_t)(n >> 32), threshold);
    }
  }
  if (r0 == 0 && has) note_word(a, 0, g2.uinteger, threshold);
}

// (c) one thread: the stream states after the payload (into the other state buffer, the value threads are still
// reading this one) and a clean rejection list for the next payload
__device__ __forceinline__ void payload_finish(const PayloadArgs& a) {
  Pcg64 g0 = a.streams[0], g1 = a.streams[1], g2 = a.streams[2];
  if (a.loc) pcg64_advance(g0, (uint64_t)(2 * a.Ng));
  if (a.M > 1) consume_words(g2, accepted_word(a.rej, a.Ng - 1) + 1);
  if (a.inv) consume_words(g1, a.Ng);
  pcg64_advance(g1, (uint64_t)(2 * a.Ng));
  a.streams_out[0] = g0;
  a.streams_out[1] = g1;
  a.streams_out[2] = g2;
  a.rej_next[0] = 0u;
}

// (b) this shard's slice of the payload
__global__ void payload_values_kernel(const PayloadArgs a) {
  if (blockIdx.x == gridDim.x - 1) {  // an extra block: its first thread advances the streams, side by side
    if (threadIdx.x == 0) payload_finish(a);
    return;
  }
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // bank indices
  if (q < a.N) {
    const long long t = accepted_word(a.rej, a.off + q);
    a.d_idx[q] = (a.M == 1) ? 0ll : (long long)(((uint64_t)a.words[t - a.off] * a.M) >> 32);
  }
  // label inversion: integers(0, 2) never rejects, value = top bit of the word
  if (a.inv && q < a.N) a.d_inv[q] = (unsigned char)(stream_word(a.streams[1], a.off + q) >> 31);
  // uniform(-1, 1) draws, kDrawsPerThread consecutive ones per thread
  const long long first = q * kDrawsPerThread;
  if (first < 2 * a.N) {
    const int cnt = (int)min((long long)kDrawsPerThread, 2 * a.N - first);
    {
      Pcg64 g = a.streams[1];
      pcg64_advance(g, (uint64_t)((a.inv ? inv_raws(g, a.Ng) : 0) + 2 * a.off + first));
      for (int i = 0; i < cnt; ++i) a.d_pos[first + i] = rng_uniform(g, -1.0, 2.0);
    }
    if (a.loc) {
      Pcg64 g = a.streams[0];
      pcg64_advance(g, (uint64_t)(2 * a.off + first));
      for (int i = 0; i < cnt; ++i) a.d_target[first + i] = (float)rng_uniform(g, -1.0, 2.0);
    }
  }
}

}  // namespace

bool payload_device_ok(unsigned long long M, long long Ng) {
  if (M < 1 || M > 0x100000000ull || Ng < 1) return false;
  const double p_reject = (double)((0x100000000ull - M) % M) / 4294967296.0;
  return p_reject * (double)(Ng + kPayloadSlack) <= 64.0;  // Poisson(64) never reaches kPayloadSlack = 1024
}

size_t payload_scratch_words(long long N) { return (size_t)(N + kPayloadSlack + 2); }

cudaError_t payload_preload() {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, payload_words_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, payload_values_kernel);
  return e;
}

int launch_device_payload(const PayloadArgs& a, cudaStream_t stream) {
  const long long raws = (a.Ng + kPayloadSlack + 1) / 2 + 1;  // raw draws behind the g2 word stream
  const long long wthreads = (raws + kRawsPerThread - 1) / kRawsPerThread;
  // 64-thread blocks (<= 56 registers per thread: 3 584 registers per block): small enough to become resident NEXT TO a
  // block of the persistent rollout kernels -- the TinyImageNetLoc kernel leaves 4 096 registers per SM, the 128- and
  // 256-thread blocks of the first version could only start after the whole window had finished
  constexpr int kPayloadThreads = 64;
  payload_words_kernel<<<(unsigned)((wthreads + kPayloadThreads - 1) / kPayloadThreads), kPayloadThreads, 0, stream>>>(a);
  static_assert(kDrawsPerThread >= 2, "N threads must cover the 2N uniform draws");
  payload_values_kernel<<<(unsigned)((a.N + kPayloadThreads - 1) / kPayloadThreads) + 1, kPayloadThreads, 0, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int)e;
  return 2;
}

}  // namespace apg

// ------------------------------------------------------------------------------------------------------
// Test hook (declared in include/apg_b200.h): one payload drawn on the device from three given stream states.
// ------------------------------------------------------------------------------------------------------
extern "C" int apg_rng_device_payload(int32_t device, apg_pcg64_state* g0, apg_pcg64_state* g1, apg_pcg64_state* g2,
                                      int64_t global_n, int64_t env_offset, int64_t num_envs, uint64_t n_items,
                                      int32_t localization, int32_t invert, int64_t* h_idx, double* h_pos, float* h_target,
                                      uint8_t* h_inv) {
  APG_CHECK(g0 && g1 && g2 && h_idx && h_pos, APG_ERR_INVALID, "apg_rng_device_payload: null argument");
  APG_CHECK(global_n >= 1 && num_envs >= 1 && env_offset >= 0 && env_offset + num_envs <= global_n, APG_ERR_INVALID,
            "apg_rng_device_payload: bad shard");
  APG_CHECK(apg::payload_device_ok(n_items, global_n), APG_ERR_INVALID,
            "apg_rng_device_payload: len(dataset) = %llu rejects too many draws for the device generator",
            (unsigned long long)n_items);
  APG_CUDA(cudaSetDevice(device));
  apg::Pcg64 st[3];
  const apg_pcg64_state* in[3] = {g0, g1, g2};
  for (int i = 0; i < 3; ++i) {
    st[i].state_hi = in[i]->state_hi; st[i].state_lo = in[i]->state_lo; st[i].inc_hi = in[i]->inc_hi;
    st[i].inc_lo = in[i]->inc_lo; st[i].has_uint32 = in[i]->has_uint32; st[i].uinteger = in[i]->uinteger;
  }
  apg::PayloadArgs a{};
  a.Ng = global_n; a.off = env_offset; a.N = num_envs; a.M = n_items; a.loc = localization; a.inv = invert;
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
  apg::Pcg64* d_st = nullptr;
  A(cudaMalloc(&d_st, 6 * sizeof(apg::Pcg64)));
  a.streams = d_st; a.streams_out = d_st + 3;
  A(cudaMalloc(&a.d_idx, num_envs * 8)); A(cudaMalloc(&a.d_pos, num_envs * 16));
  A(cudaMalloc(&a.d_target, num_envs * 8)); A(cudaMalloc(&a.d_inv, num_envs));
  A(cudaMalloc(&a.words, apg::payload_scratch_words(num_envs) * 4));
  A(cudaMalloc(&a.rej, 2 * (apg::kPayloadSlack + 1) * 4));
  if (err == cudaSuccess) {
    a.rej_next = a.rej + apg::kPayloadSlack + 1;
    A(cudaMemset(a.rej, 0, 2 * (apg::kPayloadSlack + 1) * 4));
    A(cudaMemcpy(d_st, st, sizeof(st), cudaMemcpyHostToDevice));
    const int rc = apg::launch_device_payload(a, nullptr);
    if (rc < 0) A((cudaError_t)(-rc));
    else apg::count_launch(rc);
    A(cudaDeviceSynchronize());
    A(cudaMemcpy(st, a.streams_out, sizeof(st), cudaMemcpyDeviceToHost));
    A(cudaMemcpy(h_idx, a.d_idx, num_envs * 8, cudaMemcpyDeviceToHost));
    A(cudaMemcpy(h_pos, a.d_pos, num_envs * 16, cudaMemcpyDeviceToHost));
    if (localization && h_target) A(cudaMemcpy(h_target, a.d_target, num_envs * 8, cudaMemcpyDeviceToHost));
    if (invert && h_inv) A(cudaMemcpy(h_inv, a.d_inv, num_envs, cudaMemcpyDeviceToHost));
  }
  cudaFree(d_st); cudaFree(a.d_idx); cudaFree(a.d_pos); cudaFree(a.d_target); cudaFree(a.d_inv);
  cudaFree(a.words); cudaFree(a.rej);
  if (err != cudaSuccess) return apg::fail(APG_ERR_CUDA, "apg_rng_device_payload: %s", cudaGetErrorString(err));
  apg_pcg64_state* out[3] = {g0, g1, g2};
  for (int i = 0; i < 3; ++i) {
    out[i]->state_hi = st[i].state_hi; out[i]->state_lo = st[i].state_lo; out[i]->inc_hi = st[i].inc_hi;
    out[i]->inc_lo = st[i].inc_lo; out[i]->has_uint32 = st[i].has_uint32; out[i]->uinteger = st[i].uinteger;
  }
  return APG_OK;
}
