// Unique-glimpse search of ImageLocalizationVectorEnv.reset() for sm_100a.
//
// Replaces ImagePerceptionModule.sample_unique_glimpse_positions (ap_gym/envs/image/image_perception_module.py:253-292),
// the O(P^2 * S0*S1*C) per-image search that costs the reference ~1.2 s per 64x64x3 image on a CPU core:
//   1. glimpses of the image at the P grid positions (same float64 bilinear blend as the step kernel, bit-exact);
//   2. diff[a, b] = float32 mean((g_b - g_a)^2) with numpy's pairwise summation order (verified against numpy:
//      0 + pairwise_sum over the S0*S1*C contiguous elements, 8 accumulators per <=128 block, halves split at
//      multiples of 8), uniqueness[a] = min_{b != a} diff[a, b];
//   3. the top_k most unique positions in descending order (np.argsort(-uniqueness)[:top_k]; exact ties -- which numpy
//      resolves by its unstable introsort -- are broken towards the lower index here).
// The result depends only on the bank image, so the host caches it per bank index (ap_gym_b200/image_env.py).
#include <cfloat>
#include <cmath>
#include <vector>

#include "apg_common.h"

namespace {

struct UqParams {
  int H, W, C, S0, S1, D, P, nc, bank_pitch, top_k;
  double scale, lim_x, lim_y, offy, offx, half0, half1;
  const void* bank;
  const long long* img;  // (nc) bank indices of this chunk
  const double* grid;    // (P, 2) normalised positions (x, y)
  float* G;              // (nc, P, D)
  float* U;              // (nc, P)
  int* topk;             // (nc, top_k)
};

__device__ __forceinline__ void uq_tap(double coord, double off, int len, int& k_out, double& w_out) {
  int k = (int)floor(coord + off);
  if ((double)k - off > coord) k -= 1;
  else if ((double)(k + 1) - off <= coord) k += 1;
  k = max(0, min(k, len - 2));
  k_out = k;
  w_out = coord - ((double)k - off);
}

__device__ __forceinline__ double uq_px(unsigned char u) {
  const float r = 0.003921568859368563f;
  const float uf = (float)u;
  const float q0 = __fmul_rn(uf, r);
  const float e = __fmaf_rn(-255.0f, q0, uf);
  return (double)__fmaf_rn(e, r, q0);
}
__device__ __forceinline__ double uq_px(float v) { return (double)v; }

template <typename BankT>
__global__ void __launch_bounds__(256) unique_glimpse_kernel(const UqParams p) {
  const long long total = (long long)p.nc * p.P * p.D;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(q % p.D);
    const long long r = q / p.D;
    const int pp = (int)(r % p.P);
    const int ci = (int)(r / p.P);
    const int c = d % p.C;
    const int j = (d / p.C) % p.S1;
    const int i = d / (p.C * p.S1);
    const double ccx = p.grid[2 * pp] * p.lim_x, ccy = p.grid[2 * pp + 1] * p.lim_y;
    int i0, j0;
    double wy, wx;
    uq_tap(ccy + ((double)i - p.half0) * p.scale, p.offy, p.H, i0, wy);
    uq_tap(ccx + ((double)j - p.half1) * p.scale, p.offx, p.W, j0, wx);
    const BankT* img = reinterpret_cast<const BankT*>(static_cast<const unsigned char*>(p.bank) +
                                                      (long long)p.img[ci] * p.H * p.bank_pitch);
    const int pe = p.bank_pitch / (int)sizeof(BankT);
    const BankT* rp = img + (long long)i0 * pe + j0 * p.C + c;
    const double p00 = uq_px(rp[0]), p01 = uq_px(rp[p.C]), p10 = uq_px(rp[pe]), p11 = uq_px(rp[pe + p.C]);
    const double omy = 1.0 - wy, omx = 1.0 - wx;
    double acc = 0.0 + p00 * (omy * omx);
    acc = acc + p01 * (omy * wx);
    acc = acc + p10 * (wy * omx);
    acc = acc + p11 * (wy * wx);
    p.G[q] = (float)fmin(fmax(acc, 0.0), 1.0);
  }
}

// numpy's FLOAT_pairwise_sum over the squared differences of two glimpses (float32, separately rounded ops)
__device__ float pw_sqdiff(const float* a, const float* b, int n) {
  if (n < 8) {
    float res = 0.0f;
    for (int i = 0; i < n; ++i) {
      const float d = __fsub_rn(b[i], a[i]);
      res = __fadd_rn(res, __fmul_rn(d, d));
    }
    return res;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = __fsub_rn(b[k], a[k]);
      r[k] = __fmul_rn(d, d);
    }
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = __fsub_rn(b[i + k], a[i + k]);
        r[k] = __fadd_rn(r[k], __fmul_rn(d, d));
      }
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) {
      const float d = __fsub_rn(b[i], a[i]);
      res = __fadd_rn(res, __fmul_rn(d, d));
    }
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(pw_sqdiff(a, b, n2), pw_sqdiff(a + n2, b + n2, n - n2));
}

constexpr int kTile = 16;

// block (16 x 16): thread (ty, tx) owns row a = tile_a*16 + ty and scans b = tile_b*16 + tx over all tiles
__global__ void __launch_bounds__(kTile* kTile) unique_pairs_kernel(const UqParams p) {
  extern __shared__ float sm[];
  const int Dp = p.D | 1;
  float* A = sm;
  float* B = sm + kTile * Dp;
  const int tx = threadIdx.x & (kTile - 1), ty = threadIdx.x / kTile;
  const int ci = blockIdx.y;
  const int a0 = blockIdx.x * kTile;
  const float* Gi = p.G + (size_t)ci * p.P * p.D;
  for (int q = threadIdx.x; q < kTile * p.D; q += blockDim.x) {
    const int r = q / p.D, d = q - r * p.D;
    A[r * Dp + d] = (a0 + r < p.P) ? Gi[(size_t)(a0 + r) * p.D + d] : 0.0f;
  }
  float best = INFINITY;
  const int a = a0 + ty;
  for (int b0 = 0; b0 < p.P; b0 += kTile) {
    __syncthreads();
    for (int q = threadIdx.x; q < kTile * p.D; q += blockDim.x) {
      const int r = q / p.D, d = q - r * p.D;
      B[r * Dp + d] = (b0 + r < p.P) ? Gi[(size_t)(b0 + r) * p.D + d] : 0.0f;
    }
    __syncthreads();
    const int b = b0 + tx;
    if (a < p.P && b < p.P && a != b) {
      const float s = __fadd_rn(0.0f, pw_sqdiff(A + ty * Dp, B + tx * Dp, p.D));
      best = fminf(best, __fdiv_rn(s, (float)p.D));
    }
  }
  for (int o = kTile >> 1; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (tx == 0 && a < p.P) p.U[(size_t)ci * p.P + a] = best;
}

// block per image: top_k rounds of (value desc, index asc) arg-max
__global__ void __launch_bounds__(256) unique_topk_kernel(const UqParams p) {
  extern __shared__ float su[];  // P values
  __shared__ float s_val[256];
  __shared__ int s_idx[256];
  const int ci = blockIdx.x;
  for (int q = threadIdx.x; q < p.P; q += blockDim.x) su[q] = p.U[(size_t)ci * p.P + q];
  __syncthreads();
  for (int k = 0; k < p.top_k; ++k) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int q = threadIdx.x; q < p.P; q += blockDim.x) {
      const float v = su[q];
      if (v > bv || (v == bv && q < bi)) { bv = v; bi = q; }
    }
    s_val[threadIdx.x] = bv;
    s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        const float v = s_val[threadIdx.x + o];
        const int i2 = s_idx[threadIdx.x + o];
        if (v > s_val[threadIdx.x] || (v == s_val[threadIdx.x] && i2 < s_idx[threadIdx.x])) {
          s_val[threadIdx.x] = v;
          s_idx[threadIdx.x] = i2;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const int w = s_idx[0] == 0x7fffffff ? 0 : s_idx[0];
      p.topk[(size_t)ci * p.top_k + k] = w;
      su[w] = -INFINITY;  // uniqueness values are >= 0 (or +inf for P == 1): never selected twice
    }
    __syncthreads();
  }
}

}  // namespace

// declared in apg_image.cu
extern "C" int apg_image_bank_info(apg_image_env* env, const void** d_bank, int32_t* bank_pitch, int64_t* n_items,
                                   apg_image_cfg* cfg, void** stream);

extern "C" int apg_image_unique_topk(apg_image_env* env, const int64_t* h_img_idx, int64_t n_imgs, const double* h_grid,
                                     int32_t n_positions, int32_t top_k, int32_t* h_topk) {
  APG_CHECK(env && h_img_idx && h_grid && h_topk && n_imgs >= 0 && n_positions > 0 && top_k > 0, APG_ERR_INVALID,
            "apg_image_unique_topk: bad argument");
  const void* d_bank = nullptr;
  int32_t pitch = 0;
  int64_t n_items = 0;
  apg_image_cfg c{};
  void* stream_v = nullptr;
  int rc = apg_image_bank_info(env, &d_bank, &pitch, &n_items, &c, &stream_v);
  if (rc != APG_OK) return rc;
  APG_CHECK(d_bank, APG_ERR_STATE, "apg_image_unique_topk before apg_image_set_bank");
  APG_CHECK(top_k <= n_positions, APG_ERR_INVALID, "top_k larger than the number of grid positions");
  for (int64_t i = 0; i < n_imgs; ++i)
    APG_CHECK(h_img_idx[i] >= 0 && h_img_idx[i] < n_items, APG_ERR_INVALID, "bank index out of range");
  if (n_imgs == 0) return APG_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  APG_CUDA(cudaSetDevice(c.device));
  UqParams p{};
  p.H = c.height; p.W = c.width; p.C = c.channels; p.S0 = c.sensor_h; p.S1 = c.sensor_w;
  p.D = c.sensor_h * c.sensor_w * c.channels; p.P = n_positions; p.bank_pitch = pitch; p.top_k = top_k;
  p.scale = c.sensor_scale;
  p.lim_x = ((double)c.width - 1.0) / 2.0 - ((double)c.sensor_h * c.sensor_scale - 1.0) / 2.0;
  p.lim_y = ((double)c.height - 1.0) / 2.0 - ((double)c.sensor_w * c.sensor_scale - 1.0) / 2.0;
  p.offy = ((double)c.height - 1.0) / 2.0; p.offx = ((double)c.width - 1.0) / 2.0;
  p.half0 = ((double)c.sensor_h - 1.0) / 2.0; p.half1 = ((double)c.sensor_w - 1.0) / 2.0;
  p.bank = d_bank;
  const size_t per_img = (size_t)p.P * p.D * sizeof(float);
  int64_t chunk = (int64_t)((256ull << 20) / per_img);
  if (chunk < 1) chunk = 1;
  if (chunk > n_imgs) chunk = n_imgs;
  if (chunk > 65535) chunk = 65535;
  float *dG = nullptr, *dU = nullptr;
  int* dT = nullptr;
  long long* dI = nullptr;
  double* dGrid = nullptr;
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
  A(cudaMalloc(reinterpret_cast<void**>(&dG), per_img * chunk));
  A(cudaMalloc(reinterpret_cast<void**>(&dU), (size_t)chunk * p.P * sizeof(float)));
  A(cudaMalloc(reinterpret_cast<void**>(&dT), (size_t)chunk * top_k * sizeof(int)));
  A(cudaMalloc(reinterpret_cast<void**>(&dI), (size_t)chunk * sizeof(long long)));
  A(cudaMalloc(reinterpret_cast<void**>(&dGrid), (size_t)p.P * 2 * sizeof(double)));
  if (err == cudaSuccess) A(cudaMemcpyAsync(dGrid, h_grid, (size_t)p.P * 2 * sizeof(double), cudaMemcpyHostToDevice, stream));
  const size_t smem_pairs = (size_t)2 * kTile * (p.D | 1) * sizeof(float);
  if (err == cudaSuccess && smem_pairs > 48 * 1024)
    A(cudaFuncSetAttribute(unique_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pairs));
  p.grid = dGrid; p.G = dG; p.U = dU; p.topk = dT; p.img = dI;
  for (int64_t s = 0; s < n_imgs && err == cudaSuccess; s += chunk) {
    const int nc = (int)((n_imgs - s) < chunk ? (n_imgs - s) : chunk);
    p.nc = nc;
    A(cudaMemcpyAsync(dI, h_img_idx + s, (size_t)nc * sizeof(long long), cudaMemcpyHostToDevice, stream));
    const long long total = (long long)nc * p.P * p.D;
    const int blocks = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
    if (c.bank_dtype == APG_BANK_U8) unique_glimpse_kernel<unsigned char><<<blocks, 256, 0, stream>>>(p);
    else unique_glimpse_kernel<float><<<blocks, 256, 0, stream>>>(p);
    unique_pairs_kernel<<<dim3((p.P + kTile - 1) / kTile, nc), kTile * kTile, smem_pairs, stream>>>(p);
    unique_topk_kernel<<<nc, 256, (size_t)p.P * sizeof(float), stream>>>(p);
    apg::count_launch(3);
    A(cudaGetLastError());
    A(cudaMemcpyAsync(h_topk + s * top_k, dT, (size_t)nc * top_k * sizeof(int), cudaMemcpyDeviceToHost, stream));
    A(cudaStreamSynchronize(stream));
  }
  cudaFree(dG); cudaFree(dU); cudaFree(dT); cudaFree(dI); cudaFree(dGrid);
  if (err != cudaSuccess) return apg::fail(APG_ERR_CUDA, "apg_image_unique_topk: %s", cudaGetErrorString(err));
  return APG_OK;
}
