// Image step for LARGE glimpses (one warp per environment) on uint8 banks at sensor_scale == 1 -- the
// TinyImageNetLoc shape (64x64x3 bank, 10x10x3 sensor) of BASELINE.json.  Reference path:
//   ImagePerceptionModule.step / get_glimpse      ap_gym/envs/image/image_perception_module.py:191-217,294-331
//   ImageLocalizationVectorEnv._step              ap_gym/envs/image_localization.py:151-181
//
// Two launches per vector step:
//   tma_head_kernel   thread = environment: the two scalar chains of apg_image_dev.cuh (move / loss / reward / flags /
//                     statistics) and a 32-byte tile header per glimpse (sensor centre in pixels, TMA box coordinates);
//   tma_blend_kernel  PERSISTENT warps, warp = environment, lane = (sensor column, channel).  Each warp walks its
//                     environments with a two-stage ring of shared-memory tiles: one elected lane issues ONE tensor-map
//                     copy (cp.async.bulk.tensor.2d, SASS UTMALDG) per glimpse -- box = (S0+1) rows x 48 bytes of the
//                     bank image, start aligned down to 16 bytes so that every row is exactly two 32-byte sectors -- for
//                     environment k+1 while the warp blends environment k; completion is an mbarrier transaction count,
//                     nothing in the block waits for anything but its own tile (no __syncthreads in the loop).
// Blend: the reference's 4-term float64 sum in scipy's order (bit-exact pixels).  The exact uint8 -> float64 pixel
// value float64(float32(u8) / 255) comes from a 256-entry table in shared memory, replicated 16x so that the 64-bit
// lookups of a half-warp never collide (one LDS.64 instead of the five-instruction conversion); with consecutive tap
// rows every tile row is fetched and converted once.  Environments whose taps are not consecutive (a float64 rounding
// accident at a grid line, probability ~1e-15 per tap) take a general per-output path from global memory.
#include "apg_image_tma.h"

#include <cstdint>

namespace apg_img {
namespace {

constexpr int kWarps = 8;
constexpr int kReplicas = 16;                       // table copies: lane & 15 picks its own -> conflict-free LDS.64
constexpr int kTableBytes = 256 * kReplicas * 8;    // 32 KB

struct __align__(16) TileHdr {  // 32 bytes per glimpse, head kernel -> blend kernel (through L2)
  double ccx, ccy;  // sensor centre in image-centred pixel coordinates (float64, like the reference)
  int y0;           // tensor-map row of the box: bank index * H + first tap row
  int x0;           // tensor-map byte column of the box: first tap column * C, aligned down to 16
  int idx;          // bank index (general path)
  int pad;
};

template <int S0, int S1, int C>
struct Geom {
  static constexpr int G = S1 * C;                            // floats per sensor row
  static constexpr int P = S0 * G;
  static constexpr int BOX_W = ((S1 + 1) * C + 15 + 15) & ~15;  // footprint bytes + lead of the aligned-down start
  static constexpr int BOX_H = S0 + 1;
  static constexpr int TILE_BYTES = BOX_W * BOX_H;
  static constexpr int TILE_STRIDE = (TILE_BYTES + 127) & ~127;  // tensor-map copies land on 128-byte boundaries
  static constexpr int TAPS = S0 + S1;
  static constexpr int WARP_BYTES = (2 * TILE_STRIDE + TAPS * 16 + 16 + 127) & ~127;
  static constexpr int SMEM = kTableBytes + kWarps * WARP_BYTES;
  static_assert(G <= 32 && TAPS <= 32, "one lane per (column, channel) and one lane per tap");
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// one glimpse footprint: box (BOX_W bytes x BOX_H rows) of the bank tensor at (x, y) -> shared memory, completion
// counted on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}

__device__ __forceinline__ TileHdr load_hdr(const TileHdr* g) {
  const int4 a = reinterpret_cast<const int4*>(g)[0], b = reinterpret_cast<const int4*>(g)[1];
  TileHdr h;
  h.ccx = __hiloint2double(a.y, a.x); h.ccy = __hiloint2double(a.w, a.z);
  h.y0 = b.x; h.x0 = b.y; h.idx = b.z; h.pad = b.w;
  return h;
}

// tile header of one glimpse centred at (ccx, ccy): the tap coordinates are monotone in the tap number, so the first and
// the last one decide the out-of-grid flag, and the first taps give the box origin
template <int C>
__device__ __forceinline__ void write_hdr(const ImgParams& p, TileHdr* out, double ccx, double ccy, int idx, unsigned& err) {
  const double y0 = ccy + (0.0 - p.half0) * p.scale, y1 = ccy + ((double)(p.S0 - 1) - p.half0) * p.scale;
  const double x0 = ccx + (0.0 - p.half1) * p.scale, x1 = ccx + ((double)(p.S1 - 1) - p.half1) * p.scale;
  if (y0 < -p.offy || y0 > p.offy || y1 < -p.offy || y1 > p.offy || x0 < -p.offx || x0 > p.offx || x1 < -p.offx ||
      x1 > p.offx)
    err |= APG_FLAG_OUT_OF_GRID;
  int rmin, cmin;
  double w;
  tap_row(y0, p.offy, p.H, true, rmin, w);
  tap(x0, p.offx, p.W, cmin, w);
  int4 a, b;
  a.x = __double2loint(ccx); a.y = __double2hiint(ccx); a.z = __double2loint(ccy); a.w = __double2hiint(ccy);
  b.x = idx * p.H + rmin; b.y = (cmin * C) & ~15; b.z = idx; b.w = 0;
  reinterpret_cast<int4*>(out)[0] = a;
  reinterpret_cast<int4*>(out)[1] = b;
}

template <int C>
__global__ void __launch_bounds__(256) tma_head_kernel(const ImgParams p) {
  grid_dependency_wait(p.pdl);
  __shared__ double s_stats[6];
  const bool ep_end = (p.mode == 0) && p.terminated;
  if (ep_end && threadIdx.x < 6) s_stats[threadIdx.x] = 0.0;
  if (ep_end) __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < p.N) {
    double px = 0.0, py = 0.0, tcx = 0.0, tcy = 0.0;
    int idx = 0;
    unsigned err = 0;
    env_move(p, n, px, py, tcx, tcy, idx, err);
    TileHdr* hdr = static_cast<TileHdr*>(p.t_hdr);
    write_hdr<C>(p, hdr + n, px * p.lim_x, py * p.lim_y, idx, err);
    if ((p.task == APG_TASK_LOCALIZATION) && (p.mode != 0)) write_hdr<C>(p, hdr + p.N + n, tcx, tcy, idx, err);
    env_loss(p, n, s_stats, ep_end, err);
    if (err) atomicOr(p.err, err);
  }
  if (ep_end) {
    __syncthreads();
    if (threadIdx.x < 6) atomicAdd(p.stats + threadIdx.x, s_stats[threadIdx.x]);
  }
}

// General path of one environment (any tap pattern): every lane evaluates outputs o = lane, lane + 32, ... from global
// memory with the reference's per-point taps; rows past the image (weight exactly 0, see tap_row) are read from the row
// above.  Rare by construction -- kept out of line.
template <int S0, int S1, int C>
__device__ __noinline__ void blend_env_general(const unsigned char* bank, int H, int bank_pitch, const Tap* taps, int idx,
                                               int lane, float* out) {
  using Gm = Geom<S0, S1, C>;
  const unsigned char* img = bank + (size_t)idx * H * bank_pitch;
  for (int o = lane; o < Gm::P; o += 32) {
    const int i = o / Gm::G, jc = o - i * Gm::G, j = jc / C, c = jc - j * C;
    const Tap rt = taps[i], ct = taps[S0 + j];
    const int r0 = rt.k, r1 = min(rt.k + 1, H - 1);
    const unsigned char* a = img + (size_t)r0 * bank_pitch + ct.k * C + c;
    const unsigned char* b = img + (size_t)r1 * bank_pitch + ct.k * C + c;
    const double p00 = px_to_f64(a[0]), p01 = px_to_f64(a[C]), p10 = px_to_f64(b[0]), p11 = px_to_f64(b[C]);
    const double wy = rt.w, omy = 1.0 - rt.w, wx = ct.w, omx = 1.0 - ct.w;
    double acc = p00 * (omy * omx);
    acc = acc + p01 * (omy * wx);
    acc = acc + p10 * (wy * omx);
    acc = acc + p11 * (wy * wx);
    out[o] = (float)acc;
  }
}

template <int S0, int S1, int C>
__global__ void __launch_bounds__(kWarps * 32, 4) tma_blend_kernel(const ImgParams p, const __grid_constant__ CUtensorMap tmap) {
  using Gm = Geom<S0, S1, C>;
  extern __shared__ __align__(128) unsigned char smem[];
  double* table = reinterpret_cast<double*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* wbase = smem + kTableBytes + warp * Gm::WARP_BYTES;
  unsigned char* tiles = wbase;
  Tap* taps = reinterpret_cast<Tap*>(wbase + 2 * Gm::TILE_STRIDE);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(wbase + 2 * Gm::TILE_STRIDE + Gm::TAPS * 16);

  // block prologue (independent of the previous grid, so it overlaps its tail under programmatic dependent launch):
  // the conversion table and this warp's two barriers
  for (int q = tid; q < 256 * kReplicas; q += kWarps * 32) table[q] = px_to_f64((unsigned char)(q >> 4));
  if (lane == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  grid_dependency_wait(p.pdl);

  const int n_sets = ((p.task == APG_TASK_LOCALIZATION) && (p.mode != 0)) ? 2 : 1;
  const int n_work = n_sets * p.N;
  const int stride = gridDim.x * kWarps;
  const TileHdr* hdrs = static_cast<const TileHdr*>(p.t_hdr);

  // lane = (column j, channel c) of the sensor row; lanes past the row run column G-1 and skip the store
  const bool lane_ok = lane < Gm::G;
  const int jc = lane_ok ? lane : Gm::G - 1;
  const int j = jc / C, c = jc - j * C;
  const double* tab = table + (lane & (kReplicas - 1));

  int q = blockIdx.x * kWarps + warp;
  TileHdr h0, h1;
  h0.ccx = h0.ccy = 0.0; h0.y0 = h0.x0 = h0.idx = h0.pad = 0;
  h1 = h0;
  if (q < n_work) {
    h0 = load_hdr(hdrs + q);
    if (lane == 0) {
      mbar_expect_tx(&mbar[0], Gm::TILE_BYTES);
      tma_load_2d(tiles, &tmap, h0.x0, h0.y0, &mbar[0]);
    }
  }
  if (q + stride < n_work) h1 = load_hdr(hdrs + q + stride);

  for (int it = 0; q < n_work; q += stride, ++it) {
    const int s = it & 1;
    const int qn = q + stride;
    // stage s^1 was last read in the previous iteration (program order + the __syncwarp that ends it): refill it
    if (qn < n_work && lane == 0) {
      mbar_expect_tx(&mbar[s ^ 1], Gm::TILE_BYTES);
      tma_load_2d(tiles + (s ^ 1) * Gm::TILE_STRIDE, &tmap, h1.x0, h1.y0, &mbar[s ^ 1]);
    }
    TileHdr h2 = h1;
    if (qn + stride < n_work) h2 = load_hdr(hdrs + qn + stride);  // lands while this environment is blended

    const int set = (q >= p.N) ? 1 : 0;
    const int n = q - set * p.N;
    float* out = (set == 0 ? p.glimpse : p.target_glimpse) + (size_t)n * Gm::P;

    // ---- taps: lane i < S0 -> sensor row i, lane S0 + j -> sensor column j (scipy find_indices, float64)
    int k = 0;
    if (lane < Gm::TAPS) {
      const bool is_row = lane < S0;
      const int t = is_row ? lane : lane - S0;
      const double coord = (is_row ? h0.ccy : h0.ccx) + ((double)t - (is_row ? p.half0 : p.half1)) * p.scale;
      double w;
      tap_row(coord, is_row ? p.offy : p.offx, is_row ? p.H : p.W, is_row, k, w);
      Tap tp;
      tp.w = w; tp.k = k; tp.pad = 0;
      taps[lane] = tp;
    }
    const int rmin = __shfl_sync(0xffffffffu, k, 0), cmin = __shfl_sync(0xffffffffu, k, S0);
    // fast path needs consecutive tap rows (vertical reuse) and columns inside the staged box
    const bool mine_bad = (lane < S0) ? (k - rmin != lane)
                                      : ((lane < Gm::TAPS) ? ((k - cmin) < 0 || (k - cmin) > lane - S0) : false);
    const bool slow = __any_sync(0xffffffffu, mine_bad) || p.force_slow;
    unsigned cb0 = 0u, cb1 = 0u;  // visited-mask column bits of this glimpse (columns {j0, j0+1} of every tap)
    if (lane >= S0 && lane < Gm::TAPS) {
      if (k < 32) cb0 |= 1u << k; else if (k < 64) cb1 |= 1u << (k - 32);
      if (k + 1 < 32) cb0 |= 1u << (k + 1); else if (k + 1 < 64) cb1 |= 1u << (k - 31);
    }
    cb0 = __reduce_or_sync(0xffffffffu, cb0);
    cb1 = __reduce_or_sync(0xffffffffu, cb1);
    __syncwarp();

    mbar_wait(&mbar[s], (unsigned)((it >> 1) & 1));  // this environment's tile has landed
    if (!slow) {
      const Tap ct = taps[S0 + j];
      const double wx = ct.w, omx = 1.0 - ct.w;
      const unsigned char* colp = tiles + s * Gm::TILE_STRIDE + ((cmin * C) & 15) + (ct.k - cmin) * C + c;
      double p00 = tab[(int)colp[0] * kReplicas], p01 = tab[(int)colp[C] * kReplicas];
#pragma unroll
      for (int i = 0; i < S0; ++i) {
        colp += Gm::BOX_W;
        const double p10 = tab[(int)colp[0] * kReplicas], p11 = tab[(int)colp[C] * kReplicas];
        const double wy = taps[i].w, omy = 1.0 - wy;
        double acc = p00 * (omy * omx);  // the reference's leading `0.0 +` is an exact no-op for products >= +0
        acc = acc + p01 * (omy * wx);
        acc = acc + p10 * (wy * omx);
        acc = acc + p11 * (wy * wx);
        // uint8 pixels are in [0,1] and the weights sum to 1 (+ a few ulp): float32(acc) is already clipped
        if (lane_ok) out[i * Gm::G + jc] = (float)acc;
        p00 = p10;
        p01 = p11;
      }
    } else {
      blend_env_general<S0, S1, C>(static_cast<const unsigned char*>(p.bank), p.H, p.bank_pitch, taps, h0.idx, lane, out);
    }

    // ---- visited bitmask of the observation glimpse: rows {i0, i0+1} x columns {j0, j0+1}
    if (p.track && set == 0) {
      unsigned* vm = p.visited + (size_t)n * p.H * p.words;
      if (p.mode != 0) {
        for (int w = lane; w < p.H * p.words; w += 32) vm[w] = 0u;
        __syncwarp();
      }
      if (!slow) {  // rows rmin .. rmin + S0; lane = (row, word), one RED per lane (words <= 2 on this path)
        const int r = (p.words == 2) ? (lane >> 1) : lane, w = (p.words == 2) ? (lane & 1) : 0;
        const unsigned bits = w == 0 ? cb0 : cb1;
        if (r <= S0 && rmin + r < p.H && bits) atomicOr(vm + (rmin + r) * p.words + w, bits);
      } else {
        for (int w = 0; w < p.words; ++w) {
          const unsigned bits = w == 0 ? cb0 : cb1;
          if (!bits) continue;
          for (int rr = lane; rr < 2 * S0; rr += 32) {
            const int row = (rr < S0) ? taps[rr].k : taps[rr - S0].k + 1;
            if (row < p.H) atomicOr(vm + row * p.words + w, bits);
          }
        }
      }
    }
    __syncwarp();  // every lane is done with stage s and the taps before they are rewritten
    h0 = h1;
    h1 = h2;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

bool tma_plan(TmaPlan& plan, const apg_image_cfg& cfg, int bank_pitch, int words, cudaError_t* err) {
  plan = TmaPlan{};
  *err = cudaSuccess;
  if (cfg.bank_dtype != APG_BANK_U8 || cfg.sensor_scale != 1.0 || cfg.blend_f32) return false;
  if (!(cfg.sensor_h == 10 && cfg.sensor_w == 10 && cfg.channels == 3)) return false;
  using Gm = Geom<10, 10, 3>;
  if (words > 2 || (Gm::BOX_H) * words > 32 || cfg.height < Gm::BOX_H || bank_pitch < Gm::BOX_W) return false;
  if (!encode_fn()) return false;
  plan.spec = 1;
  plan.smem = Gm::SMEM;
  plan.box_w = Gm::BOX_W;
  plan.box_h = Gm::BOX_H;
  auto kernel = tma_blend_kernel<10, 10, 3>;
  *err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
  if (*err != cudaSuccess) return false;
  int sms = 0, per_sm = 0;
  *err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
  if (*err == cudaSuccess) *err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarps * 32, plan.smem);
  if (*err != cudaSuccess || sms <= 0 || per_sm <= 0) return false;
  plan.grid = sms * per_sm;
  return true;
}

bool tma_encode(TmaPlan& plan, const void* d_bank, int bank_pitch, long long n_items, int height) {
  plan.have_map = false;
  if (!plan.spec || !encode_fn()) return false;
  const long long rows = n_items * (long long)height;
  if (rows >= (1ll << 31)) return false;  // box coordinates are int32
  const cuuint64_t dims[2] = {(cuuint64_t)bank_pitch, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)bank_pitch};
  const cuuint32_t box[2] = {(cuuint32_t)plan.box_w, (cuuint32_t)plan.box_h};
  const cuuint32_t estr[2] = {1, 1};
  // no L2 promotion: the copies are random 2-sector reads, anything wider would be wasted DRAM traffic
  const CUresult r = encode_fn()(&plan.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(d_bank), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  plan.have_map = (r == CUDA_SUCCESS);
  return plan.have_map;
}

size_t tma_header_bytes(long long n_envs) { return (size_t)2 * (size_t)n_envs * sizeof(TileHdr); }

cudaError_t tma_preload(const TmaPlan& plan) {
  if (!plan.spec) return cudaSuccess;
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, tma_head_kernel<3>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, tma_blend_kernel<10, 10, 3>);
  return e;
}

int tma_launch_step(const TmaPlan& plan, const ImgParams& p, cudaStream_t stream) {
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg{};
  cfg.stream = stream;
  cfg.attrs = at;
  cfg.numAttrs = p.pdl ? 1 : 0;
  cfg.gridDim = dim3((unsigned)((p.N + 255) / 256));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tma_head_kernel<3>, p);
  if (e != cudaSuccess) return -(int)e;
  const int n_sets = ((p.task == APG_TASK_LOCALIZATION) && (p.mode != 0)) ? 2 : 1;
  const int want = (n_sets * p.N + kWarps - 1) / kWarps;
  cfg.gridDim = dim3((unsigned)(want < plan.grid ? want : plan.grid));
  cfg.blockDim = dim3(kWarps * 32);
  cfg.dynamicSmemBytes = plan.smem;
  e = cudaLaunchKernelEx(&cfg, tma_blend_kernel<10, 10, 3>, p, plan.tmap);
  if (e != cudaSuccess) return -(int)e;
  return 2;
}

}  // namespace apg_img
