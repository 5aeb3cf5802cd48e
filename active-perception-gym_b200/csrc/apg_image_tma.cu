// Image step for LARGE glimpses on uint8 banks at sensor_scale == 1 -- the TinyImageNetLoc shape (64x64x3 bank,
// 10x10x3 sensor) of BASELINE.json.  Reference path:
//   ImagePerceptionModule.step / get_glimpse      ap_gym/envs/image/image_perception_module.py:191-217,294-331
//   ImageLocalizationVectorEnv._step              ap_gym/envs/image_localization.py:151-181
//
// ONE launch per rollout window, PERSISTENT warps (one block of 20 warps per SM).  A warp owns a group of <= 32
// consecutive environments and takes it through every step of the window:
//   phase H  lane = environment: the two scalar chains of apg_image_dev.cuh (move / loss / reward / flags / statistics)
//            and a 32-byte tile header per environment in the warp's shared memory (sensor centre in pixels, first tap
//            row / column, tensor-map box coordinates);
//   phase B  the warp blends E = 32 / S1 = 3 environments per pass, lane = (environment of the pass, sensor column): the
//            four weight products of a (row, column) are formed once and serve the C channels.  A two-stage ring of
//            shared-memory tiles per warp: lane q < E issues ONE tensor-map copy (cp.async.bulk.tensor.2d, SASS UTMALDG,
//            L2 evict-first) for the glimpse of slot q of the NEXT pass -- box = (S0 + 1 + q) rows x 48 bytes of the bank
//            image, start aligned down to 16 bytes -- while the warp blends the current pass; completion is an
//            mbarrier transaction count.  Nothing in the loop synchronises more than one warp.
//            Row shift: slot q's box starts q rows early (its own tensor map), so the tile rows the three slots read at
//            the same instruction lie 12 banks apart instead of in the same banks.
//            The float32 rows of a pass are parked in a shared-memory stage and leave as contiguous 128-bit streaming
//            stores; the visited mask takes one 64-bit RED per (slot, row).
// Blend: the reference's 4-term float64 sum in scipy's order (bit-exact pixels).  The exact uint8 -> float64 pixel
// value float64(float32(u8) / 255) comes from a 256-entry float32 table (the widening is exact), replicated 32x -- lane l
// reads copy l, i.e. bank l: one conflict-free wavefront per lookup; the 32 KB table is computed once on the host and
// lands in every block's shared memory with one bulk copy (cp.async.bulk, SASS UBLKCP).  Taps: the first tap of each
// axis is scipy's find_indices in full (phase H), the others are assumed consecutive and VERIFIED -- their weight
// w = coord - grid[k0 + t], computed exactly like the reference's, must lie in [0, 1), which is precisely the
// condition under which find_indices returns that index.  With consecutive tap rows every tile row is fetched and
// converted once per lane.  Environments that break the pattern (a float64 rounding accident at a grid line,
// probability ~1e-15 per tap) take a general per-output path from global memory.
#include "apg_image_tma.h"

#include <cstdint>
#include <vector>

namespace apg_img {
namespace {

// A/B switches (tools/build_variants.py); the defaults are the measured best (profiles/README.md, round 2)
#ifndef APG_TMA_TABLE_F32
#define APG_TMA_TABLE_F32 1  // 1: float32 table, one 4-byte bank per lane; 0: float64 table, 16 copies (equal speed)
#endif
#ifndef APG_TMA_REPLICAS
#define APG_TMA_REPLICAS (APG_TMA_TABLE_F32 ? 32 : 16)
#endif
#ifndef APG_TMA_ROWSHIFT
#define APG_TMA_ROWSHIFT 1  // 1: the tile of slot q of a pass starts q rows early (own tensor map, box height + q): -7 %
#endif
// the box starts at the first tap column aligned down to 16 bytes: tile-mode copies of a uint8 tensor need a 16-byte
// aligned start (an unaligned one raises an illegal instruction on sm_100a -- measured)
#define APG_TMA_XALIGN 16
#ifndef APG_TMA_ARITH
#define APG_TMA_ARITH 0  // 1: convert pixels arithmetically (px_to_f64) instead of the table lookup (experiment)
#endif
#ifndef APG_TMA_WARPS
#define APG_TMA_WARPS 20  // warps per block, one block per SM (96 registers per thread)
#endif
#ifndef APG_TMA_MINB
#define APG_TMA_MINB 1
#endif
#ifndef APG_TMA_BULK_OUT
#define APG_TMA_BULK_OUT 1  // 1: a staged glimpse leaves with ONE bulk store (cp.async.bulk shared -> global), 0: LDS.128 + STG.128
#endif
#ifndef APG_TMA_EARLY_ISSUE
#define APG_TMA_EARLY_ISSUE 0  // 1: the first pass's tiles are requested between the header and the loss chain of phase H
#endif
#ifndef APG_TMA_PREFETCH_IN
#define APG_TMA_PREFETCH_IN 0  // 1: prefetch.global.L2 of the next step's action / prediction at the end of phase H
#endif
#ifndef APG_TMA_BLOCKED
#define APG_TMA_BLOCKED 1  // 1: footprints come from the BLOCKED copy of the bank (8 rows x 16 bytes per 128-byte line) as 16-byte
                           //    cp.async pieces: 6.75 DRAM lines per glimpse instead of 13.75 (-8 % time, DRAM reads halved)
                           // 0: one tensor-map box per glimpse from the row-major bank (cp.async.bulk.tensor.2d)
#endif
#ifndef APG_TMA_STAGE_ROWS
#define APG_TMA_STAGE_ROWS 10  // sensor rows staged in shared memory between two 128-bit copy-outs
#endif
constexpr int kWarps = APG_TMA_WARPS;
constexpr int kReplicas = APG_TMA_REPLICAS;         // table copies: lane & (kReplicas - 1) picks its own
#if APG_TMA_TABLE_F32
typedef float TableT;
#else
typedef double TableT;
#endif
constexpr int kRowShift = APG_TMA_ROWSHIFT;
constexpr int kXAlign = APG_TMA_XALIGN;
constexpr int kTableBytes = 256 * kReplicas * (int)sizeof(TableT);

struct __align__(16) TileHdr {  // 32 bytes per glimpse, phase H -> phase B (warp-private shared memory)
  double ccx, ccy;  // sensor centre in image-centred pixel coordinates (float64, like the reference)
  int y0;           // tensor-map row of the box: bank index * H + first tap row
  int x0;           // tensor-map byte column of the box: first tap column * C, aligned down to 16
  int idx;          // bank index (general path)
  int rc;           // first tap row | first tap column << 16
};

template <int S0, int S1, int C>
struct Geom {
  static constexpr int G = S1 * C;                            // floats per sensor row
  static constexpr int P = S0 * G;
  static constexpr int BOX_W = ((S1 + 1) * C + (APG_TMA_XALIGN - 1) + 15) & ~15;  // footprint bytes + lead of the aligned-down start
  static constexpr int BOX_H = S0 + 1;
  static constexpr int TILE_BYTES = BOX_W * BOX_H;
  // tensor-map copies land on 128-byte boundaries.  Row shift (triple mode): the tile of slot q of a pass is fetched q
  // rows early (box height BOX_H + q), so the rows the lanes of the E slots read at the same instruction sit
  // q * BOX_W bytes = q * 12 banks apart instead of in the same banks (the 3-way LDS.U8 conflict of aligned tiles)
  static constexpr int TILE_ROWS_MAX = BOX_H + (kRowShift ? 32 / S1 - 1 : 0);
  static constexpr int TILE_STRIDE = (BOX_W * TILE_ROWS_MAX + 127) & ~127;
  static constexpr int TAPS = S0 + S1;
  // lane = (environment of the pass, sensor column): E environments per pass, every lane blends the C channels of its
  // column; the float32 rows go through a shared-memory stage so that the global stores are 128-bit and contiguous
  static constexpr int E = 32 / S1;
  static constexpr int STAGE_ROWS = APG_TMA_STAGE_ROWS < S0 ? APG_TMA_STAGE_ROWS : S0;
  static constexpr int STAGE_ENV = ((STAGE_ROWS * G + 31) & ~31) + 4;  // floats per environment: = 4 mod 32 (bank spread), 16-byte multiple
  static constexpr int STAGE_BYTES = E * STAGE_ENV * 4;
  static constexpr int WY_BYTES = E * S0 * 8;
  static constexpr int OFF_WY = 2 * E * TILE_STRIDE;
  static constexpr int OFF_HDR = OFF_WY + ((WY_BYTES + 15) & ~15);
  static constexpr int OFF_STAGE = OFF_HDR + 32 * 32;
  static constexpr int OFF_MBAR = OFF_STAGE + STAGE_BYTES;
  static constexpr int WARP_BYTES = (OFF_MBAR + 16 + 127) & ~127;
  static_assert(S0 <= S1, "lane (environment, column t) verifies tap row t");
  static_assert(!kRowShift || E <= 3, "row shift: one tensor map per slot of a pass");
  static_assert((STAGE_ROWS * G) % 4 == 0 && (S0 * G) % 4 == 0, "128-bit copy-out");
  static constexpr int SMEM = kTableBytes + kWarps * WARP_BYTES + 128;
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
// the same on shared-window addresses computed once (the generic -> shared conversion is not free)
__device__ __forceinline__ void mbar_expect_tx_u(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
// one glimpse footprint: box (BOX_W bytes x box rows) of the bank tensor at (x, y) -> shared memory, completion counted on
// `bar` (shared-window addresses), with an L2 cache policy (createpolicy): glimpse footprints are read once per step and the bank does not fit
// in L2 -- evict-first keeps them from displacing the visited masks, which every step re-reads
__device__ __forceinline__ void tma_load_2d_hint(unsigned dst, const CUtensorMap* map, int x, int y, unsigned bar,
                                                 unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(x), "r"(y), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void ldgsts16_hint(unsigned dst, const void* src, unsigned long long policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy) : "memory");
}
// the barrier receives one arrival from this thread once all its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// contiguous shared -> global bulk store (bytes a multiple of 16, both sides 16-byte aligned) in the thread's bulk group
__device__ __forceinline__ void bulk_store_hint(void* gdst, unsigned ssrc, unsigned bytes, unsigned long long policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(ssrc),
               "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the bulk stores of this thread have finished READING shared memory (their source may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// contiguous global -> shared bulk copy (bytes a multiple of 16), completion counted on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// (double)k for 0 <= k < 2^31 without the conversion unit: 2^52 + k is exact in float64
__device__ __forceinline__ double int_to_double(int k) {
  return __hiloint2double(0x43300000, k) - 4503599627370496.0;
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
  tap_row(x0, p.offx, p.W, true, cmin, w);
  int4 a, b;
  a.x = __double2loint(ccx); a.y = __double2hiint(ccx); a.z = __double2loint(ccy); a.w = __double2hiint(ccy);
  b.x = idx * p.H + rmin; b.y = (cmin * C) & ~(kXAlign - 1); b.z = idx; b.w = rmin | (cmin << 16);
  reinterpret_cast<int4*>(out)[0] = a;
  reinterpret_cast<int4*>(out)[1] = b;
}

// General path of one environment (any tap pattern): scipy's per-point taps, pixels straight from the bank; rows /
// columns past the image (weight exactly 0) are read from the line before.  Also updates the visited mask.  Rare by
// construction -- kept out of line.
template <int S0, int S1, int C>
__device__ __noinline__ void blend_env_general(const unsigned char* bank, int H, int W, int words, int bank_pitch, double ccx,
                                               double ccy, double half0, double half1, double scale, double offy,
                                               double offx, int idx, int lane, float* out, unsigned* vm) {
  using Gm = Geom<S0, S1, C>;
  const unsigned char* img = bank + (size_t)idx * H * bank_pitch;
  for (int o = lane; o < Gm::P; o += 32) {
    const int i = o / Gm::G, jc = o - i * Gm::G, j = jc / C, c = jc - j * C;
    int kr, kc;
    double wy, wx;
    tap(ccy + ((double)i - half0) * scale, offy, H, kr, wy);
    tap(ccx + ((double)j - half1) * scale, offx, W, kc, wx);
    const unsigned char* a = img + (size_t)kr * bank_pitch + kc * C + c;
    const unsigned char* b = a + bank_pitch;
    const double p00 = px_to_f64(a[0]), p01 = px_to_f64(a[C]), p10 = px_to_f64(b[0]), p11 = px_to_f64(b[C]);
    const double omy = 1.0 - wy, omx = 1.0 - wx;
    double acc = p00 * (omy * omx);
    acc = acc + p01 * (omy * wx);
    acc = acc + p10 * (wy * omx);
    acc = acc + p11 * (wy * wx);
    out[o] = (float)acc;
    if (vm && c == 0) {  // rows {kr, kr+1} x columns {kc, kc+1}
      for (int dr = 0; dr < 2; ++dr)
        for (int dc = 0; dc < 2; ++dc) atomicOr(vm + (kr + dr) * words + ((kc + dc) >> 5), 1u << ((kc + dc) & 31));
    }
  }
}

#if APG_TMA_ARITH
#define APG_PX(b) px_to_f64((unsigned char)(b))
#else
#define APG_PX(b) ((double)tab[(int)(b) * kReplicas])
#endif

// IH x IW = the bank image size: compile-time, like the sensor geometry -- the per-environment prologue of phase B runs
// once per glimpse on a warp with 64 registers per thread, every runtime constant it re-reads is an instruction.
template <int S0, int S1, int C, int IH, int IW>
__global__ void __launch_bounds__(kWarps * 32, APG_TMA_MINB)
tma_step_kernel(const ImgParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap1,
                const __grid_constant__ CUtensorMap tmap2, const TableT* __restrict__ g_table, const int group_size,
                const int n_groups, const int bank_rows, const unsigned char* __restrict__ bank_blk, const RolloutPlan ro) {
  using Gm = Geom<S0, S1, C>;
  extern __shared__ __align__(128) unsigned char smem[];
  TableT* table = reinterpret_cast<TableT*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* wbase = smem + kTableBytes + warp * Gm::WARP_BYTES;
  unsigned char* tiles = wbase;
  double* s_wy = reinterpret_cast<double*>(wbase + Gm::OFF_WY);  // row weights of the glimpse(s) being blended
  TileHdr* hdrs = reinterpret_cast<TileHdr*>(wbase + Gm::OFF_HDR);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(wbase + Gm::OFF_MBAR);
  unsigned long long* tbar = reinterpret_cast<unsigned long long*>(smem + kTableBytes + kWarps * Gm::WARP_BYTES);
  __shared__ double s_stats[6];

  // block prologue (independent of the previous grid, so it overlaps its tail under programmatic dependent launch):
  // barriers, then the conversion table arrives with one bulk copy
  if (tid == 0) mbar_init(tbar, 1);
  if (lane == 0) {
    mbar_init(&mbar[0], APG_TMA_BLOCKED ? 32 : 1);
    mbar_init(&mbar[1], APG_TMA_BLOCKED ? 32 : 1);
  }
  if (tid < 6) s_stats[tid] = 0.0;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(tbar, kTableBytes);
    bulk_load(table, g_table, kTableBytes, tbar);
  }
  grid_dependency_wait(p.pdl);

  // lane = (environment es of the pass, column j); lanes past E * S1 shadow the last column of the last environment
  const bool lane_ok = lane < Gm::E * S1;
  const int es_lane = lane_ok ? lane / S1 : Gm::E - 1;
  const int j = lane_ok ? lane - es_lane * S1 : S1 - 1;
  const int trow = min(j, S0 - 1);  // the tap row this lane verifies
  float* stage = reinterpret_cast<float*>(wbase + Gm::OFF_STAGE);
  // lane q < E issues the tensor-map copy of slot q of a pass (its own map: box height + q rows under row shift)
  const unsigned my_tile_u = smem_u32(wbase) + min(lane, Gm::E - 1) * Gm::TILE_STRIDE;
#if APG_TMA_BLOCKED
  constexpr int kPPR = Gm::BOX_W / 16;                            // 16-byte pieces per footprint row
  constexpr int kTailItems = (Gm::E * Gm::BOX_H - 32) * kPPR;     // pieces of the (slot, row) pairs past the 32nd
  static_assert(Gm::E * Gm::BOX_H >= 32 && kTailItems <= 32, "lane = (slot, row) pair, the tail piece by piece");
  const int iq = lane / Gm::BOX_H, ir = lane - iq * Gm::BOX_H;
  const int last_row = bank_rows - 1, ppr_bank = p.bank_pitch >> 4, last_pc = ppr_bank - 1;
#endif
  const TableT* tab = table + (lane & (kReplicas - 1));
  // sensor_scale == 1 on this path (tma_plan): tap coordinate t of an axis = centre + (t - (S-1)/2), the offsets are
  // compile-time constants for the rows and one register per lane for its column
  constexpr double kHalf0 = (S0 - 1) / 2.0, kHalf1 = (S1 - 1) / 2.0;
  const double col_off = (double)j - kHalf1, col_j = (double)j;
  const double row_off_lane = (double)trow - kHalf0, row_lane = (double)trow;
  constexpr int kWords = (IW + 31) / 32;
  static_assert(kWords <= 2 && (S0 + 1) * kWords <= 32, "visited rows x words: one lane each");
  constexpr unsigned long long wmask = (IW >= 64) ? ~0ull : ((1ull << (IW & 63)) - 1ull);
  constexpr double kOffY = (IH - 1) / 2.0, kOffX = (IW - 1) / 2.0;
  const int vis_r = (kWords == 2) ? (lane >> 1) : lane, vis_w = (kWords == 2) ? (lane & 1) : 0;
  const unsigned tiles_u = smem_u32(tiles), mbar_u = smem_u32(mbar);  // shared-window addresses of the ring
  constexpr int vis_words = IH * kWords;
  bool table_ready = false, any_ep_end = false;
  int it = 0;  // tiles consumed so far by this warp: stage = it & 1, barrier parity = (it >> 1) & 1

  // A warp takes its group of environments through ALL steps of the launch before it moves on: warps drift apart, the
  // latency-bound phase H of one hides behind the blends of the others, and launch ramp / tail are paid once per window.
  for (int g = blockIdx.x * kWarps + warp; g < n_groups; g += gridDim.x * kWarps) {
    const int env_base = p.n0 + g * group_size;
    const int cnt = min(group_size, p.n1 - env_base);
    int t = ro.t0;
    bool need_reset = ro.need_reset0 != 0;
   for (int k = 0; k < ro.n_steps; ++k) {
    StepCtx c;
    c.mode = (k == 0 && ro.first_mode == 2) ? 2 : (need_reset ? 1 : 0);
    c.t_stamp = (c.mode == 0) ? t : 0;
    const int t_after = (c.mode == 0) ? t + 1 : 0;
    c.terminated = ((c.mode == 0) && (t_after >= p.T)) ? 1 : 0;
    c.time_obs = ro.time_obs[k];
    c.action = p.action + (size_t)k * p.N * 2;
    c.pred = p.pred + (size_t)k * p.N * ((p.task == APG_TASK_CLASSIFICATION) ? p.K : 2);
    {
      const int slot = (c.mode == 1) ? ro.reset_slot[k] : 0;
      c.new_idx = ro.slot_idx[slot]; c.new_pos = ro.slot_pos[slot]; c.new_target = ro.slot_target[slot];
      c.new_inv = ro.slot_inv[slot];
    }
    c.reward2 = p.reward2 ? p.reward2 + (size_t)k * p.N : nullptr;
    c.done2 = p.done2 ? p.done2 + (size_t)k * p.N : nullptr;
    t = t_after;
    need_reset = c.terminated != 0;
    const bool ep_end = (c.mode == 0) && c.terminated;
    any_ep_end = any_ep_end || ep_end;
    const int n_sets = ((p.task == APG_TASK_LOCALIZATION) && (c.mode != 0)) ? 2 : 1;
    double tcx = 0.0, tcy = 0.0;
    int idx = 0;
    for (int set = 0; set < n_sets; ++set) {
      // the E tiles of one pass -> stage st; slot q's box starts q rows early and is q rows taller (row shift, see Geom)
      // (lane q issues the copy of slot q: the whole warp calls this; the transaction count is posted first, a copy that
      //  completes before it only drives the count negative while the arrival is still pending)
      auto issue_pass = [&](int st, int first, int nn) {
#if APG_TMA_BLOCKED
        // blocked bank (tma_encode): the 16-byte piece pc of bank row `row` lives at ((row / 8) * pieces_per_row + pc) * 128
        // + (row % 8) * 16, i.e. a 128-byte line holds 8 rows x 16 bytes.  DRAM fills whole 128-byte lines (measured:
        // tools/micro/gather_granularity.py); a footprint of 11 rows x 48 bytes touches 6.75 lines of this layout on
        // average instead of 13.75 lines of a row-major bank.  Lane l < 32 copies the PPR pieces of (slot, row) pair l
        // (iq, ir: per-lane constants), the pairs past 32 are spread over the first lanes piece by piece.  Slot q's
        // rows land q rows down in its tile (the bank spread of the row shift without fetching extra rows).  Rows /
        // pieces past the bank are clamped: they are only ever read with weight 0.
        const unsigned tile_st = tiles_u + st * (Gm::E * Gm::TILE_STRIDE);
        const unsigned long long pol = policy_evict_first();
        if (iq < nn) {
          const int2 xy = reinterpret_cast<const int2*>(hdrs + first + iq)[2];
          const int row = min(xy.x + ir, last_row), pc0 = xy.y >> 4;
          const unsigned char* src = bank_blk + ((size_t)(row >> 3) * ppr_bank + pc0) * 128 + (row & 7) * 16;
          const unsigned dst = tile_st + iq * Gm::TILE_STRIDE + (kRowShift * iq + ir) * Gm::BOX_W;
#pragma unroll
          for (int piece = 0; piece < kPPR; ++piece) ldgsts16_hint(dst + piece * 16, src + min(piece, last_pc - pc0) * 128, pol);
        }
        if (lane < kTailItems) {
          const int pair = 32 + lane / kPPR, piece = lane - (lane / kPPR) * kPPR;
          const int q = pair / Gm::BOX_H, r = pair - q * Gm::BOX_H;
          if (q < nn) {
            const int2 xy = reinterpret_cast<const int2*>(hdrs + first + q)[2];
            const int row = min(xy.x + r, last_row), pc = min((xy.y >> 4) + piece, last_pc);
            ldgsts16_hint(tile_st + q * Gm::TILE_STRIDE + (kRowShift * q + r) * Gm::BOX_W + piece * 16,
                          bank_blk + ((size_t)(row >> 3) * ppr_bank + pc) * 128 + (row & 7) * 16, pol);
          }
        }
        cp_async_arrive_noinc(mbar_u + 8 * st);  // one arrival per lane once its pieces have landed (barrier count 32)
#else
        if (lane == 0) mbar_expect_tx_u(mbar_u + 8 * st, nn * Gm::TILE_BYTES + kRowShift * Gm::BOX_W * ((nn * (nn - 1)) >> 1));
        if (lane < nn) {
          const int2 xy = reinterpret_cast<const int2*>(hdrs + first + lane)[2];
          const CUtensorMap* const my_map = (!kRowShift || lane == 0) ? &tmap : (lane == 1 ? &tmap1 : &tmap2);
          tma_load_2d_hint(my_tile_u + st * (Gm::E * Gm::TILE_STRIDE), my_map, xy.y, xy.x - kRowShift * lane, mbar_u + 8 * st,
                           policy_evict_first());
        }
#endif
      };
      // ---------------------------------------------------------------- phase H: lane = environment
      // move -> header -> (the tiles of the first pass are requested: lane q reads the header it has just written) -> loss,
      // reward, flags, statistics while they fly
      unsigned err = 0;
      const int n = env_base + lane;
      if (lane < cnt) {
        if (set == 0) {
          double px = 0.0, py = 0.0;
          env_move(p, c, n, px, py, tcx, tcy, idx, err);
          write_hdr<C>(p, hdrs + lane, px * p.lim_x, py * p.lim_y, idx, err);
        } else {
          write_hdr<C>(p, hdrs + lane, tcx, tcy, idx, err);
        }
      }
#if APG_TMA_EARLY_ISSUE
      issue_pass(it & 1, 0, min(Gm::E, cnt));
#endif
      if (lane < cnt) {
        if (set == 0) {
          env_loss(p, c, n, s_stats, ep_end, err);
#if APG_TMA_PREFETCH_IN
          // the inputs of the next step on their way into L2 (their first use is the head of the next phase H)
          if (k + 1 < ro.n_steps && p.action != nullptr) {
            const int pw = (p.task == APG_TASK_CLASSIFICATION) ? p.K : 2;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(c.action + ((size_t)p.N + n) * 2));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(c.pred + ((size_t)p.N + n) * pw));
          }
#endif
        }
        if (err) atomicOr(p.err, err);
      }
      __syncwarp();
#if !APG_TMA_EARLY_ISSUE
      issue_pass(it & 1, 0, min(Gm::E, cnt));
#endif
      if (!table_ready) {
        mbar_wait(tbar, 0);
        table_ready = true;
      }
      // ---------------------------------------------------------------- phase B: warp = E environments per pass
      // lane = (environment es of the pass, sensor column col): the weight products of a (row, column) are formed once
      // and serve the C channels; one mbarrier stage holds the E tiles of a pass.
      float* const out_base = (set == 0 ? p.glimpse : p.target_glimpse) + (size_t)env_base * Gm::P;
      unsigned* vm = (p.track && set == 0) ? p.visited + (size_t)env_base * vis_words : nullptr;
      for (int e0 = 0; e0 < cnt; e0 += Gm::E, ++it) {
        const int s = it & 1;
        const int nv = min(Gm::E, cnt - e0);
        // stage s^1 was last read in the previous pass (program order + the __syncwarp that ends it): refill it
        if (e0 + Gm::E < cnt) issue_pass(s ^ 1, e0 + Gm::E, min(Gm::E, cnt - e0 - Gm::E));
        const int es = min(es_lane, nv - 1);         // lanes past the pass recompute its last environment and store nothing
        const bool st_ok = lane_ok && es_lane < nv;
        const double2 cc = reinterpret_cast<const double2*>(hdrs + e0 + es)[0];  // sensor centre (x, y)
        const int rc = reinterpret_cast<const int*>(hdrs + e0 + es)[7];          // first taps
        const int rmin = rc & 0xffff, cmin = rc >> 16;
        // ---- taps under the consecutive-index assumption (see the file header): the weight of tap t of an axis is
        // coord_t - grid[k0 + t] with grid[k] = k - off; lane (es, col) owns column tap col and verifies row tap col
        const double gy0 = int_to_double(rmin) - kOffY, gx0 = int_to_double(cmin) - kOffX;
        const double wx = (cc.x + col_off) - (gx0 + col_j), omx = 1.0 - wx;
        const double wyl = (cc.y + row_off_lane) - (gy0 + row_lane);
        const bool bad = !(wx >= 0.0) || !(wx < 1.0) || (cmin + j > IW - 1) || !(wyl >= 0.0) || !(wyl < 1.0) ||
                         (rmin + trow > IH - 1);
        const bool slow = __any_sync(0xffffffffu, bad) || p.force_slow;
        if (st_ok && j < S0) s_wy[es * S0 + j] = wyl;
        if (vm && c.mode != 0) {
          for (int w = lane; w < nv * vis_words; w += 32) vm[w] = 0u;
        }
        __syncwarp();
        mbar_wait_u(mbar_u + 8 * s, (unsigned)((it >> 1) & 1));  // the tiles of this pass have landed
        if (!slow) {
          const unsigned char* colp = tiles + (s * Gm::E + es) * Gm::TILE_STRIDE + kRowShift * es * Gm::BOX_W +
                                      ((cmin * C) & (kXAlign - 1)) + j * C;
          const double* wyp = s_wy + es * S0;
          float* stg = stage + es * Gm::STAGE_ENV + j * C;
          double pa[C], pb[C];
#pragma unroll
          for (int ch = 0; ch < C; ++ch) { pa[ch] = APG_PX(colp[ch]); pb[ch] = APG_PX(colp[C + ch]); }
#if APG_TMA_BULK_OUT
          // the bulk stores of the previous pass have read the stage (they had its whole epilogue and this prologue)
          if (lane < Gm::E) bulk_wait_read();
          __syncwarp();
#endif
#pragma unroll
          for (int r0 = 0; r0 < S0; r0 += Gm::STAGE_ROWS) {
            constexpr int kR = Gm::STAGE_ROWS;
#pragma unroll
            for (int ii = 0; ii < kR; ++ii) {
              const int i = r0 + ii;
              if (i < S0) {
                colp += Gm::BOX_W;
                double qa[C], qb[C];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) { qa[ch] = APG_PX(colp[ch]); qb[ch] = APG_PX(colp[C + ch]); }
                const double wy = wyp[i], omy = 1.0 - wy;
                const double w00 = omy * omx, w01 = omy * wx, w10 = wy * omx, w11 = wy * wx;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                  double acc = pa[ch] * w00;  // the reference's leading `0.0 +` is an exact no-op for products >= +0
                  acc = acc + pb[ch] * w01;
                  acc = acc + qa[ch] * w10;
                  acc = acc + qb[ch] * w11;
                  // uint8 pixels are in [0,1] and the weights sum to 1 (+ a few ulp): float32(acc) is already clipped
                  if (st_ok) stg[ii * Gm::G + ch] = (float)acc;
                  pa[ch] = qa[ch];
                  pb[ch] = qb[ch];
                }
              }
            }
#if APG_TMA_BULK_OUT
            // copy-out: lane q hands the staged glimpse of slot q to the bulk-copy engine (one contiguous 16-byte
            // aligned run of P floats, L2 evict-first: the observation is not read again by the engine) -- no LDS / STG
            // wavefronts on the warp's load-store path.  The generic-proxy stores above are fenced for the async proxy.
            static_assert(kR == S0 && (Gm::P * 4) % 16 == 0 && (Gm::STAGE_ENV * 4) % 16 == 0, "whole glimpses, 16-byte runs");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane < nv) {
              bulk_store_hint(out_base + (size_t)(e0 + lane) * Gm::P, smem_u32(stage) + lane * (Gm::STAGE_ENV * 4), Gm::P * 4,
                              policy_evict_first());
              bulk_commit();
            }
#else
            __syncwarp();
            // copy-out: rows r0 .. r0 + kR of every environment of the pass, 128 bits per lane and store; streaming
            // stores (evict-first: the observation is not read again by the engine)
            static_assert(S0 % kR == 0, "whole stages");
            constexpr int n4 = kR * Gm::G / 4;
#pragma unroll
            for (int q = 0; q < Gm::E; ++q) {
              if (q < nv) {
                const float4* src = reinterpret_cast<const float4*>(stage + q * Gm::STAGE_ENV);
                float4* dst = reinterpret_cast<float4*>(out_base + (size_t)(e0 + q) * Gm::P + r0 * Gm::G);
#pragma unroll
                for (int w0 = 0; w0 < n4; w0 += 32)
                  if (w0 + 32 <= n4 || w0 + lane < n4) __stcs(dst + w0 + lane, src[w0 + lane]);
              }
            }
            __syncwarp();
#endif
          }
          // visited bitmask: rows rmin .. rmin + S0 x columns cmin .. cmin + S1; lane = (row, word), one RED per lane
          if (vm) {
            if constexpr (kWords == 2) {
              // one 64-bit RED per (slot, row): lane = (slot of the round, row), 32 / (S0 + 1) slots per round
              constexpr int RW = S0 + 1, SPR = 32 / RW;
              const int qs = lane / RW, rr = lane - qs * RW;
#pragma unroll
              for (int qb = 0; qb < Gm::E; qb += SPR) {
                const int q = qb + qs;
                if (qs < SPR && q < nv) {
                  const int rcq = reinterpret_cast<const int*>(hdrs + e0 + q)[7];
                  const int rq = rcq & 0xffff, cq = rcq >> 16;
                  const unsigned long long cb = ((((1ull << (S1 + 1)) - 1ull) << cq) & wmask);
                  if (rq + rr < IH)
                    atomicOr(reinterpret_cast<unsigned long long*>(vm + q * vis_words) + (rq + rr), cb);
                }
              }
            } else {
              for (int q = 0; q < nv; ++q) {
                const int rcq = reinterpret_cast<const int*>(hdrs + e0 + q)[7];
                const int rq = rcq & 0xffff, cq = rcq >> 16;
                const unsigned long long cb = ((((1ull << (S1 + 1)) - 1ull) << cq) & wmask);
                const unsigned bits = vis_w == 0 ? (unsigned)cb : (unsigned)(cb >> 32);
                if (vis_r <= S0 && rq + vis_r < IH && bits) atomicOr(vm + q * vis_words + (rq + vis_r) * kWords + vis_w, bits);
              }
            }
          }
        } else {
          for (int q = 0; q < nv; ++q) {
            const double2 ccq = reinterpret_cast<const double2*>(hdrs + e0 + q)[0];
            const int idxq = reinterpret_cast<const int*>(hdrs + e0 + q)[6];
            blend_env_general<S0, S1, C>(static_cast<const unsigned char*>(p.bank), p.H, p.W, p.words, p.bank_pitch, ccq.x,
                                         ccq.y, p.half0, p.half1, p.scale, p.offy, p.offx, idxq, lane,
                                         out_base + (size_t)(e0 + q) * Gm::P, vm ? vm + q * vis_words : nullptr);
          }
        }
        if (vm) vm += nv * vis_words;
        __syncwarp();  // every lane is done with stage s (and the row weights) before they are refilled
      }
    }
   }
  }
#if APG_TMA_BULK_OUT
  if (lane < Gm::E) bulk_wait_read();    // the last bulk stores must not outlive the block's shared memory
#endif
  if (!table_ready) mbar_wait(tbar, 0);  // a warp without work: the bulk copy must not outlive the block
  if (__syncthreads_or(any_ep_end ? 1 : 0)) {
    if (tid < 6) atomicAdd(p.stats + tid, s_stats[tid]);
  }
}

// row-major bank (pitch a multiple of 16, rows a multiple of 8) -> blocked copy: one thread per 16-byte piece
__global__ void bank_to_blocked_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long rows, int ppr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // destination piece index
  if (i >= rows * ppr) return;
  const long long blk = i >> 3;          // (row / 8) * ppr + pc
  const int rr = (int)(i & 7);
  const long long br = blk / ppr;
  const int pc = (int)(blk - br * ppr);
  dst[i] = src[(br * 8 + rr) * ppr + pc];
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
  if (!(cfg.sensor_h == 10 && cfg.sensor_w == 10 && cfg.channels == 3 && cfg.height == 64 && cfg.width == 64)) return false;
  using Gm = Geom<10, 10, 3>;
  if (words > 2 || (Gm::BOX_H) * words > 32 || cfg.height < Gm::BOX_H || bank_pitch < Gm::BOX_W) return false;
  if (!encode_fn()) return false;
  plan.spec = 1;
  plan.smem = Gm::SMEM;
  plan.box_w = Gm::BOX_W;
  plan.box_h = Gm::BOX_H;
  auto kernel = tma_step_kernel<10, 10, 3, 64, 64>;
  *err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
  if (*err != cudaSuccess) return false;
  int sms = 0, per_sm = 0;
  *err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
  if (*err == cudaSuccess) *err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarps * 32, plan.smem);
  if (*err != cudaSuccess || sms <= 0 || per_sm <= 0) return false;
  plan.grid = sms * per_sm;
  plan.warps = plan.grid * kWarps;
  // the conversion table float64(float32(u) / 255), 16 copies interleaved (entry u of copy r at u * 16 + r)
  std::vector<TableT> t((size_t)256 * kReplicas);
  for (int u = 0; u < 256; ++u)
    for (int r = 0; r < kReplicas; ++r) t[(size_t)u * kReplicas + r] = (TableT)((float)u / 255.0f);
  *err = cudaMalloc(reinterpret_cast<void**>(&plan.d_table), kTableBytes);
  if (*err == cudaSuccess) *err = cudaMemcpy(plan.d_table, t.data(), kTableBytes, cudaMemcpyHostToDevice);
  if (*err != cudaSuccess) return false;
  return true;
}

void tma_release(TmaPlan& plan) {
  if (plan.d_table) cudaFree(plan.d_table);
  plan.d_table = nullptr;
  if (plan.d_bank_blk) cudaFree(plan.d_bank_blk);
  plan.d_bank_blk = nullptr;
}

bool tma_encode(TmaPlan& plan, const void* d_bank, int bank_pitch, long long n_items, int height) {
  plan.have_map = false;
  if (!plan.spec || !encode_fn()) return false;
  const long long rows = n_items * (long long)height;
  if (rows >= (1ll << 31)) return false;  // box coordinates are int32
  const cuuint64_t dims[2] = {(cuuint64_t)bank_pitch, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)bank_pitch};
  const cuuint32_t estr[2] = {1, 1};
  // no L2 promotion: the copies are random 2-sector reads, anything wider would be wasted DRAM traffic
  // (tmap[q]: the box of slot q of a pass, q rows taller -- row shift, see Geom)
  plan.have_map = true;
  plan.bank_rows = (int)rows;
#if APG_TMA_BLOCKED
  if (plan.d_bank_blk) cudaFree(plan.d_bank_blk);
  plan.d_bank_blk = nullptr;
  if (rows % 8 != 0 || bank_pitch % 16 != 0) return plan.have_map = false;
  if (cudaMalloc(&plan.d_bank_blk, (size_t)rows * bank_pitch) != cudaSuccess) return plan.have_map = false;
  {
    const int ppr = bank_pitch / 16;
    const long long pieces = rows * ppr;
    bank_to_blocked_kernel<<<(unsigned)((pieces + 255) / 256), 256>>>(static_cast<const uint4*>(d_bank),
                                                                      static_cast<uint4*>(plan.d_bank_blk), rows, ppr);
    if (cudaDeviceSynchronize() != cudaSuccess) return plan.have_map = false;
  }
#endif
  for (int q = 0; q < 3; ++q) {
    const cuuint32_t box[2] = {(cuuint32_t)plan.box_w, (cuuint32_t)(plan.box_h + kRowShift * q)};
    const CUresult r = encode_fn()(&plan.tmap[q], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(d_bank), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    plan.have_map = plan.have_map && (r == CUDA_SUCCESS);
  }
  return plan.have_map;
}

cudaError_t tma_preload(const TmaPlan& plan) {
  if (!plan.spec) return cudaSuccess;
  cudaFuncAttributes a;
  return cudaFuncGetAttributes(&a, tma_step_kernel<10, 10, 3, 64, 64>);
}

int tma_launch(const TmaPlan& plan, const ImgParams& p, const RolloutPlan& ro, cudaStream_t stream) {
  cudaLaunchAttribute at[1];
  int n_at = 0;
  if (p.pdl) {
    at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n_at].val.programmaticStreamSerializationAllowed = 1;
    n_at++;
  }
  cudaLaunchConfig_t cfg{};
  cfg.stream = stream;
  cfg.attrs = at;
  cfg.numAttrs = n_at;
  // groups of <= 32 consecutive environments, one (or a few) per persistent warp
  // (every warp gets the same number of groups, give or take one short group: m rounds of <= 32 environments)
  const int n_range = p.n1 - p.n0;
  const long long m = ((long long)n_range + 32ll * plan.warps - 1) / (32ll * plan.warps);
  const long long target = m * plan.warps;
  int gs = (int)((n_range + target - 1) / target);
  if (gs > 32) gs = 32;
  if (gs < 1) gs = 1;
  const int n_groups = (n_range + gs - 1) / gs;
  const int want = (n_groups + kWarps - 1) / kWarps;
  cfg.gridDim = dim3((unsigned)(want < plan.grid ? want : plan.grid));
  cfg.blockDim = dim3(kWarps * 32);
  cfg.dynamicSmemBytes = plan.smem;
  const TableT* table = reinterpret_cast<const TableT*>(plan.d_table);
  cudaError_t e = cudaLaunchKernelEx(&cfg, tma_step_kernel<10, 10, 3, 64, 64>, p, plan.tmap[0], plan.tmap[1], plan.tmap[2], table,
                                     gs, n_groups, plan.bank_rows, static_cast<const unsigned char*>(plan.d_bank_blk), ro);
  if (e != cudaSuccess) return -(int)e;
  return 1;
}

}  // namespace apg_img
