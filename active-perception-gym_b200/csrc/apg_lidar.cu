// Fused LIDAR-localisation step for sm_100a: one launch = one vector step of
// `make_vec("LIDARLoc*-v0", N)` (SyncVectorEnv of TimeLimit(100)-wrapped LIDARLocalization2DEnv).
//
// Per environment (reference file:line in /root/reference/ap_gym):
//   * NEXT_STEP per-env autoreset: map + start cell from the env's own numpy-compatible PCG64 streams
//                                                          envs/lidar_localization2d.py:293-315,547-557 (+ gymnasium SyncVectorEnv)
//   * pose integration, collision ray, wall slide           envs/lidar_localization2d.py:317-375
//   * base reward, normalised-MSE loss, reward              envs/lidar_localization2d.py:330,377; loss_fn.py:100-110,261-267
//   * TimeLimit termination + time_step obs                 time_limit.py:113-139
//   * B-beam scan against the occupancy bitmap              envs/lidar_localization2d.py:238-277,496-536
//
// Mapping: a block owns 32 environments.  Their bit-packed maps are staged in shared memory once per step; the
// serial part of an environment's step (collision ray, then up to two wall-slide rays) runs with lane = environment,
// so 32 environments advance in parallel; the B beams then run with a group of G = min(32, pow2 >= B) lanes per
// environment, lane l casting beams l, l+G, ... and storing coalesced runs of the lidar row (see lidar_step_kernel).
//
// Ray casting = the closed-cell grid walk of oracle/lidar_oracle.c, same IEEE operations in the same order
// (file compiled with -fmad=false; the one fused operation is the explicit fma of the float64 dot product).
#include <cmath>
#include <cstring>
#include <vector>

#include "apg_common.h"
#include "apg_lidar_filter.cuh"
#include "apg_maze.cuh"
#include "apg_rng.h"
#include "apg_rooms.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kEnvsPerBlock = 32;
constexpr int kFilterQ = 8;      // sub-cells per cell and axis of the swath-pattern table (16 bytes x B x 64 entries)
constexpr int kQueueCap = 2048;  // rays per filter round of a block (a round covers kQueueCap / B environments)

struct LidarParams {
  int N, W, H, words, B, G, T, M;
  int n0, n1;          // environments [n0, n1) of the shard are stepped by this launch (n0 a multiple of 32)
  int pad, pw, prows;  // shared-memory maps carry a free border of `pad` cells: (H + 2 pad) rows of pw words
  int pmw;             // words per padded map slot: prows * pw rounded up to a multiple of 4 (16-byte bulk copies)
  int static_map, static_idx, emit_cells, sparse;
  int filter;  // 1: beams are pre-filtered (apg_lidar_filter.cuh) and only the rays that may touch a wall are walked
  int fQ;                   // sub-cells per cell and axis of the swath-pattern table
  const uint4* ftable;      // (B, fQ, fQ) swath patterns
  int mode;  // 0 = step with per-env autoreset, 2 = explicit reset (seed all streams), 3 = explicit reset (keep streams)
  float range;
  unsigned long long seed;  // mode 2: env i is seeded with seed + env_offset + i
  long long env_offset;
  const unsigned* bank;     // (M, H, words)
  const unsigned* bank_pad; // (M, pmw) the same maps in the padded shared-memory form (bank mode), or nullptr
  // procedural-dataset mode (apg_lidar_set_map_generator): every environment owns its current map, generated on the
  // device at its (auto)reset by lidar_mapgen_kernel BEFORE this kernel; map_idx then holds the dataset index (uint32)
  const unsigned* env_maps; // (N, H, words) or nullptr (bank mode)
  int* gen_count;           // length of the reset list (zeroed here for the next step's list kernel)
  const float* dirs;        // (B, 2) lidar_directions
  const float* action;
  const float* pred;
  // state
  float* pos;               // (N,2)
  float* init;              // (N,2)
  unsigned char* aliased;   // (N)
  int* elapsed;             // (N)
  int* map_idx;             // (N)
  unsigned char* prev_done; // (N)
  apg::Pcg64* g0;           // (N) env-level Generator
  apg::Pcg64* g2;           // (N) DatasetIterator Generator
  // outputs
  float* lidar;             // (N,B)
  float* odometry;          // (N,2)
  float* time_step;         // (N)
  float* map_obs;           // (N,H,W) dynamic maps only
  float* reward;
  float* base_reward;
  float* loss;
  float* target;            // (N,2)
  unsigned char* terminated;
  unsigned char* truncated;
  unsigned char* stepped;   // (N) 1 = this env took a real step (0 = autoreset step)
  long long* map_idx_out;   // (N)
  int* cells;               // (N,B,2) optional
  float* dists;             // (N,B) optional raw distances
  float* reward2;
  unsigned char* done2;
  // episode statistics (ActiveRegressionLogWrapper): running sums, per-env results at `terminated`, shard sums
  float* acc1; float* acc2; float* acc_ret;
  float* ep_avg1; float* ep_final1; float* ep_avg2; float* ep_final2; float* ep_return;
  double* stats;
  unsigned* err;
};

// One environment's occupancy bitmap in shared memory with a FREE border of `pad` cells on every side: a ray of length
// <= lidar_range that starts inside the map never leaves the padded bitmap, so the walk needs no bounds test (cells
// outside the map are free: the reference's union only has in-map boxes).  Coordinates passed to is_wall are PADDED
// cell coordinates (x + pad, y + pad).
struct MapView {
  const unsigned* bits;  // shared memory, (H + 2 pad) rows of pw words
  int pw, pad;
};

// mbarrier / bulk-copy helpers of the map staging: one cp.async.bulk (SASS UBLKCP) per environment
__device__ __forceinline__ unsigned lidar_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lidar_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lidar_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void lidar_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lidar_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lidar_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(lidar_smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void lidar_bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(lidar_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(lidar_smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ int is_wall(const MapView& m, int xp, int yp) {
  return (int)((m.bits[yp * m.pw + (xp >> 5)] >> (xp & 31)) & 1u);
}

__device__ __forceinline__ float norm2f(float a, float b) {
  return __fsqrt_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
}

struct WalkState {
  int cx, cy, n_pos, n_zero;
  bool have_first, in_iv;
  double iv_t0, fex, fey;
};

// The grid walk of scan_one from the start cell to the end of the segment (cell coordinates are padded ones).  NB = the
// ray runs along a grid line, so the cells on its other side touch it too (general, rare); NB = false drops those
// lookups from the loop.  An ordinary iteration is: next crossing, step the cell, one bit test, the crossing parameter
// of the next grid line; the interval bookkeeping of oracle/lidar_oracle.c applies only when the wall state changes or
// a lattice corner is touched (about twice per ray; ptxas predicates it, a warp-vote guard measured slower).
template <bool NB>
__device__ __forceinline__ void walk(const MapView& m, WalkState& ws, double px, double py, double dx, double dy, int sx,
                                     int sy, double inv_dx, double inv_dy, bool nbx, bool nby) {
  int cx = ws.cx, cy = ws.cy, n_pos = ws.n_pos, n_zero = ws.n_zero;
  bool have_first = ws.have_first, in_iv = ws.in_iv;
  double iv_t0 = ws.iv_t0, fex = ws.fex, fey = ws.fey;
  // The next grid line of an axis follows from the current cell: g = c + (s > 0 ? 1 : 0) - pad in MAP coordinates, an
  // exact integer, so both crossing parameters t = (g - p) * inv are simply recomputed from (cx, cy) in every iteration
  // -- the same separately rounded operations on the same values as updating them only on the axis that stepped, but
  // no loop-carried float64 state (ptxas turned the conditional updates into unconditional arithmetic + four FSEL +
  // two moves per axis: 18 of ~60 instructions per crossing).  An axis the ray does not move along gets a grid line far
  // to the right and inv = +inf: t = +inf without a select.
  constexpr int kBias = 1 << 20;
  const int offx = (sx != 0 ? (sx > 0 ? 1 : 0) - m.pad : (1 << 19)) + kBias;
  const int offy = (sy != 0 ? (sy > 0 ? 1 : 0) - m.pad : (1 << 19)) + kBias;
  const double invx = sx != 0 ? inv_dx : INFINITY, invy = sy != 0 ? inv_dy : INFINITY;
  constexpr double kMagic = 4503599627370496.0 + (double)kBias;  // 2^52 + bias: (double)(k - bias) = hilo(0x43300000, k) - kMagic
  for (;;) {
    const double gx = __hiloint2double(0x43300000, cx + offx) - kMagic;
    const double gy = __hiloint2double(0x43300000, cy + offy) - kMagic;
    const double tX = (gx - px) * invx, tY = (gy - py) * invy;
    const bool stepx = tX <= tY, stepy = tY <= tX;
    const double tN = stepx ? tX : tY;
    if (!(tN <= 1.0)) break;  // the segment ends inside the current cell
    const int ncx = cx + (stepx ? sx : 0), ncy = cy + (stepy ? sy : 0);
    bool next_wall = is_wall(m, ncx, ncy);
    if (NB) next_wall = next_wall || (nbx && is_wall(m, ncx - 1, ncy)) || (nby && is_wall(m, ncx, ncy - 1));
    bool side_wall = false;
    if (stepx && stepy) side_wall = is_wall(m, ncx, cy) || is_wall(m, cx, ncy);  // lattice corner
    if (next_wall != in_iv || side_wall) {
      if (in_iv) {
        if (!next_wall) {  // the interval closes at tN (a touching side cell adds nothing to a closed interval)
          if (tN > iv_t0) n_pos++; else n_zero++;
          in_iv = false;
        }
      } else {  // next_wall || side_wall
        if (!have_first) {  // entry vertex of the first piece: crossed coordinates are integral
          have_first = true;
          fex = stepx ? gx : px + tN * dx;
          fey = stepy ? gy : py + tN * dy;
        }
        if (next_wall) { in_iv = true; iv_t0 = tN; } else n_zero++;
      }
    }
    cx = ncx; cy = ncy;
  }
  if (in_iv) { if (iv_t0 < 1.0) n_pos++; else n_zero++; }
  ws.n_pos = n_pos; ws.n_zero = n_zero; ws.fex = fex; ws.fey = fey;
}

// oracle/lidar_oracle.c: apg_oracle_scan_one -- same arithmetic, written branch-light: the walk state is updated
// with selects, the rare cases (rays along a grid line, lattice-corner ties, start on a grid line) take side paths.
__device__ float scan_one(const MapView& m, float posx, float posy, float tx, float ty, int& cellx, int& celly) {
  const double px = posx, py = posy;
  const double dx = (double)tx - px, dy = (double)ty - py;
  const int sx = dx > 0 ? 1 : (dx < 0 ? -1 : 0), sy = dy > 0 ? 1 : (dy < 0 ? -1 : 0);
  // one IEEE reciprocal per ray and axis; every crossing parameter is (g - p) * inv (oracle/lidar_oracle.c)
  const double inv_dx = sx != 0 ? __ddiv_rn(1.0, dx) : 0.0, inv_dy = sy != 0 ? __ddiv_rn(1.0, dy) : 0.0;
  const double fpx = floor(px), fpy = floor(py);
  const bool px_int = (px == fpx), py_int = (py == fpy);
  int cx = (int)fpx + m.pad, cy = (int)fpy + m.pad;  // padded cell coordinates from here on
  if (px_int && sx < 0) cx -= 1;
  if (py_int && sy < 0) cy -= 1;
  const bool nbx = (sx == 0 && px_int), nby = (sy == 0 && py_int);
  const bool nb_any = nbx || nby;  // the ray runs along a grid line: the cells on its other side touch it too
  WalkState ws;
  ws.n_pos = 0; ws.n_zero = 0; ws.have_first = false;
  ws.in_iv = is_wall(m, cx, cy) || (nb_any && ((nbx && is_wall(m, cx - 1, cy)) || (nby && is_wall(m, cx, cy - 1))));
  ws.iv_t0 = 0.0; ws.fex = px; ws.fey = py;
  if (ws.in_iv) {
    ws.have_first = true;
  } else if (px_int || py_int) {  // start point on a grid line: the cells behind it touch the segment at t = 0
    bool touch = false;
    const int bx = cx - sx, by = cy - sy;
    if (px_int && sx != 0) touch |= is_wall(m, bx, cy) || (nby && is_wall(m, bx, cy - 1));
    if (py_int && sy != 0) touch |= is_wall(m, cx, by) || (nbx && is_wall(m, cx - 1, by));
    if (px_int && py_int && sx != 0 && sy != 0) touch |= (bool)is_wall(m, bx, by);
    if (touch) { ws.n_zero = 1; ws.have_first = true; }
  }
  ws.cx = cx; ws.cy = cy;
  // rays along a grid line (rare) take the general walk; everything else the lean one
  if (nb_any) walk<true>(m, ws, px, py, dx, dy, sx, sy, inv_dx, inv_dy, nbx, nby);
  else walk<false>(m, ws, px, py, dx, dy, sx, sy, inv_dx, inv_dy, false, false);
  const int n_pos = ws.n_pos, n_zero = ws.n_zero;
  const double fex = ws.fex, fey = ws.fey;
  const float full = norm2f(__fsub_rn(tx, posx), __fsub_rn(ty, posy));
  float dist;
  int kind;
  double c64x = 0, c64y = 0;
  float c32x = 0, c32y = 0;
  if (n_pos + n_zero == 0 || (n_pos > 0 && n_zero > 0)) {
    dist = full; kind = 0;
  } else if (n_pos == 1) {
    const double ddx = fex - px, ddy = fey - py;
    const double nrm = sqrt(__fma_rn(ddy, ddy, ddx * ddx));
    dist = (float)fmax(nrm - 1e-3, 0.0);
    kind = 1; c64x = fex; c64y = fey;
  } else if (n_zero == 1) {
    dist = 0.0f; kind = 2; c32x = posx; c32y = posy;
  } else {
    c32x = (float)fex; c32y = (float)fey;
    const float d = norm2f(__fsub_rn(c32x, posx), __fsub_rn(c32y, posy));
    dist = fmaxf(__fsub_rn(d, 1e-3f), 0.0f);
    kind = 2;
  }
  cellx = -1; celly = -1;
  if (kind != 0 && dist < full) {
    if (kind == 1) {
      double fx = floor(c64x), fy = floor(c64y);
      if (fabs(fx - c64x) < 1e-5 && tx < posx) fx -= 1.0;
      if (fabs(fy - c64y) < 1e-5 && ty < posy) fy -= 1.0;
      cellx = (int)fx; celly = (int)fy;
    } else {
      float fx = floorf(c32x), fy = floorf(c32y);
      if (fabsf(__fsub_rn(fx, c32x)) < 1e-5f && tx < posx) fx -= 1.0f;
      if (fabsf(__fsub_rn(fy, c32y)) < 1e-5f && ty < posy) fy -= 1.0f;
      cellx = (int)fx; celly = (int)fy;
    }
  }
  return dist;
}

// start cell: the k-th free cell in row-major order (np.where(map == 0), lidar_localization2d.py:303-311)
__device__ void kth_free_cell(const unsigned* bits, int W, int H, int words, unsigned k, int& col, int& row) {
  col = 0; row = 0;
  for (int y = 0; y < H; ++y) {
    for (int w = 0; w < words; ++w) {
      const int nb = min(32, W - w * 32);
      const unsigned valid = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
      const unsigned freebits = ~bits[y * words + w] & valid;
      const unsigned c = __popc(freebits);
      if (k < c) {
        row = y;
        col = w * 32 + (__fns(freebits, 0, (int)k + 1));
        return;
      }
      k -= c;
    }
  }
}

__device__ __forceinline__ unsigned count_free(const unsigned* bits, int W, int H, int words) {
  unsigned n = 0;
  for (int y = 0; y < H; ++y)
    for (int w = 0; w < words; ++w) {
      const int nb = min(32, W - w * 32);
      const unsigned valid = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
      n += __popc(~bits[y * words + w] & valid);
    }
  return n;
}

// A block owns kEnvsPerBlock = 32 consecutive environments.
//   phase 1  warp 0, lane = environment: autoreset decision; resetting environments draw their map from the env's own
//            numpy-compatible stream (SeedSequence/PCG64/Lemire, apg_rng.h);
//   phase 2  the 32 padded bitmaps arrive in shared memory, one bulk copy (cp.async.bulk, mbarrier) per environment,
//            issued in phase 1 by the lane that decided the map index;
//   phase 3  warp 0, lane = environment: start-cell draw (reset) or pose integration + collision ray + wall-slide
//            rays (step) -- 32 environments advance in parallel instead of one warp re-evaluating one environment's
//            serial ray chain 32 times -- then loss, reward, TimeLimit, episode statistics, state write-back;
//   phase 4  the block writes the map observation of the environments that reset;
//   phase 5  beams: a group of G = min(32, pow2 >= B) lanes owns one environment, lane l casts beams l, l+G, ...
//            and stores coalesced runs of the lidar row.
#ifndef APG_LIDAR_MINB
#define APG_LIDAR_MINB 8  // resident blocks per SM the register allocation aims at (64 registers per thread)
#endif
__global__ void __launch_bounds__(kThreads, APG_LIDAR_MINB) lidar_step_kernel(const LidarParams p) {
  extern __shared__ __align__(128) unsigned smem_maps[];  // kEnvsPerBlock x pmw: padded bitmaps, see MapView
  __shared__ __align__(8) unsigned long long s_bar;     // the bulk copies of the block's maps complete here
  __shared__ float s_posx[kEnvsPerBlock], s_posy[kEnvsPerBlock];
  __shared__ int s_midx[kEnvsPerBlock];
  __shared__ unsigned s_reset_mask;  // bit e: environment e of the block is reset by this step
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int env0 = p.n0 + blockIdx.x * kEnvsPerBlock;
  const int nE = min(kEnvsPerBlock, p.n1 - env0);
  const int mw = p.H * p.words;    // words of one bank map
  const int pmw = p.pmw;           // words of one padded shared-memory map slot
  // the maps are kept in the padded form as well (bank or per-environment slots): an environment's map is ONE bulk copy issued by the
  // lane that has just decided its map index (the first version re-padded 32 maps per block and step with funnel
  // shifts: 178 of 1 550 warp-instructions per env-step)
  if (warp == 0) {
    if (lane == 0) lidar_mbar_init(&s_bar, 1);
    __syncwarp();
  }
  const float Wf = (float)p.W, Hf = (float)p.H;

  // ---------------------------------------------------------------- phase 1
  bool do_reset = false;
  const int n = env0 + lane;          // warp 0's environment
  const bool act = (warp == 0) && (lane < nE);
  if (act) {
    do_reset = (p.mode != 0) || (p.prev_done[n] != 0);
    int midx;
    if (p.env_maps) {
      midx = p.map_idx[n];  // seeding, the dataset-index draw and the map itself: lidar_mapgen_kernel, already done
    } else if (do_reset) {
      apg::Pcg64 g0, g2;
      if (p.mode == 2) {
        apg::pcg64_seed(g0, p.seed + (unsigned long long)(p.env_offset + n));
        if (!p.static_map) {
          const uint64_t s2 = apg::rng_integers_0_to_2p32_inclusive(g0);  // DatasetIterator seed (:552)
          apg::pcg64_seed(g2, s2);
        } else {
          g2 = g0;
        }
        p.g0[n] = g0;
      } else {
        g2 = p.g2[n];
      }
      if (!p.static_map) midx = (int)apg::rng_integers_below(g2, (uint64_t)p.M);  // dataset_iterator.py:30
      else midx = p.static_idx;
      p.g2[n] = g2;
    } else {
      midx = p.map_idx[n];
    }
    s_midx[lane] = midx;
    lidar_bulk_load(smem_maps + (size_t)lane * pmw, p.bank_pad + (size_t)(p.env_maps ? n : midx) * pmw, (unsigned)pmw * 4u,
                    &s_bar);
  }
  if (tid == 0) lidar_mbar_expect_tx(&s_bar, (unsigned)nE * (unsigned)pmw * 4u);
  if (p.gen_count && blockIdx.x == 0 && tid == 0) *p.gen_count = 0;
  if (warp == 0) {
    const unsigned rm = __ballot_sync(0xffffffffu, act && do_reset);
    if (lane == 0) s_reset_mask = rm;
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 2: the maps have landed
  // (padded form: padded word j of padded row yp holds the map bits [32 j - pad, 32 j - pad + 32) of row yp - pad;
  //  written once per map by apg_lidar_set_map_bank / lidar_mapgen_kernel, copied here by the bulk-copy engine)
  lidar_mbar_wait(&s_bar, 0);
  // ---------------------------------------------------------------- phase 3: lane = environment
  if (act) {
    MapView mv{smem_maps + (size_t)lane * pmw, p.pw, p.pad};
    const int midx = s_midx[lane];
    // the unpadded bitmap (reset only: free-cell count / pick)
    const unsigned* gbits = p.env_maps ? p.env_maps + (size_t)n * mw : p.bank + (size_t)midx * mw;
    float posx, posy, initx, inity;
    int elapsed;
    bool aliased;
    unsigned err = 0;
    unsigned char term = 0;
    if (do_reset) {
      apg::Pcg64 g0 = p.g0[n];
      const unsigned n_free = count_free(gbits, p.W, p.H, p.words);
      const unsigned k = (unsigned)apg::rng_integers_below(g0, (uint64_t)n_free);
      int col, row;
      kth_free_cell(gbits, p.W, p.H, p.words, k, col, row);
      p.g0[n] = g0;
      posx = __fadd_rn((float)col, 0.5f);
      posy = __fadd_rn((float)row, 0.5f);
      initx = posx; inity = posy;
      aliased = true;
      elapsed = 0;
      p.reward[n] = 0.0f;
      p.base_reward[n] = 0.0f;
      p.loss[n] = 0.0f;
      p.target[2 * n] = 0.0f; p.target[2 * n + 1] = 0.0f;
      p.stepped[n] = 0;
      p.map_idx[n] = midx;
      if (p.reward2) p.reward2[n] = 0.0f;
    } else {
      posx = p.pos[2 * n]; posy = p.pos[2 * n + 1];
      initx = p.init[2 * n]; inity = p.init[2 * n + 1];
      aliased = p.aliased[n] != 0;
      elapsed = p.elapsed[n];
      float a0 = p.action[2 * n], a1 = p.action[2 * n + 1];
      const float q0 = p.pred[2 * n], q1 = p.pred[2 * n + 1];
      if (isnan(a0) || isnan(a1)) err |= APG_FLAG_NAN_ACTION;
      if (isnan(q0) || isnan(q1)) err |= APG_FLAG_NAN_PRED;
      const float lastx = posx, lasty = posy;
      const float base = __fsub_rn(0.1f, __fmul_rn(1e-3f, __fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1))));
      const float mag = norm2f(a0, a1);
      if (mag > 1.0f) { a0 = __fdiv_rn(a0, mag); a1 = __fdiv_rn(a1, mag); }
      const float tx = __fadd_rn(posx, a0), ty = __fadd_rn(posy, a1);
      float dirx = __fsub_rn(tx, posx), diry = __fsub_rn(ty, posy);
      const float total = norm2f(dirx, diry);
      if (total > 0.0f) {
        dirx = __fdiv_rn(dirx, total); diry = __fdiv_rn(diry, total);
        int c0, c1;
        const float d = scan_one(mv, posx, posy, tx, ty, c0, c1);
        posx = __fadd_rn(posx, __fmul_rn(dirx, d));
        posy = __fadd_rn(posy, __fmul_rn(diry, d));
        const float rem = __fsub_rn(total, d);
        if (rem > 1e-5f) {
          const float rvx = __fmul_rn(dirx, rem), rvy = __fmul_rn(diry, rem);
          const bool kx = rvx > 1e-5f, ky = rvy > 1e-5f;
          if (kx || ky) {
            const float v0 = (kx && ky) ? rvx : (kx ? rvx : rvy);
            const float v1 = (kx && ky) ? rvy : (kx ? rvx : rvy);
            const float d0 = scan_one(mv, posx, posy, __fadd_rn(posx, v0), __fadd_rn(posy, 0.0f), c0, c1);
            const float d1 = scan_one(mv, posx, posy, __fadd_rn(posx, 0.0f), __fadd_rn(posy, v1), c0, c1);
            if (d0 > 0.0f) {
              const float nv = norm2f(v0, 0.0f);
              posx = __fadd_rn(posx, __fmul_rn(__fdiv_rn(v0, nv), d0));
              posy = __fadd_rn(posy, __fmul_rn(__fdiv_rn(0.0f, nv), d0));
            } else {
              const float nv = norm2f(0.0f, v1);
              posx = __fadd_rn(posx, __fmul_rn(__fdiv_rn(0.0f, nv), d1));
              posy = __fadd_rn(posy, __fmul_rn(__fdiv_rn(v1, nv), d1));
            }
          }
        }
      }
      if (aliased) { initx = posx; inity = posy; }
      const bool oob = (posx < 0.0f) || (posy < 0.0f) || (posx >= Wf) || (posy >= Hf);
      posx = fminf(fmaxf(posx, 0.0f), Wf);
      posy = fminf(fmaxf(posy, 0.0f), Hf);
      aliased = false;
      const float t0 = __fsub_rn(__fmul_rn(__fdiv_rn(lastx, Wf), 2.0f), 1.0f);
      const float t1 = __fsub_rn(__fmul_rn(__fdiv_rn(lasty, Hf), 2.0f), 1.0f);
      const float e0 = __fsub_rn(q0, t0), e1 = __fsub_rn(q1, t1);
      const float mse = __fdiv_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), 2.0f);
      const float loss = __fadd_rn(__fmul_rn(mse, 3.0f), 0.0f);
      elapsed += 1;
      term = (oob || elapsed >= p.T) ? 1 : 0;
      const float rew = __fsub_rn(base, loss);
      // euclidean_distance = ||target - prediction|| (float32, no axis), mse = mean((target - prediction)^2)
      const float m1 = norm2f(e0, e1), m2 = mse;
      const bool first = (elapsed == 1);
      const float s1 = __fadd_rn(first ? 0.0f : p.acc1[n], m1);
      const float s2 = __fadd_rn(first ? 0.0f : p.acc2[n], m2);
      const float sr = __fadd_rn(first ? 0.0f : p.acc_ret[n], rew);
      if (term) {
        const float avg1 = __fdiv_rn(s1, (float)elapsed), avg2 = __fdiv_rn(s2, (float)elapsed);
        p.ep_avg1[n] = avg1; p.ep_final1[n] = m1; p.ep_avg2[n] = avg2; p.ep_final2[n] = m2; p.ep_return[n] = sr;
        atomicAdd(p.stats + 0, 1.0); atomicAdd(p.stats + 1, (double)avg1); atomicAdd(p.stats + 2, (double)m1);
        atomicAdd(p.stats + 3, (double)avg2); atomicAdd(p.stats + 4, (double)m2); atomicAdd(p.stats + 5, (double)sr);
      } else {
        p.acc1[n] = s1; p.acc2[n] = s2; p.acc_ret[n] = sr;
      }
      // `-sparse` ids: base_reward - loss * weight, weight = 1.0 on the terminating step, else 0.0
      const float rew_out = p.sparse ? __fsub_rn(base, __fmul_rn(loss, term ? 1.0f : 0.0f)) : rew;
      p.reward[n] = rew_out;
      p.base_reward[n] = base;
      p.loss[n] = loss;
      p.target[2 * n] = t0; p.target[2 * n + 1] = t1;
      p.stepped[n] = 1;
      if (p.reward2) p.reward2[n] = rew_out;
    }
    // ---- state write-back + scalar observations
    p.map_idx_out[n] = p.env_maps ? (long long)(unsigned)midx : (long long)midx;
    p.pos[2 * n] = posx; p.pos[2 * n + 1] = posy;
    p.init[2 * n] = initx; p.init[2 * n + 1] = inity;
    p.aliased[n] = aliased ? 1 : 0;
    p.elapsed[n] = elapsed;
    p.prev_done[n] = term;
    p.terminated[n] = term;
    p.truncated[n] = 0;
    if (p.done2) p.done2[n] = term;
    // TimeLimit._get_time_obs: float32(2.0 * elapsed / T - 1.0) evaluated in float64
    p.time_step[n] = (float)(2.0 * (double)elapsed / (double)p.T - 1.0);
    // odometry normalisation (:263-270)
    p.odometry[2 * n] = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__fsub_rn(posx, initx), -Wf), __fsub_rn(Wf, -Wf)), 2.0f), 1.0f);
    p.odometry[2 * n + 1] = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__fsub_rn(posy, inity), -Hf), __fsub_rn(Hf, -Hf)), 2.0f), 1.0f);
    s_posx[lane] = posx;
    s_posy[lane] = posy;
    if (err) atomicOr(p.err, err);
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 4: map observation of reset environments
  if (p.map_obs) {
    const float one = __fdiv_rn(1.0f, 255.0f);  // map.astype(f32)[..., None] / 255 -> {0, 1/255}
    const int hw = p.H * p.W;
    for (unsigned rm = s_reset_mask; rm; rm &= rm - 1u) {
      const int e = __ffs((int)rm) - 1;
      float* mo = p.map_obs + (size_t)(env0 + e) * hw;
      const unsigned* bits = smem_maps + (size_t)e * pmw;
      for (int q = tid; q < hw; q += kThreads) {
        const int y = q / p.W, x = q - y * p.W;
        const int xp = x + p.pad;
        mo[q] = ((bits[(y + p.pad) * p.pw + (xp >> 5)] >> (xp & 31)) & 1u) ? one : 0.0f;
      }
    }
  }

  // ---------------------------------------------------------------- phase 5: beams
  if (p.filter) {
    // Filtered form.  Round = kQueueCap / B environments: (a) every ray is tested against its precomputed swath pattern
    // (apg_lidar_filter.cuh, ~60 instructions) -- rays that cannot touch a wall store |target - pos| at once, the others
    // are appended to a block-wide queue (warp-aggregated); (b) the queue is walked by full warps.  In open maps half of
    // the rays never reach the exact walk, which costs ~15x the test.
    unsigned short* queue = reinterpret_cast<unsigned short*>(smem_maps + (size_t)kEnvsPerBlock * pmw + 1);
    __shared__ int s_qn;
    const int segs_per_env = (p.B + 31) >> 5;
    const int envs_per_round = max(1, kQueueCap / (segs_per_env * 32));
    for (int eb = 0; eb < nE; eb += envs_per_round) {
      if (tid == 0) s_qn = 0;
      __syncthreads();
      const int ce = min(envs_per_round, nE - eb);
      for (int el = warp; el < ce; el += kWarpsPerBlock) {
        const int e = eb + el;
        const unsigned* bits = smem_maps + (size_t)e * pmw;
        const float posx = s_posx[e], posy = s_posy[e];
        const size_t row = (size_t)(env0 + e) * p.B;
        for (int b = lane; b < segs_per_env * 32; b += 32) {
          bool hit = false;
          if (b < p.B) {
            hit = apg_lidar::may_touch_wall(bits, p.pw, p.pad, p.ftable, p.fQ, b, posx, posy);
            if (!hit) {
              const float2 dd = reinterpret_cast<const float2*>(p.dirs)[b];
              const float tx = __fadd_rn(posx, dd.x), ty = __fadd_rn(posy, dd.y);
              const float d = norm2f(__fsub_rn(tx, posx), __fsub_rn(ty, posy));
              p.lidar[row + b] = fminf(fmaxf(__fdiv_rn(d, p.range), -1.0f), 1.0f);
              if (p.cells) {
                p.cells[(row + b) * 2] = -1;
                p.cells[(row + b) * 2 + 1] = -1;
                p.dists[row + b] = d;
              }
            }
          }
          const unsigned bal = __ballot_sync(0xffffffffu, hit);
          if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_qn, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) queue[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)((el << 11) | b);
          }
        }
      }
      __syncthreads();
      const int qn = s_qn;
      for (int i = tid; i < qn; i += kThreads) {
        const int q = queue[i];
        const int e = eb + (q >> 11), b = q & 2047;
        MapView mv{smem_maps + (size_t)e * pmw, p.pw, p.pad};
        const float posx = s_posx[e], posy = s_posy[e];
        const float2 dd = reinterpret_cast<const float2*>(p.dirs)[b];
        int c0, c1;
        const float d = scan_one(mv, posx, posy, __fadd_rn(posx, dd.x), __fadd_rn(posy, dd.y), c0, c1);
        const size_t o = (size_t)(env0 + e) * p.B + b;
        p.lidar[o] = fminf(fmaxf(__fdiv_rn(d, p.range), -1.0f), 1.0f);
        if (p.cells) {
          p.cells[o * 2] = c0;
          p.cells[o * 2 + 1] = c1;
          p.dists[o] = d;
        }
      }
      __syncthreads();
    }
    return;
  }
  const int envs_per_warp = 32 / p.G;
  const int gi = lane / p.G;            // group inside the warp
  const int li = lane - gi * p.G;
  for (int e = warp * envs_per_warp + gi; e < nE; e += kWarpsPerBlock * envs_per_warp) {
    MapView mv{smem_maps + (size_t)e * pmw, p.pw, p.pad};
    const float posx = s_posx[e], posy = s_posy[e];
    const size_t row = (size_t)(env0 + e) * p.B;
    for (int b = li; b < p.B; b += p.G) {
      const float ddx = p.dirs[2 * b], ddy = p.dirs[2 * b + 1];
      int c0, c1;
      const float d = scan_one(mv, posx, posy, __fadd_rn(posx, ddx), __fadd_rn(posy, ddy), c0, c1);
      p.lidar[row + b] = fminf(fmaxf(__fdiv_rn(d, p.range), -1.0f), 1.0f);
      if (p.cells) {
        p.cells[(row + b) * 2] = c0;
        p.cells[(row + b) * 2 + 1] = c1;
        p.dists[row + b] = d;
      }
    }
  }
}

// ---- procedural-dataset mode: the reference's per-episode `next(data_loader)` on the device -------------------------
// (lidar_localization2d.py:297-298 -> dataset_iterator.py:27-32 -> floor_map_dataset_{rooms,maze}.py get_data_point)
struct MapGenParams {
  int N, W, H, words, n0, n1, mode, kind;  // kind 1 = rooms, 2 = maze
  int max_rooms, door_width, maze_cells;
  double branching_prob;
  unsigned long long seed;
  long long env_offset;
  const unsigned char* prev_done;
  apg::Pcg64* g0;
  apg::Pcg64* g2;
  int* map_idx;
  unsigned* env_maps;
  unsigned* env_maps_pad;  // (N, pmw): the same map in the padded form the step kernel stages (rows outside the map stay 0)
  int pad, pw, pmw;
  const double* qn_table;
  int* list;
  int* count;
  unsigned* scratch;    // per generator thread: rooms = two bit planes + the view stack, maze = the frame stack
  size_t scratch_words; // per thread
  unsigned* err;
};

// the environments of [n0, n1) that this step resets, compacted (map generation is long and divergent: it runs on the
// compacted list with full warps instead of on the few resetting lanes of every warp)
__global__ void lidar_reset_list_kernel(const MapGenParams p) {
  const int n = p.n0 + blockIdx.x * blockDim.x + threadIdx.x;
  const bool r = n < p.n1 && ((p.mode != 0) || (p.prev_done[n] != 0));
  const unsigned bal = __ballot_sync(0xffffffffu, r);
  if (!bal) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(p.count, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (r) p.list[base + __popc(bal & ((1u << lane) - 1u))] = n;
}

// one thread per list entry: (mode 2: seed the env's streams,) draw the dataset index from the env's DatasetIterator
// stream -- integers(0, 2**32) -- and generate that map into the env's slot
__global__ void lidar_mapgen_kernel(const MapGenParams p) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, total = gridDim.x * blockDim.x;
  const int count = *p.count;
  const int mw = p.H * p.words;
  unsigned* scr = p.scratch + (size_t)gtid * p.scratch_words;
  for (int i = gtid; i < count; i += total) {
    const int n = p.list[i];
    apg::Pcg64 g2;
    if (p.mode == 2) {
      apg::Pcg64 g0;
      apg::pcg64_seed(g0, p.seed + (unsigned long long)(p.env_offset + n));
      const uint64_t s2 = apg::rng_integers_0_to_2p32_inclusive(g0);  // DatasetIterator seed (:552)
      apg::pcg64_seed(g2, s2);
      p.g0[n] = g0;
    } else {
      g2 = p.g2[n];
    }
    const unsigned idx = (unsigned)apg::rng_integers_below(g2, 0x100000000ull);  // dataset_iterator.py:30, len == 2**32
    p.g2[n] = g2;
    p.map_idx[n] = (int)idx;
    unsigned* out = p.env_maps + (size_t)n * mw;
    if (p.kind == 1) {
      const int rc = apg_rooms::generate(idx, p.W, p.H, p.max_rooms, p.door_width, p.qn_table, scr, scr + mw,
                                         reinterpret_cast<apg_rooms::View*>(scr + 2 * mw), out);
      if (rc) atomicOr(p.err, APG_FLAG_MAPGEN);
    } else {
      apg_maze::generate(idx, p.W, p.H, p.branching_prob, out, scr);
    }
    // ... and its padded form: padded word j of padded row y + pad holds the map bits [32 j - pad, 32 j - pad + 32)
    unsigned* outp = p.env_maps_pad + (size_t)n * p.pmw + (size_t)p.pad * p.pw;
    for (int y = 0; y < p.H; ++y) {
      unsigned lo = 0u;
      for (int j = 0; j < p.pw; ++j) {
        const unsigned hi = (j < p.words) ? out[y * p.words + j] : 0u;
        outp[y * p.pw + j] = __funnelshift_l(lo, hi, p.pad);
        lo = hi;
      }
    }
  }
}

}  // namespace

// ======================================================================================================
// host side
// ======================================================================================================
struct apg_lidar_env {
  apg_lidar_cfg cfg{};
  cudaStream_t stream = nullptr;
  int words = 0, G = 0, pad = 0, pw = 0, prows = 0;
  size_t smem = 0;
  unsigned* d_bank = nullptr;
  unsigned* d_bank_pad = nullptr;  // the bank in the padded shared-memory form, pmw words per map (set_map_bank)
  unsigned* d_env_maps_pad = nullptr;  // procedural mode: (N, pmw) padded slots written by the map generator
  int pmw = 0;
  int64_t n_maps = 0;
  float* d_dirs = nullptr;
  bool have_beams = false;
  apg_lidar_buffers b{};
  uint8_t* d_aliased = nullptr;
  int32_t* d_elapsed = nullptr;
  int32_t* d_map_idx = nullptr;
  uint8_t* d_prev_done = nullptr;
  apg::Pcg64* d_g0 = nullptr;
  apg::Pcg64* d_g2 = nullptr;
  unsigned* d_err = nullptr;
  float* d_in_action = nullptr;
  float* d_in_pred = nullptr;
  float* d_acc1 = nullptr; float* d_acc2 = nullptr; float* d_acc_ret = nullptr;
  double* d_stats = nullptr;
  bool has_reset = false;
  // *_host pipelining: sub-range launches, a copy stream and its events (see apg_lidar_step_host)
  int r0 = 0, r1 = -1;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in[8] = {};
  float* lidar_out = nullptr;  // where this step's lidar observation goes (default: the engine's device buffer)
  // beam pre-filter (may_touch_wall + queue): decided per map bank -- it pays when a good part of the free cells have
  // no wall next to them (rooms), not in corridors one cell wide (mazes); APG_LIDAR_FILTER=0/1 overrides
  int filter = 0;           // the bank is open enough for the filter to pay
  bool filter_fits = false; // every beam fits the swath-pattern table (|dir| <= 5)
  uint4* d_ftable = nullptr;
  double open_fraction = 0.0;
  // procedural-dataset mode (apg_lidar_set_map_generator)
  int gen_kind = 0, gen_max_rooms = 0, gen_door = 0, gen_cells = 0, gen_threads = 0;
  double gen_branch = 1.0;
  unsigned* d_env_maps = nullptr;
  unsigned* d_gen_scratch = nullptr;
  size_t gen_scratch_words = 0;
  int* d_gen_list = nullptr;
  int* d_gen_count = nullptr;
  double* d_qn = nullptr;
};

namespace {

int lidar_launch(apg_lidar_env* e, int mode, uint64_t seed, const float* d_action, const float* d_pred, float* reward2,
                 uint8_t* done2) {
  const apg_lidar_cfg& c = e->cfg;
  LidarParams p{};
  p.N = c.num_envs; p.W = c.width; p.H = c.height; p.words = e->words; p.B = c.beam_count; p.G = e->G;
  p.pad = e->pad; p.pw = e->pw; p.prows = e->prows; p.pmw = e->pmw;
  p.bank_pad = e->gen_kind ? e->d_env_maps_pad : e->d_bank_pad;  // procedural mode: one padded slot per environment
  p.filter = (e->filter && e->filter_fits && e->d_ftable && c.beam_count <= kQueueCap) ? 1 : 0;
  p.fQ = kFilterQ; p.ftable = e->d_ftable;
  p.n0 = e->r0; p.n1 = e->r1 < 0 ? c.num_envs : e->r1;
  p.T = c.step_limit; p.M = (int)e->n_maps; p.static_map = c.static_map; p.static_idx = c.static_map_index;
  p.emit_cells = c.emit_cells; p.sparse = c.sparse_reward ? 1 : 0; p.mode = mode; p.range = c.lidar_range; p.seed = seed; p.env_offset = c.env_offset;
  p.bank = e->d_bank; p.dirs = e->d_dirs; p.action = d_action; p.pred = d_pred;
  p.pos = e->b.d_pos; p.init = e->b.d_initial_pos; p.aliased = e->d_aliased; p.elapsed = e->d_elapsed;
  p.map_idx = e->d_map_idx; p.prev_done = e->d_prev_done; p.g0 = e->d_g0; p.g2 = e->d_g2;
  p.lidar = e->lidar_out ? e->lidar_out : e->b.d_lidar; p.odometry = e->b.d_odometry; p.time_step = e->b.d_time_step; p.map_obs = e->b.d_map;
  p.reward = e->b.d_reward; p.base_reward = e->b.d_base_reward; p.loss = e->b.d_loss; p.target = e->b.d_target;
  p.terminated = e->b.d_terminated; p.truncated = e->b.d_truncated; p.stepped = e->b.d_stepped;
  p.map_idx_out = reinterpret_cast<long long*>(e->b.d_map_idx); p.cells = e->b.d_cells; p.dists = e->b.d_dists;
  p.reward2 = reward2; p.done2 = done2; p.err = e->d_err;
  p.acc1 = e->d_acc1; p.acc2 = e->d_acc2; p.acc_ret = e->d_acc_ret; p.ep_avg1 = e->b.d_ep_avg1; p.ep_final1 = e->b.d_ep_final1;
  p.ep_avg2 = e->b.d_ep_avg2; p.ep_final2 = e->b.d_ep_final2; p.ep_return = e->b.d_ep_return; p.stats = e->d_stats;
  if (e->gen_kind) {
    MapGenParams g{};
    g.N = p.N; g.W = p.W; g.H = p.H; g.words = p.words; g.n0 = p.n0; g.n1 = p.n1; g.mode = mode; g.kind = e->gen_kind;
    g.max_rooms = e->gen_max_rooms; g.door_width = e->gen_door; g.maze_cells = e->gen_cells; g.branching_prob = e->gen_branch;
    g.seed = seed; g.env_offset = c.env_offset; g.prev_done = e->d_prev_done; g.g0 = e->d_g0; g.g2 = e->d_g2;
    g.env_maps_pad = e->d_env_maps_pad; g.pad = e->pad; g.pw = e->pw; g.pmw = e->pmw;
    g.map_idx = e->d_map_idx; g.env_maps = e->d_env_maps; g.qn_table = e->d_qn; g.list = e->d_gen_list;
    g.count = e->d_gen_count; g.scratch = e->d_gen_scratch; g.scratch_words = e->gen_scratch_words; g.err = e->d_err;
    lidar_reset_list_kernel<<<apg::ceil_div(p.n1 - p.n0, 256), 256, 0, e->stream>>>(g);
    lidar_mapgen_kernel<<<e->gen_threads / 64, 64, 0, e->stream>>>(g);
    APG_CUDA(cudaGetLastError());
    apg::count_launch(2);
    p.env_maps = e->d_env_maps;
    p.gen_count = e->d_gen_count;
  }
  const int grid = apg::ceil_div(p.n1 - p.n0, kEnvsPerBlock);
  // (+ one spare word behind the bitmaps and the ray queue of the filtered beam phase)
  const size_t smem = e->smem + sizeof(unsigned) + (p.filter ? kQueueCap * sizeof(unsigned short) : 0);
  lidar_step_kernel<<<grid, kThreads, smem, e->stream>>>(p);
  APG_CUDA(cudaGetLastError());
  apg::count_launch();
  return APG_OK;
}

// device-visible alias of a page-locked host pointer (see apg_image_step_host)
template <typename T>
T* lidar_mapped_alias(T* h) {
  if (!h) return nullptr;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, h) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return (a.type == cudaMemoryTypeHost && a.devicePointer) ? static_cast<T*>(a.devicePointer) : nullptr;
}

}  // namespace

extern "C" {

int apg_lidar_create(const apg_lidar_cfg* cfg, apg_lidar_env** out) {
  APG_CHECK(cfg && out, APG_ERR_INVALID, "apg_lidar_create: null argument");
  *out = nullptr;
  APG_CHECK(cfg->num_envs > 0, APG_ERR_INVALID, "num_envs must be positive");
  APG_CHECK(cfg->width > 0 && cfg->height > 0 && cfg->width <= 4096 && cfg->height <= 4096, APG_ERR_INVALID, "bad map size");
  APG_CHECK(cfg->beam_count > 0 && cfg->beam_count <= 4096, APG_ERR_INVALID, "bad lidar_beam_count");
  APG_CHECK(cfg->lidar_range > 0, APG_ERR_INVALID, "bad lidar_range");
  APG_CHECK(cfg->step_limit > 0, APG_ERR_INVALID, "bad step_limit");
  int ndev = 0;
  cudaError_t de = cudaGetDeviceCount(&ndev);
  if (de != cudaSuccess || ndev <= 0)
    return apg::fail(APG_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                     cudaGetErrorString(de));
  APG_CHECK(cfg->device >= 0 && cfg->device < ndev, APG_ERR_INVALID, "device %d out of range", cfg->device);
  APG_CUDA(cudaSetDevice(cfg->device));
  apg_lidar_env* e = new apg_lidar_env();
  e->cfg = *cfg;
  e->words = (cfg->width + 31) / 32;
  int G = 1;
  while (G < cfg->beam_count && G < 32) G <<= 1;
  if (G < 4) G = 4;
  e->G = G;
  // free border of the shared-memory bitmaps: a ray (<= lidar_range, or <= 1 for the collision rays) that starts inside
  // the map stays within range + 1 cells of it
  e->pad = (int)std::ceil((double)(cfg->lidar_range > 1.0f ? cfg->lidar_range : 1.0f)) + 2;
  if (e->pad < apg_lidar::kColBias) e->pad = apg_lidar::kColBias;  // the 16-bit windows of the beam filter start at x - 7
  if (e->pad > 31) {
    delete e;
    return apg::fail(APG_ERR_INVALID, "lidar_range %g too large (at most 29 cells)", (double)cfg->lidar_range);
  }
  e->pw = (cfg->width + 2 * e->pad + 31) / 32;
  e->prows = cfg->height + 2 * e->pad;
  e->pmw = (e->prows * e->pw + 3) & ~3;
  e->smem = (size_t)kEnvsPerBlock * e->pmw * sizeof(unsigned);
  if (e->smem > 200 * 1024) {
    delete e;
    return apg::fail(APG_ERR_INVALID, "map too large for the shared-memory bitmaps (%zu bytes per block)", e->smem);
  }
  const size_t N = cfg->num_envs, B = cfg->beam_count;
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
  A(apg::dev_alloc(&e->b.d_lidar, N * B));
  A(apg::dev_alloc(&e->b.d_odometry, N * 2));
  A(apg::dev_alloc(&e->b.d_time_step, N));
  if (!cfg->static_map) A(apg::dev_alloc(&e->b.d_map, N * cfg->height * cfg->width));
  A(apg::dev_alloc(&e->b.d_reward, N));
  A(apg::dev_alloc(&e->b.d_base_reward, N));
  A(apg::dev_alloc(&e->b.d_loss, N));
  A(apg::dev_alloc(&e->b.d_target, N * 2));
  A(apg::dev_alloc(&e->b.d_terminated, N));
  A(apg::dev_alloc(&e->b.d_truncated, N));
  A(apg::dev_alloc(&e->b.d_stepped, N));
  A(apg::dev_alloc(&e->b.d_map_idx, N));
  A(apg::dev_alloc(&e->b.d_pos, N * 2));
  A(apg::dev_alloc(&e->b.d_initial_pos, N * 2));
  if (cfg->emit_cells) {
    A(apg::dev_alloc(&e->b.d_cells, N * B * 2));
    A(apg::dev_alloc(&e->b.d_dists, N * B));
  }
  A(apg::dev_alloc(&e->d_acc1, N)); A(apg::dev_alloc(&e->d_acc2, N)); A(apg::dev_alloc(&e->d_acc_ret, N));
  A(apg::dev_alloc(&e->b.d_ep_avg1, N)); A(apg::dev_alloc(&e->b.d_ep_final1, N)); A(apg::dev_alloc(&e->b.d_ep_avg2, N));
  A(apg::dev_alloc(&e->b.d_ep_final2, N)); A(apg::dev_alloc(&e->b.d_ep_return, N));
  A(apg::dev_alloc(&e->d_stats, (size_t)APG_NUM_STATS));
  A(apg::dev_alloc(&e->d_aliased, N));
  A(apg::dev_alloc(&e->d_elapsed, N));
  A(apg::dev_alloc(&e->d_map_idx, N));
  A(apg::dev_alloc(&e->d_prev_done, N));
  A(apg::dev_alloc(&e->d_g0, N));
  A(apg::dev_alloc(&e->d_g2, N));
  A(apg::dev_alloc(&e->d_err, 1));
  A(apg::dev_alloc(&e->d_in_action, N * 2));
  A(apg::dev_alloc(&e->d_in_pred, N * 2));
  A(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  for (auto& ev : e->ev_in) A(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  A(apg::dev_alloc(&e->d_dirs, B * 2));
  if (err == cudaSuccess && e->smem + 8192 > 48 * 1024)
    A(cudaFuncSetAttribute(lidar_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem + 8192));
  if (err != cudaSuccess) {
    apg_lidar_destroy(e);
    return apg::fail(APG_ERR_CUDA, "apg_lidar_create: allocation failed: %s", cudaGetErrorString(err));
  }
  *out = e;
  return APG_OK;
}

void apg_lidar_destroy(apg_lidar_env* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  apg::dev_free(e->d_bank); apg::dev_free(e->d_dirs);
  if (e->d_bank_pad) cudaFree(e->d_bank_pad);
  if (e->d_env_maps_pad) cudaFree(e->d_env_maps_pad);
  if (e->d_ftable) cudaFree(e->d_ftable);
  if (e->d_env_maps) cudaFree(e->d_env_maps);
  if (e->d_gen_scratch) cudaFree(e->d_gen_scratch);
  if (e->d_gen_list) cudaFree(e->d_gen_list);
  if (e->d_gen_count) cudaFree(e->d_gen_count);
  if (e->d_qn) cudaFree(e->d_qn);
  apg::dev_free(e->b.d_lidar); apg::dev_free(e->b.d_odometry); apg::dev_free(e->b.d_time_step); apg::dev_free(e->b.d_map);
  apg::dev_free(e->b.d_reward); apg::dev_free(e->b.d_base_reward); apg::dev_free(e->b.d_loss); apg::dev_free(e->b.d_target);
  apg::dev_free(e->b.d_terminated); apg::dev_free(e->b.d_truncated); apg::dev_free(e->b.d_stepped);
  apg::dev_free(e->b.d_map_idx); apg::dev_free(e->b.d_pos); apg::dev_free(e->b.d_initial_pos);
  apg::dev_free(e->b.d_cells); apg::dev_free(e->b.d_dists);
  apg::dev_free(e->d_acc1); apg::dev_free(e->d_acc2); apg::dev_free(e->d_acc_ret); apg::dev_free(e->d_stats);
  apg::dev_free(e->b.d_ep_avg1); apg::dev_free(e->b.d_ep_final1); apg::dev_free(e->b.d_ep_avg2); apg::dev_free(e->b.d_ep_final2);
  apg::dev_free(e->b.d_ep_return);
  apg::dev_free(e->d_aliased); apg::dev_free(e->d_elapsed); apg::dev_free(e->d_map_idx); apg::dev_free(e->d_prev_done);
  apg::dev_free(e->d_g0); apg::dev_free(e->d_g2); apg::dev_free(e->d_err);
  apg::dev_free(e->d_in_action); apg::dev_free(e->d_in_pred);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  for (auto& ev : e->ev_in) if (ev) cudaEventDestroy(ev);
  delete e;
}

int apg_lidar_set_stream(apg_lidar_env* e, void* cuda_stream) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  e->stream = static_cast<cudaStream_t>(cuda_stream);
  return APG_OK;
}

int apg_lidar_set_map_bank(apg_lidar_env* e, const uint32_t* h_bits, int64_t n_maps) {
  APG_CHECK(e && h_bits && n_maps > 0 && n_maps <= 0xFFFFFFFFll, APG_ERR_INVALID, "apg_lidar_set_map_bank: bad argument");
  APG_CHECK(!e->cfg.static_map || e->cfg.static_map_index < n_maps, APG_ERR_INVALID, "static_map_index out of range");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  apg::dev_free(e->d_bank);
  const size_t count = (size_t)n_maps * e->cfg.height * e->words;
  // every map needs at least one free cell (the reference would raise in integers(0, 0))
  for (int64_t m = 0; m < n_maps; ++m) {
    bool any_free = false;
    for (int y = 0; y < e->cfg.height && !any_free; ++y)
      for (int x = 0; x < e->cfg.width; ++x)
        if (!((h_bits[((size_t)m * e->cfg.height + y) * e->words + (x >> 5)] >> (x & 31)) & 1u)) { any_free = true; break; }
    APG_CHECK(any_free, APG_ERR_INVALID, "map %lld has no free cell", (long long)m);
  }
  // bits past the map width must be zero (they would be phantom walls in the padded shared-memory bitmaps)
  std::vector<uint32_t> clean(h_bits, h_bits + count);
  if (e->cfg.width & 31) {
    const uint32_t last_mask = (1u << (e->cfg.width & 31)) - 1u;
    for (size_t r = 0; r < (size_t)n_maps * e->cfg.height; ++r) clean[r * e->words + e->words - 1] &= last_mask;
  }
  APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_bank), count * sizeof(uint32_t)));
  APG_CUDA(cudaMemcpy(e->d_bank, clean.data(), count * sizeof(uint32_t), cudaMemcpyHostToDevice));
  e->n_maps = n_maps;
  {  // the padded form the step kernel stages with one bulk copy per environment: padded word j of padded row yp holds
     // the map bits [32 j - pad, 32 j - pad + 32) of row yp - pad, rows / bits outside the map are free
    std::vector<uint32_t> padded((size_t)n_maps * e->pmw, 0u);
    const int pad = e->pad;
    for (int64_t m = 0; m < n_maps; ++m)
      for (int y = 0; y < e->cfg.height; ++y) {
        const uint32_t* row = clean.data() + ((size_t)m * e->cfg.height + y) * e->words;
        uint32_t* dst = padded.data() + (size_t)m * e->pmw + (size_t)(y + pad) * e->pw;
        uint32_t lo = 0u;
        for (int j = 0; j < e->pw; ++j) {
          const uint32_t hi = j < e->words ? row[j] : 0u;
          dst[j] = (hi << pad) | (lo >> (32 - pad));
          lo = hi;
        }
      }
    if (e->d_bank_pad) { cudaFree(e->d_bank_pad); e->d_bank_pad = nullptr; }
    APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_bank_pad), padded.size() * sizeof(uint32_t)));
    APG_CUDA(cudaMemcpy(e->d_bank_pad, padded.data(), padded.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  // share of the free cells whose eight neighbours are free as well (first <= 256 maps): the beam filter's payoff
  {
    const int W = e->cfg.width, H = e->cfg.height;
    auto wall = [&](int64_t m, int x, int y) -> bool {
      if (x < 0 || y < 0 || x >= W || y >= H) return false;
      return (clean[((size_t)m * H + y) * e->words + (x >> 5)] >> (x & 31)) & 1u;
    };
    long long n_free = 0, n_open = 0;
    for (int64_t m = 0; m < n_maps && m < 256; ++m)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          if (wall(m, x, y)) continue;
          n_free++;
          bool open = true;
          for (int dy = -1; dy <= 1 && open; ++dy)
            for (int dx = -1; dx <= 1; ++dx)
              if (wall(m, x + dx, y + dy)) { open = false; break; }
          n_open += open ? 1 : 0;
        }
    e->open_fraction = n_free ? (double)n_open / (double)n_free : 0.0;
    e->filter = e->open_fraction >= 0.3 ? 1 : 0;
    if (const char* ov = getenv("APG_LIDAR_FILTER")) e->filter = atoi(ov) != 0;
  }
  return APG_OK;
}

int apg_lidar_set_beams(apg_lidar_env* e, const float* h_dirs) {
  APG_CHECK(e && h_dirs, APG_ERR_INVALID, "apg_lidar_set_beams: null argument");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaMemcpy(e->d_dirs, h_dirs, (size_t)e->cfg.beam_count * 2 * sizeof(float), cudaMemcpyHostToDevice));
  e->have_beams = true;
  // swath patterns of the beam filter (apg_lidar_filter.cuh): one 16-byte entry per (beam, sub-cell of the start cell)
  const int B = e->cfg.beam_count;
  e->filter_fits = true;
  for (int b = 0; b < B; ++b) e->filter_fits = e->filter_fits && apg_lidar::beam_pattern_fits(h_dirs[2 * b], h_dirs[2 * b + 1]);
  if (e->filter_fits) {
    static_assert(sizeof(apg_lidar::BeamPattern) == sizeof(uint4), "one uint4 per entry");
    std::vector<apg_lidar::BeamPattern> t((size_t)B * kFilterQ * kFilterQ);
    for (int b = 0; b < B; ++b)
      for (int qy = 0; qy < kFilterQ; ++qy)
        for (int qx = 0; qx < kFilterQ; ++qx)
          t[((size_t)b * kFilterQ + qy) * kFilterQ + qx] =
              apg_lidar::make_beam_pattern((double)h_dirs[2 * b], (double)h_dirs[2 * b + 1], qx, qy, kFilterQ);
    if (!e->d_ftable) APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_ftable), t.size() * sizeof(uint4)));
    APG_CUDA(cudaMemcpy(e->d_ftable, t.data(), t.size() * sizeof(uint4), cudaMemcpyHostToDevice));
  }
  return APG_OK;
}

int apg_lidar_set_map_generator(apg_lidar_env* e, int32_t kind, int32_t max_rooms, int32_t door_width,
                                double branching_prob) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CHECK(kind == APG_MAPGEN_ROOMS || kind == APG_MAPGEN_MAZE, APG_ERR_INVALID, "unknown map generator %d", kind);
  APG_CHECK(!e->cfg.static_map, APG_ERR_INVALID, "static_map takes a one-map bank (generate it with apg_maps_generate_*)");
  const int W = e->cfg.width, H = e->cfg.height;
  if (kind == APG_MAPGEN_ROOMS) {
    APG_CHECK(W == H, APG_ERR_INVALID, "rooms maps must be square (the reference transposes half of them)");
    APG_CHECK(W >= 4 && W <= 255, APG_ERR_INVALID, "bad map size");
    APG_CHECK(max_rooms >= 1 && max_rooms <= apg_rooms::kMaxRooms, APG_ERR_INVALID, "max_rooms must be in [1, %d]",
              apg_rooms::kMaxRooms);
    APG_CHECK(door_width >= 1 && door_width <= 64, APG_ERR_INVALID, "bad door_width");
  } else {
    APG_CHECK(W >= 3 && H >= 3 && W <= 255 && H <= 255 && (W & 1) && (H & 1), APG_ERR_INVALID,
              "Width and height must be odd (and at most 255).");
  }
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  const size_t N = e->cfg.num_envs, mw = (size_t)H * e->words;
  int sms = 0;
  APG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->cfg.device));
  e->gen_kind = kind; e->gen_max_rooms = max_rooms; e->gen_door = door_width; e->gen_branch = branching_prob;
  e->gen_cells = ((W - 1) / 2) * ((H - 1) / 2) + 1;
  e->gen_threads = sms * 8 * 64;  // 16 warps per SM: the generators are long serial chains, latency-bound
  e->gen_scratch_words = kind == APG_MAPGEN_ROOMS
                             ? 2 * mw + (apg_rooms::kMaxStack * sizeof(apg_rooms::View) + 3) / 4
                             : (size_t)e->gen_cells;
  if (e->d_env_maps) { cudaFree(e->d_env_maps); e->d_env_maps = nullptr; }
  if (e->d_env_maps_pad) { cudaFree(e->d_env_maps_pad); e->d_env_maps_pad = nullptr; }
  APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_env_maps_pad), N * (size_t)e->pmw * sizeof(unsigned)));
  APG_CUDA(cudaMemset(e->d_env_maps_pad, 0, N * (size_t)e->pmw * sizeof(unsigned)));
  if (e->d_gen_scratch) { cudaFree(e->d_gen_scratch); e->d_gen_scratch = nullptr; }
  APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_env_maps), N * mw * sizeof(unsigned)));
  APG_CUDA(cudaMemset(e->d_env_maps, 0, N * mw * sizeof(unsigned)));
  APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_gen_scratch), (size_t)e->gen_threads * e->gen_scratch_words * sizeof(unsigned)));
  if (!e->d_gen_list) APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_gen_list), N * sizeof(int)));
  if (!e->d_gen_count) APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_gen_count), sizeof(int)));
  APG_CUDA(cudaMemset(e->d_gen_count, 0, sizeof(int)));
  // exp(n * log(1 - p)) of numpy's binomial inversion, from the host's libm (the same one numpy calls)
  double qn[apg_rooms::kMaxRooms + 1];
  for (int k = 0; k <= apg_rooms::kMaxRooms; ++k) qn[k] = std::exp((double)k * std::log(1.0 - apg_rooms::kP));
  if (!e->d_qn) APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_qn), sizeof(qn)));
  APG_CUDA(cudaMemcpy(e->d_qn, qn, sizeof(qn), cudaMemcpyHostToDevice));
  // beam filter: rooms maps are open, maze corridors are one cell wide (see apg_lidar_set_map_bank)
  e->filter = kind == APG_MAPGEN_ROOMS ? 1 : 0;
  if (const char* ov = getenv("APG_LIDAR_FILTER")) e->filter = atoi(ov) != 0;
  e->n_maps = 0;
  e->has_reset = false;
  return APG_OK;
}

int apg_lidar_reset(apg_lidar_env* e, uint64_t seed) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CHECK((e->d_bank || e->gen_kind) && e->have_beams, APG_ERR_STATE,
            "apg_lidar_reset before set_map_bank / set_map_generator / set_beams");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  int rc = lidar_launch(e, 2, seed, nullptr, nullptr, nullptr, nullptr);
  if (rc != APG_OK) return rc;
  e->has_reset = true;
  return APG_OK;
}

int apg_lidar_reset_all(apg_lidar_env* e) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "the first reset must be seeded (the reference has no data loader before that either)");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  return lidar_launch(e, 3, 0, nullptr, nullptr, nullptr, nullptr);
}

int apg_lidar_step(apg_lidar_env* e, const float* d_action, const float* d_pred) {
  APG_CHECK(e && d_action && d_pred, APG_ERR_INVALID, "apg_lidar_step: null argument");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "step before reset (the first reset must be seeded)");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  return lidar_launch(e, 0, 0, d_action, d_pred, nullptr, nullptr);
}

int apg_lidar_rollout(apg_lidar_env* e, const float* d_actions, const float* d_preds, int32_t n_steps, float* d_rewards,
                      uint8_t* d_dones) {
  APG_CHECK(e && d_actions && d_preds && n_steps >= 0, APG_ERR_INVALID, "apg_lidar_rollout: bad argument");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "rollout before reset");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  const size_t N = e->cfg.num_envs;
  for (int k = 0; k < n_steps; ++k) {
    int rc = lidar_launch(e, 0, 0, d_actions + (size_t)k * N * 2, d_preds + (size_t)k * N * 2,
                          d_rewards ? d_rewards + (size_t)k * N : nullptr, d_dones ? d_dones + (size_t)k * N : nullptr);
    if (rc != APG_OK) return rc;
  }
  return APG_OK;
}

int apg_lidar_step_host(apg_lidar_env* e, const float* h_action, const float* h_pred, float* h_reward,
                        uint8_t* h_terminated, uint8_t* h_truncated, float* h_lidar) {
  APG_CHECK(e && h_action && h_pred, APG_ERR_INVALID, "apg_lidar_step_host: null argument");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "step before reset");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  const size_t N = e->cfg.num_envs, B = e->cfg.beam_count;
  // same scheme as apg_image_step_host: chunks whose host-to-device copies (copy stream) overlap the kernels of the
  // previous chunk; page-locked result buffers are written by the kernel directly
  float* m_reward = lidar_mapped_alias(h_reward);
  uint8_t* m_term = lidar_mapped_alias(h_terminated);
  float* m_lidar = lidar_mapped_alias(h_lidar);
  int chunks = N >= 8192 ? 2 : 1;  // measured at N = 262144 (32x32 maps, 64 beams): 1 / 2 / 4 chunks = 569 / 574 / 615 us
  if (const char* ov = getenv("APG_HOST_CHUNKS")) {  // developer switch: number of pipeline chunks of the host step
    const int v = atoi(ov);
    if (v >= 1 && v <= 8) chunks = v;
  }
  const size_t per = ((N + chunks - 1) / chunks + 1023) & ~(size_t)1023;
  e->lidar_out = m_lidar;
  int rc = APG_OK;
  for (int c = 0; c < chunks && rc == APG_OK; ++c) {
    const size_t lo = (size_t)c * per, hi = lo + per < N ? lo + per : N;
    if (lo >= hi) break;
    cudaError_t ce = cudaMemcpyAsync(e->d_in_action + lo * 2, h_action + lo * 2, (hi - lo) * 2 * sizeof(float),
                                     cudaMemcpyHostToDevice, e->copy_stream);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(e->d_in_pred + lo * 2, h_pred + lo * 2, (hi - lo) * 2 * sizeof(float), cudaMemcpyHostToDevice,
                           e->copy_stream);
    if (ce == cudaSuccess) ce = cudaEventRecord(e->ev_in[c], e->copy_stream);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->stream, e->ev_in[c], 0);
    if (ce != cudaSuccess) {
      rc = apg::fail(APG_ERR_CUDA, "apg_lidar_step_host: %s", cudaGetErrorString(ce));
      break;
    }
    e->r0 = (int)lo; e->r1 = (int)hi;
    rc = lidar_launch(e, 0, 0, e->d_in_action, e->d_in_pred, m_reward, m_term);
  }
  e->r0 = 0; e->r1 = -1;
  e->lidar_out = nullptr;
  if (rc != APG_OK) {
    cudaStreamSynchronize(e->copy_stream);
    cudaStreamSynchronize(e->stream);
    return rc;
  }
  if (h_reward && !m_reward) APG_CUDA(cudaMemcpyAsync(h_reward, e->b.d_reward, N * 4, cudaMemcpyDeviceToHost, e->stream));
  if (h_terminated && !m_term) APG_CUDA(cudaMemcpyAsync(h_terminated, e->b.d_terminated, N, cudaMemcpyDeviceToHost, e->stream));
  if (h_lidar && !m_lidar) APG_CUDA(cudaMemcpyAsync(h_lidar, e->b.d_lidar, N * B * 4, cudaMemcpyDeviceToHost, e->stream));
  if (h_truncated) memset(h_truncated, 0, N);  // TimeLimit(issue_termination=True) never truncates
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

int apg_lidar_get_buffers(apg_lidar_env* e, apg_lidar_buffers* out) {
  APG_CHECK(e && out, APG_ERR_INVALID, "null argument");
  *out = e->b;
  return APG_OK;
}

int apg_lidar_get_stats(apg_lidar_env* e, double* h_stats, int clear) {
  APG_CHECK(e && h_stats, APG_ERR_INVALID, "null argument");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaMemcpyAsync(h_stats, e->d_stats, APG_NUM_STATS * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (clear) APG_CUDA(cudaMemsetAsync(e->d_stats, 0, APG_NUM_STATS * sizeof(double), e->stream));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

int apg_lidar_poll_error(apg_lidar_env* e, uint32_t* flags, int clear) {
  APG_CHECK(e && flags, APG_ERR_INVALID, "null argument");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaMemcpyAsync(flags, e->d_err, 4, cudaMemcpyDeviceToHost, e->stream));
  if (clear) APG_CUDA(cudaMemsetAsync(e->d_err, 0, 4, e->stream));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

int apg_lidar_sync(apg_lidar_env* e) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

// ---- host-side test hook for the numpy-compatible random streams (apg_rng.h) -----------------------------
// op: 0 next64, 1 next32, 2 integers(0, arg), 3 integers(0, 2**32, endpoint=True), 4 re-seed with SeedSequence(arg),
//     5 uniform(-1, 1) (result = the float64 bit pattern), 6 advance by arg raw draws
int apg_rng_host_stream(uint64_t seed, int32_t n_ops, const int32_t* op, const uint64_t* arg, uint64_t* out) {
  APG_CHECK(n_ops >= 0 && (n_ops == 0 || (op && arg && out)), APG_ERR_INVALID, "apg_rng_host_stream: bad argument");
  apg::Pcg64 g;
  apg::pcg64_seed(g, seed);
  for (int i = 0; i < n_ops; ++i) {
    switch (op[i]) {
      case 0: out[i] = apg::pcg64_next64(g); break;
      case 1: out[i] = apg::pcg64_next32(g); break;
      case 2: out[i] = apg::rng_integers_below(g, arg[i]); break;
      case 3: out[i] = apg::rng_integers_0_to_2p32_inclusive(g); break;
      case 4: apg::pcg64_seed(g, arg[i]); out[i] = 0; break;
      case 5: { const double u = apg::rng_uniform(g, -1.0, 2.0); memcpy(&out[i], &u, 8); break; }
      case 6: apg::pcg64_advance(g, arg[i]); out[i] = 0; break;
      default: return apg::fail(APG_ERR_INVALID, "bad op %d", op[i]);
    }
  }
  return APG_OK;
}

}  // extern "C"
