// Fused image-environment step for sm_100a.
//
// One launch performs, for every environment of the shard (reference file:line in /root/reference/ap_gym):
//   * project_sphere + float64 glimpse position update + clip         util.py:94-97, envs/image/image_perception_module.py:206-208
//   * base reward -||a||*1e-3                                          envs/image/image_perception_module.py:209
//   * prediction quality + overlay stamp (pre-move position)          envs/image_classification.py:114-116, envs/image_localization.py:157-159,
//                                                                      envs/image/image_perception_module.py:219-234
//   * bilinear resample of the sensor window from the image bank      envs/image/image_perception_module.py:294-331 (scipy RGI linear, f64 blend)
//   * normalised cross-entropy / MSE loss, reward = base - loss       loss_fn.py:100-110,212-220,261-267, active_perception_vector_env.py:84-111
//   * terminated / truncated, time_step obs, NEXT_STEP all-or-nothing autoreset from a queued payload
//                                                                      envs/image/image_perception_module.py:197-217,236-251
//   * visited bitmask (image resolution) update
//
// Numerics: the file is compiled with -fmad=false, every float op below is a separately rounded IEEE op in
// the order numpy performs it (see oracle/image_oracle.py).  uint8 banks are converted with the reference's
// f32(u8)/255, reproduced exactly by a multiply + two explicit fma corrections (px_to_f64).
//
// Thread mapping: a block owns E environments; thread = environment for the scalar work and for issuing the
// asynchronous (cp.async) staging of the glimpse's bank footprint into a shared-memory tile, then warp = environment
// and lane = sensor element for the blend (see image_step_kernel).
#include <cuda_pipeline.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "apg_common.h"
#include "apg_payload.h"
#include "apg_rng.h"

#include "apg_image_dev.cuh"
#include "apg_image_small.h"
#include "apg_image_tma.h"

using namespace apg_img;

namespace {

template <typename BankT, int CT, bool VREUSE>
__device__ __forceinline__ void blend_column_f32(const BankT* __restrict__ colp, const Tap* __restrict__ tp, int tap_step,
                                                 int pitch_e, int c_rt, int n_rows, float wx, float omx,
                                                 float* __restrict__ outp, int out_step, bool store_ok) {
  const int C = CT ? CT : c_rt;
  int prev_k = -4;
  float b0 = 0.0f, b1 = 0.0f;
  for (int i = 0; i < n_rows; ++i, tp += tap_step, outp += out_step) {
    const Tap rt = *tp;
    const BankT* rp = colp + rt.k * pitch_e;
    float p00 = b0, p01 = b1;
    if (!VREUSE || rt.k != prev_k + 1) {
      p00 = px_to_f32(rp[0]);
      p01 = px_to_f32(rp[C]);
    }
    rp += pitch_e;
    const float p10 = px_to_f32(rp[0]), p11 = px_to_f32(rp[C]);
    b0 = p10;
    b1 = p11;
    prev_k = rt.k;
    const float wy = rt.wf, omy = rt.omwf;
    float acc = __fmul_rn(p00, __fmul_rn(omy, omx));
    acc = __fmaf_rn(p01, __fmul_rn(omy, wx), acc);
    acc = __fmaf_rn(p10, __fmul_rn(wy, omx), acc);
    acc = __fmaf_rn(p11, __fmul_rn(wy, wx), acc);
    if (store_ok) *outp = __saturatef(acc);
  }
}

// One lane's column of the sensor window (warp-per-environment path): walks the sensor rows i_first, i_first +
// i_step, ... of one (column, channel), blending from the staged tile.  With one sensor row per pass (VREUSE) the
// bottom taps of row i are the top taps of row i+1 unless the taps are not consecutive (clamped border / scaled
// sensor), which is uniform across the warp.  CT > 0 fixes the channel count at compile time.
template <typename BankT, int CT, bool VREUSE>
__device__ __forceinline__ void blend_column(const BankT* __restrict__ colp, const Tap* __restrict__ tp, int tap_step,
                                             int pitch_e, int c_rt, int n_rows, double wx, double omx,
                                             float* __restrict__ outp, int out_step, bool store_ok) {
  const int C = CT ? CT : c_rt;
  int prev_k = -4;
  double b0 = 0.0, b1 = 0.0;
  for (int i = 0; i < n_rows; ++i, tp += tap_step, outp += out_step) {
    const Tap rt = *tp;
    const BankT* rp = colp + rt.k * pitch_e;
    double p00 = b0, p01 = b1;
    if (!VREUSE || rt.k != prev_k + 1) {
      p00 = px_to_f64(rp[0]);
      p01 = px_to_f64(rp[C]);
    }
    rp += pitch_e;
    const double p10 = px_to_f64(rp[0]), p11 = px_to_f64(rp[C]);
    b0 = p10;
    b1 = p11;
    prev_k = rt.k;
    const double wy = rt.w, omy = 1.0 - rt.w;
    double acc = p00 * (omy * omx);  // the reference's leading `0.0 +` is an exact no-op for products >= +0
    acc = acc + p01 * (omy * wx);
    acc = acc + p10 * (wy * omx);
    acc = acc + p11 * (wy * wx);
    // uint8 pixels are in [0,1] and the weights sum to 1 (+ a few ulp): float32(acc) is already clipped
    const float v = (sizeof(BankT) == 1) ? (float)acc : (float)fmin(fmax(acc, 0.0), 1.0);
    if (store_ok) *outp = v;
  }
}

// Taps of one glimpse (-> shared memory) and the asynchronous staging of its bank footprint (thread = environment).
template <typename BankT>
__device__ __forceinline__ void stage_env(const ImgParams& p, int el, EnvHdr* s_hdr, Tap* s_taps, unsigned char* s_tile,
                                          double ccx, double ccy, int idx, unsigned& err) {
  constexpr int ES = (int)sizeof(BankT);
  const int row_bytes = p.bank_pitch;
  Tap* rows = s_taps + el;                 // tap k of this env: rows[k * E]
  Tap* cols = s_taps + (size_t)p.S0 * p.E + el;
  int rmin = 0, rmax = 0, cmin = 0, cmax = 0;
  unsigned cb0 = 0u, cb1 = 0u, nonconsec = 0u;
  for (int i = 0; i < p.S0; ++i) {
    const double y = ccy + ((double)i - p.half0) * p.scale;
    if (y < -p.offy || y > p.offy) err |= APG_FLAG_OUT_OF_GRID;
    int k;
    double w;
    tap_row(y, p.offy, p.H, ES == 1, k, w);
    if (i == 0) rmin = k;
    if (k - rmin != i) nonconsec = kNonConsec;
    rmax = min(k, p.H - 2) + 1;
    if (p.fast) { rows[i * p.E].wf = (float)w; rows[i * p.E].omwf = (float)(1.0 - w); }
    else rows[i * p.E].w = w;
    rows[i * p.E].k = k - rmin;
  }
  for (int j = 0; j < p.S1; ++j) {
    const double x = ccx + ((double)j - p.half1) * p.scale;
    if (x < -p.offx || x > p.offx) err |= APG_FLAG_OUT_OF_GRID;
    int k;
    double w;
    tap(x, p.offx, p.W, k, w);
    if (j == 0) cmin = k;
    cmax = k + 1;
    if (p.fast) { cols[j * p.E].wf = (float)w; cols[j * p.E].omwf = (float)(1.0 - w); }
    else cols[j * p.E].w = w;
    cols[j * p.E].k = k;
    if (k < 32) cb0 |= 1u << k; else if (k < 64) cb1 |= 1u << (k - 32);
    if (k + 1 < 32) cb0 |= 1u << (k + 1); else if (k + 1 < 64) cb1 |= 1u << (k - 31);
  }
  const int nrows = min(rmax - rmin + 1, p.tile_rows);
  const int first_byte = cmin * p.C * ES;
  const int lead = first_byte & 15;
  const int start = first_byte - lead;                               // aligned down to the 16-byte copy granule
  int span = (cmax + 1) * p.C * ES - start;                          // bytes needed from `start`
  span = min((span + 15) & ~15, min(row_bytes - start, p.pitch));
  EnvHdr h;
  h.lead = (unsigned short)(lead | nonconsec); h.cmin = (unsigned short)cmin; h.rmin = (unsigned short)rmin;
  h.nrows = (unsigned short)nrows;
  h.colbits[0] = cb0; h.colbits[1] = cb1;
  s_hdr[el] = h;
  const unsigned char* src = reinterpret_cast<const unsigned char*>(p.bank) + (long long)idx * p.H * row_bytes +
                             (long long)rmin * row_bytes + start;
  unsigned char* dst = s_tile + (size_t)el * p.tile_bytes;
  for (int r = 0; r < nrows; ++r) {
    for (int b = 0; b < span; b += 16) __pipeline_memcpy_async(dst + b, src + b, 16);
    src += row_bytes;
    dst += p.pitch;
  }
}

// ---- consecutive tap rows (the common case: sensor_scale == 1 away from rounding accidents) ---------------------------
// Row tap i reads tile rows i and i+1, so the bottom pixels of one sensor row are the top pixels of the next and each
// tile row is loaded and converted exactly once, with no tap-index compare in the loop.  `colp` points at tile row 0.
template <typename BankT, int CT>
__device__ __forceinline__ void blend_column_consec(const BankT* __restrict__ colp, const Tap* __restrict__ tp,
                                                    int tap_step, int pitch_e, int c_rt, int n_rows, double wx, double omx,
                                                    float* __restrict__ out, int out_step, bool store_ok) {
  const int C = CT ? CT : c_rt;
  double p00 = px_to_f64(colp[0]), p01 = px_to_f64(colp[C]);
  int o = 0;
  for (int i = 0; i < n_rows; ++i, tp += tap_step, o += out_step) {
    colp += pitch_e;
    const double p10 = px_to_f64(colp[0]), p11 = px_to_f64(colp[C]);
    const double wy = tp->w, omy = 1.0 - wy;
    double acc = p00 * (omy * omx);  // same four terms, same order as blend_column
    acc = acc + p01 * (omy * wx);
    acc = acc + p10 * (wy * omx);
    acc = acc + p11 * (wy * wx);
    const float v = (sizeof(BankT) == 1) ? (float)acc : (float)fmin(fmax(acc, 0.0), 1.0);
    if (store_ok) out[o] = v;
    p00 = p10;
    p01 = p11;
  }
}

// float32 blend, consecutive rows: separable form -- one horizontal blend per tile row, one vertical blend per sensor
// row; the uint8 scale 1/255 is folded into the column weights.  Error: a few float32 ulps of a non-negative sum.
__device__ __forceinline__ float px_raw_f32(unsigned char u) { return (float)u; }
__device__ __forceinline__ float px_raw_f32(float v) { return v; }

template <typename BankT, int CT>
__device__ __forceinline__ void blend_column_consec_f32(const BankT* __restrict__ colp, const Tap* __restrict__ tp,
                                                        int tap_step, int pitch_e, int c_rt, int n_rows, float wx,
                                                        float omx, float* __restrict__ out, int out_step, bool store_ok) {
  const int C = CT ? CT : c_rt;
  const float unit = (sizeof(BankT) == 1) ? 0.003921568859368563f : 1.0f;
  const float sx0 = __fmul_rn(omx, unit), sx1 = __fmul_rn(wx, unit);
  float hp = __fmaf_rn(px_raw_f32(colp[C]), sx1, __fmul_rn(px_raw_f32(colp[0]), sx0));
  int o = 0;
  for (int i = 0; i < n_rows; ++i, tp += tap_step, o += out_step) {
    colp += pitch_e;
    const float h = __fmaf_rn(px_raw_f32(colp[C]), sx1, __fmul_rn(px_raw_f32(colp[0]), sx0));
    const float2 wq = *reinterpret_cast<const float2*>(tp);  // (w, 1 - w)
    const float v = __saturatef(__fmaf_rn(h, wq.x, __fmul_rn(hp, wq.y)));
    if (store_ok) out[o] = v;
    hp = h;
  }
}

// ---- split phase A (warp-per-environment path) ------------------------------------------------------------------------
// The serial per-environment part of phase A (40 taps + ~35 copies for a 10x10x3 glimpse) kept one warp busy for ~40 % of
// a block's lifetime while the other warps waited at the barrier (profiles/r01_final_tinyloc.txt: 35 % of all stall
// samples).  Split: the environment's thread only does the scalars and the tile header (first / last tap per axis);
// after a barrier ALL threads compute the taps (one tap per work item) and issue the copies (one 16-byte chunk per item,
// consecutive threads -> consecutive chunks of a row).

template <typename BankT>
__device__ __forceinline__ void stage_head(const ImgParams& p, int el, EnvHdr* s_hdr, EnvCtr* s_ctr, double ccx, double ccy,
                                           int idx, unsigned& err) {
  constexpr int ES = (int)sizeof(BankT);
  // the tap coordinates are monotone in the tap number: the first and the last one decide the footprint and the
  // out-of-grid flag
  const double y0 = ccy + (0.0 - p.half0) * p.scale, y1 = ccy + ((double)(p.S0 - 1) - p.half0) * p.scale;
  const double x0 = ccx + (0.0 - p.half1) * p.scale, x1 = ccx + ((double)(p.S1 - 1) - p.half1) * p.scale;
  if (y0 < -p.offy || y0 > p.offy || y1 < -p.offy || y1 > p.offy || x0 < -p.offx || x0 > p.offx || x1 < -p.offx ||
      x1 > p.offx)
    err |= APG_FLAG_OUT_OF_GRID;
  int rmin, rlast, cmin, clast;
  double w;
  tap(y0, p.offy, p.H, rmin, w);
  tap(y1, p.offy, p.H, rlast, w);
  tap(x0, p.offx, p.W, cmin, w);
  tap(x1, p.offx, p.W, clast, w);
  const int nrows = min(rlast + 1 - rmin + 1, p.tile_rows);
  const int first_byte = cmin * p.C * ES;
  const int lead = first_byte & 15;
  const int start = first_byte - lead;
  int span = (clast + 2) * p.C * ES - start;
  span = min((span + 15) & ~15, min(p.bank_pitch - start, p.pitch));
  EnvHdr h;
  h.lead = (unsigned short)lead; h.cmin = (unsigned short)cmin; h.rmin = (unsigned short)rmin;
  h.nrows = (unsigned short)nrows;
  h.colbits[0] = 0u; h.colbits[1] = 0u;  // OR-ed together by the column-tap work items
  s_hdr[el] = h;
  EnvCtr c;
  c.ccx = ccx; c.ccy = ccy; c.span = span; c.nrows = nrows;
  c.src0 = ((long long)idx * p.H + rmin) * p.bank_pitch + start;
  s_ctr[el] = c;
}

template <typename BankT>
__device__ __forceinline__ void stage_block(const ImgParams& p, int nE, int tid, int nthreads, EnvHdr* s_hdr,
                                            const EnvCtr* s_ctr, Tap* s_taps, unsigned char* s_tile) {
  constexpr int ES = (int)sizeof(BankT);
  // copies first (longest latency): item = (environment, tile row, 16-byte chunk)
  // (integer divisions by block-uniform divisors: trunc((q + 0.5) * float(1 / d)) == q / d for every q < 2^16 and
  //  d <= 13000 -- checked exhaustively; a block holds at most 200 KB / 16 B = 12800 chunks or taps)
  const int nch = p.pitch >> 4;
  const int per_env = p.tile_rows * nch;
  const float inv_per_env = __frcp_rn((float)per_env), inv_nch = __frcp_rn((float)nch), inv_E = __frcp_rn((float)p.E);
  for (int q = tid; q < nE * per_env; q += nthreads) {
    const int el = (int)(((float)q + 0.5f) * inv_per_env);
    const int rem = q - el * per_env;
    const int r = (int)(((float)rem + 0.5f) * inv_nch);
    const int b = (rem - r * nch) << 4;
    const longlong2 cp = reinterpret_cast<const longlong2*>(s_ctr + el)[1];  // (src0, span | nrows << 32)
    const int span = (int)(cp.y & 0xffffffffll), nrows = (int)(cp.y >> 32);
    if (r < nrows && b < span) {
      const unsigned char* src = reinterpret_cast<const unsigned char*>(p.bank) + cp.x + r * p.bank_pitch + b;
      __pipeline_memcpy_async(s_tile + el * p.tile_bytes + r * p.pitch + b, src, 16);
    }
  }
  // taps: item = (tap, environment), environment fastest (conflict-free shared-memory stores)
  const int n_taps = p.S0 + p.S1;
  for (int q = tid; q < n_taps * p.E; q += nthreads) {
    const int t = (int)(((float)q + 0.5f) * inv_E);
    const int el = q - t * p.E;
    if (el >= nE) continue;
    const double2 c = reinterpret_cast<const double2*>(s_ctr + el)[0];  // (ccx, ccy)
    int k;
    double w;
    if (t < p.S0) {
      tap_row(c.y + ((double)t - p.half0) * p.scale, p.offy, p.H, ES == 1, k, w);
      k -= (int)s_hdr[el].rmin;
      if (k != t) atomicOr(reinterpret_cast<unsigned*>(&s_hdr[el]), kNonConsec);  // `lead` is the low half of word 0
    } else {
      tap(c.x + ((double)(t - p.S0) - p.half1) * p.scale, p.offx, p.W, k, w);
      unsigned b0 = 0u, b1 = 0u;
      if (k < 32) b0 |= 1u << k; else if (k < 64) b1 |= 1u << (k - 32);
      if (k + 1 < 32) b0 |= 1u << (k + 1); else if (k + 1 < 64) b1 |= 1u << (k - 31);
      if (b0) atomicOr(&s_hdr[el].colbits[0], b0);
      if (b1) atomicOr(&s_hdr[el].colbits[1], b1);
    }
    Tap* tp = s_taps + (size_t)t * p.E + el;
    if (p.fast) { tp->wf = (float)w; tp->omwf = (float)(1.0 - w); }
    else tp->w = w;
    tp->k = k;
  }
}

struct WarpLane {  // lane constants of the warp-per-environment blend
  int G, R, n_chunks, r_sub, jc0, pitch_e, first, n_rows, kind;
};

template <typename BankT>
__device__ __forceinline__ WarpLane warp_lane(const ImgParams& p, int lane) {
  WarpLane L;
  L.G = p.S1 * p.C;
  const int cpp = min(L.G, 32);
  L.R = (L.G <= 32) ? (32 / L.G) : 1;
  L.n_chunks = (L.G + 31) >> 5;
  L.r_sub = lane / cpp;
  L.jc0 = lane - L.r_sub * cpp;
  L.pitch_e = p.pitch / (int)sizeof(BankT);
  L.first = min(L.r_sub, p.S0 - 1);
  L.n_rows = (p.S0 - L.first + L.R - 1) / L.R;
  L.kind = (L.R == 1) ? ((p.C == 3) ? 3 : ((p.C == 1) ? 1 : 0)) : -1;
  return L;
}

// Rare path of the warp-per-environment blend (non-consecutive tap rows, or several sensor rows per pass): kept out of
// line so that the geometry-specialised kernels do not unroll it.
template <typename BankT>
__device__ __noinline__ void blend_column_any(int fast, int kind, const BankT* colp, const Tap* tp, int E, int pitch_e, int C,
                                              int R, int G, int S0, int r_sub, int n_rows, Tap ct, float* outp, bool col_ok) {
  if (fast) {
    const float wxf = ct.wf, omxf = ct.omwf;
    if (kind == 3)
      blend_column_f32<BankT, 3, true>(colp, tp, E, pitch_e, 3, n_rows, wxf, omxf, outp, G, col_ok);
    else if (kind == 1)
      blend_column_f32<BankT, 1, true>(colp, tp, E, pitch_e, 1, n_rows, wxf, omxf, outp, G, col_ok);
    else if (kind == 0)
      blend_column_f32<BankT, 0, true>(colp, tp, E, pitch_e, C, n_rows, wxf, omxf, outp, G, col_ok);
    else
      blend_column_f32<BankT, 0, false>(colp, tp, R * E, pitch_e, C, n_rows, wxf, omxf, outp, R * G, col_ok && r_sub < S0);
  } else {
    const double wx = ct.w, omx = 1.0 - ct.w;
    if (kind == 3)
      blend_column<BankT, 3, true>(colp, tp, E, pitch_e, 3, n_rows, wx, omx, outp, G, col_ok);
    else if (kind == 1)
      blend_column<BankT, 1, true>(colp, tp, E, pitch_e, 1, n_rows, wx, omx, outp, G, col_ok);
    else if (kind == 0)
      blend_column<BankT, 0, true>(colp, tp, E, pitch_e, C, n_rows, wx, omx, outp, G, col_ok);
    else
      blend_column<BankT, 0, false>(colp, tp, R * E, pitch_e, C, n_rows, wx, omx, outp, R * G, col_ok && r_sub < S0);
  }
}

// Blend of one environment's glimpse by one warp from its staged tile, plus its visited-mask update (set 0).
template <typename BankT>
__device__ __forceinline__ void blend_env_warp(const ImgParams& p, const WarpLane& L, int lane, int el, int n, int set,
                                               const EnvHdr* s_hdr, const Tap* s_taps, const unsigned char* s_tile,
                                               float* out) {
  const int G = L.G, R = L.R, n_chunks = L.n_chunks, r_sub = L.r_sub, jc0 = L.jc0, pitch_e = L.pitch_e;
  const int first = L.first, n_rows = L.n_rows, kind = L.kind;
  const EnvHdr h = s_hdr[el];
  const Tap* rows = s_taps + el;
  const Tap* cols = s_taps + p.S0 * p.E + el;
  const unsigned char* tile = s_tile + el * p.tile_bytes;
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int jc = (chunk << 5) + jc0;
    const bool col_ok = (jc < G) && (r_sub < R);
    const int j = (p.C == 1) ? jc : ((p.C == 3) ? (int)(((unsigned)jc * 43691u) >> 17) : jc / p.C);
    const int c = jc - j * p.C;
    const Tap ct = cols[min(j, p.S1 - 1) * p.E];
    const double wx = ct.w, omx = 1.0 - ct.w;
    // lanes without a column (or a row) run the same loads on valid addresses and skip the store
    const BankT* colp =
        reinterpret_cast<const BankT*>(tile + (h.lead & kLeadMask)) + ((ct.k - h.cmin) * p.C + (col_ok ? c : 0));
    float* outp = out + first * G + (col_ok ? jc : 0);
    const Tap* tp = rows + first * p.E;
    if (kind >= 0 && !(h.lead & kNonConsec)) {  // warp-uniform: consecutive tap rows, one sensor row per pass
      if (p.fast) {
        const float wxf = ct.wf, omxf = ct.omwf;
        if (kind == 3)
          blend_column_consec_f32<BankT, 3>(colp, tp, p.E, pitch_e, 3, n_rows, wxf, omxf, outp, G, col_ok);
        else if (kind == 1)
          blend_column_consec_f32<BankT, 1>(colp, tp, p.E, pitch_e, 1, n_rows, wxf, omxf, outp, G, col_ok);
        else
          blend_column_consec_f32<BankT, 0>(colp, tp, p.E, pitch_e, p.C, n_rows, wxf, omxf, outp, G, col_ok);
      } else if (kind == 3)
        blend_column_consec<BankT, 3>(colp, tp, p.E, pitch_e, 3, n_rows, wx, omx, outp, G, col_ok);
      else if (kind == 1)
        blend_column_consec<BankT, 1>(colp, tp, p.E, pitch_e, 1, n_rows, wx, omx, outp, G, col_ok);
      else
        blend_column_consec<BankT, 0>(colp, tp, p.E, pitch_e, p.C, n_rows, wx, omx, outp, G, col_ok);
    } else {
      blend_column_any<BankT>(p.fast, kind, colp, tp, p.E, pitch_e, p.C, R, G, p.S0, r_sub, n_rows, ct, outp, col_ok);
    }
  }
  // ---- visited bitmask of the observation glimpse: rows {i0, i0+1} x columns {j0, j0+1}
  if (p.track && set == 0) {
    unsigned* vm = p.visited + (size_t)n * p.H * p.words;
    if (p.mode != 0) {
      for (int q = lane; q < p.H * p.words; q += 32) vm[q] = 0u;
      __syncwarp();
    }
    if (!(h.lead & kNonConsec) && p.words <= 2 && (int)h.nrows * p.words <= 32) {
      // consecutive tap rows: the touched rows are rmin .. rmin + nrows - 1; lane = (row, word), one RED per lane
      const int r = (p.words == 2) ? (lane >> 1) : lane, w = (p.words == 2) ? (lane & 1) : 0;
      const unsigned bits = w == 0 ? h.colbits[0] : h.colbits[1];
      if (r < (int)h.nrows && bits) atomicOr(vm + ((int)h.rmin + r) * p.words + w, bits);
    } else
    for (int w = 0; w < p.words; ++w) {
      unsigned bits;
      if (w < 2) {
        bits = w == 0 ? h.colbits[0] : h.colbits[1];
      } else {  // images wider than 64 px
        unsigned my = 0u;
        for (int jj = lane; jj < p.S1; jj += 32) {
          const int j0 = cols[jj * p.E].k;
          if ((j0 >> 5) == w) my |= 1u << (j0 & 31);
          if (((j0 + 1) >> 5) == w) my |= 1u << ((j0 + 1) & 31);
        }
        bits = __reduce_or_sync(0xffffffffu, my);
      }
      if (bits) {
        for (int rr = lane; rr < 2 * p.S0; rr += 32) {
          const int row = h.rmin + ((rr < p.S0) ? rows[rr * p.E].k : rows[(rr - p.S0) * p.E].k + 1);
          if (row < p.H) atomicOr(vm + row * p.words + w, bits);  // no return value -> RED: nothing waits for the old word
        }
      }
    }
    __syncwarp();
  }
}

// Two-kernel variant of the warp-per-environment path (split == 2): the scalar chains and the tile headers of ALL
// environments, one thread each.  Inside the blend kernel this work occupied one warp of eight for ~18 % of a block's
// lifetime (memory round trip + serial float64 taps) with the other seven parked at a barrier; here 10^5 threads hide
// each other's latency.  The headers travel through L2 (48 bytes per environment and glimpse set).
template <typename BankT>
__global__ void __launch_bounds__(256) image_head_kernel(const ImgParams p) {
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
    env_move(p, p, n, px, py, tcx, tcy, idx, err);
    stage_head<BankT>(p, n, p.g_hdr, p.g_ctr, px * p.lim_x, py * p.lim_y, idx, err);
    if ((p.task == APG_TASK_LOCALIZATION) && (p.mode != 0))
      stage_head<BankT>(p, n, p.g_hdr + p.N, p.g_ctr + p.N, tcx, tcy, idx, err);
    env_loss(p, p, n, s_stats, ep_end, err);
    if (err) atomicOr(p.err, err);
  }
  if (ep_end) {
    __syncthreads();
    if (threadIdx.x < 6) atomicAdd(p.stats + threadIdx.x, s_stats[threadIdx.x]);
  }
}

// S0T/S1T/CT > 0 fix the sensor geometry at compile time (the benchmark shapes 10x10x3 and 5x5x1): row loops unroll, the
// lane decomposition and every output offset become constants.  <BankT, 0, 0, 0> is the generic kernel.
template <typename BankT, int S0T, int S1T, int CT>
__global__ void __launch_bounds__(kThreads, 4) image_step_kernel(const ImgParams p_in) {
  ImgParams p = p_in;
  if (S0T > 0) {
    p.S0 = S0T; p.S1 = S1T; p.C = CT; p.P = S0T * S1T * CT;
  }
  grid_dependency_wait(p.pdl);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EnvHdr* s_hdr = reinterpret_cast<EnvHdr*>(smem_raw);
  Tap* s_taps = reinterpret_cast<Tap*>(smem_raw + (size_t)p.E * sizeof(EnvHdr));
  unsigned char* s_tile = smem_raw + (size_t)p.E * sizeof(EnvHdr) + (size_t)p.E * (p.S0 + p.S1) * sizeof(Tap);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int env0 = blockIdx.x * p.E;
  const int nE = min(p.E, p.N - env0);
  const bool two_sets = (p.task == APG_TASK_LOCALIZATION) && (p.mode != 0);
  const int n_sets = two_sets ? 2 : 1;
  constexpr int ES = (int)sizeof(BankT);

  // per-thread environment state that survives across the two glimpse sets
  double px = 0.0, py = 0.0, tcx = 0.0, tcy = 0.0;
  int idx = 0;
  __shared__ double s_stats[6];
  const bool ep_end = (p.mode == 0) && p.terminated;  // this step finishes every episode of the vector env
  if (ep_end && tid < 6) s_stats[tid] = 0.0;
  if (ep_end) __syncthreads();

  for (int set = 0; set < n_sets; ++set) {
    // ------------------------------------------------------------------ phase A: thread = environment
    if (p.split == 2) {
      // the per-environment scalars and the tile headers were computed by image_head_kernel (one thread per
      // environment over the whole batch): fetch this block's headers, then taps / copies by all threads
      EnvCtr* s_ctr = reinterpret_cast<EnvCtr*>(smem_raw + p.ctr_off);
      if (tid < nE) {
        s_hdr[tid] = p.g_hdr[(size_t)set * p.N + env0 + tid];
        s_ctr[tid] = p.g_ctr[(size_t)set * p.N + env0 + tid];
      }
      __syncthreads();
      stage_block<BankT>(p, nE, tid, blockDim.x, s_hdr, s_ctr, s_taps, s_tile);
      __pipeline_commit();
    } else if (p.split) {
      EnvCtr* s_ctr = reinterpret_cast<EnvCtr*>(smem_raw + p.ctr_off);
      if (tid < nE) {  // E <= blockDim
        unsigned err = 0;
        if (set == 0) env_move(p, p, env0 + tid, px, py, tcx, tcy, idx, err);
        stage_head<BankT>(p, tid, s_hdr, s_ctr, set == 0 ? px * p.lim_x : tcx, set == 0 ? py * p.lim_y : tcy, idx, err);
        if (err) atomicOr(p.err, err);
      }
      __syncthreads();
      stage_block<BankT>(p, nE, tid, blockDim.x, s_hdr, s_ctr, s_taps, s_tile);
      __pipeline_commit();
      // loss / reward / statistics by the block's LAST threads while the copies are in flight: nothing in the block
      // depends on them before the final barrier
      const int el = tid - ((int)blockDim.x - p.E);
      if (set == 0 && el >= 0 && el < nE) {
        unsigned err = 0;
        env_loss(p, p, env0 + el, s_stats, ep_end, err);
        if (err) atomicOr(p.err, err);
      }
    } else {
      // threads [0, E): position + taps + staging; threads [E, 2E) (when the block has them): loss / reward /
      // statistics, side by side with the first group
      const int role = tid / p.E, el = tid - role * p.E;
      const int loss_role = ((int)blockDim.x >= 2 * p.E) ? 1 : 0;
      if (el < nE && role <= loss_role) {
        unsigned err = 0;
        if (role == 0) {
          if (set == 0) env_move(p, p, env0 + el, px, py, tcx, tcy, idx, err);
          stage_env<BankT>(p, el, s_hdr, s_taps, s_tile, set == 0 ? px * p.lim_x : tcx, set == 0 ? py * p.lim_y : tcy,
                           idx, err);
        }
        if (role == loss_role && set == 0) env_loss(p, p, env0 + el, s_stats, ep_end, err);
        if (err) atomicOr(p.err, err);
      }
      __pipeline_commit();
    }
    __pipeline_wait_prior(0);
    __syncthreads();
    if (ep_end && set == 0 && tid < 6 && p.split != 2) atomicAdd(p.stats + tid, s_stats[tid]);  // one RED per block and statistic

    if ((S0T > 0) ? (S0T * S1T * CT <= kThreadBlendMaxP) : (p.thread_blend != 0)) {
      // -------------------------------------------------------------- phase B (small glimpses): thread = environment
      // A 5x5 glimpse is too small to occupy a warp; every thread blends its own environment's window from its tile
      // (no redundant work, no per-environment warp overhead), parks the floats in shared memory with an odd stride
      // (conflict-free) and the block then streams them out as one contiguous, fully coalesced run.
      float* s_out = reinterpret_cast<float*>(s_tile + (size_t)p.E * p.tile_bytes);
      const int Pp = p.P | 1;  // odd stride in floats
      // `tb` threads share an environment: thread tid serves env (tid % E) and the sensor columns tid / E, + tb, ...
      const int tb = p.tb;
      const int sub = tid / p.E;
      for (int el = tid - sub * p.E; el < nE && sub < tb; el += p.E * tb) {
        const EnvHdr h = s_hdr[el];
        const Tap* rows = s_taps + el;
        const Tap* cols = s_taps + (size_t)p.S0 * p.E + el;
        const BankT* tile = reinterpret_cast<const BankT*>(s_tile + (size_t)el * p.tile_bytes + (h.lead & kLeadMask));
        const bool consec = p.tb_consec && !(h.lead & kNonConsec);  // (nearly) always: see tap_row
        const int pitch_e = p.pitch / ES;
        float* o = s_out + (size_t)el * Pp;
        for (int j = sub; j < p.S1; j += tb) {
          const Tap ct = cols[j * p.E];
          const BankT* colp = tile + (ct.k - h.cmin) * p.C;
          for (int c = 0; c < p.C; ++c) {
            if (consec && p.fast)
              blend_column_consec_f32<BankT, CT>(colp + c, rows, p.E, pitch_e, p.C, p.S0, ct.wf, ct.omwf,
                                                 o + j * p.C + c, p.S1 * p.C, true);
            else if (consec)
              blend_column_consec<BankT, CT>(colp + c, rows, p.E, pitch_e, p.C, p.S0, ct.w, 1.0 - ct.w, o + j * p.C + c,
                                             p.S1 * p.C, true);
            else if (p.fast)
              blend_column_f32<BankT, 0, true>(colp + c, rows, p.E, pitch_e, p.C, p.S0, ct.wf, ct.omwf, o + j * p.C + c,
                                               p.S1 * p.C, true);
            else
              blend_column<BankT, 0, true>(colp + c, rows, p.E, pitch_e, p.C, p.S0, ct.w, 1.0 - ct.w, o + j * p.C + c,
                                           p.S1 * p.C, true);
          }
        }
        if (p.track && set == 0 && sub == 0) {
          unsigned* vm = p.visited + (size_t)(env0 + el) * p.H * p.words;
          if (p.mode != 0)
            for (int q = 0; q < p.H * p.words; ++q) vm[q] = 0u;
          for (int w = 0; w < p.words; ++w) {
            unsigned bits = 0u;
            if (w < 2) {
              bits = w == 0 ? h.colbits[0] : h.colbits[1];
            } else {
              for (int jj = 0; jj < p.S1; ++jj) {
                const int j0 = cols[jj * p.E].k;
                if ((j0 >> 5) == w) bits |= 1u << (j0 & 31);
                if (((j0 + 1) >> 5) == w) bits |= 1u << ((j0 + 1) & 31);
              }
            }
            if (!bits) continue;
            int last = -1;  // tap rows are non-decreasing
            for (int i = 0; i < p.S0; ++i) {
              const int r0 = h.rmin + rows[i * p.E].k;
              if (r0 > last) atomicOr(vm + r0 * p.words + w, bits);
              if (r0 + 1 > last && r0 + 1 < p.H) atomicOr(vm + (r0 + 1) * p.words + w, bits);
              last = r0 + 1;
            }
          }
        }
      }
      __syncthreads();
      float* out_blk = (set == 0 ? p.glimpse : p.target_glimpse) + (size_t)env0 * p.P;
      const int total = nE * p.P;
      const float inv_P = __frcp_rn((float)p.P);  // q / P == trunc((q + 0.5) * (1 / P)) for q < 2^16
      for (int q = tid; q < total; q += blockDim.x) {
        const int el_c = (int)(((float)q + 0.5f) * inv_P);
        out_blk[q] = s_out[el_c * Pp + (q - el_c * p.P)];
      }
      if (set + 1 < n_sets) __syncthreads();
      continue;
    }

    // ------------------------------------------------------------------ phase B: warp = environment, lane = sensor element
    const WarpLane L = warp_lane<BankT>(p, lane);
    float* out_set = (set == 0 ? p.glimpse : p.target_glimpse) + (size_t)env0 * p.P;
    for (int el = warp; el < nE; el += blockDim.x >> 5)
      blend_env_warp<BankT>(p, L, lane, el, env0 + el, set, s_hdr, s_taps, s_tile, out_set + el * p.P);
    if (set + 1 < n_sets) __syncthreads();  // the tile and the taps are reused by the target glimpse
  }
}


}  // namespace

// ======================================================================================================
// host side
// ======================================================================================================
struct apg_image_env {
  apg_image_cfg cfg{};
  cudaStream_t stream = nullptr;
  int P = 0, words = 0, E = 0, bank_pitch = 0, pitch = 0, tile_rows = 0, thread_blend = 0, threads = 0, tb = 1;
  int split = 0, ctr_off = 0, spec = 0, tile_bytes = 0, tb_consec = 1, pdl = 1;
  void* d_ghdr = nullptr;
  void* d_gctr = nullptr;
  size_t smem = 0;
  // large glimpses on uint8 banks: TMA-staged persistent kernel (csrc/apg_image_tma.cu)
  apg_img::TmaPlan tma;
  // small glimpses, classification, uint8 single-channel banks: persistent rollout kernel (csrc/apg_image_small.cu)
  apg_img::SmallPlan small;
  int force_slow = 0;
  // *_host pipelining: the step kernels can be launched over a sub-range [r0, r1) of the environments, and `hold`
  // keeps the episode clock / payload ring where they are until every chunk of the step has been launched
  int r0 = 0, r1 = -1;
  bool hold = false;
  bool top_up = true;  // rollout windows replace the autoreset payloads they consume right behind their launch
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in[8] = {};
  // bank
  void* d_bank = nullptr;
  int32_t* d_bank_labels = nullptr;
  int64_t n_items = 0;
  bool own_bank = false;
  // outputs / state
  apg_image_buffers b{};
  float* d_acc1 = nullptr; float* d_acc2 = nullptr; float* d_acc_ret = nullptr;
  double* d_stats = nullptr;
  int32_t* d_st_idx = nullptr;
  int32_t* d_st_label = nullptr;
  uint8_t* d_st_inv = nullptr;
  unsigned* d_err = nullptr;
  // payload ring
  struct Slot {
    int64_t* d_idx = nullptr;
    double* d_pos = nullptr;
    float* d_target = nullptr;
    uint8_t* d_inv = nullptr;
    unsigned char* h_pin = nullptr;  // pinned staging: idx | pos | target | inv
    cudaEvent_t copied = nullptr;
    cudaEvent_t ready = nullptr;     // device draws run on the engine's own rng stream: recorded behind the draw
    bool ready_pending = false;      // ... and waited for (on the step stream) by the first launch that reads the slot
    bool drawn = false;              // drawn from the reset streams (not queued by the caller)
    int snap = 0;                    // device mode: d_streams entry the payload was drawn from
    apg::Pcg64 h_snap[3];            // host mode: the stream states before the payload
  };
  std::vector<Slot> ring;
  int ring_head = 0, ring_count = 0;
  // Device payload draws are queued on `rng_stream`, concurrently with the step kernels: a draw only depends on the
  // stream states and on the slot it overwrites having been consumed (ev_step: recorded behind every launch that reads
  // slots).  ev_topup[]: the draws queued behind window W are joined into the step stream at the end of window W + 1 --
  // they overlap a whole window, and every window still pays for (waits for) one window's worth of draws.
  cudaStream_t rng_stream = nullptr;
  cudaEvent_t ev_step = nullptr;
  bool step_dirty = false;
  cudaEvent_t ev_topup[2] = {nullptr, nullptr};
  bool topup_valid[2] = {false, false};
  int topup_phase = 0;
  // staging for *_host
  float* d_in_action = nullptr;
  float* d_in_pred = nullptr;
  // numpy-compatible reset streams (optional): autoreset payloads are drawn by the engine when the ring is empty
  apg::Pcg64 g0{}, g1{}, g2{};
  bool have_streams = false;
  // device-side payload generation (csrc/apg_payload.cu): see configure_rng
  apg::Pcg64* d_streams = nullptr;
  unsigned* d_words = nullptr;
  unsigned* d_rej = nullptr;
  bool device_rng = false;
  int cur = 0;  // d_streams entry holding the current states
  int pp = 0;   // which half of d_rej is the clean rejection list
  bool host_rng = false;  // APG_IMG_HOST_RNG=1: always draw payloads on the host
  int64_t global_n = 0, env_offset = 0;
  std::vector<int64_t> h_idx;
  std::vector<double> h_pos;
  std::vector<float> h_target;
  std::vector<uint8_t> h_inv;
  // episode clock (all-or-nothing, shared by every env of the vector env)
  int t = 0;
  bool has_reset = false;
  bool needs_reset = false;
  int pred_width = 0;
};

namespace {

size_t slot_bytes(const apg_image_env* e) {
  const size_t N = e->cfg.num_envs;
  return N * 8 + N * 16 + N * 8 + N;
}

int stage_payload(apg_image_env* e, apg_image_env::Slot& s, const int64_t* h_idx, const double* h_pos,
                  const float* h_target, const uint8_t* h_invert) {
  const size_t N = e->cfg.num_envs;
  APG_CHECK(h_idx && h_pos, APG_ERR_INVALID, "reset payload needs h_idx and h_pos");
  APG_CHECK(e->cfg.task != APG_TASK_LOCALIZATION || h_target, APG_ERR_INVALID,
            "localization reset payload needs h_target");
  APG_CHECK(!e->cfg.randomly_invert_labels || h_invert, APG_ERR_INVALID,
            "randomly_invert_labels reset payload needs h_invert");
  for (size_t i = 0; i < N; ++i)
    APG_CHECK(h_idx[i] >= 0 && h_idx[i] < e->n_items, APG_ERR_INVALID, "bank index %lld out of range [0, %lld)",
              (long long)h_idx[i], (long long)e->n_items);
  APG_CUDA(cudaEventSynchronize(s.copied));  // previous async copy out of the pinned slot is done
  unsigned char* h = s.h_pin;
  memcpy(h, h_idx, N * 8);
  memcpy(h + N * 8, h_pos, N * 16);
  if (h_target) memcpy(h + N * 24, h_target, N * 8);
  if (h_invert) memcpy(h + N * 32, h_invert, N);
  APG_CUDA(cudaMemcpyAsync(s.d_idx, h, N * 8, cudaMemcpyHostToDevice, e->stream));
  APG_CUDA(cudaMemcpyAsync(s.d_pos, h + N * 8, N * 16, cudaMemcpyHostToDevice, e->stream));
  if (h_target) APG_CUDA(cudaMemcpyAsync(s.d_target, h + N * 24, N * 8, cudaMemcpyHostToDevice, e->stream));
  if (h_invert) APG_CUDA(cudaMemcpyAsync(s.d_inv, h + N * 32, N, cudaMemcpyHostToDevice, e->stream));
  APG_CUDA(cudaEventRecord(s.copied, e->stream));
  return APG_OK;
}

// shapes served by a persistent multi-step kernel (csrc/apg_image_small.cu, csrc/apg_image_tma.cu)
bool persistent(const apg_image_env* e) { return e->small.spec || (e->tma.spec && e->tma.have_map); }

// everything of the kernel parameter block that does not depend on the step
void fill_params(const apg_image_env* e, ImgParams& p, int mode, const float* d_action, const float* d_pred,
                 const apg_image_env::Slot* slot, float* reward2, uint8_t* done2) {
  const apg_image_cfg& c = e->cfg;
  p.N = c.num_envs; p.H = c.height; p.W = c.width; p.C = c.channels; p.S0 = c.sensor_h; p.S1 = c.sensor_w;
  p.P = e->P; p.K = c.num_classes; p.task = c.task; p.invert = c.randomly_invert_labels; p.track = c.track_visited;
  p.words = e->words; p.T = c.step_limit; p.E = e->E; p.bank_pitch = e->bank_pitch; p.pitch = e->pitch; p.tile_rows = e->tile_rows; p.tile_bytes = e->tile_bytes; p.thread_blend = e->thread_blend; p.tb = e->tb; p.fast = e->cfg.blend_f32 ? 1 : 0; p.sparse = e->cfg.sparse_reward ? 1 : 0; p.split = e->split; p.pdl = e->pdl; p.tb_consec = e->tb_consec; p.g_hdr = static_cast<EnvHdr*>(e->d_ghdr); p.g_ctr = static_cast<EnvCtr*>(e->d_gctr); p.ctr_off = e->ctr_off; p.mode = mode; p.force_slow = e->force_slow; p.n0 = e->r0; p.n1 = e->r1 < 0 ? c.num_envs : e->r1;
  p.scale = c.sensor_scale;
  p.msl_x = c.max_step_length[0]; p.msl_y = c.max_step_length[1];
  // sensor_pos_lim_pixels (image_perception_module.py:419-423): x uses sensor_size[0], y uses sensor_size[1]
  p.lim_x = ((double)c.width - 1.0) / 2.0 - ((double)c.sensor_h * c.sensor_scale - 1.0) / 2.0;
  p.lim_y = ((double)c.height - 1.0) / 2.0 - ((double)c.sensor_w * c.sensor_scale - 1.0) / 2.0;
  p.offy = ((double)c.height - 1.0) / 2.0; p.offx = ((double)c.width - 1.0) / 2.0;
  p.half0 = ((double)c.sensor_h - 1.0) / 2.0; p.half1 = ((double)c.sensor_w - 1.0) / 2.0;
  // render geometry of the overlay (image_perception_module.py:167-171)
  const int render_w = c.sensor_w > 128 ? c.sensor_w : 128;
  p.rs = (double)render_w / (double)c.width;
  const int render_h = (int)nearbyint(p.rs * (double)c.height);
  p.half_wr = (double)render_w / 2.0; p.half_hr = (double)render_h / 2.0;
  if (c.task == APG_TASK_CLASSIFICATION) {
    p.ce_scale = 1.0 / (std::log((double)c.num_classes) - 0.0);
    p.ce_offset = -0.0 * p.ce_scale;
  } else {
    const double std_ = 2.0 / std::sqrt(12.0);
    const double blind = std_ * std_;
    const double s = 1.0 / (blind - 0.0);
    p.mse_scale = (float)s;   // python-float scale applied to a float32 array stays float32
    p.mse_offset = (float)(-0.0 * s);
  }
  p.bank = e->d_bank; p.bank_labels = e->d_bank_labels; p.action = d_action; p.pred = d_pred;
  if (slot) {
    p.new_idx = reinterpret_cast<const long long*>(slot->d_idx); p.new_pos = slot->d_pos;
    p.new_target = slot->d_target; p.new_inv = slot->d_inv;
  }
  p.pos = e->b.d_pos; p.st_idx = e->d_st_idx; p.st_label = e->d_st_label; p.st_target = e->b.d_cur_target_xy;
  p.st_inv = e->d_st_inv;
  p.glimpse = e->b.d_glimpse; p.target_glimpse = e->b.d_target_glimpse; p.glimpse_pos = e->b.d_glimpse_pos;
  p.time_step = e->b.d_time_step; p.inverted_label = e->b.d_inverted_label; p.reward = e->b.d_reward;
  p.base_reward = e->b.d_base_reward; p.loss = e->b.d_loss; p.target_label = e->b.d_target_label;
  p.target_xy = e->b.d_target_xy; p.terminated_out = e->b.d_terminated; p.truncated_out = e->b.d_truncated;
  p.index = reinterpret_cast<long long*>(e->b.d_index); p.reward2 = reward2; p.done2 = done2;
  p.stamp_xy = e->b.d_stamp_xy; p.stamp_q = e->b.d_stamp_q; p.visited = e->b.d_visited; p.err = e->d_err;
  p.acc1 = e->d_acc1; p.acc2 = e->d_acc2; p.acc_ret = e->d_acc_ret; p.ep_avg1 = e->b.d_ep_avg1; p.ep_final1 = e->b.d_ep_final1;
  p.ep_avg2 = e->b.d_ep_avg2; p.ep_final2 = e->b.d_ep_final2; p.ep_return = e->b.d_ep_return;
  p.first_correct = e->b.d_first_correct; p.last_incorrect = e->b.d_last_incorrect; p.stats = e->d_stats;
  p.inv_k = (float)(1.0 / (double)(c.num_classes > 0 ? c.num_classes : 1));
}

// mode: 0 move, 1 autoreset, 2 explicit reset
int launch(apg_image_env* e, int mode, const float* d_action, const float* d_pred,
           const apg_image_env::Slot* slot, float* reward2, uint8_t* done2) {
  const apg_image_cfg& c = e->cfg;
  ImgParams p{};
  fill_params(e, p, mode, d_action, d_pred, slot, reward2, done2);
  if (persistent(e)) {  // explicit reset through the persistent kernels (steps go through rollout_window)
    apg_img::RolloutPlan ro{};
    ro.n_steps = 1; ro.t0 = 0; ro.need_reset0 = 0; ro.first_mode = 2;
    ro.time_obs[0] = (float)((0.0 / (double)c.step_limit) * 2.0 - 1.0);
    ro.slot_idx[0] = reinterpret_cast<const long long*>(slot->d_idx); ro.slot_pos[0] = slot->d_pos;
    ro.slot_target[0] = slot->d_target; ro.slot_inv[0] = slot->d_inv;
    const int n = e->small.spec ? apg_img::small_launch(e->small, p, ro, e->stream) : apg_img::tma_launch(e->tma, p, ro, e->stream);
    if (n < 0) return apg::fail(APG_ERR_CUDA, "image rollout kernel: %s", cudaGetErrorString((cudaError_t)(-n)));
    apg::count_launch(n);
    e->t = 0;
    e->needs_reset = false;
    return APG_OK;
  }
  int t_after;
  if (mode == 0) {
    p.t_stamp = e->t;
    t_after = e->t + 1;
    p.terminated = (t_after >= c.step_limit) ? 1 : 0;
  } else {
    p.t_stamp = 0;
    t_after = 0;
    p.terminated = 0;
  }
  p.time_obs = (float)(((double)t_after / (double)c.step_limit) * 2.0 - 1.0);

  const int grid = apg::ceil_div(c.num_envs, e->E);
  // one launch path for every kernel of a step: plain, or with programmatic stream serialization (e->pdl)
  auto go = [&](auto kernel, int g, int threads, size_t smem) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)g); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem;
    cfg.stream = e->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = e->pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, p);
  };
  if (e->split == 2) {
    const int hgrid = apg::ceil_div(c.num_envs, 256);
    if (c.bank_dtype == APG_BANK_U8) go(image_head_kernel<unsigned char>, hgrid, 256, 0);
    else go(image_head_kernel<float>, hgrid, 256, 0);
    apg::count_launch();
  }
  if (c.bank_dtype == APG_BANK_U8) {
    if (e->spec == 1) go(image_step_kernel<unsigned char, 10, 10, 3>, grid, e->threads, e->smem);
    else if (e->spec == 2) go(image_step_kernel<unsigned char, 5, 5, 1>, grid, e->threads, e->smem);
    else go(image_step_kernel<unsigned char, 0, 0, 0>, grid, e->threads, e->smem);
  } else {
    go(image_step_kernel<float, 0, 0, 0>, grid, e->threads, e->smem);
  }
  APG_CUDA(cudaGetLastError());
  apg::count_launch();
  if (!e->hold) {
    e->t = t_after;
    e->needs_reset = (mode == 0) && p.terminated;
  }
  return APG_OK;
}

// Host-side join of the rng stream (rare paths: explicit reset, stream get/set, bank change, destroy).
int join_rng(apg_image_env* e) {
  if (e->rng_stream) APG_CUDA(cudaStreamSynchronize(e->rng_stream));
  for (auto& s : e->ring) s.ready_pending = false;
  e->topup_valid[0] = e->topup_valid[1] = false;
  return APG_OK;
}

// the step stream waits for the draw of a slot it is about to read
int wait_slot(apg_image_env* e, apg_image_env::Slot& s) {
  if (s.ready_pending) {
    APG_CUDA(cudaStreamWaitEvent(e->stream, s.ready, 0));
    s.ready_pending = false;
  }
  return APG_OK;
}

// recorded behind every launch that reads payload slots: draws that overwrite them wait for it
int mark_step(apg_image_env* e) {
  if (e->device_rng) {
    APG_CUDA(cudaEventRecord(e->ev_step, e->stream));
    e->step_dirty = true;
  }
  return APG_OK;
}

// ---- the reference's three reset-path Generators ---------------------------------------------------------------------
// Device mode (the normal case): the states live in d_streams, (ring + 1) x 3 entries -- entry `cur` holds the states
// after every payload drawn so far, and the entry a payload was drawn FROM stays intact until that payload has been
// consumed, so that the streams can be rewound to the oldest unconsumed payload (apg_image_get_streams, explicit reset).
// Everything is allocated and uploaded by configure_rng (bank / stream setup time): nothing on the step path allocates,
// synchronises or loads a kernel for the first time.
int configure_rng(apg_image_env* e) {
  e->device_rng = false;
  if (int rc = join_rng(e)) return rc;
  if (!e->have_streams || !e->d_bank) return APG_OK;
  if (e->host_rng || !apg::payload_device_ok((unsigned long long)e->n_items, e->global_n)) return APG_OK;
  const size_t n_states = 3 * (e->ring.size() + 1);
  if (!e->d_streams) {
    APG_CUDA(cudaMalloc(&e->d_streams, n_states * sizeof(apg::Pcg64)));
    APG_CUDA(apg::dev_alloc(&e->d_rej, 2 * ((size_t)apg::kPayloadSlack + 1)));
    APG_CUDA(cudaMalloc(&e->d_words, apg::payload_scratch_words(e->cfg.num_envs) * sizeof(unsigned)));
    APG_CUDA(apg::payload_preload());
  }
  apg::Pcg64 h[3] = {e->g0, e->g1, e->g2};
  APG_CUDA(cudaMemcpyAsync(e->d_streams + 3 * e->cur, h, sizeof(h), cudaMemcpyHostToDevice, e->stream));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  e->device_rng = true;
  return APG_OK;
}

// host copies of the current states (device mode: fetched from entry `cur`; synchronises the stream)
int streams_to_host(apg_image_env* e) {
  if (!e->device_rng) return APG_OK;
  if (int rc = join_rng(e)) return rc;
  apg::Pcg64 h[3];
  APG_CUDA(cudaMemcpyAsync(h, e->d_streams + 3 * e->cur, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  e->g0 = h[0]; e->g1 = h[1]; e->g2 = h[2];
  return APG_OK;
}

// Drops the queued payloads that were drawn from the streams and rewinds the streams to the state before the oldest
// of them (payloads queued by the caller are the caller's business: they make the rewind impossible).
int rewind_streams(apg_image_env* e) {
  if (e->ring_count == 0) return APG_OK;
  for (int i = 0; i < e->ring_count; ++i)
    APG_CHECK(e->ring[(e->ring_head + i) % (int)e->ring.size()].drawn, APG_ERR_STATE,
              "%d caller-queued reset payloads are pending", e->ring_count);
  const apg_image_env::Slot& first = e->ring[e->ring_head];
  if (e->device_rng) {
    e->cur = first.snap;
    // the rejection list of the half `pp` points at was cleared by the last payload that was drawn: still clean
  } else {
    e->g0 = first.h_snap[0]; e->g1 = first.h_snap[1]; e->g2 = first.h_snap[2];
  }
  e->ring_count = 0;
  return APG_OK;
}

// One autoreset payload from the reference's streams, in the reference's draw order, for the global batch; this
// shard keeps its slice.  (image_localization.py:153-156 -> dataset_iterator.py:57 -> image_perception_module.py:131-133
// -> :159-161)
int generate_payload(apg_image_env* e) {
  APG_CHECK(e->have_streams, APG_ERR_NEED_PAYLOAD, "no reset payload queued and no reset streams set");
  APG_CHECK(e->ring_count < (int)e->ring.size(), APG_ERR_STATE, "reset payload ring is full");
  apg_image_env::Slot& s = e->ring[(e->ring_head + e->ring_count) % (int)e->ring.size()];
  if (e->device_rng) {
    // drawn on the device, in-stream, bit-identical to the numpy draws below (tests/test_rng_gpu.py)
    const int n_states = (int)e->ring.size() + 1;
    const int next = (e->cur + 1) % n_states;
    apg::PayloadArgs a{};
    a.streams = e->d_streams + 3 * e->cur; a.streams_out = e->d_streams + 3 * next; a.Ng = e->global_n; a.off = e->env_offset; a.N = e->cfg.num_envs;
    a.M = (unsigned long long)e->n_items; a.loc = e->cfg.task == APG_TASK_LOCALIZATION ? 1 : 0;
    a.inv = e->cfg.randomly_invert_labels ? 1 : 0;
    a.d_idx = reinterpret_cast<long long*>(s.d_idx); a.d_pos = s.d_pos; a.d_target = s.d_target; a.d_inv = s.d_inv;
    a.words = e->d_words;
    a.rej = e->d_rej + (size_t)e->pp * (apg::kPayloadSlack + 1);
    a.rej_next = e->d_rej + (size_t)(e->pp ^ 1) * (apg::kPayloadSlack + 1);
    e->pp ^= 1;
    if (e->step_dirty) {  // the slot (and the stream-state entry) may still be read by a launch on the step stream
      APG_CUDA(cudaStreamWaitEvent(e->rng_stream, e->ev_step, 0));
      e->step_dirty = false;
    }
    const int n = apg::launch_device_payload(a, e->rng_stream);
    if (n < 0) return apg::fail(APG_ERR_CUDA, "payload kernels: %s", cudaGetErrorString((cudaError_t)(-n)));
    apg::count_launch(n);
    APG_CUDA(cudaEventRecord(s.ready, e->rng_stream));
    s.ready_pending = true;
    s.drawn = true; s.snap = e->cur;
    e->cur = next;
    e->ring_count += 1;
    return APG_OK;
  }
  s.h_snap[0] = e->g0; s.h_snap[1] = e->g1; s.h_snap[2] = e->g2;
  const int64_t N = e->cfg.num_envs, Ng = e->global_n, off = e->env_offset;
  const bool loc = e->cfg.task == APG_TASK_LOCALIZATION, inv = e->cfg.randomly_invert_labels != 0;
  e->h_idx.resize(N); e->h_pos.resize(2 * N);
  if (loc) e->h_target.resize(2 * N);
  if (inv) e->h_inv.resize(N);
  if (loc) {  // g0.uniform(-1, 1, (Ng, 2)).astype(float32)
    apg::pcg64_advance(e->g0, (uint64_t)(2 * off));
    for (int64_t i = 0; i < 2 * N; ++i) e->h_target[i] = (float)apg::rng_uniform(e->g0, -1.0, 2.0);
    apg::pcg64_advance(e->g0, (uint64_t)(2 * (Ng - off - N)));
  }
  for (int64_t j = 0; j < Ng; ++j) {  // g2.integers(0, len(dataset), Ng): rejections shift the stream, draw them all
    const uint64_t v = apg::rng_integers_below(e->g2, (uint64_t)e->n_items);
    if (j >= off && j < off + N) e->h_idx[j - off] = (int64_t)v;
  }
  if (inv) {  // g1.integers(0, 2, size=Ng) == 1
    for (int64_t j = 0; j < Ng; ++j) {
      const uint64_t v = apg::rng_integers_below(e->g1, 2);
      if (j >= off && j < off + N) e->h_inv[j - off] = (uint8_t)(v == 1);
    }
  }
  apg::pcg64_advance(e->g1, (uint64_t)(2 * off));  // g1.uniform(-1, 1, (Ng, 2))
  for (int64_t i = 0; i < 2 * N; ++i) e->h_pos[i] = apg::rng_uniform(e->g1, -1.0, 2.0);
  apg::pcg64_advance(e->g1, (uint64_t)(2 * (Ng - off - N)));
  int rc = stage_payload(e, s, e->h_idx.data(), e->h_pos.data(), loc ? e->h_target.data() : nullptr,
                         inv ? e->h_inv.data() : nullptr);
  if (rc != APG_OK) return rc;
  s.drawn = true;
  e->ring_count += 1;
  return APG_OK;
}

// n_steps consecutive steps of a persistent kernel: as few launches as the autoreset payload ring and the
// per-launch limits allow (one for the bench's 51-step windows).  Payloads of the autoresets inside a window are drawn
// in-stream before its launch.
int rollout_window(apg_image_env* e, const float* d_actions, const float* d_preds, int n_steps, float* d_rewards,
                   uint8_t* d_dones) {
  const size_t N = e->cfg.num_envs;
  const int ring = (int)e->ring.size();
  const int max_resets = ring < apg_img::kRolloutMaxSlots ? ring : apg_img::kRolloutMaxSlots;
  int done = 0;
  while (done < n_steps) {
    apg_img::RolloutPlan ro{};
    ro.t0 = e->t; ro.need_reset0 = e->needs_reset ? 1 : 0; ro.first_mode = 0;
    int t = e->t, m = 0, resets = 0;
    bool need = e->needs_reset;
    while (done + m < n_steps && m < apg_img::kRolloutMaxSteps) {
      if (need) {
        if (resets == max_resets) break;
        ro.reset_slot[m] = (unsigned char)resets++;
        t = 0; need = false;
      } else {
        t++; need = t >= e->cfg.step_limit;
      }
      ro.time_obs[m] = (float)(((double)t / (double)e->cfg.step_limit) * 2.0 - 1.0);  // t = the clock after step m
      m++;
    }
    while (e->ring_count < resets) {
      APG_CHECK(e->have_streams, APG_ERR_NEED_PAYLOAD,
                "autoreset step is due but no reset payload is queued (apg_image_queue_reset_payload)");
      int rc = generate_payload(e);
      if (rc != APG_OK) return rc;
    }
    for (int i = 0; i < resets; ++i) {
      apg_image_env::Slot& s = e->ring[(e->ring_head + i) % ring];
      if (int rc = wait_slot(e, s)) return rc;
      ro.slot_idx[i] = reinterpret_cast<const long long*>(s.d_idx); ro.slot_pos[i] = s.d_pos;
      ro.slot_target[i] = s.d_target; ro.slot_inv[i] = s.d_inv;
    }
    ro.n_steps = m;
    ImgParams p{};
    fill_params(e, p, 0, d_actions + (size_t)done * N * 2, d_preds + (size_t)done * N * e->pred_width, nullptr,
                d_rewards ? d_rewards + (size_t)done * N : nullptr, d_dones ? d_dones + (size_t)done * N : nullptr);
    const int n = e->small.spec ? apg_img::small_launch(e->small, p, ro, e->stream) : apg_img::tma_launch(e->tma, p, ro, e->stream);
    if (n < 0) return apg::fail(APG_ERR_CUDA, "image rollout kernel: %s", cudaGetErrorString((cudaError_t)(-n)));
    apg::count_launch(n);
    if (resets > 0)
      if (int rc = mark_step(e)) return rc;
    if (!e->hold) {
      int drawn = 0;
      for (int i = 0; i < resets; ++i) drawn += e->ring[(e->ring_head + i) % ring].drawn ? 1 : 0;
      e->ring_head = (e->ring_head + resets) % ring;
      e->ring_count -= resets;
      e->t = t;
      e->needs_reset = need;
      // Replace the payloads this launch consumed right behind it, in-stream: the draws of the NEXT autoresets then run
      // while the host is idle anyway instead of sitting in front of the next window's launch.  The streams run ahead
      // of the reference's position by the queued payloads; apg_image_get_streams / an explicit reset rewind them.
      if (e->device_rng && e->have_streams && e->top_up)
        for (int i = 0; i < drawn && e->ring_count < ring; ++i) {
          int rc = generate_payload(e);
          if (rc != APG_OK) return rc;
        }
    }
    done += m;
  }
  // join: the draws queued behind the PREVIOUS window have had this whole window to run next to; the step stream
  // waits for them here, so the device time of a window includes one window's worth of draws (overlapped, not skipped)
  if (e->device_rng) {
    const int ph = e->topup_phase;
    APG_CUDA(cudaEventRecord(e->ev_topup[ph], e->rng_stream));
    e->topup_valid[ph] = true;
    if (e->topup_valid[ph ^ 1]) APG_CUDA(cudaStreamWaitEvent(e->stream, e->ev_topup[ph ^ 1], 0));
    e->topup_phase = ph ^ 1;
  }
  return APG_OK;
}

int step_once(apg_image_env* e, const float* d_action, const float* d_pred, float* reward2, uint8_t* done2) {
  if (persistent(e)) return rollout_window(e, d_action, d_pred, 1, reward2, done2);
  if (e->needs_reset && e->ring_count == 0 && e->have_streams) {
    int rc = generate_payload(e);
    if (rc != APG_OK) return rc;
  }
  if (e->needs_reset) {
    APG_CHECK(e->ring_count > 0, APG_ERR_NEED_PAYLOAD,
              "autoreset step is due but no reset payload is queued (apg_image_queue_reset_payload)");
    apg_image_env::Slot& s = e->ring[e->ring_head];
    int rc = wait_slot(e, s);
    if (rc != APG_OK) return rc;
    rc = launch(e, 1, d_action, d_pred, &s, reward2, done2);
    if (rc != APG_OK) return rc;
    rc = mark_step(e);
    if (rc != APG_OK) return rc;
    if (!e->hold) {
      e->ring_head = (e->ring_head + 1) % (int)e->ring.size();
      e->ring_count -= 1;
    }
    return APG_OK;
  }
  return launch(e, 0, d_action, d_pred, nullptr, reward2, done2);
}

// advances the episode clock / payload ring by one step (what step_once does when `hold` is off)
void commit_step(apg_image_env* e) {
  if (e->needs_reset) {
    e->ring_head = (e->ring_head + 1) % (int)e->ring.size();
    e->ring_count -= 1;
    e->t = 0;
    e->needs_reset = false;
  } else {
    e->t += 1;
    e->needs_reset = e->t >= e->cfg.step_limit;
  }
}

// device-visible alias of a host pointer when the memory is page-locked (cudaHostAlloc / cudaHostRegister / torch pinned
// memory): kernels can store results straight into it (posted writes over PCIe, no device-to-host copy afterwards)
template <typename T>
T* mapped_alias(T* h) {
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

int apg_image_create(const apg_image_cfg* cfg, apg_image_env** out) {
  APG_CHECK(cfg && out, APG_ERR_INVALID, "apg_image_create: null argument");
  *out = nullptr;
  APG_CHECK(cfg->num_envs > 0, APG_ERR_INVALID, "num_envs must be positive");
  APG_CHECK(cfg->task == APG_TASK_CLASSIFICATION || cfg->task == APG_TASK_LOCALIZATION, APG_ERR_INVALID, "bad task");
  APG_CHECK(cfg->channels == 1 || cfg->channels == 3, APG_ERR_INVALID,
            "Target channels must be either 1 or 3 but is %d.", cfg->channels);
  APG_CHECK(cfg->sensor_h > 0 && cfg->sensor_w > 0 && cfg->sensor_scale > 0, APG_ERR_INVALID, "bad sensor geometry");
  APG_CHECK((double)cfg->height >= cfg->sensor_h * cfg->sensor_scale &&
                (double)cfg->width >= cfg->sensor_w * cfg->sensor_scale,
            APG_ERR_INVALID, "Image size (%d, %d) cannot be smaller than effective sensor size (%g, %g).",
            cfg->height, cfg->width, cfg->sensor_h * cfg->sensor_scale, cfg->sensor_w * cfg->sensor_scale);
  APG_CHECK(cfg->height >= 2 && cfg->width >= 2 && cfg->height <= 32768 && cfg->width <= 32768, APG_ERR_INVALID,
            "image must be between 2x2 and 32768x32768");
  APG_CHECK(cfg->step_limit > 0, APG_ERR_INVALID, "step_limit must be positive");
  APG_CHECK(cfg->bank_dtype == APG_BANK_U8 || cfg->bank_dtype == APG_BANK_F32, APG_ERR_INVALID, "bad bank dtype");
  if (cfg->task == APG_TASK_CLASSIFICATION)
    APG_CHECK(cfg->num_classes >= 2 && cfg->num_classes <= kMaxK, APG_ERR_INVALID, "num_classes must be in [2, %d]", kMaxK);
  int ndev = 0;
  cudaError_t de = cudaGetDeviceCount(&ndev);
  if (de != cudaSuccess || ndev <= 0)
    return apg::fail(APG_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                     cudaGetErrorString(de));
  APG_CHECK(cfg->device >= 0 && cfg->device < ndev, APG_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
  APG_CUDA(cudaSetDevice(cfg->device));

  apg_image_env* e = new apg_image_env();
  e->cfg = *cfg;
  const size_t N = cfg->num_envs;
  e->P = cfg->sensor_h * cfg->sensor_w * cfg->channels;
  e->words = (cfg->width + 31) / 32;
  {
    // staged tile geometry: rows [i0[0], i0[S0-1]+1], the aligned byte range covering columns [j0[0], j0[S1-1]+1]
    const int es = cfg->bank_dtype == APG_BANK_U8 ? 1 : 4;
    const int row_bytes = (cfg->width * cfg->channels * es + 15) & ~15;  // padded bank rows: 16-byte aligned
    e->bank_pitch = row_bytes;
    const int span_cols = (int)std::ceil((cfg->sensor_w - 1) * cfg->sensor_scale) + 3;
    const int span_rows = (int)std::ceil((cfg->sensor_h - 1) * cfg->sensor_scale) + 3;
    int pitch = span_cols * cfg->channels * es + 16;  // + lead bytes of the aligned-down start
    pitch = (pitch + 15) & ~15;
    if (pitch > row_bytes) pitch = row_bytes;
    e->pitch = pitch;
    e->tile_rows = span_rows < cfg->height ? span_rows : cfg->height;
    // (the row below an environment's last staged row may be read with weight 0, see tap_row: it is the next
    //  environment's first row, or the `pitch` spare bytes added once behind the last tile)
    e->tile_bytes = e->pitch * e->tile_rows;
    size_t per_env = sizeof(EnvHdr) + (size_t)(cfg->sensor_h + cfg->sensor_w) * 16 + (size_t)e->tile_bytes;
    // small glimpses (<= 96 floats, e.g. 5x5x1 / 5x5x3): one thread blends a whole environment
    e->thread_blend = e->P <= kThreadBlendMaxP ? 1 : 0;
    int E;
    if (e->thread_blend) {
      per_env += (size_t)(e->P | 1) * 4;  // parked output floats
      E = (int)((64 * 1024) / per_env);
      E = E >= 64 ? 64 : 32;  // swept on B200: 64 beats 32 and 128 for the 5x5x1 MNIST glimpse
      e->tb = cfg->sensor_w >= 2 ? 2 : 1;
      e->threads = E * e->tb;
    } else {
      // small blocks: many blocks per SM overlap each other's staging / blend phases (swept on B200: 64 envs per
      // block is best for 5x5x1 glimpses, 32 for 10x10x3 ones)
      E = (int)((72 * 1024) / per_env);
      E = E > 64 ? 64 : E;
      if (per_env > 768 && E > 32) E = 32;
      e->threads = kThreads;
      e->split = 2;  // 0: one thread per env does all of phase A; 1: taps / copies by all threads; 2: + head kernel
    }
    E = E >= 8 ? (E & ~7) : (E >= 4 ? 4 : (E >= 2 ? 2 : 1));
    e->E = E;
    e->ctr_off = (int)(per_env * E);
    if (e->split) per_env += sizeof(EnvCtr);
    e->smem = per_env * E + (size_t)e->pitch;
    if (e->smem > 200 * 1024) {
      const size_t need = e->smem;
      delete e;
      return apg::fail(APG_ERR_INVALID, "sensor window too large for the shared-memory tile (%zu bytes)", need);
    }
  }
  e->pred_width = cfg->task == APG_TASK_CLASSIFICATION ? cfg->num_classes : 2;
  const bool loc = cfg->task == APG_TASK_LOCALIZATION;
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
  A(apg::dev_alloc(&e->b.d_glimpse, N * e->P));
  if (loc) A(apg::dev_alloc(&e->b.d_target_glimpse, N * e->P));
  A(apg::dev_alloc(&e->b.d_glimpse_pos, N * 2));
  A(apg::dev_alloc(&e->b.d_time_step, N));
  if (cfg->randomly_invert_labels) A(apg::dev_alloc(&e->b.d_inverted_label, N));
  A(apg::dev_alloc(&e->b.d_reward, N));
  A(apg::dev_alloc(&e->b.d_base_reward, N));
  A(apg::dev_alloc(&e->b.d_loss, N));
  if (!loc) A(apg::dev_alloc(&e->b.d_target_label, N));
  if (loc) A(apg::dev_alloc(&e->b.d_target_xy, N * 2));
  A(apg::dev_alloc(&e->b.d_terminated, N));
  A(apg::dev_alloc(&e->b.d_truncated, N));
  A(apg::dev_alloc(&e->b.d_index, N));
  A(apg::dev_alloc(&e->b.d_pos, N * 2));
  if (loc) A(apg::dev_alloc(&e->b.d_cur_target_xy, N * 2));
  if (cfg->track_visited) {
    A(apg::dev_alloc(&e->b.d_visited, N * cfg->height * e->words));
    A(apg::dev_alloc(&e->b.d_stamp_xy, (size_t)cfg->step_limit * 2 * N));
    A(apg::dev_alloc(&e->b.d_stamp_q, (size_t)cfg->step_limit * N));
  }
  A(apg::dev_alloc(&e->d_acc1, N)); A(apg::dev_alloc(&e->d_acc2, N)); A(apg::dev_alloc(&e->d_acc_ret, N));
  A(apg::dev_alloc(&e->b.d_ep_avg1, N)); A(apg::dev_alloc(&e->b.d_ep_final1, N)); A(apg::dev_alloc(&e->b.d_ep_avg2, N));
  A(apg::dev_alloc(&e->b.d_ep_final2, N)); A(apg::dev_alloc(&e->b.d_ep_return, N));
  A(apg::dev_alloc(&e->b.d_first_correct, N)); A(apg::dev_alloc(&e->b.d_last_incorrect, N));
  A(apg::dev_alloc(&e->d_stats, (size_t)APG_NUM_STATS));
  A(apg::dev_alloc(&e->d_st_idx, N));
  A(apg::dev_alloc(&e->d_st_label, N));
  A(apg::dev_alloc(&e->d_st_inv, N));
  A(apg::dev_alloc(&e->d_err, 1));
  if (e->split == 2) {
    A(cudaMalloc(&e->d_ghdr, 2 * N * sizeof(EnvHdr)));
    A(cudaMalloc(&e->d_gctr, 2 * N * sizeof(EnvCtr)));
  }
  A(apg::dev_alloc(&e->d_in_action, N * 2));
  A(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  A(cudaStreamCreateWithFlags(&e->rng_stream, cudaStreamNonBlocking));
  A(cudaEventCreateWithFlags(&e->ev_step, cudaEventDisableTiming));
  for (auto& ev : e->ev_topup) A(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto& ev : e->ev_in) A(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  A(apg::dev_alloc(&e->d_in_pred, N * e->pred_width));
  int depth = cfg->payload_ring <= 0 ? APG_PAYLOAD_RING_DEFAULT : cfg->payload_ring;
  if (depth > APG_PAYLOAD_RING_MAX) depth = APG_PAYLOAD_RING_MAX;
  e->ring.resize(depth);
  for (auto& s : e->ring) {
    A(apg::dev_alloc(&s.d_idx, N));
    A(apg::dev_alloc(&s.d_pos, N * 2));
    A(apg::dev_alloc(&s.d_target, N * 2));
    A(apg::dev_alloc(&s.d_inv, N));
    A(cudaMallocHost(reinterpret_cast<void**>(&s.h_pin), slot_bytes(e)));
    A(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
  }
  if (err == cudaSuccess && e->smem > 48 * 1024) {
    A(cudaFuncSetAttribute(image_step_kernel<unsigned char, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
    A(cudaFuncSetAttribute(image_step_kernel<unsigned char, 10, 10, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
    A(cudaFuncSetAttribute(image_step_kernel<unsigned char, 5, 5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
    A(cudaFuncSetAttribute(image_step_kernel<float, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
  }
  // compile-time geometry for the benchmark shapes
  if (cfg->bank_dtype == APG_BANK_U8 && cfg->sensor_h == 10 && cfg->sensor_w == 10 && cfg->channels == 3) e->spec = 1;
  if (cfg->bank_dtype == APG_BANK_U8 && cfg->sensor_h == 5 && cfg->sensor_w == 5 && cfg->channels == 1) e->spec = 2;
  // developer / test switches (none of them changes results): APG_IMG_GENERIC=1 keeps every shape on the generic staged
  // kernels, APG_IMG_FORCE_SLOW=1 sends every environment of the TMA path through its general per-output branch,
  // APG_IMG_HOST_RNG=1 draws the autoreset payloads on the host, APG_IMG_PDL=0 turns programmatic dependent launch off
  const char* ov_generic = getenv("APG_IMG_GENERIC");
  const bool generic_only = ov_generic && atoi(ov_generic) != 0;
  if (generic_only) e->spec = 0;
  if (const char* ov = getenv("APG_IMG_FORCE_SLOW")) e->force_slow = atoi(ov) != 0;
  if (const char* ov = getenv("APG_IMG_HOST_RNG")) e->host_rng = atoi(ov) != 0;
  if (const char* ov = getenv("APG_IMG_PDL")) e->pdl = atoi(ov) ? 1 : 0;
  if (const char* ov = getenv("APG_IMG_TOP_UP")) e->top_up = atoi(ov) != 0;
  if (const char* ov = getenv("APG_L2_FETCH")) {  // experiment: L2 fetch granularity hint of the context (32 / 64 / 128 bytes)
    size_t before = 0, after = 0;
    cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(ov));
    cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
    fprintf(stderr, "apg: cudaLimitMaxL2FetchGranularity %zu -> %zu\n", before, after);
  }
  if (err == cudaSuccess && !generic_only && e->thread_blend && apg_img::small_plan(e->small, *cfg, e->words))
    A(apg_img::small_preload(e->small));
  if (err == cudaSuccess && !generic_only && !e->thread_blend) {
    cudaError_t terr = cudaSuccess;
    if (apg_img::tma_plan(e->tma, *cfg, e->bank_pitch, e->words, &terr)) {
      A(apg_img::tma_preload(e->tma));
    }
    A(terr);
  }
  if (err == cudaSuccess) {  // load every kernel a step can launch now (lazy module loading would do it on first use)
    cudaFuncAttributes fa;
    if (cfg->bank_dtype == APG_BANK_U8) {
      A(cudaFuncGetAttributes(&fa, image_head_kernel<unsigned char>));
      A(cudaFuncGetAttributes(&fa, image_step_kernel<unsigned char, 0, 0, 0>));
      A(cudaFuncGetAttributes(&fa, image_step_kernel<unsigned char, 10, 10, 3>));
      A(cudaFuncGetAttributes(&fa, image_step_kernel<unsigned char, 5, 5, 1>));
    } else {
      A(cudaFuncGetAttributes(&fa, image_head_kernel<float>));
      A(cudaFuncGetAttributes(&fa, image_step_kernel<float, 0, 0, 0>));
    }
  }
  if (err != cudaSuccess) {
    apg_image_destroy(e);
    return apg::fail(APG_ERR_CUDA, "apg_image_create: allocation failed: %s", cudaGetErrorString(err));
  }
  *out = e;
  return APG_OK;
}

void apg_image_destroy(apg_image_env* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  if (e->rng_stream) { cudaStreamSynchronize(e->rng_stream); cudaStreamDestroy(e->rng_stream); }
  if (e->ev_step) cudaEventDestroy(e->ev_step);
  for (auto& ev : e->ev_topup) if (ev) cudaEventDestroy(ev);
  if (e->d_streams) cudaFree(e->d_streams);
  if (e->d_words) cudaFree(e->d_words);
  if (e->d_rej) cudaFree(e->d_rej);
  if (e->d_ghdr) cudaFree(e->d_ghdr);
  if (e->d_gctr) cudaFree(e->d_gctr);
  apg_img::tma_release(e->tma);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  for (auto& ev : e->ev_in) if (ev) cudaEventDestroy(ev);
  if (e->own_bank) cudaFree(e->d_bank);
  if (e->own_bank) cudaFree(e->d_bank_labels);
  apg::dev_free(e->b.d_glimpse); apg::dev_free(e->b.d_target_glimpse); apg::dev_free(e->b.d_glimpse_pos);
  apg::dev_free(e->b.d_time_step); apg::dev_free(e->b.d_inverted_label); apg::dev_free(e->b.d_reward);
  apg::dev_free(e->b.d_base_reward); apg::dev_free(e->b.d_loss); apg::dev_free(e->b.d_target_label);
  apg::dev_free(e->b.d_target_xy); apg::dev_free(e->b.d_terminated); apg::dev_free(e->b.d_truncated);
  apg::dev_free(e->b.d_index); apg::dev_free(e->b.d_pos); apg::dev_free(e->b.d_cur_target_xy);
  apg::dev_free(e->b.d_visited); apg::dev_free(e->b.d_stamp_xy); apg::dev_free(e->b.d_stamp_q);
  apg::dev_free(e->d_acc1); apg::dev_free(e->d_acc2); apg::dev_free(e->d_acc_ret); apg::dev_free(e->d_stats);
  apg::dev_free(e->b.d_ep_avg1); apg::dev_free(e->b.d_ep_final1); apg::dev_free(e->b.d_ep_avg2); apg::dev_free(e->b.d_ep_final2);
  apg::dev_free(e->b.d_ep_return); apg::dev_free(e->b.d_first_correct); apg::dev_free(e->b.d_last_incorrect);
  apg::dev_free(e->d_st_idx); apg::dev_free(e->d_st_label); apg::dev_free(e->d_st_inv); apg::dev_free(e->d_err);
  apg::dev_free(e->d_in_action); apg::dev_free(e->d_in_pred);
  for (auto& s : e->ring) {
    apg::dev_free(s.d_idx); apg::dev_free(s.d_pos); apg::dev_free(s.d_target); apg::dev_free(s.d_inv);
    if (s.h_pin) cudaFreeHost(s.h_pin);
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.ready) cudaEventDestroy(s.ready);
  }
  delete e;
}

int apg_image_set_stream(apg_image_env* e, void* cuda_stream) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  if (int rc = join_rng(e)) return rc;
  e->stream = static_cast<cudaStream_t>(cuda_stream);
  return APG_OK;
}

int apg_image_set_bank(apg_image_env* e, const void* bank, const int32_t* labels, int64_t n_items, int on_device) {
  APG_CHECK(e && bank && labels && n_items > 0, APG_ERR_INVALID, "apg_image_set_bank: bad argument");
  APG_CHECK(n_items < (1ll << 31), APG_ERR_INVALID, "bank too large");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  if (e->own_bank) {
    cudaFree(e->d_bank);
    cudaFree(e->d_bank_labels);
  }
  e->d_bank = nullptr; e->d_bank_labels = nullptr; e->own_bank = false;
  const size_t es = e->cfg.bank_dtype == APG_BANK_U8 ? 1 : 4;
  const size_t row = (size_t)e->cfg.width * e->cfg.channels * es;   // tight row as supplied
  const size_t rows_total = (size_t)n_items * e->cfg.height;
  if (on_device) {
    APG_CHECK(row == (size_t)e->bank_pitch, APG_ERR_INVALID,
              "adopting a device bank needs 16-byte aligned rows (W*C*elem = %zu); upload from the host instead", row);
    APG_CHECK((reinterpret_cast<uintptr_t>(bank) & 15) == 0, APG_ERR_INVALID, "device bank must be 16-byte aligned");
    e->d_bank = const_cast<void*>(bank);
    e->d_bank_labels = const_cast<int32_t*>(labels);
  } else {
    // rows are re-pitched to a multiple of 16 bytes so that every footprint can be staged with 16-byte cp.async
    APG_CUDA(cudaMalloc(&e->d_bank, rows_total * e->bank_pitch + 64));
    e->own_bank = true;
    APG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->d_bank_labels), (size_t)n_items * 4));
    APG_CUDA(cudaMemset(e->d_bank, 0, rows_total * e->bank_pitch + 64));
    APG_CUDA(cudaMemcpy2D(e->d_bank, e->bank_pitch, bank, row, row, rows_total, cudaMemcpyHostToDevice));
    APG_CUDA(cudaMemcpy(e->d_bank_labels, labels, (size_t)n_items * 4, cudaMemcpyHostToDevice));
  }
  e->n_items = n_items;
  if (e->tma.spec) apg_img::tma_encode(e->tma, e->d_bank, e->bank_pitch, (long long)n_items, e->cfg.height);
  return configure_rng(e);
}

int apg_image_reset(apg_image_env* e, const int64_t* h_idx, const double* h_pos, const float* h_target,
                    const uint8_t* h_invert) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CHECK(e->d_bank, APG_ERR_STATE, "apg_image_reset before apg_image_set_bank");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  // an explicit reset drops queued autoreset payloads (the reference re-seeds its streams on reset(seed))
  if (int rc = join_rng(e)) return rc;
  e->ring_head = 0;
  e->ring_count = 0;
  apg_image_env::Slot& s = e->ring[0];
  int rc = stage_payload(e, s, h_idx, h_pos, h_target, h_invert);
  if (rc != APG_OK) return rc;
  rc = launch(e, 2, nullptr, nullptr, &s, nullptr, nullptr);
  if (rc != APG_OK) return rc;
  e->has_reset = true;
  e->needs_reset = false;
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

int apg_image_queue_reset_payload(apg_image_env* e, const int64_t* h_idx, const double* h_pos, const float* h_target,
                                  const uint8_t* h_invert) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CHECK(e->d_bank, APG_ERR_STATE, "payload queued before apg_image_set_bank");
  APG_CHECK(e->ring_count < (int)e->ring.size(), APG_ERR_STATE, "reset payload ring is full (%d pending)", e->ring_count);
  APG_CUDA(cudaSetDevice(e->cfg.device));
  apg_image_env::Slot& s = e->ring[(e->ring_head + e->ring_count) % (int)e->ring.size()];
  int rc = stage_payload(e, s, h_idx, h_pos, h_target, h_invert);
  if (rc != APG_OK) return rc;
  s.drawn = false;
  e->ring_count += 1;
  return APG_OK;
}

int apg_image_pending_payloads(apg_image_env* e) { return e ? e->ring_count : APG_ERR_INVALID; }

static void to_internal(const apg_pcg64_state* s, apg::Pcg64& g) {
  g.state_hi = s->state_hi; g.state_lo = s->state_lo; g.inc_hi = s->inc_hi; g.inc_lo = s->inc_lo;
  g.has_uint32 = s->has_uint32; g.uinteger = s->uinteger;
}
static void from_internal(const apg::Pcg64& g, apg_pcg64_state* s) {
  s->state_hi = g.state_hi; s->state_lo = g.state_lo; s->inc_hi = g.inc_hi; s->inc_lo = g.inc_lo;
  s->has_uint32 = g.has_uint32; s->uinteger = g.uinteger;
}

int apg_image_set_streams(apg_image_env* e, const apg_pcg64_state* g0, const apg_pcg64_state* g1,
                          const apg_pcg64_state* g2, int64_t global_num_envs, int64_t env_offset) {
  APG_CHECK(e && g0 && g1 && g2, APG_ERR_INVALID, "apg_image_set_streams: null argument");
  APG_CHECK(env_offset >= 0 && env_offset + e->cfg.num_envs <= global_num_envs, APG_ERR_INVALID,
            "shard [%lld, %lld) exceeds the global batch of %lld envs", (long long)env_offset,
            (long long)(env_offset + e->cfg.num_envs), (long long)global_num_envs);
  APG_CHECK(e->ring_count == 0, APG_ERR_STATE, "cannot replace the reset streams while %d payloads drawn from the old ones are queued",
            e->ring_count);
  APG_CUDA(cudaSetDevice(e->cfg.device));
  to_internal(g0, e->g0); to_internal(g1, e->g1); to_internal(g2, e->g2);
  e->global_n = global_num_envs; e->env_offset = env_offset;
  e->have_streams = true;
  return configure_rng(e);  // device mode: scratch allocation + state upload happen here, off the step path
}

int apg_image_get_streams(apg_image_env* e, apg_pcg64_state* g0, apg_pcg64_state* g1, apg_pcg64_state* g2) {
  APG_CHECK(e && g0 && g1 && g2, APG_ERR_INVALID, "apg_image_get_streams: null argument");
  APG_CHECK(e->have_streams, APG_ERR_STATE, "no reset streams set");
  {
    // payloads drawn ahead of time have advanced the streams past the reference's position: drop them and rewind
    APG_CUDA(cudaSetDevice(e->cfg.device));
    int rc = rewind_streams(e);
    if (rc != APG_OK) return rc;
    rc = streams_to_host(e);
    if (rc != APG_OK) return rc;
  }
  from_internal(e->g0, g0); from_internal(e->g1, g1); from_internal(e->g2, g2);
  return APG_OK;
}

int apg_image_prefetch_payloads(apg_image_env* e, int32_t count) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CHECK(e->d_bank, APG_ERR_STATE, "payload prefetch before apg_image_set_bank");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  for (int i = 0; i < count && e->ring_count < (int)e->ring.size(); ++i) {
    int rc = generate_payload(e);
    if (rc != APG_OK) return rc;
  }
  return APG_OK;
}

int apg_image_step(apg_image_env* e, const float* d_action, const float* d_pred) {
  APG_CHECK(e && d_action && d_pred, APG_ERR_INVALID, "apg_image_step: null argument");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "step before reset");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  return step_once(e, d_action, d_pred, nullptr, nullptr);
}

int apg_image_rollout(apg_image_env* e, const float* d_actions, const float* d_preds, int32_t n_steps,
                      float* d_rewards, uint8_t* d_dones) {
  APG_CHECK(e && d_actions && d_preds && n_steps >= 0, APG_ERR_INVALID, "apg_image_rollout: bad argument");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "rollout before reset");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  // count the autoreset steps inside the window up front so that a rollout never stops half way
  {
    int t = e->t, resets = 0;
    bool need = e->needs_reset;
    for (int k = 0; k < n_steps; ++k) {
      if (need) { resets++; t = 0; need = false; }
      else { t++; need = t >= e->cfg.step_limit; }
    }
    APG_CHECK(e->have_streams || resets <= e->ring_count, APG_ERR_NEED_PAYLOAD,
              "rollout of %d steps needs %d reset payloads, %d queued", n_steps, resets, e->ring_count);
  }
  if (persistent(e)) return rollout_window(e, d_actions, d_preds, n_steps, d_rewards, d_dones);
  const size_t N = e->cfg.num_envs;
  for (int k = 0; k < n_steps; ++k) {
    int rc = step_once(e, d_actions + (size_t)k * N * 2, d_preds + (size_t)k * N * e->pred_width,
                       d_rewards ? d_rewards + (size_t)k * N : nullptr, d_dones ? d_dones + (size_t)k * N : nullptr);
    if (rc != APG_OK) return rc;
  }
  return APG_OK;
}

int apg_image_step_host(apg_image_env* e, const float* h_action, const float* h_pred, float* h_reward,
                        uint8_t* h_terminated, uint8_t* h_truncated, float* h_glimpse) {
  APG_CHECK(e && h_action && h_pred, APG_ERR_INVALID, "apg_image_step_host: null argument");
  APG_CHECK(e->has_reset, APG_ERR_STATE, "step before reset");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  const size_t N = e->cfg.num_envs, K = e->pred_width;
  // Results go straight into page-locked caller buffers (the kernels store to their device aliases); pageable ones get
  // device-to-host copies after the last kernel.  With h_glimpse page-locked the observation of this step is written
  // to the caller's buffer INSTEAD of the engine's device buffer.
  float* m_reward = mapped_alias(h_reward);
  uint8_t* m_term = mapped_alias(h_terminated);
  float* m_glimpse = mapped_alias(h_glimpse);
  if (reinterpret_cast<uintptr_t>(m_glimpse) & 15) m_glimpse = nullptr;  // the special kernels store 128 bits at a time
  // The batch is cut into chunks: the host-to-device copy of chunk c+1 (copy stream) runs while chunk c is stepped
  // (the special-purpose kernels take an environment range; the generic staged kernels step the whole batch at once)
  const bool ranged = persistent(e);
  int chunks = 1;
  // (measured on B200: two chunks for the TMA kernel; the thread-per-environment kernel has a launch cost that a
  //  second chunk would pay twice, its single-step launch is shorter than the copy it would hide)
  if (ranged && !e->small.spec) chunks = N >= 16384 ? 2 : 1;
  if (const char* ov = getenv("APG_HOST_CHUNKS")) {  // developer switch: number of pipeline chunks of the host step
    const int v = atoi(ov);
    if (ranged && v >= 1 && v <= 8) chunks = v;
  }
  const size_t per = ((N + chunks - 1) / chunks + 1023) & ~(size_t)1023;  // chunk boundaries: multiples of 1024 envs
  float* const glimpse_dev = e->b.d_glimpse;
  if (m_glimpse) e->b.d_glimpse = m_glimpse;
  e->hold = true;
  int rc = APG_OK;
  for (int c = 0; c < chunks && rc == APG_OK; ++c) {
    const size_t lo = (size_t)c * per, hi = lo + per < N ? lo + per : N;
    if (lo >= hi) break;
    // a single chunk has nothing to overlap with: its copies go on the step stream itself (no event, no cross-stream wait)
    cudaStream_t cs = chunks == 1 ? e->stream : e->copy_stream;
    cudaError_t ce = cudaMemcpyAsync(e->d_in_action + lo * 2, h_action + lo * 2, (hi - lo) * 2 * sizeof(float),
                                     cudaMemcpyHostToDevice, cs);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(e->d_in_pred + lo * K, h_pred + lo * K, (hi - lo) * K * sizeof(float), cudaMemcpyHostToDevice, cs);
    if (ce == cudaSuccess && chunks > 1) ce = cudaEventRecord(e->ev_in[c], e->copy_stream);
    if (ce == cudaSuccess && chunks > 1) ce = cudaStreamWaitEvent(e->stream, e->ev_in[c], 0);
    if (ce != cudaSuccess) {
      rc = apg::fail(APG_ERR_CUDA, "apg_image_step_host: %s", cudaGetErrorString(ce));
      break;
    }
    e->r0 = (int)lo; e->r1 = (int)hi;
    rc = step_once(e, e->d_in_action, e->d_in_pred, m_reward, m_term);
  }
  e->hold = false;
  e->r0 = 0; e->r1 = -1;
  e->b.d_glimpse = glimpse_dev;
  if (rc != APG_OK) {
    cudaStreamSynchronize(e->copy_stream);
    cudaStreamSynchronize(e->stream);
    return rc;
  }
  commit_step(e);
  if (h_reward && !m_reward) APG_CUDA(cudaMemcpyAsync(h_reward, e->b.d_reward, N * 4, cudaMemcpyDeviceToHost, e->stream));
  if (h_terminated && !m_term) APG_CUDA(cudaMemcpyAsync(h_terminated, e->b.d_terminated, N, cudaMemcpyDeviceToHost, e->stream));
  if (h_glimpse && !m_glimpse) APG_CUDA(cudaMemcpyAsync(h_glimpse, e->b.d_glimpse, N * e->P * 4, cudaMemcpyDeviceToHost, e->stream));
  if (h_truncated) memset(h_truncated, 0, N);  // image environments never truncate (image_perception_module.py:212)
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

int apg_image_get_buffers(apg_image_env* e, apg_image_buffers* out) {
  APG_CHECK(e && out, APG_ERR_INVALID, "null argument");
  *out = e->b;
  out->stamp_count = e->needs_reset ? e->cfg.step_limit : e->t;
  out->elapsed = e->t;
  out->needs_reset = e->needs_reset ? 1 : 0;
  return APG_OK;
}

int apg_image_get_stats(apg_image_env* e, double* h_stats, int clear) {
  APG_CHECK(e && h_stats, APG_ERR_INVALID, "null argument");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaMemcpyAsync(h_stats, e->d_stats, APG_NUM_STATS * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (clear) APG_CUDA(cudaMemsetAsync(e->d_stats, 0, APG_NUM_STATS * sizeof(double), e->stream));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

int apg_image_poll_error(apg_image_env* e, uint32_t* flags, int clear) {
  APG_CHECK(e && flags, APG_ERR_INVALID, "null argument");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaMemcpyAsync(flags, e->d_err, 4, cudaMemcpyDeviceToHost, e->stream));
  if (clear) APG_CUDA(cudaMemsetAsync(e->d_err, 0, 4, e->stream));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

// internal: lets csrc/apg_unique.cu read the bank of a handle (not part of the public header)
int apg_image_bank_info(apg_image_env* e, const void** d_bank, int32_t* bank_pitch, int64_t* n_items, apg_image_cfg* cfg,
                        void** stream) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  *d_bank = e->d_bank; *bank_pitch = e->bank_pitch; *n_items = e->n_items; *cfg = e->cfg; *stream = e->stream;
  return APG_OK;
}

int apg_image_variant(apg_image_env* e) {
  if (!e) return APG_ERR_INVALID;
  if (e->tma.spec && e->tma.have_map) return 3;
  if (e->small.spec) return 4;
  return e->thread_blend ? 1 : 0;
}

int apg_image_sync(apg_image_env* e) {
  APG_CHECK(e, APG_ERR_INVALID, "null env");
  APG_CUDA(cudaSetDevice(e->cfg.device));
  APG_CUDA(cudaStreamSynchronize(e->stream));
  return APG_OK;
}

}  // extern "C"
