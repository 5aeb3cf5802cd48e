/*
 * CPU ORACLE (test infrastructure, NOT product code) for the LIDAR hot path of ap_gym 0.5.0.
 *
 * Plain-C restatement of /root/reference/ap_gym/envs/lidar_localization2d.py:
 *   __lidar_scan ................. :496-536   (shapely LineString.intersection(union of unit boxes) + type dispatch)
 *   _step pose integration/slide . :317-375
 *   target / base reward ......... :330, :377
 *   __get_obs .................... :238-277   (lidar, odometry with the __initial_pos alias quirk of :305/:343)
 * Only tests/, __graft_entry__.smoke() and the CPU-baseline legs of bench.py may load this; the product never does.
 *
 * PARITY STATUS: the python logic above is pinned against the UNMODIFIED reference run here on the test-only
 * exact-geometry `shapely` stand-in (oracle/shim/shapely, goldens in tests/golden/lidar_*.npz).  The geometry
 * itself lives in a third-party dependency that is absent from /root/reference and from this image:
 * shapely (un-pinned, pyproject.toml:20; needs >= 2.0) i.e. GEOS OverlayNG.  Its published behaviour is restated
 * as exact geometry on CLOSED unit cells; parity against real GEOS is UNPINNED (no golden vector exists anywhere).
 *
 * Geometry model (identical in oracle/shim/shapely/__init__.py and in the CUDA kernel):
 *   the segment p->t (float64 copies of the float32 end points) is walked cell by cell; the parameter of every
 *   grid-line crossing is (g - p) * (1 / d) with the reciprocal taken once per segment and axis; crossings with
 *   tX == tY are lattice corners where the
 *   two side cells touch the segment in a single point.  The walk yields the parameter intervals of the
 *   intersection with the union of closed wall cells: `n_pos` intervals of positive length, `n_zero` isolated
 *   points, and the entry vertex of the first one (crossed coordinate snapped to the integer grid line).
 *   result type      reference branch (:502-528)                               distance / contact
 *   none             else                                                      |t-p| (f32), no contact
 *   1 pos            LineString                                                f32(max(|c-p|_f64 - 1e-3, 0)), c (f64)
 *   1 zero           Point                                                     0, contact = pos
 *   >=2 same kind    MultiLineString / MultiPoint: vertices cast to float32    max(|f32(c)-p|_f32 - 1e-3f, 0), f32(c)
 *   mixed            GeometryCollection -> else                                |t-p| (f32), no contact
 *
 * Floating point: compile with -O2 -ffp-contract=off.  np.linalg.norm(x) without axis is sqrt(dot(x,x)); on the
 * machine that generated the goldens (numpy 2.3.5 / OpenBLAS 0.3.30 Haswell kernels) the float32 dot of a
 * 2-vector is x0*x0 + x1*x1 with separate roundings, the float64 dot is fma(x1, x1, x0*x0).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define EPS_D 1e-3
#define EPS_F 1e-3f

static inline int is_wall(const uint32_t* bits, int W, int H, int words, int x, int y) {
  if (x < 0 || y < 0 || x >= W || y >= H) return 0; /* the union only has in-map boxes */
  return (int)((bits[(size_t)y * words + (x >> 5)] >> (x & 31)) & 1u);
}

static inline float norm2f(float a, float b) { return sqrtf(a * a + b * b); }

/* One ray.  Returns the float32 distance, writes the contact cell (x, y) or (-1, -1). */
float apg_oracle_scan_one(const uint32_t* bits, int W, int H, int words, float posx, float posy, float tx,
                          float ty, int32_t* cell) {
  const double px = posx, py = posy;
  const double dx = (double)tx - px, dy = (double)ty - py;
  const int sx = dx > 0 ? 1 : (dx < 0 ? -1 : 0), sy = dy > 0 ? 1 : (dy < 0 ? -1 : 0);
  const double inv_dx = sx != 0 ? 1.0 / dx : 0.0, inv_dy = sy != 0 ? 1.0 / dy : 0.0;
  const int px_int = (px == floor(px)), py_int = (py == floor(py));
  int cx = (int)floor(px), cy = (int)floor(py);
  if (px_int && sx < 0) cx -= 1;
  if (py_int && sy < 0) cy -= 1;
  /* a ray running exactly along a grid line touches the cells on both sides of it */
  const int nbx = (sx == 0 && px_int), nby = (sy == 0 && py_int);
#define WALL(X, Y) (is_wall(bits, W, H, words, (X), (Y)) || (nbx && is_wall(bits, W, H, words, (X)-1, (Y))) || \
                    (nby && is_wall(bits, W, H, words, (X), (Y)-1)))
  int n_pos = 0, n_zero = 0, have_first = 0, in_iv = 0;
  double iv_t0 = 0.0, fex = 0.0, fey = 0.0;

  if (WALL(cx, cy)) {
    in_iv = 1; iv_t0 = 0.0; have_first = 1; fex = px; fey = py;
  } else {
    /* start point on a grid line: the cells behind it touch the segment at t = 0 */
    int touch = 0;
    const int bx = cx - sx, by = cy - sy;
    if (px_int && sx != 0) touch |= is_wall(bits, W, H, words, bx, cy) || (nby && is_wall(bits, W, H, words, bx, cy - 1));
    if (py_int && sy != 0) touch |= is_wall(bits, W, H, words, cx, by) || (nbx && is_wall(bits, W, H, words, cx - 1, by));
    if (px_int && py_int && sx != 0 && sy != 0) touch |= is_wall(bits, W, H, words, bx, by);
    if (touch) { n_zero = 1; have_first = 1; fex = px; fey = py; }
  }

  for (;;) {
    double tX = INFINITY, tY = INFINITY, gx = 0.0, gy = 0.0;
    if (sx != 0) { gx = (double)(sx > 0 ? cx + 1 : cx); tX = (gx - px) * inv_dx; }
    if (sy != 0) { gy = (double)(sy > 0 ? cy + 1 : cy); tY = (gy - py) * inv_dy; }
    const double tN = tX < tY ? tX : tY;
    if (!(tN <= 1.0)) break; /* the segment ends inside the current cell */
    const int stepx = tX <= tY, stepy = tY <= tX;
    const int ncx = cx + (stepx ? sx : 0), ncy = cy + (stepy ? sy : 0);
    const int next_wall = WALL(ncx, ncy);
    const int side_wall = (stepx && stepy) && (is_wall(bits, W, H, words, ncx, cy) || is_wall(bits, W, H, words, cx, ncy));
    if (in_iv) {
      if (!next_wall) { /* the interval is closed at tN (a touching side cell adds nothing to a closed interval) */
        if (tN > iv_t0) n_pos++; else n_zero++;
        in_iv = 0;
      }
    } else if (next_wall || side_wall) {
      if (!have_first) {
        have_first = 1;
        fex = stepx ? gx : px + tN * dx;
        fey = stepy ? gy : py + tN * dy;
      }
      if (next_wall) { in_iv = 1; iv_t0 = tN; } else n_zero++;
    }
    cx = ncx; cy = ncy;
  }
  if (in_iv) { if (iv_t0 < 1.0) n_pos++; else n_zero++; }
#undef WALL

  const float full = norm2f(tx - posx, ty - posy); /* np.linalg.norm(target - pos), float32 */
  float dist;
  int kind; /* 0 none, 1 float64 contact, 2 float32 contact */
  double c64x = 0, c64y = 0;
  float c32x = 0, c32y = 0;
  if (n_pos + n_zero == 0 || (n_pos > 0 && n_zero > 0)) {
    dist = full; kind = 0;
  } else if (n_pos == 1) {
    const double ddx = fex - px, ddy = fey - py;
    const double nrm = sqrt(fma(ddy, ddy, ddx * ddx));
    dist = (float)fmax(nrm - EPS_D, 0.0);
    kind = 1; c64x = fex; c64y = fey;
  } else if (n_zero == 1) {
    dist = 0.0f; kind = 2; c32x = posx; c32y = posy;
  } else {
    c32x = (float)fex; c32y = (float)fey;
    const float d = norm2f(c32x - posx, c32y - posy);
    dist = fmaxf(d - EPS_F, 0.0f);
    kind = 2;
  }
  cell[0] = -1; cell[1] = -1;
  if (kind != 0 && dist < full) {
    if (kind == 1) {
      double fx = floor(c64x), fy = floor(c64y);
      if (fabs(fx - c64x) < 1e-5 && tx < posx) fx -= 1.0;
      if (fabs(fy - c64y) < 1e-5 && ty < posy) fy -= 1.0;
      cell[0] = (int32_t)fx; cell[1] = (int32_t)fy;
    } else {
      float fx = floorf(c32x), fy = floorf(c32y);
      if (fabsf(fx - c32x) < 1e-5f && tx < posx) fx -= 1.0f;
      if (fabsf(fy - c32y) < 1e-5f && ty < posy) fy -= 1.0f;
      cell[0] = (int32_t)fx; cell[1] = (int32_t)fy;
    }
  }
  return dist;
}

void apg_oracle_scan(const uint32_t* bits, int W, int H, int words, const float* pos, const float* targets, int n,
                     float* dist, int32_t* cells) {
  for (int i = 0; i < n; ++i)
    dist[i] = apg_oracle_scan_one(bits, W, H, words, pos[0], pos[1], targets[2 * i], targets[2 * i + 1], cells + 2 * i);
}

/* LIDARLocalization2DEnv._step without the observation (:317-389).  pos/init are updated in place; `aliased`
 * is 1 while __pos and __initial_pos are the same array (from reset until the first clip rebinding).
 * Returns the out-of-bounds `terminated` flag. */
int apg_oracle_move(const uint32_t* bits, int W, int H, int words, float* pos, float* init, int* aliased,
                    const float* action, const float* pred, float* out_target, float* out_base, float* out_loss) {
  const float Wf = (float)W, Hf = (float)H;
  const float lastx = pos[0], lasty = pos[1];
  float a0 = action[0], a1 = action[1];
  *out_base = 0.1f - 1e-3f * (a0 * a0 + a1 * a1);
  const float mag = norm2f(a0, a1);
  if (mag > 1.0f) { a0 = a0 / mag; a1 = a1 / mag; }
  const float tx = pos[0] + a0, ty = pos[1] + a1;
  float dirx = tx - pos[0], diry = ty - pos[1];
  const float total = norm2f(dirx, diry);
  if (total > 0.0f) {
    dirx /= total; diry /= total;
    int32_t cell[2];
    const float d = apg_oracle_scan_one(bits, W, H, words, pos[0], pos[1], tx, ty, cell);
    pos[0] += dirx * d; pos[1] += diry * d;
    const float rem = total - d;
    if (rem > 1e-5f) {
      const float rvx = dirx * rem, rvy = diry * rem;
      const int kx = rvx > 1e-5f, ky = rvy > 1e-5f;
      if (kx || ky) {
        /* np.eye(2) * remaining_vec: rows (v0, 0), (0, v1); with one surviving component v: (v, 0), (0, v) */
        const float v0 = (kx && ky) ? rvx : (kx ? rvx : rvy);
        const float v1 = (kx && ky) ? rvy : (kx ? rvx : rvy);
        const float d0 = apg_oracle_scan_one(bits, W, H, words, pos[0], pos[1], pos[0] + v0, pos[1] + 0.0f, cell);
        const float d1 = apg_oracle_scan_one(bits, W, H, words, pos[0], pos[1], pos[0] + 0.0f, pos[1] + v1, cell);
        if (d0 > 0.0f) pos[0] += (v0 / norm2f(v0, 0.0f)) * d0, pos[1] += (0.0f / norm2f(v0, 0.0f)) * d0;
        else pos[0] += (0.0f / norm2f(0.0f, v1)) * d1, pos[1] += (v1 / norm2f(0.0f, v1)) * d1;
      }
    }
  }
  if (*aliased) { init[0] = pos[0]; init[1] = pos[1]; } /* in-place += on the shared array */
  const int oob = (pos[0] < 0.0f) || (pos[1] < 0.0f) || (pos[0] >= Wf) || (pos[1] >= Hf);
  pos[0] = fminf(fmaxf(pos[0], 0.0f), Wf); pos[1] = fminf(fmaxf(pos[1], 0.0f), Hf);
  *aliased = 0; /* np.clip rebinds __pos */
  out_target[0] = lastx / Wf * 2.0f - 1.0f;
  out_target[1] = lasty / Hf * 2.0f - 1.0f;
  const float e0 = pred[0] - out_target[0], e1 = pred[1] - out_target[1];
  *out_loss = ((e0 * e0 + e1 * e1) / 2.0f) * 3.0f + 0.0f;
  return oob;
}

/* __get_obs (:238-277): lidar (B), odometry (2), hit cells (B,2). dirs = lidar_directions (B,2) float32. */
void apg_oracle_obs(const uint32_t* bits, int W, int H, int words, const float* pos, const float* init,
                    const float* dirs, int B, float range, float* lidar, float* odo, int32_t* cells, float* dists) {
  for (int b = 0; b < B; ++b) {
    const float d = apg_oracle_scan_one(bits, W, H, words, pos[0], pos[1], pos[0] + dirs[2 * b], pos[1] + dirs[2 * b + 1],
                                        cells + 2 * b);
    dists[b] = d;
    lidar[b] = fminf(fmaxf(d / range, -1.0f), 1.0f);
  }
  const float Wf = (float)W, Hf = (float)H;
  odo[0] = ((pos[0] - init[0]) - (-Wf)) / (Wf - (-Wf)) * 2.0f - 1.0f;
  odo[1] = ((pos[1] - init[1]) - (-Hf)) / (Hf - (-Hf)) * 2.0f - 1.0f;
}

/* Batched driver used by oracle/lidar_oracle.py and the CPU baseline: one vector step for `n` environments that
 * are NOT on their autoreset step (mask[i] != 0 -> stepped).  maps: (M, H, words) bank, map_idx (n). */
void apg_oracle_step_batch(const uint32_t* bank, int W, int H, int words, const int64_t* map_idx, int n,
                           const uint8_t* mask, float* pos, float* init, int32_t* aliased, const float* actions,
                           const float* preds, const float* dirs, int B, float range, float* lidar, float* odo,
                           int32_t* cells, float* dists, float* target, float* base, float* loss, uint8_t* oob) {
  for (int i = 0; i < n; ++i) {
    if (!mask[i]) continue;
    const uint32_t* bits = bank + (size_t)map_idx[i] * H * words;
    int al = aliased[i];
    oob[i] = (uint8_t)apg_oracle_move(bits, W, H, words, pos + 2 * i, init + 2 * i, &al, actions + 2 * i, preds + 2 * i,
                                      target + 2 * i, base + i, loss + i);
    aliased[i] = al;
    apg_oracle_obs(bits, W, H, words, pos + 2 * i, init + 2 * i, dirs, B, range, lidar + (size_t)i * B, odo + 2 * i,
                   cells + (size_t)i * B * 2, dists + (size_t)i * B);
  }
}

void apg_oracle_obs_batch(const uint32_t* bank, int W, int H, int words, const int64_t* map_idx, int n,
                          const uint8_t* mask, const float* pos, const float* init, const float* dirs, int B,
                          float range, float* lidar, float* odo, int32_t* cells, float* dists) {
  for (int i = 0; i < n; ++i) {
    if (!mask[i]) continue;
    const uint32_t* bits = bank + (size_t)map_idx[i] * H * words;
    apg_oracle_obs(bits, W, H, words, pos + 2 * i, init + 2 * i, dirs, B, range, lidar + (size_t)i * B, odo + 2 * i,
                   cells + (size_t)i * B * 2, dists + (size_t)i * B);
  }
}
