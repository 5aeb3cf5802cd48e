/*
 * apg_b200.h -- C-ABI of the B200-native batched stepping engine for ap_gym (active-perception-gym 0.5.0).
 *
 * This is the drop-in boundary for the reference's vector-env hot path.  The reference is pure Python
 * (no FFI of its own), so every entry point names the Python interface it replaces; the ctypes binding a
 * maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; pointers named d_* are DEVICE pointers, h_* are HOST pointers;
 *   - every function returns 0 (APG_OK) or a negative apg_status; apg_last_error() gives the text;
 *   - all work is queued on the handle's CUDA stream (default: the legacy default stream, which is also
 *     PyTorch's default) and is asynchronous unless the function name ends in _host or says it syncs;
 *   - a handle is not thread-safe (the reference env is single-threaded too).
 *   - there is NO CPU fallback: every entry point fails with APG_ERR_CUDA when no device is usable.
 */
#ifndef APG_B200_H
#define APG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APG_VERSION 100 /* 0.1.0 */

typedef enum {
  APG_OK = 0,
  APG_ERR_INVALID = -1,      /* bad argument / configuration (reference: ValueError at construction) */
  APG_ERR_CUDA = -2,         /* CUDA runtime failure, or no device */
  APG_ERR_STATE = -3,        /* call order violated, e.g. step before reset */
  APG_ERR_NEED_PAYLOAD = -4, /* an autoreset step is due but no reset payload has been queued */
  APG_ERR_NAN_ACTION = -5,   /* reference: ValueError("NaN values detected in action.")     image_perception_module.py:204-205, lidar_localization2d.py:318-319 */
  APG_ERR_NAN_PRED = -6,     /* reference: ValueError("NaN values detected in prediction.") image_perception_module.py:194-195, lidar_localization2d.py:320-321 */
  APG_ERR_OUT_OF_GRID = -7   /* reference: scipy RegularGridInterpolator ValueError (bounds_error=True) */
} apg_status;

/* sticky device-side error bits, see apg_*_poll_error */
#define APG_FLAG_NAN_ACTION 1u
#define APG_FLAG_NAN_PRED 2u
#define APG_FLAG_OUT_OF_GRID 4u
#define APG_FLAG_MAPGEN 8u /* a procedural map exceeded the generator's limits (the reference raises for such parameters) */

/* procedural map datasets of apg_lidar_set_map_generator */
#define APG_MAPGEN_ROOMS 1 /* FloorMapDatasetRooms(width, height, max_rooms, door_width) */
#define APG_MAPGEN_MAZE 2  /* FloorMapDatasetMaze(width, height, branching_prob) */

int apg_version(void);
const char* apg_last_error(void);
/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
int64_t apg_launch_count(void);
/* sizeof() of the five structs of this header as compiled into the library, in the order apg_image_cfg,
 * apg_image_buffers, apg_lidar_cfg, apg_lidar_buffers, apg_pcg64_state: lets a binding verify its own declarations
 * (tests/test_capi_symbols.py does it for the bundled ctypes layer).  No GPU needed. */
void apg_struct_sizes(int32_t out[5]);

/* ------------------------------------------------------------------------------------------------------
 * Image environments.  Replaces, for one shard of `num_envs` environments:
 *   ImagePerceptionModule.{reset,step,get_glimpse}   ap_gym/envs/image/image_perception_module.py:120-331
 *   ImageClassificationVectorEnv._step               ap_gym/envs/image_classification.py:113-127
 *   ImageLocalizationVectorEnv._step                 ap_gym/envs/image_localization.py:151-181
 *   ActivePerceptionVectorEnv.step (loss + reward)   ap_gym/active_perception_vector_env.py:84-111
 *   CrossEntropyLossFn/MSELossFn(.normalized).numpy  ap_gym/loss_fn.py:100-110,212-220,261-267
 * ------------------------------------------------------------------------------------------------------ */
typedef struct apg_image_env apg_image_env;

#define APG_TASK_CLASSIFICATION 0
#define APG_TASK_LOCALIZATION 1
#define APG_BANK_U8 0  /* pixels stored as uint8, converted in-kernel with the reference's f32(u8)/255 */
#define APG_BANK_F32 1 /* pixels stored as float32 in [0,1] */

typedef struct {
  int32_t device;                 /* CUDA device ordinal */
  int32_t num_envs;               /* N of this shard */
  int32_t task;                   /* APG_TASK_* */
  int32_t height, width, channels;/* image bank geometry (H, W, C), C in {1, 3} */
  int32_t sensor_h, sensor_w;     /* ImagePerceptionConfig.sensor_size (S0, S1) */
  double sensor_scale;            /* ImagePerceptionConfig.sensor_scale */
  double max_step_length[2];      /* ImagePerceptionConfig.max_step_length broadcast to (x, y) */
  int32_t step_limit;             /* ImagePerceptionConfig.step_limit (T) */
  int32_t num_classes;            /* K: dataset.num_classes (classification logits width) */
  int32_t bank_dtype;             /* APG_BANK_* */
  int32_t randomly_invert_labels; /* ImagePerceptionConfig.randomly_invert_labels */
  int32_t track_visited;          /* keep the visited bitmask + overlay stamp log (reference: always on) */
  int32_t payload_ring;           /* how many future autoreset payloads may be queued (0 -> 4, max 64) */
  int32_t blend_f32;              /* 0: the reference's float64 blend (bit-exact pixels); 1: float32 blend with weights
                                     rounded once from float64 (pixels within a few float32 ulps, i.e. inside the
                                     1e-5 relative bar; about twice the throughput on large glimpses) */
  int32_t sparse_reward;          /* 1: the reference's `-sparse` ids (SparsifyVectorWrapper, sparsify_wrapper.py:74-88):
                                     reward = base_reward - loss * weight with weight = terminated, i.e. the loss is
                                     only charged by the step that ends the episode.  loss / statistics unchanged. */
} apg_image_cfg;

/* Engine-owned persistent device buffers.  Views stay valid until apg_image_destroy; contents are
 * overwritten by the next reset/step.  Shapes use N = num_envs, P = S0*S1*C. */
typedef struct {
  float* d_glimpse;          /* (N, S0, S1, C)  obs["glimpse"]                                  */
  float* d_target_glimpse;   /* (N, S0, S1, C)  obs["target_glimpse"]   (localization only)     */
  float* d_glimpse_pos;      /* (N, 2)          obs["glimpse_pos"]                              */
  float* d_time_step;        /* (N,)            obs["time_step"]                                */
  int32_t* d_inverted_label; /* (N,)            obs["inverted_label"]   (randomly_invert_labels) */
  float* d_reward;           /* (N,)  base_reward - loss                                         */
  float* d_base_reward;      /* (N,)  info["base_reward"]                                        */
  float* d_loss;             /* (N,)  info["prediction"]["loss"]                                 */
  int32_t* d_target_label;   /* (N,)  info["prediction"]["target"]  (classification)            */
  float* d_target_xy;        /* (N,2) info["prediction"]["target"]  (localization: the scored target) */
  uint8_t* d_terminated;     /* (N,)                                                             */
  uint8_t* d_truncated;      /* (N,)                                                             */
  int64_t* d_index;          /* (N,)  info["index"]                                              */
  /* state (read-only for callers) */
  double* d_pos;             /* (N, 2) float64 normalised sensor position                        */
  float* d_cur_target_xy;    /* (N, 2) current localization target (the one target_glimpse shows) */
  uint32_t* d_visited;       /* (N, H, ceil(W/32)) image-resolution visited bitmask              */
  int32_t* d_stamp_xy;       /* (T, 2, N) overlay stamp log: render-resolution rectangle centres */
  float* d_stamp_q;          /* (T, N)    overlay stamp log: clipped prediction quality          */
  /* per-env episode statistics, valid for envs whose `terminated` flag is set by the step (they replace the
   * python loops of ActiveClassification/RegressionVectorLogWrapper, active_classification_env.py:149-197,
   * active_regression_env.py:204-227): metric 1 = correct_label_prob | euclidean_distance,
   * metric 2 = accuracy | mse; avg over the episode's real steps and the final step's value */
  float* d_ep_avg1;          /* (N,) */
  float* d_ep_final1;        /* (N,) */
  float* d_ep_avg2;          /* (N,) */
  float* d_ep_final2;        /* (N,) */
  float* d_ep_return;        /* (N,) sum of rewards over the episode's real steps */
  int32_t* d_first_correct;  /* (N,) classification: first step with accuracy 1, -1 if none */
  int32_t* d_last_incorrect; /* (N,) classification: last step with accuracy 0, -1 if none */
  int32_t stamp_count;       /* number of valid stamps of the current episode (host mirror)      */
  int32_t elapsed;           /* current time step t of the (shared) episode clock                */
  int32_t needs_reset;       /* 1 when the next step is the NEXT_STEP autoreset step             */
} apg_image_buffers;

int apg_image_create(const apg_image_cfg* cfg, apg_image_env** out);
void apg_image_destroy(apg_image_env* env);
/* Use the given cudaStream_t (as an opaque pointer) for all subsequent work; NULL = legacy default stream. */
int apg_image_set_stream(apg_image_env* env, void* cuda_stream);

/* Upload (h_*) or adopt (d_*, on_device=1, caller keeps ownership and must keep it alive) the image bank:
 * n_items images of (H, W, C) in bank_dtype, and their int32 labels.
 * Replaces the per-reset `dataset.get_data_point_batch(idx)` gather (image_perception_module.py:123-129).
 * For the 64x64x3 uint8 geometry the engine also builds a blocked copy of the bank here (8 rows x 16 bytes per 128-byte
 * line, the layout its step kernel reads): an adopted bank is snapshotted, call this again after changing it. */
int apg_image_set_bank(apg_image_env* env, const void* bank, const int32_t* labels, int64_t n_items, int on_device);

/* Explicit reset: ImagePerceptionModule.reset() + (localization) the target assignment of
 * ImageLocalizationVectorEnv.reset().  The random draws are the caller's (host numpy streams):
 *   h_idx (N) bank indices, h_pos (N,2) float64 start positions, h_target (N,2) float32 targets
 *   (localization, else NULL), h_invert (N) 0/1 (randomly_invert_labels, else NULL). Synchronous copy-in. */
int apg_image_reset(apg_image_env* env, const int64_t* h_idx, const double* h_pos, const float* h_target,
                    const uint8_t* h_invert);
/* Queue the payload of a FUTURE autoreset (same arrays as above; copied, asynchronous H2D).  Payloads are
 * consumed in FIFO order by the autoreset steps; up to cfg.payload_ring may be pending. */
#define APG_PAYLOAD_RING_DEFAULT 4
#define APG_PAYLOAD_RING_MAX 64
int apg_image_queue_reset_payload(apg_image_env* env, const int64_t* h_idx, const double* h_pos,
                                  const float* h_target, const uint8_t* h_invert);
int apg_image_pending_payloads(apg_image_env* env);

/* The reference's three reset-path Generators (g0 env-level, g1 module-level, g2 dataset iterator) as
 * numpy PCG64 states.  Once set, the engine draws the payload of every autoreset itself whenever no payload is queued --
 * ON THE DEVICE, in-stream, bit-identical to numpy (csrc/apg_payload.cu; image_perception_module.py:129-133,159-161,
 * dataset_iterator.py:57, image_localization.py:153-156) -- for the GLOBAL batch of global_num_envs envs, of which this
 * shard is [env_offset, env_offset + num_envs).  The states are uploaded and every scratch buffer is allocated HERE (and
 * in apg_image_set_bank), so that no step ever allocates, synchronises or loads a kernel for the first time; banks
 * whose length would reject too many Lemire draws fall back to drawing on the host.
 * apg_image_get_streams returns the current states (e.g. to continue them in numpy for the next explicit reset); payloads
 * that were drawn ahead of time (apg_image_prefetch_payloads) are dropped and the streams rewound to the state before
 * the oldest of them, so the caller always sees the reference's stream position. */
typedef struct {
  uint64_t state_hi, state_lo; /* numpy bit_generator.state["state"]["state"] (128 bit) */
  uint64_t inc_hi, inc_lo;     /* ...["state"]["inc"] */
  uint32_t has_uint32, uinteger;
} apg_pcg64_state;
int apg_image_set_streams(apg_image_env* env, const apg_pcg64_state* g0, const apg_pcg64_state* g1,
                          const apg_pcg64_state* g2, int64_t global_num_envs, int64_t env_offset);
int apg_image_get_streams(apg_image_env* env, apg_pcg64_state* g0, apg_pcg64_state* g1, apg_pcg64_state* g2);
/* Draw and queue up to `count` autoreset payloads ahead of time from the streams (bounded by the ring). */
int apg_image_prefetch_payloads(apg_image_env* env, int32_t count);

/* One vector step with DEVICE inputs: d_action (N,2) f32, d_pred (N,K) f32 logits or (N,2) f32 coordinates.
 * One fused kernel launch; asynchronous.  Replaces ActivePerceptionVectorEnv.step end to end. */
int apg_image_step(apg_image_env* env, const float* d_action, const float* d_pred);
/* n_steps consecutive steps from stacked device inputs d_actions (n_steps, N, 2), d_preds (n_steps, N, K|2);
 * per-step reward / done flags are additionally written to d_rewards (n_steps, N) f32 and d_dones
 * (n_steps, N) u8 when these are non-NULL.  Autoreset steps consume queued payloads (drawn in-stream when the reset
 * streams are set).  Asynchronous.  Results are those of n_steps calls of apg_image_step; for the shapes served by a
 * persistent kernel (apg_image_variant 3 and 4) a window of up to 128 steps is ONE launch in which every thread /
 * warp keeps its environments' state on chip and walks them through all steps of the window. */
int apg_image_rollout(apg_image_env* env, const float* d_actions, const float* d_preds, int32_t n_steps,
                      float* d_rewards, uint8_t* d_dones);
/* One vector step with HOST inputs/outputs: copies h_action/h_pred to the device, steps, delivers reward (N) f32,
 * terminated/truncated (N) u8 and, when non-NULL, glimpse (N*P) f32 in the caller's buffers, then synchronises.  This
 * is the reference-facing call: same argument meaning as env.step({"action", "prediction"}).
 * Pipelining: the batch is cut into chunks whose host-to-device copies (a second stream) overlap the stepping of the
 * previous chunk.  PAGE-LOCKED result buffers (cudaHostAlloc / cudaHostRegister / torch pinned memory) are written by
 * the kernels directly (posted writes over PCIe, no copy afterwards); a page-locked h_glimpse receives the observation
 * INSTEAD of the engine's device buffer for that step.  Pageable buffers get ordinary device-to-host copies. */
int apg_image_step_host(apg_image_env* env, const float* h_action, const float* h_pred, float* h_reward,
                        uint8_t* h_terminated, uint8_t* h_truncated, float* h_glimpse);

/* Unique-glimpse search of ImageLocalizationVectorEnv.reset(): for each of the n_imgs bank images h_img_idx[i], the
 * top_k grid positions (indices into h_grid, (n_positions, 2) float64 normalised (x, y)) whose glimpse differs most
 * from every other grid glimpse, in descending order of uniqueness -- h_topk (n_imgs, top_k) int32.
 * Replaces ImagePerceptionModule.sample_unique_glimpse_positions up to the random pick
 * (ap_gym/envs/image/image_perception_module.py:253-277).  Synchronous. */
int apg_image_unique_topk(apg_image_env* env, const int64_t* h_img_idx, int64_t n_imgs, const double* h_grid,
                          int32_t n_positions, int32_t top_k, int32_t* h_topk);

int apg_image_get_buffers(apg_image_env* env, apg_image_buffers* out);
/* Episode statistics accumulated on the device since creation / the last clear, summed over this shard's envs
 * (the payload of the final NCCL reduce): h_stats[0] = finished episodes, [1] = sum of per-episode avg metric 1,
 * [2] = sum of final metric 1, [3] = sum of avg metric 2, [4] = sum of final metric 2, [5] = sum of episode
 * returns, [6], [7] reserved (0).  Synchronises the stream. */
#define APG_NUM_STATS 8
int apg_image_get_stats(apg_image_env* env, double* h_stats, int clear);
/* Reads (and with clear!=0 resets) the sticky device error flags; synchronises the stream. */
int apg_image_poll_error(apg_image_env* env, uint32_t* flags, int clear);
int apg_image_sync(apg_image_env* env);
/* Which step-kernel variant the handle selected for its bank / sensor geometry:
 *   0 = generic staged tiles, warp-per-env blend        (any geometry / scale / bank dtype, large glimpses)
 *   1 = generic staged tiles, thread-per-env blend      (any geometry / scale / bank dtype, glimpses <= 96 floats)
 *   3 = TMA-tile persistent kernel, whole rollout windows per launch   (uint8 64x64x3 bank, 10x10 sensor, scale 1:
 *       TinyImageNetLoc / TinyImageNet)
 *   4 = persistent thread-per-env rollout kernel        (uint8 single-channel bank up to 32x32, 5x5 sensor, scale 1,
 *       classification: MNIST) */
int apg_image_variant(apg_image_env* env);

/* ------------------------------------------------------------------------------------------------------
 * LIDAR localisation environments.  One handle = one shard of the vector env that
 * `ap_gym.make_vec("LIDARLoc{Rooms,Maze}[Static]-v0", num_envs=N)` builds, i.e. it replaces
 *   gymnasium.vector.SyncVectorEnv over N python envs (NEXT_STEP autoreset)  ap_gym/envs/registration.py:753-767
 *   TimeLimit(max_episode_steps, issue_termination=True).step/reset          ap_gym/time_limit.py:113-152
 *   ActivePerceptionEnv.step (loss + reward)                                 ap_gym/active_perception_env.py:101-121
 *   LIDARLocalization2DEnv.reset/_step/__get_obs/__lidar_scan                ap_gym/envs/lidar_localization2d.py:238-536
 *   the per-env numpy Generators of the reset path                            ap_gym/envs/lidar_localization2d.py:303-311,547-557
 * ------------------------------------------------------------------------------------------------------ */
typedef struct apg_lidar_env apg_lidar_env;

typedef struct {
  int32_t device;           /* CUDA device ordinal */
  int32_t num_envs;         /* N of this shard */
  int32_t width, height;    /* dataset.map_width / map_height */
  int32_t beam_count;       /* lidar_beam_count */
  float lidar_range;        /* lidar_range */
  int32_t step_limit;       /* TimeLimit max_episode_steps (100 for the registered ids) */
  int32_t static_map;       /* static_map */
  int32_t static_map_index; /* static_map_index */
  int32_t emit_cells;       /* also write the per-beam contact cells and raw distances (parity / rendering) */
  int64_t env_offset;       /* global index of env 0 of this shard: env i is seeded with seed + env_offset + i */
  int32_t sparse_reward;    /* 1: `-sparse` ids (SparsifyWrapper, sparsify_wrapper.py:137-150): the loss is only charged by
                               the step that terminates the env's episode */
  int32_t reserved;
} apg_lidar_cfg;

typedef struct {
  float* d_lidar;        /* (N, B)      obs["lidar"]                                   */
  float* d_odometry;     /* (N, 2)      obs["odometry"]                                */
  float* d_time_step;    /* (N,)        obs["time_step"]                               */
  float* d_map;          /* (N, H, W)   obs["map"] in {0, 1/255} (dynamic maps only; rewritten on reset) */
  float* d_reward;       /* (N,)  0 on autoreset steps                                 */
  float* d_base_reward;  /* (N,)  info["base_reward"]                                  */
  float* d_loss;         /* (N,)  info["prediction"]["loss"]                           */
  float* d_target;       /* (N,2) info["prediction"]["target"]                         */
  uint8_t* d_terminated; /* (N,)                                                       */
  uint8_t* d_truncated;  /* (N,)                                                       */
  uint8_t* d_stepped;    /* (N,)  1 = real step, 0 = autoreset step (gymnasium's `_key` info masks) */
  int64_t* d_map_idx;    /* (N,)  info["map_idx"]                                      */
  float* d_pos;          /* (N,2) state: position in map cells (x, y)                  */
  float* d_initial_pos;  /* (N,2) state: odometry origin                               */
  float* d_ep_avg1;      /* (N,) episode statistics (valid where `terminated`): avg / final euclidean_distance, */
  float* d_ep_final1;    /* (N,)   avg / final mse, episode return -- ActiveRegressionLogWrapper,               */
  float* d_ep_avg2;      /* (N,)   active_regression_env.py:141-157                                             */
  float* d_ep_final2;    /* (N,) */
  float* d_ep_return;    /* (N,) */
  int32_t* d_cells;      /* (N,B,2) contact cell (x, y) or (-1,-1) per beam (emit_cells) */
  float* d_dists;        /* (N,B)   raw float32 scan distances (emit_cells)            */
} apg_lidar_buffers;

int apg_lidar_create(const apg_lidar_cfg* cfg, apg_lidar_env** out);
void apg_lidar_destroy(apg_lidar_env* env);
int apg_lidar_set_stream(apg_lidar_env* env, void* cuda_stream);
/* Upload the bit-packed occupancy bank: h_bits (n_maps, H, ceil(W/32)) uint32, bit x&31 of word x>>5 = wall.
 * Replaces the per-episode dataset fetch + shapely union (lidar_localization2d.py:279-291,298).
 * The bank is kept a second time with a free border of max(ceil(lidar_range) + 2, 7) cells around every map -- the form
 * the step kernel stages in shared memory with one bulk copy per environment. */
int apg_lidar_set_map_bank(apg_lidar_env* env, const uint32_t* h_bits, int64_t n_maps);
/* Procedural-dataset mode, the alternative to a finite bank: the reference's datasets FloorMapDatasetRooms /
 * FloorMapDatasetMaze have len == 2**32 and every (auto)reset fetches dataset[integers(0, 2**32)] drawn from the
 * environment's own DatasetIterator stream (lidar_localization2d.py:297-298,547-557; dataset_iterator.py:27-32).  With a
 * generator set, each step compacts the environments that reset, draws their dataset indices from the per-environment
 * device streams and generates those maps on the device (csrc/apg_rooms.cuh, csrc/apg_maze.cuh: bit-identical to the
 * reference's generators) into per-environment map slots before the step kernel runs; info["map_idx"] is the dataset
 * index.  kind: APG_MAPGEN_ROOMS (max_rooms, door_width) or APG_MAPGEN_MAZE (branching_prob); width / height come from
 * the env's cfg.  Not with static_map (generate the one map with apg_maps_generate_* and upload it as a bank). */
int apg_lidar_set_map_generator(apg_lidar_env* env, int32_t kind, int32_t max_rooms, int32_t door_width,
                                double branching_prob);
/* Upload lidar_directions (B,2) float32 = (cos, sin)(linspace(-pi, pi, B, endpoint=False)) * range, computed by
 * the caller exactly like lidar_localization2d.py:180-187 (numpy float32 cos/sin). */
int apg_lidar_set_beams(apg_lidar_env* env, const float* h_dirs);
/* reset(seed=seed): seeds env i with seed + env_offset + i and resets every env.  Asynchronous. */
int apg_lidar_reset(apg_lidar_env* env, uint64_t seed);
/* reset() without a seed: every env resets now, its random streams continue where they are.  Asynchronous. */
int apg_lidar_reset_all(apg_lidar_env* env);
/* One vector step, DEVICE inputs d_action (N,2), d_pred (N,2) float32.  One kernel launch, asynchronous. */
int apg_lidar_step(apg_lidar_env* env, const float* d_action, const float* d_pred);
int apg_lidar_rollout(apg_lidar_env* env, const float* d_actions, const float* d_preds, int32_t n_steps,
                      float* d_rewards, uint8_t* d_dones);
/* One vector step with HOST buffers in and out (reward f32, terminated/truncated u8, optional lidar); syncs. */
int apg_lidar_step_host(apg_lidar_env* env, const float* h_action, const float* h_pred, float* h_reward,
                        uint8_t* h_terminated, uint8_t* h_truncated, float* h_lidar);
int apg_lidar_get_buffers(apg_lidar_env* env, apg_lidar_buffers* out);
int apg_lidar_get_stats(apg_lidar_env* env, double* h_stats, int clear); /* same layout as apg_image_get_stats */
int apg_lidar_poll_error(apg_lidar_env* env, uint32_t* flags, int clear);
int apg_lidar_sync(apg_lidar_env* env);

/* Procedural maze maps generated ON THE DEVICE, bit-identical to the reference's generator
 * FloorMapDatasetMaze.get_data_point (ap_gym/envs/floor_map/floor_map_dataset_maze.py:24-55; the recursive carve as an
 * explicit stack, the numpy draws of default_rng(index) in the same order): map i of the result is dataset[h_indices[i]],
 * bit-packed like apg_lidar_set_map_bank expects ((n_maps, height, ceil(width/32)) uint32, bit = wall).  Odd sizes up
 * to 255.  Synchronous. */
int apg_maps_generate_maze(int32_t device, const uint64_t* h_indices, int64_t n_maps, int32_t width, int32_t height,
                           double branching_prob, uint32_t* h_bits);

/* Procedural rooms maps generated ON THE DEVICE, bit-identical to FloorMapDatasetRooms.get_data_point
 * (ap_gym/envs/floor_map/floor_map_dataset_rooms.py:25-89): the recursive split_room as an explicit stack of room views,
 * numpy's binomial (inversion), choice(replace=False) (Floyd + shuffle draws) and integers (Lemire) restated on the
 * PCG64 stream of default_rng(index) (csrc/apg_rooms.cuh).  Same output layout as apg_maps_generate_maze.  Square maps
 * (the reference transposes half of them), max_rooms <= 16.  Synchronous. */
int apg_maps_generate_rooms(int32_t device, const uint64_t* h_indices, int64_t n_maps, int32_t width, int32_t height,
                            int32_t max_rooms, int32_t door_width, uint32_t* h_bits);

/* Host-side test hook for the numpy-compatible random streams used on the device reset path (csrc/apg_rng.h):
 * a Generator(PCG64(SeedSequence(seed))) driven by n_ops operations --
 * op 0 next raw uint64, 1 next buffered uint32, 2 integers(0, arg), 3 integers(0, 2**32, endpoint=True),
 * 4 re-seed with SeedSequence(arg), 5 uniform(-1, 1) (float64 bit pattern), 6 advance by arg raw draws.
 * out[i] receives the result of operation i. */
int apg_rng_host_stream(uint64_t seed, int32_t n_ops, const int32_t* op, const uint64_t* arg, uint64_t* out);

/* Test hook for the DEVICE generator of the image environments' autoreset payloads (csrc/apg_payload.cu): draws one
 * payload -- g0.uniform(-1,1,(Ng,2)).astype(f32) [localization], g2.integers(0, n_items, Ng),
 * g1.integers(0,2,Ng)==1 [invert], g1.uniform(-1,1,(Ng,2)); reference: image_localization.py:153-156,
 * dataset_iterator.py:57, image_perception_module.py:131-133,159-161 -- on `device`, copies this shard's slice
 * [env_offset, env_offset + num_envs) to the host arrays and advances the three states exactly as numpy would.
 * APG_ERR_INVALID when n_items rejects too many Lemire draws for the device path (the engine then draws on the host). */
int apg_rng_device_payload(int32_t device, apg_pcg64_state* g0, apg_pcg64_state* g1, apg_pcg64_state* g2,
                           int64_t global_n, int64_t env_offset, int64_t num_envs, uint64_t n_items,
                           int32_t localization, int32_t invert, int64_t* h_idx, double* h_pos, float* h_target,
                           uint8_t* h_inv);

/* ------------------------------------------------------------------------------------------------------
 * DLPack export of engine-owned device memory (north_star: PyTorch only sees DLPack tensors).
 * Returns a malloc'ed DLManagedTensor* (DLPack v0.8 layout) whose deleter frees only the descriptor.
 * dtype_code: 0 int, 1 uint, 2 float (DLDataTypeCode); bits: 8/32/64.
 * ------------------------------------------------------------------------------------------------------ */
void* apg_dlpack_wrap(void* d_ptr, int32_t device, int32_t dtype_code, int32_t bits, int32_t ndim,
                      const int64_t* shape);

#ifdef __cplusplus
}
#endif
#endif /* APG_B200_H */
