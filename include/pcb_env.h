/* pcb_env.h -- C ABI of libpcbgpu: batched RL-PCB placement environments on B200.
 *
 * This is the drop-in boundary for the reset/step hot path of
 * PBozmarov/RL-Environment-for-Component-Placement.  The reference exposes the
 * path as a Python gym-0.22 class per variant:
 *
 *   environment/dummy_env_square.py:37-153             (square)
 *   environment/dummy_env_rectangular.py:152-432       (rectangle)
 *   environment/dummy_env_rectangular_pin.py:396-1710  (rectangle_pin)
 *   environment/dummy_env_rectangular_pin_spatial.py:396-1661 (rectangle_spatial_pin)
 *
 * and registers it with the PPO loop through create_env(env_config)
 * (utils/agent/utils.py:394-418).  Each entry point below cites the reference
 * interface it replaces.  Conventions: every function returns 0 on success and
 * a negative pcb_status on error; the message is available from
 * pcb_last_error() (thread-local).  No exception crosses the ABI.  A handle owns
 * all of its device buffers, belongs to one GPU, and is not thread-safe.
 * reset/step are asynchronous on the caller's stream (a CUstream / cudaStream_t
 * passed as void*; NULL = legacy default stream) and never synchronise.
 */
#ifndef PCB_ENV_H_
#define PCB_ENV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCB_ABI_VERSION 1

/* Capacities of the library (csrc/pcb_layout.h).  The reference has no such limits; a
 * configuration or instance outside them is rejected by pcb_create / pcb_upload_instances
 * with PCB_ERR_INVALID and a message naming the limit -- nothing is truncated silently. */
#define PCB_MAX_GRID 64            /* height, width: 1..64 (one u32/u64 bitboard word per row) */
#define PCB_MAX_COMPONENTS 64      /* max_num_components */
#define PCB_MAX_COMPONENT_SIDE 32  /* max_component_h, max_component_w; h*w <= 255; both must
                                      fit the grid in either orientation */
#define PCB_MAX_PINS 512           /* pins per instance */
#define PCB_MAX_NETS 63            /* max_num_nets */
#define PCB_MAX_BEAM_WIDTH 4       /* reward_beam_width, reward types beam / both */
#define PCB_MAX_BEAM_NET_PINS 16   /* pins per net when the beam router runs */

typedef enum {
  PCB_OK = 0,
  PCB_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  PCB_ERR_CUDA = -2,      /* CUDA runtime error (message has the detail) */
  PCB_ERR_STATE = -3,     /* call order (e.g. step before instances uploaded) */
  PCB_ERR_NOMEM = -4
} pcb_status;

typedef enum {
  PCB_KIND_SQUARE = 0,        /* dummy_env_square.py */
  PCB_KIND_RECT = 1,          /* dummy_env_rectangular.py */
  PCB_KIND_PIN = 2,           /* dummy_env_rectangular_pin.py */
  PCB_KIND_PIN_SPATIAL = 3    /* dummy_env_rectangular_pin_spatial.py */
} pcb_kind;

typedef enum { PCB_REWARD_CENTROID = 0, PCB_REWARD_BEAM = 1, PCB_REWARD_BOTH = 2 } pcb_reward_type;
typedef enum { PCB_ACTION_TUPLE = 0, PCB_ACTION_FLAT = 1 } pcb_action_format;

/* Constructor arguments of DummyPlacementEnv plus the host-evaluated float64
 * constants of find_reward (dummy_env_rectangular_pin.py:832-909).  The host
 * side evaluates the constants with the reference's own expression order and
 * passes them in, so the device never re-derives them. */
typedef struct pcb_config {
  int32_t abi_version;          /* PCB_ABI_VERSION */
  int32_t kind;                 /* pcb_kind */
  int32_t height, width;        /* grid, 1..64 each */
  int32_t component_n;          /* square only (dummy_env_square.py:37) */
  int32_t max_num_components;   /* C: rows of all_components_feature */
  int32_t max_component_h, max_component_w; /* P = h*w rows per component (pin:481) */
  int32_t max_num_nets;         /* K = max_num_nets + 1 channels (spatial) */
  int32_t max_pins;             /* capacity of the per-instance pin table */
  int32_t reward_type;          /* pcb_reward_type */
  int32_t reward_beam_width;    /* >= 1 */
  int32_t norm_uses_fma;        /* 1: ||d|| = sqrt(fma(dy,dy,dx*dx)) like numpy+OpenBLAS ddot */
  int32_t autoreset;            /* 1: finished envs restart on their next bank slot */
  int32_t obs_ring;             /* R >= 1 observation slices [R, N, ...] */
  int32_t reserved0;
  double weight_wirelength, weight_num_intersections;     /* pin:414-415 */
  double wirelength_norm;       /* H + W (pin:879) */
  double intersections_norm;    /* min(avg area pins, avg net pins) (pin:883-896) */
  double penalty_reward;        /* not-all-placed reward (pin:899-905) */
  double penalty_wirelength;    /* info value in that case (pin:908) */
  double penalty_intersections; /* info value in that case (pin:907) */
} pcb_config_t;

/* Host-side instance bank (struct of arrays), the exchange format of
 * what reset() draws in the reference
 * (generate_instances, pin:1006-1035).  Pins are in env.pins (net-major) order. */
typedef struct pcb_instance_bank {
  int64_t count;             /* number of instances */
  const int32_t* n_comp;     /* [count] */
  const uint8_t* comp_h;     /* [count, max_num_components] */
  const uint8_t* comp_w;     /* [count, max_num_components] */
  const int32_t* n_pins;     /* [count]            (pin kinds; may be NULL otherwise) */
  const int32_t* n_nets;     /* [count] */
  const uint8_t* pin_comp;   /* [count, max_pins] component_id */
  const uint8_t* pin_rx;     /* [count, max_pins] relative_x (un-rotated) */
  const uint8_t* pin_ry;     /* [count, max_pins] relative_y */
  const uint8_t* pin_net;    /* [count, max_pins] net_id */
  const uint16_t* pin_id;    /* [count, max_pins] feature-row id (pin:1256-1258) */
} pcb_instance_bank_t;

typedef struct pcb_env pcb_env_t;

/* DLPack v0.8 subset (the standard, ABI-stable layout of dlpack.h). */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void* data; DLDevice device; int32_t ndim; DLDataType dtype;
  int64_t* shape; int64_t* strides; uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
  DLTensor dl_tensor; void* manager_ctx; void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
#endif

/* DummyPlacementEnv.__init__ (square:37, rect:152, pin:396, spatial:396) for
 * num_envs independent environments on CUDA device `device`. */
int pcb_create(const pcb_config_t* cfg, int device, int num_envs, pcb_env_t** out);
int pcb_destroy(pcb_env_t* env);

/* Replaces the RNG draw of generate_instances (rect:253-273, pin:1006-1035):
 * the instances reset() hands out.  Slot of env i, episode e = (i + e*N) % count.
 * The bank may be replaced at any time between calls (this function, pcb_generate_instances,
 * pcb_generate_instances_device): a running episode is unaffected -- everything it needs
 * was copied into the per-env state at reset -- and the next episode of an env on slot s
 * takes slot (s + N) % new_count of the new bank. */
int pcb_upload_instances(pcb_env_t* env, const pcb_instance_bank_t* host_bank);

/* Fills the bank with `count` instances drawn on the host by the library's own
 * generator (same distributions as pin:977-1264 / rect:253-273, counter-based
 * RNG instead of the global Mersenne-Twister streams).  `gen` carries the
 * ranges: {min_component_w, max_component_w, min_component_h, max_component_h,
 * min_num_components, max_num_components, min_num_nets, max_num_nets,
 * min_num_pins_per_net, max_num_pins_per_net, net_distribution, pin_spread}. */
int pcb_generate_instances(pcb_env_t* env, const int32_t gen[12], uint64_t seed, int64_t count);

/* K8: the same generator as a kernel (one thread per instance) writing the device bank in
 * place -- asynchronous on `stream`, no host round trip; call it between steps to keep
 * feeding fresh instances (the bank is (re)allocated when `count` changes). */
int pcb_generate_instances_device(pcb_env_t* env, const int32_t gen[12], uint64_t seed,
                                  int64_t count, void* stream);

/* The same generator without a handle or a GPU: fills caller-owned host arrays laid out
 * like pcb_instance_bank_t (pin arrays may be NULL for PCB_KIND_RECT).  For tests and
 * for building banks offline. */
int pcb_generate_host(int kind, int max_num_components, int max_pins, int max_num_nets,
                      const int32_t gen[12], uint64_t seed, int64_t count, int32_t* n_comp,
                      uint8_t* comp_h, uint8_t* comp_w, int32_t* n_pins, int32_t* n_nets,
                      uint8_t* pin_comp, uint8_t* pin_rx, uint8_t* pin_ry, uint8_t* pin_net,
                      uint16_t* pin_id);

/* DummyPlacementEnv.reset() (square:74, rect:310, pin:1544, spatial:1487) for
 * every env (d_reset_mask == NULL: episode 0 of each env, statistics cleared)
 * or for the envs whose mask byte is non-zero (they advance to their next
 * bank slot).  Writes the initial observation into the current slice. */
int pcb_reset(pcb_env_t* env, const uint8_t* d_reset_mask, void* stream);

/* DummyPlacementEnv.step(action) (square:115, rect:353, pin:1599, spatial:1551)
 * for all envs.  d_actions: device int32, [N,2] (square) / [N,3] (o,x,y) when
 * PCB_ACTION_TUPLE, [N] flat index when PCB_ACTION_FLAT (the decoding of
 * FlatteningActionWrapper*, env_wrappers.py:80-98, wrapper_square.py:54-69).
 * Advances the observation ring by one slice and writes obs/reward/done/info. */
int pcb_step(pcb_env_t* env, const int32_t* d_actions, int action_format, void* stream);

/* The random policy of agent/random/random_policy_*.py:11-58 (uniform over the
 * valid entries of action_mask) evaluated on device for every env.  Counter-based:
 * (seed, env index, step counter).  d_actions as for pcb_step / PCB_ACTION_TUPLE. */
int pcb_sample_random_actions(pcb_env_t* env, uint64_t seed, int32_t* d_actions, void* stream);

/* step() through HOST buffers: copies h_actions (ideally pinned) to the device,
 * steps, copies reward[N] (f64) and done[N] (u8) back and waits for them.
 * pcb_step_host_packed: the results arrive in ONE block h_result = [reward f64 x N | done u8 x N]
 * (N * 9 bytes, one allocation) with a single device-to-host copy -- the library keeps the two
 * side by side on the device. */
int pcb_step_host(pcb_env_t* env, const int32_t* h_actions, int action_format,
                  double* h_reward, uint8_t* h_done, void* stream);
int pcb_step_host_packed(pcb_env_t* env, const int32_t* h_actions, int action_format,
                         void* h_result, void* stream);

/* n_steps calls of step() for a caller that already holds the action rows (replay of a
 * recorded rollout, evaluation of a fixed plan, a policy that emits several rows at once):
 * h_actions [n_steps, N] flat or [n_steps, N, 2|3] tuples; h_reward [n_steps, N] f64 and
 * h_done [n_steps, N] u8 receive every step's results.  Every step still pays its own
 * host-to-device copy of its action row and device-to-host copy of its result rows, but
 * the copies ride on the copy engines while the kernels of neighbouring steps run, and the
 * host waits once, at the end (the pin-less kinds run all rows in one launch of the
 * multi-step kernel).  Bit-identical to n_steps x pcb_step_host.  Synchronises `stream`. */
int pcb_steps_host(pcb_env_t* env, const int32_t* h_actions, int n_steps, int action_format,
                   double* h_reward, uint8_t* h_done, void* stream);

/* Observation / result tensors as non-owning DLPack views, valid until
 * pcb_destroy.  Keys: the reference's observation keys (grid, action_mask,
 * all_components_feature, component_mask, placement_mask, all_pins_num_feature,
 * all_pins_cat_feature, pin_grid, component_grid) with shape [R, N, ...], and
 * reward, done, wirelength, num_intersections, episode_return, episode_length
 * with shape [2, N] (see pcb_current_result_half).  The deleter frees only the small manager struct. */
int pcb_get_tensor(pcb_env_t* env, const char* key, DLManagedTensor** out);

/* Index of the ring slice the last reset/step wrote. */
int pcb_current_slice(pcb_env_t* env);

/* reward, done, wirelength, num_intersections, episode_return, episode_length are
 * double-buffered ([2, N]): the step that just ran wrote half pcb_current_result_half(). */
int pcb_current_result_half(pcb_env_t* env);

/* Episode statistics accumulated on device since the last clear:
 * out = {episodes, sum_return, sum_length, sum_wirelength, sum_intersections,
 *        steps, 0, 0}.  Synchronises `stream`. */
int pcb_episode_stats(pcb_env_t* env, double out[8], int reset_counters, void* stream);

/* Copies the compact per-env state (occupancy bitboard rows, little-endian
 * 64-bit words, [N, 64]) to the host -- for tests and checkpointing. */
int pcb_get_occupancy(pcb_env_t* env, uint64_t* h_rows, void* stream);

/* Checkpoint / resume and inspection of the compact per-env state records
 * (layout: csrc/pcb_layout.h).  out = {st_stride, off_components, off_pins, off_bitboard,
 * max_num_components, max_pins, bytes per bitboard row word, rows stored}.  get/set copy
 * [N, st_stride] bytes and the step counter (which seeds the device random policy);
 * the instance bank is not part of the state -- restore it with the same upload /
 * generator call.  Both synchronise `stream`. */
int pcb_state_layout(pcb_env_t* env, int32_t out[8]);
int pcb_get_state(pcb_env_t* env, void* h_state, uint64_t* step_no, void* stream);
int pcb_set_state(pcb_env_t* env, const void* h_state, uint64_t step_no, void* stream);

/* Number of kernels this handle has launched so far (bench: gpu_launches). */
int64_t pcb_launch_count(pcb_env_t* env);

/* simulate() of agent/random/random_policy_*.py:25-58 for all envs: n_steps times
 * { pcb_sample_random_actions; pcb_step } issued from one call.  d_actions_log is
 * NULL or a device buffer [n_steps, N, 2|3] int32 that receives every action. */
int pcb_rollout_random(pcb_env_t* env, uint64_t seed, int n_steps, int32_t* d_actions_log,
                       void* stream);

/* The same rollout keeping every step's results for a rollout collector: d_reward_log
 * [n_steps, N] float64 and d_done_log [n_steps, N] uint8 (device, either may be NULL)
 * receive reward and done of each step; together with an observation ring of n_steps + 1
 * slices and d_actions_log this is the whole [T, N, ...] batch a PPO-style learner reads,
 * produced without a host round trip per step. */
int pcb_rollout_random_logged(pcb_env_t* env, uint64_t seed, int n_steps, int32_t* d_actions_log,
                              double* d_reward_log, uint8_t* d_done_log, void* stream);

/* The per-step kinds (pin, pin-spatial, boards wider than 32) run pcb_rollout_random as
 * replays of a CUDA graph of L steps (L = a multiple of the observation ring and of 4, about
 * 16 steps: step kernels, overlapped reward kernels and their events captured once), the
 * remainder with plain launches -- default on; results are identical either way.  Small
 * batches are launch-bound without it.  pcb_graph_replays: graph launches so far. */
int pcb_set_graph_rollout(pcb_env_t* env, int on);
int64_t pcb_graph_replays(pcb_env_t* env);

/* pcb_rollout_random runs the reward kernel of step t concurrently with the step kernel of
 * step t+1 (default on); 0 = strictly one after the other, as pcb_step always does. */
int pcb_set_reward_overlap(pcb_env_t* env, int on);

/* The pin-less kinds (square, rectangle; H, W <= 32) run pcb_rollout_random as ONE launch
 * of a multi-step kernel that keeps the state on chip and packs 32/LPE environments into a
 * warp (LPE = 8, 16, 32 lanes >= H) -- default on; the result is bit-identical to n_steps x
 * { pcb_sample_random_actions; pcb_step }; pcb_step itself runs one step of the same kernel
 * with the caller's actions.  pcb_set_fused_rollout(env, 0) selects the one-warp-per-env
 * kernels and per-step launches instead; returns 1 when the fused kernel will be used, 0 when not
 * (pin kinds, boards wider than 32), negative on error.  pcb_fused_rollout_times: device
 * time [ms], launches and steps of the fused kernel since the last call (needs
 * pcb_enable_kernel_timing). */
int pcb_set_fused_rollout(pcb_env_t* env, int on);
int pcb_fused_rollout_times(pcb_env_t* env, double* ms, int64_t* launches, int64_t* steps);

/* Memory-safety aid (compute-sanitizer stand-in): with PCB_GUARD=1 in the environment when
 * pcb_create runs, every device buffer of the handle is allocated between two 4 KB guard
 * bands of a known pattern; pcb_check_guards waits for the device and returns the number of
 * guard bytes that were overwritten (0 = no out-of-bounds store next to any buffer),
 * negative on error / when the handle has no guards. */
int64_t pcb_check_guards(pcb_env_t* env, void* stream);

/* Measurement aid: when on, every kernel launch is bracketed by CUDA events on the
 * launching stream.  pcb_kernel_times waits for them and returns the summed device
 * time [ms] and launch count per kernel {step, reward, reset, sample}, then clears. */
int pcb_enable_kernel_timing(pcb_env_t* env, int on);
int pcb_kernel_times(pcb_env_t* env, double ms[4], int64_t launches[4]);

/* Which kernel instantiation the handle uses: 0 = runtime-shaped (any configuration),
 * 1..99 = a built-in compile-time shape (the canonical and the reference's shipped
 * configurations), 100 = the shape this library was specialised for at build time. */
int pcb_shape_id(pcb_env_t* env);

const char* pcb_last_error(void);
int pcb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PCB_ENV_H_ */
