// pcb_api.cu -- host side of libpcbgpu: the C ABI of include/pcb_env.h.
//
// Owns the device buffers of one handle (state, instance bank, observation ring),
// packs host instances into the device record format, and launches the kernels of
// pcb_kernels.cuh.  No CPU implementation of the environment lives here: without a
// CUDA device every compute entry point fails with PCB_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/pcb_env.h"
#include "pcb_gen.h"
#include "pcb_gen_device.cuh"
#include "pcb_kernels.cuh"
#include "pcb_small.cuh"
#include "pcb_pin_small.cuh"

using namespace pcb;

static_assert(PCB_MAX_GRID == kMaxGrid && PCB_MAX_COMPONENTS == kMaxComps && PCB_MAX_PINS == kMaxPins &&
              PCB_MAX_NETS == kMaxNets - 1 && PCB_MAX_BEAM_WIDTH == kMaxBeamWidth &&
              PCB_MAX_BEAM_NET_PINS == kMaxBeamPoints + 1,
              "include/pcb_env.h documents the capacities of csrc/pcb_layout.h");

#ifndef PCB_PERSIST_CTAS
#define PCB_PERSIST_CTAS PCB_MIN_BLOCKS
#endif
#ifndef PCB_OVERLAP_REWARD
#define PCB_OVERLAP_REWARD 1
#endif

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(expr)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(PCB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                  \
  } while (0)

struct TensorDesc {
  const char* key;
  void** ptr;
  int code, bits;
  std::vector<int64_t> shape;  // full shape, [R, N, ...] or [N]
  bool ringed;
  std::vector<int64_t> strides;  // in elements; empty = contiguous
};

}  // namespace

struct pcb_env {
  pcb_config_t cfg;
  Dims d;
  int device, n_envs, ring, slice;
  bool wide;  // true: 64-bit rows, two rows per lane
  uint64_t step_no;
  int64_t launches;
  size_t smem_bytes;
  // device buffers
  uint8_t* state;
  uint8_t* bank;
  uint32_t bank_count;
  Out base;  // slice 0 pointers
  double* stats;
  int32_t* d_actions;  // staging for pcb_step_host / ping-pong of pcb_rollout_random
  int32_t* d_actions2;
  // the last step of a pcb_rollout_random call also draws the actions of the step after it;
  // the next call with the same seed on the same stream starts from them instead of
  // launching the stand-alone sampler (cleared by everything else that moves the state)
  int32_t* pend_actions;
  uint64_t pend_seed, pend_step;
  cudaStream_t pend_stream;
  // reward and done of one step sit side by side ([reward f64 x N | done u8 x N], one block
  // per result half) so that pcb_step_host brings both back with ONE device-to-host copy
  size_t res_stride;   // bytes between the two halves
  // pcb_steps_host: device staging of T action rows and T result rows, a copy stream so that
  // the upload of later rows overlaps the first steps, events per chunk of rows
  int32_t* hs_actions; double* hs_reward; uint8_t* hs_done;
  size_t hs_steps;     // capacity in steps
  cudaStream_t copy;
  std::vector<cudaEvent_t> hs_events;
  cudaEvent_t ev_host;
  // pcb_rollout_random of the per-step kinds as a CUDA graph of `graph_len` steps (the step
  // kernels, the overlapped reward kernels on the side stream and their events captured once)
  bool use_graph;
  cudaGraphExec_t graph;
  int graph_len, graph_slice;
  uint64_t graph_step0, graph_seed;      // step number and seed at capture
  cudaStream_t graph_stream;
  bool graph_overlap;
  uint64_t* d_step_base;
  int64_t graph_launches, graph_replays;
  float* ratio_lut;
  int shape_id;        // 0 = generic (DynShape), 1..4 = ShapeB/C/D/E
  uint8_t* jobs;       // routing jobs (pin kinds)
  uint32_t* job_count; // [2]
  size_t reward_smem;
  bool reward_half;      // centroid reward, max_pins <= 32: reward_half_kernel (two jobs per warp)
  int list_group;        // lanes per job of reward_list_half_kernel: 16 or 8 (reward_list_group_lanes)
  int step_group;        // ... of reward_half_kernel (reward_step_group_lanes)
  bool list_half_beam;   // reward types beam / both, max_pins <= 32: reward_list_half_kernel<G, true>
  int reward_grid;       // reward_kernel grid when it runs alone
  int reward_grid_side;  // ... when it overlaps the next step kernel (pcb_rollout_random)
  cudaStream_t side;     // stream of the overlapped reward kernels
  cudaEvent_t ev_step[2], ev_rew[2];
  bool rew_pending[2];
  bool overlap;          // pcb_rollout_random overlaps the reward kernel with the next step
  int sms;               // multiprocessors of the device
  bool fused;            // pcb_rollout_random uses rollout_small_kernel (pin-less kinds, H, W <= 32)
  // pin / pin-spatial env, H, W <= 32: rollout_pin_kernel (pcb_pin_small.cuh) -- many steps
  // per launch, 32/LPE envs per warp, terminal rewards from a job list afterwards
  bool fused_pin, fused_pin_ok;
  int list_side_mult;  // list_steps: CTAs per SM (in units of 8 one-job warps) of the side kernels
  bool state_stream;  // Params::state_stream: the state records are larger than kStateStreamBytes
  bool list_steps;  // multi-step pin path with per-step kernels filling the job list (list_steps_chunk)
  size_t pin_smem;       // dynamic shared memory per CTA (4 warps)
  uint8_t* pin_jobs[2];  // job lists of two launches in flight, pin_job_steps * N records each
  uint32_t* pin_job_total[2];
  int pin_job_steps;     // steps per launch a list is sized for
  size_t small_smem;     // its dynamic shared memory per CTA
  double fused_ms;       // timing (pcb_enable_kernel_timing): device time, launches, env-steps
  int64_t fused_launches, fused_steps;
  std::vector<TensorDesc> tensors;
  std::vector<void*> allocs;
  // PCB_GUARD=1 in the environment at pcb_create: every handle-owned buffer sits between two
  // guard bands of a known byte pattern that pcb_check_guards verifies (compute-sanitizer is not
  // available on the test pool; this catches out-of-bounds stores next to any buffer)
  bool guard;
  struct Guarded { uint8_t* base; size_t bytes; };
  std::vector<Guarded> guarded;
  // optional per-kernel CUDA-event timing (pcb_enable_kernel_timing)
  bool timing;
  struct Timed { int id; cudaEvent_t a, b; };
  std::vector<Timed> timed;       // outstanding event pairs, at most kMaxTimed
  double timed_ms[4];             // totals of the pairs already folded in
  int64_t timed_n[4];
};

namespace {

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  // many small CTAs per SM: ask for the largest shared-memory carve-out
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  return 0;
}

Out slice_ptrs(const pcb_env* e, int slice, int parity) {
  const Dims& d = e->d;
  const size_t N = (size_t)e->n_envs, s = (size_t)slice, q = (size_t)parity;
  Out o = e->base;
  o.reward = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(o.reward) + q * e->res_stride);
  o.done += q * e->res_stride;
  o.wl += q * N; o.ni += q * N;
  o.ep_return += q * N; o.ep_length += q * N;
  if (o.grid) o.grid += s * N * d.L.grid_b;
  if (o.mask) o.mask += s * N * d.L.mask_b;
  if (o.feat) o.feat += s * N * d.L.feat_n;
  if (o.comp_mask) o.comp_mask += s * N * d.C;
  if (o.place_mask) o.place_mask += s * N * d.C;
  if (o.pnum) o.pnum += s * N * d.L.pnum_n;
  if (o.pcat) o.pcat += s * N * d.L.pcat_n;
  if (o.pgrid) o.pgrid += s * N * d.L.pgrid_b;
  if (o.cgrid) o.cgrid += s * N * d.L.cgrid_b;
  return o;
}

// step_no = index of the step about to run (or just reset): its scalars go to half
// (step_no & 1); after a reset (no step yet) the "last written" half is 1.
Params make_params(const pcb_env* e, bool for_reset = false) {
  Params p;
  memset(&p, 0, sizeof p);
  p.d = e->d;
  const int parity = for_reset ? (int)((e->step_no + 1) & 1) : (int)(e->step_no & 1);
  p.out = slice_ptrs(e, e->slice, parity);
  p.state = e->state;
  p.bank = e->bank;
  p.bank_count = e->bank_count;
  p.n_envs = e->n_envs;
  p.stats = e->stats;
  p.step_no = e->step_no;
  p.ratio_lut = e->ratio_lut;
  p.jobs = e->jobs ? e->jobs + (size_t)parity * e->n_envs * job_stride(e->d.max_pins, e->d.K) : nullptr;
  p.job_count = e->job_count;
  p.job_slot_mask = 3; p.job_t = 0;
  p.state_stream = e->state_stream ? 1u : 0u;
  return p;
}

enum KernelId { K_STEP, K_TERM, K_RESET, K_SAMPLE };

// outstanding timing events are bounded: past kMaxTimed pairs the finished ones are folded
// into running totals (a caller that enables timing and never polls does not grow the list)
#ifndef PCB_LIST_SIDE_MULT
#define PCB_LIST_SIDE_MULT 16
#endif
#ifndef PCB_LIST_STEPS
#define PCB_LIST_STEPS 1  // reward types beam / both at large batches: list_steps_chunk
#endif
constexpr size_t kMaxTimed = 4096;
constexpr size_t kStateStreamBytes = 96u << 20;  // state records above this: streaming loads / stores
int fold_timed(pcb_env* e) {
  for (auto& t : e->timed) {
    CU(cudaEventSynchronize(t.b));
    float v = 0.f;
    CU(cudaEventElapsedTime(&v, t.a, t.b));
    e->timed_ms[t.id] += v;
    e->timed_n[t.id] += 1;
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  e->timed.clear();
  return 0;
}

template <typename S>
bool shape_matches(const Dims& d) {
  S sh(d);
  return d.kind == sh.kind() && d.H == sh.H() && d.W == sh.W() && d.C == sh.C() && d.mh == sh.mh() &&
         d.mw == sh.mw() && d.K == sh.K() && d.max_pins == sh.max_pins();
}

int pick_shape(const Dims& d) {
  // host-side mirror of the FixedShape parameters (pcb_kernels.cuh)
  auto is = [&](int kind, int H, int W, int C, int mh, int mw, int K, int mp) {
    return d.kind == kind && d.H == H && d.W == W && d.C == C && d.mh == mh && d.mw == mw &&
           d.K == K && d.max_pins == mp;
  };
  if (is(1, 16, 16, 6, 4, 4, 0, 0)) return 1;
  if (is(2, 32, 32, 10, 4, 4, 6, 30)) return 2;
  if (is(3, 32, 32, 10, 4, 4, 6, 30)) return 3;
  if (is(2, 64, 64, 32, 4, 4, 17, 96)) return 4;
  if (is(2, 10, 10, 5, 2, 2, 4, 18)) return 5;
  if (is(3, 10, 10, 5, 2, 2, 4, 18)) return 6;
  if (is(1, 10, 10, 20, 4, 4, 0, 0)) return 7;
  if (is(0, 10, 10, 0, 0, 0, 0, 0)) return 8;
  if (is(0, 8, 8, 0, 0, 0, 0, 0)) return 9;
#ifdef PCB_USER_SHAPE
  {
    const int u[8] = {PCB_USER_SHAPE};
    if (is(u[0], u[1], u[2], u[3], u[4], u[5], u[6], u[7])) return 100;
  }
#endif
  return 0;
}

template <int RPL, typename T, typename S>
int launch_shape(KernelId id, dim3 grid, dim3 block, size_t sm, cudaStream_t st, const Params& p) {
  switch (id) {
    case K_STEP: step_kernel<RPL, T, S><<<grid, block, sm, st>>>(p); break;
    case K_RESET: reset_kernel<RPL, T, S><<<grid, block, sm, st>>>(p); break;
    default: break;
  }
  return 0;
}

#ifdef PCB_USER_SHAPE
constexpr int kUserDims[8] = {PCB_USER_SHAPE};
constexpr bool kUserWide = kUserDims[1] > 32 || kUserDims[2] > 32;
using UserWord = std::conditional_t<kUserWide, uint64_t, uint32_t>;
constexpr int kUserRpl = kUserWide ? 2 : 1;
#endif

template <int RPL, typename T, typename S>
int set_smem_shape(size_t bytes) {
  int rc = set_smem(step_kernel<RPL, T, S>, bytes);
  if (rc) return rc;
  return set_smem(reset_kernel<RPL, T, S>, bytes);
}

int launch(pcb_env* e, KernelId id, const Params& p, cudaStream_t st, int reward_grid = 0) {
  dim3 grid((e->n_envs + kWarpsPerCta - 1) / kWarpsPerCta), block(kWarpsPerCta * 32);
#if PCB_OPT_PERSIST
  // persistent step kernel: as many CTAs as are resident at once (PCB_PERSIST_CTAS per SM)
  if (id == K_STEP) grid.x = std::min<unsigned>(grid.x, (unsigned)(e->sms * PCB_PERSIST_CTAS));
#endif
  size_t sm = id == K_SAMPLE ? 0 : e->smem_bytes;
  pcb_env::Timed tm{(int)id, nullptr, nullptr};
  if (e->timing) {
    CU(cudaEventCreate(&tm.a));
    CU(cudaEventCreate(&tm.b));
    CU(cudaEventRecord(tm.a, st));
  }
  if (id == K_TERM) {
    dim3 rgrid(reward_grid > 0 ? reward_grid : e->reward_grid);
    if (e->d.reward_type == PCB_REWARD_CENTROID) {
      if (e->reward_half) {  // two (four) jobs per warp: as many CTAs fewer for the same jobs per group
        const int groups = 32 / e->step_group;
        const dim3 hgrid((rgrid.x + groups - 1) / groups);
        if (e->step_group == 8) reward_half_kernel<8><<<hgrid, block, e->reward_smem * groups, st>>>(p);
        else reward_half_kernel<16><<<hgrid, block, e->reward_smem * groups, st>>>(p);
      } else
        reward_kernel<false><<<rgrid, block, e->reward_smem, st>>>(p);
    } else {
      beam_kernel<<<rgrid, 128, 0, st>>>(p);  // thread per (job, net); paths go into the jobs
      e->launches += 1;
      reward_kernel<true><<<rgrid, block, e->reward_smem, st>>>(p);
    }
  } else if (id == K_SAMPLE) {
    if (e->wide) sample_kernel<2, uint64_t><<<grid, block, sm, st>>>(p);
    else sample_kernel<1, uint32_t><<<grid, block, sm, st>>>(p);
  } else {
    switch (e->shape_id) {
      case 1: launch_shape<1, uint32_t, ShapeB>(id, grid, block, sm, st, p); break;
      case 2: launch_shape<1, uint32_t, ShapeC>(id, grid, block, sm, st, p); break;
      case 3: launch_shape<1, uint32_t, ShapeD>(id, grid, block, sm, st, p); break;
      case 4: launch_shape<2, uint64_t, ShapeE>(id, grid, block, sm, st, p); break;
      case 5: launch_shape<1, uint32_t, ShapeRefPin>(id, grid, block, sm, st, p); break;
      case 6: launch_shape<1, uint32_t, ShapeRefSpatial>(id, grid, block, sm, st, p); break;
      case 7: launch_shape<1, uint32_t, ShapeRefRect>(id, grid, block, sm, st, p); break;
      case 8: launch_shape<1, uint32_t, ShapeRefSquare>(id, grid, block, sm, st, p); break;
      case 9: launch_shape<1, uint32_t, ShapeA>(id, grid, block, sm, st, p); break;
#ifdef PCB_USER_SHAPE
      case 100: launch_shape<kUserRpl, UserWord, ShapeUser>(id, grid, block, sm, st, p); break;
#endif
      default:
        if (e->wide) launch_shape<2, uint64_t, DynShape>(id, grid, block, sm, st, p);
        else launch_shape<1, uint32_t, DynShape>(id, grid, block, sm, st, p);
    }
  }
  CU(cudaGetLastError());
  if (e->timing) {
    CU(cudaEventRecord(tm.b, st));
    e->timed.push_back(tm);
    if (e->timed.size() >= kMaxTimed) {
      int rc = fold_timed(e);
      if (rc) return rc;
    }
  }
  e->launches += 1;
  return 0;
}

// rollout_small_kernel: instantiations and dispatch.  LPE (lanes per env) follows from H.
template <typename S, int LPE>
int small_launch(bool configure, size_t sm, dim3 grid, dim3 block, cudaStream_t st, const RolloutParams& p) {
  if (configure) return set_smem(rollout_small_kernel<LPE, S>, sm);
  rollout_small_kernel<LPE, S><<<grid, block, sm, st>>>(p);
  return 0;
}

// Lanes per env = the smallest power of two >= H: the most envs per warp.  Spreading a small
// batch over more warps (wider groups, idle lanes) was measured and is slower: B at 4 096 envs
// 3.1e9 env-steps/s with 16 lanes per env, 1.9e9 with 32 (PCB_SMALL_LPE forces a width).
int small_pick_lpe(const pcb_env* e) {
  const int lpe = small_lanes_per_env(e->d.H);
  if (const char* f = getenv("PCB_SMALL_LPE")) {
    const int v = atoi(f);
    if ((v == 8 || v == 16 || v == 32) && v >= lpe) return v;
  }
  return lpe;
}

template <typename S>
int small_launch_lpe(int lpe, bool configure, size_t sm, dim3 grid, dim3 block, cudaStream_t st,
                     const RolloutParams& p) {
  constexpr int kMin = S::fixed ? small_lanes_per_env(S::kH) : 8;
  if (configure) {  // every instantiation the dispatcher may pick later
    int rc = 0;
    if constexpr (kMin <= 8) rc = rc ? rc : small_launch<S, 8>(true, sm, grid, block, st, p);
    if constexpr (kMin <= 16) rc = rc ? rc : small_launch<S, 16>(true, sm, grid, block, st, p);
    return rc ? rc : small_launch<S, 32>(true, sm, grid, block, st, p);
  }
  if constexpr (kMin <= 8) if (lpe == 8) return small_launch<S, 8>(false, sm, grid, block, st, p);
  if constexpr (kMin <= 16) if (lpe == 16) return small_launch<S, 16>(false, sm, grid, block, st, p);
  return small_launch<S, 32>(false, sm, grid, block, st, p);
}

int small_dispatch(pcb_env* e, bool configure, cudaStream_t st, const RolloutParams& p) {
  const int lpe = configure ? small_lanes_per_env(e->d.H) : small_pick_lpe(e), G = 32 / lpe;
  // few envs: one-warp CTAs spread evenly over the SMs (all resident at once); many: 4 warps
  const int warps = (e->n_envs + G - 1) / G;
  const int wpc = warps <= e->sms * 32 ? 1 : kWarpsPerCta;
  dim3 grid((warps + wpc - 1) / wpc), block(wpc * 32);
  const size_t sm = configure ? e->small_smem
                              : (size_t)kLutBytes + (size_t)small_env_smem(e->d.C, e->d.H * e->d.W) * wpc * G;
  switch (e->shape_id) {
    case 1: return small_launch_lpe<ShapeB>(lpe, configure, sm, grid, block, st, p);
    case 7: return small_launch_lpe<ShapeRefRect>(lpe, configure, sm, grid, block, st, p);
    case 8: return small_launch_lpe<ShapeRefSquare>(lpe, configure, sm, grid, block, st, p);
    case 9: return small_launch_lpe<ShapeA>(lpe, configure, sm, grid, block, st, p);
#ifdef PCB_USER_SHAPE
    case 100:
      if constexpr (kUserDims[0] <= 1 && !kUserWide)
        return small_launch_lpe<ShapeUser>(lpe, configure, sm, grid, block, st, p);
      break;
#endif
    default: break;
  }
  return small_launch_lpe<DynShape>(lpe, configure, sm, grid, block, st, p);
}

// rollout_pin_kernel: instantiations and dispatch (LPE follows from H)
template <typename S, int LPE>
int pin_launch(bool configure, size_t sm, dim3 grid, dim3 block, cudaStream_t st, const PinRolloutParams& p) {
  if (configure) return set_smem(rollout_pin_kernel<LPE, S>, sm);
  rollout_pin_kernel<LPE, S><<<grid, block, sm, st>>>(p);
  return 0;
}

int pin_dispatch(pcb_env* e, bool configure, cudaStream_t st, const PinRolloutParams& p) {
  const int lpe = small_lanes_per_env(e->d.H), G = 32 / lpe;
  const int warps = (e->n_envs + G - 1) / G;
  const int wpc = warps <= e->sms * 32 ? 1 : kWarpsPerCta;
  dim3 grid((warps + wpc - 1) / wpc), block(wpc * 32);
  const PinSmem L = pin_smem_layout(e->d.kind, e->d.C, e->d.L.P, e->d.max_pins, e->d.H, e->d.W, e->d.K, e->d.mh, e->d.mw);
  const size_t sm = configure ? e->pin_smem : (size_t)kLutBytes + (size_t)L.total * wpc * G;
  switch (e->shape_id) {
    case 2: return pin_launch<ShapeC, 32>(configure, sm, grid, block, st, p);
    case 3: return pin_launch<ShapeD, 32>(configure, sm, grid, block, st, p);
    case 5: return pin_launch<ShapeRefPin, 16>(configure, sm, grid, block, st, p);
    case 6: return pin_launch<ShapeRefSpatial, 16>(configure, sm, grid, block, st, p);
#ifdef PCB_USER_SHAPE
    case 100:
      if constexpr (kUserDims[0] >= 2 && !kUserWide)
        return pin_launch<ShapeUser, small_lanes_per_env(kUserDims[1])>(configure, sm, grid, block, st, p);
      break;
#endif
    default: break;
  }
  if (lpe == 8) return pin_launch<DynShape, 8>(configure, sm, grid, block, st, p);
  if (lpe == 16) return pin_launch<DynShape, 16>(configure, sm, grid, block, st, p);
  return pin_launch<DynShape, 32>(configure, sm, grid, block, st, p);
}

constexpr size_t kGuardBytes = 4096;
constexpr int kGuardByte = 0xA5;

template <typename P>
int dev_alloc(pcb_env* e, P** out, size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  const size_t pad = e->guard ? kGuardBytes : 0;
  cudaError_t err = cudaMalloc(&p, bytes + 2 * pad);
  if (err != cudaSuccess)
    return fail(err == cudaErrorMemoryAllocation ? PCB_ERR_NOMEM : PCB_ERR_CUDA,
                "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(err));
  e->allocs.push_back(p);
  uint8_t* base = reinterpret_cast<uint8_t*>(p);
  if (pad) {
    CU(cudaMemset(base, kGuardByte, bytes + 2 * pad));
    e->guarded.push_back({base, bytes});
  }
  CU(cudaMemset(base + pad, 0, bytes));
  *out = reinterpret_cast<P*>(base + pad);
  return 0;
}

int check_config(const pcb_config_t* c) {
  if (!c) return fail(PCB_ERR_INVALID, "config is NULL");
  if (c->abi_version != PCB_ABI_VERSION)
    return fail(PCB_ERR_INVALID, "ABI version %d, library is %d", c->abi_version, PCB_ABI_VERSION);
  if (c->kind < 0 || c->kind > 3) return fail(PCB_ERR_INVALID, "unknown kind %d", c->kind);
  if (c->height < 1 || c->width < 1 || c->height > kMaxGrid || c->width > kMaxGrid)
    return fail(PCB_ERR_INVALID, "grid %dx%d outside 1..%d", c->height, c->width, kMaxGrid);
  if (c->obs_ring < 1) return fail(PCB_ERR_INVALID, "obs_ring must be >= 1");
  if (c->kind == PCB_KIND_SQUARE) {
    if (c->component_n < 1 || c->component_n > c->height || c->component_n > c->width)
      return fail(PCB_ERR_INVALID, "component_n %d does not fit the grid", c->component_n);
    return 0;
  }
  if (c->max_num_components < 1 || c->max_num_components > kMaxComps)
    return fail(PCB_ERR_INVALID, "max_num_components %d outside 1..%d", c->max_num_components,
                kMaxComps);
  int big = std::max(c->max_component_h, c->max_component_w);
  if (c->max_component_h < 1 || c->max_component_w < 1 || big > std::min(c->height, c->width) ||
      big > 32)
    return fail(PCB_ERR_INVALID, "component size %dx%d must fit the grid in both orientations",
                c->max_component_h, c->max_component_w);
  if (c->kind >= PCB_KIND_PIN) {
    if (c->max_num_nets < 1 || c->max_num_nets >= kMaxNets)
      return fail(PCB_ERR_INVALID, "max_num_nets %d outside 1..%d", c->max_num_nets, kMaxNets - 1);
    if (c->max_pins < 1 || c->max_pins > kMaxPins)
      return fail(PCB_ERR_INVALID, "max_pins %d outside 1..%d", c->max_pins, kMaxPins);
    if (c->max_component_h * c->max_component_w > 255)
      return fail(PCB_ERR_INVALID, "more than 255 pins per component");
    if (c->reward_type < 0 || c->reward_type > 2)
      return fail(PCB_ERR_INVALID, "unknown reward_type %d", c->reward_type);
    if (c->reward_type != PCB_REWARD_CENTROID &&
        (c->reward_beam_width < 1 || c->reward_beam_width > kMaxBeamWidth))
      return fail(PCB_ERR_INVALID, "reward_beam_width %d outside 1..%d", c->reward_beam_width,
                  kMaxBeamWidth);
  }
  return 0;
}

void fill_dims(const pcb_config_t& c, bool fast_board, Dims& d) {
  memset(&d, 0, sizeof d);
  d.kind = c.kind; d.H = c.height; d.W = c.width;
  d.sq_n = c.component_n;
  if (c.kind != 0) { d.C = c.max_num_components; d.mh = c.max_component_h; d.mw = c.max_component_w; }
  if (c.kind >= 2) { d.K = c.max_num_nets + 1; d.max_pins = c.max_pins; }
  d.reward_type = c.reward_type; d.beam_width = c.reward_beam_width;
  d.norm_fma = c.norm_uses_fma; d.autoreset = c.autoreset;
  d.L = make_layout(d.kind, d.H, d.W, d.C, d.mh, d.mw, d.K, d.max_pins, fast_board);
  d.ww = c.weight_wirelength; d.wi = c.weight_num_intersections;
  d.wl_norm = c.wirelength_norm; d.ni_norm = c.intersections_norm;
  d.pen_reward = c.penalty_reward; d.pen_wl = c.penalty_wirelength; d.pen_ni = c.penalty_intersections;
}

void dl_deleter(DLManagedTensor* self) {
  if (!self) return;
  free(self->dl_tensor.shape);
  free(self);
}

int pack_and_upload(pcb_env* e, const pcb_instance_bank_t* b) {
  const Dims& d = e->d;
  if (d.kind == 0) return fail(PCB_ERR_INVALID, "the square env has no instances");
  if (!b || b->count < 1 || !b->n_comp || !b->comp_h || !b->comp_w)
    return fail(PCB_ERR_INVALID, "instance bank is empty or has NULL arrays");
  if (d.kind >= 2 && (!b->n_pins || !b->n_nets || !b->pin_comp || !b->pin_rx || !b->pin_ry ||
                      !b->pin_net || !b->pin_id))
    return fail(PCB_ERR_INVALID, "pin arrays missing in the instance bank");
  if (b->count > 0x7fffffffLL) return fail(PCB_ERR_INVALID, "too many instances");
  const size_t cnt = (size_t)b->count;
  std::vector<uint8_t> host(cnt * d.L.rec_stride, 0);
  std::vector<int> last(kMaxComps * 256);
  for (size_t i = 0; i < cnt; ++i) {
    uint8_t* rec = host.data() + i * d.L.rec_stride;
    EnvHdr* h = reinterpret_cast<EnvHdr*>(rec);
    int nc = b->n_comp[i];
    if (nc < 1 || nc > d.C) return fail(PCB_ERR_INVALID, "instance %zu: n_comp %d outside 1..%d", i, nc, d.C);
    h->n_comp = (uint8_t)nc;
    CompRec* comps = reinterpret_cast<CompRec*>(rec + d.L.rec_comps);
    for (int c = 0; c < nc; ++c) {
      int ch = b->comp_h[i * d.C + c], cw = b->comp_w[i * d.C + c];
      if (ch < 1 || cw < 1 || ch > d.mh || cw > d.mw)
        return fail(PCB_ERR_INVALID, "instance %zu: component %d is %dx%d, max %dx%d", i, c, ch, cw, d.mh, d.mw);
      comps[c] = CompRec{(uint8_t)ch, (uint8_t)cw, 0xFF, 0xFF};
    }
    if (d.kind < 2) continue;
    int np = b->n_pins[i], nn = b->n_nets[i];
    if (np < 0 || np > d.max_pins) return fail(PCB_ERR_INVALID, "instance %zu: n_pins %d outside 0..%d", i, np, d.max_pins);
    if (nn < 0 || nn >= d.K) return fail(PCB_ERR_INVALID, "instance %zu: n_nets %d outside 0..%d", i, nn, d.K - 1);
    h->n_pins = (uint16_t)np;
    h->n_nets = (uint8_t)nn;
    PinRec* pins = reinterpret_cast<PinRec*>(rec + d.L.rec_pins);
    std::fill(last.begin(), last.begin() + (size_t)nc * 256, -1);
    int rank[kMaxComps] = {0};
    int prev_net = 0;
    for (int q = 0; q < np; ++q) {
      size_t j = i * d.max_pins + q;
      PinRec r;
      r.comp = b->pin_comp[j]; r.rx = b->pin_rx[j]; r.ry = b->pin_ry[j]; r.net = b->pin_net[j];
      r.pin_id = b->pin_id[j]; r.flags = 0; r.rank = 0;
      if (r.comp >= nc) return fail(PCB_ERR_INVALID, "instance %zu pin %d: component %d >= %d", i, q, r.comp, nc);
      if (r.rx >= comps[r.comp].h || r.ry >= comps[r.comp].w)
        return fail(PCB_ERR_INVALID, "instance %zu pin %d: (%d,%d) outside its component", i, q, r.rx, r.ry);
      if (r.net >= nn || r.net < prev_net)
        return fail(PCB_ERR_INVALID, "instance %zu pin %d: pins must be net-major, net %d", i, q, r.net);
      prev_net = r.net;
      int max_row = d.kind == 3 ? d.C * d.L.P : d.L.P;
      if (r.pin_id >= max_row) return fail(PCB_ERR_INVALID, "instance %zu pin %d: pin_id %d >= %d", i, q, r.pin_id, max_row);
      if (rank[r.comp] >= d.L.P)
        return fail(PCB_ERR_INVALID, "instance %zu: component %d has more than %d pins", i, r.comp, d.L.P);
      r.rank = (uint8_t)rank[r.comp]++;
      // last writer of a (component, pin_id) feature row wins (update_all_pins_feature,
      // dummy_env_rectangular_pin.py:1521-1542); the spatial rows are keyed by pin_id alone
      int key = d.kind == 3 ? 0 : r.comp * 256;
      if (d.kind == 3) {
        r.flags = kPinWinner;
      } else {
        last[key + r.pin_id] = q;
      }
      pins[q] = r;
    }
    if (d.kind == 2)
      for (int q = 0; q < np; ++q)
        if (last[pins[q].comp * 256 + pins[q].pin_id] == q) pins[q].flags = kPinWinner;
    if (d.reward_type != PCB_REWARD_CENTROID) {  // the beam router keeps a net in registers
      int run = 0;
      for (int q = 0; q < np; ++q) {
        run = (q > 0 && pins[q].net == pins[q - 1].net) ? run + 1 : 1;
        if (run > kMaxBeamPoints + 1)
          return fail(PCB_ERR_INVALID, "instance %zu: net %d has more than %d pins (beam route limit)",
                      i, pins[q].net, kMaxBeamPoints + 1);
      }
    }
  }
  if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }  // it carries the old bank pointer
  if (e->bank) { cudaFree(e->bank); e->bank = nullptr; }
  void* p = nullptr;
  CU(cudaMalloc(&p, host.size()));
  e->bank = reinterpret_cast<uint8_t*>(p);
  CU(cudaMemcpy(e->bank, host.data(), host.size(), cudaMemcpyHostToDevice));
  e->bank_count = (uint32_t)cnt;
  return 0;
}

}  // namespace

extern "C" {

const char* pcb_last_error(void) { return g_err.c_str(); }
int pcb_abi_version(void) { return PCB_ABI_VERSION; }

/* 0 = runtime-shaped kernels, otherwise the id of the compiled FixedShape in use */
int pcb_shape_id(pcb_env_t* e) { return e ? e->shape_id : -1; }

int pcb_create(const pcb_config_t* cfg, int device, int num_envs, pcb_env_t** out) {
  if (!out) return fail(PCB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = check_config(cfg);
  if (rc) return rc;
  if (num_envs < 1) return fail(PCB_ERR_INVALID, "num_envs must be >= 1");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(PCB_ERR_CUDA, "CUDA device %d not present (%d devices)", device, ndev);
  CU(cudaSetDevice(device));
  pcb_env* e = new pcb_env();
  e->cfg = *cfg;
  e->device = device; e->n_envs = num_envs; e->ring = cfg->obs_ring; e->slice = 0;
  e->wide = cfg->height > 32 || cfg->width > 32;
  e->step_no = 0; e->launches = 0; e->bank = nullptr; e->bank_count = 0; e->timing = false;
  for (int i = 0; i < 4; ++i) { e->timed_ms[i] = 0.0; e->timed_n[i] = 0; }
  {
    const char* g = getenv("PCB_GUARD");
    e->guard = g && g[0] == '1';
  }
  e->ratio_lut = nullptr; e->d_actions = nullptr; e->d_actions2 = nullptr;
  e->pend_actions = nullptr; e->pend_seed = 0; e->pend_step = 0; e->pend_stream = nullptr;
  e->res_stride = 0; e->hs_actions = nullptr; e->hs_reward = nullptr; e->hs_done = nullptr;
  e->hs_steps = 0; e->copy = nullptr; e->ev_host = nullptr;
  e->use_graph = true; e->graph = nullptr; e->graph_len = 0; e->graph_slice = 0; e->graph_step0 = 0;
  e->graph_seed = 0; e->graph_stream = nullptr; e->graph_overlap = false; e->d_step_base = nullptr;
  e->graph_launches = 0; e->graph_replays = 0;
  memset(&e->base, 0, sizeof e->base);
  fill_dims(*cfg, false, e->d);
  e->shape_id = cfg->reserved0 == 1 ? 0 : pick_shape(e->d);  // reserved0 == 1 forces the generic kernels (tests)
  {
    // FixedShape::fast_board: square 16/32/64 boards are expanded straight from registers
    const int H = cfg->height, W = cfg->width;
    if (e->shape_id && H == W && (H == 16 || H == 32 || H == 64)) fill_dims(*cfg, true, e->d);
  }
  const Dims& d = e->d;
  e->smem_bytes = (size_t)kLutBytes + (size_t)d.L.sm_stride * kWarpsPerCta;
  if (e->smem_bytes > 200 * 1024) {
    const size_t need = e->smem_bytes;
    delete e;
    return fail(PCB_ERR_INVALID, "configuration needs %zu bytes of shared memory per CTA (limit 200 KB)", need);
  }
  auto cleanup = [&](int code) { pcb_destroy(e); return code; };
#define TRY(x) do { int rc_ = (x); if (rc_) return cleanup(rc_); } while (0)
  switch (e->shape_id) {
    case 1: TRY((set_smem_shape<1, uint32_t, ShapeB>(e->smem_bytes))); break;
    case 2: TRY((set_smem_shape<1, uint32_t, ShapeC>(e->smem_bytes))); break;
    case 3: TRY((set_smem_shape<1, uint32_t, ShapeD>(e->smem_bytes))); break;
    case 4: TRY((set_smem_shape<2, uint64_t, ShapeE>(e->smem_bytes))); break;
    case 5: TRY((set_smem_shape<1, uint32_t, ShapeRefPin>(e->smem_bytes))); break;
    case 6: TRY((set_smem_shape<1, uint32_t, ShapeRefSpatial>(e->smem_bytes))); break;
    case 7: TRY((set_smem_shape<1, uint32_t, ShapeRefRect>(e->smem_bytes))); break;
    case 8: TRY((set_smem_shape<1, uint32_t, ShapeRefSquare>(e->smem_bytes))); break;
    case 9: TRY((set_smem_shape<1, uint32_t, ShapeA>(e->smem_bytes))); break;
#ifdef PCB_USER_SHAPE
    case 100: TRY((set_smem_shape<kUserRpl, UserWord, ShapeUser>(e->smem_bytes))); break;
#endif
    default:
      if (e->wide) TRY((set_smem_shape<2, uint64_t, DynShape>(e->smem_bytes)));
      else TRY((set_smem_shape<1, uint32_t, DynShape>(e->smem_bytes)));
  }
  e->fused = d.kind <= 1 && !e->wide;
  e->fused_pin = false; e->fused_pin_ok = false; e->pin_smem = 0; e->list_steps = false;
  e->pin_jobs[0] = e->pin_jobs[1] = nullptr;
  e->pin_job_total[0] = e->pin_job_total[1] = nullptr; e->pin_job_steps = 0;
  e->sms = 148;
  cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, device);
  e->small_smem = 0; e->fused_ms = 0.0; e->fused_launches = 0; e->fused_steps = 0;
  if (e->fused) {
    const int per_cta = kWarpsPerCta * (32 / small_lanes_per_env(d.H));
    e->small_smem = (size_t)kLutBytes + (size_t)small_env_smem(d.C, d.H * d.W) * per_cta;
    RolloutParams none;
    memset(&none, 0, sizeof none);
    TRY(small_dispatch(e, true, nullptr, none));
  }
  if (d.kind >= 2 && !e->wide && cfg->reserved0 != 1) {
    // Small boards pack several envs into a warp; a small batch of any board is launch- and
    // latency-bound with one launch per step.  A large batch of a 32-row board keeps the
    // per-step kernels (one warp per env either way, and the reward kernel overlaps the next
    // step there instead of running after the rollout).
    const int lpe = small_lanes_per_env(d.H);
    // measured on config C (profiles/README.md): 4 096 envs 6.0e8 against 3.2e8 env-steps/s
    // per-step, 12 288 envs 6.4e8 against 6.3e8, 16 384 envs 6.2e8 against 6.7e8
    // spatial env (D): 4 096 envs 2.8e8 against 2.1e8, 8 192 envs 2.7e8 against 3.0e8
    int max_envs = d.kind == 3 ? 6144 : 12288;
    // reward types beam / both: always.  Per step their router + reward kernels are the critical
    // path (one side stream, ~100 us of latency-bound work per step); over the job list of a
    // 64-step launch the same kernels run with 64x the parallelism: C "both" at 65 536 envs
    // 4.0e8 -> 5.5e8 env-steps/s, "beam" 4.6e8 -> 6.0e8, at 4 096 envs 7.5e7 -> 4.6e8.
    // list_steps_chunk (step_kernel per step as the PRODUCER of the job list) is off by default:
    // measured slower than rollout_pin_kernel at every side-grid size (C "both" at 65 536 envs
    // 5.56e8 against 5.69e8, "beam" 6.11e8 against 6.27e8; profiles/r2h_list_steps_producer.log)
    // -- over a job list the router and the reward kernel are bound by instruction issue, so
    // they take their time from whichever stepper runs beside them.  PCB_LIST_STEPS_MIN_ENVS
    // turns it on from that batch size (tests, experiments).
    int list_min = 0x7fffffff;
    if (const char* f = getenv("PCB_LIST_STEPS_MIN_ENVS")) list_min = atoi(f);
    if (d.reward_type != PCB_REWARD_CENTROID) max_envs = 0x7fffffff;
    if (const char* f = getenv("PCB_PIN_FUSED_MAX_ENVS")) max_envs = atoi(f);
    e->fused_pin_ok = true;
    e->fused_pin = lpe < 32 || num_envs <= max_envs;
    e->list_side_mult = PCB_LIST_SIDE_MULT;
    if (const char* f = getenv("PCB_LIST_SIDE_MULT")) e->list_side_mult = std::max(1, atoi(f));  // experiments
    e->list_steps = PCB_LIST_STEPS != 0 && lpe == 32 && num_envs >= list_min && e->fused_pin &&
                    d.reward_type != PCB_REWARD_CENTROID;
    const PinSmem L = pin_smem_layout(d.kind, d.C, d.L.P, d.max_pins, d.H, d.W, d.K, d.mh, d.mw);
    e->pin_smem = (size_t)kLutBytes + (size_t)L.total * kWarpsPerCta * (32 / lpe);
    if (e->pin_smem > 200 * 1024) { e->fused_pin_ok = false; e->fused_pin = false; }
    if (e->fused_pin_ok) {
      PinRolloutParams none;
      memset(&none, 0, sizeof none);
      TRY(pin_dispatch(e, true, nullptr, none));
    }
  }
  e->jobs = nullptr; e->job_count = nullptr; e->reward_smem = 0; e->reward_grid = 1; e->reward_grid_side = 1;
  e->reward_half = false; e->list_group = 16; e->step_group = 16; e->list_half_beam = false;
  e->side = nullptr; e->rew_pending[0] = e->rew_pending[1] = false; e->overlap = PCB_OVERLAP_REWARD != 0;
  for (int i = 0; i < 2; ++i) { e->ev_step[i] = nullptr; e->ev_rew[i] = nullptr; }
  if (d.kind >= 2) {
    e->reward_smem = (size_t)((route_scratch_bytes(d.max_pins, d.K) + 15) & ~15) * kWarpsPerCta;
    TRY(set_smem(reward_kernel<false>, e->reward_smem));
    e->reward_half = PCB_REWARD_HALF != 0 && d.reward_type == PCB_REWARD_CENTROID && d.max_pins <= 32;
    // reward types beam / both over the job list of the multi-step path: two (four) jobs per warp too
    e->list_half_beam = PCB_REWARD_HALF != 0 && d.reward_type != PCB_REWARD_CENTROID && d.max_pins <= 32;
    if (e->list_half_beam) {
      e->list_group = reward_list_group_lanes(d.max_pins);
      TRY((set_smem(reward_list_half_kernel<16, true>, e->reward_smem * 2)));
      TRY((set_smem(reward_list_half_kernel<8, true>, e->reward_smem * 4)));
    }
    if (e->reward_half) {
      e->list_group = reward_list_group_lanes(d.max_pins);
      e->step_group = reward_step_group_lanes(d.kind);
      TRY(set_smem(reward_half_kernel<16>, e->reward_smem * 2));
      TRY(set_smem(reward_half_kernel<8>, e->reward_smem * 4));
      TRY(set_smem(reward_list_half_kernel<16>, e->reward_smem * 2));
      TRY(set_smem(reward_list_half_kernel<8>, e->reward_smem * 4));
    }
    TRY(set_smem(reward_kernel<true>, e->reward_smem));
    TRY(set_smem(reward_list_kernel<false>, e->reward_smem));
    TRY(set_smem(reward_list_kernel<true>, e->reward_smem));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int want = (num_envs + kWarpsPerCta - 1) / kWarpsPerCta;
#ifndef PCB_REWARD_GRID_MULT
#define PCB_REWARD_GRID_MULT 16
#endif
    e->reward_grid = std::max(1, std::min(want, sms * PCB_REWARD_GRID_MULT));
#ifndef PCB_REWARD_SIDE_MULT
#define PCB_REWARD_SIDE_MULT 4
#endif
    e->reward_grid_side = std::max(1, std::min(want, sms * PCB_REWARD_SIDE_MULT));
    int lo_pri = 0, hi_pri = 0;
    cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
    cudaError_t se = cudaStreamCreateWithPriority(&e->side, cudaStreamNonBlocking, hi_pri);
    for (int i = 0; i < 2 && se == cudaSuccess; ++i) {
      se = cudaEventCreateWithFlags(&e->ev_step[i], cudaEventDisableTiming);
      if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->ev_rew[i], cudaEventDisableTiming);
    }
    if (se != cudaSuccess) return cleanup(fail(PCB_ERR_CUDA, "side stream/events: %s", cudaGetErrorString(se)));
  }
  const size_t N = (size_t)num_envs, R = (size_t)e->ring;
  const int64_t n64 = num_envs, r64 = e->ring;
  TRY(dev_alloc(e, &e->state, N * d.L.st_stride));
  // State records of 195 MB (E, 131 072 envs) never survive in the 126 MB L2 from one step to the
  // next; loaded and stored with the streaming policy they leave it to the observation stores:
  // E 1.947e8 -> 2.008e8 env-steps/s.  Records that mostly fit are better left alone (C, 24 MB:
  // 7.27e8 -> 6.96e8 with streaming; D neutral) -- profiles/r2h_state_streaming_ab.log.
  e->state_stream = e->wide && (size_t)N * d.L.st_stride > kStateStreamBytes;
  if (const char* f = getenv("PCB_STATE_STREAM")) e->state_stream = atoi(f) != 0;  // experiments
  TRY(dev_alloc(e, &e->stats, kStatSlots * 8 * sizeof(double)));
  TRY(dev_alloc(e, &e->d_actions, N * 3 * sizeof(int32_t)));
  TRY(dev_alloc(e, &e->d_actions2, N * 3 * sizeof(int32_t)));
  TRY(dev_alloc(e, &e->d_step_base, sizeof(uint64_t)));
  if (d.kind >= 2) {
    TRY(dev_alloc(e, &e->jobs, 2 * N * (size_t)job_stride(d.max_pins, d.K)));
    TRY(dev_alloc(e, &e->job_count, 4 * sizeof(uint32_t)));
  }
  {
    // area ratio h*w/(H*W): float64 quotient rounded to float32, i.e. np.float32(reference value)
    const int nlut = d.mh * d.mw + 1;
    std::vector<float> lut(nlut > 0 ? nlut : 1, 0.f);
    for (int a = 0; a < nlut; ++a) lut[a] = (float)((double)a / (double)(d.H * d.W));
    TRY(dev_alloc(e, &e->ratio_lut, lut.size() * sizeof(float)));
    cudaError_t ce = cudaMemcpy(e->ratio_lut, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) return cleanup(fail(PCB_ERR_CUDA, "cudaMemcpy(ratio_lut): %s", cudaGetErrorString(ce)));
  }
  auto add = [&](const char* key, void** ptr, int code, int bits, std::vector<int64_t> shape, bool ringed) {
    e->tensors.push_back(TensorDesc{key, ptr, code, bits, std::move(shape), ringed});
  };
  TRY(dev_alloc(e, &e->base.grid, R * N * d.L.grid_b));
  add("grid", (void**)&e->base.grid, kDLUInt, 8, {r64, n64, d.H, d.W}, true);
  TRY(dev_alloc(e, &e->base.mask, R * N * d.L.mask_b));
  if (d.kind == 0) add("action_mask", (void**)&e->base.mask, kDLUInt, 8, {r64, n64, d.H, d.W}, true);
  else add("action_mask", (void**)&e->base.mask, kDLUInt, 8, {r64, n64, d.L.O, d.H, d.W}, true);
  if (d.kind >= 1) {
    TRY(dev_alloc(e, &e->base.feat, R * N * d.L.feat_n * sizeof(float)));
    add("all_components_feature", (void**)&e->base.feat, kDLFloat, 32, {r64, n64, d.C, d.L.F}, true);
    TRY(dev_alloc(e, &e->base.place_mask, R * N * d.C));
    add("placement_mask", (void**)&e->base.place_mask, kDLUInt, 8, {r64, n64, d.C}, true);
  }
  if (d.kind == 1) {
    TRY(dev_alloc(e, &e->base.comp_mask, R * N * d.C));
    add("component_mask", (void**)&e->base.comp_mask, kDLUInt, 8, {r64, n64, d.C}, true);
  }
  if (d.kind >= 2) {
    TRY(dev_alloc(e, &e->base.pnum, R * N * d.L.pnum_n * sizeof(int16_t)));
    TRY(dev_alloc(e, &e->base.pcat, R * N * d.L.pcat_n * sizeof(int16_t)));
    if (d.kind == 2) {
      add("all_pins_num_feature", (void**)&e->base.pnum, kDLInt, 16, {r64, n64, d.C, d.L.P, 4}, true);
      add("all_pins_cat_feature", (void**)&e->base.pcat, kDLInt, 16, {r64, n64, d.C, d.L.P, 1}, true);
    } else {
      add("all_pins_num_feature", (void**)&e->base.pnum, kDLInt, 16, {r64, n64, d.C * d.L.P + 1, 4}, true);
      add("all_pins_cat_feature", (void**)&e->base.pcat, kDLInt, 16, {r64, n64, d.C * d.L.P + 1, 2}, true);
      TRY(dev_alloc(e, &e->base.pgrid, R * N * d.L.pgrid_b));
      add("pin_grid", (void**)&e->base.pgrid, kDLUInt, 8, {r64, n64, d.H, d.W, d.K}, true);
      TRY(dev_alloc(e, &e->base.cgrid, R * N * d.L.cgrid_b));
      add("component_grid", (void**)&e->base.cgrid, kDLUInt, 8, {r64, n64, d.C, d.mh, d.mw, d.K}, true);
    }
  }
  {
    e->res_stride = (N * 9 + 15) & ~(size_t)15;
    uint8_t* blk = nullptr;
    TRY(dev_alloc(e, &blk, 2 * e->res_stride));
    e->base.reward = reinterpret_cast<double*>(blk);
    e->base.done = blk + N * sizeof(double);
    add("reward", (void**)&e->base.reward, kDLFloat, 64, {2, n64}, false);
    e->tensors.back().strides = {(int64_t)(e->res_stride / sizeof(double)), 1};
    add("done", (void**)&e->base.done, kDLUInt, 8, {2, n64}, false);
    e->tensors.back().strides = {(int64_t)e->res_stride, 1};
  }
  TRY(dev_alloc(e, &e->base.wl, 2 * N * sizeof(double)));
  add("wirelength", (void**)&e->base.wl, kDLFloat, 64, {2, n64}, false);
  TRY(dev_alloc(e, &e->base.ni, 2 * N * sizeof(double)));
  add("num_intersections", (void**)&e->base.ni, kDLFloat, 64, {2, n64}, false);
  TRY(dev_alloc(e, &e->base.ep_return, 2 * N * sizeof(double)));
  add("episode_return", (void**)&e->base.ep_return, kDLFloat, 64, {2, n64}, false);
  TRY(dev_alloc(e, &e->base.ep_length, 2 * N * sizeof(int32_t)));
  add("episode_length", (void**)&e->base.ep_length, kDLInt, 32, {2, n64}, false);
#undef TRY
  *out = e;
  return 0;
}

int pcb_destroy(pcb_env_t* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  for (int i = 0; i < 2; ++i) {
    if (e->ev_step[i]) cudaEventDestroy(e->ev_step[i]);
    if (e->ev_rew[i]) cudaEventDestroy(e->ev_rew[i]);
  }
  if (e->side) cudaStreamDestroy(e->side);
  if (e->graph) cudaGraphExecDestroy(e->graph);
  if (e->copy) cudaStreamDestroy(e->copy);
  if (e->ev_host) cudaEventDestroy(e->ev_host);
  for (cudaEvent_t ev : e->hs_events) cudaEventDestroy(ev);
  if (e->pin_jobs[0]) cudaFree(e->pin_jobs[0]);
  if (e->hs_actions) cudaFree(e->hs_actions);
  if (e->hs_reward) cudaFree(e->hs_reward);
  if (e->hs_done) cudaFree(e->hs_done);
  for (auto& t : e->timed) {  // kernel timing enabled but never polled
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  for (void* p : e->allocs) cudaFree(p);
  if (e->bank) cudaFree(e->bank);
  delete e;
  return 0;
}

int pcb_upload_instances(pcb_env_t* e, const pcb_instance_bank_t* host_bank) {
  if (!e) return fail(PCB_ERR_INVALID, "env is NULL");
  CU(cudaSetDevice(e->device));
  return pack_and_upload(e, host_bank);
}

int pcb_generate_instances(pcb_env_t* e, const int32_t gen[12], uint64_t seed, int64_t count) {
  if (!e || !gen) return fail(PCB_ERR_INVALID, "NULL argument");
  if (e->d.kind == 0) return fail(PCB_ERR_INVALID, "the square env has no instances");
  if (count < 1) return fail(PCB_ERR_INVALID, "count must be >= 1");
  CU(cudaSetDevice(e->device));
  GenBank gb;
  std::string why;
  if (!generate_bank(gen, e->d.kind, e->d.C, e->d.max_pins, e->d.K - 1, seed, count, gb, why))
    return fail(PCB_ERR_INVALID, "instance generator: %s", why.c_str());
  pcb_instance_bank_t b;
  memset(&b, 0, sizeof b);
  b.count = count;
  b.n_comp = gb.n_comp.data(); b.comp_h = gb.comp_h.data(); b.comp_w = gb.comp_w.data();
  if (e->d.kind >= 2) {
    b.n_pins = gb.n_pins.data(); b.n_nets = gb.n_nets.data();
    b.pin_comp = gb.pin_comp.data(); b.pin_rx = gb.pin_rx.data(); b.pin_ry = gb.pin_ry.data();
    b.pin_net = gb.pin_net.data(); b.pin_id = gb.pin_id.data();
  }
  return pack_and_upload(e, &b);
}

int pcb_generate_instances_device(pcb_env_t* e, const int32_t gen[12], uint64_t seed, int64_t count,
                                  void* stream) {
  if (!e || !gen) return fail(PCB_ERR_INVALID, "NULL argument");
  if (e->d.kind == 0) return fail(PCB_ERR_INVALID, "the square env has no instances");
  if (count < 1 || count > 0x7fffffffLL) return fail(PCB_ERR_INVALID, "count out of range");
  CU(cudaSetDevice(e->device));
  {  // validate the ranges with the host generator (one throw-away instance)
    GenBank gb;
    std::string why;
    if (!generate_bank(gen, e->d.kind, e->d.C, e->d.max_pins, e->d.K - 1, seed, 1, gb, why))
      return fail(PCB_ERR_INVALID, "instance generator: %s", why.c_str());
  }
  const size_t bytes = (size_t)count * e->d.L.rec_stride;
  if (!e->bank || e->bank_count != (uint32_t)count) {
    CU(cudaStreamSynchronize((cudaStream_t)stream));  // nobody may still read the old bank
    if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
    if (e->bank) { cudaFree(e->bank); e->bank = nullptr; }
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes));
    e->bank = reinterpret_cast<uint8_t*>(p);
    e->bank_count = (uint32_t)count;
  }
  GenParams g;
  g.kind = e->d.kind; g.C = e->d.C; g.P = e->d.L.P; g.max_pins = e->d.max_pins; g.K = e->d.K;
  g.rec_stride = e->d.L.rec_stride; g.rec_comps = e->d.L.rec_comps; g.rec_pins = e->d.L.rec_pins;
  g.min_w = gen[0]; g.max_w = gen[1]; g.min_h = gen[2]; g.max_h = gen[3]; g.min_c = gen[4];
  g.max_c = gen[5]; g.min_nets = gen[6]; g.max_nets = gen[7]; g.lo = gen[8]; g.hi = gen[9];
  g.nd = std::max(0, std::min(9, gen[10])); g.ps = std::max(0, std::min(9, gen[11]));
  g.seed = seed; g.count = (uint32_t)count; g.bank = e->bank;
  generate_kernel<<<(unsigned)((count + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g);
  CU(cudaGetLastError());
  e->launches += 1;
  return 0;
}

int pcb_generate_host(int kind, int max_num_components, int max_pins, int max_num_nets,
                      const int32_t gen[12], uint64_t seed, int64_t count, int32_t* n_comp,
                      uint8_t* comp_h, uint8_t* comp_w, int32_t* n_pins, int32_t* n_nets,
                      uint8_t* pin_comp, uint8_t* pin_rx, uint8_t* pin_ry, uint8_t* pin_net,
                      uint16_t* pin_id) {
  if (!gen || !n_comp || !comp_h || !comp_w) return fail(PCB_ERR_INVALID, "NULL argument");
  if (kind < 1 || kind > 3 || count < 1) return fail(PCB_ERR_INVALID, "bad kind or count");
  GenBank gb;
  std::string why;
  if (!generate_bank(gen, kind, max_num_components, max_pins, max_num_nets, seed, count, gb, why))
    return fail(PCB_ERR_INVALID, "instance generator: %s", why.c_str());
  memcpy(n_comp, gb.n_comp.data(), gb.n_comp.size() * sizeof(int32_t));
  memcpy(comp_h, gb.comp_h.data(), gb.comp_h.size());
  memcpy(comp_w, gb.comp_w.data(), gb.comp_w.size());
  if (kind >= 2) {
    if (!n_pins || !n_nets || !pin_comp || !pin_rx || !pin_ry || !pin_net || !pin_id)
      return fail(PCB_ERR_INVALID, "NULL pin array");
    memcpy(n_pins, gb.n_pins.data(), gb.n_pins.size() * sizeof(int32_t));
    memcpy(n_nets, gb.n_nets.data(), gb.n_nets.size() * sizeof(int32_t));
    memcpy(pin_comp, gb.pin_comp.data(), gb.pin_comp.size());
    memcpy(pin_rx, gb.pin_rx.data(), gb.pin_rx.size());
    memcpy(pin_ry, gb.pin_ry.data(), gb.pin_ry.size());
    memcpy(pin_net, gb.pin_net.data(), gb.pin_net.size());
    memcpy(pin_id, gb.pin_id.data(), gb.pin_id.size() * sizeof(uint16_t));
  }
  return 0;
}

static int ready(pcb_env_t* e) {
  if (!e) return fail(PCB_ERR_INVALID, "env is NULL");
  if (e->d.kind != 0 && !e->bank) return fail(PCB_ERR_STATE, "no instances uploaded (pcb_upload_instances / pcb_generate_instances)");
  return 0;
}

int pcb_reset(pcb_env_t* e, const uint8_t* d_reset_mask, void* stream) {
  int rc = ready(e);
  if (rc) return rc;
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  e->pend_actions = nullptr;
  if (!d_reset_mask) {
    e->slice = 0;
    e->step_no = 0;
    CU(cudaMemsetAsync(e->stats, 0, kStatSlots * 8 * sizeof(double), st));
    if (e->job_count) CU(cudaMemsetAsync(e->job_count, 0, 4 * sizeof(uint32_t), st));
  }
  Params p = make_params(e, true);
  p.reset_mask = d_reset_mask;
  p.reset_first = d_reset_mask ? 0 : 1;
  return launch(e, K_RESET, p, st);
}

// overlap = false: step kernel, then the reward kernel, both on `stream`.
// overlap = true (pcb_rollout_random): the reward kernel of step t runs on the side
// stream while the step kernel of step t+1 runs on `stream`; step t+2 waits for it.
static int rollout_fused(pcb_env_t* e, uint64_t seed, int n_steps, int32_t* d_actions_log,
                         double* d_reward_log, uint8_t* d_done_log, cudaStream_t st,
                         const int32_t* d_actions, int action_format);

// reward_row / done_row != nullptr: this step's reward[N] / done[N] go there instead of the
// double-buffered result tensors (pcb_rollout_random_logged)
static int step_impl(pcb_env_t* e, const int32_t* d_actions, int action_format, uint64_t seed,
                     int32_t* d_next_actions, void* stream, bool overlap,
                     double* reward_row = nullptr, uint8_t* done_row = nullptr,
                     const uint64_t* step_base = nullptr, bool last_of_call = false) {
  int rc = ready(e);
  if (rc) return rc;
  if (!d_actions) return fail(PCB_ERR_INVALID, "d_actions is NULL");
  if (action_format != PCB_ACTION_TUPLE && action_format != PCB_ACTION_FLAT)
    return fail(PCB_ERR_INVALID, "unknown action format %d", action_format);
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  e->pend_actions = nullptr;
  if (e->fused && !d_next_actions)  // pin-less kinds: the multi-env-per-warp kernel, one step
    return rollout_fused(e, seed, 1, nullptr, reward_row, done_row, st, d_actions, action_format);
  const bool routes = e->d.kind >= 2;
  const int par = (int)(e->step_no & 1);
  // the reward kernel of step t-2 used the same job buffer / scalar half
  if (routes && overlap && e->step_no >= 2 && e->rew_pending[par]) {
    CU(cudaStreamWaitEvent(st, e->ev_rew[par], 0));
    e->rew_pending[par] = false;
  }
  e->slice = (e->slice + 1) % e->ring;
  Params p = make_params(e);
  p.actions = d_actions;
  p.action_format = action_format;
  p.seed = seed;
  p.actions_out = d_next_actions;
  p.step_base = step_base;
  if (reward_row) p.out.reward = reward_row;
  if (done_row) p.out.done = done_row;
  rc = launch(e, K_STEP, p, st);
  if (rc) return rc;
  if (routes) {
    if (!overlap) {
      rc = launch(e, K_TERM, p, st);
    } else {
      CU(cudaEventRecord(e->ev_step[par], st));
      CU(cudaStreamWaitEvent(e->side, e->ev_step[par], 0));
      // no step kernel runs beside the reward kernel of a call's last step: full grid
      rc = launch(e, K_TERM, p, e->side, last_of_call ? e->reward_grid : e->reward_grid_side);
      if (rc) return rc;
      CU(cudaEventRecord(e->ev_rew[par], e->side));
      e->rew_pending[par] = true;
    }
  }
  e->step_no += 1;
  return rc;
}

// joins the overlapped reward kernels into `stream`
static int join_rewards(pcb_env_t* e, cudaStream_t st) {
  for (int i = 0; i < 2; ++i)
    if (e->rew_pending[i]) {
      CU(cudaStreamWaitEvent(st, e->ev_rew[i], 0));
      e->rew_pending[i] = false;
    }
  return 0;
}

int pcb_step(pcb_env_t* e, const int32_t* d_actions, int action_format, void* stream) {
  return step_impl(e, d_actions, action_format, 0, nullptr, stream, false);
}

int pcb_sample_random_actions(pcb_env_t* e, uint64_t seed, int32_t* d_actions, void* stream) {
  int rc = ready(e);
  if (rc) return rc;
  if (!d_actions) return fail(PCB_ERR_INVALID, "d_actions is NULL");
  CU(cudaSetDevice(e->device));
  Params p = make_params(e);
  p.seed = seed;
  p.actions_out = d_actions;
  return launch(e, K_SAMPLE, p, (cudaStream_t)stream);
}

static int step_host_impl(pcb_env_t* e, const int32_t* h_actions, int action_format, double* h_reward,
                          uint8_t* h_done, bool packed, void* stream) {
  int rc = ready(e);
  if (rc) return rc;
  if (!h_actions || !h_reward || (!packed && !h_done)) return fail(PCB_ERR_INVALID, "NULL host buffer");
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  size_t per = action_format == PCB_ACTION_FLAT ? 1 : (e->d.kind == 0 ? 2 : 3);
  CU(cudaMemcpyAsync(e->d_actions, h_actions, (size_t)e->n_envs * per * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  rc = pcb_step(e, e->d_actions, action_format, stream);
  if (rc) return rc;
  // the half the step just wrote: [reward f64 x N | done u8 x N] contiguous on the device
  const size_t N = (size_t)e->n_envs;
  const uint8_t* blk = reinterpret_cast<const uint8_t*>(e->base.reward) + (size_t)((e->step_no + 1) & 1) * e->res_stride;
  if (packed) {
    CU(cudaMemcpyAsync(h_reward, blk, N * 9, cudaMemcpyDeviceToHost, st));
  } else {
    CU(cudaMemcpyAsync(h_reward, blk, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_done, blk + N * sizeof(double), N, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  return 0;
}

int pcb_step_host(pcb_env_t* e, const int32_t* h_actions, int action_format, double* h_reward,
                  uint8_t* h_done, void* stream) {
  return step_host_impl(e, h_actions, action_format, h_reward, h_done, false, stream);
}

int pcb_step_host_packed(pcb_env_t* e, const int32_t* h_actions, int action_format, void* h_result,
                         void* stream) {
  return step_host_impl(e, h_actions, action_format, reinterpret_cast<double*>(h_result), nullptr, true, stream);
}

int pcb_get_tensor(pcb_env_t* e, const char* key, DLManagedTensor** out) {
  if (!e || !key || !out) return fail(PCB_ERR_INVALID, "NULL argument");
  for (const TensorDesc& t : e->tensors) {
    if (strcmp(t.key, key)) continue;
    DLManagedTensor* m = (DLManagedTensor*)calloc(1, sizeof(DLManagedTensor));
    int64_t* shape = (int64_t*)malloc(sizeof(int64_t) * t.shape.size());
    if (!m || !shape) { free(m); free(shape); return fail(PCB_ERR_NOMEM, "out of host memory"); }
    for (size_t i = 0; i < t.shape.size(); ++i) shape[i] = t.shape[i];
    m->dl_tensor.data = *t.ptr;
    m->dl_tensor.device = DLDevice{kDLCUDA, e->device};
    m->dl_tensor.ndim = (int32_t)t.shape.size();
    m->dl_tensor.dtype = DLDataType{(uint8_t)t.code, (uint8_t)t.bits, 1};
    m->dl_tensor.shape = shape;
    m->dl_tensor.strides = nullptr;
    if (!t.strides.empty()) {  // shape and strides share one allocation (freed by the deleter)
      int64_t* both = (int64_t*)realloc(shape, sizeof(int64_t) * 2 * t.shape.size());
      if (!both) { free(m); free(shape); return fail(PCB_ERR_NOMEM, "out of host memory"); }
      for (size_t i = 0; i < t.shape.size(); ++i) both[t.shape.size() + i] = t.strides[i];
      m->dl_tensor.shape = both;
      m->dl_tensor.strides = both + t.shape.size();
    }
    m->dl_tensor.byte_offset = 0;
    m->manager_ctx = nullptr;
    m->deleter = dl_deleter;
    *out = m;
    return 0;
  }
  return fail(PCB_ERR_INVALID, "no tensor named '%s' for this environment kind", key);
}

int pcb_current_slice(pcb_env_t* e) { return e ? e->slice : fail(PCB_ERR_INVALID, "env is NULL"); }

int pcb_current_result_half(pcb_env_t* e) {
  return e ? (int)((e->step_no + 1) & 1) : fail(PCB_ERR_INVALID, "env is NULL");
}

int pcb_episode_stats(pcb_env_t* e, double out[8], int reset_counters, void* stream) {
  if (!e || !out) return fail(PCB_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  double host[kStatSlots * 8];
  CU(cudaMemcpyAsync(host, e->stats, sizeof host, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int j = 0; j < 8; ++j) out[j] = 0.0;
  for (int s = 0; s < kStatSlots; ++s)
    for (int j = 0; j < 8; ++j) out[j] += host[s * 8 + j];
  out[5] = (double)e->step_no * (double)e->n_envs;
  if (reset_counters) CU(cudaMemsetAsync(e->stats, 0, sizeof host, st));
  return 0;
}

int pcb_get_occupancy(pcb_env_t* e, uint64_t* h_rows, void* stream) {
  if (!e || !h_rows) return fail(PCB_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const Dims& d = e->d;
  std::vector<uint8_t> host((size_t)e->n_envs * d.L.st_stride);
  CU(cudaMemcpyAsync(host.data(), e->state, host.size(), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < e->n_envs; ++i) {
    const uint8_t* occ = host.data() + (size_t)i * d.L.st_stride + d.L.st_occ;
    for (int r = 0; r < 64; ++r) {
      uint64_t v = 0;
      if (e->wide) v = reinterpret_cast<const uint64_t*>(occ)[r];
      else if (r < 32) v = reinterpret_cast<const uint32_t*>(occ)[r];
      h_rows[(size_t)i * 64 + r] = v;
    }
  }
  return 0;
}

int pcb_state_layout(pcb_env_t* e, int32_t out[8]) {
  if (!e || !out) return fail(PCB_ERR_INVALID, "NULL argument");
  const Dims& d = e->d;
  out[0] = d.L.st_stride; out[1] = d.L.st_comps; out[2] = d.L.st_pins; out[3] = d.L.st_occ;
  out[4] = d.C; out[5] = d.max_pins; out[6] = e->wide ? 8 : 4; out[7] = e->wide ? 64 : 32;
  return 0;
}

int pcb_get_state(pcb_env_t* e, void* h_state, uint64_t* step_no, void* stream) {
  if (!e || !h_state) return fail(PCB_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(h_state, e->state, (size_t)e->n_envs * e->d.L.st_stride, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (step_no) *step_no = e->step_no;
  return 0;
}

int pcb_set_state(pcb_env_t* e, const void* h_state, uint64_t step_no, void* stream) {
  if (!e || !h_state) return fail(PCB_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(e->state, h_state, (size_t)e->n_envs * e->d.L.st_stride, cudaMemcpyHostToDevice, st));
  if (e->job_count) CU(cudaMemsetAsync(e->job_count, 0, 4 * sizeof(uint32_t), st));
  CU(cudaStreamSynchronize(st));
  e->step_no = step_no;
  e->pend_actions = nullptr;
  return 0;
}

int64_t pcb_launch_count(pcb_env_t* e) { return e ? e->launches : 0; }

// the pin-less kinds: all n_steps in one launch of rollout_small_kernel (pcb_small.cuh)
// d_actions != nullptr: one step with the caller's actions (pcb_step) in the same kernel.
static int rollout_fused(pcb_env_t* e, uint64_t seed, int n_steps, int32_t* d_actions_log,
                         double* d_reward_log, uint8_t* d_done_log, cudaStream_t st,
                         const int32_t* d_actions, int action_format) {
  CU(cudaSetDevice(e->device));
  RolloutParams p;
  memset(&p, 0, sizeof p);
  p.d = e->d;
  p.base = e->base;
  p.state = e->state; p.bank = e->bank; p.ratio_lut = e->ratio_lut; p.stats = e->stats;
  p.actions = d_actions; p.action_format = action_format;
  p.action_stride = d_actions ? (size_t)e->n_envs * (action_format == PCB_ACTION_FLAT ? 1 : (e->d.kind == 0 ? 2 : 3)) : 0;
  p.res_stride = (uint32_t)e->res_stride;
  p.actions_log = d_actions_log;
  p.reward_log = d_reward_log; p.done_log = d_done_log;
  p.seed = seed; p.step_no = e->step_no; p.bank_count = e->bank_count;
  p.slot_step = e->bank_count ? (uint32_t)e->n_envs % e->bank_count : (uint32_t)e->n_envs;
  p.n_envs = e->n_envs; p.n_steps = n_steps; p.ring = e->ring; p.slice0 = e->slice;
  cudaEvent_t a = nullptr, b = nullptr;
  if (e->timing) {
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, st));
  }
  int rc = small_dispatch(e, false, st, p);
  if (rc) return rc;
  CU(cudaGetLastError());
  if (e->timing) {
    CU(cudaEventRecord(b, st));
    CU(cudaEventSynchronize(b));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    e->fused_ms += ms; e->fused_launches += 1; e->fused_steps += n_steps;
  }
  e->launches += 1;
  e->slice = (e->slice + n_steps) % e->ring;
  e->step_no += (uint64_t)n_steps;
  return 0;
}

static int restore_result_halves(pcb_env_t* e, int n_steps, const double* d_reward_log,
                                 const uint8_t* d_done_log, cudaStream_t st);
// One chunk of the multi-step pin path produced by per-step kernels: nt launches of step_kernel
// that leave their routing jobs in the chunk's job list (Params::job_slot_mask = 0, job_t = step
// of the chunk) instead of one launch of rollout_pin_kernel.  Used for the reward types beam /
// both at batch sizes where one warp per env and a launch per step is the faster stepper (the
// router and the reward kernel then run over the list beside the next chunk, as for
// rollout_pin_kernel).  `cur` holds the actions of the next step of the random policy.
static int list_steps_chunk(pcb_env_t* e, uint64_t seed, int t0, int nt, int n_steps, int buf,
                            int32_t* d_actions_log, double* d_reward_log, uint8_t* d_done_log,
                            cudaStream_t st, const int32_t* d_actions, int action_format, size_t per,
                            int32_t** cur) {
  const size_t N = (size_t)e->n_envs;
  for (int t = 0; t < nt; ++t) {
    const int g = t0 + t;  // step of the call
    e->slice = (e->slice + 1) % e->ring;
    Params p = make_params(e);
    int32_t* nxt = nullptr;
    if (d_actions) {
      p.actions = d_actions + (size_t)g * per;
      p.action_format = action_format;
    } else {
      p.actions = *cur;
      p.action_format = PCB_ACTION_TUPLE;
      if (g + 1 < n_steps)
        nxt = d_actions_log ? d_actions_log + (size_t)(g + 1) * N * 3
                            : (*cur == e->d_actions ? e->d_actions2 : e->d_actions);
    }
    p.seed = seed;
    p.actions_out = nxt;
    if (d_reward_log) p.out.reward = d_reward_log + (size_t)g * N;
    if (d_done_log) p.out.done = d_done_log + (size_t)g * N;
    p.jobs = e->pin_jobs[buf]; p.job_count = e->pin_job_total[buf];
    p.job_slot_mask = 0; p.job_t = (uint32_t)t;
    int rc = launch(e, K_STEP, p, st);
    if (rc) return rc;
    e->step_no += 1;
    if (nxt) *cur = nxt;
  }
  return 0;
}

// the pin env: n_steps in launches of up to pin_job_steps steps of rollout_pin_kernel, each
// followed by reward_list_kernel over the routing jobs it left -- on the side stream, so the
// rewards of launch k are computed while launch k+1 steps (two job lists alternate).  Cutting a
// call into more, shorter launches to overlap more was measured and is slower (C at 4 096 envs
// 6.0e8 -> 5.4e8 env-steps/s with 4 launches per 128 steps).  d_actions != nullptr: the
// caller's action rows instead of the random policy (pcb_steps_host).
static int rollout_pin_fused(pcb_env_t* e, uint64_t seed, int n_steps, int32_t* d_actions_log,
                             double* d_reward_log, uint8_t* d_done_log, cudaStream_t st,
                             const int32_t* d_actions, int action_format) {
  CU(cudaSetDevice(e->device));
  const size_t N = (size_t)e->n_envs;
  const size_t jstride = (size_t)job_stride(e->d.max_pins, e->d.K);
  // a list must hold a job per env and step (every env can finish an episode every step when
  // its instance has one component): at most 256 MB each, at most 64 steps per launch
  const int want = (int)std::max<size_t>(1, std::min<size_t>(64, (256u << 20) / (N * jstride)));
  if (!e->pin_jobs[0]) {
    const size_t list = (size_t)want * N * jstride;
    void* q = nullptr;
    cudaError_t err = cudaMalloc(&q, 2 * list + 32);
    if (err != cudaSuccess)
      return fail(err == cudaErrorMemoryAllocation ? PCB_ERR_NOMEM : PCB_ERR_CUDA, "job lists: %s", cudaGetErrorString(err));
    e->pin_jobs[0] = (uint8_t*)q;
    e->pin_jobs[1] = e->pin_jobs[0] + list;
    e->pin_job_total[0] = reinterpret_cast<uint32_t*>(e->pin_jobs[0] + 2 * list);
    e->pin_job_total[1] = e->pin_job_total[0] + 4;
    e->pin_job_steps = want;
  }
  const int chunk = e->pin_job_steps;
  // the reward kernel of launch k runs on the side stream beside launch k+1 (measured with
  // PCB_PIN_OVERLAP=0/1: C at 4 096 envs 5.9e8 -> 6.4e8, REF_pin 1.09e9 -> 1.13e9, C at 8 192
  // envs 6.4e8 -> 6.2e8)
  bool overlap = e->overlap;
  if (const char* f = getenv("PCB_PIN_OVERLAP")) overlap = atoi(f) != 0;  // experiments
  const size_t per = N * (action_format == PCB_ACTION_FLAT ? 1 : 3);
  cudaEvent_t a = nullptr, b = nullptr;
  if (e->timing) {
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, st));
  }
  int32_t* cur = nullptr;  // list_steps: the actions of the next step of the random policy
  if (e->list_steps && !d_actions) {
    cur = d_actions_log ? d_actions_log : e->d_actions;
    int rc = pcb_sample_random_actions(e, seed, cur, st);
    if (rc) return rc;
  }
  e->pend_actions = nullptr;
  int k = 0;
  for (int t0 = 0; t0 < n_steps; t0 += chunk, ++k) {
    const int nt = std::min(chunk, n_steps - t0), buf = k & 1;
    if (e->rew_pending[buf]) {  // the reward kernel of launch k-2 still reads this list
      CU(cudaStreamWaitEvent(st, e->ev_rew[buf], 0));
      e->rew_pending[buf] = false;
    }
    CU(cudaMemsetAsync(e->pin_job_total[buf], 0, sizeof(uint32_t), st));
    PinRolloutParams p;
    memset(&p, 0, sizeof p);
    p.d = e->d; p.base = e->base;
    p.state = e->state; p.bank = e->bank; p.ratio_lut = e->ratio_lut; p.stats = e->stats;
    p.actions = d_actions ? d_actions + (size_t)t0 * per : nullptr;
    p.action_format = action_format; p.action_stride = d_actions ? per : 0;
    p.res_stride = (uint32_t)e->res_stride;
    p.actions_log = d_actions_log ? d_actions_log + (size_t)t0 * N * 3 : nullptr;
    p.reward_log = d_reward_log ? d_reward_log + (size_t)t0 * N : nullptr;
    p.done_log = d_done_log ? d_done_log + (size_t)t0 * N : nullptr;
    p.seed = seed; p.step_no = e->step_no; p.bank_count = e->bank_count;
    p.slot_step = e->bank_count ? (uint32_t)e->n_envs % e->bank_count : (uint32_t)e->n_envs;
    p.n_envs = e->n_envs; p.n_steps = nt; p.ring = e->ring; p.slice0 = e->slice;
    p.jobs = e->pin_jobs[buf]; p.job_total = e->pin_job_total[buf];
    p.job_cap = (uint32_t)std::min<size_t>((size_t)e->pin_job_steps * N, 0xffffffffu);
    int rc;
    if (e->list_steps) {
      const uint64_t step0 = e->step_no;
      const int slice0 = e->slice;
      rc = list_steps_chunk(e, seed, t0, nt, n_steps, buf, d_actions_log, d_reward_log, d_done_log, st,
                            d_actions, action_format, per, &cur);
      e->step_no = step0; e->slice = slice0;  // advanced below, as for rollout_pin_kernel
    } else {
      rc = pin_dispatch(e, false, st, p);
    }
    if (rc) return rc;
    CU(cudaGetLastError());
    RewardListParams r;
    memset(&r, 0, sizeof r);
    r.d = e->d; r.base = e->base; r.res_stride = (uint32_t)e->res_stride; r.n_envs = e->n_envs;
    r.jobs = e->pin_jobs[buf]; r.job_total = e->pin_job_total[buf]; r.job_cap = p.job_cap;
    r.reward_log = p.reward_log; r.step_no = e->step_no; r.stats = e->stats;
    r.final_from = std::max(0, n_steps - 2 - t0);  // beyond nt when the call goes on after this launch
    // (list_steps: the router and the reward kernel share the SMs with the step kernels of the
    // next chunk, like the per-step reward kernel does: a small persistent grid)
    const int want_ctas = (int)std::min<size_t>(((size_t)nt * N + kWarpsPerCta - 1) / kWarpsPerCta,
                                                (size_t)e->sms * (e->list_steps ? e->list_side_mult : PCB_REWARD_GRID_MULT));
    cudaStream_t rst = overlap ? e->side : st;
    if (overlap) {
      CU(cudaEventRecord(e->ev_step[buf], st));
      CU(cudaStreamWaitEvent(e->side, e->ev_step[buf], 0));
    }
    if (e->d.reward_type == PCB_REWARD_CENTROID) {
      if (e->reward_half) {
        const int groups = 32 / e->list_group;
        const int ctas = std::max(1, (want_ctas + groups - 1) / groups);
        if (e->list_group == 8)
          reward_list_half_kernel<8><<<ctas, kWarpsPerCta * 32, e->reward_smem * groups, rst>>>(r);
        else
          reward_list_half_kernel<16><<<ctas, kWarpsPerCta * 32, e->reward_smem * groups, rst>>>(r);
      } else
        reward_list_kernel<false><<<std::max(1, want_ctas), kWarpsPerCta * 32, e->reward_smem, rst>>>(r);
    } else {
      // beam_kernel over the list: it reads its job count at job_count[step_no & 3]
      Params bp;
      memset(&bp, 0, sizeof bp);
      bp.d = e->d; bp.jobs = e->pin_jobs[buf]; bp.job_count = e->pin_job_total[buf]; bp.step_no = 0;
      beam_kernel<<<std::max(1, want_ctas), 128, 0, rst>>>(bp);
      CU(cudaGetLastError());
      e->launches += 1;
      if (e->list_half_beam) {
        const int groups = 32 / e->list_group;
        const int ctas = std::max(1, (want_ctas + groups - 1) / groups);
        if (e->list_group == 8)
          reward_list_half_kernel<8, true><<<ctas, kWarpsPerCta * 32, e->reward_smem * groups, rst>>>(r);
        else
          reward_list_half_kernel<16, true><<<ctas, kWarpsPerCta * 32, e->reward_smem * groups, rst>>>(r);
      } else
        reward_list_kernel<true><<<std::max(1, want_ctas), kWarpsPerCta * 32, e->reward_smem, rst>>>(r);
    }
    CU(cudaGetLastError());
    if (overlap) {
      CU(cudaEventRecord(e->ev_rew[buf], e->side));
      e->rew_pending[buf] = true;
    }
    e->launches += e->list_steps ? 1 : 2;  // (launch() counted the step kernels)
    e->slice = (e->slice + nt) % e->ring;
    e->step_no += (uint64_t)nt;
  }
  int rc = join_rewards(e, st);
  if (rc) return rc;
  // step_kernel writes a step's reward / done either into the log rows or into the result halves
  if (e->list_steps && (d_reward_log || d_done_log)) {
    rc = restore_result_halves(e, n_steps, d_reward_log, d_done_log, st);
    if (rc) return rc;
  }
  if (e->timing) {
    CU(cudaEventRecord(b, st));
    CU(cudaEventSynchronize(b));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    e->fused_ms += ms; e->fused_launches += 1; e->fused_steps += n_steps;
  }
  return 0;
}

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

// pcb_rollout_random of the per-step kinds as CUDA-graph replays.  One graph = L consecutive
// steps with L a multiple of the observation ring, of the 4 job counters and of 2 (result
// halves, action ping-pong): after L steps every pointer a launch carries is the same
// again, so the captured graph can be replayed whenever the handle is in the phase it was
// captured in (same ring slice, same step number modulo 4).  The one thing that must keep
// moving -- the step counter of the random policy -- is read from device memory
// (Params::step_base = steps since capture), set before a run of replays and advanced by the
// last node of the graph.
// Every replay ends with the join of its last reward kernel (nothing overlaps it), so a
// graph is as long as the call allows, between ~16 and ~64 steps; a captured graph that
// still fits is kept (capturing costs milliseconds), so the length only ever shrinks.
#ifndef PCB_GRAPH_MAX_STEPS
#define PCB_GRAPH_MAX_STEPS 64
#endif
static int graph_length(const pcb_env_t* e, int n_steps) {
  const int base = e->ring * 4 / gcd_i(e->ring, 4);  // lcm(ring, 4)
  if (base > 128) return 0;
  const int lo = base * std::max(1, 16 / base), hi = base * std::max(1, PCB_GRAPH_MAX_STEPS / base);
  if (n_steps < lo) return 0;
  if (e->graph && e->graph_len >= lo && e->graph_len <= n_steps) return e->graph_len;
  return std::min(hi, n_steps / base * base);
}

static bool graph_allowed(pcb_env_t* e, cudaStream_t st) {
  if (!e->use_graph || e->timing || e->fused) return false;
  // the default streams cannot be captured
  if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return false;  // the caller is capturing this stream itself
  }
  return true;
}

// plain steps needed before the captured graph fits the handle's phase (0 = now), -1 = never
static int graph_align_distance(const pcb_env_t* e, int L) {
  for (int dlt = 0; dlt < L; ++dlt) {
    const bool slice_ok = (e->slice + dlt) % e->ring == e->graph_slice;
    const bool step_ok = ((e->step_no + (uint64_t)dlt - e->graph_step0) & 3) == 0 &&
                         e->step_no + (uint64_t)dlt >= e->graph_step0;
    if (slice_ok && step_ok) return dlt;
  }
  return -1;
}

// captures L steps starting at the handle's current phase (actions of the first step in d_actions)
static int graph_capture(pcb_env_t* e, uint64_t seed, int L, cudaStream_t st, bool overlap) {
  if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
  int rc = join_rewards(e, st);  // no event of an earlier, un-captured launch may be waited on
  if (rc) return rc;
  const uint64_t step0 = e->step_no;
  const int slice0 = e->slice;
  const int64_t launches0 = e->launches;
  CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
  int32_t* cur = e->d_actions;
  for (int t = 0; t < L && rc == 0; ++t) {
    int32_t* nxt = cur == e->d_actions ? e->d_actions2 : e->d_actions;
    // (the graph joins its last reward kernel before it ends: nothing runs beside that one)
    rc = step_impl(e, cur, PCB_ACTION_TUPLE, seed, nxt, st, overlap, nullptr, nullptr, e->d_step_base, t + 1 == L);
    cur = nxt;
  }
  if (rc == 0) rc = join_rewards(e, st);
  if (rc == 0) {
    advance_step_base_kernel<<<1, 1, 0, st>>>(e->d_step_base, (uint64_t)L);
    if (cudaGetLastError() != cudaSuccess) rc = fail(PCB_ERR_CUDA, "advance_step_base_kernel launch failed");
  }
  cudaGraph_t g = nullptr;
  cudaError_t ce = cudaStreamEndCapture(st, &g);
  // capture only records: put the host-side bookkeeping back to where the handle really is
  e->step_no = step0; e->slice = slice0;
  e->graph_launches = e->launches - launches0;
  e->launches = launches0;
  e->rew_pending[0] = e->rew_pending[1] = false;
  if (rc == 0 && (ce != cudaSuccess || !g)) rc = fail(PCB_ERR_CUDA, "graph capture: %s", cudaGetErrorString(ce));
  if (rc == 0) {
    ce = cudaGraphInstantiate(&e->graph, g, 0);
    if (ce != cudaSuccess) { e->graph = nullptr; rc = fail(PCB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce)); }
  }
  if (g) cudaGraphDestroy(g);
  if (rc) {
    cudaGetLastError();
    e->use_graph = false;  // this configuration cannot be captured: plain launches from now on
    return rc;
  }
  e->graph_len = L; e->graph_slice = slice0; e->graph_step0 = step0; e->graph_seed = seed;
  e->graph_stream = st; e->graph_overlap = overlap;
  return 0;
}

// replays while at least L steps remain; the handle must be in the captured phase
static int graph_replay(pcb_env_t* e, int* remaining, cudaStream_t st) {
  const int L = e->graph_len;
  if (*remaining < L) return 0;
  // the graph knows nothing of reward kernels launched outside it: a plain step just before
  // may still have one in flight on the side stream, reading the job buffer the graph's
  // second step writes
  int rc = join_rewards(e, st);
  if (rc) return rc;
  set_step_base_kernel<<<1, 1, 0, st>>>(e->d_step_base, e->step_no - e->graph_step0);
  CU(cudaGetLastError());
  while (*remaining >= L) {
    CU(cudaGraphLaunch(e->graph, st));
    e->step_no += (uint64_t)L;  // L is a multiple of the ring: the slice index is unchanged
    e->launches += e->graph_launches;
    e->graph_replays += 1;
    *remaining -= L;
  }
  return 0;
}

static int restore_result_halves(pcb_env_t* e, int n_steps, const double* d_reward_log,
                                 const uint8_t* d_done_log, cudaStream_t st);

int pcb_rollout_random_logged(pcb_env_t* e, uint64_t seed, int n_steps, int32_t* d_actions_log,
                              double* d_reward_log, uint8_t* d_done_log, void* stream) {
  int rc = ready(e);
  if (rc) return rc;
  if (n_steps < 0) return fail(PCB_ERR_INVALID, "n_steps must be >= 0");
  if (n_steps == 0) return 0;
  if (e->fused)
    return rollout_fused(e, seed, n_steps, d_actions_log, d_reward_log, d_done_log, (cudaStream_t)stream,
                         nullptr, 0);
  if (e->fused_pin) {
    rc = join_rewards(e, (cudaStream_t)stream);  // reward kernels of earlier per-step calls
    if (rc) return rc;
    return rollout_pin_fused(e, seed, n_steps, d_actions_log, d_reward_log, d_done_log, (cudaStream_t)stream,
                             nullptr, 0);
  }
  const size_t N = (size_t)e->n_envs, per = N * (e->d.kind == 0 ? 2 : 3);
  cudaStream_t st = (cudaStream_t)stream;
  const int total_steps = n_steps;
  // CUDA-graph plan: `dlt` plain steps walk the handle into the phase the graph was captured
  // in (or a graph is captured right here, dlt = 0), then replays, then the remainder
  const int L = graph_length(e, n_steps);
  const bool overlap = e->overlap && e->d.kind >= 2;
  bool graphed = !d_actions_log && !d_reward_log && !d_done_log && L > 0 && n_steps >= L && graph_allowed(e, st);
  bool capture = false;
  int dlt = 0;
  if (graphed) {
    const bool usable = e->graph && e->graph_len == L && e->graph_seed == seed &&
                        e->graph_stream == st && e->graph_overlap == overlap;
    dlt = usable ? graph_align_distance(e, L) : -1;
    if (dlt < 0) { capture = true; dlt = 0; }
    if (n_steps - dlt < L) graphed = false;
  }
  // the first action comes from the stand-alone sampler; every step kernel then also
  // draws the action of the following step (same counter as the sampler would use).  The two
  // action buffers ping-pong; a replay expects its first actions in d_actions.
  int32_t* cur = d_actions_log ? d_actions_log : ((graphed && (dlt & 1)) ? e->d_actions2 : e->d_actions);
  if (e->pend_actions && e->pend_seed == seed && e->pend_step == e->step_no && e->pend_stream == st) {
    // the previous call's last step already drew them
    if (e->pend_actions != cur)
      CU(cudaMemcpyAsync(cur, e->pend_actions, per * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  } else {
    rc = pcb_sample_random_actions(e, seed, cur, stream);
    if (rc) return rc;
  }
  e->pend_actions = nullptr;
  // plain per-step launches; `cur` holds the actions of the next step throughout
  auto plain = [&](int count, bool last_of_call) -> int {
    for (int t = 0; t < count; ++t) {
      const int done = total_steps - n_steps;
      int32_t* nxt = nullptr;
      if (!(last_of_call && t + 1 == count))
        nxt = d_actions_log ? d_actions_log + (size_t)(done + 1) * per
                            : (cur == e->d_actions ? e->d_actions2 : e->d_actions);
      else  // the actions of the step after this call: kept for the next call
        nxt = cur == e->d_actions ? e->d_actions2 : e->d_actions;
      int rc2 = step_impl(e, cur, PCB_ACTION_TUPLE, seed, nxt, stream, e->overlap,
                          d_reward_log ? d_reward_log + (size_t)done * N : nullptr,
                          d_done_log ? d_done_log + (size_t)done * N : nullptr, nullptr,
                          last_of_call && t + 1 == count);
      if (rc2) return rc2;
      cur = nxt;
      n_steps -= 1;
    }
    return 0;
  };
  if (graphed) {
    if (capture && graph_capture(e, seed, L, st, overlap) != 0) graphed = false;  // plain launches then
  }
  if (graphed) {
    rc = plain(dlt, false);
    if (rc) return rc;
    if (cur != e->d_actions) return fail(PCB_ERR_STATE, "internal: action ping-pong out of phase");
    rc = graph_replay(e, &n_steps, st);
    if (rc) return rc;
  }
  rc = plain(n_steps, true);
  if (rc) return rc;
  rc = join_rewards(e, (cudaStream_t)stream);
  if (rc) return rc;
  rc = restore_result_halves(e, total_steps, d_reward_log, d_done_log, (cudaStream_t)stream);
  // (a call that ends on a replay leaves them in d_actions, like the last plain step)
  e->pend_actions = cur; e->pend_seed = seed; e->pend_step = e->step_no; e->pend_stream = st;
  return rc;
}

// the result tensors ([2] halves by step parity) describe the last two steps: copy them back
// from the [T, N] logs the steps wrote into
static int restore_result_halves(pcb_env_t* e, int n_steps, const double* d_reward_log,
                                 const uint8_t* d_done_log, cudaStream_t st) {
  const size_t N = (size_t)e->n_envs;
  for (int back = 0; back < 2 && back < n_steps; ++back) {
    const size_t t = (size_t)(n_steps - 1 - back);
    const size_t half = (size_t)((e->step_no - 1 - back) & 1) * e->res_stride;
    if (d_reward_log)
      CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(e->base.reward) + half, d_reward_log + t * N,
                         N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (d_done_log)
      CU(cudaMemcpyAsync(e->base.done + half, d_done_log + t * N, N, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

constexpr int kHostChunk = 8;  // pcb_steps_host uploads the action rows in chunks of 8 steps (after the first row)

static int ensure_host_staging(pcb_env_t* e, size_t steps) {
  if (steps <= e->hs_steps) return 0;
  CU(cudaDeviceSynchronize());
  if (e->hs_actions) cudaFree(e->hs_actions);
  if (e->hs_reward) cudaFree(e->hs_reward);
  if (e->hs_done) cudaFree(e->hs_done);
  e->hs_actions = nullptr; e->hs_reward = nullptr; e->hs_done = nullptr; e->hs_steps = 0;
  const size_t N = (size_t)e->n_envs;
  void* p = nullptr;
  cudaError_t err = cudaMalloc(&p, steps * N * 3 * sizeof(int32_t));
  if (err == cudaSuccess) { e->hs_actions = (int32_t*)p; err = cudaMalloc(&p, steps * N * sizeof(double)); }
  if (err == cudaSuccess) { e->hs_reward = (double*)p; err = cudaMalloc(&p, steps * N); }
  if (err == cudaSuccess) e->hs_done = (uint8_t*)p;
  if (err != cudaSuccess)
    return fail(err == cudaErrorMemoryAllocation ? PCB_ERR_NOMEM : PCB_ERR_CUDA,
                "pcb_steps_host staging for %zu steps: %s", steps, cudaGetErrorString(err));
  e->hs_steps = steps;
  if (!e->copy) CU(cudaStreamCreateWithFlags(&e->copy, cudaStreamNonBlocking));
  if (!e->ev_host) CU(cudaEventCreateWithFlags(&e->ev_host, cudaEventDisableTiming));
  return 0;
}

int pcb_steps_host(pcb_env_t* e, const int32_t* h_actions, int n_steps, int action_format,
                   double* h_reward, uint8_t* h_done, void* stream) {
  int rc = ready(e);
  if (rc) return rc;
  if (!h_actions || !h_reward || !h_done) return fail(PCB_ERR_INVALID, "NULL host buffer");
  if (n_steps < 0) return fail(PCB_ERR_INVALID, "n_steps must be >= 0");
  if (action_format != PCB_ACTION_TUPLE && action_format != PCB_ACTION_FLAT)
    return fail(PCB_ERR_INVALID, "unknown action format %d", action_format);
  if (n_steps == 0) return 0;
  CU(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t N = (size_t)e->n_envs, T = (size_t)n_steps;
  const size_t per = N * (action_format == PCB_ACTION_FLAT ? 1 : (e->d.kind == 0 ? 2 : 3));
  rc = ensure_host_staging(e, T);
  if (rc) return rc;
  if (e->fused || e->fused_pin) {
    // pin-less kinds (and the pin env in its multi-step kernel): every row up, one launch of
    // the multi-step kernel per chunk of steps, every result down
    CU(cudaMemcpyAsync(e->hs_actions, h_actions, T * per * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (e->fused_pin) rc = join_rewards(e, st);
    if (rc) return rc;
    rc = e->fused ? rollout_fused(e, 0, n_steps, nullptr, e->hs_reward, e->hs_done, st, e->hs_actions, action_format)
                  : rollout_pin_fused(e, 0, n_steps, nullptr, e->hs_reward, e->hs_done, st, e->hs_actions, action_format);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_reward, e->hs_reward, T * N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_done, e->hs_done, T * N, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
  }
  // Pin kinds: one step + one reward kernel per row.  The action rows travel on a copy
  // stream in chunks (the first step starts after the first chunk, not after all rows); the
  // results of step t leave on the stream its reward kernel ran on, while step t+1 runs.
  // chunk 0 is the first row alone, so that the first step waits for one row, not for eight
  auto chunk_begin = [](size_t c) { return c == 0 ? (size_t)0 : 1 + (c - 1) * (size_t)kHostChunk; };
  const size_t chunks = 1 + (T - 1 + kHostChunk - 1) / kHostChunk;
  while (e->hs_events.size() < chunks) {
    cudaEvent_t ev = nullptr;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->hs_events.push_back(ev);
  }
  CU(cudaEventRecord(e->ev_host, st));           // the copy stream starts after the caller's work
  CU(cudaStreamWaitEvent(e->copy, e->ev_host, 0));
  for (size_t c = 0; c < chunks; ++c) {
    const size_t t0 = chunk_begin(c), nt = std::min(chunk_begin(c + 1), T) - t0;
    CU(cudaMemcpyAsync(e->hs_actions + t0 * per, h_actions + t0 * per, nt * per * sizeof(int32_t),
                       cudaMemcpyHostToDevice, e->copy));
    CU(cudaEventRecord(e->hs_events[c], e->copy));
  }
  const bool routes = e->d.kind >= 2;
  const bool overlap = routes && e->overlap;
  size_t next_chunk = 0;
  for (size_t t = 0; t < T; ++t) {
    if (t == chunk_begin(next_chunk)) CU(cudaStreamWaitEvent(st, e->hs_events[next_chunk++], 0));
    double* rrow = e->hs_reward + t * N;
    uint8_t* drow = e->hs_done + t * N;
    rc = step_impl(e, e->hs_actions + t * per, action_format, 0, nullptr, stream, overlap, rrow, drow,
                   nullptr, t + 1 == T);
    if (rc) return rc;
    cudaStream_t out = overlap ? e->side : st;   // the stream on which row t becomes final
    CU(cudaMemcpyAsync(h_reward + t * N, rrow, N * sizeof(double), cudaMemcpyDeviceToHost, out));
    CU(cudaMemcpyAsync(h_done + t * N, drow, N, cudaMemcpyDeviceToHost, out));
  }
  if (overlap) {
    rc = join_rewards(e, st);                    // the reward kernels ...
    if (rc) return rc;
  }
  // (on `stream`, beside the last rows' copies to the host on the side stream)
  rc = restore_result_halves(e, n_steps, e->hs_reward, e->hs_done, st);
  if (rc) return rc;
  if (overlap) {
    CU(cudaEventRecord(e->ev_host, e->side));    // ... and the copies queued behind them
    CU(cudaStreamWaitEvent(st, e->ev_host, 0));
  }
  CU(cudaStreamSynchronize(st));
  return 0;
}

int pcb_rollout_random(pcb_env_t* e, uint64_t seed, int n_steps, int32_t* d_actions_log, void* stream) {
  return pcb_rollout_random_logged(e, seed, n_steps, d_actions_log, nullptr, nullptr, stream);
}

int pcb_set_graph_rollout(pcb_env_t* e, int on) {
  if (!e) return fail(PCB_ERR_INVALID, "env is NULL");
  e->use_graph = on != 0;
  if (!on && e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
  return 0;
}

int64_t pcb_graph_replays(pcb_env_t* e) { return e ? e->graph_replays : 0; }

int pcb_set_reward_overlap(pcb_env_t* e, int on) {
  if (!e) return fail(PCB_ERR_INVALID, "env is NULL");
  e->overlap = on != 0;
  return 0;
}

int64_t pcb_check_guards(pcb_env_t* e, void* stream) {
  if (!e) return fail(PCB_ERR_INVALID, "env is NULL");
  if (!e->guard) return fail(PCB_ERR_STATE, "the handle was not created with PCB_GUARD=1");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  CU(cudaDeviceSynchronize());
  std::vector<uint8_t> host(kGuardBytes);
  int64_t bad = 0;
  for (const pcb_env::Guarded& g : e->guarded)
    for (int side = 0; side < 2; ++side) {
      const uint8_t* src = side ? g.base + kGuardBytes + g.bytes : g.base;
      CU(cudaMemcpy(host.data(), src, kGuardBytes, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < kGuardBytes; ++i) bad += host[i] != (uint8_t)kGuardByte;
    }
  return bad;
}

int pcb_set_fused_rollout(pcb_env_t* e, int on) {
  if (!e) return fail(PCB_ERR_INVALID, "env is NULL");
  e->fused = on != 0 && e->d.kind <= 1 && !e->wide;
  e->fused_pin = on != 0 && e->fused_pin_ok;  // on: also for batches above the default threshold
  return (e->fused || e->fused_pin) ? 1 : 0;
}

int pcb_fused_rollout_times(pcb_env_t* e, double* ms, int64_t* launches, int64_t* steps) {
  if (!e || !ms || !launches || !steps) return fail(PCB_ERR_INVALID, "NULL argument");
  *ms = e->fused_ms; *launches = e->fused_launches; *steps = e->fused_steps;
  e->fused_ms = 0.0; e->fused_launches = 0; e->fused_steps = 0;
  return 0;
}

int pcb_enable_kernel_timing(pcb_env_t* e, int on) {
  if (!e) return fail(PCB_ERR_INVALID, "env is NULL");
  e->timing = on != 0;
  return 0;
}

int pcb_kernel_times(pcb_env_t* e, double ms[4], int64_t launches[4]) {
  if (!e || !ms || !launches) return fail(PCB_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(e->device));
  int rc = fold_timed(e);
  if (rc) return rc;
  for (int i = 0; i < 4; ++i) {
    ms[i] = e->timed_ms[i]; launches[i] = e->timed_n[i];
    e->timed_ms[i] = 0.0; e->timed_n[i] = 0;
  }
  e->timed.clear();
  return 0;
}

}  // extern "C"
