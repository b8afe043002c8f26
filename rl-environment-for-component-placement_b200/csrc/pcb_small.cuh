// pcb_small.cuh -- the whole random-policy rollout of the pin-less environments
// (square: dummy_env_square.py:74-153, rectangle: dummy_env_rectangular.py:310-432) in
// ONE launch: n_steps x { random policy (agent/random/random_policy_*.py:11-58) -> step ->
// episode statistics -> auto-reset -> observation write-out }; with the caller's actions
// instead of the policy the same kernel is pcb_step for these kinds (one step per launch).
//
// Why a second kernel: these boards have at most 32 rows and no routing reward, so a step
// is ~1 KB of output and a few hundred instructions.  One launch per step is launch-bound
// at small batches (4 096 envs: 6.5 us per step against 0.7 us of HBM time) and a full
// warp per env leaves half (16x16) or three quarters (8x8) of the lanes idle.  Here
//   * LPE lanes own one environment (LPE = 8, 16 or 32 >= H, lane = grid row), so a warp
//     steps 32/LPE environments; the groups of a warp run converged and differ by predicates,
//   * the state lives in registers (board row, counters) and shared memory (components)
//     for all n_steps; HBM sees it once at the start and once at the end,
//   * the next instance of an env is prefetched from the bank at the start of an episode,
//     so an auto-reset never waits for a load.
// The arithmetic is the step kernel's (pcb_kernels.cuh) and the counters of the random
// policy are the same (seed, env, step), so the result equals n_steps calls of
// { pcb_sample_random_actions; pcb_step } bit for bit (tests/test_gpu_api.py).
#ifndef PCB_SMALL_CUH_
#define PCB_SMALL_CUH_

#include "pcb_kernels.cuh"

namespace pcb {

struct RolloutParams {
  Dims d;
  Out base;              // slice 0 / half 0 of every tensor
  uint8_t* state;
  const uint8_t* bank;
  const float* ratio_lut;
  double* stats;
  const int32_t* actions;  // nullptr: the device random policy; else the caller's actions,
  int action_format;       // row t at actions + t * action_stride: [N, 2|3] tuples or [N] flat
  size_t action_stride;    // (pcb_step: one row; pcb_steps_host: n_steps rows)
  uint32_t res_stride;     // bytes between the two halves of the reward / done block
  int32_t* actions_log;  // nullptr or [n_steps, N, 2|3]
  double* reward_log;    // nullptr or [n_steps, N]
  uint8_t* done_log;     // nullptr or [n_steps, N]
  uint64_t seed;
  uint64_t step_no;      // number of the first step
  uint32_t bank_count;
  uint32_t slot_step;    // n_envs % bank_count: the next slot is a conditional subtraction away
  int n_envs, n_steps;
  int ring, slice0;      // step t writes observation slice (slice0 + 1 + t) % ring
};

// shared memory per environment: CompRec[C] + dense bit string of one board (generic widths)
PCB_HD constexpr int small_env_smem(int C, int HW) {
  return align16c(C * 4) + align16c(4 * ((HW + 31) / 32 + 2));
}

PCB_HD constexpr int small_lanes_per_env(int H) { return H <= 8 ? 8 : (H <= 16 ? 16 : 32); }

// Every shuffle / vote below is issued by the WHOLE warp with the full mask (a group mask
// that is not a compile-time constant costs a MATCH + divergence check per operation:
// measured, half of the kernel's instructions), so the control flow around them is kept
// warp-uniform and the groups only differ by predicates.  Shuffles are confined to the
// group by their width, votes by masking the ballot.
template <int LPE>
__device__ __forceinline__ bool group_any(bool pred, unsigned gmask) {
  return (__ballot_sync(kFull, pred) & gmask) != 0u;
}

template <int LPE>
__device__ __forceinline__ uint32_t rows_ahead_g(uint32_t v, int dlt, int sl) {
  const uint32_t t = __shfl_down_sync(kFull, v, dlt, LPE);
  return sl + dlt < LPE ? t : 0u;
}

// footprint_mask (pcb_kernels.cuh) for one row per lane inside an LPE-lane group.
// small = every footprint of the configuration is at most 4x4 (warp-uniform).
template <int LPE>
__device__ __forceinline__ uint32_t footprint_mask_g(uint32_t occ, int ph, int pw, int H, int W,
                                                     int sl, bool small, int pmax) {
  uint32_t v = sl < H ? (~occ & low_mask<uint32_t>(W)) : 0u;
  if (small) {
    const uint32_t a = v & (v >> 1);
    uint32_t r = pw >= 2 ? a : v;
    if (pw == 3) r = a & (v >> 2);
    if (pw == 4) r = a & (a >> 2);
    v = r;
    const uint32_t t1 = rows_ahead_g<LPE>(v, 1, sl);
    uint32_t b = ph >= 2 ? (v & t1) : v;
    const uint32_t t2 = rows_ahead_g<LPE>(ph == 4 ? b : v, 2, sl);
    if (ph >= 3) b &= t2;
    return b;
  }
  // any size: the trip counts follow the configuration, the footprint only predicates
  uint32_t acc = v;
  for (int k = 1; k < pmax; ++k)
    if (k < pw) acc &= v >> k;
  v = acc;
  for (int k = 1; k < pmax; ++k) {
    const uint32_t t = rows_ahead_g<LPE>(v, k, sl);
    if (k < ph) acc &= t;
  }
  return acc;
}

// sample_action (pcb_kernels.cuh) for every group of the warp at once: same counter, same
// draw.  The pin-less kinds use the exact method only (k-th set bit in np.argwhere order,
// k uniform in [0, total)): no data-dependent loop, so the groups stay converged.  The row
// counts of both orientations travel through one prefix sum as two 16-bit fields.
template <int LPE, typename S>
__device__ __forceinline__ void sample_action_g(const S& sh, uint32_t m0, uint32_t m1, uint64_t seed,
                                                int env, uint64_t step_no, int lane,
                                                int& ao, int& ax, int& ay) {
  const int sl = lane & (LPE - 1);
  const uint32_t c01 = (uint32_t)__popc(m0) | ((uint32_t)__popc(m1) << 16);
  uint32_t incl = c01;
#pragma unroll
  for (int off = 1; off < LPE; off <<= 1) {
    const uint32_t v = __shfl_up_sync(kFull, incl, off, LPE);
    if (sl >= off) incl += v;
  }
  const uint32_t tot = __shfl_sync(kFull, incl, LPE - 1, LPE);
  const int tot0 = (int)(tot & 0xffffu), tot1 = (int)(tot >> 16);
  const int total = sh.O() == 2 ? tot0 + tot1 : tot0;
  const uint32_t r = action_rand(seed, (uint32_t)env, step_no);
  int kth = (int)__umulhi(r, (uint32_t)total);
  const bool second = kth >= tot0;
  if (second) kth -= tot0;
  const uint32_t bits = second ? m1 : m0;
  const int inc = (int)(second ? incl >> 16 : incl & 0xffffu);
  const int excl = inc - __popc(bits);
  const unsigned who = __ballot_sync(kFull, kth >= excl && kth < inc) >> (lane & ~(LPE - 1));
  const int src = (__ffs(who) - 1) & (LPE - 1);  // row that holds the kth bit
  const int col = __shfl_sync(kFull, nth_bit32(bits, kth - excl), src, LPE);
  const bool have = total > 0;  // an empty mask reports (0, 0, 0)
  ao = have && second ? 1 : 0;
  ax = have ? src : 0;
  ay = have ? col : 0;
}

// one board (bit row per lane) -> H*W bytes of 0/1; dup_off != 0 writes a second copy at
// dst + dup_off (orientations 2, 3 of the pin env repeat 0, 1)
template <int LPE, typename S>
__device__ __forceinline__ void write_board_g(const S& sh, uint8_t* dst, uint32_t bits, uint32_t* flat,
                                              int sl, bool live, const uint2* lut, int dup_off = 0) {
  const int H = sh.H(), W = sh.W();
  if (W == 16) {
    if (live && sl < H) {
      const uint4 v = expand16(lut, bits);
      st_stream(reinterpret_cast<uint4*>(dst) + sl, v);
      if (dup_off) st_stream(reinterpret_cast<uint4*>(dst + dup_off) + sl, v);
    }
  } else if (W == 8) {
    if (live && sl < H) {
      const uint2 v = lut[bits & 0xffu];
      st_stream(reinterpret_cast<uint2*>(dst) + sl, v);
      if (dup_off) st_stream(reinterpret_cast<uint2*>(dst + dup_off) + sl, v);
    }
  } else if (W == 32) {
    if (live && sl < H) {
      const uint4 a = expand16(lut, bits), b = expand16(lut, bits >> 16);
      st_stream(reinterpret_cast<uint4*>(dst) + 2 * sl, a);
      st_stream(reinterpret_cast<uint4*>(dst) + 2 * sl + 1, b);
      if (dup_off) {
        st_stream(reinterpret_cast<uint4*>(dst + dup_off) + 2 * sl, a);
        st_stream(reinterpret_cast<uint4*>(dst + dup_off) + 2 * sl + 1, b);
      }
    }
  } else {
    // dense bit string (bit row*W + col) in shared memory, then 16 / 4 / 1 bits per lane
    const int nbits = H * W, nwords = (nbits + 31) >> 5;
    for (int i = sl; i < nwords + 1; i += LPE) flat[i] = 0u;
    __syncwarp();
    if (sl < H && bits) {
      const int pos = sl * W, lo = pos >> 5;
      const uint64_t a = (uint64_t)bits << (pos & 31);
      atomicOr(flat + lo, (uint32_t)a);
      if (a >> 32) atomicOr(flat + lo + 1, (uint32_t)(a >> 32));
    }
    __syncwarp();
    if (live) {
      if ((nbits & 15) == 0) {
        for (int q = sl; q < (nbits >> 4); q += LPE) {
          const uint4 v = expand16(lut, flat[q >> 1] >> (16 * (q & 1)));
          st_stream(reinterpret_cast<uint4*>(dst) + q, v);
          if (dup_off) st_stream(reinterpret_cast<uint4*>(dst + dup_off) + q, v);
        }
      } else if ((nbits & 3) == 0) {
        for (int q = sl; q < (nbits >> 2); q += LPE) {
          const uint32_t v = spread4((flat[q >> 3] >> (4 * (q & 7))) & 15u);
          st_stream(reinterpret_cast<uint32_t*>(dst) + q, v);
          if (dup_off) st_stream(reinterpret_cast<uint32_t*>(dst + dup_off) + q, v);
        }
      } else {
        for (int b = sl; b < nbits; b += LPE) {
          const uint8_t v = (uint8_t)((flat[b >> 5] >> (b & 31)) & 1u);
          st_stream(dst + b, v);
          if (dup_off) st_stream(dst + dup_off + b, v);
        }
      }
    }
    __syncwarp();
  }
}

template <int LPE, typename S>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
rollout_small_kernel(const __grid_constant__ RolloutParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const S sh(d);
  constexpr int G = 32 / LPE;  // environments per warp
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int sl = lane & (LPE - 1), eic = wic * G + lane / LPE;
  unsigned gmask = kFull;
  if constexpr (LPE < 32) gmask = ((1u << LPE) - 1u) << (lane & ~(LPE - 1));
  const int env_raw = blockIdx.x * ((int)(blockDim.x >> 5) * G) + eic;
  build_expand_lut(smem);
  // a group past the last env keeps running on env 0's data with every store switched off,
  // so that the warp stays converged
  const bool live = env_raw < p.n_envs;
  if (!__any_sync(kFull, live)) return;
  const int env = live ? env_raw : 0;
  const uint2* lut = reinterpret_cast<const uint2*>(smem);
  const int kind = sh.kind(), H = sh.H(), W = sh.W(), C = sh.C(), HW = H * W, O = sh.O();
  uint32_t* comps = reinterpret_cast<uint32_t*>(smem + kLutBytes + (size_t)eic * small_env_smem(C, HW));
  uint32_t* flat = comps + (align16c(C * 4) >> 2);
  const uint32_t N = (uint32_t)p.n_envs;
  const int aw = kind == 0 ? 2 : 3;
  const int pmax = kind == 0 ? d.sq_n : (d.mh > d.mw ? d.mh : d.mw);
  const bool small = pmax <= 4;

  // ---- state: HBM -> registers / shared memory (once) ----
  uint8_t* gs = p.state + (size_t)env * sh.st_stride();
  int n_comp, cur, flags;
  uint32_t hdr0, slot, ep_len;  // hdr0 = n_comp | n_nets << 8 | n_pins << 16 as stored
  double ep_ret;
  uint64_t epno;
  {
    const uint4 h4 = *reinterpret_cast<const uint4*>(gs);
    const uint4 d4 = *reinterpret_cast<const uint4*>(gs + 16);
    hdr0 = h4.x; n_comp = h4.x & 0xff; cur = h4.y & 0xff; flags = (h4.y >> 8) & 0xff;
    slot = h4.z; ep_len = h4.w;
    // the bank may have been replaced by a smaller one since this episode started
    // (pcb_upload_instances / pcb_generate_instances*): the +slot_step / subtract-once
    // update below needs slot < bank_count
    if (p.bank_count && slot >= p.bank_count) slot %= p.bank_count;
    ep_ret = __longlong_as_double((long long)(((uint64_t)d4.y << 32) | d4.x));
    epno = ((uint64_t)d4.w << 32) | d4.z;
  }
  for (int c = sl; c < C; c += LPE) comps[c] = reinterpret_cast<const uint32_t*>(gs + sh.st_comps())[c];
  uint32_t occ = reinterpret_cast<const uint32_t*>(gs + sh.st_occ())[sl];
  __syncwarp();

  // next instance of this env, fetched long before the episode ends
  constexpr int kPre = 2;  // components per lane kept in registers: C <= kPre * LPE, else re-read
  uint32_t nxt_hdr = 1u, nxt_comp[kPre] = {0u, 0u};
  auto prefetch = [&](uint32_t cur_slot) {
    uint32_t s2 = cur_slot + p.slot_step;  // (cur_slot + N) % bank_count without a division
    if (p.bank_count && s2 >= p.bank_count) s2 -= p.bank_count;
    const uint8_t* rec = p.bank + (size_t)s2 * sh.rec_stride();
    nxt_hdr = *reinterpret_cast<const uint32_t*>(rec);
#pragma unroll
    for (int k = 0; k < kPre; ++k)
      if (sl + k * LPE < C) nxt_comp[k] = reinterpret_cast<const uint32_t*>(rec + sh.rec_comps())[sl + k * LPE];
  };
  if (kind != 0) prefetch(slot);

  uint32_t m0, m1;
  auto masks = [&]() {  // warp-uniform: called by all lanes; returns "the group has a valid action"
    if (kind == 0) {
      m0 = footprint_mask_g<LPE>(occ, d.sq_n, d.sq_n, H, W, sl, small, pmax);
      m1 = 0u;
    } else {
      const bool have = cur < n_comp;
      const uint32_t r = comps[have ? cur : 0];
      const int ch = r & 0xff, cw = (r >> 8) & 0xff;
      const uint32_t a = footprint_mask_g<LPE>(occ, ch, cw, H, W, sl, small, pmax);
      const uint32_t b = footprint_mask_g<LPE>(occ, cw, ch, H, W, sl, small, pmax);
      m0 = have ? a : 0u;
      m1 = have ? b : 0u;
    }
    return group_any<LPE>((m0 | m1) != 0u, gmask);
  };
  bool any = masks();
  // square env: the mask after a reset is a constant
  uint32_t m_empty = 0u;
  bool any_empty = false;
  if (kind == 0) {
    m_empty = footprint_mask_g<LPE>(0u, d.sq_n, d.sq_n, H, W, sl, small, pmax);
    any_empty = group_any<LPE>(m_empty != 0u, gmask);
  }

  float frow[kPre][5];
  auto feature_row = [&](uint32_t r, float (&f)[5]) {
    const int h = r & 0xff, w = (r >> 8) & 0xff, cx = (r >> 16) & 0xff, cy = r >> 24;
    f[0] = (float)h; f[1] = (float)w;
    f[2] = cx == 0xFF ? -1.f : (float)cx;
    f[3] = cy == 0xFF ? -1.f : (float)cy;
    if (S::fixed && (HW & (HW - 1)) == 0) f[4] = (float)(h * w) * (1.0f / (float)HW);
    else f[4] = p.ratio_lut[h * w];
  };
  bool feat_dirty = true;  // group-uniform

  uint32_t slice = (uint32_t)p.slice0;
  for (int t = 0; t < p.n_steps; ++t) {
    const uint64_t step_no = p.step_no + (uint64_t)t;
    const uint32_t he = ((uint32_t)step_no & 1u) ? N + (uint32_t)env : (uint32_t)env;
    const size_t hoff = ((uint32_t)step_no & 1u) ? (size_t)p.res_stride : 0;  // reward | done block

    int o = 0, x = 0, y = 0, ph = d.sq_n, pw = d.sq_n;
    bool valid;
    if (p.actions) {
      const int32_t* acts = p.actions + (size_t)t * p.action_stride;
      // ---- given actions (pcb_step; FlatteningActionWrapper*: env_wrappers.py:80-98,
      // wrapper_square.py:54-69).  validate_action == mask bit, recomputed (square:155-173,
      // rect:434-457); actions outside the box are invalid ----
      if (p.action_format == 1) {
        const int a = acts[env];
        if (a < 0) { o = -1; }
        else { o = a / HW; const int rem = a - o * HW; x = rem / W; y = rem - x * W; }
      } else if (kind == 0) {
        x = acts[2 * env]; y = acts[2 * env + 1];
      } else {
        o = acts[3 * env]; x = acts[3 * env + 1]; y = acts[3 * env + 2];
      }
      bool in_range = (unsigned)o < (unsigned)O && (unsigned)x < (unsigned)H && (unsigned)y < (unsigned)W;
      if (kind != 0) {
        in_range = in_range && cur < n_comp;
        const uint32_t r = comps[in_range ? cur : 0];
        const int ch = r & 0xff, cw = (r >> 8) & 0xff;
        ph = (o & 1) ? cw : ch;
        pw = (o & 1) ? ch : cw;
      }
      const bool fits = in_range && x + ph <= H && y + pw <= W;
      const uint32_t cols = fits ? (low_mask<uint32_t>(pw) << y) : 0u;
      const bool clash = group_any<LPE>(sl >= x && sl < x + ph && (occ & cols), gmask);
      valid = fits && !clash;
    } else {
      // ---- random policy; its pick is valid exactly when the mask has a set bit, and with
      // an empty mask it reports (0,0,0), whose mask bit is 0: validate_action == `any` ----
      sample_action_g<LPE, S>(sh, m0, m1, p.seed, env, step_no, lane, o, x, y);
      valid = any;
      if (kind != 0 && valid) {
        const uint32_t r = comps[cur];
        const int ch = r & 0xff, cw = (r >> 8) & 0xff;
        ph = (o & 1) ? cw : ch;
        pw = (o & 1) ? ch : cw;
      }
    }
    if (p.actions_log && sl == 0 && live) {
      int32_t* a = p.actions_log + ((size_t)t * N + (size_t)env) * aw;
      if (kind == 0) { a[0] = x; a[1] = y; }
      else { a[0] = o; a[1] = x; a[2] = y; }
    }
    // update_grid (square:175-185, rect:459-484)
    if (valid && sl >= x && sl < x + ph) occ |= low_mask<uint32_t>(pw) << y;
    __syncwarp();  // every lane has read comps[cur]
    if (valid && kind != 0) {
      if (sl == 0) comps[cur] = (comps[cur] & 0xffffu) | ((uint32_t)x << 16) | ((uint32_t)y << 24);
#pragma unroll
      for (int k = 0; k < kPre; ++k)
        if (sl + k * LPE == cur) { frow[k][2] = (float)x; frow[k][3] = (float)y; }
      cur += 1;
    }
    ep_len += 1;
    const double reward = valid ? 1.0 : 0.0;  // square:115-153, rect:405-432
    ep_ret += reward;

    // ---- done / auto-reset (no shuffle or vote inside report / restart) ----
    auto report = [&]() {  // end of an episode: info + statistics (first report only)
      if (sl == 0 && live) {
        p.base.ep_return[he] = ep_ret;
        p.base.ep_length[he] = (int32_t)ep_len;
        if (!(flags & kFlagFinished)) {
          double* st = p.stats + (size_t)(env & (kStatSlots - 1)) * 8;
          atomicAdd(st + 0, 1.0);
          atomicAdd(st + 1, ep_ret);
          atomicAdd(st + 2, (double)ep_len);
        }
      }
      if (!d.autoreset) flags |= kFlagFinished;
    };
    auto restart = [&]() {  // reset(): next bank slot, empty board (rect:310-351)
      slot += p.slot_step;
      if (p.bank_count && slot >= p.bank_count) slot -= p.bank_count;
      epno += 1;
      if (kind != 0) {
        hdr0 = nxt_hdr; n_comp = nxt_hdr & 0xff;
#pragma unroll
        for (int k = 0; k < kPre; ++k)
          if (sl + k * LPE < C) comps[sl + k * LPE] = nxt_comp[k];
        if (C > kPre * LPE) {
          const uint8_t* rec = p.bank + (size_t)slot * sh.rec_stride();
          for (int c = sl + kPre * LPE; c < C; c += LPE)
            comps[c] = reinterpret_cast<const uint32_t*>(rec + sh.rec_comps())[c];
        }
        prefetch(slot);
      }
      cur = 0; flags = 0; ep_len = 0; ep_ret = 0.0; occ = 0u;
      feat_dirty = true;
    };
    const bool all_placed = kind != 0 && cur == n_comp;
    const bool early = !valid || all_placed;  // done without looking at the next mask
    if (early) {
      report();
      if (d.autoreset) restart();
    }
    __syncwarp();
    any = masks();
    const bool blocked = !early && !any;  // compute_if_done: nothing fits (rect:586-599)
    if (__any_sync(kFull, blocked)) {     // warp-uniform; the other groups recompute the same masks
      if (blocked) {
        report();
        if (d.autoreset) restart();
      }
      __syncwarp();
      if (kind != 0) any = masks();
      else if (blocked && d.autoreset) { m0 = m_empty; any = any_empty; }
    }
    const bool done = early || blocked;
    if (sl == 0 && live) {
      p.base.done[hoff + (size_t)env] = done ? 1 : 0;
      *reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(p.base.reward + env) + hoff) = reward;
      const size_t le = (size_t)t * N + (size_t)env;
      if (p.reward_log) p.reward_log[le] = reward;
      if (p.done_log) p.done_log[le] = done ? 1 : 0;
    }

    // ---- observation of slice (slice0 + 1 + t) % ring ----
    slice = slice + 1u == (uint32_t)p.ring ? 0u : slice + 1u;
    const uint32_t se = slice * N + (uint32_t)env;
    write_board_g<LPE, S>(sh, p.base.grid + (size_t)se * HW, occ, flat, sl, live, lut);
    uint8_t* mk = p.base.mask + (size_t)se * (uint32_t)(O * HW);
    write_board_g<LPE, S>(sh, mk, m0, flat, sl, live, lut);
    if (O >= 2) write_board_g<LPE, S>(sh, mk + HW, m1, flat, sl, live, lut);
    if (kind != 0) {
      // all_components_feature rows [h, w, x|-1, y|-1, h*w/(H*W)] (rect:53-71) live in registers
      // (lane sl owns components sl, sl + LPE) and are rebuilt only when they change: a reset
      // rewrites every row, a placement the (x, y) of one
      if (feat_dirty) {
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
          const int c = sl + k * LPE;
#pragma unroll
          for (int j = 0; j < 5; ++j) frow[k][j] = 0.f;
          if (c < n_comp) feature_row(comps[c], frow[k]);
        }
      }
      if (live) {
        float* feat = p.base.feat + (size_t)se * (uint32_t)(C * 5);
        uint8_t* pm = p.base.place_mask + (size_t)se * (uint32_t)C;
        uint8_t* cm = p.base.comp_mask + (size_t)se * (uint32_t)C;
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
          const int c = sl + k * LPE;
          if (c < C) {
            float* row = feat + c * 5;
#pragma unroll
            for (int j = 0; j < 5; ++j) __stcs(row + j, frow[k][j]);
            st_stream(pm + c, (uint8_t)(c < cur ? 1 : 0));     // rect:284
            st_stream(cm + c, (uint8_t)(c < n_comp ? 1 : 0));  // rect:296-298
          }
        }
        for (int c = sl + kPre * LPE; c < C; c += LPE) {  // more than kPre * LPE components
          float f[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
          if (c < n_comp) feature_row(comps[c], f);
          float* row = feat + c * 5;
#pragma unroll
          for (int j = 0; j < 5; ++j) __stcs(row + j, f[j]);
          st_stream(pm + c, (uint8_t)(c < cur ? 1 : 0));
          st_stream(cm + c, (uint8_t)(c < n_comp ? 1 : 0));
        }
      }
    }
    feat_dirty = false;
    __syncwarp();  // the next step's placement rewrites comps[cur]
  }

  // ---- state: registers / shared memory -> HBM (once) ----
  if (!live) return;
  if (sl == 0) {
    uint4 h4, d4;
    h4.x = hdr0; h4.y = (uint32_t)cur | ((uint32_t)flags << 8); h4.z = slot; h4.w = ep_len;
    const uint64_t rb = (uint64_t)__double_as_longlong(ep_ret);
    d4.x = (uint32_t)rb; d4.y = (uint32_t)(rb >> 32); d4.z = (uint32_t)epno; d4.w = (uint32_t)(epno >> 32);
    *reinterpret_cast<uint4*>(gs) = h4;
    *reinterpret_cast<uint4*>(gs + 16) = d4;
  }
  for (int c = sl; c < C; c += LPE) reinterpret_cast<uint32_t*>(gs + sh.st_comps())[c] = comps[c];
  reinterpret_cast<uint32_t*>(gs + sh.st_occ())[sl] = occ;
}

}  // namespace pcb
#endif
