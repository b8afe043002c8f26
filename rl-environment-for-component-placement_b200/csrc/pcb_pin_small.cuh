// pcb_pin_small.cuh -- the random-policy rollout of the PIN environments
// (dummy_env_rectangular_pin.py:1544-1710, dummy_env_rectangular_pin_spatial.py:1487-1661) as ONE launch for many steps, LPE lanes per
// environment: the layout of pcb_small.cuh extended to components with pins.
//
// Why: the per-step kernels give one warp to an environment and one launch (plus a reward
// launch) to a step.  That is right for 32x32 boards in batches of tens of thousands, and
// wrong twice elsewhere: (i) the reference's shipped configurations are 10x10 boards
// (agent/config/rectangle_pin_model.json) -- 10 of 32 lanes busy, ~700 warp-instructions
// per env-step, issue-bound at 0.14 of the HBM roofline; (ii) a small batch (4 096 envs of
// config C) is one wave of CTAs per launch, so every step costs a kernel's start-up and
// drain latency (14.5 us per step against 2.5 us of work).  Here
//   * LPE = 8 / 16 / 32 lanes own one environment (lane = board row), 32/LPE environments
//     per warp, every shuffle / vote issued by the whole warp (see pcb_small.cuh),
//   * components, pins, the NEXT instance (prefetched with cp.async at episode start) and
//     the images of the feature tensors live in shared memory for all n_steps; the feature
//     images are updated incrementally -- a placement rewrites one component row and the
//     rows of that component's pins -- and copied out every step,
//   * an episode that ends with every component placed leaves a routing job (absolute pin
//     positions) in a list; reward_list_kernel computes the terminal rewards of the whole
//     launch afterwards (find_reward, pin:832-975) and writes them into the [T, N] logs.
// The arithmetic, the counters of the random policy and the job format are those of the
// per-step kernels: a rollout gives the same observations, rewards and state bit for bit
// (tests/test_gpu_api.py::test_pin_fused_rollout_equals_per_step).
#ifndef PCB_PIN_SMALL_CUH_
#define PCB_PIN_SMALL_CUH_

#include "pcb_small.cuh"

namespace pcb {

struct PinRolloutParams {
  Dims d;
  Out base;              // slice 0 / half 0 of every tensor
  uint8_t* state;
  const uint8_t* bank;
  const float* ratio_lut;
  double* stats;
  const int32_t* actions;  // nullptr: the device random policy; else row t at actions + t * action_stride
  int action_format;
  size_t action_stride;
  uint32_t res_stride;   // bytes between the two halves of the reward | done block
  int32_t* actions_log;  // nullptr or [n_steps, N, 3]
  double* reward_log;    // nullptr or [n_steps, N]   (rows of THIS launch)
  uint8_t* done_log;     // nullptr or [n_steps, N]
  uint64_t seed;
  uint64_t step_no;      // number of the first step of this launch
  uint32_t bank_count;
  uint32_t slot_step;    // n_envs % bank_count
  int n_envs, n_steps;
  int ring, slice0;
  uint8_t* jobs;         // routing jobs of this launch, [job_cap][job_stride]
  uint32_t* job_total;   // their number (zeroed by the host before the launch)
  uint32_t job_cap;
};

// shared memory per environment: state, the next bank record, the images of the tensors that
// change a little per step (feature rows, pin rows; spatial env: pin_grid, component_grid)
struct PinSmem {
  int comps, pins, orient, nrec, pnum, pcat, feat, pgrid, cgrid, flat, total;
  int pnum_b, pcat_b, feat_b, pgrid_b, cgrid_b;  // bytes of each image = bytes of one env's slice
};
PCB_HD constexpr PinSmem pin_smem_layout(int kind, int C, int P, int max_pins, int H, int W, int K,
                                         int mh, int mw) {
  PinSmem s{};
  const int rows = kind == 3 ? C * P + 1 : C * P, F = kind == 3 ? 5 + P : 5;
  s.pnum_b = rows * 8; s.pcat_b = kind == 3 ? rows * 4 : rows * 2; s.feat_b = C * F * 4;
  s.pgrid_b = kind == 3 ? H * W * K : 0; s.cgrid_b = kind == 3 ? C * mh * mw * K : 0;
  s.comps = 0;
  s.pins = s.comps + align16c(C * 4);
  s.orient = s.pins + align16c(max_pins * 8);
  s.nrec = s.orient + (kind == 3 ? align16c(C) : 0);
  s.pnum = s.nrec + 16 + align16c(C * 4) + align16c(max_pins * 8);  // a bank record
  s.pcat = s.pnum + align16c(s.pnum_b);
  s.feat = s.pcat + align16c(s.pcat_b);
  s.pgrid = s.feat + align16c(s.feat_b);
  s.cgrid = s.pgrid + align16c(s.pgrid_b);
  s.flat = s.cgrid + align16c(s.cgrid_b);
  s.total = s.flat + align16c(4 * ((H * W + 31) / 32 + 2));
  return s;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// smem image -> global slice of one env by the LPE lanes of its group; the store width follows
// the slice size (the tensor base is 256-byte aligned and the slice offset a multiple of the size)
template <int LPE>
__device__ __forceinline__ void copy_out_g(void* dst, const void* src, int nbytes, int sl, bool live) {
  if (!live) return;
  if ((nbytes & 15) == 0) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = sl; i < (nbytes >> 4); i += LPE) st_stream(d + i, s[i]);
  } else if ((nbytes & 7) == 0) {
    const uint2* s = reinterpret_cast<const uint2*>(src);
    uint2* d = reinterpret_cast<uint2*>(dst);
    for (int i = sl; i < (nbytes >> 3); i += LPE) st_stream(d + i, s[i]);
  } else if ((nbytes & 3) == 0) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (int i = sl; i < (nbytes >> 2); i += LPE) st_stream(d + i, s[i]);
  } else {
    const uint8_t* s = reinterpret_cast<const uint8_t*>(src);
    uint8_t* d = reinterpret_cast<uint8_t*>(dst);
    for (int i = sl; i < nbytes; i += LPE) st_stream(d + i, s[i]);
  }
}

template <int LPE, typename S>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
rollout_pin_kernel(const __grid_constant__ PinRolloutParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const S sh(d);
  constexpr int G = 32 / LPE;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int sl = lane & (LPE - 1), eic = wic * G + lane / LPE;
  const int glead = lane & ~(LPE - 1);  // first lane of this group
  unsigned gmask = kFull;
  if constexpr (LPE < 32) gmask = ((1u << LPE) - 1u) << glead;
  const int env_raw = blockIdx.x * ((int)(blockDim.x >> 5) * G) + eic;
  build_expand_lut(smem);
  const bool live = env_raw < p.n_envs;
  if (!__any_sync(kFull, live)) return;
  const int env = live ? env_raw : 0;  // a group past the last env mirrors env 0, stores off
  const uint2* lut = reinterpret_cast<const uint2*>(smem);
  const int H = sh.H(), W = sh.W(), C = sh.C(), HW = H * W, P = sh.P();
  const int max_pins = sh.max_pins();
  const int kind = sh.kind(), K = sh.K(), F = sh.F(), mh = sh.mh(), mw = sh.mw();
  const PinSmem L = pin_smem_layout(kind, C, P, max_pins, H, W, K, mh, mw);
  uint8_t* es = smem + kLutBytes + (size_t)eic * L.total;
  uint32_t* comps = reinterpret_cast<uint32_t*>(es + L.comps);
  PinRec* pins = reinterpret_cast<PinRec*>(es + L.pins);
  uint8_t* nrec = es + L.nrec;
  int16_t* inum = reinterpret_cast<int16_t*>(es + L.pnum);
  int16_t* icat = reinterpret_cast<int16_t*>(es + L.pcat);
  float* ifeat = reinterpret_cast<float*>(es + L.feat);
  uint8_t* ipgrid = es + L.pgrid;    // spatial env: pin_grid (H, W, K)
  uint8_t* icgrid = es + L.cgrid;    // spatial env: component_grid (C, mh, mw, K)
  uint8_t* orient = es + L.orient;   // spatial env: orientation each component was placed with
  uint32_t* flat = reinterpret_cast<uint32_t*>(es + L.flat);
  const uint32_t N = (uint32_t)p.n_envs;
  const int pmax = d.mh > d.mw ? d.mh : d.mw;
  const bool small = pmax <= 4;
  const int rec_stride = sh.rec_stride();
  const int jstride = job_stride(max_pins, sh.K());

  // ---- state: HBM -> registers / shared memory (once) ----
  uint8_t* gs = p.state + (size_t)env * sh.st_stride();
  int n_comp, n_nets, n_pins, cur, flags;
  uint32_t slot, ep_len;
  double ep_ret;
  uint64_t epno;
  {
    const uint4 h4 = *reinterpret_cast<const uint4*>(gs);
    const uint4 d4 = *reinterpret_cast<const uint4*>(gs + 16);
    n_comp = h4.x & 0xff; n_nets = (h4.x >> 8) & 0xff; n_pins = h4.x >> 16;
    cur = h4.y & 0xff; flags = (h4.y >> 8) & 0xff;
    slot = h4.z; ep_len = h4.w;
    if (p.bank_count && slot >= p.bank_count) slot %= p.bank_count;  // the bank may have shrunk
    ep_ret = __longlong_as_double((long long)(((uint64_t)d4.y << 32) | d4.x));
    epno = ((uint64_t)d4.w << 32) | d4.z;
  }
  for (int c = sl; c < C; c += LPE) comps[c] = reinterpret_cast<const uint32_t*>(gs + sh.st_comps())[c];
  for (int q = sl; q < max_pins; q += LPE)
    reinterpret_cast<uint2*>(pins)[q] = reinterpret_cast<const uint2*>(gs + sh.st_pins())[q];
  uint32_t occ = reinterpret_cast<const uint32_t*>(gs + sh.st_occ())[sl];
  if (kind == 3)
    for (int c = sl; c < C; c += LPE) orient[c] = gs[sh.st_orient() + c];
  __syncwarp();

  // next instance of this env: requested at the start of an episode, consumed at its end
  auto prefetch = [&](uint32_t cur_slot) {
    uint32_t s2 = cur_slot + p.slot_step;
    if (p.bank_count && s2 >= p.bank_count) s2 -= p.bank_count;
    const uint8_t* rec = p.bank + (size_t)s2 * rec_stride;
    for (int i = sl; i < (rec_stride >> 4); i += LPE) cp_async16(nrec + 16 * i, rec + 16 * i);
    cp_async_commit();
  };
  prefetch(slot);

  // ---- images of the feature tensors (pin:1521-1542, 217-239; spatial:203-239, 1469-1485,
  // 1663-1697), rebuilt at a reset and updated by every placement ----
  auto feature_row = [&](int c) {  // [h, w, x|-1, y|-1, h*w/(H*W)]; the spatial env's pin ids stay
    float* row = ifeat + c * F;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f;
    if (c < n_comp) {
      const uint32_t r = comps[c];
      const int h = r & 0xff, w = (r >> 8) & 0xff, cx = (r >> 16) & 0xff, cy = r >> 24;
      f0 = (float)h; f1 = (float)w;
      f2 = cx == 0xFF ? -1.f : (float)cx;
      f3 = cx == 0xFF ? -1.f : (float)cy;
      if (S::fixed && (HW & (HW - 1)) == 0) f4 = (float)(h * w) * (1.0f / (float)HW);
      else f4 = p.ratio_lut[h * w];
    }
    row[0] = f0; row[1] = f1; row[2] = f2; row[3] = f3; row[4] = f4;
  };
  auto pin_row = [&](int q) {  // the winner of a feature row writes it (spatial: every pin, row = pin_id)
    const PinRec pr = pins[q];
    if (!(pr.flags & kPinWinner)) return;
    const uint32_t r = comps[pr.comp];
    const int cx = (r >> 16) & 0xff, cy = r >> 24;
    const int ax = cx == 0xFF ? -1 : cx + pr.rx, ay = cx == 0xFF ? -1 : cy + pr.ry;
    const int row = kind == 3 ? (int)pr.pin_id : pr.comp * P + pr.pin_id;
    *reinterpret_cast<uint2*>(inum + row * 4) =
        make_uint2((uint32_t)pr.rx | ((uint32_t)pr.ry << 16), ((uint32_t)ax & 0xffffu) | ((uint32_t)ay << 16));
    if (kind == 3) *reinterpret_cast<uint32_t*>(icat + row * 2) = (uint32_t)pr.net | ((uint32_t)pr.comp << 16);
    else icat[row] = pr.net;
  };
  // comps / pins final in shared memory (after a __syncwarp); called by the whole warp, the
  // groups with `on` unset only take part in the barriers
  auto rebuild_images = [&](bool on) {
    if (on) {
      for (int i = sl; i < (L.flat - L.pnum) >> 4; i += LPE) reinterpret_cast<uint4*>(es + L.pnum)[i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    if (on) {
      for (int c = sl; c < C; c += LPE) {
        feature_row(c);
        if (kind == 3) {  // the pin ids of the component, padded with -1 (rows of absent components stay 0)
          const float pad = c < n_comp ? -1.f : 0.f;
          for (int f = 5; f < F; ++f) ifeat[c * F + f] = pad;
        }
      }
      if (kind == 3) {
        if (sl == 0) { icat[2 * (C * P)] = -1; icat[2 * (C * P) + 1] = -1; }  // the sentinel row
        // component_grid: channel 0 all ones over the box of every component (spatial:1677-1697)
        const int cells = mh * mw;
        for (int i = sl; i < n_comp * cells; i += LPE) icgrid[i * K] = 1;
        // pin_grid channel 0: occupied cells (all of them; the placed pins clear theirs below)
        if (sl < H)
          for (uint32_t b = occ; b; b &= b - 1) ipgrid[(sl * W + (__ffs((int)b) - 1)) * K] = 1;
      }
    }
    __syncwarp();
    if (on) {
      for (int q = sl; q < n_pins; q += LPE) {
        pin_row(q);
        if (kind == 3) {
          const PinRec pr = pins[q];
          ifeat[pr.comp * F + 5 + pr.rank] = (float)pr.pin_id;
          // component_grid is drawn from the UN-rotated coordinates: undo the placement rotation
          const uint32_t r = comps[pr.comp];
          const int ch_ = r & 0xff, cw_ = (r >> 8) & 0xff, oc = orient[pr.comp];
          int rx = pr.rx, ry = pr.ry;
          if (oc == 1) { rx = ch_ - 1 - pr.ry; ry = pr.rx; }
          else if (oc == 2) { rx = ch_ - 1 - pr.rx; ry = cw_ - 1 - pr.ry; }
          else if (oc == 3) { rx = pr.ry; ry = cw_ - 1 - pr.rx; }
          icgrid[((pr.comp * mh + rx) * mw + ry) * K + pr.net + 1] = 1;
          // pin_grid (spatial:1663-1675): a placed pin clears channel 0 of its cell and sets channel net + 1
          const int cx = (r >> 16) & 0xff, cy = r >> 24;
          if (cx != 0xFF) {
            uint8_t* cell = ipgrid + ((cx + pr.rx) * W + cy + pr.ry) * K;
            cell[0] = 0;
            cell[pr.net + 1] = 1;
          }
        }
      }
    }
    __syncwarp();
  };
  rebuild_images(true);

  uint32_t m0, m1;
  auto masks = [&]() {
    const bool have = cur < n_comp;
    const uint32_t r = comps[have ? cur : 0];
    const int ch = r & 0xff, cw = (r >> 8) & 0xff;
    const uint32_t a = footprint_mask_g<LPE>(occ, ch, cw, H, W, sl, small, pmax);
    const uint32_t b = footprint_mask_g<LPE>(occ, cw, ch, H, W, sl, small, pmax);
    m0 = have ? a : 0u;
    m1 = have ? b : 0u;
    return group_any<LPE>((m0 | m1) != 0u, gmask);
  };
  bool any = masks();

  uint32_t slice = (uint32_t)p.slice0;
  for (int t = 0; t < p.n_steps; ++t) {
    const uint64_t step_no = p.step_no + (uint64_t)t;
    const uint32_t he = ((uint32_t)step_no & 1u) ? N + (uint32_t)env : (uint32_t)env;
    const size_t hoff = ((uint32_t)step_no & 1u) ? (size_t)p.res_stride : 0;

    int o = 0, x = 0, y = 0, ph = 0, pw = 0, ch = 0, cw = 0;
    bool valid;
    if (p.actions) {
      // ---- given actions (pcb_step semantics; FlatteningActionWrapperRect: env_wrappers.py:80-98) ----
      const int32_t* acts = p.actions + (size_t)t * p.action_stride;
      if (p.action_format == 1) {
        const int a = acts[env];
        if (a < 0) { o = -1; }
        else { o = a / HW; const int rem = a - o * HW; x = rem / W; y = rem - x * W; }
      } else {
        o = acts[3 * env]; x = acts[3 * env + 1]; y = acts[3 * env + 2];
      }
      bool in_range = (unsigned)o < 4u && (unsigned)x < (unsigned)H && (unsigned)y < (unsigned)W && cur < n_comp;
      const uint32_t r = comps[in_range ? cur : 0];
      ch = r & 0xff; cw = (r >> 8) & 0xff;
      ph = (o & 1) ? cw : ch;  // footprint (h,w) for o in {0,2}, (w,h) for {1,3}: pin:1755-1760
      pw = (o & 1) ? ch : cw;
      const bool fits = in_range && x + ph <= H && y + pw <= W;
      const uint32_t cols = fits ? (low_mask<uint32_t>(pw) << y) : 0u;
      const bool clash = group_any<LPE>(sl >= x && sl < x + ph && (occ & cols), gmask);
      valid = fits && !clash;
    } else {
      // ---- the random policy exactly as sample_action (pcb_kernels.cuh) draws it: up to
      // PCB_SAMPLE_TRIES uniform cells of the 4 x H x W box, the first whose mask bit is set;
      // otherwise the k-th set bit in np.argwhere order ----
      bool acc = false;
      if (__any_sync(kFull, any)) {
        for (int tr = 0; tr < PCB_SAMPLE_TRIES; ++tr) {
          const uint64_t r = action_rand64(p.seed + 0x632BE59BD9B4E019ull * (uint64_t)(tr + 1), (uint32_t)env, step_no);
          const int to = (int)((((uint32_t)r & 0x1fffffu) * 4u) >> 21);
          const int row = (int)((((uint32_t)(r >> 21) & 0x1fffffu) * (uint32_t)H) >> 21);
          const int col = (int)((((uint32_t)(r >> 42) & 0x1fffffu) * (uint32_t)W) >> 21);
          const uint32_t mine = (to & 1) ? m1 : m0;
          const uint32_t bits = __shfl_sync(kFull, mine, row, LPE);
          if (!acc && any && ((bits >> col) & 1u)) { acc = true; o = to; x = row; y = col; }
          if (__all_sync(kFull, acc || !any)) break;
        }
      }
      if (__any_sync(kFull, !acc)) {
        const uint32_t c01 = (uint32_t)__popc(m0) | ((uint32_t)__popc(m1) << 16);
        uint32_t incl = c01;
#pragma unroll
        for (int off = 1; off < LPE; off <<= 1) {
          const uint32_t v = __shfl_up_sync(kFull, incl, off, LPE);
          if (sl >= off) incl += v;
        }
        const uint32_t tot = __shfl_sync(kFull, incl, LPE - 1, LPE);
        const int tot0 = (int)(tot & 0xffffu), tot1 = (int)(tot >> 16), per_pair = tot0 + tot1;
        const int total = 2 * per_pair;
        const uint32_t r = action_rand(p.seed, (uint32_t)env, step_no);
        int kth = (int)__umulhi(r, (uint32_t)total);
        int obase = 0;
        if (kth >= per_pair) { kth -= per_pair; obase = 2; }
        const bool second = kth >= tot0;
        if (second) kth -= tot0;
        const uint32_t bits = second ? m1 : m0;
        const int inc = (int)(second ? incl >> 16 : incl & 0xffffu);
        const int excl = inc - __popc(bits);
        const unsigned who = __ballot_sync(kFull, kth >= excl && kth < inc) >> glead;
        const int src = (__ffs(who) - 1) & (LPE - 1);
        const int col = __shfl_sync(kFull, nth_bit32(bits, kth - excl), src, LPE);
        if (!acc) {
          const bool have = total > 0;
          o = have ? obase + (second ? 1 : 0) : 0;
          x = have ? src : 0;
          y = have ? col : 0;
        }
      }
      valid = any;
      if (valid) {
        const uint32_t r = comps[cur];
        ch = r & 0xff; cw = (r >> 8) & 0xff;
        ph = (o & 1) ? cw : ch;
        pw = (o & 1) ? ch : cw;
      }
    }
    if (p.actions_log && sl == 0 && live) {
      int32_t* a = p.actions_log + ((size_t)t * N + (size_t)env) * 3;
      a[0] = o; a[1] = x; a[2] = y;
    }
    // update_grid (pin:1738-1765)
    if (valid && sl >= x && sl < x + ph) occ |= low_mask<uint32_t>(pw) << y;
    __syncwarp();  // every lane has read comps[cur]
    if (valid) {
      if (sl == 0) {
        comps[cur] = (comps[cur] & 0xffffu) | ((uint32_t)x << 16) | ((uint32_t)y << 24);
        if (kind == 3) orient[cur] = (uint8_t)o;
      }
      if (kind == 3 && sl >= x && sl < x + ph)  // pin_grid channel 0 of the new footprint
        for (int j = y; j < y + pw; ++j) ipgrid[(sl * W + j) * K] = 1;
      if (o != 0) {  // Component.place_component (pin:156-204): the relative coordinates rotate in place
        for (int q = sl; q < n_pins; q += LPE) {
          PinRec pr = pins[q];
          if (pr.comp != cur) continue;
          const int rx = pr.rx, ry = pr.ry;
          int nx, ny;
          if (o == 1) { nx = ry; ny = ch - rx - 1; }
          else if (o == 2) { nx = ch - rx - 1; ny = cw - ry - 1; }
          else { nx = cw - ry - 1; ny = rx; }
          pins[q].rx = (uint8_t)nx;
          pins[q].ry = (uint8_t)ny;
        }
      }
    }
    __syncwarp();
    if (valid) {  // the images: one component row, the rows of its pins
      if (sl == 0) feature_row(cur);
      for (int q = sl; q < n_pins; q += LPE) {
        const PinRec pr = pins[q];
        if (pr.comp != cur) continue;
        pin_row(q);
        if (kind == 3) {
          uint8_t* cell = ipgrid + ((x + pr.rx) * W + y + pr.ry) * K;
          cell[0] = 0;
          cell[pr.net + 1] = 1;
        }
      }
      cur += 1;
    }
    ep_len += 1;
    __syncwarp();

    // ---- done (compute_if_done, pin:1872-1886; an invalid action ends the episode) ----
    any = masks();
    const bool done = valid ? !any : true;
    const bool route = done && cur == n_comp;  // find_reward needs the routes
    const double reward = done ? d.pen_reward : 0.0;  // pin:1669-1677, 899-909; routed: set later
    if (!route) ep_ret += reward;
    const bool first_report = !(flags & kFlagFinished);
    // routing job: positions of all pins (the job format of step_kernel / reward_kernel)
    uint32_t jidx = 0;
    if (route && sl == 0 && live) jidx = atomicAdd(p.job_total, 1u);
    jidx = __shfl_sync(kFull, jidx, 0, LPE);
    if (route && live && jidx < p.job_cap) {
      uint8_t* job = p.jobs + (size_t)jidx * jstride;
      if (sl == 0) {
        RouteJob j;
        j.env = (uint32_t)env; j.n_pins = (uint16_t)n_pins; j.n_nets = (uint8_t)n_nets;
        j.count_stats = first_report ? 1 : 0; j.ep_len = ep_len; j.pad = (uint32_t)t;
        j.ep_return = ep_ret; j.pad2 = 0.0;
        *reinterpret_cast<RouteJob*>(job) = j;
      }
      JobPin* jp = reinterpret_cast<JobPin*>(job + 32);
      for (int q = sl; q < n_pins; q += LPE) {
        const PinRec pr = pins[q];
        const uint32_t r = comps[pr.comp];
        jp[q] = JobPin{(uint8_t)(((r >> 16) & 0xff) + pr.rx), (uint8_t)((r >> 24) + pr.ry), pr.net, 0};
      }
    }
    if (done && !route && sl == 0 && live) {  // episode over without a route: penalty infos
      p.base.ep_return[he] = ep_ret;
      p.base.ep_length[he] = (int32_t)ep_len;
      p.base.wl[he] = d.pen_wl;
      p.base.ni[he] = d.pen_ni;
      if (first_report) {
        double* st = p.stats + (size_t)(env & (kStatSlots - 1)) * 8;
        atomicAdd(st + 0, 1.0);
        atomicAdd(st + 1, ep_ret);
        atomicAdd(st + 2, (double)ep_len);
        atomicAdd(st + 3, d.pen_wl);
        atomicAdd(st + 4, d.pen_ni);
      }
    }
    if (sl == 0 && live) {
      p.base.done[hoff + (size_t)env] = done ? 1 : 0;
      const size_t le = (size_t)t * N + (size_t)env;
      if (p.done_log) p.done_log[le] = done ? 1 : 0;
      if (!route) {
        *reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(p.base.reward + env) + hoff) = reward;
        if (p.reward_log) p.reward_log[le] = reward;
      }
    }
    __syncwarp();  // the job has been read out of comps / pins
    if (done) {
      if (!d.autoreset) {
        flags |= kFlagFinished;
      } else {
        // reset(): next bank slot (pin:1544-1597); its record was requested at episode start
        cp_async_wait_all();
        slot += p.slot_step;
        if (p.bank_count && slot >= p.bank_count) slot -= p.bank_count;
        epno += 1;
        cur = 0; flags = 0; ep_len = 0; ep_ret = 0.0; occ = 0u;
      }
    }
    const bool restart = done && d.autoreset;
    if (__any_sync(kFull, restart)) {  // warp-uniform; the groups that do not restart idle
      __syncwarp();
      if (restart) {
        const uint32_t h0 = *reinterpret_cast<const uint32_t*>(nrec);
        n_comp = h0 & 0xff; n_nets = (h0 >> 8) & 0xff; n_pins = h0 >> 16;
        const uint32_t* rc = reinterpret_cast<const uint32_t*>(nrec + sh.rec_comps());
        const uint2* rp = reinterpret_cast<const uint2*>(nrec + sh.rec_pins());
        for (int c = sl; c < C; c += LPE) comps[c] = rc[c];
        for (int q = sl; q < max_pins; q += LPE) reinterpret_cast<uint2*>(pins)[q] = rp[q];
        if (kind == 3)
          for (int c = sl; c < C; c += LPE) orient[c] = 0;
      }
      __syncwarp();
      if (restart) prefetch(slot);
      rebuild_images(restart);
      any = masks();
    }

    // ---- observation of slice (slice0 + 1 + t) % ring ----
    slice = slice + 1u == (uint32_t)p.ring ? 0u : slice + 1u;
    const uint32_t se = slice * N + (uint32_t)env;
    write_board_g<LPE, S>(sh, p.base.grid + (size_t)se * HW, occ, flat, sl, live, lut);
    uint8_t* mk = p.base.mask + (size_t)se * (uint32_t)(4 * HW);
    // orientations 2, 3 repeat 0, 1 (pin:1868-1869)
    if (W == 8 || W == 16 || W == 32) {
      // register -> table -> store boards: expanding twice beats storing one expansion twice
      // (measured, C at 8 192 envs: 6.4e8 against 5.6e8 env-steps/s)
      write_board_g<LPE, S>(sh, mk, m0, flat, sl, live, lut);
      write_board_g<LPE, S>(sh, mk + HW, m1, flat, sl, live, lut);
      write_board_g<LPE, S>(sh, mk + 2 * HW, m0, flat, sl, live, lut);
      write_board_g<LPE, S>(sh, mk + 3 * HW, m1, flat, sl, live, lut);
    } else {
      // dense-bit-string boards: pack once, store both copies (REF_pin: 1.07e9 -> 1.13e9)
      write_board_g<LPE, S>(sh, mk, m0, flat, sl, live, lut, 2 * HW);
      write_board_g<LPE, S>(sh, mk + HW, m1, flat, sl, live, lut, 2 * HW);
    }
    copy_out_g<LPE>(p.base.feat + (size_t)se * (uint32_t)(C * F), ifeat, L.feat_b, sl, live);
    copy_out_g<LPE>(p.base.pnum + (size_t)se * (uint32_t)(L.pnum_b >> 1), inum, L.pnum_b, sl, live);
    copy_out_g<LPE>(p.base.pcat + (size_t)se * (uint32_t)(L.pcat_b >> 1), icat, L.pcat_b, sl, live);
    if (kind == 3) {
      copy_out_g<LPE>(p.base.pgrid + (size_t)se * (uint32_t)L.pgrid_b, ipgrid, L.pgrid_b, sl, live);
      copy_out_g<LPE>(p.base.cgrid + (size_t)se * (uint32_t)L.cgrid_b, icgrid, L.cgrid_b, sl, live);
    }
    if (live) {
      uint8_t* pm = p.base.place_mask + (size_t)se * (uint32_t)C;
      for (int c = sl; c < C; c += LPE)  // pin:1577-1579, 1649-1657
        st_stream(pm + c, (uint8_t)(c >= n_comp ? 0 : (c < cur ? 2 : (c == cur ? 3 : 1))));
    }
    __syncwarp();  // the next step rewrites comps / pins / images
  }

  // ---- state: registers / shared memory -> HBM (once) ----
  cp_async_wait_all();
  if (!live) return;
  if (sl == 0) {
    uint4 h4, d4;
    h4.x = (uint32_t)n_comp | ((uint32_t)n_nets << 8) | ((uint32_t)n_pins << 16);
    h4.y = (uint32_t)cur | ((uint32_t)flags << 8); h4.z = slot; h4.w = ep_len;
    const uint64_t rb = (uint64_t)__double_as_longlong(ep_ret);
    d4.x = (uint32_t)rb; d4.y = (uint32_t)(rb >> 32); d4.z = (uint32_t)epno; d4.w = (uint32_t)(epno >> 32);
    *reinterpret_cast<uint4*>(gs) = h4;
    *reinterpret_cast<uint4*>(gs + 16) = d4;
  }
  for (int c = sl; c < C; c += LPE) reinterpret_cast<uint32_t*>(gs + sh.st_comps())[c] = comps[c];
  for (int q = sl; q < max_pins; q += LPE)
    reinterpret_cast<uint2*>(gs + sh.st_pins())[q] = reinterpret_cast<const uint2*>(pins)[q];
  reinterpret_cast<uint32_t*>(gs + sh.st_occ())[sl] = occ;
  if (kind == 3)
    for (int c = sl; c < C; c += LPE) gs[sh.st_orient() + c] = orient[c];
}

// Terminal rewards of the routing jobs one launch of rollout_pin_kernel left (find_reward,
// pin:832-975, centroid route): one warp per job, the routines of reward_kernel.  A job of step
// t (relative to the launch) writes reward_log[t][env]; jobs of the last two steps of the
// whole call also write the result tensors of their step parity, like the per-step path.
struct RewardListParams {
  Dims d;
  Out base;
  uint32_t res_stride;
  int n_envs;
  const uint8_t* jobs;
  const uint32_t* job_total;
  uint32_t job_cap;
  double* reward_log;   // rows of this launch or nullptr
  uint64_t step_no;     // number of the first step of the launch
  int final_from;       // jobs with step >= final_from also write the result halves
  double* stats;
};

template <bool BEAM>
__global__ void __launch_bounds__(kWarpsPerCta * 32, PCB_REWARD_MIN_BLOCKS)
reward_list_kernel(const __grid_constant__ RewardListParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int gw = blockIdx.x * kWarpsPerCta + wic, nw = gridDim.x * kWarpsPerCta;
  const uint32_t total = *p.job_total;
  const uint32_t count = total < p.job_cap ? total : p.job_cap;
  const int scr_bytes = route_scratch_bytes(d.max_pins, d.K);
  RouteScratch rs = route_scratch(smem + (size_t)wic * ((scr_bytes + 15) & ~15), d.max_pins, d.K);
  const uint32_t N = (uint32_t)p.n_envs;
  for (uint32_t j = gw; j < count; j += nw) {
    const uint8_t* job = p.jobs + (size_t)j * job_stride(d.max_pins, d.K);
    const RouteJob hdr = *reinterpret_cast<const RouteJob*>(job);
    const int n_pins = hdr.n_pins, n_nets = hdr.n_nets;
    const uint32_t env = hdr.env;
    const int t = (int)hdr.pad;
    route_prepare(reinterpret_cast<const JobPin*>(job + 32), rs, lane, n_pins, n_nets);
    int crossings = 0;
    double length = 0.0;
    if (!BEAM || d.reward_type == 0) {
      centroid_route(d, rs, lane, n_pins, n_nets, crossings, length);
    } else if constexpr (BEAM) {
      // the paths were routed by beam_kernel (one thread per net) into the job record
      const uint16_t* jn = reinterpret_cast<const uint16_t*>(job + job_path_offset(d.max_pins));
      for (int q = lane; q < n_pins; q += 32) rs.nxt[q] = jn[q];
      for (int k = lane; k < n_nets; k += 32) rs.start[k] = jn[d.max_pins + k];
      __syncwarp();
      beam_route(d, rs, lane, n_pins, n_nets, crossings, length);
      if (d.reward_type == 2) {  // "both": fewer crossings wins, beam on ties (pin:643-661)
        int c2; double l2;
        centroid_route(d, rs, lane, n_pins, n_nets, c2, l2);
        if (c2 < crossings) { crossings = c2; length = l2; }
      }
    }
    const double wl = length / d.wl_norm;
    const double ni = (double)crossings / d.ni_norm;
    const double reward = -1.0 * (d.ww * wl + d.wi * ni);
    if (lane == 0) {
      const double ret = hdr.ep_return + reward;
      if (p.reward_log) p.reward_log[(size_t)t * N + env] = reward;
      if (t >= p.final_from) {
        const bool odd = ((p.step_no + (uint64_t)t) & 1ull) != 0;
        const uint32_t he = odd ? N + env : env;
        const size_t hoff = odd ? (size_t)p.res_stride : 0;
        *reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(p.base.reward + env) + hoff) = reward;
        p.base.wl[he] = wl; p.base.ni[he] = ni;
        p.base.ep_return[he] = ret;
        p.base.ep_length[he] = (int32_t)hdr.ep_len;
      }
      if (hdr.count_stats) {
        double* st = p.stats + (size_t)(env & (kStatSlots - 1)) * 8;
        atomicAdd(st + 0, 1.0);
        atomicAdd(st + 1, ret);
        atomicAdd(st + 2, (double)hdr.ep_len);
        atomicAdd(st + 3, wl);
        atomicAdd(st + 4, ni);
      }
    }
    __syncwarp();
  }
}

// reward_list_kernel with two (four) jobs per warp: pcb_reward_half.cuh
template <int kG, bool BEAM = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, PCB_REWARD_MIN_BLOCKS)
reward_list_half_kernel(const __grid_constant__ RewardListParams p) {
  constexpr int kGroups = 32 / kG;
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int gl = lane & (kG - 1), grp = lane / kG, gbase = lane & ~(kG - 1);
  const uint32_t gw = blockIdx.x * kWarpsPerCta + wic, nw = gridDim.x * kWarpsPerCta;
  const uint32_t total = *p.job_total;
  const uint32_t count = total < p.job_cap ? total : p.job_cap;
  const int scr_bytes = (route_scratch_bytes(d.max_pins, d.K) + 15) & ~15;
  RouteScratch rs = route_scratch(smem + (size_t)(wic * kGroups + grp) * scr_bytes, d.max_pins, d.K);
  const uint32_t N = (uint32_t)p.n_envs;
  for (uint32_t j0 = gw * kGroups; j0 < count; j0 += nw * kGroups) {
    const uint32_t j = j0 + (uint32_t)grp;
    const bool has = j < count;
    const uint8_t* job = p.jobs + (size_t)(has ? j : j0) * job_stride(d.max_pins, d.K);
    const RouteJob hdr = *reinterpret_cast<const RouteJob*>(job);
    prefetch_job<kG>(d, p.jobs, j + nw * kGroups, count, gl);
    const GroupReward r = reward_of_group<kG, BEAM>(d, rs, job, hdr, has, gl, gbase);
    if (has && gl == 0) {
      const uint32_t env = hdr.env;
      const int t = (int)hdr.pad;
      const double ret = hdr.ep_return + r.reward;
      if (p.reward_log) p.reward_log[(size_t)t * N + env] = r.reward;
      if (t >= p.final_from) {
        const bool odd = ((p.step_no + (uint64_t)t) & 1ull) != 0;
        const uint32_t he = odd ? N + env : env;
        const size_t hoff = odd ? (size_t)p.res_stride : 0;
        *reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(p.base.reward + env) + hoff) = r.reward;
        p.base.wl[he] = r.wl; p.base.ni[he] = r.ni;
        p.base.ep_return[he] = ret;
        p.base.ep_length[he] = (int32_t)hdr.ep_len;
      }
      if (hdr.count_stats) {
        double* st = p.stats + (size_t)(env & (kStatSlots - 1)) * 8;
        atomicAdd(st + 0, 1.0);
        atomicAdd(st + 1, ret);
        atomicAdd(st + 2, (double)hdr.ep_len);
        atomicAdd(st + 3, r.wl);
        atomicAdd(st + 4, r.ni);
      }
    }
    __syncwarp();
  }
}

}  // namespace pcb
#endif
