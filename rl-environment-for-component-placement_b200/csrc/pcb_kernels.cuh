// pcb_kernels.cuh -- sm_100a kernels of the batched placement environments.
//
// One warp owns one environment.  The occupancy grid is a bitboard, one word per
// grid row, lane r holding row r (and row r+32 for grids taller than 32).  The
// valid-position mask of a (ph, pw) footprint is a sliding-window AND over the
// free bits: pw-1 shift/ANDs inside the row word, ph-1 warp shuffles across rows.
// Nothing here is a dense contraction, so no tensor cores: the kernels are bound
// by the HBM write-out of the observation slab (bit -> byte expansion + 16-byte
// coalesced stores).  Compiled with -fmad=false: every float64 product and sum
// below is individually rounded like the reference's Python/numpy arithmetic;
// the one fused multiply-add of the reference (numpy ddot inside linalg.norm) is
// written explicitly.
#ifndef PCB_KERNELS_CUH_
#define PCB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "pcb_layout.h"

namespace pcb {

constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T low_mask(int n) {
  return n >= (int)(sizeof(T) * 8) ? ~T(0) : (T(1) << n) - T(1);
}

__device__ __forceinline__ uint32_t spread4(uint32_t nib) {
  // 4 bits -> 4 bytes of 0/1 (bit i -> byte i); the partial products do not overlap.
  return (nib * 0x00204081u) & 0x01010101u;
}

__device__ __forceinline__ uint4 spread16(uint32_t b) {
  return make_uint4(spread4(b & 15u), spread4((b >> 4) & 15u), spread4((b >> 8) & 15u),
                    spread4((b >> 12) & 15u));
}

template <typename T>
__device__ __forceinline__ int popc_t(T v);
template <>
__device__ __forceinline__ int popc_t<uint32_t>(uint32_t v) { return __popc(v); }
template <>
__device__ __forceinline__ int popc_t<uint64_t>(uint64_t v) { return __popcll(v); }

// smem -> global copy of one environment's slice of a tensor, widest store the
// alignment allows (16 B when the per-env size is a multiple of 16).
__device__ __forceinline__ void copy_out(void* dst, const void* src, int nbytes, int lane) {
  uintptr_t a = reinterpret_cast<uintptr_t>(dst) | (uintptr_t)nbytes;
  if ((a & 15) == 0) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = lane; i < (nbytes >> 4); i += 32) d[i] = s[i];
  } else if ((a & 3) == 0) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (int i = lane; i < (nbytes >> 2); i += 32) d[i] = s[i];
  } else {
    const uint8_t* s = reinterpret_cast<const uint8_t*>(src);
    uint8_t* d = reinterpret_cast<uint8_t*>(dst);
    for (int i = lane; i < nbytes; i += 32) d[i] = s[i];
  }
}

__device__ __forceinline__ void zero_smem(void* p, int nbytes, int lane) {  // nbytes % 16 == 0
  uint4* d = reinterpret_cast<uint4*>(p);
  for (int i = lane; i < (nbytes >> 4); i += 32) d[i] = make_uint4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------
// bitboard: footprint mask (reference: compute_action_mask_orientation,
// dummy_env_rectangular.py:526-567 / dummy_env_rectangular_pin.py:1808-1851;
// square: dummy_env_square.py:74-113, 187-244)
//   out[r] bit j = 1  iff  r + ph <= H, j + pw <= W and rows r..r+ph-1,
//   columns j..j+pw-1 are all free.
// ---------------------------------------------------------------------------
template <int RPL, typename T>
__device__ __forceinline__ void footprint_mask(const T (&occ)[RPL], int ph, int pw, int H, int W,
                                               int lane, T (&out)[RPL]) {
  T e[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    int row = lane + 32 * k;
    T fr = row < H ? (T)(~occ[k] & low_mask<T>(W)) : T(0);
    T acc = fr;
    for (int dlt = 1; dlt < pw; ++dlt) acc &= fr >> dlt;  // zeros shift in above column W-1
    e[k] = acc;
    out[k] = acc;
  }
  for (int dlt = 1; dlt < ph; ++dlt) {
    int src = (lane + dlt) & 31;
    int hop = (lane + dlt) >> 5;  // how many 32-row slots the source row is ahead
    T got[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) got[k] = __shfl_sync(kFull, e[k], src);
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      T v = T(0);
#pragma unroll
      for (int kk = 0; kk < RPL; ++kk)
        if (k + hop == kk) v = got[kk];
      out[k] &= v;  // rows past the last slot count as blocked
    }
  }
}

// bit rows (in shared memory, one word per row) -> H*W bytes of 0/1 in global memory.
// dup_off != 0 writes a second copy at dst + dup_off (orientations 2,3 of the pin
// env repeat 0,1: dummy_env_rectangular_pin.py:1868-1869).
template <typename T>
__device__ __forceinline__ void write_bits(uint8_t* dst, const T* rows, int H, int W, int lane,
                                           int dup_off) {
  if ((W & 15) == 0) {
    const int cpr = W >> 4;  // 16-byte chunks per row: 1..4
    const int n = H * cpr;
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    uint4* e4 = reinterpret_cast<uint4*>(dst + dup_off);
    for (int q = lane; q < n; q += 32) {
      int row, part;
      if (cpr == 2) { row = q >> 1; part = q & 1; }
      else if (cpr == 4) { row = q >> 2; part = q & 3; }
      else if (cpr == 1) { row = q; part = 0; }
      else { row = q / 3; part = q - 3 * row; }
      uint4 v = spread16((uint32_t)(rows[row] >> (16 * part)) & 0xffffu);
      d4[q] = v;
      if (dup_off) e4[q] = v;
    }
  } else {
    const int n = H * W;
    for (int b = lane; b < n; b += 32) {
      int row = b / W, col = b - row * W;
      uint8_t v = (uint8_t)((rows[row] >> col) & T(1));
      dst[b] = v;
      if (dup_off) dst[b + dup_off] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// per-warp view of one environment staged in shared memory
// ---------------------------------------------------------------------------
template <typename T>
struct Env {
  uint8_t* sm;
  EnvHdr* hdr;
  EnvDyn* dyn;
  CompRec* comps;
  PinRec* pins;
  T* occ;      // 32*RPL words
  T* mrow;     // [2][32*RPL] mask rows of orientation 0 / 1
  uint8_t* stage;
  uint8_t* scr;
  __device__ __forceinline__ Env(uint8_t* base, const Dims& d, int rows)
      : sm(base),
        hdr(reinterpret_cast<EnvHdr*>(base)),
        dyn(reinterpret_cast<EnvDyn*>(base + 16)),
        comps(reinterpret_cast<CompRec*>(base + d.st_comps)),
        pins(reinterpret_cast<PinRec*>(base + d.st_pins)),
        occ(reinterpret_cast<T*>(base + d.st_occ)),
        mrow(reinterpret_cast<T*>(base + d.sm_mrow)),
        stage(base + d.sm_stage),
        scr(base + d.sm_scr) {}
};

__device__ __forceinline__ void load_state(uint8_t* sm, const uint8_t* g, int nbytes, int lane) {
  const uint4* s = reinterpret_cast<const uint4*>(g);
  uint4* d = reinterpret_cast<uint4*>(sm);
  for (int i = lane; i < (nbytes >> 4); i += 32) d[i] = s[i];
}

// ---------------------------------------------------------------------------
// observation write-out (everything the reference copies into the obs dict:
// rect:410-416, pin:1680-1687, spatial:1633-1644; layout SURVEY.md Appendix C)
// ---------------------------------------------------------------------------
template <int RPL, typename T>
__device__ __forceinline__ void write_obs(const Params& p, Env<T>& e, int env, int lane,
                                          const T (&m0)[RPL], const T (&m1)[RPL]) {
  const Dims& d = p.d;
  const int rows = 32 * RPL;
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    e.mrow[lane + 32 * k] = m0[k];
    e.mrow[rows + lane + 32 * k] = m1[k];
  }
  __syncwarp();
  const int HW = d.H * d.W;
  write_bits<T>(p.out.grid + (size_t)env * d.grid_b, e.occ, d.H, d.W, lane, 0);
  uint8_t* mk = p.out.mask + (size_t)env * d.mask_b;
  write_bits<T>(mk, e.mrow, d.H, d.W, lane, d.O == 4 ? 2 * HW : 0);
  if (d.O >= 2) write_bits<T>(mk + HW, e.mrow + rows, d.H, d.W, lane, d.O == 4 ? 2 * HW : 0);
  if (d.kind == 0) return;

  const int n_comp = e.hdr->n_comp, cur = e.hdr->cur, n_pins = e.hdr->n_pins;
  const double inv_area_den = (double)(d.H * d.W);

  // all_components_feature: [h, w, x|-1, y|-1, h*w/(H*W)] (+ pin ids, spatial)
  float* feat = p.out.feat + (size_t)env * d.feat_n;
  if (d.kind != 3) {
    for (int i = lane; i < d.feat_n; i += 32) {
      int c = i / 5, f = i - 5 * c;
      float v = 0.f;
      if (c < n_comp) {
        CompRec r = e.comps[c];
        if (f == 0) v = (float)r.h;
        else if (f == 1) v = (float)r.w;
        else if (f == 2) v = r.x == 0xFF ? -1.f : (float)r.x;
        else if (f == 3) v = r.y == 0xFF ? -1.f : (float)r.y;
        else v = (float)((double)(r.h * r.w) / inv_area_den);
      }
      feat[i] = v;
    }
  } else {
    float* st = reinterpret_cast<float*>(e.stage);
    for (int i = lane; i < d.feat_n; i += 32) {
      int c = i / d.F, f = i - d.F * c;
      float v = 0.f;
      if (c < n_comp) {
        CompRec r = e.comps[c];
        if (f == 0) v = (float)r.h;
        else if (f == 1) v = (float)r.w;
        else if (f == 2) v = r.x == 0xFF ? -1.f : (float)r.x;
        else if (f == 3) v = r.y == 0xFF ? -1.f : (float)r.y;
        else if (f == 4) v = (float)((double)(r.h * r.w) / inv_area_den);
        else v = -1.f;
      }
      st[i] = v;
    }
    __syncwarp();
    for (int q = lane; q < n_pins; q += 32) {
      PinRec pr = e.pins[q];
      st[pr.comp * d.F + 5 + pr.rank] = (float)pr.pin_id;
    }
    __syncwarp();
    copy_out(feat, st, d.feat_n * 4, lane);
    __syncwarp();
  }

  // placement_mask / component_mask
  for (int c = lane; c < d.C; c += 32) {
    if (d.kind == 1) {
      p.out.place_mask[(size_t)env * d.C + c] = c < cur ? 1 : 0;       // rect:284
      p.out.comp_mask[(size_t)env * d.C + c] = c < n_comp ? 1 : 0;     // rect:296-298
    } else {
      uint8_t v = c >= n_comp ? 0 : (c < cur ? 2 : (c == cur ? 3 : 1));  // pin:1577-1579, 1649-1657
      p.out.place_mask[(size_t)env * d.C + c] = v;
    }
  }
  if (d.kind == 1) return;

  // all_pins_num_feature / all_pins_cat_feature (pin:1521-1542, spatial:1469-1485)
  int16_t* snum = reinterpret_cast<int16_t*>(e.stage);
  int16_t* scat = snum + d.pnum_n;
  zero_smem(e.stage, ((d.pnum_n + d.pcat_n) * 2 + 15) & ~15, lane);
  __syncwarp();
  for (int q = lane; q < n_pins; q += 32) {
    PinRec pr = e.pins[q];
    if (!(pr.flags & kPinWinner)) continue;
    CompRec r = e.comps[pr.comp];
    int ax = r.x == 0xFF ? -1 : r.x + pr.rx;
    int ay = r.x == 0xFF ? -1 : r.y + pr.ry;
    int row = d.kind == 3 ? pr.pin_id : pr.comp * d.P + pr.pin_id;
    int16_t* nr = snum + row * 4;
    nr[0] = pr.rx; nr[1] = pr.ry; nr[2] = (int16_t)ax; nr[3] = (int16_t)ay;
    if (d.kind == 3) { scat[row * 2] = pr.net; scat[row * 2 + 1] = pr.comp; }
    else scat[row] = pr.net;
  }
  if (d.kind == 3 && lane == 0) { scat[d.pcat_n - 2] = -1; scat[d.pcat_n - 1] = -1; }
  __syncwarp();
  copy_out(p.out.pnum + (size_t)env * d.pnum_n, snum, d.pnum_n * 2, lane);
  copy_out(p.out.pcat + (size_t)env * d.pcat_n, scat, d.pcat_n * 2, lane);
  __syncwarp();
  if (d.kind != 3) return;

  // pin_grid (spatial:1663-1675): channel 0 = occupied cell without a pin,
  // channel net+1 = placed pin of that net.
  zero_smem(e.stage, (d.pgrid_b + 15) & ~15, lane);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    int row = lane + 32 * k;
    T bits = row < d.H ? e.occ[row] : T(0);
    while (bits) {
      int j = sizeof(T) == 8 ? __ffsll((long long)bits) - 1 : __ffs((int)bits) - 1;
      bits &= bits - 1;
      e.stage[(row * d.W + j) * d.K] = 1;
    }
  }
  __syncwarp();
  for (int q = lane; q < n_pins; q += 32) {
    PinRec pr = e.pins[q];
    CompRec r = e.comps[pr.comp];
    if (r.x == 0xFF) continue;
    uint8_t* cell = e.stage + ((r.x + pr.rx) * d.W + (r.y + pr.ry)) * d.K;
    cell[0] = 0;
    cell[pr.net + 1] = 1;
  }
  __syncwarp();
  copy_out(p.out.pgrid + (size_t)env * d.pgrid_b, e.stage, d.pgrid_b, lane);
  __syncwarp();

  // component_grid (spatial:1677-1697): drawn at reset from the UN-rotated pin
  // coordinates, i.e. from the bank record, and constant during the episode.
  zero_smem(e.stage, (d.cgrid_b + 15) & ~15, lane);
  __syncwarp();
  const int cells = d.mh * d.mw;
  for (int i = lane; i < n_comp * cells; i += 32) e.stage[i * d.K] = 1;
  __syncwarp();
  const PinRec* bp = reinterpret_cast<const PinRec*>(p.bank + (size_t)e.hdr->slot * d.rec_stride +
                                                     d.rec_pins);
  for (int q = lane; q < n_pins; q += 32) {
    PinRec pr = bp[q];
    e.stage[((pr.comp * d.mh + pr.rx) * d.mw + pr.ry) * d.K + pr.net + 1] = 1;
  }
  __syncwarp();
  copy_out(p.out.cgrid + (size_t)env * d.cgrid_b, e.stage, d.cgrid_b, lane);
  __syncwarp();
}

// Mask rows of the component to place next (both orientations); zero when none is left
// (rect:399-402, pin:1663-1666).  Returns true when at least one position is valid.
template <int RPL, typename T>
__device__ __forceinline__ bool next_masks(const Dims& d, const Env<T>& e, const T (&occ)[RPL],
                                           int cur, int n_comp, int lane, T (&m0)[RPL],
                                           T (&m1)[RPL]) {
#pragma unroll
  for (int k = 0; k < RPL; ++k) { m0[k] = T(0); m1[k] = T(0); }
  if (d.kind == 0) {
    footprint_mask<RPL, T>(occ, d.sq_n, d.sq_n, d.H, d.W, lane, m0);
  } else if (cur < n_comp) {
    CompRec r = e.comps[cur];
    footprint_mask<RPL, T>(occ, r.h, r.w, d.H, d.W, lane, m0);
    footprint_mask<RPL, T>(occ, r.w, r.h, d.H, d.W, lane, m1);
  }
  T any = T(0);
#pragma unroll
  for (int k = 0; k < RPL; ++k) any |= m0[k] | m1[k];
  return __any_sync(kFull, any != T(0));
}

// reset(): bank record -> state, empty board (rect:310-351, pin:1544-1597).
template <int RPL, typename T>
__device__ __forceinline__ void load_instance(const Params& p, Env<T>& e, uint32_t slot,
                                              uint64_t episode_no, int lane) {
  const Dims& d = p.d;
  if (d.kind != 0) {
    const uint8_t* rec = p.bank + (size_t)slot * d.rec_stride;
    // header (16 B) then comps/pins; state has the 16-byte EnvDyn in between
    if (lane == 0) *reinterpret_cast<uint4*>(e.sm) = *reinterpret_cast<const uint4*>(rec);
    const uint4* s = reinterpret_cast<const uint4*>(rec + 16);
    uint4* dd = reinterpret_cast<uint4*>(e.sm + 32);
    for (int i = lane; i < ((d.rec_stride - 16) >> 4); i += 32) dd[i] = s[i];
  }
  __syncwarp();
  if (lane == 0) {
    if (d.kind == 0) { e.hdr->n_comp = 1; e.hdr->n_nets = 0; e.hdr->n_pins = 0; e.hdr->pad = 0; }
    e.hdr->cur = 0;
    e.hdr->flags = 0;
    e.hdr->slot = slot;
    e.hdr->ep_len = 0;
    e.dyn->ep_return = 0.0;
    e.dyn->episode_no = episode_no;
  }
#pragma unroll
  for (int k = 0; k < RPL; ++k) e.occ[lane + 32 * k] = T(0);
  __syncwarp();
}

// ---------------------------------------------------------------------------
// K3 (K1/K2): step
// ---------------------------------------------------------------------------
template <int RPL, typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32) step_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerCta + wic;
  if (env >= p.n_envs) return;
  Env<T> e(smem + (size_t)wic * d.sm_stride, d, 32 * RPL);
  uint8_t* gstate = p.state + (size_t)env * d.st_stride;
  load_state(e.sm, gstate, d.st_stride, lane);
  __syncwarp();

  T occ[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) occ[k] = e.occ[lane + 32 * k];

  // decode the action (FlatteningActionWrapper*: env_wrappers.py:80-98, wrapper_square.py:54-69)
  int o = 0, x, y;
  if (p.action_format == 1) {
    int a = p.actions[env];
    int hw = d.H * d.W;
    if (a < 0) { o = -1; x = y = 0; }
    else { o = a / hw; int rem = a - o * hw; x = rem / d.W; y = rem - x * d.W; }
  } else if (d.kind == 0) {
    x = p.actions[2 * env]; y = p.actions[2 * env + 1];
  } else {
    o = p.actions[3 * env]; x = p.actions[3 * env + 1]; y = p.actions[3 * env + 2];
  }

  const int n_comp = e.hdr->n_comp;
  int cur = e.hdr->cur;
  int ph = d.sq_n, pw = d.sq_n, ch = 0, cw = 0;
  bool in_range = (unsigned)o < (unsigned)d.O && (unsigned)x < (unsigned)d.H &&
                  (unsigned)y < (unsigned)d.W;
  if (d.kind != 0) {
    in_range = in_range && cur < n_comp;
    if (in_range) {
      CompRec r = e.comps[cur];
      ch = r.h; cw = r.w;
      ph = (o & 1) ? cw : ch;  // footprint (h,w) for o in {0,2}, (w,h) for {1,3}: pin:1755-1760
      pw = (o & 1) ? ch : cw;
    }
  }
  // validate_action == mask lookup (pin:1712-1736); the mask bit is recomputed here.
  bool valid = in_range && x + ph <= d.H && y + pw <= d.W;
  T cols = T(0);
  if (valid) {
    cols = low_mask<T>(pw) << y;
    bool clash = false;
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      int row = lane + 32 * k;
      if (row >= x && row < x + ph && (occ[k] & cols)) clash = true;
    }
    valid = !__any_sync(kFull, clash);
  }

  if (valid) {
#pragma unroll
    for (int k = 0; k < RPL; ++k) {  // update_grid (pin:1738-1765)
      int row = lane + 32 * k;
      if (row >= x && row < x + ph) occ[k] |= cols;
      e.occ[row] = occ[k];
    }
    if (d.kind != 0) {
      if (lane == 0) { e.comps[cur].x = (uint8_t)x; e.comps[cur].y = (uint8_t)y; }
      if (d.kind >= 2 && o != 0) {  // Component.place_component (pin:156-204): rotate in place
        const int n_pins = e.hdr->n_pins;
        for (int q = lane; q < n_pins; q += 32) {
          PinRec pr = e.pins[q];
          if (pr.comp != cur) continue;
          int rx = pr.rx, ry = pr.ry, nx, ny;
          if (o == 1) { nx = ry; ny = ch - rx - 1; }
          else if (o == 2) { nx = ch - rx - 1; ny = cw - ry - 1; }
          else { nx = cw - ry - 1; ny = rx; }
          e.pins[q].rx = (uint8_t)nx;
          e.pins[q].ry = (uint8_t)ny;
        }
      }
      cur += 1;
    }
  }
  __syncwarp();

  T m0[RPL], m1[RPL];
  bool any = next_masks<RPL, T>(d, e, occ, cur, n_comp, lane, m0, m1);
  const bool all_placed = d.kind != 0 && cur == n_comp;
  const bool done = valid ? !any : true;  // compute_if_done (pin:1872-1886); invalid -> done
  const bool route = d.kind >= 2 && done && all_placed;  // find_reward needs the routes
  double reward;
  if (d.kind <= 1) reward = valid ? 1.0 : 0.0;      // square:115-153, rect:405-432
  else reward = done ? d.pen_reward : 0.0;          // pin:1669-1677, 899-909

  if (lane == 0) {
    e.hdr->cur = (uint8_t)cur;
    e.hdr->ep_len += 1;
    if (!route) e.dyn->ep_return += reward;
  }
  __syncwarp();
  // write the state back (only the parts a step can change)
  {
    uint4* g4 = reinterpret_cast<uint4*>(gstate);
    const uint4* s4 = reinterpret_cast<const uint4*>(e.sm);
    if (lane < 2) g4[lane] = s4[lane];
    if (valid) {
      for (int i = 2 + lane; i < (d.st_stride >> 4); i += 32) g4[i] = s4[i];
    }
  }

  if (!(done && d.autoreset)) write_obs<RPL, T>(p, e, env, lane, m0, m1);

  if (lane == 0) {
    p.out.done[env] = done ? 1 : 0;
    if (!route) {
      p.out.reward[env] = reward;
      if (d.kind >= 2 && done) { p.out.wl[env] = d.pen_wl; p.out.ni[env] = d.pen_ni; }
    }
  }
}

// ---------------------------------------------------------------------------
// K4: terminal reward (find_reward, pin:832-975)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double norm2(double dx, double dy, int use_fma) {
  // np.linalg.norm of a 2-vector = sqrt(ddot(v, v)); OpenBLAS accumulates with one fma
  return use_fma ? sqrt(fma(dy, dy, dx * dx)) : sqrt(dx * dx + dy * dy);
}

// is_intersect (pin:687-739), float64, operation order as written there.
__device__ __forceinline__ bool segments_cross(double x1, double y1, double x2, double y2,
                                               double x3, double y3, double x4, double y4) {
  if ((x1 == x3 && y1 == y3) || (x1 == x4 && y1 == y4) || (x2 == x3 && y2 == y3) ||
      (x2 == x4 && y2 == y4))
    return true;
  double dx12 = x1 - x2, dy12 = y1 - y2, dx34 = x3 - x4, dy34 = y3 - y4;
  double det = dx12 * dy34 - dy12 * dx34;
  if (det == 0.0) return false;
  double a = x1 * y2 - y1 * x2;
  double b = x3 * y4 - y3 * x4;
  double px = (a * dx34 - dx12 * b) / det;
  double py = (a * dy34 - dy12 * b) / det;
  return fmin(x1, x2) <= px && px <= fmax(x1, x2) && fmin(x3, x4) <= px && px <= fmax(x3, x4) &&
         fmin(y1, y2) <= py && py <= fmax(y1, y2) && fmin(y3, y4) <= py && py <= fmax(y3, y4);
}

struct RouteScratch {
  int16_t* px;       // [max_pins] absolute x of pin q
  int16_t* py;
  uint16_t* nxt;     // [max_pins] beam route: next pin on the path, 0xFFFF = path end
  uint16_t* start;   // [max_nets] beam route: first pin of the path of each net
  uint16_t* net_off; // [max_nets + 1]
  double* cx;        // [max_nets] centroid
  double* cy;
  double* dist;      // [max_pins] segment length of the segment headed by pin q
};

__device__ __forceinline__ RouteScratch route_scratch(uint8_t* scr, const Dims& d) {
  RouteScratch r;
  const int nets = d.K;  // K = max_num_nets + 1 >= n_nets + 1
  uint8_t* q = scr;
  r.cx = reinterpret_cast<double*>(q); q += nets * 8;
  r.cy = reinterpret_cast<double*>(q); q += nets * 8;
  r.dist = reinterpret_cast<double*>(q); q += d.max_pins * 8;
  r.px = reinterpret_cast<int16_t*>(q); q += d.max_pins * 2;
  r.py = reinterpret_cast<int16_t*>(q); q += d.max_pins * 2;
  r.net_off = reinterpret_cast<uint16_t*>(q); q += (nets + 1) * 2;
  r.start = reinterpret_cast<uint16_t*>(q); q += nets * 2;
  r.nxt = reinterpret_cast<uint16_t*>(q);
  return r;
}

__host__ __device__ inline int route_scratch_bytes(int max_pins, int K) {
  return K * 16 + max_pins * 8 + max_pins * 4 + (K + 1) * 2 + K * 2 + max_pins * 2 + 16;
}

// Centroid route (route_pins_centroid, pin:1296-1324): crossings + wirelength.
// Segment "headed" by pin q: (pin q, centroid of its net); for 2-pin nets the
// only segment is (pin 0, pin 1).
template <typename T>
__device__ __forceinline__ void centroid_route(const Dims& d, const Env<T>& e,
                                               const RouteScratch& rs, int lane, int n_pins,
                                               int n_nets, int& crossings, double& wl) {
  auto head = [&](int q, double& x2, double& y2) -> bool {
    int net = e.pins[q].net;
    int lo = rs.net_off[net], cnt = rs.net_off[net + 1] - lo;
    if (cnt == 2) {
      if (q != lo) return false;
      x2 = (double)rs.px[q + 1]; y2 = (double)rs.py[q + 1];
    } else {
      x2 = rs.cx[net]; y2 = rs.cy[net];
    }
    return true;
  };
  int cnt = 0;
  for (int s = 0; s < n_pins; ++s) {
    double sx2, sy2;
    if (!head(s, sx2, sy2)) continue;  // warp-uniform
    const double sx1 = (double)rs.px[s], sy1 = (double)rs.py[s];
    const int snet = e.pins[s].net;
    // segments of later nets start at the first pin of net snet+1
    for (int t = rs.net_off[snet + 1] + lane; t < n_pins; t += 32) {
      double tx2, ty2;
      if (!head(t, tx2, ty2)) continue;
      if (segments_cross(sx1, sy1, sx2, sy2, (double)rs.px[t], (double)rs.py[t], tx2, ty2)) ++cnt;
    }
  }
  for (int off = 16; off; off >>= 1) cnt += __shfl_xor_sync(kFull, cnt, off);
  crossings = cnt;
  // find_wirelength (pin:741-759): lengths in parallel, then the reference's sequential sum
  for (int q = lane; q < n_pins; q += 32) {
    double x2, y2, v = -1.0;
    if (head(q, x2, y2)) v = norm2((double)rs.px[q] - x2, (double)rs.py[q] - y2, d.norm_fma);
    rs.dist[q] = v;
  }
  __syncwarp();
  double acc = 0.0;
  if (lane == 0)
    for (int q = 0; q < n_pins; ++q)
      if (rs.dist[q] >= 0.0) acc += rs.dist[q];
  wl = __shfl_sync(kFull, acc, 0);
  __syncwarp();
}

}  // namespace pcb

#include "pcb_beam.cuh"

namespace pcb {

// ---------------------------------------------------------------------------
// K4 + K6: terminal kernel.  Runs after step_kernel; a warp whose environment did
// not finish exits at once.  For a finished environment: the routing reward when
// every component is placed (pin kinds), episode statistics, and -- with
// auto-reset -- the next instance from the bank plus its first observation.
// ---------------------------------------------------------------------------
template <int RPL, typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32) terminal_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerCta + wic;
  if (env >= p.n_envs) return;
  if (!p.out.done[env]) return;
  Env<T> e(smem + (size_t)wic * d.sm_stride, d, 32 * RPL);
  uint8_t* gstate = p.state + (size_t)env * d.st_stride;
  load_state(e.sm, gstate, d.st_stride, lane);
  __syncwarp();

  const int n_comp = e.hdr->n_comp, n_pins = e.hdr->n_pins, n_nets = e.hdr->n_nets;
  double reward, wl = 0.0, ni = 0.0;
  const bool route = d.kind >= 2 && e.hdr->cur == n_comp;
  if (route) {
    RouteScratch rs = route_scratch(e.scr, d);
    // get_net_pin_positions (pin:1266-1280): pins are net-major, so a net is a range
    for (int q = lane; q < n_pins; q += 32) {
      PinRec pr = e.pins[q];
      CompRec r = e.comps[pr.comp];
      rs.px[q] = (int16_t)(r.x + pr.rx);
      rs.py[q] = (int16_t)(r.y + pr.ry);
      if (q == 0 || e.pins[q - 1].net != pr.net) rs.net_off[pr.net] = (uint16_t)q;
    }
    if (lane == 0) rs.net_off[n_nets] = (uint16_t)n_pins;
    __syncwarp();
    // get_centroid (pin:1282-1294): exact integer sums, one division per coordinate
    for (int net = lane; net < n_nets; net += 32) {
      int lo = rs.net_off[net], hi = rs.net_off[net + 1], sx = 0, sy = 0;
      for (int q = lo; q < hi; ++q) { sx += rs.px[q]; sy += rs.py[q]; }
      rs.cx[net] = (double)sx / (double)(hi - lo);
      rs.cy[net] = (double)sy / (double)(hi - lo);
    }
    __syncwarp();
    int crossings = 0;
    double length = 0.0;
    if (d.reward_type == 0) {
      centroid_route<T>(d, e, rs, lane, n_pins, n_nets, crossings, length);
    } else {
      beam_paths<T>(d, e, rs, lane, n_nets);
      beam_route<T>(d, e, rs, lane, n_pins, n_nets, crossings, length);
      if (d.reward_type == 2) {  // "both": fewer crossings wins, beam on ties (pin:643-661)
        int c2; double l2;
        centroid_route<T>(d, e, rs, lane, n_pins, n_nets, c2, l2);
        if (c2 < crossings) { crossings = c2; length = l2; }
      }
    }
    wl = length / d.wl_norm;            // pin:921 / 940 / 959-961
    ni = (double)crossings / d.ni_norm; // pin:922-924
    reward = -1.0 * (d.ww * wl + d.wi * ni);
    if (lane == 0) { p.out.reward[env] = reward; p.out.wl[env] = wl; p.out.ni[env] = ni; }
  } else {
    reward = p.out.reward[env];
    if (d.kind >= 2) { wl = d.pen_wl; ni = d.pen_ni; }
  }

  const bool first_report = !(e.hdr->flags & kFlagFinished);
  if (lane == 0) {
    double ret = e.dyn->ep_return + (route ? reward : 0.0);
    p.out.ep_return[env] = ret;
    p.out.ep_length[env] = (int32_t)e.hdr->ep_len;
    if (first_report) {
      double* st = p.stats + (size_t)(env & (kStatSlots - 1)) * 8;
      atomicAdd(st + 0, 1.0);
      atomicAdd(st + 1, ret);
      atomicAdd(st + 2, (double)e.hdr->ep_len);
      if (d.kind >= 2) { atomicAdd(st + 3, wl); atomicAdd(st + 4, ni); }
    }
  }
  __syncwarp();

  if (!d.autoreset) {
    if (lane == 0 && first_report) {
      e.hdr->flags |= kFlagFinished;
      *reinterpret_cast<uint4*>(gstate) = *reinterpret_cast<const uint4*>(e.sm);
    }
    return;
  }
  // auto-reset: next bank slot of this env, first observation of the new episode
  uint32_t slot = e.hdr->slot + (uint32_t)p.n_envs;
  if (p.bank_count) slot %= p.bank_count;
  uint64_t epno = e.dyn->episode_no + 1;
  __syncwarp();
  load_instance<RPL, T>(p, e, slot, epno, lane);
  T occ[RPL], m0[RPL], m1[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) occ[k] = T(0);
  next_masks<RPL, T>(d, e, occ, 0, e.hdr->n_comp, lane, m0, m1);
  copy_out(gstate, e.sm, d.st_stride, lane);
  write_obs<RPL, T>(p, e, env, lane, m0, m1);
}

// K6: reset() -- all envs (reset_mask == nullptr) or the masked ones.
template <int RPL, typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32) reset_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerCta + wic;
  if (env >= p.n_envs) return;
  if (p.reset_mask && !p.reset_mask[env]) return;
  Env<T> e(smem + (size_t)wic * d.sm_stride, d, 32 * RPL);
  uint8_t* gstate = p.state + (size_t)env * d.st_stride;
  uint32_t slot = (uint32_t)env;
  uint64_t epno = 0;
  if (!p.reset_first) {
    const EnvHdr* gh = reinterpret_cast<const EnvHdr*>(gstate);
    const EnvDyn* gd = reinterpret_cast<const EnvDyn*>(gstate + 16);
    slot = gh->slot + (uint32_t)p.n_envs;
    epno = gd->episode_no + 1;
  }
  if (p.bank_count) slot %= p.bank_count;
  zero_smem(e.sm, d.st_stride, lane);
  __syncwarp();
  load_instance<RPL, T>(p, e, slot, epno, lane);
  T occ[RPL], m0[RPL], m1[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) occ[k] = T(0);
  next_masks<RPL, T>(d, e, occ, 0, e.hdr->n_comp, lane, m0, m1);
  copy_out(gstate, e.sm, d.st_stride, lane);
  write_obs<RPL, T>(p, e, env, lane, m0, m1);
  if (lane == 0) {
    p.out.done[env] = 0;
    p.out.reward[env] = 0.0;
  }
}

// ---------------------------------------------------------------------------
// K7: the random policy of agent/random/random_policy_*.py:11-58 -- uniform over
// np.argwhere(action_mask == 1), i.e. over the set bits of the mask bitboards in
// (orientation, row, column) order.  Philox4x32-10 keyed by the seed, counter =
// (env, step).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint32_t c0, uint32_t c1,
                                               uint32_t c2) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = 0x5043422du;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
    uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
    x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return x0;
}

template <int RPL, typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32) sample_kernel(const Params p) {
  const Dims& d = p.d;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerCta + wic;
  if (env >= p.n_envs) return;
  const uint8_t* gstate = p.state + (size_t)env * d.st_stride;
  const EnvHdr* gh = reinterpret_cast<const EnvHdr*>(gstate);
  const T* gocc = reinterpret_cast<const T*>(gstate + d.st_occ);
  T occ[RPL], m[2][RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) { occ[k] = gocc[lane + 32 * k]; m[0][k] = T(0); m[1][k] = T(0); }
  if (d.kind == 0) {
    footprint_mask<RPL, T>(occ, d.sq_n, d.sq_n, d.H, d.W, lane, m[0]);
  } else if (gh->cur < gh->n_comp) {
    CompRec r = reinterpret_cast<const CompRec*>(gstate + d.st_comps)[gh->cur];
    footprint_mask<RPL, T>(occ, r.h, r.w, d.H, d.W, lane, m[0]);
    footprint_mask<RPL, T>(occ, r.w, r.h, d.H, d.W, lane, m[1]);
  }
  // per-lane counts, warp totals per orientation
  int cnt[2];
#pragma unroll
  for (int oo = 0; oo < 2; ++oo) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < RPL; ++k) c += popc_t<T>(m[oo][k]);
    cnt[oo] = c;
  }
  int tot0 = cnt[0], tot1 = cnt[1];
  for (int off = 16; off; off >>= 1) {
    tot0 += __shfl_xor_sync(kFull, tot0, off);
    tot1 += __shfl_xor_sync(kFull, tot1, off);
  }
  const int per_pair = tot0 + tot1;
  const int total = d.O == 4 ? 2 * per_pair : (d.O == 2 ? per_pair : tot0);
  int ao = 0, ax = 0, ay = 0;
  if (total > 0) {
    uint32_t r = philox_u32(p.seed, (uint32_t)env, (uint32_t)p.step_no, (uint32_t)(p.step_no >> 32));
    int kth = (int)__umulhi(r, (uint32_t)total);  // uniform in [0, total)
    int obase = 0;
    if (kth >= per_pair && d.O == 4) { kth -= per_pair; obase = 2; }
    int oo = 0;
    if (kth >= tot0) { kth -= tot0; oo = 1; }
    // rows are visited in order 0..H-1: slot 0 lanes first, then slot 1
    int found_row = -1, found_col = 0;
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      int c = popc_t<T>(oo ? m[1][k] : m[0][k]);
      int incl = c;
      for (int off = 1; off < 32; off <<= 1) {
        int v = __shfl_up_sync(kFull, incl, off);
        if (lane >= off) incl += v;
      }
      int slot_total = __shfl_sync(kFull, incl, 31);
      int excl = incl - c;
      bool mine = found_row < 0 && kth >= excl && kth < incl;
      unsigned who = __ballot_sync(kFull, mine);
      if (who && found_row < 0) {
        int src = __ffs(who) - 1;
        int col = 0;
        if (mine) {
          T bits = oo ? m[1][k] : m[0][k];
          int skip = kth - excl;
          for (int i = 0; i < skip; ++i) bits &= bits - 1;
          col = sizeof(T) == 8 ? __ffsll((long long)bits) - 1 : __ffs((int)bits) - 1;
        }
        found_col = __shfl_sync(kFull, col, src);
        found_row = src + 32 * k;
      }
      if (found_row < 0) kth -= slot_total;
    }
    ao = obase + oo; ax = found_row; ay = found_col;
  }
  if (lane == 0) {
    if (d.kind == 0) { p.actions_out[2 * env] = ax; p.actions_out[2 * env + 1] = ay; }
    else { p.actions_out[3 * env] = ao; p.actions_out[3 * env + 1] = ax; p.actions_out[3 * env + 2] = ay; }
  }
}

}  // namespace pcb
#endif
