// pcb_layout.h -- device-side data layout of libpcbgpu (shared by host and kernels).
//
// HBM layout (DESIGN.md "Data layout"):
//   state  [N][st_stride]   one compact record per environment
//            +0   EnvHdr   (16 B)  counts, current component, bank slot, episode length
//            +16  EnvDyn   (16 B)  episode return (f64), episode number
//            +32  CompRec[C]       (h, w, x, y) u8; x = y = 0xFF until placed
//            +..  PinRec[max_pins] (comp, rx, ry, net, pin_id u16, flags, rank)
//            +..  occupancy bitboard: one word per grid row (u32 if the grid fits
//                 32x32, else u64), 32*RPL words
//            +..  spatial env only: u8 orient[C], the orientation each component was placed
//                 with (component_grid is drawn from the UN-rotated pins: the kernel rotates the
//                 state's pins back instead of re-reading the bank, so a running episode never
//                 depends on what the bank holds later)
//   bank   [count][rec_stride]  EnvHdr + CompRec[C] + PinRec[max_pins] as drawn by
//                 generate_instances (un-rotated pins); reset copies it into state
//   obs    struct-of-arrays by observation key, [R][N][...], env-major inside a key
#ifndef PCB_LAYOUT_H_
#define PCB_LAYOUT_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define PCB_HD __host__ __device__
#else
#define PCB_HD
#endif

// nvcc splits -D values at commas, so a build-time shape arrives as eight macros
#if defined(PCB_US_KIND) && !defined(PCB_USER_SHAPE)
#define PCB_USER_SHAPE PCB_US_KIND, PCB_US_H, PCB_US_W, PCB_US_C, PCB_US_MH, PCB_US_MW, PCB_US_K, PCB_US_PINS
#endif

namespace pcb {

constexpr int kMaxGrid = 64;        // H, W <= 64
constexpr int kMaxComps = 64;       // max_num_components
constexpr int kMaxPins = 512;       // pin table capacity
constexpr int kMaxNets = 64;
constexpr int kMaxBeamPoints = 15;  // pins per net - 1 when the beam route is used
constexpr int kMaxBeamWidth = 4;
#ifndef PCB_WARPS_PER_CTA
#define PCB_WARPS_PER_CTA 4
#endif
#ifndef PCB_MIN_BLOCKS
#define PCB_MIN_BLOCKS 12   // __launch_bounds__ min CTAs/SM of the step kernel (caps registers)
#endif
constexpr int kWarpsPerCta = PCB_WARPS_PER_CTA;
constexpr int kStatSlots = 32;      // spread of the episode-statistics atomics
constexpr int kPairQueue = 64;      // candidate segment pairs buffered per warp

enum : uint8_t { kFlagFinished = 1 };
enum : uint8_t { kPinWinner = 1 };  // last writer of its (component, pin_id) feature row

struct __align__(16) EnvHdr {
  uint8_t n_comp;
  uint8_t n_nets;
  uint16_t n_pins;
  uint8_t cur;      // index of the component to place next; == n_comp when all placed
  uint8_t flags;
  uint16_t pad;
  uint32_t slot;    // bank slot of the running episode
  uint32_t ep_len;  // steps taken in the running episode
};

struct __align__(16) EnvDyn {
  double ep_return;
  uint64_t episode_no;
};

struct CompRec {
  uint8_t h, w, x, y;
};

struct __align__(8) PinRec {
  uint8_t comp, rx, ry, net;
  uint16_t pin_id;
  uint8_t flags;
  uint8_t rank;  // index among the pins of its component, env.pins order
};

// Byte offsets / element counts derived from the configuration.  constexpr so that
// the shape-specialised kernels fold them at compile time and the host computes the
// very same numbers.
struct Layout {
  int O, P, F;
  int st_comps, st_pins, st_occ, st_orient, st_stride;
  int rec_comps, rec_pins, rec_stride;
  int grid_b, mask_b, feat_n, pnum_n, pcat_n, pgrid_b, cgrid_b;
  int sm_mrow, sm_stage, sm_scr, sm_stride;
  int sm_bulk;   // offset of the mask staging of the bulk-copy write-out, 0 = not used
  int pchunk;  // pin env: components whose pin feature rows are staged per pass
};

PCB_HD constexpr int align16c(int v) { return (v + 15) & ~15; }
PCB_HD constexpr int maxc(int a, int b) { return a > b ? a : b; }

// terminal-reward scratch per warp: per-pin arrays + per-net arrays + two pair queues
PCB_HD constexpr int route_scratch_bytes(int max_pins, int K) {
  return max_pins * (8 + 8 + 8 + 8 + 8 + 2 + 2 + 2 + 1 + 1) + K * (8 + 8 + 8 + 2 + 2) + 2 +
         2 * kPairQueue * 4 + 64;
}

// Pin env: the pin feature rows (C * P rows of 10 bytes) are staged in shared memory in passes
// of `pchunk` components, so that a large configuration (E: 5 KB of rows) does not set the
// CTAs per SM; a multiple of 8 components keeps every pass 16-byte aligned in the output.
PCB_HD constexpr int pin_chunk_comps(int C, int P) {
  if (C * P * 10 <= 3584) return C;
  const int by_size = 8 * maxc(1, 2048 / (P * 80));
  return by_size < C ? by_size : C;
}

// PCB_OPT_BULK: the 32x32 pin shapes stage one env's action-mask rows (orientation 0 | 1, 2 KB)
// in shared memory and hand them to the bulk-copy engine (cp.async.bulk.global.shared::cta,
// SASS UBLKCP) -- twice, because orientations 2, 3 repeat 0, 1 -- instead of 8 x 512-byte
// st.global per warp; the pin feature rows leave the same way.
#ifndef PCB_OPT_BULK
#define PCB_OPT_BULK 0
#endif
PCB_HD constexpr bool bulk_shape(int kind, int H, int W, bool fast_board) {
  return PCB_OPT_BULK != 0 && fast_board && kind >= 2 && H == 32 && W == 32;
}

// fast_board: the kernel expands bitboards straight from registers (FixedShape 16/32/64
// square grids) and needs no mask rows in shared memory.
PCB_HD constexpr Layout make_layout(int kind, int H, int W, int C, int mh, int mw, int K,
                                    int max_pins, bool fast_board) {
  Layout L{};
  const bool wide = H > 32 || W > 32;
  const int word = wide ? 8 : 4, rows = wide ? 64 : 32;
  L.O = kind == 0 ? 1 : (kind == 1 ? 2 : 4);
  L.P = mh * mw;
  L.F = kind == 3 ? 5 + L.P : 5;
  L.st_comps = 32;
  L.st_pins = L.st_comps + align16c(C * 4);
  L.st_occ = L.st_pins + align16c(max_pins * 8);
  L.st_orient = L.st_occ + rows * word;
  L.st_stride = align16c(L.st_orient + (kind == 3 ? align16c(C) : 0));
  L.rec_comps = 16;
  L.rec_pins = L.rec_comps + align16c(C * 4);
  L.rec_stride = L.rec_pins + align16c(max_pins * 8);
  L.grid_b = H * W;
  L.mask_b = L.O * H * W;
  L.feat_n = C * L.F;
  L.pnum_n = kind == 2 ? C * L.P * 4 : (kind == 3 ? (C * L.P + 1) * 4 : 0);
  L.pcat_n = kind == 2 ? C * L.P : (kind == 3 ? (C * L.P + 1) * 2 : 0);
  L.pgrid_b = kind == 3 ? H * W * K : 0;
  L.cgrid_b = kind == 3 ? C * mh * mw * K : 0;
  // staging: pin feature rows + component features side by side; the spatial boards reuse it
  L.pchunk = kind == 2 ? pin_chunk_comps(C, L.P) : C;
  const int pins_area = kind == 2 ? L.pchunk * L.P * 10 : (L.pnum_n + L.pcat_n) * 2;
  int stage = align16c(pins_area) + align16c(L.feat_n * 4) + 16;
  // pin_grid is staged in two halves of the rows
  if (kind == 3) stage = maxc(stage, maxc(align16c(((H + 1) / 2) * W * K), L.cgrid_b));
  L.sm_mrow = L.st_stride;
  // generic boards: mask rows, then a dense bit string of one board (H*W bits + slack)
  L.sm_scr = L.sm_mrow + (fast_board ? 0 : 2 * rows * word);
  L.sm_stage = L.sm_scr + (fast_board ? 0 : align16c(4 * ((H * W + 31) / 32 + 3)));
  L.sm_stride = align16c(L.sm_stage + align16c(stage) + 16);
  L.sm_bulk = 0;
  if (bulk_shape(kind, H, W, fast_board)) {
    L.sm_bulk = L.sm_stride;
    L.sm_stride += 2 * H * W;
  }
  return L;
}

// A finished pin-env episode whose components are all placed leaves a routing job for
// reward_kernel: the absolute pin positions (net-major) and the episode bookkeeping.
struct __align__(16) RouteJob {
  uint32_t env;
  uint16_t n_pins;
  uint8_t n_nets;
  uint8_t count_stats;  // 0 when the episode was already reported (autoreset off, re-stepped)
  uint32_t ep_len;
  uint32_t pad;
  double ep_return;     // return before the terminal reward
  double pad2;
};
struct JobPin { uint8_t x, y, net, pad; };
// job record: RouteJob (32 B) | JobPin[max_pins] | beam paths: nxt[max_pins] u16, start[K] u16
PCB_HD constexpr int job_path_offset(int max_pins) { return 32 + align16c(max_pins * 4); }
PCB_HD constexpr int job_stride(int max_pins, int K) {
  return job_path_offset(max_pins) + align16c(max_pins * 2 + K * 2);
}

// Everything a kernel needs to know about the configuration; passed by value.
struct Dims {
  int kind, H, W, C, K, max_pins;
  int mh, mw, sq_n;
  int reward_type, beam_width, norm_fma, autoreset;
  Layout L;
  double ww, wi, wl_norm, ni_norm, pen_reward, pen_wl, pen_ni;
};

// Pointers to the slice of every tensor the current launch writes.
struct Out {
  uint8_t* grid;
  uint8_t* mask;
  float* feat;
  uint8_t* comp_mask;
  uint8_t* place_mask;
  int16_t* pnum;
  int16_t* pcat;
  uint8_t* pgrid;
  uint8_t* cgrid;
  double* reward;
  uint8_t* done;
  double* wl;
  double* ni;
  double* ep_return;
  int32_t* ep_length;
};

struct Params {
  Dims d;
  Out out;
  uint8_t* state;
  const uint8_t* bank;
  const float* ratio_lut;  // [mh*mw + 1]: float(area / (H*W)), host-evaluated in float64
  uint32_t bank_count;
  int n_envs;
  const int32_t* actions;
  int action_format;
  const uint8_t* reset_mask;
  int reset_first;
  uint8_t* jobs;        // [N][job_stride] routing jobs of the current step
  uint32_t* job_count;  // [4] job counters, ring indexed by step number
  // step_kernel leaves its routing jobs at jobs[atomicAdd(job_count + (step_no & job_slot_mask))]
  // with RouteJob::pad = job_t.  Per-step mode: mask 3 (the ring of counters), job_t 0.  List mode
  // (the job list of a multi-step launch, filled by per-step kernels): mask 0, job_t = step of the launch.
  uint32_t job_slot_mask, job_t;
  // the state records do not fit in L2 beside the observation stream (large batches of large
  // boards): load / store them with the streaming policy, so that they leave L2 to the stores.
  // Honoured by the kernels of boards wider than 32 only (RPL == 2): on the others the extra
  // branch cost 0.9 % on config C, and their records reach that size only beyond 256 K envs.
  uint32_t state_stream;
  double* stats;      // [kStatSlots][8]
  uint64_t seed;
  uint64_t step_no;
  int32_t* actions_out;  // != nullptr: also sample the next random action of every env
  // CUDA-graph replays: a captured launch carries the step number of capture time; the steps
  // taken since then (a multiple of the graph length) are added from device memory, so the
  // counter-based random policy keeps advancing.  nullptr outside graphs.
  const uint64_t* step_base;
};

}  // namespace pcb
#endif
