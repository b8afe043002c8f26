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
//   bank   [count][rec_stride]  EnvHdr + CompRec[C] + PinRec[max_pins] as drawn by
//                 generate_instances (un-rotated pins); reset copies it into state
//   obs    struct-of-arrays by observation key, [R][N][...], env-major inside a key
#ifndef PCB_LAYOUT_H_
#define PCB_LAYOUT_H_

#include <stdint.h>

namespace pcb {

constexpr int kMaxGrid = 64;        // H, W <= 64
constexpr int kMaxComps = 64;       // max_num_components
constexpr int kMaxPins = 512;       // pin table capacity
constexpr int kMaxNets = 64;
constexpr int kMaxBeamPoints = 15;  // pins per net - 1 when the beam route is used
constexpr int kMaxBeamWidth = 4;
constexpr int kWarpsPerCta = 4;
constexpr int kStatSlots = 32;      // spread of the episode-statistics atomics

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

// Everything a kernel needs to know about the configuration; passed by value.
struct Dims {
  int kind, H, W, O, C, P, K, max_pins, F;
  int mh, mw, sq_n;
  int reward_type, beam_width, norm_fma, autoreset;
  // byte offsets inside a state record / bank record
  int st_stride, st_comps, st_pins, st_occ;
  int rec_stride, rec_comps, rec_pins;
  // per-env element counts of each observation key
  int grid_b, mask_b, feat_n, pnum_n, pcat_n, pgrid_b, cgrid_b;
  // shared memory per warp (bytes) and offsets inside it
  int sm_stride, sm_mrow, sm_stage, sm_scr;
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
  uint32_t bank_count;
  int n_envs;
  const int32_t* actions;
  int action_format;
  const uint8_t* reset_mask;
  int reset_first;
  double* stats;  // [kStatSlots][8]
  uint64_t seed;
  uint64_t step_no;
  int32_t* actions_out;
};

}  // namespace pcb
#endif
