// pcb_kernels.cuh -- sm_100a kernels of the batched placement environments.
//
// One warp owns one environment.  The occupancy grid is a bitboard, one word per
// grid row, lane r holding row r (and row r+32 for grids taller than 32).  The
// valid-position mask of a (ph, pw) footprint is a sliding-window AND over the
// free bits: log2(pw) shift/ANDs inside the row word, log2(ph) warp shuffles across
// rows.  Nothing here is a dense contraction, so no tensor cores: the kernels are
// bound by the HBM write-out of the observation slab (bit -> byte expansion in
// registers + 16-byte coalesced stores) and by instruction issue, which is why the
// canonical shapes are compiled with every dimension as a constant (FixedShape).
// Compiled with -fmad=false: every float64 product and sum below is individually
// rounded like the reference's Python/numpy arithmetic; the one fused multiply-add
// of the reference (numpy ddot inside linalg.norm) is written explicitly.
#ifndef PCB_KERNELS_CUH_
#define PCB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "pcb_layout.h"
#include "pcb_cross.h"

namespace pcb {

constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------
// shapes: every dimension either a compile-time constant or read from Dims
// ---------------------------------------------------------------------------
#define PCB_LAYOUT_FIELDS(X)                                                                   \
  X(O) X(P) X(F) X(st_comps) X(st_pins) X(st_occ) X(st_orient) X(st_stride) X(rec_comps) X(rec_pins)        \
  X(rec_stride) X(grid_b) X(mask_b) X(feat_n) X(pnum_n) X(pcat_n) X(pgrid_b) X(cgrid_b)        \
  X(sm_mrow) X(sm_stage) X(sm_scr) X(sm_stride) X(sm_bulk) X(pchunk)

template <int KIND, int HH, int WW, int CC, int MH, int MW, int KK, int MAXPINS>
struct FixedShape {
  static constexpr bool fixed = true;
  static constexpr int kH = HH;
  static constexpr bool fast_board = HH == WW && (HH == 16 || HH == 32 || HH == 64);
  static constexpr Layout LL = make_layout(KIND, HH, WW, CC, MH, MW, KK, MAXPINS, fast_board);
  __device__ __forceinline__ explicit FixedShape(const Dims&) {}
  __device__ __forceinline__ constexpr int kind() const { return KIND; }
  __device__ __forceinline__ constexpr int H() const { return HH; }
  __device__ __forceinline__ constexpr int W() const { return WW; }
  __device__ __forceinline__ constexpr int C() const { return CC; }
  __device__ __forceinline__ constexpr int K() const { return KK; }
  __device__ __forceinline__ constexpr int mh() const { return MH; }
  __device__ __forceinline__ constexpr int mw() const { return MW; }
  __device__ __forceinline__ constexpr int max_pins() const { return MAXPINS; }
#define X(f) \
  __device__ __forceinline__ constexpr int f() const { return LL.f; }
  PCB_LAYOUT_FIELDS(X)
#undef X
};

struct DynShape {
  static constexpr bool fixed = false;
  static constexpr int kH = 0;
  const Dims& d;
  __device__ __forceinline__ explicit DynShape(const Dims& dd) : d(dd) {}
  __device__ __forceinline__ int kind() const { return d.kind; }
  __device__ __forceinline__ int H() const { return d.H; }
  __device__ __forceinline__ int W() const { return d.W; }
  __device__ __forceinline__ int C() const { return d.C; }
  __device__ __forceinline__ int K() const { return d.K; }
  __device__ __forceinline__ int mh() const { return d.mh; }
  __device__ __forceinline__ int mw() const { return d.mw; }
  __device__ __forceinline__ int max_pins() const { return d.max_pins; }
#define X(f) \
  __device__ __forceinline__ int f() const { return d.L.f; }
  PCB_LAYOUT_FIELDS(X)
#undef X
};

// the canonical configurations of BASELINE.md section 2
using ShapeA = FixedShape<0, 8, 8, 0, 0, 0, 0, 0>;
using ShapeB = FixedShape<1, 16, 16, 6, 4, 4, 0, 0>;
using ShapeC = FixedShape<2, 32, 32, 10, 4, 4, 6, 30>;
using ShapeD = FixedShape<3, 32, 32, 10, 4, 4, 6, 30>;
using ShapeE = FixedShape<2, 64, 64, 32, 4, 4, 17, 96>;
// one more shape chosen at build time: -DPCB_USER_SHAPE=kind,H,W,C,mh,mw,K,max_pins
// (pcbgpu.specialize builds such a library for any configuration)
#ifdef PCB_USER_SHAPE
using ShapeUser = FixedShape<PCB_USER_SHAPE>;
#endif
// the reference's shipped 10x10 configurations (agent/config/*.json)
using ShapeRefPin = FixedShape<2, 10, 10, 5, 2, 2, 4, 18>;      // rectangle_pin_model.json & co.
using ShapeRefSpatial = FixedShape<3, 10, 10, 5, 2, 2, 4, 18>;  // rectangle_pin_spatial_model.json
using ShapeRefRect = FixedShape<1, 10, 10, 20, 4, 4, 0, 0>;     // rectangle_model.json
using ShapeRefSquare = FixedShape<0, 10, 10, 0, 0, 0, 0, 0>;    // square_model.json

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
__device__ __forceinline__ int popc_t(T v) {
  if constexpr (sizeof(T) == 8) return __popcll((unsigned long long)v);
  else return __popc((unsigned)v);
}

template <typename T>
__device__ __forceinline__ int ffs_t(T v) {  // index of the lowest set bit, v != 0
  if constexpr (sizeof(T) == 8) return __ffsll((long long)v) - 1;
  else return __ffs((int)v) - 1;
}

// position of the (n+1)-th set bit (n counted from 0); garbage when v has fewer bits
__device__ __forceinline__ int nth_bit32(uint32_t v, int n) {
  int pos = 0;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    int c = __popc(v & ((1u << w) - 1u));
    if (n >= c) { n -= c; pos += w; v >>= w; }
  }
  return pos;
}

template <typename T>
__device__ __forceinline__ int nth_bit(T v, int n) {
  if constexpr (sizeof(T) == 8) {
    unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    int cl = __popc(lo);
    return n < cl ? nth_bit32(lo, n) : 32 + nth_bit32(hi, n - cl);
  } else {
    return nth_bit32((unsigned)v, n);
  }
}

// Observation stores are streaming (st.global.cs, evict-first in L2): the slab is written
// once and never re-read by these kernels, while the per-env state and the instance bank
// ARE re-read every step and should stay L2-resident.  Measured on the bare store pattern
// of config C: 85 us with default stores, 75 us with .cs (profiles/README.md).
template <typename V>
__device__ __forceinline__ void st_stream(V* p, V v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint8_t* p, uint8_t v) {
  __stcs(reinterpret_cast<unsigned char*>(p), (unsigned char)v);
}
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) {
  __stcs(reinterpret_cast<unsigned int*>(p), (unsigned int)v);
}

// smem -> global copy of one environment's slice of a tensor, widest store the
// alignment allows (16 B when the per-env size is a multiple of 16, else 8 or 4; the staging
// buffers are 16-byte aligned).
__device__ __forceinline__ void copy_out(void* dst, const void* src, int nbytes, int lane) {
  uintptr_t a = reinterpret_cast<uintptr_t>(dst) | (uintptr_t)nbytes;
  if ((a & 15) == 0) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = lane; i < (nbytes >> 4); i += 32) st_stream(d + i, s[i]);
  } else if ((a & 7) == 0) {
    const uint2* s = reinterpret_cast<const uint2*>(src);
    uint2* d = reinterpret_cast<uint2*>(dst);
    for (int i = lane; i < (nbytes >> 3); i += 32) st_stream(d + i, s[i]);
  } else if ((a & 3) == 0) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (int i = lane; i < (nbytes >> 2); i += 32) st_stream(d + i, s[i]);
  } else {
    const uint8_t* s = reinterpret_cast<const uint8_t*>(src);
    uint8_t* d = reinterpret_cast<uint8_t*>(dst);
    for (int i = lane; i < nbytes; i += 32) st_stream(d + i, s[i]);
  }
}

__device__ __forceinline__ void zero_smem(void* p, int nbytes, int lane) {  // nbytes % 16 == 0
  uint4* d = reinterpret_cast<uint4*>(p);
  for (int i = lane; i < (nbytes >> 4); i += 32) d[i] = make_uint4(0, 0, 0, 0);
}

__device__ __forceinline__ void copy16(void* dst, const void* src, int nbytes, int lane) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (int i = lane; i < (nbytes >> 4); i += 32) d[i] = s[i];
}

// ---------------------------------------------------------------------------
// Bulk-copy write-out (sm_90+ bulk async copies, no tensor map): one thread hands a
// contiguous, 16-byte aligned run of shared memory to the copy engine, which writes it to
// global memory without occupying load/store slots of the SM.  The stores carry the same
// evict-first L2 policy as st.global.cs (st_stream above).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy
__device__ __forceinline__ void fence_smem_for_bulk() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t nbytes, uint64_t pol) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               :: "l"(gdst), "r"(s), "r"(nbytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the copy engine has READ every committed source buffer (they may be reused / the CTA may exit)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------------------
// The per-env state records are the one thing every step re-reads: 29 MB at 65 536 envs of
// config C, written by step t and read by step t+1 ~90 us later, after ~500 MB of observation
// bytes have streamed through the 126 MB L2; ncu shows 33 MB of DRAM reads per launch, i.e.
// every record comes back from DRAM, and the first use of the loaded record is the top stall.
// PCB_OPT_STATE_L2=1 gives their loads and stores an L2 evict_last policy.  Measured (round 2,
// profiles/README.md): step kernel alone 86.2 -> 89.0 us, whole step 0.819 -> 0.790 -- slower,
// so it is off; the code stays as the record of the experiment.
// ---------------------------------------------------------------------------
#ifndef PCB_OPT_STATE_L2
#define PCB_OPT_STATE_L2 0
#endif
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ld_state(const void* p, bool stream) {
  // stream: evict-first like the observation stores (Params::state_stream)
  if (PCB_OPT_STATE_L2 == 2 || PCB_OPT_STATE_L2 == 3 || stream) return __ldcs(reinterpret_cast<const uint4*>(p));  // (3 / 4: loads / stores only, experiments)
#if PCB_OPT_STATE_L2 == 1
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(l2_evict_last_policy()));
  return v;
#else
  return *reinterpret_cast<const uint4*>(p);
#endif
}
__device__ __forceinline__ void st_state(void* p, uint4 v, bool stream) {
  if (PCB_OPT_STATE_L2 == 2 || PCB_OPT_STATE_L2 == 4 || stream) { __stcs(reinterpret_cast<uint4*>(p), v); return; }
#if PCB_OPT_STATE_L2 == 1
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(l2_evict_last_policy()) : "memory");
#else
  *reinterpret_cast<uint4*>(p) = v;
#endif
}
// shared -> global copy of a state record
__device__ __forceinline__ void store_state(void* gdst, const void* ssrc, int nbytes, int lane, bool stream) {
  const uint4* s4 = reinterpret_cast<const uint4*>(ssrc);
  uint4* d = reinterpret_cast<uint4*>(gdst);
  for (int i = lane; i < (nbytes >> 4); i += 32) st_state(d + i, s4[i], stream);
}

// ---------------------------------------------------------------------------
// bitboard: footprint mask (reference: compute_action_mask_orientation,
// dummy_env_rectangular.py:526-567 / dummy_env_rectangular_pin.py:1808-1851;
// square: dummy_env_square.py:74-113, 187-244)
//   out[r] bit j = 1  iff  r + ph <= H, j + pw <= W and rows r..r+ph-1,
//   columns j..j+pw-1 are all free.
// ---------------------------------------------------------------------------
// rows shifted up by dlt: result[r] = v[r + dlt], zero past the last row
template <int RPL, typename T>
__device__ __forceinline__ void rows_ahead(const T (&v)[RPL], int dlt, int lane, T (&out)[RPL]) {
  if constexpr (RPL == 1) {
    T t = __shfl_down_sync(kFull, v[0], dlt);
    out[0] = lane + dlt < 32 ? t : T(0);
  } else {
    int src = (lane + dlt) & 31, hop = (lane + dlt) >> 5;
    T got[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) got[k] = __shfl_sync(kFull, v[k], src);
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      T r = T(0);
#pragma unroll
      for (int kk = 0; kk < RPL; ++kk)
        if (k + hop == kk) r = got[kk];
      out[k] = r;
    }
  }
}

template <int RPL, typename T>
__device__ __forceinline__ void footprint_mask(const T (&occ)[RPL], int ph, int pw, int H, int W,
                                               int lane, T (&out)[RPL]) {
  // horizontal: AND of v >> 0 .. v >> (pw-1); zeros shift in above column W-1
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    int row = lane + 32 * k;
    T v = row < H ? (T)(~occ[k] & low_mask<T>(W)) : T(0);
    if (pw <= 4) {  // the usual component sizes: at most two shift/ANDs
      T a = v & (v >> 1);
      T r = pw >= 2 ? a : v;
      if (pw == 3) r = a & (v >> 2);
      if (pw == 4) r = a & (a >> 2);
      v = r;
    } else {
      int s = 1;
      while (2 * s <= pw) { v &= v >> s; s <<= 1; }
      if (s < pw) v &= v >> (pw - s);
    }
    out[k] = v;
  }
  // vertical: AND of rows r .. r+ph-1 (rows past the board count as blocked)
  T t[RPL];
  if (ph <= 4) {
    if (ph >= 2) {
      T a[RPL];
      rows_ahead<RPL, T>(out, 1, lane, t);
#pragma unroll
      for (int k = 0; k < RPL; ++k) a[k] = out[k] & t[k];
      if (ph >= 3) {
        T src[RPL];
#pragma unroll
        for (int k = 0; k < RPL; ++k) src[k] = ph == 4 ? a[k] : out[k];
        rows_ahead<RPL, T>(src, 2, lane, t);
#pragma unroll
        for (int k = 0; k < RPL; ++k) a[k] &= t[k];
      }
#pragma unroll
      for (int k = 0; k < RPL; ++k) out[k] = a[k];
    }
  } else {
    int s = 1;
    while (2 * s <= ph) {
      rows_ahead<RPL, T>(out, s, lane, t);
#pragma unroll
      for (int k = 0; k < RPL; ++k) out[k] &= t[k];
      s <<= 1;
    }
    if (s < ph) {
      rows_ahead<RPL, T>(out, ph - s, lane, t);
#pragma unroll
      for (int k = 0; k < RPL; ++k) out[k] &= t[k];
    }
  }
}

// ---------------------------------------------------------------------------
// bit rows -> H*W bytes of 0/1 in global memory.  dup_off != 0 writes a second copy
// at dst + dup_off (orientations 2,3 of the pin env repeat 0,1: pin:1868-1869).
// Fixed 32x32 / 64x64 / 16x16 boards go register -> shuffle -> 16-byte store, every
// store instruction covering 512 contiguous bytes; other shapes read the rows from
// shared memory.
// ---------------------------------------------------------------------------
template <typename S>
struct BoardPath {
  static constexpr bool fast = false;
};
template <int KIND, int HH, int WW, int CC, int MH, int MW, int KK, int MP>
struct BoardPath<FixedShape<KIND, HH, WW, CC, MH, MW, KK, MP>> {
  static constexpr bool fast = FixedShape<KIND, HH, WW, CC, MH, MW, KK, MP>::fast_board;
};

#ifndef PCB_OPT_LUT
#define PCB_OPT_LUT 1
#endif
constexpr int kLutBytes = PCB_OPT_LUT ? 256 * 8 : 0;  // CTA-wide byte-expansion table

// 8 bits -> 8 bytes of 0/1 with one shared-memory load.  The 2 KB table lives in global
// memory (static initialiser below, L2-resident) and every CTA copies it with one
// 16-byte load + store per thread.
#define PCB_S4(n) (((n) & 1u) | (((n) & 2u) << 7) | (((n) & 4u) << 14) | (((n) & 8u) << 21))
#define PCB_E1(n) {PCB_S4((n) & 15u), PCB_S4((n) >> 4)}
#define PCB_E4(n) PCB_E1(n), PCB_E1((n) + 1), PCB_E1((n) + 2), PCB_E1((n) + 3)
#define PCB_E16(n) PCB_E4(n), PCB_E4((n) + 4), PCB_E4((n) + 8), PCB_E4((n) + 12)
#define PCB_E64(n) PCB_E16(n), PCB_E16((n) + 16), PCB_E16((n) + 32), PCB_E16((n) + 48)
__device__ const uint2 g_expand_lut[256] = {PCB_E64(0u), PCB_E64(64u), PCB_E64(128u), PCB_E64(192u)};

// phase 1 (before the CTA's other global loads are consumed): fetch; phase 2: store + barrier
__device__ __forceinline__ uint4 expand_lut_fetch() {
#if PCB_OPT_LUT
  if (threadIdx.x < 128) return reinterpret_cast<const uint4*>(g_expand_lut)[threadIdx.x];
#endif
  return make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void expand_lut_store(uint8_t* smem, uint4 v) {
#if PCB_OPT_LUT
  if (blockDim.x >= 128) {
    if (threadIdx.x < 128) reinterpret_cast<uint4*>(smem)[threadIdx.x] = v;
  } else {  // small CTAs: every thread copies several chunks
    for (int i = threadIdx.x; i < 128; i += blockDim.x)
      reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(g_expand_lut)[i];
  }
  __syncthreads();
#endif
}
__device__ __forceinline__ void build_expand_lut(uint8_t* smem) {
  expand_lut_store(smem, expand_lut_fetch());
}

__device__ __forceinline__ uint4 expand16(const uint2* lut, uint32_t b) {
#if PCB_OPT_LUT
  uint2 lo = lut[b & 0xffu], hi = lut[(b >> 8) & 0xffu];
  return make_uint4(lo.x, lo.y, hi.x, hi.y);
#else
  return spread16(b & 0xffffu);
#endif
}

template <int RPL, typename T, typename S>
__device__ __forceinline__ void write_board(const S& sh, uint8_t* dst, const T (&regs)[RPL],
                                            const T* rows, uint32_t* flat, int lane, int dup_off,
                                            const uint2* lut) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  uint4* e4 = reinterpret_cast<uint4*>(dst + dup_off);
  if constexpr (BoardPath<S>::fast) {
    if constexpr (sizeof(T) == 4 && RPL == 1) {
      if (sh.W() == 32) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t w = __shfl_sync(kFull, regs[0], (lane >> 1) + 16 * half);
          uint4 v = expand16(lut, w >> (16 * (lane & 1)));
          st_stream(d4 + half * 32 + lane, v);
          if (dup_off) st_stream(e4 + half * 32 + lane, v);
        }
      } else {  // 16x16: one 16-byte chunk per row
        if (lane < 16) {
          uint4 v = expand16(lut, (uint32_t)regs[0]);
          st_stream(d4 + lane, v);
          if (dup_off) st_stream(e4 + lane, v);
        }
      }
    } else {  // 64x64: 4 chunks per row, 256 chunks
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        T w = __shfl_sync(kFull, regs[i >> 2], (i & 3) * 8 + (lane >> 2));
        uint4 v = expand16(lut, (uint32_t)(w >> (16 * (lane & 3))));
        st_stream(d4 + i * 32 + lane, v);
        if (dup_off) st_stream(e4 + i * 32 + lane, v);
      }
    }
  } else {
    const int H = sh.H(), W = sh.W();
    if ((W & 15) == 0) {
      const int cpr = W >> 4;  // 16-byte chunks per row: 1..4
      const int n = H * cpr;
      for (int q = lane; q < n; q += 32) {
        int row, part;
        if (cpr == 2) { row = q >> 1; part = q & 1; }
        else if (cpr == 4) { row = q >> 2; part = q & 3; }
        else if (cpr == 1) { row = q; part = 0; }
        else { row = q / 3; part = q - 3 * row; }
        uint4 v = spread16((uint32_t)(rows[row] >> (16 * part)) & 0xffffu);
        st_stream(d4 + q, v);
        if (dup_off) st_stream(e4 + q, v);
      }
    } else if (((H * W) & 3) == 0) {
      // Any other width: pack the board into one dense bit string (bit row*W + col) in
      // shared memory -- each row ORs its bits into the one or two words it overlaps --
      // then every lane turns 4 (or 16) consecutive bits into bytes: no per-byte index math.
      const int nbits = H * W, nwords = (nbits + 31) >> 5;
      for (int i = lane; i < nwords + 2; i += 32) flat[i] = 0u;
      __syncwarp();
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const int row = lane + 32 * k;
        if (row < H && regs[k] != T(0)) {
          const int pos = row * W, lo = pos >> 5, sh_ = pos & 31;
          const uint64_t v = (uint64_t)regs[k];
          const uint64_t a = v << sh_;
          atomicOr(flat + lo, (uint32_t)a);
          if (a >> 32) atomicOr(flat + lo + 1, (uint32_t)(a >> 32));
          if (sizeof(T) == 8 && sh_ && (v >> (64 - sh_))) atomicOr(flat + lo + 2, (uint32_t)(v >> (64 - sh_)));
        }
      }
      __syncwarp();
      if ((nbits & 15) == 0) {
        for (int q = lane; q < (nbits >> 4); q += 32) {
          uint4 v = expand16(lut, flat[q >> 1] >> (16 * (q & 1)));
          st_stream(d4 + q, v);
          if (dup_off) st_stream(e4 + q, v);
        }
      } else {
        uint32_t* d1 = reinterpret_cast<uint32_t*>(dst);
        uint32_t* e1 = reinterpret_cast<uint32_t*>(dst + dup_off);
        for (int q = lane; q < (nbits >> 2); q += 32) {
          uint32_t v = spread4((flat[q >> 3] >> (4 * (q & 7))) & 15u);
          st_stream(d1 + q, v);
          if (dup_off) st_stream(e1 + q, v);
        }
      }
      __syncwarp();
    } else {
      // odd sizes: byte b = row * W + col; lane handles b = lane, lane + 32, ...: (row, col)
      // advance by (32 / W, 32 % W) with a carry, so only the first position needs a division
      const int n = H * W;
      const int q32 = 32 / W, r32 = 32 - q32 * W;
      int row = lane / W, col = lane - row * W;
      for (int b = lane; b < n; b += 32) {
        uint8_t v = (uint8_t)((rows[row] >> col) & T(1));
        st_stream(dst + b, v);
        if (dup_off) st_stream(dst + b + dup_off, v);
        row += q32; col += r32;
        if (col >= W) { col -= W; row += 1; }
      }
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
  uint8_t* orient;  // [C] orientation of each placed component (spatial env only)
  T* mrow;     // [2][32*RPL] mask rows of orientation 0 / 1 (generic board path only)
  uint8_t* stage;
  uint8_t* scr;
  template <typename S>
  __device__ __forceinline__ Env(uint8_t* base, const S& sh)
      : sm(base),
        hdr(reinterpret_cast<EnvHdr*>(base)),
        dyn(reinterpret_cast<EnvDyn*>(base + 16)),
        comps(reinterpret_cast<CompRec*>(base + sh.st_comps())),
        pins(reinterpret_cast<PinRec*>(base + sh.st_pins())),
        occ(reinterpret_cast<T*>(base + sh.st_occ())),
        orient(base + sh.st_orient()),
        mrow(reinterpret_cast<T*>(base + sh.sm_mrow())),
        stage(base + sh.sm_stage()),
        scr(base + sh.sm_scr()) {}
};

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

#ifndef PCB_OPT_RNG
#define PCB_OPT_RNG 1
#endif
// 1: SplitMix64 finaliser of the 64-bit counter (seed, env, step) -- ~3x fewer
// instructions than Philox4x32-10 (0), same counter-based, stateless construction.
__device__ __forceinline__ uint64_t action_rand64(uint64_t seed, uint32_t env, uint64_t step_no) {
  uint64_t z = seed + ((step_no << 32) | (uint64_t)env) * 0x9E3779B97F4A7C15ull;
  z ^= step_no >> 32;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t action_rand(uint64_t seed, uint32_t env, uint64_t step_no) {
#if PCB_OPT_RNG
  return (uint32_t)(action_rand64(seed, env, step_no) >> 32);
#else
  return philox_u32(seed, env, (uint32_t)step_no, (uint32_t)(step_no >> 32));
#endif
}

#ifndef PCB_SAMPLE_TRIES
#define PCB_SAMPLE_TRIES 6
#endif
// Uniform draw over the valid (orientation, row, column) triples.  Fast path (pin kinds):
// rejection sampling -- a uniform cell of the O x H x W box is accepted when its mask bit is
// set (exactly uniform over the valid cells; on the large, sparsely filled boards one or two
// tries suffice).  After PCB_SAMPLE_TRIES misses, and always on the small pin-less boards
// (where the multi-env-per-warp rollout kernel wants a loop-free sampler, pcb_small.cuh),
// the exact method runs: k-th set bit in np.argwhere order via a warp prefix sum.
template <int RPL, typename T, typename S>
__device__ __forceinline__ void sample_action(const S& sh, const T (&m0)[RPL], const T (&m1)[RPL],
                                              uint64_t seed, int env, uint64_t step_no, int lane,
                                              int32_t* out) {
  if (sh.kind() >= 2) {
    T anyv = T(0);
#pragma unroll
    for (int k = 0; k < RPL; ++k) anyv |= m0[k] | m1[k];
    if (__any_sync(kFull, anyv != T(0))) {
      const int H = sh.H(), W = sh.W();
      for (int t = 0; t < PCB_SAMPLE_TRIES; ++t) {
        // three independent 21-bit uniforms -> (orientation, row, column); no division
        const uint64_t r = action_rand64(seed + 0x632BE59BD9B4E019ull * (uint64_t)(t + 1), (uint32_t)env, step_no);
        const int o = (int)((((uint32_t)r & 0x1fffffu) * (uint32_t)sh.O()) >> 21);
        const int row = (int)((((uint32_t)(r >> 21) & 0x1fffffu) * (uint32_t)H) >> 21);
        const int col = (int)((((uint32_t)(r >> 42) & 0x1fffffu) * (uint32_t)W) >> 21);
        T mine = T(0);
#pragma unroll
        for (int k = 0; k < RPL; ++k)
          if ((row >> 5) == k) mine = (o & 1) ? m1[k] : m0[k];
        const T bits = __shfl_sync(kFull, mine, row & 31);
        if ((bits >> col) & T(1)) {  // warp-uniform
          if (lane == 0) {
            if (sh.kind() == 0) { out[2 * env] = row; out[2 * env + 1] = col; }
            else { out[3 * env] = o; out[3 * env + 1] = row; out[3 * env + 2] = col; }
          }
          return;
        }
      }
    }
  }
  int c0 = 0, c1 = 0;
#pragma unroll
  for (int k = 0; k < RPL; ++k) { c0 += popc_t<T>(m0[k]); c1 += popc_t<T>(m1[k]); }
  const int tot0 = __reduce_add_sync(kFull, c0), tot1 = __reduce_add_sync(kFull, c1);
  const int per_pair = tot0 + tot1;
  const int O = sh.O();
  const int total = O == 4 ? 2 * per_pair : (O == 2 ? per_pair : tot0);
  int ao = 0, ax = 0, ay = 0;
  if (total > 0) {
    uint32_t r = action_rand(seed, (uint32_t)env, step_no);
    int kth = (int)__umulhi(r, (uint32_t)total);  // uniform in [0, total)
    int obase = 0;
    if (O == 4 && kth >= per_pair) { kth -= per_pair; obase = 2; }
    const bool second = kth >= tot0;
    if (second) kth -= tot0;
    int found_row = -1, found_col = 0;
#pragma unroll
    for (int k = 0; k < RPL; ++k) {  // rows in order: slot 0 lanes 0..31, then slot 1
      const T bits = second ? m1[k] : m0[k];
      const int c = popc_t<T>(bits);
      int incl = c;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        int v = __shfl_up_sync(kFull, incl, off);
        if (lane >= off) incl += v;
      }
      const int slot_total = __shfl_sync(kFull, incl, 31);
      const int excl = incl - c;
      const bool mine = found_row < 0 && kth >= excl && kth < incl;
      const unsigned who = __ballot_sync(kFull, mine);
      const int col = nth_bit<T>(bits, kth - excl);
      if (who && found_row < 0) {
        const int src = __ffs(who) - 1;
        found_col = __shfl_sync(kFull, col, src);
        found_row = src + 32 * k;
      }
      if (found_row < 0) kth -= slot_total;
    }
    ao = obase + (second ? 1 : 0); ax = found_row; ay = found_col;
  }
  if (lane == 0) {
    if (sh.kind() == 0) { out[2 * env] = ax; out[2 * env + 1] = ay; }
    else { out[3 * env] = ao; out[3 * env + 1] = ax; out[3 * env + 2] = ay; }
  }
}

// area / (H*W) as float32(float64 quotient).  With H*W a power of two the quotient is
// exact, so a float multiply gives the same bits; otherwise the host-built table is read.
template <typename S>
__device__ __forceinline__ float area_ratio(const S& sh, const Params& p, int area) {
  const int hw = sh.H() * sh.W();
  if constexpr (S::fixed) {
    if ((hw & (hw - 1)) == 0) return (float)area * (1.0f / (float)hw);
  }
  return p.ratio_lut[area];
}

// step number the random policy draws with (see Params::step_base)
__device__ __forceinline__ uint64_t policy_step(const Params& p) {
  return p.step_no + (p.step_base ? *p.step_base : 0ull);
}

// last node of a captured rollout graph: the next replay is `n` steps further on
__global__ void advance_step_base_kernel(uint64_t* step_base, uint64_t n) { *step_base += n; }
__global__ void set_step_base_kernel(uint64_t* step_base, uint64_t v) { *step_base = v; }

// ---------------------------------------------------------------------------
// observation write-out (everything the reference copies into the obs dict:
// rect:410-416, pin:1680-1687, spatial:1633-1644; layout SURVEY.md Appendix C)
// ---------------------------------------------------------------------------
template <int RPL, typename T, typename S>
__device__ __forceinline__ void write_obs(const S& sh, const Params& p, Env<T>& e, int env,
                                          int lane, const T (&occ)[RPL], const T (&m0)[RPL],
                                          const T (&m1)[RPL]) {
  constexpr int rows = 32 * RPL;
  if constexpr (!BoardPath<S>::fast) {
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      e.occ[lane + 32 * k] = occ[k];
      e.mrow[lane + 32 * k] = m0[k];
      e.mrow[rows + lane + 32 * k] = m1[k];
    }
    __syncwarp();
  }
  const int HW = sh.H() * sh.W();
  const int O = sh.O();
  extern __shared__ __align__(16) uint8_t smem_base[];
  const uint2* lut = reinterpret_cast<const uint2*>(smem_base);
  uint32_t* flat = reinterpret_cast<uint32_t*>(e.scr);
  write_board<RPL, T, S>(sh, p.out.grid + (size_t)env * sh.grid_b(), occ, e.occ, flat, lane, 0, lut);
  uint8_t* mk = p.out.mask + (size_t)env * sh.mask_b();
  bool bulk = false;
  uint64_t pol = 0;
  if constexpr (S::fixed && RPL == 1 && sizeof(T) == 4) {
    if (sh.sm_bulk() != 0) {
      // 32x32 pin shapes: mask rows of orientation 0 | 1 -> 2 KB of bytes in shared memory,
      // then the copy engine writes them to orientations 0,1 and again to 2,3 (pin:1868-1869)
      bulk = true;
      pol = l2_evict_first_policy();
      uint4* sb = reinterpret_cast<uint4*>(e.sm + sh.sm_bulk());
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int src = (lane >> 1) + 16 * half, sft = 16 * (lane & 1);
        const uint32_t w0 = __shfl_sync(kFull, (uint32_t)m0[0], src);
        const uint32_t w1 = __shfl_sync(kFull, (uint32_t)m1[0], src);
        sb[half * 32 + lane] = expand16(lut, w0 >> sft);
        sb[64 + half * 32 + lane] = expand16(lut, w1 >> sft);
      }
      fence_smem_for_bulk();
      __syncwarp();
      if (lane == 0) {
        bulk_store(mk, sb, 2 * 1024, pol);
        bulk_store(mk + 2 * 1024, sb, 2 * 1024, pol);
      }
    }
  }
  if (!bulk) {
    write_board<RPL, T, S>(sh, mk, m0, e.mrow, flat, lane, O == 4 ? 2 * HW : 0, lut);
    if (O >= 2)
      write_board<RPL, T, S>(sh, mk + HW, m1, e.mrow + rows, flat, lane, O == 4 ? 2 * HW : 0, lut);
  }
  const int kind = sh.kind();
  if (kind == 0) return;

  const int C = sh.C(), F = sh.F();
  const int n_comp = e.hdr->n_comp, cur = e.hdr->cur, n_pins = e.hdr->n_pins;

  // Small tensors (component features, pin feature rows) are assembled in shared memory
  // and leave the SM as full coalesced chunks: scattered 2..8-byte stores each occupy a
  // request slot on the SM -> L2 path, which is the busiest unit of this kernel.
  const int pnum_n = sh.pnum_n(), pcat_n = sh.pcat_n();
  // pin env: rows of `pchunk` components per pass (all of them unless the configuration is large)
  const int pchunk = sh.pchunk();
  const int pins_bytes = kind == 2 ? ((pchunk * sh.P() * 10 + 15) & ~15) : (((pnum_n + pcat_n) * 2 + 15) & ~15);
  int16_t* snum = reinterpret_cast<int16_t*>(e.stage);
  int16_t* scat = snum + (kind == 2 ? pchunk * sh.P() * 4 : pnum_n);
  float* fdst = reinterpret_cast<float*>(e.stage + pins_bytes);
  float* feat = p.out.feat + (size_t)env * sh.feat_n();

  // phase A: zero the pin area; all_components_feature = [h, w, x|-1, y|-1, h*w/(H*W)]
  // (+ pin ids of the component padded with -1, spatial)
  if (kind == 3) zero_smem(e.stage, pins_bytes, lane);
  if (kind != 3) {
    for (int c = lane; c < C; c += 32) {  // one lane per component row (5 floats)
      float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f;
      if (c < n_comp) {
        CompRec r = e.comps[c];
        f0 = (float)r.h; f1 = (float)r.w;
        f2 = r.x == 0xFF ? -1.f : (float)r.x;
        f3 = r.y == 0xFF ? -1.f : (float)r.y;
        f4 = area_ratio(sh, p, r.h * r.w);
      }
      float* row = fdst + c * 5;
      row[0] = f0; row[1] = f1; row[2] = f2; row[3] = f3; row[4] = f4;
    }
  } else {
    // spatial (spatial:203-239): the 5 values, then the component's pin ids (written by the
    // pins in phase B) padded with -1; one lane per component row, F = 5 + P floats
    for (int c = lane; c < C; c += 32) {
      float* row = fdst + c * F;
      const bool have = c < n_comp;
      float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f;
      if (have) {
        CompRec r = e.comps[c];
        f0 = (float)r.h; f1 = (float)r.w;
        f2 = r.x == 0xFF ? -1.f : (float)r.x;
        f3 = r.y == 0xFF ? -1.f : (float)r.y;
        f4 = area_ratio(sh, p, r.h * r.w);
      }
      row[0] = f0; row[1] = f1; row[2] = f2; row[3] = f3; row[4] = f4;
      const float pad = have ? -1.f : 0.f;
      for (int f = 5; f < F; ++f) row[f] = pad;
    }
  }

  // placement_mask / component_mask (consecutive bytes: already one request)
  for (int c = lane; c < C; c += 32) {
    if (kind == 1) {
      st_stream(p.out.place_mask + (size_t)env * C + c, (uint8_t)(c < cur ? 1 : 0));    // rect:284
      st_stream(p.out.comp_mask + (size_t)env * C + c, (uint8_t)(c < n_comp ? 1 : 0));  // rect:296-298
    } else {
      uint8_t v = c >= n_comp ? 0 : (c < cur ? 2 : (c == cur ? 3 : 1));  // pin:1577-1579, 1649-1657
      st_stream(p.out.place_mask + (size_t)env * C + c, v);
    }
  }
  __syncwarp();

  // phase B: every winning pin writes its feature row (pin:1521-1542, spatial:1469-1485);
  // the spatial env also lists each component's pin ids in its feature row
  if (kind == 2) {
    const int P = sh.P();
    copy_out(feat, fdst, sh.feat_n() * 4, lane);
    for (int c0 = 0; c0 < C; c0 += pchunk) {
      const int nc = min(pchunk, C - c0), rows = nc * P;
      zero_smem(e.stage, pins_bytes, lane);
      __syncwarp();
      for (int q = lane; q < n_pins; q += 32) {
        PinRec pr = e.pins[q];
        const int c = pr.comp - c0;
        if (!(pr.flags & kPinWinner) || (unsigned)c >= (unsigned)nc) continue;
        CompRec r = e.comps[pr.comp];
        int ax = r.x == 0xFF ? -1 : r.x + pr.rx;
        int ay = r.x == 0xFF ? -1 : r.y + pr.ry;
        const int row = c * P + pr.pin_id;
        // [rel_x, rel_y, abs_x, abs_y] as int16
        uint2 v = make_uint2((uint32_t)pr.rx | ((uint32_t)pr.ry << 16),
                             ((uint32_t)ax & 0xffffu) | ((uint32_t)ay << 16));
        *reinterpret_cast<uint2*>(snum + row * 4) = v;
        scat[row] = pr.net;
      }
      if (bulk && pchunk == C && ((rows * 2) & 15) == 0) {
        // the whole slice of both pin tensors in one pass: two bulk copies instead of ~5
        // shared-memory loads + 16-byte stores per lane
        fence_smem_for_bulk();
        __syncwarp();
        if (lane == 0) {
          bulk_store(p.out.pnum + (size_t)env * pnum_n, snum, rows * 8, pol);
          bulk_store(p.out.pcat + (size_t)env * pcat_n, scat, rows * 2, pol);
        }
      } else {
        __syncwarp();
        copy_out(p.out.pnum + (size_t)env * pnum_n + c0 * P * 4, snum, rows * 8, lane);
        copy_out(p.out.pcat + (size_t)env * pcat_n + c0 * P, scat, rows * 2, lane);
        __syncwarp();
      }
    }
    if (bulk && lane == 0) bulk_commit();
    return;
  }
  if (kind == 3) {
    for (int q = lane; q < n_pins; q += 32) {
      PinRec pr = e.pins[q];
      fdst[pr.comp * F + 5 + pr.rank] = (float)pr.pin_id;
      if (!(pr.flags & kPinWinner)) continue;
      CompRec r = e.comps[pr.comp];
      int ax = r.x == 0xFF ? -1 : r.x + pr.rx;
      int ay = r.x == 0xFF ? -1 : r.y + pr.ry;
      const int row = pr.pin_id;
      uint2 v = make_uint2((uint32_t)pr.rx | ((uint32_t)pr.ry << 16),
                           ((uint32_t)ax & 0xffffu) | ((uint32_t)ay << 16));
      *reinterpret_cast<uint2*>(snum + row * 4) = v;
      *reinterpret_cast<uint32_t*>(scat + row * 2) = (uint32_t)pr.net | ((uint32_t)pr.comp << 16);
    }
    if (lane == 0) { scat[pcat_n - 2] = -1; scat[pcat_n - 1] = -1; }
    __syncwarp();
  }

  // phase C: coalesced write-out
  copy_out(feat, fdst, sh.feat_n() * 4, lane);
  if (kind == 1) return;
  copy_out(p.out.pnum + (size_t)env * pnum_n, snum, pnum_n * 2, lane);
  copy_out(p.out.pcat + (size_t)env * pcat_n, scat, pcat_n * 2, lane);
  if (kind != 3) return;
  __syncwarp();

  // pin_grid (spatial:1663-1675): channel 0 = occupied cell without a pin,
  // channel net+1 = placed pin of that net.  Staged and written in two halves of the rows,
  // which halves the largest shared-memory buffer of the kernel (CTAs per SM: 7 -> 12 on D).
  const int K = sh.K(), W = sh.W();
  const int part_rows = (sh.H() + 1) >> 1;
  for (int part = 0; part < 2; ++part) {
    const int r0 = part * part_rows, r1 = min(sh.H(), r0 + part_rows);
    const int nbytes = (r1 - r0) * W * K;
    zero_smem(e.stage, (nbytes + 15) & ~15, lane);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      const int row = lane + 32 * k;
      T bits = (row >= r0 && row < r1) ? occ[k] : T(0);
      while (bits) {
        int j = ffs_t<T>(bits);
        bits &= bits - 1;
        e.stage[((row - r0) * W + j) * K] = 1;
      }
    }
    __syncwarp();
    for (int q = lane; q < n_pins; q += 32) {
      PinRec pr = e.pins[q];
      CompRec r = e.comps[pr.comp];
      const int ax = r.x + pr.rx;
      if (r.x == 0xFF || ax < r0 || ax >= r1) continue;
      uint8_t* cell = e.stage + ((ax - r0) * W + (r.y + pr.ry)) * K;
      cell[0] = 0;
      cell[pr.net + 1] = 1;
    }
    __syncwarp();
    copy_out(p.out.pgrid + (size_t)env * sh.pgrid_b() + (size_t)r0 * W * K, e.stage, nbytes, lane);
    __syncwarp();
  }

  // component_grid (spatial:1677-1697): drawn at reset from the UN-rotated pin coordinates
  // and constant during the episode.  The state holds the pins as rotated by the placements
  // so far (place_component rotates in place, pin:156-204); orient[c] undoes that, so the
  // bank record is not read again after reset.
  zero_smem(e.stage, (sh.cgrid_b() + 15) & ~15, lane);
  __syncwarp();
  const int cells = sh.mh() * sh.mw();
  for (int i = lane; i < n_comp * cells; i += 32) e.stage[i * K] = 1;
  __syncwarp();
  for (int q = lane; q < n_pins; q += 32) {
    const PinRec pr = e.pins[q];
    const CompRec r = e.comps[pr.comp];
    const int oc = e.orient[pr.comp];
    int rx = pr.rx, ry = pr.ry;
    if (oc == 1) { rx = r.h - 1 - pr.ry; ry = pr.rx; }
    else if (oc == 2) { rx = r.h - 1 - pr.rx; ry = r.w - 1 - pr.ry; }
    else if (oc == 3) { rx = pr.ry; ry = r.w - 1 - pr.rx; }
    e.stage[((pr.comp * sh.mh() + rx) * sh.mw() + ry) * K + pr.net + 1] = 1;
  }
  __syncwarp();
  copy_out(p.out.cgrid + (size_t)env * sh.cgrid_b(), e.stage, sh.cgrid_b(), lane);
  __syncwarp();
  if (bulk && lane == 0) bulk_commit();
}

// Mask rows of the component to place next (both orientations); zero when none is left
// (rect:399-402, pin:1663-1666).  Returns true when at least one position is valid.
template <int RPL, typename T, typename S>
__device__ __forceinline__ bool next_masks(const S& sh, const Dims& d, const Env<T>& e,
                                           const T (&occ)[RPL], int cur, int n_comp, int lane,
                                           T (&m0)[RPL], T (&m1)[RPL]) {
#pragma unroll
  for (int k = 0; k < RPL; ++k) { m0[k] = T(0); m1[k] = T(0); }
  if (sh.kind() == 0) {
    footprint_mask<RPL, T>(occ, d.sq_n, d.sq_n, sh.H(), sh.W(), lane, m0);
  } else if (cur < n_comp) {
    CompRec r = e.comps[cur];
    footprint_mask<RPL, T>(occ, r.h, r.w, sh.H(), sh.W(), lane, m0);
    if (r.h == r.w) {
#pragma unroll
      for (int k = 0; k < RPL; ++k) m1[k] = m0[k];
    } else {
      footprint_mask<RPL, T>(occ, r.w, r.h, sh.H(), sh.W(), lane, m1);
    }
  }
  T any = T(0);
#pragma unroll
  for (int k = 0; k < RPL; ++k) any |= m0[k] | m1[k];
  return __any_sync(kFull, any != T(0));
}

// reset(): bank record -> state, empty board (rect:310-351, pin:1544-1597).
template <int RPL, typename T, typename S>
__device__ __forceinline__ void load_instance(const S& sh, const Params& p, Env<T>& e,
                                              uint32_t slot, uint64_t episode_no, int lane) {
  if (sh.kind() != 0) {
    const uint8_t* rec = p.bank + (size_t)slot * sh.rec_stride();
    // header (16 B) then comps/pins; state has the 16-byte EnvDyn in between
    if (lane == 0) *reinterpret_cast<uint4*>(e.sm) = *reinterpret_cast<const uint4*>(rec);
    copy16(e.sm + 32, rec + 16, sh.rec_stride() - 16, lane);
  }
  __syncwarp();
  if (lane == 0) {
    if (sh.kind() == 0) { e.hdr->n_comp = 1; e.hdr->n_nets = 0; e.hdr->n_pins = 0; e.hdr->pad = 0; }
    e.hdr->cur = 0;
    e.hdr->flags = 0;
    e.hdr->slot = slot;
    e.hdr->ep_len = 0;
    e.dyn->ep_return = 0.0;
    e.dyn->episode_no = episode_no;
  }
#pragma unroll
  for (int k = 0; k < RPL; ++k) e.occ[lane + 32 * k] = T(0);
  if (sh.kind() == 3)
    for (int c = lane; c < sh.C(); c += 32) e.orient[c] = 0;
  __syncwarp();
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

// Per-warp scratch of the reward.  A route is stored per pin: the segment "headed" by
// pin q runs from (px[q], py[q]) to (x2[q], y2[q]); box[q] is a conservative integer
// bounding box as two words with 16-bit fields: lo = x0 | y0<<16 and hi = (x1 | y1<<16) |
// kBoxGuard; kNoBoxLo in lo (and an empty hi) when pin q heads no segment.
constexpr uint32_t kBoxGuard = 0x80008000u;
constexpr uint32_t kNoBoxLo = 0x7fff7fffu;
struct RouteScratch {
  double* x2;
  double* y2;
  double* dist;
  double* cx;         // [K] centroid per net
  double* cy;
  uint2* box;         // [max_pins] (lo, guarded hi)
  uint2* seg;         // [max_pins] integer endpoints times scale: x = ax | ay << 16, y = bx | by << 16
  uint32_t* queue;    // [kPairQueue] candidate pairs (s << 16 | t) whose boxes touch
  uint32_t* hqueue;   // [kPairQueue] pairs the integer test leaves to the float64 formula
  int* nsum;          // [2 K] coordinate sums of each net (x, y)
  int16_t* px;
  int16_t* py;
  uint16_t* nxt;      // beam route: next pin on the path, 0xFFFF = path end
  uint16_t* net_off;  // [K + 1] first pin of each net (pins are net-major)
  uint16_t* start;    // [K] beam route: first pin of the path
  uint8_t* net;       // [max_pins] net of pin q
  uint8_t* scale;     // [max_pins] scale byte of seg[q] (pcb_cross.h)
};

__device__ __forceinline__ RouteScratch route_scratch(uint8_t* scr, int max_pins, int K) {
  RouteScratch r;
  uint8_t* q = scr;
  r.x2 = reinterpret_cast<double*>(q); q += max_pins * 8;
  r.y2 = reinterpret_cast<double*>(q); q += max_pins * 8;
  r.dist = reinterpret_cast<double*>(q); q += max_pins * 8;
  r.cx = reinterpret_cast<double*>(q); q += K * 8;
  r.cy = reinterpret_cast<double*>(q); q += K * 8;
  r.box = reinterpret_cast<uint2*>(q); q += max_pins * 8;
  r.seg = reinterpret_cast<uint2*>(q); q += max_pins * 8;
  r.queue = reinterpret_cast<uint32_t*>(q); q += kPairQueue * 4;
  r.hqueue = reinterpret_cast<uint32_t*>(q); q += kPairQueue * 4;
  r.nsum = reinterpret_cast<int*>(q); q += K * 8;
  r.px = reinterpret_cast<int16_t*>(q); q += max_pins * 2;
  r.py = reinterpret_cast<int16_t*>(q); q += max_pins * 2;
  r.nxt = reinterpret_cast<uint16_t*>(q); q += max_pins * 2;
  r.net_off = reinterpret_cast<uint16_t*>(q); q += (K + 1) * 2;
  r.start = reinterpret_cast<uint16_t*>(q); q += K * 2;
  r.net = q; q += max_pins;
  r.scale = q;
  return r;
}

__device__ __forceinline__ uint2 pack_box(int x0, int y0, int x1, int y1) {
  return make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16),
                    ((uint32_t)x1 | ((uint32_t)y1 << 16)) | kBoxGuard);
}
// pin -> pin segment
__device__ __forceinline__ uint2 make_box(int xa, int ya, int xb, int yb) {
  return pack_box(min(xa, xb), min(ya, yb), max(xa, xb), max(ya, yb));
}
// pin -> centroid segment.  The box only has to contain the segment: rounding the centroid
// (a multiple of 1 / pins, pins <= 512, below 64) to float moves it by < 4e-6, which cannot
// carry it across an integer it is not equal to.
__device__ __forceinline__ uint2 make_box(int xa, int ya, double cx, double cy) {
  const float fx = (float)cx, fy = (float)cy;
  return pack_box(min(xa, __float2int_rd(fx)), min(ya, __float2int_rd(fy)),
                  max(xa, __float2int_ru(fx)), max(ya, __float2int_ru(fy)));
}
__device__ __forceinline__ uint2 no_box() { return make_uint2(kNoBoxLo, kBoxGuard); }

// closed boxes a, b touch  <=>  b.hi >= a.lo and a.hi >= b.lo in both fields; the guard bit
// of a field survives the subtraction exactly when the difference is non-negative
__device__ __forceinline__ bool boxes_touch(uint2 a, uint2 b) {
  return (((b.y - a.x) & (a.y - b.x)) & kBoxGuard) == kBoxGuard;
}

// positions, net ranges, centroids (get_net_pin_positions pin:1266-1280, get_centroid :1282-1294)
__device__ __forceinline__ void route_prepare(const JobPin* pins, const RouteScratch& rs, int lane,
                                              int n_pins, int n_nets) {
  for (int q = lane; q < n_pins; q += 32) {
    JobPin jp = pins[q];
    rs.px[q] = jp.x;
    rs.py[q] = jp.y;
    rs.net[q] = jp.net;
    if (q == 0 || pins[q - 1].net != jp.net) rs.net_off[jp.net] = (uint16_t)q;
  }
  if (lane == 0) rs.net_off[n_nets] = (uint16_t)n_pins;
  __syncwarp();
  // exact integer sums, one correctly rounded division per coordinate
  for (int net = lane; net < n_nets; net += 32) {
    int lo = rs.net_off[net], hi = rs.net_off[net + 1], sx = 0, sy = 0;
    for (int q = lo; q < hi; ++q) { sx += rs.px[q]; sy += rs.py[q]; }
    rs.cx[net] = (double)sx / (double)(hi - lo);
    rs.cy[net] = (double)sy / (double)(hi - lo);
    rs.nsum[2 * net] = sx;
    rs.nsum[2 * net + 1] = sy;
  }
  __syncwarp();
}

#ifndef PCB_CROSS_INT
#define PCB_CROSS_INT 1   // 0 = round-1 crossing count (box filter, then float64 for every survivor)
#endif

// Integer record of the segment headed by pin q (pcb_cross.h): both end points multiplied by
// `scale`; nets of more than kMaxExactScale pins get scale byte 0 and their pairs go to the
// float64 formula directly.
__device__ __forceinline__ void put_segment(const RouteScratch& rs, int q, int ax, int ay, int bx,
                                            int by, int scale) {
  rs.seg[q] = make_uint2((uint32_t)ax | ((uint32_t)ay << 16), (uint32_t)bx | ((uint32_t)by << 16));
  rs.scale[q] = (uint8_t)segment_scale_byte(scale);
}

__device__ __forceinline__ int segments_cross_int(const RouteScratch& rs, uint32_t pr) {
  const int s = pr >> 16, t = pr & 0xffffu;
  const uint2 S = rs.seg[s], T = rs.seg[t];
  return segments_cross_int((int)(S.x & 0xffffu), (int)(S.x >> 16), (int)(S.y & 0xffffu),
                            (int)(S.y >> 16), rs.scale[s], (int)(T.x & 0xffffu), (int)(T.x >> 16),
                            (int)(T.y & 0xffffu), (int)(T.y >> 16), rs.scale[t]);
}

// find_num_intersection (pin:663-685) over the per-pin segment arrays: every pair of
// segments of different nets.  Pairs whose integer boxes do not touch cannot intersect (the
// reference requires the point inside both boxes); the others are queued, decided 32 at a
// time by the exact integer test, and what that leaves undecided is queued again for the
// reference's float64 formula.
__device__ __forceinline__ int count_crossings(const RouteScratch& rs, int lane, int n_pins,
                                               int n_nets) {
  int cnt = 0, qn = 0;
  auto heavy = [&](uint32_t pr) -> int {
    int s = pr >> 16, t = pr & 0xffffu;
#ifdef PCB_EXP_NO_HEAVY  // sensitivity experiment only (wrong rewards): the filter without the float64 test
    return (s + t) & 1;
#endif
    return segments_cross((double)rs.px[s], (double)rs.py[s], rs.x2[s], rs.y2[s],
                          (double)rs.px[t], (double)rs.py[t], rs.x2[t], rs.y2[t]) ? 1 : 0;
  };
#if PCB_CROSS_INT
  int hn = 0;
  const unsigned below = (1u << lane) - 1u;
  // the first n (<= 32) entries of queue[qn ...): integer test, undecided pairs to hqueue
  auto decide = [&](int n) {
    uint32_t pr = 0;
    int r = 0;
    if (lane < n) {
      pr = rs.queue[qn + lane];
      r = segments_cross_int(rs, pr);
    }
    cnt += r & 1;
    const unsigned hb = __ballot_sync(kFull, r == 2);
    if (hb != 0) {
      if (r == 2) rs.hqueue[hn + __popc(hb & below)] = pr;
      hn += __popc(hb);
      __syncwarp();
      if (hn >= 32) {
        hn -= 32;
        cnt += heavy(rs.hqueue[hn + lane]);
        __syncwarp();
      }
    }
  };
  // net b against every segment of the nets before it: pair i = (s, t) = (i % first, first + i / first)
  for (int b = 1; b < n_nets; ++b) {
    const int first = rs.net_off[b], total = first * (rs.net_off[b + 1] - first);
    // i / first through a reciprocal good to 2 ulp: (i + 0.5) / first is at least 1 / 1024 away
    // from an integer (first <= 512) and the product is off by less than 512 * 2^-21
    const float inv = __fdividef(1.0f, (float)first);
    for (int i0 = 0; i0 < total; i0 += 32) {
      const int i = min(i0 + lane, total - 1);  // lanes past the end repeat the last pair (loads stay in range)
      const int q = __float2int_rz(((float)i + 0.5f) * inv);
      const int s = i - q * first, t = first + q;
      const bool pass = boxes_touch(rs.box[t], rs.box[s]) && i0 + lane < total;
      const unsigned pb = __ballot_sync(kFull, pass);
      if (pb == 0) continue;
      if (pass) rs.queue[qn + __popc(pb & below)] = ((uint32_t)s << 16) | (uint32_t)t;
      qn += __popc(pb);
      __syncwarp();
      if (qn >= 32) {
        qn -= 32;
        decide(32);
        __syncwarp();
      }
    }
  }
  if (qn > 0) {
    const int n = qn;
    qn = 0;
    decide(n);
  }
  if (lane < hn) cnt += heavy(rs.hqueue[lane]);
  __syncwarp();
  return __reduce_add_sync(kFull, cnt);
#else
  for (int t0 = 0; t0 < n_pins; t0 += 32) {
    const int t = t0 + lane;
    uint2 mine = no_box();
    int my_hi = 0;
    if (t < n_pins) {
      mine = rs.box[t];
      if (mine.x != kNoBoxLo) my_hi = rs.net_off[rs.net[t]];  // segments of earlier nets only
    }
    const int smax = __reduce_max_sync(kFull, my_hi);
    for (int s = 0; s < smax; ++s) {
      const uint2 other = rs.box[s];
      if (other.x == kNoBoxLo) continue;  // warp-uniform
      const bool pass = s < my_hi && boxes_touch(mine, other);
      const unsigned b = __ballot_sync(kFull, pass);
      if (b == 0) continue;
      if (pass) rs.queue[qn + __popc(b & ((1u << lane) - 1u))] = ((uint32_t)s << 16) | (uint32_t)t;
      qn += __popc(b);
      __syncwarp();
      if (qn >= 32) {
        qn -= 32;
        cnt += heavy(rs.queue[qn + lane]);
        __syncwarp();
      }
    }
  }
  if (lane < qn) cnt += heavy(rs.queue[lane]);
  __syncwarp();
  return __reduce_add_sync(kFull, cnt);
#endif
}

// Centroid route (route_pins_centroid, pin:1296-1324): segment headed by pin q =
// (pin q, centroid of its net); a 2-pin net has the single segment (pin 0, pin 1).
__device__ __forceinline__ void centroid_route(const Dims& d, const RouteScratch& rs, int lane,
                                               int n_pins, int n_nets, int& crossings, double& wl) {
  for (int q = lane; q < n_pins; q += 32) {
    const int net = rs.net[q];
    const int lo = rs.net_off[net], cnt = rs.net_off[net + 1] - lo;
    double x2 = 0.0, y2 = 0.0, len = 0.0;  // + 0.0 leaves the running sum unchanged
    uint2 box = no_box();
    if (cnt != 2 || q == lo) {
      const int ix = rs.px[q], iy = rs.py[q];
      if (cnt == 2) {
        const int jx = rs.px[q + 1], jy = rs.py[q + 1];
        x2 = (double)jx; y2 = (double)jy;
        put_segment(rs, q, ix, iy, jx, jy, 1);
        box = make_box(ix, iy, jx, jy);
      } else {
        x2 = rs.cx[net]; y2 = rs.cy[net];
        if (cnt <= kMaxExactScale)
          put_segment(rs, q, ix * cnt, iy * cnt, rs.nsum[2 * net], rs.nsum[2 * net + 1], cnt);
        else
          put_segment(rs, q, 0, 0, 0, 0, 0);  // coordinates would not fit 16 bits: float64 only
        box = make_box(ix, iy, x2, y2);
      }
      const double x1 = (double)ix, y1 = (double)iy;
      len = norm2(x1 - x2, y1 - y2, d.norm_fma);
    }
    rs.x2[q] = x2; rs.y2[q] = y2; rs.box[q] = box; rs.dist[q] = len;
  }
  __syncwarp();
#ifdef PCB_EXP_NO_CROSS  // sensitivity experiment only (wrong rewards): how much of the interference is the crossing count?
  crossings = 0;
#else
  crossings = count_crossings(rs, lane, n_pins, n_nets);
#endif
  // find_wirelength (pin:741-759): the reference's sequential float64 sum, net by net (pins
  // that head no segment hold 0.0).  Every lane runs the same chain over broadcast loads.
  double acc = 0.0;
  for (int q = 0; q < n_pins; ++q) acc += rs.dist[q];
  wl = acc;
  __syncwarp();
}

}  // namespace pcb

#include "pcb_beam.cuh"

namespace pcb {

// End of an episode inside step_kernel: episode statistics (or, when the routing reward
// is still to be computed, a job for reward_kernel), and with auto-reset the next
// instance from the bank plus its first observation.
template <int RPL, typename T, typename S>
__device__ __forceinline__ void finish_episode(const S& sh, const Params& p, Env<T>& e, int env,
                                               int lane, double reward, bool route,
                                               uint8_t* gstate) {
  const Dims& d = p.d;
  const int kind = sh.kind();
  const bool first_report = !(e.hdr->flags & kFlagFinished);
  if (route) {
    // all components placed: find_reward needs the routes -> leave the pin positions
    const int n_pins = e.hdr->n_pins;
    uint32_t idx = 0;
    if (lane == 0) idx = atomicAdd(p.job_count + (p.step_no & p.job_slot_mask), 1u);
    idx = __shfl_sync(kFull, idx, 0);
    uint8_t* job = p.jobs + (size_t)idx * job_stride(sh.max_pins(), sh.K());
    if (lane == 0) {
      RouteJob j;
      j.env = (uint32_t)env; j.n_pins = (uint16_t)n_pins; j.n_nets = e.hdr->n_nets;
      j.count_stats = first_report ? 1 : 0; j.ep_len = e.hdr->ep_len; j.pad = p.job_t;
      j.ep_return = e.dyn->ep_return; j.pad2 = 0.0;
      *reinterpret_cast<RouteJob*>(job) = j;
    }
    JobPin* jp = reinterpret_cast<JobPin*>(job + 32);
    for (int q = lane; q < n_pins; q += 32) {
      PinRec pr = e.pins[q];
      CompRec r = e.comps[pr.comp];
      jp[q] = JobPin{(uint8_t)(r.x + pr.rx), (uint8_t)(r.y + pr.ry), pr.net, 0};
    }
  } else if (lane == 0) {
    double ret = e.dyn->ep_return;
    p.out.ep_return[env] = ret;
    p.out.ep_length[env] = (int32_t)e.hdr->ep_len;
    if (first_report) {
      double* st = p.stats + (size_t)(env & (kStatSlots - 1)) * 8;
      atomicAdd(st + 0, 1.0);
      atomicAdd(st + 1, ret);
      atomicAdd(st + 2, (double)e.hdr->ep_len);
      if (kind >= 2) { atomicAdd(st + 3, d.pen_wl); atomicAdd(st + 4, d.pen_ni); }
    }
  }
  __syncwarp();

  if (!d.autoreset) {
    if (lane == 0 && first_report) {
      e.hdr->flags |= kFlagFinished;
      st_state(gstate, *reinterpret_cast<const uint4*>(e.sm), (RPL == 2 && p.state_stream != 0));
    }
    return;
  }
  // auto-reset: next bank slot of this env, first observation of the new episode
  uint32_t slot = e.hdr->slot + (uint32_t)p.n_envs;
  if (p.bank_count) slot %= p.bank_count;
  uint64_t epno = e.dyn->episode_no + 1;
  __syncwarp();
  load_instance<RPL, T, S>(sh, p, e, slot, epno, lane);
  T occ[RPL], m0[RPL], m1[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) occ[k] = T(0);
  next_masks<RPL, T, S>(sh, d, e, occ, 0, e.hdr->n_comp, lane, m0, m1);
  store_state(gstate, e.sm, sh.st_stride(), lane, (RPL == 2 && p.state_stream != 0));
  write_obs<RPL, T, S>(sh, p, e, env, lane, occ, m0, m1);
  if (p.actions_out)
    sample_action<RPL, T, S>(sh, m0, m1, p.seed, env, policy_step(p) + 1, lane, p.actions_out);
}

// ---------------------------------------------------------------------------
// K1/K2/K3 (+ fused K4/K6/K7): step
// ---------------------------------------------------------------------------
#ifndef PCB_OPT_PERSIST
#define PCB_OPT_PERSIST 0
#endif
#ifndef PCB_OPT_PREFETCH
#define PCB_OPT_PREFETCH 0   // > 0: L2-prefetch the state of the CTA this many blocks ahead
#endif
// the action of one env as (o, x, y): tuple or flat (FlatteningActionWrapper*:
// env_wrappers.py:80-98, wrapper_square.py:54-69)
template <typename S>
__device__ __forceinline__ void load_action(const S& sh, const Params& p, int env, int& o, int& x, int& y) {
  o = 0; x = 0; y = 0;
  if (p.action_format == 1) {
    int a = p.actions[env];
    int hw = sh.H() * sh.W();
    if (a < 0) { o = -1; }
    else { o = a / hw; int rem = a - o * hw; x = rem / sh.W(); y = rem - x * sh.W(); }
  } else if (sh.kind() == 0) {
    x = p.actions[2 * env]; y = p.actions[2 * env + 1];
  } else {
    o = p.actions[3 * env]; x = p.actions[3 * env + 1]; y = p.actions[3 * env + 2];
  }
}

// one environment, one warp: everything of step() after the state chunk st0 (lane's 16
// bytes of the record) and the action have been requested
template <int RPL, typename T, typename S>
__device__ __forceinline__ void step_env(const S& sh, const Params& p, Env<T>& e, int env, int lane,
                                         uint4 st0, int o, int x, int y) {
  const Dims& d = p.d;
  const int kind = sh.kind();
  const int n16 = sh.st_stride() >> 4;
  uint8_t* gstate = p.state + (size_t)env * sh.st_stride();
  if (lane < n16) reinterpret_cast<uint4*>(e.sm)[lane] = st0;
  for (int i = 32 + lane; i < n16; i += 32)
    reinterpret_cast<uint4*>(e.sm)[i] = ld_state(reinterpret_cast<const uint4*>(gstate) + i, (RPL == 2 && p.state_stream != 0));
  __syncwarp();

  T occ[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) occ[k] = e.occ[lane + 32 * k];
  const int n_comp = e.hdr->n_comp;
  int cur = e.hdr->cur;
  int ph = d.sq_n, pw = d.sq_n, ch = 0, cw = 0;
  bool in_range = (unsigned)o < (unsigned)sh.O() && (unsigned)x < (unsigned)sh.H() &&
                  (unsigned)y < (unsigned)sh.W();
  if (kind != 0) {
    in_range = in_range && cur < n_comp;
    if (in_range) {
      CompRec r = e.comps[cur];
      ch = r.h; cw = r.w;
      ph = (o & 1) ? cw : ch;  // footprint (h,w) for o in {0,2}, (w,h) for {1,3}: pin:1755-1760
      pw = (o & 1) ? ch : cw;
    }
  }
  // validate_action == mask lookup (pin:1712-1736); the mask bit is recomputed here.
  bool valid = in_range && x + ph <= sh.H() && y + pw <= sh.W();
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
    if (kind != 0) {
      if (lane == 0) {
        e.comps[cur].x = (uint8_t)x; e.comps[cur].y = (uint8_t)y;
        if (kind == 3) e.orient[cur] = (uint8_t)o;
      }
      if (kind >= 2 && o != 0) {  // Component.place_component (pin:156-204): rotate in place
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
  const bool any = next_masks<RPL, T, S>(sh, d, e, occ, cur, n_comp, lane, m0, m1);
  const bool all_placed = kind != 0 && cur == n_comp;
  const bool done = valid ? !any : true;  // compute_if_done (pin:1872-1886); invalid -> done
  const bool route = kind >= 2 && done && all_placed;  // find_reward needs the routes
  double reward;
  if (kind <= 1) reward = valid ? 1.0 : 0.0;      // square:115-153, rect:405-432
  else reward = done ? d.pen_reward : 0.0;        // pin:1669-1677, 899-909

  if (lane == 0) {
    e.hdr->cur = (uint8_t)cur;
    e.hdr->ep_len += 1;
    if (!route) e.dyn->ep_return += reward;
  }
  __syncwarp();
  const bool restart = done && d.autoreset;
  if (!restart) {  // a restarting env overwrites its state below
    uint4* g4 = reinterpret_cast<uint4*>(gstate);
    const uint4* s4 = reinterpret_cast<const uint4*>(e.sm);
    if (lane < 2) st_state(g4 + lane, s4[lane], (RPL == 2 && p.state_stream != 0));
    if (valid)
      for (int i = 2 + lane; i < (sh.st_stride() >> 4); i += 32) st_state(g4 + i, s4[i], (RPL == 2 && p.state_stream != 0));
  }

  if (!restart) {
    write_obs<RPL, T, S>(sh, p, e, env, lane, occ, m0, m1);
    if (p.actions_out)  // a finished, not restarting env has no valid action: reports (0,0,0)
      sample_action<RPL, T, S>(sh, m0, m1, p.seed, env, policy_step(p) + 1, lane, p.actions_out);
  }
  if (lane == 0) {
    p.out.done[env] = done ? 1 : 0;
    if (!route) {
      p.out.reward[env] = reward;
      if (kind >= 2 && done) { p.out.wl[env] = d.pen_wl; p.out.ni[env] = d.pen_ni; }
    }
  }
  if (done) finish_episode<RPL, T, S>(sh, p, e, env, lane, reward, route, gstate);
  // shared memory handed to the copy engine must stay in place until it has been read
}

template <int RPL, typename T, typename S>
__global__ void __launch_bounds__(kWarpsPerCta * 32, PCB_MIN_BLOCKS)
step_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const S sh(d);
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int n16 = sh.st_stride() >> 4;
  Env<T> e(smem + kLutBytes + (size_t)wic * sh.sm_stride(), sh);
#if PCB_OPT_PERSIST
  // persistent CTAs: the grid fills the machine once and every warp walks over its share of
  // the environments, so the expansion table is built once per CTA instead of once per 4 envs
  build_expand_lut(smem);
  for (int env = blockIdx.x * kWarpsPerCta + wic; env < p.n_envs; env += gridDim.x * kWarpsPerCta) {
    uint4 st0 = make_uint4(0, 0, 0, 0);
    if (lane < n16) st0 = ld_state(reinterpret_cast<const uint4*>(p.state + (size_t)env * sh.st_stride()) + lane, (RPL == 2 && p.state_stream != 0));
    int o, x, y;
    load_action(sh, p, env, o, x, y);
    step_env<RPL, T, S>(sh, p, e, env, lane, st0, o, x, y);
    // shared memory handed to the copy engine must stay in place until it has been read
    if constexpr (S::fixed) if (sh.sm_bulk() != 0 && lane == 0) bulk_wait_read();
    __syncwarp();
  }
#else
  const int env = blockIdx.x * kWarpsPerCta + wic;
  const bool live = env < p.n_envs;
  // issue every global load of this warp before anything waits: table, state, action
  const uint4 lut_part = expand_lut_fetch();
#if PCB_OPT_PREFETCH
  {
    // The state records come back from DRAM every step (they do not survive ~500 MB of
    // observation stores in L2) and the first use of the loaded record is the top stall of
    // this kernel.  CTAs start roughly in blockIdx order, so the CTA PCB_OPT_PREFETCH blocks
    // further on starts about one CTA lifetime later: pull its records into L2 now.
    const int penv = env + PCB_OPT_PREFETCH * kWarpsPerCta;
    if (penv < p.n_envs && lane * 128 < sh.st_stride())
      asm volatile("prefetch.global.L2 [%0];" :: "l"(p.state + (size_t)penv * sh.st_stride() + lane * 128));
  }
#endif
  uint4 st0 = make_uint4(0, 0, 0, 0);
  if (lane < n16) st0 = ld_state(reinterpret_cast<const uint4*>(p.state + (size_t)(live ? env : 0) * sh.st_stride()) + lane, (RPL == 2 && p.state_stream != 0));
  int o = 0, x = 0, y = 0;
  if (live) load_action(sh, p, env, o, x, y);
  expand_lut_store(smem, lut_part);
  if (!live) return;
  step_env<RPL, T, S>(sh, p, e, env, lane, st0, o, x, y);
  if constexpr (S::fixed) if (sh.sm_bulk() != 0 && lane == 0) bulk_wait_read();
#endif
}

// K4: the routing reward (find_reward, pin:832-975) of the episodes that finished in this
// step with every component placed.  A small persistent grid walks the job list that
// step_kernel left; one warp per job.  BEAM = true adds the beam-search router
// (reward types "beam" / "both"), which needs far more registers.
#ifndef PCB_REWARD_MIN_BLOCKS
#define PCB_REWARD_MIN_BLOCKS 1
#endif
template <bool BEAM>
__global__ void __launch_bounds__(kWarpsPerCta * 32, PCB_REWARD_MIN_BLOCKS)
reward_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int gw = blockIdx.x * kWarpsPerCta + wic, nw = gridDim.x * kWarpsPerCta;
  const uint32_t count = p.job_count[p.step_no & 3];
  // counters form a ring of 4: step t+2 cannot start before this kernel has finished
  if (gw == 0 && lane == 0) p.job_count[(p.step_no + 2) & 3] = 0;
  const int scr_bytes = route_scratch_bytes(d.max_pins, d.K);
  RouteScratch rs = route_scratch(smem + (size_t)wic * ((scr_bytes + 15) & ~15), d.max_pins, d.K);
  for (uint32_t j = gw; j < count; j += nw) {
    const uint8_t* job = p.jobs + (size_t)j * job_stride(d.max_pins, d.K);
    const RouteJob hdr = *reinterpret_cast<const RouteJob*>(job);
    const int n_pins = hdr.n_pins, n_nets = hdr.n_nets, env = (int)hdr.env;
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
    const double wl = length / d.wl_norm;             // pin:921 / 940 / 959-961
    const double ni = (double)crossings / d.ni_norm;  // pin:922-924
    const double reward = -1.0 * (d.ww * wl + d.wi * ni);
    if (lane == 0) {
      const double ret = hdr.ep_return + reward;
      p.out.reward[env] = reward; p.out.wl[env] = wl; p.out.ni[env] = ni;
      p.out.ep_return[env] = ret;
      p.out.ep_length[env] = (int32_t)hdr.ep_len;
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

// K6: reset() -- all envs (reset_mask == nullptr) or the masked ones.
template <int RPL, typename T, typename S>
__global__ void __launch_bounds__(kWarpsPerCta * 32) reset_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const S sh(d);
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerCta + wic;
  build_expand_lut(smem);
  if (env >= p.n_envs) return;
  if (p.reset_mask && !p.reset_mask[env]) return;
  Env<T> e(smem + kLutBytes + (size_t)wic * sh.sm_stride(), sh);
  uint8_t* gstate = p.state + (size_t)env * sh.st_stride();
  uint32_t slot = (uint32_t)env;
  uint64_t epno = 0;
  if (!p.reset_first) {
    const EnvHdr* gh = reinterpret_cast<const EnvHdr*>(gstate);
    const EnvDyn* gd = reinterpret_cast<const EnvDyn*>(gstate + 16);
    slot = gh->slot + (uint32_t)p.n_envs;
    epno = gd->episode_no + 1;
  }
  if (p.bank_count) slot %= p.bank_count;
  zero_smem(e.sm, sh.st_stride(), lane);
  __syncwarp();
  load_instance<RPL, T, S>(sh, p, e, slot, epno, lane);
  T occ[RPL], m0[RPL], m1[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) occ[k] = T(0);
  next_masks<RPL, T, S>(sh, d, e, occ, 0, e.hdr->n_comp, lane, m0, m1);
  store_state(gstate, e.sm, sh.st_stride(), lane, (RPL == 2 && p.state_stream != 0));
  write_obs<RPL, T, S>(sh, p, e, env, lane, occ, m0, m1);
  if (lane == 0) {
    p.out.done[env] = 0;
    p.out.reward[env] = 0.0;
  }
  if constexpr (S::fixed) if (sh.sm_bulk() != 0 && lane == 0) bulk_wait_read();
}

// K7 as a stand-alone launch (pcb_sample_random_actions): masks recomputed from the state.
template <int RPL, typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32) sample_kernel(const Params p) {
  const Dims& d = p.d;
  const DynShape sh(d);
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerCta + wic;
  if (env >= p.n_envs) return;
  const uint8_t* gstate = p.state + (size_t)env * sh.st_stride();
  const EnvHdr* gh = reinterpret_cast<const EnvHdr*>(gstate);
  const T* gocc = reinterpret_cast<const T*>(gstate + sh.st_occ());
  T occ[RPL], m0[RPL], m1[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) { occ[k] = gocc[lane + 32 * k]; m0[k] = T(0); m1[k] = T(0); }
  if (d.kind == 0) {
    footprint_mask<RPL, T>(occ, d.sq_n, d.sq_n, d.H, d.W, lane, m0);
  } else if (gh->cur < gh->n_comp) {
    CompRec r = reinterpret_cast<const CompRec*>(gstate + sh.st_comps())[gh->cur];
    footprint_mask<RPL, T>(occ, r.h, r.w, d.H, d.W, lane, m0);
    footprint_mask<RPL, T>(occ, r.w, r.h, d.H, d.W, lane, m1);
  }
  sample_action<RPL, T, DynShape>(sh, m0, m1, p.seed, env, p.step_no, lane, p.actions_out);
}

}  // namespace pcb
#include "pcb_reward_half.cuh"

#endif
