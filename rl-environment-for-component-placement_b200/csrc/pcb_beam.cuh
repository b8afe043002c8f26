// pcb_beam.cuh -- the beam-search route of the reward (route_pins_beam_search,
// dummy_env_rectangular_pin.py:1425-1460; beam_search :1356-1423; pin_outlier :1326-1339).
//
// The reference breaks distance ties at the beam cut-off by the iteration order
// of a CPython set difference, so this file reproduces Objects/setobject.c's
// table placement for (int, int) tuples: tuple hash, linear-probe + perturb
// sequence, growth at 3/5 fill to the next power of two above 4*used, the
// copy-vs-rebuild choice of set_difference.  The executable specification is
// oracle/pyset_model.py (checked against real CPython sets in
// tests/test_host.py::test_pyset_model_matches_cpython_sets).
// One lane routes one net (nets are independent); runs on terminal steps only.
#ifndef PCB_BEAM_CUH_
#define PCB_BEAM_CUH_

namespace pcb {

constexpr int kTab = 32;  // a set of <= 18 keys never grows past 32 slots

__device__ __forceinline__ uint64_t rotl64(uint64_t v, int s) { return (v << s) | (v >> (64 - s)); }

__device__ __forceinline__ uint64_t py_tuple_hash(int x, int y) {
  const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull,
                 P5 = 2870177450012600261ull;
  uint64_t acc = P5;
  acc += (uint64_t)x * P2; acc = rotl64(acc, 31); acc *= P1;
  acc += (uint64_t)y * P2; acc = rotl64(acc, 31); acc *= P1;
  acc += 2ull ^ (P5 ^ 3527539ull);
  return acc == ~0ull ? 1546275796ull : acc;
}

// The table lives in registers: `occ` has one bit per slot, the key of slot i (a point index,
// < 16) is nibble i of (k0, k1).  Emptiness tests, the linear probe and the in-order walks
// over the live slots are bit operations on `occ`; nothing is indexed dynamically in memory.
struct PyTab {
  uint32_t occ;
  uint64_t k0, k1;
  int mask, used;
  __device__ __forceinline__ int key(int i) const {
    return (int)(((i < 16 ? k0 : k1) >> ((i & 15) * 4)) & 15ull);
  }
  __device__ __forceinline__ void put(int i, int k) {
    const int sh = (i & 15) * 4;
    if (i < 16) k0 = (k0 & ~(15ull << sh)) | ((uint64_t)k << sh);
    else k1 = (k1 & ~(15ull << sh)) | ((uint64_t)k << sh);
    occ |= 1u << i;
  }
};
static_assert(kMaxBeamPoints < 16 && kTab == 32, "PyTab packs 32 slots of 4-bit keys");

// set_add_entry's probe for a free slot: the home slot, the 9 slots after it (LINEAR_PROBES)
// when they are inside the table, then the perturbed recurrence
__device__ __forceinline__ int tab_free_slot(uint32_t occ, int mask, uint64_t h) {
  uint64_t perturb = h;
  int i = (int)(h & (uint64_t)mask);
  while (true) {
    if (!((occ >> i) & 1u)) return i;
    if (i + 9 <= mask) {
      const uint32_t free9 = (~occ >> (i + 1)) & 0x1ffu;  // slots i+1 .. i+9
      if (free9) return i + __ffs((int)free9);
    }
    perturb >>= 5;
    i = (int)(((uint64_t)i * 5 + 1 + perturb) & (uint64_t)mask);
  }
}

__device__ __forceinline__ void tab_clear(PyTab& t, int size) {
  t.occ = 0u; t.k0 = 0ull; t.k1 = 0ull;
  t.mask = size - 1;
  t.used = 0;
}

// re-insert the live keys of `t` in table order into a table of the first power
// of two > minused (set_table_resize)
__device__ __forceinline__ void tab_resize(PyTab& t, int minused, const uint64_t* h) {
  int size = 8;
  while (size <= minused) size <<= 1;
  const PyTab old = t;
  tab_clear(t, size);
  for (uint32_t live = old.occ; live; live &= live - 1) {
    const int k = old.key(__ffs((int)live) - 1);
    t.put(tab_free_slot(t.occ, t.mask, h[k]), k);
  }
  t.used = old.used;
}

__device__ __forceinline__ void tab_add(PyTab& t, int key, const uint64_t* h) {  // set_add_entry
  t.put(tab_free_slot(t.occ, t.mask, h[key]), key);
  t.used += 1;
  if (t.used * 5 >= t.mask * 3) tab_resize(t, t.used * 4, h);
}

// iteration order of (targets - seen); returns the number of keys written to out[]
__device__ __forceinline__ int tab_difference(const PyTab& targets, uint32_t seen,
                                              const uint64_t* h, int8_t* out) {
  PyTab r;
  int n = 0;
  if ((targets.used >> 2) > __popc(seen)) {
    // set_copy_and_difference: copy (set_merge into a fresh set), then discard
    tab_clear(r, 8);
    if (targets.used * 5 >= r.mask * 3) {
      int size = 8;
      while (size <= targets.used * 2) size <<= 1;
      tab_clear(r, size);
    }
    if (r.mask == targets.mask) {
      for (uint32_t live = targets.occ; live; live &= live - 1) {
        const int k = targets.key(__ffs((int)live) - 1);
        if (!((seen >> k) & 1u)) out[n++] = (int8_t)k;
      }
      return n;
    }
    for (uint32_t live = targets.occ; live; live &= live - 1) {
      const int k = targets.key(__ffs((int)live) - 1);
      r.put(tab_free_slot(r.occ, r.mask, h[k]), k);
    }
    for (uint32_t live = r.occ; live; live &= live - 1) {
      const int k = r.key(__ffs((int)live) - 1);
      if (!((seen >> k) & 1u)) out[n++] = (int8_t)k;
    }
    return n;
  }
  tab_clear(r, 8);
  for (uint32_t live = targets.occ; live; live &= live - 1) {
    const int k = targets.key(__ffs((int)live) - 1);
    if (!((seen >> k) & 1u)) tab_add(r, k, h);
  }
  for (uint32_t live = r.occ; live; live &= live - 1) out[n++] = (int8_t)r.key(__ffs((int)live) - 1);
  return n;
}

struct BeamItem {
  double prio;
  uint64_t path;  // nibble j = local index of the j-th point visited after the start point
  uint32_t seen;
};
__device__ __forceinline__ int path_at(uint64_t path, int j) { return (int)((path >> (4 * j)) & 15ull); }

// lexicographic (priority, path) order of heapq on (float, list-of-tuples) entries
__device__ __forceinline__ bool beam_less(const BeamItem& a, const BeamItem& b, int len,
                                          const int16_t* x, const int16_t* y) {
  if (a.prio != b.prio) return a.prio < b.prio;
  for (int i = 0; i < len; ++i) {
    int pa = path_at(a.path, i), pb = path_at(b.path, i);
    if (x[pa] != x[pb]) return x[pa] < x[pb];
    if (y[pa] != y[pb]) return y[pa] < y[pb];
  }
  return false;
}

// Routes one net: pins[lo .. lo+m_all) of a routing job.  Writes the path as a linked list
// over the job's pin indices (nxt[pin], 0xFFFF = end) and the index of its first pin.
__device__ __forceinline__ void beam_route_net(const JobPin* pins, int lo, int m_all, int bw,
                                               int norm_fma, uint16_t* nxt, uint16_t* start_out) {
  // get_centroid (exact integer sums, one division per coordinate), then pin_outlier:
  // the first farthest pin from the centroid
  int sxi = 0, syi = 0;
  for (int i = 0; i < m_all; ++i) { sxi += pins[lo + i].x; syi += pins[lo + i].y; }
  const double cx = (double)sxi / (double)m_all, cy = (double)syi / (double)m_all;
  int start = 0;
  {
    double best = -1.0;
    for (int i = 0; i < m_all; ++i) {
      double v = norm2((double)pins[lo + i].x - cx, (double)pins[lo + i].y - cy, norm_fma);
      if (v > best) { best = v; start = i; }
    }
  }
  // points = net pins without the start (list.remove keeps the order)
  int16_t x[kMaxBeamPoints + 1], y[kMaxBeamPoints + 1];
  uint16_t gid[kMaxBeamPoints + 1];
  uint64_t h[kMaxBeamPoints + 1];
  int m = 0;
  for (int i = 0; i < m_all; ++i) {
    if (i == start) continue;
    x[m] = pins[lo + i].x; y[m] = pins[lo + i].y; gid[m] = (uint16_t)(lo + i);
    h[m] = py_tuple_hash(x[m], y[m]);
    ++m;
  }
  const int sx = pins[lo + start].x, sy = pins[lo + start].y;
  PyTab targets;
  tab_clear(targets, 8);
  for (int k = 0; k < m; ++k) tab_add(targets, k, h);

  BeamItem q[kMaxBeamWidth * kMaxBeamWidth], nq[kMaxBeamWidth * kMaxBeamWidth];
  int qn = 1;
  q[0].prio = 0.0; q[0].seen = 0u; q[0].path = 0ull;
  int best_idx = 0;
  for (int level = 0; level <= m; ++level) {  // level == number of visited points
    // order the queue (selection of the min(bw, qn) smallest entries)
    int order[kMaxBeamWidth];
    int take = bw < qn ? bw : qn;
    uint32_t used = 0;
    for (int t = 0; t < take; ++t) {
      int bi = -1;
      for (int i = 0; i < qn; ++i) {
        if ((used >> i) & 1u) continue;
        if (bi < 0 || beam_less(q[i], q[bi], level, x, y)) bi = i;
      }
      used |= 1u << bi;
      order[t] = bi;
    }
    if (level == m) { best_idx = order[0]; break; }  // first popped complete path wins
    int nn = 0;
    for (int t = 0; t < take; ++t) {
      const BeamItem& it = q[order[t]];
      const int last = level ? path_at(it.path, level - 1) : 0;
      const int hx = level ? x[last] : sx, hy = level ? y[last] : sy;
      int8_t cand[kMaxBeamPoints + 1];
      double cd[kMaxBeamPoints + 1];
      int nc = tab_difference(targets, it.seen, h, cand);
      for (int i = 0; i < nc; ++i)
        cd[i] = norm2((double)(hx - x[cand[i]]), (double)(hy - y[cand[i]]), norm_fma);
      // stable insertion sort by distance (sorted(..., key=...))
      for (int i = 1; i < nc; ++i) {
        int8_t ck = cand[i]; double dk = cd[i];
        int j = i - 1;
        while (j >= 0 && cd[j] > dk) { cand[j + 1] = cand[j]; cd[j + 1] = cd[j]; --j; }
        cand[j + 1] = ck; cd[j + 1] = dk;
      }
      int kids = bw < nc ? bw : nc;
      for (int i = 0; i < kids; ++i) {
        BeamItem& c = nq[nn++];
        c.prio = it.prio + cd[i];
        c.seen = it.seen | (1u << cand[i]);
        c.path = it.path | ((uint64_t)cand[i] << (4 * level));
      }
    }
    for (int i = 0; i < nn; ++i) q[i] = nq[i];
    qn = nn;
  }
  *start_out = (uint16_t)(lo + start);
  uint16_t prev = (uint16_t)(lo + start);
  for (int j = 0; j < m; ++j) {
    uint16_t g = gid[path_at(q[best_idx].path, j)];
    nxt[prev] = g;
    prev = g;
  }
  nxt[prev] = 0xFFFF;
}

// Beam routes of all finished episodes of a step: ONE THREAD PER (job, net), so a warp
// routes 32 independent nets (the router is serial per net).  The paths are written into
// the job records; reward_kernel<true> picks them up.
// The cost of a net grows steeply with its pin count (2 pins: a few hundred instructions, 6
// pins: several thousand), and consecutive (job, net) items have unrelated sizes, so a warp
// would run every lane's longest path.  Each CTA therefore counting-sorts its 128 items by
// pin count in shared memory first: the lanes of a warp then route nets of (almost) the same
// size.  The order of the items does not change any result.
#ifndef PCB_BEAM_SORT
#define PCB_BEAM_SORT 1
#endif
__global__ void __launch_bounds__(128) beam_kernel(const __grid_constant__ Params p) {
  const Dims& d = p.d;
  const int nets_max = d.K - 1;
  const uint32_t count = p.job_count[p.step_no & 3];
  const uint32_t total = count * (uint32_t)nets_max;
  const int stride = job_stride(d.max_pins, d.K);
  __shared__ uint32_t s_bin[32];
  __shared__ uint32_t s_item[128];
  const int tid = threadIdx.x;
  for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {  // CTA-uniform
    uint32_t idx = base + tid;
    int lo = 0, m = 0;
    if (idx < total) {
      const uint32_t j = idx / (uint32_t)nets_max;
      const int net = (int)(idx - j * (uint32_t)nets_max);
      const uint8_t* job = p.jobs + (size_t)j * stride;
      const RouteJob* hdr = reinterpret_cast<const RouteJob*>(job);
      const int n_pins = hdr->n_pins;
      if (net < hdr->n_nets) {
        const JobPin* pins = reinterpret_cast<const JobPin*>(job + 32);
        while (lo < n_pins && pins[lo].net != net) ++lo;  // pins are net-major
        int hi = lo;
        while (hi < n_pins && pins[hi].net == net) ++hi;
        m = hi - lo;
      }
    }
#if PCB_BEAM_SORT
    // counting sort of the CTA's items by pin count, largest first
    if (tid < 32) s_bin[tid] = 0u;
    __syncthreads();
    const int key = 31 - (m < 31 ? m : 31);
    const uint32_t rank = atomicAdd(&s_bin[key], 1u);
    __syncthreads();
    if (tid < 32) {
      const uint32_t c = s_bin[tid];
      uint32_t incl = c;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, off);
        if (tid >= off) incl += v;
      }
      s_bin[tid] = incl - c;  // first position of the bin
    }
    __syncthreads();
    s_item[s_bin[key] + rank] = (uint32_t)tid | ((uint32_t)lo << 8) | ((uint32_t)m << 20);
    __syncthreads();
    const uint32_t it = s_item[tid];
    idx = base + (it & 0xffu);
    lo = (int)((it >> 8) & 0xfffu);
    m = (int)(it >> 20);
    __syncthreads();  // s_bin / s_item are reused by the next round
#endif
    if (m > 0) {
      const uint32_t j = idx / (uint32_t)nets_max;
      const int net = (int)(idx - j * (uint32_t)nets_max);
      uint8_t* job = p.jobs + (size_t)j * stride;
      const JobPin* pins = reinterpret_cast<const JobPin*>(job + 32);
      uint16_t* nxt = reinterpret_cast<uint16_t*>(job + job_path_offset(d.max_pins));
      uint16_t* start = nxt + d.max_pins;
      beam_route_net(pins, lo, m, d.beam_width, d.norm_fma, nxt, start + net);
    }
  }
}

// crossings + wirelength of the beam route: segment headed by pin q = (q, nxt[q])
__device__ __forceinline__ void beam_route(const Dims& d, const RouteScratch& rs, int lane,
                                           int n_pins, int n_nets, int& crossings, double& wl) {
  for (int q = lane; q < n_pins; q += 32) {
    const int nx = rs.nxt[q];
    double x2 = 0.0, y2 = 0.0, len = 0.0;
    uint2 box = no_box();
    if (nx != 0xFFFF) {
      const double x1 = (double)rs.px[q], y1 = (double)rs.py[q];
      x2 = (double)rs.px[nx]; y2 = (double)rs.py[nx];
      put_segment(rs, q, rs.px[q], rs.py[q], rs.px[nx], rs.py[nx], 1);
      box = make_box((int)rs.px[q], (int)rs.py[q], (int)rs.px[nx], (int)rs.py[nx]);
      len = norm2(x1 - x2, y1 - y2, d.norm_fma);
    }
    rs.x2[q] = x2; rs.y2[q] = y2; rs.box[q] = box; rs.dist[q] = len;
  }
  __syncwarp();
  crossings = count_crossings(rs, lane, n_pins, n_nets);
  // find_wirelength: nets in order, segments in path order, sequential float64 sum
  double acc = 0.0;
  if (lane == 0) {
    for (int net = 0; net < n_nets; ++net) {
      int q = rs.start[net];
      while (rs.nxt[q] != 0xFFFF) {
        acc += rs.dist[q];
        q = rs.nxt[q];
      }
    }
  }
  wl = __shfl_sync(kFull, acc, 0);
  __syncwarp();
}

}  // namespace pcb
#endif
