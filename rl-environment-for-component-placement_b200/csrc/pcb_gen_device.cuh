// pcb_gen_device.cuh -- K8: the instance generator on the device.
//
// One thread draws one instance and writes its bank record (header, components, pins,
// winner flags, ranks) straight into the device bank, so a bank can be refreshed between
// steps without a host round trip: 262 144 fresh instances per launch.  Same construction
// as the host generator (pcb_gen.cpp, i.e. the reference's generate_instances,
// dummy_env_rectangular_pin.py:977-1264, 1462-1498) and the same per-instance SplitMix64
// stream, so host and device agree except where libm / libdevice differ in the last ulp
// of exp/log/cos; parity with the reference is distributional (tests/test_gpu_gen.py).
#ifndef PCB_GEN_DEVICE_CUH_
#define PCB_GEN_DEVICE_CUH_

#include "pcb_layout.h"

namespace pcb {

struct GenParams {
  int kind, C, P, max_pins, K;     // K = max_num_nets + 1
  int rec_stride, rec_comps, rec_pins;
  int min_w, max_w, min_h, max_h, min_c, max_c, min_nets, max_nets, lo, hi, nd, ps;
  uint64_t seed;
  uint32_t count;
  uint8_t* bank;
};

struct DevRng {
  uint64_t s;
  __device__ uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  __device__ double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  __device__ int randint(int lo, int hi) { return lo + (int)__umul64hi(next(), (uint64_t)(hi - lo + 1)); }
  __device__ double normal(double mean, double sd) {
    double u1 = uniform(), u2 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    return mean + sd * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  }
};

__device__ inline uint64_t gen_mix(uint64_t a, uint64_t b) {
  DevRng r{a ^ (b * 0xD6E8FEB86659FD93ull)};
  r.next();
  return r.next();
}

// categorical draw over p[0..n) (sums to ~1), zero-probability entries skipped
__device__ inline int gen_categorical(DevRng& r, const double* p, int n) {
  double u = r.uniform(), acc = 0.0;
  int last = 0;
  for (int i = 0; i < n; ++i) {
    if (p[i] <= 0.0) continue;
    acc += p[i];
    last = i;
    if (u < acc) return i;
  }
  return last;
}

__global__ void __launch_bounds__(128) generate_kernel(const GenParams g) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.count) return;
  uint8_t* rec = g.bank + (size_t)i * g.rec_stride;
  for (int b = 0; b < g.rec_stride; b += 16) *reinterpret_cast<uint4*>(rec + b) = make_uint4(0, 0, 0, 0);
  EnvHdr* hdr = reinterpret_cast<EnvHdr*>(rec);
  CompRec* comps = reinterpret_cast<CompRec*>(rec + g.rec_comps);
  PinRec* pins = reinterpret_cast<PinRec*>(rec + g.rec_pins);
  DevRng r{gen_mix(g.seed, (uint64_t)i)};

  const int nc = r.randint(g.min_c, g.max_c);
  int area[kMaxComps];
  int total_area = 0;
  for (int c = 0; c < nc; ++c) {
    int h = r.randint(g.min_h, g.max_h);
    int w = r.randint(g.min_w, g.max_w);
    comps[c] = CompRec{(uint8_t)h, (uint8_t)w, 0xFF, 0xFF};
    area[c] = h * w;
    total_area += area[c];
  }
  hdr->n_comp = (uint8_t)nc;
  if (g.kind < 2) return;

  int nets = min(r.randint(g.min_nets, g.max_nets), total_area / 2);
  int total = min(r.randint(g.lo * nets, g.hi * nets), total_area);
  // pins -> nets: softmax of normal logits, truncated sequential multinomial for the extras
  double prob[kMaxNets], q[kMaxNets];
  int extra[kMaxNets];
  double z = 0.0;
  for (int k = 0; k < nets; ++k) {
    prob[k] = exp(r.normal(1.0 / nets, 1.0 / (g.nd + 1)));
    z += prob[k];
    extra[k] = 0;
  }
  for (int k = 0; k < nets; ++k) prob[k] /= z;
  int todo = total - g.lo * nets;
  if (g.hi > g.lo && todo > 0) {
    int cap = min(g.hi - g.lo, todo);
    for (int t = 0; t < todo; ++t) {
      double zz = 0.0;
      for (int k = 0; k < nets; ++k) { q[k] = extra[k] < cap ? prob[k] : 0.0; zz += q[k]; }
      for (int k = 0; k < nets; ++k) q[k] /= zz;
      extra[gen_categorical(r, q, nets)] += 1;
    }
  }
  // pins -> components, net by net; pins come out in net-major order
  const int k0 = g.kind == 3 ? min((int)((g.ps / 10.0) * nc) + 1, nc)
                             : min(max((int)(((g.ps + 1) / 10.0) * nc), 1), nc);
  uint8_t order[kMaxComps];   // components sorted by free space (stable, descending)
  int space[kMaxComps];       // free cells per component id
  uint8_t rank[kMaxComps];
  for (int c = 0; c < nc; ++c) { order[c] = (uint8_t)c; space[c] = area[c]; rank[c] = 0; }
  int out = 0, extras_before = 0;
  for (int k = 0; k < nets; ++k) {
    int left = g.lo + extra[k];
    // stable insertion sort of `order` by space descending (keeps the previous order on ties)
    for (int a = 1; a < nc; ++a) {
      uint8_t v = order[a];
      int b = a - 1;
      while (b >= 0 && space[order[b]] < space[v]) { order[b + 1] = order[b]; --b; }
      order[b + 1] = v;
    }
    int kk = k0 - 1, cap = 0;
    while (cap < left && kk < nc) {
      ++kk; cap = 0;
      for (int c = 0; c < kk; ++c) cap += space[order[c]];
    }
    int done = 0;
    while (left > 0) {
      int tot = 0;
      for (int c = 0; c < kk; ++c) tot += space[order[c]];
      if (tot <= 0) break;  // cannot happen for valid ranges; avoids an endless loop
      double p[kMaxComps];
      int cnt[kMaxComps];
      for (int c = 0; c < kk; ++c) { p[c] = (double)space[order[c]] / tot; cnt[c] = 0; }
      for (int t = 0; t < left; ++t) cnt[gen_categorical(r, p, kk)] += 1;
      for (int c = 0; c < kk; ++c) {
        int comp = order[c];
        int take = min(cnt[c], space[comp]);
        space[comp] -= take;
        for (int j = 0; j < take; ++j) {
          int pj = done + j;  // index of this pin inside its net
          int creation = pj < g.lo ? k * g.lo + pj : g.lo * nets + extras_before + (pj - g.lo);
          PinRec pr;
          pr.comp = (uint8_t)comp; pr.rx = 0; pr.ry = 0; pr.net = (uint8_t)k;
          pr.pin_id = (uint16_t)(g.kind == 2 ? j : creation);
          pr.flags = g.kind == 3 ? kPinWinner : 0;
          pr.rank = 0;
          pins[out + pj] = pr;
        }
        done += take;
        left -= take;
      }
    }
    out += done;
    extras_before += extra[k];
  }
  total = out;
  hdr->n_pins = (uint16_t)total;
  hdr->n_nets = (uint8_t)nets;
  // relative cells: per component, its pins in net-major order pick a random free cell
  for (int c = 0; c < nc; ++c) {
    uint64_t used[4] = {0, 0, 0, 0};
    int free_cells = area[c];
    const int w = comps[c].w;
    for (int j = 0; j < total; ++j) {
      if (pins[j].comp != c) continue;
      int pick = r.randint(0, free_cells - 1);
      int cell = 0;
      for (;; ++cell)  // pick-th free cell
        if (!((used[cell >> 6] >> (cell & 63)) & 1ull) && pick-- == 0) break;
      used[cell >> 6] |= 1ull << (cell & 63);
      --free_cells;
      pins[j].rx = (uint8_t)(cell / w);
      pins[j].ry = (uint8_t)(cell % w);
      pins[j].rank = rank[c]++;
    }
  }
  // last writer of a (component, pin_id) feature row (pin env only)
  if (g.kind == 2)
    for (int a = 0; a < total; ++a) {
      bool last = true;
      for (int b = a + 1; b < total; ++b)
        if (pins[b].comp == pins[a].comp && pins[b].pin_id == pins[a].pin_id) { last = false; break; }
      if (last) pins[a].flags = kPinWinner;
    }
}

}  // namespace pcb
#endif
