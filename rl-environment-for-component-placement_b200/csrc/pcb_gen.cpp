// pcb_gen.cpp -- see pcb_gen.h.
#include "pcb_gen.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace pcb {
namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {  // SplitMix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  int randint(int lo, int hi) {  // inclusive bounds
    return lo + (int)(((unsigned __int128)next() * (uint64_t)(hi - lo + 1)) >> 64);
  }
  double normal(double mean, double sd) {
    double u1 = uniform(), u2 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    return mean + sd * std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
  int categorical(const std::vector<double>& p) {  // p sums to ~1
    double u = uniform(), acc = 0.0;
    int last = 0;
    for (size_t i = 0; i < p.size(); ++i) {
      if (p[i] <= 0.0) continue;
      acc += p[i];
      last = (int)i;
      if (u < acc) return (int)i;
    }
    return last;
  }
};

uint64_t mix(uint64_t a, uint64_t b) {
  Rng r(a ^ (b * 0xD6E8FEB86659FD93ull));
  r.next();
  return r.next();
}

}  // namespace

bool generate_bank(const int32_t g[12], int kind, int C, int max_pins, int max_nets, uint64_t seed,
                   int64_t count, GenBank& out, std::string& why) {
  const int min_w = g[0], max_w = g[1], min_h = g[2], max_h = g[3], min_c = g[4], max_c = g[5];
  const int min_nets = g[6], max_nets_cfg = g[7], lo = g[8], hi = g[9];
  const int nd = std::max(0, std::min(9, g[10])), ps = std::max(0, std::min(9, g[11]));
  if (min_w < 1 || min_h < 1 || min_w > max_w || min_h > max_h) { why = "component size range"; return false; }
  if (min_c < 1 || min_c > max_c || max_c > C) { why = "component count range"; return false; }
  const bool pins = kind >= 2;
  if (pins) {
    if (lo < 2 || lo > hi) { why = "pins per net range"; return false; }
    if (min_nets < 1 || min_nets > max_nets_cfg || max_nets_cfg > max_nets) { why = "net count range"; return false; }
    if (hi * max_nets_cfg > max_pins && max_pins < C * max_h * max_w) {
      // capacity is min(max_ppn*max_nets, C*P); anything smaller could overflow
      why = "max_pins smaller than max_num_pins_per_net * max_num_nets"; return false;
    }
    if (lo * min_nets > min_w * min_h * min_c) { why = "min pins exceed the minimum component area"; return false; }
  }
  const size_t n = (size_t)count;
  out.n_comp.assign(n, 0);
  out.comp_h.assign(n * C, 0);
  out.comp_w.assign(n * C, 0);
  if (pins) {
    out.n_pins.assign(n, 0); out.n_nets.assign(n, 0);
    out.pin_comp.assign(n * max_pins, 0); out.pin_rx.assign(n * max_pins, 0);
    out.pin_ry.assign(n * max_pins, 0); out.pin_net.assign(n * max_pins, 0);
    out.pin_id.assign(n * max_pins, 0);
  }
  for (size_t i = 0; i < n; ++i) {
    Rng r(mix(seed, (uint64_t)i));
    const int nc = r.randint(min_c, max_c);
    out.n_comp[i] = nc;
    std::vector<int> h(nc), w(nc), area(nc);
    int total_area = 0;
    for (int c = 0; c < nc; ++c) {
      h[c] = r.randint(min_h, max_h);
      w[c] = r.randint(min_w, max_w);
      area[c] = h[c] * w[c];
      total_area += area[c];
      out.comp_h[i * C + c] = (uint8_t)h[c];
      out.comp_w[i * C + c] = (uint8_t)w[c];
    }
    if (!pins) continue;
    int nets = std::min(r.randint(min_nets, max_nets_cfg), total_area / 2);
    int total = std::min(r.randint(lo * nets, hi * nets), total_area);
    // pins -> nets: softmax of normal logits, truncated sequential multinomial for the extras
    std::vector<double> prob(nets);
    double z = 0.0;
    for (int k = 0; k < nets; ++k) { prob[k] = std::exp(r.normal(1.0 / nets, 1.0 / (nd + 1))); z += prob[k]; }
    for (int k = 0; k < nets; ++k) prob[k] /= z;
    std::vector<int> extra(nets, 0);
    int todo = total - lo * nets;
    if (hi > lo && todo > 0) {
      int cap = std::min(hi - lo, todo);
      std::vector<double> q(nets);
      for (int t = 0; t < todo; ++t) {
        double zz = 0.0;
        for (int k = 0; k < nets; ++k) { q[k] = extra[k] < cap ? prob[k] : 0.0; zz += q[k]; }
        for (int k = 0; k < nets; ++k) q[k] /= zz;
        extra[r.categorical(q)] += 1;
      }
    } else {
      total = lo * nets;  // no extras possible (the reference leaves surplus pins unassigned)
    }
    // creation index of every pin, per net in net_pins[net] order
    std::vector<std::vector<int>> net_pins(nets);
    int nxt = 0;
    for (int k = 0; k < nets; ++k) { for (int j = 0; j < lo; ++j) net_pins[k].push_back(nxt++); }
    for (int k = 0; k < nets; ++k) { for (int j = 0; j < extra[k]; ++j) net_pins[k].push_back(nxt++); }
    total = nxt;
    // pins -> components
    int k0 = kind == 3 ? std::min((int)((ps / 10.0) * nc) + 1, nc)
                       : std::min(std::max((int)(((ps + 1) / 10.0) * nc), 1), nc);
    std::vector<std::pair<int, int>> space(nc);  // (component, free cells), dict order
    for (int c = 0; c < nc; ++c) space[c] = {c, area[c]};
    std::vector<int> comp_of(total, 0), row_id(total);
    std::iota(row_id.begin(), row_id.end(), 0);
    for (int k = 0; k < nets; ++k) {
      int left = (int)net_pins[k].size();
      std::stable_sort(space.begin(), space.end(),
                       [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.second > b.second; });
      int kk = k0 - 1, cap = 0;
      while (cap < left) {
        ++kk; cap = 0;
        for (int c = 0; c < std::min(kk, nc); ++c) cap += space[c].second;
        if (kk > nc && cap < left) { why = "internal: not enough component space"; return false; }
      }
      kk = std::min(kk, nc);
      int done = 0;
      while (left > 0) {
        int tot = 0;
        for (int c = 0; c < kk; ++c) tot += space[c].second;
        std::vector<double> p(kk);
        for (int c = 0; c < kk; ++c) p[c] = (double)space[c].second / tot;
        std::vector<int> cnt(kk, 0);
        for (int t = 0; t < left; ++t) cnt[r.categorical(p)] += 1;
        for (int c = 0; c < kk; ++c) {
          int take = std::min(cnt[c], space[c].second);
          space[c].second -= take;
          for (int j = 0; j < take; ++j) {
            int pin = net_pins[k][done + j];
            comp_of[pin] = space[c].first;
            if (kind == 2) row_id[pin] = j;
          }
          done += take;
          left -= take;
        }
      }
    }
    // net-major order, then relative cells per component without replacement
    out.n_pins[i] = total;
    out.n_nets[i] = nets;
    size_t base = i * (size_t)max_pins;
    int q = 0;
    for (int k = 0; k < nets; ++k)
      for (int pin : net_pins[k]) {
        out.pin_comp[base + q] = (uint8_t)comp_of[pin];
        out.pin_net[base + q] = (uint8_t)k;
        out.pin_id[base + q] = (uint16_t)row_id[pin];
        ++q;
      }
    for (int c = 0; c < nc; ++c) {
      std::vector<int> cells(area[c]);
      std::iota(cells.begin(), cells.end(), 0);
      for (int j = 0; j < total; ++j) {
        if (out.pin_comp[base + j] != c) continue;
        int pick = r.randint(0, (int)cells.size() - 1);
        int cell = cells[pick];
        cells.erase(cells.begin() + pick);
        out.pin_rx[base + j] = (uint8_t)(cell / w[c]);
        out.pin_ry[base + j] = (uint8_t)(cell % w[c]);
      }
    }
  }
  return true;
}

}  // namespace pcb
