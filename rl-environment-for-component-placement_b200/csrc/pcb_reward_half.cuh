// reward_half_kernel: the terminal reward of the centroid route (find_reward, pin:832-975) with
// TWO jobs per warp -- lanes 0-15 one job, lanes 16-31 the next.  A job of a 32 x 32 board has
// ~20 pins, ~150 segment pairs and ~60 pairs past the box filter, so a whole warp per job runs
// most of its phases with a third of the lanes, and what the overlapped reward kernel costs
// the step kernel beside it is the time its warps stay resident.  Same routines as
// reward_kernel (pcb_kernels.cuh), restated for groups of 16 lanes: control flow is warp-uniform
// (loop bounds are the maximum over the two jobs, every vote / shuffle / barrier is issued by
// the whole warp with the full mask -- a run-time lane mask would cost a MATCH.ANY per
// operation), the per-job state (counters, queues, scratch) is per group.
// Used for the centroid reward when max_pins <= 32; anything else takes reward_kernel.
#ifndef PCB_REWARD_HALF_CUH
#define PCB_REWARD_HALF_CUH

#ifndef PCB_REWARD_HALF
#define PCB_REWARD_HALF 1
#endif

namespace pcb {

// Every routine below is a template on kG = lanes per job: 16 (two jobs per warp) or 8 (four).
// The per-step kernel runs 16: with 8 lanes the jobs of a step of 32 x 32 boards take longer than
// the pin env's step kernel they hide behind (C 0.850 -> 0.775); beside the spatial env's step
// kernel, which lasts twice as long, 8 lanes give +1.4 % on the device-timed rollout but -3.4 %
// through pcb_steps_host, where each reward kernel is followed by its copies to the host
// (profiles/r2h_reward_group_spatial_ab.log), so that stays 16 as well.  The job-list kernel of
// the multi-step path is bound by instruction issue, not by the latency of one job, and runs 8
// (reward_list_group_lanes; profiles/r2h_reward_list_group8_ab.log).
#ifndef PCB_REWARD_GROUP
#define PCB_REWARD_GROUP 16
#endif
#ifndef PCB_REWARD_GROUP_SPATIAL
#define PCB_REWARD_GROUP_SPATIAL 16
#endif
#ifndef PCB_REWARD_LIST_GROUP_SMALL
#define PCB_REWARD_LIST_GROUP_SMALL 8
#endif
static_assert(PCB_REWARD_GROUP == 16 || PCB_REWARD_GROUP == 8, "PCB_REWARD_GROUP");
static_assert(PCB_REWARD_GROUP_SPATIAL == 16 || PCB_REWARD_GROUP_SPATIAL == 8, "PCB_REWARD_GROUP_SPATIAL");
static_assert(PCB_REWARD_LIST_GROUP_SMALL == 16 || PCB_REWARD_LIST_GROUP_SMALL == 8, "PCB_REWARD_LIST_GROUP_SMALL");
// lanes per job of reward_half_kernel (kind 3 = the spatial pin env)
inline int reward_step_group_lanes(int kind) { return kind == 3 ? PCB_REWARD_GROUP_SPATIAL : PCB_REWARD_GROUP; }
// lanes per job of reward_list_half_kernel
#ifndef PCB_REWARD_LIST_SMALL_MAX_PINS
#define PCB_REWARD_LIST_SMALL_MAX_PINS 32
#endif
inline int reward_list_group_lanes(int max_pins) {
  return max_pins <= PCB_REWARD_LIST_SMALL_MAX_PINS ? PCB_REWARD_LIST_GROUP_SMALL : 16;
}

template <int kG>
__device__ __forceinline__ unsigned grp_ballot(bool pred, int gbase) {
  return (__ballot_sync(kFull, pred) >> gbase) & ((1u << kG) - 1u);
}
template <int kG>
__device__ __forceinline__ int grp_sum(int v) {
  if (kG > 8) v += __shfl_xor_sync(kFull, v, 8);
  v += __shfl_xor_sync(kFull, v, 4);
  v += __shfl_xor_sync(kFull, v, 2);
  v += __shfl_xor_sync(kFull, v, 1);
  return v;
}

// route_prepare for a group; np_u / nn_u = the larger pin / net count of the two jobs
template <int kG>
__device__ __forceinline__ void route_prepare_half(const JobPin* pins, const RouteScratch& rs, int gl,
                                                   int n_pins, int n_nets, int np_u, int nn_u) {
  for (int q0 = 0; q0 < np_u; q0 += kG) {
    const int q = q0 + gl;
    if (q < n_pins) {
      const JobPin jp = pins[q];
      rs.px[q] = jp.x;
      rs.py[q] = jp.y;
      rs.net[q] = jp.net;
      if (q == 0 || pins[q - 1].net != jp.net) rs.net_off[jp.net] = (uint16_t)q;
    }
  }
  if (gl == 0 && n_pins > 0) rs.net_off[n_nets] = (uint16_t)n_pins;
  __syncwarp();
  for (int n0 = 0; n0 < nn_u; n0 += kG) {
    const int net = n0 + gl;
    if (net < n_nets) {
      const int lo = rs.net_off[net], hi = rs.net_off[net + 1];
      int sx = 0, sy = 0;
      for (int q = lo; q < hi; ++q) { sx += rs.px[q]; sy += rs.py[q]; }
      rs.cx[net] = (double)sx / (double)(hi - lo);
      rs.cy[net] = (double)sy / (double)(hi - lo);
      rs.nsum[2 * net] = sx;
      rs.nsum[2 * net + 1] = sy;
    }
  }
  __syncwarp();
}

// count_crossings for a group of 16 lanes (see count_crossings: dense pair enumeration, integer
// box filter, exact integer test, float64 formula for what that leaves undecided)
template <int kG>
__device__ __forceinline__ int count_crossings_half(const RouteScratch& rs, int gl, int gbase,
                                                    int n_nets) {
  int cnt = 0, qn = 0, hn = 0;
  const unsigned below = (1u << gl) - 1u;
  auto heavy = [&](uint32_t pr) -> int {
    const int s = pr >> 16, t = pr & 0xffffu;
    return segments_cross((double)rs.px[s], (double)rs.py[s], rs.x2[s], rs.y2[s],
                          (double)rs.px[t], (double)rs.py[t], rs.x2[t], rs.y2[t]) ? 1 : 0;
  };
  // the first n (<= 16, per group; 0 = this group sits the round out) entries of queue[qn ...)
  auto decide = [&](int n) {
    uint32_t pr = 0;
    int r = 0;
    if (gl < n) {
      pr = rs.queue[qn + gl];
      r = segments_cross_int(rs, pr);
    }
    cnt += r & 1;
    const unsigned hb = grp_ballot<kG>(r == 2, gbase);
    if (r == 2) rs.hqueue[hn + __popc(hb & below)] = pr;
    hn += __popc(hb);
    __syncwarp();
    const bool take = hn >= kG;
    if (__any_sync(kFull, take)) {
      if (take) {
        hn -= kG;
        cnt += heavy(rs.hqueue[hn + gl]);
      }
      __syncwarp();
    }
  };
  const int nn_u = __reduce_max_sync(kFull, n_nets);
  for (int b = 1; b < nn_u; ++b) {
    int first = 0, total = 0;
    if (b < n_nets) {
      first = rs.net_off[b];
      total = first * (rs.net_off[b + 1] - first);
    }
    const int tot_u = __reduce_max_sync(kFull, total);
    const float inv = __fdividef(1.0f, (float)max(first, 1));
    for (int i0 = 0; i0 < tot_u; i0 += kG) {
      const bool live = i0 + gl < total;
      const int i = live ? i0 + gl : 0;
      const int q = __float2int_rz(((float)i + 0.5f) * inv);
      const int s = i - q * first, t = live ? first + q : 0;
      const bool pass = boxes_touch(rs.box[t], rs.box[s]) && live;
      const unsigned pb = grp_ballot<kG>(pass, gbase);
      if (pass) rs.queue[qn + __popc(pb & below)] = ((uint32_t)s << 16) | (uint32_t)t;
      qn += __popc(pb);
      __syncwarp();
      const bool full = qn >= kG;
      if (__any_sync(kFull, full)) {
        int n = 0;
        if (full) { qn -= kG; n = kG; }
        decide(n);
        __syncwarp();
      }
    }
  }
  {
    const int n = qn;
    qn = 0;
    if (__any_sync(kFull, n > 0)) decide(n);
  }
  if (gl < hn) cnt += heavy(rs.hqueue[gl]);
  __syncwarp();
  return grp_sum<kG>(cnt);
}

// centroid_route for a group
template <int kG>
__device__ __forceinline__ void centroid_route_half(const Dims& d, const RouteScratch& rs, int gl,
                                                    int gbase, int n_pins, int n_nets, int np_u,
                                                    int& crossings, double& wl) {
  for (int q0 = 0; q0 < np_u; q0 += kG) {
    const int q = q0 + gl;
    if (q < n_pins) {
      const int net = rs.net[q];
      const int lo = rs.net_off[net], cnt = rs.net_off[net + 1] - lo;
      double x2 = 0.0, y2 = 0.0, len = 0.0;
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
            put_segment(rs, q, 0, 0, 0, 0, 0);
          box = make_box(ix, iy, x2, y2);
        }
        len = norm2((double)ix - x2, (double)iy - y2, d.norm_fma);
      }
      rs.x2[q] = x2; rs.y2[q] = y2; rs.box[q] = box; rs.dist[q] = len;
    }
  }
  __syncwarp();
  crossings = count_crossings_half<kG>(rs, gl, gbase, n_nets);
  // find_wirelength (pin:741-759): sequential float64 sum over the pins of this group's job
  double acc = 0.0;
  for (int q = 0; q < np_u; ++q)
    if (q < n_pins) acc += rs.dist[q];
  wl = acc;
  __syncwarp();
}

// beam_route (pcb_beam.cuh) for a group: crossings + wirelength of the paths beam_kernel left in the
// job record; the segment headed by pin q is (q, nxt[q])
template <int kG>
__device__ __forceinline__ void beam_route_half(const Dims& d, const RouteScratch& rs, int gl, int gbase,
                                                int n_pins, int n_nets, int np_u, int& crossings,
                                                double& wl) {
  for (int q0 = 0; q0 < np_u; q0 += kG) {
    const int q = q0 + gl;
    if (q < n_pins) {
      const int nx = rs.nxt[q];
      double x2 = 0.0, y2 = 0.0, len = 0.0;
      uint2 box = no_box();
      if (nx != 0xFFFF) {
        const int ix = rs.px[q], iy = rs.py[q], jx = rs.px[nx], jy = rs.py[nx];
        x2 = (double)jx; y2 = (double)jy;
        put_segment(rs, q, ix, iy, jx, jy, 1);
        box = make_box(ix, iy, jx, jy);
        len = norm2((double)ix - x2, (double)iy - y2, d.norm_fma);
      }
      rs.x2[q] = x2; rs.y2[q] = y2; rs.box[q] = box; rs.dist[q] = len;
    }
  }
  __syncwarp();
  crossings = count_crossings_half<kG>(rs, gl, gbase, n_nets);
  // find_wirelength: nets in order, segments in path order, sequential float64 sum (the first lane
  // of each group walks its job's paths)
  double acc = 0.0;
  if (gl == 0) {
    for (int net = 0; net < n_nets; ++net) {
      int q = rs.start[net];
      while (rs.nxt[q] != 0xFFFF) {
        acc += rs.dist[q];
        q = rs.nxt[q];
      }
    }
  }
  wl = __shfl_sync(kFull, acc, gbase);
  __syncwarp();
}

struct GroupReward { double reward, wl, ni; };

// terminal reward of this group's job (has == false: the group sits the round out but takes part
// in every warp-wide operation).  BEAM: reward types "beam" / "both" as well, the paths routed by
// beam_kernel are in the job record.
template <int kG, bool BEAM = false>
__device__ __forceinline__ GroupReward reward_of_group(const Dims& d, const RouteScratch& rs,
                                                       const uint8_t* job, const RouteJob& hdr, bool has,
                                                       int gl, int gbase) {
  const int n_pins = has ? hdr.n_pins : 0, n_nets = has ? hdr.n_nets : 0;
  const int np_u = __reduce_max_sync(kFull, n_pins), nn_u = __reduce_max_sync(kFull, n_nets);
  route_prepare_half<kG>(reinterpret_cast<const JobPin*>(job + 32), rs, gl, n_pins, n_nets, np_u, nn_u);
  int crossings = 0;
  double length = 0.0;
  if (!BEAM || d.reward_type == 0) {
    centroid_route_half<kG>(d, rs, gl, gbase, n_pins, n_nets, np_u, crossings, length);
  } else if constexpr (BEAM) {
    const uint16_t* jn = reinterpret_cast<const uint16_t*>(job + job_path_offset(d.max_pins));
    for (int q = gl; q < n_pins; q += kG) rs.nxt[q] = jn[q];
    for (int k = gl; k < n_nets; k += kG) rs.start[k] = jn[d.max_pins + k];
    __syncwarp();
    beam_route_half<kG>(d, rs, gl, gbase, n_pins, n_nets, np_u, crossings, length);
    if (d.reward_type == 2) {  // "both": fewer crossings wins, beam on ties (pin:643-661)
      int c2; double l2;
      centroid_route_half<kG>(d, rs, gl, gbase, n_pins, n_nets, np_u, c2, l2);
      if (c2 < crossings) { crossings = c2; length = l2; }
    }
  }
  GroupReward r;
  r.wl = length / d.wl_norm;             // pin:921 / 940 / 959-961
  r.ni = (double)crossings / d.ni_norm;  // pin:922-924
  r.reward = -1.0 * (d.ww * r.wl + d.wi * r.ni);
  return r;
}

// the record of a group's next job: on its way to L1 while the current one is computed
template <int kG>
__device__ __forceinline__ void prefetch_job(const Dims& d, const uint8_t* jobs, uint32_t jn, uint32_t count,
                                             int gl) {
  const int line = gl * 128;
  if (jn < count && line < job_path_offset(d.max_pins))
    asm volatile("prefetch.global.L1 [%0];" ::"l"(jobs + (size_t)jn * job_stride(d.max_pins, d.K) + line));
}

template <int kG>
__global__ void __launch_bounds__(kWarpsPerCta * 32, PCB_REWARD_MIN_BLOCKS)
reward_half_kernel(const __grid_constant__ Params p) {
  constexpr int kGroups = 32 / kG;
  extern __shared__ __align__(16) uint8_t smem[];
  const Dims& d = p.d;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int gl = lane & (kG - 1), grp = lane / kG, gbase = lane & ~(kG - 1);
  const uint32_t gw = blockIdx.x * kWarpsPerCta + wic, nw = gridDim.x * kWarpsPerCta;
  const uint32_t count = p.job_count[p.step_no & 3];
  // counters form a ring of 4: step t+2 cannot start before this kernel has finished
  if (gw == 0 && lane == 0) p.job_count[(p.step_no + 2) & 3] = 0;
  const int scr_bytes = (route_scratch_bytes(d.max_pins, d.K) + 15) & ~15;
  RouteScratch rs = route_scratch(smem + (size_t)(wic * kGroups + grp) * scr_bytes, d.max_pins, d.K);
  for (uint32_t j0 = gw * kGroups; j0 < count; j0 += nw * kGroups) {
    const uint32_t j = j0 + (uint32_t)grp;
    const bool has = j < count;
    const uint8_t* job = p.jobs + (size_t)(has ? j : j0) * job_stride(d.max_pins, d.K);
    const RouteJob hdr = *reinterpret_cast<const RouteJob*>(job);
    prefetch_job<kG>(d, p.jobs, j + nw * kGroups, count, gl);
    const GroupReward r = reward_of_group<kG>(d, rs, job, hdr, has, gl, gbase);
    if (has && gl == 0) {
      const int env = (int)hdr.env;
      const double ret = hdr.ep_return + r.reward;
      p.out.reward[env] = r.reward; p.out.wl[env] = r.wl; p.out.ni[env] = r.ni;
      p.out.ep_return[env] = ret;
      p.out.ep_length[env] = (int32_t)hdr.ep_len;
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
