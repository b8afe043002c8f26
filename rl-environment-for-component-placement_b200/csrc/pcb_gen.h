// pcb_gen.h -- host-side instance generator (C++), used by pcb_generate_instances.
//
// Same construction as the reference's generate_instances
// (dummy_env_rectangular_pin.py:977-1264, 1462-1498; rect: dummy_env_rectangular.py:253-273;
// spatial differences: dummy_env_rectangular_pin_spatial.py:1102-1104, 1206-1208), with a
// counter-based RNG (one SplitMix64 stream per instance) instead of the reference's two
// process-global Mersenne-Twister streams.  Parity with the reference is therefore
// distributional, not stream-exact; stream-exact instances come from the oracle and are
// uploaded with pcb_upload_instances.
#ifndef PCB_GEN_H_
#define PCB_GEN_H_
#include <stdint.h>

#include <string>
#include <vector>

namespace pcb {

struct GenBank {
  std::vector<int32_t> n_comp, n_pins, n_nets;
  std::vector<uint8_t> comp_h, comp_w, pin_comp, pin_rx, pin_ry, pin_net;
  std::vector<uint16_t> pin_id;
};

// gen = {min_w, max_w, min_h, max_h, min_components, max_components, min_nets, max_nets,
//        min_pins_per_net, max_pins_per_net, net_distribution, pin_spread}
bool generate_bank(const int32_t gen[12], int kind, int C, int max_pins, int max_nets,
                   uint64_t seed, int64_t count, GenBank& out, std::string& why);

}  // namespace pcb
#endif
