// tableau_ref.h — noiseless reference sample of a circuit (host side, runs once per compiled sampler).
//
// In the reference this is TableauSimulator::reference_sample_circuit
// (/root/reference/src/stim/simulators/tableau_simulator.inl:1435-1438): the circuit with all noise removed is
// simulated with a stabilizer simulator whose random measurement outcomes are biased to 0; the frame sampler's
// flips are XORed onto it. north_star keeps this step on the CPU ("gets the noiseless reference sample from the
// TableauSimulator on the CPU"); this file is the stand-alone host's own inverse-tableau simulator for it.
#pragma once
#include <cstdint>
#include <vector>

#include "circuit.h"

namespace gstim {

// One byte (0/1) per measurement result, in record order.
std::vector<uint8_t> reference_sample(const Circuit &circuit);

}  // namespace gstim
