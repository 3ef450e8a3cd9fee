// lowering.h — Circuit -> flat instruction stream (program.h) for the sm_100a interpreter.
//
// No reference analogue (Stim re-walks the Circuit per batch, frame_simulator.inl:166-170).
// What the lowering does, and the reference behaviour each step preserves:
//   * REPEAT unrolled in execution order           (circuit.h:181-193 for_each_operation)
//   * MPP / SPP / MXX,MYY,MZZ decomposed            (gate_decomposition.cc:88-274, frame_simulator.inl:842-902)
//   * PAULI_CHANNEL_1/2 folded to one site + choice (distribution of tableau_simulator.h:291-324)
//   * rec[-k] resolved to absolute indices, bad lookback -> std::out_of_range (measure_record_batch.inl:83-94)
//   * bit-as-target of CX/CY -> std::invalid_argument (frame_simulator.inl:399-402, 421-424)
//   * inverted targets ignored, sweep controls are no-ops (frame_simulator.inl:146-148, 176)
//   * qubit indices compacted to the set of qubits that appear in the circuit
// RNG addressing (site / collapse-site / record counters) is part of the output; DESIGN.md §RNG.
#pragma once
#include <cstdint>
#include <vector>

#include "circuit.h"
#include "program.h"

namespace gstim {

struct Batch {
    uint32_t op = 0, flags = 0, aux = 0, extra = 0;
    uint64_t lambda = 0;  // event rate per shot, fixed point (2^-56 nat)
    uint32_t site0 = 0, csite0 = 0, rec0 = 0;
    uint32_t t1 = 0, t2 = 0, t3 = 0;
    uint32_t n_items = 0;
    std::vector<uint32_t> payload;  // op specific (see program.h)
    // XORROWS is assembled from these three at serialisation time:
    std::vector<uint32_t> dst, off, idx;
    // resources (for the hazard pass): per item, [begin,end) into res; bit31 of an entry = write.
    std::vector<uint32_t> res_off;
    std::vector<uint32_t> res;
    uint32_t words() const;
};

// Filled by serialize_program (see program.h "Noise schedule").
struct NoiseSchedule {
    std::vector<uint32_t> info;       // GSTIM_NOISE_INFO_WORDS per noise batch
    std::vector<uint32_t> n_sites;    // per noise batch
    std::vector<uint64_t> lams;       // per noise batch (fixed-point rate)
    std::vector<uint64_t> rates;      // distinct rates: 2 words each, lam and floor((2^64 - 1) / lam)
    std::vector<uint32_t> slices;     // GSTIM_SLICE_WORDS per RNG slice (program.h "Noise schedule"), in program order
};

struct LoweredCircuit {
    CircuitStats stats;
    uint32_t mode = 0;             // 0 detectors, 1 measurements
    uint32_t num_qubits = 0;       // compacted
    uint32_t rec_ring = 0;
    uint32_t num_resources = 0;
    std::vector<uint32_t> qubit_map;  // original index -> compact (logical) index or UINT32_MAX
    std::vector<uint32_t> logical_of; // physical frame row -> logical index (size Q+1; [Q] = Q, the global clock)
    std::vector<Batch> batches;
    NoiseSchedule noise;
    uint32_t max_items = 0;
    uint64_t total_items = 0;
    uint64_t num_sites = 0, num_csites = 0;
    // with_sweep lowering only (m2d.cu): absolute measurement indices XORed into every detector / observable (output id
    // d, or D + l), with multiplicity
    std::vector<std::vector<uint64_t>> out_recs;
};

// Probability -> rate key of the detector-error-model sampler's gap arithmetic (dem.cu): bit 63 = valid, (INV << 8) | SH.
uint64_t gstim_rate_key(double p);

// Pass 1: semantics. Throws std::invalid_argument / std::out_of_range like the reference would.
// with_sweep: sweep-controlled Paulis become GOP_SWEEP batches (instead of being dropped: there is no sweep data when
// sampling, frame_simulator.inl:146-148) and out_recs is filled; such a lowering is for response.cc / m2d.cu only and must
// not be serialised for the interpreter.
LoweredCircuit lower_circuit(const Circuit &c, uint32_t mode, uint32_t max_batch_words, bool with_sweep = false);

// Lowers a fragment of a circuit for the interactive simulator (flipsim.cu): no implicit reset at the start, qubit index =
// frame row (no compaction; rows [0, num_qubits)), measurement results numbered from meas0 with ABSOLUTE record indices,
// detectors numbered from 0 within the fragment, observable l at output row (fragment's detector count + l), no detector
// fusion and no layout post-pass. Batches keep the "items of a batch touch disjoint resources" property.
LoweredCircuit lower_fragment(const Circuit &c, uint32_t num_qubits, uint64_t meas0);

// Pass 2: hazard analysis for `slots` concurrent thread groups + serialisation into chunks.
std::vector<uint32_t> serialize_program(LoweredCircuit &lc, uint32_t slots, uint32_t chunk_words, GstimPlan *plan);

}  // namespace gstim
