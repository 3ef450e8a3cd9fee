// response.h — the propagated-response table of a lowered circuit (host side of the event-driven engine, sparse.cu).
//
// A Pauli-frame simulation is linear over GF(2): every output bit (detector, observable, or measurement flip) is the
// XOR of the contributions of the random bits the run drew — which noise site fired with which Pauli, which collapse
// randomisation bit was set (/root/reference/src/stim/simulators/frame_simulator.inl:173-317 only ever XORs frame rows
// into frame rows, record rows and outputs). The reference evaluates that linear map once per shot by pushing dense
// frame words through every gate. For rare noise almost all of those words are zero, so this engine evaluates the map
// the other way round: the response of every (noise site, outcome) — the set of output bits it flips — is computed once
// per circuit by propagating sensitivities BACKWARDS through the lowered program, and a shot is the XOR of the responses
// of the events that fired in it (same sampling distribution as the reference: every site is an independent
// Bernoulli(p) with the channel's own outcome distribution, exactly like RareErrorIterator over targets x shots,
// /root/reference/src/stim/util_bot/probability_util.cc:23-43).
//
// The backward pass keeps, per frame row, the sets SX[q] / SZ[q] of output bits that an X / Z flip of q at the current
// point of the program would flip, and per record slot the set SR[m]; it walks lc.batches in reverse:
//   gate with GF(2) matrix A      S_before = A^T S_after
//   MEASURE (m = x, z <- random)  SX ^= SR[slot] (^ fused detector), SZ <- {} ; the random bit's own response is SZ_after
//   XORROWS / OBS_PAULI / FEEDBACK add the destination's sensitivity to their sources
//   NOISE / CORR                  snapshot: response(outcome) = XOR of the flipped components' sets
// Collapse randomisation bits that reach an output (non-deterministic detectors, measurement sampling) become sites
// with p = 1/2; an E / ELSE_CORRELATED_ERROR chain becomes one site whose outcomes are its elements. A circuit is
// ELIGIBLE unless a chain has more than 16 elements or the table would be too large; then the interpreter (interp.cu)
// samples it.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "lowering.h"

namespace gstim {

constexpr uint32_t RESP_NONE = 0xFFFFFFFFu;      // empty slot of a table entry
constexpr uint32_t RESP_OVERFLOW = 0x80000000u;  // last slot: 0x80000000 | offset into `overflow` (count, then ids)

enum RespKind : uint32_t {
    RK_SINGLE = 0,   // one outcome
    RK_UNIFORM = 1,  // outcome = mulhi(v, n_out)            (DEPOLARIZE2: frame_simulator.inl:651-659)
    RK_THRESH = 2,   // outcome = number of thresholds <= v  (PAULI_CHANNEL_*, DEPOLARIZE1, HERALDED_*)
    RK_CHAIN = 3,    // same chooser as RK_THRESH; the outcomes are the elements of an E / ELSE_CORRELATED_ERROR chain
                     // (outcome i = noise group site_group + outcome_word[i] fired alone)
};

struct RespClass {
    uint64_t lam = 0;            // event rate of the class, fixed point (lowering.cc rate_of)
    uint32_t inv = 0, sh = 0;    // gap arithmetic: gap = (E * inv) >> sh  (E = Exp(1) in units of 2^-26 nat); inv = 0: always fires
    uint32_t kind = RK_SINGLE;
    uint32_t n_out = 1;          // outcomes per site (<= 16)
    uint32_t thr[15] = {0};      // RK_THRESH: ascending upper bounds of outcomes 0 .. n_out - 2 on a uniform u32
    uint32_t n_sites = 0;
    uint32_t entry0 = 0;         // table entry of (site s, outcome o) = entry0 + s * n_out + o
    // Dense classes (p large, e.g. the p = 1/2 collapse bits): the trials are drawn 32 at a time as packed Bernoulli words,
    // P(bit) = dense_thr / 2^32 exactly, built from 32 - ctz(dense_thr) random words by the binary expansion of the
    // probability (the bit-sliced part of biased_randomize_bits, /root/reference/src/stim/util_bot/probability_util.cc:74-132).
    // 0 = walk the class with geometric gaps.
    uint32_t dense_thr = 0;
};

struct ResponseTable {
    bool eligible = false;
    std::string why_not;             // reason when not eligible
    uint32_t n_outputs = 0;          // D + L (detector mode) or M (measurement mode)
    std::vector<RespClass> classes;
    std::vector<uint32_t> entries;   // 4 words per entry: output ids ascending, RESP_NONE padded; > 4 ids: 3 ids + overflow link
    std::vector<uint32_t> overflow;  // [count, id, id, ...] blocks
    // provenance of every site, in class-major table order (tests re-derive the responses by forward injection):
    std::vector<uint32_t> site_group;  // noise group (Philox word 0 of the interpreter's stream for it)
    std::vector<uint32_t> site_index;  // index of the site inside its group
    // outcome -> representative Pauli word v of the interpreter's chooser, per class (n_out values each), for the same tests
    std::vector<uint32_t> outcome_word;
    // with_sweep lowerings: output ids flipped by sweep bit k (XOR over every Pauli it controls), for m2d.cu
    std::vector<std::vector<uint32_t>> sweep_responses;
    uint64_t n_sites = 0, n_entries = 0;
    uint32_t max_response = 0;       // largest number of output bits of one entry
    double events_per_shot = 0;      // expected number of events per shot
    double flips_per_shot = 0;       // expected number of output bit flips per shot (before cancellation)
};

// Builds the table from the lowered batches (before or after serialisation; only lc.batches / lc.mode / lc.rec_ring are read).
// keep_conjugate: the frame simulator of the measurement converter (m2d.cu) runs with frame randomisation off, where a
// measurement / reset KEEPS the conjugate frame component instead of replacing it by random bits (the `if
// (guarantee_anticommutation_via_frame_randomization)` branches of frame_simulator.inl:173-317); no collapse sites then.
ResponseTable build_response_table(const LoweredCircuit &lc, bool keep_conjugate = false);

// Round structure of a class: entries(site s + p) == entries(site s) with `delta` added to every detector id (< D) and the
// observables unchanged, for all s in [a, a + n - p). Unrolled REPEAT blocks of QEC circuits have it (p = sites of the class
// per round, delta = detectors per round); the device table then stores head + ONE period + tail (sparse.cu PERIODIC). p = 0:
// no such structure covering at least three periods.
struct ResponsePeriod {
    uint32_t a = 0, p = 0, n = 0, delta = 0;
};
ResponsePeriod find_response_period(const ResponseTable &rt, const RespClass &c, uint32_t D);

}  // namespace gstim
