// frame_simulator_gstim.h — the C++ shim of INTEGRATION.md §1, verbatim: what a Stim maintainer adds next to
// src/stim/simulators/frame_simulator_util.h to route the four batch drivers through libgstim.so.
// Compiled against the reference's own headers and objects by `make -C oracle shim` (test infrastructure) and exercised by
// tests/test_gpu_shim.py; the product (stim_b200/) does not include this file.
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>

#include "gstim.h"  // this repo: include/gstim.h, link with -lgstim
#include "stim/circuit/circuit.h"
#include "stim/io/stim_data_formats.h"
#include "stim/mem/simd_bits.h"

namespace stim {

inline void gstim_check(int code) {
    if (code == GSTIM_OK) {
        return;
    }
    std::string msg = gstim_last_error();
    if (code == GSTIM_ERR_INVALID_ARGUMENT) {
        throw std::invalid_argument(msg);  // same types the CPU path throws
    }
    if (code == GSTIM_ERR_OUT_OF_RANGE) {
        throw std::out_of_range(msg);
    }
    throw std::runtime_error(msg);
}

inline const char *gstim_format_name(SampleFormat fmt) {
    for (const auto &kv : format_name_to_enum_map()) {
        if (kv.second.id == fmt) {
            return kv.second.name;
        }
    }
    throw std::invalid_argument("unknown sample format");
}

struct GstimSampler {
    gstim_sampler *h = nullptr;
    GstimSampler(const Circuit &circuit, int mode, uint64_t seed, int device = 0) {
        std::string text = circuit.str();  // Stim's own printer; the library re-parses the same language
        gstim_check(gstim_create_from_text(text.data(), text.size(), mode, seed, device, &h));
    }
    GstimSampler(const GstimSampler &) = delete;
    ~GstimSampler() {
        gstim_destroy(h);
    }
};

// replaces sample_batch_detection_events_writing_results_to_disk (frame_simulator_util.h:67-77)
inline void gstim_detect_to_disk(
    const Circuit &c, size_t shots, bool prepend, bool append, FILE *out, SampleFormat fmt, uint64_t seed, FILE *obs_out,
    SampleFormat obs_fmt) {
    GstimSampler s(c, GSTIM_MODE_DETECTORS, seed);
    uint32_t flags = (prepend ? GSTIM_PREPEND_OBS : 0) | (append ? GSTIM_APPEND_OBS : 0);
    fflush(out);
    if (obs_out) {
        fflush(obs_out);
    }
    gstim_check(gstim_sample_detectors_to_fd(
        s.h, shots, flags, fileno(out), gstim_format_name(fmt), obs_out ? fileno(obs_out) : -1, gstim_format_name(obs_fmt)));
}

// replaces sample_batch_measurements_writing_results_to_disk (frame_simulator_util.h:123-130)
inline void gstim_sample_to_disk(
    const Circuit &c, const simd_bits<MAX_BITWORD_WIDTH> &ref, size_t shots, FILE *out, SampleFormat fmt, uint64_t seed) {
    GstimSampler s(c, GSTIM_MODE_MEASUREMENTS, seed);
    if (ref.num_bits_padded() > 0) {
        gstim_check(gstim_set_reference_sample(s.h, ref.u8, c.count_measurements()));  // TableauSimulator result, unchanged
    }
    fflush(out);
    gstim_check(gstim_sample_measurements_to_fd(s.h, shots, fileno(out), gstim_format_name(fmt)));
}

}  // namespace stim
