// stim_gstim_main.cc — `stim detect` / `stim sample` with the reference's host code above the C ABI (TEST INFRASTRUCTURE).
//
// Everything that stays on the host in a real integration is the reference's own code, linked from the objects that
// oracle/Makefile builds out of /root/reference: the argument parser (stim/util_bot/arg_parse.h), Circuit::from_file, the
// printer Circuit::str(), the TableauSimulator reference sample, the seed handling. Only the two batch drivers are replaced
// by the shim of INTEGRATION.md §1 (frame_simulator_gstim.h), which calls libgstim.so. The flag handling below follows
// src/stim/cmd/command_detect.cc:26-79 and src/stim/cmd/command_sample.cc:27-75; tests/test_gpu_shim.py replays the
// deterministic cases of the reference's command tests through this binary.
#include <cstring>
#include <iostream>
#include <random>

#include "frame_simulator_gstim.h"
#include "stim/io/raii_file.h"
#include "stim/simulators/tableau_simulator.h"
#include "stim/util_bot/arg_parse.h"

using namespace stim;

static uint64_t seed_argument(int argc, const char **argv) {
    if (find_argument("--seed", argc, argv) != nullptr) {
        return (uint64_t)find_int64_argument("--seed", 0, 0, INT64_MAX, argc, argv);
    }
    std::random_device d;
    return ((uint64_t)d() << 32) ^ (uint64_t)d();
}

static int detect(int argc, const char **argv) {
    check_for_unknown_arguments(
        {"--seed", "--shots", "--append_observables", "--out_format", "--out", "--in", "--obs_out", "--obs_out_format"},
        {"--detect", "--prepend_observables"}, "detect", argc, argv);
    const auto &out_format = find_enum_argument("--out_format", "01", format_name_to_enum_map(), argc, argv);
    const auto &obs_out_format = find_enum_argument("--obs_out_format", "01", format_name_to_enum_map(), argc, argv);
    bool prepend = find_bool_argument("--prepend_observables", argc, argv);
    bool append = find_bool_argument("--append_observables", argc, argv);
    uint64_t shots = find_argument("--shots", argc, argv) ? (uint64_t)find_int64_argument("--shots", 1, 0, INT64_MAX, argc, argv) : 1;
    if (out_format.id == SampleFormat::SAMPLE_FORMAT_DETS && !append) {
        prepend = true;
    }
    RaiiFile in(find_open_file_argument("--in", stdin, "rb", argc, argv));
    RaiiFile out(find_open_file_argument("--out", stdout, "wb", argc, argv));
    RaiiFile obs_out(find_open_file_argument("--obs_out", stdout, "wb", argc, argv));
    if (obs_out.f == stdout) {
        obs_out.f = nullptr;
    }
    if (shots == 0) {
        return EXIT_SUCCESS;
    }
    Circuit circuit = Circuit::from_file(in.f);
    gstim_detect_to_disk(circuit, shots, prepend, append, out.f, out_format.id, seed_argument(argc, argv), obs_out.f, obs_out_format.id);
    return EXIT_SUCCESS;
}

static int sample(int argc, const char **argv) {
    check_for_unknown_arguments(
        {"--seed", "--skip_reference_sample", "--out_format", "--out", "--in", "--shots"}, {"--sample"}, "sample", argc, argv);
    const auto &out_format = find_enum_argument("--out_format", "01", format_name_to_enum_map(), argc, argv);
    bool skip_reference_sample = find_bool_argument("--skip_reference_sample", argc, argv);
    uint64_t shots = find_argument("--shots", argc, argv) ? (uint64_t)find_int64_argument("--shots", 1, 0, INT64_MAX, argc, argv) : 1;
    if (shots == 0) {
        return EXIT_SUCCESS;
    }
    RaiiFile in(find_open_file_argument("--in", stdin, "rb", argc, argv));
    RaiiFile out(find_open_file_argument("--out", stdout, "wb", argc, argv));
    Circuit circuit = Circuit::from_file(in.f);
    simd_bits<MAX_BITWORD_WIDTH> ref(0);
    if (!skip_reference_sample) {
        ref = TableauSimulator<MAX_BITWORD_WIDTH>::reference_sample_circuit(circuit);
    }
    gstim_sample_to_disk(circuit, ref, shots, out.f, out_format.id, seed_argument(argc, argv));
    return EXIT_SUCCESS;
}

int main(int argc, const char **argv) {
    try {
        if (argc >= 2 && strcmp(argv[1], "detect") == 0) {
            return detect(argc, argv);
        }
        if (argc >= 2 && strcmp(argv[1], "sample") == 0) {
            return sample(argc, argv);
        }
        std::cerr << "usage: stim_gstim detect|sample [stim's flags]\n";
        return EXIT_FAILURE;
    } catch (const std::invalid_argument &ex) {  // src/stim/main_namespaced.cc:113-122
        std::cerr << "\033[31m" << ex.what() << "\033[0m\n";
        return EXIT_FAILURE;
    } catch (const std::out_of_range &ex) {
        std::cerr << "\033[31m" << ex.what() << "\033[0m\n";
        return EXIT_FAILURE;
    } catch (const std::runtime_error &ex) {
        std::cerr << "\033[31m" << ex.what() << "\033[0m\n";
        return EXIT_FAILURE;
    }
}
