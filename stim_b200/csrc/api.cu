// api.cu — the C ABI (include/gstim.h) and the host scheduler behind it.
//
// The host scheduler replaces the batching loops of the reference
// (/root/reference/src/stim/simulators/frame_simulator_util.inl:22-35, 207-288, 291-372):
// it picks the per-block shot count from the shared-memory budget, cuts a request into chunks whose
// bit-major staging table fits a fixed HBM budget, launches interpreter + transposer per chunk on
// one stream, and drains results to the caller.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <exception>
#include <string>
#include <sys/mman.h>
#include <thread>
#include <vector>

#include "../../include/gstim.h"
#include "circuit.h"
#include "kernels.cuh"
#include "lowering.h"
#include "hostpipe.h"
#include "sparse.cuh"
#include "tableau_ref.h"
#include "writers.h"

using namespace gstim;

namespace {

thread_local std::string g_last_error = "";

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};
struct OomError : std::runtime_error {
    explicit OomError(const std::string &m) : std::runtime_error(m) {}
};
struct IoError : std::runtime_error {
    explicit IoError(const std::string &m) : std::runtime_error(m) {}
};

#define CK(expr)                                                                                          \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess) {                                                                          \
            if (_e == cudaErrorMemoryAllocation) {                                                        \
                cudaGetLastError();                                                                       \
                throw OomError(std::string("CUDA out of memory in ") + #expr);                            \
            }                                                                                             \
            throw CudaError(std::string("CUDA error '") + cudaGetErrorString(_e) + "' in " + #expr);      \
        }                                                                                                 \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) {
            return;
        }
        if (p) {
            cudaFree(p);
            p = nullptr;
            cap = 0;
        }
        CK(cudaMalloc(&p, bytes));
        cap = bytes;
    }
    ~DevBuf() {
        if (p) {
            cudaFree(p);
        }
    }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) {
            return;
        }
        if (p) {
            cudaFreeHost(p);
            p = nullptr;
            cap = 0;
        }
        CK(cudaMallocHost(&p, bytes));
        cap = bytes;
    }
    ~PinnedBuf() {
        if (p) {
            cudaFreeHost(p);
        }
    }
};

uint32_t env_u32(const char *name, uint32_t dflt) {
    const char *v = getenv(name);
    if (v == nullptr || *v == '\0') {
        return dflt;
    }
    return (uint32_t)strtoul(v, nullptr, 10);
}

}  // namespace

struct gstim_sampler {
    HostStager file_stage;  // page-locked staging pair of the file-output path (hostpipe.h)
    int device = 0;
    int mode = 0;
    uint64_t seed = 0;
    uint64_t next_col = 0;  // global 128-shot column index the next call starts at

    Circuit circuit;
    LoweredCircuit lc;
    GstimPlan plan{};
    std::vector<uint32_t> words;

    // launch configuration
    uint32_t threads = 0, pre_threads = 0, G_log2 = 0, slots = 0, K_max = 0, chunk_words = 0;
    int num_sms = 0;
    size_t smem_optin = 0;

    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    DevBuf d_dbg, d_prog, d_table, d_rec, d_rowmap, d_stage[2], d_counts;
    // noise schedule + per-CTA event scratch (interp.cu noise_prepass)
    DevBuf d_noise_info, d_rates, d_slices, d_segoff, d_ev_counts, d_ev_buf, d_ev_overflow;
    uint32_t segoff_K = 0;
    uint32_t n_noise = 0;
    uint64_t ev_total = 0;
    PinnedBuf h_stage[2];
    cudaEvent_t stage_done[2] = {nullptr, nullptr};
    cudaEvent_t stage_ready[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> events;
    std::vector<uint8_t> ref_bits;  // one byte per measurement (0/1)

    // event-driven engine (sparse.cu): present when the circuit's response table exists (response.h)
    std::unique_ptr<SparseEngine> sparse;
    std::string sparse_why;     // why the circuit is not eligible ("" when it is)
    std::string interp_why;     // why the interpreter cannot run the circuit ("" when it can)
    int engine_pref = GSTIM_ENGINE_AUTO;
    int last_engine = GSTIM_ENGINE_INTERPRETER;
    bool sparse_favoured = false;  // the cost model's choice for GSTIM_ENGINE_AUTO

    // sparse host delivery (sample_to_host_sparse): three pipeline slots
    struct SparseSlot {
        DevBuf d_rows, d_strm, d_idx, d_ctl;   // dense rows (+ obs rows), record stream, per-shot offsets, cursor + overflow flag
        PinnedBuf h_strm, h_idx, h_ctl, h_obs;
        cudaEvent_t ev_kernels = nullptr, ev_copy = nullptr;
        uint64_t first = 0, n = 0;
        int state = 0;  // 0 free, 1 kernels enqueued, 2 copy enqueued
        ~SparseSlot() {
            if (ev_kernels) {
                cudaEventDestroy(ev_kernels);
                cudaEventDestroy(ev_copy);
            }
        }
    } sparse_slot[3];
    int last_d2h_sparse = 0;

    // multi-device sampler (gstim_create_from_text_multi): samplers of the other devices; host-output calls split their
    // shots over all of them
    std::vector<gstim_sampler *> peers;

    uint64_t last_launches = 0;
    uint32_t last_K = 0;
    uint32_t K_fixed = 0;  // gstim_set_block_columns: 0 = chosen per call from the shot count
    float last_interp_ms = 0, last_transpose_ms = 0, last_call_ms = 0;
    cudaEvent_t call_start = nullptr, call_end = nullptr;

    ~gstim_sampler() {
        for (gstim_sampler *p : peers) {
            cudaSetDevice(p->device);
            delete p;
        }
        cudaSetDevice(device);
        for (auto e : events) {
            cudaEventDestroy(e);
        }
        for (auto e : stage_done) {
            if (e) {
                cudaEventDestroy(e);
            }
        }
        for (auto e : stage_ready) {
            if (e) {
                cudaEventDestroy(e);
            }
        }
        if (call_start) {
            cudaEventDestroy(call_start);
            cudaEventDestroy(call_end);
        }
        if (stream) {
            cudaStreamDestroy(stream);
        }
        if (copy_stream) {
            cudaStreamDestroy(copy_stream);
        }
    }
};

namespace {

uint32_t n_rows_of(const gstim_sampler *s) {
    return s->mode == GSTIM_MODE_DETECTORS ? s->plan.num_det + s->plan.num_obs : s->plan.num_meas;
}

// Largest single-item payload in the circuit decides the minimum chunk size.
uint32_t max_targets_in_one_item(const Circuit &c) {
    uint32_t m = 0;
    for (const auto &op : c.ops) {
        if (op.gate->cat == GateCat::REPEAT) {
            m = std::max(m, max_targets_in_one_item(c.blocks[op.block_index]));
        } else if (
            op.gate->cat == GateCat::CORR || op.gate->cat == GateCat::DETECTOR || op.gate->cat == GateCat::OBSERVABLE_INCLUDE) {
            m = std::max<uint32_t>(m, (uint32_t)op.targets.size());
        }
    }
    return m;
}

void configure(gstim_sampler *s) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, s->device));
    s->num_sms = prop.multiProcessorCount;
    s->smem_optin = prop.sharedMemPerBlockOptin;

    // chunk size: 2048 words unless a single item needs more
    uint32_t need = 4 * (max_targets_in_one_item(s->circuit) + 64);
    uint32_t chunk = 2048;
    while (chunk < need) {
        chunk <<= 1;
    }
    s->chunk_words = env_u32("GSTIM_CHUNK_WORDS", chunk);

    s->lc = lower_circuit(s->circuit, (uint32_t)s->mode, s->chunk_words - GSTIM_HDR_WORDS);

    // Every chunk hand-over of the program ring costs ~1400 cycles per shot block (barrier + ring turn-around,
    // profiles/r1_notes.md), so the shared memory that the frame columns leave unused goes into larger chunks
    // (c3: 5.9 KB spare -> 2688-word chunks, 104 instead of 155 chunks). The circuit is lowered again for that size.
    if (env_u32("GSTIM_CHUNK_WORDS", 0) == 0) {
        const uint32_t Q0 = s->lc.num_qubits, pitch0 = Q0 | 1u;
        uint32_t n_noise0 = 0;
        for (const auto &b : s->lc.batches) {
            n_noise0 += b.op == GOP_NOISE1 || b.op == GOP_NOISE2 || b.op == GOP_CORR;
        }
        const size_t fixed0 = interp_smem_bytes(pitch0, Q0, 0, s->chunk_words, n_noise0);
        const size_t per_k0 = (size_t)2 * pitch0 * 16 + 16;
        if (fixed0 + per_k0 <= s->smem_optin) {
            const size_t k0 = std::min<size_t>((s->smem_optin - fixed0) / per_k0, 32);
            const size_t spare = s->smem_optin - fixed0 - k0 * per_k0;
            const uint32_t grown = std::min<uint32_t>(8192, s->chunk_words + (uint32_t)(spare / 8) / 128 * 128);
            if (grown >= s->chunk_words + 256) {
                s->chunk_words = grown;
                s->lc = lower_circuit(s->circuit, (uint32_t)s->mode, s->chunk_words - GSTIM_HDR_WORDS);
            }
        }
    }

    // slots: cover the batch size below which 95% of all items live
    std::vector<std::pair<uint32_t, uint64_t>> sizes;
    for (const auto &b : s->lc.batches) {
        uint32_t n = (uint32_t)(b.res_off.size() - 1);
        sizes.push_back({n, n});
    }
    std::sort(sizes.begin(), sizes.end());
    uint64_t acc = 0;
    uint32_t p95 = 1;
    for (auto &e : sizes) {
        acc += e.second;
        p95 = e.first;
        if (acc * 100 >= s->lc.total_items * 95) {
            break;
        }
    }
    // (at most 640 interpreter threads: with the 128 noise producer threads the block stays at 768 threads = 80 registers
    // per thread; at 896 threads = 72 registers the opcode functions spill - c5: 7 % slower)
    uint32_t slots = std::min<uint32_t>(640, std::max<uint32_t>(32, (p95 + 31) / 32 * 32));
    slots = env_u32("GSTIM_SLOTS", slots);

    // K_max from the shared-memory budget
    uint32_t Q = s->lc.num_qubits;
    uint32_t q_pitch = Q | 1u;
    uint32_t n_noise_batches = 0;
    for (const auto &b : s->lc.batches) {
        n_noise_batches += b.op == GOP_NOISE1 || b.op == GOP_NOISE2 || b.op == GOP_CORR;
    }
    s->n_noise = n_noise_batches;
    size_t fixed = interp_smem_bytes(q_pitch, Q, 0, s->chunk_words, s->n_noise);
    size_t per_k = (size_t)2 * q_pitch * 16 + 16;
    if (fixed + per_k > s->smem_optin) {
        // The interpreter keeps the whole frame of a 128-shot column in shared memory. Such circuits are sampled by the
        // event engine (no frame at all) when they are eligible for it; the interpreter reports this when asked to run.
        s->interp_why = "Circuit frame (" + std::to_string(Q) + " active qubits) does not fit in " + std::to_string(s->smem_optin) +
                        " bytes of shared memory: the interpreter cannot sample this circuit (the event engine can, when the "
                        "circuit is eligible for it).";
        s->plan = GstimPlan{};
        s->plan.num_qubits = Q;
        s->plan.q_pitch = q_pitch;
        s->plan.num_meas = (uint32_t)s->lc.stats.num_measurements;
        s->plan.num_det = (uint32_t)s->lc.stats.num_detectors;
        s->plan.num_obs = (uint32_t)s->lc.stats.num_observables;
        s->plan.rec_ring = s->lc.rec_ring;
        s->plan.mode = s->lc.mode;
        s->plan.max_items = s->lc.max_items;
        s->plan.n_batches = (uint32_t)s->lc.batches.size();
        s->slots = slots;
        s->threads = slots;
        s->K_max = 0;
        return;
    }
    uint32_t K_max = (uint32_t)std::min<size_t>((s->smem_optin - fixed) / per_k, 32);

    // lanes per item: fill up to ~512 threads
    uint32_t target_threads = env_u32("GSTIM_THREADS", 512);
    uint32_t G_log2 = 0;
    while ((2u << G_log2) <= 32 && slots * (2u << G_log2) <= target_threads && (2u << G_log2) <= K_max) {
        G_log2++;
    }
    G_log2 = env_u32("GSTIM_G_LOG2", G_log2);
    uint32_t G = 1u << G_log2;
    K_max = K_max / G * G;
    K_max = std::min(K_max, env_u32("GSTIM_KMAX", K_max));
    K_max = std::max(K_max / G * G, G);
    s->slots = slots;
    s->G_log2 = G_log2;
    s->threads = slots * G;
    s->K_max = K_max;
    if (s->threads > 1024) {
        throw std::invalid_argument("internal: thread count exceeds 1024");
    }
    // noise producer warps (interp.cu): 128 threads next to the interpreter's keep the block at 768 threads = 80 registers
    s->pre_threads = 0;
    if (s->n_noise > 0 && s->threads + 32 <= 1024) {
        uint32_t want = env_u32("GSTIM_PRE_THREADS", 128) / 32 * 32;
        s->pre_threads = std::max<uint32_t>(32, std::min<uint32_t>(want, 1024 - s->threads));
        if (env_u32("GSTIM_PRE_THREADS", 128) == 0) {
            s->pre_threads = 0;  // timing experiments only: no noise events are produced
        }
    }
    if (s->n_noise > 0 && s->pre_threads == 0 && env_u32("GSTIM_PRE_THREADS", 128) != 0) {
        throw std::invalid_argument("internal: no room for the noise producer warps");
    }

    s->words = serialize_program(s->lc, slots, s->chunk_words, &s->plan);
    s->d_prog.ensure(s->words.size() * 4);
    CK(cudaMemcpy(s->d_prog.p, s->words.data(), s->words.size() * 4, cudaMemcpyHostToDevice));
    {
        const NoiseSchedule &ns = s->lc.noise;
        auto up = [&](DevBuf &d, const void *src, size_t bytes) {
            d.ensure(std::max<size_t>(bytes, 16));
            if (bytes) {
                CK(cudaMemcpy(d.p, src, bytes, cudaMemcpyHostToDevice));
            }
        };
        up(s->d_noise_info, ns.info.data(), ns.info.size() * 4);
        up(s->d_rates, ns.rates.data(), ns.rates.size() * 8);
        up(s->d_slices, ns.slices.data(), ns.slices.size() * 4);
        s->d_ev_overflow.ensure(16);
        CK(cudaMemset(s->d_ev_overflow.p, 0, 16));
    }
    CK(interp_set_max_smem(interp_smem_bytes(q_pitch, Q, K_max, s->chunk_words, s->n_noise)));
}

int engine_from_env() {
    const char *v = getenv("GSTIM_ENGINE");
    if (v == nullptr || *v == '\0' || strcmp(v, "auto") == 0) {
        return GSTIM_ENGINE_AUTO;
    }
    if (strcmp(v, "interp") == 0 || strcmp(v, "interpreter") == 0) {
        return GSTIM_ENGINE_INTERPRETER;
    }
    if (strcmp(v, "events") == 0 || strcmp(v, "sparse") == 0) {
        return GSTIM_ENGINE_EVENTS;
    }
    throw std::invalid_argument("GSTIM_ENGINE must be auto, interp or events.");
}

// Builds the response table (response.cc) and, when the circuit is eligible, the event-driven engine (sparse.cu).
void configure_event_engine(gstim_sampler *s) {
    s->engine_pref = engine_from_env();
    s->sparse.reset();
    s->sparse_favoured = false;
    if (env_u32("GSTIM_NO_EVENT_ENGINE", 0)) {
        s->sparse_why = "disabled by GSTIM_NO_EVENT_ENGINE";
        return;
    }
    ResponseTable rt = build_response_table(s->lc);
    if (!rt.eligible) {
        s->sparse_why = rt.why_not;
        return;
    }
    // Cost model in SM-cycles per shot (profiles/r2_notes.md): the event engine pays ~2.2 per event plus ~0.3 per flipped
    // bit; the interpreter ~0.009 per lowered item plus ~6 per noise event (its event application). Long responses
    // (measurement sampling with many live collapse bits) are what makes the interpreter the better choice.
    double noise_events = 0;
    for (size_t i = 0; i < s->lc.noise.n_sites.size(); i++) {
        const double lam = std::ldexp((double)s->lc.noise.lams[i], -56);
        noise_events += (double)s->lc.noise.n_sites[i] * (s->lc.noise.lams[i] >= (1ull << 62) ? 1.0 : -std::expm1(-lam));
    }
    const double ev_cost = rt.events_per_shot * 2.2 + rt.flips_per_shot * 0.3;
    const double interp_cost = (double)s->lc.total_items * 0.009 + noise_events * 6.0;
    const bool favoured = ev_cost < interp_cost;
    try {
        s->sparse = std::make_unique<SparseEngine>(
            std::move(rt), (uint32_t)s->mode, s->plan.num_det, s->plan.num_obs, s->plan.num_meas, s->device, env_u32("GSTIM_SLICE_EVENTS", 0),
            env_u32("GSTIM_TILE_BUFFERS", 0), env_u32("GSTIM_TABLE_COMPRESS_MB", 48));
        s->sparse_favoured = favoured;
        s->sparse_why.clear();
    } catch (const std::invalid_argument &e) {
        s->sparse_why = e.what();
    }
}

bool use_events(const gstim_sampler *s) {
    if (!s->interp_why.empty()) {
        if (s->sparse && s->engine_pref != GSTIM_ENGINE_INTERPRETER) {
            return true;
        }
        throw std::invalid_argument(s->interp_why);
    }
    if (!s->sparse || s->engine_pref == GSTIM_ENGINE_INTERPRETER) {
        return false;
    }
    return s->engine_pref == GSTIM_ENGINE_EVENTS || s->sparse_favoured;
}

uint32_t choose_K(const gstim_sampler *s, uint64_t shots) {
    if (s->K_fixed) {
        return s->K_fixed;
    }
    uint32_t G = 1u << s->G_log2;
    uint64_t cols = (shots + GSTIM_COL_SHOTS - 1) / GSTIM_COL_SHOTS;
    // aim for at least two blocks per SM before growing the block
    uint64_t k = cols / ((uint64_t)2 * s->num_sms);
    k = k / G * G;
    if (k < G) {
        k = G;
    }
    if (k > s->K_max) {
        k = s->K_max;
    }
    return (uint32_t)k;
}

cudaEvent_t get_event(gstim_sampler *s, size_t i) {
    while (s->events.size() <= i) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        s->events.push_back(e);
    }
    return s->events[i];
}

// One pass of the event engine over `shots` shots in chunks of at most `chunk_shots` (a multiple of 128): for every
// chunk, `place(first, n, &main, &main_pitch, &obs, &obs_pitch)` names the dense device rows to fill and
// `after(first, n)` runs once the kernel is enqueued on s->stream.
template <typename PLACE, typename AFTER>
void run_events(gstim_sampler *s, uint64_t shots, uint32_t layout_flags, uint64_t chunk_shots, PLACE &&place, AFTER &&after) {
    CK(cudaSetDevice(s->device));
    s->last_launches = 0;
    s->last_interp_ms = 0;
    s->last_transpose_ms = 0;
    s->last_call_ms = 0;
    s->last_engine = GSTIM_ENGINE_EVENTS;
    s->last_K = 0;
    if (shots == 0) {
        return;
    }
    SparseEngine &E = *s->sparse;
    const uint64_t cols = (shots + GSTIM_COL_SHOTS - 1) / GSTIM_COL_SHOTS;
    if (s->next_col + cols >= (1ull << 47)) {
        throw std::invalid_argument("shot offset + shots must stay below 2^54");
    }
    E.set_layout(layout_flags, s->stream);
    if (s->mode == GSTIM_MODE_MEASUREMENTS) {
        std::vector<uint8_t> packed((s->plan.num_meas + 7) / 8, 0);
        for (size_t k = 0; k < s->ref_bits.size(); k++) {
            packed[k >> 3] |= (uint8_t)(s->ref_bits[k] << (k & 7));
        }
        E.set_reference_row(s->ref_bits.empty() ? nullptr : packed.data(), packed.size());
    }
    if (!s->call_start) {
        CK(cudaEventCreate(&s->call_start));
        CK(cudaEventCreate(&s->call_end));
    }
    chunk_shots = std::max<uint64_t>(chunk_shots / GSTIM_COL_SHOTS, 1) * GSTIM_COL_SHOTS;
    const uint64_t base = s->next_col * GSTIM_COL_SHOTS;
    CK(cudaEventRecord(s->call_start, s->stream));
    size_t ev = 0;
    for (uint64_t first = 0; first < shots; first += chunk_shots) {
        const uint64_t n = std::min(chunk_shots, shots - first);
        uint8_t *main = nullptr, *obs = nullptr;
        uint64_t main_pitch = 0, obs_pitch = 0;
        place(first, n, &main, &main_pitch, &obs, &obs_pitch);
        cudaEvent_t e0 = get_event(s, ev++), e1 = get_event(s, ev++);
        CK(cudaEventRecord(e0, s->stream));
        try {
            E.launch(base + first, n, main, main_pitch, obs, obs_pitch, s->seed, s->stream);
        } catch (const std::runtime_error &e) {
            throw CudaError(e.what());
        }
        CK(cudaEventRecord(e1, s->stream));
        s->last_launches++;
        after(first, n);
    }
    CK(cudaEventRecord(s->call_end, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaEventElapsedTime(&s->last_call_ms, s->call_start, s->call_end));
    for (size_t i = 0; i + 2 <= ev; i += 2) {
        float a = 0;
        CK(cudaEventElapsedTime(&a, s->events[i], s->events[i + 1]));
        s->last_interp_ms += a;
    }
    s->next_col += cols;
}

struct RowMaps {
    std::vector<uint32_t> main, obs;
};

RowMaps detector_row_maps(const gstim_sampler *s, uint32_t flags) {
    RowMaps r;
    uint32_t D = s->plan.num_det, L = s->plan.num_obs;
    if (flags & GSTIM_PREPEND_OBS) {
        for (uint32_t l = 0; l < L; l++) {
            r.main.push_back(D + l);
        }
    }
    for (uint32_t d = 0; d < D; d++) {
        r.main.push_back(d);
    }
    if (flags & GSTIM_APPEND_OBS) {
        for (uint32_t l = 0; l < L; l++) {
            r.main.push_back(D + l);
        }
    }
    if (flags & GSTIM_SEPARATE_OBS) {
        for (uint32_t l = 0; l < L; l++) {
            r.obs.push_back(D + l);
        }
    }
    return r;
}

RowMaps measurement_row_maps(const gstim_sampler *s) {
    RowMaps r;
    uint32_t M = s->plan.num_meas;
    r.main.resize(M);
    for (uint32_t m = 0; m < M; m++) {
        uint32_t inv = m < s->ref_bits.size() && s->ref_bits[m] ? 1u : 0u;
        r.main[m] = m | (inv << 31);
    }
    return r;
}

// One pass of the sampler over `shots` shots. For every chunk, `sink(chunk_first_shot, chunk_shots,
// table, row_words)` is called after the interpreter kernel has been enqueued on s->stream.
template <typename SINK>
void run_sampler(gstim_sampler *s, uint64_t shots, SINK &&sink, uint32_t table_mb = 2048) {
    CK(cudaSetDevice(s->device));
    s->last_launches = 0;
    s->last_interp_ms = 0;
    s->last_transpose_ms = 0;
    s->last_call_ms = 0;
    s->last_engine = GSTIM_ENGINE_INTERPRETER;
    if (shots == 0) {
        return;
    }
    const uint32_t K = choose_K(s, shots);
    s->last_K = K;
    const uint32_t B = K * GSTIM_COL_SHOTS;
    const uint32_t Q = s->plan.num_qubits, q_pitch = s->plan.q_pitch;
    const size_t smem = interp_smem_bytes(q_pitch, Q, K, s->chunk_words, s->n_noise);
    const uint32_t rows = n_rows_of(s);
    const uint64_t total_blocks = (shots + B - 1) / B;

    // chunking: bit-major staging table limited to ~2 GiB
    // (host delivery passes a smaller budget: it is bound by the PCIe copy, and short chunks fill that pipeline sooner)
    const uint64_t table_budget = (uint64_t)env_u32("GSTIM_TABLE_MB", table_mb) << 20;
    const uint64_t bytes_per_block = (uint64_t)std::max<uint32_t>(rows, 1) * K * 16;
    // persistent grid: as many blocks as fit on the device at once (shared memory usually allows one per SM)
    uint32_t grid_cap = (uint32_t)s->num_sms * (uint32_t)interp_max_blocks_per_sm(s->threads + s->pre_threads, smem);
    uint64_t max_blocks = std::max<uint64_t>(table_budget / bytes_per_block, (uint64_t)grid_cap);
    // whole waves only: a chunk of 11.07 waves costs 12 (the CTAs walk their shot blocks in lock step)
    max_blocks = max_blocks / grid_cap * grid_cap;
    max_blocks = std::min(max_blocks, total_blocks);
    s->d_table.ensure(bytes_per_block * max_blocks);
    if (s->mode == GSTIM_MODE_DETECTORS) {
        s->d_rec.ensure((size_t)grid_cap * s->plan.rec_ring * K * 16);
    }

    // event scratch: one segment per noise batch, sized mean + 12 sigma + 64 of its Poisson event count
    const uint32_t n_noise = (uint32_t)s->lc.noise.n_sites.size();
    if (s->segoff_K != K) {
        std::vector<uint32_t> segoff(n_noise + 1, 0);
        uint64_t total = 0;
        for (uint32_t i = 0; i < n_noise; i++) {
            const double lam = std::ldexp((double)s->lc.noise.lams[i], -56);
            const double pr = s->lc.noise.lams[i] >= (1ull << 62) ? 1.0 : -std::expm1(-lam);
            const double sites = (double)s->lc.noise.n_sites[i] * B;
            const double mean = sites * pr;
            const double cap = std::min(sites, std::ceil(mean + 12.0 * std::sqrt(mean) + 64.0));
            segoff[i] = (uint32_t)total;
            total += (uint64_t)cap;
            if (total >= (1ull << 31)) {
                throw std::invalid_argument("noise event scratch would exceed 2^31 records per block; lower the noise or the block size");
            }
        }
        segoff[n_noise] = (uint32_t)total;
        s->d_segoff.ensure(segoff.size() * 4);
        CK(cudaMemcpy(s->d_segoff.p, segoff.data(), segoff.size() * 4, cudaMemcpyHostToDevice));
        s->segoff_K = K;
        s->ev_total = total;
    }
    s->d_ev_counts.ensure(std::max<size_t>((size_t)grid_cap * 2 * n_noise * 4, 16));
    s->d_ev_buf.ensure(std::max<size_t>((size_t)grid_cap * 2 * s->ev_total * 4, 16));

    if (s->next_col + total_blocks * K >= (1ull << 47)) {
        throw std::invalid_argument("shot offset + shots must stay below 2^54");
    }
    if (!s->call_start) {
        CK(cudaEventCreate(&s->call_start));
        CK(cudaEventCreate(&s->call_end));
    }
    CK(cudaEventRecord(s->call_start, s->stream));
    size_t ev = 0;
    uint64_t done_blocks = 0;
    while (done_blocks < total_blocks) {
        const uint64_t nb = std::min(max_blocks, total_blocks - done_blocks);
        const uint64_t first_shot = done_blocks * B;
        const uint64_t chunk_shots = std::min<uint64_t>(nb * B, shots - first_shot);
        InterpParams p{};
        p.prog = (const uint32_t *)s->d_prog.p;
        p.n_chunks = s->plan.n_chunks;
        p.chunk_words = s->chunk_words;
        p.Q = Q;
        p.q_pitch = q_pitch;
        p.K = K;
        p.G_log2 = s->G_log2;
        p.slots = s->slots;
        p.threads_interp = s->threads;
        p.n_blocks = (uint32_t)nb;
        p.n_noise = n_noise;
        p.n_rates = (uint32_t)(s->lc.noise.rates.size() / 2);
        p.noise_info = (const uint32_t *)s->d_noise_info.p;
        p.rates = (const ulonglong2 *)s->d_rates.p;
        p.slices = (const uint4 *)s->d_slices.p;
        p.n_slices = (uint32_t)(s->lc.noise.slices.size() / GSTIM_SLICE_WORDS);
        p.ev_segoff = (const uint32_t *)s->d_segoff.p;
        p.ev_counts = (uint32_t *)s->d_ev_counts.p;
        p.ev_buf = (uint32_t *)s->d_ev_buf.p;
        p.ev_overflow = (uint32_t *)s->d_ev_overflow.p;
        p.dbg_cycles = nullptr;
        p.dbg_flags = env_u32("GSTIM_DEBUG_FLAGS", 0);
        if (env_u32("GSTIM_DEBUG_CYCLES", 0)) {
            s->d_dbg.ensure(64 * 8);
            CK(cudaMemsetAsync(s->d_dbg.p, 0, 64 * 8, s->stream));
            p.dbg_cycles = (unsigned long long *)s->d_dbg.p;
        }
        p.col0_base = s->next_col + done_blocks * K;
        p.seed_lo = (uint32_t)s->seed;
        p.seed_hi = (uint32_t)(s->seed >> 32);
        p.out_k_stride = std::max<uint32_t>(rows, 1);
        p.rec_mask = s->mode == GSTIM_MODE_DETECTORS ? s->plan.rec_ring - 1 : 0xFFFFFFFFu;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(nb, grid_cap);
        if (s->mode == GSTIM_MODE_DETECTORS) {
            p.out = (uint4 *)s->d_table.p;
            p.rec = (uint4 *)s->d_rec.p;
            p.rec_k_stride = s->plan.rec_ring;
            p.rec_block_stride = 0;
            p.rec_cta_stride = (uint64_t)s->plan.rec_ring * K;
        } else {
            p.out = nullptr;
            p.rec = (uint4 *)s->d_table.p;
            p.rec_k_stride = std::max<uint32_t>(rows, 1);
            p.rec_block_stride = (uint64_t)K * std::max<uint32_t>(rows, 1);
            p.rec_cta_stride = 0;
        }
        cudaEvent_t e0 = get_event(s, ev++), e1 = get_event(s, ev++), e2 = get_event(s, ev++);
        CK(cudaEventRecord(e0, s->stream));
        // (the attribute is per kernel and device, not per sampler: another sampler may have lowered it since)
        CK(interp_set_max_smem(smem));
        CK(launch_interp(p, grid, s->threads + s->pre_threads, smem, s->stream));
        CK(cudaEventRecord(e1, s->stream));
        s->last_launches++;
        sink(first_shot, chunk_shots, (const uint32_t *)s->d_table.p, (uint64_t)std::max<uint32_t>(rows, 1));
        CK(cudaEventRecord(e2, s->stream));
        done_blocks += nb;
    }
    CK(cudaEventRecord(s->call_end, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (env_u32("GSTIM_DEBUG_CYCLES", 0) && s->d_dbg.p) {
        unsigned long long h[64];
        CK(cudaMemcpy(h, s->d_dbg.p, sizeof(h), cudaMemcpyDeviceToHost));
        static const char *names[13] = {"END", "CHUNK", "CLIFF1", "CLIFF2", "NOISE1", "NOISE2", "MEASURE", "RECZERO", "XORROWS", "OBS_PAULI", "FEEDBACK", "CORR", "QMAP"};
        unsigned long long tot = 0;
        for (int i = 0; i < 13; i++) {
            tot += h[i];
        }
        fprintf(stderr, "[gstim cycles, block 0, last launch] total %llu\n", tot);
        fprintf(stderr, "  prepass(thread 0): %llu cyc; events %llu, slices %llu, iterations %llu\n  noise apply(thread 0): prefetch+wait %llu, entry barrier %llu, flips %llu, exit barrier %llu cyc\n", h[32], h[33], h[34], h[35], h[40], h[41], h[42], h[43]);
        for (int i = 0; i < 13; i++) {
            if (h[i]) {
                fprintf(stderr, "  %-9s %10llu cyc (%5.1f%%)  %6llu batches  %8.0f cyc/batch\n", names[i], h[i], 100.0 * h[i] / tot, h[16 + i], h[16 + i] ? (double)h[i] / h[16 + i] : 0.0);
            }
        }
    }
    {
        uint32_t overflow = 0;
        CK(cudaMemcpy(&overflow, s->d_ev_overflow.p, 4, cudaMemcpyDeviceToHost));
        if (overflow) {
            CK(cudaMemset(s->d_ev_overflow.p, 0, 16));
            throw std::runtime_error("internal: noise event scratch overflowed (a > 12 sigma fluctuation); results of this call are invalid");
        }
    }
    CK(cudaEventElapsedTime(&s->last_call_ms, s->call_start, s->call_end));
    for (size_t i = 0; i + 3 <= ev; i += 3) {
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, s->events[i], s->events[i + 1]));
        CK(cudaEventElapsedTime(&b, s->events[i + 1], s->events[i + 2]));
        s->last_interp_ms += a;
        s->last_transpose_ms += b;
    }
    s->next_col += total_blocks * K;
}

void upload_row_map(gstim_sampler *s, const std::vector<uint32_t> &m, size_t offset_words) {
    if (!m.empty()) {
        CK(cudaMemcpyAsync(
            (uint32_t *)s->d_rowmap.p + offset_words, m.data(), m.size() * 4, cudaMemcpyHostToDevice, s->stream));
    }
}

void transpose_to(
    gstim_sampler *s,
    const uint32_t *table,
    uint64_t n_rows,
    const uint32_t *d_map,
    uint32_t n_bits,
    uint64_t n_shots,
    uint8_t *out,
    uint64_t pitch) {
    TransposeParams t{};
    t.table = table;
    t.n_rows = n_rows;
    t.row_map = d_map;
    t.n_bits = n_bits;
    t.n_shots = n_shots;
    t.out = out;
    t.out_pitch = pitch;
    CK(launch_transpose_b8(t, s->stream));
    if (n_bits && n_shots) {
        s->last_launches++;
    }
}

// device-resident dense b8 output
void sample_to_device(
    gstim_sampler *s,
    uint64_t shots,
    const RowMaps &maps,
    uint32_t layout_flags,
    uint8_t *main_out,
    int64_t main_stride,
    uint8_t *obs_out,
    int64_t obs_stride) {
    const uint32_t nb_main = (uint32_t)maps.main.size(), nb_obs = (uint32_t)maps.obs.size();
    const uint64_t main_pitch = main_stride ? (uint64_t)main_stride : (nb_main + 7) / 8;
    const uint64_t obs_pitch = obs_stride ? (uint64_t)obs_stride : (nb_obs + 7) / 8;
    CK(cudaSetDevice(s->device));
    if (use_events(s)) {
        // the engine writes the caller's rows directly: the output bytes are the only HBM traffic
        run_events(
            s, shots, layout_flags, 1ull << 30,
            [&](uint64_t first, uint64_t, uint8_t **m, uint64_t *mp, uint8_t **o, uint64_t *op) {
                *m = main_out && nb_main ? main_out + first * main_pitch : nullptr;
                *mp = main_pitch;
                *o = obs_out && nb_obs ? obs_out + first * obs_pitch : nullptr;
                *op = obs_pitch;
            },
            [](uint64_t, uint64_t) {});
        return;
    }
    s->d_rowmap.ensure((size_t)(nb_main + nb_obs + 1) * 4);
    upload_row_map(s, maps.main, 0);
    upload_row_map(s, maps.obs, nb_main);
    const uint32_t *dm = (const uint32_t *)s->d_rowmap.p;
    run_sampler(s, shots, [&](uint64_t first, uint64_t n, const uint32_t *table, uint64_t n_rows) {
        if (main_out && nb_main) {
            transpose_to(s, table, n_rows, dm, nb_main, n, main_out + first * main_pitch, main_pitch);
        }
        if (obs_out && nb_obs) {
            transpose_to(s, table, n_rows, dm + nb_main, nb_obs, n, obs_out + first * obs_pitch, obs_pitch);
        }
    });
}

void unpack_bits(const uint8_t *packed, size_t n_bits, uint8_t *out) {
    static const auto lut = [] {
        std::vector<uint64_t> t(256);
        for (int v = 0; v < 256; v++) {
            uint64_t w = 0;
            for (int k = 0; k < 8; k++) {
                w |= (uint64_t)((v >> k) & 1) << (8 * k);
            }
            t[v] = w;
        }
        return t;
    }();
    size_t full = n_bits / 8;
    for (size_t i = 0; i < full; i++) {
        memcpy(out + 8 * i, &lut[packed[i]], 8);
    }
    for (size_t k = full * 8; k < n_bits; k++) {
        out[k] = (packed[k >> 3] >> (k & 7)) & 1;
    }
}

// Host delivery of event-engine results as SPARSE records (sparse.cu "Sparse host delivery"): per chunk the GPU rewrites the
// dense rows as (count, offsets, values) of their non-zero bytes, only those cross PCIe, and host threads rebuild the caller's
// rows (memset + scatter). Three chunks are in flight: kernels of chunk k, the copy of chunk k - 1, the rebuild of chunk k - 2.
// Returns false (nothing sampled) when the path does not apply.
bool sample_to_host_sparse(
    gstim_sampler *s, uint64_t shots, const RowMaps &maps, uint32_t layout_flags, bool bit_packed, uint8_t *main_out, uint64_t main_pitch,
    uint8_t *obs_out, uint64_t obs_pitch) {
    const uint32_t nb_main = (uint32_t)maps.main.size(), nb_obs = (uint32_t)maps.obs.size();
    const uint64_t R = (nb_main + 7) / 8, obs_bytes = (nb_obs + 7) / 8;
    // Opt-in (GSTIM_D2H=sparse). Measured on the 16-core host of the B200 box (profiles/r2_notes.md): rebuilding c3's rows on
    // the CPU reaches 11 M shots/s against 28.5 M for the DMA engine writing dense rows - the host's cores and DRAM cannot
    // zero and patch 32.7 GB per step faster than PCIe delivers it. Hosts with many more cores may see it differently.
    const char *mode = getenv("GSTIM_D2H");
    const bool force = mode != nullptr && strcmp(mode, "sparse") == 0;
    if (!force) {
        return false;
    }
    const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (main_out == nullptr || R == 0 || R >= 65536 || shots == 0) {
        return false;
    }
    CK(cudaSetDevice(s->device));
    const uint64_t chunk = std::max<uint64_t>((((uint64_t)env_u32("GSTIM_STAGE_MB", 512) << 20) / R) / GSTIM_COL_SHOTS, 1) * GSTIM_COL_SHOTS;
    const uint64_t n_chunks = (shots + chunk - 1) / chunk;
    auto rebuild = [&](const gstim_sampler::SparseSlot &sl) {
        const unsigned long long *idx = (const unsigned long long *)sl.h_idx.p;
        const uint8_t *strm = (const uint8_t *)sl.h_strm.p;
        const unsigned nt = sl.n * R < (4u << 20) ? 1u : hw;
        auto work = [&](uint64_t a, uint64_t b) {
            for (uint64_t i = a; i < b; i++) {
                const uint8_t *rec = strm + idx[i];
                const uint32_t cnt = *reinterpret_cast<const uint16_t *>(rec);
                const uint16_t *offs = reinterpret_cast<const uint16_t *>(rec + 2);
                const uint8_t *vals = rec + 2 + 2 * (size_t)cnt;
                uint8_t *dst = main_out + (sl.first + i) * main_pitch;
                if (bit_packed) {
                    memset(dst, 0, R);
                    for (uint32_t k = 0; k < cnt; k++) {
                        dst[offs[k]] = vals[k];
                    }
                } else {
                    memset(dst, 0, nb_main);
                    for (uint32_t k = 0; k < cnt; k++) {
                        uint32_t v = vals[k];
                        uint8_t *d8 = dst + 8 * (size_t)offs[k];
                        while (v) {
                            d8[__builtin_ctz(v)] = 1;
                            v &= v - 1;
                        }
                    }
                }
                if (obs_out != nullptr && nb_obs) {
                    const uint8_t *src = (const uint8_t *)sl.h_obs.p + i * obs_bytes;
                    uint8_t *od = obs_out + (sl.first + i) * obs_pitch;
                    if (bit_packed) {
                        memcpy(od, src, obs_bytes);
                    } else {
                        unpack_bits(src, nb_obs, od);
                    }
                }
            }
        };
        if (nt == 1) {
            work(0, sl.n);
            return;
        }
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < nt; t++) {
            ts.emplace_back(work, sl.n * t / nt, sl.n * (t + 1) / nt);
        }
        for (auto &t : ts) {
            t.join();
        }
    };
    // stage B: the kernels of the slot are done -> enqueue its copies; stage C: copies done -> rebuild the rows
    auto stage_b = [&](gstim_sampler::SparseSlot &sl) {
        CK(cudaEventSynchronize(sl.ev_kernels));
        const unsigned long long used = ((const unsigned long long *)sl.h_ctl.p)[0];
        const bool overflow = ((const unsigned long long *)sl.h_ctl.p)[1] != 0;
        if (overflow) {
            // (denser than expected: this chunk goes out as dense rows)
            std::vector<uint8_t> tmp(sl.n * R);
            CK(cudaMemcpy(tmp.data(), sl.d_rows.p, sl.n * R, cudaMemcpyDeviceToHost));
            for (uint64_t i = 0; i < sl.n; i++) {
                uint8_t *dst = main_out + (sl.first + i) * main_pitch;
                if (bit_packed) {
                    memcpy(dst, tmp.data() + i * R, R);
                } else {
                    unpack_bits(tmp.data() + i * R, nb_main, dst);
                }
            }
            ((unsigned long long *)sl.h_ctl.p)[1] = 2;  // marks "main rows already delivered"
        } else {
            sl.h_strm.ensure(used + 16);
            CK(cudaMemcpyAsync(sl.h_strm.p, sl.d_strm.p, used, cudaMemcpyDeviceToHost, s->copy_stream));
            sl.h_idx.ensure(sl.n * 8);
            CK(cudaMemcpyAsync(sl.h_idx.p, sl.d_idx.p, sl.n * 8, cudaMemcpyDeviceToHost, s->copy_stream));
        }
        if (obs_out != nullptr && nb_obs) {
            sl.h_obs.ensure(sl.n * obs_bytes);
            CK(cudaMemcpyAsync(sl.h_obs.p, (const uint8_t *)sl.d_rows.p + sl.n * R, sl.n * obs_bytes, cudaMemcpyDeviceToHost, s->copy_stream));
        }
        CK(cudaEventRecord(sl.ev_copy, s->copy_stream));
        sl.state = 2;
    };
    auto stage_c = [&](gstim_sampler::SparseSlot &sl) {
        CK(cudaEventSynchronize(sl.ev_copy));
        if (((const unsigned long long *)sl.h_ctl.p)[1] == 2) {
            if (obs_out != nullptr && nb_obs) {
                for (uint64_t i = 0; i < sl.n; i++) {
                    const uint8_t *src = (const uint8_t *)sl.h_obs.p + i * obs_bytes;
                    uint8_t *od = obs_out + (sl.first + i) * obs_pitch;
                    if (bit_packed) {
                        memcpy(od, src, obs_bytes);
                    } else {
                        unpack_bits(src, nb_obs, od);
                    }
                }
            }
        } else {
            rebuild(sl);
        }
        sl.state = 0;
    };
    uint64_t k = 0;
    try {
        run_events(
            s, shots, layout_flags, chunk,
            [&](uint64_t first, uint64_t n, uint8_t **m, uint64_t *mp, uint8_t **o, uint64_t *op) {
                gstim_sampler::SparseSlot &sl = s->sparse_slot[k % 3];
                if (sl.state == 2) {
                    stage_c(sl);  // (the slot's previous chunk, k - 3)
                }
                if (!sl.ev_kernels) {
                    CK(cudaEventCreateWithFlags(&sl.ev_kernels, cudaEventDisableTiming));
                    CK(cudaEventCreateWithFlags(&sl.ev_copy, cudaEventDisableTiming));
                }
                sl.d_rows.ensure(n * (R + obs_bytes) + 16);
                sl.d_strm.ensure(n * R + 2 * n + 16);
                sl.d_idx.ensure(n * 8);
                sl.d_ctl.ensure(16);
                sl.h_ctl.ensure(16);
                sl.first = first;
                sl.n = n;
                *m = (uint8_t *)sl.d_rows.p;
                *mp = R;
                *o = nb_obs ? (uint8_t *)sl.d_rows.p + n * R : nullptr;
                *op = obs_bytes;
            },
            [&](uint64_t, uint64_t n) {
                gstim_sampler::SparseSlot &sl = s->sparse_slot[k % 3];
                CK(cudaMemsetAsync(sl.d_ctl.p, 0, 16, s->stream));
                CK(launch_compress_rows((const uint8_t *)sl.d_rows.p, R, (uint32_t)R, n, (uint8_t *)sl.d_strm.p, n * R + 2 * n,
                                        (unsigned long long *)sl.d_ctl.p, (unsigned long long *)sl.d_idx.p,
                                        (uint32_t *)((unsigned long long *)sl.d_ctl.p + 1), s->stream));
                CK(cudaMemcpyAsync(sl.h_ctl.p, sl.d_ctl.p, 16, cudaMemcpyDeviceToHost, s->stream));
                CK(cudaEventRecord(sl.ev_kernels, s->stream));
                sl.state = 1;
                s->last_launches++;
                // pipeline: copies of the previous chunk, rebuild of the one before
                if (k >= 1 && s->sparse_slot[(k - 1) % 3].state == 1) {
                    stage_b(s->sparse_slot[(k - 1) % 3]);
                }
                if (k >= 2 && s->sparse_slot[(k - 2) % 3].state == 2) {
                    stage_c(s->sparse_slot[(k - 2) % 3]);
                }
                k++;
            });
        // drain
        for (uint64_t j = (k >= 2 ? k - 2 : 0); j < k; j++) {
            gstim_sampler::SparseSlot &sl = s->sparse_slot[j % 3];
            if (sl.state == 1) {
                stage_b(sl);
            }
            if (sl.state == 2) {
                stage_c(sl);
            }
        }
    } catch (...) {
        cudaStreamSynchronize(s->copy_stream);
        cudaStreamSynchronize(s->stream);
        for (auto &sl : s->sparse_slot) {
            sl.state = 0;
        }
        throw;
    }
    (void)n_chunks;
    s->last_d2h_sparse = 1;
    return true;
}

// host output: transposed chunks are staged in device memory, copied to pinned host memory on a
// second stream (double buffered) and scattered into the caller's (possibly strided / unpacked) rows.
void sample_to_host(
    gstim_sampler *s,
    uint64_t shots,
    const RowMaps &maps,
    uint32_t layout_flags,
    bool bit_packed,
    uint8_t *main_out,
    int64_t main_stride,
    uint8_t *obs_out,
    int64_t obs_stride) {
    const uint32_t nb_main = (uint32_t)maps.main.size(), nb_obs = (uint32_t)maps.obs.size();
    const uint64_t main_bytes = (nb_main + 7) / 8, obs_bytes = (nb_obs + 7) / 8;
    const uint64_t main_row = bit_packed ? main_bytes : nb_main, obs_row = bit_packed ? obs_bytes : nb_obs;
    const uint64_t main_pitch = main_stride ? (uint64_t)main_stride : main_row;
    const uint64_t obs_pitch = obs_stride ? (uint64_t)obs_stride : obs_row;
    const uint64_t stage_pitch = main_bytes + obs_bytes;  // per shot in the staging buffers
    CK(cudaSetDevice(s->device));
    s->last_d2h_sparse = 0;
    if (use_events(s) && sample_to_host_sparse(s, shots, maps, layout_flags, bit_packed, main_out, main_pitch, obs_out, obs_pitch)) {
        return;
    }
    s->d_rowmap.ensure((size_t)(nb_main + nb_obs + 1) * 4);
    upload_row_map(s, maps.main, 0);
    upload_row_map(s, maps.obs, nb_main);
    const uint32_t *dm = (const uint32_t *)s->d_rowmap.p;
    for (int i = 0; i < 2; i++) {
        if (!s->stage_done[i]) {
            CK(cudaEventCreateWithFlags(&s->stage_done[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->stage_ready[i], cudaEventDisableTiming));
        }
    }

    // Pinned (page-locked) caller memory + packed layout: DMA straight into the caller's rows.
    auto is_pinned = [](const void *ptr) {
        if (ptr == nullptr) {
            return true;
        }
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return attr.type == cudaMemoryTypeHost;
    };
    const bool direct = bit_packed && is_pinned(main_out) && is_pinned(obs_out);
    if (!direct && main_out != nullptr && shots * main_pitch >= (64ull << 20) && env_u32("GSTIM_HUGEPAGE_HINT", 1)) {
        // A fresh pageable result array is first touched by the copy threads below: ask for transparent huge pages so that the
        // kernel zeroes and maps 2 MiB at a time instead of 4 KiB (a hint; ignored where THP is off).
        const uintptr_t a = (reinterpret_cast<uintptr_t>(main_out) + 4095) & ~(uintptr_t)4095;
        const uintptr_t b = (reinterpret_cast<uintptr_t>(main_out) + shots * main_pitch) & ~(uintptr_t)4095;
        if (b > a) {
            madvise(reinterpret_cast<void *>(a), b - a, MADV_HUGEPAGE);
        }
    }

    struct Pending {
        bool active = false;
        uint64_t first = 0, n = 0;
    } pending[2];
    auto scatter_rows = [&](const uint8_t *src, uint64_t src_bytes, uint32_t n_bits, uint8_t *dst0, uint64_t pitch, uint64_t n) {
        const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        const unsigned nt = n * src_bytes < (1u << 20) ? 1u : hw;
        auto work = [&](uint64_t a, uint64_t b) {
            for (uint64_t i = a; i < b; i++) {
                uint8_t *dst = dst0 + i * pitch;
                if (bit_packed) {
                    memcpy(dst, src + i * src_bytes, src_bytes);
                } else {
                    unpack_bits(src + i * src_bytes, n_bits, dst);
                }
            }
        };
        if (nt == 1) {
            work(0, n);
            return;
        }
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < nt; t++) {
            ts.emplace_back(work, n * t / nt, n * (t + 1) / nt);
        }
        for (auto &t : ts) {
            t.join();
        }
    };
    auto drain = [&](int b) {
        if (!pending[b].active) {
            return;
        }
        CK(cudaEventSynchronize(s->stage_done[b]));
        pending[b].active = false;
        if (direct) {
            return;
        }
        const uint8_t *h = (const uint8_t *)s->h_stage[b].p;
        const uint64_t first = pending[b].first, n = pending[b].n;
        if (main_out && nb_main) {
            scatter_rows(h, main_bytes, nb_main, main_out + first * main_pitch, main_pitch, n);
        }
        if (obs_out && nb_obs) {
            scatter_rows(h + n * main_bytes, obs_bytes, nb_obs, obs_out + first * obs_pitch, obs_pitch, n);
        }
    };

    int cur = 0;
    uint8_t *dmain = nullptr, *dobs = nullptr;
    // claims the staging buffer of the next chunk
    auto begin = [&](uint64_t n) {
        drain(cur);  // buffer about to be reused
        s->d_stage[cur].ensure(n * stage_pitch + 16);
        dmain = (uint8_t *)s->d_stage[cur].p;
        dobs = dmain + n * main_bytes;
    };
    // the chunk's rows are enqueued on s->stream: copy them out on the second stream so the next chunk's kernels
    // overlap the PCIe drain
    auto finish = [&](uint64_t first, uint64_t n) {
        CK(cudaEventRecord(s->stage_ready[cur], s->stream));
        CK(cudaStreamWaitEvent(s->copy_stream, s->stage_ready[cur], 0));
        if (direct) {
            if (main_out && nb_main) {
                CK(cudaMemcpy2DAsync(
                    main_out + first * main_pitch, main_pitch, dmain, main_bytes, main_bytes, n, cudaMemcpyDeviceToHost, s->copy_stream));
            }
            if (obs_out && nb_obs) {
                CK(cudaMemcpy2DAsync(
                    obs_out + first * obs_pitch, obs_pitch, dobs, obs_bytes, obs_bytes, n, cudaMemcpyDeviceToHost, s->copy_stream));
            }
        } else {
            s->h_stage[cur].ensure(n * stage_pitch + 16);
            CK(cudaMemcpyAsync(s->h_stage[cur].p, s->d_stage[cur].p, n * stage_pitch, cudaMemcpyDeviceToHost, s->copy_stream));
        }
        CK(cudaEventRecord(s->stage_done[cur], s->copy_stream));
        pending[cur].active = true;
        pending[cur].first = first;
        pending[cur].n = n;
        cur ^= 1;
    };
    try {
    if (use_events(s)) {
        const uint64_t chunk = std::max<uint64_t>(((uint64_t)env_u32("GSTIM_STAGE_MB", 256) << 20) / std::max<uint64_t>(stage_pitch, 1), 128);
        run_events(
            s, shots, layout_flags, chunk,
            [&](uint64_t, uint64_t n, uint8_t **m, uint64_t *mp, uint8_t **o, uint64_t *op) {
                begin(n);
                *m = nb_main ? dmain : nullptr;
                *mp = main_bytes;
                *o = nb_obs ? dobs : nullptr;
                *op = obs_bytes;
            },
            finish);
    } else {
    run_sampler(s, shots, [&](uint64_t first, uint64_t n, const uint32_t *table, uint64_t n_rows) {
        begin(n);
        if (nb_main) {
            transpose_to(s, table, n_rows, dm, nb_main, n, dmain, main_bytes);
        }
        if (nb_obs) {
            transpose_to(s, table, n_rows, dm + nb_main, nb_obs, n, dobs, obs_bytes);
        }
        finish(first, n);
    }, 512);
    }
    } catch (...) {
        // no copy may still be writing into the caller's buffer (or the staging buffers) once the error is reported
        cudaStreamSynchronize(s->copy_stream);
        pending[0].active = pending[1].active = false;
        throw;
    }
    drain(cur);
    drain(cur ^ 1);
}

// Host-output sampling of a multi-device sampler: the shots are cut into one contiguous range per device (multiples of
// 128 shots, the remainder to the last), every device samples its range of the SAME global shot index space on its own
// thread and writes its rows of the caller's arrays. With the event engine the stream is a function of (seed, global shot
// index) only, so the result is identical to a single-device call (shots shard with no inter-GPU traffic).
template <typename MAPS>
void sample_to_host_multi(
    gstim_sampler *s, uint64_t shots, MAPS &&maps_of, uint32_t layout_flags, bool bit_packed, uint8_t *main_out, int64_t main_stride,
    uint8_t *obs_out, int64_t obs_stride) {
    std::vector<gstim_sampler *> all{s};
    all.insert(all.end(), s->peers.begin(), s->peers.end());
    const uint64_t cols = (shots + GSTIM_COL_SHOTS - 1) / GSTIM_COL_SHOTS;
    const uint64_t per = cols / all.size();
    const uint64_t base_col = s->next_col;
    std::vector<std::exception_ptr> errors(all.size());
    std::vector<std::thread> threads;
    uint64_t launches = 0;
    for (size_t i = 0; i < all.size(); i++) {
        const uint64_t c0 = per * i, c1 = i + 1 == all.size() ? cols : per * (i + 1);
        const uint64_t first = c0 * GSTIM_COL_SHOTS, last = std::min<uint64_t>(c1 * GSTIM_COL_SHOTS, shots);
        if (last <= first) {
            continue;
        }
        threads.emplace_back([&, i, first, last, c0] {
            try {
                gstim_sampler *d = all[i];
                d->next_col = base_col + c0;
                d->engine_pref = s->engine_pref;
                d->ref_bits = s->ref_bits;
                RowMaps maps = maps_of(d);
                const uint64_t main_row = bit_packed ? (maps.main.size() + 7) / 8 : maps.main.size();
                const uint64_t obs_row = bit_packed ? (maps.obs.size() + 7) / 8 : maps.obs.size();
                const uint64_t mp = main_stride ? (uint64_t)main_stride : main_row, op = obs_stride ? (uint64_t)obs_stride : obs_row;
                sample_to_host(d, last - first, maps, layout_flags, bit_packed, main_out ? main_out + first * mp : nullptr, (int64_t)mp,
                               obs_out ? obs_out + first * op : nullptr, (int64_t)op);
            } catch (...) {
                errors[i] = std::current_exception();
            }
        });
    }
    for (auto &t : threads) {
        t.join();
    }
    for (auto &e : errors) {
        if (e) {
            std::rethrow_exception(e);
        }
    }
    for (gstim_sampler *d : all) {
        launches += d->last_launches;
    }
    s->last_launches = launches;
    s->next_col = base_col + cols;
    CK(cudaSetDevice(s->device));
}

void check_flag_combo(uint32_t flags) {
    int n = ((flags & GSTIM_PREPEND_OBS) != 0) + ((flags & GSTIM_APPEND_OBS) != 0) + ((flags & GSTIM_SEPARATE_OBS) != 0);
    if (n > 1) {
        throw std::out_of_range("Can't combine --prepend_observables, --append_observables, or --obs_out");
    }
}

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return GSTIM_OK;
    } catch (const std::invalid_argument &e) {
        g_last_error = e.what();
        return GSTIM_ERR_INVALID_ARGUMENT;
    } catch (const std::out_of_range &e) {
        g_last_error = e.what();
        return GSTIM_ERR_OUT_OF_RANGE;
    } catch (const OomError &e) {
        g_last_error = e.what();
        return GSTIM_ERR_OOM;
    } catch (const CudaError &e) {
        g_last_error = e.what();
        return GSTIM_ERR_CUDA;
    } catch (const IoError &e) {
        g_last_error = e.what();
        return GSTIM_ERR_IO;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return GSTIM_ERR_INTERNAL;
    }
}

void require(bool cond, const char *msg) {
    if (!cond) {
        throw std::invalid_argument(msg);
    }
}

struct FdFile {
    FILE *f = nullptr;
    explicit FdFile(int fd) {
        int d = dup(fd);
        if (d < 0) {
            throw IoError("dup() failed on the output file descriptor.");
        }
        f = fdopen(d, "wb");
        if (!f) {
            close(d);
            throw IoError("fdopen() failed on the output file descriptor.");
        }
    }
    ~FdFile() {
        if (f) {
            fclose(f);
        }
    }
};

// shot-major packed rows -> bit-major 32-bit rows, then the ptb64 writer (shots % 64 == 0 is checked by the callers)
void write_ptb64_rows(FILE *out, const uint8_t *rows, size_t row_pitch, uint64_t shots, uint64_t n_bits) {
    write_ptb64_from_rows(out, rows, row_pitch, shots, n_bits);
}

// Streams shots to a file in any format; chunked through the host sampler.
void sample_to_file(
    gstim_sampler *s,
    uint64_t shots,
    uint32_t layout_flags,
    const std::vector<uint32_t> &map,
    FILE *f,
    Format fmt,
    char p1,
    char p2,
    size_t transition,
    const std::vector<uint32_t> *obs_map,
    FILE *obs_f,
    Format obs_fmt) {
    const uint32_t nb = (uint32_t)map.size();
    const uint32_t nbo = obs_map ? (uint32_t)obs_map->size() : 0;
    if ((fmt == Format::PTB64 || (obs_f && obs_fmt == Format::PTB64)) && shots % 64 != 0) {
        throw std::invalid_argument("shots must be a multiple of 64 to use ptb64 format.");
    }
    // Host-side encoders work on b8 rows, or on the bit-major rows for ptb64.
    CK(cudaSetDevice(s->device));
    if (use_events(s)) {
        const uint64_t nbytes = (nb + 7) / 8, nbytes_o = (nbo + 7) / 8;
        // (ptb64 chunks must hold whole groups of 64 shots; 128-shot multiples do)
        const uint64_t chunk = std::max<uint64_t>((64ull << 20) / std::max<uint64_t>(nbytes + nbytes_o, 1), 128);
        std::vector<uint8_t> host;
        uint8_t *dmain = nullptr, *dobs = nullptr;
        run_events(
            s, shots, layout_flags, chunk,
            [&](uint64_t, uint64_t n, uint8_t **m, uint64_t *mp, uint8_t **o, uint64_t *op) {
                s->d_stage[0].ensure(n * (nbytes + nbytes_o) + 16);
                dmain = (uint8_t *)s->d_stage[0].p;
                dobs = dmain + n * nbytes;
                *m = nb ? dmain : nullptr;
                *mp = nbytes;
                *o = nbo ? dobs : nullptr;
                *op = nbytes_o;
            },
            [&](uint64_t, uint64_t n) {
                // rows leave through the page-locked staging pair: the DMA of one sub-chunk overlaps the (threaded) encoding of
                // the previous one; sub-chunks hold whole groups of 64 shots for ptb64
                if (obs_f && obs_map && nbytes_o) {
                    hp_staged_d2h_blocks(s->file_stage, s->stream, dobs, nbytes_o, n, 64, [&](const uint8_t *rows, uint64_t, uint64_t cnt) {
                        write_shots(obs_f, rows, nbytes_o, cnt, nbo, obs_fmt, 'L', 'L', nbo);
                    });
                } else if (obs_f && obs_map) {
                    write_shots(obs_f, (const uint8_t *)"", 0, n, 0, obs_fmt, 'L', 'L', 0);
                }
                if (nbytes) {
                    hp_staged_d2h_blocks(s->file_stage, s->stream, dmain, nbytes, n, 64, [&](const uint8_t *rows, uint64_t, uint64_t cnt) {
                        write_shots(f, rows, nbytes, cnt, nb, fmt, p1, p2, transition);
                    });
                } else {
                    write_shots(f, (const uint8_t *)"", 0, n, 0, fmt, p1, p2, transition);
                }
            });
        if (fflush(f) != 0 || (obs_f && fflush(obs_f) != 0)) {
            throw IoError("Failed to flush result data.");
        }
        return;
    }
    s->d_rowmap.ensure((size_t)(nb + nbo + 1) * 4);
    upload_row_map(s, map, 0);
    if (obs_map) {
        upload_row_map(s, *obs_map, nb);
    }
    const uint32_t *dm = (const uint32_t *)s->d_rowmap.p;
    const uint64_t nbytes = (nb + 7) / 8, nbytes_o = (nbo + 7) / 8;
    std::vector<uint8_t> host;
    std::vector<uint32_t> host_table;
    auto emit = [&](const uint32_t *table,
                    uint64_t n_rows,
                    uint64_t n,
                    const std::vector<uint32_t> &m,
                    const uint32_t *d_m,
                    uint64_t bytes,
                    FILE *out,
                    Format of,
                    char c1,
                    char c2,
                    size_t tr) {
        if (of == Format::PTB64) {
            host_table.resize((size_t)((n + 127) / 128) * n_rows * 4);
            CK(cudaMemcpyAsync(host_table.data(), table, host_table.size() * 4, cudaMemcpyDeviceToHost, s->stream));
            CK(cudaStreamSynchronize(s->stream));
            write_ptb64(out, host_table.data(), n_rows, m.data(), m.size(), n);
        } else {
            s->d_stage[0].ensure(n * bytes + 16);
            transpose_to(s, table, n_rows, d_m, (uint32_t)m.size(), n, (uint8_t *)s->d_stage[0].p, bytes);
            if (bytes) {
                hp_staged_d2h_blocks(s->file_stage, s->stream, (const uint8_t *)s->d_stage[0].p, bytes, n, 64,
                                     [&](const uint8_t *rows, uint64_t, uint64_t cnt) { write_shots(out, rows, bytes, cnt, m.size(), of, c1, c2, tr); });
            } else {
                CK(cudaStreamSynchronize(s->stream));
                write_shots(out, (const uint8_t *)"", 0, n, 0, of, c1, c2, tr);
            }
        }
    };
    run_sampler(s, shots, [&](uint64_t first, uint64_t n, const uint32_t *table, uint64_t n_rows) {
        (void)first;
        if (obs_f && obs_map) {
            emit(table, n_rows, n, *obs_map, dm + nb, nbytes_o, obs_f, obs_fmt, 'L', 'L', nbo);
        }
        emit(table, n_rows, n, map, dm, nbytes, f, fmt, p1, p2, transition);
    });
    if (fflush(f) != 0 || (obs_f && fflush(obs_f) != 0)) {
        throw IoError("Failed to flush result data.");
    }
}

void fill_table_info(const ResponseTable &rt, gstim_engine_info *out) {
    out->eligible = rt.eligible ? 1 : 0;
    out->num_classes = (uint32_t)rt.classes.size();
    out->max_response = rt.max_response;
    out->num_sites = rt.n_sites;
    out->num_entries = rt.n_entries;
    out->overflow_words = rt.overflow.size();
    out->events_per_shot = rt.events_per_shot;
    out->flips_per_shot = rt.flips_per_shot;
}

void export_table_array(const ResponseTable &rt, const std::vector<uint32_t> *slices, uint32_t n_det, int what, uint32_t *words,
                        size_t *n_words) {
    std::vector<uint32_t> tmp;
    const std::vector<uint32_t> *src = &tmp;
    switch (what) {
        case 0:
            for (const RespClass &c : rt.classes) {
                tmp.push_back((uint32_t)c.lam);
                tmp.push_back((uint32_t)(c.lam >> 32));
                tmp.push_back(c.inv);
                tmp.push_back(c.sh);
                tmp.push_back(c.kind);
                tmp.push_back(c.n_out);
                tmp.insert(tmp.end(), c.thr, c.thr + 15);
                tmp.push_back(c.n_sites);
                tmp.push_back(c.entry0);
                tmp.push_back(c.dense_thr);
            }
            break;
        case 1:
            src = &rt.entries;
            break;
        case 2:
            src = &rt.overflow;
            break;
        case 3:
            src = &rt.site_group;
            break;
        case 4:
            src = &rt.site_index;
            break;
        case 5:
            src = &rt.outcome_word;
            break;
        case 7:  // round structure per class: a, p, n, delta (response.h ResponsePeriod)
            for (const RespClass &c : rt.classes) {
                const ResponsePeriod pd = find_response_period(rt, c, n_det);
                tmp.push_back(pd.a);
                tmp.push_back(pd.p);
                tmp.push_back(pd.n);
                tmp.push_back(pd.delta);
            }
            break;
        case 6:
            if (slices == nullptr) {
                throw std::invalid_argument("slices belong to a sampler (they depend on the tile height).");
            }
            src = slices;
            break;
        default:
            throw std::invalid_argument("bad table selector.");
    }
    if (words == nullptr) {
        *n_words = src->size();
        return;
    }
    require(*n_words >= src->size(), "buffer too small.");
    if (!src->empty()) {
        memcpy(words, src->data(), src->size() * 4);
    }
    *n_words = src->size();
}

}  // namespace

// (used by dem.cu) copies one array of a response table, see gstim_get_response_table
void gstim_export_table_array(const ResponseTable &rt, const std::vector<uint32_t> *slices, uint32_t n_det, int what, uint32_t *words,
                              size_t *n_words) {
    export_table_array(rt, slices, n_det, what, words, n_words);
}

// (used by dem.cu: one error channel for the whole library)
void gstim_set_last_error(const char *msg) {
    g_last_error = msg ? msg : "";
}

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int gstim_version(void) {
    return 1;
}

const char *gstim_last_error(void) {
    return g_last_error.c_str();
}

int gstim_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int gstim_circuit_stats(const char *circuit_text, size_t text_len, gstim_stats *out) {
    return guarded([&] {
        require(circuit_text != nullptr && out != nullptr, "NULL argument.");
        Circuit c = Circuit::from_text(std::string_view(circuit_text, text_len));
        // (parse + count only: like the reference, a bad rec[-k] lookback surfaces when a sampler is compiled, not here)
        const CircuitStats st = compute_stats(c);
        std::vector<uint8_t> used(st.num_qubits, 0);
        uint64_t active = 0;
        c.for_each_instruction_once([&](const Instruction &op) {
            if (op.gate->cat == GateCat::MPAD || op.gate->targets == TR_NONE || std::string(op.gate->name) == "QUBIT_COORDS") {
                return;
            }
            for (uint32_t t : op.targets) {
                if (t != T_COMBINER && !(t & (T_REC | T_SWEEP)) && !used[t & T_VALUE_MASK]) {
                    used[t & T_VALUE_MASK] = 1;
                    active++;
                }
            }
        });
        memset(out, 0, sizeof(*out));
        out->num_qubits = st.num_qubits;
        out->num_measurements = st.num_measurements;
        out->num_detectors = st.num_detectors;
        out->num_observables = st.num_observables;
        out->max_lookback = st.max_lookback;
        out->active_qubits = active;
    });
}

int gstim_reference_sample(const char *circuit_text, size_t text_len, uint8_t *bits_out, size_t n_bits) {
    return guarded([&] {
        require(circuit_text != nullptr, "NULL argument.");
        Circuit c = Circuit::from_text(std::string_view(circuit_text, text_len));
        std::vector<uint8_t> r = reference_sample(c);
        require(r.size() == n_bits, "n_bits must equal the circuit's number of measurements.");
        require(bits_out != nullptr || n_bits == 0, "NULL output.");
        memset(bits_out, 0, (n_bits + 7) / 8);
        for (size_t k = 0; k < n_bits; k++) {
            bits_out[k >> 3] |= (uint8_t)(r[k] << (k & 7));
        }
    });
}

int gstim_lower_text(
    const char *circuit_text,
    size_t text_len,
    int mode,
    uint32_t slots,
    uint32_t chunk_words,
    uint32_t *words,
    size_t *n_words,
    uint32_t plan_out[16]) {
    return guarded([&] {
        require(circuit_text != nullptr && n_words != nullptr, "NULL argument.");
        require(mode == GSTIM_MODE_DETECTORS || mode == GSTIM_MODE_MEASUREMENTS, "bad mode.");
        require(slots >= 1, "slots must be positive.");
        Circuit c = Circuit::from_text(std::string_view(circuit_text, text_len));
        if (chunk_words == 0) {
            uint32_t need = 4 * (max_targets_in_one_item(c) + 64);
            chunk_words = 2048;
            while (chunk_words < need) {
                chunk_words <<= 1;
            }
        }
        LoweredCircuit lc = lower_circuit(c, (uint32_t)mode, chunk_words - GSTIM_HDR_WORDS);
        GstimPlan plan;
        std::vector<uint32_t> w = serialize_program(lc, slots, chunk_words, &plan);
        if (plan_out != nullptr) {
            memset(plan_out, 0, 16 * sizeof(uint32_t));
            memcpy(plan_out, &plan, sizeof(plan));
        }
        if (words == nullptr) {
            *n_words = w.size();
            return;
        }
        require(*n_words >= w.size(), "program buffer too small.");
        memcpy(words, w.data(), w.size() * 4);
        *n_words = w.size();
    });
}

int gstim_create_from_text(const char *circuit_text, size_t text_len, int mode, uint64_t seed, int device, gstim_sampler **out) {
    return guarded([&] {
        require(out != nullptr, "out must not be NULL.");
        *out = nullptr;
        require(circuit_text != nullptr, "circuit_text must not be NULL.");
        require(mode == GSTIM_MODE_DETECTORS || mode == GSTIM_MODE_MEASUREMENTS, "mode must be GSTIM_MODE_DETECTORS or GSTIM_MODE_MEASUREMENTS.");
        auto s = std::make_unique<gstim_sampler>();
        s->mode = mode;
        s->seed = seed;
        s->device = device;
        s->circuit = Circuit::from_text(std::string_view(circuit_text, text_len));
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            cudaGetLastError();
            // Validate the circuit first so argument errors surface identically with or without a GPU...
            lower_circuit(s->circuit, (uint32_t)mode, 2048 - GSTIM_HDR_WORDS);
            // ...but there is no CPU fallback.
            throw CudaError("No usable CUDA device: this library has no CPU fallback.");
        }
        require(device >= 0 && device < n, "CUDA device ordinal out of range.");
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        configure(s.get());
        configure_event_engine(s.get());
        if (!s->interp_why.empty() && !s->sparse) {
            throw std::invalid_argument(s->interp_why + " Event engine: " + s->sparse_why + ".");
        }
        *out = s.release();
    });
}

int gstim_create_from_text_multi(
    const char *circuit_text, size_t text_len, int mode, uint64_t seed, const int *devices, int n_devices, gstim_sampler **out) {
    if (out != nullptr) {
        *out = nullptr;
    }
    if (devices == nullptr || n_devices < 1 || out == nullptr) {
        return guarded([&] { require(false, "devices must name at least one CUDA device."); });
    }
    for (int i = 0; i < n_devices; i++) {
        for (int j = 0; j < i; j++) {
            if (devices[i] == devices[j]) {
                return guarded([&] { require(false, "devices must be distinct."); });
            }
        }
    }
    gstim_sampler *first = nullptr;
    int rc = gstim_create_from_text(circuit_text, text_len, mode, seed, devices[0], &first);
    if (rc != GSTIM_OK) {
        return rc;
    }
    for (int i = 1; i < n_devices; i++) {
        gstim_sampler *p = nullptr;
        rc = gstim_create_from_text(circuit_text, text_len, mode, seed, devices[i], &p);
        if (rc != GSTIM_OK) {
            gstim_destroy(first);
            return rc;
        }
        first->peers.push_back(p);
    }
    cudaSetDevice(devices[0]);
    *out = first;
    return GSTIM_OK;
}

void gstim_destroy(gstim_sampler *s) {
    if (s) {
        cudaSetDevice(s->device);
        delete s;
    }
}

int gstim_get_stats(const gstim_sampler *s, gstim_stats *out) {
    return guarded([&] {
        require(s && out, "NULL argument.");
        memset(out, 0, sizeof(*out));
        out->num_qubits = s->lc.stats.num_qubits;
        out->num_measurements = s->lc.stats.num_measurements;
        out->num_detectors = s->lc.stats.num_detectors;
        out->num_observables = s->lc.stats.num_observables;
        out->max_lookback = s->lc.stats.max_lookback;
        out->active_qubits = s->lc.num_qubits;
        out->program_words = s->words.size();
        out->num_batches = s->plan.n_batches;
        out->num_barriers = s->plan.n_barriers;
        out->num_noise_sites = s->lc.num_sites;
        out->num_collapse_sites = s->lc.num_csites;
        out->threads = s->threads;
        out->lanes_per_item = 1u << s->G_log2;
        out->slots = s->slots;
        out->max_columns = s->K_max;
        out->chunk_words = s->chunk_words;
        out->smem_bytes_max = (uint32_t)interp_smem_bytes(s->plan.q_pitch, s->plan.num_qubits, s->K_max, s->chunk_words, s->n_noise);
    });
}

int gstim_get_program(const gstim_sampler *s, uint32_t *words, size_t *n_words) {
    return guarded([&] {
        require(s && n_words, "NULL argument.");
        if (words == nullptr) {
            *n_words = s->words.size();
            return;
        }
        require(*n_words >= s->words.size(), "program buffer too small.");
        memcpy(words, s->words.data(), s->words.size() * 4);
        *n_words = s->words.size();
    });
}

int gstim_set_reference_sample(gstim_sampler *s, const uint8_t *bits, size_t n_bits) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        require(s->mode == GSTIM_MODE_MEASUREMENTS, "Reference samples only apply to measurement samplers.");
        if (bits == nullptr) {
            s->ref_bits.clear();
            return;
        }
        require(n_bits == s->plan.num_meas, "reference sample must have num_measurements bits.");
        s->ref_bits.resize(n_bits);
        for (size_t k = 0; k < n_bits; k++) {
            s->ref_bits[k] = (bits[k >> 3] >> (k & 7)) & 1;
        }
    });
}

int gstim_get_shot_offset(const gstim_sampler *s, uint64_t *offset) {
    return guarded([&] {
        require(s && offset, "NULL argument.");
        *offset = s->next_col * GSTIM_COL_SHOTS;
    });
}

int gstim_set_shot_offset(gstim_sampler *s, uint64_t offset) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        require(offset % GSTIM_COL_SHOTS == 0, "shot offset must be a multiple of 128.");
        s->next_col = offset / GSTIM_COL_SHOTS;
    });
}

int gstim_sample_detectors(
    gstim_sampler *s, uint64_t shots, uint32_t flags, void *dets_out, int64_t dets_shot_stride, void *obs_out, int64_t obs_shot_stride) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        require(s->mode == GSTIM_MODE_DETECTORS, "Not a detector sampler.");
        require(dets_shot_stride >= 0 && obs_shot_stride >= 0, "negative strides are not supported.");
        check_flag_combo(flags);
        if (!s->peers.empty()) {
            sample_to_host_multi(
                s, shots, [&](gstim_sampler *d) { return detector_row_maps(d, flags); }, flags, (flags & GSTIM_BIT_PACKED) != 0,
                (uint8_t *)dets_out, dets_shot_stride, (uint8_t *)obs_out, obs_shot_stride);
            return;
        }
        RowMaps maps = detector_row_maps(s, flags);
        sample_to_host(
            s, shots, maps, flags, (flags & GSTIM_BIT_PACKED) != 0, (uint8_t *)dets_out, dets_shot_stride, (uint8_t *)obs_out, obs_shot_stride);
    });
}

int gstim_sample_measurements(gstim_sampler *s, uint64_t shots, uint32_t flags, void *out, int64_t shot_stride) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        require(s->mode == GSTIM_MODE_MEASUREMENTS, "Not a measurement sampler.");
        require(shot_stride >= 0, "negative strides are not supported.");
        if (!s->peers.empty()) {
            sample_to_host_multi(
                s, shots, [&](gstim_sampler *d) { return measurement_row_maps(d); }, 0, (flags & GSTIM_BIT_PACKED) != 0, (uint8_t *)out,
                shot_stride, nullptr, 0);
            return;
        }
        RowMaps maps = measurement_row_maps(s);
        sample_to_host(s, shots, maps, 0, (flags & GSTIM_BIT_PACKED) != 0, (uint8_t *)out, shot_stride, nullptr, 0);
    });
}

int gstim_sample_detectors_device(
    gstim_sampler *s, uint64_t shots, uint32_t flags, void *dets_out_dev, int64_t dets_shot_stride, void *obs_out_dev, int64_t obs_shot_stride) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        require(s->mode == GSTIM_MODE_DETECTORS, "Not a detector sampler.");
        require(dets_shot_stride >= 0 && obs_shot_stride >= 0, "negative strides are not supported.");
        check_flag_combo(flags);
        RowMaps maps = detector_row_maps(s, flags);
        sample_to_device(s, shots, maps, flags, (uint8_t *)dets_out_dev, dets_shot_stride, (uint8_t *)obs_out_dev, obs_shot_stride);
    });
}

int gstim_sample_measurements_device(gstim_sampler *s, uint64_t shots, void *out_dev, int64_t shot_stride) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        require(s->mode == GSTIM_MODE_MEASUREMENTS, "Not a measurement sampler.");
        require(shot_stride >= 0, "negative strides are not supported.");
        RowMaps maps = measurement_row_maps(s);
        sample_to_device(s, shots, maps, 0, (uint8_t *)out_dev, shot_stride, nullptr, 0);
    });
}

int gstim_sample_detectors_to_fd(
    gstim_sampler *s, uint64_t shots, uint32_t flags, int fd, const char *format, int obs_fd, const char *obs_format) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        require(s->mode == GSTIM_MODE_DETECTORS, "Not a detector sampler.");
        Format fmt = parse_format(format);
        Format ofmt = obs_fd >= 0 ? parse_format(obs_format) : Format::F01;
        uint32_t f = flags & (GSTIM_PREPEND_OBS | GSTIM_APPEND_OBS);
        if (obs_fd >= 0) {
            f |= GSTIM_SEPARATE_OBS;
        }
        check_flag_combo(f);
        RowMaps maps = detector_row_maps(s, f);
        FdFile out(fd);
        std::unique_ptr<FdFile> obs;
        if (obs_fd >= 0) {
            obs = std::make_unique<FdFile>(obs_fd);
        }
        char c1 = 'D', c2 = 'L';
        size_t tr = s->plan.num_det;
        if (f & GSTIM_PREPEND_OBS) {
            c1 = 'L';
            c2 = 'D';
            tr = s->plan.num_obs;
        }
        sample_to_file(s, shots, f, maps.main, out.f, fmt, c1, c2, tr, obs ? &maps.obs : nullptr, obs ? obs->f : nullptr, ofmt);
    });
}

int gstim_sample_measurements_to_fd(gstim_sampler *s, uint64_t shots, int fd, const char *format) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        require(s->mode == GSTIM_MODE_MEASUREMENTS, "Not a measurement sampler.");
        Format fmt = parse_format(format);
        RowMaps maps = measurement_row_maps(s);
        FdFile out(fd);
        sample_to_file(s, shots, 0, maps.main, out.f, fmt, 'M', 'M', maps.main.size(), nullptr, nullptr, Format::F01);
    });
}

int gstim_write_shots_to_fd(
    const uint8_t *rows,
    size_t row_pitch,
    uint64_t shots,
    uint64_t n_bits,
    int fd,
    const char *format,
    char prefix1,
    char prefix2,
    uint64_t prefix_transition) {
    return guarded([&] {
        require(rows != nullptr || shots == 0 || n_bits == 0, "NULL rows.");
        Format fmt = parse_format(format);
        FdFile out(fd);
        if (fmt == Format::PTB64) {
            // shot-major packed rows -> bit-major 32-bit rows, then the ptb64 writer
            if (shots % 64 != 0) {
                throw std::invalid_argument("shots must be a multiple of 64 to use ptb64 format.");
            }
            write_ptb64_rows(out.f, rows, row_pitch, shots, n_bits);
        } else {
            write_shots(out.f, rows, row_pitch, shots, n_bits, fmt, prefix1, prefix2, prefix_transition);
        }
        if (fflush(out.f) != 0) {
            throw IoError("Failed to flush result data.");
        }
    });
}

int gstim_bit_counts(gstim_sampler *s, uint64_t shots, uint64_t *single_host, uint64_t *pair_host, void *single_dev, void *pair_dev) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        const bool want_pairs = pair_host != nullptr || pair_dev != nullptr;
        RowMaps maps = s->mode == GSTIM_MODE_DETECTORS ? detector_row_maps(s, GSTIM_APPEND_OBS) : measurement_row_maps(s);
        const uint32_t n_bits = (uint32_t)maps.main.size();
        CK(cudaSetDevice(s->device));
        s->d_rowmap.ensure((size_t)(n_bits + 1) * 4);
        upload_row_map(s, maps.main, 0);
        s->d_counts.ensure((size_t)std::max<uint32_t>(n_bits, 1) * 16);
        CK(cudaMemsetAsync(s->d_counts.p, 0, (size_t)std::max<uint32_t>(n_bits, 1) * 16, s->stream));
        unsigned long long *d_single = (unsigned long long *)s->d_counts.p, *d_pair = d_single + n_bits;
        if (use_events(s)) {
            // rows of one chunk stay in device staging and are reduced there
            const uint64_t nbytes = (n_bits + 7) / 8;
            const uint64_t chunk = std::max<uint64_t>((512ull << 20) / std::max<uint64_t>(nbytes, 1), 128);
            uint8_t *rows = nullptr;
            run_events(
                s, shots, s->mode == GSTIM_MODE_DETECTORS ? GSTIM_APPEND_OBS : 0u, chunk,
                [&](uint64_t, uint64_t n, uint8_t **m, uint64_t *mp, uint8_t **o, uint64_t *op) {
                    s->d_stage[0].ensure(n * nbytes + 16);
                    rows = (uint8_t *)s->d_stage[0].p;
                    *m = rows;
                    *mp = nbytes;
                    *o = nullptr;
                    *op = 0;
                },
                [&](uint64_t, uint64_t n) {
                    CK(launch_count_b8(rows, nbytes, n, n_bits, d_single, want_pairs ? d_pair : nullptr, s->stream));
                    s->last_launches++;
                });
        } else {
        run_sampler(s, shots, [&](uint64_t first, uint64_t n, const uint32_t *table, uint64_t n_rows) {
            (void)first;
            CK(launch_bit_counts(table, n_rows, n, (const uint32_t *)s->d_rowmap.p, n_bits, d_single, want_pairs ? d_pair : nullptr, s->stream));
            s->last_launches++;
        });
        }
        const size_t nb1 = (size_t)n_bits * 8, nb2 = n_bits ? (size_t)(n_bits - 1) * 8 : 0;
        if (single_dev) {
            CK(cudaMemcpyAsync(single_dev, d_single, nb1, cudaMemcpyDeviceToDevice, s->stream));
        }
        if (pair_dev && nb2) {
            CK(cudaMemcpyAsync(pair_dev, d_pair, nb2, cudaMemcpyDeviceToDevice, s->stream));
        }
        if (single_host) {
            CK(cudaMemcpyAsync(single_host, d_single, nb1, cudaMemcpyDeviceToHost, s->stream));
        }
        if (pair_host && nb2) {
            CK(cudaMemcpyAsync(pair_host, d_pair, nb2, cudaMemcpyDeviceToHost, s->stream));
        }
        CK(cudaStreamSynchronize(s->stream));
    });
}

int gstim_detector_flip_counts(gstim_sampler *s, uint64_t shots, uint64_t *counts_host, void *counts_dev) {
    if (s != nullptr && s->mode != GSTIM_MODE_DETECTORS) {
        return guarded([&] { require(false, "Not a detector sampler."); });
    }
    return gstim_bit_counts(s, shots, counts_host, nullptr, counts_dev, nullptr);
}

int gstim_measure_lop3_peak(int device, double *lane_ops_per_clk_per_sm, double *lane_ops_per_sec, double *sm_mhz) {
    return guarded([&] {
        require(lane_ops_per_clk_per_sm && lane_ops_per_sec && sm_mhz, "NULL argument.");
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        CK(measure_lop3_peak(prop.multiProcessorCount, lane_ops_per_clk_per_sm, lane_ops_per_sec, sm_mhz));
    });
}

int gstim_set_engine(gstim_sampler *s, int engine) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        require(engine == GSTIM_ENGINE_AUTO || engine == GSTIM_ENGINE_INTERPRETER || engine == GSTIM_ENGINE_EVENTS, "bad engine.");
        if (engine == GSTIM_ENGINE_EVENTS && !s->sparse) {
            throw std::invalid_argument("The event engine cannot sample this circuit: " + s->sparse_why + ".");
        }
        s->engine_pref = engine;
    });
}

int gstim_get_engine_info(const gstim_sampler *s, gstim_engine_info *out) {
    return guarded([&] {
        require(s && out, "NULL argument.");
        memset(out, 0, sizeof(*out));
        out->last_engine = s->last_engine;
        if (!s->sparse) {
            snprintf(out->why_not, sizeof(out->why_not), "%s", s->sparse_why.c_str());
            return;
        }
        fill_table_info(s->sparse->table(), out);
        out->favoured = s->sparse_favoured ? 1 : 0;
        out->tile_shots = s->sparse->tile_shots();
        out->blocks_per_sm = s->sparse->blocks_per_sm();
        out->num_slices = (uint32_t)(s->sparse->slices().size() / 4);
        out->device_entries = s->sparse->device_table_entries();
    });
}

int gstim_get_response_table(const gstim_sampler *s, int what, uint32_t *words, size_t *n_words) {
    return guarded([&] {
        require(s && n_words, "NULL argument.");
        require(s->sparse != nullptr, "The circuit has no response table.");
        export_table_array(s->sparse->table(), &s->sparse->slices(), s->mode == GSTIM_MODE_DETECTORS ? s->plan.num_det : 0, what, words, n_words);
    });
}

struct gstim_response_table {
    ResponseTable rt;
    uint32_t n_det = 0;
};

int gstim_response_table_create(const char *circuit_text, size_t text_len, int mode, gstim_response_table **out) {
    return guarded([&] {
        require(circuit_text != nullptr && out != nullptr, "NULL argument.");
        require(mode == GSTIM_MODE_DETECTORS || mode == GSTIM_MODE_MEASUREMENTS, "bad mode.");
        *out = nullptr;
        Circuit c = Circuit::from_text(std::string_view(circuit_text, text_len));
        LoweredCircuit lc = lower_circuit(c, (uint32_t)mode, 2048 - GSTIM_HDR_WORDS);
        auto t = std::make_unique<gstim_response_table>();
        t->rt = build_response_table(lc);
        t->n_det = mode == GSTIM_MODE_DETECTORS ? (uint32_t)lc.stats.num_detectors : 0;
        *out = t.release();
    });
}

void gstim_response_table_destroy(gstim_response_table *t) {
    delete t;
}

int gstim_response_table_info(const gstim_response_table *t, gstim_engine_info *out) {
    return guarded([&] {
        require(t && out, "NULL argument.");
        memset(out, 0, sizeof(*out));
        fill_table_info(t->rt, out);
        if (!t->rt.eligible) {
            snprintf(out->why_not, sizeof(out->why_not), "%s", t->rt.why_not.c_str());
        }
    });
}

int gstim_response_table_get(const gstim_response_table *t, int what, uint32_t *words, size_t *n_words) {
    return guarded([&] {
        require(t && n_words, "NULL argument.");
        require(t->rt.eligible, "The circuit has no response table.");
        export_table_array(t->rt, nullptr, t->n_det, what, words, n_words);
    });
}

int gstim_set_block_columns(gstim_sampler *s, uint32_t columns) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        const uint32_t G = 1u << s->G_log2;
        require(columns == 0 || (columns % G == 0 && columns <= s->K_max), "columns must be 0 (automatic) or a multiple of lanes_per_item up to max_columns.");
        s->K_fixed = columns;
    });
}

int gstim_last_launch_count(const gstim_sampler *s, uint64_t *launches) {
    return guarded([&] {
        require(s && launches, "NULL argument.");
        *launches = s->last_launches;
    });
}

int gstim_last_block_columns(const gstim_sampler *s, uint32_t *columns) {
    return guarded([&] {
        require(s && columns, "NULL argument.");
        *columns = s->last_K;
    });
}

int gstim_last_call_ms(const gstim_sampler *s, float *ms) {
    return guarded([&] {
        require(s && ms, "NULL argument.");
        *ms = s->last_call_ms;
    });
}

int gstim_last_kernel_ms(const gstim_sampler *s, float *interp_ms, float *transpose_ms) {
    return guarded([&] {
        require(s != nullptr, "NULL sampler.");
        if (interp_ms) {
            *interp_ms = s->last_interp_ms;
        }
        if (transpose_ms) {
            *transpose_ms = s->last_transpose_ms;
        }
    });
}

}  // extern "C"
