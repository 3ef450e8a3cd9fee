// dem.cu — detector-error-model sampler: the same Bernoulli-row -> XOR -> transpose primitives as
// the circuit sampler, for a DetectorErrorModel instead of a Circuit.
//
// Replaces DemSampler<W>::resample / sample_write (/root/reference/src/stim/simulators/dem_sampler.inl:52-130) behind
// `stim sample_dem` (/root/reference/src/stim/cmd/command_sample_dem.cc:25-93) and stim.CompiledDemSampler.sample
// (/root/reference/src/stim/simulators/dem_sampler.pybind.cc). Each error mechanism `error(p) D.. L..` is an independent
// Bernoulli(p) row over the shots; its targets' rows (detectors, observables) are XORed with it.
//
// Device side: one thread walks one error mechanism over the K * 128 shots of a shot block with geometric gaps (the
// 32-bit gap arithmetic below: floor(Exp(1) / lambda) == Geometric(p), the reference's RareErrorIterator,
// probability_util.cc:33-43) and flips the bits of the mechanism's target rows in the
// block's columns of a column-major bit table with red.global.xor (the table stays L2-resident). The table then goes
// through the same transposer / counters / writers as the circuit sampler's.
//   Shot blocks are 32 columns (4096 shots) wide, whatever the request. Philox counter of call c of error e in the shot
//   block whose first column is col0:
//       (e, 'DEMS', col0 lo, col0 hi | c << 15)  ->  four gap words (draws 4c .. 4c + 3).
//   Gap arithmetic (all integer, restated bit for bit by oracle/philox.py: exp_draw_q26 / gap_of):
//       E   = -ln(v / 2^32), v = word | 1, in units of 2^-26 nat: 256-entry log2 table (Q26 base + forward difference),
//             13-bit linear interpolation, one multiply-high by ln 2 (Q32)
//       gap = (E * INV) >> SH as a 64-bit product, INV = floor(2^32 m), SH = 58 - e for 1 / lambda = m 2^e, lambda = -log1p(-p)
//       p >= 1: INV = 0 (an event at every shot); p below 2^-58 is treated as 0.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/gstim.h"
#include "kernels.cuh"
#include "lowering.h"
#include "hostpipe.h"
#include "sparse.cuh"
#include "writers.h"

#define GSTIM_TABLE_QUAL __device__ const
#include "log2_q26_table.h"

namespace gstim {

#define GTAG_DEM 0x44454D53u
#define GSTIM_DEM_BLOCK_COLS 32u

// ------------------------------------------------------------------------------------------------
// host: the model
// ------------------------------------------------------------------------------------------------
struct DemModel {
    uint64_t num_detectors = 0, num_observables = 0;
    std::vector<double> probs;        // per error mechanism, in flattened order
    std::vector<uint32_t> tgt_off;    // CSR offsets into tgt (size E + 1)
    std::vector<uint32_t> tgt;        // detector id, or 0x80000000 | observable id
};

namespace {

struct DemReader {
    std::string_view s;
    size_t i = 0;
    DemModel m;
    uint64_t det_offset = 0;
    static constexpr size_t MAX_ERRORS = 1u << 27;

    [[noreturn]] void fail(const std::string &why) const {
        throw std::invalid_argument("Detector error model parse error: " + why);
    }
    void skip_space() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r')) {
            i++;
        }
    }
    void skip_to_eol() {
        while (i < s.size() && s[i] != '\n') {
            i++;
        }
    }
    struct Line {
        std::string name;
        std::vector<double> args;
        std::vector<std::string> targets;
        bool block_open = false;
    };
    // Reads one logical line; returns false at the end of input or at a '}'.
    bool read_line(Line &ln, bool &closed) {
        closed = false;
        while (true) {
            skip_space();
            if (i >= s.size()) {
                return false;
            }
            if (s[i] == '\n') {
                i++;
                continue;
            }
            if (s[i] == '#') {
                skip_to_eol();
                continue;
            }
            break;
        }
        if (s[i] == '}') {
            i++;
            closed = true;
            return false;
        }
        ln = Line();
        while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_')) {
            ln.name.push_back((char)tolower((unsigned char)s[i++]));
        }
        if (ln.name.empty()) {
            fail("expected an instruction name");
        }
        if (i < s.size() && s[i] == '[') {  // tag
            while (i < s.size() && s[i] != ']' && s[i] != '\n') {
                i++;
            }
            if (i >= s.size() || s[i] != ']') {
                fail("unterminated tag");
            }
            i++;
        }
        skip_space();
        if (i < s.size() && s[i] == '(') {
            i++;
            std::string num;
            while (i < s.size() && s[i] != ')' && s[i] != '\n') {
                if (s[i] == ',') {
                    if (!num.empty()) {
                        ln.args.push_back(strtod(num.c_str(), nullptr));
                    }
                    num.clear();
                } else if (s[i] != ' ' && s[i] != '\t') {
                    num.push_back(s[i]);
                }
                i++;
            }
            if (i >= s.size() || s[i] != ')') {
                fail("unterminated argument list");
            }
            i++;
            if (!num.empty()) {
                ln.args.push_back(strtod(num.c_str(), nullptr));
            }
        }
        while (true) {
            skip_space();
            if (i >= s.size() || s[i] == '\n' || s[i] == '#') {
                break;
            }
            if (s[i] == '{') {
                ln.block_open = true;
                i++;
                continue;
            }
            std::string t;
            while (i < s.size() && !isspace((unsigned char)s[i]) && s[i] != '{' && s[i] != '#') {
                t.push_back(s[i++]);
            }
            ln.targets.push_back(t);
        }
        return true;
    }
    static uint64_t parse_id(const std::string &t, size_t from) {
        if (from >= t.size()) {
            throw std::invalid_argument("Detector error model parse error: bad target '" + t + "'");
        }
        uint64_t v = 0;
        for (size_t k = from; k < t.size(); k++) {
            if (!isdigit((unsigned char)t[k])) {
                throw std::invalid_argument("Detector error model parse error: bad target '" + t + "'");
            }
            v = v * 10 + (uint64_t)(t[k] - '0');
            if (v >= (1ull << 31)) {
                throw std::invalid_argument("Detector error model parse error: target index too large in '" + t + "'");
            }
        }
        return v;
    }

    // Executes the block starting at the current position `reps` times (flattening, like
    // DetectorErrorModel::iter_flatten_error_instructions, /root/reference/src/stim/dem/detector_error_model.h).
    void run_block(bool top) {
        while (true) {
            Line ln;
            bool closed;
            if (!read_line(ln, closed)) {
                if (closed == top) {
                    fail(top ? "unmatched '}'" : "missing '}'");
                }
                return;
            }
            if (ln.name == "repeat") {
                if (ln.targets.size() != 1 || !ln.block_open) {
                    fail("repeat needs a count and a '{'");
                }
                const uint64_t reps = parse_id(ln.targets[0], 0);
                const size_t body = i;
                if (reps == 0) {
                    fail("repeat 0 is not allowed");
                }
                for (uint64_t r = 0; r < reps; r++) {
                    i = body;
                    run_block(false);
                }
                continue;
            }
            if (ln.block_open) {
                fail("unexpected '{'");
            }
            if (ln.name == "error") {
                if (ln.args.size() != 1 || !(ln.args[0] >= 0 && ln.args[0] <= 1)) {
                    fail("error needs one probability argument in [0, 1]");
                }
                if (m.probs.size() >= MAX_ERRORS) {
                    fail("more than 2^27 error mechanisms after flattening");
                }
                m.probs.push_back(ln.args[0]);
                for (const std::string &t : ln.targets) {
                    if (t == "^") {
                        continue;  // decomposition separators do not matter for sampling
                    }
                    if (t[0] == 'D' || t[0] == 'd') {
                        const uint64_t id = parse_id(t, 1) + det_offset;
                        if (id >= (1ull << 31)) {
                            fail("detector index too large");
                        }
                        m.num_detectors = std::max(m.num_detectors, id + 1);
                        m.tgt.push_back((uint32_t)id);
                    } else if (t[0] == 'L' || t[0] == 'l') {
                        const uint64_t id = parse_id(t, 1);
                        m.num_observables = std::max(m.num_observables, id + 1);
                        m.tgt.push_back(0x80000000u | (uint32_t)id);
                    } else {
                        fail("bad error target '" + t + "'");
                    }
                }
                m.tgt_off.push_back((uint32_t)m.tgt.size());
            } else if (ln.name == "detector") {
                for (const std::string &t : ln.targets) {
                    if (t[0] != 'D' && t[0] != 'd') {
                        fail("detector takes D targets");
                    }
                    m.num_detectors = std::max(m.num_detectors, parse_id(t, 1) + det_offset + 1);
                }
            } else if (ln.name == "logical_observable") {
                for (const std::string &t : ln.targets) {
                    if (t[0] != 'L' && t[0] != 'l') {
                        fail("logical_observable takes L targets");
                    }
                    m.num_observables = std::max(m.num_observables, parse_id(t, 1) + 1);
                }
            } else if (ln.name == "shift_detectors") {
                if (ln.targets.size() != 1) {
                    fail("shift_detectors needs one integer target");
                }
                det_offset += parse_id(ln.targets[0], 0);
            } else {
                fail("unknown instruction '" + ln.name + "'");
            }
        }
    }
};

}  // namespace

DemModel parse_dem(std::string_view text) {
    DemReader r;
    r.s = text;
    r.m.tgt_off.push_back(0);
    r.run_block(true);
    if (r.m.tgt.size() >= (1ull << 31)) {
        throw std::invalid_argument("Detector error model too large.");
    }
    return std::move(r.m);
}

// ------------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------------
struct DemParams {
    const uint2 *rates;       // per error: INV, SH | 0x80000000 when the error can fire
    const uint32_t *tgt_off;  // E + 1
    const uint32_t *tgt_row;  // table row of every target
    uint32_t n_errors;
    uint32_t rows;            // table rows: D + L (+ E when the errors are recorded)
    uint32_t err_row0;        // first error row, or 0xFFFFFFFF
    uint32_t K;               // 128-shot columns per shot block
    uint32_t n_blocks;
    uint64_t col0_base;       // global column of block 0
    uint32_t seed_lo, seed_hi;
    uint4 *table;             // column-major: table[column * rows + row]
};

__device__ __forceinline__ uint4 dem_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__global__ void __launch_bounds__(256, 4) gstim_dem_kernel(const DemParams p) {
    __shared__ uint2 lt[256];  // log2 table: (base, diff) in Q26
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        lt[i] = make_uint2(GSTIM_LOG2_Q26[2 * i], GSTIM_LOG2_Q26[2 * i + 1]);
    }
    const uint32_t B = p.K * GSTIM_COL_SHOTS;
    for (uint32_t g = blockIdx.x; g < p.n_blocks; g += gridDim.x) {
        uint4 *const cols = p.table + (uint64_t)g * p.K * p.rows;
        for (uint64_t i = threadIdx.x; i < (uint64_t)p.K * p.rows; i += blockDim.x) {
            cols[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();  // (the zeros of this block's columns precede every flip; also covers the table load above)
        const uint64_t col0 = p.col0_base + (uint64_t)g * p.K;
        const uint32_t c2 = (uint32_t)col0, c3 = (uint32_t)(col0 >> 32);
        for (uint32_t e = threadIdx.x; e < p.n_errors; e += blockDim.x) {
            const uint2 rate = p.rates[e];
            if (!(rate.y & 0x80000000u)) {
                continue;  // p == 0
            }
            const uint32_t inv = rate.x, sh = rate.y & 63u;
            const uint32_t t0 = p.tgt_off[e], t1 = p.tgt_off[e + 1];
            uint32_t a = 0;
            for (uint32_t call = 0;; call++) {
                const uint4 rr = dem_philox(e, GTAG_DEM, c2, c3 | (call << GSTIM_DRAW_SHIFT), p.seed_lo, p.seed_hi);
                const uint32_t words[4] = {rr.x, rr.y, rr.z, rr.w};
                bool done = false;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (done) {
                        break;
                    }
                    // exp_draw_q26 (header comment)
                    const uint32_t v = words[j] | 1u;
                    const uint32_t t = 31u - (uint32_t)__clz((int)v);
                    const uint32_t frac = (v << (31u - t)) << 1;
                    const uint2 en = lt[frac >> 24];
                    const uint32_t log2v = (t << 26) + en.x + ((en.y * ((frac >> 11) & 0x1FFFu)) >> 13);
                    const uint32_t E = __umulhi(0x80000000u - log2v, GSTIM_LN2_Q32);
                    const unsigned long long G = ((unsigned long long)E * inv) >> sh;
                    if (G >= (unsigned long long)(B - a)) {
                        done = true;
                        break;
                    }
                    a += (uint32_t)G;
                    const uint32_t shot = a++;
                    uint32_t *const col = (uint32_t *)(cols + (uint64_t)(shot >> 7) * p.rows) + ((shot >> 5) & 3u);
                    const uint32_t bit = 1u << (shot & 31u);
                    for (uint32_t k = t0; k < t1; k++) {
                        asm volatile("red.global.xor.b32 [%0], %1;" ::"l"(col + 4ull * p.tgt_row[k]), "r"(bit) : "memory");
                    }
                    if (p.err_row0 != 0xFFFFFFFFu) {
                        asm volatile("red.global.xor.b32 [%0], %1;" ::"l"(col + 4ull * (p.err_row0 + e)), "r"(bit) : "memory");
                    }
                }
                if (done) {
                    break;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace gstim

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace gstim;

namespace {

struct DevMem {
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
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            throw std::runtime_error(std::string("CUDA out of memory (") + cudaGetErrorString(e) + ")");
        }
        cap = bytes;
    }
    ~DevMem() {
        if (p) {
            cudaFree(p);
        }
    }
};

void ck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        throw std::runtime_error(std::string("CUDA error '") + cudaGetErrorString(e) + "' in " + what);
    }
}

}  // namespace

struct gstim_dem_sampler {
    int device = 0;
    uint64_t seed = 0;
    uint64_t next_col = 0;
    DemModel model;
    // Event engine (sparse.cu): a detector error model IS a response table (one site per error mechanism, its targets the
    // response). Used whenever the fired errors themselves are not asked for; built on first use, null if not possible
    // (more than 64 distinct probabilities, rows beyond shared memory).
    std::unique_ptr<SparseEngine> events;
    bool events_tried = false;
    int engine_pref = GSTIM_ENGINE_AUTO;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    DevMem d_rates, d_tgt_off, d_tgt_row, d_table, d_rowmap, d_stage;
    gstim_m2d *replay = nullptr;  // recorded errors -> detectors / observables (m2d.cu), built on first use
    HostStager host_stage;        // page-locked staging pair of the host-output paths (hostpipe.h)
    ~gstim_dem_sampler() {
        if (replay) {
            gstim_m2d_destroy(replay);
        }
        if (stream) {
            cudaStreamDestroy(stream);
        }
    }
};
gstim_m2d *gstim_m2d_from_lists(int device, uint64_t n_inputs, uint64_t D, uint64_t L, std::vector<std::vector<uint32_t>> recs);

// (gstim_last_error lives in api.cu; DEM errors are reported through the same thread-local channel)
void gstim_set_last_error(const char *msg);
void gstim_export_table_array(const gstim::ResponseTable &rt, const std::vector<uint32_t> *slices, uint32_t n_det, int what, uint32_t *words,
                              size_t *n_words);

namespace {

template <typename F>
int dem_guarded(F &&f) {
    try {
        f();
        return GSTIM_OK;
    } catch (const std::invalid_argument &e) {
        gstim_set_last_error(e.what());
        return GSTIM_ERR_INVALID_ARGUMENT;
    } catch (const std::out_of_range &e) {
        gstim_set_last_error(e.what());
        return GSTIM_ERR_OUT_OF_RANGE;
    } catch (const std::exception &e) {
        gstim_set_last_error(e.what());
        return std::string(e.what()).rfind("CUDA", 0) == 0 ? GSTIM_ERR_CUDA : GSTIM_ERR_INTERNAL;
    }
}

// The model as a response table: mechanisms grouped into classes of equal probability (ascending), model order inside a
// class, one outcome each, response = detector ids then D + observable ids (a target named twice cancels).
ResponseTable dem_response_table(const DemModel &m) {
    ResponseTable rt;
    rt.n_outputs = (uint32_t)(m.num_detectors + m.num_observables);
    const size_t E = m.probs.size();
    std::vector<uint64_t> keys(E);
    std::vector<uint32_t> order;
    for (size_t e = 0; e < E; e++) {
        keys[e] = gstim_rate_key(m.probs[e]);
        if (keys[e] != 0) {
            order.push_back((uint32_t)e);
        }
    }
    auto lam_of = [](double p) -> uint64_t {
        const float f = (float)p;
        if (f >= 1) {
            return 1ull << 62;
        }
        return (uint64_t)std::ldexp(-std::log1p(-(double)f), 56);
    };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return lam_of(m.probs[a]) < lam_of(m.probs[b]); });
    std::vector<uint32_t> ids;
    for (size_t i = 0; i < order.size();) {
        const uint32_t e0 = order[i];
        RespClass rc;
        rc.lam = lam_of(m.probs[e0]);
        rc.inv = (keys[e0] & (1ull << 63)) && (float)m.probs[e0] < 1 ? (uint32_t)((keys[e0] & ~(1ull << 63)) >> 8) : 0u;
        rc.sh = (float)m.probs[e0] < 1 ? (uint32_t)(keys[e0] & 0xFF) : 0u;
        rc.kind = RK_SINGLE;
        rc.n_out = 1;
        rc.entry0 = (uint32_t)(rt.entries.size() / 4);
        const double pr = rc.lam >= (1ull << 62) ? 1.0 : -std::expm1(-std::ldexp((double)rc.lam, -56));
        size_t j = i;
        while (j < order.size() && lam_of(m.probs[order[j]]) == rc.lam) {
            const uint32_t e = order[j];
            ids.clear();
            for (uint32_t k = m.tgt_off[e]; k < m.tgt_off[e + 1]; k++) {
                const uint32_t t = m.tgt[k];
                ids.push_back((t & 0x80000000u) ? (uint32_t)m.num_detectors + (t & 0x7FFFFFFFu) : t);
            }
            std::sort(ids.begin(), ids.end());
            size_t n = 0;
            for (size_t a = 0; a < ids.size(); a++) {  // pairs cancel
                if (a + 1 < ids.size() && ids[a] == ids[a + 1]) {
                    a++;
                    continue;
                }
                ids[n++] = ids[a];
            }
            ids.resize(n);
            uint32_t w[4] = {RESP_NONE, RESP_NONE, RESP_NONE, RESP_NONE};
            if (n <= 4) {
                for (size_t a = 0; a < n; a++) {
                    w[a] = ids[a];
                }
            } else {
                for (size_t a = 0; a < 3; a++) {
                    w[a] = ids[a];
                }
                w[3] = RESP_OVERFLOW | (uint32_t)rt.overflow.size();
                rt.overflow.push_back((uint32_t)(n - 3));
                rt.overflow.insert(rt.overflow.end(), ids.begin() + 3, ids.end());
            }
            rt.entries.insert(rt.entries.end(), w, w + 4);
            rt.site_group.push_back(e);  // (provenance: the mechanism's index in the flattened model)
            rt.site_index.push_back(0);
            rt.max_response = std::max<uint32_t>(rt.max_response, (uint32_t)n);
            rt.flips_per_shot += pr * (double)n;
            rc.n_sites++;
            j++;
        }
        rt.events_per_shot += pr * rc.n_sites;
        rt.outcome_word.push_back(0);
        rt.n_sites += rc.n_sites;
        rt.classes.push_back(rc);
        i = j;
    }
    rt.n_entries = rt.entries.size() / 4;
    rt.eligible = true;
    return rt;
}

// The event engine of a sampler, or null (more than 64 classes / rows too long / switched off).
SparseEngine *dem_events(gstim_dem_sampler *s) {
    if (s->engine_pref == GSTIM_ENGINE_INTERPRETER) {
        return nullptr;
    }
    if (!s->events_tried) {
        s->events_tried = true;
        const char *env = getenv("GSTIM_ENGINE");
        if (env != nullptr && (strcmp(env, "interp") == 0 || strcmp(env, "interpreter") == 0)) {
            return nullptr;
        }
        try {
            const uint32_t D = (uint32_t)s->model.num_detectors, L = (uint32_t)s->model.num_observables;
            s->events = std::make_unique<SparseEngine>(dem_response_table(s->model), 0u, D, L, 0u, s->device, 0u, 0u, 1u << 20);
        } catch (const std::invalid_argument &) {
            s->events.reset();
        }
    }
    return s->events.get();
}

// Event-engine pass: dense b8 rows [dets | obs appended] per chunk in device staging, handed to sink(first, n, rows, pitch).
template <typename SINK>
void dem_run_events(gstim_dem_sampler *s, SparseEngine &E, uint64_t shots, SINK &&sink) {
    ck(cudaSetDevice(s->device), "cudaSetDevice");
    if (shots == 0) {
        return;
    }
    const uint64_t cols = (shots + GSTIM_COL_SHOTS - 1) / GSTIM_COL_SHOTS;
    if (s->next_col + cols >= (1ull << 47)) {
        throw std::invalid_argument("shot offset + shots must stay below 2^54");
    }
    E.set_layout(GSTIM_APPEND_OBS, s->stream);
    const uint64_t pitch = (E.main_bits() + 7) / 8;
    const uint64_t chunk = std::max<uint64_t>(((256ull << 20) / std::max<uint64_t>(pitch, 1)) / GSTIM_COL_SHOTS, 1) * GSTIM_COL_SHOTS;
    const uint64_t base = s->next_col * GSTIM_COL_SHOTS;
    for (uint64_t first = 0; first < shots; first += chunk) {
        const uint64_t n = std::min(chunk, shots - first);
        s->d_stage.ensure(n * pitch + 16);
        try {
            E.launch(base + first, n, (uint8_t *)s->d_stage.p, pitch, nullptr, 0, s->seed, s->stream);
        } catch (const std::runtime_error &e) {
            throw std::runtime_error(std::string("CUDA: ") + e.what());
        }
        sink(first, n, (const uint8_t *)s->d_stage.p, pitch);
    }
    ck(cudaStreamSynchronize(s->stream), "cudaStreamSynchronize");
    s->next_col += cols;
}

// bits [bit0, bit0 + n_bits) of packed rows -> packed (or one byte per bit) rows of the caller
void dem_slice_rows(const uint8_t *rows, uint64_t pitch, uint64_t n, uint32_t bit0, uint32_t n_bits, bool packed, uint8_t *dst0, uint64_t dst_pitch) {
    const uint64_t out_bytes = (n_bits + 7) / 8;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *r = rows + i * pitch;
        uint8_t *dst = dst0 + i * dst_pitch;
        if (!packed) {
            for (uint32_t b = 0; b < n_bits; b++) {
                dst[b] = (r[(bit0 + b) >> 3] >> ((bit0 + b) & 7)) & 1;
            }
        } else if ((bit0 & 7) == 0) {
            memcpy(dst, r + (bit0 >> 3), out_bytes);
            if (n_bits & 7) {
                dst[out_bytes - 1] &= (uint8_t)((1u << (n_bits & 7)) - 1);
            }
        } else {
            memset(dst, 0, out_bytes);
            for (uint32_t b = 0; b < n_bits; b++) {
                dst[b >> 3] |= (uint8_t)(((r[(bit0 + b) >> 3] >> ((bit0 + b) & 7)) & 1) << (b & 7));
            }
        }
    }
}

struct DemOut {
    uint8_t *ptr = nullptr;   // host buffer or null
    int64_t stride = 0;
    uint32_t row0 = 0, n_bits = 0;
};

// Samples `shots` shots; for every chunk calls sink(first_shot, n_shots, table, rows) with the kernel enqueued.
template <typename SINK>
void dem_run(gstim_dem_sampler *s, uint64_t shots, bool record_errors, SINK &&sink) {
    ck(cudaSetDevice(s->device), "cudaSetDevice");
    if (shots == 0) {
        return;
    }
    const DemModel &m = s->model;
    const uint32_t E = (uint32_t)m.probs.size();
    const uint64_t DL = m.num_detectors + m.num_observables;
    const uint64_t rows64 = std::max<uint64_t>(DL + (record_errors ? E : 0), 1);
    if (rows64 >= (1ull << 31)) {
        throw std::invalid_argument("Detector error model has too many rows.");
    }
    const uint32_t rows = (uint32_t)rows64;
    const uint64_t cols = (shots + GSTIM_COL_SHOTS - 1) / GSTIM_COL_SHOTS;
    // A shot block is always 32 columns (4096 shots): the random stream then depends on (seed, shot offset) only.
    const uint32_t K = GSTIM_DEM_BLOCK_COLS;
    const uint64_t bytes_per_block = (uint64_t)rows * K * 16;
    const uint64_t total_blocks = (cols + K - 1) / K;
    const uint64_t budget = 1ull << 30;
    const uint64_t max_blocks = std::min<uint64_t>(std::max<uint64_t>(budget / bytes_per_block, 1), total_blocks);
    s->d_table.ensure(bytes_per_block * max_blocks);
    // table rows of the targets: detectors first, then observables
    if (s->d_tgt_row.p == nullptr) {
        std::vector<uint2> rates(std::max<uint32_t>(E, 1));
        for (uint32_t e = 0; e < E; e++) {
            const uint64_t key = gstim_rate_key(m.probs[e]);
            rates[e] = key ? make_uint2((uint32_t)(key >> 8), (uint32_t)(key & 0xFF) | 0x80000000u) : make_uint2(0, 0);
        }
        std::vector<uint32_t> rowsv(std::max<size_t>(m.tgt.size(), 1));
        for (size_t k = 0; k < m.tgt.size(); k++) {
            rowsv[k] = (m.tgt[k] & 0x80000000u) ? (uint32_t)(m.num_detectors + (m.tgt[k] & 0x7FFFFFFFu)) : m.tgt[k];
        }
        s->d_rates.ensure(rates.size() * 8);
        s->d_tgt_off.ensure(m.tgt_off.size() * 4);
        s->d_tgt_row.ensure(rowsv.size() * 4);
        ck(cudaMemcpy(s->d_rates.p, rates.data(), rates.size() * 8, cudaMemcpyHostToDevice), "upload rates");
        ck(cudaMemcpy(s->d_tgt_off.p, m.tgt_off.data(), m.tgt_off.size() * 4, cudaMemcpyHostToDevice), "upload offsets");
        ck(cudaMemcpy(s->d_tgt_row.p, rowsv.data(), rowsv.size() * 4, cudaMemcpyHostToDevice), "upload targets");
    }
    if (s->next_col + total_blocks * K >= (1ull << 47)) {
        throw std::invalid_argument("shot offset + shots must stay below 2^54");
    }
    uint64_t done = 0;
    while (done < total_blocks) {
        const uint64_t nb = std::min(max_blocks, total_blocks - done);
        DemParams p{};
        p.rates = (const uint2 *)s->d_rates.p;
        p.tgt_off = (const uint32_t *)s->d_tgt_off.p;
        p.tgt_row = (const uint32_t *)s->d_tgt_row.p;
        p.n_errors = E;
        p.rows = rows;
        p.err_row0 = record_errors ? (uint32_t)DL : 0xFFFFFFFFu;
        p.K = K;
        p.n_blocks = (uint32_t)nb;
        p.col0_base = s->next_col + done * K;
        p.seed_lo = (uint32_t)s->seed;
        p.seed_hi = (uint32_t)(s->seed >> 32);
        p.table = (uint4 *)s->d_table.p;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(nb, (uint64_t)s->num_sms * 4);
        gstim_dem_kernel<<<grid, 256, 0, s->stream>>>(p);
        ck(cudaGetLastError(), "gstim_dem_kernel launch");
        const uint64_t first = done * K * GSTIM_COL_SHOTS;
        sink(first, std::min<uint64_t>(nb * K * GSTIM_COL_SHOTS, shots - first), (const uint32_t *)s->d_table.p, (uint64_t)rows);
        done += nb;
    }
    ck(cudaStreamSynchronize(s->stream), "cudaStreamSynchronize");
    s->next_col += total_blocks * K;
}

// Transposes rows [row0, row0 + n_bits) of the table to dense b8 rows in device staging (s->d_stage), enqueued on the stream.
void dem_transpose_rows(gstim_dem_sampler *s, const uint32_t *table, uint64_t n_rows, uint64_t n, const DemOut &o) {
    const uint64_t bytes = (o.n_bits + 7) / 8;
    if (o.n_bits == 0 || n == 0) {
        return;
    }
    std::vector<uint32_t> map(o.n_bits);
    for (uint32_t b = 0; b < o.n_bits; b++) {
        map[b] = o.row0 + b;
    }
    s->d_rowmap.ensure((size_t)o.n_bits * 4);
    ck(cudaMemcpyAsync(s->d_rowmap.p, map.data(), map.size() * 4, cudaMemcpyHostToDevice, s->stream), "row map");
    s->d_stage.ensure(n * bytes + 16);
    TransposeParams t{};
    t.table = table;
    t.n_rows = n_rows;
    t.row_map = (const uint32_t *)s->d_rowmap.p;
    t.n_bits = o.n_bits;
    t.n_shots = n;
    t.out = (uint8_t *)s->d_stage.p;
    t.out_pitch = bytes;
    ck(launch_transpose_b8(t, s->stream), "transpose");
}

// ... and copies them to `host` (the file writers).
void dem_fetch(gstim_dem_sampler *s, const uint32_t *table, uint64_t n_rows, uint64_t n, const DemOut &o, std::vector<uint8_t> &host) {
    const uint64_t bytes = (o.n_bits + 7) / 8;
    host.resize(n * bytes + 1);
    if (o.n_bits == 0 || n == 0) {
        return;
    }
    dem_transpose_rows(s, table, n_rows, n, o);
    ck(cudaMemcpyAsync(host.data(), s->d_stage.p, n * bytes, cudaMemcpyDeviceToHost, s->stream), "D2H");
    ck(cudaStreamSynchronize(s->stream), "sync");
}

}  // namespace

extern "C" {

int gstim_dem_create_from_text(const char *dem_text, size_t text_len, uint64_t seed, int device, gstim_dem_sampler **out) {
    return dem_guarded([&] {
        if (out == nullptr || dem_text == nullptr) {
            throw std::invalid_argument("NULL argument.");
        }
        *out = nullptr;
        auto s = std::make_unique<gstim_dem_sampler>();
        s->model = parse_dem(std::string_view(dem_text, text_len));
        s->seed = seed;
        s->device = device;
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
            cudaGetLastError();
            throw std::runtime_error("CUDA: no usable device: this library has no CPU fallback.");
        }
        if (device < 0 || device >= n) {
            throw std::invalid_argument("CUDA device ordinal out of range.");
        }
        ck(cudaSetDevice(device), "cudaSetDevice");
        cudaDeviceProp prop;
        ck(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
        s->num_sms = prop.multiProcessorCount;
        ck(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), "cudaStreamCreate");
        *out = s.release();
    });
}

void gstim_dem_destroy(gstim_dem_sampler *s) {
    if (s) {
        cudaSetDevice(s->device);
        delete s;
    }
}

int gstim_dem_counts(const char *dem_text, size_t text_len, uint64_t *num_detectors, uint64_t *num_observables, uint64_t *num_errors) {
    return dem_guarded([&] {
        if (dem_text == nullptr) {
            throw std::invalid_argument("NULL argument.");
        }
        DemModel m = parse_dem(std::string_view(dem_text, text_len));
        if (num_detectors) {
            *num_detectors = m.num_detectors;
        }
        if (num_observables) {
            *num_observables = m.num_observables;
        }
        if (num_errors) {
            *num_errors = m.probs.size();
        }
    });
}

int gstim_dem_set_shot_offset(gstim_dem_sampler *s, uint64_t offset) {
    return dem_guarded([&] {
        if (s == nullptr || offset % GSTIM_COL_SHOTS != 0) {
            throw std::invalid_argument("shot offset must be a multiple of 128.");
        }
        s->next_col = offset / GSTIM_COL_SHOTS;
    });
}

int gstim_dem_replay(gstim_dem_sampler *s, uint64_t shots, const void *errors, int64_t errors_stride, void *dets_out, int64_t dets_stride,
                     void *obs_out, int64_t obs_stride) {
    int built = dem_guarded([&] {
        if (s == nullptr) {
            throw std::invalid_argument("NULL sampler.");
        }
        if (s->replay == nullptr) {
            const DemModel &m = s->model;
            const uint64_t D = m.num_detectors, L = m.num_observables;
            std::vector<std::vector<uint32_t>> recs(D + L);
            for (size_t e = 0; e + 1 < m.tgt_off.size(); e++) {
                for (uint32_t k = m.tgt_off[e]; k < m.tgt_off[e + 1]; k++) {
                    const uint32_t t = m.tgt[k];
                    std::vector<uint32_t> &row = recs[(t & 0x80000000u) ? D + (t & 0x7FFFFFFFu) : t];
                    if (!row.empty() && row.back() == (uint32_t)e) {
                        row.pop_back();  // a target listed twice by one mechanism cancels
                    } else {
                        row.push_back((uint32_t)e);
                    }
                }
            }
            s->replay = gstim_m2d_from_lists(s->device, m.probs.size(), D, L, std::move(recs));
        }
    });
    if (built != GSTIM_OK) {
        return built;
    }
    return gstim_m2d_convert(s->replay, shots, GSTIM_BIT_PACKED | (obs_out != nullptr ? GSTIM_SEPARATE_OBS : 0u), errors, errors_stride, nullptr, 0,
                             dets_out, dets_stride, obs_out, obs_stride);
}

int gstim_dem_sample(gstim_dem_sampler *s, uint64_t shots, uint32_t flags, void *dets_out, int64_t dets_stride, void *obs_out,
                     int64_t obs_stride, void *errs_out, int64_t errs_stride) {
    return dem_guarded([&] {
        if (s == nullptr) {
            throw std::invalid_argument("NULL sampler.");
        }
        const bool packed = (flags & GSTIM_BIT_PACKED) != 0;
        const DemModel &m = s->model;
        const uint32_t D = (uint32_t)m.num_detectors, L = (uint32_t)m.num_observables, E = (uint32_t)m.probs.size();
        DemOut outs[3] = {{(uint8_t *)dets_out, dets_stride, 0, D}, {(uint8_t *)obs_out, obs_stride, D, L}, {(uint8_t *)errs_out, errs_stride, D + L, E}};
        std::vector<uint8_t> host;
        SparseEngine *ev = errs_out == nullptr ? dem_events(s) : nullptr;
        if (ev != nullptr) {
            // rows [detectors | observables] leave the device once: by direct DMA when the caller's arrays are page-locked and
            // byte-aligned slices of the rows, else through the page-locked staging pair with the slicing / unpacking done by
            // host threads while the next sub-chunk is in flight (hostpipe.h)
            uint64_t dpitch[2];
            bool direct = packed;
            for (int k = 0; k < 2; k++) {
                const DemOut &o = outs[k];
                const uint64_t row = packed ? (o.n_bits + 7) / 8 : o.n_bits;
                dpitch[k] = o.stride ? (uint64_t)o.stride : row;
                if (o.ptr != nullptr && o.n_bits != 0) {
                    // (a slice may end inside a byte only at the end of the row, where the padding bits are zero)
                    direct = direct && (o.row0 & 7) == 0 && (((o.row0 + o.n_bits) & 7) == 0 || o.row0 + o.n_bits == D + L) && hp_is_pinned(o.ptr);
                }
            }
            if (!direct && outs[0].ptr != nullptr) {
                hp_hugepage_hint(outs[0].ptr, shots * dpitch[0]);
            }
            dem_run_events(s, *ev, shots, [&](uint64_t first, uint64_t n, const uint8_t *rows, uint64_t pitch) {
                if (direct) {
                    for (int k = 0; k < 2; k++) {
                        const DemOut &o = outs[k];
                        if (o.ptr != nullptr && o.n_bits != 0) {
                            ck(cudaMemcpy2DAsync(o.ptr + first * dpitch[k], dpitch[k], rows + (o.row0 >> 3), pitch, (o.n_bits + 7) / 8, n,
                                                 cudaMemcpyDeviceToHost, s->stream), "D2H");
                        }
                    }
                    return;
                }
                hp_staged_d2h(s->host_stage, s->stream, rows, pitch, n, [&](const uint8_t *row, uint64_t i) {
                    for (int k = 0; k < 2; k++) {
                        const DemOut &o = outs[k];
                        if (o.ptr != nullptr && o.n_bits != 0) {
                            hp_slice_row(row, o.row0, o.n_bits, packed, o.ptr + (first + i) * dpitch[k]);
                        }
                    }
                });
            });
            return;
        }
        dem_run(s, shots, errs_out != nullptr, [&](uint64_t first, uint64_t n, const uint32_t *table, uint64_t n_rows) {
            for (const DemOut &o : outs) {
                if (o.ptr == nullptr || o.n_bits == 0) {
                    continue;
                }
                // rows of this output transposed to dense b8 rows on the device, then through the page-locked staging pair:
                // host threads copy / unpack one sub-chunk while the next is in flight (error rows are 44 KB per shot for c3)
                const uint64_t bytes = (o.n_bits + 7) / 8;
                const uint64_t row = packed ? bytes : o.n_bits, pitch = o.stride ? (uint64_t)o.stride : row;
                dem_transpose_rows(s, table, n_rows, n, o);
                uint8_t *dst0 = o.ptr + first * pitch;
                if (packed && hp_is_pinned(o.ptr)) {
                    ck(cudaMemcpy2DAsync(dst0, pitch, s->d_stage.p, bytes, bytes, n, cudaMemcpyDeviceToHost, s->stream), "D2H");
                    ck(cudaStreamSynchronize(s->stream), "sync");  // (d_stage is reused by the next output)
                } else {
                    hp_staged_d2h(s->host_stage, s->stream, (const uint8_t *)s->d_stage.p, bytes, n, [&](const uint8_t *r, uint64_t i) {
                        hp_slice_row(r, 0, o.n_bits, packed, dst0 + i * pitch);
                    });
                }
            }
        });
    });
}

int gstim_dem_get_response_table(gstim_dem_sampler *s, int what, uint32_t *words, size_t *n_words) {
    return dem_guarded([&] {
        if (s == nullptr || n_words == nullptr) {
            throw std::invalid_argument("NULL argument.");
        }
        ck(cudaSetDevice(s->device), "cudaSetDevice");
        SparseEngine *ev = dem_events(s);
        if (ev == nullptr) {
            throw std::invalid_argument("This model is not sampled by the event engine.");
        }
        if (what == 8) {  // tile height
            if (words != nullptr && *n_words >= 1) {
                words[0] = ev->tile_shots();
            }
            *n_words = 1;
            return;
        }
        gstim_export_table_array(ev->table(), &ev->slices(), (uint32_t)s->model.num_detectors, what, words, n_words);
    });
}

int gstim_dem_bit_counts(gstim_dem_sampler *s, uint64_t shots, uint64_t *single_host, uint64_t *pair_host) {
    return dem_guarded([&] {
        if (s == nullptr) {
            throw std::invalid_argument("NULL sampler.");
        }
        const uint32_t n = (uint32_t)(s->model.num_detectors + s->model.num_observables);
        ck(cudaSetDevice(s->device), "cudaSetDevice");
        DevMem counts;
        counts.ensure((size_t)std::max<uint32_t>(n, 1) * 16);
        ck(cudaMemsetAsync(counts.p, 0, (size_t)std::max<uint32_t>(n, 1) * 16, s->stream), "memset");
        unsigned long long *d_single = (unsigned long long *)counts.p, *d_pair = d_single + n;
        if (SparseEngine *ev = dem_events(s)) {
            dem_run_events(s, *ev, shots, [&](uint64_t, uint64_t cnt, const uint8_t *rows, uint64_t pitch) {
                ck(launch_count_b8(rows, pitch, cnt, n, d_single, pair_host ? d_pair : nullptr, s->stream), "bit counts");
            });
        } else
        dem_run(s, shots, false, [&](uint64_t first, uint64_t cnt, const uint32_t *table, uint64_t n_rows) {
            (void)first;
            ck(launch_bit_counts(table, n_rows, cnt, nullptr, n, d_single, pair_host ? d_pair : nullptr, s->stream), "bit counts");
        });
        if (single_host && n) {
            ck(cudaMemcpy(single_host, d_single, (size_t)n * 8, cudaMemcpyDeviceToHost), "D2H");
        }
        if (pair_host && n > 1) {
            ck(cudaMemcpy(pair_host, d_pair, (size_t)(n - 1) * 8, cudaMemcpyDeviceToHost), "D2H");
        }
    });
}

int gstim_dem_sample_to_fd(gstim_dem_sampler *s, uint64_t shots, int det_fd, const char *det_format, int obs_fd, const char *obs_format,
                           int err_fd, const char *err_format) {
    return dem_guarded([&] {
        if (s == nullptr) {
            throw std::invalid_argument("NULL sampler.");
        }
        const DemModel &m = s->model;
        const uint32_t D = (uint32_t)m.num_detectors, L = (uint32_t)m.num_observables, E = (uint32_t)m.probs.size();
        struct Sink {
            int fd;
            const char *fmt;
            DemOut o;
            char prefix;
            FILE *f = nullptr;
            Format format = Format::F01;
        } sinks[3] = {{err_fd, err_format, {nullptr, 0, D + L, E}, 'M'}, {obs_fd, obs_format, {nullptr, 0, D, L}, 'L'}, {det_fd, det_format, {nullptr, 0, 0, D}, 'D'}};
        for (Sink &k : sinks) {
            if (k.fd < 0) {
                continue;
            }
            k.format = parse_format(k.fmt);
            if (k.format == Format::PTB64 && shots % 64 != 0) {
                throw std::invalid_argument("shots must be a multiple of 64 to use ptb64 format.");
            }
            int d = dup(k.fd);
            k.f = d >= 0 ? fdopen(d, "wb") : nullptr;
            if (!k.f) {
                throw std::runtime_error("fdopen() failed on an output file descriptor.");
            }
        }
        std::vector<uint8_t> host;
        std::vector<uint32_t> host_table, map;
        try {
            SparseEngine *ev = err_fd < 0 ? dem_events(s) : nullptr;
            if (ev != nullptr) {
                std::vector<uint8_t> part;
                dem_run_events(s, *ev, shots, [&](uint64_t, uint64_t n, const uint8_t *rows, uint64_t pitch) {
                    host.resize(n * pitch + 1);
                    ck(cudaMemcpyAsync(host.data(), rows, n * pitch, cudaMemcpyDeviceToHost, s->stream), "D2H");
                    ck(cudaStreamSynchronize(s->stream), "sync");
                    for (Sink &k : sinks) {
                        if (!k.f) {
                            continue;
                        }
                        const uint64_t bytes = (k.o.n_bits + 7) / 8;
                        part.assign(n * bytes + 1, 0);
                        dem_slice_rows(host.data(), pitch, n, k.o.row0, k.o.n_bits, true, part.data(), bytes);
                        if (k.format == Format::PTB64) {
                            // bit-major rows for the ptb64 writer
                            const size_t n_cols = (n + 127) / 128;
                            host_table.assign(n_cols * (size_t)std::max<uint32_t>(k.o.n_bits, 1) * 4, 0);
                            map.resize(k.o.n_bits);
                            for (uint64_t sh = 0; sh < n; sh++) {
                                for (uint32_t b = 0; b < k.o.n_bits; b++) {
                                    if ((part[sh * bytes + (b >> 3)] >> (b & 7)) & 1) {
                                        host_table[((sh >> 7) * k.o.n_bits + b) * 4 + ((sh >> 5) & 3)] |= 1u << (sh & 31);
                                    }
                                }
                            }
                            for (uint32_t b = 0; b < k.o.n_bits; b++) {
                                map[b] = b;
                            }
                            write_ptb64(k.f, host_table.data(), k.o.n_bits, map.data(), map.size(), n);
                        } else {
                            write_shots(k.f, part.data(), bytes, n, k.o.n_bits, k.format, k.prefix, k.prefix, k.o.n_bits);
                        }
                    }
                });
            } else
            dem_run(s, shots, err_fd >= 0, [&](uint64_t first, uint64_t n, const uint32_t *table, uint64_t n_rows) {
                (void)first;
                for (Sink &k : sinks) {  // (the reference writes errors, then observables, then detectors: dem_sampler.inl:98-127)
                    if (!k.f) {
                        continue;
                    }
                    if (k.format == Format::PTB64) {
                        host_table.resize((size_t)((n + 127) / 128) * n_rows * 4);
                        ck(cudaMemcpyAsync(host_table.data(), table, host_table.size() * 4, cudaMemcpyDeviceToHost, s->stream), "D2H");
                        ck(cudaStreamSynchronize(s->stream), "sync");
                        map.resize(k.o.n_bits);
                        for (uint32_t b = 0; b < k.o.n_bits; b++) {
                            map[b] = k.o.row0 + b;
                        }
                        write_ptb64(k.f, host_table.data(), n_rows, map.data(), map.size(), n);
                    } else {
                        dem_fetch(s, table, n_rows, n, k.o, host);
                        write_shots(k.f, host.data(), (k.o.n_bits + 7) / 8, n, k.o.n_bits, k.format, k.prefix, k.prefix, k.o.n_bits);
                    }
                }
            });
        } catch (...) {
            for (Sink &k : sinks) {
                if (k.f) {
                    fclose(k.f);
                }
            }
            throw;
        }
        for (Sink &k : sinks) {
            if (k.f && fclose(k.f) != 0) {
                throw std::runtime_error("Failed to flush result data.");
            }
        }
    });
}

}  // extern "C"
