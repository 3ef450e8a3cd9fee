// lowering.cc — see lowering.h.
#include "lowering.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>

namespace gstim {

uint32_t Batch::words() const {
    if (op == GOP_XORROWS) {
        return GSTIM_HDR_WORDS + (uint32_t)dst.size() + (uint32_t)off.size() + (uint32_t)idx.size();
    }
    return GSTIM_HDR_WORDS + (uint32_t)payload.size();
}

// Probability -> rate key of the 32-bit gap arithmetic of the detector-error-model sampler (dem.cu): bit 63 = valid,
// INV << 8 | SH with INV = floor(2^32 m), SH = 58 - e for 1 / lambda = m 2^e; 0 = never fires; p >= 1: INV = 0.
uint64_t gstim_rate_key(double p) {
    float f = (float)p;
    if (!(f > 0)) {
        return 0;
    }
    if (f >= 1) {
        return 1ull << 63;
    }
    const double lam = -std::log1p(-(double)f);
    int e = 0;
    const double m = std::frexp(1.0 / lam, &e);  // 1 / lambda = m * 2^e, m in [0.5, 1)
    const int sh = 58 - e;
    if (sh < 0) {
        return 0;  // p < 2^-58: not a single event in any feasible number of shots
    }
    const uint64_t inv = (uint64_t)std::floor(std::ldexp(m, 32));
    return (1ull << 63) | (inv << 8) | (uint64_t)sh;
}

namespace {

constexpr uint32_t RES_WRITE = 1u << 31;
constexpr uint32_t ITEM_X = 1u << 30;  // component flags in OBS_PAULI / FEEDBACK / CORR payload words
constexpr uint32_t ITEM_Z = 1u << 31;

// Probability -> per-shot event rate of the exponential clock in fixed point (unit 2^-56 nat, DESIGN.md
// "RNG addressing"). The reference narrows every probability to float before sampling
// (probability_util.h:47, measure_record_batch.inl:52). p >= 1 saturates: an event at every shot.
constexpr uint64_t LAM_MAX = 1ull << 62;
uint64_t rate_of(double p) {
    float f = (float)p;
    if (!(f > 0)) {
        return 0;
    }
    if (f >= 1) {
        return LAM_MAX;
    }
    double v = std::ldexp(-std::log1p(-(double)f), 56);
    if (v >= (double)LAM_MAX) {
        return LAM_MAX;
    }
    return (uint64_t)v;
}

uint32_t thr(double frac) {
    double v = std::floor(frac * 4294967296.0);
    if (!(v > 0)) {
        return 0;
    }
    if (v >= 4294967295.0) {
        return 0xFFFFFFFFu;
    }
    return (uint32_t)v;
}

struct Key {
    uint32_t op = 0, flags = 0, aux = 0, extra = 0;
    uint64_t lambda = 0;
    uint32_t t1 = 0, t2 = 0, t3 = 0;
    const uint32_t *table = nullptr;  // NOISE2 PAULI_CHANNEL_2 thresholds (15 words) or null
};

struct Lowerer {
    LoweredCircuit lc;
    uint32_t max_words;
    uint32_t Q = 0;
    uint32_t res_clock = 0, res_flag = 0, res_rec0 = 0, res_out0 = 0;
    uint32_t rec_mask = 0xFFFFFFFFu;

    // running counters (the RNG addressing contract)
    uint32_t ngroup = 0, mgroup = 0;  // noise / measure group counters
    uint64_t meas = 0;
    uint64_t det = 0;

    // current batch
    Batch cur;
    bool open = false;
    bool cur_uses_site = false, cur_uses_csite = false, cur_uses_rec = false;
    uint32_t stamp = 1;
    std::vector<uint32_t> rd_stamp, wr_stamp;

    explicit Lowerer(uint32_t max_words) : max_words(max_words) {}
    uint32_t noise_cap() const {
        const uint32_t room = max_words > 2 * GSTIM_HDR_WORDS + 15 + GSTIM_NOISE_SLICE ? max_words - 2 * GSTIM_HDR_WORDS - 15 : GSTIM_NOISE_SLICE;
        return std::min<uint32_t>(GSTIM_MAX_BATCH_ITEMS, room) / GSTIM_NOISE_SLICE * GSTIM_NOISE_SLICE;
    }

    uint32_t q_of(uint32_t target) const {
        uint32_t v = target & T_VALUE_MASK;
        return lc.qubit_map[v];
    }
    uint32_t rec_res(uint64_t m) const {
        return res_rec0 + (uint32_t)(m & rec_mask);
    }
    uint64_t rec_abs(uint32_t target, const char *gate) const {
        uint64_t k = target & T_VALUE_MASK;
        if (k == 0 || k > meas) {
            throw std::out_of_range(
                std::string("Referred to a measurement record before the beginning of time in ") + gate + ".");
        }
        return meas - k;
    }

    void flush() {
        if (!open) {
            return;
        }
        if (cur.op == GOP_XORROWS) {
            cur.n_items = (uint32_t)cur.dst.size();
        }
        lc.max_items = std::max(lc.max_items, (uint32_t)(cur.res_off.size() - 1));
        lc.total_items += cur.res_off.size() - 1;
        lc.batches.push_back(std::move(cur));
        cur = Batch();
        open = false;
        stamp++;
    }

    bool same_key(const Key &k) const {
        if (cur.op != k.op || cur.flags != k.flags || cur.aux != k.aux || cur.extra != k.extra || cur.t1 != k.t1 ||
            cur.t2 != k.t2 || cur.t3 != k.t3) {
            return false;
        }
        if (cur.lambda != k.lambda) {
            return false;
        }
        if (k.table != nullptr && memcmp(cur.payload.data(), k.table, 15 * sizeof(uint32_t)) != 0) {
            return false;
        }
        return true;
    }

    // Adds one item. `res` lists resource ids, RES_WRITE-tagged when written.
    // use_site / use_csite: the batch is tied to that noise / measure group (all items share it);
    // use_rec: items record into consecutive rows rec0 + i.
    void add(
        const Key &k,
        const uint32_t *item_words,
        uint32_t n_item_words,
        const uint32_t *res,
        uint32_t n_res,
        bool use_site,
        uint32_t site_v,
        bool use_csite,
        uint32_t csite_v,
        bool use_rec,
        uint32_t rec_v,
        bool never_merge = false) {
        bool ok = open && !never_merge && same_key(k) && cur.op != GOP_CORR;
        if (ok) {
            uint32_t n = cur.n_items;
            if ((use_site && cur.site0 != site_v) || (use_csite && cur.csite0 != csite_v) ||
                (use_rec && cur.rec0 + n != rec_v)) {
                ok = false;
            }
        }
        if (ok && (cur.words() + n_item_words + GSTIM_HDR_WORDS > max_words || cur.n_items >= GSTIM_MAX_BATCH_ITEMS)) {
            ok = false;
        }
        if (ok && (cur.op == GOP_NOISE1 || cur.op == GOP_NOISE2) && cur.n_items >= noise_cap()) {
            ok = false;  // a noise group is only ever cut at a multiple of the RNG slice size (program.h)
        }
        if (ok) {
            for (uint32_t i = 0; i < n_res; i++) {
                uint32_t r = res[i] & ~RES_WRITE;
                if (wr_stamp[r] == stamp || ((res[i] & RES_WRITE) && rd_stamp[r] == stamp)) {
                    ok = false;
                    break;
                }
            }
        }
        if (!ok) {
            flush();
            open = true;
            cur.op = k.op;
            cur.flags = k.flags;
            cur.aux = k.aux;
            cur.extra = k.extra;
            cur.lambda = k.lambda;
            cur.t1 = k.t1;
            cur.t2 = k.t2;
            cur.t3 = k.t3;
            cur.site0 = site_v;
            cur.csite0 = csite_v;
            cur.rec0 = rec_v;
            cur.res_off.push_back(0);
            if (k.table != nullptr) {
                cur.payload.assign(k.table, k.table + 15);
            }
        }
        cur.payload.insert(cur.payload.end(), item_words, item_words + n_item_words);
        for (uint32_t i = 0; i < n_res; i++) {
            uint32_t r = res[i] & ~RES_WRITE;
            if (res[i] & RES_WRITE) {
                wr_stamp[r] = stamp;
            } else {
                rd_stamp[r] = stamp;
            }
            cur.res.push_back(res[i]);
        }
        cur.res_off.push_back((uint32_t)cur.res.size());
        cur.n_items++;
    }

    // ---- detector fusion ---------------------------------------------------------------------
    // A detector that is the XOR of one result of a MEASURE batch and one earlier record row (the stabiliser-comparison
    // detectors of a memory experiment) is computed by the MEASURE item itself while the fresh result is still in
    // registers: payload of a GF_DET batch = (qubit, detector row or NONE, other record slot) per item. This removes the
    // store -> L2 -> load round trip of the result and one batch per detector layer. The XORROWS batch moves up to the
    // MEASURE batch, which is legal when nothing in between writes the record rows it reads or touches its output rows.
    void fuse_detectors() {
        if (lc.mode != 0) {
            return;
        }
        auto &B = lc.batches;
        std::vector<char> dead(B.size(), 0);
        std::vector<uint32_t> mark(lc.num_resources, 0);  // resources of the XORROWS batch under consideration
        std::vector<uint32_t> item_of(rec_mask + 1, 0), item_stamp(rec_mask + 1, 0);
        uint32_t epoch = 0;
        for (size_t x = 0; x < B.size(); x++) {
            Batch &X = B[x];
            if (X.op != GOP_XORROWS || (X.flags & GF_ACCUM) || X.dst.empty()) {
                continue;
            }
            bool pairs = true;
            for (size_t j = 0; j < X.dst.size() && pairs; j++) {
                pairs = X.off[j + 1] - X.off[j] == 2;
            }
            if (!pairs) {
                continue;
            }
            epoch++;
            for (uint32_t r : X.res) {
                mark[r & ~RES_WRITE] = epoch;
            }
            for (size_t m = x; m-- > 0;) {
                Batch &M = B[m];
                if (dead[m]) {
                    continue;
                }
                const bool candidate = M.op == GOP_MEASURE && ((M.aux >> 2) & 3u) != GK_R;
                if (candidate && try_fuse(M, X, item_of, item_stamp)) {
                    dead[x] = 1;
                    break;
                }
                // may the detector batch move above batch m?
                bool blocked = false;
                for (uint32_t r : M.res) {
                    const uint32_t id = r & ~RES_WRITE;
                    if (mark[id] == epoch && ((r & RES_WRITE) || id >= res_out0)) {
                        blocked = true;
                        break;
                    }
                }
                if (blocked) {
                    break;
                }
            }
        }
        std::vector<Batch> kept;
        kept.reserve(B.size());
        for (size_t i = 0; i < B.size(); i++) {
            if (!dead[i]) {
                kept.push_back(std::move(B[i]));
            }
        }
        B.swap(kept);
        lc.total_items = 0;
        lc.max_items = 0;
        for (const Batch &b : B) {
            lc.total_items += b.res_off.size() - 1;
            lc.max_items = std::max(lc.max_items, (uint32_t)(b.res_off.size() - 1));
        }
    }
    bool try_fuse(Batch &M, const Batch &X, std::vector<uint32_t> &item_of, std::vector<uint32_t> &item_stamp) {
        const uint32_t n = M.n_items;
        const bool fused_before = (M.flags & GF_DET) != 0;  // (a detector layer may arrive as several XORROWS batches)
        if (GSTIM_HDR_WORDS + 3 * n + GSTIM_HDR_WORDS > max_words || M.payload.size() != (fused_before ? 3 * (size_t)n : (size_t)n)) {
            return false;
        }
        static uint32_t stamp_counter = 0;
        const uint32_t st = ++stamp_counter;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t slot = (uint32_t)((M.rec0 + i) & rec_mask);
            item_of[slot] = i;
            item_stamp[slot] = st;
        }
        const uint32_t NONE = 0xFFFFFFFFu;
        std::vector<uint32_t> det(n, NONE), other(n, 0);
        if (fused_before) {
            for (uint32_t i = 0; i < n; i++) {
                det[i] = M.payload[3 * i + 1];
                other[i] = M.payload[3 * i + 2];
            }
        }
        const std::vector<uint32_t> det_before = det;
        for (size_t j = 0; j < X.dst.size(); j++) {
            const uint32_t a = X.idx[X.off[j]], b = X.idx[X.off[j] + 1];
            const bool ina = item_stamp[a] == st, inb = item_stamp[b] == st;
            if (ina == inb) {
                return false;  // none or both of the two rows come from this batch
            }
            const uint32_t i = item_of[ina ? a : b];
            if (det[i] != NONE) {
                return false;  // a result that feeds two detectors stays with XORROWS
            }
            det[i] = X.dst[j];
            other[i] = ina ? b : a;
        }
        std::vector<uint32_t> payload, res, res_off{0};
        for (uint32_t i = 0; i < n; i++) {
            payload.push_back(M.payload[fused_before ? 3 * i : i]);
            payload.push_back(det[i]);
            payload.push_back(other[i]);
            res.insert(res.end(), M.res.begin() + M.res_off[i], M.res.begin() + M.res_off[i + 1]);
            if (det[i] != NONE && det_before[i] == NONE) {
                res.push_back((res_out0 + det[i]) | RES_WRITE);
                res.push_back(res_rec0 + other[i]);
            }
            res_off.push_back((uint32_t)res.size());
        }
        M.payload.swap(payload);
        M.res.swap(res);
        M.res_off.swap(res_off);
        M.flags |= GF_DET;
        return true;
    }

    // ---- op emitters -------------------------------------------------------------------------
    void cliff1(uint32_t mat, uint32_t q) {
        Key k;
        k.op = GOP_CLIFF1;
        k.aux = mat;
        uint32_t r = q | RES_WRITE;
        add(k, &q, 1, &r, 1, false, 0, false, 0, false, 0);
    }
    void cliff2(uint32_t mat, uint32_t q1, uint32_t q2) {
        Key k;
        k.op = GOP_CLIFF2;
        k.aux = mat;
        uint32_t w = q1 | (q2 << 16);
        uint32_t r[2] = {q1 | RES_WRITE, q2 | RES_WRITE};
        add(k, &w, 1, r, 2, false, 0, false, 0, false, 0);
    }
    void feedback(uint64_t rec_index, uint32_t q, uint32_t comps) {
        Key k;
        k.op = GOP_FEEDBACK;
        uint32_t w[2] = {(uint32_t)(rec_index & rec_mask), q | comps};
        uint32_t r[2] = {rec_res(rec_index), q | RES_WRITE};
        add(k, w, 2, r, 2, false, 0, false, 0, false, 0);
    }
    // Splits a target list into maximal runs without a repeated qubit (each run = one RNG group).
    std::vector<uint32_t> run_stamp;
    uint32_t run_epoch = 0;
    template <typename KEYS>
    std::vector<std::pair<size_t, size_t>> runs_of(size_t n, KEYS &&keys_of) {
        std::vector<std::pair<size_t, size_t>> runs;
        run_stamp.resize(Q + 1, 0);
        run_epoch++;
        size_t start = 0;
        for (size_t i = 0; i < n; i++) {
            uint32_t ks[2];
            int nk = keys_of(i, ks);
            bool rep = false;
            for (int j = 0; j < nk; j++) {
                rep |= run_stamp[ks[j]] == run_epoch;
            }
            if (rep) {
                runs.push_back({start, i});
                start = i;
                run_epoch++;
            }
            for (int j = 0; j < nk; j++) {
                run_stamp[ks[j]] = run_epoch;
            }
        }
        if (start < n) {
            runs.push_back({start, n});
        }
        return runs;
    }

    // basis/kind measurement or reset of one qubit inside measure group `mg`.
    void measure(uint32_t basis, uint32_t kind, uint32_t q, uint32_t mg) {
        Key k;
        k.op = GOP_MEASURE;
        k.aux = basis | (kind << 2);
        bool records = kind != GK_R;
        uint32_t r[2] = {q | RES_WRITE, 0};
        uint32_t nr = 1;
        if (records) {
            r[1] = rec_res(meas) | RES_WRITE;
            nr = 2;
        }
        add(k, &q, 1, r, nr, false, 0, true, mg, records, (uint32_t)meas);
        if (records) {
            meas++;
        }
    }
    // One measurement-type instruction on compact qubits qs (may repeat): collapse groups + optional result noise.
    void measure_list(uint32_t basis, uint32_t kind, const std::vector<uint32_t> &qs, const std::vector<double> &args) {
        uint64_t rec_first = meas;
        for (auto run : runs_of(qs.size(), [&](size_t i, uint32_t *ks) { ks[0] = qs[i]; return 1; })) {
            uint32_t mg = mgroup++;
            for (size_t i = run.first; i < run.second; i++) {
                measure(basis, kind, qs[i], mg);
            }
        }
        if (kind != GK_R && !args.empty()) {
            rec_noise(args[0], qs, rec_first);
        }
    }
    // Result-flip noise on record rows rec_first + i with clock qubits clock_qubits[i].
    void rec_noise(double p, const std::vector<uint32_t> &clock_qubits, uint64_t rec_first) {
        uint64_t lam = rate_of(p);
        Key k;
        k.op = GOP_NOISE1;
        k.flags = GF_REC;
        k.lambda = lam;
        for (auto run : runs_of(clock_qubits.size(), [&](size_t i, uint32_t *ks) { ks[0] = clock_qubits[i]; return 1; })) {
            uint32_t g = ngroup++;
            if (lam == 0) {
                continue;
            }
            for (size_t i = run.first; i < run.second; i++) {
                uint32_t q = clock_qubits[i];
                uint32_t r[2] = {q | RES_WRITE, rec_res(rec_first + i) | RES_WRITE};
                add(k, &q, 1, r, 2, true, g, false, 0, true, (uint32_t)(rec_first + i));
            }
        }
    }
    // Single-target Pauli-choice noise over qs (may repeat). rec_first >= 0: every event also flips rec row.
    void noise1_list(uint64_t lam, uint32_t cats, uint32_t t1, uint32_t t2, uint32_t t3, const std::vector<uint32_t> &qs,
                     bool herald, uint64_t rec_first) {
        Key k;
        k.op = GOP_NOISE1;
        k.flags = herald ? GF_REC : 0;
        k.aux = cats;
        k.lambda = lam;
        k.t1 = t1;
        k.t2 = t2;
        k.t3 = t3;
        for (auto run : runs_of(qs.size(), [&](size_t i, uint32_t *ks) { ks[0] = qs[i]; return 1; })) {
            uint32_t g = ngroup++;
            if (lam == 0) {
                continue;
            }
            for (size_t i = run.first; i < run.second; i++) {
                uint32_t q = qs[i];
                uint32_t r[2] = {q | RES_WRITE, 0};
                uint32_t nr = 1;
                if (herald) {
                    r[1] = rec_res(rec_first + i) | RES_WRITE;
                    nr = 2;
                }
                add(k, &q, 1, r, nr, true, g, false, 0, herald, (uint32_t)(rec_first + i));
            }
        }
    }
    void noise2_list(uint64_t lam, const std::vector<uint32_t> &flat_pairs, const uint32_t *table, uint32_t last) {
        Key k;
        k.op = GOP_NOISE2;
        k.flags = table ? GF_TABLE : 0;
        k.aux = table ? last : 0;
        k.lambda = lam;
        k.table = table;
        size_t n = flat_pairs.size() / 2;
        for (auto run : runs_of(n, [&](size_t i, uint32_t *ks) { ks[0] = flat_pairs[2 * i]; ks[1] = flat_pairs[2 * i + 1]; return 2; })) {
            uint32_t g = ngroup++;
            if (lam == 0) {
                continue;
            }
            for (size_t i = run.first; i < run.second; i++) {
                uint32_t a = flat_pairs[2 * i], b = flat_pairs[2 * i + 1];
                uint32_t w = a | (b << 16);
                uint32_t r[2] = {a | RES_WRITE, b | RES_WRITE};
                add(k, &w, 1, r, 2, true, g, false, 0, false, 0);
            }
        }
    }
    bool with_sweep = false;
    void sweep(uint32_t sweep_index, uint32_t q, uint32_t comps) {
        Key k;
        k.op = GOP_SWEEP;
        uint32_t w[2] = {sweep_index, q | comps};
        uint32_t r = q | RES_WRITE;
        add(k, w, 2, &r, 1, false, 0, false, 0, false, 0);
    }
    void rec_zero(uint64_t rec_index) {
        Key k;
        k.op = GOP_RECZERO;
        uint32_t r = rec_res(rec_index) | RES_WRITE;
        add(k, nullptr, 0, &r, 1, false, 0, false, 0, true, (uint32_t)rec_index);
    }
    void xor_rows(uint32_t dst_row, const std::vector<uint64_t> &recs, bool accum) {
        Key k;
        k.op = GOP_XORROWS;
        k.flags = accum ? GF_ACCUM : 0;
        std::vector<uint32_t> r;
        r.push_back((res_out0 + dst_row) | RES_WRITE);
        for (uint64_t m : recs) {
            r.push_back(rec_res(m));
        }
        // XORROWS payload is kept split (dst/off/idx); make room before add() decides on merging.
        uint32_t need = 2 + (uint32_t)recs.size();
        if (open && cur.op == GOP_XORROWS && cur.words() + need + GSTIM_HDR_WORDS > max_words) {
            flush();
        }
        add(k, nullptr, 0, r.data(), (uint32_t)r.size(), false, 0, false, 0, false, 0);
        if (cur.off.empty()) {
            cur.off.push_back(0);
        }
        cur.dst.push_back(dst_row);
        for (uint64_t m : recs) {
            cur.idx.push_back((uint32_t)(m & rec_mask));
        }
        cur.off.push_back((uint32_t)cur.idx.size());
    }
    void obs_pauli(uint32_t dst_row, uint32_t q, uint32_t comps) {
        Key k;
        k.op = GOP_OBS_PAULI;
        uint32_t w[2] = {dst_row, q | comps};
        uint32_t r[2] = {(res_out0 + dst_row) | RES_WRITE, q};
        add(k, w, 2, r, 2, false, 0, false, 0, false, 0);
    }

    // ---- gate lowering -----------------------------------------------------------------------
    void do_cliff1_list(uint32_t mat, const std::vector<uint32_t> &qs) {
        for (uint32_t q : qs) {
            cliff1(mat, q);
        }
    }
    // pairs given as compact qubits or bit targets (T_REC/T_SWEEP words kept raw)
    void do_controlled(const GateInfo &g, uint32_t a, uint32_t b) {
        bool a_bit = (a & (T_REC | T_SWEEP)) != 0, b_bit = (b & (T_REC | T_SWEEP)) != 0;
        if (!a_bit && !b_bit) {
            cliff2(g.param, q_of(a), q_of(b));
            return;
        }
        std::string name = g.name;
        // Which side is the classical control, and which Pauli does it apply?
        uint32_t comps = 0;
        uint32_t bit = 0, qt = 0;
        if (name == "CX" || name == "CY") {
            if (b_bit) {
                throw std::invalid_argument(
                    std::string("Controlled ") + (name == "CX" ? "X" : "Y") + " had a bit as its target, instead of its control.");
            }
            bit = a;
            qt = b;
            comps = name == "CX" ? ITEM_X : (ITEM_X | ITEM_Z);
        } else if (name == "XCZ" || name == "YCZ") {
            if (a_bit) {
                throw std::invalid_argument(
                    std::string("Controlled ") + (name == "XCZ" ? "X" : "Y") + " had a bit as its target, instead of its control.");
            }
            bit = b;
            qt = a;
            comps = name == "XCZ" ? ITEM_X : (ITEM_X | ITEM_Z);
        } else {  // CZ: either side may be the bit; both bits -> no effect
            if (a_bit && b_bit) {
                return;
            }
            bit = a_bit ? a : b;
            qt = a_bit ? b : a;
            comps = ITEM_Z;
        }
        if (bit & T_SWEEP) {
            if (with_sweep) {
                sweep(bit & T_VALUE_MASK, q_of(qt), comps);
            }
            return;  // no sweep data when sampling (frame_simulator.inl:146-148)
        }
        feedback(rec_abs(bit, g.name), q_of(qt), comps);
    }

    void do_measure_gate(const Instruction &op) {
        uint32_t basis = op.gate->param & 3, kind = op.gate->param >> 2;
        std::vector<uint32_t> qs;
        for (uint32_t t : op.targets) {
            qs.push_back(q_of(t));  // '!' is ignored: it lives in the reference sample (frame_simulator.inl:176)
        }
        measure_list(basis, kind, qs, op.args);
    }

    void do_mpad(const std::vector<double> &args, size_t n) {
        uint64_t rec_first = meas;
        for (size_t i = 0; i < n; i++) {
            rec_zero(meas);
            meas++;
        }
        if (!args.empty()) {
            uint64_t lam = rate_of(args[0]);
            for (size_t i = 0; i < n; i++) {
                uint32_t g = ngroup++;  // all MPAD results share the global clock: one group each
                if (lam == 0) {
                    continue;
                }
                Key k;
                k.op = GOP_NOISE1;
                k.flags = GF_REC | GF_NOFRAME;
                k.lambda = lam;
                k.extra = Q + 1;  // clock = global clock (index Q)
                uint32_t r[2] = {res_clock | RES_WRITE, rec_res(rec_first + i) | RES_WRITE};
                uint32_t dummy = Q;
                add(k, &dummy, 1, r, 2, true, g, false, 0, true, (uint32_t)(rec_first + i));
            }
        }
    }

    struct Product {
        std::vector<std::pair<uint32_t, uint32_t>> terms;  // (compact qubit, xz bits: 1=x 2=z), sorted by ORIGINAL index
        std::vector<uint32_t> bits;                        // classical bit targets
    };
    // Splits a combiner-joined target list into products; multiplies same-qubit terms.
    std::vector<Product> read_products(const Instruction &op, bool allow_bits) {
        std::vector<Product> out;
        size_t k = 0;
        const auto &ts = op.targets;
        while (k < ts.size()) {
            size_t end = k + 1;
            while (end < ts.size() && ts[end] == T_COMBINER) {
                end += 2;
            }
            std::vector<std::pair<uint32_t, uint32_t>> acc;  // original qubit -> xz
            bool imag = false;
            Product p;
            for (size_t j = k; j < end; j += 2) {
                uint32_t t = ts[j];
                if (t & (T_REC | T_SWEEP)) {
                    if (!allow_bits) {
                        throw std::invalid_argument(std::string("Found an unsupported target in ") + op.gate->name + ".");
                    }
                    p.bits.push_back(t);
                    continue;
                }
                uint32_t q = t & T_VALUE_MASK;
                uint32_t xz = ((t & T_PAULI_X) ? 1u : 0u) | ((t & T_PAULI_Z) ? 2u : 0u);
                bool found = false;
                for (auto &e : acc) {
                    if (e.first == q) {
                        if (e.second != 0 && xz != 0 && e.second != xz) {
                            imag = !imag;
                        }
                        e.second ^= xz;
                        found = true;
                    }
                }
                if (!found) {
                    acc.push_back({q, xz});
                }
            }
            if (imag) {
                throw std::invalid_argument(
                    std::string("Acted on an anti-Hermitian operator (e.g. X0*Z0 instead of Y0) in ") + op.gate->name + ".");
            }
            std::sort(acc.begin(), acc.end());
            for (auto &e : acc) {
                if (e.second != 0) {
                    p.terms.push_back({lc.qubit_map[e.first], e.second});
                }
            }
            out.push_back(std::move(p));
            k = end;
        }
        return out;
    }

    // gate_decomposition.cc:88-161: conjugate each product to a Z on its first qubit, measure, undo.
    void do_mpp(const Instruction &op) {
        std::vector<uint32_t> h_xz, h_yz, cx_pairs, ms;
        std::vector<uint8_t> merged(Q, 0);
        const GateInfo *CX = find_gate("CX");
        // The CXs of one product share their target and commute; those of different products are disjoint. Emitting the
        // k-th CX of every product before any (k + 1)-th one lets them share batches (one batch per layer instead of
        // one per CX).
        auto cx_layers = [&]() {
            std::vector<size_t> next_of;  // per product (= per run of equal targets in cx_pairs): index of its next pair
            for (size_t i = 0; i < cx_pairs.size(); i += 2) {
                if (i == 0 || cx_pairs[i + 1] != cx_pairs[i - 1]) {
                    next_of.push_back(i);
                }
            }
            std::vector<size_t> end_of(next_of.begin() + (next_of.empty() ? 0 : 1), next_of.end());
            end_of.push_back(cx_pairs.size());
            for (bool any = true; any;) {
                any = false;
                for (size_t p = 0; p < next_of.size(); p++) {
                    if (next_of[p] < end_of[p]) {
                        cliff2(CX->param, cx_pairs[next_of[p]], cx_pairs[next_of[p] + 1]);
                        next_of[p] += 2;
                        any = true;
                    }
                }
            }
        };
        auto flush_group = [&]() {
            if (ms.empty()) {
                return;
            }
            do_cliff1_list(0x6, h_xz);
            do_cliff1_list(0xB, h_yz);
            cx_layers();
            measure_list(GB_Z, GK_M, ms, op.args);
            cx_layers();
            do_cliff1_list(0xB, h_yz);
            do_cliff1_list(0x6, h_xz);
            h_xz.clear();
            h_yz.clear();
            cx_pairs.clear();
            ms.clear();
            std::fill(merged.begin(), merged.end(), 0);
        };
        for (const Product &p : read_products(op, false)) {
            if (p.terms.empty()) {
                flush_group();
                do_mpad(op.args, 1);
                continue;
            }
            bool overlap = false;
            for (auto &e : p.terms) {
                overlap |= merged[e.first] != 0;
            }
            if (overlap) {
                flush_group();
            }
            bool first = true;
            for (auto &e : p.terms) {
                merged[e.first] = 1;
                if (e.second & 1) {
                    ((e.second & 2) ? h_yz : h_xz).push_back(e.first);
                }
                if (first) {
                    ms.push_back(e.first);
                    first = false;
                } else {
                    cx_pairs.push_back(e.first);
                    cx_pairs.push_back(ms.back());
                }
            }
        }
        flush_group();
    }

    // gate_decomposition.cc:163-243. SPP and SPP_DAG act identically on the frame.
    void do_spp(const Instruction &op) {
        const GateInfo *CX = find_gate("CX");
        for (const Product &p : read_products(op, true)) {
            if (p.terms.empty()) {
                continue;
            }
            std::vector<uint32_t> h_xz, h_yz;
            uint32_t focus = p.terms[0].first;
            for (auto &e : p.terms) {
                if (e.second & 1) {
                    ((e.second & 2) ? h_yz : h_xz).push_back(e.first);
                }
            }
            auto cx_layer = [&]() {
                for (size_t i = 1; i < p.terms.size(); i++) {
                    cliff2(CX->param, p.terms[i].first, focus);
                }
                for (uint32_t b : p.bits) {
                    if (b & T_SWEEP) {
                        if (with_sweep) {
                            sweep(b & T_VALUE_MASK, focus, ITEM_X);
                        }
                        continue;
                    }
                    feedback(rec_abs(b, op.gate->name), focus, ITEM_X);
                }
            };
            do_cliff1_list(0x6, h_xz);
            do_cliff1_list(0xB, h_yz);
            cx_layer();
            cliff1(0xD, focus);  // S / S_DAG : z ^= x
            cx_layer();
            do_cliff1_list(0xB, h_yz);
            do_cliff1_list(0x6, h_xz);
        }
    }

    // frame_simulator.inl:842-902 + gate_decomposition.cc:245-274.
    void do_mpair(const Instruction &op) {
        uint32_t basis = op.gate->param;
        const GateInfo *conj = find_gate(basis == GB_X ? "CX" : basis == GB_Y ? "CY" : "XCZ");
        std::vector<uint8_t> used(Q, 0);
        std::vector<uint32_t> seg;
        auto flush_seg = [&]() {
            if (seg.empty()) {
                return;
            }
            for (size_t i = 0; i < seg.size(); i += 2) {
                cliff2(conj->param, seg[i], seg[i + 1]);
            }
            std::vector<uint32_t> ms;
            for (size_t i = 0; i < seg.size(); i += 2) {
                ms.push_back(seg[i]);
            }
            measure_list(basis, GK_M, ms, op.args);
            for (size_t i = 0; i < seg.size(); i += 2) {
                cliff2(conj->param, seg[i], seg[i + 1]);
            }
            seg.clear();
            std::fill(used.begin(), used.end(), 0);
        };
        for (size_t i = 0; i < op.targets.size(); i += 2) {
            uint32_t a = q_of(op.targets[i]), b = q_of(op.targets[i + 1]);
            if (used[a] || used[b]) {
                flush_seg();
            }
            used[a] = used[b] = 1;
            seg.push_back(a);
            seg.push_back(b);
        }
        flush_seg();
    }

    std::vector<uint32_t> compact_targets(const Instruction &op) {
        std::vector<uint32_t> qs;
        for (uint32_t t : op.targets) {
            qs.push_back(q_of(t));
        }
        return qs;
    }

    void do_noise1_gate(const Instruction &op) {
        uint64_t lam = rate_of(op.args[0]);
        uint32_t cats, t1 = 0, t2 = 0, t3 = 0;
        switch (op.gate->param) {
            case 1:
                cats = 0x55;  // X X X X
                break;
            case 2:
                cats = 0xAA;  // Z
                break;
            case 3:
                cats = 0xFF;  // Y
                break;
            default:  // DEPOLARIZE1: uniform over X, Z, Y (frame_simulator.inl:636-641: 1->X, 2->Z, 3->Y)
                cats = 1u | (2u << 2) | (3u << 4) | (3u << 6);
                t1 = thr(1.0 / 3.0);
                t2 = thr(2.0 / 3.0);
                t3 = t2;
                break;
        }
        noise1_list(lam, cats, t1, t2, t3, compact_targets(op), false, 0);
    }

    void do_depolarize2(const Instruction &op) {
        noise2_list(rate_of(op.args[0]), compact_targets(op), nullptr, 0);
    }

    // One site with the channel's total probability, then a category draw: same joint distribution
    // as the ELSE_CORRELATED_ERROR chain of tableau_simulator.h:291-324.
    void do_pauli_channel_1(const Instruction &op) {
        double px = op.args[0], py = op.args[1], pz = op.args[2];
        double tot = px + py + pz;
        uint64_t lam = rate_of(std::min(tot, 1.0));
        uint32_t t1 = 0, t2 = 0, cats = 0;
        if (tot > 0) {
            t1 = thr(px / tot);
            t2 = thr((px + py) / tot);
            uint32_t last = pz > 0 ? 2u : py > 0 ? 3u : 1u;
            cats = 1u | (3u << 2) | (last << 4) | (last << 6);
        }
        noise1_list(lam, cats, t1, t2, t2, compact_targets(op), false, 0);
    }

    void do_pauli_channel_2(const Instruction &op) {
        double tot = 0;
        for (double p : op.args) {
            tot += p;
        }
        uint64_t lam = rate_of(std::min(tot, 1.0));
        uint32_t table[15];
        uint32_t last = 1;
        double cum = 0;
        for (int i = 0; i < 15; i++) {
            cum += op.args[i];
            table[i] = tot > 0 ? thr(cum / tot) : 0;
            if (op.args[i] > 0) {
                last = (uint32_t)i + 1;
            }
        }
        noise2_list(lam, compact_targets(op), table, last);
    }

    void do_corr(const Instruction &op) {
        Key k;
        k.op = GOP_CORR;
        k.flags = op.gate->param ? GF_RESET_FLAG : 0;
        k.lambda = rate_of(op.args[0]);
        std::vector<uint32_t> words, res;
        res.push_back(res_flag | RES_WRITE);
        for (uint32_t t : op.targets) {
            uint32_t q = q_of(t);
            words.push_back(q | ((t & T_PAULI_X) ? ITEM_X : 0) | ((t & T_PAULI_Z) ? ITEM_Z : 0));
            res.push_back(q | RES_WRITE);
        }
        uint32_t clock = words.empty() ? Q : (words[0] & 0xFFFFFF);
        if (words.empty()) {
            res.push_back(res_clock | RES_WRITE);
        }
        k.extra = clock;
        uint32_t g = ngroup++;
        if (k.lambda != 0 || (k.flags & GF_RESET_FLAG)) {
            flush();
            // dedupe resources (a qubit may appear twice in the Pauli list)
            std::sort(res.begin(), res.end());
            res.erase(std::unique(res.begin(), res.end()), res.end());
            add(k, words.data(), (uint32_t)words.size(), res.data(), (uint32_t)res.size(), true, g, false, 0, false, 0, true);
            cur.n_items = (uint32_t)words.size();
            flush();
        }
    }

    void do_heralded(const Instruction &op) {
        double tot;
        uint32_t t1, t2, t3;
        uint32_t cats = 1u | (2u << 2) | (3u << 4) | (0u << 6);  // X, Z, Y, I
        if (op.gate->cat == GateCat::HERALDED_ERASE) {
            tot = op.args[0];
            t1 = 1u << 30;
            t2 = 2u << 30;
            t3 = 3u << 30;
        } else {
            double hi = op.args[0], hx = op.args[1], hy = op.args[2], hz = op.args[3];
            tot = hi + hx + hy + hz;
            t1 = tot > 0 ? thr(hx / tot) : 0;
            t2 = tot > 0 ? thr((hx + hz) / tot) : 0;
            t3 = tot > 0 ? thr((hx + hz + hy) / tot) : 0;
        }
        uint64_t lam = rate_of(std::min(tot, 1.0));
        uint64_t rec_first = meas;
        for (size_t i = 0; i < op.targets.size(); i++) {
            rec_zero(meas);
            meas++;
        }
        noise1_list(lam, cats, t1, t2, t3, compact_targets(op), true, rec_first);
    }

    void do_detector(const Instruction &op) {
        if (lc.mode == 0) {
            std::vector<uint64_t> recs;
            for (uint32_t t : op.targets) {
                recs.push_back(rec_abs(t, "DETECTOR"));
            }
            if (with_sweep) {
                lc.out_recs[det] = recs;
            }
            xor_rows((uint32_t)det, recs, false);
        } else {
            for (uint32_t t : op.targets) {
                rec_abs(t, "DETECTOR");
            }
        }
        det++;
    }

    void do_observable(const Instruction &op) {
        uint32_t row = (uint32_t)lc.stats.num_detectors + (uint32_t)op.args[0];
        std::vector<uint64_t> recs;
        for (uint32_t t : op.targets) {
            if (t & T_REC) {
                recs.push_back(rec_abs(t, "OBSERVABLE_INCLUDE"));
            }
        }
        if (lc.mode != 0) {
            return;
        }
        if (with_sweep) {
            lc.out_recs[row].insert(lc.out_recs[row].end(), recs.begin(), recs.end());
        }
        if (!recs.empty()) {
            xor_rows(row, recs, true);
        }
        for (uint32_t t : op.targets) {
            if (!(t & T_REC)) {
                // X target reads the z component, Z target reads x, Y both (frame_simulator.inl:240-246)
                uint32_t comps = ((t & T_PAULI_X) ? ITEM_Z : 0) | ((t & T_PAULI_Z) ? ITEM_X : 0);
                obs_pauli(row, q_of(t), comps);
            }
        }
    }

    void do_op(const Instruction &op) {
        const GateInfo &g = *op.gate;
        switch (g.cat) {
            case GateCat::NOOP:
                break;
            case GateCat::CLIFF1:
                for (uint32_t t : op.targets) {
                    cliff1(g.param, q_of(t));
                }
                break;
            case GateCat::CLIFF2:
                for (size_t i = 0; i < op.targets.size(); i += 2) {
                    do_controlled(g, op.targets[i], op.targets[i + 1]);
                }
                break;
            case GateCat::MEASURE:
                do_measure_gate(op);
                break;
            case GateCat::MPAD:
                do_mpad(op.args, op.targets.size());
                break;
            case GateCat::MPP:
                do_mpp(op);
                break;
            case GateCat::SPP:
                do_spp(op);
                break;
            case GateCat::MPAIR:
                do_mpair(op);
                break;
            case GateCat::NOISE1:
                do_noise1_gate(op);
                break;
            case GateCat::DEPOLARIZE2:
                do_depolarize2(op);
                break;
            case GateCat::PAULI_CHANNEL_1:
                do_pauli_channel_1(op);
                break;
            case GateCat::PAULI_CHANNEL_2:
                do_pauli_channel_2(op);
                break;
            case GateCat::CORR:
                do_corr(op);
                break;
            case GateCat::HERALDED_ERASE:
            case GateCat::HERALDED_PAULI_CHANNEL_1:
                do_heralded(op);
                break;
            case GateCat::DETECTOR:
                do_detector(op);
                break;
            case GateCat::OBSERVABLE_INCLUDE:
                do_observable(op);
                break;
            case GateCat::REPEAT:
                break;
        }
    }
};

void mark_used(const Circuit &c, std::vector<uint8_t> &used) {
    for (const auto &op : c.ops) {
        if (op.gate->cat == GateCat::REPEAT) {
            mark_used(c.blocks[op.block_index], used);
            continue;
        }
        if (op.gate->cat == GateCat::MPAD || op.gate->targets == TR_NONE) {
            continue;
        }
        if (std::string(op.gate->name) == "QUBIT_COORDS") {
            continue;
        }
        for (uint32_t t : op.targets) {
            if (t == T_COMBINER || (t & (T_REC | T_SWEEP))) {
                continue;
            }
            used[t & T_VALUE_MASK] = 1;
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Layout post-pass (performance only; semantics and RNG addressing are unaffected).
//   1. physical frame rows: qubits are renumbered so that, inside every two-qubit batch, both operand
//      sets spread evenly over the 8 sixteen-byte bank groups of shared memory (row index mod 8);
//      lc.logical_of[] keeps the logical (sorted) index that addresses the Philox counters.
//   2. items of every gate batch are reordered into groups of 8 (one quarter warp) whose rows
//      are distinct mod 8 for each operand -> conflict-free LDS.128/STS.128. Groups are perfect matchings
//      of the 8x8 (residue of operand 1, residue of operand 2) multigraph.
// ---------------------------------------------------------------------------------------------
namespace {

bool is_pair_op(const Batch &b) {
    return b.op == GOP_CLIFF2 || b.op == GOP_NOISE2;
}
size_t item_skip(const Batch &b) {
    return (b.op == GOP_NOISE2 && (b.flags & GF_TABLE)) ? 15 : 0;
}

void assign_physical_rows(LoweredCircuit &lc) {
    const uint32_t Q = lc.num_qubits;
    std::vector<uint32_t> phys(Q, UINT32_MAX);
    uint32_t next = 0;
    for (const Batch &b : lc.batches) {
        if (!is_pair_op(b)) {
            continue;
        }
        size_t skip = item_skip(b);
        for (int role = 0; role < 2; role++) {
            for (size_t i = skip; i < b.payload.size(); i++) {
                uint32_t q = role == 0 ? (b.payload[i] & 0xFFFF) : (b.payload[i] >> 16);
                if (phys[q] == UINT32_MAX) {
                    phys[q] = next++;
                }
            }
        }
    }
    for (uint32_t q = 0; q < Q; q++) {
        if (phys[q] == UINT32_MAX) {
            phys[q] = next++;
        }
    }
    lc.logical_of.assign(Q + 1, Q);
    for (uint32_t q = 0; q < Q; q++) {
        lc.logical_of[phys[q]] = q;
    }
    auto lo24 = [&](uint32_t w) { return (w & 0xFF000000u) | phys[w & 0xFFFFFFu]; };
    for (Batch &b : lc.batches) {
        size_t skip = item_skip(b);
        switch (b.op) {
            case GOP_CLIFF1:
                for (auto &w : b.payload) {
                    w = phys[w];
                }
                break;
            case GOP_MEASURE: {  // physical row | logical index << 16 (the logical index addresses the collapse draws)
                const size_t stride = (b.flags & GF_DET) ? 3 : 1;
                for (size_t i = 0; i < b.payload.size(); i += stride) {
                    b.payload[i] = phys[b.payload[i]] | (b.payload[i] << 16);
                }
                break;
            }
            case GOP_NOISE1:
                if (!(b.flags & GF_NOFRAME)) {
                    for (auto &w : b.payload) {
                        w = phys[w];
                    }
                }
                break;
            case GOP_CLIFF2:
            case GOP_NOISE2:
                for (size_t i = skip; i < b.payload.size(); i++) {
                    b.payload[i] = phys[b.payload[i] & 0xFFFF] | (phys[b.payload[i] >> 16] << 16);
                }
                break;
            case GOP_OBS_PAULI:
            case GOP_FEEDBACK:
            case GOP_SWEEP:
                for (size_t i = 1; i < b.payload.size(); i += 2) {
                    b.payload[i] = lo24(b.payload[i]);
                }
                break;
            case GOP_CORR:
                for (auto &w : b.payload) {
                    w = lo24(w);
                }
                if (b.extra < Q) {
                    b.extra = phys[b.extra];
                }
                break;
            default:
                break;
        }
    }
}

// Kuhn's augmenting-path matching on the 8x8 residue multigraph.
bool try_augment(int u, const uint32_t cnt[8][8], int match_v[8], bool seen[8]) {
    for (int v = 0; v < 8; v++) {
        if (cnt[u][v] == 0 || seen[v]) {
            continue;
        }
        seen[v] = true;
        if (match_v[v] < 0 || try_augment(match_v[v], cnt, match_v, seen)) {
            match_v[v] = u;
            return true;
        }
    }
    return false;
}

void spread_banks(Batch &b) {
    // gate batches only: the item order of a noise batch is the site order its RNG slices are defined on
    const bool pairs = b.op == GOP_CLIFF2;
    if (!pairs && b.op != GOP_CLIFF1) {
        return;
    }
    const size_t skip = item_skip(b);
    const size_t n = b.payload.size() - skip;
    if (n < 9 || n != b.res_off.size() - 1) {
        return;
    }
    const uint32_t *items = b.payload.data() + skip;
    // cell[k1][k2] = stack of item indices (earliest on top)
    std::vector<uint32_t> cell[8][8];
    uint32_t cnt[8][8] = {};
    for (size_t i = n; i-- > 0;) {
        uint32_t k1 = items[i] & 7, k2 = pairs ? ((items[i] >> 16) & 7) : (uint32_t)(0);
        if (!pairs) {
            k2 = k1;  // singles: diagonal cells, a "matching" is just one item per residue
        }
        cell[k1][k2].push_back((uint32_t)i);
        cnt[k1][k2]++;
    }
    std::vector<uint32_t> order;
    order.reserve(n);
    while (order.size() < n) {
        int match_v[8];
        std::fill(match_v, match_v + 8, -1);
        // visit operand-1 residues with the most remaining items first, so heavy rows are never starved
        int us[8];
        uint32_t deg[8];
        for (int u = 0; u < 8; u++) {
            us[u] = u;
            deg[u] = 0;
            for (int v = 0; v < 8; v++) {
                deg[u] += cnt[u][v];
            }
        }
        std::sort(us, us + 8, [&](int a, int c) { return deg[a] > deg[c]; });
        for (int t = 0; t < 8; t++) {
            if (deg[us[t]] == 0) {
                continue;
            }
            bool seen[8] = {};
            try_augment(us[t], cnt, match_v, seen);
        }
        size_t before = order.size();
        bool used_u[8] = {};
        for (int v = 0; v < 8; v++) {
            if (match_v[v] >= 0) {
                int u = match_v[v];
                order.push_back(cell[u][v].back());
                cell[u][v].pop_back();
                cnt[u][v]--;
                used_u[u] = true;
            }
        }
        // fill the rest of the group (unavoidable conflicts) with items from the fullest rows
        while (order.size() - before < 8 && order.size() < n) {
            int bu = -1, bv = -1;
            uint32_t best = 0;
            for (int u = 0; u < 8; u++) {
                for (int v = 0; v < 8; v++) {
                    uint32_t score = cnt[u][v] ? cnt[u][v] + (used_u[u] ? 0 : 1000000u) : 0;
                    if (score > best) {
                        best = score;
                        bu = u;
                        bv = v;
                    }
                }
            }
            if (bu < 0) {
                break;
            }
            order.push_back(cell[bu][bv].back());
            cell[bu][bv].pop_back();
            cnt[bu][bv]--;
            used_u[bu] = true;
        }
    }
    std::vector<uint32_t> new_payload(b.payload.begin(), b.payload.begin() + (long)skip), new_res, new_off{0};
    for (uint32_t it : order) {
        new_payload.push_back(items[it]);
        new_res.insert(new_res.end(), b.res.begin() + b.res_off[it], b.res.begin() + b.res_off[it + 1]);
        new_off.push_back((uint32_t)new_res.size());
    }
    b.payload.swap(new_payload);
    b.res.swap(new_res);
    b.res_off.swap(new_off);
}

}  // namespace

LoweredCircuit lower_circuit(const Circuit &c, uint32_t mode, uint32_t max_batch_words, bool with_sweep) {
    Lowerer lw(max_batch_words);
    lw.with_sweep = with_sweep;
    LoweredCircuit &lc = lw.lc;
    lc.mode = mode;
    lc.stats = compute_stats(c);
    if (lc.stats.num_measurements >= (1ull << 32) || lc.stats.num_detectors + lc.stats.num_observables >= (1ull << 32)) {
        throw std::invalid_argument("Circuit has more than 2^32 measurements or detectors; not supported.");
    }

    // qubit compaction: every qubit value that appears as a target of any instruction except QUBIT_COORDS.
    std::vector<uint8_t> used(lc.stats.num_qubits, 0);
    mark_used(c, used);
    lc.qubit_map.assign(lc.stats.num_qubits, UINT32_MAX);
    uint32_t Q = 0;
    for (size_t q = 0; q < used.size(); q++) {
        if (used[q]) {
            lc.qubit_map[q] = Q++;
        }
    }
    if (Q > 65535) {
        throw std::invalid_argument("Circuits with more than 65535 active qubits are not supported by this build.");
    }
    lc.num_qubits = Q;
    lw.Q = Q;

    // record addressing
    uint64_t rec_slots;
    if (mode == 0) {
        // The ring must also hold every result of one instruction at once: a noisy measurement / heralded channel with
        // more targets than ring slots would alias two of its record rows (and cut its RNG group inside a slice).
        uint64_t most_results = 0;
        c.for_each_instruction_once([&](const Instruction &op) {
            switch (op.gate->cat) {
                case GateCat::MEASURE:
                case GateCat::MPAD:
                case GateCat::MPP:
                case GateCat::MPAIR:
                case GateCat::HERALDED_ERASE:
                case GateCat::HERALDED_PAULI_CHANNEL_1:
                    most_results = std::max<uint64_t>(most_results, op.targets.size());  // (an upper bound for MPP / MPAIR)
                    break;
                default:
                    break;
            }
        });
        if (lc.stats.max_lookback + most_results >= (1ull << 31)) {
            throw std::invalid_argument("Measurement record window too large for this build.");
        }
        uint32_t ring = 1;
        while (ring < lc.stats.max_lookback + most_results) {
            ring <<= 1;
        }
        lc.rec_ring = ring;
        lw.rec_mask = ring - 1;
        rec_slots = ring;
    } else {
        lc.rec_ring = 0;
        lw.rec_mask = 0xFFFFFFFFu;
        rec_slots = std::max<uint64_t>(lc.stats.num_measurements, 1);
    }
    lw.res_clock = Q;
    lw.res_flag = Q + 1;
    lw.res_rec0 = Q + 2;
    lw.res_out0 = lw.res_rec0 + (uint32_t)rec_slots;
    lc.num_resources = lw.res_out0 + (uint32_t)(lc.stats.num_detectors + lc.stats.num_observables) + 1;
    lw.rd_stamp.assign(lc.num_resources, 0);
    lw.wr_stamp.assign(lc.num_resources, 0);

    if (with_sweep) {
        lc.out_recs.assign(lc.stats.num_detectors + lc.stats.num_observables, {});
    }
    // Start of every shot: x <- 0, z <- random for all qubits (frame_simulator.inl:153-163).
    // Measure group 0 is reserved for this.
    {
        uint32_t mg = lw.mgroup++;  // measure group 0, also when the circuit has no qubits
        for (uint32_t q = 0; q < Q; q++) {
            lw.measure(GB_Z, GK_R, q, mg);
        }
    }
    // Observable rows start at zero.
    if (mode == 0) {
        for (uint64_t l = 0; l < lc.stats.num_observables; l++) {
            lw.xor_rows((uint32_t)(lc.stats.num_detectors + l), {}, false);
        }
    }
    c.for_each_operation([&](const Instruction &op) {
        lw.do_op(op);
    });
    lw.flush();
    lw.fuse_detectors();
    assign_physical_rows(lc);
    for (Batch &b : lc.batches) {
        spread_banks(b);
    }
    lc.num_sites = lw.ngroup;
    lc.num_csites = lw.mgroup;
    return std::move(lw.lc);
}

LoweredCircuit lower_fragment(const Circuit &c, uint32_t num_qubits, uint64_t meas0) {
    Lowerer lw(1u << 20);
    LoweredCircuit &lc = lw.lc;
    lc.mode = 0;
    lc.stats = compute_stats(c);
    const uint64_t total_meas = meas0 + lc.stats.num_measurements;
    if (total_meas >= (1ull << 31) || lc.stats.num_detectors + lc.stats.num_observables >= (1ull << 31)) {
        throw std::invalid_argument("Too many measurements or detectors for the interactive simulator.");
    }
    const uint32_t Q = std::max<uint32_t>(num_qubits, (uint32_t)lc.stats.num_qubits);
    if (Q > 65535) {
        throw std::invalid_argument("Circuits with more than 65535 qubits are not supported by this build.");
    }
    lc.qubit_map.resize(Q);
    for (uint32_t q = 0; q < Q; q++) {
        lc.qubit_map[q] = q;
    }
    lc.num_qubits = Q;
    lw.Q = Q;
    lc.rec_ring = 0;
    lw.rec_mask = 0xFFFFFFFFu;
    lw.meas = meas0;
    lw.res_clock = Q;
    lw.res_flag = Q + 1;
    lw.res_rec0 = Q + 2;
    lw.res_out0 = lw.res_rec0 + (uint32_t)std::max<uint64_t>(total_meas, 1);
    lc.num_resources = lw.res_out0 + (uint32_t)(lc.stats.num_detectors + lc.stats.num_observables) + 1;
    lw.rd_stamp.assign(lc.num_resources, 0);
    lw.wr_stamp.assign(lc.num_resources, 0);
    c.for_each_operation([&](const Instruction &op) {
        lw.do_op(op);
    });
    lw.flush();
    lc.num_sites = lw.ngroup;
    lc.num_csites = lw.mgroup;
    return std::move(lw.lc);
}

std::vector<uint32_t> serialize_program(LoweredCircuit &lc, uint32_t slots, uint32_t chunk_words, GstimPlan *plan) {
    std::vector<uint32_t> out;
    const uint32_t NONE = 0xFFFFFFFFu, MULTI = 0xFFFFFFFEu;
    std::vector<uint32_t> w_epoch(lc.num_resources, 0), w_slot(lc.num_resources, NONE);
    std::vector<uint32_t> r_epoch(lc.num_resources, 0), r_slot(lc.num_resources, NONE);
    uint32_t epoch = 1;
    uint32_t n_barriers = 0;

    auto put_header = [&](uint32_t op, uint32_t words) {
        size_t base = out.size();
        out.resize(base + GSTIM_HDR_WORDS, 0);
        out[base + GH_OP] = op;
        out[base + GH_WORDS] = words;
    };

    // physical row -> logical qubit table (read by the host and by oracle/program_emulator.py; the kernel skips it)
    {
        const uint32_t per = chunk_words - 2 * GSTIM_HDR_WORDS;
        for (size_t base0 = 0; base0 < lc.logical_of.size(); base0 += per) {
            uint32_t cnt = (uint32_t)std::min<size_t>(per, lc.logical_of.size() - base0);
            uint32_t pos = (uint32_t)(out.size() % chunk_words);
            if (pos + GSTIM_HDR_WORDS + cnt + GSTIM_HDR_WORDS > chunk_words) {
                put_header(GOP_NEXT_CHUNK, GSTIM_HDR_WORDS);
                out.resize((out.size() + chunk_words - 1) / chunk_words * (size_t)chunk_words, 0);
            }
            size_t hb = out.size();
            put_header(GOP_QMAP, GSTIM_HDR_WORDS + cnt);
            out[hb + GH_N] = cnt;
            out[hb + GH_EXTRA] = (uint32_t)base0;
            out.insert(out.end(), lc.logical_of.begin() + (long)base0, lc.logical_of.begin() + (long)(base0 + cnt));
        }
    }

    // Noise schedule (consumed by the kernel's event pre-pass): per noise batch an info record, per
    // physical clock row the ordered list of its noise sites.
    NoiseSchedule &ns = lc.noise;
    ns = NoiseSchedule();
    std::vector<uint32_t> group_items(lc.num_sites + 1, 0);  // sites of each noise group serialised so far
    auto rate_index = [&](uint64_t lam) -> uint32_t {
        for (size_t i = 0; i < ns.rates.size(); i += 2) {
            if (ns.rates[i] == lam) {
                return (uint32_t)(i / 2);
            }
        }
        if (ns.rates.size() / 2 >= 65536) {
            throw std::invalid_argument("Circuits with more than 65536 distinct noise probabilities are not supported by this build.");
        }
        ns.rates.push_back(lam);
        ns.rates.push_back(lam ? 0xFFFFFFFFFFFFFFFFull / lam : 0ull);
        return (uint32_t)(ns.rates.size() / 2) - 1;
    };

    bool prev_was_noise = false;
    for (Batch &b : lc.batches) {
        const bool is_noise = b.op == GOP_NOISE1 || b.op == GOP_NOISE2;
        if (is_noise) {
            epoch++;  // noise events are applied by arbitrary threads: the kernel brackets these batches with barriers
            // ... except that the exit barrier of a directly preceding noise batch already is this batch's entry barrier
            if (prev_was_noise) {
                b.flags |= GF_NOENTRY;
            } else {
                b.flags &= ~GF_NOENTRY;
            }
        }
        prev_was_noise = is_noise;
        // ---- hazard analysis: does any item need data last touched by another thread group? ----
        size_t n_haz = b.res_off.size() - 1;
        bool barrier = false;
        for (size_t i = 0; i < n_haz && !barrier; i++) {
            uint32_t slot = (uint32_t)(i % slots);
            for (uint32_t j = b.res_off[i]; j < b.res_off[i + 1]; j++) {
                uint32_t r = b.res[j] & ~RES_WRITE;
                bool wr = (b.res[j] & RES_WRITE) != 0;
                if (w_epoch[r] == epoch && w_slot[r] != slot) {
                    barrier = true;
                    break;
                }
                if (wr && r_epoch[r] == epoch && r_slot[r] != slot) {
                    barrier = true;
                    break;
                }
            }
        }
        if (barrier) {
            epoch++;
            n_barriers++;
            b.flags |= GF_BARRIER;
        } else {
            b.flags &= ~GF_BARRIER;
        }
        for (size_t i = 0; i < n_haz; i++) {
            uint32_t slot = (uint32_t)(i % slots);
            for (uint32_t j = b.res_off[i]; j < b.res_off[i + 1]; j++) {
                uint32_t r = b.res[j] & ~RES_WRITE;
                if (b.res[j] & RES_WRITE) {
                    w_epoch[r] = epoch;
                    w_slot[r] = slot;
                } else if (r_epoch[r] != epoch) {
                    r_epoch[r] = epoch;
                    r_slot[r] = slot;
                } else if (r_slot[r] != slot) {
                    r_slot[r] = MULTI;
                }
            }
        }

        if (is_noise) {
            epoch++;
        }

        // ---- serialise ----
        uint32_t words = b.words();
        uint32_t pos = (uint32_t)(out.size() % chunk_words);
        if (words + GSTIM_HDR_WORDS > chunk_words) {
            throw std::logic_error("internal: batch larger than a program chunk");
        }
        if (pos + words + GSTIM_HDR_WORDS > chunk_words) {
            put_header(GOP_NEXT_CHUNK, GSTIM_HDR_WORDS);
            out.resize((out.size() + chunk_words - 1) / chunk_words * (size_t)chunk_words, 0);
        }
        size_t base = out.size();
        out.resize(base + GSTIM_HDR_WORDS, 0);
        out[base + GH_OP] = b.op | (b.flags << 8) | (b.aux << 16);
        out[base + GH_N] = b.n_items;
        out[base + GH_WORDS] = words;
        out[base + GH_EXTRA] = b.extra;
        uint64_t lb = b.lambda;
        out[base + GH_LAMBDA_LO] = (uint32_t)lb;
        out[base + GH_LAMBDA_HI] = (uint32_t)(lb >> 32);
        out[base + GH_SITE0] = b.site0;
        out[base + GH_CSITE0] = b.csite0;
        out[base + GH_REC0] = b.rec0;
        out[base + GH_T1] = b.t1;
        out[base + GH_T2] = b.t2;
        out[base + GH_T3] = b.t3;
        if (b.op == GOP_XORROWS) {
            out.insert(out.end(), b.dst.begin(), b.dst.end());
            out.insert(out.end(), b.off.begin(), b.off.end());
            out.insert(out.end(), b.idx.begin(), b.idx.end());
        } else {
            out.insert(out.end(), b.payload.begin(), b.payload.end());
        }

        if (is_noise || b.op == GOP_CORR) {
            const uint32_t nbi = (uint32_t)(ns.info.size() / GSTIM_NOISE_INFO_WORDS);
            if (nbi >= 65536) {
                throw std::invalid_argument("Circuits with more than 65536 noise batches are not supported by this build.");
            }
            out[base + GH_CSITE0] = nbi;
            const bool table = b.op == GOP_NOISE2 && (b.flags & GF_TABLE);
            uint32_t info[GSTIM_NOISE_INFO_WORDS] = {};
            info[GNI_H0] = out[base + GH_OP];
            info[GNI_N] = b.op == GOP_CORR ? 1u : b.n_items;
            info[GNI_LAM_LO] = (uint32_t)lb;
            info[GNI_LAM_HI] = (uint32_t)(lb >> 32);
            info[GNI_GROUP] = b.site0;
            info[GNI_T1] = b.t1;
            info[GNI_T2] = b.t2;
            info[GNI_T3] = b.t3;
            info[GNI_TABLE_OFF] = table ? (uint32_t)(base + GSTIM_HDR_WORDS) : 0;
            ns.info.insert(ns.info.end(), info, info + GSTIM_NOISE_INFO_WORDS);
            ns.n_sites.push_back(info[GNI_N]);
            ns.lams.push_back(lb);
            // RNG slices of this batch (program.h "Noise schedule"): GSTIM_NOISE_SLICE consecutive sites of the group each
            if (b.site0 >= group_items.size()) {
                group_items.resize((size_t)b.site0 + 1, 0);
            }
            const uint32_t gfirst = group_items[b.site0];
            group_items[b.site0] += info[GNI_N];
            if (lb != 0) {
                const uint32_t rate = rate_index(lb);
                if (gfirst % GSTIM_NOISE_SLICE != 0) {
                    throw std::logic_error("internal: a noise group was cut inside an RNG slice");
                }
                for (uint32_t i0 = 0; i0 < info[GNI_N]; i0 += GSTIM_NOISE_SLICE) {
                    const uint32_t cnt = std::min<uint32_t>(GSTIM_NOISE_SLICE, info[GNI_N] - i0);
                    const uint32_t sl[GSTIM_SLICE_WORDS] = {b.site0, (gfirst + i0) / GSTIM_NOISE_SLICE, nbi | (rate << 16), i0 | (cnt << 11),
                                                            info[GNI_H0], b.t1, b.t2, b.t3};
                    ns.slices.insert(ns.slices.end(), sl, sl + GSTIM_SLICE_WORDS);
                }
            }
        }
    }
    put_header(GOP_END, GSTIM_HDR_WORDS);
    out.resize((out.size() + chunk_words - 1) / chunk_words * (size_t)chunk_words, 0);

    if (plan != nullptr) {
        memset(plan, 0, sizeof(*plan));
        plan->num_qubits = lc.num_qubits;
        plan->q_pitch = lc.num_qubits | 1u;
        plan->num_meas = (uint32_t)lc.stats.num_measurements;
        plan->num_det = (uint32_t)lc.stats.num_detectors;
        plan->num_obs = (uint32_t)lc.stats.num_observables;
        plan->rec_ring = lc.rec_ring;
        plan->n_words = (uint32_t)out.size();
        plan->chunk_words = chunk_words;
        plan->n_chunks = (uint32_t)(out.size() / chunk_words);
        plan->slots = slots;
        plan->mode = lc.mode;
        plan->max_items = lc.max_items;
        plan->n_batches = (uint32_t)lc.batches.size();
        plan->n_barriers = n_barriers;
    }
    return out;
}

}  // namespace gstim
