// response.cc — see response.h.
#include "response.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>

namespace gstim {

namespace {

constexpr uint32_t ITEM_X = 1u << 30;  // component flags in OBS_PAULI / FEEDBACK / CORR payload words (lowering.cc)
constexpr uint32_t ITEM_Z = 1u << 31;
constexpr uint64_t LAM_MAX = 1ull << 62;
constexpr uint64_t LAM_HALF = 49946518145322872ull;  // floor(ln 2 * 2^56): the rate of a fair coin (collapse bits)

using Set = std::vector<uint32_t>;  // sorted output ids

// a ^= b (symmetric difference of sorted sets)
void xor_into(Set &a, const Set &b, Set &tmp) {
    if (b.empty()) {
        return;
    }
    if (a.empty()) {
        a = b;
        return;
    }
    tmp.clear();
    tmp.reserve(a.size() + b.size());
    size_t i = 0, j = 0;
    while (i < a.size() && j < b.size()) {
        if (a[i] < b[j]) {
            tmp.push_back(a[i++]);
        } else if (b[j] < a[i]) {
            tmp.push_back(b[j++]);
        } else {
            i++;
            j++;
        }
    }
    tmp.insert(tmp.end(), a.begin() + i, a.end());
    tmp.insert(tmp.end(), b.begin() + j, b.end());
    a.swap(tmp);
}

void xor_one(Set &a, uint32_t v) {
    auto it = std::lower_bound(a.begin(), a.end(), v);
    if (it != a.end() && *it == v) {
        a.erase(it);
    } else {
        a.insert(it, v);
    }
}

struct ClassKey {
    uint64_t lam;
    uint32_t kind, n_out;
    uint32_t thr[15];
    uint32_t chain_off[16];  // RK_CHAIN: noise group of element i minus that of element 0
    bool operator<(const ClassKey &o) const {
        if (lam != o.lam) {
            return lam < o.lam;
        }
        if (kind != o.kind) {
            return kind < o.kind;
        }
        if (n_out != o.n_out) {
            return n_out < o.n_out;
        }
        const int c = memcmp(thr, o.thr, sizeof(thr));
        if (c != 0) {
            return c < 0;
        }
        return memcmp(chain_off, o.chain_off, sizeof(chain_off)) < 0;
    }
};

// Sites are discovered in reverse program order; every class collects its own entry words and is reversed at the end.
struct ClassAcc {
    ClassKey key;
    std::vector<uint32_t> rep_word;   // outcome -> representative chooser word
    // responses, n_out per site, in discovery (reverse) order: one flat array of ids + end offsets (a vector per response
    // meant nine million small allocations for a d = 51 memory experiment)
    std::vector<uint32_t> resp_ids;
    std::vector<uint64_t> resp_end;
    void push(const Set &s) {
        resp_ids.insert(resp_ids.end(), s.begin(), s.end());
        resp_end.push_back(resp_ids.size());
    }
    size_t n_responses() const {
        return resp_end.size();
    }
    std::vector<uint32_t> group, index;
};

// 1 / lambda = m 2^e with m in [0.5, 1): INV = floor(2^32 m), SH = 58 - e (dem.cu's gap arithmetic).
bool gap_params(uint64_t lam, uint32_t *inv, uint32_t *sh) {
    if (lam == 0) {
        return false;
    }
    if (lam >= LAM_MAX) {
        *inv = 0;
        *sh = 0;
        return true;
    }
    int e = 0;
    const double m = std::frexp(std::ldexp(1.0, 56) / (double)lam, &e);
    const int s = 58 - e;
    if (s < 0 || s > 63) {
        return false;
    }
    *inv = (uint32_t)std::min<double>(std::floor(std::ldexp(m, 32)), 4294967295.0);
    *sh = (uint32_t)s;
    return true;
}

}  // namespace

ResponseTable build_response_table(const LoweredCircuit &lc, bool keep_conjugate) {
    ResponseTable rt;
    const uint32_t D = (uint32_t)lc.stats.num_detectors, L = (uint32_t)lc.stats.num_observables;
    const uint32_t M = (uint32_t)lc.stats.num_measurements;
    const bool det_mode = lc.mode == 0;
    rt.n_outputs = det_mode ? D + L : M;
    const uint32_t Q = lc.num_qubits;
    const uint32_t rec_mask = det_mode ? lc.rec_ring - 1 : 0xFFFFFFFFu;
    const size_t n_rec = det_mode ? lc.rec_ring : std::max<uint32_t>(M, 1);

    auto fail = [&](const std::string &why) {
        rt.eligible = false;
        rt.why_not = why;
        rt.classes.clear();
        rt.entries.clear();
        rt.overflow.clear();
        rt.sweep_responses.clear();
        return rt;
    };

    // Forward pass: index of every noise batch's first site inside its noise group.
    std::vector<uint32_t> gfirst(lc.batches.size(), 0);
    {
        std::map<uint32_t, uint32_t> seen;
        for (size_t b = 0; b < lc.batches.size(); b++) {
            const Batch &B = lc.batches[b];
            if (B.op == GOP_NOISE1 || B.op == GOP_NOISE2) {
                uint32_t &n = seen[B.site0];
                gfirst[b] = n;
                n += B.n_items;
            }
        }
    }

    std::vector<Set> SX(Q + 1), SZ(Q + 1), SR(n_rec);
    std::vector<uint8_t> out_dead(rt.n_outputs + 1, 0);
    if (!det_mode) {
        for (uint32_t m = 0; m < M; m++) {
            SR[m] = {m};
        }
    }
    Set tmp, acc, comp[4], by_mask[16];
    std::map<ClassKey, size_t> class_of;
    std::vector<ClassAcc> accs;
    uint64_t total_ids = 0;
    const uint64_t MAX_IDS = 1ull << 28;

    auto get_class = [&](const ClassKey &k, const uint32_t *rep) -> ClassAcc & {
        auto it = class_of.find(k);
        if (it == class_of.end()) {
            it = class_of.emplace(k, accs.size()).first;
            accs.emplace_back();
            accs.back().key = k;
            accs.back().rep_word.assign(rep, rep + k.n_out);
        }
        return accs[it->second];
    };

    // E / ELSE_CORRELATED_ERROR chain under construction (elements arrive last to first): "the first element whose coin
    // comes up applies" (frame_simulator.inl:747-776) is ONE site with total rate sum(lambda_i) whose outcome i has
    // probability p_i prod_{j<i} (1 - p_j) - the joint distribution of the reference's per-element coins.
    struct ChainElem {
        uint64_t lam;
        uint32_t group;
        Set response;
    };
    std::vector<ChainElem> chain;
    bool chain_failed = false;
    auto finish_chain = [&]() {
        if (chain.empty()) {
            return;
        }
        std::reverse(chain.begin(), chain.end());
        if (chain.size() > 16) {
            chain_failed = true;
            chain.clear();
            return;
        }
        ClassKey k{};
        uint32_t rep[16];
        const uint32_t no = (uint32_t)chain.size();
        double survive = 1.0, cum = 0.0;
        std::vector<double> cums;
        unsigned __int128 lam_tot = 0;
        for (uint32_t i = 0; i < no; i++) {
            const double pi = chain[i].lam >= LAM_MAX ? 1.0 : -std::expm1(-std::ldexp((double)chain[i].lam, -56));
            cum += pi * survive;
            survive *= 1.0 - pi;
            cums.push_back(cum);
            lam_tot += chain[i].lam;
            rep[i] = chain[i].group - chain[0].group;  // chain classes: group offset of the element (see response.h)
        }
        k.lam = lam_tot >= LAM_MAX ? LAM_MAX : (uint64_t)lam_tot;
        k.n_out = no;
        k.kind = no == 1 ? RK_SINGLE : RK_CHAIN;
        for (uint32_t i = 0; i + 1 < no; i++) {
            const double v = std::floor(cums[i] / cum * 4294967296.0);
            k.thr[i] = v >= 4294967295.0 ? 0xFFFFFFFFu : v < 1.0 ? 1u : (uint32_t)v;
            if (i > 0 && k.thr[i] <= k.thr[i - 1]) {
                k.thr[i] = k.thr[i - 1] == 0xFFFFFFFFu ? 0xFFFFFFFFu : k.thr[i - 1] + 1;  // keep the bounds strictly ascending
            }
        }
        if (no > 1) {
            // the group offsets are part of the class identity (they tell the tests which instruction an outcome is)
            for (uint32_t i = 0; i < no; i++) {
                k.chain_off[i] = rep[i];
            }
        }
        ClassAcc &c = get_class(k, rep);
        for (uint32_t o = no; o-- > 0;) {
            total_ids += chain[o].response.size();
            c.push(chain[o].response);
        }
        c.group.push_back(chain[0].group);
        c.index.push_back(0);
        chain.clear();
    };

    for (size_t bi = lc.batches.size(); bi-- > 0;) {
        const Batch &B = lc.batches[bi];
        const uint32_t n = B.n_items;
        switch (B.op) {
            case GOP_CLIFF1: {
                const uint32_t a = B.aux & 1, b = (B.aux >> 1) & 1, c = (B.aux >> 2) & 1, d = (B.aux >> 3) & 1;
                if (a && !b && !c && d) {
                    break;
                }
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t q = B.payload[i];
                    // x' = a x ^ b z, z' = c x ^ d z  =>  SX_before = a SX' ^ c SZ', SZ_before = b SX' ^ d SZ'
                    Set nx, nz;
                    if (a) {
                        xor_into(nx, SX[q], tmp);
                    }
                    if (c) {
                        xor_into(nx, SZ[q], tmp);
                    }
                    if (b) {
                        xor_into(nz, SX[q], tmp);
                    }
                    if (d) {
                        xor_into(nz, SZ[q], tmp);
                    }
                    SX[q].swap(nx);
                    SZ[q].swap(nz);
                }
                break;
            }
            case GOP_CLIFF2: {
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t q1 = B.payload[i] & 0xFFFF, q2 = B.payload[i] >> 16;
                    Set *after[4] = {&SX[q1], &SZ[q1], &SX[q2], &SZ[q2]};
                    if (B.aux == GSTIM_MAT_CX) {
                        // x2 ^= x1, z1 ^= z2  =>  SX1 ^= SX2', SZ2 ^= SZ1'
                        xor_into(SX[q1], SX[q2], tmp);
                        xor_into(SZ[q2], SZ[q1], tmp);
                        continue;
                    }
                    Set before[4];
                    for (int j = 0; j < 4; j++) {
                        for (int k = 0; k < 4; k++) {
                            if ((B.aux >> (4 * k + j)) & 1) {  // output k reads input j
                                xor_into(before[j], *after[k], tmp);
                            }
                        }
                    }
                    for (int j = 0; j < 4; j++) {
                        after[j]->swap(before[j]);
                    }
                }
                break;
            }
            case GOP_MEASURE: {
                const uint32_t basis = B.aux & 3, kind = (B.aux >> 2) & 3;
                const uint32_t stride = (B.flags & GF_DET) ? 3 : 1;
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t q = B.payload[stride * i] & 0xFFFF, lq = B.payload[stride * i] >> 16;
                    Set sm;
                    if (kind != GK_R) {
                        const uint32_t slot = (B.rec0 + i) & rec_mask;
                        if ((B.flags & GF_DET) && B.payload[3 * i + 1] != 0xFFFFFFFFu) {
                            const uint32_t d = B.payload[3 * i + 1], other = B.payload[3 * i + 2];
                            if (!out_dead[d]) {
                                xor_one(sm, d);
                                xor_one(SR[other], d);
                            }
                            out_dead[d] = 1;
                        }
                        xor_into(sm, SR[slot], tmp);
                        SR[slot].clear();
                    }
                    if (keep_conjugate) {
                        if (basis == GB_Z) {  // m = x; x kept or cleared; z kept
                            if (kind != GK_M) {
                                SX[q].clear();
                            }
                            xor_into(SX[q], sm, tmp);
                        } else if (basis == GB_X) {
                            if (kind != GK_M) {
                                SZ[q].clear();
                            }
                            xor_into(SZ[q], sm, tmp);
                        } else if (kind == GK_M) {  // m = x ^ z, frame unchanged
                            xor_into(SX[q], sm, tmp);
                            xor_into(SZ[q], sm, tmp);
                        } else {  // m = x ^ z; x <- z, z kept
                            xor_into(SZ[q], SX[q], tmp);
                            xor_into(SZ[q], sm, tmp);
                            SX[q] = sm;
                        }
                        continue;
                    }
                    Set rr;  // response of the collapse randomisation bit
                    if (basis == GB_Z) {
                        rr.swap(SZ[q]);
                        if (kind != GK_M) {
                            SX[q].clear();
                        }
                        xor_into(SX[q], sm, tmp);
                    } else if (basis == GB_X) {
                        rr.swap(SX[q]);
                        if (kind != GK_M) {
                            SZ[q].clear();
                        }
                        xor_into(SZ[q], sm, tmp);
                    } else {
                        rr = SX[q];
                        xor_into(rr, SZ[q], tmp);
                        if (kind != GK_M) {
                            SX[q].clear();
                        }
                        xor_into(SX[q], sm, tmp);
                        SZ[q] = SX[q];
                    }
                    if (!rr.empty()) {
                        ClassKey k{};
                        k.lam = LAM_HALF;
                        k.kind = RK_SINGLE;
                        k.n_out = 1;
                        const uint32_t rep = 0;
                        ClassAcc &c = get_class(k, &rep);
                        total_ids += rr.size();
                        c.push(rr);
                        c.group.push_back(0x80000000u | B.csite0);
                        c.index.push_back(lq);
                    }
                }
                break;
            }
            case GOP_RECZERO:
                for (uint32_t i = 0; i < n; i++) {
                    SR[(B.rec0 + i) & rec_mask].clear();
                }
                break;
            case GOP_XORROWS: {
                for (size_t i = 0; i < B.dst.size(); i++) {
                    const uint32_t d = B.dst[i];
                    if (!out_dead[d]) {
                        for (uint32_t j = B.off[i]; j < B.off[i + 1]; j++) {
                            xor_one(SR[B.idx[j]], d);
                        }
                    }
                    if (!(B.flags & GF_ACCUM)) {
                        out_dead[d] = 1;
                    }
                }
                break;
            }
            case GOP_OBS_PAULI:
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t d = B.payload[2 * i], wq = B.payload[2 * i + 1], q = wq & 0xFFFFFF;
                    if (out_dead[d]) {
                        continue;
                    }
                    if (wq & ITEM_X) {
                        xor_one(SX[q], d);
                    }
                    if (wq & ITEM_Z) {
                        xor_one(SZ[q], d);
                    }
                }
                break;
            case GOP_FEEDBACK:
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t ri = B.payload[2 * i], wq = B.payload[2 * i + 1], q = wq & 0xFFFFFF;
                    if (wq & ITEM_X) {
                        xor_into(SR[ri], SX[q], tmp);
                    }
                    if (wq & ITEM_Z) {
                        xor_into(SR[ri], SZ[q], tmp);
                    }
                }
                break;
            case GOP_SWEEP:
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t k = B.payload[2 * i], wq = B.payload[2 * i + 1], q = wq & 0xFFFFFF;
                    if (rt.sweep_responses.size() <= k) {
                        rt.sweep_responses.resize(k + 1);
                    }
                    if (wq & ITEM_X) {
                        xor_into(rt.sweep_responses[k], SX[q], tmp);
                    }
                    if (wq & ITEM_Z) {
                        xor_into(rt.sweep_responses[k], SZ[q], tmp);
                    }
                }
                break;
            case GOP_NOISE1: {
                if (B.lambda == 0) {
                    break;
                }
                // outcomes = non-empty ranges of the chooser (program.h "NOISE1 aux")
                const uint64_t upper[4] = {B.t1, B.t2, B.t3, 1ull << 32};
                ClassKey k{};
                k.lam = B.lambda;
                uint32_t cats[4], rep[4], no = 0;
                uint64_t lo = 0;
                for (int j = 0; j < 4; j++) {
                    if (upper[j] > lo) {  // v in [lo, upper[j]) selects category j
                        cats[no] = (B.aux >> (2 * j)) & 3u;
                        rep[no] = (uint32_t)lo;
                        if (no > 0) {
                            k.thr[no - 1] = (uint32_t)lo;
                        }
                        no++;
                        lo = upper[j];
                    }
                }
                k.n_out = no;
                k.kind = no == 1 ? RK_SINGLE : RK_THRESH;
                ClassAcc &c = get_class(k, rep);
                for (uint32_t ii = n; ii-- > 0;) {
                    const uint32_t q = (B.flags & GF_NOFRAME) ? Q : B.payload[ii];
                    // (reverse discovery order: the outcomes of a site are pushed last-to-first and the whole list is reversed)
                    for (uint32_t o = no; o-- > 0;) {
                        acc.clear();
                        if (!(B.flags & GF_NOFRAME)) {
                            if (cats[o] & 1) {
                                xor_into(acc, SX[q], tmp);
                            }
                            if (cats[o] & 2) {
                                xor_into(acc, SZ[q], tmp);
                            }
                        }
                        if (B.flags & GF_REC) {
                            xor_into(acc, SR[(B.rec0 + ii) & rec_mask], tmp);
                        }
                        total_ids += acc.size();
                        c.push(acc);
                    }
                    c.group.push_back(B.site0);
                    c.index.push_back(gfirst[bi] + ii);
                }
                break;
            }
            case GOP_NOISE2: {
                if (B.lambda == 0) {
                    break;
                }
                const bool table = (B.flags & GF_TABLE) != 0;
                const uint32_t *items = B.payload.data() + (table ? 15 : 0);
                ClassKey k{};
                k.lam = B.lambda;
                uint32_t masks[16], rep[16], no = 0;
                if (!table) {
                    k.kind = RK_UNIFORM;
                    k.n_out = no = 15;
                    for (uint32_t o = 0; o < 15; o++) {
                        masks[o] = o + 1;  // Pauli pr = o + 1: bits x1, z1, x2, z2
                        rep[o] = (uint32_t)((((uint64_t)o << 32) + 14) / 15);  // smallest v with mulhi(v, 15) == o
                    }
                } else {
                    uint64_t lo = 0;
                    for (int j = 0; j < 16; j++) {
                        const uint64_t hi = j < 15 ? (uint64_t)B.payload[j] : (1ull << 32);
                        if (hi > lo) {
                            const uint32_t pr = j < 15 ? (uint32_t)j + 1 : B.aux;
                            const uint32_t c1 = pr >> 2, c2 = pr & 3;
                            masks[no] = (((c1 + 1) >> 1) & 1) | ((c1 >> 1) << 1) | ((((c2 + 1) >> 1) & 1) << 2) | ((c2 >> 1) << 3);
                            rep[no] = (uint32_t)lo;
                            if (no > 0) {
                                k.thr[no - 1] = (uint32_t)lo;
                            }
                            no++;
                            lo = hi;
                        }
                    }
                    k.n_out = no;
                    k.kind = no == 1 ? RK_SINGLE : RK_THRESH;
                }
                ClassAcc &c = get_class(k, rep);
                for (uint32_t ii = n; ii-- > 0;) {
                    const uint32_t q1 = items[ii] & 0xFFFF, q2 = items[ii] >> 16;
                    const Set *cs[4] = {&SX[q1], &SZ[q1], &SX[q2], &SZ[q2]};
                    if (no >= 8) {
                        // (DEPOLARIZE2) the XOR of every subset of the four sets, in Gray-code order: one merge per subset
                        by_mask[0].clear();
                        for (uint32_t g = 1, prev = 0; g < 16; g++) {
                            const uint32_t cur = g ^ (g >> 1), bit = (uint32_t)__builtin_ctz(cur ^ prev);
                            by_mask[cur] = by_mask[prev];
                            xor_into(by_mask[cur], *cs[bit], tmp);
                            prev = cur;
                        }
                        for (uint32_t o = no; o-- > 0;) {
                            const Set &r = by_mask[masks[o] & 15u];
                            total_ids += r.size();
                            c.push(r);
                        }
                    } else {
                        for (uint32_t o = no; o-- > 0;) {
                            acc.clear();
                            for (int j = 0; j < 4; j++) {
                                if ((masks[o] >> j) & 1) {
                                    xor_into(acc, *cs[j], tmp);
                                }
                            }
                            total_ids += acc.size();
                            c.push(acc);
                        }
                    }
                    c.group.push_back(B.site0);
                    c.index.push_back(gfirst[bi] + ii);
                }
                break;
            }
            case GOP_CORR: {
                if (B.lambda != 0) {
                    acc.clear();
                    for (uint32_t i = 0; i < n; i++) {
                        const uint32_t w = B.payload[i], q = w & 0xFFFFFF;
                        if (w & ITEM_X) {
                            xor_into(acc, SX[q], tmp);
                        }
                        if (w & ITEM_Z) {
                            xor_into(acc, SZ[q], tmp);
                        }
                    }
                    chain.push_back({B.lambda, B.site0, acc});
                }
                if (B.flags & GF_RESET_FLAG) {
                    finish_chain();  // CORRELATED_ERROR starts the chain that the ELSEs after it continued
                }
                break;
            }
            default:
                break;
        }
        if (total_ids > MAX_IDS) {
            return fail("response table too large");
        }
    }

    finish_chain();  // (ELSE_CORRELATED_ERRORs before any CORRELATED_ERROR)
    if (chain_failed) {
        return fail("ELSE_CORRELATED_ERROR chain with more than 16 elements");
    }
    // assemble: classes in a deterministic order (rate, then chooser), sites in program order
    std::vector<size_t> order(accs.size());
    for (size_t i = 0; i < order.size(); i++) {
        order[i] = i;
    }
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return accs[a].key < accs[b].key; });
    uint64_t n_entries = 0;
    for (size_t ci : order) {
        n_entries += accs[ci].n_responses();
    }
    if (n_entries >= (1ull << 31) / 4) {
        return fail("response table too large");
    }
    rt.entries.reserve(n_entries * 4);
    for (size_t ci : order) {
        ClassAcc &a = accs[ci];
        RespClass rc;
        rc.lam = a.key.lam;
        if (!gap_params(a.key.lam, &rc.inv, &rc.sh)) {
            continue;  // a rate too small to ever fire
        }
        rc.kind = a.key.kind;
        rc.n_out = a.key.n_out;
        memcpy(rc.thr, a.key.thr, sizeof(rc.thr));
        rc.n_sites = (uint32_t)a.group.size();
        rc.entry0 = (uint32_t)(rt.entries.size() / 4);
        const double lam = std::ldexp((double)a.key.lam, -56);
        const double p = a.key.lam >= LAM_MAX ? 1.0 : -std::expm1(-lam);
        if (p < 1.0) {
            // packed Bernoulli words pay when the words per trial (0.47 instructions per level and trial) cost less than the
            // draws of the geometric walk (46 instructions per event): p = 1/2 needs one level, p = 1/4 two, a generic p ~30
            const double t = std::floor(p * 4294967296.0 + 0.5);
            const uint32_t thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
            if (a.key.lam == LAM_HALF) {
                rc.dense_thr = 0x80000000u;  // collapse bits: a fair coin, exactly
            } else if (thr != 0) {
                const int levels = 32 - __builtin_ctz(thr);
                if (levels * 0.47 < 46.0 * p) {
                    rc.dense_thr = thr;
                }
            }
        }
        rt.events_per_shot += p * rc.n_sites;
        const uint64_t class_ids = a.resp_ids.size();
        rt.flips_per_shot += p * (double)class_ids / rc.n_out;
        for (size_t r = a.n_responses(); r-- > 0;) {  // reverse discovery order = program order, outcomes ascending
            const uint32_t *s = a.resp_ids.data() + (r ? a.resp_end[r - 1] : 0);
            const size_t sn = (size_t)(a.resp_end[r] - (r ? a.resp_end[r - 1] : 0));
            rt.max_response = std::max<uint32_t>(rt.max_response, (uint32_t)sn);
            uint32_t w[4] = {RESP_NONE, RESP_NONE, RESP_NONE, RESP_NONE};
            if (sn <= 4) {
                for (size_t j = 0; j < sn; j++) {
                    w[j] = s[j];
                }
            } else {
                for (size_t j = 0; j < 3; j++) {
                    w[j] = s[j];
                }
                if (rt.overflow.size() + sn >= (1ull << 30)) {
                    return fail("response table too large");
                }
                w[3] = RESP_OVERFLOW | (uint32_t)rt.overflow.size();
                rt.overflow.push_back((uint32_t)(sn - 3));
                rt.overflow.insert(rt.overflow.end(), s + 3, s + sn);
            }
            rt.entries.insert(rt.entries.end(), w, w + 4);
        }
        for (size_t r = a.group.size(); r-- > 0;) {
            rt.site_group.push_back(a.group[r]);
            rt.site_index.push_back(a.index[r]);
        }
        rt.outcome_word.insert(rt.outcome_word.end(), a.rep_word.begin(), a.rep_word.end());
        rt.n_sites += rc.n_sites;
        rt.classes.push_back(rc);
    }
    rt.n_entries = rt.entries.size() / 4;
    if (rt.n_outputs >= (1u << 30)) {
        return fail("too many output bits");
    }
    rt.eligible = true;
    return rt;
}

namespace {

// Output ids of one table entry (with its overflow list).
void entry_ids(const ResponseTable &rt, size_t e, std::vector<uint32_t> &ids) {
    ids.clear();
    const uint32_t *w = &rt.entries[4 * e];
    for (int j = 0; j < 4; j++) {
        if (w[j] == RESP_NONE) {
            break;
        }
        if (j == 3 && (w[j] & RESP_OVERFLOW)) {
            const uint32_t off = w[j] & 0x7FFFFFFFu, cnt = rt.overflow[off];
            ids.insert(ids.end(), rt.overflow.begin() + off + 1, rt.overflow.begin() + off + 1 + cnt);
            break;
        }
        ids.push_back(w[j]);
    }
}

// Looks for the round structure of a class: a period P and a detector shift delta with
//   entries(site s + P) == entries(site s) with delta added to every detector id (ids < D), observables unchanged,
// for all s in [a, a + n - P). Unrolled REPEAT blocks of QEC circuits have it (P = sites of the class per round, delta =
// detectors per round). Returns p = 0 when there is no such structure covering at least three periods.
}  // namespace

ResponsePeriod find_response_period(const ResponseTable &rt, const RespClass &c, uint32_t D) {
    ResponsePeriod out;
    const uint32_t n = c.n_sites, no = c.n_out;
    if (n < 6) {
        return out;
    }
    // per site: hash of its responses relative to its smallest detector id, and that id
    std::vector<uint64_t> h(n);
    std::vector<uint32_t> base(n);
    std::vector<uint32_t> ids;
    for (uint32_t s = 0; s < n; s++) {
        uint32_t b = 0xFFFFFFFFu;
        for (uint32_t o = 0; o < no; o++) {
            entry_ids(rt, (size_t)c.entry0 + (size_t)s * no + o, ids);
            for (uint32_t v : ids) {
                if (v < D) {
                    b = std::min(b, v);
                }
            }
        }
        uint64_t x = 1469598103934665603ull;
        for (uint32_t o = 0; o < no; o++) {
            entry_ids(rt, (size_t)c.entry0 + (size_t)s * no + o, ids);
            x = (x ^ (0x100u + ids.size())) * 1099511628211ull;
            for (uint32_t v : ids) {
                x = (x ^ (v < D ? v - b : 0x80000000u | v)) * 1099511628211ull;
            }
        }
        h[s] = x;
        base[s] = b;
    }
    auto match = [&](uint32_t s, uint32_t t, uint32_t delta) {
        if (h[s] != h[t]) {
            return false;
        }
        return base[s] == 0xFFFFFFFFu ? base[t] == 0xFFFFFFFFu : (base[t] != 0xFFFFFFFFu && base[t] - base[s] == delta);
    };
    const uint32_t mid = n / 2;
    uint64_t work = 0;  // (many sites of a round share their relative shape: most candidates fail within a few sites)
    uint32_t best_saving = 0;
    for (uint32_t P = 1; P <= n / 3 && mid + 2 * P <= n && work < (1ull << 27); P++) {
        if (h[mid + P] != h[mid] || base[mid] == 0xFFFFFFFFu || base[mid + P] == 0xFFFFFFFFu || base[mid + P] <= base[mid]) {
            continue;
        }
        const uint32_t delta = base[mid + P] - base[mid];
        bool ok = true;
        for (uint32_t j = 0; j < P && ok; j++) {
            ok = match(mid + j, mid + j + P, delta);
            work++;
        }
        if (!ok) {
            continue;
        }
        uint32_t a = mid, e = mid + P;  // match(s, s + P) holds for s in [a, e)
        while (a > 0 && match(a - 1, a - 1 + P, delta)) {
            a--;
        }
        while (e + P < n && match(e, e + P, delta)) {
            e++;
        }
        work += (mid - a) + (e - mid - P);
        const uint32_t covered = e + P - a;
        // (short periods exist too - neighbouring sites of a lattice row look alike - so keep the candidate that folds the
        // most sites away, and stop once most of the class is covered)
        if (covered >= 3 * P && covered - P > best_saving) {
            best_saving = covered - P;
            out.a = a;
            out.p = P;
            out.n = covered;
            out.delta = delta;
            if (best_saving >= n / 2) {
                break;
            }
        }
    }
    return out;
}

}  // namespace gstim
