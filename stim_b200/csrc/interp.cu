// interp.cu — the persistent sm_100a interpreter kernel of the Pauli-frame sampler.
//
// One thread block owns K*128 shots. The x/z frame bits of every qubit stay resident in shared memory
// as uint4 (128-shot) columns, the lowered program (program.h) is streamed through a two-stage
// shared-memory ring by bulk-async (TMA 1D, cp.async.bulk + mbarrier) copies, and every batch of the
// program is executed by all threads: item i by thread group i % slots, one lane per 128-shot column
// (or G lanes per item splitting the columns).
//
// Replaces FrameSimulator<W>::do_circuit / do_gate and the per-gate row loops
// (/root/reference/src/stim/simulators/frame_simulator.inl:166-170, 173-912), RareErrorIterator
// (/root/reference/src/stim/util_bot/probability_util.cc:23-43) and the MeasureRecordBatch window
// (/root/reference/src/stim/io/measure_record_batch.inl:49-104).
//
// Code structure: every opcode is a __noinline__ device function that returns the address of the next batch
// (so the program cursor never has to survive a call in the caller's registers or, worse, in local memory). A monolithic switch let the
// compiler hoist each case's loop invariants in front of the switch, which made every batch pay a
// few hundred instructions per warp no matter which opcode it was (profiles/r1 notes).
#include "kernels.cuh"

#define GSTIM_TABLE_QUAL __device__ const
#include "log2_table.h"

namespace gstim {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3").
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
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

// ------------------------------------------------------------------------------------------------
// shared-memory access through 32-bit shared-window addresses
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long lds64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}

// Explicit global-space accesses for the record ring / output table (a generic ST.E/LD.E takes the slow
// address-space-resolving path and holds its operand registers for a long scoreboard wait).
__device__ __forceinline__ void stg128(uint4 *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ldg128(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// A run of 32-bit words in shared memory, read with LDS (a generic pointer into the program ring would be read with
// generic loads: long-scoreboard latency on the per-batch critical path).
struct SmemWords {
    uint32_t s;
    __device__ __forceinline__ uint32_t operator[](uint32_t i) const {
        return lds32(s + 4 * i);
    }
    __device__ __forceinline__ SmemWords operator+(uint32_t n) const {
        return SmemWords{s + 4 * n};
    }
};

// mbarrier / bulk-async copy helpers (PTX ISA: mbarrier, cp.async.bulk)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (the source address is formed inside the asm block: when `base + offset` was left to the compiler, ptxas 12.9 folded it
// into a 32-bit uniform ULEA with a zeroed high half in one build of this kernel)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *base, uint64_t offset_bytes, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .u64 a;\n"
        "add.u64 a, %1, %4;\n"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [a], %2, [%3];\n"
        "}\n" ::"r"(dst),
        "l"(base), "r"(bytes), "r"(bar), "l"(offset_bytes)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// Named barriers. The interpreter warps (threads [0, T_i)) synchronise among themselves on barrier 1; the noise
// producer warps (threads [T_i, blockDim)) on barrier 6; event buffers are handed over with the arrive/sync pairs
// FULL(b) = 2 + b (producers arrive, interpreter waits) and FREE(b) = 4 + b (interpreter arrives, producers wait).
#define GSTIM_BAR_INTERP 1
#define GSTIM_BAR_FULL 2
#define GSTIM_BAR_FREE 4
#define GSTIM_BAR_PRODUCERS 6
// (barrier ids are immediates: with a register id ptxas reserves all 16 barriers and cannot pick the fast encoding)
template <int ID>
__device__ __forceinline__ void bar_sync(uint32_t count) {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_arrive(uint32_t count) {
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_sync2(uint32_t b, uint32_t count) {  // barrier ID + (b & 1)
    if (b & 1u) {
        bar_sync<ID + 1>(count);
    } else {
        bar_sync<ID>(count);
    }
}
template <int ID>
__device__ __forceinline__ void bar_arrive2(uint32_t b, uint32_t count) {
    if (b & 1u) {
        bar_arrive<ID + 1>(count);
    } else {
        bar_arrive<ID>(count);
    }
}

__device__ __forceinline__ uint4 xor4(uint4 a, uint4 b) {
    return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w);
}
__device__ __forceinline__ uint4 and4(uint4 a, uint32_t m) {
    return make_uint4(a.x & m, a.y & m, a.z & m, a.w & m);
}
__device__ __forceinline__ uint32_t bitmask(uint32_t aux, int bit) {
    return (uint32_t)0 - ((aux >> bit) & 1u);
}

// ------------------------------------------------------------------------------------------------
// Per-block context. Lives in shared memory so the opcode functions read what they need with broadcast
// loads instead of carrying it in registers.
// ------------------------------------------------------------------------------------------------
struct __align__(16) BlockCtx {
    uint32_t X_s, Z_s;        // shared-window byte addresses of the frame planes: plane[k][row] as uint4
    uint32_t flag_s;          // correlated-error flag row (K uint4)
    uint32_t lt_s;            // log2 table (512 u32)
    uint32_t needs_s;         // per rate class: min(B * rate, 2^63) (32 u64), followed by the 32 rates themselves
    uint32_t pitch_b;         // bytes between consecutive columns of a plane (q_pitch * 16)
    uint32_t K, B, G_log2, slots;
    uint32_t k0, k1;          // Philox key
    uint32_t col0_lo, col0_hi;  // global column index of this block's first column
    uint32_t rec_mask, pad0;
    uint4 *rec;               // this block's record rows
    uint4 *out;               // this block's output columns
    uint64_t rec_k_stride, out_k_stride;  // uint4 units between the 128-shot columns of a record / output row (rows are contiguous)
    const uint32_t *ev_segoff;  // event segment offsets per noise batch
    uint32_t *ev_counts;        // this CTA's event counters
    uint32_t *ev_buf;           // this CTA's event records
    // noise schedule (read by the pre-pass)
    const uint4 *slices;
    const ulonglong2 *rates;
    const uint32_t *noise_info, *prog;
    uint32_t *ev_overflow;
    uint32_t n_slices, pad5;
    uint32_t n_noise, T_i;             // T_i: interpreter threads (the remaining warps of the block produce noise events)
    uint32_t next_s;                   // shared counter the pre-pass threads claim chains from
    uint32_t ev_counts_s, ev_segoff_s; // shared-window addresses of the event counters / segment offsets (0: global)
    uint32_t stage_s;                  // 2 x GSTIM_EV_STAGE event records prefetched for the current / next noise batch
    unsigned long long *dbg;  // optional cycle counters (block 0 only)
    uint32_t dbg_flags, pad2;
    // launch-wide constants of the two role loops (interp_role / producer_role)
    uint32_t n_blocks, n_chunks, chunk_words, T_all;
    uint32_t mbar_s, ev_s, ev_total, pad4;  // ev_s: shared-window address of the event counters (two buffers), 0 if in global memory
    uint32_t *ring;
    uint64_t col0_base;
    uint4 *rec_base, *out_base;
    uint64_t rec_block_stride, rec_cta_stride;
    uint32_t *ev_counts_g, *ev_buf_g;
};

size_t interp_smem_bytes(uint32_t q_pitch, uint32_t Q, uint32_t K, uint32_t chunk_words, uint32_t n_noise) {
    (void)Q;
    size_t b = 0;
    b += (size_t)2 * K * q_pitch * 16;  // X, Z
    b += (size_t)K * 16;                // correlated-error flag row
    b += (size_t)2 * chunk_words * 4;   // program ring
    b += 512 * 4;                       // log2 table
    b += 128 * 8;                       // the first GSTIM_RATE_SMEM_MAX rates: lam, floor((2^64 - 1) / lam)
    if (n_noise <= GSTIM_EV_SMEM_MAX) {
        b += ((size_t)(3 * n_noise + 1) * 4 + 15) / 16 * 16;  // event counters (two buffers) + segment offsets
    }
    b += 2 * GSTIM_EV_STAGE * 4;        // staged event records of the current / next noise batch
    b += 32;                            // mbarriers
    b += (sizeof(BlockCtx) + 15) / 16 * 16;
    return b;
}

// Exp(1) variate from a uniform u32 in fixed point (unit 2^-56 nat): -ln((r + 1/2) / 2^32) through a
// 256-entry log2 table with linear interpolation (max error 2e-6 nat). Integer-only, so the oracle
// (oracle/philox.py: exp_draw_fx) reproduces it bit for bit. lt_s: table in shared memory: base[256], diff[256].
__device__ __forceinline__ unsigned long long exp_draw_fx(uint32_t r, uint32_t lt_s) {
    const unsigned long long v = 2ull * r + 1ull;     // odd, < 2^33
    const int t = 63 - __clzll((long long)v);         // floor(log2 v), 0..32
    const uint32_t frac = (uint32_t)(v << (32 - t));  // bits below the leading one, left aligned
    const uint32_t i = frac >> 24, f = frac & 0xFFFFFFu;
    const unsigned long long log2m = (unsigned long long)lds32(lt_s + 4 * i) + (((unsigned long long)lds32(lt_s + 1024 + 4 * i) * f) >> 24);
    const unsigned long long lv = ((unsigned long long)t << 32) + log2m;
    return ((33ull << 32) - lv) * (unsigned long long)GSTIM_LN2_Q24;
}

__device__ __forceinline__ unsigned long long sat_mul(uint32_t n, unsigned long long lam) {
    // min(n * lam, 2^63)
    const unsigned long long lo = (unsigned long long)n * lam, hi = __umul64hi((unsigned long long)n, lam);
    return (hi != 0 || lo >= (1ull << 63)) ? (1ull << 63) : lo;
}

__device__ __forceinline__ void flip_plane(const BlockCtx *bc, uint32_t plane_s, uint32_t row, uint32_t shot) {
    const uint32_t a = plane_s + (shot >> 7) * bc->pitch_b + row * 16 + ((shot >> 5) & 3) * 4;
    sts32(a, lds32(a) ^ (1u << (shot & 31)));
}
__device__ __forceinline__ void flip_rec(const BlockCtx *bc, uint32_t rec_index, uint32_t shot) {
    uint32_t *w = (uint32_t *)(bc->rec + (uint64_t)(shot >> 7) * bc->rec_k_stride + (rec_index & bc->rec_mask)) + ((shot >> 5) & 3);
    *w ^= 1u << (shot & 31);
}

// ------------------------------------------------------------------------------------------------
// Noise event pre-pass. Event positions and Pauli choices never depend on the frame, so before a shot block
// is interpreted every noise site of the whole program is sampled up front. The sites of a noise group are
// cut into slices of GSTIM_NOISE_SLICE sites (program.h "Noise schedule"); a slice x this shot block is one
// Bernoulli sequence with its own Philox stream. Threads claim slices from a shared counter and walk them with
// geometric gaps: every loop iteration is one Philox call -> two draws (gap to the slice's next event, that
// event's Pauli word), so all lanes of a warp do the same work each iteration and idle only at the very end.
// Lanes of a warp work on neighbouring slices, i.e. mostly on the same noise batch: a record takes its place
// with one shared-memory atomic on the batch's counter, and the places of neighbouring lanes are mostly
// consecutive, so the stores into this CTA's (L2-resident) scratch coalesce; one segment per noise batch.
// The interpreter then only applies flips.
//
// Distribution == RareErrorIterator (/root/reference/src/stim/util_bot/probability_util.cc:33-43):
// gaps are floor(Exp(1)/lambda) = Geometric(p), in exact integer arithmetic (unit 2^-56 nat).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t atom_add_shared(uint32_t a, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}

// floor(E / lam) with inv = floor((2^64 - 1) / lam): the multiply-high estimate is never above and at most one below.
__device__ __forceinline__ unsigned long long div_by_rate(unsigned long long E, unsigned long long lam, unsigned long long inv) {
    unsigned long long q = __umul64hi(E, inv);
    if (E - q * lam >= lam) {
        q++;
    }
    return q;
}

// Per-run arguments of the producers (the interpreter is working on another shot block at the same time).
struct PrepassRun {
    uint32_t col0_lo, col0_hi;  // first global column of the shot block
    uint32_t cnt_s;             // shared-window address of this buffer's event counters (0: use `counts`)
    uint32_t *counts;
    uint32_t *evbuf;
    uint32_t tid, threads;      // this thread's index among the threads that walk the slices, and how many there are
    uint32_t whole_block;       // 1: every thread of the block takes part (first shot block of a launch), barrier 0
};

__device__ __forceinline__ void prepass_sync(const PrepassRun &run) {
    if (run.whole_block) {
        __syncthreads();
    } else {
        bar_sync<GSTIM_BAR_PRODUCERS>(run.threads);
    }
}

__device__ __noinline__ void noise_prepass(const BlockCtx *bc, const PrepassRun run) {
    const uint32_t T_p = run.threads;
    const uint32_t tid = run.tid;
    const uint32_t next_s = bc->next_s;
    const uint32_t n_noise = bc->n_noise;
    const uint32_t cnt_s = run.cnt_s, segoff_s = bc->ev_segoff_s;
    uint32_t *evbuf = run.evbuf;
    // (rarely used pointers are re-read from the BlockCtx where they are needed: registers are what limits this loop)
    for (uint32_t i = tid; i < n_noise; i += T_p) {
        if (cnt_s) {
            sts32(cnt_s + 4 * i, 0);
        } else {
            run.counts[i] = 0;
        }
    }
    if (tid == 0) {
        sts32(next_s, 0);
    }
    prepass_sync(run);

    const uint4 *slices = bc->slices;
    const uint32_t n_slices = (bc->dbg_flags & 1u) ? 0u : bc->n_slices;
    const uint32_t B = bc->B, lt_s = bc->lt_s, rates_s = bc->needs_s;
    const uint32_t magicB = 0xFFFFFFFFu / B + 1;  // floor(a / B) == umulhi(a, magicB) for a < 2^20 (B <= 4096)
    const uint32_t k0 = bc->k0, k1 = bc->k1, col0_lo = run.col0_lo, col0_hi = run.col0_hi;

    const bool dbg_on = tid == 0 && !(bc->dbg_flags & 1u) && bc->dbg != nullptr;
    const uint32_t tA = (uint32_t)clock64();
    uint32_t n_ev = 0, n_iter = 0, n_sl = 0;

    uint32_t grp = 0, sl1 = 0, nbi = 0, item0 = 0, total = 0, a = 0, d = 0;
    uint4 bh = make_uint4(0, 0, 0, 0);  // of the slice's batch: op | flags << 8 | aux << 16, T1, T2, T3
    unsigned long long lam = 1, inv = 0;
    // the next slice is claimed and its descriptor fetched while the current one is walked
    uint32_t nidx = atom_add_shared(next_s, 1);
    uint4 nd = make_uint4(0, 0, 0, 0), nh = nd;
    if (nidx < n_slices) {
        nd = __ldg(slices + 2 * (size_t)nidx);
        nh = __ldg(slices + 2 * (size_t)nidx + 1);
    }
    bool have = false;
    // An event record is written one event late: two consecutive events of a slice that flip the same 32-bit
    // frame words (same item, same 32-shot word) both get GSTIM_EV_CONFLICT, and only those are applied with
    // shared-memory atomics (2 cycles per lane on the LSU) by the interpreter; everything else is plain.
    bool pend = false;
    uint32_t pend_rec = 0;
    const uint32_t same_word_mask = (GSTIM_EV_ITEM_MASK << GSTIM_EV_ITEM_SHIFT) | 0xFE0u;

    // record of the event at shot-site `at` of the current slice, Pauli word y
    auto make_rec = [&](uint32_t at, uint32_t y) -> uint32_t {
        const uint32_t site = __umulhi(at, magicB), shot = at - site * B;
        const uint32_t h0 = bh.x;
        const uint32_t op = h0 & 0xFF, flags = (h0 >> 8) & 0xFF, aux = h0 >> 16;
        uint32_t f = 0;  // bit0 x1, bit1 z1, bit2 x2, bit3 z2, bit4 record row
        if (op == GOP_NOISE1) {
            const uint32_t sel = y < bh.y ? 0u : y < bh.z ? 2u : y < bh.w ? 4u : 6u;
            f = (aux >> sel) & 3u;
            if (flags & GF_REC) {
                f |= 16u;
            }
        } else if (op == GOP_NOISE2) {
            if (!(flags & GF_TABLE)) {
                f = 1u + __umulhi(y, 15u);  // uniform over the 15 non-identity pairs (frame_simulator.inl:651-659)
            } else {
                const uint32_t *tab = bc->prog + __ldg(bc->noise_info + (size_t)nbi * GSTIM_NOISE_INFO_WORDS + GNI_TABLE_OFF);
                uint32_t pr = aux;
                for (uint32_t t = 0; t < 15; t++) {
                    if (y < __ldg(tab + t)) {
                        pr = t + 1;
                        break;
                    }
                }
                // index = 4*P1 + P2 with P: 0=I 1=X 2=Y 3=Z (tableau_simulator.h:307-316)
                const uint32_t c1p = pr >> 2, c2p = pr & 3u;
                f = (((c1p + 1) >> 1) & 1u) | ((c1p >> 1) << 1) | ((((c2p + 1) >> 1) & 1u) << 2) | ((c2p >> 1) << 3);
            }
        }
        return shot | ((item0 + site) << GSTIM_EV_ITEM_SHIFT) | (f << GSTIM_EV_FLIP_SHIFT);
    };
    // append a record to its noise batch's segment (lanes of a warp work on neighbouring slices, so the places they
    // get are mostly consecutive and the stores coalesce)
    auto emit = [&](uint32_t rec) {
        uint32_t seg0, cap, at;
        if (cnt_s) {
            seg0 = lds32(segoff_s + 4 * nbi);
            cap = lds32(segoff_s + 4 * nbi + 4) - seg0;
            at = atom_add_shared(cnt_s + 4 * nbi, 1u);
        } else {
            seg0 = bc->ev_segoff[nbi];
            cap = bc->ev_segoff[nbi + 1] - seg0;
            at = atomicAdd(&run.counts[nbi], 1u);
        }
        if (at < cap) {  // (an overflowing segment is reported after the pre-pass)
            evbuf[seg0 + at] = rec;
        }
    };

    while (true) {
        bool live = true;
        if (!have) {
            if (nidx >= n_slices) {
                if (!pend) {
                    break;
                }
                live = false;
            } else {
                grp = nd.x;
                sl1 = nd.y | GSTIM_SLICE_FLAG;
                nbi = nd.z & 0xFFFFu;
                item0 = nd.w & 0x7FFu;
                total = (nd.w >> 11) * B;
                bh = nh;
                a = 0;
                d = 0;
                const uint32_t ri = nd.z >> 16;
                if (ri < GSTIM_RATE_SMEM_MAX) {
                    lam = lds64(rates_s + 16 * ri);
                    inv = lds64(rates_s + 16 * ri + 8);
                } else {
                    const ulonglong2 r = __ldg(bc->rates + ri);
                    lam = r.x;
                    inv = r.y;
                }
                have = true;
                n_sl++;
                nidx = atom_add_shared(next_s, 1);
                if (nidx < n_slices) {
                    nd = __ldg(slices + 2 * (size_t)nidx);
                    nh = __ldg(slices + 2 * (size_t)nidx + 1);
                }
            }
        }
        // records to write this iteration: the pending one and the first of up to two new events
        bool has0 = false, has1 = false;
        uint32_t out0 = 0, out1 = 0;
        if (live) {
            n_iter++;
            // one Philox call = two draws (gap word, Pauli word): their clock arithmetic is independent of the
            // position in the slice, so both are computed side by side and resolved in order afterwards
            const uint4 rr = philox4x32_10(grp, sl1, col0_lo, col0_hi | (d << GSTIM_DRAW_SHIFT), k0, k1);
            d++;
            const unsigned long long G0 = div_by_rate(exp_draw_fx(rr.x, lt_s), lam, inv);
            const unsigned long long G1 = div_by_rate(exp_draw_fx(rr.z, lt_s), lam, inv);
            has0 = pend;
            out0 = pend_rec;
            pend = false;
            if (G0 >= (unsigned long long)(total - a)) {  // no further event in this slice (the second draw is dropped)
                have = false;
            } else {
                a += (uint32_t)G0;
                uint32_t r0 = make_rec(a, rr.y);
                a++;
                if (has0 && ((r0 ^ out0) & same_word_mask) == 0) {
                    r0 |= GSTIM_EV_CONFLICT;
                    out0 |= GSTIM_EV_CONFLICT;
                }
                n_ev++;
                if (G1 >= (unsigned long long)(total - a)) {
                    have = false;
                    has1 = true;
                    out1 = r0;
                } else {
                    a += (uint32_t)G1;
                    uint32_t r1 = make_rec(a, rr.w);
                    a++;
                    if (((r1 ^ r0) & same_word_mask) == 0) {
                        r1 |= GSTIM_EV_CONFLICT;
                        r0 |= GSTIM_EV_CONFLICT;
                    }
                    n_ev++;
                    has1 = true;
                    out1 = r0;
                    pend = true;
                    pend_rec = r1;
                }
            }
        } else {
            has0 = true;
            out0 = pend_rec;
            pend = false;
        }
        if (has0) {
            emit(out0);
        }
        if (has1) {
            emit(out1);
        }
    }
    if (dbg_on) {
        bc->dbg[32] += (uint32_t)clock64() - tA;
        bc->dbg[33] += n_ev;
        bc->dbg[34] += n_sl;
        bc->dbg[35] += n_iter;
    }
    prepass_sync(run);
    // clamp the event counts to their segments (an overflow invalidates the call: the host reports it)
    for (uint32_t i = tid; i < n_noise; i += T_p) {
        const uint32_t cap = segoff_s ? lds32(segoff_s + 4 * i + 4) - lds32(segoff_s + 4 * i) : bc->ev_segoff[i + 1] - bc->ev_segoff[i];
        const uint32_t c = cnt_s ? lds32(cnt_s + 4 * i) : run.counts[i];
        if (c > cap) {
            if (cnt_s) {
                sts32(cnt_s + 4 * i, cap);
            } else {
                run.counts[i] = cap;
            }
            *bc->ev_overflow = 1u;
        }
    }
}

#define SLOT_SUB                                      \
    const uint32_t G_log2 = bc->G_log2;               \
    const uint32_t sub = threadIdx.x & ((1u << G_log2) - 1); \
    const uint32_t slot = threadIdx.x >> G_log2;      \
    const uint32_t slots = bc->slots;                 \
    const uint32_t K = bc->K;                         \
    const uint32_t pitch_b = bc->pitch_b;             \
    const uint32_t X_s = bc->X_s;                     \
    const uint32_t kstep = pitch_b << G_log2

// ------------------------------------------------------------------------------------------------
// opcodes
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ const uint32_t *op_cliff1(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    SLOT_SUB;
    const uint32_t aux = hw[GH_OP] >> 16, n = hw[GH_N];
    const SmemWords pay = hw + GSTIM_HDR_WORDS;
    const uint32_t a = bitmask(aux, 0), b = bitmask(aux, 1), cc = bitmask(aux, 2), d = bitmask(aux, 3);
    const uint32_t zoff = bc->Z_s - X_s;
    for (uint32_t i = slot; i < n; i += slots) {
        uint32_t ax = X_s + pay[i] * 16 + sub * pitch_b;
        for (uint32_t k = sub; k < K; k += 1u << G_log2, ax += kstep) {
            const uint4 x = lds128(ax), z = lds128(ax + zoff);
            sts128(ax, xor4(and4(x, a), and4(z, b)));
            sts128(ax + zoff, xor4(and4(x, cc), and4(z, d)));
        }
    }
    return hdr + hw[GH_WORDS];
}

__device__ __noinline__ const uint32_t *op_cx(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    // CX: z1 ^= z2 ; x2 ^= x1   (frame_simulator.inl:387-405)
    SLOT_SUB;
    const uint32_t n = hw[GH_N];
    const SmemWords pay = hw + GSTIM_HDR_WORDS;
    const uint32_t zoff = bc->Z_s - X_s;
    for (uint32_t i = slot; i < n; i += slots) {
        const uint32_t w = pay[i];
        uint32_t a1 = X_s + (w & 0xFFFF) * 16 + sub * pitch_b;
        uint32_t a2 = X_s + (w >> 16) * 16 + sub * pitch_b;
        for (uint32_t k = sub; k < K; k += 1u << G_log2, a1 += kstep, a2 += kstep) {
            const uint4 x1 = lds128(a1), z2 = lds128(a2 + zoff), z1 = lds128(a1 + zoff), x2 = lds128(a2);
            sts128(a1 + zoff, xor4(z1, z2));
            sts128(a2, xor4(x2, x1));
        }
    }
    return hdr + hw[GH_WORDS];
}

__device__ __noinline__ const uint32_t *op_cliff2(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    SLOT_SUB;
    const uint32_t aux = hw[GH_OP] >> 16, n = hw[GH_N];
    const SmemWords pay = hw + GSTIM_HDR_WORDS;
    const uint32_t zoff = bc->Z_s - X_s;
    uint32_t m[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        m[j] = bitmask(aux, j);
    }
    for (uint32_t i = slot; i < n; i += slots) {
        const uint32_t w = pay[i];
        uint32_t a1 = X_s + (w & 0xFFFF) * 16 + sub * pitch_b;
        uint32_t a2 = X_s + (w >> 16) * 16 + sub * pitch_b;
        for (uint32_t k = sub; k < K; k += 1u << G_log2, a1 += kstep, a2 += kstep) {
            const uint4 x1 = lds128(a1), z1 = lds128(a1 + zoff), x2 = lds128(a2), z2 = lds128(a2 + zoff);
            sts128(a1, xor4(xor4(and4(x1, m[0]), and4(z1, m[1])), xor4(and4(x2, m[2]), and4(z2, m[3]))));
            sts128(a1 + zoff, xor4(xor4(and4(x1, m[4]), and4(z1, m[5])), xor4(and4(x2, m[6]), and4(z2, m[7]))));
            sts128(a2, xor4(xor4(and4(x1, m[8]), and4(z1, m[9])), xor4(and4(x2, m[10]), and4(z2, m[11]))));
            sts128(a2 + zoff, xor4(xor4(and4(x1, m[12]), and4(z1, m[13])), xor4(and4(x2, m[14]), and4(z2, m[15]))));
        }
    }
    return hdr + hw[GH_WORDS];
}

// First record and (clamped) record count of noise batch `nbi` in this CTA's event scratch.
__device__ __forceinline__ void event_segment(const BlockCtx *bc, uint32_t nbi, uint32_t &seg0, uint32_t &cnt) {
    if (bc->ev_counts_s) {
        seg0 = lds32(bc->ev_segoff_s + 4 * nbi);
        cnt = lds32(bc->ev_counts_s + 4 * nbi);
    } else {
        seg0 = bc->ev_segoff[nbi];
        cnt = bc->ev_counts[nbi];
    }
}
// Event records of noise batch `nbi` -> staging buffer (nbi & 1), asynchronously (LDGSTS): thread t copies the
// records it will apply itself, so no barrier is needed between the copy and the use. One group per call.
__device__ __forceinline__ void prefetch_events(const BlockCtx *bc, uint32_t nbi) {
    if (nbi < bc->n_noise) {
        uint32_t seg0, cnt;
        event_segment(bc, nbi, seg0, cnt);
        cnt = min(cnt, GSTIM_EV_STAGE);
        const uint32_t *ev = bc->ev_buf + seg0;
        const uint32_t st = bc->stage_s + (nbi & 1u) * (GSTIM_EV_STAGE * 4);
        for (uint32_t e = threadIdx.x; e < cnt; e += bc->T_i) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(st + 4 * e), "l"(ev + e) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void flip_atomic(uint32_t a, uint32_t bit) {
    asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(a), "r"(bit) : "memory");
}

// NOISE1 / NOISE2: apply the events the pre-pass left for this batch. Any thread applies any event, so the
// batch is bracketed by block barriers. Two events of one batch touch the same 32-bit frame word only when the
// pre-pass marked both GSTIM_EV_CONFLICT; only those use shared-memory atomics.
__device__ __noinline__ const uint32_t *op_noise(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    const uint32_t hdr_s = hw.s;
    const uint32_t h0 = lds32(hdr_s + 4 * GH_OP);
    const uint32_t flags = (h0 >> 8) & 0xFF;
    const uint32_t items_s = hdr_s + 4 * GSTIM_HDR_WORDS + (((h0 & 0xFF) == GOP_NOISE2 && (flags & GF_TABLE)) ? 60u : 0u);
    const uint32_t nbi = lds32(hdr_s + 4 * GH_CSITE0), rec0 = lds32(hdr_s + 4 * GH_REC0);
    uint32_t seg0, cnt;
    event_segment(bc, nbi, seg0, cnt);
    const uint32_t X_s = bc->X_s, zoff = bc->Z_s - bc->X_s, pitch_b = bc->pitch_b;
    const uint32_t st = bc->stage_s + (nbi & 1u) * (GSTIM_EV_STAGE * 4);
    prefetch_events(bc, nbi + 1);
    const uint32_t T_i = bc->T_i;
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this batch's records have landed
    if (!(flags & GF_NOENTRY)) {
        bar_sync<GSTIM_BAR_INTERP>(T_i);  // (a directly preceding noise batch ended with this barrier)
    }
    for (uint32_t e = threadIdx.x; e < cnt; e += T_i) {
        const uint32_t rec = e < GSTIM_EV_STAGE ? lds32(st + 4 * e) : bc->ev_buf[seg0 + e];
        const uint32_t shot = rec & ((1u << GSTIM_EV_SHOT_BITS) - 1);
        const uint32_t item = (rec >> GSTIM_EV_ITEM_SHIFT) & GSTIM_EV_ITEM_MASK;
        const uint32_t f = rec >> GSTIM_EV_FLIP_SHIFT;
        const uint32_t w = (flags & GF_NOFRAME) ? 0u : lds32(items_s + 4 * item);
        const uint32_t bit = 1u << (shot & 31);
        const uint32_t a0 = X_s + (shot >> 7) * pitch_b + ((shot >> 5) & 3) * 4;
        const uint32_t a1 = a0 + (w & 0xFFFF) * 16, a2 = a0 + (w >> 16) * 16;
        if (rec & GSTIM_EV_CONFLICT) {
            if (f & 1u) {
                flip_atomic(a1, bit);
            }
            if (f & 2u) {
                flip_atomic(a1 + zoff, bit);
            }
            if (f & 4u) {
                flip_atomic(a2, bit);
            }
            if (f & 8u) {
                flip_atomic(a2 + zoff, bit);
            }
            if (f & 16u) {
                uint32_t *rw = (uint32_t *)(bc->rec + (uint64_t)(shot >> 7) * bc->rec_k_stride + ((rec0 + item) & bc->rec_mask)) + ((shot >> 5) & 3);
                atomicXor(rw, bit);
            }
        } else {
            // all loads first, then the stores: the (up to four) words are independent
            uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            if (f & 1u) {
                w0 = lds32(a1);
            }
            if (f & 2u) {
                w1 = lds32(a1 + zoff);
            }
            if (f & 4u) {
                w2 = lds32(a2);
            }
            if (f & 8u) {
                w3 = lds32(a2 + zoff);
            }
            if (f & 1u) {
                sts32(a1, w0 ^ bit);
            }
            if (f & 2u) {
                sts32(a1 + zoff, w1 ^ bit);
            }
            if (f & 4u) {
                sts32(a2, w2 ^ bit);
            }
            if (f & 8u) {
                sts32(a2 + zoff, w3 ^ bit);
            }
            if (f & 16u) {
                uint32_t *rw = (uint32_t *)(bc->rec + (uint64_t)(shot >> 7) * bc->rec_k_stride + ((rec0 + item) & bc->rec_mask)) + ((shot >> 5) & 3);
                *rw ^= bit;
            }
        }
    }
    bar_sync<GSTIM_BAR_INTERP>(T_i);
    return hdr + hw[GH_WORDS];
}


// MEASURE body for one (basis, kind): the per-column code has no data-dependent branches left.
template <uint32_t BASIS, uint32_t KIND>
__device__ __forceinline__ void measure_items(const BlockCtx *bc, const SmemWords hw) {
    SLOT_SUB;
    const uint32_t n = hw[GH_N];
    // GF_DET: three payload words per item (qubit word, detector row or NONE, record slot to XOR with): the detector
    // is written while the fresh result is still in registers (lowering.cc: fuse_detectors)
    const bool fused = KIND != GK_R && ((hw[GH_OP] >> 8) & GF_DET) != 0;
    const uint32_t stride = fused ? 3u : 1u;
    uint4 *const out = bc->out;
    const uint64_t oks = bc->out_k_stride;
    const SmemWords pay = hw + GSTIM_HDR_WORDS;
    const uint32_t mgroup = hw[GH_CSITE0], rec0 = hw[GH_REC0];
    const uint32_t zoff = bc->Z_s - X_s;
    const uint32_t k0 = bc->k0, k1 = bc->k1;
    const uint32_t col_lo = bc->col0_lo, col_hi = bc->col0_hi, tag_hi = GTAG_COLLAPSE ^ col_hi;
    uint4 *const rec = bc->rec;
    const uint32_t rec_mask = bc->rec_mask;
    const uint64_t rks = bc->rec_k_stride;
    const uint32_t kinc = 1u << G_log2;
    const bool carry_free = col_lo + K >= col_lo;  // the K columns of the block do not cross a 2^32 column boundary
    for (uint32_t i = slot; i < n; i += slots) {
        const uint32_t w = pay[stride * i];
        const uint32_t q = w & 0xFFFFu, lq = w >> 16;  // physical frame row | logical qubit (addresses the collapse draws)
        uint4 *rrow = rec + ((rec0 + i) & rec_mask) + (uint64_t)sub * rks;
        const uint32_t det = fused ? pay[3 * i + 1] : 0xFFFFFFFFu;
        const uint4 *orec = rec + (fused ? pay[3 * i + 2] : 0u) + (uint64_t)sub * rks;
        uint4 *drow = out + (det != 0xFFFFFFFFu ? det : 0u) + (uint64_t)sub * oks;
        uint32_t ax = X_s + q * 16 + sub * pitch_b;
        for (uint32_t k = sub; k < K; k += kinc, ax += kstep, rrow += (uint64_t)kinc * rks, orec += (uint64_t)kinc * rks, drow += (uint64_t)kinc * oks) {
            uint4 prev = make_uint4(0, 0, 0, 0);
            if (det != 0xFFFFFFFFu) {
                prev = ldg128(orec);  // (issued before the Philox chain so the L2 latency hides behind it)
            }
            uint32_t c2 = col_lo + k, c3 = tag_hi;
            if (!carry_free && c2 < col_lo) {
                c3 = GTAG_COLLAPSE ^ (col_hi + 1);
            }
            const uint4 rnd = philox4x32_10(mgroup, lq, c2, c3, k0, k1);
            const uint4 zero = make_uint4(0, 0, 0, 0);
            uint4 m, nx, nz;
            if (BASIS == GB_Z) {  // frame_simulator.inl:199-208, 266-274, 306-317
                const uint4 x = lds128(ax);
                m = x;
                nx = KIND == GK_M ? x : zero;
                nz = rnd;
            } else if (BASIS == GB_X) {  // :173-182, 211-219, 277-288
                const uint4 z = lds128(ax + zoff);
                m = z;
                nz = KIND == GK_M ? z : zero;
                nx = rnd;
            } else {  // Y basis :185-196, 255-263, 291-303
                m = xor4(lds128(ax), lds128(ax + zoff));
                nz = rnd;
                nx = KIND == GK_M ? xor4(m, rnd) : rnd;
            }
            if (!(BASIS == GB_Z && KIND == GK_M)) {  // (M in the Z basis leaves x as it is)
                sts128(ax, nx);
            }
            if (!(BASIS == GB_X && KIND == GK_M)) {
                sts128(ax + zoff, nz);
            }
            if (KIND != GK_R) {
                stg128(rrow, m);
                if (det != 0xFFFFFFFFu) {
                    stg128(drow, xor4(m, prev));
                }
            }
        }
    }
}

__device__ __noinline__ const uint32_t *op_measure(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    const uint32_t aux = hw[GH_OP] >> 16;
    switch (aux & 15u) {  // basis | kind << 2
        case GB_X | (GK_M << 2): measure_items<GB_X, GK_M>(bc, hw); break;
        case GB_Y | (GK_M << 2): measure_items<GB_Y, GK_M>(bc, hw); break;
        case GB_Z | (GK_M << 2): measure_items<GB_Z, GK_M>(bc, hw); break;
        case GB_X | (GK_MR << 2): measure_items<GB_X, GK_MR>(bc, hw); break;
        case GB_Y | (GK_MR << 2): measure_items<GB_Y, GK_MR>(bc, hw); break;
        case GB_Z | (GK_MR << 2): measure_items<GB_Z, GK_MR>(bc, hw); break;
        case GB_X | (GK_R << 2): measure_items<GB_X, GK_R>(bc, hw); break;
        case GB_Y | (GK_R << 2): measure_items<GB_Y, GK_R>(bc, hw); break;
        default: measure_items<GB_Z, GK_R>(bc, hw); break;
    }
    return hdr + hw[GH_WORDS];
}

__device__ __noinline__ const uint32_t *op_reczero(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    SLOT_SUB;
    (void)pitch_b;
    (void)kstep;
    (void)X_s;
    const uint32_t n = hw[GH_N], rec0 = hw[GH_REC0];
    uint4 *const rec = bc->rec;
    const uint32_t rec_mask = bc->rec_mask;
    const uint64_t rks = bc->rec_k_stride;
    for (uint32_t i = slot; i < n; i += slots) {
        uint4 *rrow = rec + ((rec0 + i) & rec_mask);
        for (uint32_t k = sub; k < K; k += 1u << G_log2) {
            stg128(rrow + k * rks, make_uint4(0, 0, 0, 0));
        }
    }
    return hdr + hw[GH_WORDS];
}

__device__ __noinline__ const uint32_t *op_xorrows(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    SLOT_SUB;
    (void)pitch_b;
    (void)kstep;
    (void)X_s;
    const uint32_t flags = (hw[GH_OP] >> 8) & 0xFF, n = hw[GH_N];
    const SmemWords pay = hw + GSTIM_HDR_WORDS;
    const SmemWords dst = pay, off = pay + n, idx = pay + 2 * n + 1;
    const uint4 *rec = bc->rec;
    uint4 *const out = bc->out;
    const uint64_t rks = bc->rec_k_stride, oks = bc->out_k_stride;
    // Record rows live in global memory (L2): a trip keeps up to XR_COLS columns x 2 rows of loads in flight per
    // thread instead of one dependent load at a time (6 columns per trip was slower: register pressure).
    constexpr int XR_COLS = 4;
    const uint32_t G = 1u << G_log2;
    for (uint32_t i = slot; i < n; i += slots) {
        const uint32_t b0 = off[i], b1 = off[i + 1];
        uint4 *orow = out + dst[i];
        for (uint32_t k0 = sub; k0 < K; k0 += XR_COLS * G) {
            uint4 acc[XR_COLS];
#pragma unroll
            for (int u = 0; u < XR_COLS; u++) {
                acc[u] = make_uint4(0, 0, 0, 0);
            }
            uint32_t j = b0;
            for (; j + 2 <= b1; j += 2) {
                const uint4 *r0 = rec + idx[j], *r1 = rec + idx[j + 1];
                uint4 v0[XR_COLS], v1[XR_COLS];
#pragma unroll
                for (int u = 0; u < XR_COLS; u++) {
                    const uint32_t k = k0 + u * G;
                    v0[u] = k < K ? ldg128(r0 + k * rks) : make_uint4(0, 0, 0, 0);
                    v1[u] = k < K ? ldg128(r1 + k * rks) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < XR_COLS; u++) {
                    acc[u] = xor4(acc[u], xor4(v0[u], v1[u]));
                }
            }
            if (j < b1) {
                const uint4 *r0 = rec + idx[j];
#pragma unroll
                for (int u = 0; u < XR_COLS; u++) {
                    const uint32_t k = k0 + u * G;
                    if (k < K) {
                        acc[u] = xor4(acc[u], ldg128(r0 + k * rks));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < XR_COLS; u++) {
                const uint32_t k = k0 + u * G;
                if (k < K) {
                    if (flags & GF_ACCUM) {
                        acc[u] = xor4(acc[u], ldg128(orow + k * oks));
                    }
                    stg128(orow + k * oks, acc[u]);
                }
            }
        }
    }
    return hdr + hw[GH_WORDS];
}

__device__ __noinline__ const uint32_t *op_obs_pauli(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    SLOT_SUB;
    const uint32_t n = hw[GH_N];
    const SmemWords pay = hw + GSTIM_HDR_WORDS;
    const uint32_t zoff = bc->Z_s - X_s;
    uint4 *const out = bc->out;
    const uint64_t oks = bc->out_k_stride;
    for (uint32_t i = slot; i < n; i += slots) {
        const uint32_t w = pay[2 * i + 1];
        uint4 *orow = out + pay[2 * i];
        uint32_t ax = X_s + (w & 0xFFFFFF) * 16 + sub * pitch_b;
        for (uint32_t k = sub; k < K; k += 1u << G_log2, ax += kstep) {
            uint4 acc = ldg128(orow + k * oks);
            if (w & (1u << 30)) {
                acc = xor4(acc, lds128(ax));
            }
            if (w & (1u << 31)) {
                acc = xor4(acc, lds128(ax + zoff));
            }
            stg128(orow + k * oks, acc);
        }
    }
    return hdr + hw[GH_WORDS];
}

__device__ __noinline__ const uint32_t *op_feedback(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    SLOT_SUB;
    const uint32_t n = hw[GH_N];
    const SmemWords pay = hw + GSTIM_HDR_WORDS;
    const uint32_t zoff = bc->Z_s - X_s;
    const uint4 *const rec = bc->rec;
    const uint64_t rks = bc->rec_k_stride;
    for (uint32_t i = slot; i < n; i += slots) {
        const uint32_t w = pay[2 * i + 1];
        const uint4 *rrow = rec + pay[2 * i];
        uint32_t ax = X_s + (w & 0xFFFFFF) * 16 + sub * pitch_b;
        for (uint32_t k = sub; k < K; k += 1u << G_log2, ax += kstep) {
            const uint4 r = ldg128(rrow + k * rks);
            if (w & (1u << 30)) {
                sts128(ax, xor4(lds128(ax), r));
            }
            if (w & (1u << 31)) {
                sts128(ax + zoff, xor4(lds128(ax + zoff), r));
            }
        }
    }
    return hdr + hw[GH_WORDS];
}

// E / ELSE_CORRELATED_ERROR (frame_simulator.inl:747-776): one site for the whole Pauli product, masked by
// (and recorded in) the block's "already occurred" row. Executed by a single thread from the pre-sampled events.
__device__ __noinline__ const uint32_t *op_corr(const BlockCtx *bc, const uint32_t *hdr) {
    const SmemWords hw{smem_u32(hdr)};
    prefetch_events(bc, hw[GH_CSITE0] + 1);  // keeps the staging pipeline of op_noise going (this op reads its records directly)
    // This batch's own staged records are never read, but their copy group must have landed before a later batch stages into
    // the same buffer: otherwise two in-flight async copies target the same addresses (compute-sanitizer racecheck,
    // profiles/r2_notes.md). Same wait as op_noise.
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    if (threadIdx.x != 0) {
        return hdr + hw[GH_WORDS];
    }
    const uint32_t flags = (hw[GH_OP] >> 8) & 0xFF, n = hw[GH_N];
    const SmemWords pay = hw + GSTIM_HDR_WORDS;
    const uint32_t flag_s = bc->flag_s;
    if (flags & GF_RESET_FLAG) {
        for (uint32_t k = 0; k < bc->K; k++) {
            sts128(flag_s + 16 * k, make_uint4(0, 0, 0, 0));
        }
    }
    const uint32_t nbi = hw[GH_CSITE0];
    uint32_t seg0, cnt;
    event_segment(bc, nbi, seg0, cnt);
    const uint32_t *ev = bc->ev_buf + seg0;
    for (uint32_t e = 0; e < cnt; e++) {
        const uint32_t shot = ev[e] & ((1u << GSTIM_EV_SHOT_BITS) - 1);
        const uint32_t fa = flag_s + (shot >> 7) * 16 + ((shot >> 5) & 3) * 4;
        const uint32_t bit = 1u << (shot & 31);
        const uint32_t fw = lds32(fa);
        if (!(fw & bit)) {
            sts32(fa, fw | bit);
            for (uint32_t j = 0; j < n; j++) {
                const uint32_t w = pay[j];
                if (w & (1u << 30)) {
                    flip_plane(bc, bc->X_s, w & 0xFFFFFF, shot);
                }
                if (w & (1u << 31)) {
                    flip_plane(bc, bc->Z_s, w & 0xFFFFFF, shot);
                }
            }
        }
    }
    return hdr + hw[GH_WORDS];
}

// ------------------------------------------------------------------------------------------------
// The two warp roles of a block. Everything they need lives in the BlockCtx, so each role keeps only its
// own loop state in registers (the opcode functions use the full register budget).
// ------------------------------------------------------------------------------------------------
// Producer warps: the noise events of shot block run r go to event buffer r & 1, one run ahead of the interpreter.
__device__ __noinline__ void producer_role(const BlockCtx *bc) {
    const uint32_t T_all = bc->T_all, n_noise = bc->n_noise, n_blocks = bc->n_blocks;
    // (the events of the launch's first shot block were produced by the whole block, see the kernel)
    uint32_t r = 1;
    for (uint32_t g = blockIdx.x + gridDim.x; g < n_blocks; g += gridDim.x, r++) {
        const uint32_t b = r & 1u;
        if (r >= 2) {
            bar_sync2<GSTIM_BAR_FREE>(b, T_all);  // the interpreter is done with run r - 2
        }
        const uint64_t col0 = bc->col0_base + (uint64_t)g * bc->K;
        PrepassRun run;
        run.col0_lo = (uint32_t)col0;
        run.col0_hi = (uint32_t)(col0 >> 32);
        run.cnt_s = bc->ev_s ? bc->ev_s + 4 * b * n_noise : 0u;
        run.counts = bc->ev_counts_g + ((size_t)blockIdx.x * 2 + b) * n_noise;
        run.evbuf = bc->ev_buf_g + ((size_t)blockIdx.x * 2 + b) * bc->ev_total;
        run.tid = threadIdx.x - bc->T_i;
        run.threads = T_all - bc->T_i;
        run.whole_block = 0;
        noise_prepass(bc, run);
        __threadfence_block();
        bar_arrive2<GSTIM_BAR_FULL>(b, T_all);
    }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int MAX_THREADS>
__global__ void __launch_bounds__(MAX_THREADS, 1) gstim_interp_kernel(const InterpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    const uint32_t X_s = smem_u32(sp);
    sp += (size_t)p.K * p.q_pitch * 16;
    const uint32_t Z_s = smem_u32(sp);
    sp += (size_t)p.K * p.q_pitch * 16;
    const uint32_t flag_s = smem_u32(sp);
    sp += (size_t)p.K * 16;
    uint32_t *ring = (uint32_t *)sp;
    sp += (size_t)2 * p.chunk_words * 4;
    uint32_t *lt = (uint32_t *)sp;
    sp += 512 * 4;
    const uint32_t needs_s = smem_u32(sp);
    sp += 128 * 8;
    uint32_t *ev_s = (uint32_t *)sp;  // [2][n_noise] counters, [n_noise + 1] segment offsets (when they fit)
    const bool ev_in_smem = p.n_noise <= GSTIM_EV_SMEM_MAX;
    if (ev_in_smem) {
        sp += ((size_t)(3 * p.n_noise + 1) * 4 + 15) / 16 * 16;
    }
    const uint32_t stage_s = smem_u32(sp);
    sp += 2 * GSTIM_EV_STAGE * 4;
    const uint32_t mbar_s = smem_u32(sp);
    sp += 32;
    BlockCtx *bc = (BlockCtx *)sp;

    const uint32_t tid = threadIdx.x;
    const uint32_t T = p.threads_interp;   // interpreter threads; the rest of the block produces noise events
    const uint32_t T_all = blockDim.x;

    if (tid == 0) {
        mbar_init(mbar_s, 1);
        mbar_init(mbar_s + 8, 1);
        bc->next_s = mbar_s + 16;
        bc->dbg_flags = p.dbg_flags;
        bc->dbg = blockIdx.x == 0 ? p.dbg_cycles : nullptr;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        bc->X_s = X_s;
        bc->Z_s = Z_s;
        bc->flag_s = flag_s;
        bc->lt_s = smem_u32(lt);
        bc->needs_s = needs_s;
        bc->ev_segoff = ev_in_smem ? ev_s + 2 * p.n_noise : p.ev_segoff;
        bc->T_i = T;
        bc->slices = p.slices;
        bc->rates = p.rates;
        bc->ev_counts_s = 0u;
        bc->ev_segoff_s = ev_in_smem ? smem_u32(ev_s + 2 * p.n_noise) : 0u;
        bc->n_slices = p.n_slices;
        bc->n_noise = p.n_noise;
        bc->stage_s = stage_s;
        bc->noise_info = p.noise_info;
        bc->prog = p.prog;
        bc->ev_overflow = p.ev_overflow;
        bc->pitch_b = p.q_pitch * 16;
        bc->K = p.K;
        bc->B = p.K * GSTIM_COL_SHOTS;
        bc->G_log2 = p.G_log2;
        bc->slots = p.slots;
        bc->k0 = p.seed_lo;
        bc->k1 = p.seed_hi;
        bc->rec_mask = p.rec_mask;
        bc->rec_k_stride = p.rec_k_stride;
        bc->out_k_stride = p.out_k_stride;
        bc->n_blocks = p.n_blocks;
        bc->n_chunks = p.n_chunks;
        bc->chunk_words = p.chunk_words;
        bc->T_all = T_all;
        bc->mbar_s = mbar_s;
        bc->ev_s = ev_in_smem ? smem_u32(ev_s) : 0u;
        bc->ev_total = p.ev_segoff[p.n_noise];
        bc->ring = ring;
        bc->col0_base = p.col0_base;
        bc->rec_base = p.rec;
        bc->out_base = p.out;
        bc->rec_block_stride = p.rec_block_stride;
        bc->rec_cta_stride = p.rec_cta_stride;
        bc->ev_counts_g = p.ev_counts;
        bc->ev_buf_g = p.ev_buf;
    }
    for (uint32_t i = tid; i < 256; i += T_all) {
        lt[i] = GSTIM_LOG2_BASE[i];
        lt[256 + i] = GSTIM_LOG2_DIFF[i];
    }
    for (uint32_t i = tid; i < min(p.n_rates, GSTIM_RATE_SMEM_MAX); i += T_all) {
        const ulonglong2 r = p.rates[i];
        sts64(needs_s + 16 * i, r.x);
        sts64(needs_s + 16 * i + 8, r.y);
    }
    if (ev_in_smem) {
        for (uint32_t i = tid; i <= p.n_noise; i += T_all) {
            ev_s[2 * p.n_noise + i] = p.ev_segoff[i];
        }
    }
    __syncthreads();

    const bool has_producers = T_all > T;
    if (has_producers) {
        // The first shot block's noise is sampled by the whole block (24 warps instead of 4): the interpreter would
        // only wait for it anyway. From the second block on the producer warps run ahead on their own.
        const uint64_t col0 = p.col0_base + (uint64_t)blockIdx.x * p.K;
        PrepassRun run;
        run.col0_lo = (uint32_t)col0;
        run.col0_hi = (uint32_t)(col0 >> 32);
        run.cnt_s = ev_in_smem ? smem_u32(ev_s) : 0u;
        run.counts = p.ev_counts + (size_t)blockIdx.x * 2 * p.n_noise;
        run.evbuf = p.ev_buf + (size_t)blockIdx.x * 2 * p.ev_segoff[p.n_noise];
        run.tid = tid;
        run.threads = T_all;
        run.whole_block = 1;
        noise_prepass(bc, run);
        __syncthreads();
    }
    if (tid >= T) {
        producer_role(bc);
        return;
    }

    // Interpreter warps: stream the program through the ring and execute it batch by batch. This loop lives in the kernel:
    // the state that is live across the opcode calls must stay within the few registers the callees leave alone,
    // because anything pushed to local memory is reloaded at L2 latency (the frame leaves almost no L1).
    // (kernel parameters are constant-bank operands: they cost no registers)
    const uint32_t chunk_words = p.chunk_words, chunk_bytes = chunk_words * 4;
    const bool multi = p.G_log2 != 0;
    const uint32_t skipmask = p.dbg_flags >> 8;  // debug: bit (op) set -> skip that opcode
    uint32_t phase0 = 0, phase1 = 0, run_idx = 0;
#ifdef GSTIM_CYCLE_COUNTERS
    // per-opcode cycle counters (GSTIM_DEBUG_CYCLES=1): compiled in only on request, because even the test of
    // `tid == 0` costs a local-memory reload per batch when tid does not survive the opcode calls in a register
    long long dbg_t0 = clock64();
#endif

    for (uint32_t g = blockIdx.x; g < p.n_blocks; g += gridDim.x, run_idx++) {
        const uint32_t eb = run_idx & 1u;
        if (tid == 0) {
            const uint64_t col0 = p.col0_base + (uint64_t)g * p.K;
            const uint32_t n_noise = p.n_noise;
            bc->col0_lo = (uint32_t)col0;
            bc->col0_hi = (uint32_t)(col0 >> 32);
            bc->ev_counts_s = bc->ev_s ? bc->ev_s + 4 * eb * n_noise : 0u;
            bc->ev_counts = bc->ev_s ? nullptr : p.ev_counts + ((size_t)blockIdx.x * 2 + eb) * n_noise;  // (global fallback)
            bc->ev_buf = p.ev_buf + ((size_t)blockIdx.x * 2 + eb) * bc->ev_total;
            bc->rec = p.rec + (uint64_t)g * p.rec_block_stride + (uint64_t)blockIdx.x * p.rec_cta_stride;
            bc->out = p.out + (uint64_t)g * p.K * p.out_k_stride;
            // start streaming the program
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(mbar_s, chunk_bytes);
            bulk_g2s(smem_u32(ring), p.prog, 0, chunk_bytes, mbar_s);
            if (p.n_chunks > 1) {
                mbar_expect_tx(mbar_s + 8, chunk_bytes);
                bulk_g2s(smem_u32(ring + chunk_words), p.prog, chunk_bytes, chunk_bytes, mbar_s + 8);
            }
        }
        for (uint32_t k = tid; k < p.K; k += T) {
            sts128(bc->flag_s + 16 * k, make_uint4(0, 0, 0, 0));
        }
        if (has_producers && run_idx > 0) {
            bar_sync2<GSTIM_BAR_FULL>(eb, blockDim.x);  // this run's noise events are in place (also orders the bc updates above)
        } else {
            bar_sync<GSTIM_BAR_INTERP>(T);
        }
        prefetch_events(bc, 0);

        for (uint32_t chunk = 0;; chunk++) {
            const uint32_t b = chunk & 1;
            {
                const uint32_t ph = b ? phase1 : phase0;
                while (!mbar_try_wait(mbar_s + 8 * b, ph)) {
                }
                if (b) {
                    phase1 ^= 1;
                } else {
                    phase0 ^= 1;
                }
            }
            const uint32_t *pw = ring + (size_t)b * chunk_words;
            bool end = false;
#ifdef GSTIM_CYCLE_COUNTERS
            if (p.dbg_cycles != nullptr && tid == 0 && blockIdx.x == 0) {
                const long long t1 = clock64();
                p.dbg_cycles[GOP_NEXT_CHUNK] += (unsigned long long)(t1 - dbg_t0);  // chunk hand-over + ring wait
                dbg_t0 = t1;
            }
#endif
            while (true) {
                const uint32_t h0 = lds32(smem_u32(pw) + 4 * GH_OP);
                const uint32_t op = h0 & 0xFF;
                if (op == GOP_END) {
                    end = true;
                    break;
                }
                if (op == GOP_NEXT_CHUNK) {
                    break;
                }
                if (h0 & (GF_BARRIER << 8)) {
                    bar_sync<GSTIM_BAR_INTERP>(T);
                } else if (multi) {
                    __syncwarp();
                }
                switch ((skipmask >> op) & 1u ? (uint32_t)GOP_QMAP : op) {
                    case GOP_CLIFF1:
                        pw = op_cliff1(bc, pw);
                        break;
                    case GOP_CLIFF2:
                        if ((h0 >> 16) == GSTIM_MAT_CX) {
                            pw = op_cx(bc, pw);
                        } else {
                            pw = op_cliff2(bc, pw);
                        }
                        break;
                    case GOP_NOISE1:
                    case GOP_NOISE2:
                        pw = op_noise(bc, pw);
                        break;
                    case GOP_MEASURE:
                        pw = op_measure(bc, pw);
                        break;
                    case GOP_RECZERO:
                        pw = op_reczero(bc, pw);
                        break;
                    case GOP_XORROWS:
                        pw = op_xorrows(bc, pw);
                        break;
                    case GOP_OBS_PAULI:
                        pw = op_obs_pauli(bc, pw);
                        break;
                    case GOP_FEEDBACK:
                        pw = op_feedback(bc, pw);
                        break;
                    case GOP_CORR:
                        pw = op_corr(bc, pw);
                        break;
                    default:  // GOP_QMAP and unknown words are skipped
                        pw += lds32(smem_u32(pw) + 4 * GH_WORDS);
                        break;
                }
#ifdef GSTIM_CYCLE_COUNTERS
                if (p.dbg_cycles != nullptr && tid == 0 && blockIdx.x == 0) {
                    const long long t1 = clock64();
                    p.dbg_cycles[op] += (unsigned long long)(t1 - dbg_t0);
                    p.dbg_cycles[16 + op] += 1;
                    dbg_t0 = t1;
                }
#endif
            }
            bar_sync<GSTIM_BAR_INTERP>(T);  // everyone is done reading ring[b]
            if (end) {
                break;
            }
            if (tid == 0 && chunk + 2 < p.n_chunks) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(mbar_s + 8 * b, chunk_bytes);
                bulk_g2s(smem_u32(ring + (size_t)b * chunk_words), p.prog, (uint64_t)(chunk + 2) * chunk_bytes, chunk_bytes,
                         mbar_s + 8 * b);
            }
        }
        if (has_producers && (uint64_t)g + 2ull * gridDim.x < bc->n_blocks) {
            bar_arrive2<GSTIM_BAR_FREE>(eb, blockDim.x);  // event buffer eb may be refilled (for run run_idx + 2)
        }
    }
}

template <int MT>
static cudaError_t set_attr(size_t smem) {
    return cudaFuncSetAttribute(gstim_interp_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t interp_set_max_smem(size_t smem) {
    cudaError_t e;
    if ((e = set_attr<256>(smem)) != cudaSuccess) {
        return e;
    }
    if ((e = set_attr<512>(smem)) != cudaSuccess) {
        return e;
    }
    if ((e = set_attr<768>(smem)) != cudaSuccess) {
        return e;
    }
    if ((e = set_attr<896>(smem)) != cudaSuccess) {
        return e;
    }
    return set_attr<1024>(smem);
}

cudaError_t launch_interp(const InterpParams &p, uint32_t grid, uint32_t threads, size_t smem, cudaStream_t stream) {
    // the register budget follows the block size: 255 / 128 / 80 / 72 / 64 registers per thread
    if (threads <= 256) {
        gstim_interp_kernel<256><<<grid, threads, smem, stream>>>(p);
    } else if (threads <= 512) {
        gstim_interp_kernel<512><<<grid, threads, smem, stream>>>(p);
    } else if (threads <= 768) {
        gstim_interp_kernel<768><<<grid, threads, smem, stream>>>(p);
    } else if (threads <= 896) {
        gstim_interp_kernel<896><<<grid, threads, smem, stream>>>(p);
    } else {
        gstim_interp_kernel<1024><<<grid, threads, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

int interp_max_blocks_per_sm(uint32_t threads, size_t smem) {
    int n = 0;
    cudaError_t e;
    if (threads <= 256) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gstim_interp_kernel<256>, (int)threads, smem);
    } else if (threads <= 512) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gstim_interp_kernel<512>, (int)threads, smem);
    } else if (threads <= 768) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gstim_interp_kernel<768>, (int)threads, smem);
    } else if (threads <= 896) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gstim_interp_kernel<896>, (int)threads, smem);
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gstim_interp_kernel<1024>, (int)threads, smem);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return n < 1 ? 1 : n;
}

}  // namespace gstim
