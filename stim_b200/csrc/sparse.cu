// sparse.cu — the event-driven engine: samples detection events from the propagated-response table (response.h).
//
// Replaces the same reference path as the interpreter (FrameSimulator::do_circuit over a batch of shots,
// /root/reference/src/stim/simulators/frame_simulator.inl:166-170, with the transposing writer behind it,
// /root/reference/src/stim/io/measure_record_writer.h:101-166) for circuits whose response table exists (response.h
// "ELIGIBLE"). Nothing is simulated per gate: a shot's output row is the XOR of the responses of the noise events that
// fired in it, so the work per shot is O(events), not O(gates).
//
//   * A thread block owns a TILE of S = 2^k consecutive shots and keeps the tile's output image — S dense shot-major
//     rows, byte for byte what the caller's b8 array holds for those shots — in shared memory.
//   * The (site x shot) Bernoulli trials of a tile are cut into SLICES (one class, a run of consecutive sites, all S shots;
//     site-major like the reference's RareErrorIterator over targets x shots, probability_util.cc:33-43). Threads claim
//     slices from a shared counter and walk them with geometric gaps: gap = floor(Exp(1) / lambda) in the 32-bit fixed
//     point of dem.cu (exactly Geometric(p) up to the 2^-26 nat resolution of the exponential).
//   * An event picks its outcome (the channel's Pauli) from the second word of its draw, loads the 16-byte table entry of
//     (site, outcome) — up to four output bit positions, more through an overflow list — and flips those bits of its
//     shot's row with red.shared.xor.
//   * When the slice pool is dry the block stores the image to global memory with 16-byte coalesced stores (the image is
//     kept at the same 16-byte phase as its destination, so rows of odd length — c3: 1951 B — need no shuffling) and
//     moves to its next tile. There is no bit-major table and no transposer pass: HBM sees the output bytes once.
//
// RNG addressing (restated bit for bit by oracle/sparse_oracle.py): Philox4x32-10, key = seed,
//     counter = (slice index, call, tile lo, 'SP' << 20 | tile hi),  tile = global shot index / S,
// call c of a slice -> draws 2c (words 0, 1) and 2c + 1 (words 2, 3); a draw = (gap word, outcome word). The draw
// that overshoots the slice ends it; a slice also ends when its last trial fired.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "sparse.cuh"
#ifndef GSTIM_SPARSE_TIMING  // 1: block 0 prints where its warps spent their cycles (profiling builds only)
#define GSTIM_SPARSE_TIMING 0
#endif

#define GSTIM_TABLE_QUAL __device__ const
#include "log2_q26_table.h"

namespace gstim {

namespace {

constexpr uint32_t SPARSE_TAG = 0x53500000u;  // 'SP' << 20
constexpr uint32_t MAX_CLASSES = 64;          // classes are mirrored in shared memory
constexpr uint32_t MAX_BUFFERS = 8;           // tile images per block
constexpr uint32_t SPARSE_THREADS = 1024;     // 31 producer warps + 1 writer warp
#ifndef GSTIM_PRODUCER_SLEEP_NS
#define GSTIM_PRODUCER_SLEEP_NS 32
#endif
constexpr uint32_t CTL_WORDS = 8;             // per buffer: next slice, lanes that left, full seq, ready seq, phase main, phase obs

enum : uint32_t { DK_SINGLE = 0, DK_UNIFORM = 1, DK_THRESH3 = 2, DK_THRESH_N = 3 };

struct EvClass {  // 96 bytes, shared memory
    uint32_t slice_end;  // cumulative slice count up to and including this class
    uint32_t slice0;     // first slice of the class
    uint32_t per;        // sites per slice
    uint32_t n_sites;
    uint32_t entry0;
    uint32_t inv, sh;
    uint32_t kind;       // DK_*
    uint32_t n_out;
    uint32_t thr[3];     // DK_THRESH3: thresholds - 1 (outcome = number of thr[j] < word); unused ones 0xFFFFFFFF
    uint32_t thr_off;    // DK_THRESH_N: word offset of the 15 thresholds in `thr_all`
    uint32_t dense_thr;  // != 0: packed Bernoulli words with P(bit) = dense_thr / 2^32 instead of geometric gaps
    uint32_t pad[2];
    // periodic tables (PERIODIC kernels): sites [per_a + per_p, per_a + per_n) are not stored; site s of that range uses the
    // entries of site per_a + (s - per_a) % per_p with per_delta * ((s - per_a) / per_p) added to its detector ids; the
    // sites behind the range follow the first period in the table. per_p = 0: the class is stored in full.
    uint32_t per_a, per_p, per_n, per_delta;
    uint32_t pad2[4];
};

struct SparseParams {
    const EvClass *classes;
    uint32_t n_classes;
    uint32_t n_slices;
    const uint32_t *thr_all;  // 15 thresholds per class (general chooser)
    const uint4 *entries;
    const uint32_t *overflow;
    const uint8_t *init_row;  // measurement mode: reference sample (main_bytes bytes) every row starts from, or null
    uint32_t log_s;
    uint32_t n_buffers;
    uint32_t n_tiles;
    uint64_t tile0;    // global index of tile 0 of this launch
    uint64_t n_shots;  // valid shots of this launch
    uint8_t *main_out;
    uint64_t main_pitch;
    uint32_t main_bytes;
    uint8_t *obs_out;
    uint64_t obs_pitch;
    uint32_t obs_bytes;
    uint32_t obs_img_off;  // byte offset of the observable image inside a tile image (multiple of 16)
    uint32_t img_bytes;    // bytes of one tile image, both parts (multiple of 16)
    uint32_t rk[20];       // Philox round keys: (k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85), r = 0..9
    uint32_t det_lo, n_det;  // PERIODIC: table values in [det_lo, det_lo + n_det) are detectors (they take the period shift)
};

__device__ __forceinline__ uint4 sp_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ rk[2 * r];  // (precomputed round keys: bumping the key with uniform adds measured 3 % slower)
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Hand-over flags between the producer warps and the writer warp (shared memory). They are read and written with
// atomics so that the accesses are ordered by the memory model itself (and compute-sanitizer's racecheck sees them as
// synchronisation, not as data races); the polls sit behind __nanosleep, so the extra atomic traffic is negligible.
__device__ __forceinline__ uint32_t ld_volatile_shared(uint32_t *p) {  // whole warp, converged: lane 0 polls
    uint32_t v = 0;
    if ((threadIdx.x & 31u) == 0) {
        v = atomicAdd(p, 0u);
    }
    return __shfl_sync(0xFFFFFFFFu, v, 0);
}
__device__ __forceinline__ void st_volatile_shared(uint32_t *p, uint32_t v) {
    atomicExch(p, v);
}

// Copies `nbytes` bytes from shared memory (image coordinate `phase`) to dst, where (dst & 15) == phase; one warp.
// The 16-byte aligned middle goes out as ONE bulk-async copy (TMA 1-D, cp.async.bulk shared -> global) issued by lane 0:
// the warp spends a handful of instructions per tile instead of a load/store loop; the <= 15 edge bytes on either side
// are byte stores. Returns after the bulk copy has finished READING shared memory (the image may be cleared).
__device__ __forceinline__ void store_span(const uint8_t *img, uint32_t img_saddr, uint32_t phase, uint8_t *dst, uint64_t nbytes, uint32_t lane) {
    const uint64_t end = phase + nbytes;
    uint8_t *const g0 = dst - phase;  // 16-byte aligned
    const uint64_t mid0 = phase ? 16 : 0, mid1 = end & ~15ull;
    if (mid1 > mid0) {
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g0 + mid0), "r"(img_saddr + (uint32_t)mid0),
                         "r"((uint32_t)(mid1 - mid0))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        for (uint64_t b = phase + lane; b < mid0 && b < end; b += 32) {
            g0[b] = img[b];
        }
        for (uint64_t b = mid1 + lane; b < end; b += 32) {
            g0[b] = img[b];
        }
        if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        for (uint64_t b = phase + lane; b < end; b += 32) {
            g0[b] = img[b];
        }
    }
    __syncwarp();
}

// One persistent block per SM. The block keeps n_buffers tile images in shared memory; tile sequence q of the block
// (global tile blockIdx.x + q * gridDim.x) lives in buffer q % n_buffers.
//   producer warps (all warps but the last) draw TICKETS from one counter of the block: ticket T is slice T % n_slices
//     of sequence T / n_slices, so a warp goes from slice to slice without visiting sequences it has no work in and
//     without waiting for anybody; it only waits when the buffer of its ticket's sequence has not been recycled yet.
//     A finished slice is counted on its sequence once its last flips are applied (one step later, see below);
//   the writer warp waits until all n_slices slices of sequence q are counted, stores the image to global memory,
//     clears it and hands the buffer to sequence q + n_buffers.
template <bool SEPARATE, bool PERIODIC>
__global__ void __launch_bounds__(SPARSE_THREADS, 1) gstim_sparse_kernel(const __grid_constant__ SparseParams p) {
    extern __shared__ uint4 smem4[];
    uint2 *const lt = reinterpret_cast<uint2 *>(smem4);                        // 256 x (base, diff): 2 KiB
    EvClass *const cls = reinterpret_cast<EvClass *>(lt + 256);                 // MAX_CLASSES x 64 B
    uint32_t *const ctl = reinterpret_cast<uint32_t *>(cls + MAX_CLASSES);      // MAX_BUFFERS x CTL_WORDS
    uint32_t *const sink = ctl + MAX_BUFFERS * CTL_WORDS;                       // 32 words nobody reads (see flip)
    uint8_t *const img0 = reinterpret_cast<uint8_t *>(sink + 32);               // 16-byte aligned
    const uint32_t img0_saddr = (uint32_t)__cvta_generic_to_shared(img0);

    const uint32_t S = 1u << p.log_s, smask = S - 1u, NB = p.n_buffers;
    const uint32_t main_bits = p.main_bytes * 8u, obs_bits = p.obs_bytes * 8u;
    const bool main_dense = p.main_pitch == p.main_bytes, obs_dense = p.obs_pitch == p.obs_bytes;
    const uint32_t n_seq = (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;  // tiles of this block
    const uint32_t n_prod = blockDim.x - 32;

    auto phases_of = [&](uint32_t q, uint32_t *pm, uint32_t *po) {
        const uint64_t shot0 = ((uint64_t)blockIdx.x + (uint64_t)q * gridDim.x) << p.log_s;
        *pm = (p.main_out != nullptr && main_dense) ? (uint32_t)((reinterpret_cast<uintptr_t>(p.main_out) + shot0 * p.main_bytes) & 15u) : 0u;
        *po = (SEPARATE && p.obs_out != nullptr && obs_dense) ? (uint32_t)((reinterpret_cast<uintptr_t>(p.obs_out) + shot0 * p.obs_bytes) & 15u) : 0u;
    };
    auto fill_init = [&](uint8_t *img, uint32_t pm, uint32_t tid, uint32_t nt) {
        for (uint32_t i = tid; i < S * p.main_bytes; i += nt) {
            img[pm + i] = p.init_row[i % p.main_bytes];
        }
    };

    // ---- set-up by the whole block -------------------------------------------------------------------------
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        lt[i] = make_uint2(GSTIM_LOG2_Q26[2 * i], GSTIM_LOG2_Q26[2 * i + 1]);
    }
    for (uint32_t i = threadIdx.x; i < p.n_classes * (uint32_t)(sizeof(EvClass) / 4); i += blockDim.x) {
        reinterpret_cast<uint32_t *>(cls)[i] = reinterpret_cast<const uint32_t *>(p.classes)[i];
    }
    for (uint32_t i = threadIdx.x; i < NB * (p.img_bytes / 16); i += blockDim.x) {
        reinterpret_cast<uint4 *>(img0)[i] = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x < NB) {
        uint32_t *c = ctl + threadIdx.x * CTL_WORDS;
        uint32_t pm, po;
        phases_of(threadIdx.x, &pm, &po);
        c[0] = 0;                                   // (unused)
        c[1] = 0;                                   // finished slices of the sequence
        c[6] = 0;                                   // buffer 0 only: next ticket of the block
        c[2] = 0;                                   // full: sequence + 1 whose events are all applied
        c[3] = threadIdx.x < n_seq ? threadIdx.x + 1 : 0;  // ready: sequence + 1 the buffer is cleared for
        c[4] = pm;
        c[5] = po;
    }
    __syncthreads();
    if (p.init_row != nullptr) {
        for (uint32_t b = 0; b < NB && b < n_seq; b++) {
            fill_init(img0 + b * p.img_bytes, ctl[b * CTL_WORDS + 4], threadIdx.x, blockDim.x);
        }
        __syncthreads();
    }

    if (threadIdx.x >= n_prod) {
        // ---- writer warp -------------------------------------------------------------------------------------
        const uint32_t lane = threadIdx.x & 31u;
#if GSTIM_SPARSE_TIMING
        long long t_wait = 0, t_store = 0, t_zero = 0, t_a = clock64(), t_b;
#endif
        for (uint32_t q = 0; q < n_seq; q++) {
            const uint32_t b = q % NB;
            uint32_t *c = ctl + b * CTL_WORDS;
            uint8_t *img = img0 + b * p.img_bytes;
            while (p.n_slices != 0 && ld_volatile_shared(c + 2) != q + 1) {
                __nanosleep(64);
            }
            __threadfence_block();
#if GSTIM_SPARSE_TIMING
            t_b = clock64(); t_wait += t_b - t_a; t_a = t_b;
#endif
            const uint64_t shot0 = ((uint64_t)blockIdx.x + (uint64_t)q * gridDim.x) << p.log_s;
            const uint32_t n_valid = (uint32_t)min((uint64_t)S, p.n_shots - shot0);
            const uint32_t pm = c[4], po = c[5];
            if (p.main_out != nullptr && p.main_bytes) {
                if (main_dense) {
                    store_span(img, img0_saddr + b * p.img_bytes, pm, p.main_out + shot0 * p.main_bytes, (uint64_t)n_valid * p.main_bytes, lane);
                } else {
                    for (uint32_t i = lane; i < n_valid * p.main_bytes; i += 32) {
                        const uint32_t row = i / p.main_bytes, col = i - row * p.main_bytes;
                        p.main_out[(shot0 + row) * p.main_pitch + col] = img[i];
                    }
                }
            }
            if (SEPARATE && p.obs_out != nullptr && p.obs_bytes) {
                const uint8_t *oimg = img + p.obs_img_off;
                if (obs_dense) {
                    store_span(oimg, img0_saddr + b * p.img_bytes + p.obs_img_off, po, p.obs_out + shot0 * p.obs_bytes, (uint64_t)n_valid * p.obs_bytes, lane);
                } else {
                    for (uint32_t i = lane; i < n_valid * p.obs_bytes; i += 32) {
                        const uint32_t row = i / p.obs_bytes, col = i - row * p.obs_bytes;
                        p.obs_out[(shot0 + row) * p.obs_pitch + col] = oimg[i];
                    }
                }
            }
#if GSTIM_SPARSE_TIMING
            t_b = clock64(); t_store += t_b - t_a; t_a = t_b;
#endif
            if (q + NB < n_seq) {  // recycle the buffer for sequence q + NB
                __syncwarp();
#pragma unroll 8
                for (uint32_t i = lane; i < p.img_bytes / 16; i += 32) {
                    reinterpret_cast<uint4 *>(img)[i] = make_uint4(0, 0, 0, 0);
                }
                uint32_t npm, npo;
                phases_of(q + NB, &npm, &npo);
                if (p.init_row != nullptr) {
                    __syncwarp();
                    fill_init(img, npm, lane, 32);
                }
                __syncwarp();
                if (lane == 0) {
                    c[1] = 0;
                    c[4] = npm;
                    c[5] = npo;
                    __threadfence_block();
                    st_volatile_shared(c + 3, q + NB + 1);
                }
                __syncwarp();
            }
#if GSTIM_SPARSE_TIMING
            t_b = clock64(); t_zero += t_b - t_a; t_a = t_b;
#endif
        }
#if GSTIM_SPARSE_TIMING
        if (blockIdx.x == 0 && lane == 0) {
            printf("writer: n_seq %u wait %lld store %lld zero %lld cycles per tile\n", n_seq, t_wait / n_seq, t_store / n_seq, t_zero / n_seq);
        }
#endif
        return;
    }

    // ---- producer warps -------------------------------------------------------------------------------------
    // A warp walks ONE slice at a time, 64 draws per step: lane l computes Philox call (step * 32 + l) of the slice =
    // draws 2l and 2l + 1 of the step, turns their gap words into gaps, and a warp prefix sum over (gap + 1) gives
    // every draw its position in the slice - the same positions the sequential walk visits, since draw d's event sits
    // at sum_{j <= d} (gap_j + 1) - 1. Draws whose position falls beyond the slice are dropped; the first of them ends
    // the slice. All 32 lanes execute the same instructions (no divergence), valid draws only differ in predicates.
    const uint32_t lane = threadIdx.x & 31u;
    if (p.n_slices == 0) {
        return;  // nothing random: the writer stores the constant rows
    }
    uint32_t q = 0, b = 0, q_end = p.n_slices;  // sequence of the warp's ticket, its buffer, first ticket of sequence q + 1
    uint32_t *c = ctl;
    uint32_t entered_q = 0xFFFFFFFFu;  // sequence whose buffer the warp has entered
    // the slice the warp finished last is counted on its sequence (c_owe, q_owe) once its pending flips are applied
    bool owe = false;
    uint32_t *c_owe = ctl;
    uint32_t q_owe = 0;
    uint32_t rowbase_m = 0, rowbase_o = 0, c2 = 0, c3 = 0;
    const uint32_t e0 = p.n_classes > 0 ? cls[0].slice_end : 0xFFFFFFFFu, e1 = p.n_classes > 1 ? cls[1].slice_end : 0xFFFFFFFFu,
                   e2 = p.n_classes > 2 ? cls[2].slice_end : 0xFFFFFFFFu;
    // the flips of a step's events are applied one step later, so their table entries have time to arrive
    uint4 pend0 = make_uint4(RESP_NONE, RESP_NONE, RESP_NONE, RESP_NONE), pend1 = pend0;
    uint32_t pm0 = 0, po0 = 0, pm1 = 0, po1 = 0;
    uint32_t ps0 = 0, ps1 = 0;  // PERIODIC: detector-id shift of the pending events

    const uint32_t sink_saddr = (uint32_t)__cvta_generic_to_shared(sink) + lane * 4u;
    // flips output bit v of a row unless v is an empty slot / overflow link (bit 31 set): predicated, no branch
    auto flip = [&](uint32_t row_m, uint32_t row_o, uint32_t v, uint32_t shift = 0) {
        uint32_t bit;
        if (SEPARATE) {
            bit = ((v & 0x40000000u) ? row_o : row_m) + (v & 0x3FFFFFFFu);
        } else {
            bit = row_m + v;
        }
        if (PERIODIC) {
            bit += (v - p.det_lo) < p.n_det ? shift : 0u;
        }
        if (!PERIODIC) {
            // No predicate: an empty slot flips a bit of the lane's own sink word instead. (ptxas turns a predicated shared
            // atomic into a branch around it and its address arithmetic: 8 x BSSY/BRA/BSYNC per step, 10 % of the stall
            // samples were branch resolution. The folded kernels keep the predicate: the extra selects spill there.)
            const uint32_t a = (v < 0x80000000u) ? img0_saddr + ((bit >> 3) & 0x1FFFFFFCu) : sink_saddr;
            asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(a), "r"(1u << (bit & 31u)) : "memory");
        } else {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.lt.u32 p, %2, 0x80000000;\n"
                "@p red.shared.xor.b32 [%0], %1;\n"
                "}" ::"r"(img0_saddr + ((bit >> 5) << 2)),
                "r"(1u << (bit & 31u)), "r"(v)
                : "memory");
        }
    };
    auto apply = [&](const uint4 &e, uint32_t row_m, uint32_t row_o, uint32_t shift = 0) {
        flip(row_m, row_o, e.x, shift);
        flip(row_m, row_o, e.y, shift);
        flip(row_m, row_o, e.z, shift);
        flip(row_m, row_o, e.w, shift);
        if (e.w != RESP_NONE && (e.w & RESP_OVERFLOW)) {
            const uint32_t *ov = p.overflow + (e.w & 0x7FFFFFFFu);
            const uint32_t cnt = __ldg(ov);
            for (uint32_t j = 1; j <= cnt; j++) {
                flip(row_m, row_o, __ldg(ov + j), shift);
            }
        }
    };
    // table entry of (site `cs` of the class, outcome o); PERIODIC: folds the site into the stored period, *shift = what to
    // add to its detector ids
    uint32_t c_entry0 = 0, c_ebase = 0, c_s0 = 0, c_a = 0, c_p = 0, c_n = 0, c_delta = 0;
    auto entry_of = [&](uint32_t site, uint32_t n_out, uint32_t o, uint32_t *shift) -> uint32_t {
        if (!PERIODIC) {
            return c_ebase + site * n_out + o;  // (site counts from the slice's first site)
        }
        uint32_t cs = c_s0 + site;
        {
            *shift = 0;
            if (c_p != 0 && cs >= c_a + c_p) {
                if (cs < c_a + c_n) {
                    const uint32_t k = (cs - c_a) / c_p;
                    cs -= k * c_p;
                    *shift = k * c_delta;
                } else {
                    cs -= c_n - c_p;
                }
            }
        }
        return c_entry0 + cs * n_out + o;
    };

    auto settle = [&]() {  // every flip of the owed slice has been issued by this warp: count the slice
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(c_owe + 1, 1u) + 1 == p.n_slices) {
                __threadfence_block();
                st_volatile_shared(c_owe + 2, q_owe + 1);
            }
        }
        owe = false;
    };
    auto flush = [&]() {
        apply(pend0, pm0, po0, ps0);
        apply(pend1, pm1, po1, ps1);
        pend0.x = pend0.y = pend0.z = pend0.w = RESP_NONE;
        pend1 = pend0;
        if (owe) {
            settle();
        }
    };

#if GSTIM_SPARSE_TIMING
    long long t_pwait = 0, t_p0 = clock64();
    uint32_t n_pwait = 0, n_slices_done = 0, n_steps = 0;
#endif
    for (;;) {
        uint32_t T = 0;
        if (lane == 0) {
            T = atomicAdd(ctl + 6, 1u);
        }
        T = __shfl_sync(0xFFFFFFFFu, T, 0);
        while (T >= q_end) {
            q++;
            q_end += p.n_slices;
            b = b + 1 == NB ? 0 : b + 1;
        }
        if (q >= n_seq) {
            flush();
            break;
        }
#if GSTIM_SPARSE_TIMING
        n_slices_done++;
#endif
        if (q != entered_q) {
            c = ctl + b * CTL_WORDS;
            if (ld_volatile_shared(c + 3) != q + 1) {
                // the buffer has not been recycled for sequence q yet. A warp never sleeps on debts: the sequence it owes
                // a slice to may be the one this buffer is waiting for.
                flush();
#if GSTIM_SPARSE_TIMING
                const long long tw = clock64();
#endif
                while (ld_volatile_shared(c + 3) != q + 1) {
                    __nanosleep(GSTIM_PRODUCER_SLEEP_NS);
                }
#if GSTIM_SPARSE_TIMING
                t_pwait += clock64() - tw;
                n_pwait++;
#endif
            }
            __threadfence_block();
            const uint64_t gt = p.tile0 + blockIdx.x + (uint64_t)q * gridDim.x;
            c2 = (uint32_t)gt;
            c3 = SPARSE_TAG | (uint32_t)(gt >> 32);
            uint32_t pm, po;
            phases_of(q, &pm, &po);  // (recomputed here rather than read from the writer's control words)
            rowbase_m = (b * p.img_bytes + pm) * 8u;
            rowbase_o = (b * p.img_bytes + p.obs_img_off + po) * 8u;
            entered_q = q;
        }
        const uint32_t sl = T - (q_end - p.n_slices);
        uint32_t k = (sl >= e0 ? 1u : 0u) + (sl >= e1 ? 1u : 0u) + (sl >= e2 ? 1u : 0u);
        if (k == 3) {  // more than three classes (detector error models: one per distinct probability): binary search
            uint32_t hi = p.n_classes - 1;
            while (k < hi) {
                const uint32_t mid = (k + hi) >> 1;
                if (sl >= cls[mid].slice_end) {
                    k = mid + 1;
                } else {
                    hi = mid;
                }
            }
        }
        const uint4 *cw = reinterpret_cast<const uint4 *>(cls + k);
        const uint4 w0 = cw[0], w1 = cw[1], w2 = cw[2];  // (slice_end, slice0, per, n_sites) (entry0, inv, sh, kind) (n_out, thr)
        const uint32_t s0 = (sl - w0.y) * w0.z;
        const uint32_t total = min(w0.z, w0.w - s0) << p.log_s;  // trials of the slice (<= 2^30)
        const uint32_t n_out = w2.x, inv = w1.y, sh = w1.z, kind = w1.w;
        const uint32_t t0 = w2.y, t1 = w2.z, t2 = w2.w, thr_off = cls[k].thr_off, dense_thr = cls[k].dense_thr;
        c_entry0 = w1.x;
        c_s0 = s0;
        c_ebase = w1.x + s0 * n_out;
        if (PERIODIC) {
            const uint4 w4 = cw[4];
            c_a = w4.x;
            c_p = w4.y;
            c_n = w4.z;
            c_delta = w4.w;
        }

        if (dense_thr != 0) {
            flush();  // (the dense path applies its flips at once: nothing may be pending across it)
            // Dense class: the trials of the slice as packed Bernoulli words, 32 trials per lane and pass. A word with
            // P(bit) = thr / 2^32 comes from the binary expansion of thr, lowest set bit first: acc = b_i ? acc | r_i :
            // acc & r_i halves the probability and adds b_i / 2 (the bit-sliced part of biased_randomize_bits,
            // probability_util.cc:74-132; a fair coin is ONE random word). Random word i of trial word j = word (i & 3) of
            // Philox call 8 j + (i >> 2); the outcome word of the event at trial t = word 0 of call 0x80000000 | t.
            const uint32_t i0 = (uint32_t)__ffs((int)dense_thr) - 1u, n_words = (total + 31u) >> 5;
            for (uint32_t j = lane; j < n_words; j += 32) {
                uint32_t acc = 0;
                uint4 r = make_uint4(0, 0, 0, 0);
                for (uint32_t i = i0; i < 32; i++) {
                    if (i == i0 || (i & 3u) == 0) {
                        r = sp_philox(sl, 8u * j + (i >> 2), c2, c3, p.rk);
                    }
                    const uint32_t ri = (i & 3u) == 0 ? r.x : (i & 3u) == 1 ? r.y : (i & 3u) == 2 ? r.z : r.w;
                    acc = ((dense_thr >> i) & 1u) ? (acc | ri) : (acc & ri);
                }
                const uint32_t left = total - 32u * j;
                if (left < 32u) {
                    acc &= (1u << left) - 1u;
                }
                while (acc) {
                    const uint32_t t = 32u * j + (uint32_t)__ffs((int)acc) - 1u;
                    acc &= acc - 1u;
                    const uint32_t site = t >> p.log_s, shot = t & smask;
                    uint32_t o = 0;
                    if (n_out > 1) {
                        const uint32_t pw = sp_philox(sl, 0x80000000u | t, c2, c3, p.rk).x;
                        if (kind == DK_UNIFORM) {
                            o = __umulhi(pw, n_out);
                        } else if (kind == DK_THRESH3) {
                            o = (pw > t0 ? 1u : 0u) + (pw > t1 ? 1u : 0u) + (pw > t2 ? 1u : 0u);
                        } else {
                            for (uint32_t jj = 0; jj + 1 < n_out; jj++) {
                                o += pw >= __ldg(p.thr_all + thr_off + jj) ? 1u : 0u;
                            }
                        }
                    }
                    uint32_t shift = 0;
                    const uint4 e = __ldg(p.entries + entry_of(site, n_out, o, &shift));
                    apply(e, rowbase_m + shot * main_bits, rowbase_o + shot * obs_bits, shift);
                }
            }
            c_owe = c;
            q_owe = q;
            settle();
            continue;
        }

        uint32_t a0 = 0;  // trials consumed by the previous steps
        for (uint32_t call0 = 0;; call0 += 32) {
            // Last step's events are applied before this step's loads take their place: after the scan in the plain kernels
            // (the entry loads get a whole scan more to arrive: long scoreboard 7 % instead of 14 %), at the top of the step
            // in the folded ones (frees their registers earlier: fewer spills).
            if (PERIODIC) {
                apply(pend0, pm0, po0, ps0);
                apply(pend1, pm1, po1, ps1);
                if (owe) {
                    settle();
                }
            }
            const uint4 r = sp_philox(sl, call0 + lane, c2, c3, p.rk);
            const uint32_t rem = total - a0;  // > 0
            // gaps of the lane's two draws, clamped to rem (a clamped gap is an overshoot)
            uint32_t G[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t v = (h ? r.z : r.x) | 1u;
                const uint32_t t = 31u - (uint32_t)__clz((int)v);
                const uint32_t frac = (v << (31u - t)) << 1;
                const uint2 en = lt[frac >> 24];
                const uint32_t log2v = (t << 26) + en.x + ((en.y * ((frac >> 11) & 0x1FFFu)) >> 13);
                const uint32_t E = __umulhi(0x80000000u - log2v, GSTIM_LN2_Q32);
                const unsigned long long g = ((unsigned long long)E * inv) >> sh;
                G[h] = (uint32_t)min(g, (unsigned long long)rem);
            }
            // saturating prefix sum of (gap + 1) over the 64 draws of the step (saturation at rem keeps 32 bits exact for
            // every draw that is still inside the slice)
            const uint32_t s_a = min(G[0] + 1u, rem), s_ab = min(s_a + G[1] + 1u, rem);
            uint32_t incl = s_ab;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (uint32_t)d) {
                    incl = min(incl + up, rem);
                }
            }
            uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
            if (lane == 0) {
                excl = 0;
            }
            const uint32_t consumed = __shfl_sync(0xFFFFFFFFu, incl, 31);
            // event of draw h sits at trial a0 + (trials before the draw) + gap; it exists iff that is < total
            const uint32_t off0 = excl + G[0], off1 = excl + s_a + G[1];
            const bool ok0 = off0 < rem, ok1 = ok0 && off1 < rem;
            if (!PERIODIC) {
                apply(pend0, pm0, po0, ps0);
                apply(pend1, pm1, po1, ps1);
                if (owe) {
                    settle();
                }
            }
            pend0.x = pend0.y = pend0.z = pend0.w = RESP_NONE;
            pend1 = pend0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const bool ok = h ? ok1 : ok0;
                const uint32_t pos = a0 + (h ? off1 : off0), pw = h ? r.w : r.y;
                const uint32_t site = pos >> p.log_s, shot = pos & smask;
                // outcome: uniform (mulhi) or the number of thresholds <= pw (t_j hold threshold - 1; unused ones 0xFFFFFFFF)
                uint32_t o = (pw > t0 ? 1u : 0u) + (pw > t1 ? 1u : 0u) + (pw > t2 ? 1u : 0u);
                if (kind == DK_UNIFORM) {
                    o = __umulhi(pw, n_out);
                }
                if (kind == DK_THRESH_N) {
                    o = 0;
                    for (uint32_t j = 0; j + 1 < n_out; j++) {
                        o += pw >= __ldg(p.thr_all + thr_off + j) ? 1u : 0u;
                    }
                }
                if (ok) {
                    uint32_t shift = 0;
                    const uint4 e = __ldg(p.entries + entry_of(site, n_out, o, &shift));
                    if (h) {
                        pend1 = e;
                        ps1 = shift;
                    } else {
                        pend0 = e;
                        ps0 = shift;
                    }
                }
                if (h) {
                    pm1 = rowbase_m + shot * main_bits;
                    po1 = rowbase_o + shot * obs_bits;
                } else {
                    pm0 = rowbase_m + shot * main_bits;
                    po0 = rowbase_o + shot * obs_bits;
                }
            }
            a0 += consumed;
#if GSTIM_SPARSE_TIMING
            n_steps++;
#endif
            if (a0 >= total) {  // (warp-uniform) some draw of this step overshot, or the last trial fired
                break;
            }
        }
        owe = true;  // (the first step of every slice has settled the previous debt)
        c_owe = c;
        q_owe = q;
    }
#if GSTIM_SPARSE_TIMING
    if (blockIdx.x == 0 && lane == 0) {
        printf("producer warp %2u: total %lld cycles, waiting for a buffer %lld (%u waits), %u slices, %u steps\n", threadIdx.x >> 5,
               clock64() - t_p0, t_pwait, n_pwait, n_slices_done, n_steps);
    }
#endif
    // (ticket beyond the last sequence: nothing pending, nothing owed)
}

// ------------------------------------------------------------------------------------------------
// flip counts of dense b8 rows: single[b] += number of shots with bit b set, pair[b] += bit b AND bit b + 1
// (the statistics the 5-sigma tests and the per-detector count allreduce use; the reference has no such kernel — it is
// what a caller computes from the b8 array).
// A thread owns one 32-bit column of the rows (bits 32 j .. 32 j + 31) over a run of shots and counts all 32 bits at once
// in bit-sliced ("vertical") counters: eight shots are reduced by seven carry-save adders (Harley-Seal: XOR3 + MAJ3, two LOP3
// each) to residues of weight 1, 2, 4 and ONE word of weight 8, which a ripple-carry add pushes into eight bit planes; every
// 255 groups the planes and residues are folded into the global counters (3.6 word-ops per shot and column instead of 24). Rows may start at
// any byte (c3: 1951 B pitch): the word is assembled from two aligned loads with a funnel shift.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void vadd(uint32_t (&c)[8], uint32_t w) {
    uint32_t carry = w;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t t = c[k] & carry;
        c[k] ^= carry;
        carry = t;
    }
}
// Carry-save state of one column: `ones`, `twos`, `fours` hold the count of the shots seen so far modulo 8 (bit-sliced), the
// planes c[k] count whole groups of eight (weight 8 << k).
struct VCount {
    uint32_t ones = 0, twos = 0, fours = 0;
    uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
// (h, l) = a + b + c per bit: the sum bit stays at this weight, the majority carries to the next one (two LOP3)
#define GSTIM_CSA(h, l, a, b, c)         \
    {                                    \
        const uint32_t u_ = (a) ^ (b);   \
        const uint32_t c_ = (c);         \
        h = ((a) & (b)) | (u_ & c_);     \
        l = u_ ^ c_;                     \
    }
// eight shots' words at once (Harley-Seal): seven carry-save adders leave one word of weight eight for the planes
__device__ __forceinline__ void vadd8(VCount &v, const uint32_t (&w)[8]) {
    uint32_t ta, tb, fa, fb, e;
    GSTIM_CSA(ta, v.ones, v.ones, w[0], w[1]);
    GSTIM_CSA(tb, v.ones, v.ones, w[2], w[3]);
    GSTIM_CSA(fa, v.twos, v.twos, ta, tb);
    GSTIM_CSA(ta, v.ones, v.ones, w[4], w[5]);
    GSTIM_CSA(tb, v.ones, v.ones, w[6], w[7]);
    GSTIM_CSA(fb, v.twos, v.twos, ta, tb);
    GSTIM_CSA(e, v.fours, v.fours, fa, fb);
    vadd(v.c, e);
}
__device__ __forceinline__ void vflush(VCount &v, unsigned long long *dst, uint32_t bit0, uint32_t n_bits) {
#pragma unroll 4
    for (uint32_t b = 0; b < 32; b++) {
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x |= ((v.c[k] >> b) & 1u) << k;
        }
        x = 8u * x + ((v.ones >> b) & 1u) + 2u * ((v.twos >> b) & 1u) + 4u * ((v.fours >> b) & 1u);
        if (x != 0 && bit0 + b < n_bits) {
            atomicAdd(dst + bit0 + b, (unsigned long long)x);
        }
    }
    v = VCount();
}

__global__ void __launch_bounds__(512) gstim_count_b8_kernel(const uint8_t *rows, uint64_t pitch, uint64_t n_shots, uint32_t n_bits,
                                                            unsigned long long *single, unsigned long long *pair) {
    const uint32_t n_words = (n_bits + 31) / 32, n_bytes = (n_bits + 7) / 8;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_words) {
        return;
    }
    // (narrow rows: a block holds blockDim.y runs of shots side by side, so that its threads are all in use)
    const uint64_t n_runs = (uint64_t)gridDim.y * blockDim.y, ry = (uint64_t)blockIdx.y * blockDim.y + threadIdx.y;
    const uint64_t per = (n_shots + n_runs - 1) / n_runs;
    const uint64_t s0 = min(n_shots, ry * per), s1 = min(n_shots, s0 + per);
    const uint32_t valid = n_bits - 32 * j >= 32 ? 0xFFFFFFFFu : ((1u << (n_bits - 32 * j)) - 1u);
    const uint32_t bytes_here = min(4u, n_bytes - 4 * j);        // bytes of this column that exist in a row
    const bool has_next = 4 * j + 4 < n_bytes;
    VCount c1, c2;
    uint32_t pending = 0;  // groups of eight shots in the planes
    // A column with all four bytes inside the row (every thread but the one on the ragged end) is read branch-free: two aligned
    // words and a funnel shift; the first bit of the next column (for the pair counts) sits in the second word. Without
    // branches the compiler issues the sixteen loads of a group back to back — with a branch per row they went out one row at
    // a time and the kernel ran at one DRAM round trip per row (0.9 TB/s).
    const uint32_t next_mask = (pair != nullptr && has_next) ? 1u : 0u;
    auto load_full = [&](uint64_t s, uint32_t *w, uint32_t *nxt) {
        const uintptr_t ai = reinterpret_cast<uintptr_t>(rows + s * pitch + 4 * j);
        const uint32_t *al = reinterpret_cast<const uint32_t *>(ai & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(ai & 3) * 8;
        const uint32_t lo = __ldg(al), hi = __ldg(al + 1);  // (the rows' buffer has 16 spare bytes behind it)
        *w = __funnelshift_r(lo, hi, sh) & valid;
        *nxt = (hi >> sh) & next_mask;
    };
    auto load_ragged = [&](uint64_t s, uint32_t *w, uint32_t *nxt) {
        const uint8_t *a = rows + s * pitch + 4 * j;
        *w = 0;
        for (uint32_t k = 0; k < bytes_here; k++) {
            *w |= (uint32_t)a[k] << (8 * k);
        }
        *w &= valid;
        *nxt = 0;  // (nothing follows the last column)
    };
    auto add8 = [&](const uint32_t (&w)[8], const uint32_t (&nx)[8]) {
        vadd8(c1, w);
        if (pair != nullptr) {
            uint32_t pw[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                pw[u] = w[u] & ((w[u] >> 1) | (nx[u] << 31));
            }
            vadd8(c2, pw);
        }
        if (++pending == 255) {
            vflush(c1, single, 32 * j, n_bits);
            if (pair != nullptr) {
                vflush(c2, pair, 32 * j, n_bits - 1);
            }
            pending = 0;
        }
    };
    const bool full = bytes_here == 4;
    uint64_t s = s0;
    for (; s + 8 <= s1; s += 8) {  // eight rows in flight per thread (load latency), counted together (carry-save)
        uint32_t w[8], nx[8];
        if (full) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                load_full(s + u, &w[u], &nx[u]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                load_ragged(s + u, &w[u], &nx[u]);
            }
        }
        add8(w, nx);
    }
    if (s < s1) {  // the last, partial group: missing shots count as zero words
        uint32_t w[8], nx[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            w[u] = nx[u] = 0;
            if (s + u < s1) {
                if (full) {
                    load_full(s + u, &w[u], &nx[u]);
                } else {
                    load_ragged(s + u, &w[u], &nx[u]);
                }
            }
        }
        add8(w, nx);
    }
    vflush(c1, single, 32 * j, n_bits);  // (the residues may hold counts even when no group reached the planes)
    if (pair != nullptr) {
        vflush(c2, pair, 32 * j, n_bits - 1);
    }
}

// ------------------------------------------------------------------------------------------------
// Sparse host delivery: detection-event rows are mostly zero bytes (c3: 88 %), and the D2H copy of dense rows is what
// bounds the end-to-end rate. This kernel rewrites dense rows as per-shot records
//     u16 count, count x u16 byte offset, count x u8 value          (2-byte aligned, at index[shot] in the stream)
// so that only the non-zero bytes cross PCIe; host threads rebuild the caller's dense rows (api.cu). One warp per shot:
// pass A counts the non-zero bytes (ballots), lane 0 reserves the record in the stream with one atomic, pass B writes it.
// A stream that would overflow its buffer sets *overflow and the chunk is delivered densely instead.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gstim_compress_rows_kernel(const uint8_t *rows, uint64_t pitch, uint32_t row_bytes, uint64_t n_shots,
                                                                uint8_t *stream, uint64_t capacity, unsigned long long *cursor,
                                                                unsigned long long *index, uint32_t *overflow) {
    const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t shot = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; shot < n_shots; shot += warps) {
        const uint8_t *row = rows + shot * pitch;
        uint32_t cnt = 0;
        for (uint32_t c0 = 0; c0 < row_bytes; c0 += 32) {
            const uint32_t c = c0 + lane;
            const uint32_t b = c < row_bytes ? row[c] : 0u;
            cnt += __popc(__ballot_sync(0xFFFFFFFFu, b != 0));
        }
        unsigned long long o = 0;
        if (lane == 0) {
            o = atomicAdd(cursor, (unsigned long long)((2u + 3u * cnt + 1u) & ~1u));
            index[shot] = o;
        }
        o = __shfl_sync(0xFFFFFFFFu, o, 0);
        if (o + 2 + 3ull * cnt > capacity) {
            if (lane == 0) {
                *overflow = 1;
            }
            continue;
        }
        if (lane == 0) {
            *reinterpret_cast<uint16_t *>(stream + o) = (uint16_t)cnt;
        }
        uint16_t *offs = reinterpret_cast<uint16_t *>(stream + o + 2);
        uint8_t *vals = stream + o + 2 + 2ull * cnt;
        uint32_t base = 0;
        for (uint32_t c0 = 0; c0 < row_bytes; c0 += 32) {
            const uint32_t c = c0 + lane;
            const uint32_t b = c < row_bytes ? row[c] : 0u;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, b != 0);
            if (b != 0) {
                const uint32_t k = base + __popc(m & lt);
                offs[k] = (uint16_t)c;
                vals[k] = (uint8_t)b;
            }
            base += __popc(m);
        }
    }
}

void ck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
        }
        throw std::runtime_error(std::string("CUDA error '") + cudaGetErrorString(e) + "' in " + what);
    }
}

uint32_t align16(uint32_t v) {
    return (v + 15u) & ~15u;
}

constexpr size_t FIXED_SMEM = 256 * 8 + MAX_CLASSES * sizeof(EvClass) + MAX_BUFFERS * CTL_WORDS * 4 + 32 * 4;
static_assert(sizeof(EvClass) == 96 && FIXED_SMEM % 16 == 0, "shared memory layout");

}  // namespace

cudaError_t launch_count_b8(const uint8_t *rows, uint64_t pitch, uint64_t n_shots, uint32_t n_bits, unsigned long long *single,
                            unsigned long long *pair, cudaStream_t stream) {
    if (n_bits == 0 || n_shots == 0) {
        return cudaSuccess;
    }
    // (the buffer holding `rows` must have 4 readable bytes behind its last row: the library's staging buffers have 16)
    // A block covers up to 512 columns = 2 KB of a row: for rows up to that size one block reads whole rows, eight at a time,
    // so DRAM sees ~16 KB of nearly contiguous bytes per block step instead of 512-byte quarters of a row from four blocks at
    // different times (the same row-locality lesson as round 1's transposer).
    const uint32_t n_words = (n_bits + 31) / 32;
    unsigned bx = 1;
    while (bx < 512 && bx < n_words) {
        bx *= 2;
    }
    const unsigned by = 512 / bx, gx = (n_words + bx - 1) / bx;
    const uint64_t runs_wanted = std::max<uint64_t>(1, 148ull * 2048 * 2 / ((uint64_t)gx * bx));  // two waves of full occupancy
    const uint64_t run = std::max<uint64_t>(256, (n_shots + runs_wanted - 1) / runs_wanted);
    const uint64_t n_runs = (n_shots + run - 1) / run;
    dim3 grid(gx, (unsigned)((n_runs + by - 1) / by)), block(bx, by);
    gstim_count_b8_kernel<<<grid, block, 0, stream>>>(rows, pitch, n_shots, n_bits, single, pair);
    return cudaGetLastError();
}

cudaError_t launch_compress_rows(const uint8_t *rows, uint64_t pitch, uint32_t row_bytes, uint64_t n_shots, uint8_t *stream_buf,
                                 uint64_t capacity, unsigned long long *cursor, unsigned long long *index, uint32_t *overflow,
                                 cudaStream_t stream) {
    if (n_shots == 0) {
        return cudaSuccess;
    }
    const unsigned grid = (unsigned)std::min<uint64_t>((n_shots + 7) / 8, 148ull * 16);
    gstim_compress_rows_kernel<<<grid, 256, 0, stream>>>(rows, pitch, row_bytes, n_shots, stream_buf, capacity, cursor, index, overflow);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// SparseEngine
// ------------------------------------------------------------------------------------------------
using SparseEngine_Period = ResponsePeriod;

struct SparseEngine::Impl {
    int device = 0;
    int num_sms = 0;
    size_t smem_optin = 0;
    uint32_t D = 0, L = 0, M = 0, mode = 0;
    ResponseTable rt;
    uint32_t log_s = 0;
    uint32_t row_all = 0;                 // upper bound of the bytes per shot over all layouts
    uint32_t layout_flags = 0xFFFFFFFFu;  // layout the device table is encoded for
    uint32_t main_bits = 0, obs_bits = 0;
    void *d_thr = nullptr, *d_classes = nullptr, *d_entries = nullptr, *d_overflow = nullptr, *d_init = nullptr;
    uint32_t n_slices = 0;
    std::vector<uint32_t> slices_host;  // 4 words per slice (tests / oracle)
    // periodic storage of the device table (tables beyond the L2-friendly size): per class the stored period, or p = 0
    std::vector<SparseEngine_Period> periods;
    bool periodic = false;
    uint64_t device_entries = 0;  // entries in d_entries (after folding the periods)
    ~Impl() {
        for (void *p : {d_thr, d_classes, d_entries, d_overflow, d_init}) {
            if (p) {
                cudaFree(p);
            }
        }
    }
};

SparseEngine::SparseEngine(ResponseTable &&rt, uint32_t mode, uint32_t D, uint32_t L, uint32_t M, int device, uint32_t slice_events,
                           uint32_t tile_buffers, uint32_t compress_mb)
    : impl(new Impl()) {
    Impl &I = *impl;
    I.rt = std::move(rt);
    I.device = device;
    I.mode = mode;
    I.D = D;
    I.L = L;
    I.M = M;
    cudaDeviceProp prop;
    ck(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    I.num_sms = prop.multiProcessorCount;
    I.smem_optin = prop.sharedMemPerBlockOptin;

    // Tile height: the largest power of two <= 128 for which three tile images fit in shared memory (then two, then
    // one). It depends on the circuit only (not on the output layout or the device), so the random stream does too.
    I.row_all = mode == 0 ? (D + 7) / 8 + (L + 7) / 8 + 1 : (M + 7) / 8 + 1;
    const size_t budget = (size_t)227 * 1024 - FIXED_SMEM;
    if (I.smem_optin < (size_t)227 * 1024) {
        throw std::invalid_argument("device has less than 227 KiB of shared memory per block");
    }
    bool ok = false;
    uint32_t n_buffers_hint = 1;
    const uint32_t want_bufs = tile_buffers ? tile_buffers : 3;
    for (uint32_t bufs = want_bufs; bufs >= 1 && !ok; bufs--) {
        for (int ls = 7; ls >= 0; ls--) {
            if (((((size_t)I.row_all << ls) + 64) * bufs) <= budget) {
                I.log_s = (uint32_t)ls;
                n_buffers_hint = (uint32_t)std::min<size_t>(MAX_BUFFERS, budget / (((size_t)I.row_all << ls) + 64));
                ok = true;
                break;
            }
        }
    }
    if (!ok) {
        throw std::invalid_argument("output row does not fit in shared memory");
    }
    if (I.rt.classes.size() > MAX_CLASSES) {
        throw std::invalid_argument("more than 64 distinct site classes");
    }
    const uint32_t S = 1u << I.log_s;

    // Periodic storage: a table that would not stay in L2 (c5: 142 MB) is stored as head + one period + tail per class.
    I.periods.assign(I.rt.classes.size(), SparseEngine_Period());
    if (I.rt.entries.size() * 4 > ((size_t)compress_mb << 20) && mode == 0) {
        for (size_t ci = 0; ci < I.rt.classes.size(); ci++) {
            I.periods[ci] = find_response_period(I.rt, I.rt.classes[ci], D);
            I.periodic |= I.periods[ci].p != 0;
        }
    }
    // slices: runs of sites with about `slice_events` expected events per tile. A producer warp walks a slice 64 draws at
    // a time, so ~48 events per slice keep most draws of a step useful; circuits with few events per tile get smaller
    // slices so that there is work for every warp (0 = this automatic choice; it is part of the stream's definition).
    if (slice_events == 0) {
        // k steps of 64 draws hold a slice of mu events when mu + 1 + 1.5 sigma <= 64 k; take the largest k <= 3 that still
        // leaves two slices per producer warp among the tiles a block has in flight
        const double in_flight = I.rt.events_per_shot * S * n_buffers_hint;
        slice_events = 4;
        for (int k = 1; k <= 3; k++) {
            const double mu = 64.0 * k - 1.5 * std::sqrt(64.0 * k) - 1.0;
            if (in_flight / mu >= 2.0 * 31) {
                slice_events = (uint32_t)mu;
            }
        }
        if (slice_events == 4) {
            slice_events = (uint32_t)std::max(4.0, std::min(51.0, in_flight / 62.0));
        }
    }
    std::vector<EvClass> cls;
    std::vector<uint32_t> thr_all;
    for (const RespClass &c : I.rt.classes) {
        const double lam = std::ldexp((double)c.lam, -56);
        const double pr = c.inv == 0 ? 1.0 : -std::expm1(-lam);
        // (dense classes: 4096 trials per slice = four 32-trial words per lane)
        const double want = c.dense_thr ? 4096.0 / S : (double)slice_events / (pr * S);
        const uint32_t per = (uint32_t)std::max(1.0, std::min(want, (double)((1u << 30) >> I.log_s)));
        EvClass d{};
        d.slice0 = (uint32_t)(I.slices_host.size() / 4);
        d.per = per;
        d.n_sites = c.n_sites;
        const SparseEngine_Period &pd = I.periods[cls.size()];
        d.entry0 = (uint32_t)I.device_entries;  // (== c.entry0 when nothing is folded)
        d.per_a = pd.a;
        d.per_p = pd.p;
        d.per_n = pd.n;
        d.per_delta = pd.delta;
        I.device_entries += (uint64_t)(c.n_sites - (pd.p ? pd.n - pd.p : 0)) * c.n_out;
        d.inv = c.inv;
        d.sh = c.sh;
        d.n_out = c.n_out;
        d.kind = c.kind == RK_SINGLE ? DK_SINGLE : c.kind == RK_UNIFORM ? DK_UNIFORM : c.n_out <= 4 ? DK_THRESH3 : DK_THRESH_N;  // (RK_CHAIN: thresholds too)
        for (uint32_t j = 0; j < 3; j++) {  // registers of the chooser: threshold - 1 (pw > t <=> pw >= threshold), unused: never
            d.thr[j] = (d.kind == DK_THRESH3 && j + 1 < c.n_out) ? c.thr[j] - 1u : 0xFFFFFFFFu;
        }
        d.dense_thr = c.dense_thr;
        d.thr_off = (uint32_t)thr_all.size();
        thr_all.insert(thr_all.end(), c.thr, c.thr + 15);
        for (uint32_t s0 = 0; s0 < c.n_sites; s0 += per) {
            const uint32_t ns = std::min(per, c.n_sites - s0);
            I.slices_host.push_back((uint32_t)cls.size());
            I.slices_host.push_back(ns << I.log_s);
            I.slices_host.push_back(c.entry0 + s0 * c.n_out);
            I.slices_host.push_back(0);
        }
        d.slice_end = (uint32_t)(I.slices_host.size() / 4);
        cls.push_back(d);
    }
    I.n_slices = (uint32_t)(I.slices_host.size() / 4);
    ck(cudaSetDevice(device), "cudaSetDevice");
    auto up = [&](void **d, const void *src, size_t bytes) {
        ck(cudaMalloc(d, std::max<size_t>(bytes, 16)), "cudaMalloc (response table)");
        if (bytes && src) {
            ck(cudaMemcpy(*d, src, bytes, cudaMemcpyHostToDevice), "upload (response table)");
        }
    };
    up(&I.d_thr, thr_all.data(), thr_all.size() * 4);
    up(&I.d_classes, cls.data(), cls.size() * sizeof(EvClass));
    up(&I.d_overflow, nullptr, I.rt.overflow.size() * 4);
    up(&I.d_entries, nullptr, I.device_entries * 16);
    ck(cudaFuncSetAttribute(gstim_sparse_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I.smem_optin), "smem attribute");
    ck(cudaFuncSetAttribute(gstim_sparse_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I.smem_optin), "smem attribute");
    ck(cudaFuncSetAttribute(gstim_sparse_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I.smem_optin), "smem attribute");
    ck(cudaFuncSetAttribute(gstim_sparse_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I.smem_optin), "smem attribute");
}

SparseEngine::~SparseEngine() {
    delete impl;
}

const ResponseTable &SparseEngine::table() const {
    return impl->rt;
}
uint32_t SparseEngine::tile_shots() const {
    return 1u << impl->log_s;
}
uint64_t SparseEngine::device_table_entries() const {
    return impl->device_entries;
}
uint32_t SparseEngine::blocks_per_sm() const {
    return 1;
}
const std::vector<uint32_t> &SparseEngine::slices() const {
    return impl->slices_host;
}

void SparseEngine::set_reference_row(const uint8_t *packed, size_t n_bytes) {
    Impl &I = *impl;
    ck(cudaSetDevice(I.device), "cudaSetDevice");
    if (I.d_init) {
        cudaFree(I.d_init);
        I.d_init = nullptr;
    }
    bool any = false;
    for (size_t i = 0; packed != nullptr && i < n_bytes; i++) {
        any |= packed[i] != 0;
    }
    if (!any) {
        return;
    }
    ck(cudaMalloc(&I.d_init, n_bytes), "cudaMalloc");
    ck(cudaMemcpy(I.d_init, packed, n_bytes, cudaMemcpyHostToDevice), "upload reference row");
}

// Output ids -> bit positions of the requested layout (flags: GSTIM_PREPEND_OBS / GSTIM_APPEND_OBS / GSTIM_SEPARATE_OBS of gstim.h).
void SparseEngine::set_layout(uint32_t flags, cudaStream_t stream) {
    Impl &I = *impl;
    if (flags == I.layout_flags) {
        return;
    }
    const uint32_t D = I.D, L = I.L;
    const bool prepend = flags & 0x02u, append = flags & 0x04u, separate = flags & 0x08u;
    if (I.mode != 0) {
        I.main_bits = I.M;
        I.obs_bits = 0;
    } else {
        I.main_bits = D + ((prepend || append) ? L : 0);
        I.obs_bits = separate ? L : 0;
    }
    auto map = [&](uint32_t id) -> uint32_t {
        if (I.mode != 0) {
            return id;
        }
        if (id < D) {
            return prepend ? id + L : id;
        }
        const uint32_t l = id - D;
        if (prepend) {
            return l;
        }
        if (append) {
            return D + l;
        }
        if (separate) {
            return 0x40000000u | l;
        }
        return RESP_NONE;  // observables are not part of this output
    };
    // re-encode entries: dropped ids are squeezed out (slots stay ascending-then-NONE)
    std::vector<uint32_t> ent(I.rt.entries.size()), ovf(I.rt.overflow.size());
    std::vector<uint32_t> ids;
    for (size_t e = 0; e < I.rt.entries.size(); e += 4) {
        ids.clear();
        const uint32_t *w = &I.rt.entries[e];
        for (int j = 0; j < 4; j++) {
            if (w[j] == RESP_NONE) {
                break;
            }
            if (j == 3 && (w[j] & RESP_OVERFLOW)) {
                const uint32_t off = w[j] & 0x7FFFFFFFu, cnt = I.rt.overflow[off];
                for (uint32_t k = 1; k <= cnt; k++) {
                    ids.push_back(I.rt.overflow[off + k]);
                }
                break;
            }
            ids.push_back(w[j]);
        }
        size_t n = 0;
        for (uint32_t id : ids) {
            const uint32_t v = map(id);
            if (v != RESP_NONE) {
                ids[n++] = v;
            }
        }
        ids.resize(n);
        uint32_t o[4] = {RESP_NONE, RESP_NONE, RESP_NONE, RESP_NONE};
        if (n <= 4) {
            for (size_t j = 0; j < n; j++) {
                o[j] = ids[j];
            }
        } else {
            // reuse the entry's own overflow block (the mapped list is never longer than the original)
            const uint32_t off = w[3] & 0x7FFFFFFFu;
            for (size_t j = 0; j < 3; j++) {
                o[j] = ids[j];
            }
            o[3] = RESP_OVERFLOW | off;
            ovf[off] = (uint32_t)(n - 3);
            for (size_t j = 3; j < n; j++) {
                ovf[off + 1 + (j - 3)] = ids[j];
            }
        }
        memcpy(&ent[e], o, 16);
    }
    if (I.periodic) {
        // fold the periods: per class keep the sites before the end of the first period and the sites behind the range
        std::vector<uint32_t> folded;
        folded.reserve(I.device_entries * 4);
        for (size_t ci = 0; ci < I.rt.classes.size(); ci++) {
            const RespClass &c = I.rt.classes[ci];
            const SparseEngine_Period &pd = I.periods[ci];
            const uint32_t *src = ent.data() + (size_t)c.entry0 * 4;
            const size_t per_site = (size_t)c.n_out * 4;
            if (pd.p == 0) {
                folded.insert(folded.end(), src, src + c.n_sites * per_site);
            } else {
                folded.insert(folded.end(), src, src + (size_t)(pd.a + pd.p) * per_site);
                folded.insert(folded.end(), src + (size_t)(pd.a + pd.n) * per_site, src + c.n_sites * per_site);
            }
        }
        ent.swap(folded);
    }
    ck(cudaSetDevice(I.device), "cudaSetDevice");
    ck(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    if (!ent.empty()) {
        ck(cudaMemcpy(I.d_entries, ent.data(), ent.size() * 4, cudaMemcpyHostToDevice), "upload entries");
    }
    if (!ovf.empty()) {
        ck(cudaMemcpy(I.d_overflow, ovf.data(), ovf.size() * 4, cudaMemcpyHostToDevice), "upload overflow");
    }
    I.layout_flags = flags;
}

uint32_t SparseEngine::main_bits() const {
    return impl->main_bits;
}
uint32_t SparseEngine::obs_bits() const {
    return impl->obs_bits;
}

void SparseEngine::launch(uint64_t first_shot, uint64_t n_shots, uint8_t *main_out, uint64_t main_pitch, uint8_t *obs_out,
                          uint64_t obs_pitch, uint64_t seed, cudaStream_t stream) {
    Impl &I = *impl;
    if (n_shots == 0) {
        return;
    }
    const uint32_t S = 1u << I.log_s;
    if (first_shot % S != 0) {
        throw std::invalid_argument("shot offset must be a multiple of the engine's tile height (" + std::to_string(S) + " shots)");
    }
    {
        // a block numbers its (sequence, slice) tickets with 32 bits: cut launches that would run out of numbers
        const uint64_t seq_max = 0xFFFF0000ull / std::max<uint32_t>(1u, I.n_slices) - 1;
        const uint64_t shots_max = seq_max * (uint64_t)I.num_sms * S;
        if (n_shots > shots_max) {
            const uint64_t mb = ((uint64_t)I.main_bits + 7) / 8, ob = ((uint64_t)I.obs_bits + 7) / 8;
            for (uint64_t done = 0; done < n_shots; done += shots_max) {
                launch(first_shot + done, std::min(shots_max, n_shots - done), main_out ? main_out + done * (main_pitch ? main_pitch : mb) : nullptr,
                       main_pitch, obs_out ? obs_out + done * (obs_pitch ? obs_pitch : ob) : nullptr, obs_pitch, seed, stream);
            }
            return;
        }
    }
    SparseParams p{};
    p.classes = (const EvClass *)I.d_classes;
    p.n_classes = (uint32_t)I.rt.classes.size();
    p.n_slices = I.n_slices;
    p.thr_all = (const uint32_t *)I.d_thr;
    p.entries = (const uint4 *)I.d_entries;
    p.overflow = (const uint32_t *)I.d_overflow;
    p.init_row = (const uint8_t *)I.d_init;
    p.log_s = I.log_s;
    const uint64_t n_tiles = (n_shots + S - 1) / S;
    if (n_tiles >= (1ull << 31)) {
        throw std::invalid_argument("too many shots in one launch");
    }
    p.n_tiles = (uint32_t)n_tiles;
    p.tile0 = first_shot >> I.log_s;
    p.n_shots = n_shots;
    p.main_bytes = (I.main_bits + 7) / 8;
    p.obs_bytes = (I.obs_bits + 7) / 8;
    p.main_out = main_out;
    p.main_pitch = main_pitch ? main_pitch : p.main_bytes;
    p.obs_out = obs_out;
    p.obs_pitch = obs_pitch ? obs_pitch : p.obs_bytes;
    p.obs_img_off = align16(p.main_bytes * S + 16);
    p.img_bytes = p.obs_img_off + align16(p.obs_bytes * S + 16);
    for (uint32_t r = 0; r < 10; r++) {
        p.rk[2 * r] = (uint32_t)seed + r * 0x9E3779B9u;
        p.rk[2 * r + 1] = (uint32_t)(seed >> 32) + r * 0xBB67AE85u;
    }
    p.n_buffers = (uint32_t)std::min<size_t>(MAX_BUFFERS, ((size_t)227 * 1024 - FIXED_SMEM) / p.img_bytes);
    if (p.n_buffers == 0) {
        throw std::invalid_argument("internal: tile image exceeds shared memory");
    }
    const size_t smem = FIXED_SMEM + (size_t)p.n_buffers * p.img_bytes;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)I.num_sms);
    p.det_lo = (I.mode == 0 && (I.layout_flags & 0x02u)) ? I.L : 0;  // prepended observables come first
    p.n_det = I.mode == 0 ? I.D : 0;
    if (I.obs_bits) {
        if (I.periodic) {
            gstim_sparse_kernel<true, true><<<grid, SPARSE_THREADS, smem, stream>>>(p);
        } else {
            gstim_sparse_kernel<true, false><<<grid, SPARSE_THREADS, smem, stream>>>(p);
        }
    } else if (I.periodic) {
        gstim_sparse_kernel<false, true><<<grid, SPARSE_THREADS, smem, stream>>>(p);
    } else {
        gstim_sparse_kernel<false, false><<<grid, SPARSE_THREADS, smem, stream>>>(p);
    }
    ck(cudaGetLastError(), "gstim_sparse_kernel launch");
}

}  // namespace gstim
