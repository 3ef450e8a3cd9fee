// kernels.cu — sm_100a device code of the Pauli-frame sampler.
//
//   gstim_interp_kernel   persistent interpreter: one thread block owns K*128 shots, x/z frame bits of
//                         every qubit resident in shared memory, program streamed through a two-stage
//                         shared-memory ring by bulk-async (TMA 1D) copies.
//                         Replaces FrameSimulator<W>::do_circuit/do_gate and the per-gate row loops
//                         (/root/reference/src/stim/simulators/frame_simulator.inl:166-170, 173-912),
//                         RareErrorIterator (/root/reference/src/stim/util_bot/probability_util.cc:23-43)
//                         and MeasureRecordBatch (/root/reference/src/stim/io/measure_record_batch.inl).
//   gstim_transpose_kernel  bit-major rows -> dense shot-major b8 bytes.
//                         Replaces simd_bit_table::transposed + write_table_data
//                         (/root/reference/src/stim/io/measure_record_writer.h:101-166).
//   gstim_popcount_kernel per-row flip counts (feeds the optional multi-GPU allreduce).
#include "kernels.cuh"

#include <algorithm>

#define GSTIM_TABLE_QUAL __device__ const
#include "log2_table.h"

namespace gstim {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3").
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Exp(1) variate from a uniform u32 in fixed point (unit 2^-56 nat): -ln((r + 1/2) / 2^32) through a
// 256-entry log2 table with linear interpolation (max error 2e-6 nat). Integer-only, so the oracle
// (oracle/philox.py: exp_draw_fx) reproduces it bit for bit. lt = table in shared memory: base[256], diff[256].
__device__ __forceinline__ unsigned long long exp_draw_fx(uint32_t r, const uint32_t *lt) {
    const unsigned long long v = 2ull * r + 1ull;  // odd, < 2^33
    const int t = 63 - __clzll((long long)v);      // floor(log2 v), 0..32
    const uint32_t frac = (uint32_t)(v << (32 - t));  // bits below the leading one, left aligned
    const uint32_t i = frac >> 24, f = frac & 0xFFFFFFu;
    const unsigned long long log2m = (unsigned long long)lt[i] + (((unsigned long long)lt[256 + i] * f) >> 24);
    const unsigned long long lv = ((unsigned long long)t << 32) + log2m;
    return ((33ull << 32) - lv) * (unsigned long long)GSTIM_LN2_Q24;
}

// ------------------------------------------------------------------------------------------------
// mbarrier / bulk-async copy helpers (PTX ISA: mbarrier, cp.async.bulk)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
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

size_t interp_smem_bytes(uint32_t q_pitch, uint32_t Q, uint32_t K, uint32_t chunk_words, uint32_t max_items) {
    size_t b = 0;
    b += (size_t)2 * K * q_pitch * 16;          // X, Z
    b += (size_t)K * 16;                        // correlated-error flag row
    b += ((size_t)(Q + 1) * 8 + 15) / 16 * 16;  // exponential clocks (u64 fixed point)
    b += (size_t)2 * chunk_words * 4;           // program ring
    b += 512 * 4;                               // log2 table
    b += ((size_t)max_items * 2 + 15) / 16 * 16;  // event job queue
    b += 32;                                    // mbarriers + queue counters
    return b;
}

struct Ctx {
    uint4 *X, *Z, *flag;
    unsigned long long *clk;
    const uint32_t *lt;
    uint32_t K, G, sub, slot, slots, q_pitch, B;
    uint64_t col0;
    uint32_t k0, k1;  // philox key
    uint4 *rec;       // this block's record rows
    uint64_t rec_row_stride;
    uint32_t rec_mask;
    uint4 *out;  // this block's output columns
    uint64_t out_row_stride;
};

__device__ __forceinline__ void flip_frame(const Ctx &c, uint4 *plane, uint32_t q, uint32_t shot) {
    uint32_t *w = (uint32_t *)(plane + (size_t)(shot >> 7) * c.q_pitch + q) + ((shot >> 5) & 3);
    *w ^= 1u << (shot & 31);
}
__device__ __forceinline__ void flip_rec(const Ctx &c, uint32_t rec_index, uint32_t shot) {
    uint32_t *w = (uint32_t *)(c.rec + (uint64_t)(rec_index & c.rec_mask) * c.rec_row_stride + (shot >> 7)) + ((shot >> 5) & 3);
    *w ^= 1u << (shot & 31);
}

__device__ __forceinline__ unsigned long long sat_mul(uint32_t n, unsigned long long lam) {
    // min(n * lam, 2^63)
    const unsigned long long lo = (unsigned long long)n * lam, hi = __umul64hi((unsigned long long)n, lam);
    return (hi != 0 || lo >= (1ull << 63)) ? (1ull << 63) : lo;
}

// Walks the events of one noise site over the block's B shots with the exponential clock E (fixed point).
// on_event(shot, r) is called for every event with the event's Philox draw r (r.x re-arms the clock).
// Philox counter of the k-th event: (group, clock qubit | k << 16, col0 lo, TAG_EVENT ^ col0 hi).
template <typename F>
__device__ __forceinline__ void run_site(
    const Ctx &c, unsigned long long &E, unsigned long long lam, float inv_lam, uint32_t group, uint32_t cq, F &&on_event) {
    uint32_t pos = 0, kev = 0;
    while (pos < c.B) {
        const unsigned long long rem = sat_mul(c.B - pos, lam);
        if (E >= rem) {
            E -= rem;
            break;
        }
        // j = floor(E / lam), clamped to the shots left: float estimate + exact fix-up
        const uint32_t left = c.B - pos - 1;
        const float est = __ull2float_rz(E) * inv_lam;
        uint32_t j = est >= (float)left ? left : (uint32_t)est;
        while (j > 0 && (unsigned long long)j * lam > E) {
            j--;
        }
        while (j < left && (unsigned long long)(j + 1) * lam <= E) {
            j++;
        }
        const uint32_t shot = pos + j;
        const uint4 r = philox4x32_10(group, cq | (kev << 16), (uint32_t)c.col0, GTAG_EVENT ^ (uint32_t)(c.col0 >> 32), c.k0, c.k1);
        on_event(shot, r);
        E = exp_draw_fx(r.x, c.lt);
        pos = shot + 1;
        kev++;
    }
}

__global__ void __launch_bounds__(1024, 1) gstim_interp_kernel(const InterpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx c;
    c.K = p.K;
    c.q_pitch = p.q_pitch;
    c.B = p.K * GSTIM_COL_SHOTS;
    c.G = 1u << p.G_log2;
    c.sub = threadIdx.x & (c.G - 1);
    c.slot = threadIdx.x >> p.G_log2;
    c.slots = p.slots;
    c.k0 = p.seed_lo;
    c.k1 = p.seed_hi;
    c.rec_row_stride = p.rec_row_stride;
    c.rec_mask = p.rec_mask;
    c.out_row_stride = p.out_row_stride;

    unsigned char *sp = smem_raw;
    c.X = (uint4 *)sp;
    sp += (size_t)p.K * p.q_pitch * 16;
    c.Z = (uint4 *)sp;
    sp += (size_t)p.K * p.q_pitch * 16;
    c.flag = (uint4 *)sp;
    sp += (size_t)p.K * 16;
    c.clk = (unsigned long long *)sp;
    sp += ((size_t)(p.Q + 1) * 8 + 15) / 16 * 16;
    uint32_t *ring = (uint32_t *)sp;
    sp += (size_t)2 * p.chunk_words * 4;
    uint32_t *lt = (uint32_t *)sp;
    sp += 512 * 4;
    uint16_t *jobq = (uint16_t *)sp;
    sp += ((size_t)p.max_items * 2 + 15) / 16 * 16;
    uint64_t *mbar = (uint64_t *)sp;
    uint32_t *jobn = (uint32_t *)(mbar + 2);  // two alternating event-queue counters
    c.lt = lt;

    const uint32_t tid = threadIdx.x;
    const uint32_t T = blockDim.x;
    const uint32_t chunk_bytes = p.chunk_words * 4;
    const bool multi = p.G_log2 != 0;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        jobn[0] = 0;
        jobn[1] = 0;
    }
    for (uint32_t i = tid; i < 256; i += T) {
        lt[i] = GSTIM_LOG2_BASE[i];
        lt[256 + i] = GSTIM_LOG2_DIFF[i];
    }
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;
    uint32_t qpar = 0;  // which queue counter the next noise batch uses

    for (uint32_t g = blockIdx.x; g < p.n_blocks; g += gridDim.x) {
        c.col0 = p.col0_base + (uint64_t)g * p.K;
        c.rec = p.rec + (uint64_t)g * p.rec_block_stride + (uint64_t)blockIdx.x * p.rec_cta_stride;
        c.out = p.out + (uint64_t)g * p.K;

        // start streaming the program
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&mbar[0], chunk_bytes);
            bulk_g2s(ring, p.prog, chunk_bytes, &mbar[0]);
            if (p.n_chunks > 1) {
                mbar_expect_tx(&mbar[1], chunk_bytes);
                bulk_g2s(ring + p.chunk_words, p.prog + p.chunk_words, chunk_bytes, &mbar[1]);
            }
        }
        // per-qubit exponential clocks (+ the global clock at index Q)
        for (uint32_t q = tid; q <= p.Q; q += T) {
            uint4 r = philox4x32_10(p.logical_of[q], 0, (uint32_t)c.col0, GTAG_CLOCK ^ (uint32_t)(c.col0 >> 32), c.k0, c.k1);
            c.clk[q] = exp_draw_fx(r.x, lt);
        }
        for (uint32_t k = tid; k < p.K; k += T) {
            c.flag[k] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();

        for (uint32_t chunk = 0;; chunk++) {
            const uint32_t b = chunk & 1;
            {
                uint32_t ph = b ? phase1 : phase0;
                while (!mbar_try_wait(&mbar[b], ph)) {
                }
                if (b) {
                    phase1 ^= 1;
                } else {
                    phase0 ^= 1;
                }
            }
            const uint32_t *pw = ring + (size_t)b * p.chunk_words;
            uint32_t pc = 0;
            bool end = false;
            while (true) {
                const uint32_t h0 = pw[pc + GH_OP];
                const uint32_t op = h0 & 0xFF;
                if (op == GOP_END) {
                    end = true;
                    break;
                }
                if (op == GOP_NEXT_CHUNK) {
                    break;
                }
                const uint32_t flags = (h0 >> 8) & 0xFF;
                const uint32_t aux = h0 >> 16;
                const uint32_t n = pw[pc + GH_N];
                const uint32_t *pay = pw + pc + GSTIM_HDR_WORDS;
                if (flags & GF_BARRIER) {
                    __syncthreads();
                } else if (multi) {
                    __syncwarp();
                }
                switch (op) {
                    case GOP_CLIFF1: {
                        const uint32_t a = bitmask(aux, 0), bb = bitmask(aux, 1), cc = bitmask(aux, 2), d = bitmask(aux, 3);
                        for (uint32_t i = c.slot; i < n; i += c.slots) {
                            const uint32_t q = pay[i];
                            for (uint32_t k = c.sub; k < c.K; k += c.G) {
                                const size_t o = (size_t)k * c.q_pitch + q;
                                uint4 x = c.X[o], z = c.Z[o];
                                c.X[o] = xor4(and4(x, a), and4(z, bb));
                                c.Z[o] = xor4(and4(x, cc), and4(z, d));
                            }
                        }
                    } break;
                    case GOP_CLIFF2: {
                        if (aux == GSTIM_MAT_CX) {  // CX: z1 ^= z2 ; x2 ^= x1
                            for (uint32_t i = c.slot; i < n; i += c.slots) {
                                const uint32_t w = pay[i];
                                const uint32_t q1 = w & 0xFFFF, q2 = w >> 16;
#pragma unroll 2
                                for (uint32_t k = c.sub; k < c.K; k += c.G) {
                                    const size_t o1 = (size_t)k * c.q_pitch + q1, o2 = (size_t)k * c.q_pitch + q2;
                                    uint4 x1 = c.X[o1], z2 = c.Z[o2], z1 = c.Z[o1], x2 = c.X[o2];
                                    c.Z[o1] = xor4(z1, z2);
                                    c.X[o2] = xor4(x2, x1);
                                }
                            }
                        } else {
                            uint32_t m[16];
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                m[j] = bitmask(aux, j);
                            }
                            for (uint32_t i = c.slot; i < n; i += c.slots) {
                                const uint32_t w = pay[i];
                                const uint32_t q1 = w & 0xFFFF, q2 = w >> 16;
                                for (uint32_t k = c.sub; k < c.K; k += c.G) {
                                    const size_t o1 = (size_t)k * c.q_pitch + q1, o2 = (size_t)k * c.q_pitch + q2;
                                    uint4 x1 = c.X[o1], z1 = c.Z[o1], x2 = c.X[o2], z2 = c.Z[o2];
                                    c.X[o1] = xor4(xor4(and4(x1, m[0]), and4(z1, m[1])), xor4(and4(x2, m[2]), and4(z2, m[3])));
                                    c.Z[o1] = xor4(xor4(and4(x1, m[4]), and4(z1, m[5])), xor4(and4(x2, m[6]), and4(z2, m[7])));
                                    c.X[o2] = xor4(xor4(and4(x1, m[8]), and4(z1, m[9])), xor4(and4(x2, m[10]), and4(z2, m[11])));
                                    c.Z[o2] = xor4(xor4(and4(x1, m[12]), and4(z1, m[13])), xor4(and4(x2, m[14]), and4(z2, m[15])));
                                }
                            }
                        }
                    } break;
                    case GOP_NOISE1:
                    case GOP_NOISE2: {
                        // Pass A (item i -> thread group i % slots): advance each site's clock over the block's
                        // B shots; sites whose clock runs out inside the block are pushed on a block-wide queue.
                        // Pass B (any thread): drain the queue, so warps are full of event work instead of one
                        // busy lane in five. Philox draws are addressed by (group, qubit), not by thread.
                        const unsigned long long lam = ((unsigned long long)pw[pc + GH_LAMBDA_HI] << 32) | pw[pc + GH_LAMBDA_LO];
                        const unsigned long long need = sat_mul(c.B, lam);
                        const float inv_lam = 1.0f / __ull2float_rn(lam);
                        const uint32_t group = pw[pc + GH_SITE0], rec0 = pw[pc + GH_REC0];
                        const bool two = op == GOP_NOISE2;
                        const bool table = two && (flags & GF_TABLE) != 0;
                        const bool noframe = (flags & GF_NOFRAME) != 0;
                        const uint32_t clock_override = pw[pc + GH_EXTRA];
                        const uint32_t *items = table ? pay + 15 : pay;
                        uint32_t *qn = &jobn[qpar];
                        if (c.sub == 0) {
                            for (uint32_t i = c.slot; i < n; i += c.slots) {
                                const uint32_t q = noframe ? clock_override - 1 : (items[i] & 0xFFFF);
                                const unsigned long long E = c.clk[q];
                                if (E >= need) {
                                    c.clk[q] = E - need;
                                } else {
                                    jobq[atomicAdd(qn, 1u)] = (uint16_t)i;
                                }
                            }
                        }
                        __syncthreads();
                        const uint32_t njobs = *qn;
                        if (tid == 0) {
                            jobn[qpar ^ 1] = 0;  // nobody touches the other counter until the next noise batch
                        }
                        qpar ^= 1;
                        const uint32_t t1 = pw[pc + GH_T1], t2 = pw[pc + GH_T2], t3 = pw[pc + GH_T3];
                        for (uint32_t jb = tid; jb < njobs; jb += T) {
                            const uint32_t i = jobq[jb];
                            const uint32_t w = items[i];
                            const uint32_t q1 = noframe ? clock_override - 1 : (w & 0xFFFF), q2 = w >> 16;
                            unsigned long long E = c.clk[q1];
                            run_site(c, E, lam, inv_lam, group, p.logical_of[q1], [&](uint32_t shot, uint4 r) {
                                if (!two) {
                                    const uint32_t v = r.y;
                                    const uint32_t sel = v < t1 ? 0u : v < t2 ? 2u : v < t3 ? 4u : 6u;
                                    const uint32_t cat = (aux >> sel) & 3u;
                                    if (cat & 1u) {
                                        flip_frame(c, c.X, q1, shot);
                                    }
                                    if (cat & 2u) {
                                        flip_frame(c, c.Z, q1, shot);
                                    }
                                    if (flags & GF_REC) {
                                        flip_rec(c, rec0 + i, shot);
                                    }
                                } else {
                                    uint32_t fx1, fz1, fx2, fz2;
                                    if (!table) {
                                        // uniform over the 15 non-identity pairs (frame_simulator.inl:651-659)
                                        const uint32_t pr = 1u + __umulhi(r.y, 15u);
                                        fx1 = pr & 1u;
                                        fz1 = (pr >> 1) & 1u;
                                        fx2 = (pr >> 2) & 1u;
                                        fz2 = (pr >> 3) & 1u;
                                    } else {
                                        uint32_t pr = aux;
                                        for (uint32_t j = 0; j < 15; j++) {
                                            if (r.y < pay[j]) {
                                                pr = j + 1;
                                                break;
                                            }
                                        }
                                        // index = 4*P1 + P2 with P: 0=I 1=X 2=Y 3=Z (tableau_simulator.h:307-316)
                                        const uint32_t c1 = pr >> 2, c2 = pr & 3u;
                                        fx1 = ((c1 + 1) >> 1) & 1u;
                                        fz1 = c1 >> 1;
                                        fx2 = ((c2 + 1) >> 1) & 1u;
                                        fz2 = c2 >> 1;
                                    }
                                    if (fx1) {
                                        flip_frame(c, c.X, q1, shot);
                                    }
                                    if (fz1) {
                                        flip_frame(c, c.Z, q1, shot);
                                    }
                                    if (fx2) {
                                        flip_frame(c, c.X, q2, shot);
                                    }
                                    if (fz2) {
                                        flip_frame(c, c.Z, q2, shot);
                                    }
                                }
                            });
                            c.clk[q1] = E;
                        }
                        __syncthreads();  // events were applied by arbitrary threads
                    } break;
                    case GOP_MEASURE: {
                        const uint32_t basis = aux & 3u, kind = (aux >> 2) & 3u;
                        const uint32_t mgroup = pw[pc + GH_CSITE0], rec0 = pw[pc + GH_REC0];
                        for (uint32_t i = c.slot; i < n; i += c.slots) {
                            const uint32_t q = pay[i];
                            const uint32_t lq = p.logical_of[q];
                            uint4 *rrow = c.rec + (uint64_t)((rec0 + i) & c.rec_mask) * c.rec_row_stride;
                            for (uint32_t k = c.sub; k < c.K; k += c.G) {
                                const size_t o = (size_t)k * c.q_pitch + q;
                                const uint64_t col = c.col0 + k;
                                const uint4 rnd = philox4x32_10(mgroup, lq, (uint32_t)col, GTAG_COLLAPSE ^ (uint32_t)(col >> 32), c.k0, c.k1);
                                uint4 x = c.X[o], z = c.Z[o];
                                uint4 m, nx, nz;
                                const uint4 zero = make_uint4(0, 0, 0, 0);
                                if (basis == GB_Z) {  // frame_simulator.inl:199-208, 266-274, 306-317
                                    m = x;
                                    nx = kind == GK_M ? x : zero;
                                    nz = rnd;
                                } else if (basis == GB_X) {  // :173-182, 211-219, 277-288
                                    m = z;
                                    nz = kind == GK_M ? z : zero;
                                    nx = rnd;
                                } else {  // Y basis :185-196, 255-263, 291-303
                                    m = xor4(x, z);
                                    nz = rnd;
                                    nx = kind == GK_M ? xor4(m, rnd) : rnd;
                                }
                                c.X[o] = nx;
                                c.Z[o] = nz;
                                if (kind != GK_R) {
                                    rrow[k] = m;
                                }
                            }
                        }
                    } break;
                    case GOP_RECZERO: {
                        const uint32_t rec0 = pw[pc + GH_REC0];
                        for (uint32_t i = c.slot; i < n; i += c.slots) {
                            uint4 *rrow = c.rec + (uint64_t)((rec0 + i) & c.rec_mask) * c.rec_row_stride;
                            for (uint32_t k = c.sub; k < c.K; k += c.G) {
                                rrow[k] = make_uint4(0, 0, 0, 0);
                            }
                        }
                    } break;
                    case GOP_XORROWS: {
                        const uint32_t *dst = pay, *off = pay + n, *idx = pay + 2 * n + 1;
                        for (uint32_t i = c.slot; i < n; i += c.slots) {
                            const uint32_t b0 = off[i], b1 = off[i + 1];
                            uint4 *orow = c.out + (uint64_t)dst[i] * c.out_row_stride;
                            for (uint32_t k = c.sub; k < c.K; k += c.G) {
                                uint4 acc = make_uint4(0, 0, 0, 0);
                                for (uint32_t j = b0; j < b1; j++) {
                                    acc = xor4(acc, c.rec[(uint64_t)idx[j] * c.rec_row_stride + k]);
                                }
                                if (flags & GF_ACCUM) {
                                    acc = xor4(acc, orow[k]);
                                }
                                orow[k] = acc;
                            }
                        }
                    } break;
                    case GOP_OBS_PAULI: {
                        for (uint32_t i = c.slot; i < n; i += c.slots) {
                            const uint32_t w = pay[2 * i + 1];
                            const uint32_t q = w & 0xFFFFFF;
                            uint4 *orow = c.out + (uint64_t)pay[2 * i] * c.out_row_stride;
                            for (uint32_t k = c.sub; k < c.K; k += c.G) {
                                const size_t o = (size_t)k * c.q_pitch + q;
                                uint4 acc = orow[k];
                                if (w & (1u << 30)) {
                                    acc = xor4(acc, c.X[o]);
                                }
                                if (w & (1u << 31)) {
                                    acc = xor4(acc, c.Z[o]);
                                }
                                orow[k] = acc;
                            }
                        }
                    } break;
                    case GOP_FEEDBACK: {
                        for (uint32_t i = c.slot; i < n; i += c.slots) {
                            const uint32_t w = pay[2 * i + 1];
                            const uint32_t q = w & 0xFFFFFF;
                            const uint4 *rrow = c.rec + (uint64_t)pay[2 * i] * c.rec_row_stride;
                            for (uint32_t k = c.sub; k < c.K; k += c.G) {
                                const size_t o = (size_t)k * c.q_pitch + q;
                                const uint4 r = rrow[k];
                                if (w & (1u << 30)) {
                                    c.X[o] = xor4(c.X[o], r);
                                }
                                if (w & (1u << 31)) {
                                    c.Z[o] = xor4(c.Z[o], r);
                                }
                            }
                        }
                    } break;
                    case GOP_CORR: {
                        // E / ELSE_CORRELATED_ERROR (frame_simulator.inl:747-776): one site for the whole
                        // Pauli product, masked by (and recorded in) the block's "already occurred" row.
                        if (tid == 0) {
                            if (flags & GF_RESET_FLAG) {
                                for (uint32_t k = 0; k < c.K; k++) {
                                    c.flag[k] = make_uint4(0, 0, 0, 0);
                                }
                            }
                            const unsigned long long lam = ((unsigned long long)pw[pc + GH_LAMBDA_HI] << 32) | pw[pc + GH_LAMBDA_LO];
                            if (lam != 0) {
                                const uint32_t cq = pw[pc + GH_EXTRA];
                                unsigned long long E = c.clk[cq];
                                run_site(c, E, lam, 1.0f / __ull2float_rn(lam), pw[pc + GH_SITE0], p.logical_of[cq], [&](uint32_t shot, uint4 r) {
                                    uint32_t *fw = (uint32_t *)(c.flag + (shot >> 7)) + ((shot >> 5) & 3);
                                    const uint32_t bit = 1u << (shot & 31);
                                    if (!(*fw & bit)) {
                                        *fw |= bit;
                                        for (uint32_t j = 0; j < n; j++) {
                                            const uint32_t w = pay[j];
                                            if (w & (1u << 30)) {
                                                flip_frame(c, c.X, w & 0xFFFFFF, shot);
                                            }
                                            if (w & (1u << 31)) {
                                                flip_frame(c, c.Z, w & 0xFFFFFF, shot);
                                            }
                                        }
                                    }
                                });
                                c.clk[cq] = E;
                            }
                        }
                    } break;
                    default:
                        break;
                }
                pc += pw[pc + GH_WORDS];
            }
            __syncthreads();  // everyone is done reading ring[b]
            if (end) {
                break;
            }
            if (tid == 0 && chunk + 2 < p.n_chunks) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&mbar[b], chunk_bytes);
                bulk_g2s(ring + (size_t)b * p.chunk_words, p.prog + (size_t)(chunk + 2) * p.chunk_words, chunk_bytes, &mbar[b]);
            }
        }
    }
}

cudaError_t interp_set_max_smem(size_t smem) {
    return cudaFuncSetAttribute(gstim_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t launch_interp(const InterpParams &p, uint32_t grid, uint32_t threads, size_t smem, cudaStream_t stream) {
    gstim_interp_kernel<<<grid, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Output transposer. Tile = 512 shots x 1024 bits. Thread (sw, rg): sw = shot word 0..15 of the tile,
// rg = group of 32 output bits. Each thread gathers 32 source rows' words (coalesced across sw),
// transposes the 32x32 bit block in registers, parks it in shared memory, then warps stream whole
// 128-byte shot segments to the dense (arbitrarily aligned) output rows.
// ------------------------------------------------------------------------------------------------
constexpr int TP_SHOTS = 512;
constexpr int TP_BITS = 1024;
constexpr int TP_PITCH = 33;  // words per shot row in shared memory (+1 to spread banks)

__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
    // After this, a[j] bit i == (input a[i]) bit j.
#pragma unroll
    for (int st = 0; st < 5; st++) {
        const int j = 16 >> st;
        const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            if ((k & j) == 0) {
                const uint32_t t = ((a[k] >> j) ^ a[k | j]) & m;
                a[k] ^= t << j;
                a[k | j] ^= t;
            }
        }
    }
}

__global__ void __launch_bounds__(512) gstim_transpose_kernel(const TransposeParams p) {
    extern __shared__ __align__(16) uint32_t tile[];
    const uint32_t tid = threadIdx.x;
    const uint32_t sw = tid & 15, rg = tid >> 4;
    const uint64_t shot_word0 = (uint64_t)blockIdx.x * (TP_SHOTS / 32);
    const uint32_t bit0 = blockIdx.y * TP_BITS;

    uint32_t a[32];
    const uint64_t sword = shot_word0 + sw;
    const bool in_range = sword < p.row_words;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const uint32_t bit = bit0 + rg * 32 + i;
        uint32_t v = 0;
        if (bit < p.n_bits && in_range) {
            const uint32_t rm = p.row_map[bit];
            v = p.table[(uint64_t)(rm & 0x7FFFFFFFu) * p.row_words + sword];
            v ^= (uint32_t)0 - (rm >> 31);
        }
        a[i] = v;
    }
    transpose32(a);
    // shot (sw*32+s) is parked at tile row (s*16+sw): bank = (16 s + sw + rg) % 32
#pragma unroll
    for (int s = 0; s < 32; s++) {
        tile[(s * 16 + sw) * TP_PITCH + rg] = a[s];
    }
    __syncthreads();

    const uint32_t nbytes = (p.n_bits + 7) / 8;
    const uint32_t seg0 = blockIdx.y * (TP_BITS / 8);
    const uint32_t seg_len = min((uint32_t)(TP_BITS / 8), nbytes - seg0);
    const uint32_t lane = tid & 31, warp = tid >> 5;
    for (uint32_t sl = warp; sl < TP_SHOTS; sl += 16) {
        const uint64_t shot = (uint64_t)blockIdx.x * TP_SHOTS + sl;
        if (shot >= p.n_shots) {
            break;
        }
        const uint32_t *S = &tile[((sl & 31) * 16 + (sl >> 5)) * TP_PITCH];
        uint8_t *dst = p.out + shot * p.out_pitch + seg0;
        const uint32_t mis = (uint32_t)((uintptr_t)dst & 3);
        uint8_t *base = dst - mis;  // 4-byte aligned
        // aligned destination word w holds segment bytes [4w - mis, 4w - mis + 4)
        for (uint32_t w = lane; w * 4 < mis + seg_len; w += 32) {
            const uint32_t lo = w >= 1 ? S[w - 1] : 0u;
            const uint32_t hi = w < 32 ? S[w] : 0u;
            const uint32_t val = mis == 0 ? hi : __funnelshift_r(lo, hi, 8 * (4 - mis));
            const int first = (int)(4 * w) - (int)mis;  // segment byte index of this word's byte 0
            if (first >= 0 && first + 4 <= (int)seg_len) {
                *(uint32_t *)(base + 4 * w) = val;
            } else {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int sb = first + b;
                    if (sb >= 0 && sb < (int)seg_len) {
                        base[4 * w + b] = (uint8_t)(val >> (8 * b));
                    }
                }
            }
        }
    }
}

cudaError_t launch_transpose_b8(const TransposeParams &p, cudaStream_t stream) {
    if (p.n_bits == 0 || p.n_shots == 0) {
        return cudaSuccess;
    }
    dim3 grid((unsigned)((p.n_shots + TP_SHOTS - 1) / TP_SHOTS), (p.n_bits + TP_BITS - 1) / TP_BITS);
    static bool attr_set = false;
    const size_t smem = (size_t)TP_SHOTS * TP_PITCH * 4;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gstim_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            return e;
        }
        attr_set = true;
    }
    gstim_transpose_kernel<<<grid, 512, smem, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Row popcounts: one warp per (row, slice of shots).
// ------------------------------------------------------------------------------------------------
__global__ void gstim_popcount_kernel(
    const uint32_t *table, uint64_t row_words, uint32_t n_rows, uint64_t n_shots, unsigned long long *counts) {
    const uint32_t row = blockIdx.x;
    if (row >= n_rows) {
        return;
    }
    const uint64_t full_words = n_shots / 32;
    const uint32_t tail_bits = (uint32_t)(n_shots & 31);
    const uint32_t *r = table + (uint64_t)row * row_words;
    unsigned long long acc = 0;
    for (uint64_t w = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x; w < full_words; w += (uint64_t)gridDim.y * blockDim.x) {
        acc += __popc(r[w]);
    }
    if (tail_bits && blockIdx.y == 0 && threadIdx.x == 0) {
        acc += __popc(r[full_words] & ((1u << tail_bits) - 1));
    }
    for (int o = 16; o; o >>= 1) {
        acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    }
    if ((threadIdx.x & 31) == 0 && acc) {
        atomicAdd(&counts[row], acc);
    }
}

cudaError_t launch_row_popcount(
    const uint32_t *table, uint64_t row_words, uint32_t n_rows, uint64_t n_shots, unsigned long long *counts, cudaStream_t stream) {
    if (n_rows == 0 || n_shots == 0) {
        return cudaSuccess;
    }
    uint32_t slices = (uint32_t)std::min<uint64_t>(64, (n_shots / 32 + 255) / 256);
    if (slices == 0) {
        slices = 1;
    }
    dim3 grid(n_rows, slices);
    gstim_popcount_kernel<<<grid, 256, 0, stream>>>(table, row_words, n_rows, n_shots, counts);
    return cudaGetLastError();
}

}  // namespace gstim
