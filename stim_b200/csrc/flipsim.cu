// flipsim.cu — interactive Pauli-frame simulator with device-resident state.
//
// Replaces the state and the per-instruction row loops behind stim.FlipSimulator
// (/root/reference/src/stim/simulators/frame_simulator.pybind.cc:507-1561 over FrameSimulator<W>,
// /root/reference/src/stim/simulators/frame_simulator.inl:153-1113): a batch of shots whose x / z frame rows, measurement
// flip record, detector flips and observable flips stay in HBM between calls, advanced one circuit fragment at a time.
//
// A fragment (`do`) goes through the same host lowering as the bulk samplers (lowering.cc: lower_fragment — MPP / SPP /
// pair measurements decomposed, PAULI_CHANNELs folded, rec[-k] made absolute over the whole history, items grouped into
// batches that touch disjoint rows) and every batch is ONE small kernel over the HBM-resident bit tables: thread = (item,
// 128 shots). Unlike the bulk engines nothing is fused or kept on chip — the point of this API is that the caller may look
// at and edit the frame between any two instructions; throughput per instruction is one pass over the rows it touches.
//
// Tables are bit-major: row r of a table = uint32[W] words over the batch (W a multiple of 4), the layout
// get_measurement_flips / get_detector_flips / peek_pauli_flips hand out ([row, instance]).
// Noise uses the engines' 32-bit geometric gaps (one thread walks one site over the batch); collapse randomisation and
// Bernoulli masks are Philox words. RNG addressing: Philox4x32-10, key = seed, counter = (batch ordinal since creation,
// item, call or word index, tag).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/gstim.h"
#include "circuit.h"
#include "lowering.h"

#define GSTIM_TABLE_QUAL __device__ const
#include "log2_q26_table.h"

namespace gstim {

namespace {

constexpr uint32_t FS_ITEM_X = 1u << 30, FS_ITEM_Z = 1u << 31;
constexpr uint32_t TAG_COLLAPSE = 0x46534331u, TAG_NOISE = 0x46534E31u, TAG_MASK = 0x46534D31u;  // 'FSC1' 'FSN1' 'FSM1'

__device__ __forceinline__ uint4 fs_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
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

// floor(Exp(1) / lambda) in the 32-bit fixed point of dem.cu (gap = (E * inv) >> sh)
__device__ __forceinline__ unsigned long long fs_gap(uint32_t word, uint32_t inv, uint32_t sh) {
    const uint32_t v = word | 1u;
    const uint32_t t = 31u - (uint32_t)__clz((int)v);
    const uint32_t frac = (v << (31u - t)) << 1;
    const uint32_t i = frac >> 24;
    const uint32_t log2v = (t << 26) + GSTIM_LOG2_Q26[2 * i] + ((GSTIM_LOG2_Q26[2 * i + 1] * ((frac >> 11) & 0x1FFFu)) >> 13);
    const uint32_t E = __umulhi(0x80000000u - log2v, GSTIM_LN2_Q32);
    return ((unsigned long long)E * inv) >> sh;
}

struct FsState {
    uint32_t *X, *Z, *REC, *DET, *OBS, *FLAG;
    uint32_t W;          // words per row
    uint32_t batch;      // shots
    uint32_t seed_lo, seed_hi;
};

__device__ __forceinline__ uint32_t tail_mask(const FsState &st, uint32_t w) {
    const uint32_t lo = w * 32u;
    if (lo + 32u <= st.batch) {
        return 0xFFFFFFFFu;
    }
    return lo >= st.batch ? 0u : ((1u << (st.batch - lo)) - 1u);
}

__global__ void fs_cliff1(FsState st, uint32_t mat, uint32_t n, const uint32_t *items) {
    const uint32_t a = (mat & 1) ? ~0u : 0u, b = (mat & 2) ? ~0u : 0u, c = (mat & 4) ? ~0u : 0u, d = (mat & 8) ? ~0u : 0u;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (uint64_t)n * st.W; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t q = items[i / st.W], w = (uint32_t)(i % st.W);
        uint32_t *px = st.X + (size_t)q * st.W + w, *pz = st.Z + (size_t)q * st.W + w;
        const uint32_t x = *px, z = *pz;
        *px = (x & a) ^ (z & b);
        *pz = (x & c) ^ (z & d);
    }
}

__global__ void fs_cliff2(FsState st, uint32_t mat, uint32_t n, const uint32_t *items) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (uint64_t)n * st.W; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t it = items[i / st.W], w = (uint32_t)(i % st.W);
        const uint32_t q1 = it & 0xFFFF, q2 = it >> 16;
        uint32_t *p[4] = {st.X + (size_t)q1 * st.W + w, st.Z + (size_t)q1 * st.W + w, st.X + (size_t)q2 * st.W + w, st.Z + (size_t)q2 * st.W + w};
        const uint32_t v[4] = {*p[0], *p[1], *p[2], *p[3]};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                o ^= ((mat >> (4 * k + j)) & 1u) ? v[j] : 0u;
            }
            *p[k] = o;
        }
    }
}

// M / MR / R in basis X / Y / Z (frame_simulator.inl:173-317); thread = (item, group of 4 words = 128 shots)
__global__ void fs_measure(FsState st, uint32_t basis, uint32_t kind, uint32_t n, const uint32_t *items, uint32_t rec0, uint32_t randomize,
                           uint32_t ordinal) {
    const uint32_t W4 = st.W / 4;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (uint64_t)n * W4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t item = (uint32_t)(i / W4), g = (uint32_t)(i % W4);
        const uint32_t q = items[item] & 0xFFFF;
        uint4 *px = reinterpret_cast<uint4 *>(st.X + (size_t)q * st.W) + g, *pz = reinterpret_cast<uint4 *>(st.Z + (size_t)q * st.W) + g;
        uint4 x = *px, z = *pz;
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if (randomize) {
            rnd = fs_philox(ordinal, item, g, TAG_COLLAPSE, st.seed_lo, st.seed_hi);
            rnd.x &= tail_mask(st, 4 * g);
            rnd.y &= tail_mask(st, 4 * g + 1);
            rnd.z &= tail_mask(st, 4 * g + 2);
            rnd.w &= tail_mask(st, 4 * g + 3);
        }
        auto x4 = [](uint4 a, uint4 b) { return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); };
        const uint4 zero = make_uint4(0, 0, 0, 0);
        uint4 m;
        if (basis == GB_Z) {
            m = x;
            if (kind != GK_M) {
                x = zero;
            }
            z = randomize ? rnd : z;  // (no randomisation: the conjugate component is kept, like the reference)
        } else if (basis == GB_X) {
            m = z;
            if (kind != GK_M) {
                z = zero;
            }
            x = randomize ? rnd : x;
        } else {
            m = x4(x, z);
            if (randomize) {
                z = rnd;
                x = kind == GK_M ? x4(m, rnd) : rnd;
            } else if (kind != GK_M) {
                x = z;
            }
        }
        *px = x;
        *pz = z;
        if (kind != GK_R) {
            reinterpret_cast<uint4 *>(st.REC + (size_t)(rec0 + item) * st.W)[g] = m;
        }
    }
}

__global__ void fs_zero_rows(uint32_t *rows, uint64_t n_words) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
        rows[i] = 0;
    }
}

// DETECTOR / OBSERVABLE_INCLUDE rec targets: dst rows (fragment numbering: detectors first, then observables)
__global__ void fs_xorrows(FsState st, uint32_t n, const uint32_t *dst, const uint32_t *off, const uint32_t *idx, uint32_t accum,
                           uint32_t frag_dets, uint32_t det0) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (uint64_t)n * st.W; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t item = (uint32_t)(i / st.W), w = (uint32_t)(i % st.W);
        uint32_t acc = 0;
        for (uint32_t j = off[item]; j < off[item + 1]; j++) {
            acc ^= st.REC[(size_t)idx[j] * st.W + w];
        }
        const uint32_t d = dst[item];
        uint32_t *p = d < frag_dets ? st.DET + (size_t)(det0 + d) * st.W + w : st.OBS + (size_t)(d - frag_dets) * st.W + w;
        *p = accum ? (*p ^ acc) : acc;
    }
}

__global__ void fs_obs_pauli(FsState st, uint32_t n, const uint32_t *items, uint32_t frag_dets) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (uint64_t)n * st.W; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t item = (uint32_t)(i / st.W), w = (uint32_t)(i % st.W);
        const uint32_t d = items[2 * item], wq = items[2 * item + 1], q = wq & 0xFFFFFF;
        uint32_t v = 0;
        if (wq & FS_ITEM_X) {
            v ^= st.X[(size_t)q * st.W + w];
        }
        if (wq & FS_ITEM_Z) {
            v ^= st.Z[(size_t)q * st.W + w];
        }
        st.OBS[(size_t)(d - frag_dets) * st.W + w] ^= v;
    }
}

__global__ void fs_feedback(FsState st, uint32_t n, const uint32_t *items) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (uint64_t)n * st.W; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t item = (uint32_t)(i / st.W), w = (uint32_t)(i % st.W);
        const uint32_t r = st.REC[(size_t)items[2 * item] * st.W + w], wq = items[2 * item + 1], q = wq & 0xFFFFFF;
        if (wq & FS_ITEM_X) {
            st.X[(size_t)q * st.W + w] ^= r;
        }
        if (wq & FS_ITEM_Z) {
            st.Z[(size_t)q * st.W + w] ^= r;
        }
    }
}

struct FsNoise {
    uint32_t op, flags, aux, extra, t1, t2, t3, rec0, inv, sh, n;
    const uint32_t *items;   // NOISE2 with GF_TABLE: 15 thresholds first
};

// NOISE1 / NOISE2: one thread walks one site over the batch with geometric gaps; the rows of the items of a batch are
// disjoint, so plain read-modify-writes suffice.
__global__ void fs_noise(FsState st, FsNoise nz, uint32_t ordinal) {
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nz.n) {
        return;
    }
    const bool pair = nz.op == GOP_NOISE2, table = pair && (nz.flags & GF_TABLE);
    const uint32_t it = nz.items[(table ? 15u : 0u) + item];
    uint32_t pos = 0;
    for (uint32_t call = 0;; call++) {
        const uint4 r = fs_philox(ordinal, item, call, TAG_NOISE, st.seed_lo, st.seed_hi);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const unsigned long long G = fs_gap(h ? r.z : r.x, nz.inv, nz.sh);
            if (G >= (unsigned long long)(st.batch - pos)) {
                return;
            }
            pos += (uint32_t)G;
            const uint32_t v = h ? r.w : r.y, w = pos >> 5, bit = 1u << (pos & 31u);
            if (!pair) {
                const uint32_t sel = v < nz.t1 ? 0u : v < nz.t2 ? 2u : v < nz.t3 ? 4u : 6u;
                const uint32_t cat = (nz.aux >> sel) & 3u;
                if (!(nz.flags & GF_NOFRAME)) {
                    if (cat & 1) {
                        st.X[(size_t)it * st.W + w] ^= bit;
                    }
                    if (cat & 2) {
                        st.Z[(size_t)it * st.W + w] ^= bit;
                    }
                }
                if (nz.flags & GF_REC) {
                    st.REC[(size_t)(nz.rec0 + item) * st.W + w] ^= bit;
                }
            } else {
                uint32_t fx1, fz1, fx2, fz2;
                if (!table) {
                    const uint32_t pr = 1u + __umulhi(v, 15u);
                    fx1 = pr & 1;
                    fz1 = (pr >> 1) & 1;
                    fx2 = (pr >> 2) & 1;
                    fz2 = (pr >> 3) & 1;
                } else {
                    uint32_t pr = nz.aux;
                    for (uint32_t j = 0; j < 15; j++) {
                        if (v < nz.items[j]) {
                            pr = j + 1;
                            break;
                        }
                    }
                    const uint32_t c1 = pr >> 2, c2 = pr & 3;
                    fx1 = ((c1 + 1) >> 1) & 1;
                    fz1 = c1 >> 1;
                    fx2 = ((c2 + 1) >> 1) & 1;
                    fz2 = c2 >> 1;
                }
                const uint32_t q1 = it & 0xFFFF, q2 = it >> 16;
                if (fx1) {
                    st.X[(size_t)q1 * st.W + w] ^= bit;
                }
                if (fz1) {
                    st.Z[(size_t)q1 * st.W + w] ^= bit;
                }
                if (fx2) {
                    st.X[(size_t)q2 * st.W + w] ^= bit;
                }
                if (fz2) {
                    st.Z[(size_t)q2 * st.W + w] ^= bit;
                }
            }
            pos++;
            if (pos >= st.batch) {
                return;
            }
        }
    }
}

// E / ELSE_CORRELATED_ERROR (frame_simulator.inl:747-776): thread = 32 shots; a shot whose coin comes up applies the Pauli
// product unless an earlier element of the chain already did (FLAG row).
__global__ void fs_corr(FsState st, uint32_t n_targets, const uint32_t *targets, uint32_t reset_flag, uint32_t threshold, uint32_t always,
                        uint32_t ordinal) {
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < st.W; w += gridDim.x * blockDim.x) {
        uint32_t coin = 0;
        if (always) {
            coin = ~0u;
        } else {
            for (uint32_t b = 0; b < 32; b += 4) {
                const uint4 r = fs_philox(ordinal, w, b, TAG_MASK, st.seed_lo, st.seed_hi);
                coin |= (r.x < threshold ? 1u : 0u) << b | (r.y < threshold ? 1u : 0u) << (b + 1) | (r.z < threshold ? 1u : 0u) << (b + 2) |
                        (r.w < threshold ? 1u : 0u) << (b + 3);
            }
        }
        coin &= tail_mask(st, w);
        const uint32_t flag = reset_flag ? 0u : st.FLAG[w];
        const uint32_t hit = coin & ~flag;
        st.FLAG[w] = flag | hit;
        for (uint32_t t = 0; t < n_targets; t++) {
            const uint32_t wq = targets[t], q = wq & 0xFFFFFF;
            if (wq & FS_ITEM_X) {
                st.X[(size_t)q * st.W + w] ^= hit;
            }
            if (wq & FS_ITEM_Z) {
                st.Z[(size_t)q * st.W + w] ^= hit;
            }
        }
    }
}

// rows ^= mask & Bernoulli(p) words (broadcast_pauli_errors, generate_bernoulli_samples); p >= 1: mask itself
__global__ void fs_masked_flip(FsState st, uint32_t *rows, const uint32_t *mask, uint64_t n_words, uint32_t threshold, uint32_t always,
                               uint32_t ordinal, uint32_t overwrite) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t coin = ~0u;
        if (!always) {
            coin = 0;
            for (uint32_t b = 0; b < 32; b += 4) {
                const uint4 r = fs_philox(ordinal, (uint32_t)i, (uint32_t)(i >> 32) * 8u + b / 4, TAG_MASK, st.seed_lo, st.seed_hi);
                coin |= (r.x < threshold ? 1u : 0u) << b | (r.y < threshold ? 1u : 0u) << (b + 1) | (r.z < threshold ? 1u : 0u) << (b + 2) |
                        (r.w < threshold ? 1u : 0u) << (b + 3);
            }
        }
        const uint32_t v = coin & (mask ? mask[i] : ~0u);
        rows[i] = overwrite ? v : (rows[i] ^ v);
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

struct Table {
    uint32_t *p = nullptr;
    uint64_t rows = 0, cap = 0;
};

}  // namespace

}  // namespace gstim

using namespace gstim;

void gstim_set_last_error(const char *msg);

struct gstim_flipsim {
    int device = 0;
    uint32_t batch = 0, W = 0;
    bool randomize = true;
    uint64_t seed = 0;
    uint32_t ordinal = 0;  // batches executed since creation (Philox counter word 0)
    Table X, Z, REC, DET, OBS;
    uint32_t *FLAG = nullptr;
    uint32_t *d_payload = nullptr;
    size_t payload_cap = 0;
    cudaStream_t stream = nullptr;
    ~gstim_flipsim() {
        for (uint32_t *p : {X.p, Z.p, REC.p, DET.p, OBS.p, FLAG, d_payload}) {
            if (p) {
                cudaFree(p);
            }
        }
        if (stream) {
            cudaStreamDestroy(stream);
        }
    }
};

namespace {

template <typename F>
int fs_guarded(F &&f) {
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

// Grows a table to at least `rows` rows (new rows zero), keeping its contents.
void grow(gstim_flipsim *h, Table &t, uint64_t rows) {
    if (rows <= t.rows) {
        return;
    }
    if (rows > t.cap) {
        const uint64_t cap = std::max<uint64_t>({rows, t.cap * 2, 16});
        uint32_t *p = nullptr;
        ck(cudaMalloc(&p, cap * h->W * 4), "cudaMalloc");
        ck(cudaMemsetAsync(p, 0, cap * h->W * 4, h->stream), "cudaMemset");
        if (t.p) {
            ck(cudaMemcpyAsync(p, t.p, t.rows * h->W * 4, cudaMemcpyDeviceToDevice, h->stream), "copy");
            ck(cudaStreamSynchronize(h->stream), "sync");
            cudaFree(t.p);
        }
        t.p = p;
        t.cap = cap;
    }
    t.rows = rows;
}

FsState state_of(const gstim_flipsim *h) {
    FsState st{};
    st.X = h->X.p;
    st.Z = h->Z.p;
    st.REC = h->REC.p;
    st.DET = h->DET.p;
    st.OBS = h->OBS.p;
    st.FLAG = h->FLAG;
    st.W = h->W;
    st.batch = h->batch;
    st.seed_lo = (uint32_t)h->seed;
    st.seed_hi = (uint32_t)(h->seed >> 32);
    return st;
}

uint32_t grid_for(uint64_t work) {
    return (uint32_t)std::min<uint64_t>(std::max<uint64_t>((work + 255) / 256, 1), 148 * 8);
}

const uint32_t *upload(gstim_flipsim *h, const std::vector<uint32_t> &words) {
    if (words.size() > h->payload_cap) {
        // (kernels of earlier batches may still be reading the old buffer)
        ck(cudaStreamSynchronize(h->stream), "sync");
        if (h->d_payload) {
            cudaFree(h->d_payload);
        }
        h->payload_cap = std::max<size_t>(words.size() * 2, 4096);
        ck(cudaMalloc(&h->d_payload, h->payload_cap * 4), "cudaMalloc");
    } else {
        ck(cudaStreamSynchronize(h->stream), "sync");  // single payload buffer: the previous batch must be done with it
    }
    if (!words.empty()) {
        ck(cudaMemcpyAsync(h->d_payload, words.data(), words.size() * 4, cudaMemcpyHostToDevice, h->stream), "upload");
    }
    return h->d_payload;
}

void gap_params(uint64_t lam, uint32_t *inv, uint32_t *sh) {
    if (lam >= (1ull << 62)) {
        *inv = 0;
        *sh = 0;
        return;
    }
    int e = 0;
    const double m = std::frexp(std::ldexp(1.0, 56) / (double)lam, &e);
    const int s = 58 - e;
    if (s < 0 || s > 63) {
        *inv = 0xFFFFFFFFu;  // a rate too small to ever fire within a batch
        *sh = 0;
        return;
    }
    *inv = (uint32_t)std::min<double>(std::floor(std::ldexp(m, 32)), 4294967295.0);
    *sh = (uint32_t)s;
}

uint32_t threshold_of(double p) {
    const double v = std::floor(p * 4294967296.0);
    return v >= 4294967295.0 ? 0xFFFFFFFFu : v < 0 ? 0u : (uint32_t)v;
}

}  // namespace

extern "C" {

int gstim_flipsim_create(uint64_t batch_size, int disable_stabilizer_randomization, uint64_t num_qubits, uint64_t seed, int device,
                         gstim_flipsim **out) {
    return fs_guarded([&] {
        if (out == nullptr) {
            throw std::invalid_argument("NULL argument.");
        }
        *out = nullptr;
        if (batch_size == 0 || batch_size >= (1ull << 31)) {
            throw std::invalid_argument("batch_size must be between 1 and 2^31 - 1.");
        }
        if (num_qubits > 65535) {
            throw std::invalid_argument("Circuits with more than 65535 qubits are not supported by this build.");
        }
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
            cudaGetLastError();
            throw std::runtime_error("CUDA: no usable device: this library has no CPU fallback.");
        }
        if (device < 0 || device >= n) {
            throw std::invalid_argument("CUDA device ordinal out of range.");
        }
        auto h = std::make_unique<gstim_flipsim>();
        h->device = device;
        h->batch = (uint32_t)batch_size;
        h->W = (uint32_t)((batch_size + 127) / 128 * 4);
        h->randomize = !disable_stabilizer_randomization;
        h->seed = seed;
        ck(cudaSetDevice(device), "cudaSetDevice");
        ck(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate");
        ck(cudaMalloc(&h->FLAG, (size_t)h->W * 4), "cudaMalloc");
        ck(cudaMemsetAsync(h->FLAG, 0, (size_t)h->W * 4, h->stream), "memset");
        grow(h.get(), h->X, num_qubits);
        grow(h.get(), h->Z, num_qubits);
        if (h->randomize && num_qubits) {  // reset_all (frame_simulator.inl:153-163): z random
            fs_masked_flip<<<grid_for(num_qubits * h->W), 256, 0, h->stream>>>(state_of(h.get()), h->Z.p, nullptr, num_qubits * (uint64_t)h->W,
                                                                                0x80000000u, 0, h->ordinal++, 1);
            ck(cudaGetLastError(), "launch");
        }
        ck(cudaStreamSynchronize(h->stream), "sync");
        *out = h.release();
    });
}

int gstim_flipsim_copy(const gstim_flipsim *src, int copy_rng, uint64_t seed, gstim_flipsim **out) {
    return fs_guarded([&] {
        if (src == nullptr || out == nullptr) {
            throw std::invalid_argument("NULL argument.");
        }
        *out = nullptr;
        auto h = std::make_unique<gstim_flipsim>();
        h->device = src->device;
        h->batch = src->batch;
        h->W = src->W;
        h->randomize = src->randomize;
        h->seed = copy_rng ? src->seed : seed;
        h->ordinal = copy_rng ? src->ordinal : 0;
        ck(cudaSetDevice(h->device), "cudaSetDevice");
        ck(cudaStreamSynchronize(src->stream), "sync");
        ck(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate");
        ck(cudaMalloc(&h->FLAG, (size_t)h->W * 4), "cudaMalloc");
        ck(cudaMemcpyAsync(h->FLAG, src->FLAG, (size_t)h->W * 4, cudaMemcpyDeviceToDevice, h->stream), "copy");
        const Table *from[5] = {&src->X, &src->Z, &src->REC, &src->DET, &src->OBS};
        Table *to[5] = {&h->X, &h->Z, &h->REC, &h->DET, &h->OBS};
        for (int i = 0; i < 5; i++) {
            grow(h.get(), *to[i], from[i]->rows);
            if (from[i]->rows) {
                ck(cudaMemcpyAsync(to[i]->p, from[i]->p, from[i]->rows * (uint64_t)h->W * 4, cudaMemcpyDeviceToDevice, h->stream), "copy");
            }
        }
        ck(cudaStreamSynchronize(h->stream), "sync");
        *out = h.release();
    });
}

void gstim_flipsim_destroy(gstim_flipsim *h) {
    if (h) {
        cudaSetDevice(h->device);
        delete h;
    }
}

int gstim_flipsim_sizes(const gstim_flipsim *h, uint64_t *batch_size, uint64_t *num_qubits, uint64_t *num_measurements,
                        uint64_t *num_detectors, uint64_t *num_observables, uint64_t *row_words) {
    return fs_guarded([&] {
        if (h == nullptr) {
            throw std::invalid_argument("NULL simulator.");
        }
        if (batch_size) *batch_size = h->batch;
        if (num_qubits) *num_qubits = h->X.rows;
        if (num_measurements) *num_measurements = h->REC.rows;
        if (num_detectors) *num_detectors = h->DET.rows;
        if (num_observables) *num_observables = h->OBS.rows;
        if (row_words) *row_words = h->W;
    });
}

int gstim_flipsim_do_text(gstim_flipsim *h, const char *circuit_text, size_t text_len) {
    return fs_guarded([&] {
        if (h == nullptr || circuit_text == nullptr) {
            throw std::invalid_argument("NULL argument.");
        }
        ck(cudaSetDevice(h->device), "cudaSetDevice");
        Circuit c = Circuit::from_text(std::string_view(circuit_text, text_len));
        LoweredCircuit lc = lower_fragment(c, (uint32_t)h->X.rows, h->REC.rows);
        const uint64_t old_q = h->X.rows;
        grow(h, h->X, lc.num_qubits);
        grow(h, h->Z, lc.num_qubits);
        if (h->randomize && lc.num_qubits > old_q) {  // new qubits start like reset_all leaves them: z random
            const uint64_t nw = (lc.num_qubits - old_q) * (uint64_t)h->W;
            fs_masked_flip<<<grid_for(nw), 256, 0, h->stream>>>(state_of(h), h->Z.p + old_q * h->W, nullptr, nw, 0x80000000u, 0, h->ordinal++, 1);
            ck(cudaGetLastError(), "launch");
        }
        const uint32_t frag_dets = (uint32_t)lc.stats.num_detectors, det0 = (uint32_t)h->DET.rows;
        grow(h, h->REC, h->REC.rows + lc.stats.num_measurements);
        grow(h, h->DET, h->DET.rows + lc.stats.num_detectors);
        grow(h, h->OBS, std::max<uint64_t>(h->OBS.rows, lc.stats.num_observables));
        for (const Batch &B : lc.batches) {
            const FsState st = state_of(h);
            const uint32_t n = B.n_items, ord = h->ordinal++;
            switch (B.op) {
                case GOP_CLIFF1: {
                    const uint32_t *d = upload(h, B.payload);
                    fs_cliff1<<<grid_for((uint64_t)n * h->W), 256, 0, h->stream>>>(st, B.aux, n, d);
                    break;
                }
                case GOP_CLIFF2: {
                    const uint32_t *d = upload(h, B.payload);
                    fs_cliff2<<<grid_for((uint64_t)n * h->W), 256, 0, h->stream>>>(st, B.aux, n, d);
                    break;
                }
                case GOP_MEASURE: {
                    const uint32_t *d = upload(h, B.payload);
                    fs_measure<<<grid_for((uint64_t)n * h->W / 4), 256, 0, h->stream>>>(st, B.aux & 3, (B.aux >> 2) & 3, n, d, B.rec0,
                                                                                        h->randomize ? 1u : 0u, ord);
                    break;
                }
                case GOP_RECZERO:
                    fs_zero_rows<<<grid_for((uint64_t)n * h->W), 256, 0, h->stream>>>(h->REC.p + (size_t)B.rec0 * h->W, (uint64_t)n * h->W);
                    break;
                case GOP_XORROWS: {
                    std::vector<uint32_t> w;
                    w.insert(w.end(), B.dst.begin(), B.dst.end());
                    w.insert(w.end(), B.off.begin(), B.off.end());
                    w.insert(w.end(), B.idx.begin(), B.idx.end());
                    const uint32_t *d = upload(h, w);
                    const uint32_t nd = (uint32_t)B.dst.size();
                    fs_xorrows<<<grid_for((uint64_t)nd * h->W), 256, 0, h->stream>>>(st, nd, d, d + nd, d + nd + B.off.size(),
                                                                                    (B.flags & GF_ACCUM) ? 1u : 0u, frag_dets, det0);
                    break;
                }
                case GOP_OBS_PAULI: {
                    const uint32_t *d = upload(h, B.payload);
                    fs_obs_pauli<<<grid_for((uint64_t)n * h->W), 256, 0, h->stream>>>(st, n, d, frag_dets);
                    break;
                }
                case GOP_FEEDBACK: {
                    const uint32_t *d = upload(h, B.payload);
                    fs_feedback<<<grid_for((uint64_t)n * h->W), 256, 0, h->stream>>>(st, n, d);
                    break;
                }
                case GOP_NOISE1:
                case GOP_NOISE2: {
                    if (B.lambda == 0) {
                        break;
                    }
                    FsNoise nz{};
                    nz.op = B.op;
                    nz.flags = B.flags;
                    nz.aux = B.aux;
                    nz.extra = B.extra;
                    nz.t1 = B.t1;
                    nz.t2 = B.t2;
                    nz.t3 = B.t3;
                    nz.rec0 = B.rec0;
                    nz.n = n;
                    gap_params(B.lambda, &nz.inv, &nz.sh);
                    nz.items = upload(h, B.payload);
                    fs_noise<<<(n + 127) / 128, 128, 0, h->stream>>>(st, nz, ord);
                    break;
                }
                case GOP_CORR: {
                    const uint32_t *d = upload(h, B.payload);
                    const double p = B.lambda >= (1ull << 62) ? 1.0 : -std::expm1(-std::ldexp((double)B.lambda, -56));
                    fs_corr<<<grid_for(h->W), 256, 0, h->stream>>>(st, n, d, (B.flags & GF_RESET_FLAG) ? 1u : 0u, threshold_of(p),
                                                                  B.lambda >= (1ull << 62) ? 1u : 0u, ord);
                    break;
                }
                default:
                    break;
            }
            ck(cudaGetLastError(), "kernel launch");
        }
        ck(cudaStreamSynchronize(h->stream), "cudaStreamSynchronize");
    });
}

static Table *table_of(gstim_flipsim *h, int what) {
    switch (what) {
        case 0:
            return &h->X;
        case 1:
            return &h->Z;
        case 2:
            return &h->REC;
        case 3:
            return &h->DET;
        case 4:
            return &h->OBS;
        default:
            throw std::invalid_argument("bad table selector.");
    }
}

int gstim_flipsim_get_rows(gstim_flipsim *h, int what, uint64_t first_row, uint64_t n_rows, uint32_t *words_out) {
    return fs_guarded([&] {
        if (h == nullptr || (words_out == nullptr && n_rows)) {
            throw std::invalid_argument("NULL argument.");
        }
        Table *t = table_of(h, what);
        if (first_row + n_rows > t->rows) {
            throw std::out_of_range("row index out of range.");
        }
        ck(cudaSetDevice(h->device), "cudaSetDevice");
        if (n_rows) {
            ck(cudaMemcpy(words_out, t->p + first_row * h->W, n_rows * h->W * 4, cudaMemcpyDeviceToHost), "D2H");
        }
    });
}

int gstim_flipsim_set_rows(gstim_flipsim *h, int what, uint64_t first_row, uint64_t n_rows, const uint32_t *words_in, int xor_in) {
    return fs_guarded([&] {
        if (h == nullptr || (words_in == nullptr && n_rows)) {
            throw std::invalid_argument("NULL argument.");
        }
        Table *t = table_of(h, what);
        ck(cudaSetDevice(h->device), "cudaSetDevice");
        if (what == 2 && first_row == t->rows) {
            grow(h, *t, first_row + n_rows);  // append_measurement_flips
        } else if ((what == 0 || what == 1) && first_row + n_rows > t->rows) {
            grow(h, h->X, first_row + n_rows);
            grow(h, h->Z, first_row + n_rows);
        }
        if (first_row + n_rows > t->rows) {
            throw std::out_of_range("row index out of range.");
        }
        if (n_rows == 0) {
            return;
        }
        if (!xor_in) {
            ck(cudaMemcpy(t->p + first_row * h->W, words_in, n_rows * h->W * 4, cudaMemcpyHostToDevice), "H2D");
            return;
        }
        std::vector<uint32_t> w(words_in, words_in + n_rows * h->W);
        const uint32_t *d = upload(h, w);
        fs_masked_flip<<<grid_for(n_rows * h->W), 256, 0, h->stream>>>(state_of(h), t->p + first_row * h->W, d, n_rows * (uint64_t)h->W, 0, 1,
                                                                      h->ordinal++, 0);
        ck(cudaGetLastError(), "launch");
        ck(cudaStreamSynchronize(h->stream), "sync");
    });
}

int gstim_flipsim_broadcast(gstim_flipsim *h, int pauli, const uint32_t *mask_words, uint64_t n_rows, double p) {
    return fs_guarded([&] {
        if (h == nullptr || (mask_words == nullptr && n_rows)) {
            throw std::invalid_argument("NULL argument.");
        }
        if (pauli < 0 || pauli > 3) {
            throw std::invalid_argument("pauli must be 0 (I), 1 (X), 2 (Y) or 3 (Z).");
        }
        if (!(p >= 0 && p <= 1)) {
            throw std::invalid_argument("need 0 <= p <= 1");
        }
        ck(cudaSetDevice(h->device), "cudaSetDevice");
        grow(h, h->X, n_rows);
        grow(h, h->Z, n_rows);
        if (pauli == 0 || n_rows == 0 || p == 0) {
            return;
        }
        std::vector<uint32_t> w(mask_words, mask_words + n_rows * h->W);
        uint32_t *scratch = nullptr;
        const uint64_t nw = n_rows * (uint64_t)h->W;
        ck(cudaMalloc(&scratch, nw * 4), "cudaMalloc");
        const uint32_t *d = upload(h, w);
        // one Bernoulli(p) mask for the error, applied to x and / or z (Y flips both with the SAME coin)
        fs_masked_flip<<<grid_for(nw), 256, 0, h->stream>>>(state_of(h), scratch, d, nw, threshold_of(p), p >= 1 ? 1u : 0u, h->ordinal++, 1);
        if (pauli == 1 || pauli == 2) {
            fs_masked_flip<<<grid_for(nw), 256, 0, h->stream>>>(state_of(h), h->X.p, scratch, nw, 0, 1, 0, 0);
        }
        if (pauli == 3 || pauli == 2) {
            fs_masked_flip<<<grid_for(nw), 256, 0, h->stream>>>(state_of(h), h->Z.p, scratch, nw, 0, 1, 0, 0);
        }
        ck(cudaGetLastError(), "launch");
        ck(cudaStreamSynchronize(h->stream), "sync");
        cudaFree(scratch);
    });
}

int gstim_flipsim_bernoulli(gstim_flipsim *h, uint64_t n_words, double p, uint32_t *words_out) {
    return fs_guarded([&] {
        if (h == nullptr || (words_out == nullptr && n_words)) {
            throw std::invalid_argument("NULL argument.");
        }
        if (!(p >= 0 && p <= 1)) {
            throw std::invalid_argument("need 0 <= p <= 1");
        }
        if (n_words == 0) {
            return;
        }
        ck(cudaSetDevice(h->device), "cudaSetDevice");
        uint32_t *scratch = nullptr;
        ck(cudaMalloc(&scratch, n_words * 4), "cudaMalloc");
        fs_masked_flip<<<grid_for(n_words), 256, 0, h->stream>>>(state_of(h), scratch, nullptr, n_words, threshold_of(p), p >= 1 ? 1u : 0u,
                                                               h->ordinal++, 1);
        ck(cudaGetLastError(), "launch");
        ck(cudaMemcpyAsync(words_out, scratch, n_words * 4, cudaMemcpyDeviceToHost, h->stream), "D2H");
        ck(cudaStreamSynchronize(h->stream), "sync");
        cudaFree(scratch);
    });
}

int gstim_flipsim_clear(gstim_flipsim *h) {
    return fs_guarded([&] {
        if (h == nullptr) {
            throw std::invalid_argument("NULL simulator.");
        }
        ck(cudaSetDevice(h->device), "cudaSetDevice");
        h->REC.rows = 0;
        h->DET.rows = 0;
        h->OBS.rows = 0;
        for (Table *t : {&h->X, &h->Z, &h->REC, &h->DET, &h->OBS}) {
            if (t->p) {
                ck(cudaMemsetAsync(t->p, 0, t->cap * h->W * 4, h->stream), "memset");
            }
        }
        ck(cudaMemsetAsync(h->FLAG, 0, (size_t)h->W * 4, h->stream), "memset");
        if (h->randomize && h->X.rows) {
            fs_masked_flip<<<grid_for(h->X.rows * h->W), 256, 0, h->stream>>>(state_of(h), h->Z.p, nullptr, h->X.rows * (uint64_t)h->W, 0x80000000u,
                                                                             0, h->ordinal++, 1);
            ck(cudaGetLastError(), "launch");
        }
        ck(cudaStreamSynchronize(h->stream), "sync");
    });
}

}  // extern "C"
