// tableau_ref.cc — see tableau_ref.h. A from-scratch inverse-tableau stabilizer simulator.
//
// State |psi> = T|0..0>. Stored: for every qubit q the Pauli strings T^dag X_q T and T^dag Z_q T (bit-packed
// rows + sign). A gate U on the state (T <- U T) replaces the rows of its qubits by products of old rows
// (T^dag U^dag g U T); measuring Z_q is deterministic iff T^dag Z_q T has no X/Y component, in which case the
// outcome is that row's sign; otherwise the state is collapsed by Clifford operations on the INPUT side of T
// that fix |0..0>, with the random outcome forced to 0 like the reference's sign_bias = +1
// (/root/reference/src/stim/simulators/tableau_simulator.inl:1435-1438, :1370).
// Gate actions are the standard stabilizer generators of each gate (the "flow" table every Stim gate documents).
#include "tableau_ref.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>

#include "program.h"

namespace gstim {
namespace {

// single-qubit Pauli code: bit0 = x, bit1 = z  (0 I, 1 X, 2 Z, 3 Y)
inline int mul_phase(int a, int b) {
    // a * b = i^e (a ^ b)
    static const int8_t E[4][4] = {
        {0, 0, 0, 0},
        {0, 0, 3, 1},  // X*Z = -iY, X*Y = iZ
        {0, 1, 0, 3},  // Z*X = iY,  Z*Y = -iX
        {0, 3, 1, 0},  // Y*X = -iZ, Y*Z = iX
    };
    return E[a][b];
}

int code_of(char c) {
    return c == 'X' ? 1 : c == 'Z' ? 2 : c == 'Y' ? 3 : 0;
}

struct Table1 {
    uint8_t img[4], sgn[4];
};
struct Table2 {
    uint8_t img[16], sgn[16];  // code = p1 | p2 << 2
};

// From the generator images "+Y", "-Z" (image of X, image of Z).
Table1 make1(const char *fx, const char *fz) {
    Table1 t{};
    int px = code_of(fx[1]), pz = code_of(fz[1]);
    int sx = fx[0] == '-', sz = fz[0] == '-';
    t.img[0] = 0;
    t.sgn[0] = 0;
    t.img[1] = (uint8_t)px;
    t.sgn[1] = (uint8_t)sx;
    t.img[2] = (uint8_t)pz;
    t.sgn[2] = (uint8_t)sz;
    int e = 1 + mul_phase(px, pz) + 2 * (sx + sz);  // Y = i X Z
    t.img[3] = (uint8_t)(px ^ pz);
    t.sgn[3] = (uint8_t)((e >> 1) & 1);
    return t;
}

struct P2 {
    int a = 0, b = 0, e = 0;  // i^e * (a (x) b)
    void mul(const P2 &o) {
        e = (e + o.e + mul_phase(a, o.a) + mul_phase(b, o.b)) & 3;
        a ^= o.a;
        b ^= o.b;
    }
};

// From the generator images of X_, Z_, _X, _Z, e.g. "+XX", "+ZI", "+IX", "+ZZ" for CX.
Table2 make2(const char *f0, const char *f1, const char *f2, const char *f3) {
    const char *f[4] = {f0, f1, f2, f3};
    P2 g[4];
    for (int i = 0; i < 4; i++) {
        g[i].a = code_of(f[i][1]);
        g[i].b = code_of(f[i][2]);
        g[i].e = f[i][0] == '-' ? 2 : 0;
    }
    Table2 t{};
    for (int c = 0; c < 16; c++) {
        int p1 = c & 3, p2 = c >> 2;
        P2 acc;
        acc.e = (p1 == 3) + (p2 == 3);
        if (p1 & 1) acc.mul(g[0]);
        if (p1 & 2) acc.mul(g[1]);
        if (p2 & 1) acc.mul(g[2]);
        if (p2 & 2) acc.mul(g[3]);
        // acc = i^e (a (x) b) with Hermitian single-qubit Paulis: e must be even
        t.img[c] = (uint8_t)(acc.a | (acc.b << 2));
        t.sgn[c] = (uint8_t)((acc.e >> 1) & 1);
    }
    return t;
}

struct GateDef {
    const char *name, *inverse;
    const char *f[4];
};
// Forward conjugation U g U^dag of the generators (X, Z) or (X_, Z_, _X, _Z).
const GateDef DEFS1[] = {
    {"I", "I", {"+X", "+Z"}},
    {"X", "X", {"+X", "-Z"}},
    {"Y", "Y", {"-X", "-Z"}},
    {"Z", "Z", {"-X", "+Z"}},
    {"H", "H", {"+Z", "+X"}},
    {"H_XY", "H_XY", {"+Y", "-Z"}},
    {"H_YZ", "H_YZ", {"-X", "+Y"}},
    {"H_NXY", "H_NXY", {"-Y", "-Z"}},
    {"H_NXZ", "H_NXZ", {"-Z", "-X"}},
    {"H_NYZ", "H_NYZ", {"-X", "-Y"}},
    {"C_XYZ", "C_ZYX", {"+Y", "+X"}},
    {"C_NXYZ", "C_ZYNX", {"-Y", "-X"}},
    {"C_XNYZ", "C_ZNYX", {"-Y", "+X"}},
    {"C_XYNZ", "C_NZYX", {"+Y", "-X"}},
    {"C_ZYX", "C_XYZ", {"+Z", "+Y"}},
    {"C_ZYNX", "C_NXYZ", {"-Z", "+Y"}},
    {"C_ZNYX", "C_XNYZ", {"+Z", "-Y"}},
    {"C_NZYX", "C_XYNZ", {"-Z", "-Y"}},
    {"SQRT_X", "SQRT_X_DAG", {"+X", "-Y"}},
    {"SQRT_X_DAG", "SQRT_X", {"+X", "+Y"}},
    {"SQRT_Y", "SQRT_Y_DAG", {"-Z", "+X"}},
    {"SQRT_Y_DAG", "SQRT_Y", {"+Z", "-X"}},
    {"S", "S_DAG", {"+Y", "+Z"}},
    {"S_DAG", "S", {"-Y", "+Z"}},
};
const GateDef DEFS2[] = {
    {"II", "II", {"+XI", "+ZI", "+IX", "+IZ"}},
    {"CX", "CX", {"+XX", "+ZI", "+IX", "+ZZ"}},
    {"CY", "CY", {"+XY", "+ZI", "+ZX", "+ZZ"}},
    {"CZ", "CZ", {"+XZ", "+ZI", "+ZX", "+IZ"}},
    {"XCX", "XCX", {"+XI", "+ZX", "+IX", "+XZ"}},
    {"XCY", "XCY", {"+XI", "+ZY", "+XX", "+XZ"}},
    {"XCZ", "XCZ", {"+XI", "+ZZ", "+XX", "+IZ"}},
    {"YCX", "YCX", {"+XX", "+ZX", "+IX", "+YZ"}},
    {"YCY", "YCY", {"+XY", "+ZY", "+YX", "+YZ"}},
    {"YCZ", "YCZ", {"+XZ", "+ZZ", "+YX", "+IZ"}},
    {"SWAP", "SWAP", {"+IX", "+IZ", "+XI", "+ZI"}},
    {"ISWAP", "ISWAP_DAG", {"+ZY", "+IZ", "+YZ", "+ZI"}},
    {"ISWAP_DAG", "ISWAP", {"-ZY", "+IZ", "-YZ", "+ZI"}},
    {"CXSWAP", "SWAPCX", {"+XX", "+IZ", "+XI", "+ZZ"}},
    {"SWAPCX", "CXSWAP", {"+IX", "+ZZ", "+XX", "+ZI"}},
    {"CZSWAP", "CZSWAP", {"+ZX", "+IZ", "+XZ", "+ZI"}},
    {"SQRT_XX", "SQRT_XX_DAG", {"+XI", "-YX", "+IX", "-XY"}},
    {"SQRT_XX_DAG", "SQRT_XX", {"+XI", "+YX", "+IX", "+XY"}},
    {"SQRT_YY", "SQRT_YY_DAG", {"-ZY", "+XY", "-YZ", "+YX"}},
    {"SQRT_YY_DAG", "SQRT_YY", {"+ZY", "-XY", "+YZ", "-YX"}},
    {"SQRT_ZZ", "SQRT_ZZ_DAG", {"+YZ", "+ZI", "+ZY", "+IZ"}},
    {"SQRT_ZZ_DAG", "SQRT_ZZ", {"-YZ", "+ZI", "-ZY", "+IZ"}},
};

struct Tables {
    std::map<std::string, Table1> fwd1, inv1;  // inv = conjugation by the inverse gate: U^dag P U
    std::map<std::string, Table2> fwd2, inv2;
    Tables() {
        for (const auto &d : DEFS1) {
            fwd1[d.name] = make1(d.f[0], d.f[1]);
        }
        for (const auto &d : DEFS1) {
            inv1[d.name] = fwd1.at(d.inverse);
        }
        for (const auto &d : DEFS2) {
            fwd2[d.name] = make2(d.f[0], d.f[1], d.f[2], d.f[3]);
        }
        for (const auto &d : DEFS2) {
            inv2[d.name] = fwd2.at(d.inverse);
        }
    }
};
const Tables &tables() {
    static const Tables t;
    return t;
}

struct Sim {
    size_t n, nw;
    std::vector<uint64_t> X, Z;   // rows: 2q = T^dag X_q T, 2q+1 = T^dag Z_q T ; each nw words
    std::vector<uint8_t> sign;
    std::vector<uint8_t> record;

    // Words [lo[r], hi[r]) contain every non-zero word of row r (a superset is fine): the rows of a QEC circuit have local
    // support, a product only needs the words of its right-hand side (d = 51: 3-5 of 83). Not used by the step-by-step
    // path (GSTIM_TABLEAU_SLOW=1), which the tests compare against.
    std::vector<uint32_t> lo, hi;

    explicit Sim(size_t n)
        : n(n), nw((n + 63) / 64 + 1), X(2 * n * nw, 0), Z(2 * n * nw, 0), sign(2 * n, 0), lo(2 * n, 0), hi(2 * n, 0) {
        for (size_t q = 0; q < n; q++) {
            X[(2 * q) * nw + q / 64] |= 1ull << (q % 64);
            Z[(2 * q + 1) * nw + q / 64] |= 1ull << (q % 64);
            lo[2 * q] = lo[2 * q + 1] = (uint32_t)(q / 64);
            hi[2 * q] = hi[2 * q + 1] = (uint32_t)(q / 64 + 1);
        }
    }
    void touch(size_t r, size_t w) {
        lo[r] = std::min<uint32_t>(lo[r], (uint32_t)w);
        hi[r] = std::max<uint32_t>(hi[r], (uint32_t)w + 1);
    }
    void rescan(size_t r) {
        const uint64_t *x = &X[r * nw], *z = &Z[r * nw];
        size_t a = 0, b = nw;
        while (a < nw && (x[a] | z[a]) == 0) a++;
        while (b > a && (x[b - 1] | z[b - 1]) == 0) b--;
        lo[r] = (uint32_t)(a < b ? a : 0);
        hi[r] = (uint32_t)(a < b ? b : 0);
    }
    uint64_t *xr(size_t r) { return &X[r * nw]; }
    uint64_t *zr(size_t r) { return &Z[r * nw]; }

    // acc <- acc * row r ; returns the i-exponent contributed (including the row's sign)
    int mul_into(uint64_t *ax, uint64_t *az, size_t r) {
        const uint64_t *bx = xr(r), *bz = zr(r);
        long plus = 0, minus = 0;
        const size_t w0 = slow_column_ops ? 0 : lo[r], w1 = slow_column_ops ? nw : hi[r];
        for (size_t w = w0; w < w1; w++) {
            uint64_t xa = ax[w], za = az[w], xb = bx[w], zb = bz[w];
            uint64_t p = (xa & ~za & xb & zb) | (xa & za & ~xb & zb) | (~xa & za & xb & ~zb);
            uint64_t m = (xa & za & xb & ~zb) | (~xa & za & xb & zb) | (xa & ~za & ~xb & zb);
            plus += __builtin_popcountll(p);
            minus += __builtin_popcountll(m);
            ax[w] = xa ^ xb;
            az[w] = za ^ zb;
        }
        return (int)(((plus - minus) % 4 + 4 + 2 * sign[r]) & 3);
    }

    // row for sgn * (Pauli code pa on qubit a) (x) (code pb on qubit b), expressed through the current rows
    void image_row(int pa, size_t a, int pb, size_t b, int sgn, uint64_t *ox, uint64_t *oz, uint8_t &os) {
        memset(ox, 0, nw * 8);
        memset(oz, 0, nw * 8);
        int e = 2 * sgn + (pa == 3) + (pb == 3);
        if (pa & 1) e += mul_into(ox, oz, 2 * a);
        if (pa & 2) e += mul_into(ox, oz, 2 * a + 1);
        if (pb & 1) e += mul_into(ox, oz, 2 * b);
        if (pb & 2) e += mul_into(ox, oz, 2 * b + 1);
        if (e & 1) {
            throw std::logic_error("internal: non-Hermitian row in the inverse tableau");
        }
        os = (uint8_t)((e >> 1) & 1);
    }

    // Scratch rows for the generator images of one gate (no allocation per gate: a d=51 memory experiment applies
    // 6.6e5 gates). Generators that the gate maps to themselves are skipped (CX: Z of the control, X of the target).
    std::vector<uint64_t> scratch;
    uint64_t *sx(int g) { return &scratch[(size_t)(2 * g) * nw]; }
    uint64_t *sz(int g) { return &scratch[(size_t)(2 * g + 1) * nw]; }

    void gate1(const Table1 &t, size_t q) {
        if (!slow_column_ops && t.img[1] == 2 && t.img[2] == 1 && !t.sgn[1] && !t.sgn[2]) {
            // H: the two generator images trade places (a third of a surface-code circuit's gates)
            const size_t a = std::min(lo[2 * q], lo[2 * q + 1]), b = std::max(hi[2 * q], hi[2 * q + 1]);
            std::swap_ranges(xr(2 * q) + a, xr(2 * q) + b, xr(2 * q + 1) + a);
            std::swap_ranges(zr(2 * q) + a, zr(2 * q) + b, zr(2 * q + 1) + a);
            std::swap(sign[2 * q], sign[2 * q + 1]);
            std::swap(lo[2 * q], lo[2 * q + 1]);
            std::swap(hi[2 * q], hi[2 * q + 1]);
            return;
        }
        scratch.resize(8 * nw);
        uint8_t ns[2];
        bool changed[2];
        for (int g = 0; g < 2; g++) {
            int code = g == 0 ? 1 : 2;
            changed[g] = !(t.img[code] == code && t.sgn[code] == 0);
            if (changed[g]) {
                image_row(t.img[code], q, 0, q, t.sgn[code], sx(g), sz(g), ns[g]);
            }
        }
        for (int g = 0; g < 2; g++) {
            if (changed[g]) {
                memcpy(xr(2 * q + g), sx(g), nw * 8);
                memcpy(zr(2 * q + g), sz(g), nw * 8);
                sign[2 * q + g] = ns[g];
                rescan(2 * q + g);
            }
        }
    }
    void gate2(const Table2 &t, size_t a, size_t b) {
        scratch.resize(8 * nw);
        uint8_t ns[4];
        bool changed[4];
        const int codes[4] = {1, 2, 4, 8};  // X_, Z_, _X, _Z
        const size_t rows_of[4] = {2 * a, 2 * a + 1, 2 * b, 2 * b + 1};
        for (int g = 0; g < 4; g++) {
            int img = t.img[codes[g]];
            changed[g] = !(img == codes[g] && t.sgn[codes[g]] == 0);
        }
        // In-place fast path (CX, CZ, ...): a generator whose image is itself times ONE generator of the other qubit
        // that the gate leaves alone. The two rows commute (they are images of commuting generators), so the order of
        // the product does not matter and row(g) *= row(h) needs no scratch copy.
        if (!slow_column_ops) {
            for (int g = 0; g < 4; g++) {
                if (!changed[g] || t.sgn[codes[g]] != 0) {
                    continue;
                }
                const int other = t.img[codes[g]] ^ codes[g];  // the image without g itself
                int h = -1;
                for (int k = 0; k < 4; k++) {
                    if (other == codes[k]) {
                        h = k;
                    }
                }
                if ((t.img[codes[g]] & codes[g]) != codes[g] || h < 0 || (h >> 1) == (g >> 1) || changed[h]) {
                    continue;
                }
                const int e = 2 * sign[rows_of[g]] + mul_into(xr(rows_of[g]), zr(rows_of[g]), rows_of[h]);
                if (e & 1) {
                    throw std::logic_error("internal: non-Hermitian row in the inverse tableau");
                }
                sign[rows_of[g]] = (uint8_t)((e >> 1) & 1);
                if (hi[rows_of[h]] > lo[rows_of[h]]) {
                    touch(rows_of[g], lo[rows_of[h]]);
                    touch(rows_of[g], hi[rows_of[h]] - 1);
                }
                changed[g] = false;
            }
        }
        for (int g = 0; g < 4; g++) {
            int img = t.img[codes[g]];
            if (changed[g]) {
                image_row(img & 3, a, img >> 2, b, t.sgn[codes[g]], sx(g), sz(g), ns[g]);
            }
        }
        const size_t rows[4] = {2 * a, 2 * a + 1, 2 * b, 2 * b + 1};
        for (int g = 0; g < 4; g++) {
            if (changed[g]) {
                memcpy(xr(rows[g]), sx(g), nw * 8);
                memcpy(zr(rows[g]), sz(g), nw * 8);
                sign[rows[g]] = ns[g];
                rescan(rows[g]);
            }
        }
    }
    void gate1(const std::string &name, size_t q) {
        gate1(tables().inv1.at(name), q);
    }
    void gate2(const std::string &name, size_t a, size_t b) {
        gate2(tables().inv2.at(name), a, b);
    }
    void pauli(int code, size_t q) {  // X: 1, Z: 2, Y: 3 applied to the state
        if (tmode) {
            const size_t rz = phys_row(2 * q + 1), rx = phys_row(2 * q);
            if (code & 1) ST[rz / 64] ^= 1ull << (rz % 64);
            if (code & 2) ST[rx / 64] ^= 1ull << (rx % 64);
            return;
        }
        if (code & 1) sign[2 * q + 1] ^= 1;  // X flips the sign of T^dag Z_q T
        if (code & 2) sign[2 * q] ^= 1;
    }

    // ---- transposed working copy for runs of random measurements ------------------------------------------------
    // A random measurement applies a handful of input-side column operations to ALL 2n rows; in the row-major layout
    // every one of them touches 2n cache lines (d = 51: 10 402 rows, 0.6 ms per collapse, 5 201 collapses). An instruction
    // with many targets (RX on every data qubit, the first round's MR) therefore switches to a column-major copy — XT[j] /
    // ZT[j] = the bits of column j over all rows, ST = the signs as one bit vector — in which a column operation is a
    // word-parallel pass over ~2n/64 words, evaluated generically from the gate's (image, sign) table by minterms. The only
    // output-side gate needed while the copy is live is H (X-basis targets): it swaps the adjacent rows 2q, 2q+1.
    // The copy is transposed back at the end of the instruction. Same operations in the same order as the
    // step-by-step path (GSTIM_TABLEAU_SLOW=1), which the tests compare against.
    bool tmode = false, t_allowed = false;
    bool t_swapped = false;  // rows 2 t_swap_q and 2 t_swap_q + 1 are to be read as exchanged (a pending H)
    size_t t_swap_q = 0;
    size_t phys_row(size_t r) const {
        return (t_swapped && r / 2 == t_swap_q) ? (r ^ 1) : r;
    }
    // row-major cache of one 64-row block of the column-major copy (word t_blk of every column): consecutive targets sit
    // in the same block, and a collapse changes only the columns it touches
    size_t t_blk = SIZE_MAX;
    std::vector<uint64_t> t_bx, t_bz;
    void load_block(size_t blk) {
        t_bx.resize(n);
        t_bz.resize(n);
        for (size_t j = 0; j < n; j++) {
            t_bx[j] = XT[j * rw + blk];
            t_bz[j] = ZT[j * rw + blk];
        }
        t_blk = blk;
    }
    void refresh_col(size_t j) {
        if (t_blk != SIZE_MAX) {
            t_bx[j] = XT[j * rw + t_blk];
            t_bz[j] = ZT[j * rw + t_blk];
        }
    }
    int t_policy = 1;  // 0 never, 1 automatic, 2 at the first random measurement (tests)
    size_t t_randoms = 0, t_remaining = 0;
    size_t rw = 0;
    std::vector<uint64_t> XT, ZT, ST;

    bool want_transposed() {
        if (!t_allowed || t_policy == 0) {
            return false;
        }
        if (t_policy == 2) {
            return true;
        }
        // the two transpositions cost about as much as twenty collapse passes
        return ++t_randoms >= 4 && t_remaining >= 32 && n >= 256;
    }
    static void transpose64(uint64_t *a) {  // bit c of a[r] <-> bit r of a[c]
        uint64_t m = 0x00000000FFFFFFFFull;
        for (unsigned j = 32; j != 0; j >>= 1, m ^= m << j) {
            for (unsigned k = 0; k < 64; k = (k + j + 1) & ~j) {
                const uint64_t t = ((a[k] >> j) ^ a[k + j]) & m;
                a[k] ^= t << j;
                a[k + j] ^= t;
            }
        }
    }
    void transpose_to(const std::vector<uint64_t> &rows, std::vector<uint64_t> &cols) {
        cols.assign(nw * 64 * rw, 0);
        uint64_t a[64];
        for (size_t rb = 0; rb < rw; rb++) {
            for (size_t w = 0; w < nw; w++) {
                bool any = false;
                for (size_t i = 0; i < 64; i++) {
                    const size_t r = rb * 64 + i;
                    a[i] = r < 2 * n ? rows[r * nw + w] : 0;
                    any |= a[i] != 0;
                }
                if (!any) {
                    continue;
                }
                transpose64(a);
                for (size_t c = 0; c < 64; c++) {
                    cols[(w * 64 + c) * rw + rb] = a[c];
                }
            }
        }
    }
    void transpose_from(const std::vector<uint64_t> &cols, std::vector<uint64_t> &rows) {
        uint64_t a[64];
        for (size_t rb = 0; rb < rw; rb++) {
            for (size_t w = 0; w < nw; w++) {
                for (size_t c = 0; c < 64; c++) {
                    a[c] = cols[(w * 64 + c) * rw + rb];
                }
                transpose64(a);
                for (size_t i = 0; i < 64; i++) {
                    const size_t r = rb * 64 + i;
                    if (r < 2 * n) {
                        rows[r * nw + w] = a[i];
                    }
                }
            }
        }
    }
    void enter_t() {
        rw = (2 * n + 63) / 64;
        transpose_to(X, XT);
        transpose_to(Z, ZT);
        ST.assign(rw, 0);
        for (size_t r = 0; r < 2 * n; r++) {
            ST[r / 64] |= (uint64_t)(sign[r] & 1) << (r % 64);
        }
        t_blk = SIZE_MAX;
        t_swapped = false;
        tmode = true;
    }
    void exit_t() {
        if (!tmode) {
            return;
        }
        if (t_swapped) {
            swap_rows_t(t_swap_q);
            t_swapped = false;
        }
        transpose_from(XT, X);
        transpose_from(ZT, Z);
        for (size_t r = 0; r < 2 * n; r++) {
            sign[r] = (uint8_t)((ST[r / 64] >> (r % 64)) & 1);
            rescan(r);
        }
        tmode = false;
    }
    bool bit_t(const std::vector<uint64_t> &m, size_t col, size_t row) const {
        return (m[col * rw + row / 64] >> (row % 64)) & 1;
    }
    void swap_rows_t(size_t q) {  // H on qubit q: T^dag X_q T <-> T^dag Z_q T
        const size_t w = (2 * q) / 64, b = (2 * q) % 64;  // (2q is even: both rows sit in the same word)
        auto swap_in = [&](uint64_t &v) {
            const uint64_t t = ((v >> b) ^ (v >> (b + 1))) & 1ull;
            v ^= (t << b) | (t << (b + 1));
        };
        for (size_t j = 0; j < n; j++) {
            swap_in(XT[j * rw + w]);
            swap_in(ZT[j * rw + w]);
        }
        swap_in(ST[w]);
        if (w == t_blk) {
            t_blk = SIZE_MAX;
        }
    }
    // every row R <- C^dag R C for a two-column / one-column operation given by its table, all rows at once
    void col2_t(const Table2 &t, size_t k, size_t j) {
        int codes[16], nc = 0;
        for (int c = 1; c < 16; c++) {
            if (t.img[c] != c || t.sgn[c]) {
                codes[nc++] = c;
            }
        }
        uint64_t *xk = &XT[k * rw], *zk = &ZT[k * rw], *xj = &XT[j * rw], *zj = &ZT[j * rw];
        for (size_t w = 0; w < rw; w++) {
            const uint64_t v[4] = {xk[w], zk[w], xj[w], zj[w]};
            if ((v[0] | v[1] | v[2] | v[3]) == 0) {
                continue;
            }
            uint64_t d[4] = {0, 0, 0, 0}, ds = 0;
            for (int i = 0; i < nc; i++) {
                const int c = codes[i];
                uint64_t m = ~0ull;
                for (int b = 0; b < 4; b++) {
                    m &= ((c >> b) & 1) ? v[b] : ~v[b];
                }
                if (m == 0) {
                    continue;
                }
                const int delta = t.img[c] ^ c;
                for (int b = 0; b < 4; b++) {
                    if ((delta >> b) & 1) {
                        d[b] |= m;
                    }
                }
                if (t.sgn[c]) {
                    ds |= m;
                }
            }
            xk[w] = v[0] ^ d[0];
            zk[w] = v[1] ^ d[1];
            xj[w] = v[2] ^ d[2];
            zj[w] = v[3] ^ d[3];
            ST[w] ^= ds;
        }
    }
    void col1_t(const Table1 &t, size_t k) {
        uint64_t *xk = &XT[k * rw], *zk = &ZT[k * rw];
        for (size_t w = 0; w < rw; w++) {
            const uint64_t x = xk[w], z = zk[w];
            uint64_t dx = 0, dz = 0, ds = 0;
            for (int c = 1; c < 4; c++) {
                const uint64_t m = ((c & 1) ? x : ~x) & ((c & 2) ? z : ~z);
                const int delta = t.img[c] ^ c;
                if (delta & 1) dx |= m;
                if (delta & 2) dz |= m;
                if (t.sgn[c]) ds |= m;
            }
            xk[w] = x ^ dx;
            zk[w] = z ^ dz;
            ST[w] ^= ds;
        }
    }
    bool measure_z_t(size_t q) {
        const size_t r = phys_row(2 * q + 1);
        if (r / 64 != t_blk) {
            load_block(r / 64);
        }
        const unsigned rb = (unsigned)(r % 64);
        size_t k = SIZE_MAX;
        std::vector<size_t> &cols = t_cols;
        cols.clear();
        for (size_t j = 0; j < n; j++) {
            if ((t_bx[j] >> rb) & 1) {
                if (k == SIZE_MAX) {
                    k = j;
                } else {
                    cols.push_back(j);
                }
            }
        }
        if (k == SIZE_MAX) {
            return bit_t(ST, 0, r);
        }
        const Tables &tb = tables();
        const Table2 &cx = tb.inv2.at("CX"), &cz = tb.inv2.at("CZ");
        for (size_t j : cols) {
            col2_t(cx, k, j);
            refresh_col(j);
        }
        refresh_col(k);
        cols.clear();
        for (size_t j = 0; j < n; j++) {
            if (j != k && ((t_bz[j] >> rb) & 1)) {
                cols.push_back(j);
            }
        }
        for (size_t j : cols) {
            col2_t(cz, k, j);
            refresh_col(j);
        }
        if (bit_t(ZT, k, r)) {
            col1_t(tb.inv1.at("S"), k);
        }
        col1_t(tb.inv1.at("H"), k);
        if (bit_t(ST, 0, r)) {
            col1_t(tb.inv1.at("X"), k);
        }
        refresh_col(k);
        return false;
    }
    std::vector<size_t> t_cols;

    // ---- input-side column operations: T <- T C, every row R <- C^dag R C -----------------------------------
    void col1(const Table1 &t, size_t k) {
        const size_t w = k / 64;
        const uint64_t bit = 1ull << (k % 64);
        for (size_t r = 0; r < 2 * n; r++) {
            uint64_t *x = xr(r), *z = zr(r);
            int c = ((x[w] & bit) ? 1 : 0) | ((z[w] & bit) ? 2 : 0);
            if (!c) continue;
            int d = t.img[c];
            x[w] = (x[w] & ~bit) | ((d & 1) ? bit : 0);
            z[w] = (z[w] & ~bit) | ((d & 2) ? bit : 0);
            sign[r] ^= t.sgn[c];
        }
    }
    void col2(const Table2 &t, size_t k, size_t j) {
        const size_t wk = k / 64, wj = j / 64;
        const uint64_t bk = 1ull << (k % 64), bj = 1ull << (j % 64);
        for (size_t r = 0; r < 2 * n; r++) {
            uint64_t *x = xr(r), *z = zr(r);
            int c = ((x[wk] & bk) ? 1 : 0) | ((z[wk] & bk) ? 2 : 0) | ((x[wj] & bj) ? 4 : 0) | ((z[wj] & bj) ? 8 : 0);
            if (!c) continue;
            int d = t.img[c];
            x[wk] = (x[wk] & ~bk) | ((d & 1) ? bk : 0);
            z[wk] = (z[wk] & ~bk) | ((d & 2) ? bk : 0);
            x[wj] = (x[wj] & ~bj) | ((d & 4) ? bj : 0);
            z[wj] = (z[wj] & ~bj) | ((d & 8) ? bj : 0);
            sign[r] ^= t.sgn[c];
        }
    }

    // Z-basis measurement of qubit q; a random outcome is forced to 0.
    bool measure_z(size_t q) {
        if (tmode) {
            return measure_z_t(q);
        }
        const size_t r = 2 * q + 1;
        size_t k = SIZE_MAX;
        for (size_t w = slow_column_ops ? 0 : lo[r]; w < (slow_column_ops ? nw : hi[r]) && k == SIZE_MAX; w++) {
            if (xr(r)[w]) {
                k = w * 64 + (size_t)__builtin_ctzll(xr(r)[w]);
            }
        }
        if (k == SIZE_MAX) {
            return sign[r] != 0;
        }
        const Tables &tb = tables();
        if (!slow_column_ops) {
            if (want_transposed()) {
                enter_t();
                return measure_z_t(q);
            }
            collapse_fused(r, k);
            return false;
        }
        // fold the X part of the row onto column k (input-side CX with control k fixes |0..0>)
        for (size_t j = 0; j < n; j++) {
            if (j != k && (xr(r)[j / 64] >> (j % 64)) & 1) {
                col2(tb.inv2.at("CX"), k, j);
            }
        }
        // clear the Z part (input-side CZ / S fix |0..0>)
        for (size_t j = 0; j < n; j++) {
            if (j != k && (zr(r)[j / 64] >> (j % 64)) & 1) {
                col2(tb.inv2.at("CZ"), k, j);
            }
        }
        if ((zr(r)[k / 64] >> (k % 64)) & 1) {
            col1(tb.inv1.at("S"), k);
        }
        // row r is now +-X_k: collapse with an input-side H (and X for the sign) so that the outcome is 0
        col1(tb.inv1.at("H"), k);
        if (sign[r]) {
            col1(tb.inv1.at("X"), k);
        }
        return false;
    }
    // The same column operations as the loop in measure_z, applied in ONE pass over the rows: the list of
    // (gate, column) steps is fixed by row r before anything changes, every step acts on the pivot column k and at most
    // one other column, and a row's bits in those columns are all that the steps read or write. Each full pass over the
    // 2n rows of a d=51 circuit moves 2.7 MB, and a random measurement needs half a dozen of them when done one by one.
    bool slow_column_ops = false;
    struct Step {
        const Table2 *t2;
        const Table1 *t1;
        size_t j;
    };
    std::vector<Step> steps;
    void collapse_fused(size_t r, size_t k) {
        const Tables &tb = tables();
        steps.clear();
        const Table2 *cx = &tb.inv2.at("CX"), *cz = &tb.inv2.at("CZ");
        for (size_t j = 0; j < n; j++) {
            if (j != k && (xr(r)[j / 64] >> (j % 64)) & 1) {
                steps.push_back({cx, nullptr, j});
            }
        }
        // the CX steps do not change the Z part of row r outside column k (CX(k, j): z_k ^= z_j), so the CZ list can be
        // read off the row as it is now; whether S is needed depends on the updated z_k, which the pass tracks on row r
        for (size_t j = 0; j < n; j++) {
            if (j != k && (zr(r)[j / 64] >> (j % 64)) & 1) {
                steps.push_back({cz, nullptr, j});
            }
        }
        // simulate the steps on row r alone to decide the trailing single-qubit steps
        auto run = [&](size_t row, size_t upto, int &ck, uint8_t &sg) {
            const size_t wk = k / 64;
            const uint64_t bk = 1ull << (k % 64);
            uint64_t *x = xr(row), *z = zr(row);
            ck = ((x[wk] & bk) ? 1 : 0) | ((z[wk] & bk) ? 2 : 0);
            for (size_t s = 0; s < upto; s++) {
                const Step &st = steps[s];
                if (st.t2 != nullptr) {
                    const size_t wj = st.j / 64;
                    const uint64_t bj = 1ull << (st.j % 64);
                    const int cj = ((x[wj] & bj) ? 1 : 0) | ((z[wj] & bj) ? 2 : 0);
                    const int c = ck | (cj << 2);
                    if (c) {
                        const int d = st.t2->img[c];
                        sg ^= st.t2->sgn[c];
                        ck = d & 3;
                        const int dj = d >> 2;
                        x[wj] = (x[wj] & ~bj) | ((dj & 1) ? bj : 0);
                        z[wj] = (z[wj] & ~bj) | ((dj & 2) ? bj : 0);
                        if (dj) {
                            touch(row, wj);
                        }
                    }
                } else if (ck) {
                    sg ^= st.t1->sgn[ck];
                    ck = st.t1->img[ck];
                }
            }
            x[wk] = (x[wk] & ~bk) | ((ck & 1) ? bk : 0);
            z[wk] = (z[wk] & ~bk) | ((ck & 2) ? bk : 0);
            if (ck) {
                touch(row, wk);
            }
        };
        // row r first (it decides S and X), then every other row with the complete list
        const size_t n2 = steps.size();
        int ck;
        run(r, n2, ck, sign[r]);
        size_t first_tail = steps.size();
        if (ck & 2) {  // z_k still set: S
            steps.push_back({nullptr, &tb.inv1.at("S"), k});
        }
        steps.push_back({nullptr, &tb.inv1.at("H"), k});
        {
            // continue row r through the tail to learn its sign before the optional X
            const size_t wk = k / 64;
            const uint64_t bk = 1ull << (k % 64);
            int c = ((xr(r)[wk] & bk) ? 1 : 0) | ((zr(r)[wk] & bk) ? 2 : 0);
            uint8_t sg = sign[r];
            for (size_t s = first_tail; s < steps.size(); s++) {
                if (c) {
                    sg ^= steps[s].t1->sgn[c];
                    c = steps[s].t1->img[c];
                }
            }
            if (sg) {
                steps.push_back({nullptr, &tb.inv1.at("X"), k});
            }
        }
        // row r: only the tail is left; all other rows: everything
        {
            const size_t wk = k / 64;
            const uint64_t bk = 1ull << (k % 64);
            int c = ((xr(r)[wk] & bk) ? 1 : 0) | ((zr(r)[wk] & bk) ? 2 : 0);
            for (size_t s = first_tail; s < steps.size(); s++) {
                if (c) {
                    sign[r] ^= steps[s].t1->sgn[c];
                    c = steps[s].t1->img[c];
                }
            }
            xr(r)[wk] = (xr(r)[wk] & ~bk) | ((c & 1) ? bk : 0);
            zr(r)[wk] = (zr(r)[wk] & ~bk) | ((c & 2) ? bk : 0);
            touch(r, wk);
        }
        for (size_t row = 0; row < 2 * n; row++) {
            if (row != r) {
                int c;
                run(row, steps.size(), c, sign[row]);
            }
        }
    }

    void to_z_basis(uint32_t basis, size_t q) {
        if (tmode && basis == GB_Y) {
            exit_t();  // (never entered for Y-basis instructions; kept for safety)
        }
        if (tmode) {
            if (basis == GB_X) {  // H on q = the two rows of q trade places: done by renaming while the copy is live
                if (t_swapped && t_swap_q != q) {
                    swap_rows_t(t_swap_q);
                    t_swapped = false;
                }
                t_swapped = !t_swapped;
                t_swap_q = q;
            }
            return;
        }
        if (basis == GB_X) {
            gate1("H", q);
        } else if (basis == GB_Y) {
            gate1("H_YZ", q);
        }
    }
    bool measure(uint32_t basis, size_t q) {
        to_z_basis(basis, q);
        bool m = measure_z(q);
        to_z_basis(basis, q);
        return m;
    }
    void reset(uint32_t basis, size_t q) {
        to_z_basis(basis, q);
        if (measure_z(q)) {
            pauli(1, q);
        }
        to_z_basis(basis, q);
    }
};

struct Product {
    std::vector<std::pair<uint32_t, uint32_t>> terms;  // (qubit, xz code with bit0 = x, bit1 = z)
    std::vector<uint32_t> bits;
    bool sign = false;
};

std::vector<Product> read_products(const Instruction &op) {
    std::vector<Product> out;
    const auto &ts = op.targets;
    size_t k = 0;
    while (k < ts.size()) {
        size_t end = k + 1;
        while (end < ts.size() && ts[end] == T_COMBINER) {
            end += 2;
        }
        Product p;
        std::map<uint32_t, uint32_t> acc;
        int e = 0;
        for (size_t j = k; j < end; j += 2) {
            uint32_t t = ts[j];
            if (t & (T_REC | T_SWEEP)) {
                p.bits.push_back(t);
                continue;
            }
            if (t & T_INVERTED) {
                p.sign = !p.sign;
            }
            uint32_t q = t & T_VALUE_MASK;
            int c = ((t & T_PAULI_X) ? 1 : 0) | ((t & T_PAULI_Z) ? 2 : 0);
            e += mul_phase((int)acc[q], c);
            acc[q] ^= (uint32_t)c;
        }
        if (e & 1) {
            throw std::invalid_argument(std::string("Acted on an anti-Hermitian operator (e.g. X0*Z0 instead of Y0) in ") + op.gate->name + ".");
        }
        if ((e >> 1) & 1) {
            p.sign = !p.sign;
        }
        for (auto &kv : acc) {
            if (kv.second) {
                p.terms.push_back({kv.first, kv.second});
            }
        }
        out.push_back(std::move(p));
        k = end;
    }
    return out;
}

}  // namespace

std::vector<uint8_t> reference_sample(const Circuit &circuit) {
    CircuitStats stats = compute_stats(circuit);
    Sim sim(std::max<size_t>(stats.num_qubits, 1));
    {
        const char *e = getenv("GSTIM_TABLEAU_SLOW");  // differential testing of the fused column pass
        sim.slow_column_ops = e != nullptr && e[0] == '1';
        const char *t = getenv("GSTIM_TABLEAU_TRANSPOSE");  // off | force (tests); default: automatic
        sim.t_policy = t == nullptr ? 1 : std::string(t) == "off" ? 0 : std::string(t) == "force" ? 2 : 1;
    }
    auto rec_value = [&](uint32_t t, const char *gate) -> bool {
        uint64_t k = t & T_VALUE_MASK;
        if (k == 0 || k > sim.record.size()) {
            throw std::out_of_range(std::string("Referred to a measurement record before the beginning of time in ") + gate + ".");
        }
        return sim.record[sim.record.size() - k] != 0;
    };
    circuit.for_each_operation([&](const Instruction &op) {
        const GateInfo &g = *op.gate;
        const std::string name = g.name;
        switch (g.cat) {
            case GateCat::NOOP:
                if (name == "X" || name == "Y" || name == "Z") {
                    int code = name == "X" ? 1 : name == "Z" ? 2 : 3;
                    for (uint32_t t : op.targets) {
                        sim.pauli(code, t & T_VALUE_MASK);
                    }
                }
                break;
            case GateCat::CLIFF1: {
                const Table1 &tb = tables().inv1.at(name);
                for (uint32_t t : op.targets) {
                    sim.gate1(tb, t & T_VALUE_MASK);
                }
            } break;
            case GateCat::CLIFF2: {
                const Table2 &tb = tables().inv2.at(name);
                for (size_t i = 0; i < op.targets.size(); i += 2) {
                    uint32_t a = op.targets[i], b = op.targets[i + 1];
                    bool a_bit = (a & (T_REC | T_SWEEP)) != 0, b_bit = (b & (T_REC | T_SWEEP)) != 0;
                    if (!a_bit && !b_bit) {
                        sim.gate2(tb, a & T_VALUE_MASK, b & T_VALUE_MASK);
                        continue;
                    }
                    // classically controlled Pauli (bit-as-target errors are raised by the lowering)
                    uint32_t bit, q;
                    int code;
                    if (name == "CX" || name == "CY") {
                        if (b_bit) throw std::invalid_argument("Controlled gate had a bit as its target, instead of its control.");
                        bit = a, q = b, code = name == "CX" ? 1 : 3;
                    } else if (name == "XCZ" || name == "YCZ") {
                        if (a_bit) throw std::invalid_argument("Controlled gate had a bit as its target, instead of its control.");
                        bit = b, q = a, code = name == "XCZ" ? 1 : 3;
                    } else {
                        if (a_bit && b_bit) continue;
                        bit = a_bit ? a : b, q = a_bit ? b : a, code = 2;
                    }
                    if ((bit & T_SWEEP) == 0 && rec_value(bit, g.name)) {
                        sim.pauli(code, q & T_VALUE_MASK);
                    }
                }
            } break;
            case GateCat::MEASURE: {
                uint32_t basis = g.param & 3, kind = g.param >> 2;
                sim.t_allowed = basis != GB_Y;
                sim.t_randoms = 0;
                sim.t_remaining = op.targets.size();
                for (uint32_t t : op.targets) {
                    sim.t_remaining--;
                    size_t q = t & T_VALUE_MASK;
                    if (kind == GK_R) {
                        sim.reset(basis, q);
                        continue;
                    }
                    bool m = sim.measure(basis, q);
                    sim.record.push_back((uint8_t)(m ^ ((t & T_INVERTED) != 0)));
                    if (kind == GK_MR && m) {
                        // back to the +1 eigenstate of the measured basis
                        sim.to_z_basis(basis, q);
                        sim.pauli(1, q);
                        sim.to_z_basis(basis, q);
                    }
                }
                sim.exit_t();
                sim.t_allowed = false;
            } break;
            case GateCat::MPAD:
                for (uint32_t t : op.targets) {
                    sim.record.push_back((uint8_t)(t & 1));
                }
                break;
            case GateCat::MPP:
                for (const Product &p : read_products(op)) {
                    if (p.terms.empty()) {
                        sim.record.push_back((uint8_t)p.sign);
                        continue;
                    }
                    size_t first = p.terms[0].first;
                    for (auto &e : p.terms) {
                        if (e.second == 1) sim.gate1("H", e.first);
                        if (e.second == 3) sim.gate1("H_YZ", e.first);
                    }
                    for (size_t i = 1; i < p.terms.size(); i++) {
                        sim.gate2("CX", p.terms[i].first, first);
                    }
                    bool m = sim.measure_z(first);
                    sim.record.push_back((uint8_t)(m ^ p.sign));
                    for (size_t i = 1; i < p.terms.size(); i++) {
                        sim.gate2("CX", p.terms[i].first, first);
                    }
                    for (auto &e : p.terms) {
                        if (e.second == 1) sim.gate1("H", e.first);
                        if (e.second == 3) sim.gate1("H_YZ", e.first);
                    }
                }
                break;
            case GateCat::SPP:
                for (const Product &p : read_products(op)) {
                    if (p.terms.empty()) {
                        continue;
                    }
                    size_t focus = p.terms[0].first;
                    bool dag = (name == "SPP_DAG") ^ p.sign;
                    auto conj = [&]() {
                        for (auto &e : p.terms) {
                            if (e.second == 1) sim.gate1("H", e.first);
                            if (e.second == 3) sim.gate1("H_YZ", e.first);
                        }
                    };
                    auto fold = [&]() {
                        for (size_t i = 1; i < p.terms.size(); i++) {
                            sim.gate2("CX", p.terms[i].first, focus);
                        }
                        for (uint32_t b : p.bits) {
                            if (!(b & T_SWEEP) && rec_value(b, g.name)) {
                                sim.pauli(1, focus);
                            }
                        }
                    };
                    conj();
                    fold();
                    sim.gate1(dag ? "S_DAG" : "S", focus);
                    fold();
                    conj();
                }
                break;
            case GateCat::MPAIR: {
                const char *conj = g.param == GB_X ? "CX" : g.param == GB_Y ? "CY" : "XCZ";
                for (size_t i = 0; i < op.targets.size(); i += 2) {
                    size_t a = op.targets[i] & T_VALUE_MASK, b = op.targets[i + 1] & T_VALUE_MASK;
                    bool inv = ((op.targets[i] ^ op.targets[i + 1]) & T_INVERTED) != 0;
                    sim.gate2(conj, a, b);
                    bool m = sim.measure(g.param, a);
                    sim.record.push_back((uint8_t)(m ^ inv));
                    sim.gate2(conj, a, b);
                }
            } break;
            case GateCat::HERALDED_ERASE:
            case GateCat::HERALDED_PAULI_CHANNEL_1:
                for (size_t i = 0; i < op.targets.size(); i++) {
                    sim.record.push_back(0);
                }
                break;
            default:  // noise, annotations
                break;
        }
    });
    return sim.record;
}

}  // namespace gstim
