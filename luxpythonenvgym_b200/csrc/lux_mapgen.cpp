// lux_mapgen.cpp -- native, multithreaded host map generator.
//
// Same pipeline and the same IEEE-double evaluation order as luxpythonenvgym_b200/mapgen.py, i.e. as the
// reference's GameMap.generate_map (luxai2021/game/game_map.py:60-438) driven by seedrandom's ARC4 stream
// "gen_<seed>" (luxai2021/env/rng/rng.js:4, seedrandom.js:1).  Map generation stays on the host
// (BASELINE.json north_star); this is what makes 10^5..10^6 distinct seeded matches practical
// (the Python statement needs ~20 ms per map).  Output: the packed initial state of include/lux_b200.h.
// tests/test_mapgen_native.py pins it against the reference's 100 golden maps and against mapgen.py.
#include <emmintrin.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lux_b200.h"

namespace {

struct Pow2Neg { double v[160]; Pow2Neg() { for (int k = 0; k < 160; k++) v[k] = ldexp(1.0, -k); } };
const Pow2Neg POW2_NEG;

struct Arc4 {
    uint8_t s[256];
    int i = 0, j = 0;
    // the first `len` bytes of the keystream (behind seedrandom's 256 discarded ones) may have been produced ahead
    // of time (Arc4x4 below); take() serves them from `buf` and carries on from the state they left behind
    const uint8_t* buf = nullptr;
    int pos = 0, len = 0;
    Arc4() {}
    explicit Arc4(const std::string& key) {
        schedule(key);
        take(256);
    }
    void schedule(const std::string& key) {
        for (int k = 0; k < 256; k++) s[k] = (uint8_t)k;
        int jj = 0;
        const int klen = key.empty() ? 1 : (int)key.size();
        for (int k = 0; k < 256; k++) {
            int kc = key.empty() ? 0 : (uint8_t)key[k % klen];
            jj = (jj + s[k] + kc) & 255;
            uint8_t t = s[k]; s[k] = s[jj]; s[jj] = t;
        }
    }
    uint64_t take(int n) {
        uint64_t acc = 0;
        if (pos + n <= len) {
            const uint8_t* q = buf + pos;
            pos += n;
            for (int k = 0; k < n; k++) acc = (acc << 8) | q[k];
            return acc;
        }
        while (pos < len && n > 0) { acc = (acc << 8) | buf[pos++]; n--; }
        // i, j live in registers for the run (the byte stores into s[] may alias the members otherwise)
        unsigned a = (unsigned)i, b = (unsigned)j;
        for (int k = 0; k < n; k++) {
            a = (a + 1) & 255;
            const unsigned sa = s[a];
            b = (b + sa) & 255;
            const unsigned sb = s[b];
            s[a] = (uint8_t)sb; s[b] = (uint8_t)sa;
            acc = (acc << 8) | s[(sa + sb) & 255];
        }
        i = (int)a; j = (int)b;
        return acc;
    }
    // seedrandom's random(): 48 bits, topped up byte-wise to >= 52 significant bits, halved below 2^53, divided by
    // the matching power of 256.  Every intermediate of the original's double arithmetic is an exactly
    // representable integer (num < 2^53 before each "* 256", and a multiple of the bits shifted out while it is
    // halved), so the same value comes out of integer shifts and one exact scaling by a power of two.
    double next() {
        uint64_t num = take(6);
        int den_bits = 48;
        uint32_t extra = 0;
        while (num < (1ull << 52)) { num = (num + extra) << 8; den_bits += 8; extra = (uint32_t)take(1); }
        while (num >= (1ull << 53)) { num >>= 1; den_bits -= 1; extra >>= 1; }
        return (double)(num + extra) * POW2_NEG.v[den_bits];            // exact: a power of two
    }
};

// Four keystreams side by side.  One RC4 byte is a chain of dependent loads and stores through the 256-byte state
// (3.5 ns here); four independent states in one loop fill the pipeline instead (1.2 ns per byte and stream).  A map
// needs 13.2 KB of keystream on average (32 x 32; 95 % below 13.7 KB): PREFIX bytes per map are produced this way,
// the rare map that needs more (a layout that failed the resource minimum and is drawn again) continues alone.
struct Arc4x4 {
    enum { PREFIX = 14 * 1024 };
    Arc4 r[4];
    std::vector<uint8_t> stream;      // 4 x (256 + PREFIX)
    Arc4x4() : stream(4 * (256 + PREFIX)) {}
    void start(const std::string key[4], int prefix) {
        if (prefix > PREFIX) prefix = PREFIX;
        for (int q = 0; q < 4; q++) r[q].schedule(key[q]);
        // the four states side by side in one local block (the compiler can then tell the streams apart)
        uint8_t st[4][256];
        for (int q = 0; q < 4; q++) memcpy(st[q], r[q].s, 256);
        uint8_t* out[4];
        for (int q = 0; q < 4; q++) out[q] = &stream[(size_t)q * (256 + PREFIX)];
        unsigned a = 0, bs[4] = {0, 0, 0, 0};
        const int total = 256 + prefix;
        for (int k = 0; k < total; k++) {
            a = (a + 1) & 255;
#pragma GCC unroll 4
            for (int q = 0; q < 4; q++) {
                const unsigned x = st[q][a];
                bs[q] = (bs[q] + x) & 255;
                const unsigned y = st[q][bs[q]];
                st[q][a] = (uint8_t)y; st[q][bs[q]] = (uint8_t)x;
                out[q][k] = st[q][(x + y) & 255];
            }
        }
        for (int q = 0; q < 4; q++) memcpy(r[q].s, st[q], 256);
        for (int q = 0; q < 4; q++) {
            r[q].i = (int)a; r[q].j = (int)bs[q];
            r[q].buf = &stream[(size_t)q * (256 + PREFIX) + 256];      // seedrandom discards the first 256 bytes
            r[q].pos = 0; r[q].len = prefix;
        }
    }
};

struct Res { int kind = 0; int amt = 0; double fx = 0, fy = 0; };   // kind 0 = empty

const int DELTA[8][2] = {{0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}, {1, 0}, {1, 1}};
inline int sgn(double v) { return v > 0 ? 1 : (v < 0 ? -1 : 0); }

struct Field {
    int w, h;
    std::vector<int> at;          // index into pool, -1 empty
    std::vector<int> scratch;     // the drift step's output grid (swapped with `at`)
    std::vector<uint8_t> bits;    // automaton layer / per-cell resource kind of the drift step
    std::vector<Res> pool;        // deposits are moved by reference in the original; indices play that role
    Field(int w_, int h_) : w(w_), h(h_), at((size_t)w_ * h_, -1) { pool.reserve(256); }
    int& cell(int x, int y) { return at[(size_t)y * w + x]; }
};

int wood_amt(Arc4& r) { int a = 300 + (int)floor(r.next() * 100); return a < 500 ? a : 500; }
int coal_amt(Arc4& r) { return 350 + (int)floor(r.next() * 75); }
int uranium_amt(Arc4& r) { return 300 + (int)floor(r.next() * 50); }
int amount_of(int kind, Arc4& r) { return kind == LUX_RES_WOOD ? wood_amt(r) : kind == LUX_RES_COAL ? coal_amt(r) : uranium_amt(r); }

void automaton_layer(Arc4& r, double density, double spread, int w, int h, int death, int birth, std::vector<uint8_t>& g) {
    double p = density - spread / 2 + spread * r.next();
    g.assign((size_t)w * h, 0);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) g[(size_t)y * w + x] = r.next() < p ? 1 : 0;
    // two sweeps of the life-like rule, IN PLACE (a cell sees the already updated cells before it, as the
    // original's single grid does)
    for (int round = 0; round < 2; round++)
        for (int i = 1; i < h - 1; i++) {
            uint8_t* up = &g[(size_t)(i - 1) * w];
            uint8_t* me = up + w;
            uint8_t* dn = me + w;
            for (int j = 1; j < w - 1; j++) {
                const int alive = up[j - 1] + up[j] + up[j + 1] + me[j - 1] + me[j + 1] + dn[j - 1] + dn[j] + dn[j + 1];
                me[j] = me[j] ? (alive < death ? 0 : 1) : (alive > birth ? 1 : 0);
            }
        }
}

// Force contributions of a neighbour at offset (dx, dy) = (x - xx, y - yy), dx, dy in [-4, 5]:
// pow(dx / md, 2) * sgn(dx) and the same for dy, evaluated once with the very expressions of the original loop
// (so that the sums below add the same doubles in the same order).  Flat: entry (dy + 4) * 10 + (dx + 4).
// f[sign][index] = (x component, y component); sign 1 = the negated pair (x * -1.0 == -x exactly, so the table
// entry is what the original's multiplication by -1 produces, zeros included).
struct DriftTable {
    __m128d f[2][100];
    DriftTable() {
        for (int dy = -4; dy <= 5; dy++)
            for (int dx = -4; dx <= 5; dx++) {
                const int md = abs(dx) + abs(dy);
                const double fx = dx ? pow((double)dx / md, 2) * sgn(dx) : 0.0;
                const double fy = dy ? pow((double)dy / md, 2) * sgn(dy) : 0.0;
                f[0][(dy + 4) * 10 + dx + 4] = _mm_set_pd(fy, fx);
                f[1][(dy + 4) * 10 + dx + 4] = _mm_set_pd(fy * -1.0, fx * -1.0);
            }
    }
};
const DriftTable DRIFT;

// One gravity step (game_map.py:330-380): every deposit is pushed away from deposits of other kinds and pulled
// towards its own kind inside a 10 x 10 window, then moves one cell along the sign of the force.  Occupancy is kept
// as one 32-bit mask per row and kind; the window of a deposit -- at most 10 rows of 10 cells -- is cut out of them
// into a 100-bit word (bit r * 10 + c = row y0 + r, column x0 + c), whose set bits are visited in ascending
// order, i.e. row by row with ascending x: the order of the original double loop, which fixes the floating-point
// sums.  One loop per deposit instead of one per window row: a single hard-to-predict loop exit, and the table
// index of bit k is simply `base - k`.
void drift(Field& f) {
    const int w = f.w, h = f.h;
    uint32_t rows[32], rows_kind[4][32];
    for (int y = 0; y < h; y++) {
        uint32_t m = 0, mk[4] = {0, 0, 0, 0};
        for (int x = 0; x < w; x++) {
            const int id = f.cell(x, y);
            if (id >= 0) { m |= 1u << x; mk[f.pool[id].kind & 3] |= 1u << x; }
        }
        rows[y] = m;
        for (int k = 0; k < 4; k++) rows_kind[k][y] = mk[k];
    }
    // pass 1: the window words of every deposit (integer work only)
    struct Win { uint64_t all[2], same[2]; int base; Res* res; };
    static thread_local std::vector<Win> wins;
    wins.clear();
    for (int y = 0; y < h; y++)
        for (uint32_t own = rows[y]; own; own &= own - 1) {
            const int x = __builtin_ctz(own);
            Win wn;
            wn.res = &f.pool[f.cell(x, y)];
            const uint32_t* same_rows = rows_kind[wn.res->kind & 3];
            const int x0 = x - 5 < 0 ? 0 : x - 5;                                       // columns [x0, x0 + 10) cut to the map
            const int y0 = y - 5 < 0 ? 0 : y - 5, y1 = y + 5 > h ? h : y + 5;
            const uint32_t cut = (x + 5 >= 32 ? 0xffffffffu : ((1u << (x + 5)) - 1u)) >> x0;     // window columns, shifted down; <= 10 bits
            // bits 0..59: window rows 0..5, bits 60..99: rows 6..9 (second word)
            wn.all[0] = wn.all[1] = wn.same[0] = wn.same[1] = 0;
            for (int yy = y0; yy < y1; yy++) {
                const int rr = yy - y0, word = rr >= 6, sh = (word ? rr - 6 : rr) * 10;
                wn.all[word] |= (uint64_t)((rows[yy] >> x0) & cut) << sh;
                wn.same[word] |= (uint64_t)((same_rows[yy] >> x0) & cut) << sh;
            }
            // dx = x - x0 - c, dy = y - y0 - r  ->  table index (dy + 4) * 10 + dx + 4 = base - (r * 10 + c)
            wn.base = (y - y0 + 4) * 10 + (x - x0 + 4);
            wins.push_back(wn);
        }
    // pass 2: the sums, two deposits side by side (each sum is a chain of dependent additions in a fixed order; two
    // independent chains keep the adder busy).  Other kinds push (+), the same kind pulls (-): the pair (x, y) of a
    // contribution comes out of the table with its sign and is added with one two-lane addition (each lane is the
    // original's scalar sum; a zero component adds +-0.0, which leaves the sum's value -- and its sign test -- as
    // the original's skipped addition does).
#define LUX_DRIFT_STEP(bits, sm, tab, acc) do { const int k_ = __builtin_ctzll(bits); bits &= bits - 1; \
        acc = _mm_add_pd(acc, tab[(int)((sm >> k_) & 1) * 100 - k_]); } while (0)
    const size_t nw = wins.size();
    for (size_t i = 0; i < nw; i += 2) {
        const Win& A = wins[i];
        const Win& B = wins[i + 1 < nw ? i + 1 : i];
        __m128d af = _mm_setzero_pd(), bf = _mm_setzero_pd();
        for (int word = 0; word < 2; word++) {
            const __m128d* at = DRIFT.f[0] + A.base - 60 * word;
            const __m128d* bt = DRIFT.f[0] + B.base - 60 * word;
            uint64_t ab = A.all[word], bb = B.all[word];
            const uint64_t as = A.same[word], bs = B.same[word];
            while (ab && bb) { LUX_DRIFT_STEP(ab, as, at, af); LUX_DRIFT_STEP(bb, bs, bt, bf); }
            while (ab) LUX_DRIFT_STEP(ab, as, at, af);
            while (bb) LUX_DRIFT_STEP(bb, bs, bt, bf);
        }
        A.res->fx = _mm_cvtsd_f64(af); A.res->fy = _mm_cvtsd_f64(_mm_unpackhi_pd(af, af));
        if (i + 1 < nw) { B.res->fx = _mm_cvtsd_f64(bf); B.res->fy = _mm_cvtsd_f64(_mm_unpackhi_pd(bf, bf)); }
    }
#undef LUX_DRIFT_STEP
    std::vector<int>& out = f.scratch;
    out.assign((size_t)w * h, -1);
    for (int y = 0; y < h; y++)
        for (uint32_t own = rows[y]; own; own &= own - 1) {
            const int x = __builtin_ctz(own);
            const int id = f.cell(x, y);
            const Res& r = f.pool[id];
            int nx = x + sgn(r.fx), ny = y + sgn(r.fy);
            nx = nx < 0 ? 0 : (nx > w - 1 ? w - 1 : nx);
            ny = ny < 0 ? 0 : (ny > h - 1 ? h - 1 : ny);
            if (out[(size_t)ny * w + nx] < 0) out[(size_t)ny * w + nx] = id;
            else out[(size_t)y * w + x] = id;
        }
    f.at.swap(out);
}

Field layout(Arc4& r, int symmetry_vertical, int w, int h, int hw, int hh) {
    Field f(w, h);
    const struct { int kind; double density, spread; int death, birth; } layers[3] = {
        {LUX_RES_WOOD, 0.21, 0.01, 2, 4}, {LUX_RES_COAL, 0.11, 0.02, 2, 4}, {LUX_RES_URANIUM, 0.055, 0.04, 1, 6}};
    std::vector<uint8_t>& g = f.bits;
    for (auto& L : layers) {
        automaton_layer(r, L.density, L.spread, hw, hh, L.death, L.birth, g);
        for (int y = 0; y < hh; y++)
            for (int x = 0; x < hw; x++)
                if (g[(size_t)y * hw + x] == 1) {
                    Res d; d.kind = L.kind; d.amt = amount_of(L.kind, r);
                    f.pool.push_back(d);
                    f.cell(x, y) = (int)f.pool.size() - 1;
                }
    }
    for (int k = 0; k < 10; k++) drift(f);
    for (int y = 0; y < hh; y++)
        for (int x = 0; x < hw; x++) {
            int id = f.cell(x, y);
            if (id < 0) continue;
            const int kind = f.pool[id].kind;
            for (auto& d : DELTA) {
                int nx = x + d[0], ny = y + d[1];
                // the original compares nx with the half HEIGHT and ny with the half WIDTH
                if (nx < 0 || ny < 0 || nx >= hh || ny >= hw) continue;
                if (r.next() < 0.05) {
                    int amt = uranium_amt(r);
                    if (kind == LUX_RES_COAL) amt = coal_amt(r);
                    if (kind == LUX_RES_WOOD) amt = wood_amt(r);
                    Res nd; nd.kind = kind; nd.amt = amt;
                    f.pool.push_back(nd);
                    f.cell(nx, ny) = (int)f.pool.size() - 1;
                }
            }
        }
    for (int y = 0; y < hh; y++)
        for (int x = 0; x < hw; x++) {
            if (symmetry_vertical) f.cell(w - x - 1, y) = f.cell(x, y);
            else f.cell(x, h - y - 1) = f.cell(x, y);
        }
    return f;
}

bool enough(Field& f) {
    long tot[4] = {0, 0, 0, 0};
    for (int id : f.at)
        if (id >= 0) tot[f.pool[id].kind] += f.pool[id].amt;
    return tot[LUX_RES_WOOD] >= 2000 && tot[LUX_RES_COAL] >= 1500 && tot[LUX_RES_URANIUM] >= 300;
}

int generate_one(Arc4& r, int size, int u_cap, int c_cap, int t_cap, int r_cap, LuxHeader* hdr, LuxCell* cells,
                 LuxUnit* units, LuxCity* cities, uint16_t* tiles, LuxRes* res) {
    static const int SIZES[4] = {12, 16, 24, 32};
    int drawn = SIZES[(int)floor(r.next() * 4)];
    int w = size > 0 ? size : drawn, h = w;
    if (w < 4 || w > 32 || (w & 1) || u_cap < 2 || c_cap < 2 || t_cap < 2) return LUX_E_ARG;
    int hw = w, hh = h, vertical = 0;
    if (r.next() < 0.5) { vertical = 1; hw = w / 2; } else { hh = h / 2; }
    Field f = layout(r, vertical, w, h, hw, hh);
    while (!enough(f)) f = layout(r, vertical, w, h, hw, hh);

    memset(hdr, 0, sizeof *hdr);
    memset(cells, 0, sizeof(LuxCell) * (size_t)w * h);
    int nres = 0;
    auto add_resource = [&](int x, int y, int kind, int amt) -> bool {
        LuxCell& c = cells[y * w + x];
        c.amount = (uint16_t)amt;
        c.flags = (uint8_t)((c.flags & 0xFC) | kind);
        if (nres >= r_cap) return false;
        res[nres++] = (LuxRes)(y * w + x);
        return true;
    };
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int id = f.cell(x, y);
            if (id >= 0 && !add_resource(x, y, f.pool[id].kind, f.pool[id].amt)) return LUX_E_CAPACITY;
        }
    int sx = (int)floor(r.next() * (hw - 1)) + 1, sy = (int)floor(r.next() * (hh - 1)) + 1;
    while (cells[sy * w + sx].amount > 0) {
        sx = (int)floor(r.next() * (hw - 1)) + 1;
        sy = (int)floor(r.next() * (hh - 1)) + 1;
    }
    const int spawn[2][2] = {{sx, sy}, {vertical ? w - sx - 1 : sx, vertical ? sy : h - sy - 1}};
    for (int team = 0; team < 2; team++) {
        int x = spawn[team][0], y = spawn[team][1];
        LuxUnit& u = units[team];
        memset(&u, 0, sizeof u);
        u.id = team + 1; u.x = (uint8_t)x; u.y = (uint8_t)y; u.team_type = (uint8_t)team;
        LuxCity& c = cities[team];
        memset(&c, 0, sizeof c);
        c.id = team + 1; c.team = (uint8_t)team; c.ntiles = 1;
        LuxCell& cell = cells[y * w + x];
        cell.flags = (uint8_t)(cell.flags | ((team + 1) << 2));
        cell.city_slot = (uint8_t)team;
        tiles[team] = (uint16_t)(y * w + x);
    }
    int first = (int)floor(r.next() * 8), placed = 0;
    for (int k = 0; k < 7; k++) {
        const int* d = DELTA[(first + k) % 8];
        int nx = sx + d[0], ny = sy + d[1];
        int nx2 = vertical ? w - nx - 1 : nx, ny2 = vertical ? ny : h - ny - 1;
        if (nx < 0 || ny < 0 || nx >= w || ny >= h || nx2 < 0 || ny2 < 0 || nx2 >= w || ny2 >= h) continue;
        const int pts[2][2] = {{nx, ny}, {nx2, ny2}};
        for (auto& p : pts) {
            LuxCell& c = cells[p[1] * w + p[0]];
            if (c.amount == 0 && (c.flags & 0x0C) == 0) {
                placed++;
                if (!add_resource(p[0], p[1], LUX_RES_WOOD, 800)) return LUX_E_CAPACITY;
            }
        }
        if (placed == 6) break;
    }
    hdr->unit_id_count = 2; hdr->city_id_count = 2;
    // the list entries carry the cell's type and a mirror of its final amount (include/lux_b200.h LuxRes)
    for (int k = 0; k < nres; k++) {
        const LuxCell& c = cells[res[k]];
        res[k] = LUX_RES_MAKE(res[k], (uint32_t)(c.flags & 3), c.amount);
    }
    hdr->n_units = 2; hdr->n_cities = 2; hdr->n_tiles = 2; hdr->n_res = (uint16_t)nres;
    hdr->n_team_units[0] = hdr->n_team_units[1] = 1;
    hdr->n_team_tiles[0] = hdr->n_team_tiles[1] = 1;
    hdr->workers_built[0] = hdr->workers_built[1] = 1;
    hdr->winner = LUX_WIN_NONE;
    hdr->size = (uint8_t)w;
    return LUX_OK;
}

}  // namespace

extern "C" int lux_mapgen_natural_size(int64_t seed) {
    Arc4 r("gen_" + std::to_string((long long)seed));
    static const int SIZES[4] = {12, 16, 24, 32};
    return SIZES[(int)floor(r.next() * 4)];
}

// Generates maps for seeds[0..n) into HOST arrays with the section strides of LuxBatch (hdr/cells/...
// are host pointers here).  size > 0 forces width = height = size (configs["width"/"height"]);
// all maps of one batch must have that side.  threads <= 0: hardware concurrency.
extern "C" int lux_mapgen_generate_batch(const LuxBatch* out, const int64_t* seeds, int threads) {
    if (!out || !seeds || out->n_matches < 0 || !out->hdr || !out->cells || !out->units || !out->cities || !out->tiles ||
        !out->res)
        return LUX_E_ARG;
    const int n = out->n_matches, size = out->size;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    std::atomic<int> next(0), status(LUX_OK);
    // keystream produced ahead per map: what a map of this side needs (measured mean + 4 %: 32 x 32 13.7 KB)
    const int prefix = size >= 32 ? 14 * 1024 : size >= 24 ? 8 * 1024 : size >= 16 ? 4 * 1024 : 2560;
    auto work = [&]() {
        Arc4x4 quad;
        for (;;) {
            // four maps at a time: their keystreams are produced side by side, then each map is laid out from its own
            const int m0 = next.fetch_add(4);
            if (m0 >= n) break;
            const int cnt = n - m0 < 4 ? n - m0 : 4;
            std::string key[4];
            for (int q = 0; q < 4; q++) key[q] = "gen_" + std::to_string((long long)seeds[m0 + (q < cnt ? q : 0)]);
            quad.start(key, prefix);
            for (int q = 0; q < cnt; q++) {
                const int m = m0 + q;
                int rc = generate_one(quad.r[q], size, out->u_cap, out->c_cap, out->t_cap, out->r_cap, out->hdr + m,
                                      out->cells + (size_t)m * size * size, out->units + (size_t)m * out->u_cap,
                                      out->cities + (size_t)m * out->c_cap, out->tiles + (size_t)m * out->t_cap,
                                      out->res + (size_t)m * out->r_cap);
                if (rc != LUX_OK) status.store(rc);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return status.load();
}
