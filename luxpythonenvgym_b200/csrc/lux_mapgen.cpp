// lux_mapgen.cpp -- native, multithreaded host map generator.
//
// Same pipeline and the same IEEE-double evaluation order as luxpythonenvgym_b200/mapgen.py, i.e. as the
// reference's GameMap.generate_map (luxai2021/game/game_map.py:60-438) driven by seedrandom's ARC4 stream
// "gen_<seed>" (luxai2021/env/rng/rng.js:4, seedrandom.js:1).  Map generation stays on the host
// (BASELINE.json north_star); this is what makes 10^5..10^6 distinct seeded matches practical
// (the Python statement needs ~20 ms per map).  Output: the packed initial state of include/lux_b200.h.
// tests/test_mapgen_native.py pins it against the reference's 100 golden maps and against mapgen.py.
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
    explicit Arc4(const std::string& key) {
        for (int k = 0; k < 256; k++) s[k] = (uint8_t)k;
        int jj = 0;
        const int len = key.empty() ? 1 : (int)key.size();
        for (int k = 0; k < 256; k++) {
            int kc = key.empty() ? 0 : (uint8_t)key[k % len];
            jj = (jj + s[k] + kc) & 255;
            uint8_t t = s[k]; s[k] = s[jj]; s[jj] = t;
        }
        take(256);
    }
    uint64_t take(int n) {
        // i, j live in registers for the run (the byte stores into s[] may alias the members otherwise)
        uint64_t acc = 0;
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
// (so that the sums below add the same doubles in the same order).
struct DriftTable {
    double fx[10][10], fy[10][10];
    DriftTable() {
        for (int dy = -4; dy <= 5; dy++)
            for (int dx = -4; dx <= 5; dx++) {
                const int md = abs(dx) + abs(dy);
                fx[dy + 4][dx + 4] = dx ? pow((double)dx / md, 2) * sgn(dx) : 0.0;
                fy[dy + 4][dx + 4] = dy ? pow((double)dy / md, 2) * sgn(dy) : 0.0;
            }
    }
};
const DriftTable DRIFT;
const double SIGN[2] = {1.0, -1.0};

// One gravity step (game_map.py:330-380): every deposit is pushed away from deposits of other kinds and pulled
// towards its own kind inside a 10 x 10 window, then moves one cell along the sign of the force.  Occupied cells
// are found through one 32-bit occupancy mask per row; the neighbours of a window are visited row by row with
// ascending x, i.e. in the order of the original double loop, which fixes the floating-point sums.
void drift(Field& f) {
    const int w = f.w, h = f.h;
    uint32_t rows[32];
    std::vector<uint8_t>& kinds = f.bits;
    kinds.assign((size_t)w * h, 0);
    for (int y = 0; y < h; y++) {
        uint32_t m = 0;
        for (int x = 0; x < w; x++) {
            const int id = f.cell(x, y);
            if (id >= 0) { m |= 1u << x; kinds[(size_t)y * w + x] = (uint8_t)f.pool[id].kind; }
        }
        rows[y] = m;
    }
    for (int y = 0; y < h; y++)
        for (uint32_t own = rows[y]; own; own &= own - 1) {
            const int x = __builtin_ctz(own);
            Res& r = f.pool[f.cell(x, y)];
            const int kind = r.kind;
            double fx = 0, fy = 0;
            const int x0 = x - 5 < 0 ? 0 : x - 5, x1 = x + 5 > w ? w : x + 5;          // [x0, x1)
            const uint32_t window = (x1 >= 32 ? 0xffffffffu : ((1u << x1) - 1u)) & ~((1u << x0) - 1u);
            const int y0 = y - 5 < 0 ? 0 : y - 5, y1 = y + 5 > h ? h : y + 5;
            for (int yy = y0; yy < y1; yy++) {
                const uint8_t* krow = &kinds[(size_t)yy * w];
                const double* tx = DRIFT.fx[y - yy + 4] + (x + 4);                   // tx[-xx] = fx[dy + 4][x - xx + 4]
                const double* ty = DRIFT.fy[y - yy + 4] + (x + 4);
                for (uint32_t bits = rows[yy] & window; bits; bits &= bits - 1) {
                    const int xx = __builtin_ctz(bits);
                    // other kinds push (+), the same kind pulls (-); a zero component adds +-0.0, which leaves
                    // the sum's value (and its sign test) as the original's skipped addition does
                    // (the sign comes from a two-entry table: x * -1.0 == -x exactly, and no branch to mispredict)
                    const double sg = SIGN[krow[xx] == kind];
                    fx += tx[-xx] * sg;
                    fy += ty[-xx] * sg;
                }
            }
            r.fx = fx; r.fy = fy;
        }
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

int generate_one(int64_t seed, int size, int u_cap, int c_cap, int t_cap, int r_cap, LuxHeader* hdr, LuxCell* cells,
                 LuxUnit* units, LuxCity* cities, uint16_t* tiles, LuxRes* res) {
    Arc4 r("gen_" + std::to_string((long long)seed));
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
    auto work = [&]() {
        for (;;) {
            int m = next.fetch_add(1);
            if (m >= n) break;
            int rc = generate_one(seeds[m], size, out->u_cap, out->c_cap, out->t_cap, out->r_cap, out->hdr + m,
                                  out->cells + (size_t)m * size * size, out->units + (size_t)m * out->u_cap,
                                  out->cities + (size_t)m * out->c_cap, out->tiles + (size_t)m * out->t_cap,
                                  out->res + (size_t)m * out->r_cap);
            if (rc != LUX_OK) status.store(rc);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return status.load();
}
