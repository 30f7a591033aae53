// Layout generator core: one *attempt* of the reference's WorldGenerator per thread.
//
// Follows python/lle/generator/generator.py:188-228 (`_make_candidate_layout`) as called by `_try_generate(seed)`
// (:243-254) with constraint=None: agents -> exits -> lasers -> walls -> gems -> geometry check
// (candidates.py:27-41), every placement rule of placements.py / geometry.py, and - because an attempt is a pure function
// of (configuration, seed) only through Python's `random.Random` - CPython's generator itself: MT19937 seeded by
// `init_by_array` (Modules/_randommodule.c: random_seed for an int seed), `getrandbits`, `_randbelow_with_getrandbits`,
// `sample` (pool / selection-set variants and its `setsize` rule), `shuffle`, `choice`, `randint`, `choices`
// (Lib/random.py, CPython 3.12).  The layout for seed s is therefore bit-identical to the reference's.
//
// Formulation (not the reference's): no Python sets or lists of tuples.  Occupancy lives in per-row and per-column bit
// masks (H, W <= 32), so "no reserved cell on the full beam" is one mask test; populations of `sample` are never
// materialised unless CPython's pool variant needs the swap semantics; laser candidates are 16-bit (cell, direction)
// codes.  The function is __host__ __device__ so that the CPU test-suite can instantiate the same code without a GPU
// (tests/host_shim); the product only ever launches it as a kernel.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LLE_HD __host__ __device__ __forceinline__
#define LLE_HD_NOINLINE __host__ __device__ __noinline__
#else
#define LLE_HD inline
#define LLE_HD_NOINLINE
#endif

namespace llegen {

constexpr int kMaxDim = 32;      // H, W <= 32 (one u32 mask per row / column)
constexpr int kMaxCells = 1024;
constexpr int kMaxAgents = 32;
constexpr int kWork = 4 * kMaxCells;  // laser candidates: (cell, direction)

enum : int32_t { STARTS_RANDOM = 0, STARTS_EDGE = 1, STARTS_CLUSTERED = 2 };
enum : int32_t { EXITS_RANDOM = 0, EXITS_EDGE = 1, EXITS_CLUSTER = 2, EXITS_OPPOSITE = 3 };
enum : int32_t { LASERS_FREE = 0, LASERS_CROSS_AGENT = 1, LASERS_CROSS_CLUSTER = 2 };
enum : int32_t { SPAN_ANY = 0, SPAN_ACROSS = -1 };
enum : int32_t { EDGE_LEFT = 0, EDGE_RIGHT = 1, EDGE_TOP = 2, EDGE_BOTTOM = 3 };  // placements.py:87 order
enum : int32_t { DIR_N = 0, DIR_S = 1, DIR_E = 2, DIR_W = 3 };                   // placements.py:30 ALL_DIRS order

// cell codes of the output grid (include/lle_b200.h)
enum : uint8_t { CELL_FLOOR = 0, CELL_WALL = 1, CELL_EXIT = 2, CELL_GEM = 3, CELL_START = 16, CELL_SOURCE = 64 };
enum : uint8_t { LABEL_WALKABLE = 1, LABEL_INDEPENDENT = 2, LABEL_NEEDS_BLOCKER = 4 };

struct Config {  // == lle_gen_options after validation (n_walls resolved, room walls precomputed)
    int32_t width, height, n_agents;
    int32_t starts, exits;
    int32_t n_lasers, n_gems;
    int32_t laser_placement, laser_span;
    int32_t n_walls, walls_shapes;
    int32_t rooms;  // 1: walls are the structural room dividers below (generator.py:192-203)
    int32_t cluster_h, cluster_w;
    uint32_t room_rows[kMaxDim];  // bit c of room_rows[r]: (r, c) is a room wall
};

// --------------------------------------------------------------------------------------------------------------------
// CPython's random.Random
// --------------------------------------------------------------------------------------------------------------------
// `Store` holds the 624 state words (a plain array: host stack, or local memory on the device).
struct ArrayStore {
    uint32_t w[624];
    LLE_HD uint32_t& operator[](int i) { return w[i]; }
};

template <class Store>
struct PyRandomT {
    Store mt;
    int idx;  // next state word to regenerate and hand out

    // random_seed(int) -> init_by_array(key) (Modules/_randommodule.c).  `base(i)` = word i of the state after
    // init_genrand(19650218), a constant table.  init_by_array makes two passes over the state, each a dependent chain:
    // pass 1 walks i = 1..623 and then i = 1 once more, pass 2 walks i = 2..623 and then i = 1; with a key of one or two
    // 32-bit digits, `key[j] + j` alternates between k_odd (odd i) and k_even (even i).  Pass 2 cannot start before pass 1
    // has ended (it continues from pass 1's last value), and it reads every word pass 1 wrote - 2.5 KB per thread that the
    // resident threads cannot keep in cache.  So pass 1 runs twice instead: once in registers only, to reach its last
    // value, and once more in lock-step with pass 2, which then takes each pass-1 word from a register.  The state is
    // written exactly once and never read while seeding; the two chains of the second walk overlap.
    template <class Base>
    LLE_HD void seed(uint64_t s, Base base) {
        const uint32_t hi = (uint32_t)(s >> 32);
        const uint32_t k_odd = (uint32_t)s;             // key[0] + 0
        const uint32_t k_even = hi ? hi + 1u : k_odd;    // key[1] + 1 when abs(seed) has two digits, else key[0] + 0 again
        const uint32_t first = (base(1) ^ ((base(0) ^ (base(0) >> 30)) * 1664525u)) + k_odd;  // pass 1 at i = 1
        uint32_t p1 = first;
        for (int i = 2; i < 624; i += 2) {
            p1 = (base(i) ^ ((p1 ^ (p1 >> 30)) * 1664525u)) + k_even;
            p1 = (base(i + 1) ^ ((p1 ^ (p1 >> 30)) * 1664525u)) + k_odd;
        }
        const uint32_t wrapped = (first ^ ((p1 ^ (p1 >> 30)) * 1664525u)) + k_even;  // mt[0] = mt[623], then i = 1 again
        p1 = first;
        uint32_t p2 = wrapped;
        for (int i = 2; i < 624; i += 2) {
            p1 = (base(i) ^ ((p1 ^ (p1 >> 30)) * 1664525u)) + k_even;
            p2 = (p1 ^ ((p2 ^ (p2 >> 30)) * 1566083941u)) - (uint32_t)i;
            mt[i] = p2;
            p1 = (base(i + 1) ^ ((p1 ^ (p1 >> 30)) * 1664525u)) + k_odd;
            p2 = (p1 ^ ((p2 ^ (p2 >> 30)) * 1566083941u)) - (uint32_t)(i + 1);
            mt[i + 1] = p2;
        }
        mt[1] = (wrapped ^ ((p2 ^ (p2 >> 30)) * 1566083941u)) - 1u;  // pass 2 wrapped: mt[0] = mt[623], i = 1
        mt[0] = 0x80000000u;
        idx = 0;
    }
    LLE_HD static uint32_t mix(uint32_t u, uint32_t v, uint32_t far) {
        const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
        return far ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
    }
    // genrand_uint32 with the regeneration done one word at a time: the block update of the reference implementation
    // rewrites mt[0..623] in increasing order and in place, so producing word k right before it is handed out reads
    // exactly the same old / new neighbours.  An attempt that draws 80 numbers regenerates 80 words, not 624.
    LLE_HD uint32_t next() {
        const int k = idx;
        const int k1 = k == 623 ? 0 : k + 1;
        uint32_t y = mix(mt[k], mt[k1], mt[k < 227 ? k + 397 : k - 227]);
        mt[k] = y;
        idx = k1;
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    // Random._randbelow_with_getrandbits, n in [1, 2^31]
    LLE_HD uint32_t below(uint32_t n) {
        int k = 0;
        for (uint32_t t = n; t; t >>= 1) ++k;
        uint32_t r = next() >> (32 - k);
        while (r >= n) r = next() >> (32 - k);
        return r;
    }
    // floor(random() * 16) with random() = (a * 2^26 + b) / 2^53, a = next() >> 5, b = next() >> 6
    LLE_HD uint32_t random_times_16() {
        const uint32_t a = next() >> 5;
        (void)next();
        return a >> 23;
    }
};

// Random.sample's choice between its two variants (Lib/random.py: setsize = 21; if k > 5: += 4 ** ceil(log(3k, 4)))
LLE_HD bool sample_uses_pool(int n, int k) {
    int setsize = 21;
    if (k > 5) {
        int p = 1;
        while (p < 3 * k) p *= 4;  // 3k is never a power of 4, so the float log in CPython rounds the same way
        setsize += p;
    }
    return n <= setsize;
}

// --------------------------------------------------------------------------------------------------------------------
// occupancy masks
// --------------------------------------------------------------------------------------------------------------------
struct Grid {
    uint32_t row[kMaxDim];
    uint32_t col[kMaxDim];
    LLE_HD void clear() {
        for (int k = 0; k < kMaxDim; ++k) row[k] = col[k] = 0;
    }
    LLE_HD void set(int r, int c) { row[r] |= 1u << c; col[c] |= 1u << r; }
    LLE_HD bool get(int r, int c) const { return (row[r] >> c) & 1u; }
};

LLE_HD uint32_t low_mask(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }
LLE_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
LLE_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
LLE_HD int nth_set_bit(uint32_t x, int n) {
    for (; n > 0; --n) x &= x - 1;
    return ctz32(x);
}

using PyRandom = PyRandomT<ArrayStore>;

template <class Rng>
struct Attempt {
    const Config& c;
    Rng& rng;
    uint16_t* work;  // kWork entries
    int H, W;
    Grid reserved;
    uint32_t walls[kMaxDim];   // row masks
    uint32_t extra[kMaxDim];   // scratch row masks (selection sets, free sets, beams)
    int16_t agents[kMaxAgents], exits_[kMaxAgents];
    int16_t laser_cell[kMaxAgents];
    int8_t laser_dir[kMaxAgents], laser_colour[kMaxAgents];
    int16_t gems_n;
    int edge, agent_r, agent_c, exit_r, exit_c;
    int16_t lanes[kMaxAgents];

    LLE_HD Attempt(const Config& cfg, Rng& r, uint16_t* w) : c(cfg), rng(r), work(w), H(cfg.height), W(cfg.width) {}

    // ---- populations ------------------------------------------------------------------------------------------------
    // number of cells whose bit is clear in `blocked` (row masks)
    LLE_HD int count_free(const uint32_t* blocked) const {
        int n = 0;
        const uint32_t wm = low_mask(W);
        for (int r = 0; r < H; ++r) n += popc32(~blocked[r] & wm);
        return n;
    }
    // the j-th such cell in row-major order
    LLE_HD int nth_free(const uint32_t* blocked, int j) const {
        const uint32_t wm = low_mask(W);
        for (int r = 0; r < H; ++r) {
            const uint32_t f = ~blocked[r] & wm;
            const int n = popc32(f);
            if (j < n) return r * W + nth_set_bit(f, j);
            j -= n;
        }
        return -1;
    }

    // Random.sample(population, k) where population[j] = elem(j); out[] in selection order.
    template <class Elem>
    LLE_HD void sample(int n, int k, Elem elem, int16_t* out) {
        if (sample_uses_pool(n, k)) {
            for (int j = 0; j < n; ++j) work[j] = (uint16_t)elem(j);
            for (int i = 0; i < k; ++i) {
                const int j = (int)rng.below((uint32_t)(n - i));
                out[i] = (int16_t)work[j];
                work[j] = work[n - i - 1];
            }
        } else {
            uint32_t* sel = extra;  // n <= 1024 bits
            for (int t = 0; t < kMaxDim; ++t) sel[t] = 0;
            for (int i = 0; i < k; ++i) {
                int j = (int)rng.below((uint32_t)n);
                while ((sel[j >> 5] >> (j & 31)) & 1u) j = (int)rng.below((uint32_t)n);
                sel[j >> 5] |= 1u << (j & 31);
                out[i] = (int16_t)elem(j);
            }
        }
    }
    // the common case: k cells among those clear in `blocked`; the population is snapshotted first because `blocked`
    // may alias a mask the caller updates afterwards, never during the call
    LLE_HD void sample_free(const uint32_t* blocked, int n, int k, int16_t* out) {
        sample(n, k, [&](int j) { return nth_free(blocked, j); }, out);
    }
    LLE_HD static void sort_small(int16_t* a, int n) {
        for (int i = 1; i < n; ++i) {
            const int16_t v = a[i];
            int j = i - 1;
            for (; j >= 0 && a[j] > v; --j) a[j + 1] = a[j];
            a[j + 1] = v;
        }
    }
    LLE_HD void shuffle16(uint16_t* x, int n) {
        for (int i = n - 1; i >= 1; --i) {
            const int j = (int)rng.below((uint32_t)(i + 1));
            const uint16_t t = x[i];
            x[i] = x[j];
            x[j] = t;
        }
    }

    // ---- beams --------------------------------------------------------------------------------------------------------
    // length of the unobstructed beam from (r, c) towards d up to the grid boundary (geometry.py:24-43 with empty sets)
    LLE_HD int full_len(int r, int c, int d) const { return d == DIR_N ? r : d == DIR_S ? H - 1 - r : d == DIR_E ? W - 1 - c : c; }
    // does the full beam cross a cell of g?
    LLE_HD bool beam_hits(const Grid& g, int r, int c, int d) const {
        switch (d) {
            case DIR_N: return (g.col[c] & low_mask(r)) != 0;
            case DIR_S: return r + 1 < 32 ? (g.col[c] >> (r + 1)) != 0 : false;
            case DIR_E: return c + 1 < 32 ? (g.row[r] >> (c + 1)) != 0 : false;
            default: return (g.row[r] & low_mask(c)) != 0;
        }
    }
    LLE_HD void reserve_beam(int r, int c, int d) {
        const int n = full_len(r, c, d);
        const int dr = d == DIR_N ? -1 : d == DIR_S ? 1 : 0, dc = d == DIR_E ? 1 : d == DIR_W ? -1 : 0;
        for (int k = 1; k <= n; ++k) reserved.set(r + k * dr, c + k * dc);
    }
    LLE_HD bool span_ok(int len) const {  // placements.py:283-297
        if (c.laser_span == SPAN_ANY) return len >= 2;
        if (c.laser_span == SPAN_ACROSS) return true;
        return len >= c.laser_span;
    }
    LLE_HD bool candidate(int r, int q, int d) const {  // placements.py:265-280
        if (reserved.get(r, q)) return false;
        const int len = full_len(r, q, d);
        return len >= 2 && !beam_hits(reserved, r, q, d) && span_ok(len);
    }
    // is cell (pr, pc) on the full beam of (r, c, d)?
    LLE_HD bool on_beam(int pr, int pc, int r, int c, int d) const {
        switch (d) {
            case DIR_N: return pc == c && pr < r;
            case DIR_S: return pc == c && pr > r;
            case DIR_E: return pr == r && pc > c;
            default: return pr == r && pc < c;
        }
    }

    // ---- stages -------------------------------------------------------------------------------------------------------
    LLE_HD void cluster_cells(int ar, int ac, int16_t* out) const {
        int n = 0;
        for (int dr = 0; dr < c.cluster_h && n < c.n_agents; ++dr)
            for (int dc = 0; dc < c.cluster_w && n < c.n_agents; ++dc) out[n++] = (int16_t)((ar + dr) * W + ac + dc);
    }
    LLE_HD bool any_reserved(const int16_t* cells, int n) const {
        for (int k = 0; k < n; ++k)
            if (reserved.get(cells[k] / W, cells[k] % W)) return true;
        return false;
    }
    LLE_HD void reserve_cells(const int16_t* cells, int n) {
        for (int k = 0; k < n; ++k) reserved.set(cells[k] / W, cells[k] % W);
    }

    LLE_HD bool place_agents() {  // placements.py:68-125
        const int n = c.n_agents;
        edge = -1;
        agent_r = agent_c = exit_r = exit_c = -1;
        reserved.clear();
        if (c.rooms)
            for (int r = 0; r < H; ++r)
                for (int q = 0; q < W; ++q)
                    if ((c.room_rows[r] >> q) & 1u) reserved.set(r, q);  // `forbidden`
        if (c.starts == STARTS_RANDOM) {
            const int pop = count_free(reserved.row);
            if (pop < n) return false;
            sample_free(reserved.row, pop, n, agents);
        } else if (c.starts == STARTS_EDGE) {
            edge = (int)rng.below(4);
            const bool side = edge == EDGE_LEFT || edge == EDGE_RIGHT;
            const int fixed = side ? (edge == EDGE_LEFT ? 0 : W - 1) : (edge == EDGE_TOP ? 0 : H - 1);
            const uint32_t taken = side ? reserved.col[fixed] : reserved.row[fixed];
            const uint32_t valid = ~taken & low_mask(side ? H : W);
            const int pop = popc32(valid);
            if (pop < n) return false;
            sample(pop, n, [&](int j) { return nth_set_bit(valid, j); }, lanes);
            sort_small(lanes, n);
            for (int k = 0; k < n; ++k) agents[k] = (int16_t)(side ? lanes[k] * W + fixed : fixed * W + lanes[k]);
        } else {
            if (c.cluster_h > H || c.cluster_w > W) return false;
            agent_r = (int)rng.below((uint32_t)(H - c.cluster_h + 1));  // randint(0, H - ch)
            agent_c = (int)rng.below((uint32_t)(W - c.cluster_w + 1));
            cluster_cells(agent_r, agent_c, agents);
            if (any_reserved(agents, n)) return false;
        }
        reserve_cells(agents, n);
        return true;
    }

    LLE_HD bool exits_on_edge(int e, bool use_lanes) {  // placements.py:177-205
        const int n = c.n_agents;
        const bool side = e == EDGE_LEFT || e == EDGE_RIGHT;
        const int fixed = side ? (e == EDGE_LEFT ? 0 : W - 1) : (e == EDGE_TOP ? 0 : H - 1);
        int16_t ids[kMaxAgents];
        if (use_lanes) {
            for (int k = 0; k < n; ++k) ids[k] = lanes[k];
        } else {
            const int extent = side ? H : W;
            if (extent < n) return false;
            sample(extent, n, [](int j) { return j; }, ids);
            sort_small(ids, n);
        }
        for (int k = 0; k < n; ++k) exits_[k] = (int16_t)(side ? ids[k] * W + fixed : fixed * W + ids[k]);
        return true;
    }

    LLE_HD bool place_exits() {  // placements.py:133-174, 208-246
        const int n = c.n_agents;
        if (c.exits == EXITS_RANDOM) {
            const int pop = count_free(reserved.row);
            if (pop < n) return false;
            sample_free(reserved.row, pop, n, exits_);
        } else if (c.exits == EXITS_EDGE) {
            if (!exits_on_edge((int)rng.below(4), false)) return false;
        } else if (c.exits == EXITS_CLUSTER) {
            if (c.cluster_h > H || c.cluster_w > W) return false;
            bool found = false;
            for (int t = 0; t < 64 && !found; ++t) {
                const int ar = (int)rng.below((uint32_t)(H - c.cluster_h + 1));
                const int ac = (int)rng.below((uint32_t)(W - c.cluster_w + 1));
                cluster_cells(ar, ac, exits_);
                if (!any_reserved(exits_, n)) {
                    exit_r = ar;
                    exit_c = ac;
                    found = true;
                }
            }
            if (!found) return false;
        } else {
            if (edge >= 0) {
                if (!exits_on_edge(edge ^ 1, true)) return false;  // placements.py:23-28: left<->right, top<->bottom
            } else if (agent_r >= 0) {
                int ar = H - c.cluster_h - agent_r, ac = W - c.cluster_w - agent_c;
                ar = ar < H - c.cluster_h ? ar : H - c.cluster_h;
                ac = ac < W - c.cluster_w ? ac : W - c.cluster_w;
                exit_r = ar > 0 ? ar : 0;
                exit_c = ac > 0 ? ac : 0;
                cluster_cells(exit_r, exit_c, exits_);
            } else {
                return false;
            }
        }
        if (any_reserved(exits_, n)) return false;
        reserve_cells(exits_, n);
        return true;
    }

    // placements.py:300-331: shuffle, then take candidates greedily
    LLE_HD bool select_lasers(int n_cand, bool with_beam) {
        const int n = c.n_lasers;
        shuffle16(work, n_cand);
        uint32_t* beams = extra;
        for (int r = 0; r < kMaxDim; ++r) beams[r] = 0;
        int got = 0;
        for (int t = 0; t < n_cand && got < n; ++t) {
            const int cell = work[t] & 1023, d = work[t] >> 10;
            const int r = cell / W, q = cell % W;
            bool skip = (beams[r] >> q) & 1u;
            for (int s = 0; s < got && !skip; ++s) {
                const int sr = laser_cell[s] / W, sc = laser_cell[s] % W;
                skip = (sr == r && sc == q) || on_beam(sr, sc, r, q, d);
            }
            if (skip) continue;
            laser_cell[got] = (int16_t)cell;
            laser_dir[got] = (int8_t)d;
            ++got;
            const int len = full_len(r, q, d);
            const int dr = d == DIR_N ? -1 : d == DIR_S ? 1 : 0, dc = d == DIR_E ? 1 : d == DIR_W ? -1 : 0;
            for (int k = 1; k <= len; ++k) beams[r + k * dr] |= 1u << (q + k * dc);
            reserved.set(r, q);
            if (with_beam) reserve_beam(r, q, d);
        }
        return got >= n;
    }

    // placements.py:463-522
    LLE_HD bool corridor(const int16_t* slots, int d_even, int d_odd, bool fixed_is_row) {
        const int extent = fixed_is_row ? W : H;
        const int min_len = c.laser_span == SPAN_ANY ? 2 : c.laser_span == SPAN_ACROSS ? 0 : c.laser_span;
        for (int k = 0; k < c.n_lasers; ++k) {
            const int slot = slots[k];
            const int d = (k & 1) ? d_odd : d_even;
            const bool forward = d == DIR_E || d == DIR_S;
            int v;
            if (c.laser_span == SPAN_ACROSS) {
                v = forward ? 0 : extent - 1;
                if (fixed_is_row ? reserved.get(slot, v) : reserved.get(v, slot)) return false;
            } else {
                const int lo = forward ? 0 : min_len, hi = forward ? extent - min_len : extent;  // range(lo, hi)
                const uint32_t line = fixed_is_row ? reserved.row[slot] : reserved.col[slot];
                uint32_t valid = hi > lo ? (~line & low_mask(hi) & ~low_mask(lo)) : 0u;
                const int pop = popc32(valid);
                if (pop == 0) return false;
                v = nth_set_bit(valid, (int)rng.below((uint32_t)pop));
            }
            const int r = fixed_is_row ? slot : v, q = fixed_is_row ? v : slot;
            if (!span_ok(full_len(r, q, d))) return false;
            laser_cell[k] = (int16_t)(r * W + q);
            laser_dir[k] = (int8_t)d;
            reserved.set(r, q);
            reserve_beam(r, q, d);
        }
        return true;
    }

    LLE_HD bool place_lasers() {  // placements.py:334-460
        const int n = c.n_lasers;
        if (n == 0) return true;
        if (c.laser_placement == LASERS_FREE) {
            int m = 0;
            for (int r = 0; r < H; ++r)
                for (int q = 0; q < W; ++q) {
                    if (reserved.get(r, q)) continue;
                    for (int d = 0; d < 4; ++d)
                        if (candidate(r, q, d)) work[m++] = (uint16_t)((r * W + q) | (d << 10));
                }
            if (!select_lasers(m, false)) return false;
        } else if (c.laser_placement == LASERS_CROSS_AGENT) {
            // lanes are rows for the left / right edges (beams run S / N along columns), columns otherwise (E / W).
            // A beam from a band outside [min lane, max lane] runs to the boundary, so it crosses every lane: the
            // subset test of placements.py:419 always holds and is not restated.
            const bool vertical = edge == EDGE_LEFT || edge == EDGE_RIGHT;
            const int lo = lanes[0], hi = lanes[c.n_agents - 1];
            const int n_fixed = vertical ? H : W, n_other = vertical ? W : H;
            int m = 0;
            for (int band = 0; band < 2; ++band) {
                const int d = vertical ? (band ? DIR_N : DIR_S) : (band ? DIR_W : DIR_E);
                const int f0 = band ? hi + 1 : 0, f1 = band ? n_fixed : lo;
                for (int f = f0; f < f1; ++f)
                    for (int o = 0; o < n_other; ++o) {
                        const int r = vertical ? f : o, q = vertical ? o : f;
                        if (candidate(r, q, d)) work[m++] = (uint16_t)((r * W + q) | (d << 10));
                    }
            }
            if (m == 0) return false;
            if (!select_lasers(m, true)) return false;
        } else {
            if (agent_r < 0 || exit_r < 0) return false;
            const int bottom = agent_r + c.cluster_h - 1, right = agent_c + c.cluster_w - 1;
            int16_t slots[kMaxDim];
            uint16_t line[kMaxDim];
            if (exit_r - bottom - 1 >= n) {
                int m = 0;
                for (int r = bottom + 1; r < exit_r; ++r) line[m++] = (uint16_t)r;
                shuffle16(line, m);
                for (int k = 0; k < n; ++k) slots[k] = (int16_t)line[k];
                sort_small(slots, n);
                if (!corridor(slots, DIR_E, DIR_W, true)) return false;
            } else if (exit_c - right - 1 >= n) {
                int m = 0;
                for (int q = right + 1; q < exit_c; ++q) line[m++] = (uint16_t)q;
                shuffle16(line, m);
                for (int k = 0; k < n; ++k) slots[k] = (int16_t)line[k];
                sort_small(slots, n);
                if (!corridor(slots, DIR_S, DIR_N, false)) return false;
            } else {
                return false;
            }
        }
        int16_t colours[kMaxAgents];
        sample(c.n_agents, n, [](int j) { return j; }, colours);  // placements.py:366-368
        for (int k = 0; k < n; ++k) laser_colour[k] = (int8_t)colours[k];
        return true;
    }

    // geometry.py:47-99
    LLE_HD void wall_shapes() {
        // offsets (dr, dc) packed as nibbles, low nibble first: shape s has kShapeLen[s] cells
        const uint8_t kShapeLen[9] = {2, 2, 3, 3, 3, 3, 3, 3, 4};
        const uint8_t kShape[9][4] = {{0x00, 0x01, 0, 0},       {0x00, 0x10, 0, 0},       {0x00, 0x01, 0x02, 0}, {0x00, 0x10, 0x20, 0},
                                      {0x00, 0x01, 0x10, 0},    {0x00, 0x01, 0x11, 0},    {0x00, 0x10, 0x11, 0}, {0x01, 0x10, 0x11, 0},
                                      {0x00, 0x01, 0x10, 0x11}};
        const uint8_t kCum[8] = {4, 8, 9, 10, 11, 12, 13, 14};  // cumulative weights; bisect runs over the first 8
        uint32_t* left = extra;                                   // free_set
        const uint32_t wm = low_mask(W);
        int m = 0;
        for (int r = 0; r < H; ++r) {
            left[r] = ~reserved.row[r] & wm;
            for (uint32_t f = left[r]; f; f &= f - 1) work[m++] = (uint16_t)(r * W + ctz32(f));
        }
        for (int r = H; r < kMaxDim; ++r) left[r] = 0;
        shuffle16(work, m);
        int budget = c.n_walls;
        for (int t = 0; t < m && budget > 0; ++t) {
            const int ar = work[t] / W, ac = work[t] % W;
            if (!((left[ar] >> ac) & 1u)) continue;
            int pick[4];
            for (int k = 0; k < 4; ++k) {
                const uint32_t x = rng.random_times_16();
                int s = 0;
                while (s < 8 && kCum[s] <= x) ++s;  // bisect_right(cum_weights, x, 0, 8)
                pick[k] = s;
            }
            int chosen = -1;
            for (int k = 0; k < 4 && chosen < 0; ++k) {
                const int s = pick[k];
                if (kShapeLen[s] > budget) continue;
                bool ok = true;
                for (int e = 0; e < kShapeLen[s] && ok; ++e) {
                    const int r = ar + (kShape[s][e] >> 4), q = ac + (kShape[s][e] & 15);
                    ok = r < H && q < W && ((left[r] >> q) & 1u);
                }
                if (ok) chosen = s;
            }
            if (chosen < 0) {
                left[ar] &= ~(1u << ac);
                walls[ar] |= 1u << ac;
                budget -= 1;
            } else {
                for (int e = 0; e < kShapeLen[chosen]; ++e) {
                    const int r = ar + (kShape[chosen][e] >> 4), q = ac + (kShape[chosen][e] & 15);
                    left[r] &= ~(1u << q);
                    walls[r] |= 1u << q;
                }
                budget -= kShapeLen[chosen];
            }
        }
    }

    // candidates.py:27-41; also collects the lit cells (row masks) for the labels
    LLE_HD bool geometry_valid(uint32_t* lit, uint32_t lit_by[][kMaxDim]) const {
        uint32_t solid[kMaxDim];
        for (int r = 0; r < kMaxDim; ++r) solid[r] = walls[r], lit[r] = 0;
        for (int k = 0; k < c.n_lasers; ++k) solid[laser_cell[k] / W] |= 1u << (laser_cell[k] % W);
        for (int k = 0; k < c.n_lasers; ++k) {
            const int d = laser_dir[k];
            const int dr = d == DIR_N ? -1 : d == DIR_S ? 1 : 0, dc = d == DIR_E ? 1 : d == DIR_W ? -1 : 0;
            int r = laser_cell[k] / W + dr, q = laser_cell[k] % W + dc, len = 0;
            while (r >= 0 && r < H && q >= 0 && q < W && !((solid[r] >> q) & 1u)) {
                lit[r] |= 1u << q;
                if (lit_by) lit_by[k][r] |= 1u << q;
                ++len;
                r += dr;
                q += dc;
            }
            if (len < 2) return false;
        }
        for (int k = 0; k < c.n_agents; ++k)
            if ((lit[exits_[k] / W] >> (exits_[k] % W)) & 1u) return false;
        return true;
    }

    // generator.py:188-228; writes the cell grid on success
    LLE_HD bool run(uint8_t* cells) {
        for (int r = 0; r < kMaxDim; ++r) walls[r] = 0;
        if (!place_agents() || !place_exits() || !place_lasers()) return false;
        int16_t picked[kMaxCells / 2];
        int n_walls = 0;
        if (c.rooms) {
            for (int r = 0; r < H; ++r) walls[r] = c.room_rows[r];
        } else if (c.walls_shapes) {
            wall_shapes();
        } else {
            const int pop = count_free(reserved.row);
            n_walls = c.n_walls < pop ? c.n_walls : pop;
            sample_free(reserved.row, pop, n_walls, picked);
            for (int k = 0; k < n_walls; ++k) walls[picked[k] / W] |= 1u << (picked[k] % W);
        }
        uint32_t blocked[kMaxDim];
        for (int r = 0; r < kMaxDim; ++r) blocked[r] = reserved.row[r] | walls[r];
        const int pop = count_free(blocked);
        if (pop < c.n_gems) return false;
        // gems: at most area - 2 agents cells; they are written straight into the grid
        for (int k = 0; k < H * W; ++k) cells[k] = CELL_FLOOR;
        {
            // sample in chunks is not possible (one Random.sample call), so gems reuse `work` when the pool variant is
            // chosen and write through a small adapter otherwise
            const int k = c.n_gems;
            if (k > 0) {
                if (sample_uses_pool(pop, k)) {
                    for (int j = 0; j < pop; ++j) work[j] = (uint16_t)nth_free(blocked, j);
                    for (int i = 0; i < k; ++i) {
                        const int j = (int)rng.below((uint32_t)(pop - i));
                        cells[work[j]] = CELL_GEM;
                        work[j] = work[pop - i - 1];
                    }
                } else {
                    uint32_t* sel = extra;
                    for (int t = 0; t < kMaxDim; ++t) sel[t] = 0;
                    for (int i = 0; i < k; ++i) {
                        int j = (int)rng.below((uint32_t)pop);
                        while ((sel[j >> 5] >> (j & 31)) & 1u) j = (int)rng.below((uint32_t)pop);
                        sel[j >> 5] |= 1u << (j & 31);
                        cells[nth_free(blocked, j)] = CELL_GEM;
                    }
                }
            }
        }
        uint32_t lit[kMaxDim];
        if (!geometry_valid(lit, nullptr)) return false;
        for (int r = 0; r < H; ++r)
            for (uint32_t f = walls[r]; f; f &= f - 1) cells[r * W + ctz32(f)] = CELL_WALL;
        for (int k = 0; k < c.n_agents; ++k) cells[exits_[k]] = CELL_EXIT;
        for (int k = 0; k < c.n_agents; ++k) cells[agents[k]] = (uint8_t)(CELL_START + k);
        for (int k = 0; k < c.n_lasers; ++k) cells[laser_cell[k]] = (uint8_t)(CELL_SOURCE + 4 * laser_colour[k] + laser_dir[k]);
        return true;
    }

    // ---- labels (heuristic of this library, see include/lle_b200.h) ----------------------------------------------------
    // flood fill over row masks: cells reachable from `start` through cells set in `open`
    LLE_HD void flood(const uint32_t* open, int start, uint32_t* seen) const {
        for (int r = 0; r < kMaxDim; ++r) seen[r] = 0;
        seen[start / W] = 1u << (start % W);
        bool changed = true;
        while (changed) {
            changed = false;
            for (int pass = 0; pass < 2; ++pass)
                for (int k = 0; k < H; ++k) {
                    const int r = pass ? H - 1 - k : k;
                    uint32_t cur = seen[r];
                    uint32_t in = cur | (r > 0 ? seen[r - 1] : 0u) | (r + 1 < H ? seen[r + 1] : 0u);
                    in &= open[r] | cur;
                    for (;;) {  // spread along the row
                        const uint32_t wider = (in | (in << 1) | (in >> 1)) & (open[r] | cur);
                        if (wider == in) break;
                        in = wider;
                    }
                    if (in != cur) {
                        seen[r] = in;
                        changed = true;
                    }
                }
        }
    }
    // perfect matching agents -> distinct exits, adj[a] = mask of exits agent a can use (augmenting paths, breadth first)
    LLE_HD bool matchable(const uint32_t* adj) const {
        const int n = c.n_agents;
        int8_t owner[kMaxAgents], mate[kMaxAgents];
        for (int k = 0; k < n; ++k) owner[k] = mate[k] = -1;
        for (int a0 = 0; a0 < n; ++a0) {
            int8_t queue[kMaxAgents], via[kMaxAgents];
            uint32_t seen = 0;
            int head = 0, tail = 0, found = -1;
            queue[tail++] = (int8_t)a0;
            while (head < tail && found < 0) {
                const int a = queue[head++];
                for (uint32_t m = adj[a] & ~seen; m && found < 0; m &= m - 1) {
                    const int e = ctz32(m);
                    seen |= 1u << e;
                    via[e] = (int8_t)a;
                    if (owner[e] < 0) found = e;
                    else queue[tail++] = owner[e];
                }
            }
            if (found < 0) return false;
            for (int e = found;;) {
                const int a = via[e];
                const int before = mate[a];
                owner[e] = (int8_t)a;
                mate[a] = (int8_t)e;
                if (a == a0) break;
                e = before;
            }
        }
        return true;
    }
    LLE_HD uint8_t labels() {
        uint32_t lit[kMaxDim];
        uint32_t (*lit_by)[kMaxDim] = reinterpret_cast<uint32_t (*)[kMaxDim]>(work);  // n_lasers x 32 words <= 4 KB of `work`
        for (int k = 0; k < c.n_lasers; ++k)
            for (int r = 0; r < kMaxDim; ++r) lit_by[k][r] = 0;
        geometry_valid(lit, lit_by);
        uint32_t open[kMaxDim], seen[kMaxDim], adj[kMaxAgents];
        const uint32_t wm = low_mask(W);
        uint32_t solid[kMaxDim];
        for (int r = 0; r < kMaxDim; ++r) solid[r] = walls[r];
        for (int k = 0; k < c.n_lasers; ++k) solid[laser_cell[k] / W] |= 1u << (laser_cell[k] % W);
        for (int r = 0; r < kMaxDim; ++r) open[r] = r < H ? (~solid[r] & wm) : 0u;
        for (int a = 0; a < c.n_agents; ++a) {
            flood(open, agents[a], seen);
            adj[a] = 0;
            for (int e = 0; e < c.n_agents; ++e)
                if ((seen[exits_[e] / W] >> (exits_[e] % W)) & 1u) adj[a] |= 1u << e;
        }
        if (!matchable(adj)) return 0;
        for (int a = 0; a < c.n_agents; ++a) {
            for (int r = 0; r < kMaxDim; ++r) open[r] = r < H ? (~solid[r] & wm) : 0u;
            for (int k = 0; k < c.n_lasers; ++k)
                if (laser_colour[k] != a)
                    for (int r = 0; r < H; ++r) open[r] &= ~lit_by[k][r];
            flood(open, agents[a], seen);
            adj[a] = 0;
            for (int e = 0; e < c.n_agents; ++e)
                if ((seen[exits_[e] / W] >> (exits_[e] % W)) & 1u) adj[a] |= 1u << e;
        }
        return (uint8_t)(LABEL_WALKABLE | (matchable(adj) ? LABEL_INDEPENDENT : LABEL_NEEDS_BLOCKER));
    }
};

// One chain = `WorldGenerator.generate(max_attempts, seed)` (generator.py:268-284): seed once, then up to max_attempts
// attempts drawing from the same stream; max_attempts = 1 is `_try_generate(seed)`.  A layout is accepted when its label
// byte contains every bit of `require` (the place of `_accept_world`, generator.py:256-266, which draws no random number).
// Outputs: cells[H*W], *status (1 = layout, 0 = none), *label, *tries (attempts used).
struct TablePtr {  // the init_genrand table through a pointer (host); the kernel reads it from constant memory
    const uint32_t* p;
    LLE_HD uint32_t operator()(int i) const { return p[i]; }
};

template <class Rng, class Base>
LLE_HD void chain(const Config& cfg, Base mt_base, uint64_t seed, int max_attempts, uint8_t require, Rng& rng,
                  uint16_t* work, uint8_t* cells, uint8_t* status, uint8_t* label, int32_t* tries) {
    rng.seed(seed, mt_base);
    Attempt<Rng> at(cfg, rng, work);
    bool ok = false;
    uint8_t lab = 0;
    int t = 0;
    while (t < max_attempts && !ok) {
        ++t;
        if (!at.run(cells)) continue;
        lab = at.labels();
        ok = (lab & require) == require;
    }
    *status = ok ? 1 : 0;
    *label = ok ? lab : 0;
    *tries = t;
    if (!ok)
        for (int k = 0; k < cfg.height * cfg.width; ++k) cells[k] = CELL_FLOOR;
}

}  // namespace llegen
