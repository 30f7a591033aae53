// Host-only part of the layout generator shared by the library (gen_world.cu) and the CPU test shim
// (tests/host_shim/gen_host.cpp): WorldGenerator.__init__'s validation, the structural room walls, the MT19937 base table.
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lle_b200.h"
#include "gen_core.cuh"

namespace llegen {

// init_genrand(19650218): the state init_by_array starts from (Modules/_randommodule.c)
inline void mt_base_table(uint32_t* mt) {
    mt[0] = 19650218u;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
}

// placements.py:546-647 place_room_walls as row masks
inline void room_rows(const lle_gen_options& o, uint32_t* rows) {
    const int H = o.height, W = o.width;
    auto stripes = [](int total, int n, std::vector<int>& size, std::vector<int>& start, std::vector<int>& divider) {
        const int inner = total - (n - 1);
        const int base = inner >= 0 ? inner / n : 0, rem = inner >= 0 ? inner % n : 0;
        int at = 0;
        for (int k = 0; k < n; ++k) {
            size.push_back(base + (k < rem ? 1 : 0));
            start.push_back(at);
            at += size.back();
            if (k < n - 1) divider.push_back(at++);
        }
    };
    std::vector<int> rh, rs, dr, cw, cs, dc;
    stripes(H, o.n_rooms_rows, rh, rs, dr);
    stripes(W, o.n_rooms_cols, cw, cs, dc);
    for (int r = 0; r < 32; ++r) rows[r] = 0;
    const uint32_t wm = W >= 32 ? 0xffffffffu : ((1u << W) - 1u);
    for (int r : dr)
        if (r >= 0 && r < H) rows[r] = wm;
    for (int c : dc)
        if (c >= 0 && c < W)
            for (int r = 0; r < H; ++r) rows[r] |= 1u << c;
    const int half = o.door_size / 2;
    for (int r : dr)
        for (size_t k = 0; k < cw.size(); ++k) {
            const int mid = cs[k] + (cw[k] - 1) / 2;
            for (int off = -half; off < o.door_size - half; ++off) {
                const int c = mid + off;
                if (c >= cs[k] && c < cs[k] + cw[k] && r >= 0 && r < H && c >= 0 && c < W) rows[r] &= ~(1u << c);
            }
        }
    for (int c : dc)
        for (size_t k = 0; k < rh.size(); ++k) {
            const int mid = rs[k] + (rh[k] - 1) / 2;
            for (int off = -half; off < o.door_size - half; ++off) {
                const int r = mid + off;
                if (r >= rs[k] && r < rs[k] + rh[k] && r >= 0 && r < H && c >= 0 && c < W) rows[r] &= ~(1u << c);
            }
        }
}


// generator.py:116-181, same order and wording.  Returns 0 or an LLE_* code with the message in `why`.
inline int resolve_config(const lle_gen_options& o, Config& c, std::string& why) {
    auto bad = [&](const std::string& m) { why = m; return (int)LLE_INVALID_ARGUMENT; };
    auto limit = [&](const std::string& m) { why = m; return (int)LLE_LIMIT_EXCEEDED; };
    // generator.py:116-181, same order and wording
    if (o.exits == LLE_GEN_EXITS_OPPOSITE && o.starts != LLE_GEN_STARTS_EDGE && o.starts != LLE_GEN_STARTS_CLUSTERED)
        return bad("exits='opposite' requires starts='edge' or starts='clustered', not 'random'.");
    if (o.laser_placement == LLE_GEN_LASERS_CROSS_AGENT && o.starts != LLE_GEN_STARTS_EDGE) return bad("laser_placement='cross-agent' requires starts='edge'.");
    if (o.laser_placement == LLE_GEN_LASERS_CROSS_CLUSTER && o.starts != LLE_GEN_STARTS_CLUSTERED)
        return bad("laser_placement='cross-cluster' requires starts='clustered'.");
    if (o.laser_placement == LLE_GEN_LASERS_CROSS_CLUSTER && o.exits != LLE_GEN_EXITS_OPPOSITE && o.exits != LLE_GEN_EXITS_CLUSTER)
        return bad("laser_placement='cross-cluster' requires exits='opposite' or exits='cluster'.");
    if (o.laser_span != LLE_GEN_SPAN_ANY && o.laser_span != LLE_GEN_SPAN_ACROSS && o.laser_span < 2)
        return bad("laser_span must be >= 2, got " + std::to_string(o.laser_span) + ".");
    if (o.width < 1) return bad("Grid width must be >= 1. Got " + std::to_string(o.width));
    if (o.height < 1) return bad("Grid height must be >= 1. Got " + std::to_string(o.height));
    if (o.width > llegen::kMaxDim || o.height > llegen::kMaxDim) return limit("lle_gen: height and width are limited to 32");
    const int area = o.width * o.height;
    if (o.n_agents < 1) return bad("agents must be >= 1. Got " + std::to_string(o.n_agents));
    if (o.n_agents > llegen::kMaxAgents) return limit("lle_gen: at most 32 agents");
    if (o.n_lasers < 0) return bad("lasers must be >= 0. Got " + std::to_string(o.n_lasers));
    if (o.n_lasers > o.n_agents)
        return bad("lasers must be <= agents (one laser source per colour). Got lasers=" + std::to_string(o.n_lasers) + ", agents=" + std::to_string(o.n_agents) + ".");
    if (o.n_gems < 0) return bad("gems must be >= 0. Got " + std::to_string(o.n_gems));
    if (o.n_gems > area - 2 * o.n_agents)
        return bad("gems must be <= grid cells minus start and exit cells (" + std::to_string(area - 2 * o.n_agents) + "). Got gems=" + std::to_string(o.n_gems) + ".");
    if (o.starts < 0 || o.starts > 2 || o.exits < 0 || o.exits > 3 || o.laser_placement < 0 || o.laser_placement > 2)
        return bad("lle_gen: unknown starts / exits / laser_placement mode");
    std::memset(&c, 0, sizeof(c));
    c.width = o.width, c.height = o.height, c.n_agents = o.n_agents, c.starts = o.starts, c.exits = o.exits;
    c.n_lasers = o.n_lasers, c.n_gems = o.n_gems, c.laser_placement = o.laser_placement, c.laser_span = o.laser_span;
    c.walls_shapes = o.walls_shapes ? 1 : 0;
    c.cluster_h = o.cluster_h, c.cluster_w = o.cluster_w;
    const bool clustered = o.starts == LLE_GEN_STARTS_CLUSTERED || o.exits == LLE_GEN_EXITS_CLUSTER;
    if (clustered && (o.cluster_h < 1 || o.cluster_w < 1 || o.cluster_h * o.cluster_w < o.n_agents))
        return bad("lle_gen: cluster_h x cluster_w must hold n_agents cells");
    if (o.n_rooms_rows > 0) {
        if (o.n_rooms_cols < 1 || o.door_size < 0) return bad("lle_gen: rooms mode needs n_rooms_cols >= 1 and door_size >= 0");
        c.rooms = 1;
        c.n_walls = 0;
        room_rows(o, c.room_rows);
    } else {
        const int n_walls = o.n_walls == LLE_GEN_WALLS_AUTO ? area / 10 : o.n_walls;
        if (n_walls < 0) return bad("num_walls must be >= 0. Got " + std::to_string(n_walls));
        if (2 * n_walls >= area) return bad("num_walls must be < size/2. Got num_walls=" + std::to_string(n_walls) + ", size=" + std::to_string(area));
        const int needed = 2 * o.n_agents + n_walls + o.n_lasers + o.n_gems;
        if (needed > area) return bad("layout requires " + std::to_string(needed) + " unique cells, but grid has only " + std::to_string(area));
        c.n_walls = n_walls;
    }
    return 0;
}

}  // namespace llegen
