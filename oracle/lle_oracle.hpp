// ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
//
// CPU restatement of the yamoling/lle v2.11.4 `World` engine and of the two Python
// pieces every RL step pays for (layered observation, SingleObjective reward/done).
// It deliberately keeps the reference's *object* design (grid of tile variants, a beam
// `vector<bool>` shared by the tiles of one laser source, sequential
// leave -> pre_enter -> enter, the `while agent_died` re-pass) so that it is an
// independent statement of the semantics against which the bitmask CUDA path is checked.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.  The product (lle_b200/) never links or loads it.
//
// Parity pinning: the Rust crate cannot be compiled in this image (no cargo/rustc), so
// there is no oracle/_ref.  The oracle is pinned against the reference's own
// known-answer tests, transcribed in tests/test_oracle_*.py.
// Random start positions (`rand::StdRng` shuffle, utils/mod.rs:63-65) are PARITY
// UNPINNED and rejected here: every map in scope has exactly one start per agent.
//
// Every function cites the reference file:line it follows (paths relative to the
// reference repository root).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace lle_oracle {

using AgentId = size_t;
using LaserId = size_t;

// src/position.rs:8
struct Position {
    size_t i = 0, j = 0;
    bool operator==(const Position& o) const { return i == o.i && j == o.j; }
    bool operator!=(const Position& o) const { return !(*this == o); }
};

// src/action.rs:9-15 (discriminants) and :18-26 (deltas)
enum class Action : uint8_t { North = 0, South = 1, East = 2, West = 3, Stay = 4 };
inline std::pair<int, int> action_delta(Action a) {
    switch (a) {
        case Action::North: return {-1, 0};
        case Action::South: return {1, 0};
        case Action::East: return {0, 1};
        case Action::West: return {0, -1};
        default: return {0, 0};
    }
}

// src/core/tiles/direction.rs:9-29
enum class Direction : uint8_t { North = 0, East = 1, South = 2, West = 3 };
inline std::pair<int, int> direction_delta(Direction d) {
    switch (d) {
        case Direction::North: return {-1, 0};
        case Direction::East: return {0, 1};
        case Direction::South: return {1, 0};
        default: return {0, -1};
    }
}

// src/core/event.rs:4-8 ; numbering follows the python EventType enum (pyevent.rs:9-17)
enum class EventType : uint8_t { AgentExit = 0, GemCollected = 1, AgentDied = 2 };
struct WorldEvent {
    EventType type;
    AgentId agent_id;
    bool operator==(const WorldEvent& o) const { return type == o.type && agent_id == o.agent_id; }
};

// src/core/world_state.rs:5-9
struct WorldState {
    std::vector<Position> agents_positions;
    std::vector<bool> gems_collected;
    std::vector<bool> agents_alive;
    bool operator==(const WorldState& o) const {
        return agents_positions == o.agents_positions && gems_collected == o.gems_collected &&
               agents_alive == o.agents_alive;
    }
};

// src/core/parsing/errors.rs:5-70
enum class ParseErrorKind : int {
    EmptyWorld = 1,
    NoAgents,
    InvalidTile,
    InvalidFileName,
    InvalidLevel,
    NotEnoughExitTiles,
    NotEnoughStartTiles,
    DuplicateStartTile,
    InconsistentDimensions,
    InvalidLaserSourceAgentId,
    InvalidAgentId,
    InvalidDirection,
    AgentWithoutStart,
    MissingWidth,
    Unsupported,  // oracle-only: TOML / random starts are out of the pinned scope
};
struct ParseError : std::runtime_error {
    ParseErrorKind kind;
    // free-form payload (a, b, c) mirrors the numeric fields of the Rust variant
    long a = 0, b = 0, c = 0;
    ParseError(ParseErrorKind k, const std::string& msg, long a_ = 0, long b_ = 0, long c_ = 0)
        : std::runtime_error(msg), kind(k), a(a_), b(b_), c(c_) {}
};

// src/core/errors.rs:6-45
enum class RuntimeErrorKind : int {
    InvalidAction = 101,
    InvalidNumberOfGems,
    InvalidNumberOfAgents,
    InvalidAgentPosition,
    OutOfWorldPosition,
    InvalidNumberOfActions,
    InvalidWorldState,
    TileNotWalkable,
    Panic,  // a Rust panic!/expect/unwrap on this path
};
struct RuntimeWorldError : std::runtime_error {
    RuntimeErrorKind kind;
    long a = 0, b = 0, c = 0;
    RuntimeWorldError(RuntimeErrorKind k, const std::string& msg, long a_ = 0, long b_ = 0, long c_ = 0)
        : std::runtime_error(msg), kind(k), a(a_), b(b_), c(c_) {}
};

// ---- Philox4x32-10 (Salmon et al., SC'11), restated from the published algorithm.
// Pinned by the Random123 known-answer vectors in tests/test_oracle_philox.py.
struct Philox {
    static void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
    }
    static void run(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
        uint32_t k0 = key[0], k1 = key[1];
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// src/agent.rs:6-49
struct Agent {
    AgentId id;
    bool dead = false;
    bool arrived = false;
    explicit Agent(AgentId id_) : id(id_) {}
    void reset() { dead = false; arrived = false; }
    void die() { dead = true; }
    void arrive() { arrived = true; }
    bool has_arrived() const { return arrived; }
    bool is_dead() const { return dead; }
    bool is_alive() const { return !dead; }
};

// src/core/tiles/laser.rs:15-86 — state shared (Rc) by the source and all its laser tiles
struct LaserBeam {
    std::vector<bool> beam;
    bool enabled = true;
    AgentId agent_id;
    Direction direction;
    LaserId laser_id;
    LaserBeam(size_t size, AgentId agent, Direction dir, LaserId id)
        : beam(size, true), agent_id(agent), direction(dir), laser_id(id) {}
    bool is_on(size_t offset) const { return beam[offset]; }  // laser.rs:42
    void turn_on(size_t offset) {                             // laser.rs:50-55
        if (!enabled) return;
        std::fill(beam.begin() + offset, beam.end(), true);
    }
    void turn_off(size_t offset) { std::fill(beam.begin() + offset, beam.end(), false); }  // :57-59
    void enable() { enabled = true; turn_on(0); }                                        // :69-72
    void disable() { enabled = false; turn_off(0); }                                     // :74-77
};

// src/core/tiles/tile.rs:10-18 — one struct standing in for the Rust enum; `Laser`
// (laser.rs:88-92) owns the tile it wraps, so crossing beams nest.
struct Tile {
    enum Kind : uint8_t { Gem, Floor, Wall, Void, Exit, Laser, LaserSource } kind = Floor;
    std::optional<AgentId> slot;      // Floor/Exit {agent}, Gem.agent (gem.rs:8), Void.agent (void.rs:5)
    bool collected = false;           // Gem.collected (gem.rs:9)
    std::shared_ptr<LaserBeam> beam;  // Laser.beam / LaserSource.beam
    std::unique_ptr<Tile> wrapped;    // Laser.wrapped
    size_t offset = 0;                // Laser.offset

    static Tile make(Kind k) { Tile t; t.kind = k; return t; }

    // ---- Laser helpers, laser.rs:128-166
    bool laser_is_on() const { return beam->is_on(offset); }
    void laser_turn_on() {  // laser.rs:157-162 : no-op when this tile's bit is already on
        if (laser_is_on()) return;
        beam->turn_on(offset);
    }
    void laser_turn_off() { beam->turn_off(offset); }  // :164-166

    // tile.rs:21-27 ; laser.rs:173-182
    bool pre_enter(const Agent& agent) {  // false <=> Err(TileNotWalkable)
        switch (kind) {
            case Laser: {
                bool res = wrapped->pre_enter(agent);
                if (!beam->enabled) return res;
                if (agent.is_alive() && agent.id == beam->agent_id) laser_turn_off();
                return res;
            }
            case Wall:
            case LaserSource: return false;
            default: return true;
        }
    }

    // tile.rs:29-50 ; laser.rs:184-197 ; gem.rs:26-35 ; void.rs:13-22
    std::optional<WorldEvent> enter(Agent& agent) {
        switch (kind) {
            case Wall:
            case LaserSource:
                throw RuntimeWorldError(RuntimeErrorKind::Panic, "Cannot enter a wall or a laser source");
            case Exit:
                slot = agent.id;
                if (!agent.has_arrived()) {
                    agent.arrive();
                    return WorldEvent{EventType::AgentExit, agent.id};
                }
                return std::nullopt;
            case Floor: slot = agent.id; return std::nullopt;
            case Void:
                slot = agent.id;
                if (agent.is_alive()) {
                    agent.die();
                    return WorldEvent{EventType::AgentDied, agent.id};
                }
                return std::nullopt;
            case Laser:
                if (laser_is_on() && agent.id != beam->agent_id) {
                    if (agent.is_alive()) {
                        agent.die();
                        laser_turn_on();
                        return WorldEvent{EventType::AgentDied, agent.id};
                    }
                    return std::nullopt;
                }
                return wrapped->enter(agent);
            case Gem:
                slot = agent.id;
                if (!collected) {
                    collected = true;
                    return WorldEvent{EventType::GemCollected, agent.id};
                }
                return std::nullopt;
        }
        return std::nullopt;
    }

    // tile.rs:52-61 ; laser.rs:199-202
    AgentId leave() {
        switch (kind) {
            case Wall:
            case LaserSource:
                throw RuntimeWorldError(RuntimeErrorKind::Panic, "Cannot leave a wall or a laser source");
            case Laser: laser_turn_on(); return wrapped->leave();
            default: {
                if (!slot.has_value()) throw RuntimeWorldError(RuntimeErrorKind::Panic, "No agent to leave");
                AgentId a = *slot;
                slot.reset();
                return a;
            }
        }
    }

    // tile.rs:63-73
    bool is_walkable() const { return kind != Wall && kind != LaserSource; }

    // tile.rs:75-84 ; laser.rs:168-171 ; gem.rs:21-24
    void reset() {
        switch (kind) {
            case Gem: collected = false; slot.reset(); break;
            case LaserSource:
            case Wall: break;
            case Laser: laser_turn_on(); wrapped->reset(); break;
            default: slot.reset(); break;
        }
    }

    // tile.rs:86-99
    std::optional<AgentId> agent() const {
        switch (kind) {
            case Wall:
            case LaserSource: return std::nullopt;
            case Laser: return wrapped->agent();
            default: return slot;
        }
    }
    bool is_occupied() const { return agent().has_value(); }

    // laser.rs:112-118 : the gem under (possibly nested) lasers
    const Tile* gem() const {
        if (kind == Gem) return this;
        if (kind == Laser) return wrapped->gem();
        return nullptr;
    }
    Tile* gem_mut() {
        if (kind == Gem) return this;
        if (kind == Laser) return wrapped->gem_mut();
        return nullptr;
    }
};

// src/core/parsing/laser_config.rs:11-16
struct LaserConfig {
    Direction direction;
    AgentId agent_id;
    LaserId laser_id;
};

// src/utils/mod.rs:18-36
template <class T>
inline void find_duplicates_into(const std::vector<T>& input, std::vector<bool>& result) {
    result.assign(input.size(), false);
    for (size_t i = 0; i < input.size(); ++i) {
        if (!result[i]) {
            for (size_t j = i + 1; j < input.size(); ++j) {
                if (input[i] == input[j]) {
                    result[i] = true;
                    result[j] = true;
                }
            }
        }
    }
}

// view of one laser tile as `World::lasers()` hands it out (pylaser.rs:44-54 snapshot)
struct LaserView {
    Position pos;
    LaserId laser_id;
    AgentId agent_id;
    Direction direction;
    bool is_on;
    bool is_enabled;
};

class World {
  public:
    // src/core/world.rs:21-44
    size_t width = 0, height = 0;
    std::vector<std::vector<Tile>> grid;
    std::vector<Agent> agents;
    std::vector<Position> laser_source_positions;
    std::vector<Position> lasers_positions;  // HashSet in Rust (world_config.rs:204); kept row-major here
    std::vector<Position> gems_positions;
    std::vector<std::vector<Position>> random_start_positions;
    std::vector<Position> void_positions;
    std::vector<Position> exits;
    std::vector<Position> agents_positions;
    std::vector<Position> wall_positions;
    std::vector<std::vector<Action>> available_actions;
    std::vector<Position> start_positions;
    std::vector<bool> conflict_scratch;
    std::string source_text;  // what the world was parsed from (pickle/clone convenience)
    // Start sampling (random starts, TOML v2 maps).  The reference owns a rand::StdRng seeded by World::seed (world.rs:80,
    // :92-96) whose stream is PARITY UNPINNED (no lockfile, no test pins the values).  The device library defines its own
    // Philox stream (include/lle_b200.h, lle_vec_reset); the harness sets its counter words before every reset.
    uint64_t rng_seed = 0;
    uint32_t rng_env = 0, rng_t = 0, rng_epoch = 1;
    // Bookkeeping that is NOT in the reference: the move_agents pass (1, 2, ...) that emitted each
    // event of the last step(), so tests can check the device's (pass, agent) event encoding.
    std::vector<int> last_event_pass;

    size_t n_agents() const { return agents.size(); }
    size_t n_gems() const { return gems_positions.size(); }

    const Tile* at(const Position& p) const {  // world.rs:391-399
        if (p.i >= height || p.j >= width) return nullptr;
        return &grid[p.i][p.j];
    }

    // world.rs:129-139 — gems may be wrapped by lasers
    std::vector<const Tile*> gems() const {
        std::vector<const Tile*> res;
        for (const auto& pos : gems_positions) res.push_back(grid[pos.i][pos.j].gem());
        return res;
    }

    // world.rs:195-234: the current exits become floor tiles, the given positions (which must be floor tiles, possibly
    // under a laser) become exits; agents standing there stay on the tile.
    void set_exit_positions(const std::vector<Position>& new_exits) {
        if (new_exits.size() < n_agents())
            throw ParseError(ParseErrorKind::NotEnoughExitTiles, "NotEnoughExitTiles", (long)n_agents(), (long)new_exits.size());
        auto replace = [&](const Position& pos, Tile::Kind from, Tile::Kind to, const char* what) {
            Tile& tile = grid[pos.i][pos.j];
            Tile* target = &tile;
            if (tile.kind == Tile::Laser) target = tile.wrapped.get();  // laser.set_tile(...) replaces the wrapped tile (:208-213)
            if (target->kind != from) throw RuntimeWorldError(RuntimeErrorKind::Panic, what);
            Tile repl = Tile::make(to);
            repl.slot = tile.kind == Tile::Laser ? tile.agent() : target->slot;  // `agent: laser.agent()` / `agent`
            *target = std::move(repl);
        };
        for (const auto& pos : exits) replace(pos, Tile::Exit, Tile::Floor, "Tile is not an exit");
        exits = new_exits;
        for (const auto& pos : exits) replace(pos, Tile::Floor, Tile::Exit, "Tile is not a floor");
    }

    // world.rs:159-172 — only the outer laser and the one directly under it are listed
    std::vector<LaserView> lasers() const {
        std::vector<LaserView> res;
        auto view = [](const Position& p, const Tile& t) {
            return LaserView{p, t.beam->laser_id, t.beam->agent_id, t.beam->direction, t.laser_is_on(),
                             t.beam->enabled};
        };
        for (const auto& pos : lasers_positions) {
            const Tile& t = grid[pos.i][pos.j];
            if (t.kind != Tile::Laser) throw RuntimeWorldError(RuntimeErrorKind::Panic, "unreachable");
            res.push_back(view(pos, t));
            if (t.wrapped->kind == Tile::Laser) res.push_back(view(pos, *t.wrapped));
        }
        return res;
    }

    // world.rs:141-149
    std::shared_ptr<LaserBeam> source_beam(size_t idx) const {
        const Position& p = laser_source_positions.at(idx);
        return grid[p.i][p.j].beam;
    }

    // world.rs:265-275 — NOTE: counts top-level Gem tiles only (gems under a beam are ignored)
    size_t n_gems_collected() const {
        size_t res = 0;
        for (const auto& pos : gems_positions) {
            const Tile& t = grid[pos.i][pos.j];
            if (t.kind == Tile::Gem && t.collected) res++;
        }
        return res;
    }
    size_t n_agents_arrived() const {  // world.rs:277-279
        size_t n = 0;
        for (const auto& a : agents) n += a.has_arrived();
        return n;
    }

    // world.rs:343-363
    void compute_available_actions() {
        available_actions.resize(agents.size());
        for (size_t a = 0; a < agents.size(); ++a) {
            auto& acts = available_actions[a];
            acts.clear();
            acts.push_back(Action::Stay);
            if (agents[a].is_alive() && !agents[a].has_arrived()) {
                for (Action action : {Action::North, Action::East, Action::South, Action::West}) {
                    auto d = action_delta(action);
                    long i = (long)agents_positions[a].i + d.first;
                    long j = (long)agents_positions[a].j + d.second;
                    if (i < 0 || j < 0) continue;  // position.rs:55-62 -> Err
                    Position p{(size_t)i, (size_t)j};
                    const Tile* tile = at(p);
                    if (tile && tile->is_walkable() && !tile->is_occupied()) acts.push_back(action);
                }
            }
        }
    }

    // world.rs:365-378
    void solve_vertex_conflicts(std::vector<Position>& new_pos) {
        bool conflict = true;
        while (conflict) {
            conflict = false;
            find_duplicates_into(new_pos, conflict_scratch);
            for (size_t i = 0; i < conflict_scratch.size(); ++i) {
                if (conflict_scratch[i]) {
                    conflict = true;
                    new_pos[i] = agents_positions[i];
                }
            }
        }
    }

    // world.rs:411-432
    void reset() {
        for (auto& row : grid)
            for (auto& tile : row) tile.reset();
        for (auto& agent : agents) agent.reset();
        start_positions = sample_different();
        agents_positions = start_positions;
        for (size_t a = 0; a < agents.size(); ++a) {
            const Position& pos = agents_positions[a];
            if (!grid[pos.i][pos.j].pre_enter(agents[a]))
                throw RuntimeWorldError(RuntimeErrorKind::Panic, "The agent should be able to pre-enter");
        }
        for (size_t a = 0; a < agents.size(); ++a) {
            const Position& pos = agents_positions[a];
            grid[pos.i][pos.j].enter(agents[a]);
        }
        compute_available_actions();
    }

    // world.rs:435-475
    std::vector<WorldEvent> step(const std::vector<Action>& actions) {
        if (n_agents() != actions.size())
            throw RuntimeWorldError(RuntimeErrorKind::InvalidNumberOfActions, "InvalidNumberOfActions",
                                    (long)actions.size(), (long)n_agents());
        for (size_t a = 0; a < actions.size(); ++a) {
            const auto& av = available_actions[a];
            if (std::find(av.begin(), av.end(), actions[a]) == av.end())
                throw RuntimeWorldError(RuntimeErrorKind::InvalidAction, "InvalidAction", (long)a,
                                        (long)actions[a]);
        }
        std::vector<Position> new_positions;
        for (size_t a = 0; a < actions.size(); ++a) {
            auto d = action_delta(actions[a]);
            long i = (long)agents_positions[a].i + d.first, j = (long)agents_positions[a].j + d.second;
            if (i < 0 || j < 0)
                throw RuntimeWorldError(RuntimeErrorKind::OutOfWorldPosition, "OutOfWorldPosition", i, j);
            new_positions.push_back(Position{(size_t)i, (size_t)j});
        }
        solve_vertex_conflicts(new_positions);
        auto [events, agent_died] = move_agents(new_positions);
        agents_positions = new_positions;
        int pass = 1;
        last_event_pass.assign(events.size(), pass);
        while (agent_died) {
            auto [more, died2] = move_agents(new_positions);
            events.insert(events.end(), more.begin(), more.end());
            last_event_pass.insert(last_event_pass.end(), more.size(), ++pass);
            agent_died = died2;
        }
        compute_available_actions();
        return events;
    }

    // world.rs:477-505
    std::pair<std::vector<WorldEvent>, bool> move_agents(const std::vector<Position>& new_positions) {
        for (size_t a = 0; a < agents.size(); ++a) {
            if (agents[a].is_alive()) {
                const Position& pos = agents_positions[a];
                grid[pos.i][pos.j].leave();
            }
        }
        for (size_t a = 0; a < agents.size(); ++a) {
            const Position& pos = new_positions[a];
            if (!grid[pos.i][pos.j].pre_enter(agents[a]))
                throw RuntimeWorldError(RuntimeErrorKind::Panic,
                                        "When moving agents, the pre-enter should not fail");
        }
        std::vector<WorldEvent> events;
        bool agent_died = false;
        for (size_t a = 0; a < agents.size(); ++a) {
            const Position& pos = new_positions[a];
            auto ev = grid[pos.i][pos.j].enter(agents[a]);
            if (ev) {
                if (ev->type == EventType::AgentDied) agent_died = true;
                events.push_back(*ev);
            }
        }
        return {events, agent_died};
    }

    // world.rs:507-513
    WorldState get_state() const {
        WorldState s;
        s.agents_positions = agents_positions;
        for (const Tile* g : gems()) s.gems_collected.push_back(g->collected);
        for (const auto& a : agents) s.agents_alive.push_back(a.is_alive());
        return s;
    }

    // world.rs:515-597
    std::vector<WorldEvent> set_state(const WorldState& state) {
        if (state.gems_collected.size() != n_gems())
            throw RuntimeWorldError(RuntimeErrorKind::InvalidNumberOfGems, "InvalidNumberOfGems",
                                    (long)state.gems_collected.size(), (long)n_gems());
        if (state.agents_positions.size() != n_agents())
            throw RuntimeWorldError(RuntimeErrorKind::InvalidNumberOfAgents, "InvalidNumberOfAgents",
                                    (long)state.agents_positions.size(), (long)n_agents());
        std::vector<bool> dup;
        find_duplicates_into(state.agents_positions, dup);
        if (std::any_of(dup.begin(), dup.end(), [](bool b) { return b; }))
            throw RuntimeWorldError(RuntimeErrorKind::InvalidWorldState,
                                    "There are two agents at the same position");
        for (const auto& pos : state.agents_positions)
            if (pos.i >= height || pos.j >= width)
                throw RuntimeWorldError(RuntimeErrorKind::OutOfWorldPosition, "OutOfWorldPosition",
                                        (long)pos.i, (long)pos.j);
        WorldState current_state = get_state();
        for (auto& row : grid)
            for (auto& tile : row) tile.reset();
        // :550-554 — only *top-level* Gem tiles can be force-collected
        for (size_t g = 0; g < gems_positions.size(); ++g) {
            const Position& pos = gems_positions[g];
            if (state.gems_collected[g] && grid[pos.i][pos.j].kind == Tile::Gem)
                grid[pos.i][pos.j].collected = true;
        }
        // :555-569 — pre_enter sees the agents' *current* (pre-call) alive flags
        for (size_t a = 0; a < agents.size(); ++a) {
            const Position& pos = state.agents_positions[a];
            if (!grid[pos.i][pos.j].pre_enter(agents[a])) {
                set_state(current_state);  // .unwrap()
                throw RuntimeWorldError(RuntimeErrorKind::InvalidAgentPosition, "The tile is not walkable",
                                        (long)pos.i, (long)pos.j);
            }
        }
        agents_positions = state.agents_positions;
        std::vector<WorldEvent> events;
        for (size_t a = 0; a < agents.size(); ++a) {
            const Position& pos = agents_positions[a];
            agents[a].reset();
            auto ev = grid[pos.i][pos.j].enter(agents[a]);
            if (ev) events.push_back(*ev);
            if (!state.agents_alive[a]) agents[a].die();
        }
        WorldState actual = get_state();
        if (!(actual == state))
            throw RuntimeWorldError(RuntimeErrorKind::InvalidWorldState,
                                    "The given state is invalid (e.g. an agent whose alive status was "
                                    "set to `true` died).");
        compute_available_actions();
        return events;
    }

  private:
    // src/utils/mod.rs:39-86.  With one candidate per agent no RNG is consumed (:62-65 never shuffles) and the backtracking
    // and the final `result[id]` re-indexing (:82) are kept as written.  With several candidates the draw order follows
    // the device library's documented contract (attempts, Philox words, cyclic probing; after 16 dead-ended attempts the
    // fixed assignment of a bipartite matching), because rand::StdRng::shuffle cannot be reproduced.
    std::vector<Position> sample_different() const {
        const auto& starts = random_start_positions;
        size_t n = starts.size();
        std::vector<size_t> idx(n);
        for (size_t i = 0; i < n; ++i) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(),
                         [&](size_t x, size_t y) { return starts[x].size() < starts[y].size(); });
        bool random = false;
        for (const auto& c : starts) random = random || c.size() != 1;
        if (random) return sample_different_philox(idx);
        std::vector<Position> result;
        assign_positions(0, idx, starts, result);
        if (result.size() != n) throw RuntimeWorldError(RuntimeErrorKind::Panic, "Could not assign positions to agents");
        std::vector<Position> out;
        for (size_t id : idx) out.push_back(result[id]);
        return out;
    }
    std::vector<Position> sample_different_philox(const std::vector<size_t>& order) const {
        const auto& starts = random_start_positions;
        const size_t n = starts.size();
        const uint32_t key[2] = {(uint32_t)rng_seed, (uint32_t)(rng_seed >> 32)};
        for (uint32_t attempt = 0; attempt < 16; ++attempt) {
            std::vector<Position> pick(n);
            std::vector<bool> has(n, false);
            bool failed = false;
            for (size_t a : order) {
                const uint32_t ctr[4] = {rng_env, rng_t, 0x40000000u | (attempt << 8) | (uint32_t)(a >> 2), rng_epoch};
                uint32_t out[4];
                Philox::run(ctr, key, out);
                const uint32_t cnt = (uint32_t)starts[a].size();
                const uint32_t k = (uint32_t)(((uint64_t)out[a & 3] * cnt) >> 32);
                bool found = false;
                for (uint32_t probe = 0; probe < cnt && !found; ++probe) {
                    const Position& c = starts[a][(k + probe) % cnt];
                    bool taken = false;
                    for (size_t b = 0; b < n; ++b) taken = taken || (has[b] && pick[b] == c);
                    if (!taken) { pick[a] = c; has[a] = true; found = true; }
                }
                if (!found) { failed = true; break; }
            }
            if (!failed) return pick;
        }
        // last resort: a fixed assignment by augmenting paths over the agents in id order, candidates in list order
        const size_t cells = width * height;
        std::vector<long> owner(cells, -1);
        std::function<bool(size_t, std::vector<bool>&)> place = [&](size_t a, std::vector<bool>& seen) {
            for (const auto& c : starts[a]) {
                const size_t cell = c.i * width + c.j;
                if (seen[cell]) continue;
                seen[cell] = true;
                if (owner[cell] < 0 || place((size_t)owner[cell], seen)) { owner[cell] = (long)a; return true; }
            }
            return false;
        };
        for (size_t a = 0; a < n; ++a) {
            std::vector<bool> seen(cells, false);
            if (!place(a, seen)) throw RuntimeWorldError(RuntimeErrorKind::Panic, "Could not assign positions to agents");
        }
        std::vector<Position> out(n);
        for (size_t cell = 0; cell < cells; ++cell)
            if (owner[cell] >= 0) out[(size_t)owner[cell]] = Position{cell / width, cell % width};
        return out;
    }
    static bool assign_positions(size_t i, const std::vector<size_t>& idx,
                                 const std::vector<std::vector<Position>>& starts,
                                 std::vector<Position>& result) {
        if (idx.empty()) return true;
        const auto& possible = starts[idx[i]];
        for (const auto& pos : possible) {
            if (std::find(result.begin(), result.end(), pos) == result.end()) {
                result.push_back(pos);
                if (i + 1 < idx.size()) assign_positions(i + 1, idx, starts, result);
                if (result.size() == idx.size()) return true;
            }
        }
        return false;
    }
};

// ------------------------------------------------------------------------------------------
// Parsing: src/core/parsing/parser_v1.rs + world_config.rs + laser_config.rs
// ------------------------------------------------------------------------------------------
struct WorldConfig {
    size_t width = 0, height = 0;
    std::vector<Position> gems, voids, exits, walls;
    std::vector<std::vector<Position>> random_starts;
    std::vector<std::pair<Position, LaserConfig>> lasers;

    size_t n_agents() const { return random_starts.size(); }

    // world_config.rs:124-148
    void pre_validate() const {
        if (random_starts.empty()) throw ParseError(ParseErrorKind::NoAgents, "NoAgents");
        if (exits.size() < n_agents())
            throw ParseError(ParseErrorKind::NotEnoughExitTiles, "NotEnoughExitTiles", (long)n_agents(),
                             (long)exits.size());
    }
    // world_config.rs:150-170
    void post_validate() const {
        size_t total = 0;
        for (size_t a = 0; a < random_starts.size(); ++a) {
            if (random_starts[a].empty())
                throw ParseError(ParseErrorKind::AgentWithoutStart, "AgentWithoutStart", (long)a);
            total += random_starts[a].size();
        }
        if (total < n_agents())
            throw ParseError(ParseErrorKind::NotEnoughStartTiles, "NotEnoughStartTiles", (long)total,
                             (long)n_agents());
    }

    // world_config.rs:176-199 (make_grid) and :203-250 (laser_setup)
    std::vector<std::vector<Tile>> make_grid(std::vector<Position>& laser_positions) {
        std::vector<std::vector<Tile>> grid(height);
        for (auto& row : grid) {
            row.reserve(width);
            for (size_t j = 0; j < width; ++j) row.push_back(Tile::make(Tile::Floor));
        }
        for (const auto& p : gems) grid[p.i][p.j] = Tile::make(Tile::Gem);
        for (const auto& p : exits) grid[p.i][p.j] = Tile::make(Tile::Exit);
        for (const auto& p : voids) grid[p.i][p.j] = Tile::make(Tile::Void);
        for (const auto& p : walls) grid[p.i][p.j] = Tile::make(Tile::Wall);

        std::vector<std::vector<bool>> is_laser(height, std::vector<bool>(width, false));
        long w = (long)width, h = (long)height;
        for (const auto& [src_pos, cfg] : lasers) {
            // a [[lasers]] position is a plain Position in the reference (toml_laser_config.rs:9-15), not bounds-checked:
            // `grid[pos.i][pos.j] = ...` (world_config.rs:247) panics with an index error
            if (src_pos.i >= height || src_pos.j >= width)
                throw RuntimeWorldError(RuntimeErrorKind::Panic, "index out of bounds: a laser source outside the grid");
            std::vector<Position> beam_positions;
            auto delta = direction_delta(cfg.direction);
            long i = (long)src_pos.i + delta.first, j = (long)src_pos.j + delta.second;
            while (i >= 0 && j >= 0 && i < h && j < w) {
                if (!grid[i][j].is_walkable()) break;
                beam_positions.push_back(Position{(size_t)i, (size_t)j});
                i += delta.first;
                j += delta.second;
            }
            for (const auto& p : beam_positions) is_laser[p.i][p.j] = true;
            auto beam = std::make_shared<LaserBeam>(beam_positions.size(), cfg.agent_id, cfg.direction,
                                                    cfg.laser_id);  // laser_config.rs:38-46
            bool is_blocked = false;
            for (size_t k = 0; k < beam_positions.size(); ++k) {
                const Position pos = beam_positions[k];
                if (cfg.agent_id < random_starts.size()) {
                    const auto& st = random_starts[cfg.agent_id];
                    if (st.size() == 1 && std::find(st.begin(), st.end(), pos) != st.end()) is_blocked = true;
                }
                Tile laser = Tile::make(Tile::Laser);
                laser.wrapped = std::make_unique<Tile>(std::move(grid[pos.i][pos.j]));
                laser.beam = beam;
                laser.offset = k;
                if (!is_blocked) {
                    for (size_t a = 0; a < random_starts.size(); ++a) {
                        if (a == cfg.agent_id) continue;
                        auto& st = random_starts[a];
                        st.erase(std::remove(st.begin(), st.end(), pos), st.end());
                    }
                }
                grid[pos.i][pos.j] = std::move(laser);
            }
            Tile source = Tile::make(Tile::LaserSource);
            source.beam = beam;
            grid[src_pos.i][src_pos.j] = std::move(source);
        }
        laser_positions.clear();
        for (size_t i2 = 0; i2 < height; ++i2)
            for (size_t j2 = 0; j2 < width; ++j2)
                if (is_laser[i2][j2]) laser_positions.push_back(Position{i2, j2});
        return grid;
    }

    // world_config.rs:107-122 ; world.rs:48-84
    World into_world() {
        pre_validate();
        std::vector<Position> lasers_positions;
        auto grid = make_grid(lasers_positions);
        post_validate();
        World w;
        w.width = width;
        w.height = height;
        w.grid = std::move(grid);
        for (size_t a = 0; a < random_starts.size(); ++a) w.agents.emplace_back(a);
        w.gems_positions = gems;
        w.random_start_positions = random_starts;
        w.void_positions = voids;
        w.exits = exits;
        w.wall_positions = walls;
        for (const auto& l : lasers) w.laser_source_positions.push_back(l.first);
        w.lasers_positions = std::move(lasers_positions);
        w.reset();
        return w;
    }
};

inline std::string to_upper(std::string s) {
    for (auto& c : s) c = (char)std::toupper((unsigned char)c);
    return s;
}

// parser_v1.rs:132-175 ; laser_config.rs:19-35
inline WorldConfig parse_v1(const std::string& world_str) {
    WorldConfig cfg;
    bool have_width = false;
    std::istringstream lines(world_str);
    std::string raw;
    while (std::getline(lines, raw)) {
        // str::trim / split_whitespace
        std::istringstream toks(raw);
        std::vector<std::string> tokens;
        std::string tok;
        while (toks >> tok) tokens.push_back(tok);
        if (tokens.empty()) continue;
        size_t n_cols = 0;
        for (size_t col = 0; col < tokens.size(); ++col) {
            n_cols++;
            const std::string& token = tokens[col];
            Position pos{cfg.height, col};
            char c = (char)std::toupper((unsigned char)token[0]);
            switch (c) {
                case '.': break;
                case 'G': cfg.gems.push_back(pos); break;
                case '@': cfg.walls.push_back(pos); break;
                case 'X': cfg.exits.push_back(pos); break;
                case 'V': cfg.voids.push_back(pos); break;
                case 'S': {
                    std::string id = token.substr(1);
                    if (id.empty() || id.find_first_not_of("0123456789") != std::string::npos) {
                        // Rust usize::parse also accepts a leading '+'
                        if (!(id.size() > 1 && id[0] == '+' &&
                              id.find_first_not_of("0123456789", 1) == std::string::npos))
                            throw ParseError(ParseErrorKind::InvalidAgentId, "InvalidAgentId: " + id);
                    }
                    size_t agent_id = (size_t)std::stoul(id);
                    while (cfg.random_starts.size() <= agent_id) cfg.random_starts.emplace_back();
                    if (!cfg.random_starts[agent_id].empty())
                        throw ParseError(ParseErrorKind::DuplicateStartTile, "DuplicateStartTile", (long)agent_id);
                    cfg.random_starts[agent_id].push_back(pos);
                    break;
                }
                case 'L': {
                    char dch = (char)std::tolower((unsigned char)token.back());
                    Direction dir;
                    switch (dch) {
                        case 'n': dir = Direction::North; break;
                        case 'e': dir = Direction::East; break;
                        case 's': dir = Direction::South; break;
                        case 'w': dir = Direction::West; break;
                        default:  // Direction::try_from(...).unwrap() panics (laser_config.rs:20)
                            throw RuntimeWorldError(RuntimeErrorKind::Panic, "InvalidDirection: " + token);
                    }
                    std::string id = token.size() >= 2 ? token.substr(1, token.size() - 2) : std::string();
                    if (id.empty() || id.find_first_not_of("0123456789") != std::string::npos) {
                        if (!(id.size() > 1 && id[0] == '+' &&
                              id.find_first_not_of("0123456789", 1) == std::string::npos))
                            throw ParseError(ParseErrorKind::InvalidAgentId, "InvalidAgentId: " + id);
                    }
                    LaserConfig lc{dir, (size_t)std::stoul(id), cfg.lasers.size()};
                    cfg.lasers.emplace_back(pos, lc);
                    cfg.walls.push_back(pos);  // parser_v1.rs:22-25
                    break;
                }
                default:
                    throw ParseError(ParseErrorKind::InvalidTile, "InvalidTile: " + token, (long)pos.i,
                                     (long)pos.j);
            }
        }
        // parser_v1.rs:62-77
        if (have_width) {
            if (cfg.width != n_cols)
                throw ParseError(ParseErrorKind::InconsistentDimensions, "InconsistentDimensions",
                                 (long)cfg.width, (long)n_cols, (long)cfg.height);
        } else {
            cfg.width = n_cols;
            have_width = true;
        }
        cfg.height += 1;
    }
    if (cfg.height == 0) throw ParseError(ParseErrorKind::EmptyWorld, "EmptyWorld");
    return cfg;
}

// A WorldConfig written out field by field.  TOML v2 maps are deserialised by oracle/toml_config.py (tomllib + a
// restatement of src/core/parsing/toml/*.rs) and handed over in this form:
//   %LLE-CONFIG / size H W / gems n i j ... / voids n ... / exits n ... / walls n ... / agents A / A x "starts n i j ..." /
//   lasers n / n x "laser i j agent_id direction(0 N,1 E,2 S,3 W) laser_id"
inline WorldConfig parse_config_text(const std::string& text) {
    // tokens are read as strings and converted with std::stoul (iostream number parsing depends on the process locale)
    std::istringstream in(text);
    std::string word;
    WorldConfig cfg;
    auto expect = [&](const char* w) {
        if (!(in >> word) || word != w) throw std::invalid_argument(std::string("config text: expected ") + w);
    };
    auto number = [&]() -> size_t {
        if (!(in >> word)) throw std::invalid_argument("config text: truncated");
        return (size_t)std::stoul(word);
    };
    auto positions = [&](const char* name) {
        expect(name);
        std::vector<Position> out(number());
        for (auto& p : out) {
            p.i = number();
            p.j = number();
        }
        return out;
    };
    expect("%LLE-CONFIG");
    expect("size");
    cfg.height = number();
    cfg.width = number();
    cfg.gems = positions("gems");
    cfg.voids = positions("voids");
    cfg.exits = positions("exits");
    cfg.walls = positions("walls");
    expect("agents");
    const size_t n_agents = number();
    for (size_t a = 0; a < n_agents; ++a) cfg.random_starts.push_back(positions("starts"));
    expect("lasers");
    const size_t n_lasers = number();
    for (size_t k = 0; k < n_lasers; ++k) {
        expect("laser");
        Position pos;
        pos.i = number();
        pos.j = number();
        const size_t agent = number(), dir = number(), laser_id = number();
        cfg.lasers.emplace_back(pos, LaserConfig{(Direction)dir, agent, laser_id});
    }
    return cfg;
}

inline World parse_world(const std::string& text) {
    // parsing/mod.rs:14-21 tries TOML first, then v1.  TOML arrives here already deserialised (see parse_config_text).
    World w = text.rfind("%LLE-CONFIG", 0) == 0 ? parse_config_text(text).into_world() : parse_v1(text).into_world();
    w.source_text = text;
    return w;
}

// ------------------------------------------------------------------------------------------
// Python layer restated: observations.py, env/env.py, env/reward_strategy.py
// ------------------------------------------------------------------------------------------

// python/lle/observations.py:196-266 (LayeredPadded; padding_size = 0 == Layered)
struct Layered {
    size_t n_agents, A0, LASER_0, WALL, VOID, GEM, EXIT, C, H, W;
    std::vector<float> static_obs;

    explicit Layered(const World& w, size_t padding_size = 0) {
        n_agents = w.n_agents() + padding_size;  // observations.py:203
        A0 = 0;
        LASER_0 = A0 + n_agents;
        WALL = LASER_0 + n_agents;
        VOID = WALL + 1;
        GEM = VOID + 1;
        EXIT = GEM + 1;
        C = EXIT + 1;
        H = w.height;
        W = w.width;
        setup(w);
    }
    size_t size() const { return C * H * W; }
    float& px(std::vector<float>& obs, size_t c, size_t i, size_t j) const {
        if (c >= C) throw std::out_of_range("IndexError: layered channel out of range");  // numpy IndexError
        return obs[(c * H + i) * W + j];
    }
    // observations.py:216-237
    void setup(const World& w) {
        static_obs.assign(size(), 0.0f);
        for (const auto& p : w.wall_positions) px(static_obs, WALL, p.i, p.j) = 1.0f;
        for (const auto& p : w.void_positions) px(static_obs, VOID, p.i, p.j) = 1.0f;
        for (const auto& p : w.exits) px(static_obs, EXIT, p.i, p.j) = 1.0f;
        for (size_t s = 0; s < w.laser_source_positions.size(); ++s) {
            const Position& p = w.laser_source_positions[s];
            px(static_obs, LASER_0 + w.source_beam(s)->agent_id, p.i, p.j) = -1.0f;
        }
    }
    // observations.py:254-266 — one (C,H,W) copy; np.tile over agents is a pure repeat
    void observe(const World& w, float* out) const {
        std::vector<float> obs = static_obs;
        for (const auto& l : w.lasers())
            if (l.is_on) px(obs, LASER_0 + l.agent_id, l.pos.i, l.pos.j) = 1.0f;
        auto gems = w.gems();
        for (size_t g = 0; g < gems.size(); ++g)
            if (!gems[g]->collected) px(obs, GEM, w.gems_positions[g].i, w.gems_positions[g].j) = 1.0f;
        for (size_t a = 0; a < w.agents_positions.size(); ++a)
            px(obs, A0 + a, w.agents_positions[a].i, w.agents_positions[a].j) = 1.0f;
        std::memcpy(out, obs.data(), obs.size() * sizeof(float));
    }
};

// python/lle/observations.py:296-369 (PartialGenerator): one (2A+3, size, size) window per agent, centred on it.
// Channel order: agents, WALL, lasers, GEM, EXIT (no VOID layer).  Cells outside the map stay 0.
struct Partial {
    size_t A, size, center, WALL, LASER_0, GEM, EXIT, C;
    Partial(const World& w, size_t square_size) {
        if (square_size % 2 != 1) throw std::invalid_argument("Can only use odd numbers for the square size");  // :299
        A = w.n_agents();
        size = square_size;
        center = size / 2;
        WALL = A;
        LASER_0 = WALL + 1;
        GEM = LASER_0 + A;
        EXIT = GEM + 1;
        C = 2 * A + 3;
    }
    size_t floats() const { return A * C * size * size; }
    // obs[a, layer] with numpy's bounds check on the layer index
    float* layer(float* obs, size_t a, size_t c) const {
        if (c >= C) throw std::out_of_range("IndexError: partial channel out of range");
        return obs + (a * C + c) * size * size;
    }
    // :322-329
    void encode(float* layer_ptr, const Position& origin, const Position& p, float fill = 1.0f) const {
        const long i = (long)p.i - (long)origin.i + (long)center, j = (long)p.j - (long)origin.j + (long)center;
        if (0 <= i && i < (long)size && 0 <= j && j < (long)size) layer_ptr[i * (long)size + j] = fill;
    }
    // :331-350
    void observe(const World& w, float* obs) const {
        std::fill(obs, obs + floats(), 0.0f);
        const auto lasers = w.lasers();
        const auto gems = w.gems();
        for (size_t a = 0; a < A; ++a) {
            const Position& me = w.agents_positions[a];
            for (size_t a2 = 0; a2 < A; ++a2) encode(layer(obs, a, a2), me, w.agents_positions[a2]);
            for (size_t g = 0; g < gems.size(); ++g)
                if (!gems[g]->collected) encode(layer(obs, a, GEM), me, w.gems_positions[g]);
            for (const auto& p : w.exits) encode(layer(obs, a, EXIT), me, p);
            for (const auto& p : w.wall_positions) encode(layer(obs, a, WALL), me, p);
            for (const auto& l : lasers)  // _get_lasers_positions (:352-360): only colours with a lit laser are touched
                if (l.is_on) encode(layer(obs, a, LASER_0 + l.agent_id), me, l.pos);
            for (size_t s = 0; s < w.laser_source_positions.size(); ++s)
                encode(layer(obs, a, LASER_0 + w.source_beam(s)->agent_id), me, w.laser_source_positions[s], -1.0f);
        }
    }
};

// python/lle/observations.py:372-395 (AgentZeroPerspective): the layered observation, where agent n's copy has the
// agent layers 0 <-> n and the laser layers 0 <-> n swapped.  Output: (A, C, H, W), one distinct copy per agent.
struct Perspective {
    Layered base;
    explicit Perspective(const World& w) : base(w) {}
    size_t floats() const { return base.n_agents * base.size(); }
    void observe(const World& w, float* out) const {
        const size_t n = base.size(), HW = base.H * base.W;
        std::vector<float> one(n);
        base.observe(w, one.data());
        for (size_t a = 0; a < base.n_agents; ++a) std::memcpy(out + a * n, one.data(), n * sizeof(float));  // np.tile
        for (size_t a = 1; a < base.n_agents; ++a) {
            float* obs = out + a * n;
            std::swap_ranges(obs + base.A0 * HW, obs + (base.A0 + 1) * HW, obs + (base.A0 + a) * HW);
            std::swap_ranges(obs + base.LASER_0 * HW, obs + (base.LASER_0 + 1) * HW, obs + (base.LASER_0 + a) * HW);
        }
    }
};

// src/bindings/world/pyworld_state.rs:79-101
inline void state_as_array(const WorldState& s, float* out) {
    size_t k = 0;
    for (const auto& p : s.agents_positions) {
        out[k++] = (float)p.i;
        out[k++] = (float)p.j;
    }
    for (bool c : s.gems_collected) out[k++] = c ? 1.0f : 0.0f;
    for (bool a : s.agents_alive) out[k++] = a ? 1.0f : 0.0f;
}

// python/lle/observations.py:141-158 (StateGenerator.observe, one row; np.tile over agents is a pure repeat):
// float32 state, positions divided by [height, width] (int64 -> the division happens in float64) and stored back to float32
inline void state_observation(const World& w, bool normalize, float* out) {
    const WorldState s = w.get_state();
    state_as_array(s, out);
    if (normalize)
        for (size_t a = 0; a < s.agents_positions.size(); ++a) {
            out[2 * a] = (float)((double)out[2 * a] / (double)w.height);
            out[2 * a + 1] = (float)((double)out[2 * a + 1] / (double)w.width);
        }
}

enum class ObsKind : int { Layered = 0, Partial = 1, Perspective = 2, State = 3 };

// ObservationType.get_observation_generator (observations.py:66-97) for the kinds on the accelerated path.
// `floats()` is the size of one env's block in the device layout: layered = ONE (C,H,W) copy (np.tile's repeats are a
// stride-0 view), state = one row, partial / perspective = all agents' (distinct) copies.
struct ObsGen {
    ObsKind kind = ObsKind::Layered;
    int param = 0;  // layered: padding_size; partial: square size; state: 1 = normalised
    std::unique_ptr<Layered> layered;
    std::unique_ptr<Partial> partial;
    std::unique_ptr<Perspective> perspective;
    size_t state_len = 0;
    ObsGen(const World& w, ObsKind k, int prm) : kind(k), param(prm) {
        switch (k) {
            case ObsKind::Layered: layered = std::make_unique<Layered>(w, (size_t)prm); break;
            case ObsKind::Partial: partial = std::make_unique<Partial>(w, (size_t)prm); break;
            case ObsKind::Perspective: perspective = std::make_unique<Perspective>(w); break;
            case ObsKind::State: state_len = 3 * w.n_agents() + w.n_gems(); break;
        }
    }
    size_t floats() const {
        switch (kind) {
            case ObsKind::Layered: return layered->size();
            case ObsKind::Partial: return partial->floats();
            case ObsKind::Perspective: return perspective->floats();
            default: return state_len;
        }
    }
    // out[0] = number of dims of one env's block, out[1..4] = dims, out[5] = agent copies that np.tile adds in front
    // (0: the block already carries the agent dimension)
    void block_shape(long* out) const {
        for (int k = 0; k < 6; ++k) out[k] = 0;
        switch (kind) {
            case ObsKind::Layered:
                out[0] = 3; out[1] = (long)layered->C; out[2] = (long)layered->H; out[3] = (long)layered->W;
                out[5] = (long)layered->n_agents;  // np.tile(obs, (self.n_agents, ...)) with the padded agent count (:266)
                break;
            case ObsKind::Partial:
                out[0] = 4; out[1] = (long)partial->A; out[2] = (long)partial->C; out[3] = out[4] = (long)partial->size;
                break;
            case ObsKind::Perspective:
                out[0] = 4; out[1] = (long)perspective->base.n_agents; out[2] = (long)perspective->base.C;
                out[3] = (long)perspective->base.H; out[4] = (long)perspective->base.W;
                break;
            default:
                out[0] = 1; out[1] = (long)state_len; out[5] = (long)((state_len) ? 0 : 0);
                break;
        }
    }
    void setup(const World& w) {  // ObservationGenerator.reset (observations.py:239-240; called by LLE.reset, env.py:203)
        if (layered) layered->setup(w);
        if (perspective) perspective->base.setup(w);
    }
    void observe(const World& w, float* out) const {
        switch (kind) {
            case ObsKind::Layered: layered->observe(w, out); break;
            case ObsKind::Partial: partial->observe(w, out); break;
            case ObsKind::Perspective: perspective->observe(w, out); break;
            default: state_observation(w, param != 0, out); break;
        }
    }
};

constexpr float REWARD_GEM = 1.0f, REWARD_EXIT = 1.0f, REWARD_DONE = 1.0f, REWARD_DEATH = -1.0f;

// python/lle/env/extras_generators.py:75-101 (LaserSubgoal) and the position sets of
// reward_strategy.py:161-167 (PotentialShapedLLE._compute_positions_to_reward): for each chosen source, the
// positions of its laser tiles *as listed by World::lasers()* (env/utils.py:6-11), and an (agent, source) matrix of
// sticky "has stood there" flags.
struct SubgoalTracker {
    std::vector<std::vector<Position>> pos_to_reward;  // per chosen source
    std::vector<std::vector<bool>> reached;            // [agent][source]
    void init(const World& w, const std::vector<size_t>& sources) {
        pos_to_reward.clear();
        auto lasers = w.lasers();
        for (size_t src : sources) {
            std::vector<Position> cells;
            const size_t laser_id = w.source_beam(src)->laser_id;
            for (const auto& l : lasers)
                if (l.laser_id == laser_id) cells.push_back(l.pos);
            pos_to_reward.push_back(cells);
        }
        reached.assign(w.n_agents(), std::vector<bool>(sources.size(), false));
    }
    void clear() {
        for (auto& row : reached) std::fill(row.begin(), row.end(), false);
    }
    void mark(const World& w) {  // extras_generators.py:93-97 / reward_strategy.py:169-173
        for (size_t a = 0; a < w.agents_positions.size(); ++a)
            for (size_t j = 0; j < pos_to_reward.size(); ++j)
                if (std::find(pos_to_reward[j].begin(), pos_to_reward[j].end(), w.agents_positions[a]) != pos_to_reward[j].end())
                    reached[a][j] = true;
    }
    size_t size() const { return reached.size() * pos_to_reward.size(); }
    size_t count() const {
        size_t n = 0;
        for (const auto& row : reached)
            for (bool b : row) n += b;
        return n;
    }
};

// python/lle/env/env.py (LLE) + env/reward_strategy.py (SingleObjective / MultiObjective / PotentialShapedLLE)
struct Env {
    World world;
    ObsGen obs;
    std::unique_ptr<ObsGen> state_gen;  // LLE(state_type=...): `_state_generator` (env.py:86); null = the default StateGenerator
    bool multi_objective;
    bool walkable_lasers;
    size_t n_arrived = 0, n_deads = 0;
    bool done = false;
    // extras: LaserSubgoal (Builder.add_extras("laser_subgoal"), builder.py:124-150)
    bool has_extras = false;
    SubgoalTracker extras;
    // reward shaping: PotentialShapedLLE (Builder.pbrs, builder.py:77-110)
    bool has_pbrs = false;
    SubgoalTracker pbrs;
    double gamma = 0.99, reward_value = 0.5, previous_potential = 0.0;
    // LLE(randomize_lasers=True) (env.py:83, :198-200).  Python's global `random` cannot be reproduced; the colours follow
    // the device library's documented Philox contract (include/lle_b200.h, lle_vec_options.randomize_lasers), with the
    // counter words the harness keeps in world.rng_*.
    bool randomize_lasers = false;

    Env(World&& w, bool multi_obj = false, bool walkable = true)
        : world(std::move(w)), obs(world, ObsKind::Layered, 0), multi_objective(multi_obj), walkable_lasers(walkable) {}

    void set_obs(ObsKind kind, int param) { obs = ObsGen(world, kind, param); }  // Builder.obs_type (builder.py:42-49)
    void set_state_type(ObsKind kind, int param) { state_gen = std::make_unique<ObsGen>(world, kind, param); }  // Builder.state_type (:51-58)

    size_t reward_dim() const { return (multi_objective ? 4 : 1) + (multi_objective && has_pbrs ? 1 : 0); }

    void enable_extras(const std::vector<size_t>& sources) {  // LaserSubgoal.__init__ (extras_generators.py:80-91)
        has_extras = true;
        extras.init(world, sources);
    }
    void enable_pbrs(double g, double value, const std::vector<size_t>& sources) {  // reward_strategy.py:128-153
        has_pbrs = true;
        gamma = g;
        reward_value = value;
        pbrs.init(world, sources);
        previous_potential = compute_potential();
    }
    // reward_strategy.py:169-174
    double compute_potential() {
        pbrs.mark(world);
        return (double)(pbrs.size() - pbrs.count()) * reward_value;
    }
    // LaserSubgoal.compute (extras_generators.py:93-98) -> float32 [A, J]
    void compute_extras(float* out) {
        if (!has_extras) return;
        extras.mark(world);
        size_t k = 0;
        for (const auto& row : extras.reached)
            for (bool b : row) out[k++] = b ? 1.0f : 0.0f;
    }

    // reward_strategy.py:58-75 (SingleObjective) and :90-109 (MultiObjective)
    void compute_reward(const std::vector<WorldEvent>& events, float* reward) {
        if (!multi_objective) {
            double r = 0.0, death_reward = 0.0;
            for (const auto& e : events) {
                switch (e.type) {
                    case EventType::AgentDied: r += REWARD_DEATH; n_deads += 1; break;
                    case EventType::GemCollected: r += REWARD_GEM; break;
                    case EventType::AgentExit: r += REWARD_EXIT; n_arrived += 1; break;
                }
            }
            if (death_reward != 0) r = death_reward;  // dead code in the reference (:71-72)
            else if (n_arrived == world.n_agents()) r += REWARD_DONE;
            reward[0] = (float)r;
        } else {
            float r[4] = {0, 0, 0, 0};  // gem, exit, death, done
            for (const auto& e : events) {
                switch (e.type) {
                    case EventType::AgentDied: r[2] += REWARD_DEATH; n_deads += 1; break;
                    case EventType::GemCollected: r[0] += REWARD_GEM; break;
                    case EventType::AgentExit: r[1] += REWARD_EXIT; n_arrived += 1; break;
                }
            }
            if (r[2] != 0) {
                float d = r[2];
                r[0] = r[1] = r[3] = 0;
                r[2] = d;
            } else if (n_arrived == world.n_agents()) {
                r[3] += REWARD_DONE;
            }
            std::memcpy(reward, r, sizeof(r));
        }
        if (has_pbrs) {  // PotentialShapedLLE.compute_reward (reward_strategy.py:145-158)
            const double current = compute_potential();
            const double shaped = gamma * previous_potential - current;  // python floats: no fused multiply-add
            if (!multi_objective) reward[0] = reward[0] + (float)shaped;  // np.float32 += python float (NEP 50: float32 arithmetic)
            else reward[4] = (float)shaped;  // np.concat gives float64 there; the oracle's buffer is float32 (value rounded once)
            previous_potential = current;
        }
    }
    bool compute_done() const { return n_arrived == world.n_agents() || n_deads > 0; }  // env.py:253-254

    // env.py:191-203 (randomize_lasers is out of scope)
    void reset(bool lle_level = true) {
        world.reset();
        reset_strategy();
        if (has_extras) extras.clear();  // extras_generator.reset() (env.py:196)
        done = false;
        if (randomize_lasers && lle_level) {  // env.py:198-200, after world.reset(): source.set_colour(random colour)
            const uint32_t key[2] = {(uint32_t)world.rng_seed, (uint32_t)(world.rng_seed >> 32)};
            const uint32_t n_agents = (uint32_t)world.n_agents();
            for (size_t b = 0; b < world.laser_source_positions.size(); ++b) {
                const uint32_t ctr[4] = {world.rng_env, world.rng_t, 0x20000000u | (uint32_t)(b >> 2), world.rng_epoch};
                uint32_t out[4];
                Philox::run(ctr, key, out);
                world.source_beam(b)->agent_id = (size_t)(((uint64_t)out[b & 3] * n_agents) >> 32);
            }
        }
        obs.setup(world);
        if (state_gen) state_gen->setup(world);  // self._state_generator.reset() (env.py:202)
    }
    // RewardStrategy.reset (:40-42) / PotentialShapedLLE.reset (:176-180)
    void reset_strategy() {
        n_arrived = 0;
        n_deads = 0;
        if (has_pbrs) {
            pbrs.clear();
            previous_potential = compute_potential();
        }
    }
    // env.py:165-189 ; raises ValueError on a done env (:166-167)
    std::vector<WorldEvent> step(const std::vector<Action>& actions, float* reward) {
        if (done) throw std::invalid_argument("Cannot step in a done environment");
        auto events = world.step(actions);
        compute_reward(events, reward);
        done = compute_done();
        return events;
    }
    // env.py:208-216
    void set_state(const WorldState& s) {
        reset_strategy();  // with the positions the world has *before* the new state is forced (env.py:213)
        auto events = world.set_state(s);
        float scratch[5];
        compute_reward(events, scratch);
        done = compute_done();
    }
    // env.py:146-163 -> bool[A,5] indexed by Action.value
    void available_actions(uint8_t* out) const {
        size_t A = world.n_agents();
        std::memset(out, 0, A * 5);
        if (walkable_lasers) {
            for (size_t a = 0; a < A; ++a)
                for (Action act : world.available_actions[a]) out[a * 5 + (size_t)act] = 1;
            return;
        }
        auto lasers = world.lasers();
        for (size_t a = 0; a < A; ++a) {
            for (Action act : world.available_actions[a]) {
                auto d = action_delta(act);
                long ni = (long)world.agents_positions[a].i + d.first;
                long nj = (long)world.agents_positions[a].j + d.second;
                bool blocked = false;
                for (const auto& l : lasers)
                    if ((long)l.pos.i == ni && (long)l.pos.j == nj && l.agent_id != a && l.is_on) blocked = true;
                if (!blocked) out[a * 5 + (size_t)act] = 1;
            }
        }
    }
    void observe(float* out) const { obs.observe(world, out); }
    void state(float* out) const { state_as_array(world.get_state(), out); }
    // LLE.get_state (env.py:205-206): `_state_generator.get_state()` = observe()[0] (observations.py:118-119); out holds one env's
    // whole block of that generator (the caller takes the first agent's part)
    void state_observation(float* out) const { state_gen->observe(world, out); }
};

}  // namespace lle_oracle
