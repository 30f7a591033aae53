// TOML v2 maps -> RawConfig.  Restates src/core/parsing/toml/{toml_config,agent_config,position_config,toml_laser_config}.rs
// (serde models + TryInto<WorldConfig>) on top of toml_lite.hpp.
#include <algorithm>
#include <set>

#include "map_compiler.hpp"
#include "toml_lite.hpp"

namespace lle {
namespace {

using toml::Value;

struct NotV2 {};  // ParseError::NotV2: the text does not deserialise into the v2 model (toml_config.rs:117-131)

int64_t as_usize(const Value& v) {
    if (v.kind != Value::Int || v.i < 0) throw NotV2{};
    return v.i;
}

// PositionsConfig (position_config.rs:5-26) is `#[serde(untagged)]`: the first variant that deserialises wins, unknown keys
// are ignored.  IJ needs both i and j, Row needs row, Column needs col, Rect has defaults for everything (so `{}` is the
// whole map and `{ i = 3 }` is, too).
std::vector<Cell> positions_of(const Value& v, int width, int height) {
    if (v.kind != Value::Table) throw NotV2{};
    auto field = [&](const char* k) -> const Value* { return v.get(k); };
    auto usable = [&](const Value* f) { return f && f->kind == Value::Int && f->i >= 0; };
    auto oob = [](int64_t i, int64_t j) {
        return MapError(LLE_PARSE_POSITION_OUT_OF_BOUNDS, "PositionOutOfBounds { i: " + std::to_string(i) + ", j: " + std::to_string(j) + " }");
    };
    std::vector<Cell> out;
    if (usable(field("i")) && usable(field("j"))) {  // IJ (position_config.rs:31-36)
        const int64_t i = field("i")->i, j = field("j")->i;
        if (i >= height || j >= width) throw oob(i, j);
        out.push_back(Cell{(int)i, (int)j});
        return out;
    }
    if (usable(field("row"))) {  // Row (:65-70)
        const int64_t row = field("row")->i;
        if (row >= height) throw oob(row, 0);
        for (int j = 0; j < width; ++j) out.push_back(Cell{(int)row, j});
        return out;
    }
    if (usable(field("col"))) {  // Column (:71-76)
        const int64_t col = field("col")->i;
        if (col >= width) throw oob(0, col);
        for (int i = 0; i < height; ++i) out.push_back(Cell{i, (int)col});
        return out;
    }
    // Rect (:37-64); a present field of the wrong type makes the variant (the last one) fail
    int64_t i_min = 0, j_min = 0, i_max = height - 1, j_max = width - 1;
    if (const Value* f = field("i_min")) i_min = as_usize(*f);
    if (const Value* f = field("j_min")) j_min = as_usize(*f);
    if (const Value* f = field("i_max")) i_max = as_usize(*f);
    if (const Value* f = field("j_max")) j_max = as_usize(*f);
    if (i_min >= height || j_min >= width) throw oob(i_min, j_min);
    for (int64_t i = i_min; i <= i_max; ++i)
        for (int64_t j = j_min; j <= j_max; ++j) {
            if (i >= height || j >= width) throw oob(i, j);
            out.push_back(Cell{(int)i, (int)j});
        }
    return out;
}

std::vector<Cell> compute_positions(const std::vector<Value>& configs, int width, int height) {  // toml_config.rs:104-114
    std::vector<Cell> res;
    for (const auto& c : configs) {
        auto p = positions_of(c, width, height);
        res.insert(res.end(), p.begin(), p.end());
    }
    return res;
}

const std::vector<Value>& array_or_empty(const Value* v) {
    static const std::vector<Value> empty;
    if (!v) return empty;
    if (v->kind != Value::Array) throw NotV2{};
    return v->arr;
}

int direction_of(const Value& v) {  // Direction with its serde aliases (direction.rs:8-18)
    if (v.kind != Value::Str) throw NotV2{};
    const std::string& s = v.s;
    if (s == "North" || s == "N" || s == "north" || s == "n") return 0;
    if (s == "East" || s == "E" || s == "east" || s == "e") return 1;
    if (s == "South" || s == "S" || s == "south" || s == "s") return 2;
    if (s == "West" || s == "W" || s == "west" || s == "w") return 3;
    throw NotV2{};
}

}  // namespace

bool parse_toml_config(const std::string& text, RawConfig& out) {
    Value doc;
    try {
        doc = toml::Parser(text).parse();
    } catch (const toml::SyntaxError&) {
        return false;
    }
    try {
        // `#[serde(deny_unknown_fields)]` on TomlConfig (toml_config.rs:12-33) and AgentConfig (agent_config.rs:10-15):
        // an unknown key is a hard error (UnknownTomlKey), every other mismatch means "not v2"
        static const char* const top[] = {"width", "height", "n_agents", "world_string", "agents", "exits", "gems", "walls", "voids", "lasers", "starts"};
        for (const auto& kv : doc.tbl)
            if (std::find_if(std::begin(top), std::end(top), [&](const char* k) { return kv.first == k; }) == std::end(top))
                throw MapError(LLE_PARSE_UNKNOWN_TOML_KEY, "UnknownTomlKey { key: \"" + kv.first + "\" }");
        struct AgentCfg { std::vector<Value> starts; };
        std::vector<AgentCfg> agents;
        for (const auto& a : array_or_empty(doc.get("agents"))) {
            if (a.kind != Value::Table) throw NotV2{};
            AgentCfg cfg;
            for (const auto& kv : a.tbl) {
                if (kv.first != "starts" && kv.first != "start_positions")  // `alias = "start_positions"`
                    throw MapError(LLE_PARSE_UNKNOWN_TOML_KEY, "UnknownTomlKey { key: \"" + kv.first + "\" }");
                const auto& lst = array_or_empty(&kv.second);
                cfg.starts.insert(cfg.starts.end(), lst.begin(), lst.end());
            }
            if (a.get("starts") && a.get("start_positions")) throw NotV2{};  // serde: duplicate field
            agents.push_back(cfg);
        }
        int64_t width = -1, height = -1, n_agents = -1;
        if (const Value* v = doc.get("width")) width = as_usize(*v);
        if (const Value* v = doc.get("height")) height = as_usize(*v);
        if (const Value* v = doc.get("n_agents")) n_agents = as_usize(*v);
        std::vector<Value> exits = array_or_empty(doc.get("exits")), gems = array_or_empty(doc.get("gems")),
                           walls = array_or_empty(doc.get("walls")), voids = array_or_empty(doc.get("voids")),
                           starts = array_or_empty(doc.get("starts"));
        for (const auto* lst : {&exits, &gems, &walls, &voids, &starts})
            for (const auto& v : *lst)
                if (v.kind != Value::Table) throw NotV2{};
        std::vector<RawSource> sources;
        for (const auto& l : array_or_empty(doc.get("lasers"))) {  // TomlLaserConfig (toml_laser_config.rs:9-15): all four fields
            if (l.kind != Value::Table) throw NotV2{};
            const Value *d = l.get("direction"), *ag = l.get("agent"), *pos = l.get("position"), *id = l.get("laser_id");
            if (!d || !ag || !pos || !id || pos->kind != Value::Table || !pos->get("i") || !pos->get("j")) throw NotV2{};
            sources.push_back(RawSource{Cell{(int)as_usize(*pos->get("i")), (int)as_usize(*pos->get("j"))}, (int)as_usize(*ag),
                                        direction_of(*d), (int)as_usize(*id)});
        }
        const Value* ws = doc.get("world_string");
        if (ws && ws->kind != Value::Str) throw NotV2{};

        // ---- TryInto<WorldConfig> (toml_config.rs:133-180)
        if (n_agents >= 0)
            while ((int64_t)agents.size() < n_agents) agents.emplace_back();
        std::vector<Cell> ws_exits, ws_walls, ws_gems;
        std::vector<std::vector<Cell>> ws_starts;
        if (ws) {  // complete_with_world_string (toml_config.rs:36-95); note that the voids of the string are not carried over
            const RawConfig cfg = parse_v1_config(ws->s);
            if (width >= 0 && width != cfg.W)
                throw MapError(LLE_PARSE_INCONSISTENT_WORLD_STRING_WIDTH, "InconsistentWorldStringWidth { toml_width: " + std::to_string(width) +
                                                                               ", world_str_width: " + std::to_string(cfg.W) + " }");
            width = cfg.W;
            if (height >= 0 && height != cfg.H)
                throw MapError(LLE_PARSE_INCONSISTENT_WORLD_STRING_HEIGHT, "InconsistentWorldStringHeight { toml_height: " + std::to_string(height) +
                                                                                ", world_str_height: " + std::to_string(cfg.H) + " }");
            height = cfg.H;
            ws_starts = cfg.starts;
            while (agents.size() < ws_starts.size()) agents.emplace_back();
            if (n_agents >= 0 && n_agents < (int64_t)agents.size())
                throw MapError(LLE_PARSE_INCONSISTENT_NUMBER_OF_AGENTS, "InconsistentNumberOfAgents { toml_n_agents_field: " + std::to_string(n_agents) +
                                                                            ", actual_n_agents: " + std::to_string(agents.size()) + " }");
            ws_exits = cfg.exits;
            ws_walls = cfg.walls;
            ws_gems = cfg.gems;
            for (const auto& s : cfg.sources) sources.push_back(s);  // appended after the [[lasers]] of the document
        }
        if (width < 0 || height < 0) throw MapError(LLE_PARSE_EMPTY_WORLD, "EmptyWorld");
        if (width > 256 || height > 256) throw MapError(LLE_LIMIT_EXCEEDED, "map exceeds the device format (H, W <= 256)");
        const int W = (int)width, H = (int)height;
        auto with = [&](const std::vector<Value>& cfgs, const std::vector<Cell>& extra) {
            std::vector<Cell> res = compute_positions(cfgs, W, H);
            res.insert(res.end(), extra.begin(), extra.end());
            return res;
        };
        const std::vector<Cell> global_starts = compute_positions(starts, W, H);
        out = RawConfig();
        out.W = W; out.H = H;
        out.walls = with(walls, ws_walls);
        out.exits = with(exits, ws_exits);
        out.gems = with(gems, ws_gems);
        out.voids = compute_positions(voids, W, H);
        for (size_t a = 0; a < agents.size(); ++a) {  // AgentConfig::compute_start_positions (agent_config.rs:18-60)
            std::set<std::pair<int, int>> res;  // HashSet, then sorted by i * width + j: row-major order
            for (const auto& c : global_starts) res.insert({c.i, c.j});
            for (const auto& c : compute_positions(agents[a].starts, W, H)) res.insert({c.i, c.j});
            if (a < ws_starts.size())
                for (const auto& c : ws_starts[a]) res.insert({c.i, c.j});
            for (const auto& c : out.walls) res.erase({c.i, c.j});
            for (const auto& c : out.exits) res.erase({c.i, c.j});
            std::vector<Cell> lst;
            for (const auto& ij : res) lst.push_back(Cell{ij.first, ij.second});
            out.starts.push_back(lst);
        }
        out.sources = sources;
        return true;
    } catch (const NotV2&) {
        return false;
    }
}

}  // namespace lle
