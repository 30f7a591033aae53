// lle_b200 — layout generator: host side of the lle_gen_* entry points (include/lle_b200.h) and the kernel that runs
// one attempt of the reference's WorldGenerator per thread (gen_core.cuh).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lle_b200.h"
#include "gen_core.cuh"
#include "gen_config.hpp"

namespace lle {
int fail(int code, const std::string& msg);  // vec_world.cu
}
using lle::fail;

#define GEN_CUDA(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return fail(_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver ? LLE_NO_DEVICE : LLE_CUDA_ERROR, \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                               \
    } while (0)

namespace {

// init_genrand(19650218): 624 words, read with a warp-uniform index by the first seeding pass
__constant__ uint32_t c_mt_base[624];
struct ConstTable {
    __device__ __forceinline__ uint32_t operator()(int i) const { return c_mt_base[i]; }
};

// Thread per chain, grid-stride.  What an attempt costs is mostly CPython's seeding: init_by_array is 1,247 dependent steps
// over the 624-word MT19937 state (gen_core.cuh), so the kernel is latency- and issue-bound and wants many warps per SM to
// overlap those chains.  The state (2.5 KB) and the placement scratch (candidate list, masks) live in local memory, which
// the hardware interleaves across the lanes of a warp; 128-thread CTAs at <= 64 registers, `ctas_per_sm` of them resident
// (8 for small grids = 1,024 threads per SM; fewer for large grids, whose candidate lists would otherwise push the
// states out of L1/L2).  Measured alternatives: the states in lane-interleaved shared memory (92 threads per
// SM fill the 227 KB) are 3x slower - three warps cannot hide the chain.  Attempts of a warp diverge freely (rejection
// sampling, retries).
// Threads per block / minimum blocks per SM (registers): swept in round 2 on B200, M attempts/s on 5x5 / 10x10 / 8x8 lanes / 32x32:
// (128, 8) 402 / 65 / 283 / 3.1;  (128, 6) 317 / 65 / 282 / 3.0;  (128, 10) 385 / 64 / 273 / 3.1;  (128, 12) 373 / 64 / 269 / 3.0;
// (64, 16) 359 / 52 / 215 / 2.7;  (256, 4) 414 / 61 / 290 / 2.1 - occupancy is not the lever.
#ifndef LLE_GEN_THREADS
#define LLE_GEN_THREADS 128
#endif
#ifndef LLE_GEN_MIN_BLOCKS
#define LLE_GEN_MIN_BLOCKS 8
#endif
constexpr int kGenThreads = LLE_GEN_THREADS;

__global__ void __launch_bounds__(kGenThreads, LLE_GEN_MIN_BLOCKS) lle_gen_kernel(const __grid_constant__ llegen::Config cfg, const uint64_t* __restrict__ seeds,
                                                                 uint64_t first_seed, int64_t n, int max_attempts, uint8_t require,
                                                                 uint8_t* __restrict__ cells, uint8_t* __restrict__ status,
                                                                 uint8_t* __restrict__ labels, int32_t* __restrict__ tries) {
    llegen::PyRandom rng;
    alignas(8) uint16_t work[llegen::kWork];
    const int hw = cfg.height * cfg.width;
    for (int64_t i = (int64_t)blockIdx.x * kGenThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kGenThreads) {
        const uint64_t seed = seeds ? seeds[i] : first_seed + (uint64_t)i;
        llegen::chain(cfg, ConstTable{}, seed, max_attempts, require, rng, work, cells + i * hw, status + i, labels + i, tries + i);
    }
}

struct HostMT {  // CPython's generator on the host, for lle_gen_attempt_seeds only
    llegen::PyRandom r;
    explicit HostMT(uint64_t seed) {
        uint32_t base[624];
        llegen::mt_base_table(base);
        r.seed(seed, llegen::TablePtr{base});
    }
    // getrandbits(63): two 32-bit words, low word first, the last one shifted down (Modules/_randommodule.c)
    uint64_t bits63() {
        const uint64_t lo = r.next();
        const uint64_t hi = r.next() >> 1;
        return lo | (hi << 32);
    }
};

}  // namespace

struct lle_gen {
    lle_gen_options opts;
    llegen::Config cfg;
    int device = 0;
    int64_t capacity = 0, n = 0;
    uint8_t* d_cells = nullptr;
    uint8_t* d_status = nullptr;
    uint8_t* d_labels = nullptr;
    int32_t* d_tries = nullptr;
    int grid = 0;
};

extern "C" {

LLE_API void lle_gen_default_options(lle_gen_options* o) {
    std::memset(o, 0, sizeof(*o));
    o->width = o->height = 5;
    o->n_agents = 2;  // generator.py:101
    o->n_walls = LLE_GEN_WALLS_AUTO;
    o->door_size = 1;
    o->cluster_h = o->cluster_w = 1;
}

LLE_API int lle_gen_create(const lle_gen_options* opts, int32_t device, int64_t capacity, lle_gen** out) {
    if (!opts || !out || capacity < 1) return fail(LLE_INVALID_ARGUMENT, "lle_gen_create: null argument or capacity < 1");
    const lle_gen_options& o = *opts;
    llegen::Config c;
    std::string why;
    if (const int rc = llegen::resolve_config(o, c, why)) return fail(rc, why);
    const int area = o.width * o.height;
    GEN_CUDA(cudaSetDevice(device));
    auto g = new lle_gen();
    g->opts = o, g->cfg = c, g->device = device, g->capacity = capacity;
    uint32_t base[624];
    llegen::mt_base_table(base);
    cudaError_t e = cudaMemcpyToSymbol(c_mt_base, base, sizeof(base));
    if (e == cudaSuccess) e = cudaMalloc(&g->d_cells, (size_t)capacity * area);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_status, (size_t)capacity);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_labels, (size_t)capacity);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_tries, (size_t)capacity * sizeof(int32_t));
    int sms = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        lle_gen_destroy(g);
        return fail(LLE_CUDA_ERROR, std::string("lle_gen_create: ") + cudaGetErrorString(e));
    }
    // resident CTAs per SM by grid size (measured on B200: 5x5 405 M attempts/s at 8, 10x10 65 M at 6, 32x32 3.1 M at 4);
    // LLE_GEN_CTAS overrides it (development knob for A/B timing)
    int ctas = area <= 36 ? 8 : area <= 144 ? 6 : 4;
    if (const char* knob = std::getenv("LLE_GEN_CTAS"))
        if (std::atoi(knob) > 0) ctas = std::atoi(knob);
    g->grid = sms * ctas;
    *out = g;
    return LLE_OK;
}

LLE_API int lle_gen_destroy(lle_gen* g) {
    if (!g) return LLE_OK;
    cudaSetDevice(g->device);
    cudaFree(g->d_cells);
    cudaFree(g->d_status);
    cudaFree(g->d_labels);
    cudaFree(g->d_tries);
    delete g;
    return LLE_OK;
}

LLE_API int lle_gen_attempt_seeds(uint64_t seed, int64_t n, uint64_t* out) {
    if (n < 0 || (n > 0 && !out)) return fail(LLE_INVALID_ARGUMENT, "lle_gen_attempt_seeds: bad arguments");
    HostMT mt(seed);
    const uint64_t bound = 0x7fffffffffffffffull;  // sys.maxsize; bit_length 63
    for (int64_t i = 0; i < n; ++i) {
        uint64_t r = mt.bits63();
        while (r >= bound) r = mt.bits63();
        out[i] = r;
    }
    return LLE_OK;
}

LLE_API int lle_gen_run(lle_gen* g, const uint64_t* seeds_dev, uint64_t first_seed, int64_t n, int32_t max_attempts, uint32_t require,
                        void* cuda_stream) {
    if (!g || n < 0 || n > g->capacity) return fail(LLE_INVALID_ARGUMENT, "lle_gen_run: n must be in [0, capacity]");
    if (max_attempts < 1) return fail(LLE_INVALID_ARGUMENT, "max_attempts must be >= 1. Got " + std::to_string(max_attempts));  // generator.py:279-280
    if (require > 7) return fail(LLE_INVALID_ARGUMENT, "lle_gen_run: unknown label bits in require");
    GEN_CUDA(cudaSetDevice(g->device));
    g->n = n;
    if (n == 0) return LLE_OK;
    const int64_t blocks_needed = (n + kGenThreads - 1) / kGenThreads;
    const int grid = (int)(blocks_needed < g->grid ? blocks_needed : g->grid);
    lle_gen_kernel<<<grid, kGenThreads, 0, (cudaStream_t)cuda_stream>>>(g->cfg, seeds_dev, first_seed, n, max_attempts, (uint8_t)require, g->d_cells,
                                                                        g->d_status, g->d_labels, g->d_tries);
    GEN_CUDA(cudaGetLastError());
    return LLE_OK;
}

LLE_API int lle_gen_get_buffers(lle_gen* g, lle_gen_buffers* out) {
    if (!g || !out) return fail(LLE_INVALID_ARGUMENT, "lle_gen_get_buffers: null argument");
    out->capacity = g->capacity, out->n = g->n;
    out->height = g->cfg.height, out->width = g->cfg.width;
    out->cells = g->d_cells, out->status = g->d_status, out->labels = g->d_labels, out->tries = g->d_tries;
    return LLE_OK;
}

LLE_API int lle_gen_fetch(lle_gen* g, int64_t first, int64_t n, uint8_t* cells, uint8_t* status, uint8_t* labels, int32_t* tries, void* cuda_stream) {
    if (!g || first < 0 || n < 0 || first + n > g->n) return fail(LLE_INVALID_ARGUMENT, "lle_gen_fetch: the range must lie within the last run");
    GEN_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t hw = (size_t)g->cfg.height * g->cfg.width;
    if (n > 0) {
        if (cells) GEN_CUDA(cudaMemcpyAsync(cells, g->d_cells + (size_t)first * hw, (size_t)n * hw, cudaMemcpyDeviceToHost, st));
        if (status) GEN_CUDA(cudaMemcpyAsync(status, g->d_status + first, (size_t)n, cudaMemcpyDeviceToHost, st));
        if (labels) GEN_CUDA(cudaMemcpyAsync(labels, g->d_labels + first, (size_t)n, cudaMemcpyDeviceToHost, st));
        if (tries) GEN_CUDA(cudaMemcpyAsync(tries, g->d_tries + first, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    GEN_CUDA(cudaStreamSynchronize(st));
    return LLE_OK;
}

LLE_API int lle_gen_geometry_valid(const uint8_t* cells, int32_t H, int32_t W, int32_t* valid) {
    if (!cells || !valid || H < 1 || W < 1) return fail(LLE_INVALID_ARGUMENT, "lle_gen_geometry_valid: bad arguments");
    std::vector<char> lit((size_t)H * W, 0);
    *valid = 1;
    for (int r = 0; r < H && *valid; ++r)
        for (int c = 0; c < W && *valid; ++c) {
            const uint8_t v = cells[r * W + c];
            if (v < LLE_CELL_SOURCE) continue;
            const int d = (v - LLE_CELL_SOURCE) & 3;  // 0 N, 1 S, 2 E, 3 W
            const int dr = d == 0 ? -1 : d == 1 ? 1 : 0, dc = d == 2 ? 1 : d == 3 ? -1 : 0;
            int len = 0;
            for (int i = r + dr, j = c + dc; i >= 0 && i < H && j >= 0 && j < W; i += dr, j += dc) {
                const uint8_t t = cells[i * W + j];
                if (t == LLE_CELL_WALL || t >= LLE_CELL_SOURCE) break;
                lit[(size_t)i * W + j] = 1;
                ++len;
            }
            if (len < 2) *valid = 0;
        }
    for (int k = 0; k < H * W && *valid; ++k)
        if (cells[k] == LLE_CELL_EXIT && lit[k]) *valid = 0;
    return LLE_OK;
}

LLE_API int lle_gen_cells_to_text(const uint8_t* cells, int32_t height, int32_t width, char* out, size_t cap, size_t* len) {
    if (!cells || height < 1 || width < 1) return fail(LLE_INVALID_ARGUMENT, "lle_gen_cells_to_text: bad arguments");
    std::string s;
    for (int r = 0; r < height; ++r) {
        if (r) s += '\n';
        for (int c = 0; c < width; ++c) {
            if (c) s += ' ';
            const uint8_t v = cells[r * width + c];
            if (v == LLE_CELL_FLOOR) s += '.';
            else if (v == LLE_CELL_WALL) s += '@';
            else if (v == LLE_CELL_EXIT) s += 'X';
            else if (v == LLE_CELL_GEM) s += 'G';
            else if (v >= LLE_CELL_START && v < LLE_CELL_START + 32) s += "S" + std::to_string(v - LLE_CELL_START);
            else if (v >= LLE_CELL_SOURCE) s += "L" + std::to_string((v - LLE_CELL_SOURCE) / 4) + "NSEW"[(v - LLE_CELL_SOURCE) % 4];
            else return fail(LLE_INVALID_ARGUMENT, "lle_gen_cells_to_text: unknown cell code " + std::to_string(v));
        }
    }
    if (len) *len = s.size();
    if (out && cap) {
        const size_t m = s.size() < cap - 1 ? s.size() : cap - 1;
        std::memcpy(out, s.data(), m);
        out[m] = 0;
    }
    return LLE_OK;
}

}  // extern "C"
