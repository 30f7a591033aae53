// TEST INFRASTRUCTURE: a host (g++) instantiation of the generator core that the CUDA kernel runs per thread
// (lle_b200/csrc/gen_core.cuh is __host__ __device__), so that `-m "not gpu"` tests can compare the very same code with
// the oracle and the reference's golden vectors without a GPU.  The product never loads this file.
#include <cstdint>
#include <string>
#include <vector>

#include "../../lle_b200/csrc/gen_config.hpp"

extern "C" int gen_host_run(const lle_gen_options* opts, const uint64_t* seeds, int64_t n, int max_attempts, int require, uint8_t* cells,
                            uint8_t* status, uint8_t* labels, int32_t* tries, char* err, int errlen) {
    llegen::Config cfg;
    std::string why;
    if (const int rc = llegen::resolve_config(*opts, cfg, why)) {
        if (err && errlen > 0) {
            const int m = (int)why.size() < errlen - 1 ? (int)why.size() : errlen - 1;
            std::memcpy(err, why.data(), m);
            err[m] = 0;
        }
        return rc;
    }
    uint32_t base[624];
    llegen::mt_base_table(base);
    static thread_local llegen::PyRandom rng;
    alignas(8) static thread_local uint16_t work[llegen::kWork];
    const int hw = cfg.height * cfg.width;
    for (int64_t i = 0; i < n; ++i) 
        llegen::chain(cfg, llegen::TablePtr{base}, seeds[i], max_attempts, (uint8_t)require, rng, work, cells + i * hw, status + i, labels + i, tries + i);
    return 0;
}
