// Library-level entry points: thread-local error string, ABI version.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/rdeic_b200.h"

namespace rdeic {

char* err_buf() {
    static thread_local char buf[1024] = {0};
    return buf;
}

// Programmatic dependent launch is OFF by default: inside the CUDA-graph UNet step it measured 4 % SLOWER than plain
// kernel-to-kernel edges (11.30 -> 10.86 ms at batch 8, scripts/ab_unet.py on one box; neutral on the eager VAE decode):
// every kernel triggers its dependents on entry, so the next grid's CTAs queue for SMs while the current grid still has
// waves to run.  RDEIC_PDL=1 turns it back on (RDEIC_NO_PDL is still honoured).
bool pdl_enabled() {
    static const bool on = getenv("RDEIC_PDL") != nullptr && atoi(getenv("RDEIC_PDL")) != 0 && getenv("RDEIC_NO_PDL") == nullptr;
    return on;
}

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 1024, fmt, ap);
    va_end(ap);
    return 1;
}

}  // namespace rdeic

extern "C" {
const char* rdeic_last_error(void) { return rdeic::err_buf(); }
int rdeic_abi_version(void) { return 5; }
}
