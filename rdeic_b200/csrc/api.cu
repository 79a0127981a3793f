// Library-level entry points: thread-local error string, ABI version.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/rdeic_b200.h"

namespace rdeic {

char* err_buf() {
    static thread_local char buf[1024] = {0};
    return buf;
}

// Programmatic dependent launch is ON by default, with the implicit trigger (see pdl_trigger() in common.cuh for the
// measurements behind both choices).  RDEIC_PDL=0 or RDEIC_NO_PDL=1 turns it off.
bool pdl_enabled() {
    static const bool on = !(getenv("RDEIC_PDL") != nullptr && atoi(getenv("RDEIC_PDL")) == 0) && getenv("RDEIC_NO_PDL") == nullptr;
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
