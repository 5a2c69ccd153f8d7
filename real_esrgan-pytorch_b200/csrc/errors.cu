#include "errors.h"

#include "../../include/resr.h"

namespace resr {
char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace resr

extern "C" const char* resr_last_error(void) { return resr::error_buffer(); }
