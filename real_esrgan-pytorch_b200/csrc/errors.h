// Thread-local error string behind resr_last_error().
#pragma once
#include <cstdarg>
#include <cstdio>

namespace resr {
char* error_buffer();
inline int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace resr
