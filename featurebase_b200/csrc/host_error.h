// Error carrier of the host-only parts of libfbgpu (program compiler, roaring / RBF readers): an FBGPU_E_* code plus a
// message; fbgpu.cu copies it into the thread-local string behind fbgpu_last_error().
#pragma once
#include <stdarg.h>
#include <stdio.h>

namespace fbgpu {
struct Error {
    int code = 0; char msg[512] = { 0 };
    int set(int c, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(msg, sizeof msg, fmt, ap); va_end(ap); code = c; return c; }
};
}  // namespace fbgpu
