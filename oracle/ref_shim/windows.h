/* Minimal stand-in for <windows.h> so that the UNMODIFIED reference sources
 * (/root/reference/main.cuh:6, NTT.cu:6-8 and the timer sites) compile on Linux.
 * Only the three Win32 symbols the reference uses are provided.  Test infrastructure. */
#ifndef QT_REF_SHIM_WINDOWS_H
#define QT_REF_SHIM_WINDOWS_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
typedef union {
    struct { uint32_t LowPart; int32_t HighPart; };
    long long QuadPart;
} LARGE_INTEGER;
static inline int QueryPerformanceCounter(LARGE_INTEGER* t) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    t->QuadPart = (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
    return 1;
}
static inline int QueryPerformanceFrequency(LARGE_INTEGER* f) { f->QuadPart = 1000000000ll; return 1; }
#endif
