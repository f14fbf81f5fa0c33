/* oracle/tsc_hz.c -- TEST INFRASTRUCTURE ONLY: the time-stamp counter frequency of this host in Hz, measured against CLOCK_MONOTONIC
 * over 200 ms.  bench.py uses it to turn the cycle counts the reference's -DDO_TIMING build prints (rdtsc, src/GROM.c:1111-1120) into seconds. */
#include <stdint.h>
#include <stdio.h>
#include <time.h>
static inline uint64_t rdtsc(void) { uint32_t lo, hi; __asm__ __volatile__("rdtsc" : "=a"(lo), "=d"(hi)); return ((uint64_t)hi << 32) | lo; }
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(void)
{
    const double t0 = now(); const uint64_t c0 = rdtsc();
    while (now() - t0 < 0.2) { }
    const double t1 = now(); const uint64_t c1 = rdtsc();
    printf("%.0f\n", (double)(c1 - c0) / (t1 - t0));
    return 0;
}
