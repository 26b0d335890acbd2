/* A compiled caller of the C ABI, standing in for the reference-side caller of this boundary: bevy_ggrs'
 * `handle_requests` is Rust (src/schedule_systems.rs:170-289) and runs once per GGRS tick with the tick's
 * Vec<GgrsRequest>; the shim's replacement is one `bgr_handle_requests` call per tick with host arrays in and host
 * checksums out.  bench.py's `e2e` leg times THIS loop, so that the number is the boundary's and not
 * the Python interpreter's (the same loop driven through ctypes is reported beside it as `e2e_python_caller`).
 * Nothing here computes: it is a for-loop around the public entry point, built with gcc against include/. */
#include <stdint.h>
#include <string.h>
#include <time.h>

#include "../include/bevy_ggrs_b200.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ticks[i] = requests[offsets[i] .. offsets[i] + counts[i]) with session info infos[i].
 * checksums_out receives every tick's checksums back to back (checksum_counts[i] of them for tick i).
 * per_tick_s (optional) receives each call's wall time.  Returns the first non-zero status. */
__attribute__((visibility("default")))
int bgr_caller_run_ticks(bgr_engine* e, const bgr_session_info* infos, const bgr_request* requests,
                         const uint32_t* offsets, const uint32_t* counts, uint32_t n_ticks,
                         bgr_checksum* checksums_out, uint32_t checksums_cap, uint32_t* checksum_counts,
                         double* total_s, double* per_tick_s) {
    uint32_t used = 0;
    const double t0 = now_s();
    double prev = t0;
    for (uint32_t i = 0; i < n_ticks; ++i) {
        uint32_t n = 0;
        int st = bgr_handle_requests(e, &infos[i], requests + offsets[i], counts[i], checksums_out + used,
                                     checksums_cap - used, &n);
        if (st != 0) return st;
        if (n > checksums_cap - used) n = checksums_cap - used;
        checksum_counts[i] = n;
        used += n;
        if (per_tick_s) {
            const double t = now_s();
            per_tick_s[i] = t - prev;
            prev = t;
        }
    }
    *total_s = now_s() - t0;
    return 0;
}
