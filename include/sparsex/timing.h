/*
 * Wall-clock timers of the SparseX C API (replaces include/sparsex/timing.h of SparseX v1.1.0:
 * spx_timer_t and spx_timer_clear / start / pause / get_secs, used by every program under
 * src/examples and by test/src/sparsex_test.c).  Same names, same accumulate-between-start-and-pause
 * behaviour; one microsecond counter pair instead of the reference's pair of struct timeval
 * (gettimeofday is visible under strict -std=c99 as well, which the reference's programs may be built with).
 *
 * Note for timing GPU work: the spx_matvec_* calls return after the result is in the caller's
 * buffers (host vectors) or after the kernels are queued (library vectors with spx.b200.async=true);
 * spx_vec_print / element access synchronise.
 */
#ifndef SPARSEX_TIMING_H
#define SPARSEX_TIMING_H

#include <stdlib.h>
#include <sys/time.h>

typedef struct spx_timer {
    long long elapsed_us;   /* accumulated over all start/pause intervals */
    long long started_us;   /* time of the last spx_timer_start */
} spx_timer_t;

static inline long long spx_timer_now_us_(void)
{
    struct timeval tv;
    if (gettimeofday(&tv, NULL) < 0)
        exit(1);
    return (long long) tv.tv_sec * 1000000LL + tv.tv_usec;
}

static inline void spx_timer_clear(spx_timer_t *t)
{
    t->elapsed_us = 0;
    t->started_us = 0;
}

static inline void spx_timer_start(spx_timer_t *t)
{
    t->started_us = spx_timer_now_us_();
}

static inline void spx_timer_pause(spx_timer_t *t)
{
    t->elapsed_us += spx_timer_now_us_() - t->started_us;
}

static inline double spx_timer_get_secs(spx_timer_t *t)
{
    return (double) t->elapsed_us * 1e-6;
}

#endif /* SPARSEX_TIMING_H */
