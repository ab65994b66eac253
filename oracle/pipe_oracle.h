/*
 * pipe_oracle.h -- CPU oracle for the pipe_b200 parity tests.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pipe_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline, never as the product path.
 *
 * Two halves:
 *
 *  1. PLUMBING (pinned): a plain-C restatement of the reference's per-buffer
 *     hot loop -- Source.execute / Processor.execute / Sink.execute
 *     (/root/reference/pipe.go:381-471), lineExecutor / multiLineExecutor and
 *     the sync run loop (run.go:38-132,200-224), the sync fitting
 *     (internal/fitting/fitting.go:62-79) and the mock components
 *     (mock/mock.go:86-105,147-154,180-189).  It is pinned by the reference's
 *     own integer goldens (pipe_test.go:104-105,337,363,394,399,404;
 *     mock/mock_test.go:69-92,133-146,185-202) in tests/test_oracle_plumbing.py.
 *
 *  2. DSP (parity UNPINNED against the reference): gain, biquad, FIR,
 *     rational resampler and the fan-in sum do not exist in the reference
 *     (SURVEY.md section 0, D2/D3).  Their specification is this file.  It is
 *     cross-checked against scipy.signal (lfilter / upfirdn) and closed-form
 *     cases in tests/test_oracle_dsp.py, and frozen in tests/golden/.
 *     Arithmetic is float64, the type the reference allocates
 *     (pipe.go:394,437).
 *
 * Buffer layout everywhere: frame-major, channel-interleaved,
 * idx = frame * channels + channel (mock/mock.go:100-101 fills a flat index;
 * Length() is frames per channel, mock.go:95).
 */
#ifndef PIPE_ORACLE_H
#define PIPE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ DSP -- */

enum {
    ORC_STAGE_COPY = 0,     /* mock.Processor pass-through, mock.go:147-154 */
    ORC_STAGE_GAIN = 1,     /* y = g * x                                     */
    ORC_STAGE_BIQUAD = 2,   /* transposed direct form II, a0 == 1            */
    ORC_STAGE_FIR = 3,      /* y[n] = sum_k h[k] x[n-k]                      */
    ORC_STAGE_RESAMPLE = 4  /* rational up/down polyphase, integer phase acc */
};

typedef struct orc_stage {
    int32_t kind;
    int32_t n_taps;      /* FIR: taps; RESAMPLE: prototype length = up * taps_per_phase */
    int32_t up, down;    /* RESAMPLE, up <= down                                      */
    int32_t _pad;
    double gain;         /* GAIN */
    double b[3];         /* BIQUAD b0 b1 b2 */
    double a[2];         /* BIQUAD a1 a2    */
    const double *taps;  /* FIR / RESAMPLE coefficients (copied at create) */
} orc_stage;

typedef struct orc_chain orc_chain;

orc_chain *orc_chain_new(int32_t channels, int32_t n_stages, const orc_stage *stages);
void orc_chain_free(orc_chain *c);
void orc_chain_reset(orc_chain *c);
int32_t orc_chain_out_channels(const orc_chain *c);
/* Frames a call with in_frames would emit given the current resampler phase. */
int64_t orc_chain_peek_out_frames(const orc_chain *c, int64_t in_frames);
/* One buffer through every stage, state carried.  Returns output frames or -1. */
int64_t orc_chain_process(orc_chain *c, const double *in, int64_t in_frames,
                          double *out, int64_t out_capacity_frames);
/* Same arithmetic, channels split over n_threads host threads (CPU baseline). */
int64_t orc_chain_process_mt(orc_chain *c, const double *in, int64_t in_frames,
                             double *out, int64_t out_capacity_frames, int32_t n_threads);
/* Replace the parameters of one stage (same kind and sizes), state kept:
 * the analogue of a mutation landing between buffers (pipe.go:433). */
int32_t orc_chain_set_stage(orc_chain *c, int32_t idx, const orc_stage *s);

/* out[i] = sum_l in[l][i], the build-defined fan-in mixer (config 5). */
void orc_mix_sum(const double *const *inputs, int32_t n_inputs, int64_t n_values, double *out);
/* per-channel peak |x| and sum of squares over a buffer (meter sink). */
void orc_meter(const double *in, int64_t frames, int32_t channels, double *peak, double *sumsq);
/* synthetic source: x = (splitmix64(seed ^ line<<48 ^ idx) >> 40) / 2^23 - 1 */
void orc_source_fill(double *out, int64_t first_index, int64_t n_values, uint64_t seed, uint64_t line);

/* ------------------------------------------------------------- plumbing -- */

#define ORC_MAX_PROCS 8

typedef struct orc_mock_component {
    /* knobs, mock.go:23-39,60-73,130-137,160-168 */
    int32_t error_on_call, error_on_make, error_on_start, error_on_flush;
    /* observed */
    int32_t started, flushed;
    int64_t messages, samples; /* Counter, mock.go:17-21,43-46: samples = frames */
} orc_mock_component;

typedef struct orc_mock_line {
    /* mock.Source settings */
    int64_t limit;
    int32_t channels;
    int32_t n_procs;
    double value;
    int32_t sink_discard;
    int32_t _pad;
    orc_mock_component source;
    orc_mock_component procs[ORC_MAX_PROCS];
    orc_mock_component sink;
    /* mock.Sink.Counter.Values when !discard: caller-provided, may be NULL */
    double *sink_values;
    int64_t sink_values_capacity; /* in values (frames*channels) */
    int64_t sink_values_len;
} orc_mock_line;

enum {
    ORC_RUN_OK = 0,
    ORC_RUN_ERR_BIND = 1,   /* allocator error, line.go:65,73,81              */
    ORC_RUN_ERR_START = 2,  /* "error starting", run.go:201-203               */
    ORC_RUN_ERR_EXEC = 4,   /* ErrorRun.ErrExec, run.go:209-222               */
    ORC_RUN_ERR_FLUSH = 8   /* ErrorRun.ErrFlush or flush during start error  */
};

/* pipe.Run (pipe.go:90-103): all lines in one goroutine.  Returns a bitmask of
 * ORC_RUN_* describing the returned error (0 == nil). */
int32_t orc_pipe_run(int64_t buffer_size, int32_t n_lines, orc_mock_line *lines);

/* Calls SourceFunc in a loop until it errors, as mock_test.go:34-41 does. */
int32_t orc_mock_source_drain(int64_t buffer_size, orc_mock_line *line);

#ifdef __cplusplus
}
#endif
#endif
