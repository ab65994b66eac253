/*
 * pipe_b200.h -- C-ABI of the B200-native Processor hot path for pipelined/pipe.
 *
 * This is the drop-in boundary: plain C, pointers and sizes only, no torch or
 * C++ types.  Each entry point names the reference interface it stands behind
 * (paths relative to the reference repo, pipelined/pipe @7600d57).  The cgo
 * binding a pipe maintainer would add is in INTEGRATION.md and go/.
 *
 * Threading (pipe.go:425-451, run.go:173-196): the reference calls one
 * component's Start -> ProcessFunc* -> Flush from exactly one goroutine, never
 * concurrently with itself.  Accordingly a pb_chain is re-entrant per handle
 * but serves one caller at a time.  Goroutines migrate between OS threads, so
 * no entry point relies on the thread's current CUDA device: every call selects
 * the chain's device itself.
 *
 * Ownership (pipe.go:431,437): `in` is framework-owned and recycled when the
 * call returns, `out` is framework-allocated.  No entry point retains a caller
 * pointer after it returns (pb_chain_submit copies the input into the chain's
 * own pinned staging before returning unless the caller registered the memory).
 *
 * Buffer layout (mock/mock.go:95-101): frame-major, channel-interleaved,
 * value index = frame * channels + channel; "frames" is samples per channel,
 * what signal.Floating.Length() returns.
 *
 * There is no CPU fallback: every compute entry point fails with
 * PB_ERR_CUDA / PB_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef PIPE_B200_H
#define PIPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 1

/* ---- status codes: the `error` half of (int, error) in pipe.go:47,64,80 -- */
enum pb_status {
    PB_OK = 0,
    PB_ERR_INVALID = -1,     /* bad argument / descriptor                      */
    PB_ERR_CUDA = -2,        /* a CUDA call or kernel failed                    */
    PB_ERR_NO_DEVICE = -3,   /* no usable device: there is no CPU fallback      */
    PB_ERR_NOMEM = -4,
    PB_ERR_UNSUPPORTED = -5, /* e.g. resample with up > down (pipe.go:437-443)  */
    PB_ERR_CAPACITY = -6,    /* output buffer too small / batch too large       */
    PB_ERR_STATE = -7        /* call order violated (collect without submit...) */
};

enum pb_dtype { PB_F32 = 0, PB_F64 = 1 };

/* Processor kinds.  COPY is mock.Processor (mock/mock.go:147-154); the others
 * are the build-defined Processors named in BASELINE.json. */
enum pb_stage_kind {
    PB_STAGE_COPY = 0,
    PB_STAGE_GAIN = 1,     /* y = gain * x                                       */
    PB_STAGE_BIQUAD = 2,   /* TDF-II: y=b0x+s1; s1=b1x-a1y+s2; s2=b2x-a2y         */
    PB_STAGE_FIR = 3,      /* y[n] = sum_k taps[k] x[n-k]; history carried        */
    PB_STAGE_RESAMPLE = 4  /* up/down polyphase; per input frame acc+=up,
                              if acc>=down {acc-=down; emit branch up-1-acc}      */
};

typedef struct pb_stage_desc {
    int32_t kind;        /* pb_stage_kind */
    int32_t n_taps;      /* FIR: taps; RESAMPLE: prototype length = up * taps_per_phase */
    int32_t up, down;    /* RESAMPLE, 1 <= up <= down */
    int32_t _pad;
    double gain;         /* GAIN */
    double b[3];         /* BIQUAD b0 b1 b2 (a0 == 1) */
    double a[2];         /* BIQUAD a1 a2 */
    const double *taps;  /* FIR / RESAMPLE coefficients, copied at create/set */
} pb_stage_desc;

/* One contiguous run of GPU Processors in Line.Processors (line.go:14-19),
 * i.e. what a ProcessorAllocatorFunc (line.go:30) is given: bufferSize and the
 * input SignalProperties{SampleRate, Channels} (line.go:38-41). */
typedef struct pb_chain_desc {
    int32_t abi_version;    /* PB_ABI_VERSION */
    int32_t device;         /* CUDA ordinal */
    int32_t dtype;          /* pb_dtype of the buffers and of the arithmetic */
    int32_t channels;       /* SignalProperties.Channels */
    double sample_rate;     /* SignalProperties.SampleRate */
    int32_t buffer_frames;  /* bufferSize of pipe.New / pipe.Run (pipe.go:90,107) */
    int32_t max_batch;      /* max buffers per pb_chain_process_batch* call (>= 1) */
    int32_t n_stages;
    int32_t flags;          /* PB_CHAIN_* */
    const pb_stage_desc *stages;
} pb_chain_desc;

#define PB_CHAIN_METER 1u       /* fused meter sink: per-channel peak and sum of squares of the output */
#define PB_CHAIN_NO_TENSOR 2u   /* never take the tcgen05 FIR path (exact-order float32 arithmetic instead) */
#define PB_CHAIN_NO_STREAM 4u   /* never take the streaming kernels: the generic tile kernel serves FIR-less runs too */

typedef struct pb_chain pb_chain;

/* ProcessorAllocatorFunc (line.go:30, called at line.go:71): allocate every
 * buffer and table the chain needs.  Errors abort binding (line.go:72-74). */
int32_t pb_chain_create(const pb_chain_desc *desc, pb_chain **out);
/* FlushFunc (pipe.go:86; run.go:181-185): the guaranteed teardown point. */
int32_t pb_chain_destroy(pb_chain *c);
/* Zero all carried state (restart of a Pipe, pipe_test.go:124-130). */
int32_t pb_chain_reset(pb_chain *c);
/* Output SignalProperties of the run, threaded to the next stage (line.go:75). */
int32_t pb_chain_out_properties(const pb_chain *c, int32_t *channels, double *sample_rate);
/* Frames the next call would emit for in_frames, given the carried phase. */
int32_t pb_chain_peek_out_frames(const pb_chain *c, int64_t in_frames, int64_t *out_frames);

/* ProcessFunc(in, out) (int, error) (pipe.go:64, called at pipe.go:438) with
 * host buffers: H2D copy, fused kernel(s), D2H copy, synchronous.  *out_frames
 * is `processed`; a short value is how pipe.go:441-443 slices the output.
 * A buffer of 8 MiB or more passes in four pieces so that the copies and the
 * kernels overlap; the result is that of four consecutive shorter calls. */
int32_t pb_chain_process(pb_chain *c, const void *in_host, int64_t in_frames,
                         void *out_host, int64_t out_capacity_frames, int64_t *out_frames);

/* Same step with device-resident buffers on `stream` (a cudaStream_t, may be
 * NULL); asynchronous with respect to the host.  n_buffers consecutive buffers
 * are processed by one launch; buffer i has buf_frames[i] frames (only the last
 * may be short, pipe.go:404-406) and its outputs follow buffer i-1's in
 * out_dev.  buf_out_frames[i] receives the per-buffer `processed` count.
 * Stream ordering: the carried state is shared with the host-buffer paths, which run on the chain's own streams.  A
 * caller that mixes this entry point with pb_chain_process / pb_chain_submit on the same chain must call pb_chain_sync
 * in between (one component, one caller at a time: run.go:175-194). */
int32_t pb_chain_process_batch_device(pb_chain *c, const void *in_dev, const int64_t *buf_frames,
                                      int32_t n_buffers, void *out_dev, int64_t out_capacity_frames,
                                      int64_t *buf_out_frames, void *stream);

/* Wait for everything the chain enqueued on `stream` (and its own streams) and
 * report a kernel-side failure, if any, as PB_ERR_CUDA.  The device-resident
 * path is asynchronous; this is its error-collection point. */
int32_t pb_chain_sync(pb_chain *c, void *stream);

/* Pipelined host path: the analogue of the cap-1 async fitting
 * (internal/fitting/fitting.go:56-60) -- up to pb_chain_pipeline_depth()
 * submitted batches may be in flight.  submit enqueues H2D + kernels + D2H and
 * returns; collect waits for the oldest submitted batch. */
int32_t pb_chain_pipeline_depth(const pb_chain *c);
int32_t pb_chain_submit(pb_chain *c, const void *in_host, const int64_t *buf_frames, int32_t n_buffers,
                        void *out_host, int64_t out_capacity_frames);
int32_t pb_chain_collect(pb_chain *c, int64_t *buf_out_frames, int32_t n_buffers);

/* Mutations (mutable/mutable.go:40-48, applied at pipe.go:433): replace the
 * parameters of one stage; takes effect at the next process call, carried
 * state kept.  kind, n_taps and up/down must not change. */
int32_t pb_chain_set_stage(pb_chain *c, int32_t stage_index, const pb_stage_desc *stage);

/* InsertProcessor on a fused run (pipe.go:297-333; the executor side is run.go:134-169): insert `stage` at position
 * `pos` (0 .. n_stages) of the run.  The run is re-planned into fused kernels; every stage that was already there keeps
 * its carried state (FIR history, biquad state, resampler history and phase), the new stage starts from zero state like
 * a freshly allocated Processor.  Takes effect at the next process call.  (AddLine, pipe.go:260-295, needs nothing here:
 * a new Line is a new pb_chain.) */
int32_t pb_chain_insert_stage(pb_chain *c, int32_t pos, const pb_stage_desc *stage);

/* Fused meter sink (PB_CHAIN_METER): per-channel peak |y| and sum of y^2 of
 * everything emitted since create/reset, plus the frame count.  Synchronises. */
int32_t pb_chain_meter_read(pb_chain *c, double *peak, double *sumsq, int64_t *frames);

/* Which kernel family served the last process call: 0 none yet, 1 generic
 * fused tile kernel, 2 tcgen05/TMA chain kernel, 3 streaming kernels (runs without
 * FIR and resampler).  kernels = launches so far. */
int32_t pb_chain_last_path(const pb_chain *c, int32_t *path, int64_t *kernel_launches);

/* ---- Source / Sink side kernels ----------------------------------------- */

/* Synthetic Source on the device (the mock.Source analogue, mock.go:86-105,
 * with the BASELINE.md input formula instead of a constant):
 * v[i] = (splitmix64(seed ^ line<<48 ^ (first_index+i)) >> 40) / 2^23 - 1. */
int32_t pb_source_fill_device(int32_t device, int32_t dtype, void *out_dev, int64_t first_index,
                              int64_t n_values, uint64_t seed, uint64_t line, void *stream);

/* Stand-alone meter Sink (SinkFunc, pipe.go:80): per-channel peak and sum of
 * squares of a device buffer, accumulated INTO peak_dev/sumsq_dev (doubles). */
int32_t pb_meter_device(int32_t device, int32_t dtype, const void *in_dev, int64_t frames, int32_t channels,
                        double *peak_dev, double *sumsq_dev, void *stream);

/* Fan-in mixer Sink of BASELINE.json configs[4] (build-defined; the reference's
 * merger.go merges error channels): out[i] = sum_l inputs[l][i].
 * inputs may be peer-GPU pointers opened with pb_ipc_open: the sum then pulls
 * the other Lines' buffers over NVLink inside the same kernel. */
int32_t pb_mix_sum_device(int32_t device, int32_t dtype, const void *const *inputs_dev, int32_t n_inputs,
                          int64_t n_values, void *out_dev, void *stream);

/* ---- memory helpers for torch-free hosts (Go, C++) ----------------------- */
int32_t pb_device_count(int32_t *count);
int32_t pb_device_alloc(int32_t device, int64_t bytes, void **ptr_dev);
int32_t pb_device_free(int32_t device, void *ptr_dev);
int32_t pb_host_alloc_pinned(int64_t bytes, void **ptr_host);
int32_t pb_host_free_pinned(void *ptr_host);
int32_t pb_memcpy_h2d(int32_t device, void *dst_dev, const void *src_host, int64_t bytes);
int32_t pb_memcpy_d2h(int32_t device, void *dst_host, const void *src_dev, int64_t bytes);
int32_t pb_device_synchronize(int32_t device);
/* cross-process peer access for the fan-in sum: 64-byte opaque handles */
int32_t pb_ipc_export(int32_t device, void *ptr_dev, uint8_t handle[64]);
/* the handle names the allocation ptr_dev lies in and pb_ipc_open returns that allocation's base: add this offset */
int32_t pb_ipc_offset(int32_t device, void *ptr_dev, int64_t *offset);
int32_t pb_ipc_open(int32_t device, const uint8_t handle[64], void **ptr_dev);
int32_t pb_ipc_close(int32_t device, void *ptr_dev);

/* ---- diagnostics ---------------------------------------------------------- */
int32_t pb_abi_version(void);
/* Thread-local text of the last failure on the calling thread. */
const char *pb_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PIPE_B200_H */
