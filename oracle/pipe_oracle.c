/*
 * pipe_oracle.c -- CPU oracle (see pipe_oracle.h).  TEST INFRASTRUCTURE ONLY.
 *
 * Reference citations are relative to /root/reference (pipelined/pipe @7600d57).
 */
#include "pipe_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== */
/*                                   DSP                                    */
/* ======================================================================== */

typedef struct {
    orc_stage desc;
    double *taps;      /* owned copy */
    int32_t in_ch;
    /* carried state */
    double *bq_state;  /* [channels][2]  (s1, s2)                     */
    double *hist;      /* FIR: [(T-1)][C]; RESAMPLE: [(P-1)][C]        */
    int32_t hist_frames;
    int64_t acc;       /* RESAMPLE integer phase accumulator, 0 <= acc < down */
} stage_t;

struct orc_chain {
    int32_t channels;
    int32_t n_stages;
    stage_t *st;
    double *scratch[2];
    int64_t scratch_frames;
};

static int32_t stage_hist_frames(const orc_stage *s)
{
    if (s->kind == ORC_STAGE_FIR) return s->n_taps - 1;
    if (s->kind == ORC_STAGE_RESAMPLE) return s->n_taps / s->up - 1;
    return 0;
}

static int stage_valid(const orc_stage *s)
{
    switch (s->kind) {
    case ORC_STAGE_COPY:
    case ORC_STAGE_GAIN:
    case ORC_STAGE_BIQUAD:
        return 1;
    case ORC_STAGE_FIR:
        return s->n_taps >= 1 && s->taps != NULL;
    case ORC_STAGE_RESAMPLE:
        /* output can never exceed the input frame count (pipe.go:437-443:
         * the out buffer is bufferSize frames), so only up <= down */
        return s->up >= 1 && s->down >= s->up && s->n_taps >= s->up &&
               s->n_taps % s->up == 0 && s->taps != NULL;
    }
    return 0;
}

orc_chain *orc_chain_new(int32_t channels, int32_t n_stages, const orc_stage *stages)
{
    if (channels < 1 || n_stages < 0) return NULL;
    for (int i = 0; i < n_stages; i++)
        if (!stage_valid(&stages[i])) return NULL;
    orc_chain *c = (orc_chain *)calloc(1, sizeof(*c));
    c->channels = channels;
    c->n_stages = n_stages;
    c->st = (stage_t *)calloc((size_t)(n_stages > 0 ? n_stages : 1), sizeof(stage_t));
    for (int i = 0; i < n_stages; i++) {
        stage_t *s = &c->st[i];
        s->desc = stages[i];
        s->in_ch = channels;
        if (stages[i].n_taps > 0 && stages[i].taps) {
            s->taps = (double *)malloc(sizeof(double) * (size_t)stages[i].n_taps);
            memcpy(s->taps, stages[i].taps, sizeof(double) * (size_t)stages[i].n_taps);
            s->desc.taps = s->taps;
        }
        s->bq_state = (double *)calloc((size_t)channels * 2, sizeof(double));
        s->hist_frames = stage_hist_frames(&stages[i]);
        s->hist = (double *)calloc((size_t)(s->hist_frames > 0 ? s->hist_frames : 1) * (size_t)channels,
                                   sizeof(double));
    }
    return c;
}

void orc_chain_free(orc_chain *c)
{
    if (!c) return;
    for (int i = 0; i < c->n_stages; i++) {
        free(c->st[i].taps);
        free(c->st[i].bq_state);
        free(c->st[i].hist);
    }
    free(c->st);
    free(c->scratch[0]);
    free(c->scratch[1]);
    free(c);
}

void orc_chain_reset(orc_chain *c)
{
    for (int i = 0; i < c->n_stages; i++) {
        stage_t *s = &c->st[i];
        memset(s->bq_state, 0, sizeof(double) * (size_t)c->channels * 2);
        memset(s->hist, 0, sizeof(double) * (size_t)(s->hist_frames > 0 ? s->hist_frames : 1) * (size_t)c->channels);
        s->acc = 0;
    }
}

int32_t orc_chain_out_channels(const orc_chain *c) { return c->channels; }

int32_t orc_chain_set_stage(orc_chain *c, int32_t idx, const orc_stage *s)
{
    if (idx < 0 || idx >= c->n_stages) return -1;
    stage_t *t = &c->st[idx];
    if (s->kind != t->desc.kind || !stage_valid(s)) return -1;
    if (s->n_taps != t->desc.n_taps || s->up != t->desc.up || s->down != t->desc.down) {
        if (s->kind == ORC_STAGE_FIR || s->kind == ORC_STAGE_RESAMPLE) return -1;
    }
    double *taps = t->taps;
    if (taps && s->taps) memcpy(taps, s->taps, sizeof(double) * (size_t)s->n_taps);
    t->desc = *s;
    t->desc.taps = taps;
    return 0;
}

/* frames emitted by the integer phase accumulator for n input frames:
 * per input frame acc += up; if (acc >= down) { acc -= down; emit } */
static int64_t resample_count(int64_t acc, int64_t up, int64_t down, int64_t n)
{
    return (acc + n * up) / down;
}

int64_t orc_chain_peek_out_frames(const orc_chain *c, int64_t in_frames)
{
    int64_t n = in_frames;
    for (int i = 0; i < c->n_stages; i++) {
        const stage_t *s = &c->st[i];
        if (s->desc.kind == ORC_STAGE_RESAMPLE) n = resample_count(s->acc, s->desc.up, s->desc.down, n);
    }
    return n;
}

/* ---- per-stage kernels over the channel slice [c0, c1) ------------------ */

static void st_gain(const stage_t *s, const double *restrict in, int64_t n, double *restrict out,
                    int32_t C, int32_t c0, int32_t c1)
{
    const double g = (s->desc.kind == ORC_STAGE_GAIN) ? s->desc.gain : 1.0;
    for (int64_t f = 0; f < n; f++)
        for (int32_t c = c0; c < c1; c++) out[f * C + c] = g * in[f * C + c];
}

/* transposed direct form II:  y = b0 x + s1;  s1 = b1 x - a1 y + s2;  s2 = b2 x - a2 y */
static void st_biquad(stage_t *s, const double *restrict in, int64_t n, double *restrict out,
                      int32_t C, int32_t c0, int32_t c1)
{
    const double b0 = s->desc.b[0], b1 = s->desc.b[1], b2 = s->desc.b[2];
    const double a1 = s->desc.a[0], a2 = s->desc.a[1];
    for (int32_t c = c0; c < c1; c++) {
        double s1 = s->bq_state[2 * c], s2 = s->bq_state[2 * c + 1];
        for (int64_t f = 0; f < n; f++) {
            const double x = in[f * C + c];
            const double y = b0 * x + s1;
            s1 = b1 * x - a1 * y + s2;
            s2 = b2 * x - a2 * y;
            out[f * C + c] = y;
        }
        s->bq_state[2 * c] = s1;
        s->bq_state[2 * c + 1] = s2;
    }
}

static void hist_update(double *hist, int32_t H, const double *in, int64_t n, int32_t C, int32_t c0, int32_t c1)
{
    /* new hist = last H frames of (hist ++ in) */
    for (int32_t j = 0; j < H; j++) {
        const int64_t src = n - H + j; /* frame index relative to in[0] */
        for (int32_t c = c0; c < c1; c++)
            hist[(int64_t)j * C + c] = (src >= 0) ? in[src * C + c] : hist[(int64_t)(H + src) * C + c];
    }
}

/* y[n] = sum_{k=0}^{T-1} h[k] x[n-k], taps summed in ascending k */
static void st_fir(stage_t *s, const double *restrict in, int64_t n, double *restrict out,
                   int32_t C, int32_t c0, int32_t c1)
{
    const int32_t T = s->desc.n_taps, H = T - 1;
    const double *restrict h = s->taps;
    for (int64_t f = 0; f < n; f++) {
        double *restrict o = out + f * C;
        for (int32_t c = c0; c < c1; c++) o[c] = 0.0;
        for (int32_t k = 0; k < T; k++) {
            const int64_t idx = f - k;
            const double *restrict row = (idx >= 0) ? in + idx * C : s->hist + (H + idx) * C;
            const double hk = h[k];
            for (int32_t c = c0; c < c1; c++) o[c] += hk * row[c];
        }
    }
    /* hist is only read for idx < 0, i.e. rows this update does not overwrite
     * before they are consumed: the update runs after the whole buffer. */
    hist_update(s->hist, H, in, n, C, c0, c1);
}

/* Rational resampler, polyphase, integer phase accumulator.
 * Per input frame i: acc += up; if (acc >= down) { acc -= down; emit
 *   out = sum_{k<P} h[(up-1-acc) + k*up] * x[i-k] }.
 * Equivalent to scipy.signal.upfirdn([0]+h, x, up, down)[m+1] for a stream
 * starting at acc = 0. */
static int64_t st_resample(stage_t *s, const double *restrict in, int64_t n, double *restrict out,
                           int32_t C, int32_t c0, int32_t c1)
{
    const int64_t up = s->desc.up, down = s->desc.down;
    const int32_t P = (int32_t)(s->desc.n_taps / up), H = P - 1;
    const double *restrict h = s->taps;
    int64_t acc = s->acc, m = 0;
    for (int64_t i = 0; i < n; i++) {
        acc += up;
        if (acc < down) continue;
        acc -= down;
        const int64_t p = up - 1 - acc;
        double *restrict o = out + m * C;
        for (int32_t c = c0; c < c1; c++) o[c] = 0.0;
        for (int32_t k = 0; k < P; k++) {
            const int64_t idx = i - k;
            const double *restrict row = (idx >= 0) ? in + idx * C : s->hist + (H + idx) * C;
            const double hk = h[p + (int64_t)k * up];
            for (int32_t c = c0; c < c1; c++) o[c] += hk * row[c];
        }
        m++;
    }
    hist_update(s->hist, H, in, n, C, c0, c1);
    return m; /* acc is committed by the caller once (same for every slice) */
}

static void ensure_scratch(orc_chain *c, int64_t frames)
{
    if (frames <= c->scratch_frames) return;
    for (int i = 0; i < 2; i++) {
        free(c->scratch[i]);
        c->scratch[i] = (double *)malloc(sizeof(double) * (size_t)frames * (size_t)c->channels);
    }
    c->scratch_frames = frames;
}

/* run every stage on the channel slice; returns output frames */
static int64_t chain_slice(orc_chain *c, const double *in, int64_t n, double *out, int32_t c0, int32_t c1)
{
    const int32_t C = c->channels;
    const double *src = in;
    int which = 0;
    if (c->n_stages == 0) {
        for (int64_t f = 0; f < n; f++)
            for (int32_t ch = c0; ch < c1; ch++) out[f * C + ch] = in[f * C + ch];
        return n;
    }
    for (int i = 0; i < c->n_stages; i++) {
        stage_t *s = &c->st[i];
        double *dst = (i == c->n_stages - 1) ? out : c->scratch[which];
        switch (s->desc.kind) {
        case ORC_STAGE_COPY:
        case ORC_STAGE_GAIN: st_gain(s, src, n, dst, C, c0, c1); break;
        case ORC_STAGE_BIQUAD: st_biquad(s, src, n, dst, C, c0, c1); break;
        case ORC_STAGE_FIR: st_fir(s, src, n, dst, C, c0, c1); break;
        case ORC_STAGE_RESAMPLE: n = st_resample(s, src, n, dst, C, c0, c1); break;
        }
        src = dst;
        which ^= 1;
    }
    return n;
}

static void commit_acc(orc_chain *c, int64_t in_frames)
{
    int64_t n = in_frames;
    for (int i = 0; i < c->n_stages; i++) {
        stage_t *s = &c->st[i];
        if (s->desc.kind != ORC_STAGE_RESAMPLE) continue;
        const int64_t tot = s->acc + n * s->desc.up;
        n = tot / s->desc.down;
        s->acc = tot % s->desc.down;
    }
}

int64_t orc_chain_process(orc_chain *c, const double *in, int64_t in_frames, double *out,
                          int64_t out_capacity_frames)
{
    if (in_frames < 0) return -1;
    const int64_t expect = orc_chain_peek_out_frames(c, in_frames);
    if (expect > out_capacity_frames) return -1;
    ensure_scratch(c, in_frames > 0 ? in_frames : 1);
    const int64_t got = chain_slice(c, in, in_frames, out, 0, c->channels);
    commit_acc(c, in_frames);
    return got == expect ? got : -1;
}

typedef struct {
    orc_chain *c;
    const double *in;
    double *out;
    int64_t n;
    int32_t c0, c1;
    int64_t got;
} slice_job;

static void *slice_main(void *p)
{
    slice_job *j = (slice_job *)p;
    j->got = chain_slice(j->c, j->in, j->n, j->out, j->c0, j->c1);
    return NULL;
}

int64_t orc_chain_process_mt(orc_chain *c, const double *in, int64_t in_frames, double *out,
                             int64_t out_capacity_frames, int32_t n_threads)
{
    if (n_threads <= 1 || c->channels < 2) return orc_chain_process(c, in, in_frames, out, out_capacity_frames);
    if (n_threads > c->channels) n_threads = c->channels;
    if (n_threads > 256) n_threads = 256;
    const int64_t expect = orc_chain_peek_out_frames(c, in_frames);
    if (in_frames < 0 || expect > out_capacity_frames) return -1;
    ensure_scratch(c, in_frames > 0 ? in_frames : 1);
    pthread_t th[256];
    slice_job jobs[256];
    for (int t = 0; t < n_threads; t++) {
        jobs[t].c = c;
        jobs[t].in = in;
        jobs[t].out = out;
        jobs[t].n = in_frames;
        jobs[t].c0 = (int32_t)((int64_t)c->channels * t / n_threads);
        jobs[t].c1 = (int32_t)((int64_t)c->channels * (t + 1) / n_threads);
        jobs[t].got = -1;
        pthread_create(&th[t], NULL, slice_main, &jobs[t]);
    }
    int64_t got = expect;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].got != expect) got = -1;
    }
    commit_acc(c, in_frames);
    return got;
}

void orc_mix_sum(const double *const *inputs, int32_t n_inputs, int64_t n_values, double *out)
{
    for (int64_t i = 0; i < n_values; i++) {
        double acc = 0.0;
        for (int32_t l = 0; l < n_inputs; l++) acc += inputs[l][i];
        out[i] = acc;
    }
}

void orc_meter(const double *in, int64_t frames, int32_t channels, double *peak, double *sumsq)
{
    for (int32_t c = 0; c < channels; c++) {
        double p = 0.0, s = 0.0;
        for (int64_t f = 0; f < frames; f++) {
            const double v = in[f * channels + c];
            const double a = v < 0 ? -v : v;
            if (a > p) p = a;
            s += v * v;
        }
        peak[c] = p;
        sumsq[c] = s;
    }
}

static uint64_t splitmix64(uint64_t v)
{
    uint64_t z = v + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

void orc_source_fill(double *out, int64_t first_index, int64_t n_values, uint64_t seed, uint64_t line)
{
    for (int64_t i = 0; i < n_values; i++) {
        const uint64_t z = splitmix64(seed ^ (line << 48) ^ (uint64_t)(first_index + i));
        out[i] = (double)(z >> 40) / 8388608.0 - 1.0;
    }
}

/* ======================================================================== */
/*                                PLUMBING                                  */
/* ======================================================================== */

enum { RES_NIL = 0, RES_EOF = 1, RES_ERR = 2 };

/* syncFitting, internal/fitting/fitting.go:39-42,62-79 */
typedef struct {
    int closed;
    int64_t frames;   /* Message.Signal.Length() */
    double *data;     /* Message.Signal            */
} fitting_t;

static int fit_send(fitting_t *f, double *data, int64_t frames)
{
    if (f->closed) return 0; /* fitting.go:63-65 */
    f->data = data;
    f->frames = frames;
    return 1;
}
static int fit_receive(fitting_t *f, double **data, int64_t *frames)
{
    *data = f->data;
    *frames = f->frames;
    return !f->closed; /* fitting.go:70-75 */
}

enum { EX_SOURCE, EX_PROC, EX_SINK };

typedef struct {
    int kind;
    orc_mock_component *m;
    orc_mock_line *line;
    fitting_t *in, *out;
    double *outbuf; /* pool allocator Float64(): channels*bufferSize, pipe.go:490-492 */
    int64_t buffer_size;
} exec_t;

typedef struct {
    exec_t ex[ORC_MAX_PROCS + 2];
    fitting_t fit[ORC_MAX_PROCS + 1];
    int n_ex;
    int started; /* lineExecutor.started, run.go:24 */
    orc_mock_line *line;
} line_exec_t;

/* mock.Source SourceFunc, mock.go:86-105 */
static int mock_source_func(orc_mock_line *l, double *out, int64_t buffer_size, int64_t *read)
{
    orc_mock_component *m = &l->source;
    if (m->error_on_call) return RES_ERR;
    if (m->samples == l->limit) return RES_EOF;
    int64_t r = buffer_size;
    const int64_t left = l->limit - m->samples;
    if (left < r) r = left;
    for (int64_t i = 0; i < r * l->channels; i++) out[i] = l->value;
    m->messages++;
    m->samples += r;
    *read = r;
    return RES_NIL;
}

/* Source.execute, pipe.go:381-413 (no mutations in the oracle) */
static int source_execute(exec_t *e)
{
    int64_t read = 0;
    const int r = mock_source_func(e->line, e->outbuf, e->buffer_size, &read);
    if (r != RES_NIL) {
        e->out->closed = 1; /* pipe.go:400 */
        return r;
    }
    /* pipe.go:404-406: short read => Slice(0, read) */
    if (!fit_send(e->out, e->outbuf, read)) {
        e->out->closed = 1;
        return RES_EOF;
    }
    return RES_NIL;
}

/* Processor.execute, pipe.go:425-451 with mock.Processor, mock.go:147-154 */
static int proc_execute(exec_t *e)
{
    double *in;
    int64_t frames;
    if (!fit_receive(e->in, &in, &frames)) {
        e->out->closed = 1; /* pipe.go:428 */
        return RES_EOF;
    }
    if (e->m->error_on_call) {
        e->out->closed = 1; /* pipe.go:439 */
        return RES_ERR;
    }
    /* signal.FloatingAsFloating(in, out): copy min(len) frames */
    int64_t n = frames < e->buffer_size ? frames : e->buffer_size;
    memcpy(e->outbuf, in, sizeof(double) * (size_t)(n * e->line->channels));
    e->m->messages++;
    e->m->samples += n;
    if (!fit_send(e->out, e->outbuf, n)) { /* pipe.go:441-449 */
        e->out->closed = 1;
        return RES_EOF;
    }
    return RES_NIL;
}

/* Sink.execute, pipe.go:459-471 with mock.Sink, mock.go:180-189 */
static int sink_execute(exec_t *e)
{
    double *in;
    int64_t frames;
    if (!fit_receive(e->in, &in, &frames)) return RES_EOF;
    if (e->m->error_on_call) return RES_ERR;
    orc_mock_line *l = e->line;
    if (!l->sink_discard && l->sink_values) {
        const int64_t nv = frames * l->channels;
        if (l->sink_values_len + nv <= l->sink_values_capacity) {
            memcpy(l->sink_values + l->sink_values_len, in, sizeof(double) * (size_t)nv);
            l->sink_values_len += nv;
        }
    }
    e->m->messages++;
    e->m->samples += frames;
    return RES_NIL;
}

static int exec_execute(exec_t *e)
{
    switch (e->kind) {
    case EX_SOURCE: return source_execute(e);
    case EX_PROC: return proc_execute(e);
    default: return sink_execute(e);
    }
}

/* lineExecutor.execute, run.go:38-52 */
static int line_execute(line_exec_t *le)
{
    int err = RES_NIL;
    for (int i = 0; i < le->started; i++) {
        err = exec_execute(&le->ex[i]);
        if (err == RES_NIL) continue;
        if (err == RES_EOF) continue; /* keep going to propagate EOF */
        return err;
    }
    return err;
}

/* lineExecutor.flushHook, run.go:54-62: only the first `started` components */
static int line_flush(line_exec_t *le)
{
    int errs = 0;
    for (int i = 0; i < le->started; i++) {
        le->ex[i].m->flushed = 1; /* mock.go:49-52 */
        if (le->ex[i].m->error_on_flush) errs++;
    }
    return errs;
}

/* lineExecutor.startHook, run.go:64-74 */
static int line_start(line_exec_t *le)
{
    for (int i = 0; i < le->n_ex; i++) {
        le->ex[i].m->started = 1; /* mock.go:55-58 sets Started before returning the error */
        if (le->ex[i].m->error_on_start) return 1;
        le->started++;
    }
    return 0;
}

int32_t orc_pipe_run(int64_t buffer_size, int32_t n_lines, orc_mock_line *lines)
{
    if (n_lines < 1 || buffer_size < 1) return ORC_RUN_ERR_BIND;
    /* Line.route, line.go:62-90: allocators run in order; first error aborts */
    for (int l = 0; l < n_lines; l++) {
        if (lines[l].n_procs > ORC_MAX_PROCS) return ORC_RUN_ERR_BIND;
        if (lines[l].source.error_on_make) return ORC_RUN_ERR_BIND;
        for (int p = 0; p < lines[l].n_procs; p++)
            if (lines[l].procs[p].error_on_make) return ORC_RUN_ERR_BIND;
        if (lines[l].sink.error_on_make) return ORC_RUN_ERR_BIND;
    }
    line_exec_t *les = (line_exec_t *)calloc((size_t)n_lines, sizeof(line_exec_t));
    int *alive = (int *)calloc((size_t)n_lines, sizeof(int)); /* index list, run.go:124 */
    int n_alive = n_lines;
    /* route.connect + route.executor, line.go:92-122 */
    for (int l = 0; l < n_lines; l++) {
        line_exec_t *le = &les[l];
        orc_mock_line *ln = &lines[l];
        le->line = ln;
        le->n_ex = ln->n_procs + 2;
        const size_t bufsz = sizeof(double) * (size_t)buffer_size * (size_t)(ln->channels > 0 ? ln->channels : 1);
        for (int i = 0; i < le->n_ex; i++) {
            exec_t *e = &le->ex[i];
            e->line = ln;
            e->buffer_size = buffer_size;
            e->kind = (i == 0) ? EX_SOURCE : (i == le->n_ex - 1 ? EX_SINK : EX_PROC);
            e->m = (i == 0) ? &ln->source : (i == le->n_ex - 1 ? &ln->sink : &ln->procs[i - 1]);
            e->in = (i > 0) ? &le->fit[i - 1] : NULL;
            e->out = (i < le->n_ex - 1) ? &le->fit[i] : NULL;
            e->outbuf = (i < le->n_ex - 1) ? (double *)malloc(bufsz) : NULL;
        }
        alive[l] = l;
    }

    int32_t ret = ORC_RUN_OK;
    /* multiLineExecutor.startHook, run.go:78-99 */
    int start_err = 0;
    for (int l = 0; l < n_lines; l++) {
        if (line_start(&les[l])) {
            start_err = 1;
            break;
        }
    }
    if (start_err) {
        ret = ORC_RUN_ERR_START; /* run.go:201-203 */
        int ferr = 0;
        for (int l = 0; l < n_lines; l++) ferr += line_flush(&les[l]); /* run.go:94 */
        if (ferr) ret |= ORC_RUN_ERR_FLUSH;
        goto done;
    }

    /* run loop, run.go:215-222 with multiLineExecutor.execute, run.go:113-132 */
    int err_exec = RES_NIL;
    int exec_is_flush_err = 0;
    while (err_exec == RES_NIL) {
        int err = RES_NIL;
        for (int i = 0; i < n_alive;) {
            err = line_execute(&les[alive[i]]);
            if (err == RES_NIL) {
                i++;
                continue;
            }
            if (err == RES_EOF) {
                if (line_flush(&les[alive[i]])) { /* run.go:121-123: returned before removal */
                    err = RES_ERR;
                    exec_is_flush_err = 1;
                    break;
                }
                for (int k = i; k + 1 < n_alive; k++) alive[k] = alive[k + 1];
                n_alive--;
                if (n_alive > 0) continue;
            }
            break;
        }
        if (n_alive > 0 && err == RES_EOF) err = RES_NIL; /* loop fell off the end after removals */
        err_exec = err;
    }
    if (err_exec == RES_ERR) ret |= exec_is_flush_err ? (ORC_RUN_ERR_EXEC | ORC_RUN_ERR_FLUSH) : ORC_RUN_ERR_EXEC;
    /* deferred flushHook, run.go:204-213: whatever lines are still registered */
    {
        int ferr = 0;
        for (int i = 0; i < n_alive; i++) ferr += line_flush(&les[alive[i]]);
        if (ferr) ret |= ORC_RUN_ERR_FLUSH;
    }
done:
    for (int l = 0; l < n_lines; l++)
        for (int i = 0; i < les[l].n_ex; i++) free(les[l].ex[i].outbuf);
    free(les);
    free(alive);
    return ret;
}

int32_t orc_mock_source_drain(int64_t buffer_size, orc_mock_line *line)
{
    double *buf = (double *)malloc(sizeof(double) * (size_t)buffer_size * (size_t)(line->channels > 0 ? line->channels : 1));
    int r;
    for (;;) {
        int64_t read = 0;
        r = mock_source_func(line, buf, buffer_size, &read);
        if (r != RES_NIL) break;
    }
    free(buf);
    return r == RES_EOF ? 0 : 1;
}
