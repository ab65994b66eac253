"""The reference's plugin boundary end to end: mock.Source -> GPU Processor (one fused chain behind
pipe.Processor) -> mock.Sink, driven by the host mirror of pipe.Run / pipe.New, against the oracle."""
import numpy as np
import pytest

import _oracle as orc
from pipe_b200 import design, gpu, pipe
from pipe_b200.pipe import mock

pytestmark = pytest.mark.gpu
BUFFER_SIZE = 512


def test_config1_mock_source_gpu_passthrough_mock_sink():
    # configs[0] and pipe_test.go:82-106: 2 ch float64, 512-frame buffers, 862 messages / 441,344 frames
    src, snk = mock.Source(limit=862 * BUFFER_SIZE, channels=2, value=0.5), mock.Sink(discard=True)
    cp = gpu.ChainProcessor([gpu.copy()], dtype=np.float64)
    pipe.run(BUFFER_SIZE, pipe.Line(source=src.source(), processors=pipe.processors(cp.processor()), sink=snk.sink()))
    assert (src.counter.messages, src.counter.samples) == (862, 441344)
    assert (cp.messages, cp.samples) == (862, 441344)
    assert (snk.counter.messages, snk.counter.samples) == (862, 441344)


@pytest.mark.parametrize("limit,msgs", [(1040, 3), (1640, 4), (3048, 6), (4096, 8)])
def test_short_final_buffer_counts_and_values(limit, msgs):
    # pipe_test.go:337,363,394,404 with the GPU Processor in place of mock.Processor
    src, snk = mock.Source(limit=limit, channels=1, value=1.0), mock.Sink(discard=False)
    cp = gpu.ChainProcessor([gpu.copy()], dtype=np.float64)
    pipe.run(BUFFER_SIZE, pipe.Line(source=src.source(), processors=pipe.processors(cp.processor()), sink=snk.sink()))
    assert (cp.messages, cp.samples) == (msgs, limit)
    assert (snk.counter.messages, snk.counter.samples) == (msgs, limit)
    assert np.array_equal(snk.values, np.ones((limit, 1)))


@pytest.mark.parametrize("mode", ["run", "async"])
def test_chain4_through_the_pipe_matches_the_oracle(mode):
    ch, bs, limit = 16, 1024, 3 * 1024 + 100
    stages = design.config_stages("chain4")

    def fill(out, first_frame):
        out[:] = orc.source_fill(first_frame * ch, out.size).reshape(out.shape)

    src = mock.Source(limit=limit, channels=ch, sample_rate=48000.0, fill=fill)
    snk = mock.Sink(discard=False)
    line = pipe.Line(source=src.source(), processors=pipe.processors(gpu.chain(stages, dtype=np.float32)), sink=snk.sink())
    if mode == "run":
        pipe.run(bs, line, dtype=np.float32)
    else:
        pipe.new(bs, line, dtype=np.float32).start().wait()
    cpu = orc.Chain(ch, stages)
    x = orc.source_fill(0, limit * ch).reshape(limit, ch)
    refs = [cpu.process(x[i:i + bs]) for i in range(0, limit, bs)]
    ref = np.concatenate(refs)
    assert snk.counter.messages == 4
    assert snk.counter.samples == len(ref) == (limit * 147) // 160        # bit-exact frame bookkeeping
    y = snk.values
    err = np.abs(y - ref).max(axis=0)
    assert (err <= 1e-6 * np.abs(ref).max(axis=0)).all()


def test_output_signal_properties_are_threaded_to_the_sink():
    seen = {}

    def sink_alloc(buffer_size, props):
        seen["props"] = props
        return pipe.Sink(sink_func=lambda b: None)

    src = mock.Source(limit=10, channels=4, sample_rate=48000.0)
    pipe.run(16, pipe.Line(source=src.source(), processors=pipe.processors(gpu.chain(design.config_stages("chain4"))),
                           sink=sink_alloc))
    assert seen["props"].channels == 4
    assert abs(seen["props"].sample_rate - 44100.0) < 1e-9   # line.go:75


def test_gpu_allocator_error_aborts_binding():
    bad = gpu.chain([gpu.resample(3, 2, np.ones(6))])
    src, snk = mock.Source(limit=10, channels=1), mock.Sink()
    with pytest.raises(Exception) as e:
        pipe.new(16, pipe.Line(source=src.source(), processors=pipe.processors(bad), sink=snk.sink()))
    assert "processor" in str(e.value) and "UNSUPPORTED" in str(e.value)   # line.go:72-74


def test_pipe_with_gpu_chain_restarts_from_zero_state():
    # TestReset (pipe_test.go:107-130): Start / Wait, source.Reset(), Start again -- the chain handle stays bound across the
    # runs (FlushFunc must not free it) and StartFunc zeroes the carried state, so both runs give the oracle's first-run output
    ch, bs, limit = 16, 512, 3 * 512 + 77
    stages = design.config_stages("chain4")

    def fill(out, first_frame):
        out[:] = orc.source_fill(first_frame * ch, out.size).reshape(out.shape)

    src = mock.Source(limit=limit, channels=ch, sample_rate=48000.0, fill=fill)
    snk = mock.Sink(discard=False)
    cp = gpu.ChainProcessor(stages, dtype=np.float32)
    p = pipe.new(bs, pipe.Line(source=src.source(), processors=pipe.processors(cp.processor()), sink=snk.sink()), dtype=np.float32)
    x = orc.source_fill(0, limit * ch).reshape(limit, ch)
    cpu = orc.Chain(ch, stages)
    ref = np.concatenate([cpu.process(x[i:i + bs]) for i in range(0, limit, bs)])
    p.start().wait()
    first = np.array(snk.values)
    src.reset()
    p.start().wait()
    both = np.array(snk.values)
    assert cp.starts == 2 and cp.chain is not None
    assert len(both) == 2 * len(first) == 2 * len(ref)
    for run in (first, both[len(first):]):
        assert (np.abs(run - ref).max(axis=0) <= 1e-6 * np.abs(ref).max(axis=0)).all()
    cp.close()
    assert cp.chain is None


def test_float64_pipe_reaches_the_tcgen05_kernel_through_compute_dtype():
    # the reference always allocates float64 buffers (pipe.go:394,437); compute_dtype=float32 converts in the marshalling copy,
    # which is how a float64 pipe reaches K2: 128 channels x 1600-frame buffers through pipe.Line / gpu.chain, last_path == 2
    ch, bs, limit = 128, 1600, 4 * 1600
    stages = design.config_stages("chain4")

    def fill(out, first_frame):
        out[:] = orc.source_fill(first_frame * ch, out.size).reshape(out.shape)

    src = mock.Source(limit=limit, channels=ch, sample_rate=48000.0, fill=fill)
    snk = mock.Sink(discard=False)
    cp = gpu.ChainProcessor(stages, dtype=np.float64, compute_dtype=np.float32)
    pipe.run(bs, pipe.Line(source=src.source(), processors=pipe.processors(cp.processor()), sink=snk.sink()), dtype=np.float64)
    assert cp.chain.last_path()[0] == 2
    x = orc.source_fill(0, limit * ch).reshape(limit, ch)
    cpu = orc.Chain(ch, stages)
    ref = np.concatenate([cpu.process(x[i:i + bs]) for i in range(0, limit, bs)])
    y = snk.values
    assert y.dtype == np.float64 and len(y) == len(ref) == (limit * 147) // 160
    for i in range(0, len(ref), len(ref) // 4):
        seg, rseg = y[i:i + len(ref) // 4], ref[i:i + len(ref) // 4]
        assert (np.abs(seg - rseg).max(axis=0) <= 1e-6 * np.abs(rseg).max(axis=0)).all()
