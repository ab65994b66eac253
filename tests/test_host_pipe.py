"""Host logic (pipe_b200.pipe, the mirror of the reference API) on CPU with mock components,
checked against the reference's own goldens (cited per test, paths relative to /root/reference)
and against the C oracle's independent restatement of the same plumbing."""
import numpy as np
import pytest

import _oracle as orc
from pipe_b200 import pipe
from pipe_b200.pipe import mock

BUFFER_SIZE = 512  # pipe_test.go:17


def mock_line(limit, channels=1, discard=True, **kw):
    src = mock.Source(limit=limit, channels=channels, **{k[4:]: v for k, v in kw.items() if k.startswith("src_")})
    proc = mock.Processor(**{k[5:]: v for k, v in kw.items() if k.startswith("proc_")})
    snk = mock.Sink(discard=discard, **{k[5:]: v for k, v in kw.items() if k.startswith("sink_")})
    line = pipe.Line(source=src.source(), processors=pipe.processors(proc.processor()), sink=snk.sink())
    return line, src, proc, snk


def test_simple_pipe_async_counts():
    # pipe_test.go:82-106
    line, src, proc, snk = mock_line(862 * BUFFER_SIZE, channels=2)
    pipe.new(BUFFER_SIZE, line).start().wait()
    assert src.counter.messages == 862 and src.counter.samples == 862 * BUFFER_SIZE
    assert snk.counter.messages == 862 and snk.counter.samples == 441344


def test_reset_second_start():
    # pipe_test.go:108-131
    src, snk = mock.Source(limit=862 * BUFFER_SIZE, channels=2), mock.Sink(discard=True)
    p = pipe.new(BUFFER_SIZE, pipe.Line(source=src.source(), sink=snk.sink()))
    p.start().wait()
    assert src.counter.messages == 862
    src.reset()
    p.start().wait()
    assert snk.counter.messages == 2 * 862 and snk.counter.samples == 2 * 862 * BUFFER_SIZE


@pytest.mark.parametrize("limits,expected", [
    ((1040,), ((3, 1040),)), ((1040, 1640), ((3, 1040), (4, 1640))),
    ((3048, 1640, 4096), ((6, 3048), (4, 1640), (8, 4096)))])
def test_run_lines_short_final_buffer(limits, expected):
    # pipe_test.go:330-436, and the oracle's restatement must agree
    ms = [mock_line(n) for n in limits]
    pipe.run(BUFFER_SIZE, *[m[0] for m in ms])
    ret, ols = orc.pipe_run(BUFFER_SIZE, [orc.mock_line(limit=n, channels=1) for n in limits])
    assert ret == orc.RUN_OK
    for (_, src, proc, snk), (msgs, frames), ol in zip(ms, expected, ols):
        for comp in (src, proc, snk):
            assert (comp.counter.messages, comp.counter.samples, comp.flushed) == (msgs, frames, True)
        assert (ol.sink.messages, ol.sink.samples) == (msgs, frames)


def _flags(src, proc, snk):
    return (src.started, proc.started, snk.started, src.flushed, proc.flushed, snk.flushed)


@pytest.mark.parametrize("flush_err", [False, True])
def test_two_lines_processor_start_error(flush_err):
    # pipe_test.go:228-306
    mock_error = Exception("mock error")
    l1 = mock_line(1040, discard=False, **({"src_error_on_flush": mock_error} if flush_err else {}))
    l2 = mock_line(1040, discard=False, proc_error_on_start=mock_error)
    with pytest.raises(pipe.ErrorStart) as e:
        pipe.run(BUFFER_SIZE, l1[0], l2[0])
    assert pipe.errors_is(e.value, mock_error)
    assert _flags(*l1[1:]) == (True, True, True, True, True, True)
    assert _flags(*l2[1:]) == (True, True, False, True, False, False)


def test_single_line_processor_start_error():
    # pipe_test.go:307-329
    l = mock_line(1040, discard=False, proc_error_on_start=Exception("mock error"))
    with pytest.raises(pipe.ErrorStart):
        pipe.run(BUFFER_SIZE, l[0])
    assert _flags(*l[1:]) == (True, True, False, True, False, False)


def test_single_processor_error_surfaces_and_everything_flushes():
    # pipe_test.go:437-457
    mock_error = Exception("mock error")
    l = mock_line(1040, proc_error_on_call=mock_error)
    with pytest.raises(pipe.ErrorRun) as e:
        pipe.run(BUFFER_SIZE, l[0])
    assert pipe.errors_is(e.value, mock_error)
    assert (l[1].flushed, l[2].flushed, l[3].flushed) == (True, True, True)


@pytest.mark.parametrize("which", ["src", "proc", "sink"])
def test_binding_errors(which):
    # pipe_test.go:21-80
    err = Exception("binding error")
    l = mock_line(0, **{f"{which}_error_on_make": err})
    with pytest.raises(Exception) as e:
        pipe.new(BUFFER_SIZE, l[0])
    assert pipe.errors_is(e.value, err)


def test_zero_value_source_runs_clean():
    # line_test.go:11-19
    src, snk = mock.Source(), mock.Sink()
    pipe.run(BUFFER_SIZE, pipe.Line(source=src.source(), sink=snk.sink()))
    assert snk.counter.messages == 0


@pytest.mark.parametrize("limit,bs,calls", [(11, 5, 3), (2500, 5, 500)])
def test_mock_source_call_counts(limit, bs, calls):
    # mock/mock_test.go:69-92
    m = mock.Source(limit=limit, channels=2, value=1.0, sample_rate=44100)
    src = m.source()(bs)
    buf = np.empty((bs, 2))
    with pytest.raises(pipe.EOF):
        while True:
            src.source_func(buf)
    assert m.counter.messages == calls and m.counter.samples == limit


@pytest.mark.parametrize("vals", [[1, 1, 1, 1], [1, 1, 1, 1, 2, 2, 2, 2]])
def test_mock_processor_and_sink_values(vals):
    # mock/mock_test.go:133-146,185-202
    x = np.asarray(vals, dtype=np.float64).reshape(-1, 1)
    proc = mock.Processor().processor()(0, pipe.SignalProperties())
    out = np.empty_like(x)
    assert proc.process_func(x, out) == len(x) and np.array_equal(out, x)
    m = mock.Sink(discard=False)
    m.sink()(0, pipe.SignalProperties(channels=1)).sink_func(x)
    assert np.array_equal(m.values, x)


def test_async_error_cancels_and_surfaces():
    mock_error = Exception("mock error")
    l = mock_line(100 * BUFFER_SIZE, proc_error_on_call=mock_error)
    with pytest.raises(Exception) as e:
        pipe.new(BUFFER_SIZE, l[0]).start().wait()
    assert pipe.errors_is(e.value, mock_error)
    assert l[1].flushed and l[3].flushed
