"""Pin the oracle's plumbing restatement against the reference's own goldens.

Every expected number below is copied from a reference test assertion (cited
per test, paths relative to /root/reference); none is produced by this repo.
"""
import numpy as np
import pytest

import _oracle as orc

BUFFER_SIZE = 512  # pipe_test.go:17


def test_simple_pipe_counts():
    # pipe_test.go:82-106: 2 ch, Limit 862*512 -> 862 messages, 441,344 frames
    ret, (l,) = orc.pipe_run(BUFFER_SIZE, [orc.mock_line(limit=862 * BUFFER_SIZE, channels=2)])
    assert ret == orc.RUN_OK
    assert l.source.messages == 862
    assert l.source.samples == 862 * BUFFER_SIZE == 441344
    assert l.sink.messages == 862 and l.sink.samples == 441344


def test_reset_doubles_sink_totals():
    # pipe_test.go:108-131: second run after source.Reset() -> sink sees 2*862 messages
    line = orc.mock_line(limit=862 * BUFFER_SIZE, channels=2, n_procs=0)
    ret, (l,) = orc.pipe_run(BUFFER_SIZE, [line])
    assert ret == orc.RUN_OK and l.sink.messages == 862
    l.source.messages = 0  # mock.go:112-118 Reset(): Counter = Counter{}
    l.source.samples = 0
    ret, (l2,) = orc.pipe_run(BUFFER_SIZE, [l])
    assert ret == orc.RUN_OK
    assert l2.sink.messages == 2 * 862
    assert l2.sink.samples == 2 * 862 * BUFFER_SIZE


def test_multiple_lines_shared_context():
    # pipe_test.go:156-189
    ret, ls = orc.pipe_run(BUFFER_SIZE, [orc.mock_line(limit=862 * BUFFER_SIZE, channels=2, n_procs=0)
                                         for _ in range(2)])
    assert ret == orc.RUN_OK
    for l in ls:
        assert l.source.messages == 862 and l.source.samples == 862 * BUFFER_SIZE


@pytest.mark.parametrize("limits,expected", [
    ((1040,), ((3, 1040),)),                               # pipe_test.go:330-349
    ((1040, 1640), ((3, 1040), (4, 1640))),                # pipe_test.go:350-385
    ((3048, 1640, 4096), ((6, 3048), (4, 1640), (8, 4096))),  # pipe_test.go:386-436
])
def test_lines_short_final_buffer(limits, expected):
    ret, ls = orc.pipe_run(BUFFER_SIZE, [orc.mock_line(limit=n, channels=1) for n in limits])
    assert ret == orc.RUN_OK
    for l, (msgs, frames) in zip(ls, expected):
        # assertLine, pipe_test.go:203-210: source, processor and sink agree
        for comp in (l.source, l.procs[0], l.sink):
            assert comp.messages == msgs
            assert comp.samples == frames
            assert comp.flushed == 1


def _flags(l):
    return (l.source.started, l.procs[0].started, l.sink.started,
            l.source.flushed, l.procs[0].flushed, l.sink.flushed)


def test_two_lines_processor_start_error_source_flush_error():
    # pipe_test.go:228-268
    ret, (l1, l2) = orc.pipe_run(BUFFER_SIZE, [
        orc.mock_line(limit=1040, channels=1, discard=False, source_error_on_flush=1),
        orc.mock_line(limit=1040, channels=1, discard=False, proc0_error_on_start=1),
    ])
    assert ret & orc.RUN_ERR_START and ret & orc.RUN_ERR_FLUSH
    assert _flags(l1) == (1, 1, 1, 1, 1, 1)
    assert _flags(l2) == (1, 1, 0, 1, 0, 0)


def test_two_lines_processor_start_error():
    # pipe_test.go:269-306
    ret, (l1, l2) = orc.pipe_run(BUFFER_SIZE, [
        orc.mock_line(limit=1040, channels=1, discard=False),
        orc.mock_line(limit=1040, channels=1, discard=False, proc0_error_on_start=1),
    ])
    assert ret == orc.RUN_ERR_START
    assert _flags(l1) == (1, 1, 1, 1, 1, 1)
    assert _flags(l2) == (1, 1, 0, 1, 0, 0)
    assert l1.source.messages == 0  # nothing executed


def test_single_line_processor_start_error():
    # pipe_test.go:307-329
    ret, (l,) = orc.pipe_run(BUFFER_SIZE, [orc.mock_line(limit=1040, channels=1, discard=False,
                                                         proc0_error_on_start=1)])
    assert ret == orc.RUN_ERR_START
    assert _flags(l) == (1, 1, 0, 1, 0, 0)


def test_single_processor_error_still_flushes():
    # pipe_test.go:437-457: errors.Is(err, mockError) and all three flushed
    ret, (l,) = orc.pipe_run(BUFFER_SIZE, [orc.mock_line(limit=1040, channels=1, proc0_error_on_call=1)])
    assert ret == orc.RUN_ERR_EXEC
    assert (l.source.flushed, l.procs[0].flushed, l.sink.flushed) == (1, 1, 1)


@pytest.mark.parametrize("knob", ["source_error_on_make", "proc0_error_on_make", "sink_error_on_make"])
def test_binding_errors(knob):
    # pipe_test.go:21-80: allocator errors abort before anything starts
    ret, (l,) = orc.pipe_run(BUFFER_SIZE, [orc.mock_line(limit=0, channels=1, **{knob: 1})])
    assert ret == orc.RUN_ERR_BIND
    assert (l.source.started, l.sink.started) == (0, 0)


def test_zero_limit_line_runs_clean():
    # line_test.go:11-19
    ret, (l,) = orc.pipe_run(BUFFER_SIZE, [orc.mock_line(limit=0, channels=0, n_procs=0)])
    assert ret == orc.RUN_OK and l.sink.messages == 0


@pytest.mark.parametrize("limit,buffer_size,calls", [(11, 5, 3), (2500, 5, 500)])
def test_mock_source_call_counts(limit, buffer_size, calls):
    # mock/mock_test.go:69-92
    line = orc.mock_line(limit=limit, channels=2, value=1.0)
    assert orc.mock_source_drain(buffer_size, line) == 0
    assert line.source.messages == calls
    assert line.source.samples == limit


def test_mock_source_error_on_call():
    # mock/mock_test.go:93-98
    line = orc.mock_line(limit=10, channels=1, source_error_on_call=1)
    assert orc.mock_source_drain(5, line) == 1
    assert line.source.messages == 0


@pytest.mark.parametrize("values", [[1, 1, 1, 1], [1, 1, 1, 1, 2, 2, 2, 2]])
def test_passthrough_values(values):
    # mock/mock_test.go:133-146 (Processor) and :185-202 (Sink): identity on
    # a 1-channel buffer (the case labelled "2 channels" allocates Channels: 1).
    x = np.asarray(values, dtype=np.float64).reshape(-1, 1)
    y = orc.Chain(1, [{"kind": "copy"}]).process(x)
    assert np.array_equal(y, x)
    # through the plumbing with a constant source, captured by a non-discard sink
    ret, (l,) = orc.pipe_run(4, [orc.mock_line(limit=len(values), channels=1, value=float(values[0]),
                                                discard=False, capture=len(values))])
    assert ret == orc.RUN_OK
    assert l.sink_values_len == len(values)
    assert np.array_equal(l._buf[:len(values)], np.full(len(values), float(values[0])))
