"""Host-side mirror of the reference API for the Processor hot path.

The reference is Go and this image has no Go toolchain, so the host logic that
sits above the C-ABI is restated here with the reference's names, argument
meaning and error behaviour (paths relative to the reference repo):

    Line{Source, Processors, Sink}         line.go:14-19
    Source/Processor/Sink + *Func types    pipe.go:32-86
    *AllocatorFunc                          line.go:21-35
    SignalProperties                        line.go:38-41
    run(buffer_size, *lines)   == pipe.Run  pipe.go:90-103, run.go:200-224
    new(buffer_size, *lines)   == pipe.New  pipe.go:107-126
    Pipe.start / wait                       pipe.go:197-257, run.go:173-196
    sync / async fittings                   internal/fitting/fitting.go:50-104

Buffers (signal.Floating) are numpy arrays of shape (frames, channels), the
frame-major interleaved layout of the reference; `len(buf)` is Length().
Only what the per-buffer path needs is mirrored: no AddLine / InsertProcessor
(control plane, SURVEY.md section 2 "OUT OF SCOPE").
"""
from __future__ import annotations

import queue
import threading
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np


class EOF(Exception):
    """io.EOF: the clean end-of-stream sentinel (run.go:44,120,191,218)."""


class ErrorRun(Exception):
    """error.go:11-35: execution and/or flush failed after a successful start."""

    def __init__(self, err_exec: Optional[BaseException], err_flush: Optional[BaseException]):
        self.err_exec, self.err_flush = err_exec, err_flush
        super().__init__(f"execute error: {err_exec!r}, flush error: {err_flush!r}")

    def is_(self, err: BaseException) -> bool:  # error.go:28-36 ErrorRun.Is
        return any(_chain_has(e, err) for e in (self.err_exec, self.err_flush) if e is not None)


class ErrorStart(Exception):
    """'error starting: %w' (run.go:202)."""


def _chain_has(e: BaseException, target: BaseException) -> bool:
    while e is not None:
        if e is target:
            return True
        if isinstance(e, ErrorList) and any(_chain_has(x, target) for x in e.errors):
            return True
        e = e.__cause__
    return False


def errors_is(err: Optional[BaseException], target: BaseException) -> bool:
    if err is None:
        return False
    if isinstance(err, ErrorRun):
        return err.is_(target)
    return _chain_has(err, target)


class ErrorList(Exception):
    """execErrors (error.go:38-57)."""

    def __init__(self, errors):
        self.errors = list(errors)
        super().__init__(",".join(str(e) for e in self.errors))


def _wrap(msg: str, cause: BaseException) -> Exception:
    e = Exception(f"{msg}: {cause}")
    e.__cause__ = cause
    return e


@dataclass
class SignalProperties:  # line.go:38-41
    sample_rate: float = 0.0
    channels: int = 0


SourceFunc = Callable[[np.ndarray], int]                 # fills out, returns frames read; raises EOF
ProcessFunc = Callable[[np.ndarray, np.ndarray], int]    # (in, out) -> frames processed
SinkFunc = Callable[[np.ndarray], None]
HookFunc = Optional[Callable[[], None]]


@dataclass
class Source:  # pipe.go:35-43
    source_func: SourceFunc = None
    start_func: HookFunc = None
    flush_func: HookFunc = None
    props: SignalProperties = field(default_factory=SignalProperties)


@dataclass
class Processor:  # pipe.go:52-60
    process_func: ProcessFunc = None
    start_func: HookFunc = None
    flush_func: HookFunc = None
    props: SignalProperties = field(default_factory=SignalProperties)


@dataclass
class Sink:  # pipe.go:69-76
    sink_func: SinkFunc = None
    start_func: HookFunc = None
    flush_func: HookFunc = None


SourceAllocatorFunc = Callable[[int], Source]                            # (buffer_size)
ProcessorAllocatorFunc = Callable[[int, SignalProperties], Processor]    # (buffer_size, input props)
SinkAllocatorFunc = Callable[[int, SignalProperties], Sink]


@dataclass
class Line:  # line.go:14-19
    source: SourceAllocatorFunc = None
    processors: Sequence[ProcessorAllocatorFunc] = ()
    sink: SinkAllocatorFunc = None
    sync: bool = False  # a mutable Context on the Line means single-goroutine execution (pipe.go:135-139)


def processors(*procs: ProcessorAllocatorFunc):  # pipe.go:368-370
    return list(procs)


# ------------------------------------------------------------------ fittings --

@dataclass
class Message:  # fitting.go:12-15
    signal: np.ndarray = None


class SyncFitting:  # fitting.go:39-42,62-79
    def __init__(self):
        self.closed, self.message = False, None

    def send(self, m: Message) -> bool:
        if self.closed:
            return False
        self.message = m
        return True

    def receive(self):
        return self.message, not self.closed

    def close(self):
        self.closed = True


class AsyncFitting:  # fitting.go:44-47,56-60,81-104: chan Message, cap 1
    _CLOSED = object()

    def __init__(self, cancel: threading.Event):
        self.q: queue.Queue = queue.Queue(maxsize=1)
        self.cancel = cancel

    def send(self, m: Message) -> bool:
        while not self.cancel.is_set():
            try:
                self.q.put(m, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def receive(self):
        while not self.cancel.is_set():
            try:
                m = self.q.get(timeout=0.05)
            except queue.Empty:
                continue
            if m is self._CLOSED:
                self.q.put(m)  # stay closed for any later receive
                return None, False
            return m, True
        return None, False

    def close(self):
        while True:
            try:
                self.q.put(self._CLOSED, timeout=0.05)
                return
            except queue.Full:
                if self.cancel.is_set():
                    return


# ----------------------------------------------------------------- executors --

class _SourceExec:
    def __init__(self, src: Source, buffer_size: int, dtype):
        self.c, self.buffer_size, self.dtype, self.out = src, buffer_size, dtype, None

    def start_hook(self):
        if self.c.start_func:
            self.c.start_func()

    def flush_hook(self):
        if self.c.flush_func:
            self.c.flush_func()

    def execute(self):  # Source.execute, pipe.go:381-413
        output = np.empty((self.buffer_size, max(self.c.props.channels, 0)), dtype=self.dtype)
        try:
            read = self.c.source_func(output)
        except BaseException:
            self.out.close()
            raise
        if read != len(output):
            output = output[:read]
        if not self.out.send(Message(output)):
            self.out.close()
            raise EOF()


class _ProcExec:
    def __init__(self, proc: Processor, buffer_size: int, dtype):
        self.c, self.buffer_size, self.dtype, self.inp, self.out = proc, buffer_size, dtype, None, None

    start_hook = _SourceExec.start_hook
    flush_hook = _SourceExec.flush_hook

    def execute(self):  # Processor.execute, pipe.go:425-451
        m, ok = self.inp.receive()
        if not ok:
            self.out.close()
            raise EOF()
        output = np.empty((self.buffer_size, self.c.props.channels), dtype=self.dtype)
        try:
            processed = self.c.process_func(m.signal, output)
        except BaseException:
            self.out.close()
            raise
        if processed != self.buffer_size:
            output = output[:processed]
        if not self.out.send(Message(output)):
            self.out.close()
            raise EOF()


class _SinkExec:
    def __init__(self, sink: Sink):
        self.c, self.inp = sink, None

    start_hook = _SourceExec.start_hook
    flush_hook = _SourceExec.flush_hook

    def execute(self):  # Sink.execute, pipe.go:459-471
        m, ok = self.inp.receive()
        if not ok:
            raise EOF()
        self.c.sink_func(m.signal)


class _LineExecutor:  # run.go:20-74
    def __init__(self, executors):
        self.executors, self.started = executors, 0

    def execute(self):
        err = None
        for i in range(self.started):
            try:
                self.executors[i].execute()
                err = None
            except EOF as e:
                err = e  # continue execution to propagate EOF
        if err is not None:
            raise err

    def flush_hook(self):
        errs = []
        for i in range(self.started):
            try:
                self.executors[i].flush_hook()
            except Exception as e:
                errs.append(e)
        if errs:
            raise ErrorList(errs)

    def start_hook(self):
        for e in self.executors:
            try:
                e.start_hook()
            except Exception as ex:
                raise ErrorList([ex])
            self.started += 1


def _bind(line: Line, buffer_size: int, dtype):  # Line.route, line.go:62-90
    try:
        src = line.source(buffer_size)
    except Exception as e:
        raise _wrap("source", e)
    prev = src.props
    procs = []
    for alloc in line.processors:
        try:
            p = alloc(buffer_size, prev)
        except Exception as e:
            raise _wrap("processor", e)
        prev = p.props
        procs.append(p)
    try:
        snk = line.sink(buffer_size, prev)
    except Exception as e:
        raise _wrap("sink", e)
    execs = [_SourceExec(src, buffer_size, dtype)] + [_ProcExec(p, buffer_size, dtype) for p in procs] + [_SinkExec(snk)]
    return execs


def _connect(execs, make_fitting):  # route.connect, line.go:92-104
    for a, b in zip(execs[:-1], execs[1:]):
        f = make_fitting()
        a.out, b.inp = f, f


def run(buffer_size: int, *lines: Line, dtype=np.float64) -> None:
    """pipe.Run (pipe.go:90-103): every line in the calling thread, one buffer per line per iteration."""
    les = []
    for l in lines:
        execs = _bind(l, buffer_size, dtype)
        _connect(execs, SyncFitting)
        les.append(_LineExecutor(execs))
    # multiLineExecutor.startHook, run.go:78-99
    start_err = None
    for le in les:
        try:
            le.start_hook()
        except ErrorList as e:
            start_err = e
            break
    if start_err is not None:
        err = ErrorStart(f"error starting lines: {start_err}")
        err.__cause__ = start_err
        flush_errs = []
        for le in les:
            try:
                le.flush_hook()
            except ErrorList as fe:
                flush_errs.append(fe)
        if flush_errs:
            err2 = ErrorStart(f"error flushing lines: {ErrorList(flush_errs)} during start error: {err}")
            err2.__cause__ = ErrorList([ErrorList(flush_errs), start_err])
            raise err2
        raise err
    # run loop, run.go:215-222 over multiLineExecutor.execute, run.go:113-132
    alive = list(les)
    err_exec = None
    while err_exec is None and alive:
        i = 0
        while i < len(alive):
            try:
                alive[i].execute()
                i += 1
            except EOF:
                try:
                    alive[i].flush_hook()
                except ErrorList as fe:
                    err_exec = fe  # returned before the line is removed (run.go:121-123)
                    break
                alive.pop(i)
            except Exception as e:
                err_exec = e
                break
    if err_exec is not None:
        err_exec = _wrap("error running", err_exec)
    err_flush = None
    flush_errs = []
    for le in alive:  # deferred flushHook, run.go:204-213
        try:
            le.flush_hook()
        except ErrorList as fe:
            flush_errs.append(fe)
    if flush_errs:
        err_flush = _wrap("error flushing", ErrorList(flush_errs))
    if err_exec is None and err_flush is None:
        return
    raise ErrorRun(err_exec, err_flush)


class Pipe:
    """pipe.New + Start + Wait for immutable (async) lines: one thread per component,
    cap-1 queues between them (pipe.go:107-126,172-214; run.go:173-196)."""

    def __init__(self, buffer_size: int, lines: Sequence[Line], dtype=np.float64):
        if not lines:
            raise ValueError("pipe without lines")  # pipe.go:108-110 panics
        self.buffer_size, self.dtype = buffer_size, dtype
        self.routes = [_bind(l, buffer_size, dtype) for l in lines]
        self._threads, self._errors, self._cancel = [], [], None

    def start(self) -> "Pipe":
        self._cancel = threading.Event()
        self._errors, self._threads = [], []
        lock = threading.Lock()
        for execs in self.routes:
            _connect(execs, lambda: AsyncFitting(self._cancel))
            for ex in execs:
                t = threading.Thread(target=self._component_main, args=(ex, lock), daemon=True)
                self._threads.append(t)
        for t in self._threads:
            t.start()
        return self

    def _component_main(self, ex, lock):  # start(), run.go:173-196
        def fail(e):
            with lock:
                self._errors.append(e)
            self._cancel.set()  # first error cancels the context (pipe.go:230-237)
        try:
            ex.start_hook()
        except Exception as e:
            fail(_wrap("error starting", e))
            return
        try:
            while True:
                ex.execute()
        except EOF:
            pass
        except Exception as e:
            fail(_wrap("error running", e))
        try:
            ex.flush_hook()
        except Exception as e:
            fail(_wrap("error flushing", e))

    def wait(self) -> None:  # pipe.Wait, pipe.go:250-257: first error wins
        for t in self._threads:
            t.join()
        if self._errors:
            raise self._errors[0]


def new(buffer_size: int, *lines: Line, dtype=np.float64) -> Pipe:
    return Pipe(buffer_size, lines, dtype=dtype)


# -------------------------------------------------------------------- mocks --

class mock:
    """mock package (mock/mock.go): the reference's only components."""

    @dataclass
    class Counter:  # mock.go:17-21,43-46
        messages: int = 0
        samples: int = 0
        values: Optional[np.ndarray] = None

        def advance(self, size: int):
            self.messages += 1
            self.samples += size

    @dataclass
    class Source:  # mock.go:61-109
        limit: int = 0
        value: float = 0.0
        channels: int = 0
        sample_rate: float = 0.0
        error_on_call: Optional[Exception] = None
        error_on_make: Optional[Exception] = None
        error_on_start: Optional[Exception] = None
        error_on_flush: Optional[Exception] = None
        fill: Optional[Callable[[np.ndarray, int], None]] = None  # (out[:read], first_frame); default: constant value
        started: bool = False
        flushed: bool = False
        counter: "mock.Counter" = None

        def __post_init__(self):
            self.counter = mock.Counter()

        def reset(self):  # mock.go:112-118
            self.counter = mock.Counter()

        def source(self) -> SourceAllocatorFunc:
            def alloc(buffer_size: int) -> Source:
                if self.error_on_make:
                    raise self.error_on_make

                def source_func(out: np.ndarray) -> int:
                    if self.error_on_call:
                        raise self.error_on_call
                    if self.counter.samples == self.limit:
                        raise EOF()
                    read = min(len(out), self.limit - self.counter.samples)
                    if self.fill is not None:
                        self.fill(out[:read], self.counter.samples)
                    else:
                        out[:read] = self.value
                    self.counter.advance(read)
                    return read
                return Source(source_func, _hook(self, "started", "error_on_start"), _hook(self, "flushed", "error_on_flush"),
                              SignalProperties(self.sample_rate, self.channels))
            return alloc

    @dataclass
    class Processor:  # mock.go:130-157
        error_on_call: Optional[Exception] = None
        error_on_make: Optional[Exception] = None
        error_on_start: Optional[Exception] = None
        error_on_flush: Optional[Exception] = None
        started: bool = False
        flushed: bool = False
        counter: "mock.Counter" = None

        def __post_init__(self):
            self.counter = mock.Counter()

        def processor(self) -> ProcessorAllocatorFunc:
            def alloc(buffer_size: int, props: SignalProperties) -> Processor:
                if self.error_on_make:
                    raise self.error_on_make

                def process_func(inp: np.ndarray, out: np.ndarray) -> int:
                    if self.error_on_call:
                        raise self.error_on_call
                    n = min(len(inp), len(out))  # signal.FloatingAsFloating
                    out[:n] = inp[:n]
                    self.counter.advance(n)
                    return n
                return Processor(process_func, _hook(self, "started", "error_on_start"),
                                 _hook(self, "flushed", "error_on_flush"), props)
            return alloc

    @dataclass
    class Sink:  # mock.go:160-192
        discard: bool = False
        error_on_call: Optional[Exception] = None
        error_on_make: Optional[Exception] = None
        error_on_start: Optional[Exception] = None
        error_on_flush: Optional[Exception] = None
        started: bool = False
        flushed: bool = False
        counter: "mock.Counter" = None

        def __post_init__(self):
            self.counter = mock.Counter()
            self._chunks = []

        @property
        def values(self) -> np.ndarray:
            return np.concatenate(self._chunks) if self._chunks else np.zeros((0, 0))

        def sink(self) -> SinkAllocatorFunc:
            def alloc(buffer_size: int, props: SignalProperties) -> Sink:
                if self.error_on_make:
                    raise self.error_on_make

                def sink_func(inp: np.ndarray) -> None:
                    if self.error_on_call:
                        raise self.error_on_call
                    if not self.discard:
                        self._chunks.append(np.array(inp, copy=True))
                    self.counter.advance(len(inp))
                return Sink(sink_func, _hook(self, "started", "error_on_start"), _hook(self, "flushed", "error_on_flush"))
            return alloc


def _hook(obj, flag: str, err_attr: str):
    def hook():
        setattr(obj, flag, True)  # mock.go:49-58: the flag is set before the error is returned
        err = getattr(obj, err_attr)
        if err:
            raise err
    return hook
