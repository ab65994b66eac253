// pipe.hpp -- C++ host mirror of the reference API for the Processor hot path, above the C-ABI (include/pipe_b200.h).
//
// The reference (pipelined/pipe) is Go and this image has no Go toolchain, so the host side that a Go maintainer would keep
// (Line / Source / Processor / Sink, the allocator funcs, pipe.Run, pipe.New + Start + Wait, the mock components) is restated
// here in C++17 with the reference's names, argument meaning and error behaviour.  Header-only; a program links against
// libpipe_b200.so only if it uses pipe::gpu::Chain.  Citations are paths in the reference repository:
//
//   Line{Source, Processors, Sink}            line.go:14-19
//   Source / Processor / Sink + *Func types    pipe.go:32-86
//   *AllocatorFunc                             line.go:21-35
//   SignalProperties                           line.go:38-41
//   Run(bufferSize, lines...)                  pipe.go:90-103, run.go:78-132,200-224   (every line in the calling thread)
//   New(bufferSize, lines...).Start().Wait()   pipe.go:107-126,197-257, run.go:173-196 (one thread per component, cap-1 channels)
//   sync / async fittings                      internal/fitting/fitting.go:39-104
//   mock::Source / Processor / Sink            mock/mock.go:61-192
//
// Go's (int, error) becomes pipe::Result{n, err}; io.EOF is pipe::Error::Eof().  A signal.Floating is pipe::Floating<T>:
// frame-major, channel-interleaved (value index = frame * channels + channel, mock/mock.go:95-101); Length() is frames.
// The reference allocates float64 buffers (pipe.go:394,437); T = float selects the float32 path the GPU chain is fastest on.
// Only what the per-buffer path needs is mirrored: no AddLine / InsertProcessor (control plane, SURVEY.md section 8 "next").
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/pipe_b200.h"

namespace pipe {

// ------------------------------------------------------------------------------------------------------------- errors --
struct Error {
    bool set = false, eof = false;
    std::string msg;
    static Error None() { return {}; }
    static Error Eof() { return {true, true, "EOF"}; }
    static Error New(std::string m) { return {true, false, std::move(m)}; }
    explicit operator bool() const { return set; }
    Error Wrap(const std::string &what) const { return {true, false, what + ": " + msg}; }  // fmt.Errorf("...: %w", err)
};
struct Result {
    int n = 0;
    Error err;
};

// ErrorRun (error.go:11-35): execution and/or flush failed after a successful start
struct RunError {
    Error exec, flush;
    explicit operator bool() const { return exec.set || flush.set; }
    std::string String() const { return "execute error: " + (exec.set ? exec.msg : "<nil>") + ", flush error: " + (flush.set ? flush.msg : "<nil>"); }
};

struct SignalProperties {  // line.go:38-41
    double SampleRate = 0;
    int Channels = 0;
};

// ------------------------------------------------------------------------------------------------------------- buffers --
template <typename T>
struct Floating {  // what the hot path uses of signal.Floating: Channels, Length, Capacity, Slice, flat access
    int channels = 0;
    std::shared_ptr<std::vector<T>> store;
    int length = 0;  // frames
    static Floating Alloc(int channels, int frames)
    {
        Floating f;
        f.channels = channels;
        f.store = std::make_shared<std::vector<T>>((size_t)std::max(channels, 0) * (size_t)frames);
        f.length = frames;
        return f;
    }
    int Channels() const { return channels; }
    int Length() const { return length; }
    int Len() const { return length * channels; }
    T *Data() { return store ? store->data() : nullptr; }
    const T *Data() const { return store ? store->data() : nullptr; }
    T Sample(int i) const { return (*store)[(size_t)i]; }
    void SetSample(int i, T v) { (*store)[(size_t)i] = v; }
    Floating Slice(int frames) const  // Slice(0, frames): pipe.go:404-406,441-443
    {
        Floating f = *this;
        f.length = std::min(frames, length);
        return f;
    }
};

// ---------------------------------------------------------------------------------------------------------- components --
using HookFunc = std::function<Error()>;

template <typename T>
struct Source {  // pipe.go:35-43
    std::function<Result(Floating<T> &out)> SourceFunc;
    HookFunc StartFunc, FlushFunc;
    SignalProperties Props;
};
template <typename T>
struct Processor {  // pipe.go:52-60
    std::function<Result(const Floating<T> &in, Floating<T> &out)> ProcessFunc;
    HookFunc StartFunc, FlushFunc;
    SignalProperties Props;
};
template <typename T>
struct Sink {  // pipe.go:69-76
    std::function<Error(const Floating<T> &in)> SinkFunc;
    HookFunc StartFunc, FlushFunc;
};

template <typename T>
using SourceAllocatorFunc = std::function<Error(int bufferSize, Source<T> &out)>;
template <typename T>
using ProcessorAllocatorFunc = std::function<Error(int bufferSize, SignalProperties input, Processor<T> &out)>;
template <typename T>
using SinkAllocatorFunc = std::function<Error(int bufferSize, SignalProperties input, Sink<T> &out)>;

template <typename T>
struct Line {  // line.go:14-19
    SourceAllocatorFunc<T> Source;
    std::vector<ProcessorAllocatorFunc<T>> Processors;
    SinkAllocatorFunc<T> Sink;
};

// ------------------------------------------------------------------------------------------------------------- fittings --
namespace detail {

template <typename T>
struct Message {  // fitting.go:12-15
    Floating<T> signal;
};

template <typename T>
struct Fitting {
    virtual ~Fitting() = default;
    virtual bool Send(Message<T> m) = 0;                 // false: the receiver is gone / the context was cancelled
    virtual bool Receive(Message<T> &m) = 0;             // false: closed
    virtual void Close() = 0;
};

template <typename T>
struct SyncFitting : Fitting<T> {  // fitting.go:39-42,62-79
    bool closed = false;
    Message<T> slot;
    bool Send(Message<T> m) override
    {
        if (closed) return false;
        slot = std::move(m);
        return true;
    }
    bool Receive(Message<T> &m) override
    {
        m = slot;
        return !closed;
    }
    void Close() override { closed = true; }
};

template <typename T>
struct AsyncFitting : Fitting<T> {  // fitting.go:44-47,56-60,81-104: chan Message with capacity 1
    std::mutex mu;
    std::condition_variable cv;
    std::optional<Message<T>> slot;
    bool closed = false;
    std::atomic<bool> *cancel;
    explicit AsyncFitting(std::atomic<bool> *c) : cancel(c) {}
    bool Send(Message<T> m) override
    {
        std::unique_lock<std::mutex> lk(mu);
        while (slot.has_value() && !cancel->load()) cv.wait_for(lk, std::chrono::milliseconds(20));
        if (cancel->load()) return false;
        slot = std::move(m);
        cv.notify_all();
        return true;
    }
    bool Receive(Message<T> &m) override
    {
        std::unique_lock<std::mutex> lk(mu);
        while (!slot.has_value() && !closed && !cancel->load()) cv.wait_for(lk, std::chrono::milliseconds(20));
        if (slot.has_value()) {
            m = std::move(*slot);
            slot.reset();
            cv.notify_all();
            return true;
        }
        return false;
    }
    void Close() override
    {
        std::lock_guard<std::mutex> lk(mu);
        closed = true;
        cv.notify_all();
    }
};

// ------------------------------------------------------------------------------------------------------------ executors --
template <typename T>
struct Executor {
    virtual ~Executor() = default;
    virtual Error Execute() = 0;
    HookFunc start, flush;
    Error StartHook() { return start ? start() : Error::None(); }
    Error FlushHook() { return flush ? flush() : Error::None(); }
    std::shared_ptr<Fitting<T>> in, out;
};

template <typename T>
struct SourceExec : Executor<T> {  // Source.execute, pipe.go:381-413
    Source<T> c;
    int bufferSize;
    Error Execute() override
    {
        Floating<T> output = Floating<T>::Alloc(c.Props.Channels, bufferSize);
        Result r = c.SourceFunc(output);
        if (r.err) {
            this->out->Close();
            return r.err;
        }
        if (r.n != output.Length()) output = output.Slice(r.n);
        if (!this->out->Send({output})) {
            this->out->Close();
            return Error::Eof();
        }
        return Error::None();
    }
};

template <typename T>
struct ProcExec : Executor<T> {  // Processor.execute, pipe.go:425-451
    Processor<T> c;
    int bufferSize;
    Error Execute() override
    {
        Message<T> m;
        if (!this->in->Receive(m)) {
            this->out->Close();
            return Error::Eof();
        }
        Floating<T> output = Floating<T>::Alloc(c.Props.Channels, bufferSize);
        Result r = c.ProcessFunc(m.signal, output);
        if (r.err) {
            this->out->Close();
            return r.err;
        }
        if (r.n != bufferSize) output = output.Slice(r.n);
        if (!this->out->Send({output})) {
            this->out->Close();
            return Error::Eof();
        }
        return Error::None();
    }
};

template <typename T>
struct SinkExec : Executor<T> {  // Sink.execute, pipe.go:459-471
    Sink<T> c;
    Error Execute() override
    {
        Message<T> m;
        if (!this->in->Receive(m)) return Error::Eof();
        return c.SinkFunc(m.signal);
    }
};

template <typename T>
using Route = std::vector<std::shared_ptr<Executor<T>>>;

template <typename T>
Error Bind(const Line<T> &line, int bufferSize, Route<T> &route)  // Line.route, line.go:62-90
{
    auto se = std::make_shared<SourceExec<T>>();
    se->bufferSize = bufferSize;
    if (Error e = line.Source(bufferSize, se->c)) return e.Wrap("source");
    se->start = se->c.StartFunc;
    se->flush = se->c.FlushFunc;
    route.push_back(se);
    SignalProperties prev = se->c.Props;
    for (const auto &alloc : line.Processors) {
        auto pe = std::make_shared<ProcExec<T>>();
        pe->bufferSize = bufferSize;
        if (Error e = alloc(bufferSize, prev, pe->c)) return e.Wrap("processor");
        pe->start = pe->c.StartFunc;
        pe->flush = pe->c.FlushFunc;
        prev = pe->c.Props;  // the output properties become the next stage's input (line.go:75)
        route.push_back(pe);
    }
    auto ke = std::make_shared<SinkExec<T>>();
    if (Error e = line.Sink(bufferSize, prev, ke->c)) return e.Wrap("sink");
    ke->start = ke->c.StartFunc;
    ke->flush = ke->c.FlushFunc;
    route.push_back(ke);
    return Error::None();
}

template <typename T, typename Make>
void Connect(Route<T> &route, Make make)  // route.connect, line.go:92-104
{
    for (size_t i = 0; i + 1 < route.size(); i++) {
        std::shared_ptr<Fitting<T>> f = make();
        route[i]->out = f;
        route[i + 1]->in = f;
    }
}

template <typename T>
struct LineExecutor {  // run.go:20-74
    Route<T> executors;
    int started = 0;
    Error Execute()
    {
        Error err;
        for (int i = 0; i < started; i++) {
            Error e = executors[(size_t)i]->Execute();
            if (!e) err = Error::None();
            else if (e.eof) err = e;  // continue execution to propagate EOF (run.go:44)
            else return e;
        }
        return err;
    }
    Error StartHook()
    {
        for (auto &e : executors) {
            if (Error err = e->StartHook()) return err;
            started++;
        }
        return Error::None();
    }
    Error FlushHook()  // only what was started is flushed (run.go:60-74)
    {
        std::string all;
        for (int i = 0; i < started; i++)
            if (Error err = executors[(size_t)i]->FlushHook()) all += (all.empty() ? "" : ",") + err.msg;
        return all.empty() ? Error::None() : Error::New(all);
    }
};

}  // namespace detail

// ------------------------------------------------------------------------------------------------------------- pipe.Run --
// pipe.Run (pipe.go:90-103): every line in the calling thread, one buffer per line per iteration (multiLineExecutor.execute,
// run.go:113-132); a line is flushed and removed at EOF; the rest is flushed by the deferred hook (run.go:204-213).
template <typename T>
RunError Run(int bufferSize, const std::vector<Line<T>> &lines, Error *startError = nullptr)
{
    std::vector<detail::LineExecutor<T>> les(lines.size());
    for (size_t i = 0; i < lines.size(); i++) {
        if (Error e = detail::Bind(lines[i], bufferSize, les[i].executors)) {
            if (startError) *startError = e;
            return {e, {}};
        }
        detail::Connect<T>(les[i].executors, [] { return std::make_shared<detail::SyncFitting<T>>(); });
    }
    for (auto &le : les)  // multiLineExecutor.startHook, run.go:78-99
        if (Error e = le.StartHook()) {
            Error se = e.Wrap("error starting lines");
            std::string fl;
            for (auto &l2 : les)
                if (Error fe = l2.FlushHook()) fl += (fl.empty() ? "" : ",") + fe.msg;
            if (!fl.empty()) se = Error::New("error flushing lines: " + fl + " during start error: " + se.msg);
            if (startError) *startError = se;
            return {se, {}};
        }
    std::vector<detail::LineExecutor<T> *> alive;
    for (auto &le : les) alive.push_back(&le);
    RunError out;
    while (!out.exec.set && !alive.empty()) {
        for (size_t i = 0; i < alive.size();) {
            Error e = alive[i]->Execute();
            if (!e) {
                i++;
            } else if (e.eof) {
                if (Error fe = alive[i]->FlushHook()) {  // returned before the line is removed (run.go:121-123)
                    out.exec = fe;
                    break;
                }
                alive.erase(alive.begin() + (long)i);
            } else {
                out.exec = e;
                break;
            }
        }
    }
    if (out.exec.set) out.exec = out.exec.Wrap("error running");
    std::string fl;
    for (auto *le : alive)
        if (Error fe = le->FlushHook()) fl += (fl.empty() ? "" : ",") + fe.msg;
    if (!fl.empty()) out.flush = Error::New("error flushing: " + fl);
    return out;
}

// ------------------------------------------------------------------------------------------------ pipe.New / Start / Wait --
// Immutable (async) lines: one thread per component, cap-1 channels between them (pipe.go:107-126,172-214; run.go:173-196);
// the first error cancels the context (pipe.go:230-237).
template <typename T>
class Pipe {
public:
    static Error New(int bufferSize, const std::vector<Line<T>> &lines, std::unique_ptr<Pipe> &out)
    {
        if (lines.empty()) return Error::New("pipe without lines");  // pipe.go:108-110 panics
        auto p = std::unique_ptr<Pipe>(new Pipe());
        p->routes_.resize(lines.size());
        for (size_t i = 0; i < lines.size(); i++)
            if (Error e = detail::Bind(lines[i], bufferSize, p->routes_[i])) return e;
        out = std::move(p);
        return Error::None();
    }
    Pipe &Start()
    {
        cancel_.store(false);
        errors_.clear();
        for (auto &r : routes_) {
            detail::Connect<T>(r, [this] { return std::make_shared<detail::AsyncFitting<T>>(&cancel_); });
            for (auto &ex : r) threads_.emplace_back([this, ex] { Component(*ex); });
        }
        return *this;
    }
    // pipe.Wait: the first error, nil when every component reached EOF cleanly
    Error Wait()
    {
        for (auto &t : threads_) t.join();
        threads_.clear();
        return errors_.empty() ? Error::None() : errors_.front();
    }

    ~Pipe()  // a Pipe dropped without Wait cancels its components and joins them
    {
        cancel_.store(true);
        for (auto &t : threads_)
            if (t.joinable()) t.join();
    }

private:
    Pipe() = default;
    void Fail(Error e)
    {
        std::lock_guard<std::mutex> lk(mu_);
        errors_.push_back(std::move(e));
        cancel_.store(true);
    }
    void Component(detail::Executor<T> &ex)  // start(), run.go:173-196
    {
        if (Error e = ex.StartHook()) {
            if (ex.out) ex.out->Close();
            return Fail(e.Wrap("error starting"));
        }
        Error err;
        while (!err && !cancel_.load()) err = ex.Execute();
        Error fe = ex.FlushHook();  // deferred: always runs once the component has started
        if (err && !err.eof) Fail(err.Wrap("error running"));
        if (fe) Fail(fe.Wrap("error flushing"));
    }
    std::vector<detail::Route<T>> routes_;
    std::vector<std::thread> threads_;
    std::vector<Error> errors_;
    std::mutex mu_;
    std::atomic<bool> cancel_{false};
};

// ---------------------------------------------------------------------------------------------------------------- mock --
namespace mock {

struct Counter {  // mock.go:17-21,43-46
    int Messages = 0;
    int64_t Samples = 0;  // frames
    void advance(int size)
    {
        Messages++;
        Samples += size;
    }
};

struct Flusher {  // mock.go:49-58: the flag is set before the error is returned
    bool Started = false, Flushed = false;
    Error ErrorOnStart, ErrorOnFlush;
    HookFunc StartHook()
    {
        return [this] {
            Started = true;
            return ErrorOnStart;
        };
    }
    HookFunc FlushHook()
    {
        return [this] {
            Flushed = true;
            return ErrorOnFlush;
        };
    }
};

template <typename T>
struct Source : Flusher {  // mock.go:61-109
    int64_t Limit = 0;
    T Value = 0;
    int Channels = 0;
    double SampleRate = 0;
    Error ErrorOnCall, ErrorOnMake;
    std::function<void(Floating<T> &out, int frames, int64_t firstFrame)> Fill;  // default: the constant Value
    Counter counter;
    SourceAllocatorFunc<T> Allocator()
    {
        return [this](int, pipe::Source<T> &s) -> Error {
            if (ErrorOnMake) return ErrorOnMake;
            s.Props = {SampleRate, Channels};
            s.StartFunc = StartHook();
            s.FlushFunc = FlushHook();
            s.SourceFunc = [this](Floating<T> &out) -> Result {
                if (ErrorOnCall) return {0, ErrorOnCall};
                if (counter.Samples == Limit) return {0, Error::Eof()};
                const int read = (int)std::min<int64_t>(out.Length(), Limit - counter.Samples);
                if (Fill) Fill(out, read, counter.Samples);
                else std::fill(out.Data(), out.Data() + (size_t)read * (size_t)out.Channels(), Value);
                counter.advance(read);
                return {read, Error::None()};
            };
            return Error::None();
        };
    }
};

template <typename T>
struct Processor : Flusher {  // mock.go:130-157
    Error ErrorOnCall, ErrorOnMake;
    Counter counter;
    ProcessorAllocatorFunc<T> Allocator()
    {
        return [this](int, SignalProperties in, pipe::Processor<T> &p) -> Error {
            if (ErrorOnMake) return ErrorOnMake;
            p.Props = in;
            p.StartFunc = StartHook();
            p.FlushFunc = FlushHook();
            p.ProcessFunc = [this](const Floating<T> &in, Floating<T> &out) -> Result {
                if (ErrorOnCall) return {0, ErrorOnCall};
                const int n = std::min(in.Length(), out.Length());  // signal.FloatingAsFloating
                std::copy(in.Data(), in.Data() + (size_t)n * (size_t)in.Channels(), out.Data());
                counter.advance(n);
                return {n, Error::None()};
            };
            return Error::None();
        };
    }
};

template <typename T>
struct Sink : Flusher {  // mock.go:160-192
    bool Discard = false;
    Error ErrorOnCall, ErrorOnMake;
    Counter counter;
    std::vector<T> Values;  // interleaved, appended per message
    int Channels = 0;
    SinkAllocatorFunc<T> Allocator()
    {
        return [this](int, SignalProperties in, pipe::Sink<T> &s) -> Error {
            if (ErrorOnMake) return ErrorOnMake;
            Channels = in.Channels;
            s.StartFunc = StartHook();
            s.FlushFunc = FlushHook();
            s.SinkFunc = [this](const Floating<T> &in) -> Error {
                if (ErrorOnCall) return ErrorOnCall;
                if (!Discard) Values.insert(Values.end(), in.Data(), in.Data() + in.Len());
                counter.advance(in.Length());
                return Error::None();
            };
            return Error::None();
        };
    }
};

}  // namespace mock

// ----------------------------------------------------------------------------------------------------------------- gpu --
namespace gpu {

// One contiguous run of GPU Processors in Line.Processors as ONE pipe::Processor: the allocator creates the pb_chain
// (line.go:71 -> pb_chain_create), ProcessFunc is one call across the C-ABI per buffer (pipe.go:438 -> pb_chain_process).
// The reference binds components once and a Pipe may be started again after Wait (pipe_test.go:107-130), so the handle lives
// as long as the Processor's closures: StartFunc of a restart zeroes the carried state (pb_chain_reset), FlushFunc
// (run.go:181-185) only drains the device, and the last closure to go destroys the chain.  There is no CPU fallback: without a usable sm_100 device the
// allocator fails and pipe::Run / Pipe::New return that error, as line.go:72-74 does for any allocator error.
struct Stage {
    pb_stage_desc d{};
    std::vector<double> taps;
    static Stage Copy() { Stage s; s.d.kind = PB_STAGE_COPY; return s; }
    static Stage Gain(double g) { Stage s; s.d.kind = PB_STAGE_GAIN; s.d.gain = g; return s; }
    static Stage Biquad(const double (&b)[3], const double (&a)[2])
    {
        Stage s;
        s.d.kind = PB_STAGE_BIQUAD;
        std::copy(b, b + 3, s.d.b);
        std::copy(a, a + 2, s.d.a);
        return s;
    }
    static Stage Fir(std::vector<double> taps)
    {
        Stage s;
        s.d.kind = PB_STAGE_FIR;
        s.taps = std::move(taps);
        s.d.n_taps = (int32_t)s.taps.size();
        return s;
    }
    static Stage Resample(int up, int down, std::vector<double> prototype)
    {
        Stage s;
        s.d.kind = PB_STAGE_RESAMPLE;
        s.d.up = up;
        s.d.down = down;
        s.taps = std::move(prototype);
        s.d.n_taps = (int32_t)s.taps.size();
        return s;
    }
};

struct ChainHandle {  // shared by the three funcs of the Processor; destroyed with the last reference
    pb_chain *h = nullptr;
    int starts = 0;
    ~ChainHandle()
    {
        if (h) pb_chain_destroy(h);
    }
};

template <typename T>
ProcessorAllocatorFunc<T> Chain(std::vector<Stage> stages, int device = 0, unsigned flags = 0)
{
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "signal.Floating is float32 or float64");
    return [stages = std::move(stages), device, flags](int bufferSize, SignalProperties in, Processor<T> &p) -> Error {
        std::vector<pb_stage_desc> descs;
        for (const Stage &s : stages) {
            pb_stage_desc d = s.d;
            d.taps = s.taps.empty() ? nullptr : s.taps.data();
            descs.push_back(d);
        }
        pb_chain_desc cd{};
        cd.abi_version = PB_ABI_VERSION;
        cd.device = device;
        cd.dtype = std::is_same<T, float>::value ? PB_F32 : PB_F64;
        cd.channels = in.Channels;
        cd.sample_rate = in.SampleRate;
        cd.buffer_frames = bufferSize;
        cd.max_batch = 1;
        cd.n_stages = (int32_t)descs.size();
        cd.flags = (int32_t)flags;
        cd.stages = descs.data();
        auto ch = std::make_shared<ChainHandle>();
        if (pb_chain_create(&cd, &ch->h) != PB_OK) return Error::New(std::string("pb_chain_create: ") + pb_last_error());
        int32_t oc = 0;
        double osr = 0;
        pb_chain_out_properties(ch->h, &oc, &osr);
        p.Props = {osr, oc};
        p.ProcessFunc = [ch](const Floating<T> &in, Floating<T> &out) -> Result {
            int64_t got = 0;
            if (pb_chain_process(ch->h, in.Data(), in.Length(), out.Data(), out.Length(), &got) != PB_OK)
                return {0, Error::New(std::string("pb_chain_process: ") + pb_last_error())};
            return {(int)got, Error::None()};
        };
        p.StartFunc = [ch]() -> Error {  // a restarted Pipe begins from zero state
            if (ch->starts++ > 0 && pb_chain_reset(ch->h) != PB_OK) return Error::New(std::string("pb_chain_reset: ") + pb_last_error());
            return Error::None();
        };
        p.FlushFunc = [ch]() -> Error {  // everything enqueued has completed; the chain stays bound for the next Start
            if (pb_chain_sync(ch->h, nullptr) != PB_OK) return Error::New(std::string("pb_chain_sync: ") + pb_last_error());
            return Error::None();
        };
        return Error::None();
    };
}

}  // namespace gpu

}  // namespace pipe
