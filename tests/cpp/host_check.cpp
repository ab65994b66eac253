// host_check.cpp -- driver for tests/test_cpp_host.py: the C++ host mirror (pipe_b200/host/pipe.hpp) against the reference's
// own plumbing goldens (mode "plumbing", no GPU) and through the C-ABI on a GPU (mode "gpu").  Prints key=value lines.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../pipe_b200/host/pipe.hpp"

using namespace pipe;

template <typename T>
static Line<T> mock_line(mock::Source<T> &s, mock::Processor<T> &p, mock::Sink<T> &k)
{
    Line<T> l;
    l.Source = s.Allocator();
    l.Processors = {p.Allocator()};
    l.Sink = k.Allocator();
    return l;
}

static int plumbing()
{
    {  // pipe_test.go:84-105: 2 ch, Limit 862*512, bufferSize 512 -> 862 messages, 441344 frames
        mock::Source<double> s;
        s.Limit = 862 * 512;
        s.Channels = 2;
        s.Value = 0.5;
        mock::Processor<double> p;
        mock::Sink<double> k;
        k.Discard = true;
        RunError e = Run<double>(512, {mock_line(s, p, k)});
        printf("golden862.err=%d\ngolden862.messages=%d\ngolden862.samples=%lld\ngolden862.proc_messages=%d\n", (int)(bool)e,
               k.counter.Messages, (long long)k.counter.Samples, p.counter.Messages);
        printf("golden862.hooks=%d%d%d%d%d%d\n", s.Started, s.Flushed, p.Started, p.Flushed, k.Started, k.Flushed);
    }
    for (int limit : {1040, 1640, 3048, 4096}) {  // pipe_test.go:337,363,394,399,404
        mock::Source<double> s;
        s.Limit = limit;
        s.Channels = 1;
        mock::Processor<double> p;
        mock::Sink<double> k;
        k.Discard = true;
        RunError e = Run<double>(512, {mock_line(s, p, k)});
        printf("limit%d.err=%d\nlimit%d.messages=%d\nlimit%d.samples=%lld\n", limit, (int)(bool)e, limit, k.counter.Messages, limit,
               (long long)k.counter.Samples);
    }
    {  // mock_test.go:69-92: 11 frames / buffer 5 -> 3 calls; values pass through (mock_test.go:133-146)
        mock::Source<double> s;
        s.Limit = 11;
        s.Channels = 2;
        s.Value = 1.0;
        mock::Processor<double> p;
        mock::Sink<double> k;
        RunError e = Run<double>(5, {mock_line(s, p, k)});
        bool ones = k.Values.size() == 22;
        for (double v : k.Values) ones = ones && v == 1.0;
        printf("values.err=%d\nvalues.messages=%d\nvalues.all_ones=%d\n", (int)(bool)e, k.counter.Messages, (int)ones);
    }
    {  // two lines in one Run (multiLineExecutor, run.go:113-132): the shorter one is flushed and removed first
        mock::Source<double> s1, s2;
        s1.Limit = 1040;
        s1.Channels = 1;
        s2.Limit = 3048;
        s2.Channels = 3;
        mock::Processor<double> p1, p2;
        mock::Sink<double> k1, k2;
        k1.Discard = k2.Discard = true;
        RunError e = Run<double>(512, {mock_line(s1, p1, k1), mock_line(s2, p2, k2)});
        printf("two.err=%d\ntwo.messages=%d,%d\ntwo.samples=%lld,%lld\ntwo.flushed=%d%d\n", (int)(bool)e, k1.counter.Messages,
               k2.counter.Messages, (long long)k1.counter.Samples, (long long)k2.counter.Samples, k1.Flushed, k2.Flushed);
    }
    {  // error surfacing (pipe_test.go:437-457): a ProcessFunc error ends the run wrapped as "error running: ..."
        mock::Source<double> s;
        s.Limit = 2048;
        s.Channels = 1;
        mock::Processor<double> p;
        p.ErrorOnCall = Error::New("mock error");
        mock::Sink<double> k;
        RunError e = Run<double>(512, {mock_line(s, p, k)});
        printf("procerr.exec=%s\nprocerr.flushed=%d%d%d\n", e.exec.msg.c_str(), s.Flushed, p.Flushed, k.Flushed);
    }
    {  // allocator error aborts binding (line.go:72-74), nothing is started
        mock::Source<double> s;
        s.Limit = 10;
        s.Channels = 1;
        mock::Processor<double> p;
        p.ErrorOnMake = Error::New("mock error");
        mock::Sink<double> k;
        Error se;
        RunError e = Run<double>(512, {mock_line(s, p, k)}, &se);
        printf("makeerr.msg=%s\nmakeerr.started=%d\n", se.msg.c_str(), (int)s.Started);
    }
    {  // start error: what was started is flushed (run.go:78-99)
        mock::Source<double> s;
        s.Limit = 10;
        s.Channels = 1;
        mock::Processor<double> p;
        p.ErrorOnStart = Error::New("mock error");
        mock::Sink<double> k;
        Error se;
        Run<double>(512, {mock_line(s, p, k)}, &se);
        printf("starterr.msg=%s\nstarterr.flags=%d%d%d%d%d%d\n", se.msg.c_str(), s.Started, s.Flushed, p.Started, p.Flushed, k.Started,
               k.Flushed);
    }
    {  // pipe.New + Start + Wait: one thread per component, cap-1 channels (the same golden)
        mock::Source<double> s;
        s.Limit = 862 * 512;
        s.Channels = 2;
        mock::Processor<double> p;
        mock::Sink<double> k;
        k.Discard = true;
        std::unique_ptr<Pipe<double>> pp;
        Error ne = Pipe<double>::New(512, {mock_line(s, p, k)}, pp);
        Error we = ne ? ne : pp->Start().Wait();
        printf("async.err=%d\nasync.messages=%d\nasync.samples=%lld\nasync.flushed=%d%d%d\n", (int)(bool)we, k.counter.Messages,
               (long long)k.counter.Samples, s.Flushed, p.Flushed, k.Flushed);
    }
    {  // async error: the failing component cancels the others
        mock::Source<double> s;
        s.Limit = 1 << 20;
        s.Channels = 1;
        mock::Processor<double> p;
        mock::Sink<double> k;
        k.Discard = true;
        k.ErrorOnCall = Error::New("mock error");
        std::unique_ptr<Pipe<double>> pp;
        Pipe<double>::New(512, {mock_line(s, p, k)}, pp);
        Error we = pp->Start().Wait();
        printf("asyncerr.msg=%s\n", we.msg.c_str());
    }
    return 0;
}

// stages file: one stage per line -- "gain g" | "biquad b0 b1 b2 a1 a2" | "fir n t..." | "resample up down n t..." | "copy"
static std::vector<gpu::Stage> read_stages(const char *path)
{
    std::vector<gpu::Stage> st;
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream is(line);
        std::string kind;
        if (!(is >> kind)) continue;
        if (kind == "copy") {
            st.push_back(gpu::Stage::Copy());
        } else if (kind == "gain") {
            double g;
            is >> g;
            st.push_back(gpu::Stage::Gain(g));
        } else if (kind == "biquad") {
            double b[3], a[2];
            is >> b[0] >> b[1] >> b[2] >> a[0] >> a[1];
            st.push_back(gpu::Stage::Biquad(b, a));
        } else if (kind == "fir" || kind == "resample") {
            int up = 1, down = 1, n = 0;
            if (kind == "resample") is >> up >> down;
            is >> n;
            std::vector<double> t((size_t)n);
            for (double &v : t) is >> v;
            st.push_back(kind == "fir" ? gpu::Stage::Fir(t) : gpu::Stage::Resample(up, down, t));
        }
    }
    return st;
}

// gpu <stages.txt> <in.f32> <out.f32> <channels> <frames> <bufferSize> <sampleRate>
static int gpu_mode(char **a)
{
    const int channels = atoi(a[3]), buffer = atoi(a[5]);
    const long long frames = atoll(a[4]);
    std::vector<float> x((size_t)frames * (size_t)channels);
    {
        std::ifstream f(a[1], std::ios::binary);
        f.read(reinterpret_cast<char *>(x.data()), (std::streamsize)(x.size() * sizeof(float)));
    }
    mock::Source<float> s;
    s.Limit = frames;
    s.Channels = channels;
    s.SampleRate = atof(a[6]);
    s.Fill = [&](Floating<float> &out, int n, int64_t first) {
        std::memcpy(out.Data(), x.data() + (size_t)first * (size_t)channels, (size_t)n * (size_t)channels * sizeof(float));
    };
    mock::Sink<float> k;
    std::vector<int> lens;
    Line<float> l;
    l.Source = s.Allocator();
    l.Processors = {gpu::Chain<float>(read_stages(a[0]))};
    auto inner = k.Allocator();
    l.Sink = [&](int bs, SignalProperties in, Sink<float> &out) -> Error {
        Error e = inner(bs, in, out);
        auto f = out.SinkFunc;
        out.SinkFunc = [f, &lens](const Floating<float> &m) {
            lens.push_back(m.Length());
            return f(m);
        };
        printf("out.channels=%d\nout.sample_rate=%.6f\n", in.Channels, in.SampleRate);
        return e;
    };
    Error se;
    RunError e = Run<float>(buffer, {l}, &se);
    printf("gpu.err=%d\ngpu.msg=%s%s\ngpu.messages=%d\ngpu.samples=%lld\ngpu.lens=", (int)(bool)e, e.exec.msg.c_str(), se.msg.c_str(),
           k.counter.Messages, (long long)k.counter.Samples);
    for (int n : lens) printf("%d,", n);
    printf("\n");
    std::ofstream o(a[2], std::ios::binary);
    o.write(reinterpret_cast<const char *>(k.Values.data()), (std::streamsize)(k.Values.size() * sizeof(float)));
    return e ? 1 : 0;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !strcmp(argv[1], "plumbing")) return plumbing();
    if (argc >= 9 && !strcmp(argv[1], "gpu")) return gpu_mode(argv + 2);
    fprintf(stderr, "usage: host_check plumbing | gpu <stages> <in> <out> <channels> <frames> <buffer> <rate>\n");
    return 2;
}
