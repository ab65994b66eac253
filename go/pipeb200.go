// Package pipeb200 puts a fused CUDA chain (libpipe_b200.so, include/pipe_b200.h) behind
// pipelined/pipe's Processor plugin boundary.
//
// STATUS: UNCOMPILED AND UNTESTED (round 2: float32-compute mode, restart-safe lifecycle, per-call error text).  The build image has no Go toolchain and no copy of
// pipelined.dev/signal v0.10.0, so this file has never been through `go build`.  It is the
// binding a pipe maintainer would add; the same C-ABI is exercised from Python (ctypes) and
// C++ in this repository's tests.
//
// Usage:
//
//	line := pipe.Line{
//	    Source: src.Source(),
//	    Processors: pipe.Processors(pipeb200.Chain(0,
//	        pipeb200.Gain(0.8), pipeb200.FIR(taps), pipeb200.Biquad(b, a), pipeb200.Resample(147, 160, proto))),
//	    Sink: sink.Sink(),
//	}
//
// A contiguous run of GPU stages is ONE pipe.Processor, so Processor.execute (pipe.go:425-451)
// crosses cgo once per buffer for the whole run.
package pipeb200

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -L${SRCDIR}/../pipe_b200 -lpipe_b200 -Wl,-rpath,${SRCDIR}/../pipe_b200
#include <stdlib.h>
#include "pipe_b200.h"
*/
import "C"

import (
	"context"
	"fmt"
	"runtime"
	"sync"
	"unsafe"

	"pipelined.dev/pipe"
	"pipelined.dev/pipe/mutable"
	"pipelined.dev/signal"
)

// Stage describes one Processor of the fused run.
type Stage struct {
	kind     C.int32_t
	gain     float64
	b        [3]float64
	a        [2]float64
	taps     []float64
	up, down int
}

func Copy() Stage                      { return Stage{kind: C.PB_STAGE_COPY} }
func Gain(g float64) Stage             { return Stage{kind: C.PB_STAGE_GAIN, gain: g} }
func Biquad(b [3]float64, a [2]float64) Stage { return Stage{kind: C.PB_STAGE_BIQUAD, b: b, a: a} }
func FIR(taps []float64) Stage         { return Stage{kind: C.PB_STAGE_FIR, taps: taps} }
func Resample(up, down int, proto []float64) Stage {
	return Stage{kind: C.PB_STAGE_RESAMPLE, up: up, down: down, taps: proto}
}

// call runs one C-ABI call and, on failure, reads the thread-local pb_last_error() on the SAME OS thread: goroutines migrate
// between threads, so the message must be fetched before the goroutine can be rescheduled (ADVICE r1).
func call(f func() C.int32_t) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != C.PB_OK {
		return fmt.Errorf("pipe_b200 error %d: %s", int(rc), C.GoString(C.pb_last_error()))
	}
	return nil
}

// Options of a fused chain.
type Options struct {
	Device int
	// Float32 selects float32 arithmetic on the device.  pipe always hands over float64 buffers (pipe.go:394,437); with
	// Float32 the marshalling copy converts on the way in and out.  This is what reaches the tcgen05 kernel (the 4-stage
	// chain at channel counts that are multiples of 128) and the float32 streaming kernels; without it every chain runs
	// the float64 kernels.  See INTEGRATION.md "which kernel a chain reaches".
	Float32 bool
	Flags   uint32 // PB_CHAIN_*
	// Workers is the number of goroutines the marshalling copies (ReadFloat64 / WriteFloat64, and the float32 narrowing and
	// widening) are cut over, by frame ranges (signal.Floating.Slice gives views of the same backing buffer).  One 4096 x 1024
	// float64 buffer is 32 MiB each way: a single goroutine moves that at memcpy speed (~3 ms per direction), several times
	// the device's share of the call.  0 = min(GOMAXPROCS, 8).
	Workers int
}

// parallelFrames runs f over [0, frames) cut into `workers` contiguous frame ranges, one goroutine each.
func parallelFrames(workers, frames int, f func(lo, hi int)) {
	if workers <= 1 || frames < 4*workers {
		f(0, frames)
		return
	}
	var wg sync.WaitGroup
	per := (frames + workers - 1) / workers
	for lo := 0; lo < frames; lo += per {
		hi := lo + per
		if hi > frames {
			hi = frames
		}
		wg.Add(1)
		go func(lo, hi int) {
			defer wg.Done()
			f(lo, hi)
		}(lo, hi)
	}
	wg.Wait()
}

// Chain returns the ProcessorAllocatorFunc (line.go:30) for a run of GPU stages on `device`, float64 arithmetic.
func Chain(device int, stages ...Stage) pipe.ProcessorAllocatorFunc {
	return ChainWith(Options{Device: device}, stages...)
}

// ChainWith is Chain with options.
func ChainWith(opt Options, stages ...Stage) pipe.ProcessorAllocatorFunc {
	return func(mctx mutable.Context, bufferSize int, props pipe.SignalProperties) (pipe.Processor, error) {
		// The descriptors (and the taps they point to) must live in C memory for the duration of
		// pb_chain_create only: the library copies every coefficient and keeps no caller pointer.
		cst := (*[1 << 16]C.pb_stage_desc)(C.calloc(C.size_t(len(stages)), C.size_t(unsafe.Sizeof(C.pb_stage_desc{}))))
		defer C.free(unsafe.Pointer(cst))
		var ctaps []unsafe.Pointer
		defer func() {
			for _, p := range ctaps {
				C.free(p)
			}
		}()
		for i, s := range stages {
			d := &cst[i]
			d.kind, d.gain = s.kind, C.double(s.gain)
			for k := 0; k < 3; k++ {
				d.b[k] = C.double(s.b[k])
			}
			for k := 0; k < 2; k++ {
				d.a[k] = C.double(s.a[k])
			}
			d.up, d.down, d.n_taps = C.int32_t(s.up), C.int32_t(s.down), C.int32_t(len(s.taps))
			if len(s.taps) > 0 {
				p := C.malloc(C.size_t(8 * len(s.taps)))
				copy((*[1 << 28]float64)(p)[:len(s.taps)], s.taps)
				ctaps = append(ctaps, p)
				d.taps = (*C.double)(p)
			}
		}
		dtype := C.int32_t(C.PB_F64)
		if opt.Float32 {
			dtype = C.PB_F32
		}
		desc := C.pb_chain_desc{
			abi_version: C.PB_ABI_VERSION, device: C.int32_t(opt.Device), dtype: dtype,
			channels: C.int32_t(props.Channels), sample_rate: C.double(props.SampleRate),
			buffer_frames: C.int32_t(bufferSize), max_batch: 1, n_stages: C.int32_t(len(stages)),
			flags: C.int32_t(opt.Flags), stages: &cst[0],
		}
		var h *C.pb_chain
		if err := call(func() C.int32_t { return C.pb_chain_create(&desc, &h) }); err != nil {
			return pipe.Processor{}, err // aborts binding, line.go:72-74
		}
		var outCh C.int32_t
		var outRate C.double
		C.pb_chain_out_properties(h, &outCh, &outRate)

		// Staging in C memory: Go pointers must not be retained by C, and signal.Floating gives
		// no access to its backing slice, so samples are marshalled through WriteFloat64/ReadFloat64
		// (mock_test.go:120,128 show the same calls).  The staging is PINNED (pb_host_alloc_pinned): the library then
		// copies straight from it instead of through its own pinned bounce buffer.
		n := bufferSize * props.Channels
		elem := 8
		if opt.Float32 {
			elem = 4
		}
		var pin, pout unsafe.Pointer
		if err := call(func() C.int32_t { return C.pb_host_alloc_pinned(C.int64_t(n*elem), &pin) }); err != nil {
			C.pb_chain_destroy(h)
			return pipe.Processor{}, err
		}
		if err := call(func() C.int32_t { return C.pb_host_alloc_pinned(C.int64_t(n*elem), &pout) }); err != nil {
			C.pb_host_free_pinned(pin)
			C.pb_chain_destroy(h)
			return pipe.Processor{}, err
		}
		in64 := (*[1 << 28]float64)(pin)[:n:n] // views of the same staging
		out64 := (*[1 << 28]float64)(pout)[:n:n]
		in32 := (*[1 << 29]float32)(pin)[:n:n]
		out32 := (*[1 << 29]float32)(pout)[:n:n]
		var tmp []float64 // float32 mode: ReadFloat64 / WriteFloat64 want []float64
		if opt.Float32 {
			tmp = make([]float64, n)
		}
		starts := 0
		workers := opt.Workers
		if workers <= 0 {
			workers = runtime.GOMAXPROCS(0)
			if workers > 8 {
				workers = 8
			}
		}

		// The chain lives as long as the Processor: pipe binds components once (pipe.New) and a Pipe may be started again
		// after Wait (TestReset, pipe_test.go:107-130).  The finalizer releases the device resources.
		type owner struct{ h *C.pb_chain }
		own := &owner{h}
		runtime.SetFinalizer(own, func(o *owner) {
			C.pb_chain_destroy(o.h)
			C.pb_host_free_pinned(pin)
			C.pb_host_free_pinned(pout)
		})

		return pipe.Processor{
			SignalProperties: pipe.SignalProperties{Channels: int(outCh), SampleRate: signal.Frequency(outRate)},
			StartFunc: func(context.Context) error { // a restarted Pipe begins from zero state
				defer runtime.KeepAlive(own)
				starts++
				if starts == 1 {
					return nil
				}
				return call(func() C.int32_t { return C.pb_chain_reset(h) })
			},
			ProcessFunc: func(in, out signal.Floating) (int, error) {
				defer runtime.KeepAlive(own)
				frames := in.Length()
				ch := props.Channels
				parallelFrames(workers, frames, func(lo, hi int) {
					src := in.Slice(lo, hi)
					if opt.Float32 {
						t := tmp[lo*ch : hi*ch]
						signal.ReadFloat64(src, t)
						d := in32[lo*ch : hi*ch]
						for i, v := range t {
							d[i] = float32(v)
						}
					} else {
						signal.ReadFloat64(src, in64[lo*ch:hi*ch])
					}
				})
				var got C.int64_t
				// goroutines migrate between OS threads; the library selects its device on every call
				if err := call(func() C.int32_t {
					return C.pb_chain_process(h, pin, C.int64_t(frames), pout, C.int64_t(bufferSize), &got)
				}); err != nil {
					return 0, err // closes the sender and ends the run, pipe.go:438-440
				}
				och := int(outCh)
				parallelFrames(workers, int(got), func(lo, hi int) {
					dst := out.Slice(lo, hi) // a view of out's backing buffer (out arrives at full length, pipe.go:437)
					if opt.Float32 {
						t := tmp[lo*och : hi*och]
						for i, v := range out32[lo*och : hi*och] {
							t[i] = float64(v)
						}
						signal.WriteFloat64(t, dst)
					} else {
						signal.WriteFloat64(out64[lo*och:hi*och], dst)
					}
				})
				return int(got), nil // a short count slices the output, pipe.go:441-443
			},
			FlushFunc: func(context.Context) error { // run.go:181-185: everything enqueued has completed; the chain stays bound
				defer runtime.KeepAlive(own)
				runtime.KeepAlive(mctx)
				return call(func() C.int32_t { return C.pb_chain_sync(h, nil) })
			},
		}, nil
	}
}
