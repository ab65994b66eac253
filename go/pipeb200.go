// Package pipeb200 puts a fused CUDA chain (libpipe_b200.so, include/pipe_b200.h) behind
// pipelined/pipe's Processor plugin boundary.
//
// STATUS: UNCOMPILED AND UNTESTED.  The build image has no Go toolchain and no copy of
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

func lastError(code C.int32_t) error {
	return fmt.Errorf("pipe_b200 error %d: %s", int(code), C.GoString(C.pb_last_error()))
}

// Chain returns the ProcessorAllocatorFunc (line.go:30) for a run of GPU stages on `device`.
func Chain(device int, stages ...Stage) pipe.ProcessorAllocatorFunc {
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
		desc := C.pb_chain_desc{
			abi_version: C.PB_ABI_VERSION, device: C.int32_t(device), dtype: C.PB_F64, // pipe allocates Float64 (pipe.go:394,437)
			channels: C.int32_t(props.Channels), sample_rate: C.double(props.SampleRate),
			buffer_frames: C.int32_t(bufferSize), max_batch: 1, n_stages: C.int32_t(len(stages)), stages: &cst[0],
		}
		var h *C.pb_chain
		if rc := C.pb_chain_create(&desc, &h); rc != C.PB_OK {
			return pipe.Processor{}, lastError(rc) // aborts binding, line.go:72-74
		}
		var outCh C.int32_t
		var outRate C.double
		C.pb_chain_out_properties(h, &outCh, &outRate)

		// Staging in C memory: Go pointers must not be retained by C, and signal.Floating gives
		// no access to its backing slice, so samples are marshalled through WriteFloat64/ReadFloat64
		// (mock_test.go:120,128 show the same calls).
		n := bufferSize * props.Channels
		cin := (*[1 << 28]float64)(C.malloc(C.size_t(8 * n)))[:n:n]
		cout := (*[1 << 28]float64)(C.malloc(C.size_t(8 * n)))[:n:n]

		return pipe.Processor{
			SignalProperties: pipe.SignalProperties{Channels: int(outCh), SampleRate: signal.Frequency(outRate)},
			ProcessFunc: func(in, out signal.Floating) (int, error) {
				frames := in.Length()
				signal.ReadFloat64(in, cin[:frames*props.Channels])
				var got C.int64_t
				// goroutines migrate between OS threads; the library selects its device on every call
				rc := C.pb_chain_process(h, unsafe.Pointer(&cin[0]), C.int64_t(frames),
					unsafe.Pointer(&cout[0]), C.int64_t(bufferSize), &got)
				if rc != C.PB_OK {
					return 0, lastError(rc) // closes the sender and ends the run, pipe.go:438-440
				}
				signal.WriteFloat64(cout[:int(got)*int(outCh)], out)
				return int(got), nil // a short count slices the output, pipe.go:441-443
			},
			FlushFunc: func(context.Context) error { // guaranteed teardown, run.go:181-185
				C.free(unsafe.Pointer(&cin[0]))
				C.free(unsafe.Pointer(&cout[0]))
				if rc := C.pb_chain_destroy(h); rc != C.PB_OK {
					return lastError(rc)
				}
				runtime.KeepAlive(mctx)
				return nil
			},
		}, nil
	}
}
