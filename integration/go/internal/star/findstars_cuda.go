//go:build cuda

package star

/*
#cgo LDFLAGS: -lnightlight_cuda
#include "nightlight_cuda.h"
*/
import "C"

import (
	"runtime"
	"unsafe"

	"github.com/mlnoga/nightlight/internal/cuda"
	"github.com/mlnoga/nightlight/internal/stats"
)

// FindStars runs the candidate scan on the GPU and the sparse per-star steps in the library
// (replaces findstars.go:59-100). star.Star has the memory layout of C.nl_star.
func FindStars(data []float32, width int32, location, scale, starSig, bpSigma, starInOut float32, radius int32,
	medianDiffStats *stats.Stats) (stars []Star, sumOfShifts, avgHFR float32) {
	sd := float32(0)
	if medianDiffStats != nil {
		sd = medianDiffStats.StdDev()
	} else {
		bpSigma = 0 // the reference's fallback draws a random 1 % sample (findstars.go:139-150): not reproducible
	}
	buf := make([]Star, len(data)/100+1024)
	var n C.int32_t
	var sos, hfr C.float
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	ctx := cuda.Get()
	defer cuda.Put(ctx)
	if rc := C.nl_find_stars((*C.nl_ctx)(ctx.H), (*C.float)(&data[0]), C.int32_t(len(data)), C.int32_t(width), C.float(location),
		C.float(scale), C.float(starSig), C.float(bpSigma), C.float(starInOut), C.int32_t(radius), C.float(sd),
		(*C.nl_star)(unsafe.Pointer(&buf[0])), C.int32_t(len(buf)), &n, &sos, &hfr); rc != 0 {
		panic(cuda.LastErr())
	}
	if int(n) > len(buf) {
		n = C.int32_t(len(buf))
	}
	return buf[:n], float32(sos), float32(hfr)
}
