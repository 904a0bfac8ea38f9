//go:build cuda

package pre

/*
#cgo LDFLAGS: -lnightlight_cuda
#include "nightlight_cuda.h"
*/
import "C"

import (
	"runtime"

	"github.com/mlnoga/nightlight/internal/cuda"
	"github.com/mlnoga/nightlight/internal/stats"
)

// BadPixelMap on the GPU (replaces badpixels.go:32-51): 3x3 median filter, difference, its statistics and the ordered
// scan for outliers run on the device; only the index list and four scalars come back.
func BadPixelMap(data []float32, width int32, sigmaLow, sigmaHigh float32) (bpm []int32, medianDiffStats *stats.Stats) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	ctx := cuda.Get()
	defer cuda.Put(ctx)
	bpm = make([]int32, len(data)/100+1024)
	var n C.int64_t
	var st [4]C.float
	for {
		if rc := C.nl_bad_pixel_map((*C.nl_ctx)(ctx.H), (*C.float)(&data[0]), C.int64_t(len(data)), C.int32_t(width),
			C.float(sigmaLow), C.float(sigmaHigh), (*C.int32_t)(&bpm[0]), C.int64_t(len(bpm)), &n, &st[0]); rc != 0 {
			panic(cuda.LastErr())
		}
		if int(n) <= len(bpm) {
			break
		}
		bpm = make([]int32, int(n))
	}
	// a Stats that already knows min, mean, max and stdDev (NewStatsWithMMM, stats.go:66-68, plus the std deviation)
	medianDiffStats = stats.NewStatsWithMMMStdDev(nil, 0, float32(st[0]), float32(st[2]), float32(st[1]), float32(st[3]))
	return bpm[:n], medianDiffStats
}
