//go:build cuda

package stats

/*
#cgo LDFLAGS: -lnightlight_cuda
#include "nightlight_cuda.h"
*/
import "C"

import (
	"runtime"

	"github.com/mlnoga/nightlight/internal/cuda"
)

// EstimateNoise on the GPU (replaces noise_amd64.go:25-43 / noise.go:32-55). The context pool selects the numerics
// (AVX2 lane order or pure Go) once, from cpuid, so results equal what this host computed before.
func EstimateNoise(data []float32, width int32) float32 {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	ctx := cuda.Get()
	defer cuda.Put(ctx)
	var noise C.float
	if rc := C.nl_estimate_noise((*C.nl_ctx)(ctx.H), (*C.float)(&data[0]), C.int32_t(len(data)), C.int32_t(width), &noise); rc != 0 {
		panic(cuda.LastErr())
	}
	return float32(noise)
}

// calcAll fills min, mean, max and stdDev in one call (replaces calcMinMeanMax + calcVariance,
// stats_amd64.go:24-45; Stats.Min/Max/Mean/StdDev, stats.go:102-153, set haveMMM and haveStdDev from it).
func calcAll(data []float32) (min, mean, max, stdDev float32) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	ctx := cuda.Get()
	defer cuda.Put(ctx)
	var st [4]C.float
	if rc := C.nl_stats((*C.nl_ctx)(ctx.H), (*C.float)(&data[0]), C.int64_t(len(data)), &st[0]); rc != 0 {
		panic(cuda.LastErr())
	}
	return float32(st[0]), float32(st[1]), float32(st[2]), float32(st[3])
}
