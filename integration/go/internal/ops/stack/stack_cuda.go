//go:build cuda

package stack

/*
#cgo LDFLAGS: -lnightlight_cuda
#include <stdlib.h>
#include "nightlight_cuda.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"

	"github.com/mlnoga/nightlight/internal/cuda"
	"github.com/mlnoga/nightlight/internal/fits"
	"github.com/mlnoga/nightlight/internal/ops"
)

// Apply stacks a set of light frames on the GPUs (replaces stack.go:115-227; same log lines and errors).
// One call into the library: all frame pointers in, one image out.  The image's rows are dealt to the devices of
// cuda.PerDevice() -- the analogue of the fan-out over pixel ranges at stack.go:134-147 -- and every device pipelines its
// rows in stripes (upload, stack and download overlap).
func (op *OpStack) Apply(f []*fits.Image, c *ops.Context) (result *fits.Image, err error) {
	mode := op.Mode
	if mode < StMedian || mode > StAuto {
		return nil, errors.New("invalid stacking mode")
	}
	if mode == StAuto {
		mode = autoSelectStackingMode(len(f))
	}
	fmt.Fprintf(c.Log, "Stacking %d frames with stacking mode %d and sigma low %g high %g:\n",
		len(f), mode, op.SigmaLow, op.SigmaHigh)

	weights, err := getWeights(f, op.Weighting) // unchanged Go (stack.go:231-270)
	if err != nil {
		return nil, err
	}
	if mode == StMADSigma && weights != nil {
		panic("MADSigma stacking with weights is still unimplemented")
	}

	runtime.LockOSThread() // nl_last_error is per thread
	defer runtime.UnlockOSThread()

	// The frames are read by DMA straight from the Go slices: page-lock them for the duration of the call (the Go
	// heap does not move objects) -- pageable memory uploads at roughly half the PCIe rate -- and keep the
	// collector from freeing them with a Pinner while C holds their addresses.
	var pinner runtime.Pinner
	defer pinner.Unpin()
	n := len(f)
	ptrs := (*[1 << 28]*C.float)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))[:n:n]
	defer C.free(unsafe.Pointer(&ptrs[0]))
	for i, img := range f {
		pinner.Pin(&img.Data[0])
		if cuda.Pin(img.Data) {
			defer cuda.Unpin(img.Data)
		}
		ptrs[i] = (*C.float)(&img.Data[0])
	}
	data := make([]float32, len(f[0].Data))
	pinner.Pin(&data[0])
	if cuda.Pin(data) {
		defer cuda.Unpin(data)
	}

	ctxs := cuda.PerDevice()
	handles := (*[1 << 10]*C.nl_ctx)(C.malloc(C.size_t(len(ctxs)) * C.size_t(unsafe.Sizeof(uintptr(0)))))[:len(ctxs):len(ctxs)]
	defer C.free(unsafe.Pointer(&handles[0]))
	for i, x := range ctxs {
		handles[i] = (*C.nl_ctx)(x.H)
	}
	var w *C.float
	if weights != nil {
		w = (*C.float)(&weights[0])
	}
	var numClippedLow, numClippedHigh C.int64_t
	if rc := C.nl_stack_apply_multi(&handles[0], C.int32_t(len(ctxs)), &ptrs[0], C.int32_t(n), C.int64_t(len(data)),
		C.int64_t(f[0].Naxisn[0]), 8, C.int32_t(mode), w, C.float(op.SigmaLow), C.float(op.SigmaHigh), C.float(op.RefFrameLoc),
		(*C.float)(&data[0]), &numClippedLow, &numClippedHigh); rc != 0 {
		return nil, cuda.LastErr()
	}

	if mode >= StSigma {
		fmt.Fprintf(c.Log, "Clipped low %d (%.2f%%) high %d (%.2f%%)\n",
			numClippedLow, float32(numClippedLow)*100.0/(float32(len(data)*len(f))),
			numClippedHigh, float32(numClippedHigh)*100.0/(float32(len(data)*len(f))))
	}
	exposureSum := float32(0)
	for _, l := range f {
		exposureSum += l.Exposure
	}
	stack := fits.NewImageFromNaxisn(f[0].Naxisn, data)
	stack.Exposure = exposureSum
	return stack, nil
}

// FindSigmasAndStack is the reference's commented-out goal-seek (stackfindsigma.go:27-170) over frames that stay
// resident in one stack job: count-only trial stacks, then one stack at the sigmas found.
func FindSigmasAndStack(f []*fits.Image, mode StackMode, weights []float32, refMedian, stClipPercLow, stClipPercHigh float32) (
	result *fits.Image, numClippedLow, numClippedHigh int64, sigmaLow, sigmaHigh float32, err error) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	ctx := cuda.Get()
	defer cuda.Put(ctx)
	var job *C.nl_stack_job
	if rc := C.nl_stack_begin((*C.nl_ctx)(ctx.H), C.int32_t(len(f)), C.int64_t(len(f[0].Data)), &job); rc != 0 {
		return nil, 0, 0, 0, 0, cuda.LastErr()
	}
	defer C.nl_stack_end(job)
	for i, img := range f { // [][]float32 cannot cross cgo: one pointer-free slice per call
		if rc := C.nl_stack_put_frame(job, C.int32_t(i), (*C.float)(&img.Data[0]), C.int64_t(len(img.Data))); rc != 0 {
			return nil, 0, 0, 0, 0, cuda.LastErr()
		}
	}
	data := make([]float32, len(f[0].Data))
	var w *C.float
	if weights != nil {
		w = (*C.float)(&weights[0])
	}
	var lo, hi C.int64_t
	var sl, sh C.float
	var trials C.int32_t
	if rc := C.nl_find_sigmas_and_stack(job, C.int32_t(mode), w, C.float(refMedian), C.float(stClipPercLow), C.float(stClipPercHigh),
		(*C.float)(&data[0]), &lo, &hi, &sl, &sh, &trials); rc != 0 {
		return nil, 0, 0, 0, 0, cuda.LastErr()
	}
	return fits.NewImageFromNaxisn(f[0].Naxisn, data), int64(lo), int64(hi), float32(sl), float32(sh), nil
}
