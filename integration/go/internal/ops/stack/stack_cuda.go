//go:build cuda

package stack

/*
#cgo LDFLAGS: -lnightlight_cuda
#include "nightlight_cuda.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"

	"github.com/mlnoga/nightlight/internal/cuda"
	"github.com/mlnoga/nightlight/internal/fits"
	"github.com/mlnoga/nightlight/internal/ops"
)

// Apply stacks a set of light frames on the GPU (replaces stack.go:115-227; same log lines and errors).
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
	ctx := cuda.Get()
	defer cuda.Put(ctx)

	var job *C.nl_stack_job
	if rc := C.nl_stack_begin((*C.nl_ctx)(ctx.H), C.int32_t(len(f)), C.int64_t(len(f[0].Data)), &job); rc != 0 {
		return nil, cuda.LastErr()
	}
	defer C.nl_stack_end(job)
	for i, img := range f { // [][]float32 cannot cross cgo: one pointer-free slice per call
		if rc := C.nl_stack_put_frame(job, C.int32_t(i), (*C.float)(&img.Data[0]), C.int64_t(len(img.Data))); rc != 0 {
			return nil, cuda.LastErr()
		}
	}
	data := make([]float32, len(f[0].Data))
	var w *C.float
	if weights != nil {
		w = (*C.float)(&weights[0])
	}
	var numClippedLow, numClippedHigh C.int64_t
	if rc := C.nl_stack_run(job, C.int32_t(mode), w, C.float(op.SigmaLow), C.float(op.SigmaHigh), C.float(op.RefFrameLoc),
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
