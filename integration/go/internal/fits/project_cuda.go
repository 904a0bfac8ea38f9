//go:build cuda

package fits

/*
#cgo LDFLAGS: -lnightlight_cuda
#include "nightlight_cuda.h"
*/
import "C"

import (
	"runtime"

	"github.com/mlnoga/nightlight/internal/cuda"
	"github.com/mlnoga/nightlight/internal/star"
)

// Project resamples the image into the destination geometry on the GPU (replaces project.go:26-76).
func (img *Image) Project(destNaxisn []int32, trans star.Transform2D, outOfBounds float32) (res *Image, err error) {
	t := [6]C.float{C.float(trans.A), C.float(trans.B), C.float(trans.C), C.float(trans.D), C.float(trans.E), C.float(trans.F)}
	res = NewImageFromNaxisn(destNaxisn, nil)
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	ctx := cuda.Get()
	defer cuda.Put(ctx)
	rc := C.nl_project((*C.nl_ctx)(ctx.H), (*C.float)(&img.Data[0]), C.int32_t(img.Naxisn[0]), C.int32_t(img.Naxisn[1]),
		(*C.float)(&res.Data[0]), C.int32_t(destNaxisn[0]), C.int32_t(destNaxisn[1]), &t[0], C.float(outOfBounds))
	if rc != 0 {
		return nil, cuda.LastErr() // NL_E_SINGULAR: "Matrix has no inverse" (coord.go:160-163)
	}
	res.ID, res.FileName, res.Exposure, res.Trans = img.ID, img.FileName, img.Exposure, trans
	return res, nil
}
