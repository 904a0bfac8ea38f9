//go:build cuda

// Package cuda binds libnightlight_cuda.so (include/nightlight_cuda.h).
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -lnightlight_cuda
#include "nightlight_cuda.h"
*/
import "C"

import (
	"errors"
	"os"
	"strconv"
	"sync"
	"unsafe"
)

// Ctx is one CUDA device plus one stream; calls on a Ctx are serialised on its stream.
type Ctx struct{ H *C.nl_ctx }

func device() int {
	if v, err := strconv.Atoi(os.Getenv("NIGHTLIGHT_CUDA_DEVICE")); err == nil {
		return v
	}
	return 0
}

// LastErr returns the calling thread's last library error. Callers must have locked the OS thread
// (runtime.LockOSThread) between the failing call and this one.
func LastErr() error { return errors.New(C.GoString(C.nl_last_error())) }

var pool = sync.Pool{New: func() any {
	var h *C.nl_ctx
	if rc := C.nl_ctx_create(C.int(device()), &h); rc != 0 {
		panic(LastErr()) // no CUDA device: there is no CPU fallback in a cuda build
	}
	return &Ctx{H: h}
}}

// Get hands out a context for the calling goroutine (MaterializeAll runs up to c.MaxThreads of them).
func Get() *Ctx  { return pool.Get().(*Ctx) }
func Put(c *Ctx) { pool.Put(c) }

// Pin page-locks a pixel slice in place so uploads run at full PCIe speed (Go's heap does not move
// objects); Unpin must be called before the slice is dropped.
func Pin(data []float32) {
	if len(data) > 0 {
		C.nl_host_register(unsafe.Pointer(&data[0]), C.int64_t(4*len(data)))
	}
}
func Unpin(data []float32) {
	if len(data) > 0 {
		C.nl_host_unregister(unsafe.Pointer(&data[0]))
	}
}
