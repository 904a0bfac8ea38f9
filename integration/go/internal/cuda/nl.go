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
	"runtime"
	"strconv"
	"sync"
	"unsafe"
)

// Ctx is one CUDA device plus one stream; calls on a Ctx are serialised on its stream.
type Ctx struct {
	H      *C.nl_ctx
	Device int
}

// LastErr returns the calling thread's last library error. Callers must have locked the OS thread
// (runtime.LockOSThread) between the failing call and this one.
func LastErr() error { return errors.New(C.GoString(C.nl_last_error())) }

// Devices lists the CUDA devices the process may use: NIGHTLIGHT_CUDA_DEVICES="0,1,2,3" (default: all).
func Devices() []int {
	if v := os.Getenv("NIGHTLIGHT_CUDA_DEVICES"); v != "" {
		var out []int
		start := 0
		for i := 0; i <= len(v); i++ {
			if i == len(v) || v[i] == ',' {
				if d, err := strconv.Atoi(v[start:i]); err == nil {
					out = append(out, d)
				}
				start = i + 1
			}
		}
		if len(out) > 0 {
			return out
		}
	}
	var n C.int
	if rc := C.nl_device_count(&n); rc != 0 || n == 0 {
		panic(LastErr()) // no CUDA device: there is no CPU fallback in a cuda build
	}
	out := make([]int, int(n))
	for i := range out {
		out[i] = i
	}
	return out
}

func newCtx(device int) *Ctx {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	var h *C.nl_ctx
	if rc := C.nl_ctx_create(C.int(device), &h); rc != 0 {
		panic(LastErr())
	}
	c := &Ctx{H: h, Device: device}
	// a context owns a stream, scratch and staging buffers and pinned memory: never leave its release to chance
	runtime.SetFinalizer(c, func(c *Ctx) { c.Close() })
	return c
}

// Close releases the context's device and pinned memory.
func (c *Ctx) Close() {
	if c.H != nil {
		C.nl_ctx_destroy(c.H)
		c.H = nil
		runtime.SetFinalizer(c, nil)
	}
}

// A bounded free list instead of a sync.Pool: the garbage collector may drop pooled objects at any time, and a
// dropped context would leak its CUDA resources and cost a fresh nl_ctx_create (cudaMalloc, stream) on the next Get.
var (
	mu   sync.Mutex
	free []*Ctx
	next int // round robin over Devices() for new contexts
)

// MaxIdle bounds the idle contexts kept for reuse (set it to ops.Context.MaxThreads).
var MaxIdle = 16

// Get hands out a context for the calling goroutine (MaterializeAll runs up to c.MaxThreads of them).
func Get() *Ctx {
	mu.Lock()
	if n := len(free); n > 0 {
		c := free[n-1]
		free = free[:n-1]
		mu.Unlock()
		return c
	}
	devs := Devices()
	d := devs[next%len(devs)]
	next++
	mu.Unlock()
	return newCtx(d)
}

// Put returns a context; beyond MaxIdle idle ones it is closed.
func Put(c *Ctx) {
	mu.Lock()
	if len(free) < MaxIdle {
		free = append(free, c)
		mu.Unlock()
		return
	}
	mu.Unlock()
	c.Close()
}

// CloseAll releases every idle context (call at process exit).
func CloseAll() {
	mu.Lock()
	idle := free
	free = nil
	mu.Unlock()
	for _, c := range idle {
		c.Close()
	}
}

// PerDevice returns one context per device of Devices(), created once and kept for the life of the process: the
// multi-device stack (nl_stack_apply_multi) keeps its stripe lanes inside them between calls.
var perDevice struct {
	once sync.Once
	ctx  []*Ctx
}

func PerDevice() []*Ctx {
	perDevice.once.Do(func() {
		for _, d := range Devices() {
			perDevice.ctx = append(perDevice.ctx, newCtx(d))
		}
	})
	return perDevice.ctx
}

// Pin page-locks a pixel slice in place so uploads run at full PCIe speed (Go's heap does not move
// objects); Unpin must be called before the slice is dropped.
func Pin(data []float32) bool {
	if len(data) == 0 {
		return false
	}
	return C.nl_host_register(unsafe.Pointer(&data[0]), C.int64_t(4*len(data))) == 0
}
func Unpin(data []float32) {
	if len(data) > 0 {
		C.nl_host_unregister(unsafe.Pointer(&data[0]))
	}
}
