package mp3

/*
#include <stdlib.h>
#include "mp3host.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

// Engine is one or more B200s behind include/mp3host.h: device engines, host parsing threads, pinned arenas.
type Engine struct{ h *C.mp3_engine }

// EngineOptions mirrors mp3_engine_opts.
type EngineOptions struct {
	Devices     []int // CUDA ordinals; empty = device 0.  DecodeBatch deals its streams to them, DecodeStreamSplit a frame range each
	HostThreads int   // parsing threads of DecodeBatch (0 = all cores)
	ChunkFrames int   // Decoder decode-ahead per GPU call (0 = 256)
	TrimGapless bool  // DecodeBatch: drop the LAME encoder delay / padding (lameinfo TotalDelay / TotalPadding)
}

// NewEngine creates an engine.  It fails without a CUDA device: there is no CPU decode path.
func NewEngine(o EngineOptions) (*Engine, error) {
	var co C.mp3_engine_opts
	if len(o.Devices) > C.MP3_MAX_DEVICES {
		return nil, fmt.Errorf("mp3: at most %d devices", int(C.MP3_MAX_DEVICES))
	}
	co.n_devices = C.int(len(o.Devices))
	for i, d := range o.Devices {
		co.devices[i] = C.int(d)
	}
	co.host_threads = C.int(o.HostThreads)
	co.chunk_frames = C.uint32_t(o.ChunkFrames)
	if o.TrimGapless {
		co.trim_gapless = 1
	}
	var h *C.mp3_engine
	if rc := C.mp3_engine_create(&co, &h); rc != C.MP3_OK {
		return nil, fmt.Errorf("mp3: no usable CUDA device (mp3_engine_create: %d); there is no CPU decode path", int(rc))
	}
	e := &Engine{h: h}
	runtime.SetFinalizer(e, (*Engine).Close)
	return e, nil
}

// Close destroys the engine; decoders created from it must be closed first.
func (e *Engine) Close() {
	if e.h != nil {
		C.mp3_engine_destroy(e.h)
		e.h = nil
	}
}

var (
	defaultOnce sync.Once
	defaultEng  *Engine
	defaultErr  error
)

// DefaultEngine is the engine package-level NewDecoder uses: device 0, created on first use.
func DefaultEngine() (*Engine, error) {
	defaultOnce.Do(func() { defaultEng, defaultErr = NewEngine(EngineOptions{}) })
	return defaultEng, defaultErr
}

// StreamResult is one stream of a batch: what io.ReadAll(NewDecoder(stream)) returns in the reference.
type StreamResult struct {
	PCM        []byte // 16-bit LE stereo; a view into the engine's pinned buffer, valid until the next batch call
	SampleRate int    // of the first frame (decode.go:377-381); 0 if the stream failed to open
	Frames     int64
	Err        error // nil: clean end of stream; otherwise the reference's NewDecoder / Read error (PCM holds what came before it)
}

// DecodeBatch decodes many independent streams in one call (north star: the batch entry point).  Streams are dealt
// to the engine's devices in contiguous blocks balanced by bytes; no data moves between devices.
func (e *Engine) DecodeBatch(streams [][]byte) ([]StreamResult, error) {
	n := len(streams)
	if n == 0 {
		return nil, nil
	}
	// cgo: C may not keep Go pointers, and a Go slice of Go pointers may not be passed at all — pin the streams for the call
	var pin runtime.Pinner
	defer pin.Unpin()
	ptrs := (*[1 << 30]*C.uint8_t)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))[:n:n]
	lens := (*[1 << 30]C.size_t)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(C.size_t(0)))))[:n:n]
	defer C.free(unsafe.Pointer(&ptrs[0]))
	defer C.free(unsafe.Pointer(&lens[0]))
	for i, s := range streams {
		if len(s) > 0 {
			pin.Pin(&s[0])
			ptrs[i] = (*C.uint8_t)(unsafe.Pointer(&s[0]))
		}
		lens[i] = C.size_t(len(s))
	}
	res := make([]C.mp3_stream_result, n)
	var base *C.uint8_t
	if rc := C.mp3_decode_batch(e.h, &ptrs[0], &lens[0], C.size_t(n), &res[0], &base, nil); rc != C.MP3_OK {
		return nil, fmt.Errorf("mp3: DecodeBatch: %s", C.GoString(C.mp3_engine_last_error(e.h)))
	}
	out := make([]StreamResult, n)
	for i := range res {
		out[i] = StreamResult{SampleRate: int(res[i].sample_rate), Frames: int64(res[i].frames), Err: errorOf(res[i].status)}
		if res[i].pcm_bytes > 0 {
			out[i].PCM = unsafe.Slice((*byte)(unsafe.Add(unsafe.Pointer(base), res[i].pcm_offset)), int(res[i].pcm_bytes))
		}
	}
	return out, nil
}

// StreamIndex is the frame index of one stream (the reference builds the same table for Seek, decode.go:154-216).
type StreamIndex struct {
	h    *C.mp3_stream_index
	data unsafe.Pointer
}

// NewStreamIndex indexes a stream held in memory.
func NewStreamIndex(stream []byte) (*StreamIndex, error) {
	ix := &StreamIndex{data: C.CBytes(stream)}
	if rc := C.mp3_stream_index_create((*C.uint8_t)(ix.data), C.size_t(len(stream)), &ix.h); rc != C.MP3_OK {
		C.free(ix.data)
		return nil, errorOf(rc)
	}
	runtime.SetFinalizer(ix, (*StreamIndex).Close)
	return ix, nil
}

func (ix *StreamIndex) Close() {
	if ix.h != nil {
		C.mp3_stream_index_free(ix.h)
		C.free(ix.data)
		ix.h = nil
	}
}

// Frames returns the number of frames.
func (ix *StreamIndex) Frames() int64 { return int64(C.mp3_stream_index_frames(ix.h)) }

// DecodeFrames returns the PCM of frames [f0, f1), byte-identical to that stretch of a linear decode, decoded on its
// own on device slot `slot` (lead-in and halo are handled inside the library).
func (e *Engine) DecodeFrames(ix *StreamIndex, f0, f1 int64, slot int) ([]byte, error) {
	want := int64(C.mp3_stream_index_pcm_bytes(ix.h, C.int64_t(f0), C.int64_t(f1)))
	if want < 0 {
		return nil, fmt.Errorf("mp3: frame range [%d, %d) outside the stream", f0, f1)
	}
	if want == 0 {
		return nil, nil
	}
	pcm := make([]byte, want)
	var got C.int64_t
	rc := C.mp3_decode_frames(e.h, C.int(slot), ix.h, C.int64_t(f0), C.int64_t(f1), (*C.uint8_t)(unsafe.Pointer(&pcm[0])), &got)
	if rc == C.MP3_ERR_DEVICE {
		return nil, fmt.Errorf("mp3: DecodeFrames: %s", C.GoString(C.mp3_engine_last_error(e.h)))
	}
	return pcm[:got], errorOf(rc)
}

// DecodeStreamSplit decodes one long stream as one frame range per device of the engine, concurrently; the result is
// byte-identical to the linear decode.  The slice views the engine's pinned buffer (valid until the next batch call).
func (e *Engine) DecodeStreamSplit(ix *StreamIndex) ([]byte, error) {
	var base *C.uint8_t
	var n C.int64_t
	rc := C.mp3_decode_stream_split(e.h, ix.h, &base, &n, nil)
	if rc == C.MP3_ERR_DEVICE || rc == C.MP3_ERR_INVALID {
		return nil, fmt.Errorf("mp3: DecodeStreamSplit: %s", C.GoString(C.mp3_engine_last_error(e.h)))
	}
	if n == 0 {
		return nil, errorOf(rc)
	}
	return unsafe.Slice((*byte)(unsafe.Pointer(base)), int(n)), errorOf(rc)
}
