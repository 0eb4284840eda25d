// Package mp3 is the drop-in for github.com/llehouerou/go-mp3's package mp3 on top of the B200 engine.
//
// NOT COMPILED IN THE BUILD IMAGE (no Go toolchain there).  Build with
//   CGO_CFLAGS=-I<repo>/include  CGO_LDFLAGS="-L<repo>/go-mp3_b200 -lmp3host -Wl,-rpath,<repo>/go-mp3_b200"
//
// It keeps the reference's exported surface — NewDecoder, (*Decoder).Read / Seek / SampleRate / Length /
// BytesPerFrame / Duration / Position / Remaining / Progress / SamplePosition / SampleCount / SeekToSample / Skip /
// SeekToTime (decode.go:70-341, 361-388), the always-16-bit-stereo output contract (decode.go:356-360) — and adds
// DecodeBatch, DecodeFrames and DecodeStreamSplit (engine.go).  The serial stream work (tags, header sync, side info,
// reservoir resolution, frame index) is done by the validated host stage behind include/mp3host.h
// (go-mp3_b200/csrc/host/, tested against the oracle on fixtures, synthetic, fuzzed and malformed streams); the
// pure-Go sketch of that stage is go/mp3gpu/bitslice.go.  Every PCM byte comes from the CUDA kernels: there is no CPU
// decode path, and NewDecoder fails without a GPU.
package mp3

/*
#include <stdlib.h>
#include <string.h>
#include "mp3host.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"runtime"
	"time"
	"unsafe"
)

// Errors of the reference that callers can test for.
var (
	// ErrSeekNotSupported mirrors errors.New("mp3: seek not supported on non-seekable source") (decode.go:291,323).
	ErrSeekNotSupported = errors.New("mp3: seek not supported on non-seekable source")
	// ErrInvalidWhence mirrors errors.New("mp3: invalid whence") (decode.go:104).
	ErrInvalidWhence = errors.New("mp3: invalid whence")
)

// errorOf maps a status code of include/mp3host.h to the reference's error value (decode.go:48-63: the EOF family is
// io.EOF; everything else carries the reference's message).
func errorOf(code C.int) error {
	switch code {
	case C.MP3_OK:
		return nil
	case C.MP3_EOF, C.MP3_ERR_UNEXPECTED_EOF, C.MP3_ERR_SYNC_LIMIT:
		return io.EOF
	case C.MP3_ERR_SEEK_UNSUPPORTED:
		return ErrSeekNotSupported
	case C.MP3_ERR_WHENCE:
		return ErrInvalidWhence
	}
	return errors.New(C.GoString(C.mp3_error_string(code)))
}

// Decoder is a MP3-decoded stream.  It decodes its MP3 source and its Read yields 16-bit little-endian, 2-channel
// PCM (decode.go:18-43).  Not safe for concurrent use, like the reference's (decode.go:31-33).
type Decoder struct {
	eng  *Engine
	h    *C.mp3_decoder
	data unsafe.Pointer // the source's bytes in C memory: the host stage keeps pointers into them
}

// NewDecoder decodes the given io.Reader and returns a decoded stream (decode.go:361-388).  The source is read to its
// end up front (the host stage parses ahead of Read by opts.chunk_frames frames per GPU call); a source that
// implements io.Seeker gives a seekable Decoder with a Length, anything else a Decoder whose Length is -1 and whose
// Seek methods fail, exactly as the reference's do.  Uses the package-level default engine (device 0).
func NewDecoder(r io.Reader) (*Decoder, error) {
	e, err := DefaultEngine()
	if err != nil {
		return nil, err
	}
	return e.NewDecoder(r, 0)
}

// NewDecoder is NewDecoder on device slot `slot` of this engine.
func (e *Engine) NewDecoder(r io.Reader, slot int) (*Decoder, error) {
	_, seekable := r.(io.Seeker)
	src, err := io.ReadAll(r)
	if err != nil {
		return nil, err
	}
	d := &Decoder{eng: e}
	if len(src) > 0 {
		d.data = C.malloc(C.size_t(len(src)))
		C.memcpy(d.data, unsafe.Pointer(&src[0]), C.size_t(len(src)))
	}
	var code C.int
	sk := C.int(0)
	if seekable {
		sk = 1
	}
	d.h = C.mp3_new_decoder_on(e.h, C.int(slot), (*C.uint8_t)(d.data), C.size_t(len(src)), sk, &code)
	if d.h == nil {
		C.free(d.data)
		if code == C.MP3_ERR_DEVICE {
			return nil, fmt.Errorf("mp3: %s", C.GoString(C.mp3_engine_last_error(e.h)))
		}
		return nil, errorOf(code)
	}
	runtime.SetFinalizer(d, (*Decoder).Close)
	return d, nil
}

// Close releases the decoder's native state (the reference's Decoder has none; calling it is optional).
func (d *Decoder) Close() {
	if d.h != nil {
		C.mp3_decoder_free(d.h)
		C.free(d.data)
		d.h, d.data = nil, nil
	}
}

// Read is io.Reader's Read (decode.go:70-80): at most the rest of one frame per call.
func (d *Decoder) Read(buf []byte) (int, error) {
	if len(buf) == 0 {
		return 0, nil
	}
	var code C.int
	n := C.mp3_decoder_read(d.h, (*C.uint8_t)(unsafe.Pointer(&buf[0])), C.size_t(len(buf)), &code)
	if n == 0 {
		return 0, errorOf(code)
	}
	return int(n), nil
}

// Seek is io.Seeker's Seek (decode.go:89-145).  Seek panics on a non-seekable source in the reference; here it
// returns ErrSeekNotSupported.
func (d *Decoder) Seek(offset int64, whence int) (int64, error) {
	var code C.int
	pos := C.mp3_decoder_seek(d.h, C.int64_t(offset), C.int(whence), &code)
	if code != C.MP3_OK {
		return 0, errorOf(code)
	}
	return int64(pos), nil
}

// SampleRate returns the sample rate like 44100 (decode.go:150-152).
func (d *Decoder) SampleRate() int { return int(C.mp3_decoder_sample_rate(d.h)) }

// Length returns the total size in bytes, -1 when the source is not seekable (decode.go:224-226).
func (d *Decoder) Length() int64 { return int64(C.mp3_decoder_length(d.h)) }

// BytesPerFrame returns the decoded bytes per MP3 frame (decode.go:230-232).
func (d *Decoder) BytesPerFrame() int64 { return int64(C.mp3_decoder_bytes_per_frame(d.h)) }

// Duration returns the total duration, -1 when the source is not seekable (decode.go:236-241).
func (d *Decoder) Duration() time.Duration { return time.Duration(C.mp3_decoder_duration_ns(d.h)) }

// Position returns the current playback position (decode.go:244-246).
func (d *Decoder) Position() time.Duration { return time.Duration(C.mp3_decoder_position_ns(d.h)) }

// Remaining returns the remaining duration, -1 when the source is not seekable (decode.go:250-256).
func (d *Decoder) Remaining() time.Duration { return time.Duration(C.mp3_decoder_remaining_ns(d.h)) }

// Progress returns the playback progress in [0, 1], -1 when the source is not seekable (decode.go:260-268).
func (d *Decoder) Progress() float64 { return float64(C.mp3_decoder_progress(d.h)) }

// SamplePosition returns the current position in samples (decode.go:272-274).
func (d *Decoder) SamplePosition() int64 { return int64(C.mp3_decoder_sample_position(d.h)) }

// SampleCount returns the total number of samples, -1 when the source is not seekable (decode.go:278-283).
func (d *Decoder) SampleCount() int64 { return int64(C.mp3_decoder_sample_count(d.h)) }

// SeekToSample seeks to a sample position, clamped to the stream (decode.go:288-307).
func (d *Decoder) SeekToSample(sample int64) error {
	return errorOf(C.mp3_decoder_seek_to_sample(d.h, C.int64_t(sample)))
}

// Skip moves the position by a duration, negative to go back (decode.go:313-315).
func (d *Decoder) Skip(delta time.Duration) error {
	return errorOf(C.mp3_decoder_skip(d.h, C.int64_t(delta)))
}

// SeekToTime seeks to a time position, clamped to the stream and aligned to a sample (decode.go:320-341).
func (d *Decoder) SeekToTime(t time.Duration) error {
	return errorOf(C.mp3_decoder_seek_to_time(d.h, C.int64_t(t)))
}
