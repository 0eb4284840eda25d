// Package mp3gpu is the cgo binding of the B200 granule-decode engine (include/mp3gpu.h).
//
// NOT COMPILED IN THE BUILD IMAGE (no Go toolchain there); it is the binding a go-mp3 maintainer adds as
// internal/mp3gpu.  Build: CGO_CFLAGS=-I<repo>/include CGO_LDFLAGS="-L<repo>/go-mp3_b200 -lmp3gpu".
//
// It replaces, for a whole batch of frames at once, what Decoder.readFrame does per frame after the
// serial parse (decode.go:45-67): the scalefactor/Huffman half of maindata.Read
// (internal/maindata/maindata.go:119-288, internal/maindata/huffman.go:27-138) and (*Frame).Decode()
// (internal/frame/frame.go:121-138).
package mp3gpu

/*
#include <stdlib.h>
#include "mp3gpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"unsafe"
)

// Unit mirrors mp3gpu_unit (32 bytes): one (granule, channel) bit-slice plus its side info.
type Unit struct {
	BitStart  uint64 // absolute bit index into mainData of the unit's first part2 bit
	BufEndRel int32  // (end of the frame's logical reservoir buffer) - BitStart, in bits
	W0        uint32 // part2_3_length:12 | big_values:9 | global_gain:8 | win_switch:1 | block_type:2
	W1        uint32 // scalefac_compress:9 | table_select[3]:5 each | region0_count:4 | region1_count:4
	W2        uint32 // subblock_gain[3]:3 each | preflag | scalefac_scale | count1table_select | scfsi:4 |
	//                  lsf | sfreq:2 | mode:2 | mode_ext:2 | gr | ch | valid | zero_state | mixed_block_flag
	_ [2]uint32
}

const (
	W2Valid     = 1 << 25
	W2ZeroState = 1 << 26
	// BytesPerGranule is the PCM produced per granule: 576 stereo samples x 2 channels x int16.
	BytesPerGranule = 2304
)

// Engine owns one CUDA device context (tables, streams, workspace).  Not safe for concurrent use,
// like *mp3.Decoder (decode.go:31-33).
type Engine struct{ ctx *C.mp3gpu_ctx }

// NewEngine creates an engine on a CUDA device.  There is no CPU fallback: it fails without a GPU.
func NewEngine(device int) (*Engine, error) {
	var ctx *C.mp3gpu_ctx
	opts := C.mp3gpu_opts{abi_version: C.MP3GPU_ABI_VERSION}
	if rc := C.mp3gpu_create(C.int(device), &opts, &ctx); rc != C.MP3GPU_OK {
		return nil, fmt.Errorf("mp3: mp3gpu_create failed (%d)", int(rc))
	}
	return &Engine{ctx: ctx}, nil
}

func (e *Engine) Close() { C.mp3gpu_destroy(e.ctx); e.ctx = nil }

// Decode runs K1..K4 for len(units)/2 granules.  mainData must carry 64 bytes of zero padding after its
// logical length n; pcm receives len(units)/2 * BytesPerGranule bytes (16-bit LE, L R interleaved).
// cgo rule: the C side does not retain the Go pointers beyond the call.
func (e *Engine) Decode(mainData []byte, n int, units []Unit, pcm []byte) error {
	g := len(units) / 2
	if g == 0 {
		return nil
	}
	if len(pcm) < g*BytesPerGranule || len(mainData) < n+64 {
		return errors.New("mp3: short buffer")
	}
	rc := C.mp3gpu_decode(e.ctx, (*C.uint8_t)(unsafe.Pointer(&mainData[0])), C.size_t(n),
		(*C.mp3gpu_unit)(unsafe.Pointer(&units[0])), C.size_t(g), (*C.int16_t)(unsafe.Pointer(&pcm[0])))
	if rc != C.MP3GPU_OK {
		return fmt.Errorf("mp3: device decode failed: %s", C.GoString(C.mp3gpu_last_error(e.ctx)))
	}
	return nil
}
