// Package lameinfo is the drop-in for github.com/llehouerou/go-mp3/lameinfo over the C++ mirror behind
// include/mp3host.h (go-mp3_b200/csrc/host/lameinfo.h, tested with the reference's own 16 tests, tests/test_lameinfo.py).
// NOT COMPILED IN THE BUILD IMAGE (no Go toolchain there).  lameinfo is pure metadata and never touches the device; a
// maintainer may equally keep the reference's own pure-Go package — it is unchanged by the drop-in.
package lameinfo

/*
#include "mp3host.h"
*/
import "C"

import (
	"errors"
	"io"
	"unsafe"
)

// Info contains the parsed LAME/Xing header information (lameinfo.go:20-49).
type Info struct {
	IsXing         bool
	Flags          uint32
	FrameCount     uint32
	ByteCount      uint32
	TOC            [100]byte
	VBRScale       uint32
	LAMEVersion    string
	EncoderDelay   uint16
	EncoderPadding uint16
}

// Flag constants for the Flags field (lameinfo.go:52-57).
const (
	FlagFrameCount = 0x0001
	FlagByteCount  = 0x0002
	FlagTOC        = 0x0004
	FlagVBRScale   = 0x0008
)

// DecoderDelay is the standard decoder delay for MP3 decoders (lameinfo.go:86).
const DecoderDelay = 529

// ErrNoXingHeader is returned when no Xing/Info header is found (lameinfo.go:111).
var ErrNoXingHeader = errors.New("lameinfo: no Xing/Info header found")

func (i *Info) HasFrameCount() bool { return i.Flags&FlagFrameCount != 0 }
func (i *Info) HasByteCount() bool  { return i.Flags&FlagByteCount != 0 }
func (i *Info) HasTOC() bool        { return i.Flags&FlagTOC != 0 }
func (i *Info) HasVBRScale() bool   { return i.Flags&FlagVBRScale != 0 }
func (i *Info) HasLAMEInfo() bool   { return i.LAMEVersion != "" }

// TotalDelay returns the samples to skip at the start for gapless playback (lameinfo.go:88-93).
func (i *Info) TotalDelay() int {
	if !i.HasLAMEInfo() {
		return DecoderDelay
	}
	return int(i.EncoderDelay) + DecoderDelay
}

// TotalPadding returns the samples to trim from the end for gapless playback (lameinfo.go:97-108).
func (i *Info) TotalPadding() int {
	if p := int(i.EncoderPadding) - DecoderDelay; i.HasLAMEInfo() && p > 0 {
		return p
	}
	return 0
}

// TOCOffset returns the byte offset (from the first audio frame) at which `fraction` (0..1) of the playing time has passed,
// interpolated in the TOC (coarse seek for sources without a frame index); -1 if the tag has no TOC.  streamBytes is used
// when the tag carries no byte count.  Not part of the reference, which parses the TOC and never uses it.
func (i *Info) TOCOffset(fraction float64, streamBytes uint64) int64 {
	var c C.mp3_lame_info
	c.flags = C.uint32_t(i.Flags)
	c.byte_count = C.uint32_t(i.ByteCount)
	for k := 0; k < 100; k++ {
		c.toc[k] = C.uint8_t(i.TOC[k])
	}
	return int64(C.mp3_lameinfo_toc_offset(&c, C.double(fraction), C.uint64_t(streamBytes)))
}

func fromC(rc C.int, c *C.mp3_lame_info) (*Info, error) {
	switch rc {
	case C.MP3_OK:
	case C.MP3_EOF:
		return nil, io.EOF
	case C.MP3_ERR_UNEXPECTED_EOF:
		return nil, io.ErrUnexpectedEOF
	default:
		return nil, ErrNoXingHeader
	}
	info := &Info{IsXing: c.is_xing != 0, Flags: uint32(c.flags), FrameCount: uint32(c.frame_count), ByteCount: uint32(c.byte_count),
		VBRScale: uint32(c.vbr_scale), EncoderDelay: uint16(c.encoder_delay), EncoderPadding: uint16(c.encoder_padding)}
	for k := range info.TOC {
		info.TOC[k] = byte(c.toc[k])
	}
	if c.has_lame_info != 0 {
		info.LAMEVersion = string(C.GoBytes(unsafe.Pointer(&c.lame_version[0]), 9)) // the 9 bytes as they are, NULs included
	}
	return info, nil
}

// Parse extracts the LAME/Xing header of the first audio frame (lameinfo.go:139-270).
func Parse(frame []byte) (*Info, error) {
	var c C.mp3_lame_info
	var p *C.uint8_t
	if len(frame) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&frame[0]))
	}
	return fromC(C.mp3_lameinfo_parse(p, C.size_t(len(frame)), &c), &c)
}

// ParseFromReader reads the first MP3 frame from a reader positioned at it and parses it (lameinfo.go:288-328).
func ParseFromReader(r io.Reader) (*Info, error) {
	buf := make([]byte, 2881) // the largest Layer I/II/III frame (lameinfo.go:331-386) is 2,880 bytes + padding
	n, err := io.ReadFull(r, buf)
	if err != nil && err != io.ErrUnexpectedEOF && err != io.EOF {
		return nil, err
	}
	var c C.mp3_lame_info
	var p *C.uint8_t
	if n > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&buf[0]))
	}
	return fromC(C.mp3_lameinfo_parse_from_reader(p, C.size_t(n), &c), &c)
}
