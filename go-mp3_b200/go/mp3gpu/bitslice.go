package mp3gpu

// Bit-slice construction: the part of frame.Read that stays in Go.  A maintainer calls AppendFrame once per
// parsed frame, right where maindata.Read used to be called (internal/frame/frame.go:101-108); `si` is the
// *sideinfo.SideInfo of the frame, `own` its main-data bytes (framesize - sideinfo - 4 - CRC,
// maindata.go:88-102).  The C++ mirror of this file, with the same logic and tests against the reference's
// behaviour, is go-mp3_b200/csrc/host/stream_parser.h.

// SideInfo is the subset of internal/sideinfo.SideInfo (sideinfo.go:33-55) the engine needs.
type SideInfo struct {
	MainDataBegin    int
	Scfsi            [2][4]int
	Part2_3Length    [2][2]int
	BigValues        [2][2]int
	GlobalGain       [2][2]int
	ScalefacCompress [2][2]int
	WinSwitchFlag    [2][2]int
	BlockType        [2][2]int
	MixedBlockFlag   [2][2]int
	TableSelect      [2][2][3]int
	SubblockGain     [2][2][3]int
	Region0Count     [2][2]int
	Region1Count     [2][2]int
	Preflag          [2][2]int
	ScalefacScale    [2][2]int
	Count1TableSelect [2][2]int
}

// Header is the subset of frameheader.FrameHeader accessors used here.
type Header struct {
	LSF, Sfreq, Mode, ModeExt, Granules, Channels int
}

// Slicer holds the reservoir state of one stream: the previous frame's logical buffer is the window
// M[winStart:len(M)] (maindata.go:310-322 by induction; SURVEY.md 8b').
type Slicer struct {
	M        []byte
	Units    []Unit
	havePrev bool
	winStart int
}

// Reset is d.frame = nil (decode.go:47,108): the next frame ignores main_data_begin and starts from zero state.
func (s *Slicer) Reset() { s.havePrev = false }

// Rebase drops the bytes in front of the current reservoir window once the units appended so far have been submitted
// (mp3_decoder::fill in csrc/host/mp3host.cc does the same), so that a streaming decoder's buffer stays bounded:
// later units' BitStart are relative to the new M.
func (s *Slicer) Rebase() {
	s.M = append(s.M[:0], s.M[s.winStart:]...)
	s.winStart = 0
	s.Units = s.Units[:0]
}

// AppendFrame resolves the frame's reservoir window and appends its units.  part2Reads(gr, ch) returns the sizes of
// the individual Bits(n) reads of the unit's scalefactors in stream order (a pure function of side info; only
// consulted for zero-length units, whose cursor stays after the scalefactor bits: maindata/huffman.go:29-34).
// Every read is refused on its own when it would cross the buffer end (bits.go:45-60), exactly like
// ScalefacCursor in csrc/host/stream_parser.h — NOT all-or-nothing for the unit.
func (s *Slicer) AppendFrame(h Header, si *SideInfo, own []byte, part2Reads func(gr, ch int) []int) {
	mEnd := len(s.M)
	start := mEnd
	if s.havePrev {
		if si.MainDataBegin > mEnd-s.winStart {
			start = s.winStart // reservoir underflow: bits.Append(prev, own), parsed from bit 0 (maindata.go:295-308)
		} else {
			start = mEnd - si.MainDataBegin
		}
	}
	s.M = append(s.M, own...)
	total := (len(s.M) - start) * 8
	cursor := 0
	for gr := 0; gr < h.Granules; gr++ {
		for ch := 0; ch < 2; ch++ {
			var u Unit
			if ch < h.Channels {
				scfsi := 0
				for b := 0; b < 4; b++ {
					scfsi |= si.Scfsi[ch][b] << b
				}
				u.BitStart = uint64(start*8 + cursor)
				u.BufEndRel = int32(total - cursor)
				u.W0 = uint32(si.Part2_3Length[gr][ch]) | uint32(si.BigValues[gr][ch])<<12 | uint32(si.GlobalGain[gr][ch])<<21 |
					uint32(si.WinSwitchFlag[gr][ch])<<29 | uint32(si.BlockType[gr][ch])<<30
				u.W1 = uint32(si.ScalefacCompress[gr][ch]) | uint32(si.TableSelect[gr][ch][0])<<9 | uint32(si.TableSelect[gr][ch][1])<<14 |
					uint32(si.TableSelect[gr][ch][2])<<19 | uint32(si.Region0Count[gr][ch])<<24 | uint32(si.Region1Count[gr][ch])<<28
				u.W2 = uint32(si.SubblockGain[gr][ch][0]) | uint32(si.SubblockGain[gr][ch][1])<<3 | uint32(si.SubblockGain[gr][ch][2])<<6 |
					uint32(si.Preflag[gr][ch])<<9 | uint32(si.ScalefacScale[gr][ch])<<10 | uint32(si.Count1TableSelect[gr][ch])<<11 |
					uint32(scfsi)<<12 | uint32(h.LSF)<<16 | uint32(h.Sfreq)<<17 | uint32(h.Mode)<<19 | uint32(h.ModeExt)<<21 |
					uint32(gr)<<23 | uint32(ch)<<24 | W2Valid | uint32(si.MixedBlockFlag[gr][ch])<<27
				if !s.havePrev && gr == 0 {
					u.W2 |= W2ZeroState
				}
				if si.Part2_3Length[gr][ch] != 0 {
					cursor += si.Part2_3Length[gr][ch] // SetPos(bitPosEnd+1), maindata/huffman.go:136
				} else {
					for _, n := range part2Reads(gr, ch) {
						if n > 0 && cursor+n <= total {
							cursor += n
						}
					}
				}
			}
			s.Units = append(s.Units, u)
		}
	}
	s.havePrev = true
	s.winStart = start
}
