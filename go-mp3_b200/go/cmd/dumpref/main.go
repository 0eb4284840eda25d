// dumpref — run the UNMODIFIED reference decoder over files and print what the oracle's golden file pins.
//
// Copy this directory into a checkout of github.com/llehouerou/go-mp3 as cmd/dumpref and run
//
//	go run ./cmd/dumpref example/classic_lame.mp3 example/mpeg2.mp3
//
// then compare each line with tests/golden/oracle_pcm.json of this repository ("pcm_sha256", "pcm_bytes",
// "sample_rate").  Equality pins the C oracle (and therefore the exact GPU build, which is bit-identical to
// it) to the real Go implementation on amd64.  No Go toolchain exists in the image this repository was built in.
package main

import (
	"crypto/sha256"
	"fmt"
	"io"
	"os"

	mp3 "github.com/llehouerou/go-mp3"
)

func main() {
	for _, path := range os.Args[1:] {
		f, err := os.Open(path)
		if err != nil {
			fmt.Fprintln(os.Stderr, err)
			os.Exit(1)
		}
		d, err := mp3.NewDecoder(f)
		if err != nil {
			fmt.Printf("%s open_error=%q\n", path, err)
			continue
		}
		pcm, err := io.ReadAll(d)
		fmt.Printf("%s pcm_bytes=%d sample_rate=%d pcm_sha256=%x err=%v\n", path, len(pcm), d.SampleRate(), sha256.Sum256(pcm), err)
		f.Close()
	}
}
