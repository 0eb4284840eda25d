// bench_batch_test.go — the CPU baseline the north star names, for a Go-equipped reviewer: the reference decoder,
// one goroutine per stream across all cores (NewDecoder + io.ReadAll, as bench_test.go:40-55), reported in the
// unit bench.py uses (stereo Msamples/s).  Drop into a checkout of the reference next to bench_test.go; the
// stream files are written by `python tools/synth/dump_streams.py DIR N` in this repository.
package mp3

import (
	"bytes"
	"io"
	"os"
	"path/filepath"
	"runtime"
	"sync"
	"testing"
)

func BenchmarkDecodeBatch(b *testing.B) {
	files, _ := filepath.Glob(filepath.Join(os.Getenv("MP3_STREAM_DIR"), "*.mp3"))
	if len(files) == 0 {
		b.Skip("set MP3_STREAM_DIR")
	}
	var streams [][]byte
	for _, f := range files {
		d, err := os.ReadFile(f)
		if err != nil {
			b.Fatal(err)
		}
		streams = append(streams, d)
	}
	b.ResetTimer()
	var samples int64
	for i := 0; i < b.N; i++ {
		var wg sync.WaitGroup
		var mu sync.Mutex
		work := make(chan []byte)
		for w := 0; w < runtime.NumCPU(); w++ {
			wg.Add(1)
			go func() {
				defer wg.Done()
				for s := range work {
					d, err := NewDecoder(bytes.NewReader(s))
					if err != nil {
						continue
					}
					pcm, _ := io.ReadAll(d)
					mu.Lock()
					samples += int64(len(pcm) / 4)
					mu.Unlock()
				}
			}()
		}
		for _, s := range streams {
			work <- s
		}
		close(work)
		wg.Wait()
	}
	b.ReportMetric(float64(samples)/b.Elapsed().Seconds()/1e6, "Msamples/s")
	b.ReportMetric(float64(runtime.NumCPU()), "cores")
}
