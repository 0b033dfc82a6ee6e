// Package runner — safetensors reader for B200Runtime.loadWeights: walks every *.safetensors file of an HF checkpoint
// directory and hands each tensor to the engine as bf16 bit patterns under its checkpoint name (hb_model_tensor_set).
// Pure Go (no cgo): format = 8-byte little-endian header length, JSON header {name: {dtype, shape, data_offsets}},
// raw little-endian data.  helix_b200/weights_io.py is the executable mirror (tests/test_weights_io_cpu.py).
package runner

import (
	"encoding/binary"
	"encoding/json"
	"fmt"
	"io"
	"math"
	"os"
	"path/filepath"
	"sort"
)

type stEntry struct {
	DType   string  `json:"dtype"`
	Shape   []int64 `json:"shape"`
	Offsets [2]int64 `json:"data_offsets"`
}

func f32ToBF16(f float32) uint16 {
	u := math.Float32bits(f)
	u += ((u >> 16) & 1) + 0x7FFF // round to nearest even
	return uint16(u >> 16)
}

func f16ToF32(h uint16) float32 {
	sign := uint32(h>>15) << 31
	exp := uint32(h>>10) & 0x1F
	man := uint32(h) & 0x3FF
	switch {
	case exp == 0 && man == 0:
		return math.Float32frombits(sign)
	case exp == 0: // subnormal: normalise
		e := uint32(127 - 15 + 1)
		for man&0x400 == 0 {
			man <<= 1
			e--
		}
		return math.Float32frombits(sign | e<<23 | (man&0x3FF)<<13)
	case exp == 0x1F:
		return math.Float32frombits(sign | 0xFF<<23 | man<<13)
	}
	return math.Float32frombits(sign | (exp+127-15)<<23 | man<<13)
}

// forEachSafetensor calls fn(name, bf16 bits) for every tensor in dir/*.safetensors (sorted by file, then by offset).
func forEachSafetensor(dir string, fn func(name string, bf16 []uint16) error) error {
	files, err := filepath.Glob(filepath.Join(dir, "*.safetensors"))
	if err != nil {
		return err
	}
	if len(files) == 0 {
		return fmt.Errorf("helix-b200: no *.safetensors in %s", dir)
	}
	sort.Strings(files)
	for _, path := range files {
		if err := readSafetensorsFile(path, fn); err != nil {
			return fmt.Errorf("%s: %w", path, err)
		}
	}
	return nil
}

func readSafetensorsFile(path string, fn func(name string, bf16 []uint16) error) error {
	f, err := os.Open(path)
	if err != nil {
		return err
	}
	defer f.Close()
	var lenBuf [8]byte
	if _, err := io.ReadFull(f, lenBuf[:]); err != nil {
		return err
	}
	hlen := int64(binary.LittleEndian.Uint64(lenBuf[:]))
	if hlen <= 0 || hlen > 100<<20 {
		return fmt.Errorf("implausible header length %d", hlen)
	}
	hdr := make([]byte, hlen)
	if _, err := io.ReadFull(f, hdr); err != nil {
		return err
	}
	raw := map[string]json.RawMessage{}
	if err := json.Unmarshal(hdr, &raw); err != nil {
		return err
	}
	type named struct {
		name string
		e    stEntry
	}
	var entries []named
	for name, msg := range raw {
		if name == "__metadata__" {
			continue
		}
		var e stEntry
		if err := json.Unmarshal(msg, &e); err != nil {
			return fmt.Errorf("tensor %s: %w", name, err)
		}
		entries = append(entries, named{name, e})
	}
	sort.Slice(entries, func(i, j int) bool { return entries[i].e.Offsets[0] < entries[j].e.Offsets[0] })
	base := 8 + hlen
	for _, it := range entries {
		n := int64(1)
		for _, d := range it.e.Shape {
			n *= d
		}
		var width int64
		switch it.e.DType {
		case "BF16", "F16":
			width = 2
		case "F32":
			width = 4
		default:
			return fmt.Errorf("tensor %s: dtype %s not supported (BF16 / F16 / F32)", it.name, it.e.DType)
		}
		if it.e.Offsets[1]-it.e.Offsets[0] != n*width || n == 0 {
			return fmt.Errorf("tensor %s: shape and data_offsets disagree", it.name)
		}
		buf := make([]byte, n*width)
		if _, err := f.ReadAt(buf, base+it.e.Offsets[0]); err != nil {
			return err
		}
		out := make([]uint16, n)
		switch it.e.DType {
		case "BF16":
			for i := range out {
				out[i] = binary.LittleEndian.Uint16(buf[2*i:])
			}
		case "F16":
			for i := range out {
				out[i] = f32ToBF16(f16ToF32(binary.LittleEndian.Uint16(buf[2*i:])))
			}
		case "F32":
			for i := range out {
				out[i] = f32ToBF16(math.Float32frombits(binary.LittleEndian.Uint32(buf[4*i:])))
			}
		}
		if err := fn(it.name, out); err != nil {
			return fmt.Errorf("tensor %s: %w", it.name, err)
		}
	}
	return nil
}
