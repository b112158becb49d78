// kaamer-golden — dumps reference outputs of the search hot path as JSON golden vectors.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain, and the reference ships no
// tests or golden vectors for this path (SURVEY.md §4), which is why the parity of the CUDA path is
// pinned on a CPU restatement only.  This command is the one step that turns "parity unpinned" into
// "parity pinned": run it once on any box with Go >= 1.12 and the kaamer module cache,
//
//     cd <kaamer checkout> && cp -r <this repo>/go/cmd/kaamer-golden cmd/ && \
//     go run ./cmd/kaamer-golden > <this repo>/tests/golden/ref_vectors.json
//
// and commit the file: tests/test_reference_vectors.py (skipped while the file is absent) checks the
// oracle AND the GPU path against every vector.  It calls the reference's own functions, unmodified:
//   kvstore.NewAATable / K_.EncodeKmer            pkg/kvstore/k_store.go:39-117
//   search.GetORFs / SetBestStartCodon            pkg/search/dna.go:65-272
//   QueryResult.FilterResults                     pkg/search/search.go:189-220
//   search.FormatPositionsToString                pkg/search/search.go:694-742
//   align.Align (biogo SWAffine, BLOSUM62, -11)   pkg/align/align.go:46-161
//   biogo matrix.BLOSUM62 itself (the gap row that decides kaamer_gpu_set_align_model's default)
package main

import (
	"encoding/json"
	"os"

	"github.com/biogo/biogo/align/matrix"
	"github.com/zorino/kaamer/pkg/align"
	"github.com/zorino/kaamer/pkg/kvstore"
	"github.com/zorino/kaamer/pkg/search"
)

type alnVec struct {
	Query, Subject string
	NumberOfAA     uint64
	Result         align.AlignmentResult
}

type orfVec struct {
	DNA  string
	ORFs []search.ORF
}

type golden struct {
	Blosum62     [][]int // biogo matrix.BLOSUM62, order "-ABCDEFGHIJKLMNPQRSTVWXYZ*": row 0 = gap costs
	EncodeKmer   map[string]uint32
	ORFs         []orfVec
	Alignments   []alnVec
	Positions    map[string]string
	FilterCounts []int
}

func main() {
	g := golden{EncodeKmer: map[string]uint32{}, Positions: map[string]string{}}
	g.Blosum62 = matrix.BLOSUM62

	k := kvstore.K_{}
	k.NewAATable() // (method set as in pkg/kvstore/k_store.go; adjust if the receiver differs)
	for _, kmer := range []string{"AAAAAAA", "YYYYYYY", "WWWWWWW", "ACDEFGH", "MKTAYIA", "MELPNIM", "AAAAAAX", "AXAAAAA", "AAAAAA*"} {
		g.EncodeKmer[kmer] = k.EncodeKmer(kmer)
	}

	for _, dna := range []string{
		"atg" + repeat("gct", 20) + "taa",
		"ttgacgtnacgtaaatgcccgggtttaaacccgggtttaaaatgaaatttcccgggaaataa" + repeat("gat", 30) + "tag",
	} {
		g.ORFs = append(g.ORFs, orfVec{DNA: dna, ORFs: search.GetORFs(dna, 11)})
	}

	stats := kvstore.KStats{NumberOfAA: 3500000}
	pairs := [][2]string{
		{"MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAPILSRVGDGTQDNLSGAEKAVQ", "MKTAYIAKQRQISFVKSHFSRQAPILSRVGDGTQDNLSGAEKAVQVKVKALPDAQFEVV"},
		{"MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ", "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ"},
		{"MKTUYIAKQRQISFVKSHFSRQ", "MKTUUIAKQRQISFuKSHFSRQ*"},
		{"WWWWWWWWWWWWWW", "AAAAAAAAAAAAAAAAAAAAAAAA"},
		{"MKTAYIAKQRQISFVKSHFSRQ1", "MKTAYIAKQRQISFVKSHFSRQ"},
	}
	for _, p := range pairs {
		r, _ := align.Align(p[0], p[1], stats, "blosum62", 11, 1)
		g.Alignments = append(g.Alignments, alnVec{p[0], p[1], stats.NumberOfAA, r})
	}

	for name, pos := range map[string][]bool{
		"run_to_end": {false, true, true, false, true, true, true},
		"single":     {false, true, false, false},
		"all":        {true, true, true, true},
	} {
		g.Positions[name] = search.FormatPositionsToString(pos, false)
		g.Positions[name+"_aln"] = search.FormatPositionsToString(pos, true)
	}

	enc := json.NewEncoder(os.Stdout)
	enc.SetIndent("", " ")
	enc.Encode(g)
}

func repeat(s string, n int) string {
	out := ""
	for i := 0; i < n; i++ {
		out += s
	}
	return out
}
