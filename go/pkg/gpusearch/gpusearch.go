// Package gpusearch binds libkaamer_gpu.so (include/kaamer_gpu.h) for kaamer's Go host code.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain.  The same C ABI
// is exercised from Python/ctypes by tests/ (every symbol, struct sizes, error behaviour), so
// this file only has to move bytes.  It is what a kaamer maintainer adds under pkg/ to route
// pkg/search through the GPU; see INTEGRATION.md for the three call-site edits.
//
// Replaces, per batch of queries:
//   K_.CreateBytesKey                 pkg/kvstore/k_store.go:66
//   SearchResults.KmerSearch          pkg/search/search.go:414
//   sortMapByValue                    pkg/search/search.go:132
//   QueryResult.FilterResults         pkg/search/search.go:189
//   GetORFs / SetBestStartCodon       pkg/search/dna.go:65,198
//   align.Align                       pkg/align/align.go:46
package gpusearch

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../kaamer_b200 -lkaamer_gpu -Wl,-rpath,${SRCDIR}/../../../kaamer_b200
#include <stdlib.h>
#include "kaamer_gpu.h"
*/
import "C"

import (
	"errors"
	"unsafe"
)

// Index is the device-resident k-mer index: the search-path analogue of the read-only
// kvstore.KVStores opened once at server start (api/server.go:65).
type Index struct{ h *C.kaamer_gpu_t }

func lastErr(rc C.int) error {
	return errors.New("kaamer_gpu: " + C.GoString(C.kaamer_gpu_last_error()))
}

// Open loads a .kidx file exported by `kaamer-db -make ... -gpu-index` (go/pkg/makedb).
func Open(path string, device int) (*Index, error) {
	cp := C.CString(path)
	defer C.free(unsafe.Pointer(cp))
	var h *C.kaamer_gpu_t
	if rc := C.kaamer_gpu_open(cp, C.int(device), &h); rc != C.KAAMER_OK {
		return nil, lastErr(rc)
	}
	return &Index{h}, nil
}

func (ix *Index) Close() { C.kaamer_gpu_close(ix.h); ix.h = nil }

// Options = the SearchOptions fields that reach the hot path (pkg/search/search.go:56-71).
type Options struct {
	MaxResults       int
	MinKMatch        int64
	MinKRatio        float64
	ExtractPositions bool
}

func (o Options) c() C.kaamer_opts {
	var c C.kaamer_opts
	c.min_kmatch = C.int64_t(o.MinKMatch)
	c.min_kratio = C.double(o.MinKRatio)
	c.max_results = C.int32_t(o.MaxResults)
	if o.ExtractPositions {
		c.want_positions = 1
	}
	return c
}

// Hit mirrors search.Hit (pkg/search/search.go:111-115) without the alignment pointer.
type Hit struct {
	Key       uint32
	Kmatch    int64
	Positions []bool // PositionHits[Key] (search.go:442-452), nil unless requested
}

// Row is one query (protein search) or one surviving ORF (nucleotide search).
type Row struct {
	Query      int // index of the query / contig in the batch
	SizeInKmer int
	Hits       []Hit
	// nucleotide rows: Location after SetBestStartCodon and the (trimmed) ORF sequence
	Start, End int
	PlusStrand bool
	Sequence   string
}

// pack concatenates sequences into the (residues, offsets) form of the C ABI.
func pack(seqs []string) ([]byte, []C.uint64_t) {
	n := 0
	for _, s := range seqs {
		n += len(s)
	}
	res := make([]byte, 0, n+1)
	off := make([]C.uint64_t, len(seqs)+1)
	for i, s := range seqs {
		res = append(res, s...)
		off[i+1] = C.uint64_t(len(res))
	}
	if len(res) == 0 {
		res = append(res, 0)
	}
	return res, off
}

func rows(h *C.kaamer_hits, nt bool) []Row {
	n := int(h.n_rows)
	nh := int(h.n_hits)
	hitOff := unsafe.Slice((*uint64)(unsafe.Pointer(h.hit_off)), n+1)
	size := unsafe.Slice((*int32)(unsafe.Pointer(h.size_in_kmer)), n)
	subj := unsafe.Slice((*uint32)(unsafe.Pointer(h.subject_id)), nh)
	km := unsafe.Slice((*uint32)(unsafe.Pointer(h.kmatch)), nh)
	var posOff []uint64
	var pos []byte
	if h.pos_off != nil {
		posOff = unsafe.Slice((*uint64)(unsafe.Pointer(h.pos_off)), nh+1)
		pos = unsafe.Slice((*byte)(unsafe.Pointer(h.pos)), int(posOff[nh]))
	}
	out := make([]Row, n)
	for i := 0; i < n; i++ {
		r := Row{Query: i, SizeInKmer: int(size[i])}
		for k := hitOff[i]; k < hitOff[i+1]; k++ {
			ht := Hit{Key: subj[k], Kmatch: int64(km[k])}
			if posOff != nil {
				p := pos[posOff[k]:posOff[k+1]]
				ht.Positions = make([]bool, len(p))
				for j, b := range p {
					ht.Positions[j] = b != 0
				}
			}
			r.Hits = append(r.Hits, ht)
		}
		if nt {
			contig := unsafe.Slice((*uint32)(unsafe.Pointer(h.row_contig)), n)
			st := unsafe.Slice((*int64)(unsafe.Pointer(h.row_start)), n)
			en := unsafe.Slice((*int64)(unsafe.Pointer(h.row_end)), n)
			pl := unsafe.Slice((*byte)(unsafe.Pointer(h.row_plus)), n)
			so := unsafe.Slice((*uint64)(unsafe.Pointer(h.row_seq_off)), n+1)
			sq := unsafe.Slice((*byte)(unsafe.Pointer(h.row_seq)), int(so[n]))
			r.Query = int(contig[i])
			r.Start, r.End, r.PlusStrand = int(st[i]), int(en[i]), pl[i] != 0
			r.Sequence = string(sq[so[i]:so[i+1]])
		}
		out[i] = r
	}
	return out
}

// SearchProteins replaces the body of the per-query worker loop of ProteinSearch
// (pkg/search/search_protein.go:70-114) for a whole batch.
func (ix *Index) SearchProteins(seqs []string, o Options) ([]Row, error) {
	res, off := pack(seqs)
	co := o.c()
	var h *C.kaamer_hits
	rc := C.kaamer_gpu_search_proteins(ix.h, (*C.uint8_t)(unsafe.Pointer(&res[0])), &off[0],
		C.uint32_t(len(seqs)), &co, &h)
	if rc != C.KAAMER_OK {
		return nil, lastErr(rc)
	}
	defer C.kaamer_gpu_free_hits(h)
	return rows(h, false), nil
}

// SearchNucleotide replaces GetORFs + the per-ORF worker loop of NucleotideSearch / FastqSearch
// (pkg/search/search_nucleotide.go:61-140, search_fastq.go:61-140) for a batch of contigs/reads.
func (ix *Index) SearchNucleotide(contigs []string, o Options) ([]Row, error) {
	res, off := pack(contigs)
	co := o.c()
	var h *C.kaamer_hits
	rc := C.kaamer_gpu_search_nucleotide(ix.h, (*C.uint8_t)(unsafe.Pointer(&res[0])), &off[0],
		C.uint32_t(len(contigs)), &co, &h)
	if rc != C.KAAMER_OK {
		return nil, lastErr(rc)
	}
	defer C.kaamer_gpu_free_hits(h)
	return rows(h, true), nil
}

// Alignment mirrors align.AlignmentResult (pkg/align/align.go:25-40) minus AlnString.
type Alignment struct {
	Identity, Similarity                               float32
	Length, Mismatches, GapOpenings, Raw               int
	BitScore, EValue                                   float64
	QueryStart, QueryEnd, SubjectStart, SubjectEnd     int
	IllegalLetter                                      bool
}

// Align replaces align.Align for pairs (query index, subject protein id); lambda/K/gap values
// come from align.GetMatrixScores (pkg/align/matrixScores.go:110-120).
func (ix *Index) Align(queries []string, pairQuery, pairSubject []uint32, lambda, k float64,
	gapOpen, gapExtend int, numberOfAA uint64) ([]Alignment, error) {
	if len(pairQuery) == 0 {
		return nil, nil
	}
	res, off := pack(queries)
	ao := C.kaamer_aln_opts{lambda: C.double(lambda), K: C.double(k), gap_open: C.int32_t(gapOpen),
		gap_extend: C.int32_t(gapExtend), number_of_aa: C.uint64_t(numberOfAA)}
	out := make([]C.kaamer_aln, len(pairQuery))
	rc := C.kaamer_gpu_align(ix.h, (*C.uint8_t)(unsafe.Pointer(&res[0])), &off[0],
		(*C.uint32_t)(unsafe.Pointer(&pairQuery[0])), (*C.uint32_t)(unsafe.Pointer(&pairSubject[0])),
		C.uint32_t(len(pairQuery)), &ao, &out[0])
	if rc != C.KAAMER_OK {
		return nil, lastErr(rc)
	}
	r := make([]Alignment, len(out))
	for i, a := range out {
		r[i] = Alignment{float32(a.identity), float32(a.similarity), int(a.length), int(a.mismatches),
			int(a.gap_openings), int(a.raw), float64(a.bitscore), float64(a.evalue), int(a.query_start),
			int(a.query_end), int(a.subject_start), int(a.subject_end), a.status != 0}
	}
	return r, nil
}

// AlignText is Align plus AlignmentResult.AlnString (pkg/align/align.go:69-103) of every pair:
// "<gapped query>\n<match line>\n<gapped subject>", what the JSON writer marshals (search.go:497-503).
func (ix *Index) AlignText(queries []string, pairQuery, pairSubject []uint32, lambda, k float64,
	gapOpen, gapExtend int, numberOfAA uint64) ([]Alignment, []string, error) {
	if len(pairQuery) == 0 {
		return nil, nil, nil
	}
	res, off := pack(queries)
	ao := C.kaamer_aln_opts{lambda: C.double(lambda), K: C.double(k), gap_open: C.int32_t(gapOpen),
		gap_extend: C.int32_t(gapExtend), number_of_aa: C.uint64_t(numberOfAA)}
	out := make([]C.kaamer_aln, len(pairQuery))
	var text *C.kaamer_aln_text
	rc := C.kaamer_gpu_align_text(ix.h, (*C.uint8_t)(unsafe.Pointer(&res[0])), &off[0],
		(*C.uint32_t)(unsafe.Pointer(&pairQuery[0])), (*C.uint32_t)(unsafe.Pointer(&pairSubject[0])),
		C.uint32_t(len(pairQuery)), &ao, &out[0], &text)
	if rc != C.KAAMER_OK {
		return nil, nil, lastErr(rc)
	}
	defer C.kaamer_gpu_free_aln_text(text)
	n := len(out)
	offs := unsafe.Slice((*uint64)(unsafe.Pointer(text.off)), n+1)
	blob := unsafe.Slice((*byte)(unsafe.Pointer(text.text)), int(offs[n]))
	r := make([]Alignment, n)
	s := make([]string, n)
	for i, a := range out {
		r[i] = Alignment{float32(a.identity), float32(a.similarity), int(a.length), int(a.mismatches),
			int(a.gap_openings), int(a.raw), float64(a.bitscore), float64(a.evalue), int(a.query_start),
			int(a.query_end), int(a.subject_start), int(a.subject_end), a.status != 0}
		s[i] = string(blob[offs[i]:offs[i+1]])
	}
	return r, s, nil
}

// Ticket names a batch in flight (kaamer_gpu_search_proteins_submit).  The residues live in C memory
// (kaamer_gpu_pinned_alloc) owned by the Ticket until Wait returns: cgo forbids C code to keep Go pointers.
type Ticket struct {
	id  C.int32_t
	res unsafe.Pointer
	off unsafe.Pointer
	nq  int
}

// Submit enqueues a batch and returns at once; at most two batches may be in flight per Index.  A dispatcher
// goroutine calls Submit for batch k+1 before Wait for batch k: the copies and the host work of one batch then
// hide under the kernels of the other (the reference overlaps nbOfThreads queries, api/server.go:55-59).
func (ix *Index) Submit(seqs []string, o Options) (*Ticket, error) {
	res, off := pack(seqs)
	var pres, poff unsafe.Pointer
	if rc := C.kaamer_gpu_pinned_alloc(C.uint64_t(len(res)+16), &pres); rc != C.KAAMER_OK {
		return nil, lastErr(rc)
	}
	if rc := C.kaamer_gpu_pinned_alloc(C.uint64_t(8*len(off)), &poff); rc != C.KAAMER_OK {
		C.kaamer_gpu_pinned_free(pres)
		return nil, lastErr(rc)
	}
	copy(unsafe.Slice((*byte)(pres), len(res)), res)
	copy(unsafe.Slice((*C.uint64_t)(poff), len(off)), off)
	t := &Ticket{res: pres, off: poff, nq: len(seqs)}
	co := o.c()
	if rc := C.kaamer_gpu_search_proteins_submit(ix.h, (*C.uint8_t)(pres), (*C.uint64_t)(poff),
		C.uint32_t(len(seqs)), &co, &t.id); rc != C.KAAMER_OK {
		C.kaamer_gpu_pinned_free(pres)
		C.kaamer_gpu_pinned_free(poff)
		return nil, lastErr(rc)
	}
	return t, nil
}

// Wait blocks until the hits of the batch are in host memory (one event synchronisation).
func (ix *Index) Wait(t *Ticket) ([]Row, error) {
	var hits *C.kaamer_hits
	rc := C.kaamer_gpu_search_proteins_wait(ix.h, t.id, &hits)
	C.kaamer_gpu_pinned_free(t.res)
	C.kaamer_gpu_pinned_free(t.off)
	if rc != C.KAAMER_OK {
		return nil, lastErr(rc)
	}
	defer C.kaamer_gpu_free_hits(hits)
	return rows(hits, false), nil
}

// SetAlignModel replaces the DP model of Align (explicit opt-in, SURVEY §8f-4): matrix in biogo
// alphabet.Protein order "-ABCDEFGHIJKLMNPQRSTVWXYZ*" with row/column 0 = per-residue gap cost, e.g.
// biogo's own matrix.BLOSUM62 flattened, and SWAffine.GapOpen.  nil restores the reference's model.
func (ix *Index) SetAlignModel(matrix [][]int, gapOpen int) error {
	if matrix == nil {
		if rc := C.kaamer_gpu_set_align_model(ix.h, nil); rc != C.KAAMER_OK {
			return lastErr(rc)
		}
		return nil
	}
	var m C.kaamer_aln_model
	for i := 0; i < 26; i++ {
		for j := 0; j < 26; j++ {
			m.matrix[i*26+j] = C.int8_t(matrix[i][j])
		}
	}
	m.gap_open = C.int32_t(gapOpen)
	if rc := C.kaamer_gpu_set_align_model(ix.h, &m); rc != C.KAAMER_OK {
		return lastErr(rc)
	}
	return nil
}

// FormatTSV writes the `-fmt tsv` rows of a whole batch (search.go:505-606) in one native call; the caller
// does one w.Write.  names: Query.Name per query of the batch.
func (ix *Index) FormatTSV(hits *C.kaamer_hits, aln []C.kaamer_aln, names []string, seqOff []C.uint64_t,
	isProtein, withPositions, withAnnotations bool) ([]byte, error) {
	nres, noff := pack(names)
	var alnp *C.kaamer_aln
	if len(aln) > 0 {
		alnp = &aln[0]
	}
	var out *C.char
	var n C.uint64_t
	b2i := func(b bool) C.int {
		if b {
			return 1
		}
		return 0
	}
	nres = append(nres, 0)
	rc := C.kaamer_host_format_tsv(ix.h, hits, alnp, (*C.char)(unsafe.Pointer(&nres[0])), &noff[0], &seqOff[0],
		b2i(isProtein), b2i(withPositions), b2i(withAnnotations), &out, &n)
	if rc != C.KAAMER_OK {
		return nil, lastErr(rc)
	}
	defer C.kaamer_host_free_text(out)
	return C.GoBytes(unsafe.Pointer(out), C.int(n)), nil
}
