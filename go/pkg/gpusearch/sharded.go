// Sharded index (mode P, DESIGN.md §7) for a kaamer server that drives several GPUs from one
// process: every GPU holds one key range of the .kidx file, all ranges are attached to every
// handle, and each handle then answers whole batches on its own — its search kernels read the
// table entries and posting lists of the other GPUs through NVLink.
//
// NOT COMPILED IN THIS REPOSITORY'S CI (no Go toolchain in the build image); the same C calls are
// exercised by tests/test_gpu_peer.py::test_open_shard_from_kidx_file through ctypes.
package gpusearch

/*
#include <unistd.h>
#include "kaamer_gpu.h"
*/
import "C"

import (
	"sync/atomic"
	"unsafe"
)

// ShardedIndex deals batches round-robin to the per-GPU handles; every handle sees the whole
// key space, so any of them can answer any batch.
type ShardedIndex struct {
	shards []*Index
	next   uint32
}

// OpenSharded loads the key ranges of path onto devices (at most KAAMER_MAX_PEER_SHARDS) and
// attaches them to each other.  Replaces kvstore.KVStoresNew(..., readOnly) (api/server.go:65)
// for databases that do not fit the HBM of one GPU.
func OpenSharded(path string, devices []int) (*ShardedIndex, error) {
	n := len(devices)
	cp := C.CString(path)
	defer C.free(unsafe.Pointer(cp))
	fences := make([]C.uint64_t, n+1)
	if rc := C.kaamer_gpu_kidx_fences(cp, C.int(n), &fences[0]); rc != C.KAAMER_OK {
		return nil, lastErr(rc)
	}
	s := &ShardedIndex{}
	exports := make([]C.kaamer_shard_handle, n)
	closeFds := func() {
		for i := range exports {
			if exports[i].table_fd >= 0 {
				C.close(C.int(exports[i].table_fd))
				C.close(C.int(exports[i].postings_fd))
			}
		}
	}
	for i, dev := range devices {
		var h *C.kaamer_gpu_t
		if rc := C.kaamer_gpu_open_shard(cp, C.int(dev), fences[i], fences[i+1], &h); rc != C.KAAMER_OK {
			s.Close()
			return nil, lastErr(rc)
		}
		s.shards = append(s.shards, &Index{h})
		exports[i].table_fd, exports[i].postings_fd = -1, -1
		if rc := C.kaamer_gpu_shard_export(h, &exports[i]); rc != C.KAAMER_OK {
			closeFds()
			s.Close()
			return nil, lastErr(rc)
		}
	}
	defer closeFds() // same process: shards are attached by pointer, the descriptors are not needed
	for _, ix := range s.shards {
		// the 14.5 GB table is replicated on every GPU, only the posting lists stay sharded
		if rc := C.kaamer_gpu_attach_shards(ix.h, &exports[0], C.int(n), C.KAAMER_ATTACH_REPLICATE_TABLE); rc != C.KAAMER_OK {
			s.Close()
			return nil, lastErr(rc)
		}
	}
	return s, nil
}

func (s *ShardedIndex) pick() *Index {
	return s.shards[int(atomic.AddUint32(&s.next, 1))%len(s.shards)]
}

func (s *ShardedIndex) SearchProteins(seqs []string, o Options) ([]Row, error) {
	return s.pick().SearchProteins(seqs, o)
}

func (s *ShardedIndex) SearchNucleotide(contigs []string, o Options) ([]Row, error) {
	return s.pick().SearchNucleotide(contigs, o)
}

// Close detaches every handle before any shard memory is released.
func (s *ShardedIndex) Close() {
	for _, ix := range s.shards {
		C.kaamer_gpu_detach_shards(ix.h)
	}
	for _, ix := range s.shards {
		ix.Close()
	}
	s.shards = nil
}
