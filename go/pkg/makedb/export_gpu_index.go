// Exporter of the flat device index (.kidx, DESIGN.md §1) from a kaamer-db-built database.
//
// NOT COMPILED HERE (no Go toolchain in the build image).  A maintainer adds this file to
// pkg/makedb and a `-gpu-index <file>` flag to cmd/kaamer-db (cmd/kaamer-db/main.go:132-165).
// It walks kmer_store -> kcomb_store exactly as KmerSearch does per query k-mer
// (pkg/search/search.go:421-429), so the exported posting lists ARE what the Go search sees.
package makedb

import (
	"bufio"
	"encoding/binary"
	"os"

	"github.com/dgraph-io/badger/v3"
	"github.com/golang/protobuf/proto"
	"github.com/zorino/kaamer/pkg/kvstore"
)

type kidxHeader struct {
	Magic                                          [8]byte // "KIDX0001"
	Version, K                                     uint32
	NKeys, NPostings, NProteins, NAA, NKmers       uint64
	MaxProteinID, Flags                            uint32
	NResidues                                      uint64
	Pad                                            [56]byte
}

func pad64(w *bufio.Writer, n int) {
	if r := n % 64; r != 0 {
		w.Write(make([]byte, 64-r))
	}
}

// ExportGPUIndex streams the indexed kmer_store in key order (badger iterates big-endian keys
// in ascending numeric order, pkg/kvstore/k_store.go:69-70) and resolves every comb id.
func ExportGPUIndex(dbPath string, out string) error {
	kv := kvstore.KVStoresNew(dbPath, 4, 1000, false, true)
	defer kv.Close()
	var keys []uint32
	offsets := []uint64{0}
	var postings []uint32
	maxID := uint32(0)
	err := kv.KmerStore.DB.View(func(txn *badger.Txn) error {
		it := txn.NewIterator(badger.DefaultIteratorOptions)
		defer it.Close()
		for it.Rewind(); it.Valid(); it.Next() {
			item := it.Item()
			k := item.Key()
			if len(k) != 4 {
				continue
			}
			combID, err := item.ValueCopy(nil)
			if err != nil || len(combID) < 1 {
				continue // search.go:421-425
			}
			val, err := kv.KCombStore.GetValueFromBadger(combID)
			if err != nil {
				continue
			}
			kc := &kvstore.KComb{}
			proto.Unmarshal(val, kc)
			keys = append(keys, binary.BigEndian.Uint32(k))
			for _, id := range kc.ProteinKeys { // already unique, descending (kv_store.go:284-305)
				postings = append(postings, id)
				if id > maxID {
					maxID = id
				}
			}
			offsets = append(offsets, uint64(len(postings)))
		}
		return nil
	})
	if err != nil {
		return err
	}
	// KStats (ProteinStore["db_stats"], pkg/makedb/inputFASTA.go:166-178)
	stats := kvstore.KStats{}
	if v, e := kv.ProteinStore.GetValueFromBadger([]byte("db_stats")); e == nil {
		proto.Unmarshal(v, &stats)
	}
	// protein table indexed by id (subject sequences for the alignment stage)
	protOff := make([]uint64, int(maxID)+2)
	var residues []byte
	for id := uint32(0); id <= maxID; id++ {
		key := make([]byte, 4)
		binary.BigEndian.PutUint32(key, id)
		if v, e := kv.ProteinStore.GetValueFromBadger(key); e == nil {
			p := &kvstore.Protein{}
			proto.Unmarshal(v, p)
			residues = append(residues, p.Sequence...)
		}
		protOff[id+1] = uint64(len(residues))
	}
	f, err := os.Create(out)
	if err != nil {
		return err
	}
	defer f.Close()
	w := bufio.NewWriterSize(f, 1<<20)
	defer w.Flush()
	h := kidxHeader{Version: 1, K: 7, NKeys: uint64(len(keys)), NPostings: uint64(len(postings)),
		NProteins: stats.NumberOfProteins, NAA: stats.NumberOfAA, NKmers: stats.NumberOfKmers,
		MaxProteinID: maxID, Flags: 1, NResidues: uint64(len(residues))}
	copy(h.Magic[:], "KIDX0001")
	binary.Write(w, binary.LittleEndian, &h)
	binary.Write(w, binary.LittleEndian, keys)
	pad64(w, 4*len(keys))
	binary.Write(w, binary.LittleEndian, offsets)
	pad64(w, 8*len(offsets))
	binary.Write(w, binary.LittleEndian, postings)
	pad64(w, 4*len(postings))
	binary.Write(w, binary.LittleEndian, protOff)
	pad64(w, 8*len(protOff))
	w.Write(residues)
	pad64(w, len(residues))
	return nil
}
