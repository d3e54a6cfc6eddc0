// Native index builder (host): see smr_build.cpp.  Stands in for build_index (src/sortmerna/indexdb.cpp:1119-2095).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace smr {

struct BuildOptions {
  uint32_t lnwin = 18;      // -L   (options.hpp: seed_win_len)
  uint32_t interval = 1;    // -interval (options.hpp:588)
  uint32_t max_pos = 10000; // -max_pos  (options.hpp:589); 0 = keep every position
  double max_mb = 3072;     // -m        (options.hpp:586): estimated MB per index part
  uint32_t threads = 0;     // T: 2T worker threads (forward / reverse tries of the 9-mers k % T == t); 0 = cores / 8, 1..8
};

struct BuildReport {
  uint32_t parts = 0;
  uint64_t numseq = 0, windows = 0, unique_lmers = 0, trie_nodes = 0, bytes_written = 0;
};

// One record of the reference FASTA as the builder sees it.
struct RefRecord { std::string name; size_t rec_start = 0, rec_end = 0; std::vector<uint8_t> seq, seq04; };
std::string parse_reference_fasta(const std::string& fasta, uint32_t pread, bool want04, std::vector<RefRecord>& recs, double freq[4], uint64_t& full_len,
                                  size_t& file_size);
std::string next_index_part(const std::vector<RefRecord>& recs, size_t first, uint32_t pread, double max_mb, std::vector<size_t>& members, size_t& next,
                            uint64_t& start_part, uint64_t& seq_part_size);

// Writes <prefix>.kmer_P.dat, .bursttrie_P.dat, .pos_P.dat for every part P and <prefix>.stats.
// Returns the empty string on success, else the error text (the reference prints it and exits).
std::string build_index_files(const std::string& fasta, const std::string& prefix, const BuildOptions& opt, BuildReport* report);

}  // namespace smr
