// TEST INFRASTRUCTURE (oracle): prints what the UNMODIFIED reference's Read::toBinString() (src/sortmerna/read.cpp:429-462)
// produces for the alignments described on stdin, so tests can pin sortmerna_b200/csrc/smr_blob.cpp against the real
// serializer.  Linked against the reference's own object files (oracle/Makefile.ref, target _ref/blob_ref); contains no
// reference code -- it only includes the reference's headers and calls its classes.
//
// stdin (binary): u32 nreads, u32 slots, i32 num_alignments, u64 ncigar, then nreads x {28-byte smr_read_result},
//                 nreads*slots x {40-byte smr_aln}, ncigar x u32, nreads*4 x u32 (denovo counters)
// stdout (binary): per read u64 length + blob bytes
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "read.hpp"

#pragma pack(push, 1)
struct Res { uint32_t lastIndex, lastPart, hit_seeds, min_index, max_index, n_align; uint16_t max_SW_count; uint8_t is_done, is_hit; };
struct Aln { uint32_t cigar_off, cigar_len, ref_num; int32_t ref_begin1, ref_end1, read_begin1, read_end1; uint32_t readlen; uint16_t score1, part, index_num; uint8_t strand, pad; };
#pragma pack(pop)

static bool rd(void* p, size_t n) { return fread(p, 1, n, stdin) == n; }

int main() {
  uint32_t nreads, slots; int32_t num_alignments; uint64_t ncig;
  if (!rd(&nreads, 4) || !rd(&slots, 4) || !rd(&num_alignments, 4) || !rd(&ncig, 8)) return 1;
  std::vector<Res> res(nreads); std::vector<Aln> alns((size_t)nreads * slots); std::vector<uint32_t> cig(ncig), dn((size_t)nreads * 4);
  if (nreads && !rd(res.data(), res.size() * sizeof(Res))) return 1;
  if (!alns.empty() && !rd(alns.data(), alns.size() * sizeof(Aln))) return 1;
  if (ncig && !rd(cig.data(), ncig * 4)) return 1;
  if (nreads && !rd(dn.data(), dn.size() * 4)) return 1;
  for (uint32_t r = 0; r < nreads; ++r) {
    Read read;
    read.lastIndex = res[r].lastIndex; read.lastPart = res[r].lastPart;
    read.c_yid_ycov = dn[r * 4 + 0]; read.n_yid_ncov = dn[r * 4 + 1]; read.n_nid_ycov = dn[r * 4 + 2]; read.n_denovo = dn[r * 4 + 3];
    read.is_done = res[r].is_done != 0; read.is_hit = res[r].is_hit != 0; read.null_align_output = false;
    read.max_SW_count = res[r].max_SW_count;
    read.num_alignments = num_alignments > 0 ? num_alignments : 0;
    read.hit_seeds = res[r].hit_seeds;
    read.alignment.min_index = res[r].min_index; read.alignment.max_index = res[r].max_index;
    for (uint32_t k = 0; k < res[r].n_align; ++k) {
      const Aln& a = alns[(size_t)r * slots + k];
      s_align2 s;
      s.cigar.assign(cig.begin() + a.cigar_off, cig.begin() + a.cigar_off + a.cigar_len);
      s.ref_num = a.ref_num; s.ref_begin1 = a.ref_begin1; s.ref_end1 = a.ref_end1; s.read_begin1 = a.read_begin1; s.read_end1 = a.read_end1;
      s.readlen = a.readlen; s.score1 = a.score1; s.part = a.part; s.index_num = a.index_num; s.strand = a.strand != 0;
      read.alignment.alignv.push_back(s);
    }
    const std::string b = read.toBinString();
    const uint64_t len = b.size();
    fwrite(&len, 8, 1, stdout);
    if (len) fwrite(b.data(), 1, len, stdout);
  }
  return 0;
}
