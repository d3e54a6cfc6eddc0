// Reference-side binding of libsmr_b200.so: a replacement for the reference's
//   void align(Readfeed&, Readstats&, Index&, KeyValueDatabase&, Runopts&)          (src/sortmerna/processor.cpp:173-285)
// written against the reference's own headers and nothing else -- the translation unit a maintainer adds to the host program
// (INTEGRATION.md).  Everything around it stays the reference's: CLI / Runopts, Readfeed, Refstats (ALP, minimal_score),
// References, KeyValueDatabase, writeSummary / writeReports / denovo_stats.  oracle/Makefile.ref links it with the reference's
// unmodified objects into oracle/_ref/sortmerna_gpu (the symbol `align` of processor.o is weakened, this one wins).
//
// Differences to the CPU driver, all invisible to the rest of the program:
//   * every (index, part) is made resident on the GPU once; reads are streamed ONCE (the reference re-reads them per index part);
//   * reads are handed over in batches; the per-read KVDB blob is produced by smr_pack_kvdb_blobs (byte-identical to
//     Read::toBinString) and stored under the same key (read.id);
//   * Readstats counters come back from the library (num_aligned, reads_matched_per_db, num_short of the last index pass).
#include <cstdint>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "common.hpp"
#include "index.hpp"
#include "kvdb.hpp"
#include "options.hpp"
#include "read.hpp"
#include "readfeed.hpp"
#include "readstats.hpp"
#include "references.hpp"
#include "refstats.hpp"
#include "smr_b200.h"

namespace {

std::vector<char> slurp(const std::string& path) {
  std::ifstream in(path, std::ios::binary | std::ios::ate);
  if (!in) { ERR("cannot open index file ", path); exit(EXIT_FAILURE); }
  std::vector<char> v((size_t)in.tellg());
  in.seekg(0);
  if (!v.empty()) in.read(v.data(), (std::streamsize)v.size());
  return v;
}

void die(smr_ctx* ctx, const char* what) {
  ERR(what, ": ", smr_last_error(ctx));
  exit(EXIT_FAILURE);
}

}  // namespace

void align(Readfeed& readfeed, Readstats& readstats, Index& /*index: the library keeps its own resident form*/, KeyValueDatabase& kvdb, Runopts& opts)
{
  INFO("==== Starting alignment (libsmr_b200) ====");
  if (opts.num_alignments == 0) { ERR("'-num_alignments 0' is not supported by the GPU path"); exit(EXIT_FAILURE); }
  // one context per GPU (SMR_GPUS, default 1; never more than the devices present): reads shard by record, every GPU holds the
  // whole index, the only cross-GPU state are the Readstats counters (summed below) -- SURVEY 8(e)
  int ngpu = 1;
  if (const char* e = getenv("SMR_GPUS")) ngpu = std::max(1, std::min(atoi(e), smr_device_count()));
  std::vector<smr_ctx*> ctxs((size_t)ngpu, nullptr);
  for (int g = 0; g < ngpu; ++g)
    if (smr_init(g, &ctxs[g]) != SMR_OK) { ERR("no usable CUDA device ", g, ": the GPU alignment path has no CPU fallback"); exit(EXIT_FAILURE); }

  Refstats refstats(opts, readstats);            // unchanged: .stats, Gumbel parameters, minimal_score (refstats.cpp:103-276)
  References refs;
  smr_params p{};
  p.match = opts.match; p.mismatch = opts.mismatch; p.score_N = opts.score_N; p.gap_open = opts.gap_open; p.gap_ext = opts.gap_extension;
  p.num_seeds = (int32_t)opts.num_seeds; p.min_lis = (int32_t)opts.min_lis; p.edges = (int32_t)opts.edges; p.edges_is_percent = opts.is_as_percent ? 1 : 0;
  p.num_alignments = (int32_t)opts.num_alignments; p.is_best = opts.is_best ? 1 : 0;
  p.is_forward = opts.is_forward ? 1 : 0; p.is_reverse = opts.is_reverse ? 1 : 0; p.is_full_search = opts.is_full_search ? 1 : 0;
  p.minoccur = (int32_t)opts.minoccur;
  for (smr_ctx* ctx : ctxs) if (smr_set_params(ctx, &p) != SMR_OK) die(ctx, "smr_set_params");

  // every (index, part) becomes resident once per GPU (the reference loads / unloads them one at a time, processor.cpp:216-262)
  for (size_t i = 0; i < opts.indexfiles.size(); ++i)
    for (uint16_t part = 0; part < refstats.num_index_parts[i]; ++part) {
      const std::string pfx = opts.indexfiles[i].second, sfx = "_" + std::to_string(part) + ".dat";
      const std::vector<char> kmer = slurp(pfx + ".kmer" + sfx), trie = slurp(pfx + ".bursttrie" + sfx), pos = slurp(pfx + ".pos" + sfx);
      refs.load((uint32_t)i, part, opts, refstats);                    // unchanged References::load: sequences in the 0-4 alphabet
      std::string cat; std::vector<uint64_t> off(1, 0);
      for (auto& r : refs.buffer) { cat += r.sequence; off.push_back(cat.size()); }
      const uint32_t skip[3] = {opts.skiplengths[i][0], opts.skiplengths[i][1], opts.skiplengths[i][2]};
      for (smr_ctx* ctx : ctxs)
        if (smr_load_index_part(ctx, (uint32_t)i, part, kmer.data(), kmer.size(), trie.data(), trie.size(), pos.data(), pos.size(),
                                (const uint8_t*)cat.data(), off.data(), (uint32_t)refs.buffer.size(), refstats.lnwin[i], refstats.minimal_score[i], skip) != SMR_OK)
          die(ctx, "smr_load_index_part");
      refs.unload();
    }

  // batches of reads from the unchanged Readfeed; read ids ("<file>_<n>") stay the KVDB keys.  Up to one batch per GPU is in
  // flight; results are stored in batch order.
  const uint32_t slots = (uint32_t)std::max<int32_t>(1, (int32_t)opts.num_alignments);
  const size_t nrefs = opts.indexfiles.size();
  uint32_t batch_reads = 1u << 20;
  if (const char* e = getenv("SMR_BATCH_READS")) batch_reads = (uint32_t)std::max(1, atoi(e));
  struct Batch {
    std::vector<std::string> ids; std::string seqcat; std::vector<uint64_t> off{0};
    std::vector<smr_read_result> res; std::vector<smr_aln> alns; std::vector<uint32_t> cigars; uint64_t used = 0; std::vector<uint64_t> cnt;
    std::string blobs; std::vector<uint64_t> boff;
  };
  std::vector<Batch> pending;
  pending.emplace_back();
  std::vector<uint64_t> total(SMR_CNT_FIXED + nrefs, 0);
  auto run_batch = [&](smr_ctx* ctx, Batch& b) {
    const uint32_t n = (uint32_t)b.ids.size();
    b.res.resize(n); b.alns.resize((size_t)n * slots); b.cigars.resize((size_t)64 * n * slots + 4096); b.cnt.assign(SMR_CNT_FIXED + nrefs, 0);
    if (smr_align_batch(ctx, (const uint8_t*)b.seqcat.data(), b.off.data(), n, b.res.data(), b.alns.data(), b.cigars.data(), b.cigars.size(), &b.used,
                        b.cnt.data(), (uint32_t)b.cnt.size()) != SMR_OK) die(ctx, "smr_align_batch");
    b.boff.resize((size_t)n + 1);
    smr_pack_kvdb_blobs(b.res.data(), b.alns.data(), b.cigars.data(), n, slots, (int32_t)opts.num_alignments, nullptr, nullptr, 0, b.boff.data());
    b.blobs.assign(b.boff[n], '\0');
    if (smr_pack_kvdb_blobs(b.res.data(), b.alns.data(), b.cigars.data(), n, slots, (int32_t)opts.num_alignments, nullptr, (uint8_t*)&b.blobs[0], b.blobs.size(),
                            b.boff.data()) != SMR_OK) die(ctx, "smr_pack_kvdb_blobs");
  };
  auto flush = [&]() {                       // run the pending batches (one per GPU, concurrently), then store their results in order
    if (pending.back().ids.empty()) pending.pop_back();
    if (pending.empty()) { pending.emplace_back(); return; }
    std::vector<std::thread> workers;
    for (size_t k = 1; k < pending.size(); ++k) workers.emplace_back(run_batch, ctxs[k], std::ref(pending[k]));
    run_batch(ctxs[0], pending[0]);
    for (auto& t : workers) t.join();
    for (Batch& b : pending) {
      for (uint32_t r = 0; r < (uint32_t)b.ids.size(); ++r)
        if (b.boff[r + 1] > b.boff[r]) kvdb.put(b.ids[r], b.blobs.substr(b.boff[r], b.boff[r + 1] - b.boff[r]));   // == kvdb.put(read.id, read.toBinString())
      for (size_t k = 0; k < b.cnt.size(); ++k) total[k] += b.cnt[k];
    }
    pending.clear();
    pending.emplace_back();
  };

  readfeed.init_reading();
  std::string readstr;
  for (int id = 0; id < (int)opts.num_proc_thread; ++id) {        // the reference's per-processor feed order (processor.cpp:104-160)
    int idx = id * (int)readfeed.num_sense;
    for (; readfeed.next(idx, readstr);) {
      Read read(readstr);
      read.init(opts);
      if (!read.isEmpty && read.isValid) {                         // too-short reads go along: the library counts them (num_short) and never aligns them
        Batch& b = pending.back();
        b.ids.push_back(read.id);
        for (char c : read.sequence) b.seqcat.push_back((char)nt_table[(int)((unsigned char)c & 0x7F)]);   // 0..3, 4 = ambiguous (common.hpp:68-77)
        b.off.push_back(b.seqcat.size());
        if (b.ids.size() == batch_reads) { if ((int)pending.size() == ngpu) flush(); else pending.emplace_back(); }
      }
      readstr.resize(0);
      // Known deviation (paired files only): the reference `continue`s past its file switch for a read it does not process in
      // the current index pass (too short, or already is_done from an earlier index: processor.cpp:116-124 vs :160), so from
      // then on it draws mates from the wrong file and stops when either file ends -- which reads are searched against which
      // index then depends on the thread count and on earlier results.  Here every read is searched against every index.
      if (opts.is_paired) idx ^= 1;
    }
  }
  flush();
  readfeed.rewind_in();
  readfeed.init_vzlib_in();

  readstats.num_aligned.store(total[SMR_CNT_NUM_ALIGNED], std::memory_order_relaxed);
  readstats.num_short.store(total[SMR_CNT_NUM_SHORT], std::memory_order_relaxed);
  for (size_t i = 0; i < nrefs; ++i) readstats.reads_matched_per_db[i] += total[SMR_CNT_FIXED + i];
  for (smr_ctx* ctx : ctxs) smr_destroy(ctx);
  INFO("==== Done alignment (libsmr_b200) ====\n");

  readstats.set_is_set_aligned_id_cov();
  readstats.store_to_db(kvdb);
}
