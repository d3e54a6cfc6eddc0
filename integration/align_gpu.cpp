// Reference-side binding of libsmr_b200.so: a replacement for the reference's
//   void align(Readfeed&, Readstats&, Index&, KeyValueDatabase&, Runopts&)          (src/sortmerna/processor.cpp:173-285)
// written against the reference's own headers and nothing else -- the translation unit a maintainer adds to the host program
// (INTEGRATION.md).  Everything around it stays the reference's: CLI / Runopts, Readfeed, Refstats (ALP, minimal_score),
// References, KeyValueDatabase, writeSummary / writeReports / denovo_stats.  oracle/Makefile.ref links it with the reference's
// unmodified objects into oracle/_ref/sortmerna_gpu (the symbol `align` of processor.o is weakened, this one wins).
//
// Differences to the CPU driver, all invisible to the rest of the program:
//   * every (index, part) is made resident on the GPU once; reads are streamed ONCE (the reference re-reads them per index part);
//   * the feed is drained by several parser threads (one per group of the reference's "processors": each processor id owns
//     its own split file, processor.cpp:104-160), which encode reads into batches; a bounded queue hands every full batch to
//     one worker thread per context (SMR_GPUS devices x SMR_CTX_PER_GPU contexts), so parsing, H2D / kernels / D2H and the KVDB
//     stores overlap;
//   * the per-read KVDB blob is produced by smr_pack_kvdb_blobs (byte-identical to Read::toBinString) and stored under the same
//     key (read.id);
//   * Readstats counters come back from the library (num_aligned, reads_matched_per_db, num_short of the last index pass).
// Resume (-task with a KVDB that already holds results) is not supported by this path: every read is aligned again (the CPU
// driver consults Read::load_db first, processor.cpp:116-126).
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <fstream>
#include <memory>
#include <mutex>
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

struct Batch {
  std::vector<std::string> ids; std::string seqcat; std::vector<uint64_t> off{0};
  std::vector<smr_read_result> res; std::vector<smr_aln> alns; std::vector<uint32_t> cigars; uint64_t used = 0; std::vector<uint64_t> cnt;
  std::string blobs; std::vector<uint64_t> boff;
  void reset() { ids.clear(); seqcat.clear(); off.assign(1, 0); used = 0; }
};

// full batches on their way to the GPU workers, empty ones on their way back to the parsers (bounded: host memory stays
// at a few batches however long the input is)
class BatchQueue {
  std::mutex m; std::condition_variable cv; std::deque<std::unique_ptr<Batch>> q; bool closed = false;
 public:
  void push(std::unique_ptr<Batch> b) { { std::lock_guard<std::mutex> l(m); q.push_back(std::move(b)); } cv.notify_one(); }
  std::unique_ptr<Batch> pop() {   // nullptr once closed and drained
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [&] { return !q.empty() || closed; });
    if (q.empty()) return nullptr;
    std::unique_ptr<Batch> b = std::move(q.front()); q.pop_front();
    return b;
  }
  void close() { { std::lock_guard<std::mutex> l(m); closed = true; } cv.notify_all(); }
};

}  // namespace

void align(Readfeed& readfeed, Readstats& readstats, Index& /*index: the library keeps its own resident form*/, KeyValueDatabase& kvdb, Runopts& opts)
{
  INFO("==== Starting alignment (libsmr_b200) ====");
  const auto t_start = std::chrono::high_resolution_clock::now();
  // one context per GPU (SMR_GPUS, default 1; never more than the devices present): reads shard by record, every GPU holds the
  // whole index, the only cross-GPU state are the Readstats counters (summed below) -- SURVEY 8(e)
  int ngpu = 1;
  if (const char* e = getenv("SMR_GPUS")) ngpu = std::max(1, std::min(atoi(e), smr_device_count()));
  // two contexts per GPU (SMR_CTX_PER_GPU): while one batch is in its kernels, the other context copies its batch in / its results
  // out and packs the KVDB blobs (the library serialises the kernel sections of contexts that share a device)
  int cpg = 2;
  if (const char* e = getenv("SMR_CTX_PER_GPU")) cpg = std::max(1, std::min(atoi(e), 4));
  const int nctx = ngpu * cpg;
  std::vector<smr_ctx*> ctxs((size_t)nctx, nullptr);
  for (int c = 0; c < nctx; ++c)
    if (smr_init(c / cpg, &ctxs[c]) != SMR_OK) { ERR("no usable CUDA device ", c / cpg, ": the GPU alignment path has no CPU fallback"); exit(EXIT_FAILURE); }

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
  // SMR_INDEX_DEVICE=1: the library indexes the reference FASTA itself on the GPU (smr_build_index_device) with the options of this
  // run instead of reading the index files -- same resident arrays when the files were built with the same options.
  const bool index_on_device = getenv("SMR_INDEX_DEVICE") && atoi(getenv("SMR_INDEX_DEVICE")) != 0;
  for (size_t i = 0; i < opts.indexfiles.size(); ++i) {
    const uint32_t skip[3] = {opts.skiplengths[i][0], opts.skiplengths[i][1], opts.skiplengths[i][2]};
    if (index_on_device) {
      std::vector<std::thread> loaders;
      for (smr_ctx* ctx : ctxs)
        loaders.emplace_back([&, ctx] {
          uint32_t nparts = 0;
          if (smr_build_index_device(ctx, (uint32_t)i, opts.indexfiles[i].first.c_str(), refstats.lnwin[i], opts.interval, opts.max_pos, opts.max_file_size, skip,
                                     refstats.minimal_score[i], &nparts, nullptr) != SMR_OK)
            die(ctx, "smr_build_index_device");
          if (nparts != refstats.num_index_parts[i]) { ERR("index ", i, ": ", nparts, " parts built on the device, ", refstats.num_index_parts[i], " in the .stats file (different -m ?)"); exit(EXIT_FAILURE); }
        });
      for (auto& t : loaders) t.join();
      continue;
    }
    for (uint16_t part = 0; part < refstats.num_index_parts[i]; ++part) {
      const std::string pfx = opts.indexfiles[i].second, sfx = "_" + std::to_string(part) + ".dat";
      const std::vector<char> kmer = slurp(pfx + ".kmer" + sfx), trie = slurp(pfx + ".bursttrie" + sfx), pos = slurp(pfx + ".pos" + sfx);
      refs.load((uint32_t)i, part, opts, refstats);                    // unchanged References::load: sequences in the 0-4 alphabet
      std::string cat; std::vector<uint64_t> off(1, 0);
      for (auto& r : refs.buffer) { cat += r.sequence; off.push_back(cat.size()); }
      std::vector<std::thread> loaders;                                // one host thread per GPU: flattening + upload run side by side
      for (smr_ctx* ctx : ctxs)
        loaders.emplace_back([&, ctx] {
          if (smr_load_index_part(ctx, (uint32_t)i, part, kmer.data(), kmer.size(), trie.data(), trie.size(), pos.data(), pos.size(),
                                  (const uint8_t*)cat.data(), off.data(), (uint32_t)refs.buffer.size(), refstats.lnwin[i], refstats.minimal_score[i], skip) != SMR_OK)
            die(ctx, "smr_load_index_part");
        });
      for (auto& t : loaders) t.join();
      refs.unload();
    }
  }

  const auto t_loaded = std::chrono::high_resolution_clock::now();
  // batches of reads from the unchanged Readfeed; read ids ("<file>_<n>") stay the KVDB keys
  const size_t nrefs = opts.indexfiles.size();
  uint32_t batch_reads = 1u << 19;
  if (const char* e = getenv("SMR_BATCH_READS")) batch_reads = (uint32_t)std::max(1, atoi(e));
  BatchQueue full, empty;

  std::vector<uint64_t> total(SMR_CNT_FIXED + nrefs, 0);
  std::mutex total_m;

  auto run_batch = [&](smr_ctx* ctx, Batch& b) {
    const uint32_t n = (uint32_t)b.ids.size();
    for (;;) {
      const uint32_t slots = smr_aln_slots(ctx);       // num_alignments, or the stride of the all-alignments mode (-num_alignments 0)
      b.res.resize(n); b.alns.resize((size_t)n * slots); b.cnt.assign(SMR_CNT_FIXED + nrefs, 0);
      if (b.cigars.size() < (size_t)16 * n * slots + 4096) b.cigars.resize((size_t)16 * n * slots + 4096);
      const int rc = smr_align_batch(ctx, (const uint8_t*)b.seqcat.data(), b.off.data(), n, b.res.data(), b.alns.data(), b.cigars.data(), b.cigars.size(), &b.used,
                                     b.cnt.data(), (uint32_t)b.cnt.size());
      if (rc == SMR_ERR_CAPACITY && opts.num_alignments == 0 && smr_aln_slots_needed(ctx) > slots) {   // a read stored more alignments than the stride
        if (smr_set_aln_slots(ctx, smr_aln_slots_needed(ctx)) != SMR_OK) die(ctx, "smr_set_aln_slots");
        continue;
      }
      if (rc == SMR_ERR_CAPACITY && b.cigars.size() < ((size_t)1 << 32)) { b.cigars.resize(b.cigars.size() * 2); continue; }   // CIGAR pool too small: grow, again
      if (rc != SMR_OK) die(ctx, "smr_align_batch");
      b.boff.resize((size_t)n + 1);
      smr_pack_kvdb_blobs(b.res.data(), b.alns.data(), b.cigars.data(), n, slots, (int32_t)opts.num_alignments, nullptr, nullptr, 0, b.boff.data());
      b.blobs.assign(b.boff[n], '\0');
      if (smr_pack_kvdb_blobs(b.res.data(), b.alns.data(), b.cigars.data(), n, slots, (int32_t)opts.num_alignments, nullptr, (uint8_t*)&b.blobs[0], b.blobs.size(),
                              b.boff.data()) != SMR_OK) die(ctx, "smr_pack_kvdb_blobs");
      break;
    }
    for (uint32_t r = 0; r < n; ++r)
      if (b.boff[r + 1] > b.boff[r]) kvdb.put(b.ids[r], b.blobs.substr(b.boff[r], b.boff[r + 1] - b.boff[r]));   // == kvdb.put(read.id, read.toBinString())
    std::lock_guard<std::mutex> l(total_m);
    for (size_t k = 0; k < b.cnt.size(); ++k) total[k] += b.cnt[k];
  };

  std::vector<std::thread> workers;
  for (int g = 0; g < nctx; ++g)
    workers.emplace_back([&, g] {
      while (std::unique_ptr<Batch> b = full.pop()) { run_batch(ctxs[g], *b); b->reset(); empty.push(std::move(b)); }
    });

  readfeed.init_reading();
  // parser threads: processor id t, t + T, t + 2T, ... each in the reference's own feed order (processor.cpp:104-160)
  const int nproc = (int)opts.num_proc_thread;
  int nparse = std::max(1, std::min(nproc, (int)std::thread::hardware_concurrency()));
  if (const char* e = getenv("SMR_PARSE_THREADS")) nparse = std::max(1, std::min(nproc, atoi(e)));
  for (int k = 0; k < nparse + 2 * nctx; ++k) empty.push(std::make_unique<Batch>());   // one per parser + two per context (one running, one queued)
  std::vector<std::thread> parsers;
  for (int t = 0; t < nparse; ++t)
    parsers.emplace_back([&, t] {
      std::unique_ptr<Batch> cur = empty.pop();
      std::string readstr;
      for (int id = t; id < nproc; id += nparse) {
        int idx = id * (int)readfeed.num_sense;
        for (; readfeed.next(idx, readstr);) {
          Read read(readstr);
          read.init(opts);
          if (!read.isEmpty && read.isValid) {                       // too-short reads go along: the library counts them (num_short) and never aligns them
            cur->ids.push_back(read.id);
            const size_t o = cur->seqcat.size(), len = read.sequence.size();
            cur->seqcat.resize(o + len);
            for (size_t k = 0; k < len; ++k) cur->seqcat[o + k] = (char)nt_table[(int)((unsigned char)read.sequence[k] & 0x7F)];   // 0..3, 4 = ambiguous (common.hpp:68-77)
            cur->off.push_back(o + len);
            if (cur->ids.size() == batch_reads) { full.push(std::move(cur)); cur = empty.pop(); }
          }
          readstr.resize(0);
          // Known deviation (paired files only): the reference `continue`s past its file switch for a read it does not process in
          // the current index pass (too short, or already is_done from an earlier index: processor.cpp:116-124 vs :160), so from
          // then on it draws mates from the wrong file and stops when either file ends -- which reads are searched against which
          // index then depends on the thread count and on earlier results.  Here every read is searched against every index.
          if (opts.is_paired) idx ^= 1;
        }
      }
      if (!cur->ids.empty()) full.push(std::move(cur)); else empty.push(std::move(cur));
    });
  for (auto& t : parsers) t.join();
  full.close();
  for (auto& t : workers) t.join();
  readfeed.rewind_in();
  readfeed.init_vzlib_in();

  readstats.num_aligned.store(total[SMR_CNT_NUM_ALIGNED], std::memory_order_relaxed);
  readstats.num_short.store(total[SMR_CNT_NUM_SHORT], std::memory_order_relaxed);
  for (size_t i = 0; i < nrefs; ++i) readstats.reads_matched_per_db[i] += total[SMR_CNT_FIXED + i];
  for (smr_ctx* ctx : ctxs) smr_destroy(ctx);
  {
    const auto t_done = std::chrono::high_resolution_clock::now();
    const std::chrono::duration<double> el_load = t_loaded - t_start, el_all = t_done - t_start;
    INFO("index + references resident on ", ngpu, " GPU(s) in ", el_load.count(), " sec; reads streamed, aligned and stored in ", el_all.count() - el_load.count(), " sec");
    INFO("==== Done alignment in ", el_all.count(), " sec ====\n");     // the line the reference prints (processor.cpp:280)
  }

  readstats.set_is_set_aligned_id_cov();
  readstats.store_to_db(kvdb);
}
