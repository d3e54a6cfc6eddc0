// Native index builder: FASTA -> <prefix>.kmer_P.dat / .bursttrie_P.dat / .pos_P.dat / .stats, the on-disk format that
// Index::load (src/sortmerna/index.cpp:143-357), Refstats::load (refstats.cpp:103-190) and smr_load_index_part read.
// Stands in for build_index (src/sortmerna/indexdb.cpp:1119-2095).  Same content as the reference's builder:
//   * the same (L+1)-mer windows, alphabet map and part splitting rule (indexdb.cpp:1343-1435),
//   * the same mini burst tries -- shape and bucket order are those of sequential insertion with the burst rule of
//     insert_prefix (indexdb.cpp:147-304), so the .bursttrie stream is byte-identical up to the id words,
//   * the same 9-mer occurrence counts incl. the "already counted by the forward window" rule (indexdb.cpp:1457-1464),
//   * the same position lists, capped at max_pos in scan order (indexdb.cpp:318-348).
// What differs: the reference numbers the unique L-mers with a CMPH minimal perfect hash (arbitrary bijection onto [0,N),
// indexdb.cpp:1571-1590,1715-1717); here ids are given in order of first occurrence -- one pass, no key file, no second scan
// of the FASTA.  Nothing downstream depends on the numbering (ids only address the positions table).
#include "smr_build.h"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <deque>
#include <exception>
#include <fstream>
#include <thread>

namespace smr {
namespace {

constexpr uint32_t kNoneB = 0xFFFFFFFFu;
constexpr uint32_t kBurstBytes = 128;   // THRESHOLD, include/indexdb.hpp:60
constexpr uint32_t kEntryBytes = 8;     // ENTRYSIZE, include/indexdb.hpp:57

// letter -> 2-bit code of the index builder (map_nt, indexdb.cpp:83-109): everything that is not listed is 0
struct NtMap {
  uint8_t m[256];
  NtMap() {
    memset(m, 0, sizeof(m));
    for (const char* p = "BCDWYbcwy"; *p; ++p) m[(uint8_t)*p] = 1;
    for (const char* p = "GKSXgksx"; *p; ++p) m[(uint8_t)*p] = 2;
    for (const char* p = "TUtu"; *p; ++p) m[(uint8_t)*p] = 3;
  }
};
const NtMap kMap;

struct Elem { uint8_t flag = 0; uint32_t ref = 0; };   // flag 1: node index, 2: bucket index
struct Node { Elem e[4]; };
struct BEntry { uint32_t tail, id; };

struct TriePool {
  std::vector<Node> nodes;
  std::vector<std::vector<BEntry>> buckets;
  uint32_t new_node() { nodes.emplace_back(); return (uint32_t)nodes.size() - 1; }
  uint32_t new_bucket() { buckets.emplace_back(); return (uint32_t)buckets.size() - 1; }
};


// search_burst_trie + insert_prefix (indexdb.cpp:402-451,147-304) in one walk.
// c(k) = k-th character of the (partialwin+1)-character tail.  Returns the id stored with the entry.
//   forward trie: id_in == kNoneB -> the entry takes the id of an existing entry with the same L-mer (all but the last tail
//   character equal), else next_id++;   reverse trie: id_in is the id of the window.
template <class CharAt>
uint32_t trie_add(TriePool& tp, uint32_t& root, const CharAt& c, const uint32_t tail_len, const uint32_t burst_depth,
                  const uint32_t id_in, uint32_t& next_id) {
  if (root == kNoneB) root = tp.new_node();
  uint32_t node = root, depth = 1;
  uint32_t ch = c(0);
  while (tp.nodes[node].e[ch].flag == 1) {
    node = tp.nodes[node].e[ch].ref;
    ch = c(depth++);
  }
  const uint32_t s = tail_len - depth;
  uint32_t encode = 0;
  for (uint32_t i = 0; i < s; ++i) encode |= c(depth + i) << (2 * i);
  uint32_t id = id_in;
  if (tp.nodes[node].e[ch].flag == 2) {
    const uint32_t msk = (1u << (2 * (s - 1))) - 1;
    for (const BEntry& be : tp.buckets[tp.nodes[node].e[ch].ref]) {
      if ((be.tail & msk) == (encode & msk)) {
        if (id_in == kNoneB) id = be.id;
        if (be.tail == encode) return be.id;   // the (L+1)-mer is there already (duplicates not allowed, :1471-1473)
      }
    }
  }
  if (id == kNoneB) id = next_id++;
  if (tp.nodes[node].e[ch].flag == 0) {
    const uint32_t b = tp.new_bucket();
    tp.nodes[node].e[ch].flag = 2;
    tp.nodes[node].e[ch].ref = b;
  }
  const uint32_t bi = tp.nodes[node].e[ch].ref;
  tp.buckets[bi].push_back(BEntry{encode, id});
  // burst (indexdb.cpp:221-299): only buckets above depth L+1-L/2-3, once the bucket exceeds THRESHOLD bytes
  if (depth < burst_depth && tp.buckets[bi].size() * kEntryBytes > kBurstBytes) {
    const uint32_t child = tp.new_node();
    std::vector<BEntry> old;
    old.swap(tp.buckets[bi]);
    for (const BEntry& be : old) {
      Elem& ce = tp.nodes[child].e[be.tail & 3u];
      if (ce.flag == 0) { ce.flag = 2; ce.ref = tp.new_bucket(); }
      tp.buckets[ce.ref].push_back(BEntry{be.tail >> 2, be.id});
    }
    tp.nodes[node].e[ch].flag = 1;
    tp.nodes[node].e[ch].ref = child;
  }
  return id;
}

// traversetrie (indexdb.cpp:528-587): bytes the reference's loader allocates = nodes * 4 * sizeof(NodeElement) + bucket bytes
uint32_t trie_bytes(const TriePool& tp, uint32_t node) {
  uint32_t total = 4 * 16;
  for (int k = 0; k < 4; ++k) {
    const Elem& e = tp.nodes[node].e[k];
    if (e.flag == 1) total += trie_bytes(tp, e.ref);
    else if (e.flag == 2) total += (uint32_t)tp.buckets[e.ref].size() * kEntryBytes;
  }
  return total;
}

// load_index (indexdb.cpp:769-861): breadth-first stream of flags and buckets
void trie_stream(const TriePool& tp, uint32_t root, std::vector<uint8_t>& out) {
  std::deque<Elem> fifo;
  for (int k = 0; k < 4; ++k) { fifo.push_back(tp.nodes[root].e[k]); out.push_back(tp.nodes[root].e[k].flag); }
  while (!fifo.empty()) {
    const Elem e = fifo.front(); fifo.pop_front();
    if (e.flag == 1) {
      for (int k = 0; k < 4; ++k) { fifo.push_back(tp.nodes[e.ref].e[k]); out.push_back(tp.nodes[e.ref].e[k].flag); }
    } else if (e.flag == 2) {
      const std::vector<BEntry>& b = tp.buckets[e.ref];
      const uint32_t bytes = (uint32_t)b.size() * kEntryBytes;
      const size_t o = out.size();
      out.resize(o + 4 + bytes);
      memcpy(&out[o], &bytes, 4);
      memcpy(&out[o + 4], b.data(), bytes);
    }
  }
}

bool write_file(const std::string& path, const void* p, size_t n, std::string& err) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { err = "cannot open " + path + " for writing: " + strerror(errno); return false; }
  const bool ok = n == 0 || fwrite(p, 1, n, f) == n;
  if (fclose(f) != 0 || !ok) { err = "short write to " + path; return false; }
  return true;
}

}  // namespace

// STEP 1 of build_index (indexdb.cpp:1188-1271): records, nucleotide distribution, total length.  seq = the builder's 2-bit codes
// (map_nt); seq04 (optional) = the same characters in the aligner's 0..4 alphabet (nt_table, common.hpp:68-77).
std::string parse_reference_fasta(const std::string& fasta, uint32_t pread, bool want04, std::vector<RefRecord>& recs, double freq[4], uint64_t& full_len,
                                  size_t& file_size) {
  static const struct Nt04 { uint8_t m[256]; Nt04() { memset(m, 4, sizeof(m)); const char* a = "AaCcGgTtUu"; const uint8_t v[10] = {0, 0, 1, 1, 2, 2, 3, 3, 3, 3}; for (int i = 0; i < 10; ++i) m[(uint8_t)a[i]] = v[i]; } } k04;
  std::vector<char> file;
  {
    std::ifstream in(fasta, std::ios::binary | std::ios::ate);
    if (!in) return "Could not open file: " + fasta;
    file.resize((size_t)in.tellg());
    in.seekg(0);
    if (!file.empty()) in.read(file.data(), (std::streamsize)file.size());
  }
  size_t o = 0;
  const size_t n = file.size();
  file_size = n;
  if (n == 0) return "empty reference file";
  while (o < n) {
    RefRecord r;
    r.rec_start = o;
    if (file[o] != '>') return "Each read header of the database fasta file must begin with '>'; check sequence # " + std::to_string(2 * recs.size());
    ++o;
    bool stop = false;
    while (o < n && file[o] != '\n') {
      const char c = file[o++];
      if (c != ' ' && c != '\t' && !stop) r.name.push_back(c); else stop = true;
    }
    if (o < n) ++o;   // the newline
    while (o < n && file[o] != '>') {
      const char c = file[o++];
      if (c != '\n' && c != ' ') {
        r.seq.push_back(kMap.m[(uint8_t)c]);
        if (want04) r.seq04.push_back(k04.m[(uint8_t)c]);
        if (c != 'N') freq[kMap.m[(uint8_t)c]] += 1.0;
      }
    }
    r.rec_end = o;
    full_len += r.seq.size();
    if (r.seq.size() < pread)
      return "At least one of your sequences is shorter than the seed length " + std::to_string(pread) +
             ", please filter out all sequences shorter than " + std::to_string(pread) + " to continue index construction.";
    recs.push_back(std::move(r));
  }
  return std::string();
}

// which sequences go into the part that starts at record `first` (indexdb.cpp:1381-1431): 9.5e-6 MB per window, a sequence that does
// not fit alone is skipped.  members empty + no error: nothing but oversized sequences was left.
std::string next_index_part(const std::vector<RefRecord>& recs, size_t first, uint32_t pread, double max_mb, std::vector<size_t>& members, size_t& next,
                            uint64_t& start_part, uint64_t& seq_part_size) {
  members.clear();
  double index_size = 0;
  next = first;
  seq_part_size = 0;
  start_part = recs[first].rec_start;
  for (; next < recs.size(); ++next) {
    const double est = (double)(recs[next].seq.size() - pread + 1) * 9.5e-6;
    if (est > max_mb) continue;
    if (index_size + est > max_mb) break;
    index_size += est;
    seq_part_size = recs[next].rec_end - start_part;
    members.push_back(next);
  }
  if (members.empty() && next < recs.size())
    return "no index was created, all of your sequences are too large to be indexed with the current memory limit";
  return std::string();
}

std::string build_index_files(const std::string& fasta, const std::string& prefix, const BuildOptions& opt, BuildReport* rep) {
  const uint32_t L = opt.lnwin, pread = L + 1, half = L / 2;
  if (L < 8 || L > 26 || (L & 1)) return "unsupported seed length";
  if (opt.interval == 0) return "interval must be >= 1";
  std::string err;
  std::vector<RefRecord> recs;
  double freq[4] = {0, 0, 0, 0};
  uint64_t full_len = 0;
  size_t n = 0;
  if (!(err = parse_reference_fasta(fasta, pread, false, recs, freq, full_len, n)).empty()) return err;
  const uint32_t limit = 1u << L;
  const uint32_t mask32 = limit - 1;
  const uint32_t burst_depth = pread - half - 3;
  uint32_t nthreads = opt.threads;
  if (nthreads == 0) { const uint32_t hw = std::thread::hardware_concurrency(); nthreads = hw / 8 < 1 ? 1 : (hw / 8 > 8 ? 8 : hw / 8); }
  if (nthreads > 64) nthreads = 64;
  struct PartStat { uint64_t start_part, seq_part_size; uint32_t numseq_part, pad; };
  std::vector<PartStat> parts;
  BuildReport report;
  size_t first = 0;
  uint16_t part_num = 0;
  while (first < recs.size()) {
    std::vector<size_t> members;
    size_t next = first;
    uint64_t seq_part_size = 0, start_part = 0;
    if (!(err = next_index_part(recs, first, pread, opt.max_mb, members, next, start_part, seq_part_size)).empty()) return err;
    if (members.empty()) break;   // only oversized sequences were left
    // Mini tries are independent per 9-mer, and forward / reverse tries are independent of each other apart from the ids:
    // worker t of T owns the 9-mers k with k % T == t, once for the forward and once for the reverse tries (2T threads).
    // Forward workers number new L-mers locally (made global by a per-worker base afterwards -- still a bijection onto
    // [0,N)); reverse workers store the window index of the inserting window and take that window's id at the end.
    const uint32_t T = nthreads;
    std::vector<TriePool> poolF(T), poolR(T);
    std::vector<uint32_t> localN(T, 0);
    std::vector<uint32_t> rootF(limit, kNoneB), rootR(limit, kNoneB), count(limit, 0);
    size_t total_win = 0;
    for (size_t m : members) total_win += (recs[m].seq.size() - pread + opt.interval) / opt.interval;
    if (total_win >= 0xFFFFFFFFull) return "more than 2^32 windows in one index part";
    std::vector<uint32_t> win_id(total_win);
    std::vector<uint8_t> win_owner(total_win);
    // fn(w, kf, kr, tf): every window of the part in scan order; tf = pointer to character `half` of the window
    auto scan = [&](auto&& fn) {
      size_t w = 0;
      for (size_t m : members) {
        const uint8_t* s = recs[m].seq.data();
        const uint32_t len = (uint32_t)recs[m].seq.size();
        const uint32_t numwin = (len - pread + opt.interval) / opt.interval;
        uint32_t kf = 0, kr = 0;
        for (uint32_t j = 0; j < half; ++j) { kf = (kf << 2) | s[j]; kr = (kr << 2) | s[half + 1 + j]; }
        uint32_t pos = 0;
        for (uint32_t j = 0; j < numwin; ++j, ++w) {
          fn(w, kf, kr, s + pos + half);
          if (j != numwin - 1)
            for (uint32_t sh = 0; sh < opt.interval; ++sh) {
              kf = ((kf << 2) & mask32) | s[pos + half];
              kr = ((kr << 2) & mask32) | s[pos + half + 1 + half];
              ++pos;
            }
        }
      }
    };
    {
      std::vector<std::thread> pool;
      for (uint32_t t = 0; t < T; ++t) {
        pool.emplace_back([&, t]() {     // forward (L+1)-mers: prefix 9-mer -> tail s[pos+half .. pos+L]
          TriePool& tp = poolF[t];
          uint32_t next_id = 0;
          scan([&](size_t w, uint32_t kf, uint32_t, const uint8_t* tf) {
            if (kf % T != t) return;
            win_id[w] = trie_add(tp, rootF[kf], [tf](uint32_t k) -> uint32_t { return tf[k]; }, half + 1, burst_depth, kNoneB, next_id);
            win_owner[w] = (uint8_t)t;
          });
          localN[t] = next_id;
        });
        pool.emplace_back([&, t]() {     // reverse (L+1)-mers: suffix 9-mer s[pos+half+1 .. pos+L] -> tail s[pos+half], .., s[pos]
          TriePool& tp = poolR[t];
          uint32_t unused = 0;
          scan([&](size_t w, uint32_t, uint32_t kr, const uint8_t* tf) {
            if (kr % T != t) return;
            trie_add(tp, rootR[kr], [tf](uint32_t k) -> uint32_t { return *(tf - k); }, half + 1, burst_depth, (uint32_t)w, unused);
          });
        });
      }
      // 9-mer occurrence counts (indexdb.cpp:1457-1464): order dependent ("already counted by the forward window"), sequential
      std::vector<bool> by_forward(limit, false);
      scan([&](size_t, uint32_t kf, uint32_t kr, const uint8_t*) {
        count[kf]++;
        by_forward[kf] = true;
        if (!by_forward[kr]) count[kr]++;
      });
      for (auto& th : pool) th.join();
    }
    std::vector<uint32_t> base(T + 1, 0);
    for (uint32_t t = 0; t < T; ++t) base[t + 1] = base[t] + localN[t];
    const uint32_t next_id = base[T];
    {
      std::vector<std::thread> pool;
      for (uint32_t t = 0; t < T; ++t)
        pool.emplace_back([&, t]() {
          for (auto& b : poolF[t].buckets) for (BEntry& e : b) e.id += base[t];
          const size_t lo = total_win * t / T, hi = total_win * (t + 1) / T;
          for (size_t w = lo; w < hi; ++w) win_id[w] += base[win_owner[w]];
        });
      for (auto& th : pool) th.join();
      pool.clear();
      for (uint32_t t = 0; t < T; ++t)
        pool.emplace_back([&, t]() { for (auto& b : poolR[t].buckets) for (BEntry& e : b) e.id = win_id[e.id]; });
      for (auto& th : pool) th.join();
    }
    // positions (add_kmer_to_table, indexdb.cpp:318-348): scan order, at most max_pos per id (0 = all)
    std::vector<uint32_t> psize(next_id, 0);
    for (uint32_t id : win_id) if (opt.max_pos == 0 || psize[id] < opt.max_pos) psize[id]++;
    std::vector<uint64_t> poff((size_t)next_id + 1, 0);
    for (uint32_t i = 0; i < next_id; ++i) poff[i + 1] = poff[i] + psize[i];
    std::vector<uint32_t> posbuf;   // file image: u32 N, then per id: u32 size, size x {pos, seq}
    posbuf.resize(1 + (size_t)next_id + 2 * poff[next_id]);
    posbuf[0] = next_id;
    std::vector<uint64_t> wr((size_t)next_id);
    for (uint32_t i = 0; i < next_id; ++i) {
      const uint64_t at = 1 + (uint64_t)i + 2 * poff[i];
      posbuf[at] = psize[i];
      wr[i] = at + 1;
    }
    std::fill(psize.begin(), psize.end(), 0);
    {
      size_t w = 0;
      uint32_t seqno = 0;
      for (size_t m : members) {
        const uint32_t len = (uint32_t)recs[m].seq.size();
        const uint32_t numwin = (len - pread + opt.interval) / opt.interval;
        for (uint32_t j = 0; j < numwin; ++j, ++w) {
          const uint32_t id = win_id[w];
          if (opt.max_pos != 0 && psize[id] == opt.max_pos) continue;
          posbuf[wr[id]] = j * opt.interval;
          posbuf[wr[id] + 1] = seqno;
          wr[id] += 2;
          psize[id]++;
        }
        ++seqno;
      }
    }
    // files
    const std::string ps = std::to_string(part_num);
    if (!write_file(prefix + ".kmer_" + ps + ".dat", count.data(), (size_t)limit * 4, err)) return err;
    // .bursttrie: 9-mers in order; T contiguous ranges are streamed in parallel and written back to back
    std::vector<std::vector<uint8_t>> streams(T);
    {
      std::vector<std::thread> pool;
      for (uint32_t t = 0; t < T; ++t)
        pool.emplace_back([&, t]() {
          std::vector<uint8_t>& stream = streams[t];
          const uint32_t lo = (uint32_t)((uint64_t)limit * t / T), hi = (uint32_t)((uint64_t)limit * (t + 1) / T);
          for (uint32_t i = lo; i < hi; ++i) {
            const TriePool& pf = poolF[i % T];
            const TriePool& pr = poolR[i % T];
            const uint32_t sz[2] = {rootF[i] != kNoneB ? trie_bytes(pf, rootF[i]) : 0u, rootR[i] != kNoneB ? trie_bytes(pr, rootR[i]) : 0u};
            const size_t at = stream.size();
            stream.resize(at + 8);
            memcpy(&stream[at], sz, 8);
            if (rootF[i] != kNoneB) trie_stream(pf, rootF[i], stream);
            if (rootR[i] != kNoneB) trie_stream(pr, rootR[i], stream);
          }
        });
      for (auto& th : pool) th.join();
    }
    size_t stream_bytes = 0, trie_nodes = 0;
    for (uint32_t t = 0; t < T; ++t) { stream_bytes += streams[t].size(); trie_nodes += poolF[t].nodes.size() + poolR[t].nodes.size(); }
    {
      const std::string path = prefix + ".bursttrie_" + ps + ".dat";
      FILE* f = fopen(path.c_str(), "wb");
      if (!f) return "cannot open " + path + " for writing: " + strerror(errno);
      bool ok = true;
      for (uint32_t t = 0; t < T && ok; ++t) ok = streams[t].empty() || fwrite(streams[t].data(), 1, streams[t].size(), f) == streams[t].size();
      if (fclose(f) != 0 || !ok) return "short write to " + path;
    }
    if (!write_file(prefix + ".pos_" + ps + ".dat", posbuf.data(), posbuf.size() * 4, err)) return err;
    parts.push_back(PartStat{start_part, seq_part_size, (uint32_t)members.size(), 0});
    report.unique_lmers += next_id;
    report.windows += total_win;
    report.trie_nodes += trie_nodes;
    report.bytes_written += (uint64_t)limit * 4 + stream_bytes + posbuf.size() * 4;
    ++part_num;
    first = next;
  }
  if (part_num == 0) return "no index was created";
  // .stats (indexdb.cpp:2020-2080)
  std::vector<uint8_t> st;
  auto put = [&st](const void* p, size_t k) { const uint8_t* b = (const uint8_t*)p; st.insert(st.end(), b, b + k); };
  const uint64_t filesize = n;
  put(&filesize, 8);
  const uint32_t fasta_len = (uint32_t)fasta.size() + 1;
  put(&fasta_len, 4);
  put(fasta.c_str(), fasta_len);
  const double tot = freq[0] + freq[1] + freq[2] + freq[3];
  double bf[4] = {freq[0] / tot, freq[1] / tot, freq[2] / tot, freq[3] / tot};
  put(bf, 32);
  put(&full_len, 8);
  put(&L, 4);
  const uint64_t numseq = recs.size();
  put(&numseq, 8);
  put(&part_num, 2);
  for (const PartStat& p : parts) put(&p, sizeof(PartStat));
  const uint32_t num_sq = (uint32_t)recs.size();
  put(&num_sq, 4);
  for (const RefRecord& r : recs) {
    const uint32_t len_id = (uint32_t)r.name.size(), slen = (uint32_t)r.seq.size();
    put(&len_id, 4);
    put(r.name.data(), len_id);
    put(&slen, 4);
  }
  if (!write_file(prefix + ".stats", st.data(), st.size(), err)) return err;
  report.parts = part_num;
  report.numseq = recs.size();
  report.bytes_written += st.size();
  if (rep) *rep = report;
  return std::string();
}

}  // namespace smr

// C ABI (include/smr_b200.h)
extern "C" int smr_build_index(const char* fasta_path, const char* out_prefix, uint32_t lnwin, uint32_t interval, uint32_t max_pos,
                               double max_mb, uint32_t threads, uint64_t* report6, char* err, size_t err_cap) {
  if (!fasta_path || !out_prefix) return 2;   // SMR_ERR_ARG
  smr::BuildOptions o;
  o.lnwin = lnwin; o.interval = interval; o.max_pos = max_pos; o.max_mb = max_mb; o.threads = threads;
  smr::BuildReport rep;
  std::string e;
  try {
    e = smr::build_index_files(fasta_path, out_prefix, o, &rep);
  } catch (const std::exception& ex) {
    e = std::string("index build failed: ") + ex.what();
  }
  if (!e.empty()) {
    if (err && err_cap) { strncpy(err, e.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
    return 3;   // SMR_ERR_INDEX
  }
  if (report6) {
    report6[0] = rep.parts; report6[1] = rep.numseq; report6[2] = rep.windows; report6[3] = rep.unique_lmers;
    report6[4] = rep.trie_nodes; report6[5] = rep.bytes_written;
  }
  return 0;
}
