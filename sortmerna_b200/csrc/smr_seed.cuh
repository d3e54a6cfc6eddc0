// Seed search: one warp per read, one lane per 18-mer window, both strands, all three passes.
//
// Stands in for the window loop of traverse() (src/sortmerna/paralleltraversal.cpp:124-250), the
// bit-vector construction init_win_f/init_win_r (src/sortmerna/bitvector.cpp:56-132), Read::hashKmer
// (src/sortmerna/read.cpp:601-611) and the trie x Levenshtein-automaton DFS traversetrie_align
// (src/sortmerna/traverse_bursttrie.cpp:100-298).
//
// B200 mapping: the index is a set of flat HBM arrays (smr_index.h); a window search is a chain of
// dependent 8/32-byte sector reads (lookup -> root node -> child nodes -> bucket entries), so the
// kernel is HBM/L2-latency bound and is parallelised over (read, strand, window) -- 90 windows per
// 150-nt read -- with warp ballot/shuffle compaction of the hits into a per-read region.
// All windows of all passes are searched up front (a window's hits depend only on (read, strand,
// position, index)); the candidate kernel later replays the reference's pass order on them.
#pragma once
#include "smr_dev.cuh"
#include "smr_levbits.h"

namespace smr {

// ---------------------------------------------------------------------------------------------
// The reference drives the trie DFS with a table-driven universal Levenshtein automaton for d=1
// (traverse_bursttrie.cpp:68-98, bit-vectors from bitvector.cpp:56-132).  That automaton accepts at
// depth d (d+1 text characters consumed) exactly when the edit distance between those d+1 characters
// and the 9-nt half window is <= 1, is "alive" exactly when some prefix of the half window is within
// distance 1 of the text so far, and reaches state 9 at depth 8 exactly on an exact match (checked
// exhaustively against the table: tests/test_lev_equivalence.py, and by the GPU-vs-oracle window tests;
// the oracle keeps the table).  The kernels therefore evaluate those three predicates with a few
// bit-parallel operations on 2-bit packed strings instead of walking a table:
//   P = the half window (pw chars), T = the text (trie path letters + bucket tail, pw+1 chars),
//   both packed first character in the LOWEST two bits (the bucket tails' own layout).
// ---------------------------------------------------------------------------------------------
// reverse the order of the pw 2-bit characters of v
__device__ __forceinline__ uint32_t rev_chars(uint32_t v, uint32_t pw) {
  uint32_t x = __brev(v) >> (32 - 2 * pw);
  return ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
}

// 2-bit packed reads, first base most significant: base i of a read lives in word i>>4 at bit 30-2*(i&15).
// Returns the L-mer starting at forward position p as a 2L-bit integer (first base most significant).
__device__ __forceinline__ uint64_t window_fwd(const uint32_t* __restrict__ pk, uint32_t p, uint32_t L) {
  const uint32_t wi = p >> 4, s = 2 * (p & 15);
  const uint64_t hi = ((uint64_t)__ldg(pk + wi) << 32) | __ldg(pk + wi + 1);
  const uint64_t lo = (uint64_t)__ldg(pk + wi + 2) << 32;
  const uint64_t v = s ? ((hi << s) | (lo >> (64 - s))) : hi;
  return v >> (64 - 2 * L);
}
// reverse complement of a 2L-bit L-mer (Read::revIntStr, read.cpp:350-357, on the 0..3 alphabet)
__device__ __forceinline__ uint64_t revcomp_bits(uint64_t v, uint32_t L) {
  uint64_t x = __brevll(v) >> (64 - 2 * L);
  x = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
  return ~x & ((1ull << (2 * L)) - 1);
}

// Per-lane hit buffer (ids of one window, needed for the per-window de-duplication,
// traverse_bursttrie.cpp:265-279).  Slot k lives at buf[k*stride].
struct LaneHits {
  uint32_t* buf; uint32_t stride, cap, n; bool overflow;
};

struct SeedStats { uint32_t nodes, buckets, entries; };

// the reference's per-entry side effects (traverse_bursttrie.cpp:249-281) applied to a classification code;
// returns true when the window search ends on a 0-error match
__device__ __forceinline__ bool apply_entry(uint32_t code, uint32_t id, bool full_search, LaneHits& lh) {
  const uint32_t d1 = code & 3u;
  if (d1 == 0) return false;
  const bool z = (code & 4u) && !full_search;
  if (d1 == 2 && z) { lh.n = 1; lh.buf[0] = id; lh.overflow = false; return true; }                  // :256-262
  for (uint32_t f = 0; f < lh.n && f < lh.cap; ++f) if (lh.buf[f * lh.stride] == id) return false;   // duplicate: :265-277
  if (lh.n < lh.cap) lh.buf[lh.n * lh.stride] = id; else lh.overflow = true;
  lh.n++;
  if (d1 == 1 && z) { lh.n = 1; lh.buf[0] = id; lh.overflow = false; return true; }                  // 0-error one step after the push
  return false;
}

// Sequential DFS of one mini burst trie for one window (traverse_bursttrie.cpp:100-298): used for the few
// windows that overflow the cooperative buffers, and by the unit-test entry point.
template <bool INSTR>
__device__ bool walk_trie(const DevIndex& ix, uint32_t root, uint32_t P, bool full_search, LaneHits& hits, SeedStats& st) {
  const uint32_t pw = ix.partialwin;
  uint32_t stk_node[16];
  uint8_t stk_letter[16];
  uint32_t depth = 0, node = root, letter = 0, path = 0;
  uint4 na = __ldg(ix.nodes + 2 * (size_t)node), nb = __ldg(ix.nodes + 2 * (size_t)node + 1);
  if (INSTR) st.nodes++;
  for (;;) {
    if (letter == 4) {
      if (depth == 0) return false;
      --depth;
      node = stk_node[depth]; letter = stk_letter[depth];
      path &= (1u << (2 * depth)) - 1u;
      na = __ldg(ix.nodes + 2 * (size_t)node); nb = __ldg(ix.nodes + 2 * (size_t)node + 1);
      continue;
    }
    uint32_t w0, w1;
    switch (letter) {
      case 0: w0 = na.x; w1 = na.y; break;
      case 1: w0 = na.z; w1 = na.w; break;
      case 2: w0 = nb.x; w1 = nb.y; break;
      default: w0 = nb.z; w1 = nb.w; break;
    }
    const uint32_t flag = w0 & 3u;
    const uint32_t tp = path | (letter << (2 * depth));
    if (flag == 0 || !viable_bits(P, tp, depth + 1)) { ++letter; continue; }   // empty element or automaton dead (:122-147)
    if (flag == 1) {
      stk_node[depth] = node; stk_letter[depth] = (uint8_t)(letter + 1);
      path = tp; ++depth; node = w1; letter = 0;
      na = __ldg(ix.nodes + 2 * (size_t)node); nb = __ldg(ix.nodes + 2 * (size_t)node + 1);
      if (INSTR) st.nodes++;
      continue;
    }
    const uint32_t cnt = w0 >> 2;                                               // bucket (:176-292)
    if (INSTR) st.buckets++;
    const uint2* __restrict__ e = ix.entries + w1;
    for (uint32_t k = 0; k < cnt; ++k) {
      const uint2 en = __ldg(e + k);
      if (INSTR) st.entries++;
      if (apply_entry(classify_bits(P, tp | (en.x << (2 * (depth + 1))), pw), en.y, full_search, hits)) return true;
    }
    ++letter;
  }
}

// ---------------------------------------------------------------------------------------------
// Cooperative sub-search of 32 windows at once: one window per lane for the short, divergent trie NODE
// walk; then the visited buckets are cut into tasks of <= kTaskEntries entries, compacted across the
// warp, and classified one task per lane (uniform arithmetic, no table); finally each lane replays the
// reference's DFS-order semantics (0-error exit, per-window de-duplication) over the codes of its window.
// ---------------------------------------------------------------------------------------------
constexpr int kTaskMax = 256;      // tasks per 32-window round kept in shared memory
constexpr int kTaskEntries = 8;    // entries per task
constexpr int kEntryCap = 512;     // classification codes per round kept in shared memory
constexpr uint32_t kTaskEnd = 0xFFFFu;

struct CoopSmem {                  // per warp, ~4.4 KB
  uint2 flat[kTaskMax];            // {first entry index, cnt | (depth+1)<<4 | path<<8}
  uint16_t start[kTaskMax];        // first code slot of the task
  uint16_t next[kTaskMax];         // next task of the same window, DFS order (kTaskEnd = last)
  uint8_t lane[kTaskMax];          // owning window (lane)
  uint8_t code[kEntryCap];
  uint32_t P[32];
  uint32_t ntask, nentry;
};

// trie NODE walk only: records the buckets the DFS would visit, in DFS order, as a linked list of tasks of
// <= kTaskEntries entries.  Returns false if the shared buffers overflowed (the caller then searches this
// window with walk_trie instead).
template <bool INSTR>
__device__ bool walk_nodes(const DevIndex& ix, uint32_t root, uint32_t P, CoopSmem& sm, unsigned lane, uint32_t& head, SeedStats& st) {
  uint32_t stk_node[16];
  uint8_t stk_letter[16];
  uint32_t depth = 0, node = root, letter = 0, path = 0, last = kTaskEnd;
  uint4 na = __ldg(ix.nodes + 2 * (size_t)node), nb = __ldg(ix.nodes + 2 * (size_t)node + 1);
  if (INSTR) st.nodes++;
  for (;;) {
    if (letter == 4) {
      if (depth == 0) return true;
      --depth;
      node = stk_node[depth]; letter = stk_letter[depth];
      path &= (1u << (2 * depth)) - 1u;
      na = __ldg(ix.nodes + 2 * (size_t)node); nb = __ldg(ix.nodes + 2 * (size_t)node + 1);
      continue;
    }
    uint32_t w0, w1;
    switch (letter) {
      case 0: w0 = na.x; w1 = na.y; break;
      case 1: w0 = na.z; w1 = na.w; break;
      case 2: w0 = nb.x; w1 = nb.y; break;
      default: w0 = nb.z; w1 = nb.w; break;
    }
    const uint32_t flag = w0 & 3u;
    const uint32_t tp = path | (letter << (2 * depth));
    if (flag == 0 || !viable_bits(P, tp, depth + 1)) { ++letter; continue; }
    if (flag == 1) {
      stk_node[depth] = node; stk_letter[depth] = (uint8_t)(letter + 1);
      path = tp; ++depth; node = w1; letter = 0;
      na = __ldg(ix.nodes + 2 * (size_t)node); nb = __ldg(ix.nodes + 2 * (size_t)node + 1);
      if (INSTR) st.nodes++;
      continue;
    }
    if (INSTR) st.buckets++;
    const uint32_t cnt = w0 >> 2;
    if (INSTR) st.entries += cnt;
    for (uint32_t c0 = 0; c0 < cnt; c0 += kTaskEntries) {
      const uint32_t n = min(cnt - c0, (uint32_t)kTaskEntries);
      const uint32_t slot = atomicAdd(&sm.ntask, 1u), es = atomicAdd(&sm.nentry, n);
      if (slot >= (uint32_t)kTaskMax || es + n > (uint32_t)kEntryCap) {
        if (slot < (uint32_t)kTaskMax) { sm.flat[slot] = make_uint2(0u, 0u); sm.start[slot] = 0; sm.next[slot] = (uint16_t)kTaskEnd; sm.lane[slot] = 0; }
        return false;
      }
      sm.flat[slot] = make_uint2(w1 + c0, n | ((depth + 1) << 4) | (tp << 8));
      sm.start[slot] = (uint16_t)es; sm.next[slot] = (uint16_t)kTaskEnd; sm.lane[slot] = (uint8_t)lane;
      if (last == kTaskEnd) head = slot; else sm.next[last] = (uint16_t)slot;
      last = slot;
    }
    ++letter;
  }
}

// One sub-search (forward or mirror) for the 32 windows of a round.  `root`/`P` are per lane (root ==
// kNoneDev: lane idle).  Appends to lh with the reference's de-duplication; sets zero.
template <bool INSTR>
__device__ void coop_subsearch(const DevIndex& ix, CoopSmem& sm, const uint32_t root, const uint32_t P, const bool full_search, LaneHits& lh,
                               bool& zero, SeedStats& st) {
  const unsigned lane = lane_id();
  const uint32_t pw = ix.partialwin;
  sm.P[lane] = P;
  if (lane == 0) { sm.ntask = 0; sm.nentry = 0; }
  __syncwarp();
  uint32_t head = kTaskEnd;
  bool slow = false;
  if (root != kNoneDev) slow = !walk_nodes<INSTR>(ix, root, P, sm, lane, head, st);
  __syncwarp();
  const uint32_t ntask = min(sm.ntask, (uint32_t)kTaskMax);
  // classify: one task (<= kTaskEntries consecutive entries of one bucket) per lane; pure arithmetic
  for (uint32_t t = lane; t < ntask; t += 32) {
    const uint2 tk = sm.flat[t];
    const uint32_t cnt = tk.y & 0xFu, sh = 2 * ((tk.y >> 4) & 0xFu), path = tk.y >> 8, Pw = sm.P[sm.lane[t]];
    const uint32_t es = sm.start[t];
    if (es + cnt > (uint32_t)kEntryCap) continue;   // belongs to a window that overflowed: it is redone below
    const uint2* __restrict__ ep = ix.entries + tk.x;
#pragma unroll 4
    for (uint32_t k = 0; k < cnt; ++k) sm.code[es + k] = (uint8_t)classify_bits(Pw, path | (__ldg(ep + k).x << sh), pw);
  }
  __syncwarp();
  // replay the DFS-order semantics per window over its chain of tasks
  if (!slow) {
    for (uint32_t t = head; t != kTaskEnd && !zero; t = sm.next[t]) {
      const uint2 tk = sm.flat[t];
      const uint32_t cnt = tk.y & 0xFu, es = sm.start[t];
      for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t c = sm.code[es + k];
        if ((c & 3u) == 0) continue;
        if (apply_entry(c, __ldg(ix.entries + tk.x + k).y, full_search, lh)) { zero = true; break; }
      }
    }
  } else {
    zero = walk_trie<INSTR>(ix, root, P, full_search, lh, st);
  }
  __syncwarp();
}

// both sub-searches of one window, sequentially by one lane (paralleltraversal.cpp:129-249); V = the lnwin-mer
template <bool INSTR>
__device__ bool seed_window(const DevIndex& ix, uint64_t V, bool full_search, LaneHits& hits, SeedStats& st) {
  const uint32_t pw = ix.partialwin;
  const uint32_t keyf = (uint32_t)(V >> (2 * pw)), keyr = (uint32_t)(V & ((1ull << (2 * pw)) - 1));
  hits.n = 0;
  bool zero = false;
  const uint32_t rootF = __ldg(&ix.lookup[keyf]).x;                               // :161
  if (rootF != kNoneDev) zero = walk_trie<INSTR>(ix, rootF, rev_chars(keyr, pw), full_search, hits, st);
  if (!zero) {                                                                    // :188
    const uint32_t rootR = __ldg(&ix.lookup[keyr]).y;                             // :215
    if (rootR != kNoneDev) zero = walk_trie<INSTR>(ix, rootR, keyf, full_search, hits, st);
  }
  return zero;
}

// ---------------------------------------------------------------------------------------------
// batch layout on the device
// ---------------------------------------------------------------------------------------------
struct DevBatch {
  uint32_t nreads;
  const uint8_t* seq04;      // 0..4, concatenated
  const uint32_t* seq_off;   // [nreads+1]
  const uint32_t* pk03;      // 2-bit packed, N->A(0) (seqToIntStr, read.cpp:334-347)
  const uint32_t* pk03alt;   // 2-bit packed, N->T(3): its reverse complement is the "N->A" reverse strand
  const uint32_t* pk_off;    // [nreads+1] word offsets
  const uint8_t* has_n;      // [nreads]
  uint32_t hit_scale;        // read r owns hits[hit_base(r) .. +hit_cap(r)): see hit_base()/hit_cap()
  uint32_t seq_base0;        // seq_off of the first read of the chunk, r0 = first read of the chunk
  uint32_t r0;
  uint2* hits;               // {id, win_pos | variant<<24}; one region set per (index,part): part p starts at p*hits_stride
  size_t hits_stride;        // entries per part
  uint32_t cnt_stride;       // hit_cnt of (part p, read r) lives at hit_cnt[p*cnt_stride + (r - r0)]
  uint32_t* cost;            // [chunk] estimated candidate work of a read (sum of position-list lengths of its hits, all parts)
  uint32_t* bins;            // [kCostBins * cnt_stride] reads of this chunk binned by log2(cost): heaviest-first schedule
  uint32_t* bin_count;       // [kCostBins]
  uint16_t* hit_db;          // [nreads] index_num of the first accepted alignment (reads_matched_per_db)
  uint32_t* hit_cnt;         // [nreads]
  uint32_t* flags;           // [nreads] overflow flags
  ReadState* state;          // [nreads]
  unsigned long long* counters;  // [dcCount + n_index_files]
};

// per-read hit region: capacity proportional to the read length (2 hits per nucleotide covers the
// ~1 window per 3 nt x 2-3 strand variants with ~2-3 ids per window), scaled up on a retry
__device__ __forceinline__ uint32_t hit_cap(const DevBatch& b, uint32_t r) { return b.hit_scale * (2u * (b.seq_off[r + 1] - b.seq_off[r]) + 32u); }
__device__ __forceinline__ size_t hit_base(const DevBatch& b, uint32_t part, uint32_t r) {
  return (size_t)part * b.hits_stride + (size_t)b.hit_scale * (2ull * (b.seq_off[r] - b.seq_base0) + 32ull * (r - b.r0));
}
constexpr int kCostBins = 24;

// 2-bit packing of a batch: one warp per read
__global__ void pack_reads_kernel(DevBatch b, uint32_t* pk03, uint32_t* pk03alt, uint8_t* has_n) {
  // grid-stride over reads; each warp packs one read
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (uint32_t r = b.r0 + warp; r < b.r0 + b.nreads; r += nwarps) {
    const uint32_t o = b.seq_off[r], len = b.seq_off[r + 1] - o;
    const uint32_t wo = b.pk_off[r], nw = b.pk_off[r + 1] - wo;
    bool anyn = false;
    for (uint32_t w = lane; w < nw; w += 32) {
      uint32_t a = 0, t = 0;
      for (uint32_t k = 0; k < 16; ++k) {
        const uint32_t i = w * 16 + k;
        uint32_t c = i < len ? b.seq04[o + i] : 0u;
        uint32_t ca = c, ct = c;
        if (c > 3) { ca = 0; ct = 3; anyn = anyn || (i < len); }
        a |= ca << (30 - 2 * k); t |= ct << (30 - 2 * k);
      }
      pk03[wo + w] = a;
      if (pk03alt) pk03alt[wo + w] = t;
    }
    anyn = __any_sync(kFull, anyn);
    if (lane == 0) has_n[r] = anyn ? 1 : 0;
  }
}

constexpr int kSeedWarpsPerCta = 4;
constexpr int kLaneHitCap = 128;  // ids per window in the per-warp HBM scratch (x scale on a retry)

// The seed kernel.  grid-stride over reads, one warp per read.
//   lane_hits_g: per-lane id buffers [total warps][cap_g][32]
template <bool INSTR>
__global__ void __launch_bounds__(kSeedWarpsPerCta * 32, 8)
seed_kernel(DevIndex ix, DevBatch b, DevParams prm, uint32_t* lane_hits_g, uint32_t cap_g) {
  __shared__ CoopSmem s_coop[kSeedWarpsPerCta];
  const unsigned lane = lane_id();
  const uint32_t wic = threadIdx.x >> 5;
  const uint32_t warp = blockIdx.x * kSeedWarpsPerCta + wic, nwarps = gridDim.x * kSeedWarpsPerCta;
  const uint32_t L = ix.lnwin, pw = ix.partialwin;
  const bool full = prm.is_full_search != 0;
  CoopSmem& sm = s_coop[wic];
  LaneHits lh;
  lh.buf = lane_hits_g + (size_t)warp * cap_g * 32 + lane; lh.stride = 32; lh.cap = cap_g;   // per-window ids live in a per-warp HBM scratch
  SeedStats st{0, 0, 0};
  uint32_t n_windows = 0, n_short = 0;
  const bool single = (prm.is_forward != 0) != (prm.is_reverse != 0);
  const bool do_fwd = !(single && prm.is_reverse), do_rev = !(single && prm.is_forward);
  // window positions: union of the three pass grids (paralleltraversal.cpp:118-124,262-277).  With the
  // default 18/9/3 every pass position is a multiple of the last shift.
  const uint32_t s0 = ix.skip[0], s1 = ix.skip[1], s2 = ix.skip[2];
  const uint32_t step = (s0 % s2 == 0 && s1 % s2 == 0) ? s2 : 1u;

  for (uint32_t r = b.r0 + warp; r < b.r0 + b.nreads; r += nwarps) {
    const uint32_t len = b.seq_off[r + 1] - b.seq_off[r];
    const uint32_t cnt_idx = ix.slot * b.cnt_stride + (r - b.r0);
    if (lane == 0) b.hit_cnt[cnt_idx] = 0;
    if (len < L) { n_short += (lane == 0 && ix.is_last); continue; }           // processor.cpp:109-114 (reset per pass, :228)
    // (reads that become is_done in an earlier part are still searched here: parts are seeded before the
    //  candidate kernel replays them read-major; the candidate kernel skips them, processor.cpp:120-126)
    if (b.flags[r]) continue;                                                   // scratch overflow earlier: the read is redone by the retry
    const bool hasn = b.has_n[r] != 0;
    const uint32_t* pk = b.pk03 + b.pk_off[r];
    const uint32_t* pka = hasn ? b.pk03alt + b.pk_off[r] : pk;
    const uint32_t npos = (len - L) / step + 1;          // positions q*step, q < npos
    const size_t region = hit_base(b, ix.slot, r); const uint32_t region_cap = hit_cap(b, r);
    uint32_t total = 0, flags = 0, cost = 0;
    const uint32_t nvar = hasn ? 3u : 2u;
    for (uint32_t var = 0; var < nvar; ++var) {
      if (var == kVarFwd && !do_fwd) continue;
      if (var != kVarFwd && !do_rev) continue;
      for (uint32_t q0 = 0; q0 < npos; q0 += 32) {
        const uint32_t q = q0 + lane, p = q * step;
        bool active = q < npos;
        if (active && step == 1) active = (p % s0 == 0) || (p % s1 == 0) || (p % s2 == 0);
        lh.n = 0; lh.overflow = false;
        uint64_t V = 0;
        uint32_t keyf = 0, keyr = 0, rootF = kNoneDev, rootR = kNoneDev;
        if (active) {
          if (var == kVarFwd) V = window_fwd(pk, p, L);
          else V = revcomp_bits(window_fwd(var == kVarRevT ? pk : pka, len - p - L, L), L);
          keyf = (uint32_t)(V >> (2 * pw)); keyr = (uint32_t)(V & ((1ull << (2 * pw)) - 1));
          rootF = __ldg(&ix.lookup[keyf]).x;                                      // paralleltraversal.cpp:161
          ++n_windows;
        }
        bool zero = false;
        // forward sub-search for all 32 windows, then the mirror sub-search for those without a 0-error hit (:188)
        coop_subsearch<INSTR>(ix, sm, rootF, rev_chars(keyr, pw), full, lh, zero, st);   // P = w[9..18) ascending
        if (active && !zero) rootR = __ldg(&ix.lookup[keyr]).y;                   // :215
        coop_subsearch<INSTR>(ix, sm, rootR, keyf, full, lh, zero, st);                   // P = w[8..0] descending
        if (lh.overflow) flags |= kOvfSeedLane;
        const uint32_t n = lh.overflow ? 0u : lh.n;
        const uint32_t incl = warp_incl_scan_u32(n), tot = __shfl_sync(kFull, incl, 31);
        if (total + tot > region_cap) { flags |= kOvfSeedRegion; }
        else {
          const size_t base = region + total + incl - n;
          for (uint32_t k = 0; k < n; ++k) {
            const uint32_t id = lh.buf[k * lh.stride];
            b.hits[base + k] = make_uint2(id, p | (var << 24));
            cost += __ldg(ix.pos_off + id + 1) - __ldg(ix.pos_off + id);
          }
        }
        total += tot;
        __syncwarp();
      }
    }
    flags = __reduce_or_sync(kFull, flags);
    cost = warp_sum_u32(min(cost, 1u << 24));
    if (lane == 0) {
      b.hit_cnt[cnt_idx] = (flags & kOvfSeedRegion) ? 0u : total;
      if (flags) atomicOr(&b.flags[r], flags);
      if (total) atomicAdd(&b.cost[r - b.r0], max(cost, 1u));
    }
  }
  // instrumentation + num_short (processor.cpp:113)
  const uint32_t ns = warp_sum_u32(n_short);
  if (lane == 0 && ns) atomicAdd(&b.counters[dcNumShort], (unsigned long long)ns);
  if (INSTR) {
    const uint64_t w = warp_sum_u64(n_windows), nn = warp_sum_u64(st.nodes), nb = warp_sum_u64(st.buckets), ne = warp_sum_u64(st.entries);
    if (lane == 0) {
      atomicAdd(&b.counters[dcWindows], (unsigned long long)w); atomicAdd(&b.counters[dcNodes], (unsigned long long)nn);
      atomicAdd(&b.counters[dcBuckets], (unsigned long long)nb); atomicAdd(&b.counters[dcEntries], (unsigned long long)ne);
    }
  }
}

// heaviest-first schedule for the candidate kernel: a few reads carry thousands of Smith-Waterman calls
// (16S/23S conserved regions vote for thousands of references), so reads are binned by log2 of their
// estimated work and the persistent warps drain the bins from the heaviest down.
__global__ void bin_kernel(DevBatch b) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.nreads) return;
  const uint32_t c = b.cost[i];
  if (c == 0 || b.flags[b.r0 + i]) return;
  const uint32_t bin = min((uint32_t)(kCostBins - 1), 31u - (uint32_t)__clz(c));
  const uint32_t slot = atomicAdd(&b.bin_count[bin], 1u);
  b.bins[(size_t)bin * b.cnt_stride + slot] = b.r0 + i;
}

// unit-test kernel: explicit windows through the SAME cooperative path as seed_kernel (mode 0), or through
// the per-lane DFS fallback only (mode 1) (smr_debug_seed_windows)
__global__ void __launch_bounds__(kSeedWarpsPerCta * 32)
seed_debug_kernel(DevIndex ix, const uint8_t* seq03, const uint32_t* seq_off, const uint32_t* win_read, const uint32_t* win_pos, uint32_t nwin,
                  uint32_t* ids, uint32_t cap, uint32_t* counts, uint8_t* zero, int full_search, int mode) {
  __shared__ CoopSmem s_coop[kSeedWarpsPerCta];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;   // whole warps stay together: nwin is padded by the launch
  const bool active = k < nwin;
  const uint32_t pw = ix.partialwin;
  uint64_t V = 0;
  if (active) {
    const uint8_t* sq = seq03 + seq_off[win_read[k]] + win_pos[k];
    for (uint32_t i = 0; i < ix.lnwin; ++i) V = (V << 2) | (sq[i] & 3u);
  }
  LaneHits lh; lh.buf = ids + (size_t)(active ? k : 0) * cap; lh.stride = 1; lh.cap = active ? cap : 0; lh.n = 0; lh.overflow = false;
  SeedStats st{0, 0, 0};
  bool z = false;
  const bool full = full_search != 0;
  if (mode == 0) {
    CoopSmem& sm = s_coop[threadIdx.x >> 5];
    const uint32_t keyf = (uint32_t)(V >> (2 * pw)), keyr = (uint32_t)(V & ((1ull << (2 * pw)) - 1));
    const uint32_t rootF = active ? __ldg(&ix.lookup[keyf]).x : kNoneDev;
    coop_subsearch<false>(ix, sm, rootF, rev_chars(keyr, pw), full, lh, z, st);
    const uint32_t rootR = (active && !z) ? __ldg(&ix.lookup[keyr]).y : kNoneDev;
    coop_subsearch<false>(ix, sm, rootR, keyf, full, lh, z, st);
  } else if (active) z = seed_window<false>(ix, V, full, lh, st);
  if (active) { counts[k] = lh.n; zero[k] = z ? 1 : 0; }
}

}  // namespace smr
