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
// How a window is searched here.
//
// The reference walks the mini burst trie of the window's exact 9-mer half in lock step with a table-driven
// universal Levenshtein automaton for d=1 (traverse_bursttrie.cpp:68-298, bit-vectors bitvector.cpp:56-132),
// pruning sub-tries the automaton rejects.  That automaton accepts at depth d exactly when the d+1 text
// characters read so far are within one edit of the 9-nt other half, and reaches state 9 at depth 8 exactly
// on an exact match (proved against the table and against dynamic-programming edit distance on millions of
// cases: tests/test_lev_equivalence.py; the oracle keeps the table).  Pointer chasing and per-window DFS
// stacks are the wrong shape for a GPU (the first version of this kernel ran with 2 of 32 lanes active), so:
//   * at load time every mini trie is flattened into ONE contiguous list of its entries in the DFS order of
//     the reference, each entry carrying its full text = trie path letters + bucket tail (smr_index.h);
//   * a window search classifies EVERY entry of the list of its 9-mer with three bit-parallel predicates
//     (smr_levbits.h) -- entries in sub-tries the reference would have pruned simply classify as "no match",
//     so the outcome is the same, and the ~2.4x more entries cost less than the pruning did;
//   * the 32 windows of a round are searched together: their lists form one entry stream, one entry per lane
//     per step (coalesced 8-byte loads, no divergence); the few matching entries are compacted with
//     ballot/popc into a shared list and each lane then replays the reference's order-dependent rules
//     (0-error exit, per-window de-duplication, traverse_bursttrie.cpp:249-281) over its own matches.
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

struct SeedStats { uint32_t entries, lists; };   // list entries classified; non-empty flat lists scanned (the "buckets" of SURVEY 8(d): one per sub-search)

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

// class of a window position = the first pass whose grid contains it (paralleltraversal.cpp:118-131)
__device__ __forceinline__ uint32_t pass_class(uint32_t p, uint32_t s0, uint32_t s1, uint32_t s2) {
  if (p % s0 == 0) return 0;
  if (p % s1 == 0) return 1;
  if (p % s2 == 0) return 2;
  return 3;
}

constexpr int kAccCap = 192;   // matching entries buffered per flush

struct CoopSmem {              // per warp
  uint32_t id[kAccCap];
  uint8_t meta[kAccCap];       // owner lane | code << 5
};

// One sub-search (forward: trie_F list of the first half, pattern = second half; or mirror) for the 32 windows
// of a round.  off/cnt: the lane's list in ix.flist (cnt == 0: lane idle).  Appends to lh; sets zero.
template <bool INSTR>
__device__ void coop_flat(const DevIndex& ix, CoopSmem& sm, const uint32_t off, const uint32_t cnt, const uint32_t P, const bool full_search,
                          LaneHits& lh, bool& zero, SeedStats& st) {
  const unsigned lane = lane_id();
  const uint32_t pw = ix.partialwin;
  const uint32_t incl = warp_incl_scan_u32(cnt), E = __shfl_sync(kFull, incl, 31), excl = incl - cnt;
  if (INSTR) { st.entries += cnt; st.lists += cnt ? 1u : 0u; }
  uint32_t nacc = 0;
  for (uint32_t e0 = 0; e0 < E; e0 += 32) {
    const uint32_t e = e0 + lane;
    uint32_t lo = 0;                         // owner = first lane whose inclusive sum exceeds e
#pragma unroll
    for (int stp = 16; stp > 0; stp >>= 1) { const uint32_t v = __shfl_sync(kFull, incl, lo + stp - 1); if (v <= e) lo += stp; }
    lo = min(lo, 31u);
    const uint32_t ex_own = __shfl_sync(kFull, excl, lo), off_own = __shfl_sync(kFull, off, lo), P_own = __shfl_sync(kFull, P, lo);
    uint32_t code = 0, id = 0;
    if (e < E) {
      const uint2 en = __ldg(ix.flist + off_own + (e - ex_own));
      code = classify_bits(P_own, en.x, pw); id = en.y;
    }
    const unsigned m = __ballot_sync(kFull, (code & 3u) != 0);
    if ((code & 3u) != 0) { const uint32_t slot = nacc + __popc(m & ((1u << lane) - 1)); sm.id[slot] = id; sm.meta[slot] = (uint8_t)(lo | (code << 5)); }
    nacc += __popc(m);
    if (nacc + 32 > (uint32_t)kAccCap || e0 + 32 >= E) {   // flush: every lane replays its own matches, in list order
      __syncwarp();
      for (uint32_t i = 0; i < nacc && !zero; ++i) {
        const uint32_t mt = sm.meta[i];
        if ((mt & 31u) == lane) zero = apply_entry(mt >> 5, sm.id[i], full_search, lh);
      }
      nacc = 0;
      __syncwarp();
    }
  }
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
  SeedStats st{0, 0};
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
    // the windows of all strand variants form one sequence (variant-major), 32 per round
    const uint32_t nvar = hasn ? 3u : 2u;
    const uint32_t v_lo = do_fwd ? 0u : 1u, v_hi = do_rev ? nvar : 1u;      // variants searched: [v_lo, v_hi)
    const uint32_t nq = (v_hi - v_lo) * npos;
    for (uint32_t q0 = 0; q0 < nq; q0 += 32) {
      const uint32_t qq = q0 + lane;
      const uint32_t var = v_lo + qq / npos, p = (qq % npos) * step;
      bool active = qq < nq;
      if (active && step == 1) active = (p % s0 == 0) || (p % s1 == 0) || (p % s2 == 0);
      lh.n = 0; lh.overflow = false;
      uint32_t keyf = 0, keyr = 0;
      uint4 lk_f = make_uint4(0, 0, 0, 0);
      if (active) {
        uint64_t V;
        if (var == kVarFwd) V = window_fwd(pk, p, L);
        else V = revcomp_bits(window_fwd(var == kVarRevT ? pk : pka, len - p - L, L), L);
        keyf = (uint32_t)(V >> (2 * pw)); keyr = (uint32_t)(V & ((1ull << (2 * pw)) - 1));
        lk_f = __ldg(&ix.flookup[keyf]);                                        // paralleltraversal.cpp:161
        ++n_windows;
      }
      bool zero = false;
      // sub-search (a): exact first half, <= 1 error in the second half (P = w[9..18) ascending)
      coop_flat<INSTR>(ix, sm, lk_f.x, lk_f.y, rev_chars(keyr, pw), full, lh, zero, st);
      // sub-search (b), only without a 0-error hit (:188): exact second half, <= 1 error in the reversed first half
      uint4 lk_r = make_uint4(0, 0, 0, 0);
      if (active && !zero) lk_r = __ldg(&ix.flookup[keyr]);                      // :215
      coop_flat<INSTR>(ix, sm, lk_r.z, lk_r.w, keyf, full, lh, zero, st);
      if (lh.overflow) flags |= kOvfSeedLane;
      const uint32_t n = lh.overflow ? 0u : lh.n;
      const uint32_t incl = warp_incl_scan_u32(n), tot = __shfl_sync(kFull, incl, 31);
      if (total + tot > region_cap) { flags |= kOvfSeedRegion; }
      else {
        const size_t base = region + total + incl - n;
        for (uint32_t k = 0; k < n; ++k) {
          const uint32_t id = lh.buf[k * lh.stride];
          b.hits[base + k] = make_uint2(id, p | (var << 24) | (pass_class(p, s0, s1, s2) << 28));
          cost += __ldg(ix.pos_off + id + 1) - __ldg(ix.pos_off + id);
        }
      }
      total += tot;
      __syncwarp();
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
    const uint64_t w = warp_sum_u64(n_windows), ne = warp_sum_u64(st.entries), nl = warp_sum_u64(st.lists);
    if (lane == 0) { atomicAdd(&b.counters[dcWindows], (unsigned long long)w); atomicAdd(&b.counters[dcEntries], (unsigned long long)ne); atomicAdd(&b.counters[dcBuckets], (unsigned long long)nl); }
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

// unit-test kernel: explicit windows through the SAME cooperative path as seed_kernel (mode 0), or by one
// lane scanning its own list sequentially (mode 1) (smr_debug_seed_windows)
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
  SeedStats st{0, 0};
  bool z = false;
  const bool full = full_search != 0;
  const uint32_t keyf = (uint32_t)(V >> (2 * pw)), keyr = (uint32_t)(V & ((1ull << (2 * pw)) - 1));
  const uint32_t Pf = rev_chars(keyr, pw), Pr = keyf;
  if (mode == 0) {
    CoopSmem& sm = s_coop[threadIdx.x >> 5];
    const uint4 lf = active ? __ldg(&ix.flookup[keyf]) : make_uint4(0, 0, 0, 0);
    coop_flat<false>(ix, sm, lf.x, lf.y, Pf, full, lh, z, st);
    const uint4 lr = (active && !z) ? __ldg(&ix.flookup[keyr]) : make_uint4(0, 0, 0, 0);
    coop_flat<false>(ix, sm, lr.z, lr.w, Pr, full, lh, z, st);
  } else if (active) {
    const uint4 lf = __ldg(&ix.flookup[keyf]);
    for (uint32_t i = 0; i < lf.y && !z; ++i) { const uint2 en = __ldg(ix.flist + lf.x + i); z = apply_entry(classify_bits(Pf, en.x, pw), en.y, full, lh); }
    if (!z) {
      const uint4 lr = __ldg(&ix.flookup[keyr]);
      for (uint32_t i = 0; i < lr.w && !z; ++i) { const uint2 en = __ldg(ix.flist + lr.z + i); z = apply_entry(classify_bits(Pr, en.x, pw), en.y, full, lh); }
    }
  }
  if (active) { counts[k] = lh.n; zero[k] = z ? 1 : 0; }
}

}  // namespace smr
