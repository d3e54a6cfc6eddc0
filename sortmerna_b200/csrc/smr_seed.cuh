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

namespace smr {

// Universal Levenshtein automaton d=1 as used by the reference (traverse_bursttrie.cpp:68-98):
// 15 states (14 = dead); table t, row bv, column state at kLevOff[t] + bv*14 + state.
__constant__ uint8_t c_lev[420] = {
    // t = 0 : 16 rows
    3,14,14,14,14,14,14,14,14,14,14,14,14,14,  3,14,14,14,14,14,14,14,14,14,14,14,14,14,
    7,14,14,14,4,4,4,4,14,14,14,14,14,14,      7,14,14,14,4,4,4,4,14,14,14,14,14,14,
    0,14,2,2,14,14,2,2,14,14,14,14,14,14,      0,14,2,2,14,14,2,2,14,14,14,14,14,14,
    0,14,2,2,4,4,6,6,14,14,14,14,14,14,        0,14,2,2,4,4,6,6,14,14,14,14,14,14,
    3,1,14,1,14,1,14,1,14,14,14,14,14,14,      3,1,14,1,14,1,14,1,14,14,14,14,14,14,
    7,1,14,1,4,5,4,5,14,14,14,14,14,14,        7,1,14,1,4,5,4,5,14,14,14,14,14,14,
    0,1,2,3,14,1,2,3,14,14,14,14,14,14,        0,1,2,3,14,1,2,3,14,14,14,14,14,14,
    0,1,2,3,4,5,6,7,14,14,14,14,14,14,         0,1,2,3,4,5,6,7,14,14,14,14,14,14,
    // t = 1 : 8 rows
    3,14,14,14,14,14,14,14,14,14,14,14,14,14,  13,14,14,14,10,10,10,10,14,14,14,14,14,14,
    8,14,2,2,14,14,2,2,14,14,14,14,14,14,      8,14,2,2,10,10,12,12,14,14,14,14,14,14,
    3,1,14,1,14,1,14,1,14,14,14,14,14,14,      13,1,14,1,10,11,10,11,14,14,14,14,14,14,
    8,1,2,3,14,1,2,3,14,14,14,14,14,14,        8,1,2,3,10,11,12,13,14,14,14,14,14,14,
    // t = 2 : 4 rows
    12,14,14,14,14,14,14,14,12,14,14,14,14,14, 9,14,10,10,14,14,10,10,9,14,14,14,10,10,
    12,1,14,1,14,1,14,1,12,14,14,1,14,1,       9,1,10,12,14,1,10,12,9,14,14,1,10,12,
    // t = 3 : 2 rows
    10,14,14,14,14,14,14,14,14,10,14,14,14,14, 10,10,14,10,14,10,14,10,14,10,14,14,10,14};

__device__ __forceinline__ uint32_t lev_off(uint32_t t) { return t == 0 ? 0u : (t == 1 ? 224u : (t == 2 ? 336u : 392u)); }

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

// characteristic bit masks of a 9-nt half window: for letter c, bit (pw+1-k) of field c is set iff
// p[k]==c, so that the reference's bit-vector row d (bitvector.cpp:56-132) is (M_c >> (pw-1-d)) & 15.
__device__ __forceinline__ uint64_t build_masks(uint32_t half, uint32_t pw, bool ascending) {
  // ascending: p[k] = base k of `half` (first base most significant); else p[k] = base pw-1-k
  uint64_t M = 0;
  for (uint32_t k = 0; k < pw; ++k) {
    const uint32_t c = ascending ? (half >> (2 * (pw - 1 - k))) & 3u : (half >> (2 * k)) & 3u;
    M |= 1ull << (16 * c + pw + 1 - k);
  }
  return M;
}

__device__ __forceinline__ uint32_t lev_next(const uint8_t* __restrict__ s_lev, uint64_t M, uint32_t pw, uint32_t depth,
                                             uint32_t c, uint32_t lev) {
  const uint32_t mc = (uint32_t)(M >> (16 * c)) & 0xFFFFu;
  if (depth < pw - 2) return s_lev[((mc >> (pw - 1 - depth)) & 15u) * 14u + lev];            // traverse_bursttrie.cpp:131-135
  const uint32_t t = 3 - pw + depth;                                                            // :136-139
  return s_lev[lev_off(t) + ((mc >> 2) & ((2u << (pw - depth)) - 1u)) * 14u + lev];
}

// DFS of one mini burst trie in lock step with the automaton (traverse_bursttrie.cpp:100-298).
// Returns true when a 0-error match ended the search of this window (accept_zero_kmer).
template <bool INSTR>
__device__ bool walk_trie(const DevIndex& ix, const uint8_t* __restrict__ s_lev, uint32_t root, uint64_t M, bool full_search,
                          LaneHits& hits, SeedStats& st) {
  const uint32_t pw = ix.partialwin;
  uint32_t stk_node[16];
  uint8_t stk_meta[16];
  uint32_t depth = 0, node = root, lev_in = 0, letter = 0;
  uint4 na = __ldg(ix.nodes + 2 * (size_t)node), nb = __ldg(ix.nodes + 2 * (size_t)node + 1);
  if (INSTR) st.nodes++;
  for (;;) {
    if (letter == 4) {
      if (depth == 0) return false;
      --depth;
      node = stk_node[depth]; letter = stk_meta[depth] & 7u; lev_in = stk_meta[depth] >> 4;
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
    if (flag == 0) { ++letter; continue; }
    const uint32_t lev = lev_next(s_lev, M, pw, depth, letter, lev_in);
    if (lev == 14) { ++letter; continue; }
    if (flag == 1) {
      stk_node[depth] = node; stk_meta[depth] = (uint8_t)((letter + 1) | (lev_in << 4));
      ++depth; node = w1; lev_in = lev; letter = 0;
      na = __ldg(ix.nodes + 2 * (size_t)node); nb = __ldg(ix.nodes + 2 * (size_t)node + 1);
      if (INSTR) st.nodes++;
      continue;
    }
    // bucket (:176-292)
    const uint32_t cnt = w0 >> 2, nchars = pw - depth;
    if (INSTR) st.buckets++;
    const uint2* __restrict__ e = ix.entries + w1;
    for (uint32_t k = 0; k < cnt; ++k) {
      const uint2 en = __ldg(e + k);
      if (INSTR) st.entries++;
      uint32_t tail = en.x, depth_b = depth, s = lev;
      bool local_accept = false, zero = false;
      for (uint32_t j = 0; j < nchars; ++j) {
        ++depth_b;
        s = lev_next(s_lev, M, pw, depth_b, tail & 3u, s);
        if (s == 14) break;
        if (depth_b >= pw - 2) {
          if (s >= 8) local_accept = true;                          // 1-error match (:232-235)
          if (depth_b == pw - 1 && s == 9) zero = !full_search;     // 0-error match (:237-246)
        }
        if (local_accept) {
          if (zero) {                                               // :256-262
            hits.n = 1; hits.buf[0] = en.y; hits.overflow = false;
            return true;
          }
          bool dup = false;                                         // :265-277
          for (uint32_t f = 0; f < hits.n && f < hits.cap; ++f) if (hits.buf[f * hits.stride] == en.y) { dup = true; break; }
          if (dup) break;
          if (hits.n < hits.cap) hits.buf[hits.n * hits.stride] = en.y; else hits.overflow = true;
          hits.n++;
        }
        tail >>= 2;
      }
    }
    ++letter;
  }
}

// both sub-searches of one window (paralleltraversal.cpp:129-249); V = the lnwin-mer, first base most significant
template <bool INSTR>
__device__ bool seed_window(const DevIndex& ix, const uint8_t* __restrict__ s_lev, uint64_t V, bool full_search, LaneHits& hits,
                            SeedStats& st) {
  const uint32_t pw = ix.partialwin;
  const uint32_t keyf = (uint32_t)(V >> (2 * pw)), keyr = (uint32_t)(V & ((1ull << (2 * pw)) - 1));
  hits.n = 0;
  bool zero = false;
  const uint32_t rootF = __ldg(&ix.lookup[keyf]).x;                               // :161
  if (rootF != kNoneDev) zero = walk_trie<INSTR>(ix, s_lev, rootF, build_masks(keyr, pw, true), full_search, hits, st);
  if (!zero) {                                                                    // :188
    const uint32_t rootR = __ldg(&ix.lookup[keyr]).y;                             // :215
    if (rootR != kNoneDev) zero = walk_trie<INSTR>(ix, s_lev, rootR, build_masks(keyf, pw, false), full_search, hits, st);
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

constexpr int kSeedWarpsPerCta = 8;
constexpr int kLaneHitCap = 16;   // per-window hit slots in shared memory (fast path)

// The seed kernel.  grid-stride over reads, one warp per read.
//   lane_hits_g: optional global per-lane buffers (retry path) [nwarps_total][cap_g][32]; nullptr -> shared memory, kLaneHitCap
template <bool INSTR>
__global__ void __launch_bounds__(kSeedWarpsPerCta * 32)
seed_kernel(DevIndex ix, DevBatch b, DevParams prm, uint32_t* lane_hits_g, uint32_t cap_g) {
  __shared__ uint8_t s_lev[420];
  __shared__ uint32_t s_hits[kSeedWarpsPerCta][kLaneHitCap][32];
  for (int i = threadIdx.x; i < 420; i += blockDim.x) s_lev[i] = c_lev[i];
  __syncthreads();
  const unsigned lane = lane_id();
  const uint32_t wic = threadIdx.x >> 5;
  const uint32_t warp = blockIdx.x * kSeedWarpsPerCta + wic, nwarps = gridDim.x * kSeedWarpsPerCta;
  const uint32_t L = ix.lnwin;
  LaneHits lh;
  if (lane_hits_g) { lh.buf = lane_hits_g + (size_t)warp * cap_g * 32 + lane; lh.stride = 32; lh.cap = cap_g; }
  else { lh.buf = &s_hits[wic][0][lane]; lh.stride = 32; lh.cap = kLaneHitCap; }
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
        if (active) {
          uint64_t V;
          if (var == kVarFwd) V = window_fwd(pk, p, L);
          else V = revcomp_bits(window_fwd(var == kVarRevT ? pk : pka, len - p - L, L), L);
          seed_window<INSTR>(ix, s_lev, V, prm.is_full_search != 0, lh, st);
          ++n_windows;
        }
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

// unit-test kernel: explicit windows, one lane per window (smr_debug_seed_windows)
__global__ void seed_debug_kernel(DevIndex ix, const uint8_t* seq03, const uint32_t* seq_off, const uint32_t* win_read,
                                  const uint32_t* win_pos, uint32_t nwin, uint32_t* ids, uint32_t cap, uint32_t* counts,
                                  uint8_t* zero, int full_search) {
  __shared__ uint8_t s_lev[420];
  for (int i = threadIdx.x; i < 420; i += blockDim.x) s_lev[i] = c_lev[i];
  __syncthreads();
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nwin) return;
  const uint8_t* s = seq03 + seq_off[win_read[k]] + win_pos[k];
  uint64_t V = 0;
  for (uint32_t i = 0; i < ix.lnwin; ++i) V = (V << 2) | (s[i] & 3u);
  LaneHits lh; lh.buf = ids + (size_t)k * cap; lh.stride = 1; lh.cap = cap; lh.n = 0; lh.overflow = false;
  SeedStats st{0, 0, 0};
  const bool z = seed_window<false>(ix, s_lev, V, full_search != 0, lh, st);
  counts[k] = lh.n; zero[k] = z ? 1 : 0;
}

}  // namespace smr
