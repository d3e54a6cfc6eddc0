/*
 * oracle/smr_oracle.h -- C interface of the CPU restatement of SortMeRNA's per-read alignment
 * hot path.  TEST INFRASTRUCTURE ONLY: nothing under sortmerna_b200/ may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * The result structs deliberately have the same layout as include/smr_b200.h so that the parity
 * tests can compare the GPU path and the oracle field by field.
 */
#ifndef SMR_ORACLE_H
#define SMR_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_index ora_index; /* one loaded (index, part) */

typedef struct {
  int32_t match, mismatch, score_N, gap_open, gap_ext; /* options.cpp:1684-1708 */
  int32_t num_seeds, min_lis, edges, edges_is_percent; /* options.cpp:1725-1738 */
  int32_t num_alignments, is_best;                     /* include/options.hpp:495,567 */
  int32_t is_forward, is_reverse, is_full_search;      /* processor.cpp:130-135, traverse_bursttrie.cpp:244 */
  int32_t minoccur;                                    /* include/options.hpp:572 (constant 0) */
} ora_params;

/* same layout as smr_read_result (include/smr_b200.h) */
typedef struct {
  uint32_t lastIndex, lastPart; /* read.cpp:435-436 */
  uint32_t hit_seeds;           /* read.cpp:446 */
  uint32_t min_index, max_index;/* alignment_struct2, ssw.hpp:157-159 */
  uint32_t n_align;             /* alignv.size() */
  uint16_t max_SW_count;        /* read.cpp:444 */
  uint8_t is_done, is_hit;      /* read.cpp:441-442 */
} ora_read_result;

/* same layout as smr_aln (include/smr_b200.h); one per stored alignment (s_align2, ssw.hpp:44-56) */
typedef struct {
  uint32_t cigar_off, cigar_len; /* into the cigar pool; BAM-style len<<4|op, op 0=M 1=I 2=D */
  uint32_t ref_num;
  int32_t ref_begin1, ref_end1, read_begin1, read_end1;
  uint32_t readlen;
  uint16_t score1, part, index_num;
  uint8_t strand, pad;
} ora_aln;

/* counters the reference keeps in Readstats (readstats.hpp:77-84) that this path mutates */
typedef struct {
  uint64_t num_aligned;
  uint64_t num_short_last; /* num_short is reset per index pass (processor.cpp:228): value of the LAST pass */
  uint64_t sw_calls;       /* instrumentation: number of ssw_align-equivalent calls */
  uint64_t sw_cells;       /* instrumentation: sum refLen*readLen of those calls */
  uint64_t windows;        /* instrumentation: seed windows searched */
  uint64_t trie_nodes;     /* instrumentation: trie nodes visited */
  uint64_t bucket_entries; /* instrumentation: bucket entries visited */
  uint64_t buckets;        /* instrumentation: buckets visited */
  uint64_t pos_entries;    /* instrumentation: position entries touched once per compute_lis call */
} ora_counters;

/* index.cpp:143-357 -- parse <prefix>.{kmer,bursttrie,pos}_<part>.dat */
ora_index* ora_index_load(const char* prefix, uint32_t part, uint32_t lnwin, char* err, size_t errlen);
void ora_index_free(ora_index*);
uint32_t ora_index_num_ids(const ora_index*);
/* structure statistics: out[0]=non-empty 9-mers, [1]=trie nodes, [2]=buckets, [3]=bucket entries,
   [4]=ids, [5]=positions, [6]=max bucket bytes, [7]=max positions per id */
void ora_index_stats(const ora_index*, uint64_t out[8]);

/* One window, both sub-searches (paralleltraversal.cpp:129-249 + traverse_bursttrie.cpp:100-298).
   seq03: read (or its reverse complement) in the 0..3 alphabet.  Returns number of id hits written
   (ids[] in id_hits order), *accept_zero = the accept_zero_kmer flag after the window. */
int ora_seed_window(const ora_index*, const uint8_t* seq03, uint32_t win_pos, int is_full_search,
                    int minoccur, uint32_t* ids, int max_ids, int* accept_zero);

/* ssw_init + ssw_align(flag=2, filters, filterd=0, maskLen=0) restated (ssw.c:788-941).
   read/ref in 0..4; mat = 5x5 row-major mat[ref*5+read].  Returns 0 on success.
   out[0]=score1 out[1]=ref_begin1 out[2]=ref_end1 out[3]=read_begin1 out[4]=read_end1 out[5]=cigarLen */
int ora_ssw_align(const int8_t* read, int32_t readLen, const int8_t* ref, int32_t refLen,
                  const int8_t* mat, int32_t gap_open, int32_t gap_ext, uint16_t filters,
                  int32_t out[6], uint32_t* cigar, int32_t max_cigar);

/*
 * The align() driver (processor.cpp:173-285) restated read-batch-wise: for every index (in order)
 * and every read, run align2's per-read body with KVDB-equivalent carry-over between index passes.
 *   idx[k]            k-th (index,part) in --ref order; index_num[k]/part[k] its numbers
 *   refseq[k]/refoff[k]  numeric (0..4) reference sequences of that part, concatenated + offsets (nref+1)
 *   minimal_score[k]  refstats.minimal_score[index_num[k]]; skiplengths[k*3..] the three pass shifts
 *   is_last_idx: computed as k == nidx-1
 *   reads: 0..4 (4 = ambiguous), concatenated + offsets (nreads+1)
 * Outputs: res[nreads], alns[nreads*slots] (slots = num_alignments, or the stride set by ora_set_aln_slots when it is 0:
 * "all alignments", alignment.cpp:420-424; rc 2 + ora_aln_slots_needed() when a read stored more), cigar pool (u32), per-db match counters.
 */
int ora_align(const ora_index* const* idx, const uint16_t* index_num, const uint16_t* part, uint32_t nidx,
              uint32_t n_index_files,
              const uint8_t* const* refseq, const uint64_t* const* refoff, const uint32_t* nref,
              const uint32_t* minimal_score, const uint32_t* skiplengths,
              const ora_params* prm,
              const uint8_t* reads, const uint64_t* readoff, uint32_t nreads,
              ora_read_result* res, ora_aln* alns, uint32_t* cigar_pool, uint64_t cigar_cap,
              uint64_t* cigar_used, uint64_t* reads_matched_per_db /* n_index_files */,
              ora_counters* counters, int nthreads);

void ora_set_aln_slots(uint32_t slots);
uint32_t ora_aln_slots_needed(void);

#ifdef __cplusplus
}
#endif
#endif
