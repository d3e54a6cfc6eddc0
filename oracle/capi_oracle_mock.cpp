// TEST INFRASTRUCTURE ONLY (oracle/): the entry points of include/smr_b200.h that the reference-side binding
// (integration/align_gpu.cpp) uses, implemented on top of the CPU oracle (smr_oracle.cpp) instead of the GPU library.
// Purpose: on a box WITHOUT a GPU the binding TU itself (feed order, KVDB keys and blobs, Readstats plumbing, index / reference
// hand-over) can be linked with the reference's unmodified host program and checked against the reference's own output.
// It is never built into, loaded by or shipped with the product: sortmerna_b200/libsmr_b200.so is the only implementation of the
// C ABI; this file only lets tests/test_integration_binding.py exercise align_gpu.cpp where no CUDA device exists.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include <vector>

#include "../include/smr_b200.h"
#include "smr_oracle.h"

struct smr_ctx {
  std::string err, dir;
  ora_params prm{};
  struct Part { ora_index* ix; uint16_t index_num, part; std::vector<uint8_t> refseq; std::vector<uint64_t> refoff; uint32_t nref, ms; uint32_t skip[3]; };
  std::vector<Part> parts;
  uint32_t n_index_files = 0;
  uint32_t all_slots = 16, need_slots = 0;
};

extern "C" {

int smr_device_count(void) { const char* e = getenv("SMR_MOCK_DEVICES"); return e ? atoi(e) : 1; }

int smr_init(int, smr_ctx** out) {
  smr_ctx* c = new smr_ctx;
  char tmpl[] = "/tmp/smr_mock_XXXXXX";
  if (!mkdtemp(tmpl)) { delete c; return SMR_ERR_NO_DEVICE; }
  c->dir = tmpl;
  *out = c;
  return SMR_OK;
}
void smr_destroy(smr_ctx* c) {
  if (!c) return;
  for (auto& p : c->parts) ora_index_free(p.ix);
  std::string cmd = "rm -rf '" + c->dir + "'";
  if (system(cmd.c_str())) {}
  delete c;
}
const char* smr_last_error(const smr_ctx* c) { return c ? c->err.c_str() : "null context"; }

int smr_set_params(smr_ctx* c, const smr_params* p) {
  static_assert(sizeof(smr_params) == sizeof(ora_params), "parameter blocks differ");
  memcpy(&c->prm, p, sizeof(ora_params));
  return SMR_OK;
}

int smr_set_instrumentation(smr_ctx* c, int) { return c ? SMR_OK : SMR_ERR_ARG; }
int smr_set_aln_slots(smr_ctx* c, uint32_t slots) { if (!c || !slots) return SMR_ERR_ARG; c->all_slots = slots; return SMR_OK; }
uint32_t smr_aln_slots(const smr_ctx* c) { return c->prm.num_alignments > 0 ? (uint32_t)c->prm.num_alignments : c->all_slots; }
uint32_t smr_aln_slots_needed(const smr_ctx* c) { return c->need_slots; }

int smr_load_index_part(smr_ctx* c, uint32_t index_num, uint32_t part, const void* kmer, size_t kb, const void* trie, size_t tb, const void* pos, size_t pb,
                        const uint8_t* refseq, const uint64_t* ref_off, uint32_t nref, uint32_t lnwin, uint32_t minimal_score, const uint32_t skip[3]) {
  const std::string pfx = c->dir + "/i" + std::to_string(index_num), sfx = "_" + std::to_string(part) + ".dat";
  const struct { const char* ext; const void* p; size_t n; } files[3] = {{".kmer", kmer, kb}, {".bursttrie", trie, tb}, {".pos", pos, pb}};
  for (auto& f : files) {
    FILE* fp = fopen((pfx + f.ext + sfx).c_str(), "wb");
    if (!fp || fwrite(f.p, 1, f.n, fp) != f.n) { c->err = "mock: cannot stage index file"; if (fp) fclose(fp); return SMR_ERR_INDEX; }
    fclose(fp);
  }
  char err[512] = {0};
  ora_index* ix = ora_index_load(pfx.c_str(), part, lnwin, err, sizeof(err));
  if (!ix) { c->err = err; return SMR_ERR_INDEX; }
  smr_ctx::Part pt;
  pt.ix = ix; pt.index_num = (uint16_t)index_num; pt.part = (uint16_t)part; pt.nref = nref; pt.ms = minimal_score;
  pt.refseq.assign(refseq, refseq + ref_off[nref]); pt.refoff.assign(ref_off, ref_off + nref + 1);
  memcpy(pt.skip, skip, 12);
  c->parts.push_back(std::move(pt));
  c->n_index_files = std::max(c->n_index_files, index_num + 1);
  return SMR_OK;
}

// the stand-in has no device: the binding's SMR_INDEX_DEVICE=1 path is a GPU test (tests/test_gpu_integration.py)
int smr_build_index_device(smr_ctx* c, uint32_t, const char*, uint32_t, uint32_t, uint32_t, double, const uint32_t*, uint32_t, uint32_t*, uint64_t*) {
  if (c) c->err = "mock: no device index builder";
  return SMR_ERR_UNSUPPORTED;
}

int smr_align_batch(smr_ctx* c, const uint8_t* seq_cat, const uint64_t* seq_off, uint32_t nreads, smr_read_result* results, smr_aln* alns,
                    uint32_t* cigar_pool, uint64_t cigar_cap, uint64_t* cigar_used, uint64_t* counters, uint32_t n_counters) {
  static_assert(sizeof(smr_read_result) == sizeof(ora_read_result) && sizeof(smr_aln) == sizeof(ora_aln), "result layouts differ");
  const uint32_t n = (uint32_t)c->parts.size();
  std::vector<const ora_index*> ix(n); std::vector<uint16_t> inum(n), ipart(n); std::vector<const uint8_t*> rs(n); std::vector<const uint64_t*> ro(n);
  std::vector<uint32_t> nref(n), ms(n), skip(3 * n);
  for (uint32_t k = 0; k < n; ++k) {
    auto& p = c->parts[k];
    ix[k] = p.ix; inum[k] = p.index_num; ipart[k] = p.part; rs[k] = p.refseq.data(); ro[k] = p.refoff.data(); nref[k] = p.nref; ms[k] = p.ms;
    memcpy(&skip[3 * k], p.skip, 12);
  }
  std::vector<uint64_t> matched(c->n_index_files, 0);
  ora_counters oc{};
  uint64_t used = 0;
  ora_set_aln_slots(c->all_slots);
  const int rc = ora_align(ix.data(), inum.data(), ipart.data(), n, c->n_index_files, rs.data(), ro.data(), nref.data(), ms.data(), skip.data(), &c->prm,
                           seq_cat, seq_off, nreads, (ora_read_result*)results, (ora_aln*)alns, cigar_pool, cigar_cap, &used, matched.data(), &oc, 4);
  if (rc == 2) c->need_slots = ora_aln_slots_needed();
  if (rc != 0) { c->err = "mock: ora_align failed"; return SMR_ERR_CAPACITY; }
  if (cigar_used) *cigar_used = used;
  if (counters) {
    if (n_counters > SMR_CNT_NUM_ALIGNED) counters[SMR_CNT_NUM_ALIGNED] += oc.num_aligned;
    if (n_counters > SMR_CNT_NUM_SHORT) counters[SMR_CNT_NUM_SHORT] += oc.num_short_last;
    for (uint32_t i = 0; i < c->n_index_files && SMR_CNT_FIXED + i < n_counters; ++i) counters[SMR_CNT_FIXED + i] += matched[i];
  }
  return SMR_OK;
}

}  // extern "C"
