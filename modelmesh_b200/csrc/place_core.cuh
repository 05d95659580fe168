// place_core.cuh — the placement decision in rank space, written once for two cooperative shapes:
//   * Coop32<NWL>: one 32-lane warp per decision on sm_100a.  The decision's exclusion-bitmap row has been staged in
//     shared memory by a TMA bulk copy; lane l owns the NWL consecutive 32-bit words [l*NWL, (l+1)*NWL) of it in
//     registers (a contiguous range of NWL*32 ranks), reductions are REDUX / SHFL / VOTE.
//   * Coop1: a single "lane" holding the whole row; compiled by g++ into the CPU-only test harness (tests/emul) so the
//     bitmask formulation can be checked against the oracle without a GPU.  It is NOT part of the shipped library.
//
// What is computed (reference: CacheMissForwardingLB.getNext, ModelMesh.java:4776-5004; quirk labels N1.. are those of
// SURVEY.md §8a / oracle/mm_oracle.cpp):
//   F      = cand[type] & ~excl[model] & ~extra (& ~replicaset-excluded, retried without if empty)   MM:4760-4805
//   best   = first set bit of F  (= argmin under PLACEMENT_ORDER, ranks are assigned at commit)       MM:4806
//   non-simple (a)/(b) preferred-instance handling as first-set / range queries                        MM:4822-4887
//   cut    = first rank in S whose walk test fails (MM:4913-4928, literal N2 semantics)               MM:4901-4937
//   shortlist = {best} ∪ (S below cut); rpm filter (MM:4957-4980); hash-indexed pick (MM:4981-4986, N4)
//
// The routine is written to be issue-efficient on the GPU: every pass over the row costs NWL warp instructions per
// operation, so passes are few (load, find-first, restrict, classify, count) and single-rank questions ("is self in the
// filtered set?") are answered in O(1) from the staged row instead of by a pass.
#pragma once
#include <stdint.h>

#include "../../include/mmplace.h"

#if defined(__CUDACC__)
#define MMP_HD __host__ __device__ __forceinline__
#define MMP_D __device__ __forceinline__
#else
#define MMP_HD inline
#endif

namespace mmp {

static constexpr uint32_t NONE_RANK = 0x7fffffffu;
static constexpr int32_t TARGET_INVALID = -3;  // malformed decision (bad model/self index or no fresh row for a non-live self)

struct RankRow {  // one per PLACEMENT_ORDER rank, 32 bytes
  int64_t lru;    // published lruTime (IR:37)
  int64_t rem;    // getRemaining() (IR:203-205)
  int32_t count;  // IR:39
  int32_t rpm;    // IR:51
  int32_t idx;    // instance index
  uint32_t flags; // bit0: isFull(rem)
};
struct WordSumI { int32_t lo, hi; };  // min/max count over the 32 ranks of a bitmap word
struct WordSumL { int64_t lo, hi; };  // min/max lruTime
struct FreshRow { int64_t lru, rem; int32_t count, rpm; };  // getFreshInstanceRecord() (MM:5369-5386), what the walk reads of it

struct SnapshotView {  // pointers into HBM (or host vectors in the CPU harness)
  int32_t n_ranks, row_words, n_models, max_instances;
  int32_t any_rs, n_type_ids;
  int64_t min_space;
  const uint32_t *excl;        // [n_models][row_words] loaded ∪ failed, bit = rank
  const uint32_t *cand;        // [n_slots][row_words]  allowed(type) ∧ active
  const uint32_t *candx;       // [n_slots][row_words]  cand ∧ ¬(likely-replaced replicaset members)  (MM:4769-4770)
  const uint32_t *pref;        // [n_slots][row_words]
  const uint8_t *has_pref;     // [n_slots]
  const uint16_t *type_slot;   // [n_type_ids]
  const uint32_t *full;        // [row_words] isFull(remaining) (MM:4640-4642)
  const RankRow *rows;         // [n_ranks]
  const int32_t *rank_of;      // [max_instances]
  const WordSumI *csum;        // [row_words]
  const WordSumL *lsum;        // [row_words]
  const mmp_model_row *models; // [n_models]
};

// ---- Java-semantics helpers (wrapping arithmetic, truncating division, saturating double->int) ----
MMP_HD int64_t jsub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
MMP_HD int32_t jaddi(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
MMP_HD int32_t jmuli(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
MMP_HD int64_t age_of(int64_t t, int64_t now) { return t == 0 ? 0 : jsub(now, t); }  // MM:4162-4164
MMP_HD int32_t jd2i(double d) {
  if (d != d) return 0;
  if (d >= 2147483647.0) return 2147483647;
  if (d <= -2147483648.0) return (int32_t)0x80000000;
  return (int32_t)d;
}
MMP_HD double jmul_d(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);  // a plain IEEE multiply, never contracted
#else
  volatile double r = a * b;
  return r;
#endif
}
MMP_HD uint64_t hash64(uint64_t seed, uint64_t decision_id) {  // replaces ThreadLocalRandom (N4); same as the oracle's
  uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (decision_id + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
// index in [0, n) from the hash: multiply-shift on the high 32 bits (cheap on the GPU; the contract of N4)
MMP_HD uint32_t hash_index(uint64_t seed, uint64_t decision_id, uint32_t n) {
  return (uint32_t)(((hash64(seed, decision_id) >> 32) * (uint64_t)n) >> 32);
}
MMP_HD int ffs32(uint32_t x) {  // index of lowest set bit, x != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
MMP_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
// position of the (n+1)-th set bit of w (n < popc(w))
MMP_HD int nth_bit(uint32_t w, uint32_t n) {
#if defined(__CUDA_ARCH__)
  return (int)__fns(w, 0, (int)n + 1);
#else
  for (uint32_t i = 0; i < n; i++) w &= w - 1;
  return __builtin_ctz(w);
#endif
}
// 0xffffffff << t with t clamped to [0, 32] (32 -> 0).  PTX shl clamps the shift amount, C does not.
MMP_HD uint32_t shl_ones(int32_t t) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  uint32_t tt = (uint32_t)max(t, 0);
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(0xffffffffu), "r"(tt));
  return r;
#else
  return t <= 0 ? 0xffffffffu : (t >= 32 ? 0u : (0xffffffffu << t));
#endif
}
// bits of the word whose first rank is `rank_base` that lie strictly above rank `lo` / strictly below rank `hi`
MMP_HD uint32_t mask_above(uint32_t rank_base, uint32_t lo) { return shl_ones((int32_t)lo - (int32_t)rank_base + 1); }
MMP_HD uint32_t mask_below(uint32_t rank_base, uint32_t hi) { return ~shl_ones((int32_t)hi - (int32_t)rank_base); }

// ---- the single-lane cooperative shape (CPU harness) ----
struct Coop1 {
  static constexpr int L = 1;
  static constexpr int NW_CAP = 2048;  // 65536 instances
  int nwl_;
  explicit Coop1(int row_words) : nwl_(row_words) {}
  MMP_HD int nwl() const { return nwl_; }
  MMP_HD int lane() const { return 0; }
  MMP_HD uint32_t wbase() const { return 0; }
  MMP_HD uint32_t rmin(uint32_t x) const { return x; }
  MMP_HD uint32_t rsum(uint32_t x) const { return x; }
  MMP_HD int32_t rmin_i(int32_t x) const { return x; }
  MMP_HD bool rany(bool p) const { return p; }
  MMP_HD uint32_t exscan(uint32_t) const { return 0; }
  template <class F> MMP_HD uint32_t eval_word(uint32_t wi, int32_t n_ranks, F &&f) const {
    uint32_t m = 0;
    for (int b = 0; b < 32; b++) {
      uint32_t r = wi * 32 + b;
      if ((int32_t)r < n_ranks && f(r)) m |= 1u << b;
    }
    return m;
  }
  MMP_HD void load_andnot(uint32_t *f, const uint32_t *a, const uint32_t *b) const {
    for (int k = 0; k < nwl_; k++) f[k] = a[k] & ~b[k];
  }
  MMP_HD void store_row(uint32_t *dst, const uint32_t *f) const {
    for (int k = 0; k < nwl_; k++) dst[k] = f[k];
  }
};

#if defined(__CUDACC__)
// ---- the warp cooperative shape: lane l owns words [l*NWL, (l+1)*NWL) ----
template <int NWL_>
struct Coop32 {
  static constexpr int L = 32;
  static constexpr int NW_CAP = NWL_;
  int lane_;
  MMP_D Coop32() : lane_(threadIdx.x & 31) {}
  MMP_D int nwl() const { return NWL_; }
  MMP_D int lane() const { return lane_; }
  MMP_D uint32_t wbase() const { return (uint32_t)lane_ * NWL_; }
  MMP_D uint32_t rmin(uint32_t x) const { return __reduce_min_sync(0xffffffffu, x); }
  MMP_D uint32_t rsum(uint32_t x) const { return __reduce_add_sync(0xffffffffu, x); }
  MMP_D int32_t rmin_i(int32_t x) const { return __reduce_min_sync(0xffffffffu, x); }
  MMP_D bool rany(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
  MMP_D uint32_t exscan(uint32_t x) const {
    uint32_t v = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane_ >= o) v += t;
    }
    return v - x;
  }
  template <class F> MMP_D uint32_t eval_word(uint32_t wi, int32_t n_ranks, F &&f) const {
    uint32_t r = wi * 32 + lane_;
    bool p = (int32_t)r < n_ranks && f(r);
    return __ballot_sync(0xffffffffu, p);
  }
  // f = a & ~b over this lane's words; a: type mask in global memory (L1/L2 resident), b: the staged exclusion row
  // (shared memory).  Widest vector the lane's byte offset (lane*NWL*4) allows.
  MMP_D void load_andnot(uint32_t *f, const uint32_t *a, const uint32_t *b) const {
    const uint32_t w0 = wbase();
    if constexpr (NWL_ % 4 == 0) {
#pragma unroll
      for (int k = 0; k < NWL_; k += 4) {
        uint4 x = __ldg(reinterpret_cast<const uint4 *>(a + w0 + k));
        uint4 y = *reinterpret_cast<const uint4 *>(b + w0 + k);
        f[k] = x.x & ~y.x; f[k + 1] = x.y & ~y.y; f[k + 2] = x.z & ~y.z; f[k + 3] = x.w & ~y.w;
      }
    } else if constexpr (NWL_ % 2 == 0) {
#pragma unroll
      for (int k = 0; k < NWL_; k += 2) {
        uint2 x = __ldg(reinterpret_cast<const uint2 *>(a + w0 + k));
        uint2 y = *reinterpret_cast<const uint2 *>(b + w0 + k);
        f[k] = x.x & ~y.x; f[k + 1] = x.y & ~y.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NWL_; k++) f[k] = __ldg(a + w0 + k) & ~b[w0 + k];
    }
  }
  MMP_D void store_row(uint32_t *dst, const uint32_t *f) const {
    const uint32_t w0 = wbase();
#pragma unroll
    for (int k = 0; k < NWL_; k++) dst[w0 + k] = f[k];
  }
};
#endif

#if defined(__CUDA_ARCH__)
#define MMP_UNROLL _Pragma("unroll")
#define MMP_FOR_K(co, k) _Pragma("unroll") for (int k = 0; k < C::NW_CAP; ++k)
#define MMP_FOR_K_DESC(co, k) _Pragma("unroll") for (int k = C::NW_CAP - 1; k >= 0; --k)
#else
#define MMP_UNROLL
#define MMP_FOR_K(co, k) for (int k = 0; k < (co).nwl(); ++k)
#define MMP_FOR_K_DESC(co, k) for (int k = (co).nwl() - 1; k >= 0; --k)
#endif

// rank of the first set bit of f & m(k, rank_base) over the whole row, NONE_RANK if none
template <class C, class M> MMP_HD uint32_t first_set_where(const C &co, const uint32_t *f, M &&m) {
  const uint32_t rb0 = co.wbase() * 32u;
  uint32_t w = 0, kk = 0;
  MMP_FOR_K_DESC(co, k) {
    uint32_t x = f[k] & m(k, rb0 + (uint32_t)k * 32u);
    if (x) { w = x; kk = (uint32_t)k; }
  }
  uint32_t r = w ? rb0 + kk * 32u + (uint32_t)ffs32(w) : NONE_RANK;
  return co.rmin(r);
}
template <class C> MMP_HD uint32_t first_set(const C &co, const uint32_t *f) {
  return first_set_where(co, f, [](int, uint32_t) { return 0xffffffffu; });
}
template <class C> MMP_HD void clear_bit(const C &co, uint32_t *f, uint32_t rank) {
  const uint32_t w = rank >> 5;
  MMP_FOR_K(co, k) { if (co.wbase() + (uint32_t)k == w) f[k] &= ~(1u << (rank & 31)); }
}
// rank of the kth (0-based) set bit in ascending rank order; `mine` = popcount of this lane's words; kth < total
template <class C> MMP_HD uint32_t select_kth(const C &co, const uint32_t *f, uint32_t mine, uint32_t kth) {
  const uint32_t pre = co.exscan(mine);
  uint32_t r = NONE_RANK;
  if (kth >= pre && kth < pre + mine) {
    uint32_t rem = kth - pre;
    bool found = false;
    MMP_FOR_K(co, k) {
      uint32_t pc = (uint32_t)popc32(f[k]);
      if (!found) {
        if (rem < pc) { r = (co.wbase() + (uint32_t)k) * 32u + (uint32_t)nth_bit(f[k], rem); found = true; }
        else rem -= pc;
      }
    }
  }
  return co.rmin(r);
}

// First rank in f whose per-rank predicate holds.  `cls(wi)` classifies a 32-rank word from its min/max summary:
// 0 = no rank violates, 1 = every rank violates, 2 = mixed; `eval(rank)` is the exact per-rank test; `word(wi)` returns
// word wi of f to every lane (recomputed from the staged row, so no register indexing).  Words are resolved in
// ascending order and the search stops at the first hit, so the common case (a sorted fleet: one mixed word) costs
// one cooperative word evaluation.
template <class C, class CLS, class EV, class WORD>
MMP_HD uint32_t first_violator(const C &co, const uint32_t *f, int32_t n_ranks, CLS &&cls, EV &&eval, WORD &&word) {
  const uint32_t w0 = co.wbase();
  uint32_t A = NONE_RANK, M = NONE_RANK;
  MMP_FOR_K_DESC(co, k) {  // descending: the lowest class-1 hit and the lowest mixed word of this lane are what remain
    if (f[k]) {
      int c = cls(w0 + (uint32_t)k);
      if (c == 1) A = (w0 + (uint32_t)k) * 32u + (uint32_t)ffs32(f[k]);
      else if (c == 2) M = w0 + (uint32_t)k;
    }
  }
  uint32_t Amin = co.rmin(A), Mmin = co.rmin(M);
  while (Mmin != NONE_RANK && Mmin * 32u < Amin) {
    uint32_t vm = co.eval_word(Mmin, n_ranks, eval) & word(Mmin);
    if (vm) { uint32_t r = Mmin * 32u + (uint32_t)ffs32(vm); if (r < Amin) Amin = r; break; }
    if (M == Mmin) {  // this lane owned it: advance to its next mixed word
      M = NONE_RANK;
      MMP_FOR_K_DESC(co, k) {
        uint32_t wi = w0 + (uint32_t)k;
        if (wi > Mmin && f[k] && cls(wi) == 2) M = wi;
      }
    }
    Mmin = co.rmin(M);
  }
  return Amin;
}

struct DecideOut {
  int32_t target, n_candidates;
  int32_t best, n_remaining, pick_index, flags, cut_rank, best_rank;
};

// rpm-filter predicate of MM:4966-4974 for one recorded rpm
struct RpmFilter {
  int64_t ago;
  int32_t min_load, m11, m15;
  MMP_HD void init(int32_t min_rpm, int64_t last_used_ago) {
    ago = last_used_ago;
    min_load = min_rpm > 100 ? min_rpm : 100;                       // Math.max(100, instReqLoad.min())
    m11 = jd2i(jmul_d(1.1, (double)min_load));
    m15 = jd2i(jmul_d(1.5, (double)min_load));
  }
  MMP_HD bool drop(int32_t rpm) const {
    return rpm >= 100 && ((ago < -1000 && rpm > m11) || (ago < 5000 && rpm > m15) ||
                          (ago < 720000 && rpm > jmuli(min_load, 3)) || (ago < 86400000 && rpm > jmuli(min_load, 4)));
  }
};

// Everything about one decision that does not need its bitmap row (72 bytes).  k_place lets lane j prepare the
// context of decision j of a 32-decision batch (the dependent gathers overlap across lanes) and stages it in shared memory.
struct DecisionCtx {
  mmp_decision_in d;
  int64_t last_used;
  FreshRow fr;         // the caller's fresh record (MM:5369), or its published row with rpm 0 (N7)
  int32_t self_rank;
  int32_t slot;        // type-constraint mask slot, -1 = malformed decision
};

MMP_HD void prepare_ctx(const SnapshotView &s, const mmp_decision_in &d, const FreshRow *fresh_tab, int32_t n_fresh,
                        DecisionCtx &c) {
  c.d = d; c.slot = -1; c.self_rank = -1; c.last_used = 0;
  c.fr.lru = 0; c.fr.rem = 0; c.fr.count = 0; c.fr.rpm = 0;
  if (d.model < 0 || d.model >= s.n_models || d.self < 0 || d.self >= s.max_instances) return;
  const mmp_model_row mr = s.models[d.model];
  const int tid = mr.type_id < s.n_type_ids ? mr.type_id : 0;
  c.last_used = (d.flags & MMP_DF_MODEL_LAST_USED) ? mr.last_used : d.last_used;
  c.self_rank = s.rank_of[d.self];
  if (d.fresh >= 0 && d.fresh < n_fresh) c.fr = fresh_tab[d.fresh];
  else if (c.self_rank >= 0) { const RankRow sr = s.rows[c.self_rank]; c.fr.lru = sr.lru; c.fr.rem = sr.rem; c.fr.count = sr.count; c.fr.rpm = 0; }
  else return;
  c.slot = s.type_slot[tid];
}

// One getNext.  f: caller-provided array of C::NW_CAP words (registers on the GPU).  erow: this decision's exclusion
// row, readable by every lane (shared memory on the GPU) until the routine returns.
// cand_rows (optional): [2][row_words] receives the candidate mask (other than best) and the survivor mask.
template <class C>
MMP_HD void decide_ctx(const SnapshotView &s, const DecisionCtx &c, const uint32_t *erow, const int32_t *extra, int64_t now,
                       uint64_t seed, uint64_t decision_id, const C &co, uint32_t *f, DecideOut &o, uint32_t *cand_rows) {
  o.target = MMP_TARGET_NONE; o.n_candidates = 0; o.best = -1; o.n_remaining = 0; o.pick_index = 0; o.flags = 0;
  o.cut_rank = (int32_t)NONE_RANK; o.best_rank = -1;
  const int RW = s.row_words;
  if (c.slot < 0) { o.target = TARGET_INVALID; return; }
  const mmp_decision_in &d = c.d;
  const int slot = c.slot;
  const bool favour_self = (d.flags & MMP_DF_FAVOUR_SELF) != 0;
  const int32_t self_rank = c.self_rank;
  const FreshRow fr = c.fr;
  const uint32_t *CAND = s.cand + (size_t)slot * RW;
  const uint32_t *P = s.pref + (size_t)slot * RW;
  const int n_extra = d.extra_n < 16 ? d.extra_n : 16;
  const uint32_t w0 = co.wbase();

  // ---- filter (MM:4760-4771): candx already excludes likely-replaced replicaset members ----
  const uint32_t *CX = s.any_rs ? s.candx + (size_t)slot * RW : CAND;
  auto apply_extra = [&]() {
    for (int e = 0; e < n_extra; e++) {
      int32_t x = extra[d.extra_off + e];
      if (x >= 0 && x < s.max_instances) { int32_t r = s.rank_of[x]; if (r >= 0) clear_bit(co, f, (uint32_t)r); }
    }
  };
  co.load_andnot(f, CX, erow);
  if (n_extra) apply_extra();
  uint32_t b = first_set(co, f);
  if (b == NONE_RANK && s.any_rs) {
    // MM:4798-4802: nothing survives; run the filter again without the replicaset exclusion
    o.flags |= MMP_TF_RS_RETRY;
    CX = CAND;
    co.load_andnot(f, CX, erow);
    if (n_extra) apply_extra();
    b = first_set(co, f);
  }
  if (b == NONE_RANK) return;  // null
  // word wi of the filtered set F, recomputed from the staged row for every lane (uniform loads, no register indexing)
  auto f_word = [&](uint32_t wi) -> uint32_t {
    uint32_t m = CX[wi] & ~erow[wi];
    for (int e = 0; e < n_extra; e++) {
      int32_t x = extra[d.extra_off + e];
      if (x >= 0 && x < s.max_instances) { int32_t r = s.rank_of[x]; if (r >= 0 && (uint32_t)(r >> 5) == wi) m &= ~(1u << (r & 31)); }
    }
    return m;
  };
  auto in_filter = [&](uint32_t r) -> bool { return (f_word(r >> 5) >> (r & 31)) & 1u; };
  auto pref_bit = [&](uint32_t r) -> bool { return (P[r >> 5] >> (r & 31)) & 1u; };

  const RankRow rb = s.rows[b];  // bestEntry.getValue()
  bool us = rb.idx == d.self;    // excluded self never passes the filter, so !excludeSelf is implied
  int64_t best_rem = us ? fr.rem : rb.rem, best_lru = us ? fr.lru : rb.lru;
  int32_t best_count = us ? fr.count : rb.count, best_rpm = us ? fr.rpm : rb.rpm, best_idx = rb.idx;
  uint32_t best_rank = b;
  const bool best_full = best_rem < s.min_space;
  if (best_full) o.flags |= MMP_TF_BEST_FULL;
  const bool has_pref = s.has_pref[slot] != 0;
  bool simple = !has_pref || pref_bit(b);
  uint32_t lo = b, hi = NONE_RANK;
  bool use_pref = has_pref && simple;  // best is preferred: preference is treated as required (MM:4905-4907)
  o.best = best_idx; o.best_rank = (int32_t)b;

  if (!simple) {
    if (!best_full) {
      // non-simple (a) MM:4828-4852: first later entry that is preferred, unless a full one comes first
      uint32_t p1 = first_set_where(co, f, [&](int k, uint32_t rbk) { return P[w0 + k] & mask_above(rbk, b); });
      uint32_t f1 = first_set_where(co, f, [&](int k, uint32_t rbk) { return s.full[w0 + k] & ~P[w0 + k] & mask_above(rbk, b); });
      if (p1 < f1) {
        const RankRow rp = s.rows[p1];
        best_rank = p1; best_idx = rp.idx; best_rem = rp.rem; best_lru = rp.lru; best_count = rp.count; best_rpm = rp.rpm;
        us = rp.idx == d.self;
        lo = p1; use_pref = true;
        o.best = best_idx; o.best_rank = (int32_t)p1;
      } else hi = f1;
      simple = true;
    } else {
      // non-simple (b) MM:4853-4887
      const int64_t oldest = best_lru, a4 = age_of(oldest, now) / 4;
      auto viol = [&](int64_t l) { int64_t diff = jsub(l, oldest); return diff > 120000 && diff > a4; };
      MMP_FOR_K(co, k) { f[k] &= mask_above((w0 + (uint32_t)k) * 32u, b); }
      const uint32_t kb = first_violator(co, f, s.n_ranks,
          [&](uint32_t wi) { WordSumL m = s.lsum[wi]; return !viol(m.hi) ? 0 : (viol(m.lo) ? 1 : 2); },
          [&](uint32_t r) { return viol(s.rows[r].lru); },
          [&](uint32_t wi) { return f_word(wi) & mask_above(wi * 32u, b); });
      bool anyp = false;
      MMP_FOR_K(co, k) { if (f[k] & P[w0 + k] & mask_below((w0 + (uint32_t)k) * 32u, kb)) anyp = true; }
      if (co.rany(anyp)) {
        // only preferred instances within the age distance are candidates; each records its own published rpm
        o.flags |= MMP_TF_PREF_B;
        MMP_FOR_K(co, k) { f[k] &= P[w0 + k] & mask_below((w0 + (uint32_t)k) * 32u, kb); }
        if (cand_rows) co.store_row(cand_rows, f);
        const bool self_in = self_rank >= 0 && (uint32_t)self_rank > b && (uint32_t)self_rank < kb && pref_bit((uint32_t)self_rank) &&
                             in_filter((uint32_t)self_rank);
        if (self_in && favour_self) { o.flags |= MMP_TF_FAVOUR_EXIT; return; }  // N8: returns null
        uint32_t mine = 0;
        MMP_FOR_K(co, k) { mine += (uint32_t)popc32(f[k]); }
        const int32_t ccount = (int32_t)co.rsum(mine);
        o.n_candidates = ccount;
        uint32_t chosen;
        if (ccount == 1) { chosen = first_set(co, f); o.n_remaining = 1; }
        else {
          int32_t remaining = ccount;
          const int64_t ago = age_of(c.last_used, now);
          if (ago < 432000000LL) {
            int32_t mn = 2147483647;
            MMP_FOR_K(co, k) { uint32_t w = f[k], wi = w0 + (uint32_t)k; while (w) { int bt = ffs32(w); w &= w - 1; int32_t v = s.rows[wi * 32 + bt].rpm; if (v < mn) mn = v; } }
            RpmFilter rf; rf.init(co.rmin_i(mn), ago);
            mine = 0;
            MMP_FOR_K(co, k) {
              uint32_t w = f[k], wi = w0 + (uint32_t)k;
              while (w) { int bt = ffs32(w); w &= w - 1; if (rf.drop(s.rows[wi * 32 + bt].rpm)) f[k] &= ~(1u << bt); }
              mine += (uint32_t)popc32(f[k]);
            }
            remaining = (int32_t)co.rsum(mine);
          }
          uint32_t index = remaining == 1 ? 0u : hash_index(seed, decision_id, (uint32_t)remaining);
          chosen = select_kth(co, f, mine, index);
          o.n_remaining = remaining; o.pick_index = (int32_t)index;
        }
        if (cand_rows) co.store_row(cand_rows + RW, f);
        int32_t cidx = s.rows[chosen].idx;
        o.target = (!favour_self && cidx == d.self) ? MMP_TARGET_SELF : cidx;
        return;
      }
      hi = kb; simple = true;  // no preferred in range: rewind, "no preference" logic over the replayed prefix
    }
  }
  // ---- simple case MM:4889-4938 ----
  o.flags |= MMP_TF_SIMPLE;
  if (us && favour_self) { o.flags |= MMP_TF_FAVOUR_EXIT; o.target = MMP_TARGET_SELF; return; }
  // S = F restricted to ranks in (lo, hi) and, when preference is binding, to preferred instances
  MMP_FOR_K(co, k) {
    const uint32_t rbk = (w0 + (uint32_t)k) * 32u;
    uint32_t m = mask_above(rbk, lo) & mask_below(rbk, hi);
    if (use_pref) m &= P[w0 + k];
    f[k] &= m;
  }
  const bool self_in_s = self_rank >= 0 && (uint32_t)self_rank > lo && (uint32_t)self_rank < hi &&
                         (!use_pref || pref_bit((uint32_t)self_rank)) && in_filter((uint32_t)self_rank);
  auto s_word = [&](uint32_t wi) -> uint32_t {  // word wi of S for every lane
    uint32_t m = f_word(wi) & mask_above(wi * 32u, lo) & mask_below(wi * 32u, hi);
    if (use_pref) m &= P[wi];
    return m;
  };
  const int64_t oldest = best_lru;
  bool c_self, self_viol;
  uint32_t cut_others = NONE_RANK;
  // the non-self walk test reads the caller's fresh record (N2), so it is one constant per decision
  if (best_full) {
    const int64_t a10 = age_of(oldest, now) / 10;
    int64_t df = jsub(fr.lru, oldest), ds = jsub(rb.lru, oldest);
    c_self = df > 45000 && df > a10;
    self_viol = ds > 45000 && ds > a10;
  } else {
    const int64_t q = best_rem >> 2;
    c_self = fr.rem < s.min_space || fr.rem < q;
    self_viol = rb.rem < s.min_space || rb.rem < q;
  }
  const int32_t thr = jaddi(best_count, best_count >> 2);
  auto cv = [&](int32_t cnt) { return cnt >= 10 && cnt > thr; };  // MM:4924-4927, always on the candidate's own count
  if (!best_full && self_in_s && cv(s.rows[self_rank].count)) self_viol = true;
  if (c_self) {
    // every non-self candidate fails: the walk stops at the first member of S other than self
    cut_others = first_set(co, f);
    if (self_in_s && cut_others == (uint32_t)self_rank)
      cut_others = first_set_where(co, f, [&](int, uint32_t rbk) { return mask_above(rbk, (uint32_t)self_rank); });
  } else if (!best_full) {
    // a self member that fails the count test is reported here too; it then also sets self_viol: same cut
    cut_others = first_violator(co, f, s.n_ranks,
        [&](uint32_t wi) { WordSumI m = s.csum[wi]; return !cv(m.hi) ? 0 : (cv(m.lo) ? 1 : 2); },
        [&](uint32_t r) { return cv(s.rows[r].count); }, s_word);
  }
  const uint32_t cut_self = (self_in_s && self_viol) ? (uint32_t)self_rank : NONE_RANK;
  const uint32_t cut = cut_others < cut_self ? cut_others : cut_self;
  o.cut_rank = (int32_t)cut;
  const bool self_in_sl = self_in_s && (uint32_t)self_rank < cut;
  if (favour_self && self_in_sl) { o.flags |= MMP_TF_FAVOUR_EXIT; o.target = MMP_TARGET_SELF; return; }
  uint32_t mine = 0;
  MMP_FOR_K(co, k) { f[k] &= mask_below((w0 + (uint32_t)k) * 32u, cut); mine += (uint32_t)popc32(f[k]); }  // f = candidates other than best
  const int32_t n_in = (int32_t)co.rsum(mine);
  const int32_t n_others = n_in - (self_in_sl ? 1 : 0);
  const int32_t ccount = 1 + n_in;
  o.n_candidates = ccount;
  if (cand_rows) co.store_row(cand_rows, f);
  bool keep_best = true, keep_others = true, keep_self = true;
  int32_t remaining = ccount;
  uint32_t index = 0;
  if (ccount > 1) {
    const int64_t ago = age_of(c.last_used, now);
    if (ago < 432000000LL) {  // FIVE_DAYS_MS
      int32_t mn = best_rpm;
      if (n_others > 0 && fr.rpm < mn) mn = fr.rpm;
      if (self_in_sl && rb.rpm < mn) mn = rb.rpm;
      RpmFilter rf; rf.init(mn, ago);
      keep_best = !rf.drop(best_rpm); keep_others = !rf.drop(fr.rpm); keep_self = !rf.drop(rb.rpm);
      remaining = (keep_best ? 1 : 0) + (keep_others ? n_others : 0) + ((self_in_sl && keep_self) ? 1 : 0);
    }
    index = remaining == 1 ? 0u : hash_index(seed, decision_id, (uint32_t)remaining);
  }
  o.n_remaining = remaining; o.pick_index = (int32_t)index;
  o.flags |= (keep_best ? MMP_TF_KEEP_BEST : 0) | (keep_others ? MMP_TF_KEEP_OTHERS : 0) | (keep_self ? MMP_TF_KEEP_SELF : 0);
  // survivors in rank order: best first (its rank precedes all of S), then S below the cut
  if (cand_rows) {  // trace only: materialise the survivor mask
    const uint32_t sw = (uint32_t)self_rank >> 5, sb = 1u << (self_rank & 31);
    MMP_FOR_K(co, k) {
      uint32_t v = keep_others ? f[k] : 0u;
      if (self_in_sl && w0 + (uint32_t)k == sw) v = keep_self ? (v | sb) : (v & ~sb);
      cand_rows[RW + w0 + k] = v;
    }
  }
  uint32_t chosen_rank;
  uint32_t kth = index;
  if (keep_best && kth == 0) chosen_rank = best_rank;
  else {
    if (keep_best) kth--;
    if (!keep_others) chosen_rank = (uint32_t)self_rank;  // the only other survivor can be the self candidate
    else {
      if (self_in_sl && !keep_self) {
        // skip the self candidate: it sits at position ps among the set bits of f
        uint32_t below = 0;
        MMP_FOR_K(co, k) { below += (uint32_t)popc32(f[k] & mask_below((w0 + (uint32_t)k) * 32u, (uint32_t)self_rank)); }
        const uint32_t ps = co.rsum(below);
        if (kth >= ps) kth++;
      }
      chosen_rank = select_kth(co, f, mine, kth);
    }
  }
  const int32_t cidx = chosen_rank == best_rank ? best_idx : s.rows[chosen_rank].idx;
  o.target = (!favour_self && cidx == d.self) ? MMP_TARGET_SELF : cidx;
}

}  // namespace mmp
