// place_core.cuh — the placement decision in rank space, written once for two cooperative shapes:
//   * Coop32<V,NJ>: one 32-lane warp per decision on sm_100a; each lane holds NJ vectors of V 32-bit words of the
//     decision's exclusion-bitmap row in registers (V = 4: 128-bit loads), reductions are REDUX / SHFL / VOTE.
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
  const uint32_t *pref;        // [n_slots][row_words]
  const uint8_t *has_pref;     // [n_slots]
  const uint16_t *type_slot;   // [n_type_ids]
  const uint32_t *rs;          // [row_words] likely-replaced replicaset members (MM:4769-4770)
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
// bits of word `wi` whose rank is > lo (lo = NONE_RANK never used here)
MMP_HD uint32_t mask_above(uint32_t wi, uint32_t lo) {
  int32_t d = (int32_t)lo - (int32_t)(wi * 32u);
  return d < 0 ? 0xffffffffu : (d >= 31 ? 0u : (0xfffffffeu << d));
}
// bits of word `wi` whose rank is < hi (hi = NONE_RANK: all)
MMP_HD uint32_t mask_below(uint32_t wi, uint32_t hi) {
  int64_t d = (int64_t)hi - (int64_t)wi * 32;
  return d <= 0 ? 0u : (d >= 32 ? 0xffffffffu : ((1u << (int)d) - 1u));
}

// ---- the single-lane cooperative shape (CPU harness) ----
struct Coop1 {
  static constexpr int V = 1;
  static constexpr int L = 1;
  static constexpr int NW_CAP = 2048;  // 65536 instances
  int nj_;
  explicit Coop1(int row_words) : nj_(row_words) {}
  MMP_HD int nj() const { return nj_; }
  MMP_HD int lane() const { return 0; }
  MMP_HD uint32_t rmin(uint32_t x) const { return x; }
  MMP_HD uint32_t ror(uint32_t x) const { return x; }
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
  // f[k] = a[wi] & ~b[wi]
  MMP_HD void load_andnot(uint32_t *f, const uint32_t *a, const uint32_t *b, int row_words) const {
    for (int k = 0; k < row_words; k++) f[k] = a[k] & ~b[k];
  }
  MMP_HD void store_row(uint32_t *dst, const uint32_t *f, int row_words) const {
    for (int k = 0; k < row_words; k++) dst[k] = f[k];
  }
};

#if defined(__CUDACC__)
// ---- the warp cooperative shape ----
template <int V_, int NJ_>
struct Coop32 {
  static constexpr int V = V_;
  static constexpr int L = 32;
  static constexpr int NW_CAP = V_ * NJ_;
  int lane_;
  MMP_D Coop32() : lane_(threadIdx.x & 31) {}
  MMP_D int nj() const { return NJ_; }
  MMP_D int lane() const { return lane_; }
  MMP_D uint32_t rmin(uint32_t x) const { return __reduce_min_sync(0xffffffffu, x); }
  MMP_D uint32_t ror(uint32_t x) const { return __reduce_or_sync(0xffffffffu, x); }
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
  MMP_D void load_andnot(uint32_t *f, const uint32_t *a, const uint32_t *b, int row_words) const {
    const int nvec = row_words / V_;
#pragma unroll
    for (int j = 0; j < NJ_; j++) {
      int q = j * 32 + lane_;
      if (q < nvec) {
        if constexpr (V_ == 4) {
          // exclusion row: streamed once -> bypass L1; type mask: re-used by every decision of the type -> cached
          uint4 e, c = __ldg(reinterpret_cast<const uint4 *>(a) + q);
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                       : "l"(reinterpret_cast<const uint4 *>(b) + q));
          f[j * 4 + 0] = c.x & ~e.x; f[j * 4 + 1] = c.y & ~e.y; f[j * 4 + 2] = c.z & ~e.z; f[j * 4 + 3] = c.w & ~e.w;
        } else {
#pragma unroll
          for (int v = 0; v < V_; v++) {
            uint32_t e, c = __ldg(a + q * V_ + v);
            asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(e) : "l"(b + q * V_ + v));
            f[j * V_ + v] = c & ~e;
          }
        }
      } else {
#pragma unroll
        for (int v = 0; v < V_; v++) f[j * V_ + v] = 0;
      }
    }
  }
  MMP_D void store_row(uint32_t *dst, const uint32_t *f, int row_words) const {
    const int nvec = row_words / V_;
#pragma unroll
    for (int j = 0; j < NJ_; j++) {
      int q = j * 32 + lane_;
      if (q < nvec) {
#pragma unroll
        for (int v = 0; v < V_; v++) dst[q * V_ + v] = f[j * V_ + v];
      }
    }
  }
  // f = raw words of a row staged in shared memory (conflict-free: consecutive lanes read consecutive vectors)
  MMP_D void load_smem(uint32_t *f, const uint32_t *row_s, int row_words) const {
    const int nvec = row_words / V_;
#pragma unroll
    for (int j = 0; j < NJ_; j++) {
      int q = j * 32 + lane_;
      if (q < nvec) {
        if constexpr (V_ == 4) {
          uint4 e = reinterpret_cast<const uint4 *>(row_s)[q];
          f[j * 4 + 0] = e.x; f[j * 4 + 1] = e.y; f[j * 4 + 2] = e.z; f[j * 4 + 3] = e.w;
        } else {
#pragma unroll
          for (int v = 0; v < V_; v++) f[j * V_ + v] = row_s[q * V_ + v];
        }
      } else {
#pragma unroll
        for (int v = 0; v < V_; v++) f[j * V_ + v] = 0;
      }
    }
  }
  // f = cand & ~f  (f pre-filled with the exclusion words)
  MMP_D void combine_cand(uint32_t *f, const uint32_t *cand_row, int row_words) const {
    const int nvec = row_words / V_;
#pragma unroll
    for (int j = 0; j < NJ_; j++) {
      int q = j * 32 + lane_;
      if (q < nvec) {
        if constexpr (V_ == 4) {
          uint4 c = __ldg(reinterpret_cast<const uint4 *>(cand_row) + q);
          f[j * 4 + 0] = c.x & ~f[j * 4 + 0]; f[j * 4 + 1] = c.y & ~f[j * 4 + 1];
          f[j * 4 + 2] = c.z & ~f[j * 4 + 2]; f[j * 4 + 3] = c.w & ~f[j * 4 + 3];
        } else {
#pragma unroll
          for (int v = 0; v < V_; v++) f[j * V_ + v] = __ldg(cand_row + q * V_ + v) & ~f[j * V_ + v];
        }
      }
    }
  }
};
#endif

// local slot k = j*V + v  <->  bitmap word wi = (j*L + lane)*V + v   (ascending in k for a fixed lane)
template <class C> MMP_HD uint32_t word_index(const C &co, int k) {
  return (uint32_t)(((k / C::V) * C::L + co.lane()) * C::V + (k % C::V));
}
#if defined(__CUDA_ARCH__)
#define MMP_UNROLL _Pragma("unroll")
#define MMP_FOR_K(co, k) _Pragma("unroll") for (int k = 0; k < C::NW_CAP; ++k) if (k < (co).nj() * C::V)
#define MMP_FOR_J(co, j) _Pragma("unroll") for (int j = 0; j < C::NW_CAP / C::V; ++j) if (j < (co).nj())
#else
#define MMP_UNROLL
#define MMP_FOR_K(co, k) for (int k = 0; k < (co).nj() * C::V; ++k)
#define MMP_FOR_J(co, j) for (int j = 0; j < (co).nj(); ++j)
#endif

template <class C> MMP_HD uint32_t first_set(const C &co, const uint32_t *f) {
  uint32_t r = NONE_RANK;
  MMP_FOR_K(co, k) { if (r == NONE_RANK && f[k]) r = word_index(co, k) * 32 + ffs32(f[k]); }
  return co.rmin(r);
}
// first set bit of f & m(k, wi)
template <class C, class M> MMP_HD uint32_t first_set_where(const C &co, const uint32_t *f, M &&m) {
  uint32_t r = NONE_RANK;
  MMP_FOR_K(co, k) {
    if (r == NONE_RANK && f[k]) {
      uint32_t wi = word_index(co, k);
      uint32_t w = f[k] & m(wi);
      if (w) r = wi * 32 + ffs32(w);
    }
  }
  return co.rmin(r);
}
template <class C> MMP_HD bool test_bit(const C &co, const uint32_t *f, uint32_t rank) {
  bool p = false;
  MMP_FOR_K(co, k) { if (word_index(co, k) == (rank >> 5)) p = (f[k] >> (rank & 31)) & 1u; }
  return co.rany(p);
}
template <class C> MMP_HD void clear_bit(const C &co, uint32_t *f, uint32_t rank) {
  MMP_FOR_K(co, k) { if (word_index(co, k) == (rank >> 5)) f[k] &= ~(1u << (rank & 31)); }
}
template <class C> MMP_HD uint32_t popc_all(const C &co, const uint32_t *f) {
  uint32_t c = 0;
  MMP_FOR_K(co, k) { c += popc32(f[k]); }
  return co.rsum(c);
}
// rank of the kth (0-based) set bit in ascending rank order; k < popc_all(f)
template <class C> MMP_HD uint32_t select_kth(const C &co, const uint32_t *f, uint32_t kth) {
  uint32_t result = NONE_RANK;
  bool done = false;
  MMP_FOR_J(co, j) {
    if (!done) {
      uint32_t c = 0;
      MMP_UNROLL
      for (int v = 0; v < C::V; v++) c += popc32(f[j * C::V + v]);
      uint32_t pre = co.exscan(c), tot = co.rsum(c);
      if (kth < tot) {
        uint32_t mine = NONE_RANK;
        if (kth >= pre && kth < pre + c) {
          uint32_t rem = kth - pre;
          MMP_UNROLL
          for (int v = 0; v < C::V; v++) {
            uint32_t w = f[j * C::V + v];
            uint32_t pc = popc32(w);
            if (mine == NONE_RANK) {
              if (rem < pc) {
                for (uint32_t i = 0; i < rem; i++) w &= w - 1;
                mine = word_index(co, j * C::V + v) * 32 + ffs32(w);
              } else rem -= pc;
            }
          }
        }
        result = co.rmin(mine);
        done = true;
      } else kth -= tot;
    }
  }
  return result;
}

// First rank in f whose per-rank predicate holds.  `cls(wi)` classifies a 32-rank word from its min/max summary:
// 0 = no rank violates, 1 = every rank violates, 2 = mixed; `eval(rank)` is the exact per-rank test.  Words are
// resolved in ascending order and the search stops at the first hit, so the common case (a sorted fleet: one mixed
// word) costs one cooperative word evaluation.
template <class C, class CLS, class EV>
MMP_HD uint32_t first_violator(const C &co, const uint32_t *f, int32_t n_ranks, CLS &&cls, EV &&eval) {
  uint32_t A = NONE_RANK, M = NONE_RANK;
  MMP_FOR_K(co, k) {
    if (f[k] && A == NONE_RANK) {
      uint32_t wi = word_index(co, k);
      int c = cls(wi);
      if (c == 1) A = wi * 32 + ffs32(f[k]);
      else if (c == 2 && M == NONE_RANK) M = wi;
    }
  }
  uint32_t Amin = co.rmin(A), Mmin = co.rmin(M);
  while (Mmin != NONE_RANK && Mmin * 32 < Amin) {
    uint32_t mine = 0;
    MMP_FOR_K(co, k) { if (word_index(co, k) == Mmin) mine = f[k]; }
    uint32_t sw = co.ror(mine);
    uint32_t vm = co.eval_word(Mmin, n_ranks, eval) & sw;
    if (vm) { uint32_t r = Mmin * 32 + ffs32(vm); if (r < Amin) Amin = r; break; }
    if (M == Mmin) {
      M = NONE_RANK;
      MMP_FOR_K(co, k) {
        uint32_t wi = word_index(co, k);
        if (M == NONE_RANK && wi > Mmin && f[k] && cls(wi) == 2) M = wi;
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

// Everything about one decision that does not need its bitmap row (72 bytes).  k_place_ring lets lane j prepare the
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

// One getNext.  f is the caller-provided register/stack array of C::NW_CAP words; load_row(f, cand_row) must fill it
// with cand_row & ~(the decision's exclusion row).
// cand_rows (optional): [2][row_words] receives the candidate mask (other than best) and the survivor mask.
template <class C, class LOADER>
MMP_HD void decide_ctx(const SnapshotView &s, const DecisionCtx &c, const int32_t *extra, int64_t now, uint64_t seed,
                       uint64_t decision_id, const C &co, uint32_t *f, LOADER &&load_row, DecideOut &o, uint32_t *cand_rows) {
  o.target = MMP_TARGET_NONE; o.n_candidates = 0; o.best = -1; o.n_remaining = 0; o.pick_index = 0; o.flags = 0;
  o.cut_rank = (int32_t)NONE_RANK; o.best_rank = -1;
  const int RW = s.row_words;
  if (c.slot < 0) { o.target = TARGET_INVALID; return; }
  const mmp_decision_in &d = c.d;
  const int slot = c.slot;
  const int64_t last_used = c.last_used;
  const bool favour_self = (d.flags & MMP_DF_FAVOUR_SELF) != 0;
  const int32_t self_rank = c.self_rank;
  const FreshRow fr = c.fr;

  // ---- filter (MM:4760-4771) ----
  load_row(f, s.cand + (size_t)slot * RW);
  for (int e = 0; e < d.extra_n && e < 16; e++) {
    int32_t x = extra[d.extra_off + e];
    if (x >= 0 && x < s.max_instances) { int32_t r = s.rank_of[x]; if (r >= 0) clear_bit(co, f, (uint32_t)r); }
  }
  if (s.any_rs) {
    bool any = false;
    MMP_FOR_K(co, k) { if (f[k] & ~s.rs[word_index(co, k)]) any = true; }
    if (co.rany(any)) { MMP_FOR_K(co, k) { f[k] &= ~s.rs[word_index(co, k)]; } }
    else o.flags |= MMP_TF_RS_RETRY;  // MM:4798-4802: run the filter again without the replicaset exclusion
  }
  const uint32_t b = first_set(co, f);
  if (b == NONE_RANK) return;  // null
  const RankRow rb = s.rows[b];  // bestEntry.getValue()
  bool us = rb.idx == d.self;    // excluded self never passes the filter, so !excludeSelf is implied
  int64_t best_rem = us ? fr.rem : rb.rem, best_lru = us ? fr.lru : rb.lru;
  int32_t best_count = us ? fr.count : rb.count, best_rpm = us ? fr.rpm : rb.rpm, best_idx = rb.idx;
  uint32_t best_rank = b;
  const bool best_full = best_rem < s.min_space;
  if (best_full) o.flags |= MMP_TF_BEST_FULL;
  const bool has_pref = s.has_pref[slot] != 0;
  const uint32_t *P = s.pref + (size_t)slot * RW;
  bool simple = !has_pref || ((P[b >> 5] >> (b & 31)) & 1u);
  uint32_t lo = b, hi = NONE_RANK;
  bool use_pref = has_pref && simple;  // best is preferred: preference is treated as required (MM:4905-4907)
  o.best = best_idx; o.best_rank = (int32_t)b;

  if (!simple) {
    if (!best_full) {
      // non-simple (a) MM:4828-4852: first later entry that is preferred, unless a full one comes first
      uint32_t p1 = first_set_where(co, f, [&](uint32_t wi) { return P[wi] & mask_above(wi, b); });
      uint32_t f1 = first_set_where(co, f, [&](uint32_t wi) { return s.full[wi] & ~P[wi] & mask_above(wi, b); });
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
      MMP_FOR_K(co, k) { f[k] &= mask_above(word_index(co, k), b); }
      uint32_t kb = first_violator(co, f, s.n_ranks,
          [&](uint32_t wi) { WordSumL m = s.lsum[wi]; return !viol(m.hi) ? 0 : (viol(m.lo) ? 1 : 2); },
          [&](uint32_t r) { return viol(s.rows[r].lru); });
      bool anyp = false;
      MMP_FOR_K(co, k) { uint32_t wi = word_index(co, k); if (f[k] & P[wi] & mask_below(wi, kb)) anyp = true; }
      if (co.rany(anyp)) {
        // only preferred instances within the age distance are candidates; each records its own published rpm
        o.flags |= MMP_TF_PREF_B;
        MMP_FOR_K(co, k) { uint32_t wi = word_index(co, k); f[k] &= P[wi] & mask_below(wi, kb); }
        if (cand_rows) co.store_row(cand_rows, f, RW);
        const bool self_in = self_rank >= 0 && test_bit(co, f, (uint32_t)self_rank);
        if (self_in && favour_self) { o.flags |= MMP_TF_FAVOUR_EXIT; return; }  // N8: returns null
        const int32_t ccount = (int32_t)popc_all(co, f);
        o.n_candidates = ccount;
        uint32_t chosen;
        if (ccount == 1) { chosen = first_set(co, f); o.n_remaining = 1; }
        else {
          int32_t remaining = ccount;
          const int64_t ago = age_of(last_used, now);
          if (ago < 432000000LL) {
            int32_t mn = 2147483647;
            MMP_FOR_K(co, k) { uint32_t w = f[k], wi = word_index(co, k); while (w) { int bt = ffs32(w); w &= w - 1; int32_t v = s.rows[wi * 32 + bt].rpm; if (v < mn) mn = v; } }
            RpmFilter rf; rf.init(co.rmin_i(mn), ago);
            MMP_FOR_K(co, k) { uint32_t w = f[k], wi = word_index(co, k); while (w) { int bt = ffs32(w); w &= w - 1; if (rf.drop(s.rows[wi * 32 + bt].rpm)) f[k] &= ~(1u << bt); } }
            remaining = (int32_t)popc_all(co, f);
          }
          uint32_t index = remaining == 1 ? 0u : hash_index(seed, decision_id, (uint32_t)remaining);
          chosen = select_kth(co, f, index);
          o.n_remaining = remaining; o.pick_index = (int32_t)index;
        }
        if (cand_rows) co.store_row(cand_rows + RW, f, RW);
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
  MMP_FOR_K(co, k) {
    uint32_t wi = word_index(co, k);
    uint32_t m = mask_above(wi, lo) & mask_below(wi, hi);
    if (use_pref) m &= P[wi];
    f[k] &= m;
  }
  const bool self_in_s = self_rank >= 0 && test_bit(co, f, (uint32_t)self_rank);
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
  if (self_in_s) clear_bit(co, f, (uint32_t)self_rank);  // others only
  const int32_t thr = jaddi(best_count, best_count >> 2);
  auto cv = [&](int32_t c) { return c >= 10 && c > thr; };  // MM:4924-4927, always on the candidate's own count
  if (!best_full && self_in_s && cv(s.rows[self_rank].count)) self_viol = true;
  if (c_self) cut_others = first_set(co, f);
  else if (!best_full) {
    cut_others = first_violator(co, f, s.n_ranks,
        [&](uint32_t wi) { WordSumI m = s.csum[wi]; return !cv(m.hi) ? 0 : (cv(m.lo) ? 1 : 2); },
        [&](uint32_t r) { return cv(s.rows[r].count); });
  }
  const uint32_t cut_self = (self_in_s && self_viol) ? (uint32_t)self_rank : NONE_RANK;
  const uint32_t cut = cut_others < cut_self ? cut_others : cut_self;
  o.cut_rank = (int32_t)cut;
  const bool self_in_sl = self_in_s && (uint32_t)self_rank < cut;
  if (favour_self && self_in_sl) { o.flags |= MMP_TF_FAVOUR_EXIT; o.target = MMP_TARGET_SELF; return; }
  MMP_FOR_K(co, k) { f[k] &= mask_below(word_index(co, k), cut); }
  const int32_t n_others = (int32_t)popc_all(co, f);
  const int32_t ccount = 1 + n_others + (self_in_sl ? 1 : 0);
  o.n_candidates = ccount;
  const uint32_t self_w = (uint32_t)self_rank >> 5, self_b = 1u << (self_rank & 31);
  if (self_in_sl) { MMP_FOR_K(co, k) { if (word_index(co, k) == self_w) f[k] |= self_b; } }  // f = candidates other than best
  if (cand_rows) co.store_row(cand_rows, f, RW);
  bool keep_best = true, keep_others = true, keep_self = true;
  int32_t remaining = ccount;
  uint32_t index = 0;
  if (ccount > 1) {
    const int64_t ago = age_of(last_used, now);
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
  uint32_t chosen_rank;
  uint32_t kth = index;
  if (!keep_others) {
    const bool ks = self_in_sl && keep_self;
    MMP_FOR_K(co, k) { f[k] = (ks && word_index(co, k) == self_w) ? self_b : 0u; }
  } else if (self_in_sl && !keep_self) {
    MMP_FOR_K(co, k) { if (word_index(co, k) == self_w) f[k] &= ~self_b; }
  }
  if (cand_rows) co.store_row(cand_rows + RW, f, RW);
  if (keep_best && kth == 0) chosen_rank = best_rank;
  else { if (keep_best) kth--; chosen_rank = select_kth(co, f, kth); }
  const int32_t cidx = chosen_rank == best_rank ? best_idx : s.rows[chosen_rank].idx;
  o.target = (!favour_self && cidx == d.self) ? MMP_TARGET_SELF : cidx;
}

// Convenience form: context prepared inline, exclusion row read from global memory.
template <class C>
MMP_HD void decide(const SnapshotView &s, const mmp_decision_in &d, const FreshRow *fresh_tab, int32_t n_fresh,
                   const int32_t *extra, int64_t now, uint64_t seed, uint64_t decision_id, const C &co, uint32_t *f,
                   DecideOut &o, uint32_t *cand_rows) {
  DecisionCtx c;
  prepare_ctx(s, d, fresh_tab, n_fresh, c);
  const uint32_t *erow = s.excl + (size_t)(c.slot >= 0 ? d.model : 0) * s.row_words;
  decide_ctx(s, c, extra, now, seed, decision_id, co, f,
             [&](uint32_t *ff, const uint32_t *cand_row) { co.load_andnot(ff, cand_row, erow, s.row_words); }, o, cand_rows);
}

}  // namespace mmp
